// RNN-T lattice dynamic program (Graves 2012): alpha / beta anti-diagonal wavefronts and the
// arc-occupancy coefficients that turn softmax rows into dL/dlogits.
//
// One CTA per (utterance, direction); thread u owns lattice column u.  All per-cell arrays use the
// diagonal layout X[b][d=t+u][u] (common.cuh) so every wavefront step touches contiguous memory.
// The recurrences are latency-bound by construction (T+U-1 dependent steps, 2B independent chains);
// log-probabilities for the next kPrefetch diagonals are fetched ahead of the dependent chain.
//
// Replaces (by analogy; the reference has no RNN-T loss) the library call inside
// loss/ctc_loss.py:46-48,96-101.
#include "launch.h"
#include "ptx.cuh"

namespace rnnt {

namespace {

// Precision.  alpha / beta are sums of up to T + U log-probabilities: at T = 500, U = 100, V = 1024 they reach
// -4000, where one fp32 ulp is 4.9e-4 -- and the occupancies exp(alpha + beta + lp - lnP) inherit that *absolute*
// error as a *relative* one, accumulated over ~600 dependent steps (measured: 1.6e-3 .. 2e-3 relative error of df,
// dg, dW at the BASELINE shapes against an fp64 evaluation, tests/test_gpu_fullsize.py).  The state is therefore
// carried and stored in fp64; only the bounded correction ln(1 + e^-|a-b|) <= ln 2 is evaluated in fp32
// (ex2.approx / lg2.approx, absolute error ~1e-7).  Per cell and step this costs three DADDs and two conversions
// on top of the fp32 chain.
__device__ __forceinline__ double logaddexp_d(double a, double b) {
  // branch-free: the sign of a - b picks the maximum, |a - b| goes through the fp32 correction.  Both operands are
  // always finite (log(0) is the large negative kNegD, never -inf), so a - b is never NaN.
  const double d = a - b;
  const double m = __double2hiint(d) < 0 ? b : a;
  const float ad = static_cast<float>(fabs(d));             // huge differences become +inf: ex2(-inf) = 0
  return m + static_cast<double>(kLn2 * lg2f(1.0f + ex2f(-ad * kLog2e)));
}

constexpr double kNegD = -1.0e30;

// NC = lattice columns per thread (column u = threadIdx.x + k * blockDim.x): 1 up to 1024 columns, 2 / 4 beyond.
// kMaxT = launch bound: the usual U + 1 <= 256 gets a 256-thread variant whose register budget holds the prefetch
// registers without spills.
//
// The dependent chain of one wavefront step is  neighbour value (shared memory) -> logaddexp -> two adds -> shared
// memory -> barrier.  Everything else is kept off it and free of branches: the arc log-probabilities are prefetched
// kPrefetch diagonals ahead with "no arc" already encoded as kNeg (cells outside the lattice, the last frame's blank
// arc, the last column's label arc), stores are predicated, and the two special cells (alpha's final cell, beta's
// first) are handled by selects.
template <int NC, int kMaxT>
__global__ void __launch_bounds__(kMaxT, 1)
lattice_alpha_beta_kernel(Lattice L, const float* __restrict__ lpb, const float* __restrict__ lpl,
                          double* __restrict__ alpha, double* __restrict__ beta, float* __restrict__ loss,
                          float* __restrict__ lnp_beta, double* __restrict__ lnp64) {
  constexpr int kPrefetch = NC == 1 ? 8 : (NC == 2 ? 4 : 2);
  extern __shared__ double sbuf[];  // 2 x (NC * blockDim.x + 2)
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int T = L.f_lens[b];
  const int U = L.y_lens[b];
  const int U1 = L.U1max;
  const int ncols = NC * blockDim.x;
  const int stride = ncols + 2;
  double* s0 = sbuf;
  double* s1 = sbuf + stride;
  const size_t base = static_cast<size_t>(b) * L.D * U1;
  const float* pb = lpb + base;
  const float* pl = lpl + base;
  const int dlast = T - 1 + U;

  for (int i = threadIdx.x; i < 2 * stride; i += blockDim.x) sbuf[i] = kNegD;
  __syncthreads();

  // arc log-probabilities of cell (t = d - u, u) on diagonal d: eb = blank arc (t,u) -> (t+1,u), el = label arc
  // (t,u) -> (t,u+1); kNeg where the cell or the arc does not exist
  auto fetch = [&](int d, int u, float& eb, float& el) {
    const int t = d - u;
    const bool ok = u <= U && t >= 0 && t < T && d >= 0 && d <= dlast;
    eb = (ok && t + 1 < T) ? pb[static_cast<size_t>(d) * U1 + u] : kNeg;
    el = (ok && u < U) ? pl[static_cast<size_t>(d) * U1 + u] : kNeg;
  };

  float cb[NC][kPrefetch], cl[NC][kPrefetch];

  if (!is_beta) {
    double* out = alpha + base;
    // s?[u+1] holds the label-arc contribution arriving at column u+1; s?[0] stays kNeg.
    double a_cur[NC], a_fin = 0.0;
#pragma unroll
    for (int c = 0; c < NC; ++c) a_cur[c] = (c == 0 && threadIdx.x == 0) ? 0.0 : kNegD;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) fetch(i, threadIdx.x + c * blockDim.x, cb[c][i], cl[c][i]);
    for (int d0 = 0; d0 <= dlast; d0 += kPrefetch) {
      float nb[NC][kPrefetch], nl[NC][kPrefetch];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < kPrefetch; ++i) fetch(d0 + kPrefetch + i, threadIdx.x + c * blockDim.x, nb[c][i], nl[c][i]);
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const int d = d0 + i;
        if (d <= dlast) {  // uniform across the CTA
          double* sw = (d & 1) ? s1 : s0;
          double ob[NC];
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int u = threadIdx.x + c * blockDim.x;
            const int t = d - u;
            if (u <= U && t >= 0 && t < T) out[static_cast<size_t>(d) * U1 + u] = a_cur[c];   // predicated store
            a_fin = (d == dlast && u == U) ? a_cur[c] : a_fin;
            ob[c] = a_cur[c] + static_cast<double>(cb[c][i]);
            sw[u + 1] = a_cur[c] + static_cast<double>(cl[c][i]);
          }
          __syncthreads();
#pragma unroll
          for (int c = 0; c < NC; ++c) a_cur[c] = logaddexp_d(ob[c], sw[threadIdx.x + c * blockDim.x]);
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < kPrefetch; ++i) { cb[c][i] = nb[c][i]; cl[c][i] = nl[c][i]; }
    }
    // ln P = alpha[T-1, U] + lp_blank[T-1, U]: the thread that owns column U
#pragma unroll
    for (int c = 0; c < NC; ++c) {
      if (static_cast<int>(threadIdx.x + c * blockDim.x) == U) {
        const double lnp = a_fin + static_cast<double>(pb[static_cast<size_t>(dlast) * U1 + U]);
        loss[b] = static_cast<float>(-lnp);
        lnp64[b] = lnp;
      }
    }
  } else {
    double* out = beta + base;
    // s?[u] holds beta of the previous (d+1) diagonal at column u.
    double b_prev[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) b_prev[c] = kNegD;
    const double lpb_final = static_cast<double>(pb[static_cast<size_t>(dlast) * U1 + U]);   // beta[T-1, U]
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) fetch(dlast - i, threadIdx.x + c * blockDim.x, cb[c][i], cl[c][i]);
    for (int d0 = dlast; d0 >= 0; d0 -= kPrefetch) {
      float nb[NC][kPrefetch], nl[NC][kPrefetch];
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < kPrefetch; ++i) fetch(d0 - kPrefetch - i, threadIdx.x + c * blockDim.x, nb[c][i], nl[c][i]);
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const int d = d0 - i;
        if (d >= 0) {  // uniform across the CTA
          const double* sr = ((d + 1) & 1) ? s1 : s0;
          double* sw = (d & 1) ? s1 : s0;
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const int u = threadIdx.x + c * blockDim.x;
            const int t = d - u;
            double bv = logaddexp_d(b_prev[c] + static_cast<double>(cb[c][i]), sr[u + 1] + static_cast<double>(cl[c][i]));
            bv = (d == dlast && u == U) ? lpb_final : bv;
            if (u <= U && t >= 0 && t < T) out[static_cast<size_t>(d) * U1 + u] = bv;            // predicated store
            b_prev[c] = bv;
            sw[u] = bv;
          }
          __syncthreads();
        }
      }
#pragma unroll
      for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int i = 0; i < kPrefetch; ++i) { cb[c][i] = nb[c][i]; cl[c][i] = nl[c][i]; }
    }
    if (threadIdx.x == 0) lnp_beta[b] = static_cast<float>(b_prev[0]);   // beta[0, 0]
  }
}

__global__ void lattice_coefs_kernel(Lattice L, const float* __restrict__ lpb, const float* __restrict__ lpl,
                                     const double* __restrict__ alpha, const double* __restrict__ beta,
                                     const double* __restrict__ lnp64, float* __restrict__ c1,
                                     float* __restrict__ c2) {
  const int b = blockIdx.y;
  const int T = L.f_lens[b];
  const int U = L.y_lens[b];
  const int cells = L.D * L.U1max;
  const double lnp = lnp64[b];
  const size_t base = static_cast<size_t>(b) * cells;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
    const int d = i / L.U1max;
    const int u = i - d * L.U1max;
    const int t = d - u;
    float v1 = 0.0f, v2 = 0.0f;
    if (u <= U && t >= 0 && t < T) {
      const double a = alpha[base + i] - lnp;   // the large magnitudes cancel in fp64; the exponent is O(1..100)
      const double xb = static_cast<double>(lpb[base + i]);
      if (t + 1 < T) {
        v1 = ex2f(static_cast<float>(a + xb + beta[base + i + L.U1max]) * kLog2e);
      } else if (u == U) {
        v1 = ex2f(static_cast<float>(a + xb) * kLog2e);
      }
      if (u < U) v2 = ex2f(static_cast<float>(a + static_cast<double>(lpl[base + i]) + beta[base + i + L.U1max + 1]) * kLog2e);
    }
    c1[base + i] = v1;
    c2[base + i] = v2;
  }
}

// ---- which tiles does the backward pass need? ---------------------------------------------------------------------
// d loss / d logits of a lattice cell is  c0 * softmax - [blank] c1 - [label] c2  with the arc occupancies c1, c2 and
// c0 = c1 + c2: a tile whose 128 cells all have c0 == 0 -- the alignment never gets there: exp(alpha + beta - lnP)
// underflows, as it does for a fifth of the lattice of a 500 x 101 utterance -- contributes exact zeros to every gradient,
// so the backward pass walks the compacted list of the other tiles.  With eps > 0 tiles whose largest occupancy is below
// eps are dropped too (off by default).
__global__ void tile_activity_kernel(Lattice L, const float* __restrict__ c1, const float* __restrict__ c2,
                                     const float* __restrict__ grad_loss, float eps, int* __restrict__ flags) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= L.n_tiles_total) return;
  const TileInfo ti = decode_tile(L, warp);
  float m = 0.0f;
#pragma unroll
  for (int k = 0; k < kTileRows / 32; ++k) {
    const int r = lane + 32 * k;
    const int t = ti.t0 + (r >> 3), u = ti.u0 + (r & 7);
    if (t < ti.T && u <= ti.U) {
      const size_t i = diag_index(L, ti.b, t, u);
      m = fmaxf(m, c1[i] + c2[i]);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) flags[warp] = (m * fabsf(grad_loss[ti.b]) > eps) ? 1 : 0;
}

__global__ void __launch_bounds__(1024, 1)
compact_tiles_kernel(const int* __restrict__ flags, int n_tiles, int* __restrict__ active, int* __restrict__ n_active) {
  // one block; every thread scans kItems consecutive flags, so a batch of 13 k tiles takes a single pass
  constexpr int kItems = 16;
  __shared__ int warp_sums[32];
  __shared__ int base;
  if (threadIdx.x == 0) base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int start = 0; start < n_tiles; start += 1024 * kItems) {
    const int i0 = start + threadIdx.x * kItems;
    int f[kItems];
    int mine = 0;
#pragma unroll
    for (int k = 0; k < kItems; ++k) { f[k] = (i0 + k < n_tiles) ? flags[i0 + k] : 0; mine += f[k]; }
    int incl = mine;                                    // inclusive scan of the per-thread counts inside the warp
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int ws = warp_sums[lane];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, ws, o);
        if (lane >= o) ws += v;
      }
      warp_sums[lane] = ws;                             // inclusive over the warps
    }
    __syncthreads();
    int pos = base + (warp > 0 ? warp_sums[warp - 1] : 0) + incl - mine;
#pragma unroll
    for (int k = 0; k < kItems; ++k)
      if (f[k]) active[pos++] = i0 + k;
    __syncthreads();
    if (threadIdx.x == 0) base += warp_sums[31];
    __syncthreads();
  }
  if (threadIdx.x == 0) { n_active[0] = base; n_active[1] = n_tiles; }
}

// natural [B][Tmax][U1max] <-> diagonal [B][D][U1max] layout (only the explicit-logits entry point needs it)
__global__ void nat_to_diag_kernel(Lattice L, const float* __restrict__ a_nat, const float* __restrict__ b_nat,
                                   float* __restrict__ a_diag, float* __restrict__ b_diag) {
  const int b = blockIdx.y;
  const int n = L.Tmax * L.U1max;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i / L.U1max, u = i - t * L.U1max;
    const size_t di = diag_index(L, b, t, u);
    a_diag[di] = a_nat[static_cast<size_t>(b) * n + i];
    b_diag[di] = b_nat[static_cast<size_t>(b) * n + i];
  }
}
__global__ void diag_to_nat_kernel(Lattice L, const float* __restrict__ a_diag, const float* __restrict__ b_diag,
                                   float* __restrict__ a_nat, float* __restrict__ b_nat) {
  const int b = blockIdx.y;
  const int n = L.Tmax * L.U1max;
  const int T = L.f_lens[b], U = L.y_lens[b];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int t = i / L.U1max, u = i - t * L.U1max;
    const bool ok = t < T && u <= U;
    const size_t di = diag_index(L, b, t, u);
    a_nat[static_cast<size_t>(b) * n + i] = ok ? a_diag[di] : 0.0f;
    b_nat[static_cast<size_t>(b) * n + i] = ok ? b_diag[di] : 0.0f;
  }
}

}  // namespace

void launch_tile_activity(const Lattice& L, const float* c1, const float* c2, const float* grad_loss, float eps, int* flags,
                          cudaStream_t s) {
  if (L.n_tiles_total <= 0) return;
  const int blocks = (L.n_tiles_total * 32 + 255) / 256;
  tile_activity_kernel<<<blocks, 256, 0, s>>>(L, c1, c2, grad_loss, eps, flags);
}
void launch_compact_tiles(const int* flags, int n_tiles, int* active, int* n_active, cudaStream_t s) {
  compact_tiles_kernel<<<1, 1024, 0, s>>>(flags, n_tiles, active, n_active);
}

void launch_nat_to_diag(const Lattice& L, const float* a_nat, const float* b_nat, float* a_diag, float* b_diag,
                        cudaStream_t s) {
  int bx = (L.Tmax * L.U1max + 255) / 256;
  if (bx > 64) bx = 64;
  nat_to_diag_kernel<<<dim3(bx, L.B), 256, 0, s>>>(L, a_nat, b_nat, a_diag, b_diag);
}
void launch_diag_to_nat(const Lattice& L, const float* a_diag, const float* b_diag, float* a_nat, float* b_nat,
                        cudaStream_t s) {
  int bx = (L.Tmax * L.U1max + 255) / 256;
  if (bx > 64) bx = 64;
  diag_to_nat_kernel<<<dim3(bx, L.B), 256, 0, s>>>(L, a_diag, b_diag, a_nat, b_nat);
}

int lattice_max_columns() { return 4096; }

void launch_lattice_alpha_beta(const Lattice& L, const float* lpb, const float* lpl, double* alpha, double* beta,
                               float* loss, float* lnp_beta, double* lnp64, cudaStream_t s) {
  const int nc = L.U1max <= 1024 ? 1 : (L.U1max <= 2048 ? 2 : 4);
  int threads = (((L.U1max + nc - 1) / nc + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  const size_t smem = 2 * (static_cast<size_t>(nc) * threads + 2) * sizeof(double);
  const dim3 grid(L.B, 2);
  if (nc == 1 && threads <= 256) {
    lattice_alpha_beta_kernel<1, 256><<<grid, threads, smem, s>>>(L, lpb, lpl, alpha, beta, loss, lnp_beta, lnp64);
  } else if (nc == 1) {
    lattice_alpha_beta_kernel<1, 1024><<<grid, threads, smem, s>>>(L, lpb, lpl, alpha, beta, loss, lnp_beta, lnp64);
  } else if (nc == 2) {
    lattice_alpha_beta_kernel<2, 1024><<<grid, threads, smem, s>>>(L, lpb, lpl, alpha, beta, loss, lnp_beta, lnp64);
  } else {
    // 2 x 4098 doubles = 65.6 KB of dynamic shared memory: above the 48 KB default (idempotent, per device)
    cudaFuncSetAttribute(lattice_alpha_beta_kernel<4, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    lattice_alpha_beta_kernel<4, 1024><<<grid, threads, smem, s>>>(L, lpb, lpl, alpha, beta, loss, lnp_beta, lnp64);
  }
}

void launch_lattice_coefs(const Lattice& L, const float* lpb, const float* lpl, const double* alpha,
                          const double* beta, const double* lnp64, float* c1, float* c2, cudaStream_t s) {
  const int cells = L.D * L.U1max;
  int bx = (cells + 255) / 256;
  if (bx > 64) bx = 64;
  lattice_coefs_kernel<<<dim3(bx, L.B), 256, 0, s>>>(L, lpb, lpl, alpha, beta, lnp64, c1, c2);
}

}  // namespace rnnt
