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

constexpr int kPrefetch = 8;

__device__ __forceinline__ float logaddexp_f(float a, float b) {
  const float m = fmaxf(a, b);
  const float n = fminf(a, b);
  return m + kLn2 * lg2f(1.0f + ex2f((n - m) * kLog2e));
}

__global__ void __launch_bounds__(1024, 1)
lattice_alpha_beta_kernel(Lattice L, const float* __restrict__ lpb, const float* __restrict__ lpl,
                          float* __restrict__ alpha, float* __restrict__ beta, float* __restrict__ loss,
                          float* __restrict__ lnp_beta) {
  extern __shared__ float sbuf[];  // 2 x (blockDim.x + 2)
  const int b = blockIdx.x;
  const bool is_beta = blockIdx.y == 1;
  const int T = L.f_lens[b];
  const int U = L.y_lens[b];
  const int u = threadIdx.x;
  const int stride = blockDim.x + 2;
  float* s0 = sbuf;
  float* s1 = sbuf + stride;
  const size_t base = static_cast<size_t>(b) * L.D * L.U1max;
  const float* pb = lpb + base;
  const float* pl = lpl + base;
  const int dlast = T - 1 + U;
  const bool col_ok = u <= U;

  for (int i = threadIdx.x; i < 2 * stride; i += blockDim.x) sbuf[i] = kNeg;
  __syncthreads();

  float cb[kPrefetch], cl[kPrefetch];

  if (!is_beta) {
    float* out = alpha + base;
    // s?[u+1] holds the label-arc contribution arriving at column u+1; s?[0] stays kNeg.
    float a_cur = (u == 0) ? 0.0f : kNeg;
#pragma unroll
    for (int i = 0; i < kPrefetch; ++i) {
      const int d = i, t = d - u;
      const bool ok = col_ok && t >= 0 && t < T && d <= dlast;
      cb[i] = ok ? pb[static_cast<size_t>(d) * L.U1max + u] : kNeg;
      cl[i] = (ok && u < U) ? pl[static_cast<size_t>(d) * L.U1max + u] : kNeg;
    }
    for (int d0 = 0; d0 <= dlast; d0 += kPrefetch) {
      float nb[kPrefetch], nl[kPrefetch];
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const int d = d0 + kPrefetch + i, t = d - u;
        const bool ok = col_ok && t >= 0 && t < T && d <= dlast;
        nb[i] = ok ? pb[static_cast<size_t>(d) * L.U1max + u] : kNeg;
        nl[i] = (ok && u < U) ? pl[static_cast<size_t>(d) * L.U1max + u] : kNeg;
      }
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const int d = d0 + i;
        if (d <= dlast) {  // uniform across the CTA
          const int t = d - u;
          const bool valid = col_ok && t >= 0 && t < T;
          float* sw = (d & 1) ? s1 : s0;
          float ob = kNeg, ol = kNeg;
          if (valid) {
            out[static_cast<size_t>(d) * L.U1max + u] = a_cur;
            if (t + 1 < T) ob = a_cur + cb[i];
            if (u < U) ol = a_cur + cl[i];
            if (t == T - 1 && u == U) loss[b] = -(a_cur + cb[i]);
          }
          sw[u + 1] = ol;
          __syncthreads();
          a_cur = logaddexp_f(ob, sw[u]);
        }
      }
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) { cb[i] = nb[i]; cl[i] = nl[i]; }
    }
  } else {
    float* out = beta + base;
    // s?[u] holds beta of the previous (d+1) diagonal at column u.
    float b_prev = kNeg;
#pragma unroll
    for (int i = 0; i < kPrefetch; ++i) {
      const int d = dlast - i, t = d - u;
      const bool ok = col_ok && t >= 0 && t < T && d >= 0;
      cb[i] = ok ? pb[static_cast<size_t>(d) * L.U1max + u] : kNeg;
      cl[i] = (ok && u < U) ? pl[static_cast<size_t>(d) * L.U1max + u] : kNeg;
    }
    for (int d0 = dlast; d0 >= 0; d0 -= kPrefetch) {
      float nb[kPrefetch], nl[kPrefetch];
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const int d = d0 - kPrefetch - i, t = d - u;
        const bool ok = col_ok && t >= 0 && t < T && d >= 0;
        nb[i] = ok ? pb[static_cast<size_t>(d) * L.U1max + u] : kNeg;
        nl[i] = (ok && u < U) ? pl[static_cast<size_t>(d) * L.U1max + u] : kNeg;
      }
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) {
        const int d = d0 - i;
        if (d >= 0) {  // uniform across the CTA
          const int t = d - u;
          const bool valid = col_ok && t >= 0 && t < T;
          const float* sr = ((d + 1) & 1) ? s1 : s0;
          float* sw = (d & 1) ? s1 : s0;
          float bv = kNeg;
          if (valid) {
            if (t == T - 1 && u == U) {
              bv = cb[i];
            } else {
              const float x = (t + 1 < T) ? b_prev + cb[i] : kNeg;
              const float y = (u < U) ? sr[u + 1] + cl[i] : kNeg;
              bv = logaddexp_f(x, y);
            }
            out[static_cast<size_t>(d) * L.U1max + u] = bv;
            if (d == 0) lnp_beta[b] = bv;
          }
          b_prev = bv;
          sw[u] = bv;
          __syncthreads();
        }
      }
#pragma unroll
      for (int i = 0; i < kPrefetch; ++i) { cb[i] = nb[i]; cl[i] = nl[i]; }
    }
  }
}

__global__ void lattice_coefs_kernel(Lattice L, const float* __restrict__ lpb, const float* __restrict__ lpl,
                                     const float* __restrict__ alpha, const float* __restrict__ beta,
                                     const float* __restrict__ loss, float* __restrict__ c1,
                                     float* __restrict__ c2) {
  const int b = blockIdx.y;
  const int T = L.f_lens[b];
  const int U = L.y_lens[b];
  const int cells = L.D * L.U1max;
  const float lnp = -loss[b];
  const size_t base = static_cast<size_t>(b) * cells;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cells; i += gridDim.x * blockDim.x) {
    const int d = i / L.U1max;
    const int u = i - d * L.U1max;
    const int t = d - u;
    float v1 = 0.0f, v2 = 0.0f;
    if (u <= U && t >= 0 && t < T) {
      const float a = alpha[base + i];
      const float xb = lpb[base + i];
      if (t + 1 < T) {
        v1 = ex2f((a + xb + beta[base + i + L.U1max] - lnp) * kLog2e);
      } else if (u == U) {
        v1 = ex2f((a + xb - lnp) * kLog2e);
      }
      if (u < U) v2 = ex2f((a + lpl[base + i] + beta[base + i + L.U1max + 1] - lnp) * kLog2e);
    }
    c1[base + i] = v1;
    c2[base + i] = v2;
  }
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

void launch_lattice_alpha_beta(const Lattice& L, const float* lpb, const float* lpl, float* alpha, float* beta,
                               float* loss, float* lnp_beta, cudaStream_t s) {
  int threads = ((L.U1max + 31) / 32) * 32;
  if (threads < 32) threads = 32;
  const size_t smem = 2 * (threads + 2) * sizeof(float);
  lattice_alpha_beta_kernel<<<dim3(L.B, 2), threads, smem, s>>>(L, lpb, lpl, alpha, beta, loss, lnp_beta);
}

void launch_lattice_coefs(const Lattice& L, const float* lpb, const float* lpl, const float* alpha,
                          const float* beta, const float* loss, float* c1, float* c2, cudaStream_t s) {
  const int cells = L.D * L.U1max;
  int bx = (cells + 255) / 256;
  if (bx > 64) bx = 64;
  lattice_coefs_kernel<<<dim3(bx, L.B), 256, 0, s>>>(L, lpb, lpl, alpha, beta, loss, c1, c2);
}

}  // namespace rnnt
