// Whole RNN-T greedy decode loop in ONE cooperative launch (SURVEY.md 8(a8) + 8(f) rank 2).
//
// The reference-style decoder (src/myrtlespeech/post_process/ctc_greedy_decoder.py:77-92 is the CTC analog) runs one
// host iteration per emitted symbol / frame.  Here the loop itself lives on the device.  One decode step is three
// small GEMMs over the whole batch, each followed by a grid-wide barrier:
//
//   L  prediction-network LSTM cell    gates[b, :] = table[label_b, :] + h_b . W_hh^T          (K = Hp, N = 4 Hp)
//   P  projection + joint activation   g[b, :] = h_b . W_p^T + b_p ; hj = bf16(tanh(f[b, t_b] + bf16(g)))  (K = Hp, N = H)
//   J  joint projection + argmax       k_b = argmax_v (hj_b . W^T + bias)                       (K = H,  N = V)
//
// `table[v] = W_ih . emb[v] + b_ih + b_hh` folds the embedding lookup and the input half of the cell into one gather
// (row V is the start-of-sequence input).  Work is split across CTAs along N: every CTA keeps its slice of the three
// weight matrices RESIDENT in shared memory for the whole decode (loaded once by TMA), the batch is the M = 128
// dimension of a tcgen05.mma (bf16 in, fp32 accumulate in TMEM), and only the activations (h, hj: B x Hp / B x H bf16,
// L2-resident) are streamed per step through a TMA ring.  The per-CTA cell state c, the CTA's slice of h and of g, and a
// redundant copy of the per-utterance bookkeeping (frame, symbols at this frame, count) live in shared memory; the
// cross-CTA argmax is an atomicMax on a packed (ordered logit, ~index) key.  Steps in which no utterance emitted a
// symbol skip the L phase and the P GEMM (g is unchanged; only the frame moved).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "launch.h"
#include "ptx.cuh"

namespace rnnt {
namespace {

constexpr int kBM = 128;              // batch rows per M tile == TMEM lanes
constexpr int kBK = 64;               // bf16 per k-block (one 128-byte swizzle row)
constexpr int kAStage = kBM * kBK * 2;  // 16 KB
constexpr int kDecThreads = 192;      // warp 0: TMA, warp 1: MMA issue, warps 2..5: epilogue
constexpr int kDecTmemCols = 256;
constexpr int kMaxStages = 12;

__device__ __forceinline__ void fence_proxy_async_global_() { asm volatile("fence.proxy.async.global;" ::: "memory"); }
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_add_u32(unsigned* ptr, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_cg_u64(const unsigned long long* ptr) {
  unsigned long long v;
  asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ bool mbar_try_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_addr(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait_addr(bar, parity)) {
    if (++spins > (1u << 26)) __trap();
  }
}
__device__ __forceinline__ void umma_commit_addr(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint32_t ordered_bits(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
// Gate non-linearities of the cell: ex2.approx + rcp.approx (absolute error ~1e-7, far below the bf16 rounding of h);
// the libm versions cost ~40 dependent instructions each on an epilogue warp that runs alone on its scheduler.
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return fmaf(2.0f, sigmoid_f(2.0f * x), -1.0f); }
__device__ __forceinline__ float bf16_round_f(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

struct Pipe {
  uint8_t* ring;
  uint64_t* full;
  uint64_t* empty;
  uint64_t* tfull;
  int n_stages;
  uint32_t tmem;
  uint32_t it;     // k-blocks pushed through the ring so far (every thread keeps the same count)
  uint32_t n_acc;  // accumulator tiles produced so far
};

// One [128 x n] = A[128 x K] . Wres[n x K]^T tile: warp 0 streams A through the ring, warp 1 issues the MMAs against
// the resident weight slice, everybody else returns immediately and waits on `tfull` in its epilogue.
__device__ __forceinline__ void gemm_tile(Pipe& pp, const CUtensorMap* tm_a, int a_row0, int kb, uint32_t w_smem, int n,
                                          int warp, int lane) {
  if (warp == 0) {
    if (lane == 0) {
      fence_proxy_async_global_();  // other CTAs' generic-proxy writes (ordered by the grid barrier) -> async proxy
      for (int k = 0; k < kb; ++k) {
        const uint32_t it = pp.it + k;
        const int s = it % pp.n_stages;
        const uint32_t ph = (it / pp.n_stages) & 1;
        mbar_wait(&pp.empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&pp.full[s], kAStage);
        tma_load_2d(pp.ring + s * kAStage, tm_a, &pp.full[s], k * kBK, a_row0);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(kBM, n, false, false);
    for (int k = 0; k < kb; ++k) {
      const uint32_t it = pp.it + k;
      const int s = it % pp.n_stages;
      const uint32_t ph = (it / pp.n_stages) & 1;
      mbar_wait(&pp.full[s], ph);
      tc_fence_after();
      if (elect_one()) {  // converged warp + elect.sync: no divergence waterfall around the tcgen05 instructions
        const uint32_t a_addr = smem_u32(pp.ring + s * kAStage);
        const uint32_t b_addr = w_smem + k * n * 128;
#pragma unroll
        for (int kk = 0; kk < kBK / 16; ++kk) {
          const uint64_t ad = make_smem_desc_sw128(a_addr + kk * 32, 16, 1024);
          const uint64_t bd = make_smem_desc_sw128(b_addr + kk * 32, 16, 1024);
          umma_bf16(pp.tmem, ad, bd, idesc, (k | kk) != 0 ? 1u : 0u);
        }
        umma_commit(&pp.empty[s]);
        if (k == kb - 1) umma_commit(pp.tfull);
      }
      __syncwarp();
    }
  }
  pp.it += kb;
}

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    red_release_add_u32(counter, 1u);
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target) {
      if (++spins > (1u << 25)) __trap();  // a protocol bug traps instead of hanging the GPU
    }
  }
  __syncthreads();
}

}  // namespace

__global__ void __launch_bounds__(kDecThreads, 1)
greedy_decode_kernel(const __grid_constant__ CUtensorMap tm_hj, const __grid_constant__ CUtensorMap tm_hbuf,
                     const __grid_constant__ CUtensorMap tm_wj, const __grid_constant__ CUtensorMap tm_wl,
                     const __grid_constant__ CUtensorMap tm_wp, const DecodeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* wj = smem + p.o_wj;
  uint8_t* wl = smem + p.o_wl;
  uint8_t* wp = smem + p.o_wp;
  float* c_s = reinterpret_cast<float*>(smem + p.o_c);                      // [Bp][nu]   cell state
  __nv_bfloat16* h_s = reinterpret_cast<__nv_bfloat16*>(smem + p.o_h);      // [Bp][nu]   this CTA's units of h
  float* g_s = reinterpret_cast<float*>(smem + p.o_g);                      // [Bp][nP]   this CTA's columns of bf16(g)
  int* s_t = reinterpret_cast<int*>(smem + p.o_state);                      // current frame
  int* s_em = s_t + p.B;                                                    // symbols emitted at this frame
  int* s_n = s_em + p.B;                                                    // symbols emitted so far
  int* s_lab = s_n + p.B;                                                   // label to feed the LSTM, -1 = none
  int* s_len = s_lab + p.B;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.o_bars);
  uint64_t* full_bar = bars;                    // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;      // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;  // [1]
  uint64_t* w_bar = bars + 2 * kMaxStages + 1;  // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x;
  const unsigned G = gridDim.x;
  const bool in_j = cta < p.nslJ, in_l = cta < p.nslL, in_p = cta < p.nslP;
  const int n_mt = p.Bp / kBM;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_hj); prefetch_tmap(&tm_hbuf); prefetch_tmap(&tm_wj); prefetch_tmap(&tm_wl); prefetch_tmap(&tm_wp);
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(tfull_bar, 1);
    mbar_init(w_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, kDecTmemCols);
    tmem_relinquish();
  }
  for (int b = threadIdx.x; b < p.B; b += kDecThreads) {
    s_t[b] = 0; s_em[b] = 0; s_n[b] = 0; s_lab[b] = p.V;  // start of sequence: every utterance steps the LSTM once
    s_len[b] = p.lens[b];
  }
  for (int i = threadIdx.x; i < p.Bp * p.nu; i += kDecThreads) { c_s[i] = 0.0f; h_s[i] = __float2bfloat16(0.0f); }
  for (int i = threadIdx.x; i < p.Bp * p.nP; i += kDecThreads) g_s[i] = 0.0f;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();

  // resident weight slices: loaded once, one barrier
  if (threadIdx.x == 0) {
    uint32_t bytes = 0;
    if (in_j) bytes += static_cast<uint32_t>(p.kbJ) * p.nJ * 128;
    if (in_l) bytes += static_cast<uint32_t>(p.kbL) * p.nL * 128;
    if (in_p) bytes += static_cast<uint32_t>(p.kbP) * p.nP * 128;
    if (bytes) mbar_arrive_expect_tx(w_bar, bytes); else mbar_arrive(w_bar);
    if (in_j) for (int k = 0; k < p.kbJ; ++k) tma_load_2d(wj + k * p.nJ * 128, &tm_wj, w_bar, k * kBK, cta * p.nJ);
    if (in_l) for (int k = 0; k < p.kbL; ++k) tma_load_2d(wl + k * p.nL * 128, &tm_wl, w_bar, k * kBK, cta * p.nL);
    if (in_p) for (int k = 0; k < p.kbP; ++k) tma_load_2d(wp + k * p.nP * 128, &tm_wp, w_bar, k * kBK, cta * p.nP);
  }
  if (warp == 1) { mbar_wait(w_bar, 0); tc_fence_after(); }

  Pipe pp;
  pp.ring = smem; pp.full = full_bar; pp.empty = empty_bar; pp.tfull = tfull_bar;
  pp.n_stages = p.n_stages; pp.tmem = *tmem_slot; pp.it = 0; pp.n_acc = 0;

  const bool epi = warp >= 2;
  const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
  const int r = quad * 32 + lane;            // row inside the M tile == TMEM lane
  const uint32_t lane_taddr = pp.tmem + (static_cast<uint32_t>(quad * 32) << 16);

  unsigned n_bar = 0;
  int par = 0;          // h buffer that holds the current hidden state
  bool any_sym = true;  // start-of-sequence step
  const int u0 = cta * p.nu, n0p = cta * p.nP, v0 = cta * p.nJ;
  const size_t gate_pitch = static_cast<size_t>(4) * p.Hp;

  for (int step = 0; step <= p.max_steps; ++step) {
    // ------------------------------------------------------------------ L: LSTM cell
    if (any_sym) {
      if (in_l) {
        for (int mt = 0; mt < n_mt; ++mt) {
          gemm_tile(pp, &tm_hbuf, par * p.Bp + mt * kBM, p.kbL, smem_u32(wl), p.nL, warp, lane);
          if (epi) {
            mbar_wait(tfull_bar, pp.n_acc & 1);
            tc_fence_after();
            const int b = mt * kBM + r;
            const int lab = b < p.B ? s_lab[b] : -1;
            const float* tb = p.table + static_cast<size_t>(lab < 0 ? 0 : lab) * gate_pitch;
            for (int g = 0; g < (p.nL + 31) / 32; ++g) {
              uint32_t raw[32];
              tmem_ld32(lane_taddr + g * 32, raw);
              tmem_ld_wait();
              if (lab >= 0) {
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                  const int ul = g * 8 + jj, u = u0 + ul;
                  if (ul < p.nu && u < p.Hp) {
                    const float ai = __uint_as_float(raw[4 * jj + 0]) + __ldg(tb + u);
                    const float af = __uint_as_float(raw[4 * jj + 1]) + __ldg(tb + p.Hp + u);
                    const float ag = __uint_as_float(raw[4 * jj + 2]) + __ldg(tb + 2 * p.Hp + u);
                    const float ao = __uint_as_float(raw[4 * jj + 3]) + __ldg(tb + 3 * p.Hp + u);
                    const float c_new = sigmoid_f(af) * c_s[b * p.nu + ul] + sigmoid_f(ai) * tanh_f(ag);
                    c_s[b * p.nu + ul] = c_new;
                    h_s[b * p.nu + ul] = __float2bfloat16_rn(sigmoid_f(ao) * tanh_f(c_new));
                  }
                }
              }
            }
            if (b < p.B) {
              // publish this CTA's units of h (updated or carried over) into the other h buffer
              __nv_bfloat16* dst = p.hbuf + (static_cast<size_t>(par ^ 1) * p.Bp + b) * p.Hp + u0;
              for (int ul = 0; ul < p.nu; ul += 4)
                if (u0 + ul < p.Hp) *reinterpret_cast<uint2*>(dst + ul) = *reinterpret_cast<const uint2*>(h_s + b * p.nu + ul);
            }
            __threadfence();
            fence_proxy_async_global_();
            tc_fence_before();
          }
          __syncthreads();
          tc_fence_after();
          ++pp.n_acc;
        }
      }
      grid_barrier(p.gbar, G * (++n_bar));
      par ^= 1;
    }
    // ------------------------------------------------------------------ P: projection + tanh(f + g)
    if (in_p) {
      for (int mt = 0; mt < n_mt; ++mt) {
        if (any_sym) gemm_tile(pp, &tm_hbuf, par * p.Bp + mt * kBM, p.kbP, smem_u32(wp), p.nP, warp, lane);
        if (epi) {
          const int b = mt * kBM + r;
          if (any_sym) {
            mbar_wait(tfull_bar, pp.n_acc & 1);
            tc_fence_after();
            for (int g = 0; g < (p.nP + 31) / 32; ++g) {
              uint32_t raw[32];
              tmem_ld32(lane_taddr + g * 32, raw);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) {
                const int col = g * 32 + i, n = n0p + col;
                if (col < p.nP && n < p.H)
                  g_s[b * p.nP + col] = bf16_round_f(__uint_as_float(raw[i]) + (p.bias_p ? __ldg(p.bias_p + n) : 0.0f));
              }
            }
            tc_fence_before();
          }
          if (b < p.B) {
            int t = s_t[b];
            if (t > p.Tmax - 1) t = p.Tmax - 1;
            const __nv_bfloat16* fr = p.f + (static_cast<size_t>(b) * p.Tmax + t) * p.H + n0p;
            __nv_bfloat16* dst = p.hj + static_cast<size_t>(b) * p.H + n0p;
            for (int col = 0; col < p.nP; col += 8) {
              if (n0p + col < p.H) {  // H % 8 == 0: whole 16-byte vectors
                const uint4 fv = __ldg(reinterpret_cast<const uint4*>(fr + col));
                const uint32_t fw[4] = {fv.x, fv.y, fv.z, fv.w};
                const float* gs = g_s + b * p.nP + col;
                uint32_t o[4];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  o[e] = pack_bf16x2(tanh_approx(bf16lo(fw[e]) + gs[2 * e]), tanh_approx(bf16hi(fw[e]) + gs[2 * e + 1]));
                *reinterpret_cast<uint4*>(dst + col) = make_uint4(o[0], o[1], o[2], o[3]);
              }
            }
          }
          __threadfence();
          fence_proxy_async_global_();
        }
        if (any_sym) {
          __syncthreads();
          tc_fence_after();
          ++pp.n_acc;
        }
      }
    }
    grid_barrier(p.gbar, G * (++n_bar));
    // ------------------------------------------------------------------ J: joint projection + argmax
    unsigned long long* amax = p.amax + static_cast<size_t>(step & 1) * p.Bp;
    if (in_j) {
      for (int mt = 0; mt < n_mt; ++mt) {
        gemm_tile(pp, &tm_hj, mt * kBM, p.kbJ, smem_u32(wj), p.nJ, warp, lane);
        if (epi) {
          mbar_wait(tfull_bar, pp.n_acc & 1);
          tc_fence_after();
          const int b = mt * kBM + r;
          float best = -INFINITY;
          int best_v = -1;
          for (int g = 0; g < (p.nJ + 31) / 32; ++g) {
            uint32_t raw[32];
            tmem_ld32(lane_taddr + g * 32, raw);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int col = g * 32 + i, v = v0 + col;
              if (col < p.nJ && v < p.V) {
                const float z = __uint_as_float(raw[i]) + (p.bias_j ? __ldg(p.bias_j + v) : 0.0f);
                if (z > best || best_v < 0) { if (z == z) { best = z; best_v = v; } }  // ascending v: lowest index wins ties
              }
            }
          }
          if (b < p.B && best_v >= 0) {
            const unsigned long long key =
                (static_cast<unsigned long long>(ordered_bits(best)) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(best_v));
            atomicMax(amax + b, key);
          }
          tc_fence_before();
        }
        __syncthreads();
        tc_fence_after();
        ++pp.n_acc;
      }
    }
    grid_barrier(p.gbar, G * (++n_bar));
    // ------------------------------------------------------------------ bookkeeping (identical in every CTA)
    int active = 0, sym_any = 0;
    for (int b = threadIdx.x; b < p.B; b += kDecThreads) {
      const unsigned long long key = ld_cg_u64(amax + b);
      int t = s_t[b], lab = -1;
      if (t < s_len[b]) {
        const int k = key ? static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key)) : p.blank;
        int em = s_em[b];
        if (k != p.blank) {
          const int n = s_n[b];
          if (cta == 0 && n < p.sym_cap) p.sym[static_cast<size_t>(b) * p.sym_cap + n] = k;
          s_n[b] = n + 1;
          ++em;
          lab = k;
        }
        if (lab < 0 || em >= p.S) { ++t; em = 0; }
        s_t[b] = t; s_em[b] = em;
      }
      s_lab[b] = lab;
      active |= (t < s_len[b]) ? 1 : 0;
      sym_any |= (lab >= 0) ? 1 : 0;
      if (cta == 0) p.amax[static_cast<size_t>((step + 1) & 1) * p.Bp + b] = 0ull;  // next step's keys
    }
    const int any_active = __syncthreads_or(active);
    any_sym = __syncthreads_or(sym_any) != 0;
    if (!any_active) break;
  }

  if (cta == 0)
    for (int b = threadIdx.x; b < p.B; b += kDecThreads) p.n_sym[b] = s_n[b] < p.sym_cap ? s_n[b] : p.sym_cap;
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(pp.tmem, kDecTmemCols);
  }
}

// W_hh [4 Hp][Hp] (torch gate order i, f, g, o) -> rows ordered [slice][unit in slice][gate], so that a CTA's slice is
// contiguous and the four gates of one hidden unit sit in adjacent accumulator columns.  Rows past Hp are zero.
__global__ void permute_whh_kernel(const __nv_bfloat16* __restrict__ W, __nv_bfloat16* __restrict__ out, int Hp, int nu,
                                   int n_rows) {
  const int R = blockIdx.x;
  if (R >= n_rows) return;
  const int nL = 4 * nu;
  const int slice = R / nL, rem = R - slice * nL;
  const int j = rem >> 2, gate = rem & 3;
  const int u = slice * nu + j;
  __nv_bfloat16* dst = out + static_cast<size_t>(R) * Hp;
  if (u < Hp) {
    const __nv_bfloat16* src = W + (static_cast<size_t>(gate) * Hp + u) * Hp;
    for (int i = threadIdx.x; i < Hp; i += blockDim.x) dst[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < Hp; i += blockDim.x) dst[i] = __float2bfloat16(0.0f);
  }
}


// ====================================================================================================================
// Cluster variant (default): one thread-block cluster per 16 utterances, no grid-wide synchronisation at all.
//
// The three per-step products are turned around: the WEIGHTS are the M side of the MMA (128 output features per tile,
// streamed from L2 by TMA -- they are static, so the producer warp runs ahead of the recurrence and the ring is always
// full), the cluster's 16 utterances are the N side and their activations (h: 16 x Hp, hj: 16 x H, bf16) stay RESIDENT in
// every CTA's shared memory.  CTA `rank` of a cluster of C owns 1/C of the rows of each weight matrix; after each
// product it writes its slice of the new activation straight into the operand buffer (128-byte-swizzled K-major
// layout, its slice is a contiguous run of 2 KB k-blocks) and pushes that run into the other C-1 CTAs with one
// cp.async.bulk shared::cta -> shared::cluster each, completing on the RECEIVER's mbarrier; the per-utterance argmax is
// exchanged the same way with st.async.  The MMA warp therefore waits only on transaction barriers, never on a
// software barrier, and the W_hh . h product of step s+1 is issued while the argmax of step s is still in flight (it
// does not depend on the emitted label -- only its epilogue does).
// ====================================================================================================================
__device__ unsigned long long g_dec_prof[16];  // cycles per epilogue stage, summed over steps (cluster 0, rank 0)
#define DSTAMP(i)                                                \
  do {                                                           \
    if (prof) { const long long now_ = clock64(); acc[i] += now_ - last; last = now_; } \
  } while (0)

constexpr int kNB = 16;               // utterances per cluster == UMMA N
constexpr int kNH = kNB / 2;           // utterances per epilogue warp
constexpr int kKBlk = kNB * 128;      // bytes of one k-block of an activation buffer (16 rows x 128 B)
constexpr int kCThreads = 352;        // warp 0: TMA, warps 1 and 10: MMA issue, warps 2..9: epilogue
constexpr int kMma2Warp = 10;
constexpr int kEpiThreads = 256;
constexpr int kMaxTiles = 10;         // accumulator tiles (L of every layer + P + J) per CTA
constexpr int kMaxLayers = 3;         // LSTM layers of the prediction network
constexpr int kMaxLTiles = 4;         // gate tiles per layer and CTA (32 hidden units each)
constexpr int kTileCols = 2 * kNB;    // TMEM columns per accumulator tile: one partial accumulator per issuing warp.
                                      // (Spreading one warp's MMAs over four accumulators was measured: no change --
                                      // the cost per MMA is issue latency of the thread, not an accumulator dependency.)

// This thread's 8 utterances of one accumulator tile: the sum of the two issuing warps' partial accumulators (the
// second one was never written when the product has a single k-block).
__device__ __forceinline__ void ld_acc(uint32_t taddr, bool two, float (&v)[kNH]) {
  uint32_t a[kNH], b[kNH];
  tmem_ld8(taddr, a);
  tmem_ld8(taddr + kNB, b);
  tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < kNH; ++i) v[i] = two ? __uint_as_float(a[i]) + __uint_as_float(b[i]) : __uint_as_float(a[i]);
}

__device__ __forceinline__ uint32_t act_offset(int n, int utt) {  // byte offset of element (utt, feature n), SW128 K-major
  return static_cast<uint32_t>((n >> 6) * kKBlk + utt * 128 + ((((n & 63) >> 3) ^ (utt & 7)) << 4) + (n & 7) * 2);
}

__global__ void __launch_bounds__(kCThreads, 1)
greedy_decode_cluster_kernel(const __grid_constant__ CUtensorMap tm_wj, const __grid_constant__ CUtensorMap tm_wl,
                             const __grid_constant__ CUtensorMap tm_wu, const __grid_constant__ CUtensorMap tm_wp,
                             const ClusterDecodeArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* ring = smem;
  uint8_t* hj = smem + p.o_hj;
  // per LSTM layer l: [h buffer 0][h buffer 1][c: up x 16 f32][own units of h: up x 16 bf16], `layer_stride` apart
  const int n_h_bytes = (p.C * p.up / 64) * kKBlk;
  auto hbuf = [&](int l, int i) { return smem + p.o_layers + l * p.layer_stride + i * n_h_bytes; };
  auto cbuf = [&](int l) { return reinterpret_cast<float*>(smem + p.o_layers + l * p.layer_stride + 2 * n_h_bytes); };
  auto hownbuf = [&](int l) {
    return reinterpret_cast<__nv_bfloat16*>(smem + p.o_layers + l * p.layer_stride + 2 * n_h_bytes + p.up * kNB * 4);
  };
  float* gates = reinterpret_cast<float*>(smem + p.o_gates);                 // [mtL][4][32][16]
  unsigned long long* amax_s = reinterpret_cast<unsigned long long*>(smem + p.o_amax);  // [C][16]
  unsigned long long* part = reinterpret_cast<unsigned long long*>(smem + p.o_part);    // [4][16]
  int* s_t = reinterpret_cast<int*>(smem + p.o_state);
  int* s_em = s_t + kNB;
  int* s_n = s_em + kNB;
  int* s_lab = s_n + kNB;
  int* s_len = s_lab + kNB;
  volatile int* s_go = s_len + kNB;  // [2]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + p.o_bars);
  uint64_t* full_bar = bars;                          // [kMaxStages]
  uint64_t* empty_bar = bars + kMaxStages;            // [kMaxStages]
  uint64_t* tfull_bar = bars + 2 * kMaxStages;        // [kMaxTiles]
  uint64_t* hfull_bar = tfull_bar + kMaxTiles;        // [kMaxLayers][2]
  uint64_t* hjfull_bar = hfull_bar + 2 * kMaxLayers;
  uint64_t* amaxfull_bar = hjfull_bar + 1;
  uint64_t* step_bar = amaxfull_bar + 1;
  uint64_t* fin_bar = step_bar + 1;
  uint64_t* res_bar = fin_bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_bar + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int C = p.C;
  const int b0 = (blockIdx.x / C) * kNB;
  const int NL = p.NL;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_wj); prefetch_tmap(&tm_wl); prefetch_tmap(&tm_wu); prefetch_tmap(&tm_wp);
    for (int i = 0; i < kMaxStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < kMaxTiles; ++i) mbar_init(&tfull_bar[i], 2);   // one arrival per issuing warp
    for (int i = 0; i < 2 * kMaxLayers; ++i) mbar_init(&hfull_bar[i], 1);
    mbar_init(hjfull_bar, 1); mbar_init(amaxfull_bar, 1); mbar_init(step_bar, 1); mbar_init(fin_bar, 2);
    mbar_init(res_bar, 128);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, p.tmem_cols);
    tmem_relinquish();
  }
  for (int i = threadIdx.x; i < NL * p.layer_stride / 4; i += kCThreads)   // h_{-1} = 0, c_{-1} = 0 for every layer
    reinterpret_cast<uint32_t*>(smem + p.o_layers)[i] = 0u;
  if (threadIdx.x < kNB) {
    const int b = b0 + threadIdx.x;
    s_t[threadIdx.x] = 0; s_em[threadIdx.x] = 0; s_n[threadIdx.x] = 0;
    s_len[threadIdx.x] = b < p.B ? p.lens[b] : 0;
    s_lab[threadIdx.x] = b < p.B ? p.V : -1;  // start of sequence: every utterance steps the LSTM once
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // every CTA's barriers and buffers exist before anybody pushes into them
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int slotP0 = NL * p.mtL, slotJ0 = NL * p.mtL + p.mtP;   // accumulator slots: layer l tile m = l mtL + m, then P, J

  if (warp == 0) {
    // ------------------------------------------------------------------ weight stream (runs ahead of the recurrence)
    if (lane == 0) {
      uint32_t it = 0;
      const bool prof = p.prof && blockIdx.x == 0;
      long long w_empty = 0, w_step = 0;
      auto push = [&](const CUtensorMap* tm, int row0, int kb, int k0 = 0) {   // k-blocks [k0, kb) of one 128-row tile
        for (int k = k0; k < kb; ++k, ++it) {
          const int st = it % p.n_stages;
          const long long t0 = prof ? clock64() : 0;
          mbar_wait(&empty_bar[st], ((it / p.n_stages) & 1) ^ 1);
          if (prof) w_empty += clock64() - t0;
          mbar_arrive_expect_tx(&full_bar[st], kAStage);
          tma_load_2d(ring + st * kAStage, tm, &full_bar[st], k * kBK, row0);
        }
      };
      // Order of the stream == order of the MMA warps: L0(0); then per step L1(s).., P(s), L0(s+1) tile 0, J(s), L0(s+1) tiles 1..
      auto push_l = [&](int m) { push(&tm_wl, static_cast<int>(rank) * 4 * p.up + m * kBM, p.kbHp); };
      for (int m = 0; m < p.mtL; ++m) push_l(m);
      for (int s = 0; s <= p.max_steps; ++s) {
        const long long t0 = prof ? clock64() : 0;
        if (s > 0) { mbar_wait(step_bar, (s - 1) & 1); if (!s_go[(s - 1) & 1]) break; }
        if (prof) w_step += clock64() - t0;
        for (int l = 1; l < NL; ++l)
          for (int m = 0; m < p.mtL; ++m) push(&tm_wu, ((l - 1) * C + static_cast<int>(rank)) * 4 * p.up + m * kBM, 2 * p.kbHp);
        for (int m = 0; m < p.mtP; ++m) push(&tm_wp, static_cast<int>(rank) * p.RP + m * kBM, p.kbHp, m == 0 ? p.res_p : 0);
        if (!p.l_late) push_l(0);
        for (int m = 0; m < p.mtJ; ++m) push(&tm_wj, static_cast<int>(rank) * p.RJ + m * kBM, p.kbH, m == 0 ? p.res_j : 0);
        for (int m = p.l_late ? 0 : 1; m < p.mtL; ++m) push_l(m);
      }
      if (prof) { g_dec_prof[10] = w_empty; g_dec_prof[11] = w_step; }
    }
  } else if (warp == 1 || warp == kMma2Warp) {
    // ------------------------------------------------------------------ MMA issue (two warps)
    // Two warps issue, k-blocks interleaved, each into its own accumulator (columns [0,16) / [16,32) of the tile's TMEM
    // slot; the epilogue adds the two; tfull barriers count two arrivals).  The MMAs themselves are free here (skipping
    // them changes the step by 3 %, `decode_prof` bit 1): what bounds this loop is the weight stream -- TMA latency
    // under load (~3 k cycles) against the bytes the ring keeps in flight -- so the second warp only shortens the
    // wait -> fence -> issue -> commit turnaround of a ring slot (scripts/micro/mma_small.cu: ~107 cycles per MMA for
    // one thread with the ring's bookkeeping, 43 without).  Converged warp + elect.sync throughout: under
    // `if (lane == 0)` ptxas wraps every UTCHMMA in a divergence waterfall (DESIGN.md finding 5.2; 49 -> 35 ms here).
    const int w = warp == 1 ? 0 : 1;
    const uint32_t idesc = make_idesc_bf16(kBM, kNB, false, false);
    const bool prof = p.prof && blockIdx.x == 0 && lane == 0 && w == 0;
    long long w_full = 0, w_dep = 0;
    const uint64_t ad0 = make_smem_desc_sw128(smem_u32(ring), 16, 1024);
    const uint32_t fb0 = smem_u32(&full_bar[0]), eb0 = smem_u32(&empty_bar[0]);
    int st = 0;
    uint32_t ph = 0, fb = fb0, eb = eb0;
    uint64_t ad = ad0;
    // One accumulator tile over kb + kb1 k-blocks: the first kb against activation buffer b_base, the rest against b_base1
    // (upper LSTM layers: h of the layer below, then the layer's own previous h).
    auto tile = [&](uint32_t b_base, int kb0, int slot, uint32_t b_base1 = 0, int kb1 = 0, int res_k = 0, int res_col = 0) {
      const uint64_t bd0 = make_smem_desc_sw128(b_base, 16, 1024);
      const uint64_t bd1 = make_smem_desc_sw128(b_base1, 16, 1024) - static_cast<uint64_t>(kb0) * (kKBlk >> 4);
      const uint32_t d_tmem = tmem + slot * kTileCols + w * kNB, tf = smem_u32(&tfull_bar[slot]);
      const int kb = kb0 + kb1;
      if (res_k > 0) {   // leading k-blocks whose weights are resident in TMEM: A operand from TMEM, no ring slot
        if (elect_one()) {
          for (int k = w; k < res_k; k += 2) {
            const uint64_t bd = bd0 + static_cast<uint64_t>(k) * (kKBlk >> 4);
#pragma unroll
            for (int kk = 0; kk < kBK / 16; ++kk)
              umma_bf16_ts(d_tmem, tmem + res_col + (k * 4 + kk) * 8, bd + 2 * kk, idesc, (k > w || kk != 0) ? 1u : 0u);
          }
        }
        __syncwarp();
      }
      for (int k = res_k; k < kb; ++k) {
        if ((k & 1) == w) {
          {
            const long long t0 = prof ? clock64() : 0;   // try_wait itself suspends the warp: time the whole wait
            mbar_wait_addr(fb, ph);
            if (prof) w_full += clock64() - t0;
          }
          tc_fence_after();
          const uint64_t bd = (k < kb0 ? bd0 : bd1) + static_cast<uint64_t>(k) * (kKBlk >> 4);
          if (elect_one()) {
            if (!(p.prof & 2)) {   // bring-up switch: prof & 2 skips the MMAs (timing experiment, results are garbage)
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk)
                umma_bf16(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k > w || kk != 0) ? 1u : 0u);
            }
            umma_commit_addr(eb);
          }
          __syncwarp();
        }
        if (++st == p.n_stages) { st = 0; ph ^= 1; ad = ad0; fb = fb0; eb = eb0; }
        else { ad += kAStage >> 4; fb += 8; eb += 8; }
      }
      if (elect_one()) umma_commit_addr(tf);   // this warp's share of the tile (possibly empty when kb == 1) is issued
      __syncwarp();
    };
    if (p.res_p || p.res_j) { mbar_wait(res_bar, 0); tc_fence_after(); }
    // W_hh . h(s) -- the product of step s+1's cell -- needs h(s) but not the label emitted at step s (only its epilogue
    // does), so it is issued speculatively inside step s, in the two gaps in which these warps would otherwise wait for an
    // exchange: tile 0 while the hj slices travel, the other tiles while the argmax keys travel.
    for (int m = 0; m < p.mtL; ++m)   // step 0: h(-1) = 0
      tile(smem_u32(hbuf(0, 0)), p.kbHp, m);
    for (int s = 0; s <= p.max_steps; ++s) {
      const int cur = s & 1, nxt = cur ^ 1;
      const uint32_t hpar = (s >> 1) & 1;   // every h buffer is refilled every other step
      long long t0 = prof ? clock64() : 0;
      if (s > 0) { mbar_wait(step_bar, (s - 1) & 1); if (!s_go[(s - 1) & 1]) break; }
      for (int l = 1; l < NL; ++l) {        // upper layers: W_ih . h_{l-1}(s) + W_hh . h_l(s-1)
        mbar_wait(&hfull_bar[2 * (l - 1) + nxt], hpar);
        tc_fence_after();
        for (int m = 0; m < p.mtL; ++m)
          tile(smem_u32(hbuf(l - 1, nxt)), p.kbHp, l * p.mtL + m, smem_u32(hbuf(l, cur)), p.kbHp);
      }
      mbar_wait(&hfull_bar[2 * (NL - 1) + nxt], hpar);
      if (prof) w_dep += clock64() - t0;
      tc_fence_after();
      for (int m = 0; m < p.mtP; ++m)
        tile(smem_u32(hbuf(NL - 1, nxt)), p.kbHp, slotP0 + m, 0, 0, m == 0 ? p.res_p : 0, p.res_col);
      if (!p.l_late) tile(smem_u32(hbuf(0, nxt)), p.kbHp, 0);
      t0 = prof ? clock64() : 0;
      mbar_wait(hjfull_bar, s & 1);
      if (prof) w_dep += clock64() - t0;
      tc_fence_after();
      for (int m = 0; m < p.mtJ; ++m) tile(smem_u32(hj), p.kbH, slotJ0 + m, 0, 0, m == 0 ? p.res_j : 0, p.res_j_col);
      for (int m = p.l_late ? 0 : 1; m < p.mtL; ++m) tile(smem_u32(hbuf(0, nxt)), p.kbHp, m);
    }
    if (prof) { g_dec_prof[12] = w_full; g_dec_prof[13] = w_dep; }
    if (elect_one()) umma_commit(fin_bar);
    __syncwarp();
    mbar_wait(fin_bar, 0);
    tc_fence_after();
  } else {
    // ------------------------------------------------------------------ epilogue warps
    // Eight warps: two per TMEM lane quadrant, each taking one half of the cluster's 16 utterances (an epilogue warp runs
    // alone on its scheduler and is latency-bound at ~5 cycles per instruction, so the work is spread, not vectorised).
    const int quad = warp & 3;
    const int half = (warp - 2) >> 2;            // utterances [8 half, 8 half + 8)
    const int n0 = half * kNH;
    const int row = quad * 32 + lane;            // TMEM lane == weight row inside the tile
    const int et = (warp - 2) * 32 + lane;       // 0..255
    const uint32_t lane_taddr = tmem + (static_cast<uint32_t>(quad * 32) << 16) + n0;
    const size_t gate_pitch = static_cast<size_t>(4) * p.Hp;
    const uint32_t slice_h = static_cast<uint32_t>(p.up / 64) * kKBlk;
    const uint32_t slice_p = static_cast<uint32_t>(p.RP / 64) * kKBlk;

    if ((p.res_p || p.res_j) && et < 128) {
      // Weights that stay in TMEM for the whole decode: lane = output row, column c holds the bf16 pair (k = 2c, 2c + 1),
      // 8 columns per K = 16 MMA.  Rows beyond the matrix and columns beyond K are zero.
      auto load_rows = [&](const __nv_bfloat16* wr, bool row_ok, int K, int n_kb, int col0) {
        for (int c = 0; c < n_kb * 4; ++c) {   // 16 K-elements = 8 columns per store
          uint32_t v[8];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int k0 = c * 16 + q * 8;
            uint4 x = make_uint4(0, 0, 0, 0);
            if (row_ok && k0 < K) x = __ldg(reinterpret_cast<const uint4*>(wr + k0));   // K % 8 == 0
            v[4 * q] = x.x; v[4 * q + 1] = x.y; v[4 * q + 2] = x.z; v[4 * q + 3] = x.w;
          }
          tmem_st8(tmem + (static_cast<uint32_t>(quad * 32) << 16) + col0 + c * 8, v);
        }
      };
      if (p.res_p) {   // the first res_p k-blocks of this CTA's rows of W_proj
        const int n_row = static_cast<int>(rank) * p.RP + row;
        const bool row_ok = row < p.RP && n_row < p.H;
        load_rows(p.w_proj + static_cast<size_t>(row_ok ? n_row : 0) * p.Hp, row_ok, p.Hp, p.res_p, p.res_col);
      }
      if (p.res_j) {   // the first res_j k-blocks of this CTA's first vocabulary tile (on the critical path of every step)
        const int v_row = static_cast<int>(rank) * p.RJ + row;
        const bool row_ok = row < p.RJ && v_row < p.V;
        load_rows(p.w_joint + static_cast<size_t>(row_ok ? v_row : 0) * p.H, row_ok, p.H, p.res_j, p.res_j_col);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(res_bar);
    }
    const bool prof = p.prof && blockIdx.x == 0 && et == 0;
    long long acc[10] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long last = clock64();
    int n_steps = 0;
    for (int s = 0; s <= p.max_steps; ++s) {
      const int nxt = (s + 1) & 1;
      const uint32_t par = s & 1;
      ++n_steps;
      // ---------------------------------------------------------------- L: gates -> cell -> h, layer by layer
      for (int l = 0; l < NL; ++l) {
        float* c_s = cbuf(l);
        __nv_bfloat16* hown = hownbuf(l);
        uint8_t* hdst = hbuf(l, nxt);
        // input half of every gate tile of this layer, fetched before the first accumulator is waited for
        float tva[kMaxLTiles][kNH];
#pragma unroll
        for (int m = 0; m < kMaxLTiles; ++m) {
          if (m < p.mtL) {
            const int u = static_cast<int>(rank) * p.up + m * 32 + lane;   // this thread: gate `quad` of unit u
            if (l == 0) {   // embedding + input half of the cell: one table row per emitted label
#pragma unroll
              for (int i = 0; i < kNH; ++i) {
                const int lab = s_lab[n0 + i];
                tva[m][i] = (lab >= 0 && u < p.Hp) ? __ldg(p.table + static_cast<size_t>(lab) * gate_pitch + quad * p.Hp + u) : 0.0f;
              }
            } else {
              const float bu = u < p.Hp ? __ldg(p.bias_up + static_cast<size_t>(l - 1) * gate_pitch + quad * p.Hp + u) : 0.0f;
#pragma unroll
              for (int i = 0; i < kNH; ++i) tva[m][i] = bu;
            }
          }
        }
        if (l == 0) DSTAMP(0);   // table gather issued
#pragma unroll
        for (int m = 0; m < kMaxLTiles; ++m) {
          if (m >= p.mtL) break;
          const float (&tv)[kNH] = tva[m];
          mbar_wait(&tfull_bar[l * p.mtL + m], par);
          if (l == 0 && m == 0) DSTAMP(1);   // waited for the L accumulator
          tc_fence_after();
          float raw[kNH];
          ld_acc(lane_taddr + (l * p.mtL + m) * kTileCols, l > 0 || p.kbHp > 1, raw);
          float a[kNH];
#pragma unroll
          for (int i = 0; i < kNH; ++i) {
            const float x = raw[i] + tv[i];
            if (p.cell == 0) a[i] = quad == 2 ? tanh_f(x) : sigmoid_f(x);   // LSTM: i, f, g, o
            else a[i] = quad < 2 ? sigmoid_f(x) : x;                        // GRU: r, z, hidden and input halves of n
          }
          float4* gdst = reinterpret_cast<float4*>(gates + ((m * 4 + quad) * 32 + lane) * kNB + n0);
          gdst[0] = make_float4(a[0], a[1], a[2], a[3]);
          gdst[1] = make_float4(a[4], a[5], a[6], a[7]);
        }
        tc_fence_before();
        named_bar_sync(1, kEpiThreads);
        for (int idx = et; idx < p.mtL * 32 * kNB; idx += kEpiThreads) {   // (tile m, unit j, utterance n), n fastest
          const int m = idx >> 9, j = (idx >> 4) & 31, n = idx & 15;
          const int ul = m * 32 + j, n_feat = static_cast<int>(rank) * p.up + ul;
          __nv_bfloat16 hv;
          if (s_lab[n] >= 0) {
            const float* gm = gates + (m * 4 * 32 + j) * kNB + n;
            const float g0 = gm[0], g1 = gm[32 * kNB], g2 = gm[2 * 32 * kNB], g3 = gm[3 * 32 * kNB];
            if (p.cell == 0) {   // LSTM: c' = f c + i g, h' = o tanh(c')
              const float c_new = g1 * c_s[ul * kNB + n] + g0 * g2;
              c_s[ul * kNB + n] = c_new;
              hv = __float2bfloat16_rn(g3 * tanh_f(c_new));
            } else {             // GRU: n = tanh(n_x + r n_h), h' = (1 - z) n + z h; c_s keeps h in fp32
              const float nn = tanh_f(g3 + g0 * g2);
              const float h_new = nn + g1 * (c_s[ul * kNB + n] - nn);
              c_s[ul * kNB + n] = h_new;
              hv = __float2bfloat16_rn(h_new);
            }
            hown[ul * kNB + n] = hv;
          } else {
            hv = hown[ul * kNB + n];
          }
          if (n_feat >= p.Hp) hv = __float2bfloat16(0.0f);
          *reinterpret_cast<__nv_bfloat16*>(hdst + act_offset(n_feat, n)) = hv;
        }
        fence_proxy_async_smem();
        named_bar_sync(1, kEpiThreads);   // also: nobody still reads `gates` when the next layer overwrites it
        {   // one bulk copy per peer, issued by C - 1 different threads; the own barrier expects the peers' slices
          uint64_t* hf = &hfull_bar[2 * l + nxt];
          if (et == 0) mbar_arrive_expect_tx(hf, static_cast<uint32_t>(C - 1) * slice_h);
          if (et >= 32 && et < 31 + C) {
            const uint32_t src = smem_u32(hdst) + rank * slice_h, dst = (rank + (et - 31)) % C;
            bulk_copy_to_cluster(mapa_u32(src, dst), src, slice_h, mapa_u32(smem_u32(hf), dst));
          }
        }
      }
      DSTAMP(2);  // L epilogue + push issued
      // ---------------------------------------------------------------- P: g -> hj = tanh(f + g)
      for (int m = 0; m < p.mtP; ++m) {
        const int lr = m * kBM + row, n_feat = static_cast<int>(rank) * p.RP + lr;
        const bool mine = lr < p.RP, valid = mine && n_feat < p.H;
        float fv[kNH];
#pragma unroll
        for (int i = 0; i < kNH; ++i) {
          const int b = b0 + n0 + i;
          int t = s_t[n0 + i];
          t = t < p.Tmax ? t : p.Tmax - 1;
          fv[i] = (valid && b < p.B)
                      ? __bfloat162float(__ldg(p.f + (static_cast<size_t>(b) * p.Tmax + t) * p.H + n_feat)) : 0.0f;
        }
        const float bp = (valid && p.bias_p) ? __ldg(p.bias_p + n_feat) : 0.0f;
        mbar_wait(&tfull_bar[slotP0 + m], par);
        if (m == p.mtP - 1) DSTAMP(3);  // waited for the P accumulator (h exchange + P product)
        tc_fence_after();
        float raw[kNH];
        ld_acc(lane_taddr + (slotP0 + m) * kTileCols, p.kbHp > 1, raw);
        if (mine) {
#pragma unroll
          for (int i = 0; i < kNH; ++i) {
            const float g = bf16_round_f(raw[i] + bp);
            const float hv = valid ? tanh_approx(fv[i] + g) : 0.0f;
            *reinterpret_cast<__nv_bfloat16*>(hj + act_offset(n_feat, n0 + i)) = __float2bfloat16_rn(hv);
          }
        }
      }
      tc_fence_before();
      fence_proxy_async_smem();
      named_bar_sync(1, kEpiThreads);
      if (et == 0) mbar_arrive_expect_tx(hjfull_bar, static_cast<uint32_t>(C - 1) * slice_p);
      if (et >= 32 && et < 31 + C) {
        const uint32_t src = smem_u32(hj) + rank * slice_p, dst = (rank + (et - 31)) % C;
        bulk_copy_to_cluster(mapa_u32(src, dst), src, slice_p, mapa_u32(smem_u32(hjfull_bar), dst));
      }
      DSTAMP(4);  // P epilogue + push issued
      // ---------------------------------------------------------------- J: logits -> argmax
      unsigned long long best = 0ull;  // lane i < 8 keeps the best key of utterance n0 + i over this warp's rows
      for (int m = 0; m < p.mtJ; ++m) {
        const int lr = m * kBM + row, v = static_cast<int>(rank) * p.RJ + lr;
        const bool valid = lr < p.RJ && v < p.V;
        const float bj = (valid && p.bias_j) ? __ldg(p.bias_j + v) : 0.0f;
        mbar_wait(&tfull_bar[slotJ0 + m], par);
        if (m == p.mtJ - 1) DSTAMP(5);  // waited for the J accumulator (hj exchange + J product)
        tc_fence_after();
        float raw[kNH];
        ld_acc(lane_taddr + (slotJ0 + m) * kTileCols, p.kbH > 1, raw);
#pragma unroll
        for (int i = 0; i < kNH; ++i) {
          const float z = raw[i] + bj;
          const uint32_t ob = (valid && z == z) ? ordered_bits(z) : 0u;
          const uint32_t mx = __reduce_max_sync(0xffffffffu, ob);
          const uint32_t mv = __reduce_min_sync(0xffffffffu, (ob == mx && ob != 0u) ? static_cast<uint32_t>(v) : 0xFFFFFFFFu);
          const unsigned long long key = mx ? ((static_cast<unsigned long long>(mx) << 32) | (0xFFFFFFFFu - mv)) : 0ull;
          if (lane == i && key > best) best = key;
        }
      }
      if (lane < kNH) part[quad * kNB + n0 + lane] = best;
      tc_fence_before();
      named_bar_sync(1, kEpiThreads);
      if (et < kNB) {
        unsigned long long key = part[et];
#pragma unroll
        for (int q = 1; q < 4; ++q) { const unsigned long long o = part[q * kNB + et]; key = o > key ? o : key; }
        const uint32_t slot = smem_u32(amax_s + rank * kNB + et), bar = smem_u32(amaxfull_bar);
        for (int d = 0; d < C; ++d) st_async_u64(mapa_u32(slot, d), key, mapa_u32(bar, d));
      }
      if (et == 0) mbar_arrive_expect_tx(amaxfull_bar, static_cast<uint32_t>(C) * kNB * 8);
      DSTAMP(6);  // J epilogue + keys sent
      mbar_wait(amaxfull_bar, par);
      DSTAMP(7);  // waited for everybody's keys
      // ---------------------------------------------------------------- bookkeeping (replicated in every CTA)
      if (warp == 2) {
        int active = 0;
        if (lane < kNB) {
          unsigned long long key = 0ull;
          for (int r = 0; r < C; ++r) { const unsigned long long o = amax_s[r * kNB + lane]; key = o > key ? o : key; }
          int t = s_t[lane], lab = -1;
          if (t < s_len[lane]) {
            const int k = key ? static_cast<int>(0xFFFFFFFFu - static_cast<uint32_t>(key)) : p.blank;
            int em = s_em[lane];
            if (k != p.blank) {
              const int n = s_n[lane];
              if (rank == 0 && n < p.sym_cap) p.sym[static_cast<size_t>(b0 + lane) * p.sym_cap + n] = k;
              s_n[lane] = n + 1;
              ++em;
              lab = k;
            }
            if (lab < 0 || em >= p.S) { ++t; em = 0; }
            s_t[lane] = t; s_em[lane] = em;
          }
          s_lab[lane] = lab;
          active = t < s_len[lane] ? 1 : 0;
        }
        const unsigned any = __ballot_sync(0xffffffffu, active);
        if (lane == 0) s_go[par] = any != 0u;
      }
      named_bar_sync(1, kEpiThreads);
      if (et == 0) mbar_arrive(step_bar);
      DSTAMP(8);  // bookkeeping
      if (!s_go[par]) break;
    }
    if (prof) {
      for (int i = 0; i < 9; ++i) g_dec_prof[i] = static_cast<unsigned long long>(acc[i]);
      g_dec_prof[9] = static_cast<unsigned long long>(n_steps);
    }
    if (rank == 0 && et < kNB && b0 + et < p.B) p.n_sym[b0 + et] = s_n[et] < p.sym_cap ? s_n[et] : p.sym_cap;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // nobody leaves while a peer could still address its shared memory
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, p.tmem_cols);
  }
}

// W_hh [4 Hp][Hp] (torch gate order) -> rows [rank][tile m][gate][32 units]: CTA `rank` owns units
// [rank*up, rank*up + up), one 128-row tile carries the four gates of 32 units, one gate per TMEM lane quadrant.
__global__ void permute_whh_cluster_kernel(const __nv_bfloat16* __restrict__ W, __nv_bfloat16* __restrict__ out, int Hp,
                                           int up, int n_rows) {
  const int R = blockIdx.x;
  if (R >= n_rows) return;
  const int per_rank = 4 * up;
  const int rank = R / per_rank, rem = R - rank * per_rank;
  const int m = rem >> 7, gate = (rem >> 5) & 3, j = rem & 31;
  const int u = rank * up + m * 32 + j;
  __nv_bfloat16* dst = out + static_cast<size_t>(R) * Hp;
  if (u < Hp) {
    const __nv_bfloat16* src = W + (static_cast<size_t>(gate) * Hp + u) * Hp;
    for (int i = threadIdx.x; i < Hp; i += blockDim.x) dst[i] = src[i];
  } else {
    for (int i = threadIdx.x; i < Hp; i += blockDim.x) dst[i] = __float2bfloat16(0.0f);
  }
}

// Upper LSTM layers: W_upper [NL-1][4 Hp][2 Hp] = [W_ih_l | W_hh_l] (torch layouts) -> rows [l-1][rank][tile][gate][32 units] as
// above, columns [W_ih_l | 0-pad to kb*64 | W_hh_l | 0-pad to kb*64] so that both halves start on a k-block boundary.
__global__ void permute_wup_cluster_kernel(const __nv_bfloat16* __restrict__ W, __nv_bfloat16* __restrict__ out, int Hp,
                                           int up, int C, int kb, int n_rows) {
  const int R = blockIdx.x;
  if (R >= n_rows) return;
  const int per_rank = 4 * up, per_layer = C * per_rank;
  const int l1 = R / per_layer, r2 = R - l1 * per_layer;
  const int rank = r2 / per_rank, rem = r2 - rank * per_rank;
  const int m = rem >> 7, gate = (rem >> 5) & 3, j = rem & 31;
  const int u = rank * up + m * 32 + j;
  const int half_cols = kb * 64;
  __nv_bfloat16* dst = out + static_cast<size_t>(R) * 2 * half_cols;
  const __nv_bfloat16* src = W + (static_cast<size_t>(l1) * 4 * Hp + static_cast<size_t>(gate) * Hp + u) * 2 * Hp;
  for (int c = threadIdx.x; c < 2 * half_cols; c += blockDim.x) {
    const int half = c / half_cols, cc = c - half * half_cols;
    dst[c] = (u < Hp && cc < Hp) ? src[half * Hp + cc] : __float2bfloat16(0.0f);
  }
}

namespace {
bool g_decode_cooperative = true;
}
void set_decode_cooperative(int v) { g_decode_cooperative = v != 0; }

void launch_permute_whh(const __nv_bfloat16* W, __nv_bfloat16* out, int Hp, int nu, int n_rows, cudaStream_t s) {
  permute_whh_kernel<<<n_rows, 128, 0, s>>>(W, out, Hp, nu, n_rows);
}

int max_ctas_greedy_decode(int smem_bytes) {
  int dev = 0, sms = 0, per_sm = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) return 0;
  if (cudaFuncSetAttribute(greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, greedy_decode_kernel, kDecThreads, smem_bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return per_sm > 0 ? sms : 0;  // one CTA per SM
}

cudaError_t launch_greedy_decode(const CUtensorMap& tm_hj, const CUtensorMap& tm_hbuf, const CUtensorMap& tm_wj,
                                 const CUtensorMap& tm_wl, const CUtensorMap& tm_wp, const DecodeArgs& a, int n_ctas,
                                 int smem_bytes, cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(greedy_decode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(kDecThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // every CTA waits on every other one: co-residency is required
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_decode_cooperative ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, greedy_decode_kernel, tm_hj, tm_hbuf, tm_wj, tm_wl, tm_wp, a);
}


void launch_permute_whh_cluster(const __nv_bfloat16* W, __nv_bfloat16* out, int Hp, int up, int n_rows, cudaStream_t s) {
  permute_whh_cluster_kernel<<<n_rows, 128, 0, s>>>(W, out, Hp, up, n_rows);
}

void launch_permute_wup_cluster(const __nv_bfloat16* W, __nv_bfloat16* out, int Hp, int up, int C, int kb, int n_rows,
                                cudaStream_t s) {
  permute_wup_cluster_kernel<<<n_rows, 128, 0, s>>>(W, out, Hp, up, C, kb, n_rows);
}

int read_decode_prof(unsigned long long* out, int n) {
  if (n > 16) n = 16;
  return cudaMemcpyFromSymbol(out, g_dec_prof, sizeof(unsigned long long) * n) == cudaSuccess ? n : -1;
}

int max_clusters_greedy_decode(int smem_bytes, int C) {
  if (cudaFuncSetAttribute(greedy_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  if (C > 8) cudaFuncSetAttribute(greedy_decode_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C);
  cfg.blockDim = dim3(kCThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, greedy_decode_cluster_kernel, &cfg) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

cudaError_t launch_greedy_decode_cluster(const CUtensorMap& tm_wj, const CUtensorMap& tm_wl, const CUtensorMap& tm_wu,
                                         const CUtensorMap& tm_wp, const ClusterDecodeArgs& a, int n_clusters, int smem_bytes,
                                         cudaStream_t s) {
  cudaError_t e = cudaFuncSetAttribute(greedy_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  if (e != cudaSuccess) return e;
  if (a.C > 8) cudaFuncSetAttribute(greedy_decode_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_clusters * a.C);
  cfg.blockDim = dim3(kCThreads);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = a.C; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, greedy_decode_cluster_kernel, tm_wj, tm_wl, tm_wu, tm_wp, a);
}

}  // namespace rnnt
