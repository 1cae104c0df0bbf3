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
constexpr int kMaxStages = 8;

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
__device__ __forceinline__ uint32_t ordered_bits(float x) {
  const uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float sigmoid_f(float x) { return 1.0f / (1.0f + expf(-x)); }
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
      if (lane == 0) {
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
                    const float c_new = sigmoid_f(af) * c_s[b * p.nu + ul] + sigmoid_f(ai) * tanhf(ag);
                    c_s[b * p.nu + ul] = c_new;
                    h_s[b * p.nu + ul] = __float2bfloat16_rn(sigmoid_f(ao) * tanhf(c_new));
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

}  // namespace rnnt
