// RNN-T joint network kernels for sm_100a.
//
//   hgen      h = bf16(tanh(f_t + g_u))                          (MUFU-bound elementwise, L2-resident slab)
//   slab_gemm D = A . B^T on tcgen05 with TMEM accumulators, TMA-fed 128B-swizzled operands, and one of
//             three fused epilogues:
//               kFwd : logits -> online log-softmax; keeps lse, lp_blank, lp_label (no BTUV tensor)
//               kDz  : logits recompute -> dz = softmax*c0 - [blank]c1 - [label]c2 -> bf16 slab (+ db)
//               kDh  : dh = dz . W -> dpre = dh*(1-h^2) -> tile-reduced red.add into df, dg
//   dw        dW += dz^T . h   (both operands MN-major straight from the row-major slabs, split-K)
//
// Warp roles in the tcgen05 kernels (192 threads): warp 0 = TMA producer, warp 1 = TMEM allocator +
// single-thread MMA issuer, warps 2..5 = epilogue (one TMEM lane quadrant each).  The N extent is walked in
// chunks of <=256 columns; TMEM holds two chunk accumulators so the epilogue of chunk j overlaps the MMAs
// of chunk j+1.
//
// Replaces (by analogy) model/fully_connected.py:133-166 (Linear over (x, lens)),
// loss/ctc_loss.py:45,95 (LogSoftmax inside the loss) and their autograd.
#include <math.h>

#include "launch.h"
#include "ptx.cuh"

namespace rnnt {

namespace {

constexpr int kBM = 128;                    // rows per tile == TMEM lanes
constexpr int kBK = 64;                     // bf16 per k-block (one 128-byte swizzle row)
constexpr int kNCmax = 256;                 // max columns per chunk (UMMA N)
constexpr int kAStage = kBM * kBK * 2;            // 16 KB: this CTA's 128 rows
constexpr int kBStage = (kNCmax / 2) * kBK * 2;   // 16 KB: this CTA's half of the chunk's N rows (CTA pair)
constexpr int kStageBytes = kAStage + kBStage;
constexpr int kThreads = 192;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kMaxBiasCols = 8192;          // V (rounded up to chunks) supported by the smem bias table (32 KB)
constexpr int kDhPitch = 68;                // fp32 pitch of the dpre tile (16-byte aligned rows, conflict-free 128-bit access)

enum Epi { kFwd = 0, kDz = 1, kDh = 2 };
int g_gemm_dbg = 0;

// bring-up profiling (gemm_dbg & 4): %globaltimer stamps of the last slab_gemm launch, 8 per CTA
__device__ unsigned long long g_prof[160 * 8];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PROF(slot) do { if (p.dbg & 4) g_prof[blockIdx.x * 8 + (slot)] = gtimer(); } while (0)

template <int EPI> struct Cfg;
template <> struct Cfg<kFwd> { static constexpr int stages = 6; static constexpr int extra = kMaxBiasCols * 4; };
template <> struct Cfg<kDz>  { static constexpr int stages = 4; static constexpr int extra = kMaxBiasCols * 4 + kBM * kNCmax * 2; };
template <> struct Cfg<kDh>  { static constexpr int stages = 4; static constexpr int extra = 2 * kBM * kDhPitch * 4; };

template <int EPI> constexpr int smem_total() { return Cfg<EPI>::stages * kStageBytes + Cfg<EPI>::extra + 1024 + 256; }

struct GemmArgs {
  Lattice L;
  int tile0;      // first global tile handled by this launch; CTA x handles tile0 + x, slab rows [128x, 128x+128)
  int n_tiles_total;  // tiles in the whole batch; a CTA whose tile is beyond it is the "ghost" half of an odd pair
  int dbg;            // bring-up switches (0 in production): 1 = skip epilogue work, 2 = skip TMA after the first ring fill
  int n_total;    // valid N extent (V or H)
  int nc;         // columns per chunk, multiple of 32, <= 256
  int n_chunks;
  int k_blocks;
  int blank;
  int Umax;
  int Vp;
  int H;
  // kFwd / kDz
  const float* bias;
  const int* y;
  float* lse_tile;
  float* lpb;
  float* lpl;
  const float* c1;
  const float* c2;
  const float* grad_loss;
  float* db;
  // kDh
  const __nv_bfloat16* hslab;
  float* df;
  float* dg;
};

__device__ __forceinline__ float pick32(const float (&v)[32], int idx) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < 32; ++i) r = (idx == i) ? v[i] : r;
  return r;
}

template <int EPI>
__global__ void __launch_bounds__(kThreads, 1)
slab_gemm_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                 const __grid_constant__ CUtensorMap tm_out, const GemmArgs p) {
  constexpr int kStages = Cfg<EPI>::stages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint8_t* extra = smem + kStages * kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(extra + Cfg<EPI>::extra);
  uint64_t* full_bar = bars;                    // [kStages]
  uint64_t* empty_bar = bars + kStages;         // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;     // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  TileInfo* s_ti = reinterpret_cast<TileInfo*>(tmem_slot + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m = blockIdx.x;
  // CTA pair: the even CTA ("leader") issues every MMA for both (M = 256: 128 rows from each CTA's smem, each
  // CTA supplies half of the chunk's N rows of B); accumulators land in each CTA's own TMEM.
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const bool ghost = (p.tile0 + m) >= p.n_tiles_total;

  if (threadIdx.x == 0) {
    PROF(0);
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    if (EPI == kDz) prefetch_tmap(&tm_out);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    fence_barrier_init();
    *s_ti = decode_tile(p.L, p.tile0 + m);
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const TileInfo ti = *s_ti;
  if (threadIdx.x == 0) PROF(1);

  const int nc_half = p.nc >> 1;
  const uint32_t b_bytes = static_cast<uint32_t>(nc_half) * kBK * 2;

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int it = 0;
      for (int j = 0; j < p.n_chunks; ++j) {
        for (int k = 0; k < p.k_blocks; ++k, ++it) {
          const int s = it % kStages;
          const uint32_t ph = (it / kStages) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          if ((p.dbg & 2) && it >= kStages) { if (leader) mbar_arrive(&full_bar[s]); continue; }
          // the leader's barrier collects the bytes of both CTAs' loads
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * (kAStage + b_bytes));
          uint8_t* sa = stage_base + s * kStageBytes;
          tma_load_2d_pair(sa, &tm_a, &full_bar[s], k * kBK, m * kBM);
          tma_load_2d_pair(sa + kAStage, &tm_b, &full_bar[s], k * kBK, j * p.nc + static_cast<int>(rank) * nc_half);
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -----------------
    const uint32_t idesc = make_idesc_bf16(2 * kBM, p.nc, false, false);
    int it = 0;
    for (int j = 0; leader && j < p.n_chunks; ++j) {
      const int buf = j & 1;
      mbar_wait(&tempty_bar[buf], ((j >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * kNCmax;
      for (int k = 0; k < p.k_blocks; ++k, ++it) {
        const int s = it % kStages;
        const uint32_t ph = (it / kStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          if (it == 0) PROF(2);
          const uint32_t a_addr = smem_u32(stage_base + s * kStageBytes);
          const uint32_t b_addr = a_addr + kAStage;
#pragma unroll
          for (int kk = 0; kk < kBK / 16; ++kk) {
            const uint64_t ad = make_smem_desc_sw128(a_addr + kk * 32, 16, 1024);
            const uint64_t bd = make_smem_desc_sw128(b_addr + kk * 32, 16, 1024);
            umma_bf16_pair(d_tmem, ad, bd, idesc, (k | kk) != 0 ? 1u : 0u);
          }
          umma_commit_pair(&empty_bar[s], 3);                           // frees the slot in both CTAs
          if (k == p.k_blocks - 1) umma_commit_pair(&tfull_bar[buf], 3);  // accumulator ready in both CTAs
        }
        __syncwarp();
      }
    }
    if (lane == 0) PROF(3);
  } else {
    // ------------------------------- epilogue ------------------------------------
    const int quad = warp & 3;              // TMEM lane quadrant this warp may read
    const int r = quad * 32 + lane;         // tile row == TMEM lane
    const int et = (warp - 2) * 32 + lane;  // epilogue thread id 0..127
    const int dt = r >> 3, du = r & 7;
    const int t = ti.t0 + dt, u = ti.u0 + du;
    const bool valid = !ghost && (t < ti.T) && (u <= ti.U);
    const size_t grow = static_cast<size_t>(p.tile0 + m) * kBM + r;
    const size_t didx = valid ? diag_index(p.L, ti.b, t, u) : 0;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);

    if constexpr (EPI == kFwd || EPI == kDz) {
      float* sbias = reinterpret_cast<float*>(extra);
      const int ncols = p.n_chunks * p.nc;
      for (int c = et; c < ncols; c += kEpiThreads)
        sbias[c] = (c < p.n_total) ? (p.bias ? p.bias[c] * kLog2e : 0.0f) : -INFINITY;
      if constexpr (EPI == kDz) {
        uint4* st = reinterpret_cast<uint4*>(extra + kMaxBiasCols * 4);
        for (int i = et; i < kBM * kNCmax * 2 / 16; i += kEpiThreads) st[i] = make_uint4(0, 0, 0, 0);
      }
      named_bar_sync(1, kEpiThreads);
      const int label = (valid && u < ti.U) ? p.y[static_cast<size_t>(ti.b) * p.Umax + u] : -1;

      if constexpr (EPI == kFwd) {
        float mx = -INFINITY, sum = 0.0f, zb = 0.0f, zl = 0.0f;
        for (int j = 0; j < p.n_chunks; ++j) {
          const int buf = j & 1;
          mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
          tc_fence_after();
          if (et == 0 && j == 0) PROF(4);
          if (et == 0 && j == p.n_chunks - 1) PROF(5);
          for (int g = 0; g < ((p.dbg & 1) ? 0 : p.nc / 32); ++g) {
            uint32_t raw[32];
            tmem_ld32(lane_taddr + buf * kNCmax + g * 32, raw);
            tmem_ld_wait();
            const int c0 = j * p.nc + g * 32;
            float v[32];
            const float4* bp = reinterpret_cast<const float4*>(sbias + c0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = bp[q];
              v[4 * q + 0] = fmaf(__uint_as_float(raw[4 * q + 0]), kLog2e, bb.x);
              v[4 * q + 1] = fmaf(__uint_as_float(raw[4 * q + 1]), kLog2e, bb.y);
              v[4 * q + 2] = fmaf(__uint_as_float(raw[4 * q + 2]), kLog2e, bb.z);
              v[4 * q + 3] = fmaf(__uint_as_float(raw[4 * q + 3]), kLog2e, bb.w);
            }
            float gm = v[0];
#pragma unroll
            for (int i = 1; i < 32; ++i) gm = fmaxf(gm, v[i]);
            const float mn = fmaxf(mx, gm);
            sum *= ex2f(mx - mn);
#pragma unroll
            for (int i = 0; i < 32; ++i) sum += ex2f(v[i] - mn);
            mx = mn;
            if (static_cast<unsigned>(p.blank - c0) < 32u) zb = pick32(v, p.blank - c0);
            if (static_cast<unsigned>(label - c0) < 32u) zl = pick32(v, label - c0);
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_even_cta(&tempty_bar[buf]);
        }
        const float lse2 = mx + lg2f(sum);
        if (!ghost) p.lse_tile[grow] = valid ? lse2 * kLn2 : 0.0f;
        if (valid) {
          p.lpb[didx] = (zb - lse2) * kLn2;
          p.lpl[didx] = (u < ti.U) ? (zl - lse2) * kLn2 : kNeg;
        }
        if (et == 0) PROF(6);
      } else {
        // ---- kDz ----
        uint8_t* stage_out = extra + kMaxBiasCols * 4;
        float c1g = 0.0f, c2g = 0.0f, lse2 = 1.0e30f, lpb_r = 0.0f, lpl_r = 0.0f;
        if (valid) {
          const float gl = p.grad_loss[ti.b];
          c1g = p.c1[didx] * gl;
          c2g = p.c2[didx] * gl;
          lse2 = p.lse_tile[grow] * kLog2e;
          lpb_r = p.lpb[didx];
          lpl_r = (u < ti.U) ? p.lpl[didx] : 0.0f;
        }
        const float c0g = c1g + c2g;
        const int n_box = (p.nc + 63) / 64;
        for (int j = 0; j < p.n_chunks; ++j) {
          const int buf = j & 1;
          mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
          tc_fence_after();
          for (int g = 0; g < p.nc / 32; ++g) {
            uint32_t raw[32];
            tmem_ld32(lane_taddr + buf * kNCmax + g * 32, raw);
            tmem_ld_wait();
            const int c0 = j * p.nc + g * 32;
            const float4* bp = reinterpret_cast<const float4*>(sbias + c0);
            uint32_t pk[16];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 bb = bp[q];
              const float d0 = ex2f(fmaf(__uint_as_float(raw[4 * q + 0]), kLog2e, bb.x) - lse2) * c0g;
              const float d1 = ex2f(fmaf(__uint_as_float(raw[4 * q + 1]), kLog2e, bb.y) - lse2) * c0g;
              const float d2 = ex2f(fmaf(__uint_as_float(raw[4 * q + 2]), kLog2e, bb.z) - lse2) * c0g;
              const float d3 = ex2f(fmaf(__uint_as_float(raw[4 * q + 3]), kLog2e, bb.w) - lse2) * c0g;
              pk[2 * q + 0] = pack_bf16x2(d0, d1);
              pk[2 * q + 1] = pack_bf16x2(d2, d3);
            }
            // staging layout == TMA 128B-swizzled boxes of [128 rows x 64 cols]
            uint8_t* box = stage_out + (g >> 1) * (kBM * 128) + r * 128;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int chunk16 = (g & 1) * 4 + q;
              *reinterpret_cast<uint4*>(box + ((chunk16 ^ (r & 7)) << 4)) =
                  make_uint4(pk[4 * q + 0], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_even_cta(&tempty_bar[buf]);
          // exact values for the two special columns of this row (avoids a bf16 read-modify-write)
          {
            const int cb = p.blank - j * p.nc;
            if (static_cast<unsigned>(cb) < static_cast<unsigned>(p.nc)) {
              const float dv = c0g * ex2f(lpb_r * kLog2e) - c1g;
              uint8_t* a = stage_out + (cb >> 6) * (kBM * 128) + r * 128 + ((((cb & 63) >> 3) ^ (r & 7)) << 4) + (cb & 7) * 2;
              *reinterpret_cast<__nv_bfloat16*>(a) = __float2bfloat16_rn(dv);
            }
            const int cl = label - j * p.nc;
            if (label >= 0 && static_cast<unsigned>(cl) < static_cast<unsigned>(p.nc)) {
              const float dv = c0g * ex2f(lpl_r * kLog2e) - c2g;
              uint8_t* a = stage_out + (cl >> 6) * (kBM * 128) + r * 128 + ((((cl & 63) >> 3) ^ (r & 7)) << 4) + (cl & 7) * 2;
              *reinterpret_cast<__nv_bfloat16*>(a) = __float2bfloat16_rn(dv);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1, kEpiThreads);
          if (et == 0 && !ghost) {
            for (int bx = 0; bx < n_box; ++bx) {
              const int col = j * p.nc + bx * 64;
              if (col < p.Vp) tma_store_2d(&tm_out, stage_out + bx * (kBM * 128), col, m * kBM);
            }
            tma_store_commit();
          }
          // db: column sums of the staged tile (thread et owns columns 2et, 2et+1 of the chunk)
          {
            const int cc = 2 * et;
            if (cc < p.nc && !ghost) {
              float s0 = 0.0f, s1 = 0.0f;
              const uint8_t* colp = stage_out + (cc >> 6) * (kBM * 128) + (cc & 7) * 2;
              const int ch = (cc & 63) >> 3;
#pragma unroll 8
              for (int rr = 0; rr < kBM; ++rr) {
                const uint32_t w = *reinterpret_cast<const uint32_t*>(colp + rr * 128 + ((ch ^ (rr & 7)) << 4));
                s0 += bf16lo(w);
                s1 += bf16hi(w);
              }
              const int gc = j * p.nc + cc;
              if (gc < p.n_total) red_add_f32(p.db + gc, s0);
              if (gc + 1 < p.n_total) red_add_f32(p.db + gc + 1, s1);
            }
          }
          if (et == 0) tma_store_wait_read0();
          named_bar_sync(1, kEpiThreads);
        }
        if (et == 0) tma_store_wait_all0();
      }
    } else {
      // ---- kDh ----
      // Per 64-column sub-chunk: dpre = dh * (1 - h^2) is written to a double-buffered fp32 smem tile
      // (row = lattice cell), then reduced over the 8 label positions (-> df) and the 16 frames (-> dg) of
      // the tile with 128-bit smem reads and red.global.add.v4.  h for the NEXT sub-chunk is fetched before
      // the current one is processed, so its L2 latency hides behind the arithmetic.
      float* tile_base = reinterpret_cast<float*>(extra);
      const __nv_bfloat16* hrow = p.hslab + (static_cast<size_t>(m) * kBM + r) * p.H;
      const int n_sub = (p.nc + 63) / 64;
      const int total_sub = p.n_chunks * n_sub;
      auto load_h = [&](int s_idx, uint4 (&hv)[8]) {
        const int jj = s_idx / n_sub, ss = s_idx - jj * n_sub;
#pragma unroll
        for (int gg = 0; gg < 2; ++gg) {
          const int g = ss * 2 + gg;
          const int c0 = jj * p.nc + g * 32;
          const uint4* hp = reinterpret_cast<const uint4*>(hrow + c0);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            hv[gg * 4 + q] = (g * 32 < p.nc && c0 + 8 * q < p.H) ? __ldg(hp + q) : make_uint4(0, 0, 0, 0);
        }
      };
      uint4 hcur[8];
      load_h(0, hcur);
      int s_idx = 0;
      for (int j = 0; j < p.n_chunks; ++j) {
        const int buf = j & 1;
        mbar_wait(&tfull_bar[buf], (j >> 1) & 1);
        tc_fence_after();
        for (int sub = 0; sub < n_sub; ++sub, ++s_idx) {
          uint4 hnext[8];
          if (s_idx + 1 < total_sub) {
            load_h(s_idx + 1, hnext);
          } else {
#pragma unroll
            for (int q = 0; q < 8; ++q) hnext[q] = make_uint4(0, 0, 0, 0);
          }
          float* tile = tile_base + (s_idx & 1) * (kBM * kDhPitch);
#pragma unroll
          for (int gg = 0; gg < 2; ++gg) {
            const int g = sub * 2 + gg;
            const int c0 = j * p.nc + g * 32;
            float4* trow = reinterpret_cast<float4*>(tile + r * kDhPitch + gg * 32);
            if (g * 32 < p.nc && c0 < p.H) {
              uint32_t raw[32];
              tmem_ld32(lane_taddr + buf * kNCmax + g * 32, raw);
              tmem_ld_wait();
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint4 hq = hcur[gg * 4 + q];
                const uint32_t w[4] = {hq.x, hq.y, hq.z, hq.w};
                float o[8];
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  const float h0 = bf16lo(w[e]), h1 = bf16hi(w[e]);
                  const float d0 = __uint_as_float(raw[8 * q + 2 * e]);
                  const float d1 = __uint_as_float(raw[8 * q + 2 * e + 1]);
                  o[2 * e] = fmaf(-h0 * h0, d0, d0);
                  o[2 * e + 1] = fmaf(-h1 * h1, d1, d1);
                }
                trow[2 * q] = make_float4(o[0], o[1], o[2], o[3]);
                trow[2 * q + 1] = make_float4(o[4], o[5], o[6], o[7]);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 8; ++i) trow[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
          }
          if (sub == n_sub - 1) {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive_even_cta(&tempty_bar[buf]);
          }
          named_bar_sync(1, kEpiThreads);  // tile[s_idx & 1] complete; also fences the buffer written two subs ago
          const int c4 = et & 15;
          const int colbase = j * p.nc + sub * 64 + 4 * c4;
          if (!ghost && sub * 64 + 4 * c4 < p.nc && colbase < p.H) {  // a ghost CTA's rows are uninitialised memory
            const float* tcol = tile + 4 * c4;
#pragma unroll
            for (int k = 0; k < 2; ++k) {  // df: sum over the 8 label positions of frame a
              const int a = (et >> 4) + 8 * k;
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int c = 0; c < kTU; ++c) {
                const float4 v = *reinterpret_cast<const float4*>(tcol + (a * kTU + c) * kDhPitch);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
              }
              if (ti.t0 + a < ti.T)
                red_add_v4_f32(p.df + (static_cast<size_t>(ti.b) * p.L.Tmax + ti.t0 + a) * p.H + colbase, acc.x, acc.y,
                               acc.z, acc.w);
            }
            {  // dg: sum over the 16 frames of label position c
              const int c = et >> 4;
              float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
              for (int a = 0; a < kTT; ++a) {
                const float4 v = *reinterpret_cast<const float4*>(tcol + (a * kTU + c) * kDhPitch);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
              }
              if (ti.u0 + c <= ti.U)
                red_add_v4_f32(p.dg + (static_cast<size_t>(ti.b) * p.L.U1max + ti.u0 + c) * p.H + colbase, acc.x, acc.y,
                               acc.z, acc.w);
            }
          }
#pragma unroll
          for (int q = 0; q < 8; ++q) hcur[q] = hnext[q];
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers until here
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
    if (lane == 0) PROF(7);
  }
}

// ------------------------------------------------------------------------------------------------
// dW += dz^T . h  : M = V (256 per CTA pair, 128 per CTA), N = H (256 per tile), K = slab rows (64 per k-block).
// Both operands are read MN-major straight from the row-major slabs ([rows][Vp] and [rows][H]) as
// 128B-swizzled TMA boxes of 64 elements x 64 rows.  Work items (output tile, k-block) are dealt out evenly
// to the CTA pairs (split-K); each run of items on one output tile ends with a red.add flush into dW.
// ------------------------------------------------------------------------------------------------
constexpr int kDwStages = 6;
constexpr int kDwBox = 64 * 128;  // 8 KB: 64 rows x 128 B
constexpr int kDwAStage = 2 * kDwBox;   // this CTA's 128 of the pair's 256 V rows
constexpr int kDwBStage = 2 * kDwBox;   // this CTA's 128 of the tile's 256 H columns
constexpr int kDwStageBytes = kDwAStage + kDwBStage;
constexpr int kDwSmem = kDwStages * kDwStageBytes + 1024 + 256;

struct DwArgs {
  float* dW;
  int V, H;
  int n_vt, n_ht;   // output tiles along V (256 per pair) and H (256)
  int nkb;          // k-blocks (64 slab rows) in this slab
  int per_pair;     // work items (tile, k-block) per CTA pair
  int total;        // n_vt * n_ht * nkb
};

__global__ void __launch_bounds__(kThreads, 1)
dw_kernel(const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_h, const DwArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDwStages * kDwStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + kDwStages;
  uint64_t* tfull_bar = bars + 2 * kDwStages;
  uint64_t* tempty_bar = bars + 2 * kDwStages + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kDwStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int begin = pair * p.per_pair;
  const int end = min(begin + p.per_pair, p.total);

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_dz);
    prefetch_tmap(&tm_h);
    for (int i = 0; i < kDwStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2cta(tmem_slot, kTmemCols);
    tmem_relinquish_2cta();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int item = begin; item < end; ++item, ++it) {
        const int tile = item / p.nkb, kb = item - tile * p.nkb;
        const int vm = tile / p.n_ht, hn = tile - vm * p.n_ht;
        const int s = it % kDwStages;
        const uint32_t ph = (it / kDwStages) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * kDwStageBytes);
        uint8_t* sa = smem + s * kDwStageBytes;
        const int v0 = vm * 256 + static_cast<int>(rank) * 128;
        const int h0 = hn * 256 + static_cast<int>(rank) * 128;
#pragma unroll
        for (int q = 0; q < 2; ++q) tma_load_2d_pair(sa + q * kDwBox, &tm_dz, &full_bar[s], v0 + q * 64, kb * 64);
#pragma unroll
        for (int q = 0; q < 2; ++q)
          tma_load_2d_pair(sa + kDwAStage + q * kDwBox, &tm_h, &full_bar[s], h0 + q * 64, kb * 64);
      }
    }
  } else if (warp == 1) {
    const uint32_t idesc = make_idesc_bf16(256, 256, true, true);
    // MN-major SW128 (verified on hardware): lbo = distance between 64-element M/N groups (one TMA box),
    // sbo = distance between 8-row k groups
    const uint32_t lbo = static_cast<uint32_t>(kDwBox);
    const uint32_t sbo = 1024u;
    int it = 0, run = 0;
    int item = begin;
    while (leader && item < end) {
      const int tile = item / p.nkb;
      const int run_end = min(end, (tile + 1) * p.nkb);
      const int buf = run & 1;
      mbar_wait(&tempty_bar[buf], ((run >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + buf * 256;
      bool first = true;
      for (; item < run_end; ++item, ++it) {
        const int s = it % kDwStages;
        const uint32_t ph = (it / kDwStages) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_addr = smem_u32(smem + s * kDwStageBytes);
          const uint32_t b_addr = a_addr + kDwAStage;
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {  // 16 slab rows per MMA
            const uint64_t ad = make_smem_desc_sw128(a_addr + kk * 2048, lbo, sbo);
            const uint64_t bd = make_smem_desc_sw128(b_addr + kk * 2048, lbo, sbo);
            umma_bf16_pair(d_tmem, ad, bd, idesc, (first && kk == 0) ? 0u : 1u);
          }
          umma_commit_pair(&empty_bar[s], 3);
          if (item == run_end - 1) umma_commit_pair(&tfull_bar[buf], 3);
        }
        first = false;
        __syncwarp();
      }
      ++run;
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    int run = 0;
    int item = begin;
    while (item < end) {
      const int tile = item / p.nkb;
      const int run_end = min(end, (tile + 1) * p.nkb);
      const int vm = tile / p.n_ht, hn = tile - vm * p.n_ht;
      const int buf = run & 1;
      mbar_wait(&tfull_bar[buf], (run >> 1) & 1);
      tc_fence_after();
      const int v = vm * 256 + static_cast<int>(rank) * 128 + r;
      float* out = p.dW + static_cast<size_t>(v) * p.H + hn * 256;
#pragma unroll 1
      for (int g = 0; g < 8; ++g) {
        uint32_t raw[32];
        tmem_ld32(lane_taddr + buf * 256 + g * 32, raw);
        tmem_ld_wait();
        if (v < p.V) {
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int hcol = hn * 256 + g * 32 + 4 * q;
            if (hcol + 3 < p.H) {
              red_add_v4_f32(out + g * 32 + 4 * q, __uint_as_float(raw[4 * q]), __uint_as_float(raw[4 * q + 1]),
                             __uint_as_float(raw[4 * q + 2]), __uint_as_float(raw[4 * q + 3]));
            } else {
#pragma unroll
              for (int e = 0; e < 4; ++e)
                if (hcol + e < p.H) red_add_f32(out + g * 32 + 4 * q + e, __uint_as_float(raw[4 * q + e]));
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_even_cta(&tempty_bar[buf]);
      item = run_end;
      ++run;
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
// hgen: one warp per lattice row, 16-byte vectors along H.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
hgen_kernel(Lattice L, const __nv_bfloat16* __restrict__ f, const __nv_bfloat16* __restrict__ g,
            __nv_bfloat16* __restrict__ hslab, int tile0, int H) {
  __shared__ TileInfo s_ti;
  const int rows_per_block = 8;
  const int row0 = blockIdx.x * rows_per_block;  // slab row
  const int tile = row0 / kTileRows;
  if (threadIdx.x == 0) s_ti = decode_tile(L, tile0 + tile);
  __syncthreads();
  const TileInfo ti = s_ti;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = row0 + warp;
  const int r = row - tile * kTileRows;
  const int t = ti.t0 + (r >> 3), u = ti.u0 + (r & 7);
  const bool valid = t < ti.T && u <= ti.U;
  uint4* out = reinterpret_cast<uint4*>(hslab + static_cast<size_t>(row) * H);
  const int nvec = H >> 3;
  if (!valid) {
    for (int i = lane; i < nvec; i += 32) out[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  const uint4* fp = reinterpret_cast<const uint4*>(f + (static_cast<size_t>(ti.b) * L.Tmax + t) * H);
  const uint4* gp = reinterpret_cast<const uint4*>(g + (static_cast<size_t>(ti.b) * L.U1max + u) * H);
  for (int i = lane; i < nvec; i += 32) {
    const uint4 a = __ldg(fp + i), b = __ldg(gp + i);
    const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
    uint32_t o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = bf16lo(aw[e]) + bf16lo(bw[e]);
      const float x1 = bf16hi(aw[e]) + bf16hi(bw[e]);
      o[e] = pack_bf16x2(tanh_approx(x0), tanh_approx(x1));
    }
    out[i] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// Wt has ceil(H / 32) * 32 rows (rows >= H are zero).  perm != 0: inside every group of 32 rows, row 8 a + 2 c + e holds
// column 8 c + 2 a + e of W (a, c = 0..3, e = 0, 1) -- the backward mega-kernel reads its dh accumulator in the 16x256b
// fragment layout, where a thread owns accumulator columns 8 a + 2 c + {0, 1}: with the rows of the B operand permuted like
// this those are eight CONTIGUOUS columns of H (8 c .. 8 c + 7), one 16-byte access for h, df and dg each.
__global__ void transpose_w_kernel(const __nv_bfloat16* __restrict__ W, __nv_bfloat16* __restrict__ Wt, int V,
                                   int H, int Vp, int perm) {
  __shared__ __nv_bfloat16 t[32][33];
  const int v0 = blockIdx.x * 32, h0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int v = v0 + i, h = h0 + threadIdx.x;
    t[i][threadIdx.x] = (v < V && h < H) ? W[static_cast<size_t>(v) * H + h] : __float2bfloat16(0.0f);
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int v = v0 + threadIdx.x;
    const int src = perm ? (((i >> 1) & 3) << 3 | (i >> 3) << 1 | (i & 1)) : i;
    if (v < Vp) Wt[static_cast<size_t>(h0 + i) * Vp + v] = t[threadIdx.x][src];
  }
}

// Greedy-decode joint step: one CTA per utterance; warps stride over vocabulary rows.
__global__ void __launch_bounds__(256)
greedy_argmax_kernel(const __nv_bfloat16* __restrict__ f, const __nv_bfloat16* __restrict__ g,
                     const __nv_bfloat16* __restrict__ W, const float* __restrict__ bias,
                     const int* __restrict__ t_idx, int* __restrict__ out_k, int Tmax, int V, int H) {
  extern __shared__ float sh[];  // H floats of h, then 8 (value, index) pairs
  const int b = blockIdx.x;
  const int t = t_idx[b];
  if (t < 0) { if (threadIdx.x == 0) out_k[b] = -1; return; }
  const __nv_bfloat16* fr = f + (static_cast<size_t>(b) * Tmax + t) * H;
  const __nv_bfloat16* gr = g + static_cast<size_t>(b) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float x = __bfloat162float(fr[i]) + __bfloat162float(gr[i]);
    sh[i] = __bfloat162float(__float2bfloat16_rn(tanh_approx(x)));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  float best = -INFINITY;
  int best_k = 0x7fffffff;
  for (int v = warp; v < V; v += nwarp) {
    const __nv_bfloat16* wr = W + static_cast<size_t>(v) * H;
    float acc = 0.0f;
    for (int i = lane * 8; i < H; i += 256) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(wr + i));
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc = fmaf(bf16lo(ww[e]), sh[i + 2 * e], acc);
        acc = fmaf(bf16hi(ww[e]), sh[i + 2 * e + 1], acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    acc += bias ? bias[v] : 0.0f;
    if (acc > best || (acc == best && v < best_k)) { best = acc; best_k = v; }
  }
  float* rv = sh + H;
  int* ri = reinterpret_cast<int*>(rv + nwarp);
  if (lane == 0) { rv[warp] = best; ri[warp] = best_k; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarp; ++w)
      if (rv[w] > best || (rv[w] == best && ri[w] < best_k)) { best = rv[w]; best_k = ri[w]; }
    out_k[b] = best_k;
  }
}

// One greedy decode step with the bookkeeping fused in: for every utterance still inside its frames, the joint argmax at
// its current frame, then (one thread) emit / count / advance exactly as the reference-style host loop did.
// State arrays are device int32 and updated in place.
__global__ void __launch_bounds__(256)
greedy_step_kernel(const __nv_bfloat16* __restrict__ f, const float* __restrict__ g, const __nv_bfloat16* __restrict__ W,
                   const float* __restrict__ bias, const int* __restrict__ lens, int* __restrict__ t_cur,
                   int* __restrict__ emitted, int* __restrict__ n_sym, int* __restrict__ sym, int sym_cap,
                   int* __restrict__ is_sym, int* __restrict__ label, int* __restrict__ active, int Tmax, int V, int H,
                   int blank, int max_symbols) {
  extern __shared__ float sh[];  // H floats of h, then per-warp (value, index) pairs
  const int b = blockIdx.x;
  const int t = t_cur[b];
  if (t >= lens[b]) {            // finished utterance: nothing changes
    if (threadIdx.x == 0) { is_sym[b] = 0; label[b] = 0; active[b] = 0; }
    return;
  }
  const __nv_bfloat16* fr = f + (static_cast<size_t>(b) * Tmax + t) * H;
  const float* gr = g + static_cast<size_t>(b) * H;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float x = __bfloat162float(fr[i]) + __bfloat162float(__float2bfloat16_rn(gr[i]));
    sh[i] = __bfloat162float(__float2bfloat16_rn(tanh_approx(x)));
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  float best = -INFINITY;
  int best_k = 0x7fffffff;
  for (int v = warp; v < V; v += nwarp) {
    const __nv_bfloat16* wr = W + static_cast<size_t>(v) * H;
    float acc = 0.0f;
    for (int i = lane * 8; i < H; i += 256) {
      const uint4 w = __ldg(reinterpret_cast<const uint4*>(wr + i));
      const uint32_t ww[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc = fmaf(bf16lo(ww[e]), sh[i + 2 * e], acc);
        acc = fmaf(bf16hi(ww[e]), sh[i + 2 * e + 1], acc);
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    acc += bias ? bias[v] : 0.0f;
    if (acc > best || (acc == best && v < best_k)) { best = acc; best_k = v; }
  }
  float* rv = sh + H;
  int* ri = reinterpret_cast<int*>(rv + nwarp);
  if (lane == 0) { rv[warp] = best; ri[warp] = best_k; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < nwarp; ++w)
      if (rv[w] > best || (rv[w] == best && ri[w] < best_k)) { best = rv[w]; best_k = ri[w]; }
    const bool is = best_k != blank;
    int em = emitted[b];
    if (is) {
      const int n = n_sym[b];
      sym[static_cast<size_t>(b) * sym_cap + (n < sym_cap ? n : sym_cap - 1)] = best_k;
      n_sym[b] = n + 1;
      ++em;
    }
    int tn = t;
    if (!is || em >= max_symbols) { ++tn; em = 0; }
    t_cur[b] = tn;
    emitted[b] = em;
    is_sym[b] = is ? 1 : 0;
    label[b] = is ? best_k : 0;
    active[b] = tn < lens[b] ? 1 : 0;
  }
}

template <int EPI>
void launch_slab_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& tout, const GemmArgs& a,
                      int n_tiles, cudaStream_t s) {
  static bool configured[kMaxDevices] = {};
  if (bool& c = configured[current_device()]; !c) {
    cudaFuncSetAttribute(slab_gemm_kernel<EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_total<EPI>());
    c = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((n_tiles + 1) / 2 * 2);  // CTA pairs; an odd tail gets a ghost CTA
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem_total<EPI>();
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, slab_gemm_kernel<EPI>, ta, tb, tout, a);
}

}  // namespace

void set_gemm_dbg(int v) { g_gemm_dbg = v; }
int get_gemm_dbg() { return g_gemm_dbg; }
int read_gemm_prof(unsigned long long* out, int n) {
  if (n > 160 * 8) n = 160 * 8;
  return cudaMemcpyFromSymbol(out, g_prof, sizeof(unsigned long long) * n) == cudaSuccess ? n : -1;
}
int smem_bytes_fwd(int) { return smem_total<kFwd>(); }
int smem_bytes_dz(int) { return smem_total<kDz>(); }
int smem_bytes_dh() { return smem_total<kDh>(); }
int smem_bytes_dw() { return kDwSmem; }

void launch_hgen(const Lattice& L, const __nv_bfloat16* f, const __nv_bfloat16* g, __nv_bfloat16* hslab, int tile0,
                 int n_tiles, int H, cudaStream_t s) {
  hgen_kernel<<<n_tiles * (kTileRows / 8), 256, 0, s>>>(L, f, g, hslab, tile0, H);
}

void launch_transpose_w(const __nv_bfloat16* W, __nv_bfloat16* Wt, int V, int H, int Vp, int perm, cudaStream_t s) {
  dim3 grid((Vp + 31) / 32, (H + 31) / 32);
  transpose_w_kernel<<<grid, dim3(32, 8), 0, s>>>(W, Wt, V, H, Vp, perm);
}

void launch_joint_fwd(const Lattice& L, const JointDims& d, const CUtensorMap& tm_h, const CUtensorMap& tm_w,
                      const FwdArgs& a, int tile0, int n_tiles, int nc, cudaStream_t s) {
  GemmArgs g{};
  g.dbg = g_gemm_dbg;
  g.L = L; g.tile0 = tile0; g.n_tiles_total = L.n_tiles_total; g.n_total = d.V; g.nc = nc; g.n_chunks = (d.V + nc - 1) / nc;
  g.k_blocks = (d.H + kBK - 1) / kBK; g.blank = d.blank; g.Umax = d.Umax; g.Vp = d.Vp; g.H = d.H;
  g.bias = a.bias; g.y = a.y; g.lse_tile = a.lse_tile; g.lpb = a.lpb; g.lpl = a.lpl;
  launch_slab_gemm<kFwd>(tm_h, tm_w, tm_h, g, n_tiles, s);
}

void launch_joint_dz(const Lattice& L, const JointDims& d, const CUtensorMap& tm_h, const CUtensorMap& tm_w,
                     const CUtensorMap& tm_dz_store, const DzArgs& a, int tile0, int n_tiles, int nc,
                     cudaStream_t s) {
  GemmArgs g{};
  g.L = L; g.tile0 = tile0; g.n_tiles_total = L.n_tiles_total; g.n_total = d.V; g.nc = nc; g.n_chunks = (d.V + nc - 1) / nc;
  g.k_blocks = (d.H + kBK - 1) / kBK; g.blank = d.blank; g.Umax = d.Umax; g.Vp = d.Vp; g.H = d.H;
  g.bias = a.bias; g.y = a.y; g.lse_tile = const_cast<float*>(a.lse_tile);
  g.lpb = const_cast<float*>(a.lpb); g.lpl = const_cast<float*>(a.lpl);
  g.c1 = a.c1; g.c2 = a.c2; g.grad_loss = a.grad_loss; g.db = a.db;
  launch_slab_gemm<kDz>(tm_h, tm_w, tm_dz_store, g, n_tiles, s);
}

void launch_joint_dh(const Lattice& L, const JointDims& d, const CUtensorMap& tm_dz, const CUtensorMap& tm_wt,
                     const DhArgs& a, int tile0, int n_tiles, int nc, cudaStream_t s) {
  GemmArgs g{};
  g.L = L; g.tile0 = tile0; g.n_tiles_total = L.n_tiles_total; g.n_total = d.H; g.nc = nc; g.n_chunks = (d.H + nc - 1) / nc;
  g.k_blocks = (d.Vp + kBK - 1) / kBK; g.blank = d.blank; g.Umax = d.Umax; g.Vp = d.Vp; g.H = d.H;
  g.hslab = a.hslab; g.df = a.df; g.dg = a.dg;
  launch_slab_gemm<kDh>(tm_dz, tm_wt, tm_dz, g, n_tiles, s);
}

void launch_joint_dw(const JointDims& d, const CUtensorMap& tm_dz_mn, const CUtensorMap& tm_h_mn, float* dW,
                     int n_tiles, int n_ctas, cudaStream_t s) {
  static bool configured[kMaxDevices] = {};
  if (bool& c = configured[current_device()]; !c) {
    cudaFuncSetAttribute(dw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kDwSmem);
    c = true;
  }
  DwArgs a{};
  a.dW = dW; a.V = d.V; a.H = d.H;
  a.n_vt = (d.V + 255) / 256; a.n_ht = (d.H + 255) / 256;
  a.nkb = n_tiles * 2;  // 128 rows per tile, 64 per k-block
  a.total = a.n_vt * a.n_ht * a.nkb;
  int n_pairs = n_ctas / 2;
  if (n_pairs < 1) n_pairs = 1;
  if (n_pairs > a.total) n_pairs = a.total;
  a.per_pair = (a.total + n_pairs - 1) / n_pairs;
  n_pairs = (a.total + a.per_pair - 1) / a.per_pair;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(2 * n_pairs);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kDwSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, dw_kernel, tm_dz_mn, tm_h_mn, a);
}

void launch_greedy_step(const __nv_bfloat16* f, const float* g, const __nv_bfloat16* W, const float* bias, const int* lens,
                        int* t_cur, int* emitted, int* n_sym, int* sym, int sym_cap, int* is_sym, int* label, int* active,
                        int B, int Tmax, int V, int H, int blank, int max_symbols, cudaStream_t s) {
  const size_t smem = (H + 16) * sizeof(float) + 64;
  greedy_step_kernel<<<B, 256, smem, s>>>(f, g, W, bias, lens, t_cur, emitted, n_sym, sym, sym_cap, is_sym, label, active,
                                         Tmax, V, H, blank, max_symbols);
}

void launch_greedy_argmax(const __nv_bfloat16* f, const __nv_bfloat16* g, const __nv_bfloat16* W, const float* bias,
                          const int* t_idx, int* out_k, int B, int Tmax, int V, int H, cudaStream_t s) {
  const size_t smem = (H + 16) * sizeof(float) + 64;
  greedy_argmax_kernel<<<B, 256, smem, s>>>(f, g, W, bias, t_idx, out_k, Tmax, V, H);
}

}  // namespace rnnt
