// Persistent RNN-T joint kernels for sm_100a: one launch walks every lattice tile of the batch.
//
//   fwd_persist_kernel   h = bf16(tanh(f_t + g_u)) produced in-kernel by four "hgen" warps one tile ahead
//                        (into a per-CTA, double-buffered, L2-resident scratch), logits = h . W^T on tcgen05
//                        (CTA pair, M = 256, TMEM double-buffered 256-column chunks), online log-softmax in
//                        the epilogue warps; keeps only lse, lp_blank, lp_label.
//
// Per-launch costs of the slab kernels in joint.cu (launch gap, barrier init, TMEM allocation, pipeline
// fill, un-overlapped last epilogue: ~11 us per 38 us launch, measured with %globaltimer stamps) are paid
// once per step here, and the tanh pass (MUFU-bound) hides behind the tensor pipe.
//
// Warp roles (320 threads): 0 = TMA producer, 1 = TMEM allocator + MMA issuer (leader CTA only),
// 2..5 = epilogue (TMEM lane quadrant = warp & 3), 6..9 = hgen.
#include <math.h>

#include "launch.h"
#include "ptx.cuh"

namespace rnnt {

namespace {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kNCmax = 256;
constexpr int kAStage = kBM * kBK * 2;
constexpr int kBStage = (kNCmax / 2) * kBK * 2;
constexpr int kStageBytes = kAStage + kBStage;
constexpr int kPThreads = 320;
constexpr int kEpiThreads = 128;
constexpr int kHgenThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kMaxBiasCols = 2048;

// bring-up profiling (gemm_dbg & 4): per-CTA wait-cycle counters of the last persistent launch
//   [0] MMA loop cycles  [1] MMA waiting on full (TMA-starved)  [2] MMA waiting on tempty (epilogue-starved)
//   [3] MMA loop ns (%globaltimer)  [4] TMA waiting on hfull (hgen-starved)  [5] TMA waiting on empty
//   [6] hgen busy cycles  [7] epilogue busy cycles (between tfull and tempty arrive)
__device__ unsigned long long g_pprof[160 * 8];
__device__ __forceinline__ unsigned long long gtimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define PCNT_BEGIN(var) long long var##_t0 = 0; if (p.dbg & 4) var##_t0 = clock64()
#define PCNT_END(var, acc) do { if (p.dbg & 4) acc += clock64() - var##_t0; } while (0)

__device__ __forceinline__ float pick32(const float (&v)[32], int idx) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < 32; ++i) r = (idx == i) ? v[i] : r;
  return r;
}

__device__ __forceinline__ void st_cg_u4(void* p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
               : "memory");
}
// generic-proxy global writes -> visible to later async-proxy (TMA) reads
__device__ __forceinline__ void fence_proxy_async_global() { asm volatile("fence.proxy.async.global;" ::: "memory"); }

// h rows of one tile -> dst[128][H] (row = dt * 8 + du); rows outside the utterance's lattice are zero.
// Called by the 128 hgen threads (ht = 0..127); each owns 8-column vectors cv = ht, ht + 128, ...
__device__ __forceinline__ void hgen_tile(const TileInfo& ti, const __nv_bfloat16* __restrict__ f,
                                          const __nv_bfloat16* __restrict__ g, __nv_bfloat16* dst, int H, int Tmax,
                                          int U1max, int ht) {
  const int nvec = H >> 3;
  int nt = ti.T - ti.t0; nt = nt < 0 ? 0 : (nt > kTT ? kTT : nt);
  int nu = ti.U + 1 - ti.u0; nu = nu < 0 ? 0 : (nu > kTU ? kTU : nu);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int cv = ht; cv < nvec; cv += kHgenThreads) {
    float gv[kTU][8];
#pragma unroll
    for (int du = 0; du < kTU; ++du) {
      uint4 q = zero;
      if (du < nu) q = __ldg(reinterpret_cast<const uint4*>(g + (static_cast<size_t>(ti.b) * U1max + ti.u0 + du) * H) + cv);
      const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
      for (int e = 0; e < 4; ++e) { gv[du][2 * e] = bf16lo(w[e]); gv[du][2 * e + 1] = bf16hi(w[e]); }
    }
    uint4 fq = zero;
    if (nt > 0) fq = __ldg(reinterpret_cast<const uint4*>(f + (static_cast<size_t>(ti.b) * Tmax + ti.t0) * H) + cv);
#pragma unroll 1
    for (int dt = 0; dt < kTT; ++dt) {
      uint4 fnext = zero;
      if (dt + 1 < nt)
        fnext = __ldg(reinterpret_cast<const uint4*>(f + (static_cast<size_t>(ti.b) * Tmax + ti.t0 + dt + 1) * H) + cv);
      __nv_bfloat16* orow = dst + static_cast<size_t>(dt * kTU) * H + cv * 8;
      if (dt < nt) {
        const uint32_t w[4] = {fq.x, fq.y, fq.z, fq.w};
        float fv[8];
#pragma unroll
        for (int e = 0; e < 4; ++e) { fv[2 * e] = bf16lo(w[e]); fv[2 * e + 1] = bf16hi(w[e]); }
#pragma unroll
        for (int du = 0; du < kTU; ++du) {
          uint4 o = zero;
          if (du < nu) {
            o.x = pack_bf16x2(tanh_approx(fv[0] + gv[du][0]), tanh_approx(fv[1] + gv[du][1]));
            o.y = pack_bf16x2(tanh_approx(fv[2] + gv[du][2]), tanh_approx(fv[3] + gv[du][3]));
            o.z = pack_bf16x2(tanh_approx(fv[4] + gv[du][4]), tanh_approx(fv[5] + gv[du][5]));
            o.w = pack_bf16x2(tanh_approx(fv[6] + gv[du][6]), tanh_approx(fv[7] + gv[du][7]));
          }
          st_cg_u4(orow + static_cast<size_t>(du) * H, o);
        }
      } else {
#pragma unroll
        for (int du = 0; du < kTU; ++du) st_cg_u4(orow + static_cast<size_t>(du) * H, zero);
      }
      fq = fnext;
    }
  }
}

constexpr int kFwdStages = 6;
constexpr int kFwdSmem = kFwdStages * kStageBytes + kMaxBiasCols * 4 + 1024 + 256;

__global__ void __launch_bounds__(kPThreads, 1)
fwd_persist_kernel(const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b, const FwdPArgs p) {
  constexpr int kStages = kFwdStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* sbias = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes + kMaxBiasCols * 4);
  uint64_t* full_bar = bars;                      // [kStages] TMA -> MMA (leader's copy collects both CTAs' bytes)
  uint64_t* empty_bar = bars + kStages;           // [kStages] MMA -> TMA (commit, both CTAs)
  uint64_t* tfull_bar = bars + 2 * kStages;       // [2] MMA -> epilogue (commit, both CTAs)
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2] epilogue -> MMA (8 warps arrive on the leader's copy)
  uint64_t* hfull_bar = bars + 2 * kStages + 4;   // [2] hgen -> TMA (128 threads, local)
  uint64_t* hempty_bar = bars + 2 * kStages + 6;  // [2] MMA -> hgen (commit, both CTAs)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 8);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_pairs = gridDim.x >> 1;
  const int n_ptiles = (p.n_tiles_total + 1) >> 1;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_a);
    prefetch_tmap(&tm_b);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8);
      mbar_init(&hfull_bar[i], kHgenThreads); mbar_init(&hempty_bar[i], 1);
    }
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

  const int nc_half = p.nc >> 1;
  const uint32_t b_bytes = static_cast<uint32_t>(nc_half) * kBK * 2;
  const int a_row0 = blockIdx.x * 2 * kBM;  // this CTA's two scratch tiles in the tm_a row space

  if (warp == 0) {
    // ------------------------------- TMA producer -------------------------------
    if (lane == 0) {
      int gi = 0, it = 0;
      long long w_hfull = 0, w_empty = 0;
      for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++it) {
        const int hb = it & 1;
        { PCNT_BEGIN(a); mbar_wait(&hfull_bar[hb], (it >> 1) & 1); PCNT_END(a, w_hfull); }
        for (int j = 0; j < p.n_chunks; ++j) {
          for (int k = 0; k < p.k_blocks; ++k, ++gi) {
            const int s = gi % kStages;
            const uint32_t ph = (gi / kStages) & 1;
            { PCNT_BEGIN(a); mbar_wait(&empty_bar[s], ph ^ 1); PCNT_END(a, w_empty); }
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * (kAStage + b_bytes));
            uint8_t* sa = stage_base + s * kStageBytes;
            tma_load_2d_pair(sa, &tm_a, &full_bar[s], k * kBK, a_row0 + hb * kBM);
            tma_load_2d_pair(sa + kAStage, &tm_b, &full_bar[s], k * kBK, j * p.nc + static_cast<int>(rank) * nc_half);
          }
        }
      }
      if (p.dbg & 4) { g_pprof[blockIdx.x * 8 + 4] = w_hfull; g_pprof[blockIdx.x * 8 + 5] = w_empty; }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -----------------
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(2 * kBM, p.nc, false, false);
      int gi = 0, gc = 0, it = 0;
      long long w_full = 0, w_tempty = 0;
      const long long c_begin = clock64();
      const unsigned long long ns_begin = gtimer_ns();
      for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++it) {
        for (int j = 0; j < p.n_chunks; ++j, ++gc) {
          const int buf = gc & 1;
          { PCNT_BEGIN(a); mbar_wait(&tempty_bar[buf], ((gc >> 1) & 1) ^ 1); PCNT_END(a, w_tempty); }
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kNCmax;
          for (int k = 0; k < p.k_blocks; ++k, ++gi) {
            const int s = gi % kStages;
            const uint32_t ph = (gi / kStages) & 1;
            { PCNT_BEGIN(a); mbar_wait(&full_bar[s], ph); PCNT_END(a, w_full); }
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_addr = smem_u32(stage_base + s * kStageBytes);
              const uint64_t ad = make_smem_desc_sw128(a_addr, 16, 1024);
              const uint64_t bd = make_smem_desc_sw128(a_addr + kAStage, 16, 1024);
#pragma unroll
              for (int kk = 0; kk < kBK / 16; ++kk)
                umma_bf16_pair(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
              umma_commit_pair(&empty_bar[s], 3);
              if (k == p.k_blocks - 1) {
                umma_commit_pair(&tfull_bar[buf], 3);
                if (j == p.n_chunks - 1) umma_commit_pair(&hempty_bar[it & 1], 3);  // scratch tile fully consumed
              }
            }
            __syncwarp();
          }
        }
      }
      if ((p.dbg & 4) && lane == 0) {
        g_pprof[blockIdx.x * 8 + 0] = clock64() - c_begin;
        g_pprof[blockIdx.x * 8 + 1] = w_full;
        g_pprof[blockIdx.x * 8 + 2] = w_tempty;
        g_pprof[blockIdx.x * 8 + 3] = gtimer_ns() - ns_begin;
      }
    }
  } else if (warp < 6) {
    // ------------------------------- epilogue ------------------------------------
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const int et = (warp - 2) * 32 + lane;
    const int dt = r >> 3, du = r & 7;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
    const int ncols = p.n_chunks * p.nc;
    for (int c = et; c < ncols; c += kEpiThreads)
      sbias[c] = (c < p.V) ? (p.bias ? p.bias[c] * kLog2e : 0.0f) : -INFINITY;
    named_bar_sync(1, kEpiThreads);

    int gc = 0;
    long long busy = 0;
    for (int pt = pair; pt < n_ptiles; pt += n_pairs) {
      const int tile = 2 * pt + static_cast<int>(rank);
      const bool ghost = tile >= p.n_tiles_total;
      const TileInfo ti = decode_tile(p.L, ghost ? p.n_tiles_total - 1 : tile);
      const int t = ti.t0 + dt, u = ti.u0 + du;
      const bool valid = !ghost && (t < ti.T) && (u <= ti.U);
      const size_t grow = static_cast<size_t>(tile) * kBM + r;
      const size_t didx = valid ? diag_index(p.L, ti.b, t, u) : 0;
      const int label = (valid && u < ti.U) ? p.y[static_cast<size_t>(ti.b) * p.Umax + u] : -1;
      float mx = -INFINITY, sum = 0.0f, zb = 0.0f, zl = 0.0f;
      for (int j = 0; j < p.n_chunks; ++j, ++gc) {
        const int buf = gc & 1;
        mbar_wait(&tfull_bar[buf], (gc >> 1) & 1);
        tc_fence_after();
        PCNT_BEGIN(e);
        for (int g = 0; g < p.nc / 32; ++g) {
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
        PCNT_END(e, busy);
      }
      const float lse2 = mx + lg2f(sum);
      if (!ghost) p.lse_tile[grow] = valid ? lse2 * kLn2 : 0.0f;
      if (valid) {
        p.lpb[didx] = (zb - lse2) * kLn2;
        p.lpl[didx] = (u < ti.U) ? (zl - lse2) * kLn2 : kNeg;
      }
    }
    if ((p.dbg & 4) && et == 0) g_pprof[blockIdx.x * 8 + 7] = busy;
  } else {
    // ------------------------------- hgen -----------------------------------------
    const int ht = threadIdx.x - 192;
    __nv_bfloat16* my_scratch = p.hscratch + static_cast<size_t>(a_row0) * p.H;
    int it = 0;
    long long busy = 0;
    for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++it) {
      const int hb = it & 1;
      const int tile = 2 * pt + static_cast<int>(rank);
      mbar_wait(&hempty_bar[hb], ((it >> 1) & 1) ^ 1);
      PCNT_BEGIN(h);
      if (tile < p.n_tiles_total) {
        const TileInfo ti = decode_tile(p.L, tile);
        hgen_tile(ti, p.f, p.g, my_scratch + static_cast<size_t>(hb) * kBM * p.H, p.H, p.L.Tmax, p.L.U1max, ht);
      }
      __threadfence();
      fence_proxy_async_global();
      mbar_arrive(&hfull_bar[hb]);
      PCNT_END(h, busy);
    }
    if ((p.dbg & 4) && ht == 0) g_pprof[blockIdx.x * 8 + 6] = busy;
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2cta(tmem_base, kTmemCols);
  }
}


// =================================================================================================
// Backward "mega-kernel": one launch per step.
//
//   producer pairs [0, P)      per pair-tile (256 lattice rows): hgen -> dz pass (logits recompute, softmax,
//                              dz = c0*softmax - [blank]c1 - [label]c2 -> bf16 ring slot, + db) -> dh pass
//                              (dh = dz . W, dpre = dh (1 - h^2), tile-reduced red.add into df / dg).
//                              The pair's h and dz tiles live in an L2-resident ring slot (NS slots per pair).
//   consumer pairs [P, P + C)  own one 256 (V) x 512 (H) fp32 block of dW in TMEM for the WHOLE step and
//                              stream every ring slot of their K-group through  dW += dz^T . h
//                              (both operands MN-major straight out of the row-major slots).
//
// Cross-CTA flow control is two global counters per ring slot: `ready` (+1 by each producer CTA when its half
// of the slot is complete and globally visible) and `done` (+1 by each consumer pair once the slot's data
// has landed in its smem).  Every wait is bounded and traps instead of hanging.
// =================================================================================================
constexpr int kBwdStages = 4;                       // producer: 4 x 32 KB;  consumer: 4 x 48 KB
constexpr int kDhPitch = 68;
constexpr int kUnionBytes = 2 * kBM * kDhPitch * 4;  // dh fp32 tiles (69,632 B) >= dz staging (65,536 B)
constexpr int kCStageBytes = 3 * 16384;
constexpr int kMaxNS = 4;
constexpr int kMaxVChunks = 8;
constexpr int kProdSmem = kBwdStages * kStageBytes + kUnionBytes + kMaxBiasCols * 4;
constexpr int kConsSmem = kBwdStages * kCStageBytes;
constexpr int kMegaSmem = (kProdSmem > kConsSmem ? kProdSmem : kConsSmem) + 1024 + 512;

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* ptr) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ptr) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu_add_u32(unsigned* ptr, unsigned v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(ptr), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_counter_ge(const unsigned* ptr, unsigned target) {
  unsigned spins = 0;
  while (ld_acquire_gpu_u32(ptr) < target) {
    __nanosleep(100);
    if (++spins > (1u << 23)) __trap();
  }
}
__device__ __forceinline__ uint4 ld_cg_u4(const void* ptr) {
  uint4 v;
  asm volatile("ld.global.cg.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(ptr));
  return v;
}
__device__ __forceinline__ void tma_store_wait_all1() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }

__global__ void __launch_bounds__(kPThreads, 1)
bwd_mega_kernel(const __grid_constant__ CUtensorMap tm_h, const __grid_constant__ CUtensorMap tm_w,
                const __grid_constant__ CUtensorMap tm_dz, const __grid_constant__ CUtensorMap tm_wt,
                const __grid_constant__ CUtensorMap tm_dz_mn, const __grid_constant__ CUtensorMap tm_h_mn,
                const BwdPArgs p) {
  constexpr int kStages = kBwdStages;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  constexpr int kDataBytes = (kProdSmem > kConsSmem ? kProdSmem : kConsSmem);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDataBytes);
  uint64_t* full_bar = bars;                      // [4]
  uint64_t* empty_bar = bars + 4;                 // [4]
  uint64_t* tfull_bar = bars + 8;                 // [2]
  uint64_t* tempty_bar = bars + 10;               // [2]
  uint64_t* hfull_bar = bars + 12;                // [kMaxNS] hgen -> TMA / epilogue (128 threads, local)
  uint64_t* hfree_bar = bars + 16;                // [kMaxNS] epilogue (dh pass finished with the slot's h) -> hgen
  uint64_t* dzr_bar = bars + 20;                  // [kMaxVChunks] dz chunk stored and visible -> TMA (dh pass)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 28);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1;
  const int n_ptiles = (p.n_tiles_total + 1) >> 1;
  const bool is_producer = pair < p.P;

  if (threadIdx.x == 0) {
    prefetch_tmap(&tm_h); prefetch_tmap(&tm_w); prefetch_tmap(&tm_dz);
    prefetch_tmap(&tm_wt); prefetch_tmap(&tm_dz_mn); prefetch_tmap(&tm_h_mn);
    for (int i = 0; i < kStages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 8); }
    for (int i = 0; i < kMaxNS; ++i) { mbar_init(&hfull_bar[i], kHgenThreads); mbar_init(&hfree_bar[i], 4); }
    for (int i = 0; i < kMaxVChunks; ++i) mbar_init(&dzr_bar[i], 1);
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

  if (is_producer) {
    // ===========================================================================================
    //                                       PRODUCER
    // ===========================================================================================
    uint8_t* stage_base = smem;
    uint8_t* uni = smem + kStages * kStageBytes;                 // dz staging / dh fp32 tiles
    float* sbias = reinterpret_cast<float*>(uni + kUnionBytes);
    const int ncv_half = p.nc_v >> 1, nch_half = p.nc_h >> 1;
    const uint32_t bv_bytes = static_cast<uint32_t>(ncv_half) * kBK * 2;
    const uint32_t bh_bytes = static_cast<uint32_t>(nch_half) * kBK * 2;

    if (warp == 0) {
      // ------------------------------- TMA producer -------------------------------
      if (lane == 0) {
        int gi = 0, it = 0;
        long long w_hfull = 0, w_empty = 0, w_dzr = 0;
        for (int pt = pair; pt < n_ptiles; pt += p.P, ++it) {
          const int slot = it % p.NS, use = it / p.NS;
          const int ring_row = ((pair * p.NS + slot) * 2 + static_cast<int>(rank)) * kBM;
          { PCNT_BEGIN(a); mbar_wait(&hfull_bar[slot], use & 1); PCNT_END(a, w_hfull); }
          for (int j = 0; j < p.n_chunks_v; ++j) {
            for (int k = 0; k < p.kb_h; ++k, ++gi) {
              const int s = gi % kStages;
              const uint32_t ph = (gi / kStages) & 1;
              { PCNT_BEGIN(a); mbar_wait(&empty_bar[s], ph ^ 1); PCNT_END(a, w_empty); }
              if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * (kAStage + bv_bytes));
              uint8_t* sa = stage_base + s * kStageBytes;
              tma_load_2d_pair(sa, &tm_h, &full_bar[s], k * kBK, ring_row);
              tma_load_2d_pair(sa + kAStage, &tm_w, &full_bar[s], k * kBK, j * p.nc_v + static_cast<int>(rank) * ncv_half);
            }
          }
          int dz_ready = -1;
          for (int j = 0; j < p.n_chunks_h; ++j) {
            for (int k = 0; k < p.kb_v; ++k, ++gi) {
              if (j == 0) {
                int cj = (k * kBK + kBK - 1) / p.nc_v;
                if (cj > p.n_chunks_v - 1) cj = p.n_chunks_v - 1;
                while (dz_ready < cj) {
                  ++dz_ready;
                  PCNT_BEGIN(a); mbar_wait(&dzr_bar[dz_ready], it & 1); PCNT_END(a, w_dzr);
                }
              }
              const int s = gi % kStages;
              const uint32_t ph = (gi / kStages) & 1;
              { PCNT_BEGIN(a); mbar_wait(&empty_bar[s], ph ^ 1); PCNT_END(a, w_empty); }
              if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * (kAStage + bh_bytes));
              uint8_t* sa = stage_base + s * kStageBytes;
              tma_load_2d_pair(sa, &tm_dz, &full_bar[s], k * kBK, ring_row);
              tma_load_2d_pair(sa + kAStage, &tm_wt, &full_bar[s], k * kBK, j * p.nc_h + static_cast<int>(rank) * nch_half);
            }
          }
        }
        if (p.dbg & 4) {
          g_pprof[blockIdx.x * 8 + 4] = w_hfull; g_pprof[blockIdx.x * 8 + 5] = w_empty; g_pprof[blockIdx.x * 8 + 6] = w_dzr;
        }
      }
    } else if (warp == 1) {
      // ------------------------------- MMA issuer (leader CTA only) -----------------
      if (leader) {
        const uint32_t idesc_v = make_idesc_bf16(2 * kBM, p.nc_v, false, false);
        const uint32_t idesc_h = make_idesc_bf16(2 * kBM, p.nc_h, false, false);
        int gi = 0, gc = 0;
        long long w_full = 0, w_tempty = 0;
        const long long c_begin = clock64();
        const unsigned long long ns_begin = gtimer_ns();
        for (int pt = pair; pt < n_ptiles; pt += p.P) {
          for (int pass = 0; pass < 2; ++pass) {
            const int n_chunks = pass == 0 ? p.n_chunks_v : p.n_chunks_h;
            const int k_blocks = pass == 0 ? p.kb_h : p.kb_v;
            const uint32_t idesc = pass == 0 ? idesc_v : idesc_h;
            for (int j = 0; j < n_chunks; ++j, ++gc) {
              const int buf = gc & 1;
              { PCNT_BEGIN(a); mbar_wait(&tempty_bar[buf], ((gc >> 1) & 1) ^ 1); PCNT_END(a, w_tempty); }
              tc_fence_after();
              const uint32_t d_tmem = tmem_base + buf * kNCmax;
              for (int k = 0; k < k_blocks; ++k, ++gi) {
                const int s = gi % kStages;
                const uint32_t ph = (gi / kStages) & 1;
                { PCNT_BEGIN(a); mbar_wait(&full_bar[s], ph); PCNT_END(a, w_full); }
                tc_fence_after();
                if (lane == 0) {
                  const uint32_t a_addr = smem_u32(stage_base + s * kStageBytes);
                  const uint64_t ad = make_smem_desc_sw128(a_addr, 16, 1024);
                  const uint64_t bd = make_smem_desc_sw128(a_addr + kAStage, 16, 1024);
#pragma unroll
                  for (int kk = 0; kk < kBK / 16; ++kk)
                    umma_bf16_pair(d_tmem, ad + 2 * kk, bd + 2 * kk, idesc, (k | kk) != 0 ? 1u : 0u);
                  umma_commit_pair(&empty_bar[s], 3);
                  if (k == k_blocks - 1) umma_commit_pair(&tfull_bar[buf], 3);
                }
                __syncwarp();
              }
            }
          }
        }
        if ((p.dbg & 4) && lane == 0) {
          g_pprof[blockIdx.x * 8 + 0] = clock64() - c_begin;
          g_pprof[blockIdx.x * 8 + 1] = w_full;
          g_pprof[blockIdx.x * 8 + 2] = w_tempty;
          g_pprof[blockIdx.x * 8 + 3] = gtimer_ns() - ns_begin;
        }
      }
    } else if (warp < 6) {
      // ------------------------------- epilogue ------------------------------------
      const int quad = warp & 3;
      const int r = quad * 32 + lane;
      const int et = (warp - 2) * 32 + lane;
      const int dt = r >> 3, du = r & 7;
      const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
      const int ncols_v = p.n_chunks_v * p.nc_v;
      for (int c = et; c < ncols_v; c += kEpiThreads)
        sbias[c] = (c < p.V) ? (p.bias ? p.bias[c] * kLog2e : 0.0f) : -INFINITY;
      named_bar_sync(1, kEpiThreads);

      int gc = 0, it = 0;
      for (int pt = pair; pt < n_ptiles; pt += p.P, ++it) {
        const int slot = it % p.NS, use = it / p.NS;
        const int ring_row = ((pair * p.NS + slot) * 2 + static_cast<int>(rank)) * kBM;
        const int tile = 2 * pt + static_cast<int>(rank);
        const bool ghost = tile >= p.n_tiles_total;
        const TileInfo ti = decode_tile(p.L, ghost ? p.n_tiles_total - 1 : tile);
        const int t = ti.t0 + dt, u = ti.u0 + du;
        const bool valid = !ghost && (t < ti.T) && (u <= ti.U);
        const size_t grow = static_cast<size_t>(tile) * kBM + r;
        const size_t didx = valid ? diag_index(p.L, ti.b, t, u) : 0;
        const int label = (valid && u < ti.U) ? p.y[static_cast<size_t>(ti.b) * p.Umax + u] : -1;
        mbar_wait(&hfull_bar[slot], use & 1);  // acquire the hgen warps' writes of this slot's h

        // ---- dz pass ----
        {
          uint8_t* stage_out = uni;
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
          const int n_box = (p.nc_v + 63) / 64;
          for (int j = 0; j < p.n_chunks_v; ++j, ++gc) {
            const int buf = gc & 1;
            mbar_wait(&tfull_bar[buf], (gc >> 1) & 1);
            tc_fence_after();
            if (p.nc_v & 32) {  // odd number of 32-column groups: keep the last box's upper half defined (zero)
              const int g = p.nc_v / 32;
              uint8_t* box = stage_out + (g >> 1) * (kBM * 128) + r * 128;
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(box + (((4 + q) ^ (r & 7)) << 4)) = make_uint4(0, 0, 0, 0);
            }
            for (int g = 0; g < p.nc_v / 32; ++g) {
              uint32_t raw[32];
              tmem_ld32(lane_taddr + buf * kNCmax + g * 32, raw);
              tmem_ld_wait();
              const int c0 = j * p.nc_v + g * 32;
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
            {
              const int cb = p.blank - j * p.nc_v;
              if (static_cast<unsigned>(cb) < static_cast<unsigned>(p.nc_v)) {
                const float dv = c0g * ex2f(lpb_r * kLog2e) - c1g;
                uint8_t* a = stage_out + (cb >> 6) * (kBM * 128) + r * 128 + ((((cb & 63) >> 3) ^ (r & 7)) << 4) + (cb & 7) * 2;
                *reinterpret_cast<__nv_bfloat16*>(a) = __float2bfloat16_rn(dv);
              }
              const int cl = label - j * p.nc_v;
              if (label >= 0 && static_cast<unsigned>(cl) < static_cast<unsigned>(p.nc_v)) {
                const float dv = c0g * ex2f(lpl_r * kLog2e) - c2g;
                uint8_t* a = stage_out + (cl >> 6) * (kBM * 128) + r * 128 + ((((cl & 63) >> 3) ^ (r & 7)) << 4) + (cl & 7) * 2;
                *reinterpret_cast<__nv_bfloat16*>(a) = __float2bfloat16_rn(dv);
              }
            }
            fence_proxy_async_smem();
            named_bar_sync(1, kEpiThreads);
            if (et == 0) {
              for (int bx = 0; bx < n_box; ++bx) {
                const int col = j * p.nc_v + bx * 64;
                if (col < p.Vp) tma_store_2d(&tm_dz, stage_out + bx * (kBM * 128), col, ring_row);
              }
              tma_store_commit();
            }
            {
              const int cc = 2 * et;
              if (cc < p.nc_v && !ghost) {
                float s0 = 0.0f, s1 = 0.0f;
                const uint8_t* colp = stage_out + (cc >> 6) * (kBM * 128) + (cc & 7) * 2;
                const int ch = (cc & 63) >> 3;
#pragma unroll 8
                for (int rr = 0; rr < kBM; ++rr) {
                  const uint32_t w = *reinterpret_cast<const uint32_t*>(colp + rr * 128 + ((ch ^ (rr & 7)) << 4));
                  s0 += bf16lo(w);
                  s1 += bf16hi(w);
                }
                const int gcol = j * p.nc_v + cc;
                if (gcol < p.V) red_add_f32(p.db + gcol, s0);
                if (gcol + 1 < p.V) red_add_f32(p.db + gcol + 1, s1);
              }
            }
            if (et == 0) {
              if (j > 0) { tma_store_wait_all1(); mbar_arrive(&dzr_bar[j - 1]); }  // chunk j-1 is in global memory
              tma_store_wait_read0();
            }
            named_bar_sync(1, kEpiThreads);
          }
          if (et == 0) {
            tma_store_wait_all0();
            mbar_arrive(&dzr_bar[p.n_chunks_v - 1]);
            __threadfence();
            red_release_gpu_add_u32(p.ready + pair * p.NS + slot, 1u);  // this CTA's half of the slot is complete
          }
        }

        // ---- dh pass ----
        {
          float* tile_base = reinterpret_cast<float*>(uni);
          const __nv_bfloat16* hrow = p.h_ring + (static_cast<size_t>(ring_row) + r) * p.H;
          const int n_sub = (p.nc_h + 63) / 64;
          const int total_sub = p.n_chunks_h * n_sub;
          auto load_h = [&](int s_idx, uint4 (&hv)[8]) {
            const int jj = s_idx / n_sub, ss = s_idx - jj * n_sub;
#pragma unroll
            for (int gg = 0; gg < 2; ++gg) {
              const int g = ss * 2 + gg;
              const int c0 = jj * p.nc_h + g * 32;
              const uint4* hp = reinterpret_cast<const uint4*>(hrow + c0);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                hv[gg * 4 + q] = (g * 32 < p.nc_h && c0 + 8 * q < p.H) ? ld_cg_u4(hp + q) : make_uint4(0, 0, 0, 0);
            }
          };
          uint4 hcur[8];
          load_h(0, hcur);
          int s_idx = 0;
          for (int j = 0; j < p.n_chunks_h; ++j, ++gc) {
            const int buf = gc & 1;
            mbar_wait(&tfull_bar[buf], (gc >> 1) & 1);
            tc_fence_after();
            for (int sub = 0; sub < n_sub; ++sub, ++s_idx) {
              uint4 hnext[8];
              if (s_idx + 1 < total_sub) {
                load_h(s_idx + 1, hnext);
              } else {
#pragma unroll
                for (int q = 0; q < 8; ++q) hnext[q] = make_uint4(0, 0, 0, 0);
              }
              float* tile_s = tile_base + (s_idx & 1) * (kBM * kDhPitch);
#pragma unroll
              for (int gg = 0; gg < 2; ++gg) {
                const int g = sub * 2 + gg;
                const int c0 = j * p.nc_h + g * 32;
                float4* trow = reinterpret_cast<float4*>(tile_s + r * kDhPitch + gg * 32);
                if (g * 32 < p.nc_h && c0 < p.H) {
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
              named_bar_sync(1, kEpiThreads);
              const int c4 = et & 15;
              const int colbase = j * p.nc_h + sub * 64 + 4 * c4;
              if (!ghost && sub * 64 + 4 * c4 < p.nc_h && colbase < p.H) {
                const float* tcol = tile_s + 4 * c4;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                  const int a = (et >> 4) + 8 * k;
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int c = 0; c < kTU; ++c) {
                    const float4 v = *reinterpret_cast<const float4*>(tcol + (a * kTU + c) * kDhPitch);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                  }
                  if (ti.t0 + a < ti.T)
                    red_add_v4_f32(p.df + (static_cast<size_t>(ti.b) * p.L.Tmax + ti.t0 + a) * p.H + colbase, acc.x,
                                   acc.y, acc.z, acc.w);
                }
                {
                  const int c = et >> 4;
                  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                  for (int a = 0; a < kTT; ++a) {
                    const float4 v = *reinterpret_cast<const float4*>(tcol + (a * kTU + c) * kDhPitch);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                  }
                  if (ti.u0 + c <= ti.U)
                    red_add_v4_f32(p.dg + (static_cast<size_t>(ti.b) * p.L.U1max + ti.u0 + c) * p.H + colbase, acc.x,
                                   acc.y, acc.z, acc.w);
                }
              }
#pragma unroll
              for (int q = 0; q < 8; ++q) hcur[q] = hnext[q];
            }
          }
          named_bar_sync(1, kEpiThreads);  // the fp32 tiles are free again before the next tile's dz staging
          if (lane == 0) mbar_arrive(&hfree_bar[slot]);
        }
      }
    } else {
      // ------------------------------- hgen -----------------------------------------
      const int ht = threadIdx.x - 192;
      int it = 0;
      for (int pt = pair; pt < n_ptiles; pt += p.P, ++it) {
        const int slot = it % p.NS, use = it / p.NS;
        const int ring_row = ((pair * p.NS + slot) * 2 + static_cast<int>(rank)) * kBM;
        const int tile = 2 * pt + static_cast<int>(rank);
        if (use > 0) {
          mbar_wait(&hfree_bar[slot], (use - 1) & 1);                     // local dh pass done with the slot
          if (ht == 0) wait_counter_ge(p.done + pair * p.NS + slot, static_cast<unsigned>(p.n_out * use));
          named_bar_sync(2, kHgenThreads);                                 // consumers done with the slot
        }
        TileInfo ti;
        if (tile < p.n_tiles_total) {
          ti = decode_tile(p.L, tile);
        } else {
          ti.b = 0; ti.t0 = 0; ti.u0 = 0; ti.T = 0; ti.U = -1;             // ghost half: all-zero rows
        }
        hgen_tile(ti, p.f, p.g, p.h_ring + static_cast<size_t>(ring_row) * p.H, p.H, p.L.Tmax, p.L.U1max, ht);
        __threadfence();
        fence_proxy_async_global();
        mbar_arrive(&hfull_bar[slot]);
      }
    }
  } else if (pair < p.P + p.C) {
    // ===========================================================================================
    //                                       CONSUMER
    // ===========================================================================================
    const int c = pair - p.P;
    const int kg = c / p.n_out;
    const int otile = c - kg * p.n_out;
    const int vt = otile / p.n_ht, hn = otile - vt * p.n_ht;
    const bool two = (p.H - hn * 512) > 256;
    const int v0 = vt * 256 + static_cast<int>(rank) * 128;    // this CTA's 128 V rows of the pair's 256
    const int hx0 = hn * 512 + static_cast<int>(rank) * 128;   // this CTA's half of accumulator X's 256 H columns
    const int hy0 = hx0 + 256;
    const uint32_t stage_tx = 2u * (16384u + (two ? 32768u : 16384u));
    int n_mine = 0;
    for (int pt = kg; pt < n_ptiles; pt += p.KG) ++n_mine;

    if (warp == 0) {
      if (lane == 0) {
        int gi = 0;
        long long w_ready = 0, w_empty = 0;
        for (int pt = kg; pt < n_ptiles; pt += p.KG) {
          const int pp = pt % p.P, it = pt / p.P;
          const int slot = it % p.NS, use = it / p.NS;
          { PCNT_BEGIN(a); wait_counter_ge(p.ready + pp * p.NS + slot, 2u * static_cast<unsigned>(use + 1)); PCNT_END(a, w_ready); }
          fence_proxy_async_global();
          const int row0 = (pp * p.NS + slot) * 2 * kBM;
          for (int kb = 0; kb < 4; ++kb, ++gi) {
            const int s = gi % kStages;
            const uint32_t ph = (gi / kStages) & 1;
            { PCNT_BEGIN(a); mbar_wait(&empty_bar[s], ph ^ 1); PCNT_END(a, w_empty); }
            if (leader) mbar_arrive_expect_tx(&full_bar[s], stage_tx);
            uint8_t* sa = smem + s * kCStageBytes;
            const int rr = row0 + kb * 64;
#pragma unroll
            for (int q = 0; q < 2; ++q) tma_load_2d_pair(sa + q * 8192, &tm_dz_mn, &full_bar[s], v0 + q * 64, rr);
#pragma unroll
            for (int q = 0; q < 2; ++q) tma_load_2d_pair(sa + 16384 + q * 8192, &tm_h_mn, &full_bar[s], hx0 + q * 64, rr);
            if (two) {
#pragma unroll
              for (int q = 0; q < 2; ++q) tma_load_2d_pair(sa + 32768 + q * 8192, &tm_h_mn, &full_bar[s], hy0 + q * 64, rr);
            }
          }
        }
        if (p.dbg & 4) { g_pprof[blockIdx.x * 8 + 4] = w_ready; g_pprof[blockIdx.x * 8 + 5] = w_empty; }
      }
    } else if (warp == 1) {
      if (leader) {
        const uint32_t idesc = make_idesc_bf16(256, 256, true, true);
        int gi = 0;
        long long w_full = 0;
        const long long c_begin = clock64();
        const unsigned long long ns_begin = gtimer_ns();
        for (int pt = kg; pt < n_ptiles; pt += p.KG) {
          const int pp = pt % p.P, it = pt / p.P;
          const int slot = it % p.NS;
          for (int kb = 0; kb < 4; ++kb, ++gi) {
            const int s = gi % kStages;
            const uint32_t ph = (gi / kStages) & 1;
            { PCNT_BEGIN(a); mbar_wait(&full_bar[s], ph); PCNT_END(a, w_full); }
            tc_fence_after();
            if (lane == 0) {
              const uint32_t a_addr = smem_u32(smem + s * kCStageBytes);
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t ad = make_smem_desc_sw128(a_addr + kk * 2048, 8192, 1024);
                const uint64_t bx = make_smem_desc_sw128(a_addr + 16384 + kk * 2048, 8192, 1024);
                umma_bf16_pair(tmem_base, ad, bx, idesc, (gi | kk) != 0 ? 1u : 0u);
                if (two) {
                  const uint64_t by = make_smem_desc_sw128(a_addr + 32768 + kk * 2048, 8192, 1024);
                  umma_bf16_pair(tmem_base + 256, ad, by, idesc, (gi | kk) != 0 ? 1u : 0u);
                }
              }
              umma_commit_pair(&empty_bar[s], 3);
              // the slot's rows have landed in smem (both CTAs' bytes are counted on this barrier)
              if (kb == 3) red_release_gpu_add_u32(p.done + pp * p.NS + slot, 1u);
            }
            __syncwarp();
          }
        }
        if (lane == 0 && n_mine > 0) umma_commit_pair(&tfull_bar[0], 3);
        if ((p.dbg & 4) && lane == 0) {
          g_pprof[blockIdx.x * 8 + 0] = clock64() - c_begin;
          g_pprof[blockIdx.x * 8 + 1] = w_full;
          g_pprof[blockIdx.x * 8 + 3] = gtimer_ns() - ns_begin;
        }
      }
    } else if (warp < 6) {
      if (n_mine > 0) {
        const int quad = warp & 3;
        const int r = quad * 32 + lane;
        const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quad * 32) << 16);
        mbar_wait(&tfull_bar[0], 0);
        tc_fence_after();
        const int v = v0 + r;
        const int hbase = hn * 512;
        float* out = p.dW + static_cast<size_t>(v) * p.H + hbase;
        const int n_groups = two ? 16 : 8;
#pragma unroll 1
        for (int g = 0; g < n_groups; ++g) {
          uint32_t raw[32];
          tmem_ld32(lane_taddr + g * 32, raw);
          tmem_ld_wait();
          if (v < p.V) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const int hcol = hbase + g * 32 + 4 * q;
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
      }
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

}  // namespace

int smem_bytes_fwd_persist() { return kFwdSmem; }
int read_persist_prof(unsigned long long* out, int n) {
  if (n > 160 * 8) n = 160 * 8;
  return cudaMemcpyFromSymbol(out, g_pprof, sizeof(unsigned long long) * n) == cudaSuccess ? n : -1;
}

void launch_fwd_persist(const CUtensorMap& tm_hscratch, const CUtensorMap& tm_w, const FwdPArgs& a, int n_ctas,
                        cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(fwd_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kFwdSmem);
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = kFwdSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, fwd_persist_kernel, tm_hscratch, tm_w, a);
}


int smem_bytes_bwd_mega() { return kMegaSmem; }

void launch_bwd_mega(const CUtensorMap& tm_h, const CUtensorMap& tm_w, const CUtensorMap& tm_dz, const CUtensorMap& tm_wt,
                     const CUtensorMap& tm_dz_mn, const CUtensorMap& tm_h_mn, const BwdPArgs& a, int n_ctas,
                     cudaStream_t s) {
  static bool configured = false;
  if (!configured) {
    cudaFuncSetAttribute(bwd_mega_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kMegaSmem);
    configured = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(kPThreads);
  cfg.dynamicSmemBytes = kMegaSmem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, bwd_mega_kernel, tm_h, tm_w, tm_dz, tm_wt, tm_dz_mn, tm_h_mn, a);
}

}  // namespace rnnt
