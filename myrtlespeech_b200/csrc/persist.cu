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
      for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++it) {
        const int hb = it & 1;
        mbar_wait(&hfull_bar[hb], (it >> 1) & 1);
        for (int j = 0; j < p.n_chunks; ++j) {
          for (int k = 0; k < p.k_blocks; ++k, ++gi) {
            const int s = gi % kStages;
            const uint32_t ph = (gi / kStages) & 1;
            mbar_wait(&empty_bar[s], ph ^ 1);
            if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * (kAStage + b_bytes));
            uint8_t* sa = stage_base + s * kStageBytes;
            tma_load_2d_pair(sa, &tm_a, &full_bar[s], k * kBK, a_row0 + hb * kBM);
            tma_load_2d_pair(sa + kAStage, &tm_b, &full_bar[s], k * kBK, j * p.nc + static_cast<int>(rank) * nc_half);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------- MMA issuer (leader CTA only) -----------------
    if (leader) {
      const uint32_t idesc = make_idesc_bf16(2 * kBM, p.nc, false, false);
      int gi = 0, gc = 0, it = 0;
      for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++it) {
        for (int j = 0; j < p.n_chunks; ++j, ++gc) {
          const int buf = gc & 1;
          mbar_wait(&tempty_bar[buf], ((gc >> 1) & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + buf * kNCmax;
          for (int k = 0; k < p.k_blocks; ++k, ++gi) {
            const int s = gi % kStages;
            const uint32_t ph = (gi / kStages) & 1;
            mbar_wait(&full_bar[s], ph);
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
      }
      const float lse2 = mx + lg2f(sum);
      if (!ghost) p.lse_tile[grow] = valid ? lse2 * kLn2 : 0.0f;
      if (valid) {
        p.lpb[didx] = (zb - lse2) * kLn2;
        p.lpl[didx] = (u < ti.U) ? (zl - lse2) * kLn2 : kNeg;
      }
    }
  } else {
    // ------------------------------- hgen -----------------------------------------
    const int ht = threadIdx.x - 192;
    __nv_bfloat16* my_scratch = p.hscratch + static_cast<size_t>(a_row0) * p.H;
    int it = 0;
    for (int pt = pair; pt < n_ptiles; pt += n_pairs, ++it) {
      const int hb = it & 1;
      const int tile = 2 * pt + static_cast<int>(rank);
      mbar_wait(&hempty_bar[hb], ((it >> 1) & 1) ^ 1);
      if (tile < p.n_tiles_total) {
        const TileInfo ti = decode_tile(p.L, tile);
        hgen_tile(ti, p.f, p.g, my_scratch + static_cast<size_t>(hb) * kBM * p.H, p.H, p.L.Tmax, p.L.U1max, ht);
      }
      __threadfence();
      fence_proxy_async_global();
      mbar_arrive(&hfull_bar[hb]);
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

}  // namespace rnnt
