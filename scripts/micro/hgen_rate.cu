// hgen in isolation: how many SM cycles does it take four warps to produce one 128-row x H tile of
// h = bf16(tanh(f_t + g_u)) into an L2-resident scratch, for the loop shapes tried in persist.cu::hgen_tile?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hgen_rate hgen_rate.cu && ./hgen_rate
#include <cstdio>
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

constexpr int kTT = 16, kTU = 8, kHgenThreads = 128;
__device__ __forceinline__ float tanh_approx(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ float bf16lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }
__device__ __forceinline__ void st_cg_u4(void* p, uint4 v) {
  asm volatile("st.global.cg.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void st_plain_u4(void* p, uint4 v) { *reinterpret_cast<uint4*>(p) = v; }

struct TileInfo { int b, t0, u0, T, U; };

// MODE 0: the shipped loop.  MODE 1: no per-position branches (select instead).  MODE 2: MODE 1 + plain stores.
// MODE 3: two label positions' tanh chains interleaved by hand (MUFU issue alternates with FADD / F2FP of the other chain).
template <int MODE>
__device__ __forceinline__ void hgen_tile(const TileInfo& ti, const __nv_bfloat16* __restrict__ f, const __nv_bfloat16* __restrict__ g,
                                          __nv_bfloat16* dst, int H, int Tmax, int U1max, int ht) {
  const int nvec = H >> 3;
  int nt = ti.T - ti.t0; nt = nt < 0 ? 0 : (nt > kTT ? kTT : nt);
  int nu = ti.U + 1 - ti.u0; nu = nu < 0 ? 0 : (nu > kTU ? kTU : nu);
  const uint4 zero = make_uint4(0, 0, 0, 0);
  int cv0 = ht, cv_step = kHgenThreads, dt0 = 0, dt1 = kTT;
  if (nvec < kHgenThreads) {
    const int n_groups = kHgenThreads / nvec;
    const int fpg = (kTT + n_groups - 1) / n_groups;
    const int grp = ht / nvec;
    cv0 = ht - grp * nvec; cv_step = nvec; dt0 = grp * fpg; dt1 = dt0 + fpg < kTT ? dt0 + fpg : kTT;
    if (grp >= n_groups || dt0 >= kTT) return;
  }
  for (int cv = cv0; cv < nvec; cv += cv_step) {
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
    if (dt0 < nt) fq = __ldg(reinterpret_cast<const uint4*>(f + (static_cast<size_t>(ti.b) * Tmax + ti.t0 + dt0) * H) + cv);
#pragma unroll 1
    for (int dt = dt0; dt < dt1; ++dt) {
      uint4 fnext = zero;
      if (dt + 1 < nt && dt + 1 < dt1)
        fnext = __ldg(reinterpret_cast<const uint4*>(f + (static_cast<size_t>(ti.b) * Tmax + ti.t0 + dt + 1) * H) + cv);
      __nv_bfloat16* orow = dst + static_cast<size_t>(dt * kTU) * H + cv * 8;
      const uint32_t w[4] = {fq.x, fq.y, fq.z, fq.w};
      float fv[8];
#pragma unroll
      for (int e = 0; e < 4; ++e) { fv[2 * e] = bf16lo(w[e]); fv[2 * e + 1] = bf16hi(w[e]); }
      if (MODE == 0) {
        if (dt < nt) {
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
      } else if (MODE >= 3) {
        // batches of NB label positions: all adds, then all tanh, then pack + store (MUFU latency paid once per batch)
        constexpr int NB = MODE == 3 ? 2 : (MODE == 4 ? 4 : 8);
        const bool row_ok = dt < nt;
#pragma unroll
        for (int d0 = 0; d0 < kTU; d0 += NB) {
          float t[NB][8];
#pragma unroll
          for (int k = 0; k < NB; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) t[k][e] = fv[e] + gv[d0 + k][e];
#pragma unroll
          for (int k = 0; k < NB; ++k)
#pragma unroll
            for (int e = 0; e < 8; ++e) t[k][e] = tanh_approx(t[k][e]);
#pragma unroll
          for (int k = 0; k < NB; ++k) {
            uint4 o;
            o.x = pack_bf16x2(t[k][0], t[k][1]); o.y = pack_bf16x2(t[k][2], t[k][3]);
            o.z = pack_bf16x2(t[k][4], t[k][5]); o.w = pack_bf16x2(t[k][6], t[k][7]);
            if (!(row_ok && d0 + k < nu)) o = zero;
            st_cg_u4(orow + static_cast<size_t>(d0 + k) * H, o);
          }
        }
      } else {
        const bool row_ok = dt < nt;
#pragma unroll
        for (int du = 0; du < kTU; ++du) {
          const bool ok = row_ok && du < nu;
          uint4 o;
          o.x = pack_bf16x2(tanh_approx(fv[0] + gv[du][0]), tanh_approx(fv[1] + gv[du][1]));
          o.y = pack_bf16x2(tanh_approx(fv[2] + gv[du][2]), tanh_approx(fv[3] + gv[du][3]));
          o.z = pack_bf16x2(tanh_approx(fv[4] + gv[du][4]), tanh_approx(fv[5] + gv[du][5]));
          o.w = pack_bf16x2(tanh_approx(fv[6] + gv[du][6]), tanh_approx(fv[7] + gv[du][7]));
          if (!ok) o = zero;
          if (MODE == 1) st_cg_u4(orow + static_cast<size_t>(du) * H, o); else st_plain_u4(orow + static_cast<size_t>(du) * H, o);
        }
      }
      fq = fnext;
    }
  }
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) run(const __nv_bfloat16* f, const __nv_bfloat16* g, __nv_bfloat16* scratch, int H, int T, int U,
                                            int B, int tiles_per_cta, int hgen_warps, unsigned long long* cyc) {
  const int warp = threadIdx.x >> 5;
  if (warp >= hgen_warps) return;
  const int ht = threadIdx.x;                       // hgen thread index (0 .. 32 * hgen_warps)
  const int n_tb = (T + kTT - 1) / kTT, n_ub = (U + 1 + kTU - 1) / kTU;
  const long long t0 = clock64();
  for (int i = 0; i < tiles_per_cta; ++i) {
    const int tile = blockIdx.x + i * gridDim.x;
    const int local = tile % (n_tb * n_ub);
    TileInfo ti{(tile / (n_tb * n_ub)) % B, (local / n_ub) * kTT, (local % n_ub) * kTU, T, U};
    __nv_bfloat16* dst = scratch + (static_cast<size_t>(blockIdx.x) * 2 + (i & 1)) * 128 * H;
    // with more than four warps the extra warps take a second share of the tile's column vectors / frames
    if (hgen_warps <= 4) {
      hgen_tile<MODE>(ti, f, g, dst, H, T, U + 1, ht);
    } else {
      // 8 warps: warps 0-3 frames 0..7, warps 4-7 frames 8..15 (H >= 1024) -- emulate by halving the tile per group
      TileInfo th = ti;
      const int grp = ht >> 7;
      th.t0 = ti.t0 + grp * 8; th.T = min(ti.T, th.t0 + 8);
      hgen_tile<MODE>(th, f, g, dst + static_cast<size_t>(grp) * 64 * H, H, T, U + 1, ht & 127);
    }
    __threadfence();
    asm volatile("fence.proxy.async.global;" ::: "memory");
    asm volatile("bar.sync 1, %0;" ::"r"(hgen_warps * 32));
  }
  if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
  const int B = 4, T = 500, U = 100;
  for (int H : {512, 1024}) {
    __nv_bfloat16 *f, *g, *scratch; unsigned long long* cyc;
    cudaMalloc(&f, sizeof(__nv_bfloat16) * B * T * H); cudaMalloc(&g, sizeof(__nv_bfloat16) * B * (U + 1) * H);
    cudaMalloc(&scratch, sizeof(__nv_bfloat16) * 148 * 2 * 128 * H); cudaMalloc(&cyc, 148 * 8);
    cudaMemset(f, 0x3c, sizeof(__nv_bfloat16) * B * T * H); cudaMemset(g, 0x3b, sizeof(__nv_bfloat16) * B * (U + 1) * H);
    const int tiles = 40;
    for (int warps : {4}) {
      for (int mode = 0; mode < 6; ++mode) {
        for (int rep = 0; rep < 2; ++rep) {
          if (mode == 0) run<0><<<148, 512>>>(f, g, scratch, H, T, U, B, tiles, warps, cyc);
          if (mode == 1) run<1><<<148, 512>>>(f, g, scratch, H, T, U, B, tiles, warps, cyc);
          if (mode == 2) run<2><<<148, 512>>>(f, g, scratch, H, T, U, B, tiles, warps, cyc);
          if (mode == 3) run<3><<<148, 512>>>(f, g, scratch, H, T, U, B, tiles, warps, cyc);
          if (mode == 4) run<4><<<148, 512>>>(f, g, scratch, H, T, U, B, tiles, warps, cyc);
          if (mode == 5) run<5><<<148, 512>>>(f, g, scratch, H, T, U, B, tiles, warps, cyc);
          cudaDeviceSynchronize();
        }
        unsigned long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double s = 0; for (auto v : h) s += v;
        printf("H=%4d hgen warps %d mode %d: %.0f cycles per 128-row tile (%d tanh; MUFU floor %.0f)  %s\n", H, warps, mode,
               s / 148 / tiles, 128 * H, 128.0 * H / 16, cudaGetErrorString(cudaGetLastError()));
      }
    }
    cudaFree(f); cudaFree(g); cudaFree(scratch); cudaFree(cyc);
  }
  return 0;
}
