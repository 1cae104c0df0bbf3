// Microbenchmark 2: tcgen05.mma rate under realistic operand addressing, and L2 -> smem TMA bandwidth.
//   part A: M=256 (CTA pair) / M=128 (single CTA) x N x K=16 bf16 MMAs; same or alternating accumulators;
//           one smem address or a ring of stage addresses; K-major or MN-major operands.
//   part B: TMA-only streaming of 16 KB 128B-swizzled boxes out of an L2-resident buffer, all SMs.
// Values are not checked (smem is zero-filled); only time is measured.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include "../../myrtlespeech_b200/csrc/ptx.cuh"
using namespace rnnt;

struct Cfg { int N; int alt_d; int ring; int mn_major; int n_instr; int burst; };

template <int PAIR>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Cfg c, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint64_t dummy_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t rank = 0;
  if (PAIR) rank = cluster_ctarank();
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); mbar_init(&dummy_bar, 1); fence_barrier_init(); }
  if (warp == 0) {
    if constexpr (PAIR) { tmem_alloc_2cta(&tmem_slot, 512); tmem_relinquish_2cta(); } else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); if (PAIR) cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(PAIR ? 256 : 128, c.N, c.mn_major, c.mn_major);
    const uint32_t base = smem_u32(smem);
    long long t0 = clock64();
    if (PAIR && lane == 0 && c.burst > 0) {
      if constexpr (PAIR) {
      // real-kernel issue pattern: per k-block, `burst` MMAs back to back from one descriptor base (+32 B per k16), then a commit
      for (int i = 0; i < c.n_instr; i += c.burst) {
        const int stage = c.ring ? ((i / c.burst) % 6) : 0;
        const uint32_t a_addr = base + stage * 32768, b_addr = a_addr + 16384;
        const uint64_t ad0 = make_smem_desc_sw128(a_addr, 16, 1024), bd0 = make_smem_desc_sw128(b_addr, 16, 1024);
        const uint32_t d = tmem + (c.alt_d ? (((i / c.burst) & 1) * 256) : 0);
        if (c.burst == 4) {
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) umma_bf16_pair(d, ad0 + kk * 2, bd0 + kk * 2, idesc, 1u);
        } else {
#pragma unroll
          for (int kk = 0; kk < 8; ++kk) umma_bf16_pair(d, ad0 + (kk & 3) * 2, bd0 + (kk & 3) * 2, idesc, 1u);
        }
        if (c.burst != 9) umma_commit_pair(&dummy_bar, 1);
      }
      umma_commit_pair(&done_bar, 1);
      }
    } else if (lane == 0) {
      for (int i = 0; i < c.n_instr; ++i) {
        const int kk = i & 3;
        const int stage = c.ring ? ((i >> 2) % 6) : 0;
        const uint32_t a_addr = base + stage * 32768, b_addr = a_addr + 16384;
        uint64_t ad, bd;
        if (c.mn_major) {
          ad = make_smem_desc_sw128(a_addr + (c.ring ? kk * 2048 : 0), 8192, 1024);
          bd = make_smem_desc_sw128(b_addr + (c.ring ? kk * 2048 : 0), 8192, 1024);
        } else {
          ad = make_smem_desc_sw128(a_addr + (c.ring ? kk * 32 : 0), 16, 1024);
          bd = make_smem_desc_sw128(b_addr + (c.ring ? kk * 32 : 0), 16, 1024);
        }
        const uint32_t d = tmem + (c.alt_d ? ((i & 1) * 256) : 0);
        if constexpr (PAIR) umma_bf16_pair(d, ad, bd, idesc, 1u); else umma_bf16(d, ad, bd, idesc, 1u);
      }
      if constexpr (PAIR) umma_commit_pair(&done_bar, 1); else umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    long long t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); if (PAIR) cluster_sync_all();
  if (warp == 0) { tc_fence_after(); if constexpr (PAIR) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}

// ---- part B: TMA streaming ---------------------------------------------------------------------
constexpr int kStages = 6;
constexpr int kBox = 16384;
__global__ void __launch_bounds__(64, 1)
tma_bw_kernel(const __grid_constant__ CUtensorMap tm, int rows_per_cta, int shared_rows, int n_iter, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full[kStages];
  if (threadIdx.x == 0) { for (int i = 0; i < kStages; ++i) mbar_init(&full[i], 1); fence_barrier_init(); }
  __syncthreads();
  if (threadIdx.x == 0) {
    const long long t0 = clock64();
    // private region: rows [bid*rows_per_cta, +rows_per_cta); shared_rows > 0: every CTA reads rows [0, shared_rows)
    const int row_base = shared_rows > 0 ? 0 : blockIdx.x * rows_per_cta;
    const int span = shared_rows > 0 ? shared_rows : rows_per_cta;
    int issued = 0, waited = 0;
    for (; issued < kStages && issued < n_iter; ++issued) {
      mbar_arrive_expect_tx(&full[issued], 2 * kBox);
      const int r = row_base + (issued * 128) % span;
      tma_load_2d(smem + issued * 2 * kBox, &tm, &full[issued], ((issued * 7) & 15) * 64, r);
      tma_load_2d(smem + issued * 2 * kBox + kBox, &tm, &full[issued], ((issued * 7 + 3) & 15) * 64, r);
    }
    for (; waited < n_iter; ++waited) {
      const int s = waited % kStages;
      mbar_wait(&full[s], (waited / kStages) & 1);
      if (issued < n_iter) {
        mbar_arrive_expect_tx(&full[s], 2 * kBox);
        const int r = row_base + (issued * 128) % span;
        tma_load_2d(smem + s * 2 * kBox, &tm, &full[s], ((issued * 7) & 15) * 64, r);
        tma_load_2d(smem + s * 2 * kBox + kBox, &tm, &full[s], ((issued * 7 + 3) & 15) * 64, r);
        ++issued;
      }
    }
    out_cycles[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int n_sm = 148;
  long long* d_out; cudaMalloc(&d_out, sizeof(long long) * 148);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(mma_rate_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct Run { int pair; Cfg c; };
  Run runs[] = {
    {1, {256, 0, 0, 0, 2048, 0}}, {1, {256, 0, 1, 0, 2048, 0}}, {1, {256, 0, 1, 1, 2048, 0}},
    {1, {256, 0, 0, 0, 2048, 4}}, {1, {256, 0, 1, 0, 2048, 4}}, {1, {256, 1, 1, 0, 2048, 4}}, {1, {256, 0, 1, 0, 2048, 8}},
    {1, {128, 0, 1, 0, 2048, 0}}, {1, {128, 0, 1, 0, 2048, 4}}, {1, {64, 0, 1, 0, 2048, 4}},
    {0, {256, 0, 0, 0, 2048, 0}}, {0, {128, 0, 0, 0, 2048, 0}},
  };
  for (int grid_mode = 0; grid_mode < 2; ++grid_mode) {
    for (auto& r : runs) {
      const int grid = grid_mode == 0 ? (r.pair ? 2 : 1) : n_sm;
      cudaLaunchConfig_t lc{}; lc.gridDim = dim3(grid); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem; lc.stream = 0;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = r.pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      cudaMemset(d_out, 0, sizeof(long long) * 148);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaError_t le = r.pair ? cudaLaunchKernelEx(&lc, mma_rate_kernel<1>, r.c, d_out)
                                : cudaLaunchKernelEx(&lc, mma_rate_kernel<0>, r.c, d_out);
        if (le != cudaSuccess) printf("launch error %s\n", cudaGetErrorString(le));
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
      const double macs_per_sm = 128.0 * r.c.N * 16 * r.c.n_instr;
      printf("grid=%3d pair=%d N=%3d altD=%d ring=%d mn=%d burst=%d : %7.1f cyc/instr  %6.0f MAC/clk/SM  kernel %.1f us\n", grid, r.pair,
             r.c.N, r.c.alt_d, r.c.ring, r.c.mn_major, r.c.burst, (double)mx / r.c.n_instr, macs_per_sm / (double)mx, ms * 1e3);
      fflush(stdout);
    }
  }

  // ---- part B ----
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  { void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q); enc = (PFN_cuTensorMapEncodeTiled_v12000)p; }
  const uint64_t cols = 1024, rows = 148ull * 128;   // 38.8 MB of bf16, L2-resident
  void* buf; cudaMalloc(&buf, cols * rows * 2); cudaMemset(buf, 0, cols * rows * 2);
  CUtensorMap tm;
  cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {cols * 2}; cuuint32_t box[2] = {64, 128}; cuuint32_t es[2] = {1, 1};
  CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
  cudaFuncSetAttribute(tma_bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  struct BRun { int rows_per_cta; int shared_rows; const char* name; };
  BRun bruns[] = {{128, 0, "private 256KB/CTA"}, {0, 1024, "all CTAs read the same 2MB"}, {128, 0, "private again"}};
  for (auto& b : bruns) {
    const int n_iter = 4096;
    for (int rep = 0; rep < 2; ++rep) {
      tma_bw_kernel<<<n_sm, 64, smem>>>(tm, b.rows_per_cta, b.shared_rows, n_iter, d_out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
    }
    long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0; double avg = 0; for (int i = 0; i < n_sm; ++i) { if (h[i] > mx) mx = h[i]; avg += h[i]; }
    avg /= n_sm;
    printf("TMA stream [%s]: %.1f B/clk/SM (slowest CTA), %.1f B/clk/SM (mean); chip %.0f B/clk\n", b.name,
           (double)n_iter * 32768 / mx, (double)n_iter * 32768 / avg, 148.0 * n_iter * 32768 / mx);
  }
  return 0;
}
