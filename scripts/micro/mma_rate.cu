// Microbenchmark: tcgen05.mma issue->complete rate for several configurations (no TMA; smem contents are
// whatever they are -- we measure time, not values).  One CTA per SM (or pair), 148 CTAs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../myrtlespeech_b200/csrc/ptx.cuh"
using namespace rnnt;

struct Cfg { int pair; int N; int alt_d; int mn_major; int n_instr; int k_adv; };

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Cfg c, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = c.pair ? cluster_ctarank() : 0;
  // zero operands so no NaN/denormal oddities
  for (int i = threadIdx.x; i < 160 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done_bar, 1); fence_barrier_init(); }
  if (warp == 0) { if (c.pair) { tmem_alloc_2cta(&tmem_slot, 512); tmem_relinquish_2cta(); } else { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); } }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); if (c.pair) cluster_sync_all(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  long long t0 = 0, t1 = 0;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = make_idesc_bf16(c.pair ? 256 : 128, c.N, c.mn_major, c.mn_major);
    const uint32_t a_addr = smem_u32(smem), b_addr = a_addr + 64 * 1024;
    t0 = clock64();
    if (lane == 0 && c.k_adv < 0) {
      const uint64_t ad = make_smem_desc_sw128(a_addr, 16, 1024), bd = make_smem_desc_sw128(b_addr, 16, 1024);
      for (int i = 0; i < c.n_instr; i += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { if (c.pair) umma_bf16_pair(tmem, ad, bd, idesc, 1u); else umma_bf16(tmem, ad, bd, idesc, 1u); }
      }
      if (c.pair) umma_commit_pair(&done_bar, 1); else umma_commit(&done_bar);
    } else if (lane == 0) {
      for (int i = 0; i < c.n_instr; ++i) {
        const int kk = i & 3;
        const int stage = (i >> 2) & 1;
        uint64_t ad, bd;
        if (c.mn_major) {
          ad = make_smem_desc_sw128(a_addr + stage * 32768 + kk * 2048, 8192, 1024);
          bd = make_smem_desc_sw128(b_addr + stage * 32768 + kk * 2048, 8192, 1024);
        } else {
          ad = make_smem_desc_sw128(a_addr + stage * 16384 + kk * c.k_adv, 16, 1024);
          bd = make_smem_desc_sw128(b_addr + stage * 32768 + kk * c.k_adv, 16, 1024);
        }
        const uint32_t d = tmem + (c.alt_d ? ((i & 1) * 256) : 0);
        if (c.pair) umma_bf16_pair(d, ad, bd, idesc, 1u); else umma_bf16(d, ad, bd, idesc, 1u);
      }
      if (c.pair) umma_commit_pair(&done_bar, 1); else umma_commit(&done_bar);
    }
    __syncwarp();
    mbar_wait(&done_bar, 0);
    t1 = clock64();
    if (lane == 0) out_cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before(); __syncthreads(); if (c.pair) cluster_sync_all();
  if (warp == 0) { tc_fence_after(); if (c.pair) tmem_dealloc_2cta(tmem, 512); else tmem_dealloc(tmem, 512); }
}

int main() {
  int n_sm = 148;
  long long* d_out; cudaMalloc(&d_out, sizeof(long long) * 148);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  Cfg cfgs[] = {
    {0, 256, 0, 0, 1024, 32}, {0, 256, 1, 0, 1024, 32},
    {0, 256, 0, 0, 1024, -1}, {0, 128, 0, 0, 1024, -1}, {0, 64, 0, 0, 1024, -1}, {0, 32, 0, 0, 1024, -1},
    {1, 256, 0, 0, 1024, -1}, {1, 192, 0, 0, 1024, -1}, {1, 128, 0, 0, 1024, -1}, {1, 64, 0, 0, 1024, -1}, {1, 32, 0, 0, 1024, -1},
    {0, 256, 0, 1, 1024, 32},
  };
  for (int grid_mode = 0; grid_mode < 2; ++grid_mode) {
    for (auto& c : cfgs) {
      int grid = grid_mode == 0 ? (c.pair ? 2 : 1) : n_sm;
      cudaLaunchConfig_t lc{}; lc.gridDim = dim3(grid); lc.blockDim = dim3(128); lc.dynamicSmemBytes = smem; lc.stream = 0;
      cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
      at[0].val.clusterDim.x = c.pair ? 2 : 1; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      lc.attrs = at; lc.numAttrs = 1;
      cudaMemset(d_out, 0, sizeof(long long) * 148);
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        cudaError_t le = cudaSuccess;
        if (c.pair) le = cudaLaunchKernelEx(&lc, mma_rate_kernel, c, d_out);
        else { mma_rate_kernel<<<grid, 128, smem>>>(c, d_out); le = cudaGetLastError(); }
        if (le != cudaSuccess) { printf("launch error %s\n", cudaGetErrorString(le)); }
        cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("CUDA error %s\n", cudaGetErrorString(e)); return 1; }
      }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      long long h[148]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0; for (int i = 0; i < grid; ++i) if (h[i] > mx) mx = h[i];
      const double macs_per_sm = 128.0 * c.N * 16 * c.n_instr;
      printf("grid=%3d pair=%d N=%3d altD=%d mn=%d kadv=%2d : %7.1f cyc/instr  %6.0f MAC/clk/SM  kernel %.1f us\n", grid, c.pair, c.N,
             c.alt_d, c.mn_major, c.k_adv, (double)mx / c.n_instr, macs_per_sm / (double)mx, ms * 1e3);
    }
  }
  return 0;
}
