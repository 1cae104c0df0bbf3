// Semantics check: cp.async.bulk.tensor ... cta_group::2 ... multicast::cluster in a cluster of 4 CTAs (two
// cta_group::2 pairs).  Each CTA loads a quarter of a 256-row operand and multicasts it to the CTA with the same
// rank-in-pair of the other pair; the full barrier lives in each pair's even CTA and must see the bytes that land
// in both CTAs of the pair (its own two quarters x 2 CTAs).
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../myrtlespeech_b200/csrc/ptx.cuh"
using namespace rnnt;

__device__ __forceinline__ void tma_load_2d_pair_mcast(uint32_t smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                       uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

__global__ void __cluster_dims__(4, 1, 1) __launch_bounds__(128, 1)
mcast_kernel(const __grid_constant__ CUtensorMap tm, const __nv_bfloat16* src, int* result) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t full_bar;
  const uint32_t r4 = cluster_ctarank();
  const uint32_t half = r4 & 1, q = r4 >> 1;
  for (int i = threadIdx.x; i < 16384 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0xFFFFFFFFu;
  if (threadIdx.x == 0) { mbar_init(&full_bar, 1); fence_barrier_init(); }
  fence_proxy_async_smem();
  __syncthreads();
  cluster_sync_all();
  if (threadIdx.x == 0) {
    if (half == 0) mbar_arrive_expect_tx(&full_bar, 2 * 16384);   // both CTAs of the pair receive 2 x 8 KB
    const uint16_t mask = static_cast<uint16_t>((1u << half) | (1u << (half + 2)));
    tma_load_2d_pair_mcast(smem_u32(smem) + q * 8192, &tm, &full_bar, 0, half * 128 + q * 64, mask);
  }
  int ok_wait = 1;
  if (half == 0) {
    unsigned spins = 0;
    while (!mbar_try_wait(&full_bar, 0)) { if (++spins > (1u << 22)) { ok_wait = 0; break; } }
  }
  cluster_sync_all();   // the odd CTAs learn about completion through the cluster barrier (test only)
  __syncthreads();
  // verify: smem row rr (0..127) of this CTA == src row half*128 + rr, 64 bf16 per row, 128B swizzle
  int bad = 0;
  for (int i = threadIdx.x; i < 128 * 64; i += blockDim.x) {
    const int rr = i >> 6, c = i & 63;
    const int chunk = (c >> 3) ^ (rr & 7);
    const __nv_bfloat16 got = *reinterpret_cast<const __nv_bfloat16*>(smem + rr * 128 + chunk * 16 + (c & 7) * 2);
    const __nv_bfloat16 want = src[(half * 128 + rr) * 64 + c];
    if (__bfloat162float(got) != __bfloat162float(want)) ++bad;
  }
  atomicAdd(&result[blockIdx.x * 2], bad);
  if (threadIdx.x == 0) result[blockIdx.x * 2 + 1] = ok_wait;
}

int main() {
  PFN_cuTensorMapEncodeTiled_v12000 enc = nullptr;
  { void* p = nullptr; cudaDriverEntryPointQueryResult qr;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr); enc = (PFN_cuTensorMapEncodeTiled_v12000)p; }
  const int rows = 256, cols = 64;
  __nv_bfloat16* h = new __nv_bfloat16[rows * cols];
  for (int i = 0; i < rows * cols; ++i) h[i] = __float2bfloat16((float)((i * 7) % 251));
  __nv_bfloat16* d; cudaMalloc(&d, rows * cols * 2); cudaMemcpy(d, h, rows * cols * 2, cudaMemcpyHostToDevice);
  int* res; cudaMalloc(&res, 64 * sizeof(int)); cudaMemset(res, 0, 64 * sizeof(int));
  CUtensorMap tm;
  cuuint64_t dims[2] = {cols, rows}; cuuint64_t strides[1] = {cols * 2}; cuuint32_t box[2] = {64, 64}; cuuint32_t es[2] = {1, 1};
  CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) { printf("encode failed %d\n", (int)cr); return 1; }
  cudaFuncSetAttribute(mcast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  mcast_kernel<<<8, 128, 64 * 1024>>>(tm, d, res);
  cudaError_t e = cudaDeviceSynchronize();
  printf("kernel: %s\n", cudaGetErrorString(e));
  int hr[64]; cudaMemcpy(hr, res, sizeof(hr), cudaMemcpyDeviceToHost);
  for (int b = 0; b < 8; ++b) printf("cta %d (rank %d): mismatches %d, leader wait ok %d\n", b, b & 3, hr[b * 2], hr[b * 2 + 1]);
  return 0;
}
