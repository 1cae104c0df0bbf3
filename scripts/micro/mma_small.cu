// Microbenchmark 3: cost of one tcgen05.mma (cta_group::1, bf16, K = 16) at small N, for M = 128 and M = 64, issued
// by a converged warp through elect.sync.  Answers: is the ~110 cycles per MMA seen in the decode kernel (N = 16) a
// fixed cost of the instruction, and does M = 64 halve it?  Values are not checked; only time is measured.
#include <cstdio>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../../myrtlespeech_b200/csrc/ptx.cuh"
using namespace rnnt;

struct Cfg { int M; int N; int alt_d; int n_instr; int n_issuers; int commit_each; int wait_each; };

__global__ void __launch_bounds__(128, 1) mma_small_kernel(Cfg c, long long* out_cycles) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t done_bar[2];
  __shared__ uint64_t dummy_bar[8];
  __shared__ uint64_t ready_bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 192 * 1024 / 16; i += blockDim.x) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0) { mbar_init(&done_bar[0], 1); mbar_init(&done_bar[1], 1); for (int i = 0; i < 8; ++i) mbar_init(&dummy_bar[i], 1); mbar_init(&ready_bar, 1); mbar_arrive(&ready_bar); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (warp >= 1 && warp <= c.n_issuers) {
    const int w = warp - 1;
    const uint32_t idesc = make_idesc_bf16(c.M, c.N, false, false);
    const uint32_t base = smem_u32(smem);
    const uint64_t ad0 = make_smem_desc_sw128(base, 16, 1024), bd0 = make_smem_desc_sw128(base + 6 * 16384, 16, 1024);
    const long long t0 = clock64();
    for (int i = 0; i < c.n_instr; i += 4) {
      const int stage = (i >> 2) % 6;
      const uint64_t ad = ad0 + stage * (16384 >> 4), bd = bd0 + ((i >> 2) & 7) * (2048 >> 4);
      if (c.wait_each) { mbar_wait(&ready_bar, 0); tc_fence_after(); }   // already complete: cost of the poll + fence
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_bf16(tmem + w * 256 + (c.alt_d ? kk * 32 : 0), ad + 2 * kk, bd + 2 * kk, idesc, 1u);
        if (c.commit_each) umma_commit(&dummy_bar[(i >> 2) & 7]);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&done_bar[w]);
    __syncwarp();
    mbar_wait(&done_bar[w], 0);
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out_cycles[w] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d_out; cudaMalloc(&d_out, sizeof(long long) * 4);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(mma_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int Ms[] = {128, 64}, Ns[] = {16, 32, 64, 128, 256};
  for (int mode = 0; mode < 4; ++mode)
    for (int M : Ms)
      for (int N : Ns)
        for (int alt = 0; alt < 1; ++alt) {
          if (mode && N > 16) continue;
          const int issuers = 1;
          Cfg c{M, N, alt, 4096, issuers, mode & 1, mode >> 1};
          for (int rep = 0; rep < 2; ++rep) {
            mma_small_kernel<<<1, 128, smem>>>(c, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error %s (M=%d N=%d)\n", cudaGetErrorString(e), M, N); return 1; }
          }
          long long h[2]; cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
          const long long mx = issuers == 2 && h[1] > h[0] ? h[1] : h[0];
          printf("commit/k-block=%d wait+fence/k-block=%d M=%3d N=%3d altD=%d : %6.1f cyc/instr per issuer (%.1f cyc per MMA on the SM)\n", c.commit_each, c.wait_each, M, N, alt,
                 (double)mx / c.n_instr, (double)mx / c.n_instr / issuers);
        }
  return 0;
}
