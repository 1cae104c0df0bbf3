// Throughput of the dh epilogue's two halving shuffle butterflies (df over 8 label positions, dg over 4 frames) as a
// function of the warps per SM that run them concurrently.  Build: nvcc -arch=sm_100a -O3 -o butterfly_rate butterfly_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float shfl_xor_f(float v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
__device__ __forceinline__ void reduce_over_positions(const float (&v)[32], int lane, float (&out)[4]) {
  float a[16], b[8];
  const bool h4 = lane & 4, h2 = lane & 2, h1 = lane & 1;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float keep = h4 ? v[16 + i] : v[i], send = h4 ? v[i] : v[16 + i]; a[i] = keep + shfl_xor_f(send, 4); }
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float keep = h2 ? a[8 + i] : a[i], send = h2 ? a[i] : a[8 + i]; b[i] = keep + shfl_xor_f(send, 2); }
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float keep = h1 ? b[4 + i] : b[i], send = h1 ? b[i] : b[4 + i]; out[i] = keep + shfl_xor_f(send, 1); }
}
__device__ __forceinline__ void reduce_over_frames(const float (&v)[32], int lane, float (&out)[8]) {
  float a[16];
  const bool h16 = lane & 16, h8 = lane & 8;
#pragma unroll
  for (int i = 0; i < 16; ++i) { const float keep = h16 ? v[16 + i] : v[i], send = h16 ? v[i] : v[16 + i]; a[i] = keep + shfl_xor_f(send, 16); }
#pragma unroll
  for (int i = 0; i < 8; ++i) { const float keep = h8 ? a[8 + i] : a[i], send = h8 ? a[i] : a[8 + i]; out[i] = keep + shfl_xor_f(send, 8); }
}
// mode 0: both butterflies; 1: only the FSEL/FADD part (shuffles replaced by moves); 2: plain xor butterflies (5 x 32 shuffles, no selects)
template <int kMode>
__global__ void k(float* out, long long* cyc, int iters) {
  const int lane = threadIdx.x & 31;
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = static_cast<float>(threadIdx.x * 32 + i) * 1e-3f;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (kMode == 0) {
      float o4[4], o8[8];
      reduce_over_positions(v, lane, o4);
      reduce_over_frames(v, lane, o8);
      acc += o4[0] + o4[1] + o4[2] + o4[3] + o8[0] + o8[1] + o8[2] + o8[3] + o8[4] + o8[5] + o8[6] + o8[7];
    } else if (kMode == 1) {
      const bool h4 = lane & 4;
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 16; ++i) { const float keep = h4 ? v[16 + i] : v[i], send = h4 ? v[i] : v[16 + i]; s += keep * 1.0001f + send; }
      acc += s;
    } else {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) s += shfl_xor_f(v[i], 4);
      acc += s;
    }
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] += acc * 1e-9f;
  }
  const long long t1 = clock64();
  if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int kMode>
void run(const char* name, int warps, float* out, long long* cyc) {
  const int iters = 2000;
  k<kMode><<<148, warps * 32>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  long long h[148 * 32];
  cudaMemcpy(h, cyc, sizeof(long long) * 148 * warps, cudaMemcpyDeviceToHost);
  double s = 0;
  for (int i = 0; i < 148 * warps; ++i) s += h[i];
  printf("%-34s warps/SM %2d: %7.1f cycles per iteration per warp\n", name, warps, s / (148.0 * warps) / iters);
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 148 * 32 * 8);
  for (int w : {4, 8, 12, 16}) run<0>("df + dg butterflies (52 SHFL)", w, out, cyc);
  for (int w : {4, 8, 16}) run<1>("selects + adds only (no SHFL)", w, out, cyc);
  for (int w : {4, 8, 16}) run<2>("32 SHFL + 32 FADD", w, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
