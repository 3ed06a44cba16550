// Scratch micro-benchmark (not part of the library): aggregate L2 -> SM read bandwidth on this B200 for a working
// set that fits the 126 MB L2, with plain 16-byte read-only loads from every SM — the "third roofline" the fused
// gather kernels run into.    nvcc -O3 -gencode arch=compute_100a,code=sm_100a l2bw.cu -o l2bw
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

__global__ void read_kernel(const float4* __restrict__ p, size_t n_vec, int iters, float* sink) {
  float acc = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int it = 0; it < iters; ++it)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += stride) {
      const float4 v = __ldg(p + i);
      acc += v.x + v.y + v.z + v.w;
    }
  if (acc == 123.456f) *sink = acc;
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  printf("%s, %d SMs, L2 %.0f MB\n", prop.name, prop.multiProcessorCount, prop.l2CacheSize / 1048576.0);
  float* sink;
  cudaMalloc(&sink, 4);
  const int mbs[] = {8, 16, 32, 64, 96, 256, 1024};
  for (int mb : mbs) {
    const size_t bytes = (size_t)mb << 20;
    float4* buf;
    cudaMalloc(&buf, bytes);
    cudaMemset(buf, 0, bytes);
    const int iters = mb <= 96 ? 40 : 4;
    for (int bps : {2, 8}) {
      const int grid = prop.multiProcessorCount * bps;
      read_kernel<<<grid, 512>>>(buf, bytes / 16, 2, sink);  // warm the L2
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0), cudaEventCreate(&e1);
      cudaEventRecord(e0);
      read_kernel<<<grid, 512>>>(buf, bytes / 16, iters, sink);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("working set %5d MB, %d CTAs/SM x 512 thr: %.0f GB/s\n", mb, bps, (double)bytes * iters / (ms * 1e-3) / 1e9);
    }
    cudaFree(buf);
  }
  return 0;
}
