// Memory-system ceilings that bound the RoI kernels (run on the GPU box; not product code):
//   * L2->SM read bandwidth on an L2-resident buffer (the RoI forward re-reads the 9.8 MB feature map ~4x
//     per output byte; the warp-per-cell backward re-read dY rows 3.6x),
//   * DRAM read-only, write-only (st.cs) and copy bandwidth on buffers far larger than L2.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o benchmarks/_membw benchmarks/membw.cu && benchmarks/_membw
#include <cuda_runtime.h>
#include <stdio.h>

__global__ void read_kernel(const float4* __restrict__ p, size_t n, int reps, float* sink) {
  float4 acc = make_float4(0, 0, 0, 0);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (int r = 0; r < reps; ++r)
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const float4 v = __ldg(p + i);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  if (acc.x + acc.y + acc.z + acc.w == 1.2345f) *sink = acc.x;
}
__global__ void write_kernel(float4* __restrict__ p, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(p + i, make_float4(1.f, 2.f, 3.f, 4.f));
}
__global__ void copy_kernel(const float4* __restrict__ a, float4* __restrict__ b, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(b + i, __ldg(a + i));
}

template <typename F> float time_ms(F f, int iters) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); f(); cudaDeviceSynchronize();
  cudaEventRecord(a); for (int i = 0; i < iters; ++i) f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms / iters;
}

int main() {
  cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
  const int sms = pr.multiProcessorCount;
  float *big_a, *big_b, *sink; const size_t big = (size_t)4 << 30;
  cudaMalloc(&big_a, big); cudaMalloc(&big_b, big); cudaMalloc(&sink, 4);
  cudaMemset(big_a, 0, big); cudaMemset(big_b, 0, big);
  printf("{\"device\": \"%s\", \"sms\": %d", pr.name, sms);
  const size_t l2_sizes[] = {8u << 20, 16u << 20, 32u << 20, 64u << 20};
  for (size_t bytes : l2_sizes) {
    const size_t n = bytes / 16; const int reps = 64;
    float ms = time_ms([&] { read_kernel<<<sms * 8, 256>>>((const float4*)big_a, n, reps, sink); }, 5);
    printf(", \"l2_read_%zuMB_GBps\": %.0f", bytes >> 20, (double)bytes * reps / ms / 1e6);
  }
  {
    const size_t n = big / 16;
    float ms = time_ms([&] { read_kernel<<<sms * 8, 256>>>((const float4*)big_a, n, 1, sink); }, 5);
    printf(", \"dram_read_GBps\": %.0f", (double)big / ms / 1e6);
    ms = time_ms([&] { write_kernel<<<sms * 8, 256>>>((float4*)big_b, n); }, 5);
    printf(", \"dram_write_GBps\": %.0f", (double)big / ms / 1e6);
    ms = time_ms([&] { copy_kernel<<<sms * 8, 256>>>((const float4*)big_a, (float4*)big_b, n); }, 5);
    printf(", \"dram_copy_GBps\": %.0f", 2.0 * big / ms / 1e6);
  }
  printf("}\n");
  return 0;
}
