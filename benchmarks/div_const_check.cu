// Exhaustive checks over all 2^32 float bit patterns: frcnn::div_const<10> / <5> (csrc/common.cuh) against __fdiv_rn,
// and frcnn::np_expf_mid against frcnn::np_expf on its range |x| <= EXP_MID_LIMIT.
//   nvcc -gencode arch=compute_100a,code=sm_100a -I include -I faster_rcnn_b200/csrc -o /tmp/div_check benchmarks/div_const_check.cu && /tmp/div_check
#include <cstdio>

#include "common.cuh"

template <int D>
__global__ void check(unsigned long long* bad, unsigned* first_bad) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
    const float x = __uint_as_float((unsigned)b);
    const unsigned got = __float_as_uint(frcnn::div_const<D>(x)), want = __float_as_uint(__fdiv_rn(x, (float)D));
    const bool both_nan = (got & 0x7fffffffu) > 0x7f800000u && (want & 0x7fffffffu) > 0x7f800000u;
    if (got != want && !both_nan) { ++local; atomicMin(first_bad, (unsigned)b); }
  }
  if (local) atomicAdd(bad, local);
}

__global__ void check_exp(unsigned long long* bad, unsigned* first_bad) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  for (unsigned long long b = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; b < (1ull << 32); b += stride) {
    const float x = __uint_as_float((unsigned)b);
    if (!(fabsf(x) <= frcnn::EXP_MID_LIMIT)) continue;
    if (__float_as_uint(frcnn::np_expf_mid(x)) != __float_as_uint(frcnn::np_expf(x))) { ++local; atomicMin(first_bad, (unsigned)b); }
  }
  if (local) atomicAdd(bad, local);
}

int main() {
  unsigned long long* bad;
  unsigned* first;
  cudaMallocManaged(&bad, 16);
  cudaMallocManaged(&first, 8);
  int rc = 0;
  for (int d = 0; d < 2; ++d) {
    *bad = 0;
    *first = 0xffffffffu;
    if (d == 0) check<10><<<148 * 8, 256>>>(bad, first);
    else check<5><<<148 * 8, 256>>>(bad, first);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("DIVCHECK cuda error\n"); return 2; }
    printf("DIVCHECK D=%d mismatches=%llu first=0x%08x\n", d == 0 ? 10 : 5, *bad, *first);
    if (*bad) rc = 1;
  }
  *bad = 0;
  *first = 0xffffffffu;
  check_exp<<<148 * 8, 256>>>(bad, first);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("DIVCHECK cuda error\n"); return 2; }
  printf("EXPCHECK np_expf_mid vs np_expf mismatches=%llu first=0x%08x\n", *bad, *first);
  if (*bad) rc = 1;
  return rc;
}
