// Checks tf32_rna_bits() (common.cuh) against the hardware's cvt.rna.tf32.f32 on every 509th fp32 bit pattern.
// nvcc -gencode arch=compute_100a,code=sm_100a -o tools/tf32_rna_check tools/tf32_rna_check.cu && tools/tf32_rna_check
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void check(unsigned long long* bad, unsigned long long* seen) {
  unsigned long long b = 0, n = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < (1ull << 32) / 509; i += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t bits = (uint32_t)(i * 509);
    if (((bits >> 23) & 0xFF) == 0xFF) continue;   // Inf / NaN
    uint32_t hw;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hw) : "f"(__uint_as_float(bits)));
    const uint32_t sw = (bits + 0x1000u) & 0xFFFFE000u;
    // the hardware leaves the low 13 bits unspecified in principle: compare the TF32 fields only
    if ((hw & 0xFFFFE000u) != sw) ++b;
    ++n;
  }
  atomicAdd(bad, b);
  atomicAdd(seen, n);
}
int main() {
  unsigned long long *d, h[2] = {0, 0};
  cudaMalloc(&d, 16);
  cudaMemset(d, 0, 16);
  check<<<148 * 8, 256>>>(d, d + 1);
  cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("tf32 rna bit trick vs cvt.rna.tf32.f32: %llu mismatches in %llu finite patterns\n", h[0], h[1]);
  return h[0] != 0;
}
