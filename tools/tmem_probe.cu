// Probe: register <-> (lane, column) mapping of tcgen05.st.16x256b.x2, read back with tcgen05.ld.32x32b (thread = lane).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int t = threadIdx.x;
  if (t < 32) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  if (t < 32) {
    // clear 32 lanes x 16 columns
    uint32_t z[16];
    for (int i = 0; i < 16; ++i) z[i] = 0xFFFFu;
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(base),
                 "r"(z[0]), "r"(z[1]), "r"(z[2]), "r"(z[3]), "r"(z[4]), "r"(z[5]), "r"(z[6]), "r"(z[7]), "r"(z[8]), "r"(z[9]), "r"(z[10]),
                 "r"(z[11]), "r"(z[12]), "r"(z[13]), "r"(z[14]), "r"(z[15]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t r[8];
    for (int i = 0; i < 8; ++i) r[i] = 100u * t + i;           // first half: lanes 0-15
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base), "r"(r[0]), "r"(r[1]), "r"(r[2]),
                 "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    for (int i = 0; i < 8; ++i) r[i] = 100u * t + i + 50000u;  // second half: lanes 16-31
    asm volatile("tcgen05.st.sync.aligned.16x256b.x2.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(base + (16u << 16)), "r"(r[0]), "r"(r[1]),
                 "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(base) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; ++i) out[t * 16 + i] = v[i];
  }
  __syncthreads();
  if (t < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 32 * 16 * 4);
  probe<<<1, 128>>>(d);
  uint32_t h[512];
  cudaError_t e = cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("status %s\n", cudaGetErrorString(e));
  for (int lane = 0; lane < 32; ++lane) {
    printf("lane %2d:", lane);
    for (int c = 0; c < 16; ++c) printf(" %5u", h[lane * 16 + c]);
    printf("\n");
  }
  return 0;
}
