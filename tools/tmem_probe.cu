// Probe: register <-> (lane, column) mapping of tcgen05.st.16x256b.x1 (read back with 32x32b.x8).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void probe(uint32_t* out) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  // every warp w stores into its own subpartition (lanes 32w..32w+31): two 16-lane halves
  for (int half = 0; half < 2; ++half) {
    uint32_t r0 = (threadIdx.x << 8) | (half << 4) | 0, r1 = (threadIdx.x << 8) | (half << 4) | 1,
             r2 = (threadIdx.x << 8) | (half << 4) | 2, r3 = (threadIdx.x << 8) | (half << 4) | 3;
    uint32_t taddr = base + ((uint32_t)(warp * 32 + half * 16) << 16);
    asm volatile("tcgen05.st.sync.aligned.16x256b.x1.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr), "r"(r0), "r"(r1), "r"(r2), "r"(r3) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t v[8];
  uint32_t taddr = base + ((uint32_t)(warp * 32) << 16);
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  for (int c = 0; c < 8; ++c) out[threadIdx.x * 8 + c] = v[c];
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(base) : "memory");
}
int main() {
  uint32_t* d; cudaMalloc(&d, 128 * 8 * 4);
  probe<<<1, 128>>>(d);
  uint32_t h[128 * 8]; cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  printf("lane: col0..col7 as (src thread, half, reg)\n");
  for (int l = 0; l < 40; ++l) {
    printf("lane %3d:", l);
    for (int c = 0; c < 8; ++c) { uint32_t v = h[l * 8 + c]; printf(" (T%u,h%u,r%u)", v >> 8, (v >> 4) & 1, v & 15); }
    printf("\n");
  }
  return 0;
}
