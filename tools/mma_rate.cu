// Microbenchmarks that decide the decode-kernel design (development tool):
//   1. issue rate of legacy mma.sync on sm_100a: HMMA m16n8k16 (f16) vs QMMA m16n8k32 (e4m3)
//   2. fragment layout of ldmatrix.m16n16.x1.trans.b8
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int KIND, int CHAINS>
__global__ void mma_loop(int iters, unsigned long long* out, float* sink) {
  float d[CHAINS][4];
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) d[c][0] = d[c][1] = d[c][2] = d[c][3] = 0.f;
  uint32_t a[4] = {threadIdx.x, 2, 3, 4}, b0 = 5, b1 = 6;
  unsigned long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) {
      if (KIND == 0)
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      else if (KIND == 2) {
        int* di = reinterpret_cast<int*>(d[c]);
        asm volatile(
            "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+r"(di[0]), "+r"(di[1]), "+r"(di[2]), "+r"(di[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
      } else
        asm volatile(
            "mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
            "{%0,%1,%2,%3};"
            : "+f"(d[c][0]), "+f"(d[c][1]), "+f"(d[c][2]), "+f"(d[c][3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
  }
  unsigned long long t1 = clock64();
  float s = 0;
#pragma unroll
  for (int c = 0; c < CHAINS; ++c) s += d[c][0] + d[c][1] + d[c][2] + d[c][3];
  if (s == 123.f) sink[0] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0;
}

__global__ void ldsm_probe(uint32_t* out) {
  __shared__ __align__(128) uint8_t m[16 * 16];
  for (int i = threadIdx.x; i < 256; i += 32) m[i] = (uint8_t)i;  // value = row*16 + col
  __syncwarp();
  uint32_t r0, r1;
  uint32_t addr = (uint32_t)__cvta_generic_to_shared(&m[(threadIdx.x & 15) * 16]);
  asm volatile("ldmatrix.sync.aligned.m16n16.x1.trans.shared.b8 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
  out[threadIdx.x * 2] = r0;
  out[threadIdx.x * 2 + 1] = r1;
}

template <int KIND, int CHAINS>
void run(const char* name, int warps) {
  unsigned long long* out;
  float* sink;
  cudaMalloc(&out, 8);
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  mma_loop<KIND, CHAINS><<<148, warps * 32>>>(iters, out, sink);
  cudaDeviceSynchronize();
  mma_loop<KIND, CHAINS><<<148, warps * 32>>>(iters, out, sink);
  cudaDeviceSynchronize();
  unsigned long long h;
  cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%s warps/CTA=%d chains=%d: %.2f cycles per MMA per warp, %.2f cycles per MMA per SMSP\n", name, warps,
         CHAINS, (double)h / iters / CHAINS, (double)h / iters / CHAINS / ((warps + 3) / 4));
}

int main() {
  run<0, 8>("HMMA.16816.f16 ", 4);
  run<0, 8>("HMMA.16816.f16 ", 8);
  run<0, 8>("HMMA.16816.f16 ", 16);
  run<0, 1>("HMMA.16816.f16 (dependent chain)", 4);
  run<1, 8>("QMMA.16832.e4m3", 4);
  run<1, 8>("QMMA.16832.e4m3", 8);
  run<1, 8>("QMMA.16832.e4m3", 16);
  run<1, 1>("QMMA.16832.e4m3 (dependent chain)", 4);
  run<2, 8>("IMMA.16832.u8.s8", 4);
  run<2, 8>("IMMA.16832.u8.s8", 8);
  run<2, 8>("IMMA.16832.u8.s8", 16);
  run<2, 4>("IMMA.16832.u8.s8 (4 chains)", 4);
  run<2, 1>("IMMA.16832.u8.s8 (dependent chain)", 4);
  uint32_t* d;
  cudaMalloc(&d, 64 * 4);
  ldsm_probe<<<1, 32>>>(d);
  uint32_t h[64];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("ldmatrix.m16n16.x1.trans.b8: source byte value = row*16+col; per thread (r0 bytes | r1 bytes) as (row,col)\n");
  for (int t = 0; t < 32; ++t) {
    printf("T%02d:", t);
    for (int r = 0; r < 2; ++r) {
      for (int b = 0; b < 4; ++b) {
        int v = (h[t * 2 + r] >> (8 * b)) & 0xFF;
        printf(" (%d,%d)", v / 16, v % 16);
      }
      printf(" |");
    }
    printf("\n");
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
