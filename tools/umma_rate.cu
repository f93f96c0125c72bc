// Microbenchmark (development tool): what one issuing thread pays per tcgen05.mma / tcgen05.commit, against the UMMA N
// (token block of the prefill kernel).  One CTA; operands are zeroed shared-memory tiles in the prefill kernel's
// layouts (A: MN-major SWIZZLE_128B, B: K-major SWIZZLE_128B); a "stage" = 4 x (M=128, N, K=16) + `commits` commits;
// every `inflight` stages the thread waits for the oldest commit (as the A-buffer ring of gemm_tc.cu does).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../chatglm_q_b200/csrc/ptx.cuh"
using namespace cgq;

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return static_cast<uint64_t>((saddr >> 4) & 0x3FFF) | (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFF) << 16) |
         (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46) | (2ull << 61);
}

__global__ void __launch_bounds__(128, 1) probe(int N, int stages, int commits, int inflight, int accs, int Mrows, int warpwide, unsigned long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - ptx::smem_u32(smem_raw));
  __shared__ uint32_t slot;
  __shared__ uint64_t bars[16];
  for (int i = threadIdx.x; i < (16384 + 32768) / 16; i += blockDim.x) reinterpret_cast<uint4*>(gen)[i] = make_uint4(0, 0, 0, 0);
  if (threadIdx.x == 0)
    for (int i = 0; i < 16; ++i) ptx::mbar_init(&bars[i], 1);
  ptx::fence_mbar_init();
  ptx::fence_proxy_async_smem();
  if (threadIdx.x < 32) {
    ptx::tmem_alloc(&slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warpwide && threadIdx.x < 32) {
    // every lane of warp 0 runs the loop, lane 0 issues (predicated inside the asm)
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (0u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
                           (static_cast<uint32_t>(Mrows >> 4) << 24);
    const uint64_t adesc = make_desc(base, 8 * 1024, 1024), bdesc = make_desc(base + 16384, 16, 1024);
    const uint32_t issue = threadIdx.x == 0 ? 1u : 0u;
    int ph = 0, b = 0;
    const long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      if (s >= inflight) ptx::mbar_wait(&bars[b], ph);
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) ptx::umma_f16_ss_warp(tmem, adesc + k4 * 128, bdesc + k4 * 2, idesc, 1u, issue);
      ptx::umma_commit_warp(&bars[b], issue);
      if (++b == inflight) {
        b = 0;
        if (s >= inflight) ph ^= 1;
      }
    }
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = out[1] = t1 - t0;
  } else if (!warpwide && threadIdx.x == 0) {
    const uint32_t idesc = (1u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (0u << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
                           (static_cast<uint32_t>(Mrows >> 4) << 24);
    const uint64_t adesc = make_desc(base, 8 * 1024, 1024), bdesc = make_desc(base + 16384, 16, 1024);
    int ph[16] = {0};
    const long long t0 = clock64();
    for (int s = 0; s < stages; ++s) {
      const int b = s % inflight;
      if (s >= inflight) {
        ptx::mbar_wait(&bars[b], ph[b]);
        ph[b] ^= 1;
      }
#pragma unroll
      if (accs == 1) {                 // constant operands: nothing but the four instructions
        ptx::umma_f16_ss(tmem, adesc, bdesc, idesc, 1u);
        ptx::umma_f16_ss(tmem, adesc, bdesc, idesc, 1u);
        ptx::umma_f16_ss(tmem, adesc, bdesc, idesc, 1u);
        ptx::umma_f16_ss(tmem, adesc, bdesc, idesc, 1u);
      } else {
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4)   // two independent accumulators, N columns apart
          ptx::umma_f16_ss(tmem + static_cast<uint32_t>((k4 & 1) * N), adesc, bdesc, idesc, 1u);
      }
      for (int c = 0; c < commits; ++c) ptx::umma_commit(&bars[c == 0 ? b : 15]);
    }
    const long long t1 = clock64();
    for (int b = 0; b < inflight && b < stages; ++b) ptx::mbar_wait(&bars[b], ph[b]);
    const long long t2 = clock64();
    out[0] = t1 - t0;
    out[1] = t2 - t0;
  }
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc(tmem, 512);
}

int main() {
  unsigned long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int stages = 512;
  for (int warpwide : {0, 1})
  for (int Mrows : {128})
  for (int N : {16, 128, 256})
    for (int accs : {1})
      for (int commits : {1}) {
        const int inflight = 4;
        probe<<<1, 128, 64 * 1024>>>(N, stages, commits, inflight, accs, Mrows, warpwide, d);
        cudaDeviceSynchronize();
        probe<<<1, 128, 64 * 1024>>>(N, stages, commits, inflight, accs, Mrows, warpwide, d);
        cudaDeviceSynchronize();
        unsigned long long h[2];
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("%s M=%3d N=%3d accumulators=%d: %.0f cycles per stage (4 UMMA k16 + 1 commit) issued, %.0f incl. drain  (%s)\n", warpwide ? "all lanes run the loop, lane 0 issues:" : "if (lane == 0) branch:", Mrows, N,
               accs, (double)h[0] / stages, (double)h[1] / stages, cudaGetErrorString(cudaGetLastError()));
      }
  return 0;
}
