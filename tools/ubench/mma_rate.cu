// Micro-benchmark: cycles per tcgen05.mma for small-N shapes (A from TMEM or smem), one or two issuing warps.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_rate.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../nejm-brain-to-text_b200/csrc/sm100.cuh"
using namespace b2t;
__device__ __forceinline__ void umma_ts(uint32_t d, uint32_t a, uint64_t bd, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(d), "r"(a), "l"(bd), "r"(idesc), "r"(acc) : "memory");
}
template <int M, int N, bool TS, int ISSUERS>
__global__ void k(long long* out, int reps) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)raw + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar[2];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += blockDim.x) ((uint32_t*)smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tb = slot;
  constexpr uint32_t idesc = umma_idesc_bf16(M, N, 0, 0);
  if (warp < ISSUERS && (threadIdx.x & 31) == 0) {
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 16384);
    long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      const uint32_t d = tb + 384 + warp * 64 + (r & 3) * 16 * 0;
      if (TS) umma_ts(d, tb + ((r * 8) % 256), umma_smem_desc(sb + (r & 3) * 32, 16, 1024), idesc, r != 0);
      else umma_bf16(d, umma_smem_desc(sa + (r & 3) * 32, 16, 1024), umma_smem_desc(sb + (r & 3) * 32, 16, 1024), idesc, r != 0);
    }
    long long t1 = clock64();
    umma_commit(&bar[warp]);
    mbar_wait(&bar[warp], 0);
    long long t2 = clock64();
    out[warp * 2] = t1 - t0; out[warp * 2 + 1] = t2 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tb); }
}
template <int M, int N, bool TS, int ISSUERS>
void run(const char* name) {
  long long* d; cudaMalloc(&d, 64);
  auto kern = k<M, N, TS, ISSUERS>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const int reps = 480;
  for (int it = 0; it < 2; ++it) kern<<<1, 128, 64 * 1024>>>(d, reps);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4] = {0, 0, 0, 0};
  cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
  printf("%-34s issue %.1f cyc/mma, issue+complete %.1f cyc/mma%s (%s)\n", name, (double)h[0] / reps, (double)h[1] / reps,
         ISSUERS == 2 ? " [per issuer; 2 issuers]" : "", cudaGetErrorString(e));
  cudaFree(d);
}
int main() {
  run<128, 16, true, 1>("M128 N16  A=TMEM");
  run<128, 16, false, 1>("M128 N16  A=smem");
  run<64, 16, false, 1>("M64  N16  A=smem");
  run<64, 8, false, 1>("M64  N8   A=smem");
  run<128, 32, true, 1>("M128 N32  A=TMEM");
  run<128, 64, true, 1>("M128 N64  A=TMEM");
  run<128, 128, true, 1>("M128 N128 A=TMEM");
  run<128, 128, false, 1>("M128 N128 A=smem");
  run<128, 256, false, 1>("M128 N256 A=smem");
  run<128, 16, true, 2>("M128 N16  A=TMEM x2 issuers");
  return 0;
}
