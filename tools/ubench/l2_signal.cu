// One-way / round-trip latency of a word handed from one SM to another through L2 on B200, for the store / load flavours
// the recurrence kernels could use.  CTA 0 and CTA k ping-pong a counter ITERS times; cycles per round trip are reported.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_signal l2_signal.cu && ./l2_signal
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

enum { ST_RELAXED = 0, ST_VOLATILE = 1, ST_RELEASE = 2, ATOM_EXCH = 3, RED_ADD = 4, ST_RELAXED_V4 = 5 };
enum { LD_RELAXED = 0, LD_CG = 1, LD_VOLATILE = 2, LD_ACQUIRE = 3 };

template <int ST> __device__ __forceinline__ void put(uint32_t* p, uint32_t v) {
  if (ST == ST_RELAXED) asm volatile("st.relaxed.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (ST == ST_VOLATILE) asm volatile("st.volatile.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (ST == ST_RELEASE) asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
  if (ST == ATOM_EXCH) { uint32_t o; asm volatile("atom.relaxed.gpu.global.exch.b32 %0, [%1], %2;" : "=r"(o) : "l"(p), "r"(v) : "memory"); }
  if (ST == RED_ADD) asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
  if (ST == ST_RELAXED_V4) asm volatile("st.relaxed.gpu.global.v4.b32 [%0], {%1,%1,%1,%1};" ::"l"(p), "r"(v) : "memory");
}
template <int LD> __device__ __forceinline__ uint32_t get(const uint32_t* p) {
  uint32_t v;
  if (LD == LD_RELAXED) asm volatile("ld.relaxed.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if (LD == LD_CG) asm volatile("ld.global.cg.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if (LD == LD_VOLATILE) asm volatile("ld.volatile.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  if (LD == LD_ACQUIRE) asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// extra = number of additional 16-byte stores per thread (256 threads) issued by the sender right BEFORE the signalling store
// (models the stash / partial traffic that shares the SM's store path with the signal)
template <int ST, int LD>
__global__ void pingpong(uint32_t* a, uint32_t* b, uint4* junk, int iters, int peer, int extra, long long* out) {
  const bool ping = blockIdx.x == 0, pong = blockIdx.x == peer;
  if (!ping && !pong) return;
  uint4* myjunk = junk + (size_t)blockIdx.x * 256 * 64;
  long long t0 = 0;
  for (int i = 1; i <= iters; ++i) {
    if (i == 2 && threadIdx.x == 0) t0 = clock64();
    if (ping) {
      for (int k = 0; k < extra; ++k) myjunk[(size_t)k * 256 + threadIdx.x] = make_uint4(i, i, i, i);
      if (threadIdx.x == 0) {
        put<ST>(a, i);
        while (get<LD>(b) != (uint32_t)i) {}
      }
    } else {
      if (threadIdx.x == 0) while (get<LD>(a) != (uint32_t)i) {}
      __syncthreads();
      for (int k = 0; k < extra; ++k) myjunk[(size_t)k * 256 + threadIdx.x] = make_uint4(i, i, i, i);
      if (threadIdx.x == 0) put<ST>(b, i);
    }
    __syncthreads();
  }
  if (ping && threadIdx.x == 0) out[0] = (clock64() - t0) / (iters - 1);
}

template <int ST, int LD>
static void run(const char* name, int peer, int extra) {
  uint32_t *a, *b; uint4* junk; long long* out;
  cudaMalloc(&a, 256); cudaMalloc(&b, 256); cudaMalloc(&junk, (size_t)148 * 256 * 64 * 16); cudaMalloc(&out, 8);
  cudaMemset(a, 0, 256); cudaMemset(b, 0, 256);
  pingpong<ST, LD><<<148, 256>>>(a, b, junk, 2000, peer, extra, out);
  long long h = 0;
  cudaError_t e = cudaMemcpy(&h, out, 8, cudaMemcpyDeviceToHost);
  printf("%-34s peer CTA %3d  extra stores/thread %2d : %6lld cycles per round trip (%s)\n", name, peer, extra, h, cudaGetErrorString(e));
  cudaFree(a); cudaFree(b); cudaFree(junk); cudaFree(out);
}

int main() {
  for (int peer : {1, 2, 75, 147}) {
    run<ST_RELAXED, LD_RELAXED>("st.relaxed.gpu / ld.relaxed.gpu", peer, 0);
    run<ST_RELAXED, LD_CG>("st.relaxed.gpu / ld.cg", peer, 0);
  }
  run<ST_VOLATILE, LD_VOLATILE>("st.volatile / ld.volatile", 75, 0);
  run<ST_RELEASE, LD_ACQUIRE>("st.release.gpu / ld.acquire.gpu", 75, 0);
  run<ATOM_EXCH, LD_RELAXED>("atom.exch / ld.relaxed.gpu", 75, 0);
  run<ST_RELAXED_V4, LD_RELAXED>("st.relaxed.gpu.v4 / ld.relaxed", 75, 0);
  for (int extra : {1, 4, 8, 24}) {
    run<ST_RELAXED, LD_RELAXED>("st.relaxed / ld.relaxed + traffic", 75, extra);
    run<ATOM_EXCH, LD_RELAXED>("atom.exch / ld.relaxed + traffic", 75, extra);
  }
  return 0;
}
