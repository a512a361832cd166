// Host-side construction of TMA tensor maps (bf16, SWIZZLE_128B) without linking libcuda:
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint at first use.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

namespace b2t {

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PFN_encodeTiled get_encode_tiled() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
      fprintf(stderr, "b2t: cuTensorMapEncodeTiled unavailable (%d)\n", (int)e);
      return nullptr;
    }
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

// 4-D bf16 tensor map.  dims[0] is the contiguous dimension (elements); strides_elems[i] is the
// stride of dims[i+1] in elements (must be a multiple of 8 = 16 bytes).  box[i] <= 256, box[0] == 64
// (128 bytes, the SWIZZLE_128B span).  Out-of-bounds elements read as zero.
inline int make_tmap_bf16_4d(CUtensorMap* out, const void* ptr, const uint64_t dims[4], const uint64_t strides_elems[3],
                             const uint32_t box[4]) {
  PFN_encodeTiled fn = get_encode_tiled();
  if (!fn) return -1;
  cuuint64_t gdim[4] = {dims[0], dims[1], dims[2], dims[3]};
  cuuint64_t gstr[3] = {strides_elems[0] * 2, strides_elems[1] * 2, strides_elems[2] * 2};
  cuuint32_t bx[4] = {box[0], box[1], box[2], box[3]};
  cuuint32_t es[4] = {1, 1, 1, 1};
  for (int i = 0; i < 3; ++i) {
    if (gstr[i] % 16 != 0) {
      fprintf(stderr, "b2t: tensor-map stride %d (%llu B) not a multiple of 16\n", i, (unsigned long long)gstr[i]);
      return -2;
    }
  }
  if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
    fprintf(stderr, "b2t: tensor-map base not 16-byte aligned\n");
    return -3;
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), gdim, gstr, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "b2t: cuTensorMapEncodeTiled failed (%d) dims=%llu,%llu,%llu,%llu box=%u,%u,%u,%u\n", (int)r,
            (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)dims[2],
            (unsigned long long)dims[3], box[0], box[1], box[2], box[3]);
    return -4;
  }
  return 0;
}

}  // namespace b2t
