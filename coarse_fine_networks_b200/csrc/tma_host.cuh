// Host side of the TMA (tensor-map) path of the tcgen05 GEMM kernels: building CUtensorMap descriptors without linking
// libcuda (the driver entry point is resolved at run time through the runtime API), and the device-side issue wrappers.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

typedef CUresult (*cf_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                       const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline cf_encode_tiled_fn cf_get_encode_tiled() {
    static cf_encode_tiled_fn fn = nullptr;
    static int tried = 0;
    if (!tried) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (cf_encode_tiled_fn)p;
        tried = 1;
    }
    return fn;
}

// Row tensor [B][R][K] fp32 (channels-last activations: a "row" is one (t,h,w) position).  TMA needs every global stride to
// be a multiple of 16 bytes, which K = 54 (216-byte rows) is not: `fold` consecutive rows are then presented as one row of
// fold*K floats (fold = 2 for even K, 4 otherwise; needs R % fold == 0), which is the same memory.
//   fold == 1 : dims (K, R, B),          box (32, rows, 1)      -> one [rows][32 floats] k-chunk per copy (pitch 128 B)
//   fold  > 1 : dims (fold*K, R/fold, B), box (fold*K, rows/fold, 1) -> the whole [rows][K] tile per copy (pitch K*4 B)
// (rows = box_rows: 128 for the forward / data-gradient GEMM tiles, the row-block size of the weight-gradient kernel)
// Coordinates past an extent (rows beyond the sample, channels beyond K) arrive as zeros and still count towards the
// transaction bytes.  Returns false when the shape cannot be described (caller keeps the register-load producers).
static inline bool cf_make_row_tmap(CUtensorMap* tm, const float* base, int B, long long R, int K, int fold, int box_rows = 128) {
    cf_encode_tiled_fn enc = cf_get_encode_tiled();
    if (!enc || !base || (((uintptr_t)base) & 15)) return false;
    if (fold < 1 || R % fold != 0) return false;
    const unsigned long long kf = (unsigned long long)K * fold, rf = (unsigned long long)(R / fold);
    if ((kf * 4) % 16 != 0 || ((unsigned long long)R * K * 4) % 16 != 0) return false;
    if (fold > 1 && kf > 256) return false;                       // box dimensions are limited to 256 elements
    cuuint64_t gdim[3] = {kf, rf, (cuuint64_t)B};
    cuuint64_t gstr[2] = {kf * 4, (cuuint64_t)R * K * 4};
    if (box_rows % fold != 0) return false;
    cuuint32_t box[3] = {fold == 1 ? 32u : (cuuint32_t)kf, (cuuint32_t)(box_rows / fold), 1u};
    cuuint32_t est[3] = {1u, 1u, 1u};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void*)base, gdim, gstr, box, est, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

#ifdef __CUDACC__
// cp.async.bulk.tensor (TMA tile load) of one 3-D box into shared memory, completion on an mbarrier (transaction bytes)
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* tm, int c0, int c1, int c2, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
                     (uint32_t)__cvta_generic_to_shared(dst)),
                 "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"((uint32_t)__cvta_generic_to_shared(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(tm) : "memory");
}
#endif
