// tma_host.cu -- host-side creation of TMA tensor maps.  The driver entry point is resolved at run time
// through the CUDA runtime so the library does not link libcuda (it must load on a box without a driver
// for the CPU-side symbol tests).
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>

#include "engine_internal.h"

namespace wrnn {

// 2-D row-major fp16 tensor [rows][cols]; box = box_rows x box_cols elements; 128-byte swizzle
// (box_cols * 2 bytes must be 128).  Out-of-bounds rows are zero-filled by the hardware.
cudaError_t make_tmap_f16_2d(void* tmap_out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows,
                             uint32_t box_cols) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    cuuint64_t dims[2] = {cols, rows};
    cuuint64_t strides[1] = {cols * 2};
    cuuint32_t box[2] = {box_cols, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = encode(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}


// The same matrix seen as [cols/64 k-blocks][rows][64]: one box = box_kb k-blocks of box_rows x 64 elements, delivered as
// box_kb consecutive SWIZZLE_128B tiles (a whole group of k-blocks per TMA operation: an operation has a fixed cost of
// ~0.3 us whatever its size).
cudaError_t make_tmap_f16_kblocks(void* tmap_out, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows, uint32_t box_kb) {
    static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    }
    cuuint64_t dims[3] = {64, rows, cols / 64};
    cuuint64_t strides[2] = {cols * 2, 128};
    cuuint32_t box[3] = {64, box_rows, box_kb};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = encode(reinterpret_cast<CUtensorMap*>(tmap_out), CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base),
                        dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace wrnn
