// Tensor maps (TMA descriptors) of the image windows the tile kernel stages (kernels_tile.cuh).  The encoder is a driver
// function; the library links the CUDA runtime statically and nothing of the driver, so it is looked up at run time.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string.h>

#include "host_common.cuh"

namespace pytvb {

typedef CUresult (*TmapEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                      const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline TmapEncodeTiledFn tmap_encoder() {
    static TmapEncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<TmapEncodeTiledFn>(p);
    }();
    return fn;
}

template <typename T> struct TmapType;
template <> struct TmapType<float> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT32; };
template <> struct TmapType<double> { static constexpr CUtensorMapDataType v = CU_TENSOR_MAP_DATA_TYPE_FLOAT64; };

// Map of `planes` planes of (M, Ni, Nj) starting at `base`, box = (1, boxM, boxI, boxJ) elements; cells outside the tensor
// are filled with zeros.  `base` null: an all-zero map (never dereferenced by the kernel).
template <typename T>
inline int make_image_tmap(CUtensorMap* out, const T* base, long long planes, int M, int Ni, int Nj, int boxM, int boxI, int boxJ) {
    memset(out, 0, sizeof(*out));
    if (!base) return PYTVB_OK;
    TmapEncodeTiledFn enc = tmap_encoder();
    PYTVB_REQUIRE(enc != nullptr, "the CUDA driver does not export cuTensorMapEncodeTiled");
    const cuuint64_t dims[4] = {(cuuint64_t)Nj, (cuuint64_t)Ni, (cuuint64_t)M, (cuuint64_t)planes};
    const cuuint64_t strides[3] = {(cuuint64_t)Nj * sizeof(T), (cuuint64_t)Ni * Nj * sizeof(T), (cuuint64_t)M * Ni * Nj * sizeof(T)};
    const cuuint32_t box[4] = {(cuuint32_t)boxJ, (cuuint32_t)boxI, (cuuint32_t)boxM, 1u};
    const cuuint32_t estr[4] = {1u, 1u, 1u, 1u};
    const CUresult r = enc(out, TmapType<T>::v, 4, const_cast<T*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    PYTVB_REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) for a (%lld, %d, %d, %d) image, box (%d, %d, %d)", (int)r, planes, M, Ni, Nj, boxM, boxI,
                  boxJ);
    return PYTVB_OK;
}

}  // namespace pytvb
