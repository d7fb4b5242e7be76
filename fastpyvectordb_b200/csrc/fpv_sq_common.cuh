// Shared by the SIMT scalar-quantizer scans (fpv_sq.cu) and the finish kernels of the tensor-core path
// (fpv_sq_mma.cu): the per-element arithmetic must be the SAME code so that both paths return the same bits.
#pragma once
#include "fpv_common.cuh"

namespace fpv {

__device__ __forceinline__ float u8f(uint32_t w, int b) {   // 2^23 + byte b of w, as float (exact)
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + b));
}

template <int KIND>
__device__ __forceinline__ void sq_word(uint32_t w, const float4& a, const float4& b, const float4& c, float& acc, float& nrm) {
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float f = u8f(w, i);
        if (KIND == FPV_SQ_L2) {
            float t = (av[i] - f) * bv[i];
            acc = fmaf(t, t, acc);
        } else {
            float dec = fmaf(f - 8388608.0f, av[i], bv[i]);
            acc = fmaf(dec, cv[i], acc);
            if (KIND == FPV_SQ_COSINE) nrm = fmaf(dec, dec, nrm);
        }
    }
}

// consts [q][3][Dp] for `kind` (fpv_sq.cu: sq_prep_kernel); Dp = D rounded up to 16
int sq_prep_launch(int kind, const uint8_t* qcodes, int64_t q, int D, int Dp, const float* mn, const float* sc, float* consts,
                   cudaStream_t st);

}  // namespace fpv
