// tcgen05 relative-position attention (placeholder until the kernels land: reports "not eligible").
#include "common.cuh"

int tgan_relattn_fwd_tc(const void*, int64_t, const void*, const void*, int64_t, const void*, int64_t, const float*,
                        const float*, const uint8_t*, void*, int64_t, float*, int, int, int, int, int, int, float,
                        float, uint64_t, uint64_t, cudaStream_t) {
    tgan_set_error("tgan_relattn_fwd: tcgen05 kernel not available for this shape");
    return -1;
}
int tgan_relattn_bwd_tc(const void*, int64_t, const void*, const void*, int64_t, const void*, int64_t, const float*,
                        const float*, const uint8_t*, const void*, const void*, int64_t, const float*, float*, void*,
                        void*, void*, int64_t, float*, int64_t, float*, float*, int, int, int, int, int, int, float,
                        float, uint64_t, uint64_t, cudaStream_t) {
    tgan_set_error("tgan_relattn_bwd: tcgen05 kernel not available for this shape");
    return -1;
}
