// SIMT (FFMA) GEMM with fused epilogues.  This is the fp32-mode contraction (1e-4 parity with the reference's
// fp32 path) and the fallback for shapes the tcgen05 kernel does not take (tiny decode GEMMs, odd leading
// dimensions).  128x128x16 block tile, 256 threads, 8x8 register tile, fp32 accumulation.
#include "common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4, NT = 256;

template <typename TA, typename TC, bool TRANSA, bool TRANSB>
__global__ void __launch_bounds__(NT) gemm_simt_kernel(int M, int N, int K, const TA* __restrict__ A, int64_t lda,
                                                       const TA* __restrict__ B, int64_t ldb, TC* __restrict__ C,
                                                       int64_t ldc, EpiParams ep) {
    __shared__ float As[BK][BM + PAD];
    __shared__ float Bs[BK][BN + PAD];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    const int ty = tid / 16, tx = tid % 16;

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[8], rb[8];
    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            int e = tid + NT * t;
            int m, k;
            if (TRANSA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < M && gk < K) v = to_f(TRANSA ? A[(int64_t)gk * lda + gm] : A[(int64_t)gm * lda + gk]);
            ra[t] = v;
            int n;
            if (TRANSB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
            int gn = n0 + n;
            gk = k0 + k;
            v = 0.f;
            if (gn < N && gk < K) v = to_f(TRANSB ? B[(int64_t)gn * ldb + gk] : B[(int64_t)gk * ldb + gn]);
            rb[t] = v;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            int e = tid + NT * t;
            int m, k, n;
            if (TRANSA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
            As[k][m] = ra[t];
            if (TRANSB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
            Bs[k][n] = rb[t];
        }
    };

    load_tiles(0);
    for (int k0 = 0; k0 < K; k0 += BK) {
        store_tiles();
        __syncthreads();
        if (k0 + BK < K) load_tiles(k0 + BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[8], b[8];
            *(float4*)&a[0] = *(const float4*)&As[kk][ty * 4];
            *(float4*)&a[4] = *(const float4*)&As[kk][64 + ty * 4];
            *(float4*)&b[0] = *(const float4*)&Bs[kk][tx * 4];
            *(float4*)&b[4] = *(const float4*)&Bs[kk][64 + tx * 4];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }

#pragma unroll
    for (int i = 0; i < 8; ++i) {
        int gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int gn = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
            if (gn >= N) continue;
            float v = acc[i][j] * ep.alpha;
            const int flags = ep.flags;
            if (flags & TGAN_EPI_BIAS) v += ep.bias[gn];
            if (flags & TGAN_EPI_RELU) v = fmaxf(v, 0.f);
            if (flags & (TGAN_EPI_MASK_POS | TGAN_EPI_ADD_AUX)) {
                float a = ep.aux_is_f32 ? ((const float*)ep.aux)[(int64_t)gm * ep.ldaux + gn]
                                        : to_f(((const TA*)ep.aux)[(int64_t)gm * ep.ldaux + gn]);
                if (flags & TGAN_EPI_MASK_POS) v = a > 0.f ? v : 0.f;
                if (flags & TGAN_EPI_DROPOUT)
                    v = dropout_keep_k(step_fold(ep.drop_key), (uint64_t)gm * ldc + gn, ep.drop_thresh) ? v * ep.drop_scale : 0.f;
                if (flags & TGAN_EPI_ADD_AUX) v += a;
            } else if (flags & TGAN_EPI_DROPOUT) {
                v = dropout_keep_k(step_fold(ep.drop_key), (uint64_t)gm * ldc + gn, ep.drop_thresh) ? v * ep.drop_scale : 0.f;
            }
            TC* cp = C + (int64_t)gm * ldc + gn;
            if constexpr (sizeof(TC) == 4) {
                // fp32 accumulation is a reduction: weight-gradient GEMMs of concurrent streams may target the same rows
                if (flags & TGAN_EPI_ACCUM) atomicAdd(reinterpret_cast<float*>(cp), v);
                else *cp = from_f<TC>(v);
            } else {
                if (flags & TGAN_EPI_ACCUM) v += to_f(*cp);
                *cp = from_f<TC>(v);
            }
        }
    }
}

template <typename TA, typename TC>
int launch(int transA, int transB, int M, int N, int K, const void* A, int64_t lda, const void* B, int64_t ldb,
           void* C, int64_t ldc, const EpiParams& ep, cudaStream_t st) {
    dim3 grid(ceil_div(N, BN), ceil_div(M, BM));
#define TGAN_GEMM_CASE(TA_, TB_)                                                                              \
    gemm_simt_kernel<TA, TC, TA_, TB_><<<grid, NT, 0, st>>>(M, N, K, (const TA*)A, lda, (const TA*)B, ldb, \
                                                             (TC*)C, ldc, ep)
    if (!transA && !transB) TGAN_GEMM_CASE(false, false);
    else if (!transA && transB) TGAN_GEMM_CASE(false, true);
    else if (transA && !transB) TGAN_GEMM_CASE(true, false);
    else TGAN_GEMM_CASE(true, true);
#undef TGAN_GEMM_CASE
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

}  // namespace

int tgan_gemm_simt(int dtype_ab, int dtype_c, int transA, int transB, int M, int N, int K, const void* A,
                   int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias, const void* aux,
                   int64_t ldaux, int flags, float alpha, float drop_p, uint64_t seed, uint64_t site,
                   cudaStream_t st) {
    if (M <= 0 || N <= 0) return 0;
    EpiParams ep;
    ep.bias = bias; ep.aux = aux; ep.ldaux = ldaux; ep.flags = flags; ep.alpha = alpha;
    ep.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    ep.drop_thresh = dropout_thresh(drop_p);
    ep.drop_key = dropout_key(seed, site);
    ep.aux_is_f32 = (flags & TGAN_EPI_AUX_F32) || dtype_ab == TGAN_F32;
    if (drop_p <= 0.f) ep.flags &= ~TGAN_EPI_DROPOUT;
    if (dtype_ab == TGAN_F32) {
        TGAN_CHECK_ARG(dtype_c == TGAN_F32, "tgan_gemm: fp32 operands need an fp32 output");
        return launch<float, float>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, ep, st);
    }
    if (dtype_c == TGAN_F32) return launch<bf16, float>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, ep, st);
    return launch<bf16, bf16>(transA, transB, M, N, K, A, lda, B, ldb, C, ldc, ep, st);
}

// ---- public dispatcher ------------------------------------------------------------------------------------
extern "C" int tgan_gemm(int dtype_ab, int dtype_c, int transA, int transB, int M, int N, int K, const void* A,
                         int64_t lda, const void* B, int64_t ldb, void* C, int64_t ldc, const float* bias,
                         const void* aux, int64_t ldaux, int epi_flags, float alpha, float drop_p, uint64_t seed,
                         uint64_t site, int impl, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    TGAN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "tgan_gemm: negative dimension");
    TGAN_CHECK_ARG(!(epi_flags & TGAN_EPI_BIAS) || bias, "tgan_gemm: BIAS flag without bias pointer");
    TGAN_CHECK_ARG(!(epi_flags & (TGAN_EPI_MASK_POS | TGAN_EPI_ADD_AUX)) || aux, "tgan_gemm: aux flag without aux pointer");
    if (impl != TGAN_IMPL_SIMT && dtype_ab == TGAN_BF16) {
        int rc = tgan_gemm_tc(dtype_c, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux, epi_flags,
                              alpha, drop_p, seed, site, impl == TGAN_IMPL_TC, st);
        if (rc >= 0) return rc;
        if (impl == TGAN_IMPL_TC) return 3;  // message set by tgan_gemm_tc
    } else if (impl == TGAN_IMPL_TC) {
        tgan_set_error("tgan_gemm: tcgen05 path needs bf16 operands");
        return 3;
    }
    return tgan_gemm_simt(dtype_ab, dtype_c, transA, transB, M, N, K, A, lda, B, ldb, C, ldc, bias, aux, ldaux,
                          epi_flags, alpha, drop_p, seed, site, st);
}

int tgan_set_step_ctr_gemm_simt(const void* p) { return tgan_set_step_ctr_local(p); }
