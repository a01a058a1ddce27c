// BERT discriminator attention tiles on the tensor cores (bf16 operands, fp32 accumulation): value, input gradient
// and forward tangent.  Reference arithmetic: HuggingFace modeling_bert.py BertSelfAttention (the library the reference
// calls at transformer_gan.py:403-416); same contracts as the SIMT kernels in bert_ops.cu, which stay for the fp32
// parity mode and for head sizes that are not multiples of 16.
//
// One CTA = one (sequence, head): T <= 64 tokens x d_head <= 64.  The SIMT version kept fp32 tiles in shared memory and
// was bound by its shared-memory loads (8 LDS per 16 FMAs: 280 us per call at 512 sequences x 12 heads).  A 64-token
// tile is too small for a tcgen05 pipeline (one M = 64 MMA group per product, TMEM round trips in between), so these
// kernels use warp-level mma.sync.m16n8k16 with the operands staged once as bf16 ([64][72] tiles, ldmatrix):
//   warp w owns query rows 16w .. 16w+15: S = Q K^T, softmax / dropout on the accumulator fragments, and the P V
//   product straight from those fragments (the C layout of S is the A layout of the next product);
//   the key-side gradients (dK = dS^T Q, dV = P~^T dO) contract over ALL query rows: dS / P~ go through shared
//   memory once and each warp then owns 16 key rows (transposed ldmatrix).
// Dropout masks: the same stateless hash as the SIMT kernels (forward, dgrad and JVP regenerate identical masks).
#include "common.cuh"

namespace {
constexpr int MT = 64;    // tokens per tile
constexpr int MLD = 72;   // bf16 row pitch of the shared tiles: 144 bytes, ldmatrix rows land in distinct bank groups
constexpr int MTHREADS = 128;

struct MmaAttnArgs {
    int B, heads, T, dh, H;
    float scale, drop_scale;
    uint32_t thresh, key;
};

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float x, float y) {
    __nv_bfloat162 v = __floats2bfloat162_rn(x, y);
    return *reinterpret_cast<uint32_t*>(&v);
}

// global [T rows, dh cols] (row pitch ldg) -> shared [64][MLD] bf16, zero beyond (T, dh)
__device__ __forceinline__ void load_tile_bf16(bf16* s, const bf16* g, int64_t ldg, int T, int dh) {
    for (int idx = threadIdx.x; idx < MT * 8; idx += MTHREADS) {
        const int r = idx >> 3, c = (idx & 7) * 8;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (r < T && c < dh) v = *reinterpret_cast<const uint4*>(g + (int64_t)r * ldg + c);
        *reinterpret_cast<uint4*>(s + r * MLD + c) = v;
    }
}

// acc[nt] (16 rows of this warp x n-tile nt of 8) += A[rows r0.., k] * Bnk[n][k]^T   (both row-major [.][k], k < kdim)
__device__ __forceinline__ void mm_rows_nt(float (&acc)[8][4], const bf16* A, int r0, const bf16* Bnk, int kdim, int lane) {
    const uint32_t a_base = smem_addr(A + (r0 + (lane & 15)) * MLD + ((lane >> 4) << 3));
    const uint32_t b_base = smem_addr(Bnk + ((lane & 7) + ((lane >> 4) << 3)) * MLD + (((lane >> 3) & 1) << 3));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        if (16 * kk >= kdim) break;
        uint32_t a[4];
        ldsm_x4(a, a_base + 32 * kk);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            uint32_t b[4];
            ldsm_x4(b, b_base + (16 * np * MLD + 16 * kk) * 2);
            mma16816(acc[2 * np], a, b[0], b[1]);
            mma16816(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}
// out[nt] += Afrag(k-step kk) * Bkn[k][n]   (B row-major [k][n], n < ndim; A given as 4 k-steps of fragments)
__device__ __forceinline__ void mm_frag_kn(float (&out)[8][4], const uint32_t (&af)[4][4], const bf16* Bkn, int ndim, int lane) {
    const uint32_t b_base = smem_addr(Bkn + ((lane & 7) + (((lane >> 3) & 1) << 3)) * MLD + ((lane >> 4) << 3));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            if (16 * np >= ndim) break;
            uint32_t b[4];
            ldsm_x4_t(b, b_base + (16 * kk * MLD + 16 * np) * 2);
            mma16816(out[2 * np], af[kk], b[0], b[1]);
            mma16816(out[2 * np + 1], af[kk], b[2], b[3]);
        }
    }
}
// out[nt] += At[k][rows r0..]^T * Bkn[k][n]   (A stored TRANSPOSED: row-major [k][m]; k < 64, n < ndim)
__device__ __forceinline__ void mm_tn(float (&out)[8][4], const bf16* At, int r0, const bf16* Bkn, int ndim, int lane) {
    const uint32_t a_base = smem_addr(At + ((lane & 7) + ((lane >> 4) << 3)) * MLD + r0 + (((lane >> 3) & 1) << 3));
    const uint32_t b_base = smem_addr(Bkn + ((lane & 7) + (((lane >> 3) & 1) << 3)) * MLD + ((lane >> 4) << 3));
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        uint32_t a[4];
        ldsm_x4_t(a, a_base + 16 * kk * MLD * 2);
#pragma unroll
        for (int np = 0; np < 4; ++np) {
            if (16 * np >= ndim) break;
            uint32_t b[4];
            ldsm_x4_t(b, b_base + (16 * kk * MLD + 16 * np) * 2);
            mma16816(out[2 * np], a, b[0], b[1]);
            mma16816(out[2 * np + 1], a, b[2], b[3]);
        }
    }
}
__device__ __forceinline__ void zero8(float (&a)[8][4]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[i][j] = 0.f;
}
// accumulator fragments (16 rows x 64) -> A fragments of the next product (k = the 64 columns)
__device__ __forceinline__ void to_afrag(uint32_t (&af)[4][4], const float (&c)[8][4]) {
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
        af[kk][0] = pack2(c[2 * kk][0], c[2 * kk][1]);
        af[kk][1] = pack2(c[2 * kk][2], c[2 * kk][3]);
        af[kk][2] = pack2(c[2 * kk + 1][0], c[2 * kk + 1][1]);
        af[kk][3] = pack2(c[2 * kk + 1][2], c[2 * kk + 1][3]);
    }
}
// accumulator fragments -> global rows (row r0 + g / + 8, columns 8 nt + 2 t), scaled
__device__ __forceinline__ void store_rows(bf16* g, int64_t ldg, const float (&c)[8][4], int r0, int T, int dh, float mul,
                                           int lane) {
    const int gq = lane >> 2, t2 = (lane & 3) * 2;
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        if (8 * nt >= dh) break;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int r = r0 + gq + 8 * h;
            if (r < T)
                *reinterpret_cast<uint32_t*>(g + (int64_t)r * ldg + 8 * nt + t2) = pack2(c[nt][2 * h] * mul, c[nt][2 * h + 1] * mul);
        }
    }
}
__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}
__device__ __forceinline__ bool keep_elem(const MmaAttnArgs& a, uint32_t key, int bh, int i, int j) {
    if (!a.thresh) return true;
    return dropout_keep_k(key, ((uint64_t)bh * a.T + i) * a.T + j, a.thresh);
}

// ---------------------------------------------------------------------------------------------------------------
// forward: ctx = drop(softmax(Q K^T * scale)) V ;  lse saved
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MTHREADS)
bert_attn_fwd_mma(const bf16* __restrict__ qkv, int64_t ldq, bf16* __restrict__ ctx, int64_t ldc, float* __restrict__ lse,
                  MmaAttnArgs a) {
    __shared__ __align__(16) bf16 sQ[MT * MLD], sK[MT * MLD], sV[MT * MLD];
    const int bh = blockIdx.x, b = bh / a.heads, h = bh % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t2 = (lane & 3) * 2;
    const bf16* base = qkv + (int64_t)b * a.T * ldq + h * a.dh;
    load_tile_bf16(sQ, base, ldq, a.T, a.dh);
    load_tile_bf16(sK, base + a.H, ldq, a.T, a.dh);
    load_tile_bf16(sV, base + 2 * a.H, ldq, a.T, a.dh);
    __syncthreads();
    const int r0 = 16 * warp;
    float s[8][4];
    zero8(s);
    mm_rows_nt(s, sQ, r0, sK, a.dh, lane);
    const uint32_t key = step_fold(a.key);
    float m[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 8 * nt + t2 + (e & 1);
            s[nt][e] = j < a.T ? s[nt][e] * a.scale : -INFINITY;
            m[e >> 1] = fmaxf(m[e >> 1], s[nt][e]);
        }
    m[0] = quad_max(m[0]); m[1] = quad_max(m[1]);
    float l[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float p = __expf(s[nt][e] - m[e >> 1]);
            s[nt][e] = p;
            l[e >> 1] += p;
        }
    l[0] = quad_sum(l[0]); l[1] = quad_sum(l[1]);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int i = r0 + gq + 8 * hh;
        if ((lane & 3) == 0 && i < a.T) lse[(int64_t)bh * a.T + i] = m[hh] + __logf(l[hh]);
    }
    const float inv[2] = {a.drop_scale / l[0], a.drop_scale / l[1]};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = r0 + gq + 8 * (e >> 1), j = 8 * nt + t2 + (e & 1);
            s[nt][e] = keep_elem(a, key, bh, i, j) ? s[nt][e] * inv[e >> 1] : 0.f;
        }
    uint32_t pf[4][4];
    to_afrag(pf, s);
    float o[8][4];
    zero8(o);
    mm_frag_kn(o, pf, sV, a.dh, lane);
    store_rows(ctx + (int64_t)b * a.T * ldc + h * a.dh, ldc, o, r0, a.T, a.dh, 1.f, lane);
}

// P (un-dropped probabilities) of this warp's 16 rows from the saved lse: S -> exp(S * scale - lse); 0 outside (T, T)
__device__ __forceinline__ void probs_rows(float (&s)[8][4], const float* __restrict__ lse_bh, int r0, const MmaAttnArgs& a,
                                           int lane) {
    const int gq = lane >> 2, t2 = (lane & 3) * 2;
    float L[2];
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        const int i = r0 + gq + 8 * hh;
        L[hh] = i < a.T ? lse_bh[i] : INFINITY;
    }
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 8 * nt + t2 + (e & 1);
            s[nt][e] = j < a.T ? __expf(s[nt][e] * a.scale - L[e >> 1]) : 0.f;
        }
}

// ---------------------------------------------------------------------------------------------------------------
// dgrad: dqkv from dctx.   dP~ = dO V^T ;  dS = P (drop'(dP~) - delta) ;  dQ = dS K s ;  dK = dS^T Q s ;  dV = P~^T dO
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MTHREADS)
bert_attn_bwd_mma(const bf16* __restrict__ qkv, int64_t ldq, const bf16* __restrict__ dctx, int64_t ldc,
                  const float* __restrict__ lse, bf16* __restrict__ dqkv, int64_t lddq, MmaAttnArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    bf16* sQ = reinterpret_cast<bf16*>(smraw);
    bf16 *sK = sQ + MT * MLD, *sV = sK + MT * MLD, *sG = sV + MT * MLD, *sDS = sG + MT * MLD, *sPT = sDS + MT * MLD;
    const int bh = blockIdx.x, b = bh / a.heads, h = bh % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t2 = (lane & 3) * 2;
    const bf16* base = qkv + (int64_t)b * a.T * ldq + h * a.dh;
    load_tile_bf16(sQ, base, ldq, a.T, a.dh);
    load_tile_bf16(sK, base + a.H, ldq, a.T, a.dh);
    load_tile_bf16(sV, base + 2 * a.H, ldq, a.T, a.dh);
    load_tile_bf16(sG, dctx + (int64_t)b * a.T * ldc + h * a.dh, ldc, a.T, a.dh);
    __syncthreads();
    const int r0 = 16 * warp;
    const uint32_t key = step_fold(a.key);
    float p[8][4], d[8][4];
    zero8(p);
    mm_rows_nt(p, sQ, r0, sK, a.dh, lane);
    probs_rows(p, lse + (int64_t)bh * a.T, r0, a, lane);
    zero8(d);
    mm_rows_nt(d, sG, r0, sV, a.dh, lane);   // dP~[i][j] = dO_i . V_j
    float delta[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = r0 + gq + 8 * (e >> 1), j = 8 * nt + t2 + (e & 1);
            const bool keep = keep_elem(a, key, bh, i, j);
            const float dpk = keep ? d[nt][e] * a.drop_scale : 0.f;   // gradient w.r.t. the un-dropped probability
            delta[e >> 1] = fmaf(p[nt][e], dpk, delta[e >> 1]);
            d[nt][e] = dpk;
            // p keeps P; the dropped weights P~ are formed below
            if (!keep) d[nt][e] = 0.f;
        }
    delta[0] = quad_sum(delta[0]); delta[1] = quad_sum(delta[1]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = r0 + gq + 8 * (e >> 1), j = 8 * nt + t2 + (e & 1);
            const float pj = p[nt][e];
            d[nt][e] = pj * (d[nt][e] - delta[e >> 1]);                        // dS
            p[nt][e] = keep_elem(a, key, bh, i, j) ? pj * a.drop_scale : 0.f;  // P~
        }
    // dS / P~ -> shared (bf16) for the key-side products; dQ straight from the fragments
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            const int i = r0 + gq + 8 * hh, j = 8 * nt + t2;
            *reinterpret_cast<uint32_t*>(sDS + i * MLD + j) = pack2(d[nt][2 * hh], d[nt][2 * hh + 1]);
            *reinterpret_cast<uint32_t*>(sPT + i * MLD + j) = pack2(p[nt][2 * hh], p[nt][2 * hh + 1]);
        }
    bf16* obase = dqkv + (int64_t)b * a.T * lddq + h * a.dh;
    {
        uint32_t df[4][4];
        to_afrag(df, d);
        float o[8][4];
        zero8(o);
        mm_frag_kn(o, df, sK, a.dh, lane);                                    // dQ = dS K
        store_rows(obase, lddq, o, r0, a.T, a.dh, a.scale, lane);
    }
    __syncthreads();
    {   // this warp now owns KEY rows r0 .. r0 + 15
        float o[8][4];
        zero8(o);
        mm_tn(o, sDS, r0, sQ, a.dh, lane);                                    // dK = dS^T Q
        store_rows(obase + a.H, lddq, o, r0, a.T, a.dh, a.scale, lane);
        zero8(o);
        mm_tn(o, sPT, r0, sG, a.dh, lane);                                    // dV = P~^T dO
        store_rows(obase + 2 * a.H, lddq, o, r0, a.T, a.dh, 1.f, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------------
// JVP: ctx_dot from qkv_dot.  Sd = (Qd K^T + Q Kd^T) s ;  Pd = P (Sd - rowsum(P Sd)) ;  ctx_d = drop(Pd) V + drop(P) Vd
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(MTHREADS)
bert_attn_jvp_mma(const bf16* __restrict__ qkv, int64_t ldq, const bf16* __restrict__ qkvd, int64_t ldqd,
                  const float* __restrict__ lse, bf16* __restrict__ ctxd, int64_t ldc, MmaAttnArgs a) {
    extern __shared__ __align__(16) uint8_t smraw[];
    bf16* sQ = reinterpret_cast<bf16*>(smraw);
    bf16 *sK = sQ + MT * MLD, *sV = sK + MT * MLD, *sQd = sV + MT * MLD, *sKd = sQd + MT * MLD, *sVd = sKd + MT * MLD;
    const int bh = blockIdx.x, b = bh / a.heads, h = bh % a.heads;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, gq = lane >> 2, t2 = (lane & 3) * 2;
    const bf16* base = qkv + (int64_t)b * a.T * ldq + h * a.dh;
    const bf16* based = qkvd + (int64_t)b * a.T * ldqd + h * a.dh;
    load_tile_bf16(sQ, base, ldq, a.T, a.dh);
    load_tile_bf16(sK, base + a.H, ldq, a.T, a.dh);
    load_tile_bf16(sV, base + 2 * a.H, ldq, a.T, a.dh);
    load_tile_bf16(sQd, based, ldqd, a.T, a.dh);
    load_tile_bf16(sKd, based + a.H, ldqd, a.T, a.dh);
    load_tile_bf16(sVd, based + 2 * a.H, ldqd, a.T, a.dh);
    __syncthreads();
    const int r0 = 16 * warp;
    const uint32_t key = step_fold(a.key);
    float p[8][4], d[8][4];
    zero8(p);
    mm_rows_nt(p, sQ, r0, sK, a.dh, lane);
    probs_rows(p, lse + (int64_t)bh * a.T, r0, a, lane);
    zero8(d);
    mm_rows_nt(d, sQd, r0, sK, a.dh, lane);
    mm_rows_nt(d, sQ, r0, sKd, a.dh, lane);
    float dot[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            d[nt][e] *= a.scale;
            dot[e >> 1] = fmaf(p[nt][e], d[nt][e], dot[e >> 1]);
        }
    dot[0] = quad_sum(dot[0]); dot[1] = quad_sum(dot[1]);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int i = r0 + gq + 8 * (e >> 1), j = 8 * nt + t2 + (e & 1);
            const float mk = keep_elem(a, key, bh, i, j) ? a.drop_scale : 0.f;
            const float pj = p[nt][e];
            d[nt][e] = pj * (d[nt][e] - dot[e >> 1]) * mk;   // drop(Pd)
            p[nt][e] = pj * mk;                              // drop(P)
        }
    uint32_t df[4][4], pf[4][4];
    to_afrag(df, d);
    to_afrag(pf, p);
    float o[8][4];
    zero8(o);
    mm_frag_kn(o, df, sV, a.dh, lane);
    mm_frag_kn(o, pf, sVd, a.dh, lane);
    store_rows(ctxd + (int64_t)b * a.T * ldc + h * a.dh, ldc, o, r0, a.T, a.dh, 1.f, lane);
}

MmaAttnArgs make_margs(int B, int heads, int T, int dh, float drop_p, uint64_t seed, uint64_t site) {
    MmaAttnArgs a;
    a.B = B; a.heads = heads; a.T = T; a.dh = dh; a.H = heads * dh;
    a.scale = 1.f / sqrtf((float)dh);
    a.drop_scale = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    a.thresh = drop_p > 0.f ? dropout_thresh16(drop_p) : 0u;
    a.key = dropout_key(seed, site);
    return a;
}
constexpr int TILE_BYTES = MT * MLD * 2;
}  // namespace

int tgan_set_step_ctr_bert_mma(const void* p) { return tgan_set_step_ctr_local(p); }

// Eligibility (checked by the callers in bert_ops.cu): bf16, T <= 64, d_head a multiple of 16 <= 64, 16-byte aligned rows.
int tgan_bert_attn_fwd_mma(const void* qkv, int64_t ldq, void* ctx, int64_t ldc, float* lse, int B, int heads, int T, int dh,
                           float drop_p, uint64_t seed, uint64_t site, cudaStream_t st) {
    MmaAttnArgs a = make_margs(B, heads, T, dh, drop_p, seed, site);
    bert_attn_fwd_mma<<<B * heads, MTHREADS, 0, st>>>((const bf16*)qkv, ldq, (bf16*)ctx, ldc, lse, a);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

int tgan_bert_attn_bwd_mma(const void* qkv, int64_t ldq, const void* dctx, int64_t ldc, const float* lse, void* dqkv,
                           int64_t lddq, int B, int heads, int T, int dh, float drop_p, uint64_t seed, uint64_t site,
                           cudaStream_t st) {
    MmaAttnArgs a = make_margs(B, heads, T, dh, drop_p, seed, site);
    TGAN_CUDA_OK(cudaFuncSetAttribute(bert_attn_bwd_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * TILE_BYTES));
    bert_attn_bwd_mma<<<B * heads, MTHREADS, 6 * TILE_BYTES, st>>>((const bf16*)qkv, ldq, (const bf16*)dctx, ldc, lse,
                                                                   (bf16*)dqkv, lddq, a);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

int tgan_bert_attn_jvp_mma(const void* qkv, int64_t ldq, const void* qkvd, int64_t ldqd, const float* lse, void* ctxd,
                           int64_t ldc, int B, int heads, int T, int dh, float drop_p, uint64_t seed, uint64_t site,
                           cudaStream_t st) {
    MmaAttnArgs a = make_margs(B, heads, T, dh, drop_p, seed, site);
    TGAN_CUDA_OK(cudaFuncSetAttribute(bert_attn_jvp_mma, cudaFuncAttributeMaxDynamicSharedMemorySize, 6 * TILE_BYTES));
    bert_attn_jvp_mma<<<B * heads, MTHREADS, 6 * TILE_BYTES, st>>>((const bf16*)qkv, ldq, (const bf16*)qkvd, ldqd, lse,
                                                                   (bf16*)ctxd, ldc, a);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}
