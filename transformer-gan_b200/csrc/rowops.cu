// Bandwidth-bound row-wise kernels of the Transformer-XL hot path: embedding, positional embedding, LayerNorm,
// dropout, cross-entropy, Gumbel-softmax straight-through, bias-gradient column sums, dtype/pad converts,
// parameter packing and the fused clip+Adam update.  All are HBM-bound: one warp per row, 16-byte vector
// accesses, no shared-memory staging needed (no reuse).
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

// ------------------------------------------------------------------------------------------------------------
// error plumbing / bookkeeping
// ------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
unsigned long long g_tgan_launches = 0;

void tgan_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
extern "C" const char* tgan_last_error(void) { return g_err; }
extern "C" int tgan_version(void) { return 200; }  // 2xx: round-2 surface (split single-token backward, BERT ops, sampling, batching, LAMB, NCCL buckets)
extern "C" unsigned long long tgan_launch_count(void) { return g_tgan_launches; }
extern "C" int tgan_set_step_counter(const void* dev_u32) {
    int rc = tgan_set_step_ctr_local(dev_u32);
    rc |= tgan_set_step_ctr_gemm_simt(dev_u32) | tgan_set_step_ctr_gemm_tc(dev_u32) | tgan_set_step_ctr_relattn_simt(dev_u32) |
          tgan_set_step_ctr_relattn_decode(dev_u32) | tgan_set_step_ctr_bert(dev_u32) | tgan_set_step_ctr_sampling(dev_u32) |
          tgan_set_step_ctr_relattn_fwd_tc(dev_u32) | tgan_set_step_ctr_relattn_bwd_tc(dev_u32);
    if (rc) { tgan_set_error("tgan_set_step_counter: cudaMemcpyToSymbol failed"); return 2; }
    return 0;
}

#define DISPATCH_T(dtype, ...)                                  \
    do {                                                        \
        if ((dtype) == TGAN_F32) { typedef float T; __VA_ARGS__; } \
        else if ((dtype) == TGAN_BF16) { typedef bf16 T; __VA_ARGS__; } \
        else { tgan_set_error("bad dtype %d", (int)(dtype)); return 1; } \
    } while (0)

namespace {
constexpr int WPB = 8;  // warps per block for the warp-per-row kernels

// ------------------------------------------------------------------------------------------------------------
// embedding forward: out[row, c] = drop(E[ids[row], c] * scale)         mem_transformer.py:329-339, 557
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void embed_fwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ E, int64_t lde,
                                 T* __restrict__ out, int64_t ldo, int rows, int D, int DP, float scale,
                                 float drop_scale, uint32_t thresh, uint64_t seed, uint64_t site) {
    int row = blockIdx.x * WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const T* e = E + ids[row] * lde;
    T* o = out + (int64_t)row * ldo;
    for (int c0 = lane * 8; c0 < DP; c0 += 256) {
        float v[8];
        load8(e + c0, v);
        uint32_t keep = thresh ? dropout_keep8(seed, site, (uint64_t)row * ldo + c0, thresh) : 0xffu;
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = (c0 + t < D && ((keep >> t) & 1)) ? v[t] * scale * drop_scale : 0.f;
        store8(o + c0, v);
    }
}

// embedding backward.  dE[v, c] += scale * sum_{rows: ids[row] == v} dropmask(dout[row, c]).
// grid = (column blocks of 64, row slabs); each CTA accumulates its slab into a shared-memory table [V][64] with
// shared-memory reductions (bank = column: conflict-free within a warp) and flushes non-zero entries with one
// global reduction each.
template <typename T>
__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const T* __restrict__ dout, int64_t ldo,
                                 float* __restrict__ dE, int64_t ldde, int rows, int V, int D, float scale,
                                 float drop_scale, uint32_t thresh, uint64_t seed, uint64_t site, int rows_per_block) {
    extern __shared__ float tab[];  // [V][64]
    const int c0 = blockIdx.x * 64;
    for (int e = threadIdx.x; e < V * 64; e += blockDim.x) tab[e] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    const uint32_t key = step_fold(dropout_key(seed, site));
    const int c = c0 + 2 * lane;
    for (int row = r0 + warp; row < r1; row += nwarps) {
        const int v = (int)ids[row];
        float g0 = c < D ? to_f(dout[(int64_t)row * ldo + c]) : 0.f;
        float g1 = c + 1 < D ? to_f(dout[(int64_t)row * ldo + c + 1]) : 0.f;
        if (thresh) {
            g0 = dropout_keep_k(key, (uint64_t)row * ldo + c, thresh) ? g0 * drop_scale : 0.f;
            g1 = dropout_keep_k(key, (uint64_t)row * ldo + c + 1, thresh) ? g1 * drop_scale : 0.f;
        }
        atomicAdd(&tab[v * 64 + 2 * lane], g0);
        atomicAdd(&tab[v * 64 + 2 * lane + 1], g1);
    }
    __syncthreads();
    for (int e = threadIdx.x; e < V * 64; e += blockDim.x) {
        const float t = tab[e];
        const int cc = c0 + (e & 63);
        if (t != 0.f && cc < D) atomicAdd(&dE[(int64_t)(e >> 6) * ldde + cc], scale * t);
    }
}

// ------------------------------------------------------------------------------------------------------------
// positional embedding                                             mem_transformer.py:13-23, 550-558
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void pos_emb_kernel(const float* __restrict__ inv_freq, T* __restrict__ pe, int64_t ld, int klen, int D,
                               int DP, int clamp_len, float drop_scale, uint32_t thresh, uint64_t seed,
                               uint64_t site) {
    int p = blockIdx.x;
    float d = (float)(klen - 1 - p);
    if (clamp_len > 0) d = fminf(d, (float)clamp_len);
    int half = D / 2;
    for (int c = threadIdx.x; c < DP; c += blockDim.x) {
        float v = 0.f;
        if (c < D) {
            float ang = d * inv_freq[c < half ? c : c - half];
            v = c < half ? sinf(ang) : cosf(ang);
            if (thresh) v = dropout_keep(seed, site, (uint64_t)p * ld + c, thresh) ? v * drop_scale : 0.f;
        }
        pe[(int64_t)p * ld + c] = from_f<T>(v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// LayerNorm forward (warp per row; z fp32 -> y T)                   mem_transformer.py:58, 255
// Each lane owns the 4-column groups lane*4 + 128*u: the row is read ONCE with 16-byte loads, kept in registers for
// the mean / variance / normalise passes, and written with 8- (bf16) or 16-byte (fp32) stores.
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load4(const float* p, float* o) {
    float4 a = *reinterpret_cast<const float4*>(p);
    o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* p, float* o) {
    uint2 r = *reinterpret_cast<const uint2*>(p);
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
    float2 f0 = __bfloat1622float2(h[0]), f1 = __bfloat1622float2(h[1]);
    o[0] = f0.x; o[1] = f0.y; o[2] = f1.x; o[3] = f1.y;
}
__device__ __forceinline__ void store4(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void store4(bf16* p, const float* v) {
    uint2 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
    h[0] = __floats2bfloat162_rn(v[0], v[1]);
    h[1] = __floats2bfloat162_rn(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = r;
}

template <typename T, int MAXU>
__global__ void ln_fwd_kernel(const float* __restrict__ z, int64_t ldz, T* __restrict__ y, int64_t ldy,
                              const float* __restrict__ gamma, const float* __restrict__ beta,
                              float* __restrict__ mean, float* __restrict__ rstd, int rows, int D, int DP, int pad_one,
                              float eps) {
    int row = blockIdx.x * WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* zr = z + (int64_t)row * ldz;
    float v[MAXU][4];
    float s = 0.f;
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < DP) {
            load4(zr + c, v[u]);
#pragma unroll
            for (int t = 0; t < 4; ++t) { v[u][t] = c + t < D ? v[u][t] : 0.f; s += v[u][t]; }
        } else {
#pragma unroll
            for (int t = 0; t < 4; ++t) v[u][t] = 0.f;
        }
    }
    const float mu = warp_sum(s) / D;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int c = lane * 4 + 128 * u;
#pragma unroll
        for (int t = 0; t < 4; ++t) { const float d = c + t < D ? v[u][t] - mu : 0.f; q += d * d; }
    }
    const float rs = rsqrtf(warp_sum(q) / D + eps);
    if (lane == 0) { mean[row] = mu; rstd[row] = rs; }
    T* yr = y + (int64_t)row * ldy;
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < DP) {
            float g[4], bt[4], o[4];
            load4(gamma + c, g); load4(beta + c, bt);
#pragma unroll
            for (int t = 0; t < 4; ++t)
                o[t] = c + t < D ? (v[u][t] - mu) * rs * g[t] + bt[t] : ((pad_one && c + t == D) ? 1.f : 0.f);
            store4(yr + c, o);
        }
    }
}

// LayerNorm backward.  One warp per row (same column ownership as the forward); dgamma/dbeta partials are reduced
// per block in shared memory and added to global memory with one atomic per column per block.
template <typename T, int MAXU>
__global__ void ln_bwd_kernel(const T* __restrict__ dy, int64_t lddy, const float* __restrict__ z, int64_t ldz,
                              const float* __restrict__ gamma, const float* __restrict__ mean,
                              const float* __restrict__ rstd, T* __restrict__ dz, int64_t lddz,
                              T* __restrict__ dzd, int64_t lddd, float* __restrict__ dgamma,
                              float* __restrict__ dbeta, float* __restrict__ dsum, int rows, int D, int DP,
                              int rows_per_block, float drop_scale, uint32_t thresh, uint64_t seed, uint64_t site) {
    extern __shared__ float sm[];  // [3][DP]
    float* s_dg = sm;
    float* s_db = sm + DP;
    float* s_dd = sm + 2 * DP;
    for (int c = threadIdx.x; c < 3 * DP; c += blockDim.x) sm[c] = 0.f;
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t key = step_fold(dropout_key(seed, site));
    float pg[MAXU][4], pb[MAXU][4], pd[MAXU][4], gm[MAXU][4];
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < DP) load4(gamma + c, gm[u]);
#pragma unroll
        for (int t = 0; t < 4; ++t) { pg[u][t] = pb[u][t] = pd[u][t] = 0.f; if (c >= DP) gm[u][t] = 0.f; }
    }
    const int r_begin = blockIdx.x * rows_per_block, r_end = min(rows, r_begin + rows_per_block);
    for (int row = r_begin + warp; row < r_end; row += WPB) {
        const T* dyr = dy + (int64_t)row * lddy;
        const float* zr = z + (int64_t)row * ldz;
        const float mu = mean[row], rs = rstd[row];
        float g[MAXU][4], xh[MAXU][4];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int u = 0; u < MAXU; ++u) {
            const int c = lane * 4 + 128 * u;
            if (c < DP) {
                float d4[4], z4[4];
                load4(dyr + c, d4);
                load4(zr + c, z4);
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const bool ok = c + t < D;
                    const float x = ok ? (z4[t] - mu) * rs : 0.f;
                    const float dv = ok ? d4[t] : 0.f;
                    pg[u][t] += dv * x;
                    pb[u][t] += dv;
                    const float gg = dv * gm[u][t];
                    g[u][t] = gg; xh[u][t] = x;
                    s1 += gg; s2 += gg * x;
                }
            } else {
#pragma unroll
                for (int t = 0; t < 4; ++t) { g[u][t] = 0.f; xh[u][t] = 0.f; }
            }
        }
        s1 = warp_sum(s1) / D; s2 = warp_sum(s2) / D;
#pragma unroll
        for (int u = 0; u < MAXU; ++u) {
            const int c = lane * 4 + 128 * u;
            if (c < DP) {
                float o[4], od[4];
                uint32_t keep = 0xfu;
                if (thresh && dzd) {
                    const uint64_t e0 = (uint64_t)row * lddd + c;  // multiple of 4: two hash pairs
                    const uint32_t h0 = dropout_hash_pair(key, e0 >> 1), h1 = dropout_hash_pair(key, (e0 >> 1) + 1);
                    keep = (uint32_t)((h0 & 0xffffu) >= thresh) | ((uint32_t)((h0 >> 16) >= thresh) << 1) |
                           ((uint32_t)((h1 & 0xffffu) >= thresh) << 2) | ((uint32_t)((h1 >> 16) >= thresh) << 3);
                }
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    const float val = c + t < D ? rs * (g[u][t] - s1 - xh[u][t] * s2) : 0.f;
                    o[t] = val;
                    od[t] = ((keep >> t) & 1) ? val * drop_scale : 0.f;
                    pd[u][t] += dzd ? od[t] : val;
                }
                store4(dz + (int64_t)row * lddz + c, o);
                if (dzd) store4(dzd + (int64_t)row * lddd + c, od);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < MAXU; ++u) {
        const int c = lane * 4 + 128 * u;
        if (c < DP) {
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                atomicAdd(&s_dg[c + t], pg[u][t]);
                atomicAdd(&s_db[c + t], pb[u][t]);
                if (dsum) atomicAdd(&s_dd[c + t], pd[u][t]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        atomicAdd(&dgamma[c], s_dg[c]);
        atomicAdd(&dbeta[c], s_db[c]);
        if (dsum) atomicAdd(&dsum[c], s_dd[c]);
    }
}

// ------------------------------------------------------------------------------------------------------------
// dropout (out of place or in place)
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void dropout_kernel(const T* __restrict__ src, int64_t lds, T* __restrict__ dst, int64_t ldd, int rows,
                               int cols8, float drop_scale, uint32_t thresh, uint64_t seed, uint64_t site) {
    int64_t total = (int64_t)rows * cols8;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = e / cols8;
        int c0 = (int)(e % cols8) * 8;
        float v[8];
        load8(src + row * lds + c0, v);
        uint32_t keep = thresh ? dropout_keep8(seed, site, (uint64_t)row * ldd + c0, thresh) : 0xffu;
#pragma unroll
        for (int t = 0; t < 8; ++t) v[t] = ((keep >> t) & 1) ? v[t] * drop_scale : 0.f;
        store8(dst + row * ldd + c0, v);
    }
}

// ------------------------------------------------------------------------------------------------------------
// cross-entropy over the small vocabulary (warp per row)            proj_adaptive_softmax.py:75-84
// ------------------------------------------------------------------------------------------------------------
__global__ void ce_fwd_kernel(const float* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ target,
                              float* __restrict__ nll, float* __restrict__ lse, int rows, int V) {
    int row = blockIdx.x * WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* l = logits + (int64_t)row * ldl;
    float m = -INFINITY;
    for (int c = lane; c < V; c += 32) m = fmaxf(m, l[c]);
    m = warp_max(m);
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(l[c] - m);
    s = warp_sum(s);
    if (lane == 0) {
        float ls = m + logf(s);
        lse[row] = ls;
        nll[row] = ls - l[target[row]];
    }
}

template <typename T>
__global__ void ce_bwd_kernel(const float* __restrict__ logits, int64_t ldl, const int64_t* __restrict__ target,
                              const float* __restrict__ lse, const float* __restrict__ dnll, T* __restrict__ dl,
                              int64_t ldd, int rows, int V, int VP) {
    int row = blockIdx.x * WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const float* l = logits + (int64_t)row * ldl;
    const float ls = lse[row], g = dnll[row];
    const int tg = (int)target[row];
    for (int c0 = lane * 8; c0 < VP; c0 += 256) {
        float o[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) {
            int c = c0 + t;
            o[t] = c < V ? (expf(l[c] - ls) - (c == tg ? 1.f : 0.f)) * g : 0.f;
        }
        store8(dl + (int64_t)row * ldd + c0, o);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Gumbel-softmax straight-through (warp per row)                    mem_transformer.py:609-628
// ------------------------------------------------------------------------------------------------------------
__global__ void gumbel_fwd_kernel(const float* __restrict__ logits, int64_t ldl, const float* __restrict__ U,
                                  int64_t ldu, float tau, const float* __restrict__ tau_dev, float* __restrict__ y,
                                  int64_t ldy, float* __restrict__ st, int64_t lds, int64_t* __restrict__ ids, int rows,
                                  int V, uint64_t seed, uint64_t site) {
    int row = blockIdx.x * WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (tau_dev) tau = __ldg(tau_dev);
    const float* l = logits + (int64_t)row * ldl;
    float* yr = y + (int64_t)row * ldy;
    // pass 1: perturbed, tempered logits (kept in y), running max / first argmax
    float m = -INFINITY;
    int am = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
        float u;
        if (U) u = U[(int64_t)row * ldu + c];
        else {
            uint64_t e = (uint64_t)row * V + c;
            Philox4 r = philox4x32_10(seed, step_fold_site(site), e >> 2);
            uint32_t bits = (e & 3) == 0 ? r.x : (e & 3) == 1 ? r.y : (e & 3) == 2 ? r.z : r.w;
            u = (bits >> 8) * (1.0f / 16777216.0f);  // [0, 1) like torch.rand
        }
        float g = -logf(-logf(u + 1e-20f) + 1e-20f);
        float t = (l[c] + g) / tau;
        yr[c] = t;
        if (t > m) { m = t; am = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, m, o);
        int oa = __shfl_xor_sync(0xffffffffu, am, o);
        if (om > m || (om == m && oa < am)) { m = om; am = oa; }
    }
    float s = 0.f;
    for (int c = lane; c < V; c += 32) s += expf(yr[c] - m);
    s = warp_sum(s);
    const float inv = 1.f / s;
    // the reference takes argmax of the softmax output; exp() is monotone so the first maximal tempered logit
    // is also the first maximal probability unless two probabilities round to the same float -- then torch's
    // max returns the first index of the rounded maximum: replicate by comparing the rounded values.
    float pm = -1.f;
    int pa = 0x7fffffff;
    for (int c = lane; c < V; c += 32) {
        float p = expf(yr[c] - m) * inv;
        yr[c] = p;
        if (p > pm) { pm = p; pa = c; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        float om = __shfl_xor_sync(0xffffffffu, pm, o);
        int oa = __shfl_xor_sync(0xffffffffu, pa, o);
        if (om > pm || (om == pm && oa < pa)) { pm = om; pa = oa; }
    }
    if (lane == 0 && ids) ids[row] = pa;
    if (st) {
        float* sr = st + (int64_t)row * lds;
        for (int c = lane; c < V; c += 32) {
            float p = yr[c];
            sr[c] = ((c == pa ? 1.f : 0.f) - p) + p;  // (y_hard - y).detach() + y, same fp32 rounding
        }
    }
}

__global__ void gumbel_bwd_kernel(const float* __restrict__ y, int64_t ldy, const float* __restrict__ dst,
                                  int64_t lds, float tau, const float* __restrict__ tau_dev, float* __restrict__ dl,
                                  int64_t ldd, int rows, int V) {
    int row = blockIdx.x * WPB + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    if (tau_dev) tau = __ldg(tau_dev);
    const float* yr = y + (int64_t)row * ldy;
    const float* dr = dst + (int64_t)row * lds;
    float dot = 0.f;
    for (int c = lane; c < V; c += 32) dot += yr[c] * dr[c];
    dot = warp_sum(dot);
    const float it = 1.f / tau;
    for (int c = lane; c < V; c += 32) dl[(int64_t)row * ldd + c] = it * yr[c] * (dr[c] - dot);
}

// ------------------------------------------------------------------------------------------------------------
// column sums (bias gradients): out[n] += sum_m x[m, n]
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void colsum_kernel(const T* __restrict__ x, int64_t ld, float* __restrict__ out, int rows, int cols,
                              int rows_per_block) {
    // block = 32 x 8 threads: threadIdx.x -> column, threadIdx.y -> row phase
    __shared__ float part[8][33];
    const int c = blockIdx.x * 32 + threadIdx.x;
    const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    float s = 0.f;
    if (c < cols)
        for (int r = r0 + threadIdx.y; r < r1; r += 8) s += to_f(x[(int64_t)r * ld + c]);
    part[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0 && c < cols) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += part[i][threadIdx.x];
        atomicAdd(&out[c], t);
    }
}

// ------------------------------------------------------------------------------------------------------------
// convert / pad
// ------------------------------------------------------------------------------------------------------------
template <typename TS, typename TD>
__global__ void convert_kernel(const TS* __restrict__ src, int64_t lds, TD* __restrict__ dst, int64_t ldd,
                               int64_t rows, int cols, int cols_pad) {
    int64_t total = rows * cols_pad;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t r = e / cols_pad;
        int c = (int)(e % cols_pad);
        dst[r * ldd + c] = from_f<TD>(c < cols ? to_f(src[r * lds + c]) : 0.f);
    }
}

// ------------------------------------------------------------------------------------------------------------
// parameter packing / gradient unpacking (descriptor table, one launch)
// ------------------------------------------------------------------------------------------------------------
// block = 32 x 8 threads; 32-bit index arithmetic (every tensor of the model has < 2^31 elements); rows are walked by
// thread rows so both sides of the copy are coalesced, transposed destinations go through a 32 x 32 shared-memory tile
template <typename T>
__global__ void __launch_bounds__(256) pack_kernel(T* __restrict__ mat, float* __restrict__ vec,
                                                   const int64_t* __restrict__ desc) {
    const int64_t* d = desc + 12 * blockIdx.y;
    const float* __restrict__ src = reinterpret_cast<const float*>(d[0]);
    const int64_t dst_off = d[1], ld = d[4];
    const int rows = (int)d[2], cols = (int)d[3], rg = (int)d[5], rgp = (int)d[6], cg = (int)d[7], cgp = (int)d[8];
    const bool tr = d[9] != 0, is_vec = d[10] != 0;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    if (!tr) {
        for (int r = blockIdx.x * 8 + ty; r < rows; r += gridDim.x * 8) {
            const int64_t o = dst_off + (int64_t)((r / rg) * rgp + r % rg) * ld;
            const float* sr = src + (int64_t)r * cols;
            for (int c = tx; c < cols; c += 32) {
                const int cp = (c / cg) * cgp + c % cg;
                if (is_vec) vec[o + cp] = sr[c];
                else mat[o + cp] = from_f<T>(sr[c]);
            }
        }
        return;
    }
    __shared__ float tile[32][33];
    const int tiles_c = (cols + 31) >> 5, tiles = ((rows + 31) >> 5) * tiles_c;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int r0 = (t / tiles_c) << 5, c0 = (t % tiles_c) << 5;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int r = r0 + ty + 8 * k, c = c0 + tx;
            tile[ty + 8 * k][tx] = (r < rows && c < cols) ? src[(int64_t)r * cols + c] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int c = c0 + ty + 8 * k, r = r0 + tx;
            if (r < rows && c < cols) {
                const int cp = (c / cg) * cgp + c % cg, rp = (r / rg) * rgp + r % rg;
                mat[dst_off + (int64_t)cp * ld + rp] = from_f<T>(tile[tx][ty + 8 * k]);
            }
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ mat, const float* __restrict__ vec,
                                                     const int64_t* __restrict__ desc, int accumulate) {
    const int64_t* d = desc + 12 * blockIdx.y;
    float* __restrict__ dst = reinterpret_cast<float*>(d[0]);
    const int64_t off = d[1], ld = d[4];
    const int rows = (int)d[2], cols = (int)d[3], rg = (int)d[5], rgp = (int)d[6], cg = (int)d[7], cgp = (int)d[8];
    const float* __restrict__ srcp = d[10] != 0 ? vec : mat;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = blockIdx.x * 8 + ty; r < rows; r += gridDim.x * 8) {
        const float* sr = srcp + off + (int64_t)((r / rg) * rgp + r % rg) * ld;
        float* dr = dst + (int64_t)r * cols;
        for (int c = tx; c < cols; c += 32) {
            const float v = sr[(c / cg) * cgp + c % cg];
            dr[c] = accumulate ? dr[c] + v : v;
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// optimizer side: sum of squares, fused clip + Adam
// ------------------------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
    float s = 0.f;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float v = x[e];
        s += v * v;
    }
    s = warp_sum(s);
    __shared__ float part[32];
    if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        float t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
        t = warp_sum(t);
        if (threadIdx.x == 0) atomicAdd(out, t);
    }
}

__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                            float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                            float bc1, float bc2, const float* __restrict__ gnorm_sq, float clip, float grad_scale) {
    float cs = grad_scale;
    if (gnorm_sq && clip > 0.f) {
        float nrm = sqrtf(*gnorm_sq) * grad_scale;
        float coef = clip / (nrm + 1e-6f);  // torch.nn.utils.clip_grad_norm_
        if (coef < 1.f) cs *= coef;
    }
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        float gr = g[e] * cs;
        float pv = p[e];
        if (wd != 0.f) gr += wd * pv;
        float mm = b1 * m[e] + (1.f - b1) * gr;
        float vv = b2 * v[e] + (1.f - b2) * gr * gr;
        m[e] = mm; v[e] = vv;
        float denom = sqrtf(vv) / sqrtf(bc2) + eps;  // torch.optim.Adam
        p[e] = pv - (lr / bc1) * mm / denom;
    }
}

// LAMB (lamb.py:57-118; "paper v3": no bias correction) over a flat parameter buffer cut into per-tensor chunks.
// chunk table: int64 [n_chunks][3] = (tensor id, first element, element count <= LAMB_CHUNK).
// stage 1: moments, adam_step = m / (sqrt(v) + eps) + wd * p (kept in `upd`), per-tensor sums of p^2 and adam_step^2;
// stage 2: p -= lr * trust * adam_step with trust = clamp(|p|, 0, 10) / (|adam_step| + eps), 1 when either norm is 0.
constexpr int LAMB_CHUNK = 16384;
__global__ void __launch_bounds__(256)
lamb_stage1_kernel(const float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                   float* __restrict__ upd, const int64_t* __restrict__ chunks, float* __restrict__ norms, float b1,
                   float b2, float eps, float wd, const float* __restrict__ gnorm_sq, float clip, float grad_scale) {
    __shared__ float red[2][8];
    const int64_t tid = chunks[3 * blockIdx.x], e0 = chunks[3 * blockIdx.x + 1], cnt = chunks[3 * blockIdx.x + 2];
    float cs = grad_scale;
    if (gnorm_sq && clip > 0.f) {
        const float coef = clip / (sqrtf(*gnorm_sq) * grad_scale + 1e-6f);  // torch.nn.utils.clip_grad_norm_
        if (coef < 1.f) cs *= coef;
    }
    float sp = 0.f, su = 0.f;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) {
        const int64_t e = e0 + i;
        const float gr = g[e] * cs, pv = p[e];
        const float mm = b1 * m[e] + (1.f - b1) * gr;
        const float vv = b2 * v[e] + (1.f - b2) * gr * gr;
        m[e] = mm; v[e] = vv;
        const float a = mm / (sqrtf(vv) + eps) + wd * pv;
        upd[e] = a;
        sp = fmaf(pv, pv, sp);
        su = fmaf(a, a, su);
    }
    sp = warp_sum(sp); su = warp_sum(su);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { red[0][warp] = sp; red[1][warp] = su; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        atomicAdd(&norms[2 * tid + threadIdx.x], t);
    }
}

__global__ void __launch_bounds__(256)
lamb_stage2_kernel(float* __restrict__ p, const float* __restrict__ upd, const int64_t* __restrict__ chunks,
                   const float* __restrict__ norms, float lr, float eps, int adam) {
    const int64_t tid = chunks[3 * blockIdx.x], e0 = chunks[3 * blockIdx.x + 1], cnt = chunks[3 * blockIdx.x + 2];
    const float wn = fminf(sqrtf(norms[2 * tid]), 10.f), un = sqrtf(norms[2 * tid + 1]);
    const float trust = (adam || wn == 0.f || un == 0.f) ? 1.f : wn / (un + eps);
    const float step = lr * trust;
    for (int64_t i = threadIdx.x; i < cnt; i += blockDim.x) p[e0 + i] -= step * upd[e0 + i];
}

inline int grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    int64_t cap = 148LL * 16;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}
}  // namespace

// ------------------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------------------
#define ST ((cudaStream_t)stream)

extern "C" int tgan_embed_fwd(int dtype, const int64_t* ids, const void* E, int64_t lde, void* out, int64_t ldo,
                              int rows, int D, int DP, float scale, float drop_p, uint64_t seed, uint64_t site,
                              void* stream) {
    if (rows <= 0) return 0;
    TGAN_CHECK_ARG(DP % 8 == 0 && lde % 8 == 0 && ldo % 8 == 0 && D <= DP, "tgan_embed_fwd: DP/ld must be multiples of 8");
    uint32_t th = drop_p > 0.f ? dropout_thresh(drop_p) : 0u;
    float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    DISPATCH_T(dtype, (embed_fwd_kernel<T><<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(
                          ids, (const T*)E, lde, (T*)out, ldo, rows, D, DP, scale, ds, th, seed, site)));
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_embed_bwd(int dtype, const int64_t* ids, const void* dout, int64_t ldo, float* dE, int64_t ldde,
                              int rows, int V, int D, float scale, float drop_p, uint64_t seed, uint64_t site,
                              void* stream) {
    if (rows <= 0) return 0;
    const size_t smem = (size_t)V * 64 * sizeof(float);
    TGAN_CHECK_ARG(smem <= 200 * 1024, "tgan_embed_bwd: vocabulary too large for the shared-memory table (V <= 800)");
    uint32_t th = drop_p > 0.f ? dropout_thresh(drop_p) : 0u;
    float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    int slabs = min(ceil_div(rows, 256), 32);
    int rpb = ceil_div(rows, slabs);
    slabs = ceil_div(rows, rpb);
    dim3 grid(ceil_div(D, 64), slabs);
    if (dtype == TGAN_F32) {
        TGAN_CUDA_OK(cudaFuncSetAttribute(embed_bwd_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        embed_bwd_kernel<float><<<grid, 512, smem, ST>>>(ids, (const float*)dout, ldo, dE, ldde, rows, V, D, scale, ds, th, seed, site, rpb);
    } else if (dtype == TGAN_BF16) {
        TGAN_CUDA_OK(cudaFuncSetAttribute(embed_bwd_kernel<bf16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        embed_bwd_kernel<bf16><<<grid, 512, smem, ST>>>(ids, (const bf16*)dout, ldo, dE, ldde, rows, V, D, scale, ds, th, seed, site, rpb);
    } else { tgan_set_error("bad dtype %d", dtype); return 1; }
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_pos_emb(int dtype, const float* inv_freq, void* pe, int64_t ld, int klen, int D, int DP,
                            int clamp_len, float drop_p, uint64_t seed, uint64_t site, void* stream) {
    if (klen <= 0) return 0;
    uint32_t th = drop_p > 0.f ? dropout_thresh(drop_p) : 0u;
    float ds = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
    DISPATCH_T(dtype, (pos_emb_kernel<T><<<klen, 128, 0, ST>>>(inv_freq, (T*)pe, ld, klen, D, DP, clamp_len, ds, th,
                                                                seed, site)));
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_ln_fwd_eps(int dtype, const float* z, int64_t ldz, void* y, int64_t ldy, const float* gamma,
                               const float* beta, float* mean, float* rstd, int rows, int D, int DP, int pad_one,
                               float eps, void* stream) {
    if (rows <= 0) return 0;
    TGAN_CHECK_ARG(DP % 8 == 0 && ldz % 4 == 0 && ldy % 8 == 0 && D <= DP, "tgan_ln_fwd: alignment");
    TGAN_CHECK_ARG(!pad_one || D < DP, "tgan_ln_fwd: pad_one needs a pad lane (D < DP)");
    TGAN_CHECK_ARG(DP <= 1024 && (((uintptr_t)gamma | (uintptr_t)beta | (uintptr_t)z) & 15) == 0 && ((uintptr_t)y & 7) == 0,
                   "tgan_ln_fwd: DP <= 1024, 16-byte aligned z / gamma / beta");
    if (DP <= 512) {
        DISPATCH_T(dtype, (ln_fwd_kernel<T, 4><<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(z, ldz, (T*)y, ldy, gamma, beta,
                                                                                          mean, rstd, rows, D, DP, pad_one, eps)));
    } else {
        DISPATCH_T(dtype, (ln_fwd_kernel<T, 8><<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(z, ldz, (T*)y, ldy, gamma, beta,
                                                                                          mean, rstd, rows, D, DP, pad_one, eps)));
    }
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_ln_fwd(int dtype, const float* z, int64_t ldz, void* y, int64_t ldy, const float* gamma,
                           const float* beta, float* mean, float* rstd, int rows, int D, int DP, int pad_one,
                           void* stream) {
    return tgan_ln_fwd_eps(dtype, z, ldz, y, ldy, gamma, beta, mean, rstd, rows, D, DP, pad_one, 1e-5f, stream);
}

extern "C" int tgan_ln_bwd(int dtype, const void* dy, int64_t lddy, const float* z, int64_t ldz, const float* gamma,
                           const float* mean, const float* rstd, void* dz, int64_t lddz, void* dz_drop, int64_t lddd,
                           float* dgamma, float* dbeta, float* dsum, int rows, int D, int DP, float drop_p,
                           uint64_t seed, uint64_t site, void* stream) {
    if (rows <= 0) return 0;
    TGAN_CHECK_ARG(DP % 8 == 0 && DP <= 1024 && lddy % 8 == 0 && lddz % 8 == 0 && (!dz_drop || lddd % 8 == 0) &&
                       ldz % 4 == 0 && (((uintptr_t)gamma | (uintptr_t)z) & 15) == 0,
                   "tgan_ln_bwd: DP <= 1024, multiples of 8, 16-byte aligned z / gamma");
    uint32_t th = (drop_p > 0.f && dz_drop) ? dropout_thresh(drop_p) : 0u;
    float ds = (drop_p > 0.f && dz_drop) ? 1.f / (1.f - drop_p) : 1.f;
    int blocks = min(ceil_div(rows, WPB), 148 * 4);  // (4 rows per warp for small calls was measured: 14 -> 19 us at 512 rows)
    int rpb = ceil_div(rows, blocks);
    blocks = ceil_div(rows, rpb);
    size_t smem = 3 * DP * sizeof(float);
    if (DP <= 512) {
        DISPATCH_T(dtype, (ln_bwd_kernel<T, 4><<<blocks, WPB * 32, smem, ST>>>(
                              (const T*)dy, lddy, z, ldz, gamma, mean, rstd, (T*)dz, lddz, (T*)dz_drop, lddd, dgamma,
                              dbeta, dsum, rows, D, DP, rpb, ds, th, seed, site)));
    } else {
        DISPATCH_T(dtype, (ln_bwd_kernel<T, 8><<<blocks, WPB * 32, smem, ST>>>(
                              (const T*)dy, lddy, z, ldz, gamma, mean, rstd, (T*)dz, lddz, (T*)dz_drop, lddd, dgamma,
                              dbeta, dsum, rows, D, DP, rpb, ds, th, seed, site)));
    }
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_dropout(int dtype, const void* src, int64_t lds, void* dst, int64_t ldd, int rows, int cols,
                            float p, uint64_t seed, uint64_t site, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    TGAN_CHECK_ARG(cols % 8 == 0 && lds % 8 == 0 && ldd % 8 == 0, "tgan_dropout: cols/ld must be multiples of 8");
    uint32_t th = p > 0.f ? dropout_thresh(p) : 0u;
    float ds = p > 0.f ? 1.f / (1.f - p) : 1.f;
    int64_t n = (int64_t)rows * (cols / 8);
    DISPATCH_T(dtype, (dropout_kernel<T><<<grid_for(n, 256), 256, 0, ST>>>((const T*)src, lds, (T*)dst, ldd, rows,
                                                                            cols / 8, ds, th, seed, site)));
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_ce_fwd(const float* logits, int64_t ldl, const int64_t* target, float* nll, float* lse, int rows,
                           int V, void* stream) {
    if (rows <= 0) return 0;
    ce_fwd_kernel<<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(logits, ldl, target, nll, lse, rows, V);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_ce_bwd(int dtype, const float* logits, int64_t ldl, const int64_t* target, const float* lse,
                           const float* dnll, void* dlogits, int64_t ldd, int rows, int V, int VP, void* stream) {
    if (rows <= 0) return 0;
    TGAN_CHECK_ARG(VP % 8 == 0 && ldd % 8 == 0, "tgan_ce_bwd: VP/ldd must be multiples of 8");
    DISPATCH_T(dtype, (ce_bwd_kernel<T><<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(logits, ldl, target, lse, dnll,
                                                                                   (T*)dlogits, ldd, rows, V, VP)));
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_gumbel_st_fwd(const float* logits, int64_t ldl, const float* U, int64_t ldu, float tau,
                                  const float* tau_dev, float* y, int64_t ldy, float* st, int64_t lds, int64_t* ids,
                                  int rows, int V, uint64_t seed, uint64_t site, void* stream) {
    if (rows <= 0) return 0;
    TGAN_CHECK_ARG(y != nullptr && (tau_dev != nullptr || tau > 0.f), "tgan_gumbel_st_fwd: y buffer and tau > 0 required");
    gumbel_fwd_kernel<<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(logits, ldl, U, ldu, tau, tau_dev, y, ldy, st, lds, ids,
                                                                 rows, V, seed, site);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_gumbel_st_bwd(const float* y, int64_t ldy, const float* dst, int64_t lds, float tau,
                                  const float* tau_dev, float* dlogits, int64_t ldd, int rows, int V, void* stream) {
    if (rows <= 0) return 0;
    gumbel_bwd_kernel<<<ceil_div(rows, WPB), WPB * 32, 0, ST>>>(y, ldy, dst, lds, tau, tau_dev, dlogits, ldd, rows, V);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_colsum(int dtype, const void* x, int64_t ld, float* out, int rows, int cols, void* stream) {
    if (rows <= 0 || cols <= 0) return 0;
    int by = min(ceil_div(rows, 64), 64);
    int rpb = ceil_div(rows, by);
    by = ceil_div(rows, rpb);
    dim3 grid(ceil_div(cols, 32), by), block(32, 8);
    DISPATCH_T(dtype, (colsum_kernel<T><<<grid, block, 0, ST>>>((const T*)x, ld, out, rows, cols, rpb)));
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_convert(int dtype_src, const void* src, int64_t lds, int dtype_dst, void* dst, int64_t ldd,
                            int64_t rows, int cols, int cols_pad, void* stream) {
    if (rows <= 0 || cols_pad <= 0) return 0;
    int g = grid_for(rows * cols_pad, 256);
#define CONV(TS, TD) convert_kernel<TS, TD><<<g, 256, 0, ST>>>((const TS*)src, lds, (TD*)dst, ldd, rows, cols, cols_pad)
    if (dtype_src == TGAN_F32 && dtype_dst == TGAN_F32) CONV(float, float);
    else if (dtype_src == TGAN_F32 && dtype_dst == TGAN_BF16) CONV(float, bf16);
    else if (dtype_src == TGAN_BF16 && dtype_dst == TGAN_F32) CONV(bf16, float);
    else if (dtype_src == TGAN_BF16 && dtype_dst == TGAN_BF16) CONV(bf16, bf16);
    else { tgan_set_error("tgan_convert: bad dtype"); return 1; }
#undef CONV
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_pack_params(int dtype, void* packed_mat, float* packed_vec, const int64_t* desc, int n_desc,
                                int64_t max_elems, void* stream) {
    if (n_desc <= 0) return 0;
    dim3 grid(grid_for(max_elems, 256) > 64 ? 64 : grid_for(max_elems, 256), n_desc);
    DISPATCH_T(dtype, (pack_kernel<T><<<grid, 256, 0, ST>>>((T*)packed_mat, packed_vec, desc)));
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_unpack_grads(const float* padded_mat, const float* padded_vec, const int64_t* desc, int n_desc,
                                 int64_t max_elems, int accumulate, void* stream) {
    if (n_desc <= 0) return 0;
    dim3 grid(grid_for(max_elems, 256) > 64 ? 64 : grid_for(max_elems, 256), n_desc);
    unpack_kernel<<<grid, 256, 0, ST>>>(padded_mat, padded_vec, desc, accumulate);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_sumsq(const float* x, int64_t n, float* out, void* stream) {
    if (n <= 0) return 0;
    sumsq_kernel<<<grid_for(n, 256) > 296 ? 296 : grid_for(n, 256), 256, 0, ST>>>(x, n, out);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_adam_step(float* param, const float* grad, float* m, float* v, int64_t n, float lr, float beta1,
                              float beta2, float eps, float weight_decay, int step, const float* gnorm_sq, float clip,
                              float grad_scale, void* stream) {
    if (n <= 0) return 0;
    float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
    adam_kernel<<<grid_for(n, 256), 256, 0, ST>>>(param, grad, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2,
                                                   gnorm_sq, clip, grad_scale);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// LAMB: lamb.py:57-118 (+ clip_grad_norm_, train.py:914-921) on flat buffers
// ------------------------------------------------------------------------------------------------------------
extern "C" int tgan_lamb_step(float* param, const float* grad, float* m, float* v, float* upd, const int64_t* chunks,
                              int n_chunks, float* norms, int n_tensors, float lr, float beta1, float beta2, float eps,
                              float weight_decay, const float* gnorm_sq, float clip, float grad_scale, int adam,
                              void* stream) {
    if (n_chunks <= 0) return 0;
    TGAN_CHECK_ARG(param && grad && m && v && upd && chunks && norms && n_tensors > 0, "tgan_lamb_step: null argument");
    TGAN_CUDA_OK(cudaMemsetAsync(norms, 0, sizeof(float) * 2 * (size_t)n_tensors, ST));
    lamb_stage1_kernel<<<n_chunks, 256, 0, ST>>>(param, grad, m, v, upd, chunks, norms, beta1, beta2, eps, weight_decay,
                                                  gnorm_sq, clip, grad_scale);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    lamb_stage2_kernel<<<n_chunks, 256, 0, ST>>>(param, upd, chunks, norms, lr, eps, adam);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}
