// Micro-benchmark: sustained issue/execution rate of tcgen05.mma (bf16, cta_group::1) on one SM for the shapes the
// attention / GEMM kernels use.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17
//   -I transformer-gan_b200/csrc tools/mma_rate.cu -o gpurun_out/mma_rate ; run on the B200 box.
// Prints cycles per MMA instruction for: accumulate chains into ONE TMEM tile vs round-robin over several tiles,
// SS vs TS (A from TMEM), K-major vs MN-major A, N = 32 .. 256, M = 64 / 128.
#include <cstdio>
#include <cuda_runtime.h>

#include "tc_common.cuh"

void tgan_set_error(const char*, ...) {}
unsigned long long g_tgan_launches;
namespace tc {
EncodeTiledFn get_encode_fn() { return nullptr; }
int make_tmap_2d(CUtensorMap*, const void*, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t) { return 1; }
int make_tmap_3d(CUtensorMap*, const void*, uint64_t, uint64_t, uint64_t, uint64_t, uint64_t, uint32_t, uint32_t, uint32_t) { return 1; }
int sm_count() { return 148; }
}  // namespace tc
using namespace tc;

struct Cfg {
    int M, N, a_mn, b_mn, ts, ntiles, count;  // ntiles: accumulators used round-robin
};

__global__ void __launch_bounds__(128, 1) mma_rate_kernel(Cfg c, long long* out) {
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const uint32_t bar = base + 160 * 1024, tptr = bar + 16;
    volatile uint32_t* tp = reinterpret_cast<volatile uint32_t*>(gbase + 160 * 1024 + 16);
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 160 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(gbase)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 0) tmem_alloc(tptr, 512);
    fence_proxy_async_smem();
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tm = *tp;
    if (warp == 0) {
        const uint32_t idesc = umma_idesc_bf16(c.M, c.N, c.a_mn, c.b_mn);
        const uint64_t da = c.a_mn ? umma_smem_desc(base, 8192, 1024) : umma_smem_desc(base, 16, 1024);
        const uint64_t db = c.b_mn ? umma_smem_desc(base + 65536, 8192, 1024) : umma_smem_desc(base + 65536, 16, 1024);
        const uint32_t astep = c.a_mn ? 128 : 2, bstep = c.b_mn ? 128 : 2;
        long long t0 = 0, t1 = 0;
        for (int rep = 0; rep < 2; ++rep) {  // rep 0 warms up
            t0 = clock64();
            if (elect_one()) {
                for (int i = 0; i < c.count; ++i) {
                    const uint32_t d = tm + (i % c.ntiles) * c.N;
                    const int k = i & 3;
                    if (c.ts) umma_bf16_ts(d, tm + 448 + 8 * k, db + bstep * k, idesc, 1);
                    else umma_bf16(d, da + astep * k, db + bstep * k, idesc, 1);
                }
                umma_commit(bar);
            }
            __syncwarp();
            mbar_wait(bar, rep & 1);
            t1 = clock64();
        }
        if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) { tcgen05_fence_after(); tmem_dealloc(tm, 512); }
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 148 * sizeof(long long));
    const int SMEM = 160 * 1024 + 64 + 1024;
    cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    const Cfg cfgs[] = {
        {128, 64, 0, 0, 0, 1, 256},  {128, 64, 0, 0, 0, 2, 256},  {128, 64, 0, 0, 0, 4, 256},
        {128, 32, 0, 0, 0, 1, 256},  {128, 32, 0, 0, 0, 4, 256},  {128, 128, 0, 0, 0, 1, 256},
        {128, 128, 0, 0, 0, 2, 256}, {128, 256, 0, 0, 0, 1, 256}, {128, 256, 0, 0, 0, 2, 256 - 0},
        {128, 64, 0, 1, 0, 1, 256},  {128, 64, 0, 1, 0, 4, 256},  {64, 64, 1, 1, 0, 1, 256},
        {64, 64, 1, 1, 0, 4, 256},   {128, 64, 1, 1, 0, 1, 256},  {128, 64, 0, 1, 1, 1, 256},
        {128, 64, 0, 1, 1, 4, 256},  {64, 64, 0, 1, 0, 1, 256},   {128, 16, 0, 0, 0, 1, 256},
    };
    for (int grid : {1, 148}) {
        printf("grid %d\n", grid);
        for (const Cfg& c : cfgs) {
            if (c.ntiles * c.N > 448) continue;
            mma_rate_kernel<<<grid, 128, SMEM>>>(c, d_out);
            cudaError_t e = cudaDeviceSynchronize();
            long long h[148];
            cudaMemcpy(h, d_out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
            printf("M=%3d N=%3d A:%s B:%s %s accumulators=%d : %7.1f clk / MMA  (%s)\n", c.M, c.N, c.a_mn ? "MN" : "K ",
                   c.b_mn ? "MN" : "K ", c.ts ? "TS" : "SS", c.ntiles, (double)mx / c.count, cudaGetErrorString(e));
        }
    }
    return 0;
}
