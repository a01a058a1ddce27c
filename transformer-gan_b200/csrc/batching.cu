// Device-side batch assembly over an HBM-resident token corpus.
// Reference: MusicDataset.get_iterator / get_dis_iterator / eval_iterator (model/data_utils.py:226-304, :307-368,
// :370-434): per batch the host walks `batch_size` column trackers (sequence index, position), copies up to `bptt`
// tokens per column out of per-sequence CPU tensors into pinned [bptt, batch] LongTensors, and ships them to the GPU.
// Here the corpus (every sequence with its start token, data_utils.py:121-141) lives in HBM as one int32 array; a batch
// is ONE launch of a single CTA:
//   plan    thread 0 replays the tracker walk in column order -- it is sequential by definition (`next_idx` is handed
//           out in column order, :257-262) but touches global memory only when a column moves to a new sequence; the
//           per-column state and the current sequence lengths are staged in shared memory by all threads first;
//   gather  all threads fill data / target ([bptt, B] int64, batch index fastest = coalesced stores), pad_id beyond
//           n_new, and the per-column reset flags.
// Bytes: 2 * bptt * B * 8 written, bptt * B * 4 read (+1 token per column): ~1.3 MB for 128 x 512 -- launch-latency
// sized, far below any roofline; the point is that no token crosses PCIe and no host loop runs per batch.
#include <algorithm>

#include "common.cuh"

namespace {
constexpr int BT = 1024;

struct Plan {
    int64_t* src;   // [B] first token of the column's span in the corpus, -1: nothing
    int32_t* n_new; // [B]
};

__global__ void __launch_bounds__(BT)
batch_next_kernel(const int32_t* __restrict__ corpus, const int64_t* __restrict__ seq_off,
                  const int32_t* __restrict__ seq_len, const int32_t* __restrict__ perm, int n_seq,
                  int32_t* __restrict__ tracker, int64_t* __restrict__ plan_src, int32_t* __restrict__ plan_n,
                  int64_t* __restrict__ data, int64_t* __restrict__ target, uint8_t* __restrict__ reset,
                  int32_t* __restrict__ n_tokens, int bptt, int B, int64_t pad_id) {
    extern __shared__ __align__(16) int32_t sh[];  // idx[B], pos[B], cur_len[B], n[B]; then int64 src[B]
    int32_t* s_idx = sh;
    int32_t* s_pos = sh + B;
    int32_t* s_len = sh + 2 * B;
    int32_t* s_n = sh + 3 * B;
    int64_t* s_src = reinterpret_cast<int64_t*>(sh + 4 * B);  // 16 * B bytes in: 8-byte aligned
    for (int i = threadIdx.x; i < B; i += BT) {
        const int idx = tracker[i];
        s_idx[i] = idx;
        s_pos[i] = tracker[B + i];
        s_len[i] = idx < n_seq ? seq_len[perm[idx]] : 0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int next_idx = tracker[2 * B];
        int total = 0;
        for (int i = 0; i < B; ++i) {  // data_utils.py:250-283 (random_crop off)
            int idx = s_idx[i], pos = s_pos[i], len = s_len[i];
            int n_new = 0;
            int64_t src = -1;
            uint8_t rs = 0;
            while (idx < n_seq) {
                if (pos + 1 >= len) {          // :256-262: sequence exhausted -> next one in permutation order
                    idx = next_idx++;
                    pos = 0;
                    rs = 1;
                    len = idx < n_seq ? seq_len[perm[idx]] : 0;
                    continue;
                }
                n_new = min(len - 1 - pos, bptt);  // :272
                src = seq_off[perm[idx]] + pos;
                pos += n_new;
                break;
            }
            s_idx[i] = idx; s_pos[i] = pos; s_n[i] = n_new; s_src[i] = src;
            reset[i] = rs;
            total += n_new;
        }
        tracker[2 * B] = next_idx;
        *n_tokens = total;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < B; i += BT) {
        tracker[i] = s_idx[i];
        tracker[B + i] = s_pos[i];
        if (plan_src) { plan_src[i] = s_src[i]; plan_n[i] = s_n[i]; }
    }
    for (int e = threadIdx.x; e < bptt * B; e += BT) {
        const int t = e / B, i = e - t * B;
        const bool live = t < s_n[i];
        const int64_t p = s_src[i] + t;
        data[e] = live ? (int64_t)corpus[p] : pad_id;
        target[e] = live ? (int64_t)corpus[p + 1] : pad_id;
    }
}

// gather for a plan computed elsewhere (eval_iterator's closed form, get_dis_iterator's host-drawn offsets)
__global__ void __launch_bounds__(256)
batch_gather_kernel(const int32_t* __restrict__ corpus, const int64_t* __restrict__ plan_src,
                    const int32_t* __restrict__ plan_n, int64_t* __restrict__ data, int64_t* __restrict__ target,
                    int bptt, int B, int64_t pad_id) {
    const int64_t n = (int64_t)bptt * B;
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int t = (int)(e / B), i = (int)(e - (int64_t)t * B);
        const bool live = t < plan_n[i];
        const int64_t p = plan_src[i] + t;
        data[e] = live ? (int64_t)corpus[p] : pad_id;
        if (target) target[e] = live ? (int64_t)corpus[p + 1] : pad_id;
    }
}
}  // namespace

extern "C" int tgan_batch_next(const int32_t* corpus, const int64_t* seq_off, const int32_t* seq_len, const int32_t* perm,
                               int n_seq, int32_t* tracker, int64_t* plan_src, int32_t* plan_n, int64_t* data,
                               int64_t* target, uint8_t* reset, int32_t* n_tokens, int bptt, int B, int64_t pad_id,
                               void* stream) {
    TGAN_CHECK_ARG(corpus && seq_off && seq_len && perm && tracker && data && target && reset && n_tokens,
                   "tgan_batch_next: null argument");
    TGAN_CHECK_ARG(n_seq > 0 && bptt > 0 && B > 0 && B <= 8192, "tgan_batch_next: bad dims (batch <= 8192)");
    const size_t smem = (size_t)4 * B * sizeof(int32_t) + (size_t)B * sizeof(int64_t);
    static bool attr = false;
    if (!attr) {
        TGAN_CUDA_OK(cudaFuncSetAttribute(batch_next_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192 * 24 + 8));
        attr = true;
    }
    batch_next_kernel<<<1, BT, smem, (cudaStream_t)stream>>>(corpus, seq_off, seq_len, perm, n_seq, tracker, plan_src,
                                                             plan_n, data, target, reset, n_tokens, bptt, B, pad_id);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}

extern "C" int tgan_batch_gather(const int32_t* corpus, const int64_t* plan_src, const int32_t* plan_n, int64_t* data,
                                 int64_t* target, int bptt, int B, int64_t pad_id, void* stream) {
    TGAN_CHECK_ARG(corpus && plan_src && plan_n && data, "tgan_batch_gather: null argument");
    TGAN_CHECK_ARG(bptt > 0 && B > 0, "tgan_batch_gather: bad dims");
    const int64_t n = (int64_t)bptt * B;
    const int grid = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
    batch_gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(corpus, plan_src, plan_n, data, target, bptt, B, pad_id);
    TGAN_COUNT_LAUNCH();
    TGAN_LAUNCH_OK();
    return 0;
}
