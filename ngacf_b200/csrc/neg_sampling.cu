// NegSampling training / SampledNeg evaluation pieces (reference train_eval_Gowalla.py:36-88,193-270):
//   sampler: K distinct negatives per (user, positive) row, never materialising the negative sets (loadGowalla.py:56-60,80-83,101-105);
//   BCE-with-logits loss + gradient for labels [1,0,..,0] (torch.nn.BCEWithLogitsLoss, run_Gowalla.py:110);
//   rank metrics HR@k / NDCG@k of the positive among its K sampled negatives (graphattention/evaluation.py).
// Propagation, pair scoring, gradient scatter and Adam are the kernels of the PairSampling path.
#include "common.cuh"

namespace ngacf {

// one thread per row; draws until K distinct candidates are accepted (oracle/port.py:sample_negs is the specification)
__global__ void __launch_bounds__(128) sample_negs_kernel(const int* __restrict__ rows_user, const int* __restrict__ rows_item,
                                                          const int* __restrict__ all_ptr, const int* __restrict__ all_rank,
                                                          const int* __restrict__ pool, int P, int64_t row_begin, int64_t n,
                                                          const int64_t* __restrict__ row_dev, uint32_t k0, uint32_t k1, uint32_t epoch, int K,
                                                          uint32_t tag, int64_t col_stride, int64_t* __restrict__ users, int64_t* __restrict__ items) {
    const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= n) return;
    if (row_dev) { row_begin += row_dev[0]; epoch += (uint32_t)row_dev[1]; }
    const uint64_t r = (uint64_t)(row_begin + b);
    const int u = rows_user[r];
    const int beg = all_ptr[u], deg = all_ptr[u + 1] - beg;
    const int nneg = P - deg;
    // element (row b, column j): row-major b*(K+1) + j, or column-major j*col_stride + b (one column = one contiguous batch of pairs)
    const int64_t o0 = col_stride > 0 ? b : b * (K + 1), os = col_stride > 0 ? col_stride : 1;
    int64_t* out_u = users + o0;
    int64_t* out_i = items + o0;
    for (int j = 0; j <= K; ++j) out_u[j * os] = u;
    out_i[0] = rows_item[r];
    if (nneg < K) {                       // random.sample would raise; flagged to the host as -1
        for (int j = 1; j <= K; ++j) out_i[j * os] = -1;
        return;
    }
    uint32_t w[4];
    int got = 0;
    for (uint32_t d = 0; got < K; ++d) {
        if ((d & 3u) == 0u) philox4x32_10((uint32_t)r, (uint32_t)(r >> 32), epoch, tag | ((d >> 2) << 16), k0, k1, w);
        const int k = (int)__umulhi(w[d & 3u], (uint32_t)nneg);
        int lo = 0, hi = deg;             // smallest j with rank[j]-j > k
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (all_rank[beg + mid] - mid <= k) lo = mid + 1; else hi = mid;
        }
        const int64_t cand = pool[k + lo];
        bool dup = false;
        for (int j = 1; j <= got; ++j) dup |= out_i[j * os] == cand;
        if (dup) continue;
        out_i[++got * os] = cand;
    }
}

// single-block loss (deterministic tree): loss = mean(softplus(x) - y x), y = 1 for every (K+1)-th element starting at 0
__global__ void __launch_bounds__(1024) bce_logits_kernel(const float* __restrict__ x, int64_t n, int group, float* __restrict__ loss,
                                                          float* __restrict__ dx) {
    __shared__ float red[32];
    float local = 0.f;
    const float inv = 1.0f / (float)n;
    for (int64_t e = threadIdx.x; e < n; e += blockDim.x) {
        const float v = x[e];
        const float y = (group > 0 ? (e % group) == 0 : e < (int64_t)(-group)) ? 1.f : 0.f;
        const float sp = fmaxf(v, 0.f) + log1pf(__expf(-fabsf(v)));       // softplus, stable
        local += sp - y * v;
        if (dx) dx[e] = (1.0f / (1.0f + __expf(-v)) - y) * inv;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
        float v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0 && loss) *loss = v * inv;
    }
}

// one warp per row of `group` candidate scores: rank of column 0 = number of strictly larger scores; sums[0] += hit, sums[1] += ndcg
// (double atomics: order-independent up to fp64 rounding of at most n terms in [0,1])
__global__ void __launch_bounds__(256) rank_metrics_kernel(const float* __restrict__ scores, int64_t n, int group, int top_k,
                                                           double* __restrict__ sums) {
    const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= n) return;
    const float* s = scores + row * group;
    const float s0 = s[0];
    int rank = 0;
    for (int j = 1 + lane; j < group; j += 32) rank += s[j] > s0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) rank += __shfl_xor_sync(0xffffffffu, rank, o);
    if (lane == 0 && rank < top_k) {
        atomicAdd(sums + 0, 1.0);
        atomicAdd(sums + 1, 1.0 / log2((double)rank + 2.0));
    }
}

}  // namespace ngacf

using namespace ngacf;

extern "C" int ngacf_sample_negs(const int32_t* rows_user, const int32_t* rows_item, const int32_t* all_ptr, const int32_t* all_rank,
                                 const int32_t* pool, int32_t P, int64_t row_begin, int64_t row_end, const int64_t* row_dev, uint64_t seed,
                                 uint32_t epoch, int32_t K, uint32_t tag, int64_t col_stride, int64_t* users, int64_t* items, void* stream) {
    NGACF_REQUIRE(rows_user && rows_item && all_ptr && all_rank && pool && users && items, "sample_negs: null argument");
    NGACF_REQUIRE(row_end >= row_begin && P > 0 && K >= 1 && K <= 4096 && tag < 0x10000u, "sample_negs: bad range / K / tag");
    const int64_t n = row_end - row_begin;
    NGACF_REQUIRE(col_stride == 0 || col_stride >= n, "sample_negs: column stride smaller than the number of rows");
    if (n == 0) return NGACF_OK;
    sample_negs_kernel<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(rows_user, rows_item, all_ptr, all_rank, pool, P, row_begin, n, row_dev,
                                                                          (uint32_t)seed, (uint32_t)(seed >> 32), epoch, K, tag, col_stride, users, items);
    return check_launch("sample_negs");
}

extern "C" int ngacf_bce_logits_loss(const float* scores, int64_t n, int32_t group, float* loss, float* dscore, void* stream) {
    NGACF_REQUIRE(scores && n > 0 && group != 0, "bce_logits_loss: null/empty argument");
    bce_logits_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(scores, n, group, loss, dscore);
    return check_launch("bce_logits_loss");
}

extern "C" int ngacf_rank_metrics(const float* scores, int64_t n_rows, int32_t group, int32_t top_k, double* sums, void* stream) {
    NGACF_REQUIRE(scores && sums && n_rows >= 0 && group >= 1 && top_k >= 1, "rank_metrics: bad argument");
    if (n_rows == 0) return NGACF_OK;
    rank_metrics_kernel<<<ceil_div(n_rows * 32, 256), 256, 0, (cudaStream_t)stream>>>(scores, n_rows, group, top_k, sums);
    return check_launch("rank_metrics");
}
