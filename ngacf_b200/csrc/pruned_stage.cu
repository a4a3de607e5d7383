// Pruned output stage of a TRAINING step (exact).
//
// The loss of a PairSampling step reads the last stage's output only at the batch rows (graphattention/SPUIGACF.py:49-52:
// features[userIdx], features[itemIdx]), so
//   forward : that stage's aggregation is needed for the batch rows only ("active" rows);
//   backward: G = dL/dZ_last is non-zero on the active rows only, hence Ghat = dN = 0 elsewhere and d s_e vanishes on every
//             edge without an active endpoint.
// The reference computes the full stage and gathers afterwards (train_eval_Gowalla.py:131-137); everything skipped here is an
// exact zero, so losses, gradients and updated parameters are those of the full computation up to the order of additions.
//
//   mark_active   stamp[users[b]] = stamp[U+items[b]] = v.  A node is active for a propagation iff its stamp equals the
//                 propagation's value (the trainer uses the dropout call index: unique per propagation, never cleared).
//   active_plan   (a) compacted list of the tasks of active rows, (b) one activity bit per adjacency position
//                 (bit p = neighbour adj_idx[p] is active).
//   prep (list)   Ghat / dN of the active rows only -- nothing else of them is ever read.
//   user rows     an active user visits its whole row; any other user visits only its edges to active items (read off the
//                 activity bits: no adjacency or row is touched for the rest) and stores (d s_e, e*keep) for the item pass.
//   item rows     dS[i] = sum of the stored d s_e (all edges of an active item; the edges to active users otherwise);
//                 Ghat[u] rows are gathered for active users only.
// The visited edges are compacted 16 at a time; per-edge scalars (logit, weight, d s_e) are computed once per edge by the
// lane that owns it.  H = 1 (the single-head output stage of SPUIGACF / SPUIMultiGACF).
#include "common.cuh"

namespace ngacf {

__global__ void mark_active_kernel(int* __restrict__ stamp, const int64_t* __restrict__ users, const int64_t* __restrict__ items, int B, int U,
                                   int val, const int64_t* __restrict__ val_dev, int* __restrict__ list_count) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0 && list_count) { list_count[0] = 0; list_count[1] = 0; }     // the plan kernel that follows appends from zero
    if (b >= B) return;
    const int v = active_value(val, val_dev);           // same value from every writer: benign
    stamp[users[b]] = v;
    if (items[b] >= 0) stamp[(int64_t)U + items[b]] = v;
}

// One launch: blocks [0, task_blocks) compact the tasks of active rows, the remaining blocks write the activity bits.
// Two lists in one buffer: user tasks from the front (count[0] entries), item tasks from the back (count[1] entries:
// list[T-1], list[T-2], ...).  List order is arrival order: harmless, every task writes its own row (long rows combine in slot
// order whatever the arrival).
__device__ __forceinline__ int listed_task(const int* __restrict__ list, int T, int n_users, int g) {
    return __ldg(list + (g < n_users ? g : T - 1 - (g - n_users)));
}

__global__ void __launch_bounds__(256) active_plan_kernel(const int4* __restrict__ tasks, int T, int T_users, int task_blocks,
                                                          const int* __restrict__ adj_idx, int64_t n_adj, const int* __restrict__ stamp, int val,
                                                          const int64_t* __restrict__ val_dev, int* __restrict__ list, int* __restrict__ count,
                                                          uint32_t* __restrict__ bits) {
    const int av = active_value(val, val_dev);
    const int lane = threadIdx.x & 31;
    if ((int)blockIdx.x < task_blocks) {
        const int t = blockIdx.x * blockDim.x + threadIdx.x;
        const bool a = t < T && __ldg(stamp + __ldg(tasks + t).x) == av;
        const bool item = t >= T_users;
#pragma unroll
        for (int side = 0; side < 2; ++side) {               // a warp may straddle the user / item boundary
            const unsigned m = __ballot_sync(0xffffffffu, a && (item == (side == 1)));
            if (m == 0) continue;
            const int leader = __ffs(m) - 1;
            int base = 0;
            if (lane == leader) base = atomicAdd(count + side, __popc(m));
            base = __shfl_sync(0xffffffffu, base, leader);
            if (a && (item == (side == 1))) {
                const int pos = base + __popc(m & ((1u << lane) - 1u));
                list[side == 0 ? pos : T - 1 - pos] = t;
            }
        }
    } else {
        const int64_t p = (int64_t)(blockIdx.x - task_blocks) * blockDim.x + threadIdx.x;
        const bool a = p < n_adj && __ldg(stamp + ld_stream_i32(adj_idx + p)) == av;
        const unsigned w = __ballot_sync(0xffffffffu, a);
        if (lane == 0 && p < n_adj) bits[p >> 5] = w;
    }
}

__global__ void __launch_bounds__(256) prep_list_kernel(const int4* __restrict__ tasks, int T, const int* __restrict__ list, const int* __restrict__ count,
                                                        const float* __restrict__ G, const float* __restrict__ Z, const float* __restrict__ h,
                                                        const float* __restrict__ norm, float* __restrict__ Ghat, float* __restrict__ dN) {
    const int n_users = __ldg(count), n_list = n_users + __ldg(count + 1);
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int stride = (gridDim.x * blockDim.x) >> 4;
    for (int g = (blockIdx.x * blockDim.x + threadIdx.x) >> 4; g < n_list; g += stride) {
        const int64_t n = __ldg(tasks + listed_task(list, T, n_users, g)).x;      // the chunks of a long row write the same values
        const float4 gg = ld_stream4(G + n * D + lane16 * 4);
        const float4 z = ld_stream4(Z + n * D + lane16 * 4);
        const float4 hh = ld_stream4(h + n * D + lane16 * 4);
        const float nr = __ldg(norm + n);
        const float inv = nr != 0.f ? 1.0f / nr : 0.f;
        float part = gg.x * (z.x - hh.x) + gg.y * (z.y - hh.y) + gg.z * (z.z - hh.z) + gg.w * (z.w - hh.w);
        part = head_reduce<1>(part, gm);
        st_stream4(Ghat + n * D + lane16 * 4, make_float4(gg.x * inv, gg.y * inv, gg.z * inv, gg.w * inv));
        if (lane16 == 0) dN[n] = -part * inv;
    }
}

// position (0-based) of the n-th set bit of x, n < popc(x)
__device__ __forceinline__ int nth_set_bit(unsigned x, int n) {
    int pos = 0;
    int c = __popc(x & 0xFFFFu);
    if (n >= c) { n -= c; pos += 16; x >>= 16; }
    c = __popc(x & 0xFFu);
    if (n >= c) { n -= c; pos += 8; x >>= 8; }
    c = __popc(x & 0xFu);
    if (n >= c) { n -= c; pos += 4; x >>= 4; }
    c = __popc(x & 0x3u);
    if (n >= c) { n -= c; pos += 2; x >>= 2; }
    if (n >= (int)(x & 1u)) pos += 1;
    return pos;
}

// the (<= 5) activity words of the adjacency range [beg,end) of a task (<= CHUNK = 128 positions), clipped to the range
struct RangeBits {
    unsigned W[5];
    int pre[6];
    int w0;
    __device__ __forceinline__ void load(const uint32_t* __restrict__ bits, int beg, int end, bool all, int lane16, unsigned gm) {
        w0 = beg >> 5;
        unsigned word = 0;
        const int w = w0 + lane16, lo = w << 5;
        if (lane16 < 5 && lo < end) {
            unsigned x = all ? 0xFFFFFFFFu : __ldg(bits + w);
            if (lo < beg) x &= 0xFFFFFFFFu << (beg - lo);
            if (lo + 32 > end) x &= 0xFFFFFFFFu >> (lo + 32 - end);
            word = x;
        }
        pre[0] = 0;
#pragma unroll
        for (int q = 0; q < 5; ++q) {
            W[q] = __shfl_sync(gm, word, q, 16);
            pre[q + 1] = pre[q] + __popc(W[q]);
        }
    }
    __device__ __forceinline__ int total() const { return pre[5]; }
    // adjacency position of the k-th set bit, k < total()
    __device__ __forceinline__ int position(int k) const {
        const int q = (k >= pre[1]) + (k >= pre[2]) + (k >= pre[3]) + (k >= pre[4]);
        const unsigned ww = q == 0 ? W[0] : q == 1 ? W[1] : q == 2 ? W[2] : q == 3 ? W[3] : W[4];
        const int pk = q == 0 ? 0 : q == 1 ? pre[1] : q == 2 ? pre[2] : q == 3 ? pre[3] : pre[4];
        return ((w0 + q) << 5) + nth_set_bit(ww, k - pk);
    }
};

// the visit rounds of one user task.  AROW: the user itself is active (whole row, both dot products, neighbours may be inactive)
template <bool DROP, bool AROW>
__device__ __forceinline__ void users_rounds(const RangeBits& rb, const int total, const int node, const int lane16, const unsigned gm,
                                             const int* __restrict__ adj_idx, const int* __restrict__ adj_eid,
                                             const float* __restrict__ Ghat, const float* __restrict__ dN, const float* __restrict__ h,
                                             const float* __restrict__ s, const uint8_t* __restrict__ edgemask, const float scale,
                                             const uint32_t* __restrict__ bits, float2* __restrict__ ds_store, float4& acc, float& dS_l) {
    const float sn = __ldg(s + node);
    const float sc = DROP ? scale : 1.f;
    const float4 hn = ld_stream4(h + (int64_t)node * D + lane16 * 4);
    float4 ghn = make_float4(0.f, 0.f, 0.f, 0.f);
    float dNn = 0.f;
    if (AROW) {
        ghn = ld_stream4(Ghat + (int64_t)node * D + lane16 * 4);
        dNn = __ldg(dN + node);
    }
    for (int r0 = 0; r0 < total; r0 += 16) {
        const int k = r0 + lane16;
        const bool mine = k < total;
        int m_c = 0, eid_c = 0, nb_act = 1;
        float s_c = 0.f, dN_c = 0.f;
        unsigned mk_c = 1u;
        if (mine) {
            const int pos = rb.position(k);
            m_c = ld_stream_i32(adj_idx + pos);
            eid_c = ld_stream_i32(adj_eid + pos);
            if (AROW) nb_act = (int)((__ldg(bits + (pos >> 5)) >> (pos & 31)) & 1u);   // Ghat / dN rows of inactive neighbours are stale
            s_c = __ldg(s + m_c);
            if (nb_act) dN_c = __ldg(dN + m_c);
            if (DROP) mk_c = edgemask[eid_c];
        }
        const float x_l = sn + s_c;
        const float e_l = edge_weight(x_l);
        const float keepsc_l = DROP ? ((mk_c & 1u) ? sc : 0.f) : 1.f;
        const float et_l = e_l * keepsc_l;
        float det_l = 0.f;
        const int cnt = min(16, total - r0);
        // visits in blocks of 4: the row gathers of a block are all in flight before the first is consumed (predicated loads for
        // the tail, no remainder loop)
        for (int j0 = 0; j0 < cnt; j0 += 4) {
            float4 g4[4], hm[4];
            float et[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int j = (j0 + q) & 15;
                const int m = __shfl_sync(gm, m_c, j, 16);
                et[q] = __shfl_sync(gm, et_l, j, 16);
                bool take = j0 + q < cnt;
                bool take_g = take;
                if (AROW) take_g = take && __shfl_sync(gm, nb_act, j, 16) != 0;     // a stale row may hold anything: do not read it
                g4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                hm[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (take_g) g4[q] = ld_gather4(Ghat + (int64_t)m * D + lane16 * 4);
                if (AROW && take) hm[q] = ld_gather4(h + (int64_t)m * D + lane16 * 4);
                if (!take) et[q] = 0.f;
            }
            NGACF_ISSUE_FENCE4(g4);
            if (AROW) NGACF_ISSUE_FENCE4(hm);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc.x = fmaf(et[q], g4[q].x, acc.x); acc.y = fmaf(et[q], g4[q].y, acc.y);
                acc.z = fmaf(et[q], g4[q].z, acc.z); acc.w = fmaf(et[q], g4[q].w, acc.w);
                float part = g4[q].x * hn.x + g4[q].y * hn.y + g4[q].z * hn.z + g4[q].w * hn.w;
                if (AROW) part += ghn.x * hm[q].x + ghn.y * hm[q].y + ghn.z * hm[q].z + ghn.w * hm[q].w;
                const float det = head_reduce<1>(part, gm);
                if (lane16 == j0 + q) det_l = det;
            }
        }
        if (mine) {
            const float de = fmaf(det_l, keepsc_l, dNn + dN_c);
            const float ds = de * (-e_l) * (x_l > 0.f ? 1.f : LRELU_ALPHA);
            ds_store[eid_c] = make_float2(ds, et_l);
            dS_l += ds;
        }
    }
}

template <bool DROP>
__global__ void __launch_bounds__(256) stage_bwd_users_active_kernel(const int4* __restrict__ tasks, int T_begin, int T_end,
                                                                     const int* __restrict__ adj_ptr, const int* __restrict__ adj_idx,
                                                                     const int* __restrict__ adj_eid, const int* __restrict__ long_first_slot,
                                                                     int* long_counter, float* scratch, const float* __restrict__ G,
                                                                     const float* __restrict__ Ghat, const float* __restrict__ dN,
                                                                     const float* __restrict__ h, const float* __restrict__ s,
                                                                     const uint8_t* __restrict__ edgemask, float scale,
                                                                     const float* const* __restrict__ wtab, int U,
                                                                     const int* __restrict__ stamp, int stamp_val, const int64_t* __restrict__ stamp_dev,
                                                                     const uint32_t* __restrict__ bits, float2* __restrict__ ds_store,
                                                                     float* __restrict__ dh, float* __restrict__ dS) {
    const int t = T_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    if (t >= T_end) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    const bool a_n = __ldg(stamp + node) == active_value(stamp_val, stamp_dev);
    if (a_n) return;                                     // the rows of active users are walked by stage_bwd_users_arow_kernel
    RangeBits rb;
    rb.load(bits, beg, end, false, lane16, gm);          // this user's edges to active items
    const int total = rb.total();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float dS_l = 0.f;
    if (total > 0)
        users_rounds<DROP, false>(rb, total, node, lane16, gm, adj_idx, adj_eid, Ghat, dN, h, s, edgemask, scale, bits, ds_store, acc, dS_l);
    float dSacc = head_reduce<1>(dS_l, gm);
    if (lid >= 0) {
        float sums[1] = {dSacc};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<1, 1, true>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, acc, sums)) return;
        dSacc = sums[0];
    }
    const float* ap = wtab[2] + (node >= U ? D : 0) + lane16 * 4;        // G = 0 on an inactive row
    const float4 a4 = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), __ldg(ap + 3));
    st_stream4(dh + (int64_t)node * D + lane16 * 4,
               make_float4(acc.x + dSacc * a4.x, acc.y + dSacc * a4.y, acc.z + dSacc * a4.z, acc.w + dSacc * a4.w));
    if (lane16 == 0) dS[node] = dSacc;
}

// ------------------------------------------------------------------------------------------------
// One CTA (8 groups of 16 lanes) per listed task: group g owns the g-th 16-edge batch of the task, so the eight batches of a
// 128-edge task run side by side instead of one after the other -- with a few thousand active rows per step the length of the
// longest dependent chain, not throughput, decides how long these kernels take (a 128-edge task walked by one group was
// ~35 us for an active user, ~12 us in the forward; CUDA-graph timing, scripts/probe/time_active_ranges.py).
// The eight partial sums meet in shared memory in batch order (deterministic).
// ------------------------------------------------------------------------------------------------
constexpr int CTA_GROUPS = CHUNK / 16;      // 8
constexpr int CTA_THREADS = CTA_GROUPS * 16;

__device__ __forceinline__ void cta_sum_groups(float4 (*part_acc)[16], float* part_s, const int grp, const int lane16, float4& acc, float& sum) {
    part_acc[grp][lane16] = acc;
    if (lane16 == 0) part_s[grp] = sum;
    __syncthreads();
    if (grp == 0) {
        float4 t = part_acc[0][lane16];
        float ts = part_s[0];
#pragma unroll
        for (int q = 1; q < CTA_GROUPS; ++q) {
            const float4 v = part_acc[q][lane16];
            t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
            ts += part_s[q];
        }
        acc = t;
        sum = ts;
    }
}

// Long rows in the CTA-per-task kernels: group 0 (which holds the task's sums) publishes them to the task's slot; the CTA that
// arrives last sums the slots with all eight groups (group g takes slots g, g+8, ... in order, four in flight; the eight group
// sums meet in group order) instead of one group walking up to 111 slots.  Called by every thread; returns true for the last
// CTA with the totals in group 0.
__device__ __forceinline__ bool cta_long_row(const int lid, const int chunk, const int* __restrict__ long_first_slot, int* long_counter,
                                             float* scratch, float4 (*part_acc)[16], float* part_s, int* flag, const int grp, const int lane16,
                                             const unsigned gm, float4& acc, float& sum) {
    const int first = long_first_slot[lid];
    const int nslots = long_first_slot[lid + 1] - first;
    if (grp == 0) {
        float* slot = scratch + (size_t)(first + chunk) * SCRATCH_STRIDE;
        *reinterpret_cast<float4*>(slot + lane16 * 4) = acc;
        if (lane16 == 0) slot[D] = sum;
        fence_acq_rel_gpu();
        __syncwarp(gm);
        if (lane16 == 0) *flag = atomicAdd(long_counter + lid, 1) == nslots - 1;
    }
    __syncthreads();
    if (!*flag) return false;                       // block-uniform
    fence_acq_rel_gpu();
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    float ts = 0.f;
    for (int c0 = grp; c0 < nslots; c0 += 4 * CTA_GROUPS) {
        float4 v[4];
        float sv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int c = c0 + q * CTA_GROUPS;
            const bool in = c < nslots;
            const float* sl = scratch + (size_t)(first + (in ? c : 0)) * SCRATCH_STRIDE;
            v[q] = in ? ld_cg4(sl + lane16 * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
            sv[q] = in ? __ldcg(sl + D) : 0.f;
        }
        NGACF_ISSUE_FENCE4(v);
#pragma unroll
        for (int q = 0; q < 4; ++q) { t.x += v[q].x; t.y += v[q].y; t.z += v[q].z; t.w += v[q].w; ts += sv[q]; }
    }
    cta_sum_groups(part_acc, part_s, grp, lane16, t, ts);
    if (threadIdx.x == 0) long_counter[lid] = 0;     // re-arm for the next launch
    acc = t;
    sum = ts;
    return true;
}

// forward of the listed rows (H = 1): same arithmetic as aggregate_fwd's task body
template <bool DROP>
__global__ void __launch_bounds__(CTA_THREADS) aggregate_fwd_cta_kernel(const int4* __restrict__ tasks, int T, const int* __restrict__ task_list,
                                                                        const int* __restrict__ task_count, const int* __restrict__ adj_ptr,
                                                                        const int* __restrict__ adj_idx, const int* __restrict__ adj_eid,
                                                                        const int* __restrict__ long_first_slot, int* long_counter,
                                                                        float* scratch, const float* __restrict__ h, const float* __restrict__ s,
                                                                        const uint8_t* __restrict__ edgemask, float scale,
                                                                        float* __restrict__ Z, float* __restrict__ norm) {
    __shared__ float4 part_acc[CTA_GROUPS][16];
    __shared__ float part_s[CTA_GROUPS];
    __shared__ int flag;
    const int n_users = __ldg(task_count), n_list = n_users + __ldg(task_count + 1);
    const int grp = threadIdx.x >> 4, lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    for (int g = blockIdx.x; g < n_list; g += gridDim.x) {
        const int4 tk = __ldg(tasks + listed_task(task_list, T, n_users, g));
        const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
        const float sn = __ldg(s + node);
        const int base = beg + grp * 16, idx = base + lane16;
        const int cnt = max(0, min(16, end - base));
        int m_l = 0;
        float w_l = 0.f, wd_l = 0.f;
        if (idx < end) {
            m_l = ld_stream_i32(adj_idx + idx);
            unsigned mk = 1u;
            if (DROP) mk = edgemask[ld_stream_i32(adj_eid + idx)];
            w_l = edge_weight(sn + __ldg(s + m_l));
            wd_l = DROP ? ((mk & 1u) ? w_l * scale : 0.f) : w_l;
        }
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = 0; j0 < cnt; j0 += 8) {
            float4 hm[8];
            float wd[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = (j0 + q) & 15;
                const int m = __shfl_sync(gm, m_l, j, 16);
                wd[q] = __shfl_sync(gm, wd_l, j, 16);
                hm[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + q < cnt) hm[q] = ld_gather4(h + (int64_t)m * D + lane16 * 4);
                else wd[q] = 0.f;
            }
            NGACF_ISSUE_FENCE8(hm);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc.x = fmaf(wd[q], hm[q].x, acc.x); acc.y = fmaf(wd[q], hm[q].y, acc.y);
                acc.z = fmaf(wd[q], hm[q].z, acc.z); acc.w = fmaf(wd[q], hm[q].w, acc.w);
            }
        }
        float rs = head_reduce<1>(w_l, gm);
        cta_sum_groups(part_acc, part_s, grp, lane16, acc, rs);
        bool last = true;
        if (lid >= 0)      // block-uniform
            last = cta_long_row(lid, (beg - __ldg(adj_ptr + node)) / CHUNK, long_first_slot, long_counter, scratch, part_acc, part_s, &flag, grp,
                                lane16, gm, acc, rs);
        if (last && grp == 0) {
            const float inv = rs != 0.f ? 1.0f / rs : 0.f;
            const float4 hn = ld_stream4(h + (int64_t)node * D + lane16 * 4);
            st_stream4(Z + (int64_t)node * D + lane16 * 4,
                       make_float4(fmaf(acc.x, inv, hn.x), fmaf(acc.y, inv, hn.y), fmaf(acc.z, inv, hn.z), fmaf(acc.w, inv, hn.w)));
            if (lane16 == 0) norm[node] = rs;
        }
        __syncthreads();          // the partial-sum buffers are reused by the next listed task
    }
}

// user pass of the ACTIVE users (their listed tasks): whole row, both dot products; inactive neighbours have Ghat = dN = 0
template <bool DROP>
__global__ void __launch_bounds__(CTA_THREADS) stage_bwd_users_arow_kernel(const int4* __restrict__ tasks, const int* __restrict__ task_list,
                                                                           const int* __restrict__ task_count, const int* __restrict__ adj_ptr,
                                                                           const int* __restrict__ adj_idx, const int* __restrict__ adj_eid,
                                                                           const int* __restrict__ long_first_slot, int* long_counter,
                                                                           float* scratch, const float* __restrict__ G,
                                                                           const float* __restrict__ Ghat, const float* __restrict__ dN,
                                                                           const float* __restrict__ h, const float* __restrict__ s,
                                                                           const uint8_t* __restrict__ edgemask, float scale,
                                                                           const float* const* __restrict__ wtab, int U,
                                                                           const uint32_t* __restrict__ bits, float2* __restrict__ ds_store,
                                                                           float* __restrict__ dh, float* __restrict__ dS) {
    __shared__ float4 part_acc[CTA_GROUPS][16];
    __shared__ float part_s[CTA_GROUPS];
    __shared__ int flag;
    const int n_list = __ldg(task_count);                          // the user tasks sit at the front of the list
    const int grp = threadIdx.x >> 4, lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const float sc = DROP ? scale : 1.f;
    for (int g = blockIdx.x; g < n_list; g += gridDim.x) {
        const int4 tk = __ldg(tasks + __ldg(task_list + g));
        const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
        const float sn = __ldg(s + node);
        const float4 hn = ld_stream4(h + (int64_t)node * D + lane16 * 4);
        const float4 ghn = ld_stream4(Ghat + (int64_t)node * D + lane16 * 4);
        const float dNn = __ldg(dN + node);
        const int base = beg + grp * 16, idx = base + lane16;
        const int cnt = max(0, min(16, end - base));
        const bool mine = idx < end;
        int m_c = 0, eid_c = 0, nb_act = 0;
        float s_c = 0.f, dN_c = 0.f;
        unsigned mk_c = 1u;
        if (mine) {
            m_c = ld_stream_i32(adj_idx + idx);
            eid_c = ld_stream_i32(adj_eid + idx);
            nb_act = (int)((__ldg(bits + (idx >> 5)) >> (idx & 31)) & 1u);
            s_c = __ldg(s + m_c);
            if (nb_act) dN_c = __ldg(dN + m_c);
            if (DROP) mk_c = edgemask[eid_c];
        }
        const float x_l = sn + s_c;
        const float e_l = edge_weight(x_l);
        const float keepsc_l = DROP ? ((mk_c & 1u) ? sc : 0.f) : 1.f;
        const float et_l = e_l * keepsc_l;
        float det_l = 0.f;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int j0 = 0; j0 < cnt; j0 += 8) {
            float4 g4[8], hm[8];
            float et[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = (j0 + q) & 15;
                const int m = __shfl_sync(gm, m_c, j, 16);
                et[q] = __shfl_sync(gm, et_l, j, 16);
                const int na = __shfl_sync(gm, nb_act, j, 16);
                const bool take = j0 + q < cnt;
                g4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                hm[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (take && na) g4[q] = ld_gather4(Ghat + (int64_t)m * D + lane16 * 4);     // a stale row is never read
                if (take) hm[q] = ld_gather4(h + (int64_t)m * D + lane16 * 4);
                else et[q] = 0.f;
            }
            NGACF_ISSUE_FENCE8(g4);
            NGACF_ISSUE_FENCE8(hm);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc.x = fmaf(et[q], g4[q].x, acc.x); acc.y = fmaf(et[q], g4[q].y, acc.y);
                acc.z = fmaf(et[q], g4[q].z, acc.z); acc.w = fmaf(et[q], g4[q].w, acc.w);
                float part = g4[q].x * hn.x + g4[q].y * hn.y + g4[q].z * hn.z + g4[q].w * hn.w
                           + ghn.x * hm[q].x + ghn.y * hm[q].y + ghn.z * hm[q].z + ghn.w * hm[q].w;
                const float det = head_reduce<1>(part, gm);
                if (lane16 == j0 + q) det_l = det;
            }
        }
        float dS_l = 0.f;
        if (mine) {
            const float de = fmaf(det_l, keepsc_l, dNn + dN_c);
            const float ds = de * (-e_l) * (x_l > 0.f ? 1.f : LRELU_ALPHA);
            ds_store[eid_c] = make_float2(ds, et_l);
            dS_l = ds;
        }
        float dSacc = head_reduce<1>(dS_l, gm);
        cta_sum_groups(part_acc, part_s, grp, lane16, acc, dSacc);
        bool last = true;
        if (lid >= 0)
            last = cta_long_row(lid, (beg - __ldg(adj_ptr + node)) / CHUNK, long_first_slot, long_counter, scratch, part_acc, part_s, &flag, grp,
                                lane16, gm, acc, dSacc);
        if (last && grp == 0) {
            const float4 gn = ld_stream4(G + (int64_t)node * D + lane16 * 4);
            const float* ap = wtab[2] + lane16 * 4;            // user half of a
            const float4 a4 = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), __ldg(ap + 3));
            st_stream4(dh + (int64_t)node * D + lane16 * 4, make_float4(gn.x + acc.x + dSacc * a4.x, gn.y + acc.y + dSacc * a4.y,
                                                                       gn.z + acc.z + dSacc * a4.z, gn.w + acc.w + dSacc * a4.w));
            if (lane16 == 0) dS[node] = dSacc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256) stage_bwd_items_active_kernel(const int4* __restrict__ tasks, int T_begin, int T_end,
                                                                     const int* __restrict__ adj_ptr, const int* __restrict__ adj_idx,
                                                                     const int* __restrict__ adj_eid, const int* __restrict__ long_first_slot,
                                                                     int* long_counter, float* scratch, const float* __restrict__ G,
                                                                     const float* __restrict__ Ghat, const float* const* __restrict__ wtab, int U,
                                                                     const int* __restrict__ stamp, int stamp_val, const int64_t* __restrict__ stamp_dev,
                                                                     const uint32_t* __restrict__ bits, const float2* __restrict__ ds_store,
                                                                     float* __restrict__ dh, float* __restrict__ dS) {
    const int t = T_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    if (t >= T_end) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    const bool a_n = __ldg(stamp + node) == active_value(stamp_val, stamp_dev);
    RangeBits rb;
    rb.load(bits, beg, end, false, lane16, gm);           // the item's edges to ACTIVE users
    const int total = rb.total();
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float dS_l = 0.f;
    if (a_n) {      // every edge of an active item carries a d s_e (stored by the user pass); <= CHUNK = 8 x 16 edges per task
        int eid[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int idx = beg + q * 16 + lane16;
            eid[q] = idx < end ? ld_stream_i32(adj_eid + idx) : -1;
        }
        asm volatile("" : "+r"(eid[0]), "+r"(eid[1]), "+r"(eid[2]), "+r"(eid[3]), "+r"(eid[4]), "+r"(eid[5]), "+r"(eid[6]), "+r"(eid[7]));
        float dsv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) dsv[q] = eid[q] >= 0 ? __ldg(ds_store + eid[q]).x : 0.f;
        asm volatile("" : "+f"(dsv[0]), "+f"(dsv[1]), "+f"(dsv[2]), "+f"(dsv[3]), "+f"(dsv[4]), "+f"(dsv[5]), "+f"(dsv[6]), "+f"(dsv[7]));
#pragma unroll
        for (int q = 0; q < 8; ++q) dS_l += dsv[q];
    }
    for (int r0 = 0; r0 < total; r0 += 16) {
        const int k = r0 + lane16;
        int m_c = 0;
        float y_c = 0.f;
        if (k < total) {
            const int pos = rb.position(k);
            m_c = ld_stream_i32(adj_idx + pos);
            const float2 pr = __ldg(ds_store + ld_stream_i32(adj_eid + pos));
            y_c = pr.y;
            if (!a_n) dS_l += pr.x;
        }
        const int cnt = min(16, total - r0);
        for (int j0 = 0; j0 < cnt; j0 += 8) {
            float4 g4[8];
            float y[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int j = (j0 + q) & 15;
                const int m = __shfl_sync(gm, m_c, j, 16);
                y[q] = __shfl_sync(gm, y_c, j, 16);
                g4[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + q < cnt) g4[q] = ld_gather4(Ghat + (int64_t)m * D + lane16 * 4);
                else y[q] = 0.f;
            }
            NGACF_ISSUE_FENCE8(g4);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                acc.x = fmaf(y[q], g4[q].x, acc.x); acc.y = fmaf(y[q], g4[q].y, acc.y);
                acc.z = fmaf(y[q], g4[q].z, acc.z); acc.w = fmaf(y[q], g4[q].w, acc.w);
            }
        }
    }
    float dSacc = head_reduce<1>(dS_l, gm);
    if (lid >= 0) {
        float sums[1] = {dSacc};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<1, 1, true>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, acc, sums)) return;
        dSacc = sums[0];
    }
    float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
    if (a_n) gn = ld_stream4(G + (int64_t)node * D + lane16 * 4);
    const float* ap = wtab[2] + (node >= U ? D : 0) + lane16 * 4;
    const float4 a4 = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), __ldg(ap + 3));
    st_stream4(dh + (int64_t)node * D + lane16 * 4,
               make_float4(gn.x + acc.x + dSacc * a4.x, gn.y + acc.y + dSacc * a4.y, gn.z + acc.z + dSacc * a4.z, gn.w + acc.w + dSacc * a4.w));
    if (lane16 == 0) dS[node] = dSacc;
}

}  // namespace ngacf

using namespace ngacf;

extern "C" int ngacf_mark_active(int32_t* stamp, const int64_t* users, const int64_t* items, int32_t B, int32_t U, int32_t val,
                                 const int64_t* val_dev, int32_t* task_count, void* stream) {
    NGACF_REQUIRE(stamp && users && items && B >= 0 && U >= 0, "mark_active: bad argument");
    const int blocks = B > 0 ? ceil_div(B, 256) : 1;
    mark_active_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(stamp, users, items, B, U, val, val_dev, task_count);
    return check_launch("mark_active");
}

extern "C" int ngacf_active_plan(const int32_t* stamp, int32_t val, const int64_t* val_dev, const int32_t* tasks, int32_t T, int32_t T_users,
                                 const int32_t* adj_idx, int64_t n_adj, int32_t* task_list, int32_t* task_count, uint32_t* edge_bits,
                                 void* stream) {
    NGACF_REQUIRE(stamp && tasks && adj_idx && task_list && task_count && edge_bits && T >= 0 && T_users >= 0 && T_users <= T && n_adj >= 0,
                  "active_plan: bad argument");
    const int task_blocks = ceil_div(T, 256), bit_blocks = ceil_div(n_adj, 256);
    if (task_blocks + bit_blocks == 0) return NGACF_OK;
    active_plan_kernel<<<task_blocks + bit_blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int4*>(tasks), T, T_users, task_blocks, adj_idx,
                                                                                   n_adj, stamp, val, val_dev, task_list, task_count, edge_bits);
    return check_launch("active_plan");
}

extern "C" int ngacf_stage_bwd_prep_active(const int32_t* tasks, int32_t T, const int32_t* task_list, const int32_t* task_count, const float* G,
                                           const float* Z, const float* h, const float* norm, int32_t H, float* Ghat, float* dN, void* stream) {
    NGACF_REQUIRE(tasks && task_list && task_count && G && Z && h && norm && Ghat && dN && T >= 0, "stage_bwd_prep_active: null argument");
    if (H != 1) {
        set_error("stage_bwd_prep_active: the pruned pass exists for the single-head output stage only (H = %d)", H);
        return NGACF_ERR_UNSUPPORTED;
    }
    if (T == 0) return NGACF_OK;
    int blocks = ceil_div((int64_t)T * 16, 256);
    if (blocks > 148 * 4) blocks = 148 * 4;
    prep_list_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int4*>(tasks), T, task_list, task_count, G, Z, h, norm, Ghat, dN);
    return check_launch("stage_bwd_prep_active");
}

extern "C" int ngacf_stage_bwd_edges_active(int32_t mode, const int32_t* tasks, int32_t T_begin, int32_t T_end, const int32_t* task_list,
                                            const int32_t* task_count, const int32_t* adj_ptr, const int32_t* adj_idx, const int32_t* adj_eid,
                                            const int32_t* long_first_slot, int32_t* long_counter, float* scratch, const float* G,
                                            const float* Ghat, const float* dN, const float* h, const float* s, int32_t H,
                                            const uint8_t* edgemask, float scale, const float* const* wtab, int32_t U, const int32_t* stamp,
                                            int32_t stamp_val, const int64_t* stamp_dev, const uint32_t* edge_bits, float* ds_store, float* dh,
                                            float* dS, void* stream) {
    NGACF_REQUIRE(tasks && task_list && task_count && adj_ptr && adj_idx && adj_eid && G && Ghat && dN && h && s && wtab && ds_store && dh && dS &&
                      stamp && edge_bits,
                  "stage_bwd_edges_active: null argument");
    NGACF_REQUIRE((mode == 0 || mode == 1) && T_end >= T_begin, "stage_bwd_edges_active: bad mode/range");
    if (H != 1) {
        set_error("stage_bwd_edges_active: the pruned pass exists for the single-head output stage only (H = %d)", H);
        return NGACF_ERR_UNSUPPORTED;
    }
    if (T_end == T_begin) return NGACF_OK;
    const int blocks = ceil_div((int64_t)(T_end - T_begin) * 16, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const int4* tk = reinterpret_cast<const int4*>(tasks);
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(stage_bwd_users_active_kernel<true>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(stage_bwd_users_active_kernel<false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
        cudaFuncSetAttribute(stage_bwd_items_active_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 0);
    });
    if (mode == 0) {
        // rows of the active users: one CTA per listed task (grid-stride over the list, whose length lives on the device)
        const int arow_blocks = 148 * 2;
        if (edgemask) {
            stage_bwd_users_active_kernel<true><<<blocks, 256, 0, st>>>(tk, T_begin, T_end, adj_ptr, adj_idx, adj_eid, long_first_slot, long_counter,
                                                                        scratch, G, Ghat, dN, h, s, edgemask, scale, wtab, U, stamp, stamp_val,
                                                                        stamp_dev, edge_bits, reinterpret_cast<float2*>(ds_store), dh, dS);
            stage_bwd_users_arow_kernel<true><<<arow_blocks, CTA_THREADS, 0, st>>>(tk, task_list, task_count, adj_ptr, adj_idx, adj_eid,
                                                                                   long_first_slot, long_counter, scratch, G, Ghat, dN, h, s, edgemask,
                                                                                   scale, wtab, U, edge_bits, reinterpret_cast<float2*>(ds_store), dh, dS);
        } else {
            stage_bwd_users_active_kernel<false><<<blocks, 256, 0, st>>>(tk, T_begin, T_end, adj_ptr, adj_idx, adj_eid, long_first_slot, long_counter,
                                                                         scratch, G, Ghat, dN, h, s, edgemask, scale, wtab, U, stamp, stamp_val,
                                                                         stamp_dev, edge_bits, reinterpret_cast<float2*>(ds_store), dh, dS);
            stage_bwd_users_arow_kernel<false><<<arow_blocks, CTA_THREADS, 0, st>>>(tk, task_list, task_count, adj_ptr, adj_idx, adj_eid,
                                                                                    long_first_slot, long_counter, scratch, G, Ghat, dN, h, s, edgemask,
                                                                                    scale, wtab, U, edge_bits, reinterpret_cast<float2*>(ds_store), dh, dS);
        }
    } else {
        stage_bwd_items_active_kernel<<<blocks, 256, 0, st>>>(tk, T_begin, T_end, adj_ptr, adj_idx, adj_eid, long_first_slot, long_counter, scratch,
                                                              G, Ghat, wtab, U, stamp, stamp_val, stamp_dev, edge_bits,
                                                              reinterpret_cast<const float2*>(ds_store), dh, dS);
    }
    return check_launch("stage_bwd_edges_active");
}

// forward of the listed rows, H = 1: one CTA per task (csrc/propagate_fwd.cu launches the group-per-task list kernel for H = 8)
int ngacf_launch_aggregate_fwd_cta(const int32_t* tasks, int32_t T, const int32_t* task_list, const int32_t* task_count, const int32_t* adj_ptr,
                                   const int32_t* adj_idx, const int32_t* adj_eid, const int32_t* long_first_slot, int32_t* long_counter,
                                   float* scratch, const float* h, const float* s, const uint8_t* edgemask, float scale, float* Z, float* norm,
                                   cudaStream_t st) {
    int blocks = T < 148 * 8 ? T : 148 * 8;
    const int4* tk = reinterpret_cast<const int4*>(tasks);
    if (edgemask)
        aggregate_fwd_cta_kernel<true><<<blocks, CTA_THREADS, 0, st>>>(tk, T, task_list, task_count, adj_ptr, adj_idx, adj_eid, long_first_slot,
                                                                       long_counter, scratch, h, s, edgemask, scale, Z, norm);
    else
        aggregate_fwd_cta_kernel<false><<<blocks, CTA_THREADS, 0, st>>>(tk, T, task_list, task_count, adj_ptr, adj_idx, adj_eid, long_first_slot,
                                                                        long_counter, scratch, h, s, edgemask, scale, Z, norm);
    return check_launch("aggregate_fwd_active");
}
