// SpGraphAttentionLayer (graphattention/SPGA.py:358-421 of the reference): the single-table cousin of the
// bipartite layer.  One W for every node, logit of the DIRECTED edge n -> m = a_src . h[n] + a_dst . h[m] (not symmetric),
// row-normalised only, no residual:
//     e_nm = exp(-LeakyReLU(p[n] + q[m])),  R[n] = sum_m e_nm,  out[n] = sum_m drop(e_nm) h[m] / R[n]        (:397-413)
// The adjacency is the N x N matrix of the user-item graph (SPGACF / SPGAMGP pass it as `mask`): symmetric bipartite pattern, with
// or without the diagonal -- i.e. the unified adjacency ngacf_graph_build already emits, plus an optional self edge per node.
// Adjacency POSITION p (0 .. 2E-1) identifies a directed edge; its reverse sits at rev[p].  Masks / stored pairs are indexed by
// position, the self edge of node n by 2E + n.
// Backward (closed form; the reference uses autograd through SpecialSpmmFunction, SPGA.py:423-444), G = dL/dout:
//     Ghat[n] = G[n]/R[n],  dN[n] = -(G[n].out[n])/R[n]
//     row pass  (edges n -> m): d e~ = Ghat[n].h[m];  d e = keep*scale*d e~ + dN[n];  d s = d e (-e) LeakyReLU'(x);  dP[n] = sum_m d s
//               stores (d s, e*keep*scale) per position
//     col pass  (edges n -> m seen from m through rev): dh[m] = sum_n e~_nm Ghat[n] + dP[m] a_src + dQ[m] a_dst,  dQ[m] = sum_n d s_nm
//     da_src = sum_n dP[n] h[n],  da_dst = sum_n dQ[n] h[n]  (node_logits_bwd);  dW, dX: ngacf_transform_bwd on dh.
#include "common.cuh"

namespace ngacf {

template <int H>
__device__ __forceinline__ int sp_head(int lane16) { return H == 8 ? (lane16 >> 1) : 0; }
template <int H>
__device__ __forceinline__ bool sp_writer(int lane16) { return H == 8 ? ((lane16 & 1) == 0) : (lane16 == 0); }

// p[n,k] = a_k[:DH] . h[n, head k],  q[n,k] = a_k[DH:] . h[n, head k]
template <int H>
__global__ void __launch_bounds__(256) node_logits_kernel(const float* __restrict__ h, const float* const* __restrict__ wtab, int64_t N,
                                                          float* __restrict__ p, float* __restrict__ q) {
    constexpr int DH = D / H;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (n >= N) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int head = sp_head<H>(lane16);
    const float* a = wtab[2 * H + head];
    const int c = (lane16 * 4) % DH;
    const float4 hv = ld_stream4(h + n * D + lane16 * 4);
    float ps = hv.x * __ldg(a + c) + hv.y * __ldg(a + c + 1) + hv.z * __ldg(a + c + 2) + hv.w * __ldg(a + c + 3);
    float qs = hv.x * __ldg(a + DH + c) + hv.y * __ldg(a + DH + c + 1) + hv.z * __ldg(a + DH + c + 2) + hv.w * __ldg(a + DH + c + 3);
    ps = head_reduce<H>(ps, gm);
    qs = head_reduce<H>(qs, gm);
    if (sp_writer<H>(lane16)) { p[n * H + head] = ps; q[n * H + head] = qs; }
}

// per-block partial sums of da: part[b][0:64] = sum_n dP[n,head(c)] h[n,c], part[b][64:128] = the same with dQ (fixed order)
template <int H>
__global__ void __launch_bounds__(256) node_logits_bwd_kernel(const float* __restrict__ h, const float* __restrict__ dP, const float* __restrict__ dQ,
                                                              int64_t N, float* __restrict__ part) {
    __shared__ float red[2][16][D];
    const int lane16 = threadIdx.x & 15, grp = threadIdx.x >> 4;
    const int head = sp_head<H>(lane16);
    float4 as = make_float4(0.f, 0.f, 0.f, 0.f), ad = as;
    for (int64_t n = (int64_t)blockIdx.x * 16 + grp; n < N; n += (int64_t)gridDim.x * 16) {
        const float4 hv = ld_stream4(h + n * D + lane16 * 4);
        const float dp = __ldg(dP + n * H + head), dq = __ldg(dQ + n * H + head);
        as.x = fmaf(dp, hv.x, as.x); as.y = fmaf(dp, hv.y, as.y); as.z = fmaf(dp, hv.z, as.z); as.w = fmaf(dp, hv.w, as.w);
        ad.x = fmaf(dq, hv.x, ad.x); ad.y = fmaf(dq, hv.y, ad.y); ad.z = fmaf(dq, hv.z, ad.z); ad.w = fmaf(dq, hv.w, ad.w);
    }
    *reinterpret_cast<float4*>(&red[0][grp][lane16 * 4]) = as;
    *reinterpret_cast<float4*>(&red[1][grp][lane16 * 4]) = ad;
    __syncthreads();
    if (threadIdx.x < 2 * D) {
        const int w = threadIdx.x / D, c = threadIdx.x % D;
        float sum = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) sum += red[w][g][c];
        part[(size_t)blockIdx.x * 2 * D + w * D + c] = sum;
    }
}

template <int H, bool DROP>
__global__ void __launch_bounds__(256) spgat_aggregate_fwd_kernel(const int4* __restrict__ tasks, int T, const int* __restrict__ adj_ptr,
                                                                  const int* __restrict__ adj_idx, const int* __restrict__ long_first_slot,
                                                                  int* long_counter, float* scratch, const float* __restrict__ h,
                                                                  const float* __restrict__ p, const float* __restrict__ q,
                                                                  const uint8_t* __restrict__ emask, float scale, int self_loops, int n_adj,
                                                                  float* __restrict__ Z, float* __restrict__ norm) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (t >= T) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int head = sp_head<H>(lane16);
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    const float pn = __ldg(p + (int64_t)node * H + head);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float rs = 0.f;
    for (int base = beg; base < end; base += 16) {
        const int idx = base + lane16;
        int m_l = 0;
        unsigned mk_l = 0xFFu;
        if (idx < end) {
            m_l = ld_stream_i32(adj_idx + idx);
            if (DROP) mk_l = emask[idx];
        }
        const int cnt = min(16, end - base);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const int m = __shfl_sync(gm, m_l, j, 16);
            const float qm = __ldg(q + (int64_t)m * H + head);
            const float4 hm = ld_gather4(h + (int64_t)m * D + lane16 * 4);
            const float w = edge_weight(pn + qm);
            rs += w;
            float wd = w;
            if (DROP) {
                const unsigned mk = __shfl_sync(gm, mk_l, j, 16);
                wd = ((mk >> head) & 1u) ? w * scale : 0.f;
            }
            acc.x = fmaf(wd, hm.x, acc.x); acc.y = fmaf(wd, hm.y, acc.y);
            acc.z = fmaf(wd, hm.z, acc.z); acc.w = fmaf(wd, hm.w, acc.w);
        }
    }
    if (lid >= 0) {
        float sums[1] = {rs};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<H, 1>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, acc, sums)) return;
        rs = sums[0];
    }
    if (self_loops) {     // the edge n -> n (mask / pair index 2E + n)
        const float4 hn = ld_stream4(h + (int64_t)node * D + lane16 * 4);
        const float w = edge_weight(pn + __ldg(q + (int64_t)node * H + head));
        rs += w;
        float wd = w;
        if (DROP) wd = ((emask[(int64_t)n_adj + node] >> head) & 1u) ? w * scale : 0.f;
        acc.x = fmaf(wd, hn.x, acc.x); acc.y = fmaf(wd, hn.y, acc.y); acc.z = fmaf(wd, hn.z, acc.z); acc.w = fmaf(wd, hn.w, acc.w);
    }
    const float inv = rs != 0.f ? 1.0f / rs : 0.f;      // a node without any edge: the reference divides 0 by 0; 0 here
    st_stream4(Z + (int64_t)node * D + lane16 * 4, make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv));
    if (sp_writer<H>(lane16)) norm[(int64_t)node * H + head] = rs;
}

// row pass: d s of every outgoing edge, dP[n]; Ghat[n] for the column pass
template <int H, bool DROP>
__global__ void __launch_bounds__(256) spgat_bwd_rows_kernel(const int4* __restrict__ tasks, int T, const int* __restrict__ adj_ptr,
                                                             const int* __restrict__ adj_idx, const int* __restrict__ long_first_slot,
                                                             int* long_counter, float* scratch, const float* __restrict__ G,
                                                             const float* __restrict__ Z, const float* __restrict__ norm,
                                                             const float* __restrict__ h, const float* __restrict__ p, const float* __restrict__ q,
                                                             const uint8_t* __restrict__ emask, float scale, int self_loops, int n_adj,
                                                             float* __restrict__ Ghat, float2* __restrict__ pairs, float* __restrict__ dP) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (t >= T) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int head = sp_head<H>(lane16);
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    const float sc = DROP ? scale : 1.f;
    const float nr = __ldg(norm + (int64_t)node * H + head);
    const float inv = nr != 0.f ? 1.0f / nr : 0.f;
    const float4 g = ld_stream4(G + (int64_t)node * D + lane16 * 4);
    const float4 z = ld_stream4(Z + (int64_t)node * D + lane16 * 4);
    const float4 ghn = make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv);
    const float dNn = -head_reduce<H>(g.x * z.x + g.y * z.y + g.z * z.z + g.w * z.w, gm) * inv;
    const float pn = __ldg(p + (int64_t)node * H + head);
    float dPacc = 0.f;
    for (int base = beg; base < end; base += 16) {
        const int idx = base + lane16;
        int m_l = 0;
        unsigned mk_l = 0xFFu;
        if (idx < end) {
            m_l = ld_stream_i32(adj_idx + idx);
            if (DROP) mk_l = emask[idx];
        }
        const int cnt = min(16, end - base);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const int m = __shfl_sync(gm, m_l, j, 16);
            const float qm = __ldg(q + (int64_t)m * H + head);
            const float4 hm = ld_gather4(h + (int64_t)m * D + lane16 * 4);
            const float x = pn + qm;
            const float e = edge_weight(x);
            float keepsc = sc;
            if (DROP) {
                const unsigned mk = __shfl_sync(gm, mk_l, j, 16);
                keepsc = ((mk >> head) & 1u) ? sc : 0.f;
            }
            const float det = head_reduce<H>(ghn.x * hm.x + ghn.y * hm.y + ghn.z * hm.z + ghn.w * hm.w, gm);
            const float ds = fmaf(det, keepsc, dNn) * (-e) * (x > 0.f ? 1.f : LRELU_ALPHA);
            if (sp_writer<H>(lane16)) pairs[(int64_t)(base + j) * H + head] = make_float2(ds, e * keepsc);
            dPacc += ds;
        }
    }
    float4 none = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lid >= 0) {
        float sums[1] = {dPacc};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<H, 1>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, none, sums)) return;
        dPacc = sums[0];
    }
    if (self_loops) {
        const float4 hn = ld_stream4(h + (int64_t)node * D + lane16 * 4);
        const float x = pn + __ldg(q + (int64_t)node * H + head);
        const float e = edge_weight(x);
        float keepsc = sc;
        if (DROP) keepsc = ((emask[(int64_t)n_adj + node] >> head) & 1u) ? sc : 0.f;
        const float det = head_reduce<H>(ghn.x * hn.x + ghn.y * hn.y + ghn.z * hn.z + ghn.w * hn.w, gm);
        const float ds = fmaf(det, keepsc, dNn) * (-e) * (x > 0.f ? 1.f : LRELU_ALPHA);
        if (sp_writer<H>(lane16)) pairs[((int64_t)n_adj + node) * H + head] = make_float2(ds, e * keepsc);
        dPacc += ds;
    }
    st_stream4(Ghat + (int64_t)node * D + lane16 * 4, ghn);
    if (sp_writer<H>(lane16)) dP[(int64_t)node * H + head] = dPacc;
}

// column pass: incoming edges of m are the reverses of its outgoing ones (symmetric pattern)
template <int H>
__global__ void __launch_bounds__(256) spgat_bwd_cols_kernel(const int4* __restrict__ tasks, int T, const int* __restrict__ adj_ptr,
                                                             const int* __restrict__ adj_idx, const int* __restrict__ rev,
                                                             const int* __restrict__ long_first_slot, int* long_counter, float* scratch,
                                                             const float* __restrict__ Ghat, const float2* __restrict__ pairs,
                                                             const float* __restrict__ dP, const float* const* __restrict__ wtab,
                                                             int self_loops, int n_adj, float* __restrict__ dh, float* __restrict__ dQ) {
    constexpr int DH = D / H;
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (t >= T) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int head = sp_head<H>(lane16);
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float dQacc = 0.f;
    for (int base = beg; base < end; base += 16) {
        const int idx = base + lane16;
        int n_l = 0, r_l = 0;
        if (idx < end) {
            n_l = ld_stream_i32(adj_idx + idx);
            r_l = ld_stream_i32(rev + idx);
        }
        const int cnt = min(16, end - base);
#pragma unroll 4
        for (int j = 0; j < cnt; ++j) {
            const int n = __shfl_sync(gm, n_l, j, 16);
            const int r = __shfl_sync(gm, r_l, j, 16);
            const float2 pr = __ldg(pairs + (int64_t)r * H + head);
            const float4 g4 = ld_gather4(Ghat + (int64_t)n * D + lane16 * 4);
            acc.x = fmaf(pr.y, g4.x, acc.x); acc.y = fmaf(pr.y, g4.y, acc.y);
            acc.z = fmaf(pr.y, g4.z, acc.z); acc.w = fmaf(pr.y, g4.w, acc.w);
            dQacc += pr.x;
        }
    }
    if (lid >= 0) {
        float sums[1] = {dQacc};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<H, 1>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, acc, sums)) return;
        dQacc = sums[0];
    }
    if (self_loops) {
        const float2 pr = __ldg(pairs + ((int64_t)n_adj + node) * H + head);
        const float4 g4 = ld_stream4(Ghat + (int64_t)node * D + lane16 * 4);
        acc.x = fmaf(pr.y, g4.x, acc.x); acc.y = fmaf(pr.y, g4.y, acc.y); acc.z = fmaf(pr.y, g4.z, acc.z); acc.w = fmaf(pr.y, g4.w, acc.w);
        dQacc += pr.x;
    }
    const float dp = __ldg(dP + (int64_t)node * H + head);
    const float* a = wtab[2 * H + head];
    const int c = (lane16 * 4) % DH;
    st_stream4(dh + (int64_t)node * D + lane16 * 4,
               make_float4(acc.x + dp * __ldg(a + c) + dQacc * __ldg(a + DH + c), acc.y + dp * __ldg(a + c + 1) + dQacc * __ldg(a + DH + c + 1),
                           acc.z + dp * __ldg(a + c + 2) + dQacc * __ldg(a + DH + c + 2), acc.w + dp * __ldg(a + c + 3) + dQacc * __ldg(a + DH + c + 3)));
    if (sp_writer<H>(lane16)) dQ[(int64_t)node * H + head] = dQacc;
}

}  // namespace ngacf

using namespace ngacf;

extern "C" int ngacf_node_logits(const float* h, const float* const* wtab, int32_t H, int64_t N, float* p, float* q, void* stream) {
    NGACF_REQUIRE(h && wtab && p && q && N >= 0 && (H == 1 || H == 8), "node_logits: bad argument");
    if (N == 0) return NGACF_OK;
    const int blocks = ceil_div(N * 16, 256);
    if (H == 8) node_logits_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(h, wtab, N, p, q);
    else        node_logits_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(h, wtab, N, p, q);
    return check_launch("node_logits");
}

extern "C" int ngacf_node_logits_bwd(const float* h, const float* dP, const float* dQ, int32_t H, int64_t N, float* partials, int32_t n_blocks,
                                     void* stream) {
    NGACF_REQUIRE(h && dP && dQ && partials && N >= 0 && n_blocks > 0 && (H == 1 || H == 8), "node_logits_bwd: bad argument");
    if (H == 8) node_logits_bwd_kernel<8><<<n_blocks, 256, 0, (cudaStream_t)stream>>>(h, dP, dQ, N, partials);
    else        node_logits_bwd_kernel<1><<<n_blocks, 256, 0, (cudaStream_t)stream>>>(h, dP, dQ, N, partials);
    return check_launch("node_logits_bwd");
}

extern "C" int ngacf_spgat_aggregate_fwd(const int32_t* tasks, int32_t T, const int32_t* adj_ptr, const int32_t* adj_idx,
                                         const int32_t* long_first_slot, int32_t* long_counter, float* scratch, const float* h, const float* p,
                                         const float* q, int32_t H, const uint8_t* emask, float scale, int32_t self_loops, int32_t n_adj,
                                         float* Z, float* norm, void* stream) {
    NGACF_REQUIRE(tasks && adj_ptr && adj_idx && h && p && q && Z && norm && T > 0 && (H == 1 || H == 8), "spgat_aggregate_fwd: bad argument");
    const int blocks = ceil_div((int64_t)T * 16, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const int4* tk = reinterpret_cast<const int4*>(tasks);
#define LAUNCH(HH, DR) spgat_aggregate_fwd_kernel<HH, DR><<<blocks, 256, 0, st>>>(tk, T, adj_ptr, adj_idx, long_first_slot, long_counter, scratch, h, p, q, emask, scale, self_loops, n_adj, Z, norm)
    if (H == 8) { if (emask) LAUNCH(8, true); else LAUNCH(8, false); }
    else        { if (emask) LAUNCH(1, true); else LAUNCH(1, false); }
#undef LAUNCH
    return check_launch("spgat_aggregate_fwd");
}

extern "C" int ngacf_spgat_bwd(const int32_t* tasks, int32_t T, const int32_t* adj_ptr, const int32_t* adj_idx, const int32_t* rev,
                               const int32_t* long_first_slot, int32_t* long_counter, float* scratch, const float* G, const float* Z,
                               const float* norm, const float* h, const float* p, const float* q, int32_t H, const uint8_t* emask, float scale,
                               int32_t self_loops, int32_t n_adj, const float* const* wtab, float* Ghat, float* pairs, float* dP, float* dQ,
                               float* dh, void* stream) {
    NGACF_REQUIRE(tasks && adj_ptr && adj_idx && rev && G && Z && norm && h && p && q && wtab && Ghat && pairs && dP && dQ && dh && T > 0 &&
                      (H == 1 || H == 8),
                  "spgat_bwd: bad argument");
    const int blocks = ceil_div((int64_t)T * 16, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const int4* tk = reinterpret_cast<const int4*>(tasks);
    float2* pr = reinterpret_cast<float2*>(pairs);
#define ROWS(HH, DR) spgat_bwd_rows_kernel<HH, DR><<<blocks, 256, 0, st>>>(tk, T, adj_ptr, adj_idx, long_first_slot, long_counter, scratch, G, Z, norm, h, p, q, emask, scale, self_loops, n_adj, Ghat, pr, dP)
    if (H == 8) { if (emask) ROWS(8, true); else ROWS(8, false); }
    else        { if (emask) ROWS(1, true); else ROWS(1, false); }
#undef ROWS
    if (H == 8) spgat_bwd_cols_kernel<8><<<blocks, 256, 0, st>>>(tk, T, adj_ptr, adj_idx, rev, long_first_slot, long_counter, scratch, Ghat, pr, dP, wtab, self_loops, n_adj, dh, dQ);
    else        spgat_bwd_cols_kernel<1><<<blocks, 256, 0, st>>>(tk, T, adj_ptr, adj_idx, rev, long_first_slot, long_counter, scratch, Ghat, pr, dP, wtab, self_loops, n_adj, dh, dQ);
    return check_launch("spgat_bwd");
}
