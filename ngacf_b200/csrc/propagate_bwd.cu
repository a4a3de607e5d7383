// Backward of one SpUIGAT stage in closed form (the reference relies on autograd through
// graphattention/SPUIGACF.py:340-400).  Unified-node formulation, G = dL/dZ:
//   Ghat[n] = G[n]/norm[n]                      dN[n] = -(G[n].(Z[n]-h[n]))/norm[n]            (per head)
//   d e~(n,m) = Ghat[n].h[m] + Ghat[m].h[n]     d e = keep*scale*d e~ + dN[n] + dN[m]
//   d s = d e * (-e) * LeakyReLU'(s[n]+s[m])    dS[n] = sum_m d s(n,m)
//   dh[n] = G[n] + sum_m drop(e) Ghat[m] + dS[n] (x) a_side
//   dW = Xd^T dh,  da = sum_n dS[n] (x) h[n],  dX = dh W^T (then dropout mask and ELU' of the producer)
// User rows walk the CSR half and store d s per edge; item rows walk the CSC half and read it back
// through adj_eid, so neither side needs atomics.
#include "common.cuh"

namespace ngacf {

template <int H>
__device__ __forceinline__ int lane_head_b(int lane16) { return H == 8 ? (lane16 >> 1) : 0; }
template <int H>
__device__ __forceinline__ bool head_writer(int lane16) { return H == 8 ? ((lane16 & 1) == 0) : (lane16 == 0); }

// ------------------------------------------------------------------------------------------------
template <int H>
__global__ void __launch_bounds__(256) stage_bwd_prep_kernel(const float* __restrict__ G, const float* __restrict__ Z,
                                                             const float* __restrict__ h, const float* __restrict__ norm, int64_t N,
                                                             float* __restrict__ Ghat, float* __restrict__ dN) {
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    if (n >= N) return;
    const int head = lane_head_b<H>(lane16);
    const float4 g = ld_stream4(G + n * D + lane16 * 4);
    const float4 z = ld_stream4(Z + n * D + lane16 * 4);
    const float4 hh = ld_stream4(h + n * D + lane16 * 4);
    const float nr = __ldg(norm + n * H + head);
    const float inv = nr != 0.f ? 1.0f / nr : 0.f;
    float part = g.x * (z.x - hh.x) + g.y * (z.y - hh.y) + g.z * (z.z - hh.z) + g.w * (z.w - hh.w);
    part = head_reduce<H>(part, gm);
    st_stream4(Ghat + n * D + lane16 * 4, make_float4(g.x * inv, g.y * inv, g.z * inv, g.w * inv));
    if (head_writer<H>(lane16)) dN[n * H + head] = -part * inv;
}

// ------------------------------------------------------------------------------------------------
template <int H, int MODE, bool DROP, bool PARTIAL>
__global__ void __launch_bounds__(256, (MODE == 1 && H == 1) ? 5 : 0) stage_bwd_edges_kernel(const int4* __restrict__ tasks, int T_begin, int T_end,
                                                              const int* __restrict__ adj_ptr, const int* __restrict__ adj_idx,
                                                              const int* __restrict__ adj_eid, const int* __restrict__ long_first_slot,
                                                              int* long_counter, float* scratch, const float* __restrict__ G,
                                                              const float* __restrict__ Ghat, const float* __restrict__ dN,
                                                              const float* __restrict__ h, const float* __restrict__ s,
                                                              const uint8_t* __restrict__ edgemask, float scale,
                                                              const float* const* __restrict__ wtab, int U,
                                                              float* ds_store, float* __restrict__ dh, float* __restrict__ dS) {
    constexpr int DH = D / H;
    const int t = T_begin + ((blockIdx.x * blockDim.x + threadIdx.x) >> 4);
    if (t >= T_end) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int head = lane_head_b<H>(lane16);
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    // H = 1: the user pass stores (d s_e, e*keep) per edge and the item pass reads the pair -- no logit gather, exp or mask byte
    // there (in-box A/B: 60 -> 47 us per launch; for H = 8 the 64-byte pairs cost more than they save, profiles/r1e_ab_prefetch.txt)
    constexpr bool PAIR = (H == 1);
    constexpr bool NEED_MASK = DROP && !(PAIR && MODE == 1);
    const float sn = (PAIR && MODE == 1) ? 0.f : __ldg(s + (int64_t)node * H + head);
    const float sc = DROP ? scale : 1.f;
    float4 ghn = make_float4(0.f, 0.f, 0.f, 0.f), hn = ghn;
    float dNn = 0.f;
    if (MODE == 0) {
        ghn = ld_stream4(Ghat + (int64_t)node * D + lane16 * 4);
        hn = ld_stream4(h + (int64_t)node * D + lane16 * 4);
        dNn = __ldg(dN + (int64_t)node * H + head);
    }
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    float dSacc = 0.f;
    // software-pipelined adjacency: batch b+1's neighbour/edge ids are requested before batch b's gathers, its mask bytes
    // right after them (same scheme as aggregate_fwd_kernel)
    int m_l = 0, eid_l = 0;
    unsigned mk_l = 0xFFu;
    if (beg + lane16 < end) {
        m_l = ld_stream_i32(adj_idx + beg + lane16);
        eid_l = ld_stream_i32(adj_eid + beg + lane16);
        if (NEED_MASK) mk_l = edgemask[eid_l];
    }
    for (int base = beg; base < end; base += 16) {
        const int nidx = base + 16 + lane16;
        int m_n = 0, eid_n = 0;
        unsigned mk_n = 0xFFu;
        if (nidx < end) {
            m_n = ld_stream_i32(adj_idx + nidx);
            eid_n = ld_stream_i32(adj_eid + nidx);
        }
        const int cnt = min(16, end - base);
#pragma unroll (MODE == 0 ? 4 : 8)
        for (int j = 0; j < cnt; ++j) {
            const int m = __shfl_sync(gm, m_l, j, 16);
            const int eid = __shfl_sync(gm, eid_l, j, 16);
            if (PAIR && MODE == 1) {
                const float4 g4 = ld_gather4(Ghat + (int64_t)m * D + lane16 * 4);
                const float2 pr = __ldg(reinterpret_cast<const float2*>(ds_store) + eid);
                acc.x = fmaf(pr.y, g4.x, acc.x); acc.y = fmaf(pr.y, g4.y, acc.y);
                acc.z = fmaf(pr.y, g4.z, acc.z); acc.w = fmaf(pr.y, g4.w, acc.w);
                dSacc += pr.x;
                continue;
            }
            const float sm = __ldg(s + (int64_t)m * H + head);
            const float4 gm4 = ld_gather4(Ghat + (int64_t)m * D + lane16 * 4);
            const float x = sn + sm;
            const float e = edge_weight(x);
            float keepsc = sc;
            if (DROP) {
                const unsigned mk = __shfl_sync(gm, mk_l, j, 16);
                keepsc = ((mk >> head) & 1u) ? sc : 0.f;
            }
            const float et = e * keepsc;
            acc.x = fmaf(et, gm4.x, acc.x); acc.y = fmaf(et, gm4.y, acc.y);
            acc.z = fmaf(et, gm4.z, acc.z); acc.w = fmaf(et, gm4.w, acc.w);
            float ds;
            if (MODE == 0) {
                const float4 hm = ld_gather4(h + (int64_t)m * D + lane16 * 4);
                float part = ghn.x * hm.x + ghn.y * hm.y + ghn.z * hm.z + ghn.w * hm.w
                           + gm4.x * hn.x + gm4.y * hn.y + gm4.z * hn.z + gm4.w * hn.w;
                const float det = head_reduce<H>(part, gm);
                const float de = fmaf(det, keepsc, dNn + __ldg(dN + (int64_t)m * H + head));
                ds = de * (-e) * (x > 0.f ? 1.f : LRELU_ALPHA);
                if (head_writer<H>(lane16)) {
                    if (PAIR) reinterpret_cast<float2*>(ds_store)[eid] = make_float2(ds, et);
                    else ds_store[(int64_t)eid * H + head] = ds;
                }
            } else {
                ds = __ldg(ds_store + (int64_t)eid * H + head);
            }
            dSacc += ds;
        }
        if (NEED_MASK && nidx < end) mk_n = edgemask[eid_n];
        m_l = m_n; eid_l = eid_n; mk_l = mk_n;
    }
    if (lid >= 0) {
        float sums[1] = {dSacc};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<H, 1>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, acc, sums)) return;
        dSacc = sums[0];
    }
    if (PARTIAL) {     // multi-GPU: raw partial sums over this rank's slice of the row; reduced across ranks, then stage_bwd_finalize
        st_stream4(dh + (int64_t)node * D + lane16 * 4, acc);
        if (head_writer<H>(lane16)) dS[(int64_t)node * H + head] = dSacc;
        return;
    }
    const float4 gn = ld_stream4(G + (int64_t)node * D + lane16 * 4);
    const float* ap = wtab[2 * H + head] + (node >= U ? DH : 0) + ((lane16 * 4) % DH);
    const float4 a4 = make_float4(__ldg(ap), __ldg(ap + 1), __ldg(ap + 2), __ldg(ap + 3));
    st_stream4(dh + (int64_t)node * D + lane16 * 4,
               make_float4(gn.x + acc.x + dSacc * a4.x, gn.y + acc.y + dSacc * a4.y, gn.z + acc.z + dSacc * a4.z, gn.w + acc.w + dSacc * a4.w));
    if (head_writer<H>(lane16)) dS[(int64_t)node * H + head] = dSacc;
}

// dh[n] = G[n] + P[n] + dS[n] (x) a_side for rows whose partials (P in dh, dS) were reduced across ranks
template <int H>
__global__ void __launch_bounds__(256) stage_bwd_finalize_kernel(float* __restrict__ dh, const float* __restrict__ dS,
                                                                 const float* __restrict__ G, const float* const* __restrict__ wtab,
                                                                 int item_side, int64_t n_rows) {
    constexpr int DH = D / H;
    const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (n >= n_rows) return;
    const int lane16 = threadIdx.x & 15;
    const int head = lane_head_b<H>(lane16);
    const float d = dS[n * H + head];
    const float* ap = wtab[2 * H + head] + (item_side ? DH : 0) + ((lane16 * 4) % DH);
    const float4 p = ld_stream4(dh + n * D + lane16 * 4), g = ld_stream4(G + n * D + lane16 * 4);
    st_stream4(dh + n * D + lane16 * 4, make_float4(g.x + p.x + d * __ldg(ap), g.y + p.y + d * __ldg(ap + 1), g.z + p.z + d * __ldg(ap + 2),
                                                    g.w + p.w + d * __ldg(ap + 3)));
}

// ------------------------------------------------------------------------------------------------
// dense backward: dX = dh W^T (+ mask/ELU'), partial dW = Xd^T dh, partial da = W^T-contracted Q = Xd^T dS
// persistent blocks per side, 64-row tiles, three CTAs per SM; deterministic two-pass reduction of the partials.
//   * h = Xd W, hence da[c] = sum_r dS[r,head(c)] h[r,c] = sum_k W[k,c] Q[k,head(c)]: the h tile is never read
//   * ELU'(z) is recovered from the recomputed stage input e = ELU(z) already in shared memory (1 for e > 0, e + 1 otherwise)
//   * lane = (8 row/k groups) x (4 column groups): both GEMMs read 8 + 4 distinct 16-byte words per warp and step,
//     conflict-free with the 68-float row stride
// ------------------------------------------------------------------------------------------------
constexpr int TB_TM = 64;
constexpr int TB_XS = 68;
constexpr int TB_SMEM_FLOATS = 64 * 64 + 2 * TB_TM * TB_XS + TB_TM * 8 + 2 * TB_TM;   // W^T, dh tile, Xd tile, dS tile, mask words
constexpr size_t TB_SMEM = (size_t)TB_SMEM_FLOATS * sizeof(float);
constexpr int TB_PART = 64 * 64 + 64;   // floats per block partial
constexpr int TB_CTAS_PER_SM = 3;     // in-box A/B: 3 CTAs (85 registers, no spills) beat 4 (64 registers, spills)

template <int H>
__global__ void __launch_bounds__(256, TB_CTAS_PER_SM) transform_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dS,
                                                            const float* __restrict__ Xu,
                                                            const float* __restrict__ Xi, int apply_elu,
                                                            const uint64_t* __restrict__ featmask, float scale,
                                                            const float* const* __restrict__ wtab, int U, int I, int nb_u,
                                                            float* __restrict__ dXu, float* __restrict__ dXi, int accumulate_dx,
                                                            float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    float* WTs = smem;                       // [c][k] = W[k][c]
    float* dhs = smem + 64 * 64;             // [64][68]
    float* Xds = dhs + TB_TM * TB_XS;        // [64][68]
    float* dSs = Xds + TB_TM * TB_XS;        // [64][H]
    uint64_t* msk = reinterpret_cast<uint64_t*>(dSs + TB_TM * 8);   // [64]
    constexpr int DH = D / H;
    const bool item_side = (int)blockIdx.x >= nb_u;
    const int bs = item_side ? blockIdx.x - nb_u : blockIdx.x;
    const int nbs = item_side ? gridDim.x - nb_u : nb_u;
    const int rows_side = item_side ? I : U;
    const float* X = item_side ? Xi : Xu;
    float* dX = item_side ? dXi : dXu;
    const int64_t node_off = item_side ? U : 0;
    const int tiles = (rows_side + TB_TM - 1) / TB_TM;
    const float inv_scale = featmask ? 1.f / scale : 1.f;

    // WTs[c*64+k] = W[k][c]
    {
        const float* const* wptr = wtab + (item_side ? H : 0);
        for (int idx = threadIdx.x; idx < 64 * 16; idx += blockDim.x) {
            const int k = idx >> 4, c = (idx & 15) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(wptr[c / DH] + k * DH + (c % DH)));
            WTs[(c + 0) * 64 + k] = v.x; WTs[(c + 1) * 64 + k] = v.y; WTs[(c + 2) * 64 + k] = v.z; WTs[(c + 3) * 64 + k] = v.w;
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g8 = (warp & 1) * 8 + (lane >> 2);      // 0..15: row group of GEMM (a), k group of GEMM (b)
    const int c4 = (warp >> 1) * 4 + (lane & 3);      // 0..15: output column group (4 columns) of both GEMMs
    float accW[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) accW[i][0] = accW[i][1] = accW[i][2] = accW[i][3] = 0.f;
    const int qh = c4 & 7;
    const bool qhi = (c4 >> 3) != 0;
    float q[4] = {0.f, 0.f, 0.f, 0.f};   // H=1: Q[4*g8 + j]; H=8: q[0..1] = Q[4*g8 + 2*(c4>>3) + j][c4 & 7]

    for (int tile = bs; tile < tiles; tile += nbs) {
        const int row0 = tile * TB_TM;
        const int nrows = min(TB_TM, rows_side - row0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < TB_TM * 16; idx += 256) {
            const int r = idx >> 4, qd = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), x = v;
            if (r < nrows) {
                v = ld_stream4(dh + (node_off + row0 + r) * D + qd * 4);
                // same expression as transform_fwd: recomputes the dropped, activated stage input
                x = ld_stream4(X + (int64_t)(row0 + r) * D + qd * 4);
                if (apply_elu) { x.x = elu(x.x); x.y = elu(x.y); x.z = elu(x.z); x.w = elu(x.w); }
                if (featmask) {
                    uint32_t m = (uint32_t)(featmask[node_off + row0 + r] >> (qd * 4)) & 0xFu;
                    x.x = (m & 1u) ? x.x * scale : 0.f; x.y = (m & 2u) ? x.y * scale : 0.f;
                    x.z = (m & 4u) ? x.z * scale : 0.f; x.w = (m & 8u) ? x.w * scale : 0.f;
                }
            }
            *reinterpret_cast<float4*>(dhs + r * TB_XS + qd * 4) = v;
            *reinterpret_cast<float4*>(Xds + r * TB_XS + qd * 4) = x;
        }
        for (int idx = threadIdx.x; idx < TB_TM * H; idx += 256)
            dSs[idx] = idx < nrows * H ? __ldg(dS + (node_off + row0) * H + idx) : 0.f;
        if (threadIdx.x < TB_TM)
            msk[threadIdx.x] = (featmask && (int)threadIdx.x < nrows) ? featmask[node_off + row0 + threadIdx.x] : ~0ull;
        __syncthreads();

        // (a) dXd[r][k] = sum_c dh[r][c] * W[k][c]; thread = rows g8 + 16i, k = 4*c4 .. 4*c4+3
        {
            float acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
#pragma unroll 4
            for (int cc = 0; cc < 16; ++cc) {
                const float4 w0 = *reinterpret_cast<const float4*>(WTs + (cc * 4 + 0) * 64 + c4 * 4);
                const float4 w1 = *reinterpret_cast<const float4*>(WTs + (cc * 4 + 1) * 64 + c4 * 4);
                const float4 w2 = *reinterpret_cast<const float4*>(WTs + (cc * 4 + 2) * 64 + c4 * 4);
                const float4 w3 = *reinterpret_cast<const float4*>(WTs + (cc * 4 + 3) * 64 + c4 * 4);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 x = *reinterpret_cast<const float4*>(dhs + (g8 + 16 * i) * TB_XS + cc * 4);
                    acc[i][0] = fmaf(x.x, w0.x, acc[i][0]); acc[i][1] = fmaf(x.x, w0.y, acc[i][1]);
                    acc[i][2] = fmaf(x.x, w0.z, acc[i][2]); acc[i][3] = fmaf(x.x, w0.w, acc[i][3]);
                    acc[i][0] = fmaf(x.y, w1.x, acc[i][0]); acc[i][1] = fmaf(x.y, w1.y, acc[i][1]);
                    acc[i][2] = fmaf(x.y, w1.z, acc[i][2]); acc[i][3] = fmaf(x.y, w1.w, acc[i][3]);
                    acc[i][0] = fmaf(x.z, w2.x, acc[i][0]); acc[i][1] = fmaf(x.z, w2.y, acc[i][1]);
                    acc[i][2] = fmaf(x.z, w2.z, acc[i][2]); acc[i][3] = fmaf(x.z, w2.w, acc[i][3]);
                    acc[i][0] = fmaf(x.w, w3.x, acc[i][0]); acc[i][1] = fmaf(x.w, w3.y, acc[i][1]);
                    acc[i][2] = fmaf(x.w, w3.z, acc[i][2]); acc[i][3] = fmaf(x.w, w3.w, acc[i][3]);
                }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = g8 + 16 * i;
                if (r >= nrows) continue;
                float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
                if (featmask) {
                    const uint32_t m = (uint32_t)(msk[r] >> (c4 * 4)) & 0xFu;
                    v.x = (m & 1u) ? v.x * scale : 0.f; v.y = (m & 2u) ? v.y * scale : 0.f;
                    v.z = (m & 4u) ? v.z * scale : 0.f; v.w = (m & 8u) ? v.w * scale : 0.f;
                }
                if (apply_elu) {   // the stage input was ELU(Zprev): chain through ELU' = 1 (e > 0) or e + 1 = exp(z)
                    const float4 e = *reinterpret_cast<const float4*>(Xds + r * TB_XS + c4 * 4);
                    const float ex = e.x * inv_scale, ey = e.y * inv_scale, ez = e.z * inv_scale, ew = e.w * inv_scale;
                    v.x *= ex > 0.f ? 1.f : ex + 1.f; v.y *= ey > 0.f ? 1.f : ey + 1.f;
                    v.z *= ez > 0.f ? 1.f : ez + 1.f; v.w *= ew > 0.f ? 1.f : ew + 1.f;
                }
                float* dst = dX + (int64_t)(row0 + r) * D + c4 * 4;
                if (accumulate_dx) {
                    const float4 o = *reinterpret_cast<const float4*>(dst);
                    v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
                }
                *reinterpret_cast<float4*>(dst) = v;
            }
        }
        // (b) dW[k][c] += sum_r Xd[r][k] * dh[r][c], Q[k][hd] += sum_r Xd[r][k] * dS[r][hd]; thread = k 4*g8.., c 4*c4..
        //     (rows >= nrows are zero in all three tiles)
#pragma unroll 4
        for (int r = 0; r < TB_TM; ++r) {
            const float4 xs = *reinterpret_cast<const float4*>(Xds + r * TB_XS + g8 * 4);
            const float4 ds = *reinterpret_cast<const float4*>(dhs + r * TB_XS + c4 * 4);
            accW[0][0] = fmaf(xs.x, ds.x, accW[0][0]); accW[0][1] = fmaf(xs.x, ds.y, accW[0][1]);
            accW[0][2] = fmaf(xs.x, ds.z, accW[0][2]); accW[0][3] = fmaf(xs.x, ds.w, accW[0][3]);
            accW[1][0] = fmaf(xs.y, ds.x, accW[1][0]); accW[1][1] = fmaf(xs.y, ds.y, accW[1][1]);
            accW[1][2] = fmaf(xs.y, ds.z, accW[1][2]); accW[1][3] = fmaf(xs.y, ds.w, accW[1][3]);
            accW[2][0] = fmaf(xs.z, ds.x, accW[2][0]); accW[2][1] = fmaf(xs.z, ds.y, accW[2][1]);
            accW[2][2] = fmaf(xs.z, ds.z, accW[2][2]); accW[2][3] = fmaf(xs.z, ds.w, accW[2][3]);
            accW[3][0] = fmaf(xs.w, ds.x, accW[3][0]); accW[3][1] = fmaf(xs.w, ds.y, accW[3][1]);
            accW[3][2] = fmaf(xs.w, ds.z, accW[3][2]); accW[3][3] = fmaf(xs.w, ds.w, accW[3][3]);
            if (H == 1) {
                const float d = dSs[r];
                q[0] = fmaf(xs.x, d, q[0]); q[1] = fmaf(xs.y, d, q[1]); q[2] = fmaf(xs.z, d, q[2]); q[3] = fmaf(xs.w, d, q[3]);
            } else {
                const float d = dSs[r * 8 + qh];
                q[0] = fmaf(qhi ? xs.z : xs.x, d, q[0]);
                q[1] = fmaf(qhi ? xs.w : xs.y, d, q[1]);
            }
        }
    }
    float* part = partials + (size_t)blockIdx.x * TB_PART;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + (g8 * 4 + i) * 64 + c4 * 4) = make_float4(accW[i][0], accW[i][1], accW[i][2], accW[i][3]);
    // da partial of this block: da[c] = sum_k W[k][c] * Q[k][head(c)]
    __syncthreads();
    float* Qs = dhs;     // [64][H]
    if (H == 1) {
        if (c4 == 0) { Qs[g8 * 4 + 0] = q[0]; Qs[g8 * 4 + 1] = q[1]; Qs[g8 * 4 + 2] = q[2]; Qs[g8 * 4 + 3] = q[3]; }
    } else {
        const int k0 = g8 * 4 + (qhi ? 2 : 0);
        Qs[(k0 + 0) * 8 + qh] = q[0];
        Qs[(k0 + 1) * 8 + qh] = q[1];
    }
    __syncthreads();
    if (threadIdx.x < 64) {
        const int c = threadIdx.x, hd = c / DH;
        float a = 0.f;
#pragma unroll 8
        for (int k = 0; k < 64; ++k) a = fmaf(WTs[c * 64 + k], Qs[k * H + hd], a);
        part[64 * 64 + c] = a;
    }
}

// sums the block partials of each side and writes the per-head gradient tensors.  One CTA = 32 consecutive outputs x 8
// groups of partials (coalesced 128-byte loads, fixed summation order -> deterministic)
template <int H>
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partials, int nb_u, int nb_total,
                                                              float* const* __restrict__ gtab, int accumulate) {
    constexpr int DH = D / H;
    __shared__ float red[8][33];
    const int og = threadIdx.x & 31, pg = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + og;                 // TB_PART is a multiple of 32
    const int side = idx / TB_PART, o = idx % TB_PART;
    const int b0 = side ? nb_u : 0, b1 = side ? nb_total : nb_u;
    float sum = 0.f;
    for (int b = b0 + pg; b < b1; b += 8) sum += partials[(size_t)b * TB_PART + o];
    red[pg][og] = sum;
    __syncthreads();
    if (pg != 0 || b1 <= b0) return;      // b1 <= b0: this call covered one side only (multi-GPU row sharding)
    sum = ((red[0][og] + red[1][og]) + (red[2][og] + red[3][og])) + ((red[4][og] + red[5][og]) + (red[6][og] + red[7][og]));
    float* dst;
    if (o < 64 * 64) {
        const int k = o >> 6, c = o & 63;
        dst = gtab[side * H + c / DH] + k * DH + (c % DH);
    } else {
        const int c = o - 64 * 64;
        dst = gtab[2 * H + c / DH] + side * DH + (c % DH);
    }
    *dst = accumulate ? *dst + sum : sum;
}

}  // namespace ngacf

using namespace ngacf;

extern "C" int ngacf_stage_bwd_prep(const float* G, const float* Z, const float* h, const float* norm, int32_t H, int64_t N, float* Ghat,
                                    float* dN, void* stream) {
    NGACF_REQUIRE(G && Z && h && norm && Ghat && dN && N >= 0, "stage_bwd_prep: null argument");
    if (N == 0) return NGACF_OK;
    NGACF_REQUIRE(H == 1 || H == 8, "stage_bwd_prep: H must be 1 or 8");
    const int blocks = ceil_div(N * 16, 256);
    if (H == 8) stage_bwd_prep_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(G, Z, h, norm, N, Ghat, dN);
    else        stage_bwd_prep_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(G, Z, h, norm, N, Ghat, dN);
    return check_launch("stage_bwd_prep");
}

extern "C" int ngacf_stage_bwd_finalize(float* dh, const float* dS, const float* G, const float* const* wtab, int32_t H, int32_t item_side,
                                        int64_t n_rows, void* stream) {
    NGACF_REQUIRE(dh && dS && G && wtab && n_rows >= 0 && (H == 1 || H == 8), "stage_bwd_finalize: bad argument");
    if (n_rows == 0) return NGACF_OK;
    const int blocks = ceil_div(n_rows * 16, 256);
    if (H == 8) stage_bwd_finalize_kernel<8><<<blocks, 256, 0, (cudaStream_t)stream>>>(dh, dS, G, wtab, item_side, n_rows);
    else        stage_bwd_finalize_kernel<1><<<blocks, 256, 0, (cudaStream_t)stream>>>(dh, dS, G, wtab, item_side, n_rows);
    return check_launch("stage_bwd_finalize");
}

extern "C" int ngacf_stage_bwd_edges(int32_t mode, const int32_t* tasks, int32_t T_begin, int32_t T_end, const int32_t* adj_ptr,
                                     const int32_t* adj_idx, const int32_t* adj_eid, const int32_t* long_first_slot,
                                     int32_t* long_counter, float* scratch, const float* G, const float* Ghat, const float* dN,
                                     const float* h, const float* s, int32_t H, const uint8_t* edgemask, float scale,
                                     const float* const* wtab, int32_t U, float* ds_store, float* dh, float* dS, int32_t partial,
                                     void* stream) {
    NGACF_REQUIRE(tasks && adj_ptr && adj_idx && adj_eid && G && Ghat && dN && h && s && wtab && ds_store && dh && dS,
                  "stage_bwd_edges: null argument");
    NGACF_REQUIRE((mode == 0 || mode == 1) && (H == 1 || H == 8) && T_end >= T_begin, "stage_bwd_edges: bad mode/H/range");
    if (T_end == T_begin) return NGACF_OK;
    const int blocks = ceil_div((int64_t)(T_end - T_begin) * 16, 256);
    cudaStream_t st = (cudaStream_t)stream;
    const int4* tk = reinterpret_cast<const int4*>(tasks);
    static PerDeviceOnce once;
    once.run([] {
#define CARVE(HH, MM, DR) cudaFuncSetAttribute(stage_bwd_edges_kernel<HH, MM, DR, false>, cudaFuncAttributePreferredSharedMemoryCarveout, 0)
        CARVE(8, 0, true); CARVE(8, 0, false); CARVE(8, 1, true); CARVE(8, 1, false);
        CARVE(1, 0, true); CARVE(1, 0, false); CARVE(1, 1, true); CARVE(1, 1, false);
#undef CARVE
    });
#define LAUNCH(HH, MM, DR, PA)                                                                                                              \
    stage_bwd_edges_kernel<HH, MM, DR, PA><<<blocks, 256, 0, st>>>(tk, T_begin, T_end, adj_ptr, adj_idx, adj_eid, long_first_slot, long_counter, \
                                                                   scratch, G, Ghat, dN, h, s, edgemask, scale, wtab, U, ds_store, dh, dS)
    const bool dr = edgemask != nullptr;
    NGACF_REQUIRE(!(partial && mode == 0), "stage_bwd_edges: partial sums are produced by the item pass (mode 1) only");
    if (H == 8) {
        if (mode == 0) { if (dr) LAUNCH(8, 0, true, false); else LAUNCH(8, 0, false, false); }
        else if (partial) { if (dr) LAUNCH(8, 1, true, true); else LAUNCH(8, 1, false, true); }
        else           { if (dr) LAUNCH(8, 1, true, false); else LAUNCH(8, 1, false, false); }
    } else {
        if (mode == 0) { if (dr) LAUNCH(1, 0, true, false); else LAUNCH(1, 0, false, false); }
        else if (partial) { if (dr) LAUNCH(1, 1, true, true); else LAUNCH(1, 1, false, true); }
        else           { if (dr) LAUNCH(1, 1, true, false); else LAUNCH(1, 1, false, false); }
    }
#undef LAUNCH
    return check_launch("stage_bwd_edges");
}

static void transform_bwd_grid(int U, int I, int* nb_u, int* nb_i) {
    const int tiles_u = ceil_div(U, TB_TM), tiles_i = ceil_div(I, TB_TM);
    const int budget = TB_CTAS_PER_SM * 148;   // resident CTAs (54 KB of shared memory each; register-limited)
    int bu = (int)((int64_t)budget * tiles_u / (tiles_u + tiles_i > 0 ? tiles_u + tiles_i : 1));
    if (bu < 1) bu = 1;
    if (bu > tiles_u) bu = tiles_u;      // 0 when this call has no user rows
    int bi = budget - bu;
    if (bi < 1) bi = 1;
    if (bi > tiles_i) bi = tiles_i;
    *nb_u = bu;
    *nb_i = bi;
}

extern "C" size_t ngacf_transform_bwd_workspace_bytes(int32_t U, int32_t I) {
    int bu, bi;
    transform_bwd_grid(U, I, &bu, &bi);
    return (size_t)(bu + bi) * TB_PART * sizeof(float);
}

extern "C" int ngacf_transform_bwd(const float* dh, const float* dS, const float* h, const float* Xu, const float* Xi, int32_t apply_elu,
                                   const uint64_t* featmask, float scale, const float* const* wtab, float* const* gtab, int32_t H,
                                   int32_t U, int32_t I, float* dXu, float* dXi, int32_t accumulate_dx, int32_t accumulate_dw,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    NGACF_REQUIRE(dh && dS && wtab && gtab && workspace && U >= 0 && I >= 0 && (U == 0 || (Xu && dXu)) && (I == 0 || (Xi && dXi)),
                  "transform_bwd: null argument");
    if (U + I == 0) return NGACF_OK;
    NGACF_REQUIRE(H == 1 || H == 8, "transform_bwd: H must be 1 or 8");
    if (workspace_bytes < ngacf_transform_bwd_workspace_bytes(U, I)) {
        set_error("transform_bwd: workspace too small");
        return NGACF_ERR_WORKSPACE;
    }
    int bu, bi;
    if (dense_on_tensor_cores()) {
        cudaStream_t st = (cudaStream_t)stream;
        transform_bwd_tc(dh, dS, Xu, Xi, apply_elu, featmask, scale, wtab, H, U, I, dXu, dXi, accumulate_dx, (float*)workspace, &bu, &bi, st);
        if (H == 8) reduce_partials_kernel<8><<<2 * TB_PART / 32, 256, 0, st>>>((float*)workspace, bu, bu + bi, gtab, accumulate_dw);
        else        reduce_partials_kernel<1><<<2 * TB_PART / 32, 256, 0, st>>>((float*)workspace, bu, bu + bi, gtab, accumulate_dw);
        return check_launch("transform_bwd(tc)");
    }
    transform_bwd_grid(U, I, &bu, &bi);
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(transform_bwd_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM);
        cudaFuncSetAttribute(transform_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TB_SMEM);
    });
    cudaStream_t st = (cudaStream_t)stream;
    float* partials = (float*)workspace;
    if (H == 8) {
        transform_bwd_kernel<8><<<bu + bi, 256, TB_SMEM, st>>>(dh, dS, Xu, Xi, apply_elu, featmask, scale, wtab, U, I, bu, dXu, dXi, accumulate_dx, partials);
        reduce_partials_kernel<8><<<2 * TB_PART / 32, 256, 0, st>>>(partials, bu, bu + bi, gtab, accumulate_dw);
    } else {
        transform_bwd_kernel<1><<<bu + bi, 256, TB_SMEM, st>>>(dh, dS, Xu, Xi, apply_elu, featmask, scale, wtab, U, I, bu, dXu, dXi, accumulate_dx, partials);
        reduce_partials_kernel<1><<<2 * TB_PART / 32, 256, 0, st>>>(partials, bu, bu + bi, gtab, accumulate_dw);
    }
    return check_launch("transform_bwd");
}
