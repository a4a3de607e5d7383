// Laplacian propagation of SPUIGAGPCF: GPLayer.forward = torch.sparse.mm(laplacianMat + selfLoop, features)
// (graphattention/SPUIGACF.py:174-185), the normalised adjacency being buildLaplacianMat's 'norm_adj' / 'mean_adj'
// (data/loadGowalla.py:197-227).  The Laplacian of a bipartite interaction graph has the sparsity pattern of the unified node
// adjacency the attention kernels already walk, plus a diagonal, and it is symmetric -- so it is stored as one value per
// UNDIRECTED edge (indexed by CSR edge id) and one per node, the product runs over the same degree-bucketed task list with the
// same deterministic long-row combine, and its transpose (the backward) is the same call.
//   Y[n] = diag[n] * X[n] + sum_{m in adj(n)} val[eid(n,m)] * X[m]
#include "common.cuh"

namespace ngacf {

__global__ void __launch_bounds__(256) spmm_sym_kernel(const int4* __restrict__ tasks, int T, const int* __restrict__ adj_ptr,
                                                       const int* __restrict__ adj_idx, const int* __restrict__ adj_eid,
                                                       const int* __restrict__ long_first_slot, int* long_counter, float* scratch,
                                                       const float* __restrict__ val, const float* __restrict__ diag,
                                                       const float* __restrict__ X, float* __restrict__ Y) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (t >= T) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int4 tk = __ldg(tasks + t);
    const int node = tk.x, beg = tk.y, end = tk.z, lid = tk.w;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 16) {
        const int idx = base + lane16;
        int m_l = 0;
        float w_l = 0.f;
        if (idx < end) {
            m_l = ld_stream_i32(adj_idx + idx);
            w_l = __ldg(val + ld_stream_i32(adj_eid + idx));
        }
        const int cnt = min(16, end - base);
#pragma unroll 8
        for (int j = 0; j < cnt; ++j) {
            const int m = __shfl_sync(gm, m_l, j, 16);
            const float w = __shfl_sync(gm, w_l, j, 16);
            const float4 x = ld_gather4(X + (int64_t)m * D + lane16 * 4);
            acc.x = fmaf(w, x.x, acc.x); acc.y = fmaf(w, x.y, acc.y);
            acc.z = fmaf(w, x.z, acc.z); acc.w = fmaf(w, x.w, acc.w);
        }
    }
    if (lid >= 0) {
        float sums[1] = {0.f};
        const int chunk = (beg - __ldg(adj_ptr + node)) / CHUNK;
        if (!long_row_combine<1, 1>(lid, chunk, long_first_slot, long_counter, scratch, lane16, gm, acc, sums)) return;
    }
    const float dg = __ldg(diag + node);
    const float4 xn = ld_stream4(X + (int64_t)node * D + lane16 * 4);
    st_stream4(Y + (int64_t)node * D + lane16 * 4,
               make_float4(fmaf(dg, xn.x, acc.x), fmaf(dg, xn.y, acc.y), fmaf(dg, xn.z, acc.z), fmaf(dg, xn.w, acc.w)));
}

}  // namespace ngacf

using namespace ngacf;

extern "C" int ngacf_spmm_sym(const int32_t* tasks, int32_t T, const int32_t* adj_ptr, const int32_t* adj_idx, const int32_t* adj_eid,
                              const int32_t* long_first_slot, int32_t* long_counter, float* scratch, const float* val, const float* diag,
                              const float* X, float* Y, void* stream) {
    NGACF_REQUIRE(tasks && adj_ptr && adj_idx && adj_eid && val && diag && X && Y && T > 0, "spmm_sym: null/empty argument");
    NGACF_REQUIRE(X != Y, "spmm_sym: in-place product is not supported");
    spmm_sym_kernel<<<ceil_div((int64_t)T * 16, 256), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const int4*>(tasks), T, adj_ptr, adj_idx,
                                                                                       adj_eid, long_first_slot, long_counter, scratch, val, diag, X, Y);
    return check_launch("spmm_sym");
}
