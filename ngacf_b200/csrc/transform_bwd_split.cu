// Dense backward of a stage, split in two kernels so that only dX sits on the backward chain:
//   transform_bwd_dx : dX = dh W^T, then dropout mask / ELU' of the producer (-> G of the previous stage, or
//                      the embedding gradients).  Same 128-row tile / 8x4 register tile as transform_fwd.
//   transform_bwd_dw : per-CTA partial dW = Xd^T dh and V = Xd^T dS over persistent row tiles; a deterministic
//                      second pass sums the partials and forms da = W^T-contraction of V (da[c] = sum_i W[i][c] V[i][head(c)],
//                      which equals sum_n dS[n,head] h[n,c] because h = Xd W) -- so h is never re-read.
// The weight gradients are only needed by Adam, so transform_bwd_dw runs on a separate stream, overlapped with the
// gather kernels of the next stage (FFMA-bound vs memory-bound).
// Reference: autograd of torch.mm at graphattention/SPUIGACF.py:356-357 and of the logit :359-361.
#include "common.cuh"

namespace ngacf {

constexpr int DX_TM = 128;
constexpr int DX_XS = 68;
constexpr size_t DX_SMEM = (size_t)(64 * 64 + DX_TM * DX_XS) * sizeof(float);

template <int H>
__global__ void __launch_bounds__(256, 4) transform_bwd_dx_kernel(const float* __restrict__ dh, const float* __restrict__ Xu,
                                                                  const float* __restrict__ Xi, int apply_elu,
                                                                  const uint64_t* __restrict__ featmask, float scale,
                                                                  const float* const* __restrict__ wtab, int U, int I, int tiles_u,
                                                                  float* __restrict__ dXu, float* __restrict__ dXi, int accumulate) {
    extern __shared__ __align__(16) float smem[];
    float* WTs = smem;                   // [c][k] = W[k][c]
    float* Ds = smem + 64 * 64;          // dh tile [128][68]
    constexpr int DH = D / H;
    const bool item_side = (int)blockIdx.x >= tiles_u;
    const int tile = item_side ? blockIdx.x - tiles_u : blockIdx.x;
    const int rows_side = item_side ? I : U;
    const float* X = item_side ? Xi : Xu;
    float* dX = item_side ? dXi : dXu;
    const int64_t node0 = (item_side ? (int64_t)U : 0) + (int64_t)tile * DX_TM;
    const int row0 = tile * DX_TM;
    const int nrows = min(DX_TM, rows_side - row0);
    {
        const float* const* wptr = wtab + (item_side ? H : 0);
        for (int idx = threadIdx.x; idx < 64 * 16; idx += blockDim.x) {
            const int k = idx >> 4, c = (idx & 15) * 4;
            const float4 v = __ldg(reinterpret_cast<const float4*>(wptr[c / DH] + k * DH + (c % DH)));
            WTs[(c + 0) * 64 + k] = v.x; WTs[(c + 1) * 64 + k] = v.y; WTs[(c + 2) * 64 + k] = v.z; WTs[(c + 3) * 64 + k] = v.w;
        }
    }
    for (int idx = threadIdx.x; idx < DX_TM * 16; idx += 256) {
        const int r = idx >> 4, q = idx & 15;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (r < nrows) v = ld_stream4(dh + (node0 + r) * D + q * 4);
        *reinterpret_cast<float4*>(Ds + r * DX_XS + q * 4) = v;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f; }
#pragma unroll 4
    for (int c4 = 0; c4 < 16; ++c4) {
        const float4 w0 = *reinterpret_cast<const float4*>(WTs + (c4 * 4 + 0) * 64 + tx * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(WTs + (c4 * 4 + 1) * 64 + tx * 4);
        const float4 w2 = *reinterpret_cast<const float4*>(WTs + (c4 * 4 + 2) * 64 + tx * 4);
        const float4 w3 = *reinterpret_cast<const float4*>(WTs + (c4 * 4 + 3) * 64 + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 x = *reinterpret_cast<const float4*>(Ds + (ty + 16 * i) * DX_XS + c4 * 4);
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
    for (int i = 0; i < 8; ++i) {
        const int r = ty + 16 * i;
        if (r >= nrows) continue;
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        if (featmask) {
            const uint32_t m = (uint32_t)(featmask[node0 + r] >> (tx * 4)) & 0xFu;
            v.x = (m & 1u) ? v.x * scale : 0.f; v.y = (m & 2u) ? v.y * scale : 0.f;
            v.z = (m & 4u) ? v.z * scale : 0.f; v.w = (m & 8u) ? v.w * scale : 0.f;
        }
        float* dst = dX + (int64_t)(row0 + r) * D + tx * 4;
        if (apply_elu) {   // the stage input was ELU(Zprev): chain through ELU'
            const float4 z = ld_stream4(X + (int64_t)(row0 + r) * D + tx * 4);
            v.x *= elu_grad(z.x); v.y *= elu_grad(z.y); v.z *= elu_grad(z.z); v.w *= elu_grad(z.w);
        }
        if (accumulate) {
            const float4 o = *reinterpret_cast<const float4*>(dst);
            v.x += o.x; v.y += o.y; v.z += o.z; v.w += o.w;
        }
        *reinterpret_cast<float4*>(dst) = v;
    }
}

// ------------------------------------------------------------------------------------------------
constexpr int DW_TM = 128;
constexpr int DW_XS = 68;
constexpr int DW_PART = 64 * 64 + 64 * 8;     // floats per CTA partial: dW (64x64) + V (64 x up to 8 heads)
constexpr size_t DW_SMEM = (size_t)(2 * DW_TM * DW_XS + DW_TM * 8) * sizeof(float);

template <int H>
__global__ void __launch_bounds__(256, 3) transform_bwd_dw_kernel(const float* __restrict__ dh, const float* __restrict__ dS,
                                                                  const float* __restrict__ Xu, const float* __restrict__ Xi, int apply_elu,
                                                                  const uint64_t* __restrict__ featmask, float scale, int U, int I, int nb_u,
                                                                  float* __restrict__ partials) {
    extern __shared__ __align__(16) float smem[];
    float* dhs = smem;                       // [128][68]
    float* Xds = dhs + DW_TM * DW_XS;        // [128][68]
    float* dSs = Xds + DW_TM * DW_XS;        // [128][8]
    const bool item_side = (int)blockIdx.x >= nb_u;
    const int bs = item_side ? blockIdx.x - nb_u : blockIdx.x;
    const int nbs = item_side ? gridDim.x - nb_u : nb_u;
    const int rows_side = item_side ? I : U;
    const float* X = item_side ? Xi : Xu;
    const int64_t node_off = item_side ? U : 0;
    const int tiles = (rows_side + DW_TM - 1) / DW_TM;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    float accW[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i) accW[i][0] = accW[i][1] = accW[i][2] = accW[i][3] = 0.f;
    // V[i][k]: thread -> feature i = tid & 63, heads kg*2, kg*2+1 (H=8) or head 0 by threads < 64 (H=1)
    const int vi = threadIdx.x & 63, vk = (threadIdx.x >> 6) * 2;
    float accV0 = 0.f, accV1 = 0.f;

    for (int tile = bs; tile < tiles; tile += nbs) {
        const int row0 = tile * DW_TM;
        const int nrows = min(DW_TM, rows_side - row0);
        __syncthreads();
        for (int idx = threadIdx.x; idx < DW_TM * 16; idx += 256) {
            const int r = idx >> 4, q = idx & 15;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), x = v;
            if (r < nrows) {
                v = ld_stream4(dh + (node_off + row0 + r) * D + q * 4);
                x = ld_stream4(X + (int64_t)(row0 + r) * D + q * 4);     // recompute the dropped, activated stage input
                if (apply_elu) { x.x = elu(x.x); x.y = elu(x.y); x.z = elu(x.z); x.w = elu(x.w); }
                if (featmask) {
                    const uint32_t m = (uint32_t)(featmask[node_off + row0 + r] >> (q * 4)) & 0xFu;
                    x.x = (m & 1u) ? x.x * scale : 0.f; x.y = (m & 2u) ? x.y * scale : 0.f;
                    x.z = (m & 4u) ? x.z * scale : 0.f; x.w = (m & 8u) ? x.w * scale : 0.f;
                }
            }
            *reinterpret_cast<float4*>(dhs + r * DW_XS + q * 4) = v;
            *reinterpret_cast<float4*>(Xds + r * DW_XS + q * 4) = x;
        }
        for (int idx = threadIdx.x; idx < DW_TM * H; idx += 256) {
            const int r = idx / H, k = idx % H;
            dSs[r * 8 + k] = r < nrows ? __ldg(dS + (node_off + row0 + r) * H + k) : 0.f;
        }
        __syncthreads();
        // dW[k][c] += sum_r Xd[r][k] dh[r][c]; thread = k 4ty..4ty+3, c 4tx..4tx+3 (rows >= nrows are zero)
#pragma unroll 4
        for (int r = 0; r < DW_TM; ++r) {
            const float4 xs = *reinterpret_cast<const float4*>(Xds + r * DW_XS + ty * 4);
            const float4 ds = *reinterpret_cast<const float4*>(dhs + r * DW_XS + tx * 4);
            accW[0][0] = fmaf(xs.x, ds.x, accW[0][0]); accW[0][1] = fmaf(xs.x, ds.y, accW[0][1]);
            accW[0][2] = fmaf(xs.x, ds.z, accW[0][2]); accW[0][3] = fmaf(xs.x, ds.w, accW[0][3]);
            accW[1][0] = fmaf(xs.y, ds.x, accW[1][0]); accW[1][1] = fmaf(xs.y, ds.y, accW[1][1]);
            accW[1][2] = fmaf(xs.y, ds.z, accW[1][2]); accW[1][3] = fmaf(xs.y, ds.w, accW[1][3]);
            accW[2][0] = fmaf(xs.z, ds.x, accW[2][0]); accW[2][1] = fmaf(xs.z, ds.y, accW[2][1]);
            accW[2][2] = fmaf(xs.z, ds.z, accW[2][2]); accW[2][3] = fmaf(xs.z, ds.w, accW[2][3]);
            accW[3][0] = fmaf(xs.w, ds.x, accW[3][0]); accW[3][1] = fmaf(xs.w, ds.y, accW[3][1]);
            accW[3][2] = fmaf(xs.w, ds.z, accW[3][2]); accW[3][3] = fmaf(xs.w, ds.w, accW[3][3]);
        }
        // V[i][k] += sum_r Xd[r][i] dS[r][k]
        if (H == 8) {
#pragma unroll 4
            for (int r = 0; r < DW_TM; ++r) {
                const float x = Xds[r * DW_XS + vi];
                const float2 d2 = *reinterpret_cast<const float2*>(dSs + r * 8 + vk);
                accV0 = fmaf(x, d2.x, accV0);
                accV1 = fmaf(x, d2.y, accV1);
            }
        } else if (threadIdx.x < 64) {
#pragma unroll 4
            for (int r = 0; r < DW_TM; ++r) accV0 = fmaf(Xds[r * DW_XS + vi], dSs[r * 8], accV0);
        }
    }
    float* part = partials + (size_t)blockIdx.x * DW_PART;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        *reinterpret_cast<float4*>(part + (ty * 4 + i) * 64 + tx * 4) = make_float4(accW[i][0], accW[i][1], accW[i][2], accW[i][3]);
    if (H == 8) {
        part[64 * 64 + vi * 8 + vk] = accV0;
        part[64 * 64 + vi * 8 + vk + 1] = accV1;
    } else if (threadIdx.x < 64) {
        part[64 * 64 + vi * 8] = accV0;
    }
}

// pass 2: fixed-order sums over the CTAs of each side; dW through gtab, V into a small scratch for pass 3.
// One CTA = 32 consecutive outputs x 8 groups of partials: every load is a coalesced 128-byte row of one partial, each
// thread adds its ~1/8 of the partials in order, the 8 group sums are combined in order -> deterministic.
template <int H>
__global__ void __launch_bounds__(256) dw_reduce_kernel(const float* __restrict__ partials, int nb_u, int nb_total, float* const* __restrict__ gtab,
                                                        float* __restrict__ vsum /* [2][64*8] */, int accumulate) {
    constexpr int DH = D / H;
    __shared__ float red[8][33];
    const int og = threadIdx.x & 31, pg = threadIdx.x >> 5;
    const int idx = blockIdx.x * 32 + og;                 // DW_PART is a multiple of 32: a CTA never straddles the two sides
    const int side = idx / DW_PART, o = idx % DW_PART;
    const int b0 = side ? nb_u : 0, b1 = side ? nb_total : nb_u;
    float sum = 0.f;
    for (int b = b0 + pg; b < b1; b += 8) sum += partials[(size_t)b * DW_PART + o];
    red[pg][og] = sum;
    __syncthreads();
    if (pg != 0 || b1 <= b0) return;
    if (o >= 64 * 64 && ((o - 64 * 64) & 7) >= H) return;
    sum = ((red[0][og] + red[1][og]) + (red[2][og] + red[3][og])) + ((red[4][og] + red[5][og]) + (red[6][og] + red[7][og]));
    if (o < 64 * 64) {
        const int k = o >> 6, c = o & 63;
        float* dst = gtab[side * H + c / DH] + k * DH + (c % DH);
        *dst = accumulate ? *dst + sum : sum;
    } else {
        vsum[side * 512 + (o - 64 * 64)] = sum;
    }
}

// pass 3: da[c] = sum_i W_side[i][c] * V[i][head(c)]
template <int H>
__global__ void da_kernel(const float* __restrict__ vsum, const float* const* __restrict__ wtab, float* const* __restrict__ gtab,
                          int has_u, int has_i, int accumulate) {
    constexpr int DH = D / H;
    const int t = threadIdx.x;          // 128 threads: side = t / 64, column c = t % 64
    const int side = t >> 6, c = t & 63;
    if ((side == 0 && !has_u) || (side == 1 && !has_i)) return;
    const int head = c / DH, j = c % DH;
    const float* W = wtab[side * H + head];
    float sum = 0.f;
    for (int i = 0; i < 64; ++i) sum = fmaf(__ldg(W + i * DH + j), vsum[side * 512 + i * 8 + head], sum);
    float* dst = gtab[2 * H + head] + side * DH + j;
    *dst = accumulate ? *dst + sum : sum;
}

}  // namespace ngacf

using namespace ngacf;

static void dw_grid(int U, int I, int* nb_u, int* nb_i) {
    const int tiles_u = ceil_div(U, DW_TM), tiles_i = ceil_div(I, DW_TM);
    const int budget = 3 * 148;   // three resident CTAs per SM
    int bu = (int)((int64_t)budget * tiles_u / (tiles_u + tiles_i > 0 ? tiles_u + tiles_i : 1));
    if (bu < 1) bu = 1;
    if (bu > tiles_u) bu = tiles_u;
    int bi = budget - bu;
    if (bi < 1) bi = 1;
    if (bi > tiles_i) bi = tiles_i;
    *nb_u = bu;
    *nb_i = bi;
}

extern "C" int ngacf_transform_bwd_dx(const float* dh, const float* Xu, const float* Xi, int32_t apply_elu, const uint64_t* featmask, float scale,
                                      const float* const* wtab, int32_t H, int32_t U, int32_t I, float* dXu, float* dXi, int32_t accumulate,
                                      void* stream) {
    NGACF_REQUIRE(dh && wtab && U >= 0 && I >= 0 && (U == 0 || dXu) && (I == 0 || dXi) && (!apply_elu || ((U == 0 || Xu) && (I == 0 || Xi))),
                  "transform_bwd_dx: null argument");
    NGACF_REQUIRE(H == 1 || H == 8, "transform_bwd_dx: H must be 1 or 8");
    if (U + I == 0) return NGACF_OK;
    if (dense_on_tensor_cores()) {
        transform_bwd_dx_tc(dh, Xu, Xi, apply_elu, featmask, scale, wtab, H, U, I, dXu, dXi, accumulate, (cudaStream_t)stream);
        return check_launch("transform_bwd_dx(tc)");
    }
    const int tiles_u = ceil_div(U, DX_TM), tiles_i = ceil_div(I, DX_TM);
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(transform_bwd_dx_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DX_SMEM);
        cudaFuncSetAttribute(transform_bwd_dx_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DX_SMEM);
    });
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 8) transform_bwd_dx_kernel<8><<<tiles_u + tiles_i, 256, DX_SMEM, st>>>(dh, Xu, Xi, apply_elu, featmask, scale, wtab, U, I, tiles_u, dXu, dXi, accumulate);
    else        transform_bwd_dx_kernel<1><<<tiles_u + tiles_i, 256, DX_SMEM, st>>>(dh, Xu, Xi, apply_elu, featmask, scale, wtab, U, I, tiles_u, dXu, dXi, accumulate);
    return check_launch("transform_bwd_dx");
}

extern "C" size_t ngacf_transform_bwd_dw_workspace_bytes(int32_t U, int32_t I) {
    int bu, bi;
    dw_grid(U, I, &bu, &bi);
    return (size_t)(bu + bi) * DW_PART * sizeof(float) + 2 * 512 * sizeof(float);
}

extern "C" int ngacf_transform_bwd_dw(const float* dh, const float* dS, const float* Xu, const float* Xi, int32_t apply_elu,
                                      const uint64_t* featmask, float scale, const float* const* wtab, float* const* gtab, int32_t H, int32_t U,
                                      int32_t I, int32_t accumulate, void* workspace, size_t workspace_bytes, void* stream) {
    NGACF_REQUIRE(dh && dS && wtab && gtab && workspace && U >= 0 && I >= 0 && (U == 0 || Xu) && (I == 0 || Xi), "transform_bwd_dw: null argument");
    NGACF_REQUIRE(H == 1 || H == 8, "transform_bwd_dw: H must be 1 or 8");
    if (U + I == 0) return NGACF_OK;
    if (workspace_bytes < ngacf_transform_bwd_dw_workspace_bytes(U, I)) { set_error("transform_bwd_dw: workspace too small"); return NGACF_ERR_WORKSPACE; }
    int bu, bi;
    dw_grid(U, I, &bu, &bi);
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(transform_bwd_dw_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SMEM);
        cudaFuncSetAttribute(transform_bwd_dw_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)DW_SMEM);
    });
    cudaStream_t st = (cudaStream_t)stream;
    float* partials = (float*)workspace;
    float* vsum = partials + (size_t)(bu + bi) * DW_PART;
    if (H == 8) {
        transform_bwd_dw_kernel<8><<<bu + bi, 256, DW_SMEM, st>>>(dh, dS, Xu, Xi, apply_elu, featmask, scale, U, I, bu, partials);
        dw_reduce_kernel<8><<<2 * DW_PART / 32, 256, 0, st>>>(partials, bu, bu + bi, gtab, vsum, accumulate);
        da_kernel<8><<<1, 128, 0, st>>>(vsum, wtab, gtab, bu > 0, bi > 0, accumulate);
    } else {
        transform_bwd_dw_kernel<1><<<bu + bi, 256, DW_SMEM, st>>>(dh, dS, Xu, Xi, apply_elu, featmask, scale, U, I, bu, partials);
        dw_reduce_kernel<1><<<2 * DW_PART / 32, 256, 0, st>>>(partials, bu, bu + bi, gtab, vsum, accumulate);
        da_kernel<1><<<1, 128, 0, st>>>(vsum, wtab, gtab, bu > 0, bi > 0, accumulate);
    }
    return check_launch("transform_bwd_dw");
}
