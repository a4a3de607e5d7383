// Dense stage transforms on the 5th-generation tensor cores (tcgen05.mma kind::tf32, fp32 accumulators in TMEM).
//   forward   h = drop(act(X)) Wcat, s = per-head a . h          (SPUIGACF.py:30-39,356-361)
//   backward  dX = (dh Wcat^T) * mask/(1-p) * ELU'(Zprev)          (autograd of the above w.r.t. the stage input)
// Both are (rows x 64) x (64 x 64) products with K = 64: far too thin for the FFMA pipe to keep up with HBM
// (profiles/r1e_ab_prefetch.txt), so the product runs as "3xTF32": every fp32 operand is split into a TF32 head and a
// TF32 tail (x = hi + lo, |lo| <= 2^-11 |x|) and D = Ahi Bhi + Alo Bhi + Ahi Blo is accumulated in fp32.  The dropped
// lo*lo term and the tail rounding are ~2^-22 relative per product -- fp32-class accuracy (parity bar: 1e-4).
//
// One CTA = 256 threads = one 128-row tile at a time (persistent over its side's tiles), two CTAs per SM:
//   16-byte loads (next tile prefetched into registers) -> activation / dropout / split -> K-major SWIZZLE_NONE operand
//   image in shared memory -> one elected thread issues 24 tcgen05.mma (8 k-steps x 3 terms) -> tcgen05.commit -> mbarrier
//   -> thread (row = TMEM lane, column half) reads 32 accumulators with tcgen05.ld -> epilogue -> staged through the (now
//   free) operand buffer for 128-byte row-segment stores.
#include <cuda_bf16.h>
#include "common.cuh"

namespace ngacf {
namespace tcx {

constexpr int TM = 128;
constexpr int PANEL_A = TM * 16;          // one 16-byte k-chunk of all 128 rows
constexpr int A_HALF = 16 * PANEL_A;      // 32 KB: hi (or lo) image of the row tile, 16 k-chunks
constexpr int PANEL_B = 64 * 16;
constexpr int B_HALF = 16 * PANEL_B;      // 16 KB
constexpr int THREADS = 256;
constexpr size_t SMEM_BYTES = 2 * A_HALF + 2 * B_HALF + 64 * 4 /*a vector*/ + 2 * TM * 4 /*logit partials*/ + 64 /*barrier + tmem slot*/ + 128 /*alignment*/;
constexpr uint32_t TMEM_COLS = 64;
// instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major, N = 64, M = 128
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE matrix descriptor: core matrix = 8 rows x 16 B; LBO = distance between the two k-chunks of one
// MMA (one panel), SBO = distance between 8-row groups (128 B).  Same encoding as eval_topk_tc.cu.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t panel_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((panel_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* r) {
    uint32_t u[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]), "=r"(u[17]), "=r"(u[18]),
          "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]),
          "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __uint_as_float(u[i]);
}

// x = hi + lo with hi, lo representable in TF32 (round-to-nearest both times)
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(x));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(x - hi));
    lo = __uint_as_float(l);
}
__device__ __forceinline__ void split4(const float4 x, float4& hi, float4& lo) {
    split_tf32(x.x, hi.x, lo.x); split_tf32(x.y, hi.y, lo.y); split_tf32(x.z, hi.z, lo.z); split_tf32(x.w, hi.w, lo.w);
}

// coalesced-enough tile load: warp w (of 8) owns rows 16w..16w+15; one instruction = 8 rows x 4 chunks (64-byte row segments in
// global memory, 4 x 128 contiguous bytes in the operand image: no bank conflicts)
__device__ __forceinline__ bool item_side_of(unsigned block, int nb_u) { return (int)block >= nb_u; }
__device__ __forceinline__ void tile_rows(int warp, int lane, int it, int& r, int& q) {
    r = 16 * warp + 8 * (it >> 2) + (lane & 7);
    q = 4 * (it & 3) + (lane >> 3);
}

#ifdef NGACF_DENSE_TRACE
// phase timeline of CTA 0 and CTA 150 of the forward kernel (debug builds): [cta][event]; events: 0 start, 1 prologue done,
// then per tile (up to 4): 2+5t image stored, 3+5t MMAs issued, 4+5t accumulator ready, 5+5t epilogue math done, 6+5t stores issued
__device__ long long g_dense_trace[2][32];
#define DTRACE(ev) do { if ((blockIdx.x == 0 || blockIdx.x == 150) && threadIdx.x == 0 && (ev) < 32) g_dense_trace[blockIdx.x ? 1 : 0][ev] = clock64(); } while (0)
#else
#define DTRACE(ev) do { } while (0)
#endif

// MODE 0: forward (A = drop(act(X)), B[n][k] = Wcat[k][n], epilogue h, s)
// MODE 1: backward dX (A = dh, B[n][k] = Wcat[n][k], epilogue mask/ELU'/accumulate)
template <int H, int MODE>
__global__ void __launch_bounds__(THREADS, 2) transform_tc_kernel(const float* __restrict__ Au, const float* __restrict__ Ai,   // fwd: X; bwd: dh rows
                                                                  const float* __restrict__ Zu, const float* __restrict__ Zi,   // bwd: stage input (ELU'), else null
                                                                  int apply_elu, const uint64_t* __restrict__ featmask, float scale,
                                                                  const float* const* __restrict__ wtab, int U, int I, int nb_u,
                                                                  float* __restrict__ Ou, float* __restrict__ Oi,               // fwd: h rows; bwd: dX
                                                                  float* __restrict__ s, int accumulate) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* sAhi = base;
    uint8_t* sAlo = base + A_HALF;
    uint8_t* sBhi = base + 2 * A_HALF;
    uint8_t* sBlo = sBhi + B_HALF;
    float* av = reinterpret_cast<float*>(sBlo + B_HALF);      // [64] logit vector of this side
    float* sp = av + 64;                                      // [2][128] H=1: per-row partial logits of the two column halves
    uint64_t* bar = reinterpret_cast<uint64_t*>(sp + 2 * TM);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    constexpr int DH = D / H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool item_side = (int)blockIdx.x >= nb_u;
    const int bs = item_side ? blockIdx.x - nb_u : blockIdx.x;
    const int nbs = item_side ? gridDim.x - nb_u : nb_u;
    const int rows_side = item_side ? I : U;
    const float* A = item_side ? Ai : Au;
    const float* Zin = item_side ? Zi : Zu;
    float* O = item_side ? Oi : Ou;
    const int64_t node_off = item_side ? U : 0;
    const int tiles = (rows_side + TM - 1) / TM;

    DTRACE(0);
    int tcount = 0;
    // ---- the weight block sits behind the pointer table: two dependent round trips, so the pointers go first ----
    const float* wp[4];
    {
        const float* const* wptr = wtab + (item_side_of(blockIdx.x, nb_u) ? H : 0);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int idx = threadIdx.x + j * THREADS;
            const int col = MODE == 0 ? (idx & 63) : (idx & 15) * 4;
            wp[j] = wptr[col / (D / H)];
        }
    }
    // ---- then the first tile's rows ----
    float4 v[8];
    uint64_t mw[2] = {0ull, 0ull};
    auto request = [&](int tile) {
        const int row0 = tile * TM;
        const int nrows = min(TM, rows_side - row0);
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int r, q;
            tile_rows(warp, lane, it, r, q);
            v[it] = r < nrows ? ld_stream4(A + (int64_t)(row0 + r) * D + q * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (MODE == 0 && featmask) {
#pragma unroll
            for (int rg = 0; rg < 2; ++rg) {
                const int r = 16 * warp + 8 * rg + (lane & 7);
                mw[rg] = r < nrows ? featmask[node_off + row0 + r] : 0ull;
            }
        }
    };
    if (bs < tiles) request(bs);
    DTRACE(27);

    // ---- one-time: barrier, TMEM, operand B (the side's 64x64 weight block, split) ----
    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    DTRACE(28);
    {
        if (MODE == 0) {
            // B[n][kk = k] = Wcat[k][n]: one item = (column n, four consecutive k) = one 16-byte slot of the K-major image, so the
            // transposed image is written with conflict-free 16-byte stores (it was 32 scalar stores per thread, 8-way conflicts);
            // the four scalar loads of an item are coalesced over n
            float wv[4][4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + j * THREADS, n = idx & 63, kg = idx >> 6;
                const float* src = wp[j] + (4 * kg) * DH + (n % DH);
#pragma unroll
                for (int i = 0; i < 4; ++i) wv[j][i] = __ldg(src + i * DH);
            }
            DTRACE(29);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + j * THREADS, n = idx & 63, kg = idx >> 6;
                float4 hi, lo;
                split4(make_float4(wv[j][0], wv[j][1], wv[j][2], wv[j][3]), hi, lo);
                const uint32_t o = (uint32_t)kg * PANEL_B + (uint32_t)n * 16;
                *reinterpret_cast<float4*>(sBhi + o) = hi;
                *reinterpret_cast<float4*>(sBlo + o) = lo;
            }
        } else {
            float4 w[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int idx = tid + j * THREADS, k = idx >> 4, c = (idx & 15) * 4;
                w[j] = __ldg(reinterpret_cast<const float4*>(wp[j] + k * DH + (c % DH)));     // Wcat[k][c..c+3]
            }
            DTRACE(29);
#pragma unroll
            for (int j = 0; j < 4; ++j) {           // B[n = k][kk = c..c+3]
                const int idx = tid + j * THREADS, k = idx >> 4, c = (idx & 15) * 4;
                float4 hi, lo;
                split4(w[j], hi, lo);
                const uint32_t o = (uint32_t)(c >> 2) * PANEL_B + (uint32_t)k * 16;
                *reinterpret_cast<float4*>(sBhi + o) = hi;
                *reinterpret_cast<float4*>(sBlo + o) = lo;
            }
        }
        if (MODE == 0 && tid < 64) av[tid] = __ldg(wtab[2 * H + tid / DH] + (item_side ? DH : 0) + tid % DH);
    }
    DTRACE(30);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    uint32_t phase = 0;
    DTRACE(1);
    const uint64_t dAhi = smem_desc(smem_u32(sAhi), PANEL_A), dAlo = smem_desc(smem_u32(sAlo), PANEL_A);
    const uint64_t dBhi = smem_desc(smem_u32(sBhi), PANEL_B), dBlo = smem_desc(smem_u32(sBlo), PANEL_B);
    const int quarter = warp & 3, half = warp >> 2;      // TMEM lanes 32*quarter.., accumulator columns 32*half..

    for (int tile = bs; tile < tiles; tile += nbs) {
        const int row0 = tile * TM;
        const int nrows = min(TM, rows_side - row0);
        // ---- A image of this tile from the prefetched registers ----
#pragma unroll
        for (int it = 0; it < 8; ++it) {
            int r, q;
            tile_rows(warp, lane, it, r, q);
            float4 x = v[it];
            if (MODE == 0) {
                if (apply_elu) { x.x = elu(x.x); x.y = elu(x.y); x.z = elu(x.z); x.w = elu(x.w); }
                if (featmask) {
                    const uint32_t m = (uint32_t)(mw[it >> 2] >> (q * 4)) & 0xFu;
                    x.x = (m & 1u) ? x.x * scale : 0.f; x.y = (m & 2u) ? x.y * scale : 0.f;
                    x.z = (m & 4u) ? x.z * scale : 0.f; x.w = (m & 8u) ? x.w * scale : 0.f;
                }
            }
            float4 hi, lo;
            split4(x, hi, lo);
            const uint32_t o = (uint32_t)q * PANEL_A + (uint32_t)r * 16;
            *reinterpret_cast<float4*>(sAhi + o) = hi;
            *reinterpret_cast<float4*>(sAlo + o) = lo;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        DTRACE(2 + 5 * tcount);
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            // the two 2^-11 terms first, hi.hi last: every accumulation rounds at the accumulator's current magnitude
            for (int term = 0; term < 3; ++term) {
                const uint64_t a0 = term == 0 ? dAlo : dAhi;
                const uint64_t b0 = term == 1 ? dBlo : dBhi;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)          // K = 8 TF32 per MMA = two 16-byte k-chunk panels (start address field += bytes/16)
                    umma_tf32(tmem_base, a0 + (uint64_t)(ks * 2 * PANEL_A / 16), b0 + (uint64_t)(ks * 2 * PANEL_B / 16), (term | ks) ? 1u : 0u);
            }
            umma_commit(smem_u32(bar));
            DTRACE(3 + 5 * tcount);
        }
        // the next tile's rows -- and what this tile's epilogue needs from global memory -- travel while the tensor core works
        if (tile + nbs < tiles) request(tile + nbs);
        float4 zz[8];
        uint32_t em[8];
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int r = 32 * quarter + 4 * j + (lane >> 3), q = 8 * half + (lane & 7);
                const bool ok = r < nrows;
                em[j] = (featmask && ok) ? (uint32_t)(featmask[node_off + row0 + r] >> (q * 4)) & 0xFu : 0xFu;
                zz[j] = (apply_elu && ok) ? ld_stream4(Zin + (int64_t)(row0 + r) * D + q * 4) : make_float4(1.f, 1.f, 1.f, 1.f);
            }
        }
        mbar_wait(smem_u32(bar), phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        DTRACE(4 + 5 * tcount);

        // ---- epilogue: thread = (row = TMEM lane 32*quarter + lane, 32 accumulator columns 32*half..) ----
        float acc[32];
        tmem_ld32(tmem_base + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(32 * half), acc);
        const int rl = 32 * quarter + lane;
        if (MODE == 0) {
            if (H == 8) {
                float sv[4];
#pragma unroll
                for (int hh = 0; hh < 4; ++hh) {
                    const float* ap = av + 32 * half + hh * 8;
                    const float4 a0 = *reinterpret_cast<const float4*>(ap), a1 = *reinterpret_cast<const float4*>(ap + 4);
                    float p = acc[hh * 8] * a0.x;
                    p = fmaf(acc[hh * 8 + 1], a0.y, p); p = fmaf(acc[hh * 8 + 2], a0.z, p); p = fmaf(acc[hh * 8 + 3], a0.w, p);
                    p = fmaf(acc[hh * 8 + 4], a1.x, p); p = fmaf(acc[hh * 8 + 5], a1.y, p); p = fmaf(acc[hh * 8 + 6], a1.z, p);
                    p = fmaf(acc[hh * 8 + 7], a1.w, p);
                    sv[hh] = p;
                }
                if (rl < nrows) *reinterpret_cast<float4*>(s + (node_off + row0 + rl) * 8 + 4 * half) = make_float4(sv[0], sv[1], sv[2], sv[3]);
            } else {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
                for (int c = 0; c < 32; c += 4) {
                    const float4 a = *reinterpret_cast<const float4*>(av + 32 * half + c);
                    p0 = fmaf(acc[c], a.x, p0); p1 = fmaf(acc[c + 1], a.y, p1); p2 = fmaf(acc[c + 2], a.z, p2); p3 = fmaf(acc[c + 3], a.w, p3);
                }
                sp[half * TM + rl] = (p0 + p1) + (p2 + p3);
            }
        }
        // stage this warp's 32 rows x 128 bytes through its slice of the (retired) A-hi image: row r at r*128, 16-byte chunk c at
        // slot c ^ (r & 7) -> conflict-free for the row-per-thread writes and the 4-rows-per-instruction reads
        DTRACE(5 + 5 * tcount);
        uint8_t* stg = sAhi + warp * 4096;
#pragma unroll
        for (int c = 0; c < 8; ++c)
            *reinterpret_cast<float4*>(stg + lane * 128 + ((c ^ (lane & 7)) * 16)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int rr = 4 * j + (lane >> 3), qq = lane & 7;
            const int r = 32 * quarter + rr, q = 8 * half + qq;
            if (r >= nrows) continue;
            float4 o4 = *reinterpret_cast<const float4*>(stg + rr * 128 + ((qq ^ (rr & 7)) * 16));
            float* dst = O + (int64_t)(row0 + r) * D + q * 4;
            if (MODE == 1) {
                if (featmask) {
                    const uint32_t m = em[j];
                    o4.x = (m & 1u) ? o4.x * scale : 0.f; o4.y = (m & 2u) ? o4.y * scale : 0.f;
                    o4.z = (m & 4u) ? o4.z * scale : 0.f; o4.w = (m & 8u) ? o4.w * scale : 0.f;
                }
                if (apply_elu) {
                    const float4 z = zz[j];
                    o4.x *= elu_grad(z.x); o4.y *= elu_grad(z.y); o4.z *= elu_grad(z.z); o4.w *= elu_grad(z.w);
                }
                if (accumulate) {
                    const float4 o = *reinterpret_cast<const float4*>(dst);
                    o4.x += o.x; o4.y += o.y; o4.z += o.z; o4.w += o.w;
                }
                *reinterpret_cast<float4*>(dst) = o4;
            } else {
                st_stream4(dst, o4);
            }
        }
        // the staging reads and the TMEM reads of this tile must retire before the next tile's image / MMAs overwrite them
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        DTRACE(6 + 5 * tcount);
        ++tcount;
        if (MODE == 0 && H == 1 && tid < nrows) s[node_off + row0 + tid] = sp[tid] + sp[TM + tid];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace tcx
}  // namespace ngacf
#ifdef NGACF_DENSE_TRACE
extern "C" int ngacf_debug_dense_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, ngacf::tcx::g_dense_trace, sizeof(ngacf::tcx::g_dense_trace));
}
#endif
namespace ngacf {
namespace tcx {

static void side_grid(int U, int I, int* nb_u, int* nb_i) {
    const int tiles_u = ceil_div(U, TM), tiles_i = ceil_div(I, TM);
    const int budget = 2 * 148;
    int bu = (int)((int64_t)budget * tiles_u / (tiles_u + tiles_i > 0 ? tiles_u + tiles_i : 1));
    if (bu < 1) bu = 1;
    if (bu > tiles_u) bu = tiles_u;
    int bi = budget - bu;
    if (bi < 1) bi = 1;
    if (bi > tiles_i) bi = tiles_i;
    *nb_u = bu;
    *nb_i = bi;
}

template <int H, int MODE>
static void launch(const float* Au, const float* Ai, const float* Zu, const float* Zi, int apply_elu, const uint64_t* featmask, float scale,
                   const float* const* wtab, int U, int I, float* Ou, float* Oi, float* s, int accumulate, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(transform_tc_kernel<H, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    });
    int bu, bi;
    side_grid(U, I, &bu, &bi);
    transform_tc_kernel<H, MODE><<<bu + bi, THREADS, SMEM_BYTES, st>>>(Au, Ai, Zu, Zi, apply_elu, featmask, scale, wtab, U, I, bu, Ou, Oi, s, accumulate);
}


// =================================================================================================
// Fused dense backward of one stage on the tensor cores (replaces the CUDA-core transform_bwd_kernel):
//   dX  = (dh W^T) * F,  F = mask/(1-p) * ELU'(Zprev)                       D1[128 rows x 64]   per tile
//   dW += Xd^T dh,  Q += Xd^T dS  (da[c] = sum_k W[k][c] Q[k][head(c)])     D2[64(+64) x 80]    accumulated over the CTA's tiles
// One 512-thread CTA per SM, persistent over its side's 128-row tiles.
//
// Operand format.  The dW product contracts over ROWS, i.e. both of its operands are MN-major in their natural row-major
// tiles.  kind::tf32 returns zeros for MN-major operands on sm_100a (scripts/probe/run_probe.py, profiles/r1e_umma_probe.txt),
// kind::f16 supports them, so this kernel splits every fp32 value into THREE bf16 terms x = h + m + l (8+8+8 mantissa bits,
// exact) and accumulates the six products with weight >= 2^-16 (hh, hm, mh, mm, hl, lh): dropped terms are <= 2^-23 relative,
// fp32-class accuracy at the bf16 rate (6 bf16 MMAs cost what 3 tf32 MMAs do).
//
// The two row-major tiles are written ONCE, as SWIZZLE_NONE images [16-byte chunk = 8 features][row][16 B]: read K-major they
// are the A operand of the dX product (K = feature), read MN-major (LBO/SBO roles swapped) they are the A and B operands of
// the dW product (K = row) -- no transposed copy.  The dS tile is appended to the dh image as chunk 8, so Q comes out as
// accumulator columns 64..71 of the same MMAs.  M = 128 for both products (the M = 64 accumulator layout is not row = lane):
// for dW only lanes 0..63 (k) are meaningful, lanes 64..127 come from reading past the Xd image (ignored).
// =================================================================================================
namespace bwd {
constexpr int THREADS = 512;
constexpr int PANEL = TM * 16;                       // 2048: one 16-byte chunk column of 128 rows
constexpr int DH_TERM = 10 * PANEL;                  // dh (8 chunks) + dS (chunk 8) + pad (chunk 9): one bf16 term
constexpr int XD_TERM = 8 * PANEL;
constexpr int WPANEL = 64 * 16;                      // W image of the dX product: [chunk = 8 c][row = k][16 B]
constexpr int W_TERM = 8 * WPANEL;
constexpr int OFF_DH = 0;                            // [h | m | l]
constexpr int OFF_XD = 3 * DH_TERM;                  // [h | m | l]
constexpr int OFF_W = OFF_XD + 3 * XD_TERM;          // [h | m | l]; also the slack the M = 128 reads of the last Xd term run into
constexpr int OFF_F = OFF_W + 3 * W_TERM;            // fp32 factor tile in the epilogue's staging layout (M = 128 slack too)
constexpr int OFF_MISC = OFF_F + TM * 64 * 4;
constexpr size_t SMEM_BYTES = OFF_MISC + 64 * 8 * 4 /*Q*/ + 64 /*barrier, tmem slot*/ + 128 /*alignment*/;
constexpr uint32_t TMEM_COLS = 256;                  // D1: columns 0..63, D2 (h.h): 64..143, D2s (corrections): 144..223
constexpr uint32_t IDESC_DX = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);     // bf16, K-major, N = 64
constexpr uint32_t IDESC_DW = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(80 >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
constexpr int PART = 64 * 64 + 64;                   // floats per CTA partial (same layout as propagate_bwd.cu: dW then da)
// term pairs ordered by weight: l.h, h.l, m.m (2^-16), m.h, h.m (2^-8), h.h
__device__ constexpr int TERM_A[6] = {2, 0, 1, 1, 0, 0};
__device__ constexpr int TERM_B[6] = {0, 2, 1, 0, 1, 0};

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
}
// MN-major SWIZZLE_NONE descriptor over the same image: SBO = distance between 16-byte chunks along M/N (one panel),
// LBO = distance between the two 8-row groups of one MMA along K (128 B)
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((PANEL >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* r) {
    uint32_t u[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]), "=r"(u[9]),
          "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) r[i] = __uint_as_float(u[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* r) {
    uint32_t u[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = __uint_as_float(u[i]);
}
// x = h + m + l, three bf16 terms (round-to-nearest, residuals exact in fp32); returns the terms of two values packed lo|hi
__device__ __forceinline__ void split3_pair(float x0, float x1, uint32_t& h, uint32_t& m, uint32_t& l) {
    // packed conversions (cvt.rn.bf16x2.f32): low half = first value; a bf16 widens to fp32 by a 16-bit shift
    __nv_bfloat162 t = __floats2bfloat162_rn(x0, x1);
    h = *reinterpret_cast<uint32_t*>(&t);
    const float r0 = x0 - __uint_as_float(h << 16), r1 = x1 - __uint_as_float(h & 0xFFFF0000u);
    t = __floats2bfloat162_rn(r0, r1);
    m = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(r0 - __uint_as_float(m << 16), r1 - __uint_as_float(m & 0xFFFF0000u));
    l = *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void split3x8(const float4 a, const float4 b, uint4& h, uint4& m, uint4& l) {
    split3_pair(a.x, a.y, h.x, m.x, l.x); split3_pair(a.z, a.w, h.y, m.y, l.y);
    split3_pair(b.x, b.y, h.z, m.z, l.z); split3_pair(b.z, b.w, h.w, m.w, l.w);
}
// staging / factor-tile layout: slice of epilogue warp (quarter, cq) = 32 rows x 64 B; 16-byte chunk qq of row rr at slot qq ^ ((rr>>1)&3)
__device__ __forceinline__ uint32_t stage_off(int r, int q) {
    const int rr = r & 31, quarter = r >> 5, cq = q >> 2, qq = q & 3;
    return (uint32_t)((quarter + 4 * cq) * 2048 + rr * 64 + ((qq ^ ((rr >> 1) & 3)) * 16));
}

#ifdef NGACF_DENSE_TRACE
// phase timeline of CTA 0 and CTA 100 of the fused backward kernel: 0 start, 1 prologue done, per tile t (up to 6): 2+4t images stored,
// 3+4t MMAs issued, 4+4t accumulators ready, 5+4t epilogue done; 30 = partials written
__device__ long long g_dense_bwd_trace[2][32];
#define BTRACE(ev) do { if ((blockIdx.x == 0 || blockIdx.x == 100) && threadIdx.x == 0 && (ev) < 32) g_dense_bwd_trace[blockIdx.x ? 1 : 0][ev] = clock64(); } while (0)
#else
#define BTRACE(ev) do { } while (0)
#endif

template <int H>
__global__ void __launch_bounds__(THREADS, 1) transform_bwd_tc_kernel(const float* __restrict__ dh, const float* __restrict__ dS,
                                                                      const float* __restrict__ Xu, const float* __restrict__ Xi, int apply_elu,
                                                                      const uint64_t* __restrict__ featmask, float scale,
                                                                      const float* const* __restrict__ wtab, int U, int I, int nb_u,
                                                                      float* __restrict__ dXu, float* __restrict__ dXi, int accumulate_dx,
                                                                      float* __restrict__ partials) {
    BTRACE(0);
    int tcount = 0;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~(uintptr_t)127);
    uint8_t* sDH = base + OFF_DH;
    uint8_t* sXD = base + OFF_XD;
    uint8_t* sW = base + OFF_W;
    uint8_t* sF = base + OFF_F;
    float* Qs = reinterpret_cast<float*>(base + OFF_MISC);        // [64][8]
    uint64_t* bar = reinterpret_cast<uint64_t*>(Qs + 64 * 8);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);

    constexpr int DH_ = D / H;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool item_side = (int)blockIdx.x >= nb_u;
    const int bs = item_side ? blockIdx.x - nb_u : blockIdx.x;
    const int nbs = item_side ? gridDim.x - nb_u : nb_u;
    const int rows_side = item_side ? I : U;
    const float* X = item_side ? Xi : Xu;
    float* dX = item_side ? dXi : dXu;
    const int64_t node_off = item_side ? U : 0;
    const int tiles = (rows_side + TM - 1) / TM;

    // the weight block sits behind the pointer table (two dependent round trips): its pointer is requested before anything else
    const float* wsrc = (wtab + (item_side ? H : 0))[((tid & 7) * 8) / DH_];
    // tile loads: warp w (of 16) owns rows 8w..8w+7; item `it` = those 8 rows x feature chunks (8 floats) 4it..4it+3
    const int lr = 8 * warp + (lane & 7);            // this thread's row within the tile (fixed)
    float4 vd[4], vx[4], vs0 = make_float4(0.f, 0.f, 0.f, 0.f), vs1 = vs0;
    uint64_t mw = 0ull;
    auto request = [&](int tile) {
        const int row0 = tile * TM;
        const bool ok = row0 + lr < rows_side;
        const int64_t node = node_off + row0 + lr;
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int q8 = 4 * it + (lane >> 3);
            const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
            vd[2 * it] = ok ? ld_stream4(dh + node * D + q8 * 8) : z4;
            vd[2 * it + 1] = ok ? ld_stream4(dh + node * D + q8 * 8 + 4) : z4;
            vx[2 * it] = ok ? ld_stream4(X + (int64_t)(row0 + lr) * D + q8 * 8) : z4;
            vx[2 * it + 1] = ok ? ld_stream4(X + (int64_t)(row0 + lr) * D + q8 * 8 + 4) : z4;
        }
        mw = (featmask && ok) ? featmask[node] : (ok ? ~0ull : 0ull);
        // dS: one 16-byte bf16 chunk per row, built by threads 0..127 (row = tid)
        vs0 = make_float4(0.f, 0.f, 0.f, 0.f);
        vs1 = vs0;
        if (tid < TM && row0 + tid < rows_side) {
            if (H == 8) {
                vs0 = ld_stream4(dS + (node_off + row0 + tid) * 8);
                vs1 = ld_stream4(dS + (node_off + row0 + tid) * 8 + 4);
            } else {
                vs0.x = __ldg(dS + node_off + row0 + tid);
            }
        }
    };
    if (bs < tiles) request(bs);

    if (tid == 0) {
        mbar_init(smem_u32(bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    {   // W image of the dX product: B[n = k][kk = c] = Wcat[k][c], one (row k, 8-feature chunk) item per thread
        const int k = tid >> 3, c = (tid & 7) * 8;
        const float* src = wsrc + k * DH_ + (c % DH_);               // DH is 8 or 64: the 8 values are contiguous
        const float4 w0 = __ldg(reinterpret_cast<const float4*>(src)), w1 = __ldg(reinterpret_cast<const float4*>(src + 4));
        uint4 h, m, l;
        split3x8(w0, w1, h, m, l);
        const uint32_t o = (uint32_t)(c >> 3) * WPANEL + (uint32_t)k * 16;
        *reinterpret_cast<uint4*>(sW + o) = h;
        *reinterpret_cast<uint4*>(sW + W_TERM + o) = m;
        *reinterpret_cast<uint4*>(sW + 2 * W_TERM + o) = l;
    }
    // chunk 9 of the dh image (accumulator columns 72..79) is never written by the tile loop: clear it once
    for (int i = tid; i < 3 * (PANEL / 16); i += THREADS)
        *reinterpret_cast<uint4*>(sDH + (i / (PANEL / 16)) * DH_TERM + 9 * PANEL + (i % (PANEL / 16)) * 16) = make_uint4(0u, 0u, 0u, 0u);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tD1 = tmem_base, tD2 = tmem_base + 64, tD2s = tmem_base + 144;
    uint32_t phase = 0;
    const uint64_t dDHk = smem_desc(smem_u32(sDH), PANEL), dWk = smem_desc(smem_u32(sW), WPANEL);      // K-major views (dX product)
    const uint64_t dXDm = smem_desc_mn(smem_u32(sXD)), dDHm = smem_desc_mn(smem_u32(sDH));                // MN-major views (dW product)
    const int quarter = warp & 3, cq = warp >> 2;     // epilogue: TMEM lanes 32*quarter.., D1 columns 16*cq..
    const float sc = featmask ? scale : 1.f;
    bool first = true;
    BTRACE(1);

    for (int tile = bs; tile < tiles; tile += nbs) {
        const int row0 = tile * TM;
        const int nrows = min(TM, rows_side - row0);
        // ---- operand images + factor tile from the prefetched registers ----
#pragma unroll
        for (int it = 0; it < 2; ++it) {
            const int q8 = 4 * it + (lane >> 3);
            const uint32_t o = (uint32_t)q8 * PANEL + (uint32_t)lr * 16;
            uint4 h, m, l;
            split3x8(vd[2 * it], vd[2 * it + 1], h, m, l);
            *reinterpret_cast<uint4*>(sDH + o) = h;
            *reinterpret_cast<uint4*>(sDH + DH_TERM + o) = m;
            *reinterpret_cast<uint4*>(sDH + 2 * DH_TERM + o) = l;
            // recomputed stage input (same expressions as the forward) and the dX factor mask/(1-p) * ELU'(z)
            float4 xs[2];
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const float4 z = vx[2 * it + j];
                float4 x = z, f = make_float4(sc, sc, sc, sc);
                if (apply_elu) {
                    x.x = elu(z.x); x.y = elu(z.y); x.z = elu(z.z); x.w = elu(z.w);
                    f.x *= z.x > 0.f ? 1.f : x.x + 1.f; f.y *= z.y > 0.f ? 1.f : x.y + 1.f;
                    f.z *= z.z > 0.f ? 1.f : x.z + 1.f; f.w *= z.w > 0.f ? 1.f : x.w + 1.f;
                }
                const int q = 2 * q8 + j;
                const uint32_t mk = (uint32_t)(mw >> (q * 4)) & 0xFu;
                x.x = (mk & 1u) ? x.x * sc : 0.f; x.y = (mk & 2u) ? x.y * sc : 0.f; x.z = (mk & 4u) ? x.z * sc : 0.f; x.w = (mk & 8u) ? x.w * sc : 0.f;
                f.x = (mk & 1u) ? f.x : 0.f; f.y = (mk & 2u) ? f.y : 0.f; f.z = (mk & 4u) ? f.z : 0.f; f.w = (mk & 8u) ? f.w : 0.f;
                xs[j] = x;
                *reinterpret_cast<float4*>(sF + stage_off(lr, q)) = f;
            }
            split3x8(xs[0], xs[1], h, m, l);
            *reinterpret_cast<uint4*>(sXD + o) = h;
            *reinterpret_cast<uint4*>(sXD + XD_TERM + o) = m;
            *reinterpret_cast<uint4*>(sXD + 2 * XD_TERM + o) = l;
        }
        if (tid < TM) {
            uint4 h, m, l;
            split3x8(vs0, vs1, h, m, l);
            const uint32_t o = (uint32_t)8 * PANEL + (uint32_t)tid * 16;
            *reinterpret_cast<uint4*>(sDH + o) = h;
            *reinterpret_cast<uint4*>(sDH + DH_TERM + o) = m;
            *reinterpret_cast<uint4*>(sDH + 2 * DH_TERM + o) = l;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        BTRACE(2 + 4 * tcount);
        // (two issuing threads with accumulators of their own were measured: no gain -- the 72 MMAs of a tile are tensor-pipe time,
        // ~59 cycles each, not issue time)
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // the six term pairs (a, b) with a + b <= 2 (0 = h, 1 = m, 2 = l), smallest products first: every accumulation rounds at
            // the accumulator's current magnitude; descriptors = base + constant (start address field is in 16-byte units)
            // dX: D1 = sum DH_a W_b   (K-major: K = feature, 16 per MMA = two chunk panels)
#pragma unroll
            for (int pr = 0; pr < 6; ++pr) {
                const int ta = TERM_A[pr], tb = TERM_B[pr];
#pragma unroll
                for (int ks = 0; ks < 4; ++ks)
                    umma(tD1, dDHk + (uint64_t)((ta * DH_TERM + ks * 2 * PANEL) >> 4), dWk + (uint64_t)((tb * W_TERM + ks * 2 * WPANEL) >> 4), IDESC_DX,
                         (pr | ks) ? 1u : 0u);
            }
            // dW | Q: D2 += sum XD_a^T [DH|DS]_b   (both MN-major: K = rows, 16 per MMA = two 8-row groups)
#pragma unroll
            // D2 lives across the CTA's tiles, so the five correction products (<= 2^-8 of h.h) get their own accumulator D2s
            // and meet h.h only once, in the final fp32 add
            for (int pr = 0; pr < 6; ++pr) {
                const int ta = TERM_A[pr], tb = TERM_B[pr];
                const uint32_t dst = pr == 5 ? tD2 : tD2s;
#pragma unroll
                for (int ks = 0; ks < 8; ++ks)
                    umma(dst, dXDm + (uint64_t)((ta * XD_TERM + ks * 256) >> 4), dDHm + (uint64_t)((tb * DH_TERM + ks * 256) >> 4), IDESC_DW,
                         (!first || ((pr != 0 && pr != 5) | ks)) ? 1u : 0u);
            }
            umma_commit(smem_u32(bar));
            BTRACE(3 + 4 * tcount);
        }
        first = false;
        if (tile + nbs < tiles) request(tile + nbs);
        mbar_wait(smem_u32(bar), phase);
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        BTRACE(4 + 4 * tcount);

        // ---- dX epilogue: thread = (row = TMEM lane 32*quarter + lane, 16 columns 16*cq..) ----
        float acc[16];
        tmem_ld16(tD1 + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(16 * cq), acc);
        uint8_t* stg = sDH + warp * 2048;            // the dh images are retired: 16 slices x 2 KB = the first 32 KB of that region
#pragma unroll
        for (int c = 0; c < 4; ++c)
            *reinterpret_cast<float4*>(stg + lane * 64 + ((c ^ ((lane >> 1) & 3)) * 16)) = make_float4(acc[4 * c], acc[4 * c + 1], acc[4 * c + 2], acc[4 * c + 3]);
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int rr = 8 * j + (lane >> 2), qq = lane & 3;
            const int r = 32 * quarter + rr, q = 4 * cq + qq;
            if (r >= nrows) continue;
            const uint32_t so = (uint32_t)(rr * 64 + ((qq ^ ((rr >> 1) & 3)) * 16));
            float4 o4 = *reinterpret_cast<const float4*>(stg + so);
            const float4 f = *reinterpret_cast<const float4*>(sF + warp * 2048 + so);
            o4.x *= f.x; o4.y *= f.y; o4.z *= f.z; o4.w *= f.w;
            float* dst = dX + (int64_t)(row0 + r) * D + q * 4;
            if (accumulate_dx) {
                const float4 o = *reinterpret_cast<const float4*>(dst);
                o4.x += o.x; o4.y += o.y; o4.z += o.z; o4.w += o.w;
            }
            *reinterpret_cast<float4*>(dst) = o4;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        BTRACE(5 + 4 * tcount);
        ++tcount;
    }

    // ---- CTA partial: dW[k][c] (lanes 0..63 of D2, columns 0..63) and da through Q (columns 64..71) ----
    float* part = partials + (size_t)blockIdx.x * PART;
    if (first) {          // no tile (cannot happen with the grid below; keeps the reduction well defined)
        for (int i = tid; i < PART; i += THREADS) part[i] = 0.f;
    } else {
        if (quarter < 2) {
            float acc[16];
            const int k = 32 * quarter + lane;
            float accs[16];
            tmem_ld16(tD2 + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(16 * cq), acc);
            tmem_ld16(tD2s + ((uint32_t)(32 * quarter) << 16) + (uint32_t)(16 * cq), accs);
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<float4*>(part + k * 64 + 16 * cq + 4 * c) =
                    make_float4(acc[4 * c] + accs[4 * c], acc[4 * c + 1] + accs[4 * c + 1], acc[4 * c + 2] + accs[4 * c + 2], acc[4 * c + 3] + accs[4 * c + 3]);
            if (cq == 0) {
                float qv[8], qs[8];
                tmem_ld8(tD2 + ((uint32_t)(32 * quarter) << 16) + 64u, qv);
                tmem_ld8(tD2s + ((uint32_t)(32 * quarter) << 16) + 64u, qs);
#pragma unroll
                for (int j = 0; j < 8; ++j) Qs[k * 8 + j] = qv[j] + qs[j];
            }
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid < 64) {
            const int c = tid, hd = c / DH_;
            float a = 0.f;
#pragma unroll 8
            for (int k = 0; k < 64; ++k) {
                const uint32_t o = (uint32_t)(c >> 3) * WPANEL + (uint32_t)k * 16 + (uint32_t)(c & 7) * 2;
                const float w = (__bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sW + o)) +
                                 __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sW + W_TERM + o))) +
                                __bfloat162float(*reinterpret_cast<const __nv_bfloat16*>(sW + 2 * W_TERM + o));     // h + m + l = W[k][c]
                a = fmaf(w, Qs[k * 8 + hd], a);
            }
            part[64 * 64 + c] = a;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    BTRACE(30);
    if (warp == 0) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

}  // namespace bwd
}  // namespace tcx
}  // namespace ngacf
#ifdef NGACF_DENSE_TRACE
extern "C" int ngacf_debug_dense_bwd_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, ngacf::tcx::bwd::g_dense_bwd_trace, sizeof(ngacf::tcx::bwd::g_dense_bwd_trace));
}
#endif
namespace ngacf {
namespace tcx {
namespace bwd {

static void side_grid1(int U, int I, int* nb_u, int* nb_i) {
    const int tiles_u = ceil_div(U, TM), tiles_i = ceil_div(I, TM);
    const int budget = 148;
    int bu = (int)((int64_t)budget * tiles_u / (tiles_u + tiles_i > 0 ? tiles_u + tiles_i : 1));
    if (bu < 1) bu = 1;
    if (bu > tiles_u) bu = tiles_u;
    int bi = budget - bu;
    if (bi < 1) bi = 1;
    if (bi > tiles_i) bi = tiles_i;
    *nb_u = bu;
    *nb_i = bi;
}
}  // namespace bwd

}  // namespace tcx

// called by ngacf_transform_fwd / ngacf_transform_bwd_dx (propagate_fwd.cu, transform_bwd_split.cu)
void transform_fwd_tc(const float* Xu, const float* Xi, int apply_elu, const uint64_t* featmask, float scale, const float* const* wtab, int H,
                      int U, int I, float* h, float* s, cudaStream_t st) {
    float* hi = h + (int64_t)U * D;
    if (H == 8) tcx::launch<8, 0>(Xu, Xi, nullptr, nullptr, apply_elu, featmask, scale, wtab, U, I, h, hi, s, 0, st);
    else        tcx::launch<1, 0>(Xu, Xi, nullptr, nullptr, apply_elu, featmask, scale, wtab, U, I, h, hi, s, 0, st);
}
void transform_bwd_dx_tc(const float* dh, const float* Zu, const float* Zi, int apply_elu, const uint64_t* featmask, float scale,
                         const float* const* wtab, int H, int U, int I, float* dXu, float* dXi, int accumulate, cudaStream_t st) {
    const float* dhi = dh + (int64_t)U * D;
    if (H == 8) tcx::launch<8, 1>(dh, dhi, Zu, Zi, apply_elu, featmask, scale, wtab, U, I, dXu, dXi, nullptr, accumulate, st);
    else        tcx::launch<1, 1>(dh, dhi, Zu, Zi, apply_elu, featmask, scale, wtab, U, I, dXu, dXi, nullptr, accumulate, st);
}

// fused dense backward on the tensor cores; writes (*nb_u + *nb_i) CTA partials of 64*64+64 floats into `partials`
void transform_bwd_tc(const float* dh, const float* dS, const float* Xu, const float* Xi, int apply_elu, const uint64_t* featmask, float scale,
                      const float* const* wtab, int H, int U, int I, float* dXu, float* dXi, int accumulate_dx, float* partials,
                      int* nb_u, int* nb_i, cudaStream_t st) {
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(tcx::bwd::transform_bwd_tc_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcx::bwd::SMEM_BYTES);
        cudaFuncSetAttribute(tcx::bwd::transform_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tcx::bwd::SMEM_BYTES);
    });
    tcx::bwd::side_grid1(U, I, nb_u, nb_i);
    const int grid = *nb_u + *nb_i;
    if (H == 8)
        tcx::bwd::transform_bwd_tc_kernel<8><<<grid, tcx::bwd::THREADS, tcx::bwd::SMEM_BYTES, st>>>(dh, dS, Xu, Xi, apply_elu, featmask, scale, wtab, U, I, *nb_u,
                                                                                                   dXu, dXi, accumulate_dx, partials);
    else
        tcx::bwd::transform_bwd_tc_kernel<1><<<grid, tcx::bwd::THREADS, tcx::bwd::SMEM_BYTES, st>>>(dh, dS, Xu, Xi, apply_elu, featmask, scale, wtab, U, I, *nb_u,
                                                                                                   dXu, dXi, accumulate_dx, partials);
}

}  // namespace ngacf
