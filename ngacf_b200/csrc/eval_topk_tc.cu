// AllNeg evaluation, tensor-core path: users x items score contraction on the 5th-gen tensor cores
// (tcgen05.mma, bf16 hi/lo split = "bf16x3", fp32 accumulators in TMEM) with the per-row top-K fused
// on the accumulators as they are read back with tcgen05.ld -- the (users x items) score matrix never
// touches shared memory or HBM.  The kernel keeps K'=32 approximate candidates per user; a second
// kernel re-scores them exactly (fp32, specified summation tree), orders them (score desc, id asc),
// and proves with an error bound that no non-candidate can enter the top-20; rows that fail the proof
// are flagged and recomputed by the caller through the exact CUDA-core entry point.
// Replaces train_eval_Gowalla.py:300-341,370-385 of the reference.
//
// Kernel structure (one CTA = 128 users x one SEGMENT of the item tiles; two CTAs per SM overlap):
//   warps 0-3, 6-9  epilogue: eight warps; thread (quadrant, lane) owns user row r = TMEM lane r and one HALF of every tile's
//              128 columns (warps 0-3 columns 0-63, warps 6-9 columns 64-127); tcgen05.ld 16 columns at a time, candidate
//              mask (item pool minus train positives), threshold test; survivors are APPENDED to a small per-thread buffer and
//              merged into the thread's sorted 16-entry list only when a buffer fills (round 1 inserted every survivor at once:
//              the 32-step insertion was run by the whole warp for almost every 32-column chunk, ~200 of ~380 instructions)
//   warp 4     producer: one 32 KB cp.async.bulk (TMA engine, 1-D) per item tile, pre-tiled in HBM in
//              the exact UMMA shared-memory image (K-major, no swizzle) by prep_items_kernel
//   warp 5     TMEM allocator + single-thread MMA issuer: 12 tcgen05.mma (4 k-steps x {hi.hi, hi.lo, lo.hi})
//              per tile into a double-buffered 128x128 fp32 accumulator
// 2-D decomposition: grid = (user blocks, S segments); S grows as the user count shrinks (a rank of an 8-GPU evaluation holds
// 30 user blocks: without the split 30 CTAs would walk all 321 tiles serially on a 148-SM part).  Every (user, segment, column
// half) keeps its own 16 candidates and threshold; the re-score kernel merges the 2*S lists of a user.
#include <cuda_bf16.h>
#include "common.cuh"

namespace ngacf {
namespace tc {

constexpr int TM = 128;                 // users per CTA (UMMA M)
constexpr int TN = 128;                 // items per tile (UMMA N)
constexpr int KP = 24;                  // approximate candidates kept per LIST (user x segment x column half): > 20, so that a list
                                        // holding the whole top-20 still has its threshold below the 20th exact score
constexpr int CBUF = 16;                // append-buffer entries per epilogue thread
constexpr int EPI = 256;                // epilogue threads per CTA
constexpr int MAX_LISTS = 32;           // 2 * S <= 32 lists per user
constexpr int K = NGACF_TOPK;
constexpr int PANEL = TN * 16;          // bytes of one k-chunk panel: 128 rows x 16 B
constexpr int HALF_BYTES = 8 * PANEL;   // 16 KB: one operand half (hi or lo), 8 k-chunks
constexpr int TILE_BYTES = 2 * HALF_BYTES;   // 32 KB per item tile (hi + lo)
constexpr int THREADS = 320;
constexpr int STAGES = 2;                // item-tile ring in shared memory
constexpr size_t SMEM_BYTES = 2 * HALF_BYTES /*A*/ + STAGES * TILE_BYTES /*B ring*/ + 16 * EPI * 4 /*score staging*/ + 128 /*barriers*/ + 128 /*align*/;
constexpr float GUARD = 1e-4f;          // |approx - exact| <= GUARD * |u| * max|i|  (bf16x3: ~6e-5 worst case)

// instruction descriptor: D=f32, A=B=bf16, K-major both, N=128, M=128  (cute::UMMA::InstrDescriptor bit layout)
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, SWIZZLE_NONE shared-memory matrix descriptor: core matrix = 8 rows x 16 B (contiguous 128 B);
// LBO = byte distance between the two k-chunks of one MMA (= one panel), SBO = distance between 8-row groups.
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((PANEL >> 4) & 0x3FFF) << 16;     // leading byte offset
    d |= (uint64_t)((128 >> 4) & 0x3FFF) << 32;       // stride byte offset
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    return d;                                         // base_offset 0, layout_type 0 (no swizzle)
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
// try_wait with a suspend-time hint: the thread is parked in hardware until the phase completes (or the hint expires) instead of
// re-issuing the test.  Without it a third of ALL issued instructions of this kernel were the spin loops of the producer thread,
// the MMA thread and the epilogue warps (ncu source page, profiles/r2_score_topk_tc_source.txt), stealing issue slots from the
// epilogue warps that share their schedulers.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}"
        ::"r"(bar), "r"(parity), "r"(1000000u) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(IDESC), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void split_bf16x8(const float (&x)[8], uint4& hi, uint4& lo) {
    __nv_bfloat16 h[8], l[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        h[k] = __float2bfloat16_rn(x[k]);
        l[k] = __float2bfloat16_rn(x[k] - __bfloat162float(h[k]));
    }
    hi = *reinterpret_cast<uint4*>(h);
    lo = *reinterpret_cast<uint4*>(l);
}

// ------------------------------------------------------------------------------------------------
// item table -> pre-tiled UMMA smem images (hi | lo) in HBM, pool bit words, max item norm
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_items_kernel(const float* __restrict__ F, int U, int I, const uint8_t* __restrict__ in_pool,
                                                         uint8_t* __restrict__ img, uint32_t* __restrict__ pool_bits,
                                                         unsigned int* __restrict__ maxnorm_bits, int n_tiles) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (item, k-chunk)
    const int64_t item = idx >> 3;
    const int c = (int)(idx & 7);
    if (item >= (int64_t)n_tiles * TN) return;
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = 0.f;
    if (item < I) {
        const float4 a = ld_stream4(F + (U + item) * D + c * 8), b = ld_stream4(F + (U + item) * D + c * 8 + 4);
        x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
    }
    uint4 hi, lo;
    split_bf16x8(x, hi, lo);
    const int64_t tile = item / TN;
    const int r = (int)(item % TN);
    uint8_t* base = img + tile * TILE_BYTES + (size_t)c * PANEL + (size_t)r * 16;
    *reinterpret_cast<uint4*>(base) = hi;
    *reinterpret_cast<uint4*>(base + HALF_BYTES) = lo;
    float ss = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) ss = fmaf(x[k], x[k], ss);
    ss += __shfl_xor_sync(0xffffffffu, ss, 1);
    ss += __shfl_xor_sync(0xffffffffu, ss, 2);
    ss += __shfl_xor_sync(0xffffffffu, ss, 4);
    const unsigned pool = (item < I && in_pool[item]) ? 1u : 0u;
    // lanes 0,8,16,24 hold four consecutive items; a 32-item word spans 8 warps -> atomicOr (bit set is order independent)
    if (c == 0) {
        atomicMax(maxnorm_bits, __float_as_uint(sqrtf(ss)));
        if (pool) atomicOr(pool_bits + (item >> 5), 1u << (item & 31));
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 2) score_topk_tc_kernel(const float* __restrict__ F, int U, const int* __restrict__ users,
                                                                   int n_users, const int* __restrict__ train_ptr,
                                                                   const int* __restrict__ train_items, const uint32_t* __restrict__ pool_bits,
                                                                   const uint8_t* __restrict__ img, int n_tiles, int S,
                                                                   float* __restrict__ gbuf_s, int* __restrict__ gbuf_i,
                                                                   int* __restrict__ cand_ids, float* __restrict__ cand_thr) {
    // no pointer arithmetic through integers here: the compiler must keep the shared address space (STS/LDS), the staging
    // stores of round 1 were generic ST.E because the base pointer had been aligned by hand through uintptr_t
    extern __shared__ __align__(128) unsigned char smem[];      // SWIZZLE_NONE operands: 16-byte alignment would do
    unsigned char* sA = smem;                                   // hi | lo, 32 KB
    unsigned char* sB = sA + 2 * HALF_BYTES;                    // STAGES x (hi | lo), 32 KB each
    float4* Vs4 = reinterpret_cast<float4*>(sB + STAGES * TILE_BYTES);   // [4 column quads][EPI threads] x float4: one accumulator piece
    uint64_t* bars = reinterpret_cast<uint64_t*>(Vs4 + 4 * EPI); // full[2], empty[2], tmem_full[2], tmem_empty[2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
    const uint32_t bar_full0 = smem_u32(bars + 0), bar_empty0 = smem_u32(bars + 2);
    const uint32_t bar_tfull0 = smem_u32(bars + 4), bar_tempty0 = smem_u32(bars + 6);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int u0 = blockIdx.x * TM;
    const int seg = blockIdx.y;
    const int t0 = (int)((int64_t)seg * n_tiles / S), t1 = (int)((int64_t)(seg + 1) * n_tiles / S);   // this CTA's item tiles
    const int nt = t1 - t0;

    // ---- one-time setup: user rows -> bf16 hi/lo UMMA image (generic-proxy stores) ----
    for (int idx = tid; idx < TM * 8; idx += THREADS) {
        const int r = idx >> 3, c = idx & 7;
        float x[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = 0.f;
        if (u0 + r < n_users) {
            const int64_t u = users[u0 + r];
            const float4 a = ld_gather4(F + u * D + c * 8), b = ld_gather4(F + u * D + c * 8 + 4);
            x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
        }
        uint4 hi, lo;
        split_bf16x8(x, hi, lo);
        *reinterpret_cast<uint4*>(sA + (size_t)c * PANEL + (size_t)r * 16) = hi;
        *reinterpret_cast<uint4*>(sA + HALF_BYTES + (size_t)c * PANEL + (size_t)r * 16) = lo;
    }
    if (tid == 0) {
        mbar_init(bar_full0, 1);
        mbar_init(bar_full0 + 8, 1);
        mbar_init(bar_empty0, 1);
        mbar_init(bar_empty0 + 8, 1);
        mbar_init(bar_tfull0, 1);
        mbar_init(bar_tfull0 + 8, 1);
        mbar_init(bar_tempty0, 8);                   // eight epilogue warps drain an accumulator
        mbar_init(bar_tempty0 + 8, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {   // TMEM: 256 columns = two 128-column fp32 accumulators
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(256u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // A image visible to the tensor-core (async) proxy
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ================= producer =================
        if (lane == 0) {
            for (int lt = 0; lt < nt; ++lt) {
                const int st = lt & 1;
                mbar_wait(bar_empty0 + 8 * st, (uint32_t)(((lt >> 1) & 1) ^ 1));   // ring slot free (MMAs of tile lt-2 retired)
                mbar_expect_tx(bar_full0 + 8 * st, TILE_BYTES);
                bulk_g2s(smem_u32(sB) + st * TILE_BYTES, img + (size_t)(t0 + lt) * TILE_BYTES, TILE_BYTES, bar_full0 + 8 * st);
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            const uint32_t aH = smem_u32(sA), aL = aH + HALF_BYTES;
            for (int lt = 0; lt < nt; ++lt) {
                const int buf = lt & 1;
                const uint32_t bH = smem_u32(sB) + buf * TILE_BYTES, bL = bH + HALF_BYTES;
                mbar_wait(bar_full0 + 8 * buf, (uint32_t)((lt >> 1) & 1));                 // tile landed in its ring slot
                mbar_wait(bar_tempty0 + 8 * buf, (uint32_t)(((lt >> 1) & 1) ^ 1));         // accumulator drained by the epilogue
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(buf * TN);
#pragma unroll
                for (int term = 0; term < 3; ++term) {
                    const uint32_t a0 = term == 2 ? aL : aH;
                    const uint32_t b0 = term == 1 ? bL : bH;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks)            // K = 16 per MMA = two k-chunk panels
                        umma_bf16(d, smem_desc(a0 + ks * 2 * PANEL), smem_desc(b0 + ks * 2 * PANEL), (term | ks) ? 1u : 0u);
                }
                umma_commit(bar_empty0 + 8 * buf);            // ring slot reusable once these MMAs retire
                umma_commit(bar_tfull0 + 8 * buf);            // accumulator ready for the epilogue
            }
        }
    } else {
        // ================= epilogue: thread = (user row = TMEM lane, column half) =================
        const int half = warp >= 6 ? 1 : 0;
        const int quad = warp & 3;                            // TMEM lanes 32*quad .. 32*quad+31 are the ones this warp may read
        const int r = quad * 32 + lane;
        const int et = half * TM + r;                         // epilogue thread id 0..255
        const int uslot = u0 + r;
        const int user = uslot < n_users ? users[uslot] : -1;
        int cur = 0;
        const int tend = user >= 0 ? train_ptr[user + 1] : 0;
        if (user >= 0) {                                      // first train item of this user inside the segment (lower bound)
            int lo = train_ptr[user], hi = tend;
            const int first_item = t0 * TN;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (train_items[mid] < first_item) lo = mid + 1; else hi = mid;
            }
            cur = lo;
        }
        int nxt = (user >= 0 && cur < tend) ? train_items[cur] : 0x7fffffff;
        // candidate list: KP (score, id) pairs sorted by descending score, entirely in registers
        float ls[KP];
        int li[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) { ls[k] = -INFINITY; li[k] = -1; }
        float thr = -INFINITY;                                // == ls[KP-1]
        // append buffer of this thread (global memory, interleaved over the CTA's epilogue threads: coalesced, L1/L2 resident)
        const size_t cta_lin = (size_t)blockIdx.y * gridDim.x + blockIdx.x;
        float* bS = gbuf_s + cta_lin * (CBUF * EPI) + et;
        int* bI = gbuf_i + cta_lin * (CBUF * EPI) + et;
        int cnt = 0;
        auto flush = [&]() {                                  // merge the buffered survivors into the sorted list
            for (int e = 0; e < cnt; ++e) {
                const float sc = bS[e * EPI];
                const int id = bI[e * EPI];
                if (sc > thr) {
#pragma unroll
                    for (int k = KP - 1; k > 0; --k) {
                        const bool shift = ls[k - 1] < sc;
                        const bool here = ls[k] < sc;
                        li[k] = shift ? li[k - 1] : (here ? id : li[k]);
                        ls[k] = shift ? ls[k - 1] : (here ? sc : ls[k]);
                    }
                    if (ls[0] < sc) { ls[0] = sc; li[0] = id; }
                    thr = ls[KP - 1];
                }
            }
            cnt = 0;
        };
        for (int lt = 0; lt < nt; ++lt) {
            const int buf = lt & 1;
            const int item0 = (t0 + lt) * TN + half * 64;     // first item of this thread's 64 columns
            unsigned tw0 = 0, tw1 = 0;                        // train positives of this user inside those 64 columns
            const int tile_end = (t0 + lt + 1) * TN;
            while (nxt < tile_end) {                          // nxt = the user's next train item, loaded ahead of its tile
                const int off = nxt - item0;
                if (off >= 0 && off < 64) {
                    const unsigned bit = 1u << (off & 31);
                    tw0 |= off < 32 ? bit : 0u;
                    tw1 |= off >= 32 ? bit : 0u;
                }
                ++cur;
                nxt = cur < tend ? train_items[cur] : 0x7fffffff;
            }
            const unsigned pw0 = user >= 0 ? (__ldg(pool_bits + (t0 + lt) * 4 + half * 2) & ~tw0) : 0u;
            const unsigned pw1 = user >= 0 ? (__ldg(pool_bits + (t0 + lt) * 4 + half * 2 + 1) & ~tw1) : 0u;
            mbar_wait(bar_tfull0 + 8 * buf, (uint32_t)((lt >> 1) & 1));
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
            for (int pc = 0; pc < 4; ++pc) {                  // four 16-column pieces of this thread's 64 columns
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * TN + half * 64 + pc * 16), v);
                const unsigned allowed = ((pc < 2 ? pw0 : pw1) >> ((pc & 1) * 16)) & 0xFFFFu;
                // threshold filter on the registers; the survivors are picked up again from a shared-memory copy so that the
                // append code exists once instead of 16 times
                unsigned hit = 0;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)               // staging: four 16-byte stores (thread stride 16 B: conflict-free)
                    Vs4[q4 * EPI + et] = make_float4(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]), __uint_as_float(v[4 * q4 + 2]),
                                                     __uint_as_float(v[4 * q4 + 3]));
#pragma unroll
                for (int j = 0; j < 16; ++j) hit |= (__uint_as_float(v[j]) > thr ? 1u : 0u) << j;
                hit &= allowed;
                const int n = __popc(hit);                    // <= 16 = CBUF: fits after a flush
                if (__any_sync(0xffffffffu, cnt + n > CBUF)) flush();
                while (hit) {                                 // lanes append their own survivors in parallel
                    const int j = __ffs(hit) - 1;
                    hit &= hit - 1;
                    bS[cnt * EPI] = reinterpret_cast<const float*>(Vs4 + (j >> 2) * EPI + et)[j & 3];
                    bI[cnt * EPI] = item0 + pc * 16 + j;
                    ++cnt;
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty0 + 8 * buf);
        }
        flush();
        if (user >= 0) {
            const size_t list = ((size_t)uslot * S + seg) * 2 + half;
#pragma unroll
            for (int k = 0; k < KP; ++k) cand_ids[list * KP + k] = li[k];
            cand_thr[list] = li[KP - 1] >= 0 ? thr : -INFINITY;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 5) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------
// exact re-score of the candidates + order + proof
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float dot64_tree_g(float4 a, float4 b, unsigned gm) {
    float p0 = __fmul_rn(a.x, b.x), p1 = __fmul_rn(a.y, b.y), p2 = __fmul_rn(a.z, b.z), p3 = __fmul_rn(a.w, b.w);
    float v = __fadd_rn(__fadd_rn(p0, p1), __fadd_rn(p2, p3));
    v = __fadd_rn(v, __shfl_xor_sync(gm, v, 1, 16));
    v = __fadd_rn(v, __shfl_xor_sync(gm, v, 2, 16));
    v = __fadd_rn(v, __shfl_xor_sync(gm, v, 4, 16));
    v = __fadd_rn(v, __shfl_xor_sync(gm, v, 8, 16));
    return v;
}

// One 16-lane group per user: the user's 2*S candidate lists (KP entries each, disjoint item ranges) are re-scored exactly,
// ranked under (score desc, id asc), and the top-20 is accepted only if no non-candidate can beat its last entry: a
// non-candidate of list l has approx <= thr[l], hence exact <= thr[l] + delta.
constexpr int RS_GROUPS = 4;            // users per CTA (64 threads; 24 KB of candidate scores / ids)
__global__ void __launch_bounds__(RS_GROUPS * 16) rescore_kernel(const float* __restrict__ F, int U, const int* __restrict__ users, int n_users,
                                                                const int* __restrict__ cand_ids, const float* __restrict__ cand_thr, int n_lists,
                                                                const unsigned int* __restrict__ maxnorm_bits, int* __restrict__ top_ids,
                                                                float* __restrict__ top_scores, int* __restrict__ fallback) {
    __shared__ float sc_s[RS_GROUPS][MAX_LISTS * KP];
    __shared__ int id_s[RS_GROUPS][MAX_LISTS * KP];
    const int grp = threadIdx.x >> 4;
    const int j = blockIdx.x * RS_GROUPS + grp;
    if (j >= n_users) return;                                  // whole groups leave together; no block-level barrier below
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int NC = n_lists * KP;
    const int64_t u = users[j];
    const float4 fu = ld_gather4(F + u * D + lane16 * 4);
    const float unorm = sqrtf(dot64_tree_g(fu, fu, gm));
    float* scs = sc_s[grp];
    int* ids = id_s[grp];
    for (int c = lane16; c < NC; c += 16) ids[c] = cand_ids[(int64_t)j * NC + c];
    __syncwarp(gm);
    for (int c0 = 0; c0 < NC; c0 += 4) {                        // exact scores, four candidate rows in flight
        float4 fi[4];
        int id[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            id[q] = c0 + q < NC ? ids[c0 + q] : -1;
            fi[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (id[q] >= 0) fi[q] = ld_gather4(F + (int64_t)(U + id[q]) * D + lane16 * 4);     // group-uniform
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const float sv = dot64_tree_g(fu, fi[q], gm);
            if (lane16 == 0 && c0 + q < NC) scs[c0 + q] = id[q] >= 0 ? sv : -INFINITY;
        }
    }
    __syncwarp(gm);
    // rank of my candidates under (score desc, id asc); ids are distinct (the lists cover disjoint item ranges)
    int nvalid = 0;
    for (int c = 0; c < NC; ++c) nvalid += ids[c] >= 0 ? 1 : 0;
    float tau = -INFINITY;
    for (int c = lane16; c < NC; c += 16) {
        const int id = ids[c];
        if (id < 0) continue;
        const float sv = scs[c];
        int rank = 0;
        for (int o = 0; o < NC; ++o) {
            const int io = ids[o];
            const float so = scs[o];
            rank += (io >= 0 && (so > sv || (so == sv && io < id))) ? 1 : 0;
        }
        if (rank < K) { top_ids[(int64_t)j * K + rank] = id; top_scores[(int64_t)j * K + rank] = sv; }
        if (rank == K - 1) tau = sv;
    }
    for (int k = nvalid + lane16; k < K; k += 16) { top_ids[(int64_t)j * K + k] = -1; top_scores[(int64_t)j * K + k] = 0.f; }
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) tau = fmaxf(tau, __shfl_xor_sync(gm, tau, o, 16));
    float thr = -INFINITY;
    for (int l = lane16; l < n_lists; l += 16) thr = fmaxf(thr, cand_thr[(int64_t)j * n_lists + l]);
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) thr = fmaxf(thr, __shfl_xor_sync(gm, thr, o, 16));
    const float delta = GUARD * unorm * __uint_as_float(*maxnorm_bits);
    const bool ok = (thr == -INFINITY) || (nvalid >= K && thr + delta < tau);
    if (lane16 == 0) fallback[j] = ok ? 0 : 1;
}

}  // namespace tc
}  // namespace ngacf

using namespace ngacf;

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// item-tile segments per user block: fill the 2 x 148 CTA slots when there are few user blocks (a rank of a sharded evaluation)
static int plan_segments(int n_users, int n_tiles) {
    const int blocks = (n_users + tc::TM - 1) / tc::TM;
    int S = blocks > 0 ? (2 * 148) / blocks : 1;
    if (S < 1) S = 1;
    if (S > tc::MAX_LISTS / 2) S = tc::MAX_LISTS / 2;
    if (S > n_tiles) S = n_tiles > 0 ? n_tiles : 1;
    return S;
}

extern "C" size_t ngacf_score_topk_tc_workspace_bytes(int32_t I, int32_t n_users) {
    const size_t n_tiles = (size_t)(I + tc::TN - 1) / tc::TN;
    const int S = plan_segments(n_users, (int)n_tiles);
    const size_t lists = (size_t)n_users * 2 * S;
    const size_t ctas = (size_t)((n_users + tc::TM - 1) / tc::TM) * S;
    return al256(n_tiles * tc::TILE_BYTES) + al256(n_tiles * 4 * 4) + 256 + al256(lists * tc::KP * 4) + al256(lists * 4) +
           2 * al256(ctas * tc::CBUF * tc::EPI * 4) + 1024;
}

extern "C" int ngacf_score_topk_tc(const float* F, int32_t U, int32_t I, const int32_t* users, int32_t n_users, const int32_t* train_ptr,
                                   const int32_t* train_items, const uint8_t* in_pool, int32_t* top_ids, float* top_scores,
                                   int32_t* fallback, void* workspace, size_t workspace_bytes, void* stream) {
    NGACF_REQUIRE(F && users && train_ptr && train_items && in_pool && top_ids && top_scores && fallback && workspace && U > 0 && I > 0,
                  "score_topk_tc: null/empty argument");
    if (workspace_bytes < ngacf_score_topk_tc_workspace_bytes(I, n_users)) { set_error("score_topk_tc: workspace too small"); return NGACF_ERR_WORKSPACE; }
    if (n_users == 0) return NGACF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (I + tc::TN - 1) / tc::TN;
    const int S = plan_segments(n_users, n_tiles);
    const int blocks = ceil_div(n_users, tc::TM);
    const size_t lists = (size_t)n_users * 2 * S;
    const size_t ctas = (size_t)blocks * S;
    char* w = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint8_t* img = (uint8_t*)w;                         w += al256((size_t)n_tiles * tc::TILE_BYTES);
    uint32_t* pool_bits = (uint32_t*)w;                 w += al256((size_t)n_tiles * 4 * 4);
    unsigned int* maxnorm = (unsigned int*)w;           w += 256;
    int* cand_ids = (int*)w;                            w += al256(lists * tc::KP * 4);
    float* cand_thr = (float*)w;                        w += al256(lists * 4);
    float* gbuf_s = (float*)w;                          w += al256(ctas * tc::CBUF * tc::EPI * 4);
    int* gbuf_i = (int*)w;
    cudaMemsetAsync(pool_bits, 0, (size_t)n_tiles * 4 * 4 + 256 + 256, st);     // pool bits + max norm (contiguous)
    tc::prep_items_kernel<<<ceil_div((int64_t)n_tiles * tc::TN * 8, 256), 256, 0, st>>>(F, U, I, in_pool, img, pool_bits, maxnorm, n_tiles);
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(tc::score_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    });
    tc::score_topk_tc_kernel<<<dim3(blocks, S), tc::THREADS, tc::SMEM_BYTES, st>>>(F, U, users, n_users, train_ptr, train_items, pool_bits, img,
                                                                                  n_tiles, S, gbuf_s, gbuf_i, cand_ids, cand_thr);
    tc::rescore_kernel<<<ceil_div(n_users, tc::RS_GROUPS), tc::RS_GROUPS * 16, 0, st>>>(F, U, users, n_users, cand_ids, cand_thr, 2 * S, maxnorm,
                                                                                       top_ids, top_scores, fallback);
    return check_launch("score_topk_tc");
}
