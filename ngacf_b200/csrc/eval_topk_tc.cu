// AllNeg evaluation, tensor-core path: users x items score contraction on the 5th-gen tensor cores
// (tcgen05.mma, bf16 hi/lo split = "bf16x3", fp32 accumulators in TMEM) with the per-row top-K fused
// on the accumulators as they are read back with tcgen05.ld -- the (users x items) score matrix never
// touches shared memory or HBM.  The kernel hands ~24 approximate candidates per user to a second
// kernel, which re-scores them exactly (fp32, specified summation tree), orders them (score desc, id asc),
// and proves with an error bound that no non-candidate can enter the top-20; rows that fail the proof
// are flagged and recomputed by the caller through the exact CUDA-core entry point.
// Replaces train_eval_Gowalla.py:300-341,370-385 of the reference.
//
// Kernel structure (one CTA = 128 users x one SEGMENT of the item tiles; two CTAs per SM overlap):
//   warps 0-3, 6-9  epilogue: eight warps; thread (quadrant, lane) owns user row r = TMEM lane r and one HALF of every tile's
//              128 columns (warps 0-3 columns 0-63, warps 6-9 columns 64-127) = one LIST.  Per 16 columns: tcgen05.ld (the next
//              piece's load in flight), threshold test (2 instructions per accumulator), allowed-column word (item pool minus
//              the user's train positives, from a bit matrix built once per evaluator), survivors appended to the list's LOG
//              in global memory.  A list keeps the KP best SCORES sorted in registers (no ids: they stay in the log); the log
//              is merged into them on a fixed tile schedule (1, 2, 4, 8, ...), the same for all eight warps, because a merge
//              stalls the CTA's whole TMA -> MMA -> epilogue chain.  The filter threshold is shared by all lists of a user
//              through gthr[user] (atomicMax; a union's KP-th best bounds every part's).  At the end the two lists of a row
//              cut their logs at the KP-th best of their union.
//   warp 4     producer: one 32 KB cp.async.bulk (TMA engine, 1-D) per item tile, pre-tiled in HBM in
//              the exact UMMA shared-memory image (K-major, no swizzle) by prep_items_kernel
//   warp 5     TMEM allocator + single-thread MMA issuer: 12 tcgen05.mma (4 k-steps x {hi.hi, hi.lo, lo.hi})
//              per tile into a double-buffered 128x128 fp32 accumulator; descriptors precomputed
// 2-D decomposition: grid = (user blocks, S segments); S grows as the user count shrinks (a rank of an 8-GPU evaluation holds
// 30 user blocks: without the split 30 CTAs would walk all 321 tiles serially on a 148-SM part).
// History of the epilogue (scripts/probe/trace_topk.py prints the pipeline timeline of a debug build): round 1 inserted every
// survivor into a sorted (score, id) list at once; round 2a buffered survivors per thread and merged when a lane's buffer filled --
// eight warps stalling the pipeline at eight different tiles, 40 % of the kernel; the train-positive walk was a chain of dependent
// global loads per tile (users with 1000 train items: +0.3 ms for their CTA).  Tried and dropped: more segments than CTA slots
// (every list's own threshold is the KP-th best of its columns only: 143 candidates per user reach the re-score kernel with ten
// lists), N = 64 MMA groups with a hand-over per column half (the single issuing thread becomes the bottleneck: +7 %).
#include <cstdlib>
#include <cuda_bf16.h>
#include "common.cuh"

namespace ngacf {
namespace tc {

constexpr int TM = 128;                 // users per CTA (UMMA M)
constexpr int TN = 128;                 // items per tile (UMMA N)
constexpr int KP = 24;                  // approximate candidates kept per LIST (user x segment x column half): > 20, so that a list
                                        // holding the whole top-20 still has its threshold below the 20th exact score
constexpr int KU = 32;                  // rank of the shared filter threshold inside the union of a row's two sorted lists (KP..2 KP).  It is also
                                        // the rank of the final candidate cut, i.e. the proof's margin: the K-th EXACT score must clear the KU-th
                                        // approximate one by the bf16x3 error bound.  KU = KP = 24 is 5 % faster (0.72 vs 0.76 ms) but leaves four
                                        // ranks of margin: 1 of 52,639 amazon-book-shape users failed the proof -- and one fallback row costs 6 ms
                                        // in the exact kernel.  With 12 ranks the failure probability falls by ~(delta/gap)^8 (none observed).
constexpr int CBUF = 384;               // survivor LOG entries per epilogue thread (append-only; ~150 used on the gowalla shape; overflow -> exact fallback)
constexpr int EPI = 256;                // epilogue threads per CTA
constexpr int MAX_LISTS = 8;            // 2 * S <= 8 lists per user (plan_topk: at most four segments); sizes the re-score kernel's shared arrays
constexpr int K = NGACF_TOPK;
constexpr int PANEL = TN * 16;          // bytes of one k-chunk panel: 128 rows x 16 B
constexpr int HALF_BYTES = 8 * PANEL;   // 16 KB: one operand half (hi or lo), 8 k-chunks
constexpr int TILE_BYTES = 2 * HALF_BYTES;   // 32 KB per item tile (hi + lo)
constexpr int THREADS = 320;
constexpr int STAGES = 2;                // item-tile ring in shared memory
constexpr size_t SMEM_BYTES = 2 * HALF_BYTES /*A*/ + STAGES * TILE_BYTES /*B ring*/ + 16 * EPI * 4 /*score staging*/ + 128 /*barriers*/ + 128 /*align*/;
// the survivor-log pool (resident_ctas()) holds two buffers per SM: a third resident CTA would spin for a buffer forever
static_assert(3 * SMEM_BYTES > 228 * 1024, "at most two CTAs of score_topk_tc_kernel may be resident per SM");
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

__device__ __forceinline__ unsigned int smid_u32() { unsigned int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }

// order-preserving float <-> uint (atomicMax on thresholds); 0 decodes to -inf, so a zero-filled array is "no threshold yet"
__device__ __forceinline__ unsigned int thr_encode(float f) {
    const unsigned int b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float thr_decode(unsigned int e) {
    if (e == 0u) return -INFINITY;
    return __uint_as_float((e & 0x80000000u) ? (e & 0x7fffffffu) : ~e);
}

// split form for software pipelining: the registers of an issued load are only defined after tmem_ld_wait() -- which also names
// them, so that no use can be scheduled above the wait
__device__ __forceinline__ void tmem_ld16_issue(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]),
                   "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :: "memory");
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
// allowed-column mask matrix [tile][user slot][half] (uint64 each): pool bits of the tile, minus the user's train positives.
// Depends on (users, train set, pool) only: built once per evaluator, reused by every evaluation (reuse_mask).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) mask_fill_kernel(const uint32_t* __restrict__ pool_bits, int n_tiles, int n_slots, int n_users,
                                                        uint4* __restrict__ amask) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // (tile, user slot)
    if (idx >= (int64_t)n_tiles * n_slots) return;
    const int tile = (int)(idx / n_slots), slot = (int)(idx % n_slots);
    uint4 w = make_uint4(0u, 0u, 0u, 0u);
    if (slot < n_users) w = __ldg(reinterpret_cast<const uint4*>(pool_bits) + tile);
    amask[idx] = w;
}
__global__ void __launch_bounds__(256) mask_clear_train_kernel(const int* __restrict__ users, int n_users, const int* __restrict__ train_ptr,
                                                               const int* __restrict__ train_items, int n_slots, uint32_t* __restrict__ amask) {
    const int slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;           // one warp per user
    if (slot >= n_users) return;
    const int lane = threadIdx.x & 31;
    const int u = users[slot];
    for (int e = train_ptr[u] + lane; e < train_ptr[u + 1]; e += 32) {
        const int item = train_items[e];
        atomicAnd(amask + ((size_t)(item / TN) * n_slots + slot) * 4 + ((item % TN) >> 5), ~(1u << (item & 31)));
    }
}

#ifdef NGACF_TOPK_TRACE
// pipeline timeline of two CTAs (debug builds only): [cta slot][event][tile]
__device__ long long g_topk_trace[2][8][512];
__device__ long long g_topk_cta[4096][4];
__device__ unsigned long long g_rescore_stat[4];   // sum of kept candidates, sum of listed (id >= 0) candidates, users, max kept      // per CTA: globaltimer at start / end of the epilogue of warp 0, SM id, log entries of (warp 0, lane 0)
__device__ __forceinline__ long long gtimer() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
__device__ __forceinline__ int smid() { int r; asm volatile("mov.u32 %0, %%smid;" : "=r"(r)); return r; }
#define TRACE(ev, lt) do { if ((blockIdx.x == 0 || blockIdx.x == 200) && blockIdx.y == 0 && (lt) < 512) g_topk_trace[blockIdx.x ? 1 : 0][ev][lt] = clock64(); } while (0)
#else
#define TRACE(ev, lt) do { } while (0)
#endif

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(THREADS, 2) score_topk_tc_kernel(const float* __restrict__ F, int U, const int* __restrict__ users,
                                                                   int n_users, const uint2* __restrict__ amask,
                                                                   const uint8_t* __restrict__ img, int tile_begin, int tile_count, int S,
                                                                   int list_base, int lists_per_user,
                                                                   uint2* __restrict__ gbuf, int* buf_locks, int n_bufs, unsigned int* gthr,
                                                                   int* __restrict__ cand_ids, float* __restrict__ cand_sc,
                                                                   float* __restrict__ cand_thr) {
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
    const int t0 = tile_begin + (int)((int64_t)seg * tile_count / S), t1 = tile_begin + (int)((int64_t)(seg + 1) * tile_count / S);   // this CTA's item tiles
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
    if (tid == 32) {
        // survivor-log buffer of this CTA: one of n_bufs (= resident CTA slots of the device), claimed for the CTA's lifetime.  The
        // logs are only read by their own CTA, so the workspace holds one per RESIDENT CTA, not one per (user block, segment).
        int b = (int)((smid_u32() * 2u) % (unsigned)n_bufs);
        while (atomicCAS(buf_locks + b, 0, 1) != 0) b = b + 1 == n_bufs ? 0 : b + 1;
        tmem_slot[1] = (uint32_t)b;
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
                TRACE(0, lt);
                mbar_expect_tx(bar_full0 + 8 * st, TILE_BYTES);
                bulk_g2s(smem_u32(sB) + st * TILE_BYTES, img + (size_t)(t0 + lt) * TILE_BYTES, TILE_BYTES, bar_full0 + 8 * st);
            }
        }
    } else if (warp == 5) {
        // ================= MMA issuer =================
        if (lane == 0) {
            // all 24 operand descriptors are loop invariants (two ring slots): built once, the issue loop is then 12 back-to-back
            // MMAs -- it was ~90 cycles per MMA of descriptor arithmetic, longer than the MMA itself, and sits on the critical path
            // accumulator drained -> MMAs issued -> accumulator full -> epilogue
            uint64_t dA[2][4], dB[2][2][4];
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
                dA[0][ks] = smem_desc(smem_u32(sA) + ks * 2 * PANEL);
                dA[1][ks] = smem_desc(smem_u32(sA) + HALF_BYTES + ks * 2 * PANEL);
#pragma unroll
                for (int b = 0; b < 2; ++b) {
                    dB[b][0][ks] = smem_desc(smem_u32(sB) + b * TILE_BYTES + ks * 2 * PANEL);
                    dB[b][1][ks] = smem_desc(smem_u32(sB) + b * TILE_BYTES + HALF_BYTES + ks * 2 * PANEL);
                }
            }
            auto issue = [&](int lt, const uint64_t (&bd)[2][4], int buf) {
                mbar_wait(bar_full0 + 8 * buf, (uint32_t)((lt >> 1) & 1));                 // tile landed in its ring slot
                TRACE(1, lt);
                mbar_wait(bar_tempty0 + 8 * buf, (uint32_t)(((lt >> 1) & 1) ^ 1));         // accumulator drained by the epilogue
                TRACE(2, lt);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t d = tmem_base + (uint32_t)(buf * TN);
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(d, dA[0][ks], bd[0][ks], ks ? 1u : 0u);      // hi . hi   (K = 16 per MMA)
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(d, dA[0][ks], bd[1][ks], 1u);                // hi . lo
#pragma unroll
                for (int ks = 0; ks < 4; ++ks) umma_bf16(d, dA[1][ks], bd[0][ks], 1u);                // lo . hi
                umma_commit(bar_empty0 + 8 * buf);            // ring slot reusable once these MMAs retire
                umma_commit(bar_tfull0 + 8 * buf);            // accumulator ready for the epilogue
                TRACE(3, lt);
            };
            for (int lt = 0; lt < nt; lt += 2) {
                issue(lt, dB[0], 0);
                if (lt + 1 < nt) issue(lt + 1, dB[1], 1);
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
        // The KP best SCORES so far, sorted descending, entirely in registers; thr = ls[KP-1] filters the accumulators.  The item ids
        // are not carried through the sort: every survivor is appended to this thread's LOG in global memory (append-only,
        // interleaved over the CTA's epilogue threads: a warp's entries of one index share a 256-byte segment), and the final
        // candidates are the log entries with score >= the final threshold.
        float ls[KP];
#pragma unroll
        for (int k = 0; k < KP; ++k) ls[k] = -INFINITY;
        float thr = -INFINITY;                                // == ls[KP-1]
        // FILTER threshold: the maximum of the thr of ALL lists of this user -- both column halves, every item segment, whichever
        // CTA runs them -- kept in gthr[user slot] (atomicMax on an order-preserving encoding).  The KP-th best of a union is >= the
        // KP-th best of any part, so every list's thr is a valid bound for all of them; stale reads are valid too (it only grows).
        // A user's lists together log ~24 ln(I/24) survivors instead of that per list, and a CTA that starts late (second wave,
        // other segment) starts with a warm threshold -- which is what makes the segment split cheap.
        unsigned int* gT = gthr + uslot;
        float T = thr_decode(__ldcg(gT));                     // a CTA that starts late starts warm
        unsigned int tpre = 0u;                               // gthr value requested one tile ahead
        uint2* bE = gbuf + (size_t)tmem_slot[1] * (CBUF * EPI) + et;
        int cnt = 0, done = 0, overflow = 0;                  // log entries written / already merged into ls
#ifdef NGACF_TOPK_TRACE
        int n_flush = 0;
        const long long t_begin = gtimer();
#endif
        // merge the log entries [done, cnt) into the sorted scores (warp-collective).  Runs on a FIXED tile schedule, the same for
        // all eight epilogue warps of the CTA: a merge takes thousands of cycles during which the warp does not drain its part of
        // the accumulator, i.e. the whole CTA pipeline waits -- with per-warp triggers (a lane's buffer filling up) the eight warps
        // stalled the pipeline at eight different tiles per round (measured: 40 % of the kernel, scripts/probe/trace_topk.py).
        auto flush = [&]() {
#ifdef NGACF_TOPK_TRACE
            ++n_flush;
#endif
            // every lane walks ITS OWN new entries (the logs of the 32 lanes drift apart by dozens of entries: a common index range
            // was three times longer than any lane's share), sixteen loads in flight (the entries were written long ago: L2 latency)
            const int nnew = cnt - done;
            const int nmax = __reduce_max_sync(0xffffffffu, nnew);
#pragma unroll 1
            for (int i0 = 0; i0 < nmax; i0 += 16) {
                float sb[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) sb[q] = i0 + q < nnew ? __uint_as_float(bE[(size_t)(done + i0 + q) * EPI].x) : -INFINITY;
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float sc = sb[q];
                    if (__any_sync(0xffffffffu, sc > thr)) {              // most logged entries are below the threshold by now
                        // sorted insertion without predicates: ls'[k] = max(ls[k], min(ls[k-1], sc)) -- unchanged where sc <= ls[k],
                        // sc where ls[k] < sc <= ls[k-1], the shifted ls[k-1] above; a lane whose sc <= thr changes nothing
#pragma unroll
                        for (int k = KP - 1; k > 0; --k) ls[k] = fmaxf(ls[k], fminf(ls[k - 1], sc));
                        ls[0] = fmaxf(ls[0], sc);
                        thr = ls[KP - 1];
                    }
                }
            }
            done = cnt;
        };
        // After a scheduled merge the two lists of a user row (same CTA, warps w and w + 6 - ...) swap their KP sorted scores through
        // shared memory and both take the KU-th best of the UNION as filter threshold: max(thr0, thr1) is the
        // ~2 KP-th best of the row -- about half the survivors to log and merge afterwards.  All eight epilogue
        // warps run this at the same tiles (named barrier 1).
        auto exchange = [&]() {
            // through the staging area: a thread owns 16 floats there (nobody else touches them), so the KP scores go in two rounds;
            // the second barrier of a round keeps the partner from overwriting what is still being read
            float b[KP];
            static_assert(KP <= 32 && KP > 16, "two rounds of 16 floats");
#pragma unroll
            for (int r0 = 0; r0 < KP; r0 += 16) {
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    if (r0 + 4 * q4 < KP) Vs4[q4 * EPI + et] = make_float4(ls[r0 + 4 * q4], ls[r0 + 4 * q4 + 1], ls[r0 + 4 * q4 + 2], ls[r0 + 4 * q4 + 3]);
                asm volatile("bar.sync 1, 256;" ::: "memory");
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    if (r0 + 4 * q4 < KP) {
                        const float4 o = Vs4[q4 * EPI + (et ^ TM)];
                        b[r0 + 4 * q4] = o.x; b[r0 + 4 * q4 + 1] = o.y; b[r0 + 4 * q4 + 2] = o.z; b[r0 + 4 * q4 + 3] = o.w;
                    }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
            // KU-th of the union of two sorted lists = max_i min(a[i-1], b[KU-i-1]), i scores from this list and KU - i from the
            // other (both <= KP: the lists are truncated, which can only lower the value -- still a valid bound)
            float t_union = -INFINITY;
#pragma unroll
            for (int i = (KU > KP ? KU - KP : 0); i <= (KP < KU ? KP : KU); ++i)
                t_union = fmaxf(t_union, fminf(i > 0 ? ls[i > 0 ? i - 1 : 0] : INFINITY, i < KU ? b[i < KU ? KU - i - 1 : 0] : INFINITY));
            // (own thr as well: T >= thr keeps "at most KP log entries above T" true for the final cut; with evenly mixed halves
            // thr is the row's ~2 KP-th best, below t_union)
            t_union = fmaxf(t_union, thr);
            T = fmaxf(T, t_union);
            if (t_union > -INFINITY) T = fmaxf(T, thr_decode(atomicMax(gT, thr_encode(t_union))));
        };
        // allowed columns (item pool minus this user's train positives) of this thread's 64 columns: one 8-byte word per tile from
        // the mask matrix [tile][user slot][half] (coalesced over the warp), requested one tile ahead.  Rows past n_users are zero.
        const size_t mstride = (size_t)gridDim.x * TM * 2;
        const uint2* mrow = amask + (size_t)uslot * 2 + half;
        uint2 mnext = nt > 0 ? __ldg(mrow + (size_t)t0 * mstride) : make_uint2(0u, 0u);
        for (int lt = 0; lt < nt; ++lt) {
            const int buf = lt & 1;
            const int item0 = (t0 + lt) * TN + half * 64;     // first item of this thread's 64 columns
            const unsigned pw0 = mnext.x, pw1 = mnext.y;
            if (lt + 1 < nt) mnext = __ldg(mrow + (size_t)(t0 + lt + 1) * mstride);
            T = fmaxf(T, thr_decode(tpre));                   // what the user's other lists have reached meanwhile
            tpre = __ldcg(gT);
            // merge schedule: tiles 1, 2, 4, 8, 16, ...: a merge costs ~10k cycles of a stalled CTA pipeline, a survivor logged
            // because the threshold is stale ~0.3k; doubling intervals are near the optimum of that trade
            // -- and only when a lane has enough to merge: a list that runs under a warm shared threshold logs a handful of survivors
            if (lt > 0 && (lt & (lt - 1)) == 0) {
                if (__reduce_max_sync(0xffffffffu, cnt - done) >= 8) flush();
                if (warp == 0 && lane == 0) TRACE(6, lt);
                exchange();
                if (warp == 0 && lane == 0) TRACE(7, lt);
            }
            mbar_wait(bar_tfull0 + 8 * buf, (uint32_t)((lt >> 1) & 1));
            if (warp == 0 && lane == 0) TRACE(4, lt);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            // four 16-column pieces of this thread's 64 columns, software-pipelined: the tcgen05.ld of piece pc+1 is in flight while
            // piece pc is filtered (the epilogue is a latency chain per warp, and the CTA pipeline waits for its slowest warp)
            const uint32_t taddr0 = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(buf * TN + half * 64);
            uint32_t va[16], vb[16];
            tmem_ld16_issue(taddr0, va);
            auto piece = [&](const uint32_t (&v)[16], int pc) {
                // threshold filter: two instructions per accumulator (FSETP, predicated OR of the column's bit), four short chains
                unsigned h0 = 0, h1 = 0, h2 = 0, h3 = 0;
#define NGACF_HIT(h, j, bit) asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p or.b32 %0, %0, " #bit ";\n\t}" : "+r"(h) : "f"(__uint_as_float(v[j])), "f"(T))
                NGACF_HIT(h0, 0, 0x1); NGACF_HIT(h1, 4, 0x10); NGACF_HIT(h2, 8, 0x100); NGACF_HIT(h3, 12, 0x1000);
                NGACF_HIT(h0, 1, 0x2); NGACF_HIT(h1, 5, 0x20); NGACF_HIT(h2, 9, 0x200); NGACF_HIT(h3, 13, 0x2000);
                NGACF_HIT(h0, 2, 0x4); NGACF_HIT(h1, 6, 0x40); NGACF_HIT(h2, 10, 0x400); NGACF_HIT(h3, 14, 0x4000);
                NGACF_HIT(h0, 3, 0x8); NGACF_HIT(h1, 7, 0x80); NGACF_HIT(h2, 11, 0x800); NGACF_HIT(h3, 15, 0x8000);
#undef NGACF_HIT
                unsigned hit = (h0 | h1) | (h2 | h3);
                hit &= ((pc < 2 ? pw0 : pw1) >> ((pc & 1) * 16)) & 0xFFFFu;      // item pool minus this user's train positives
                if (__any_sync(0xffffffffu, hit != 0)) {
                    // the survivors are picked up from a shared-memory copy (own slots only: thread stride 16 B, conflict-free), so
                    // that the append code exists once instead of 16 times and no register is indexed dynamically
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        Vs4[q4 * EPI + et] = make_float4(__uint_as_float(v[4 * q4]), __uint_as_float(v[4 * q4 + 1]), __uint_as_float(v[4 * q4 + 2]),
                                                         __uint_as_float(v[4 * q4 + 3]));
                    while (hit) {                             // lanes append their own survivors in parallel
                        const int j = __ffs(hit) - 1;
                        hit &= hit - 1;
                        if (cnt < CBUF) {
                            bE[(size_t)cnt * EPI] = make_uint2(reinterpret_cast<const uint32_t*>(Vs4 + (j >> 2) * EPI + et)[j & 3], (uint32_t)(item0 + pc * 16 + j));
                            ++cnt;
                        } else {
                            overflow = 1;                     // pathological score order: this list goes to the exact fallback
                        }
                    }
                }
            };
            tmem_ld_wait(va);
            tmem_ld16_issue(taddr0 + 16, vb);
            piece(va, 0);
            tmem_ld_wait(vb);
            tmem_ld16_issue(taddr0 + 32, va);
            piece(vb, 1);
            tmem_ld_wait(va);
            tmem_ld16_issue(taddr0 + 48, vb);
            piece(va, 2);
            tmem_ld_wait(vb);
            piece(vb, 3);
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty0 + 8 * buf);
            if (warp == 0 && lane == 0) TRACE(5, lt);
        }
        // final threshold of this CTA's two lists of a user: the KU-th best of the UNION of their sorted scores (at least KU
        // items reach it)
        flush();
        exchange();
        T = fmaxf(T, thr_decode(__ldcg(gT)));
        // candidates = the logged survivors at or above the final (shared) threshold: at most KP per list, since the entries above
        // T >= thr are part of this list's own top KP (ties beyond KP stay out: they are <= cand_thr)
        {
            const size_t list = (size_t)(user >= 0 ? uslot : 0) * lists_per_user + list_base + seg * 2 + half;
            const int e_hi = __reduce_max_sync(0xffffffffu, cnt);
            int emitted = 0, tie_left = KP;               // entries equal to the threshold may only fill what the larger ones leave
#pragma unroll
            for (int k = 0; k < KP; ++k) tie_left -= ls[k] > T ? 1 : 0;
#pragma unroll 1
            for (int e0 = 0; e0 < e_hi; e0 += 16) {           // sixteen log entries in flight (it was one: a chain of L2 round trips)
                uint2 ent[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) ent[q] = e0 + q < cnt ? bE[(size_t)(e0 + q) * EPI] : make_uint2(0xff800000u, 0xffffffffu);
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                    const float sc = __uint_as_float(ent[q].x);
                    const bool tie = sc == T && tie_left > 0;
                    if (user >= 0 && emitted < KP && (sc > T || tie)) {       // padding entries are -inf: never above T, a tie only
                        cand_sc[list * KP + emitted] = sc;                    // while T = -inf, where their id (-1) is padding too
                        cand_ids[list * KP + emitted++] = (int)ent[q].y;
                        tie_left -= tie ? 1 : 0;
                    }
                }
            }
            if (user >= 0) {
                for (int k = emitted; k < KP; ++k) cand_ids[list * KP + k] = -1;
                cand_thr[list] = overflow ? INFINITY : T;     // neither half's list full: T = -inf (there is no non-candidate)
            }
        }
#ifdef NGACF_TOPK_TRACE
        if (warp == 0 && lane == 0 && blockIdx.y * gridDim.x + blockIdx.x < 4096) {
            const int ci = blockIdx.y * gridDim.x + blockIdx.x;
            g_topk_cta[ci][0] = t_begin; g_topk_cta[ci][1] = gtimer(); g_topk_cta[ci][2] = smid(); g_topk_cta[ci][3] = cnt;
        }
#endif
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid == 32) atomicExch(buf_locks + (int)tmem_slot[1], 0);      // every log read of this CTA is behind the barrier
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
constexpr int RS_GROUPS = 4;            // users per CTA (64 threads; 6 KB of candidate scores / ids)
__global__ void __launch_bounds__(RS_GROUPS * 16) rescore_kernel(const float* __restrict__ F, int U, const int* __restrict__ users, int n_users,
                                                                const int* __restrict__ cand_ids, const float* __restrict__ cand_sc,
                                                                const float* __restrict__ cand_thr, const unsigned int* __restrict__ gthr,
                                                                int n_lists, const unsigned int* __restrict__ maxnorm_bits, int* __restrict__ top_ids,
                                                                float* __restrict__ top_scores, int* __restrict__ fallback) {
    __shared__ float sc_s[RS_GROUPS][MAX_LISTS * KP];
    __shared__ int id_s[RS_GROUPS][MAX_LISTS * KP];
    const int grp = threadIdx.x >> 4;
    const int j = blockIdx.x * RS_GROUPS + grp;
    if (j >= n_users) return;                                  // whole groups leave together; no block-level barrier below
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int64_t u = users[j];
    const float4 fu = ld_gather4(F + u * D + lane16 * 4);
    const float unorm = sqrtf(dot64_tree_g(fu, fu, gm));
    float* scs = sc_s[grp];
    int* ids = id_s[grp];
    // Candidates = the listed items whose approximate score reaches the user's FINAL shared threshold.  A list that finished early
    // emitted against the threshold of its time -- up to KP entries each, most of them far below the final one; everything not
    // kept here (never logged, logged but not emitted, emitted below the final threshold) has approx <= T_final.  The lists are
    // mostly padding (-1) as well: compact first (the ranking below is quadratic in the count).
    const float T_final = thr_decode(gthr[j]);
    const int NC_all = n_lists * KP;
    int NC = 0;
    const int gshift = (threadIdx.x & 16);
    for (int c0 = 0; c0 < NC_all; c0 += 16) {
        int id = c0 + lane16 < NC_all ? cand_ids[(int64_t)j * NC_all + c0 + lane16] : -1;
        if (id >= 0 && cand_sc[(int64_t)j * NC_all + c0 + lane16] < T_final) id = -1;
        const unsigned m = (__ballot_sync(gm, id >= 0) >> gshift) & 0xFFFFu;
        if (id >= 0) ids[NC + __popc(m & ((1u << lane16) - 1u))] = id;
        NC += __popc(m);
    }
    __syncwarp(gm);
#ifdef NGACF_TOPK_TRACE
    if (lane16 == 0) { atomicAdd(&g_rescore_stat[0], (unsigned long long)NC); atomicAdd(&g_rescore_stat[2], 1ull); }
#endif
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
    float thr = T_final;                                         // a list's cand_thr is <= T_final, or +inf if its log overflowed
    for (int l = lane16; l < n_lists; l += 16) thr = fmaxf(thr, cand_thr[(int64_t)j * n_lists + l]);
#pragma unroll
    for (int o = 1; o < 16; o <<= 1) thr = fmaxf(thr, __shfl_xor_sync(gm, thr, o, 16));
    const float delta = GUARD * unorm * __uint_as_float(*maxnorm_bits);
    const bool ok = (thr == -INFINITY) || (nvalid >= K && thr + delta < tau);
#ifdef NGACF_TOPK_TRACE
    if (lane16 == 0 && !ok) {
        if (nvalid < K) atomicAdd(&g_rescore_stat[1], 1ull); else atomicAdd(&g_rescore_stat[3], 1ull);
        if (j < 4) printf("user slot %d: NC %d nvalid %d T_final %g thr %g tau %g delta %g\n", j, NC, nvalid, T_final, thr, tau, delta);
    }
#endif
    if (lane16 == 0) {
        fallback[j] = ok ? 0 : 1;
        if (!ok) atomicAdd(fallback + n_users, 1);              // number of flagged rows: the caller reads one int instead of scanning
    }
}

}  // namespace tc
}  // namespace ngacf

using namespace ngacf;

static inline size_t al256(size_t x) { return (x + 255) / 256 * 256; }

// item-tile segments per user block, for the few user blocks of a rank of a sharded evaluation: about one CTA per SM, at most four
// segments.  Filling all 2 x 148 CTA slots does not pay: a user's lists share their filter threshold (gthr), but each list's own
// threshold is the KP-th best of ITS columns only, so the shared value loosens with the number of lists (143 candidates per user reach
// the re-score kernel with ten lists, 43 with four), and every CTA pays its pipeline fill and merges again.  Measured on the shard of
// an 8-GPU evaluation (scripts/probe/shard_eval.py): 30 user blocks S = 9 -> 0.45 ms, S = 4 -> 0.29 ms; 52 blocks S = 5 -> 0.60,
// S = 2 -> 0.58.
// survivor-log buffers = CTA slots that can be resident at once (two CTAs per SM: shared memory)
static int resident_ctas() {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return 2 * (sms > 0 ? sms : 148);
}

struct TopkPlan { int S, lists_per_user; };
static TopkPlan plan_topk(int n_users, int n_tiles) {
    TopkPlan p;
    const int blocks = (n_users + tc::TM - 1) / tc::TM;
    int S = blocks > 0 ? 128 / blocks : 1;
    if (S > 4) S = 4;
    if (const char* v = getenv("NGACF_TOPK_SEGMENTS")) { if (*v) S = atoi(v); }     // tuning knob (scripts/probe/shard_eval.py)
    if (S > n_tiles / 16) S = n_tiles / 16;
    if (S < 1) S = 1;
    if (S > tc::MAX_LISTS / 2) S = tc::MAX_LISTS / 2;
    p.S = S;
    p.lists_per_user = 2 * S;
    return p;
}

#ifdef NGACF_TOPK_TRACE
extern "C" int ngacf_debug_topk_trace(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, tc::g_topk_trace, sizeof(tc::g_topk_trace));
}
extern "C" int ngacf_debug_rescore_stat(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, tc::g_rescore_stat, sizeof(tc::g_rescore_stat));
}
extern "C" int ngacf_debug_topk_cta(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, tc::g_topk_cta, sizeof(tc::g_topk_cta));
}
#endif

extern "C" size_t ngacf_score_topk_tc_workspace_bytes(int32_t I, int32_t n_users) {
    const size_t n_tiles = (size_t)(I + tc::TN - 1) / tc::TN;
    const TopkPlan plan = plan_topk(n_users, (int)n_tiles);
    const size_t lists = (size_t)n_users * plan.lists_per_user;
    const size_t slots = (size_t)((n_users + tc::TM - 1) / tc::TM) * tc::TM;
    return al256(n_tiles * tc::TILE_BYTES) + al256(n_tiles * 4 * 4) + 256 + 2 * al256(lists * tc::KP * 4) + al256(lists * 4) +
           al256((size_t)resident_ctas() * tc::CBUF * tc::EPI * 8) + al256((size_t)resident_ctas() * 4) + al256(slots * 4) +
           al256(n_tiles * slots * 16) + 1024;
}

extern "C" int ngacf_score_topk_tc(const float* F, int32_t U, int32_t I, const int32_t* users, int32_t n_users, const int32_t* train_ptr,
                                   const int32_t* train_items, const uint8_t* in_pool, int32_t* top_ids, float* top_scores,
                                   int32_t* fallback, int32_t reuse_mask, void* workspace, size_t workspace_bytes, void* stream) {
    NGACF_REQUIRE(F && users && train_ptr && train_items && in_pool && top_ids && top_scores && fallback && workspace && U > 0 && I > 0,
                  "score_topk_tc: null/empty argument");
    if (workspace_bytes < ngacf_score_topk_tc_workspace_bytes(I, n_users)) { set_error("score_topk_tc: workspace too small"); return NGACF_ERR_WORKSPACE; }
    if (n_users == 0) return NGACF_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int n_tiles = (I + tc::TN - 1) / tc::TN;
    const TopkPlan plan = plan_topk(n_users, n_tiles);
    const int blocks = ceil_div(n_users, tc::TM);
    const size_t lists = (size_t)n_users * plan.lists_per_user;
    char* w = (char*)(((uintptr_t)workspace + 255) & ~(uintptr_t)255);
    uint8_t* img = (uint8_t*)w;                         w += al256((size_t)n_tiles * tc::TILE_BYTES);
    uint32_t* pool_bits = (uint32_t*)w;                 w += al256((size_t)n_tiles * 4 * 4);
    unsigned int* maxnorm = (unsigned int*)w;           w += 256;
    int* cand_ids = (int*)w;                            w += al256(lists * tc::KP * 4);
    float* cand_sc = (float*)w;                         w += al256(lists * tc::KP * 4);
    float* cand_thr = (float*)w;                        w += al256(lists * 4);
    const int n_bufs = resident_ctas();
    uint2* gbuf = (uint2*)w;                            w += al256((size_t)n_bufs * tc::CBUF * tc::EPI * 8);
    int* buf_locks = (int*)w;                           w += al256((size_t)n_bufs * 4);
    unsigned int* gthr = (unsigned int*)w;              w += al256((size_t)blocks * tc::TM * 4);
    uint4* amask = (uint4*)w;                           // [tile][user slot]: persists in the caller's workspace between calls
    const int n_slots = blocks * tc::TM;
    cudaMemsetAsync(pool_bits, 0, (size_t)n_tiles * 4 * 4 + 256 + 256, st);     // pool bits + max norm (contiguous)
    cudaMemsetAsync(fallback + n_users, 0, sizeof(int32_t), st);
    cudaMemsetAsync(buf_locks, 0, al256((size_t)n_bufs * 4) + (size_t)blocks * tc::TM * 4, st);   // free buffers; "no threshold yet" for every user
    tc::prep_items_kernel<<<ceil_div((int64_t)n_tiles * tc::TN * 8, 256), 256, 0, st>>>(F, U, I, in_pool, img, pool_bits, maxnorm, n_tiles);
    if (!reuse_mask) {
        tc::mask_fill_kernel<<<ceil_div((int64_t)n_tiles * n_slots, 256), 256, 0, st>>>(pool_bits, n_tiles, n_slots, n_users, amask);
        tc::mask_clear_train_kernel<<<ceil_div((int64_t)n_users * 32, 256), 256, 0, st>>>(users, n_users, train_ptr, train_items, n_slots,
                                                                                          reinterpret_cast<uint32_t*>(amask));
    }
    static PerDeviceOnce once;
    once.run([] {
        cudaFuncSetAttribute(tc::score_topk_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc::SMEM_BYTES);
    });
    tc::score_topk_tc_kernel<<<dim3(blocks, plan.S), tc::THREADS, tc::SMEM_BYTES, st>>>(F, U, users, n_users, reinterpret_cast<const uint2*>(amask), img,
                                                                                       0, n_tiles, plan.S, 0, plan.lists_per_user, gbuf, buf_locks,
                                                                                       n_bufs, gthr, cand_ids, cand_sc, cand_thr);
    tc::rescore_kernel<<<ceil_div(n_users, tc::RS_GROUPS), tc::RS_GROUPS * 16, 0, st>>>(F, U, users, n_users, cand_ids, cand_sc, cand_thr, gthr, plan.lists_per_user, maxnorm,
                                                                                       top_ids, top_scores, fallback);
    return check_launch("score_topk_tc");
}
