// Graph builder: COO (u,i) int64 -> coalesced CSR + CSC + CSC->CSR permutation + unified node
// adjacency + degree-bucketed task list.  Replaces data/loadGowalla.py:179-186,218-219,229-253
// (scipy COO -> torch COO -> coalesce) and the per-forward COO rebuilds of SPUIGACF.py:365,377.
//
// All primitives are hand-written and deterministic: an LSD radix sort (8-bit digits, stable
// match_any ranking), a multi-level exclusive scan, boundary-detection row pointers.
#include "common.cuh"

namespace ngacf {

// ------------------------------------------------------------------------------------------------
// exclusive scan (int32), tile = 256 threads x 8
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__global__ void __launch_bounds__(SCAN_THREADS) scan_tile_kernel(const int* in, int* out /* may alias in */,
                                                                 int* __restrict__ tile_sums, int64_t n) {
    __shared__ int warp_sums[SCAN_THREADS / 32];
    int64_t base = (int64_t)blockIdx.x * SCAN_TILE + (int64_t)threadIdx.x * SCAN_ITEMS;
    int v[SCAN_ITEMS];
    int local = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        local += v[k];
    }
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0;
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < SCAN_THREADS / 32) warp_sums[lane] = wi - w;   // exclusive
        if (lane == SCAN_THREADS / 32 - 1 && tile_sums) tile_sums[blockIdx.x] = wi;
    }
    __syncthreads();
    int run = warp_sums[warp] + incl - local;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
}

__global__ void scan_add_kernel(int* __restrict__ out, const int* __restrict__ tile_offsets, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += tile_offsets[i / SCAN_TILE];
}

// scratch must hold scan_scratch_ints(n) ints
static size_t scan_scratch_ints(int64_t n) {
    size_t total = 0;
    while (n > SCAN_TILE) {
        n = (n + SCAN_TILE - 1) / SCAN_TILE;
        total += (size_t)n + 1;
    }
    return total + 2;
}

static void exclusive_scan_i32(const int* in, int* out, int64_t n, int* scratch, cudaStream_t st) {
    if (n <= 0) return;
    int tiles = ceil_div(n, SCAN_TILE);
    if (tiles == 1) {
        scan_tile_kernel<<<1, SCAN_THREADS, 0, st>>>(in, out, nullptr, n);
        return;
    }
    int* sums = scratch;
    scan_tile_kernel<<<tiles, SCAN_THREADS, 0, st>>>(in, out, sums, n);
    exclusive_scan_i32(sums, sums, tiles, scratch + tiles + 1, st);
    scan_add_kernel<<<ceil_div(n, 256), 256, 0, st>>>(out, sums, n);
}

// ------------------------------------------------------------------------------------------------
// LSD radix sort, uint64 keys + optional uint32 payload, 8-bit digits, stable
// ------------------------------------------------------------------------------------------------
constexpr int RS_WARPS = 8;
constexpr int RS_ROUNDS = 8;                          // keys per lane
constexpr int RS_TILE = RS_WARPS * 32 * RS_ROUNDS;    // 2048 keys per block

__device__ __forceinline__ void rs_warp_hist(const uint64_t* __restrict__ keys, int64_t n, int64_t warp_base, int shift,
                                             int* hist /* smem [256] of this warp */, int lane) {
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t idx = warp_base + r * 32 + lane;
        int digit = idx < n ? (int)((keys[idx] >> shift) & 0xFF) : 256;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        if (digit < 256 && lane == __ffs(peers) - 1) hist[digit] += __popc(peers);
        __syncwarp();
    }
}

__global__ void __launch_bounds__(RS_WARPS * 32) rs_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                                int* __restrict__ counts /* [256][nblocks] */, int nblocks) {
    __shared__ int hist[RS_WARPS][256];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    int64_t warp_base = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * 32 * RS_ROUNDS;
    rs_warp_hist(keys, n, warp_base, shift, hist[warp], lane);
    __syncthreads();
    int d = threadIdx.x;   // 256 threads == 256 digits
    int total = 0;
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) total += hist[w][d];
    counts[(int64_t)d * nblocks + blockIdx.x] = total;
}

__global__ void __launch_bounds__(RS_WARPS * 32) rs_scatter_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals,
                                                                   uint64_t* __restrict__ keys_out, uint32_t* __restrict__ vals_out,
                                                                   int64_t n, int shift, const int* __restrict__ offsets, int nblocks) {
    __shared__ int hist[RS_WARPS][256];
    int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += blockDim.x) (&hist[0][0])[i] = 0;
    __syncthreads();
    int64_t warp_base = (int64_t)blockIdx.x * RS_TILE + (int64_t)warp * 32 * RS_ROUNDS;
    rs_warp_hist(keys, n, warp_base, shift, hist[warp], lane);
    __syncthreads();
    {   // per-digit running offset across the warps of this block (stable: warp order = key order)
        int d = threadIdx.x;
        int run = offsets[(int64_t)d * nblocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            int c = hist[w][d];
            hist[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
    int* off = hist[warp];
    for (int r = 0; r < RS_ROUNDS; ++r) {
        int64_t idx = warp_base + r * 32 + lane;
        bool valid = idx < n;
        uint64_t key = valid ? keys[idx] : 0;
        int digit = valid ? (int)((key >> shift) & 0xFF) : 256;
        unsigned peers = __match_any_sync(0xffffffffu, digit);
        int rank = __popc(peers & ((1u << lane) - 1u));
        int pos = 0;
        if (valid) pos = off[digit] + rank;
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) off[digit] += __popc(peers);
        __syncwarp();
        if (valid) {
            keys_out[pos] = key;
            if (vals) vals_out[pos] = vals[idx];
        }
    }
}

struct SortBufs {
    uint64_t *ka, *kb;
    uint32_t *va, *vb;   // may be null
    int* counts;         // [256 * nblocks + 1]
    int* scan_scratch;
};

// returns 0 if the sorted data ends in (ka,va), 1 if in (kb,vb)
static int radix_sort_u64(const SortBufs& b, int64_t n, int bits, cudaStream_t st) {
    if (n <= 0) return 0;
    int nblocks = ceil_div(n, RS_TILE);
    int passes = (bits + 7) / 8;
    if (passes < 1) passes = 1;
    uint64_t *kin = b.ka, *kout = b.kb;
    uint32_t *vin = b.va, *vout = b.vb;
    for (int p = 0; p < passes; ++p) {
        int shift = 8 * p;
        rs_hist_kernel<<<nblocks, RS_WARPS * 32, 0, st>>>(kin, n, shift, b.counts, nblocks);
        exclusive_scan_i32(b.counts, b.counts, (int64_t)256 * nblocks, b.scan_scratch, st);
        rs_scatter_kernel<<<nblocks, RS_WARPS * 32, 0, st>>>(kin, vin, kout, vout, n, shift, b.counts, nblocks);
        uint64_t* tk = kin; kin = kout; kout = tk;
        uint32_t* tv = vin; vin = vout; vout = tv;
    }
    return passes & 1;
}

// ------------------------------------------------------------------------------------------------
// build kernels
// ------------------------------------------------------------------------------------------------
__global__ void make_keys_kernel(const int64_t* __restrict__ u, const int64_t* __restrict__ i, int64_t n, int U, int I,
                                 uint64_t* __restrict__ keys, int* __restrict__ counts) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    int64_t uu = u[e], ii = i[e];
    bool bad = uu < 0 || uu >= U || ii < 0 || ii >= I;
    if (bad) {
        atomicAdd(&counts[5], 1);
        uu = 0; ii = 0;   // keeps the sort well-defined; the host rejects the build when counts[5] != 0
    }
    keys[e] = (uint64_t)uu * (uint64_t)I + (uint64_t)ii;
}

__global__ void flag_heads_kernel(const uint64_t* __restrict__ keys, int64_t n, int* __restrict__ flags) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e > n) return;
    if (e == n) { flags[e] = 0; return; }   // zero tail so the scan yields the total at [n]
    flags[e] = (e == 0 || keys[e] != keys[e - 1]) ? 1 : 0;
}

// compact unique sorted keys; emits CSR column indices, the CSC sort keys (i*U+u) and payloads (CSR id)
__global__ void compact_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ flags, const int* __restrict__ pos,
                               int64_t n, int U, int I, int* __restrict__ eu, int* __restrict__ colidx,
                               uint64_t* __restrict__ csc_keys, uint32_t* __restrict__ csc_vals, int* __restrict__ counts) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) {
        if (e == n) counts[0] = pos[n];
        return;
    }
    if (!flags[e]) return;
    int p = pos[e];
    uint64_t k = keys[e];
    int uu = (int)(k / (uint64_t)I), ii = (int)(k % (uint64_t)I);
    eu[p] = uu;
    colidx[p] = ii;
    csc_keys[p] = (uint64_t)ii * (uint64_t)U + (uint64_t)uu;
    csc_vals[p] = (uint32_t)p;
}

// ptr[r] = first position whose row >= r, for sorted `rows`; E read from counts[0]
__global__ void row_pointers_kernel(const int* __restrict__ rows, const int* __restrict__ counts, int R, int* __restrict__ ptr) {
    int E = counts[0];
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (E == 0) {
        if (e <= R) ptr[e] = 0;
        return;
    }
    if (e >= E) return;
    int cur = rows[e];
    int prev = e > 0 ? rows[e - 1] : -1;
    for (int r = prev + 1; r <= cur; ++r) ptr[r] = (int)e;
    if (e == E - 1)
        for (int r = cur + 1; r <= R; ++r) ptr[r] = E;
}

__global__ void csc_unpack_kernel(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ vals, const int* __restrict__ counts,
                                  int U, int* __restrict__ rowidx, int* __restrict__ perm, int* __restrict__ cols) {
    int E = counts[0];
    int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= E) return;
    uint64_t k = keys[c];
    rowidx[c] = (int)(k % (uint64_t)U);
    cols[c] = (int)(k / (uint64_t)U);
    perm[c] = (int)vals[c];
}

__global__ void unify_kernel(const int* __restrict__ rowptr, const int* __restrict__ colidx, const int* __restrict__ colptr,
                             const int* __restrict__ rowidx, const int* __restrict__ perm, int* __restrict__ counts, int U, int I,
                             int* __restrict__ adj_ptr, int* __restrict__ adj_idx, int* __restrict__ adj_eid, int* __restrict__ ntasks /* [N+1] */) {
    int E = counts[0];
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int N = U + I;
    if (t <= N) {
        int p = t < U ? rowptr[t] : E + colptr[t - U];
        adj_ptr[t] = p;
        if (t < N) {
            int q = (t + 1 < U) ? rowptr[t + 1] : E + colptr[t + 1 - U];
            int deg = q - p;
            ntasks[t] = deg <= CHUNK ? 1 : (deg + CHUNK - 1) / CHUNK;
            if (t < U && deg == 0) atomicAdd(&counts[4], 1);
        } else {
            ntasks[t] = 0;
        }
    }
    if (t < E) {
        adj_idx[t] = U + colidx[t];
        adj_eid[t] = (int)t;
        adj_idx[E + t] = rowidx[t];
        adj_eid[E + t] = perm[t];
    }
}

// long rows: flag + slot counts
__global__ void long_flags_kernel(const int* __restrict__ ntasks, int N, int* __restrict__ is_long /* [N+1] */, int* __restrict__ nslots /* [N+1] */) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > N) return;
    int k = n < N ? ntasks[n] : 0;
    is_long[n] = k > 1 ? 1 : 0;
    nslots[n] = k > 1 ? k : 0;
}

// emit unsorted tasks + sort keys.  key = side (bit 16) | (CHUNK - len) so that an ascending stable sort
// gives users first, longest first, node order inside equal lengths.
__global__ void emit_tasks_kernel(const int* __restrict__ adj_ptr, const int* __restrict__ task_off, const int* __restrict__ long_id,
                                  const int* __restrict__ slot_off, const int* __restrict__ is_long, int U, int N,
                                  int4* __restrict__ tasks, uint64_t* __restrict__ keys, uint32_t* __restrict__ vals,
                                  int* __restrict__ long_first_slot, int* __restrict__ long_counter, int* __restrict__ counts) {
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n > N) return;
    if (n == N) {
        counts[1] = task_off[N];
        counts[2] = long_id[N];
        counts[3] = slot_off[N];
        long_first_slot[long_id[N]] = slot_off[N];
        return;
    }
    int beg = adj_ptr[n], end = adj_ptr[n + 1];
    int t0 = task_off[n];
    int k = task_off[n + 1] - t0;
    int lid = is_long[n] ? long_id[n] : -1;
    if (lid >= 0) {
        long_first_slot[lid] = slot_off[n];
        long_counter[lid] = 0;
    }
    for (int c = 0; c < k; ++c) {
        int b = beg + c * CHUNK;
        int e = (k == 1) ? end : min(end, b + CHUNK);
        tasks[t0 + c] = make_int4((int)n, b, e, lid);
        keys[t0 + c] = ((uint64_t)(n >= U ? 1 : 0) << 16) | (uint64_t)(CHUNK - (e - b));
        vals[t0 + c] = (uint32_t)(t0 + c);
    }
}

__global__ void gather_tasks_kernel(const int4* __restrict__ src, const uint32_t* __restrict__ order, const uint64_t* __restrict__ keys,
                                    const int* __restrict__ counts, int4* __restrict__ dst, int* __restrict__ counts_out) {
    int T = counts[1];
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= T) return;
    dst[t] = src[order[t]];
    // number of user-side tasks = first position whose side bit is set
    bool side = (keys[t] >> 16) & 1;
    bool prev_side = t > 0 ? ((keys[t - 1] >> 16) & 1) : false;
    if (side && !prev_side) counts_out[6] = (int)t;
    if (t == T - 1 && !side) counts_out[6] = T;
}

// entries past the data-dependent count (counts[which]) get a key above every real key
__global__ void pad_tail_kernel(uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, const int* __restrict__ counts, int which,
                                uint64_t pad_key, int64_t n) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n || e < counts[which]) return;
    keys[e] = pad_key;
    vals[e] = 0xffffffffu;
}

}  // namespace ngacf

using namespace ngacf;

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

extern "C" size_t ngacf_graph_build_workspace_bytes(int64_t E_in, int32_t U, int32_t I) {
    int64_t N = (int64_t)U + I;
    int64_t cap_tasks = N + 2 * E_in / CHUNK + 2;
    int64_t n_sort = E_in > cap_tasks ? E_in : cap_tasks;
    int64_t nblocks = (n_sort + RS_TILE - 1) / RS_TILE + 1;
    size_t b = 0;
    b += 2 * align_up((size_t)n_sort * 8);                 // keys A/B
    b += 2 * align_up((size_t)n_sort * 4);                 // vals A/B
    b += align_up((size_t)(256 * nblocks + 1) * 4);        // digit counts
    b += align_up(scan_scratch_ints(256 * nblocks + n_sort + N + 8) * 4 * 2);
    b += 2 * align_up((size_t)(E_in + 1) * 4);             // flags, positions
    b += 2 * align_up((size_t)(E_in + 1) * 4);             // eu, csc cols
    b += 6 * align_up((size_t)(N + 2) * 4);                // ntasks, task_off, is_long, long_id, nslots, slot_off
    b += align_up((size_t)cap_tasks * 16);                 // unsorted tasks
    return b + 4096;
}

extern "C" int ngacf_graph_build(const int64_t* coo_u, const int64_t* coo_i, int64_t E_in, int32_t U, int32_t I,
                                 int32_t* rowptr, int32_t* colidx, int32_t* colptr, int32_t* rowidx, int32_t* perm,
                                 int32_t* adj_ptr, int32_t* adj_idx, int32_t* adj_eid, int32_t* tasks,
                                 int32_t* long_first_slot, int32_t* long_counter, int32_t* counts, void* workspace,
                                 size_t workspace_bytes, void* stream) {
    NGACF_REQUIRE(U > 0 && I > 0 && E_in >= 0, "graph_build: bad sizes U=%d I=%d E=%lld", U, I, (long long)E_in);
    NGACF_REQUIRE(E_in < (int64_t)1 << 30, "graph_build: E_in too large for int32 edge ids");
    NGACF_REQUIRE((int64_t)U + I < (int64_t)1 << 30, "graph_build: too many nodes");
    if (workspace_bytes < ngacf_graph_build_workspace_bytes(E_in, U, I)) {
        set_error("graph_build: workspace too small");
        return NGACF_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t N = (int64_t)U + I;
    const int64_t cap_tasks = N + 2 * E_in / CHUNK + 2;
    const int64_t n_sort = E_in > cap_tasks ? E_in : cap_tasks;
    const int64_t nblocks = (n_sort + RS_TILE - 1) / RS_TILE + 1;

    char* w = (char*)workspace;
    auto take = [&](size_t bytes) { char* p = w; w += align_up(bytes); return p; };
    SortBufs sb;
    sb.ka = (uint64_t*)take((size_t)n_sort * 8);
    sb.kb = (uint64_t*)take((size_t)n_sort * 8);
    sb.va = (uint32_t*)take((size_t)n_sort * 4);
    sb.vb = (uint32_t*)take((size_t)n_sort * 4);
    sb.counts = (int*)take((size_t)(256 * nblocks + 1) * 4);
    size_t scan_ints = scan_scratch_ints(256 * nblocks + n_sort + N + 8);
    sb.scan_scratch = (int*)take(scan_ints * 4 * 2);
    int* flags = (int*)take((size_t)(E_in + 1) * 4);
    int* pos = (int*)take((size_t)(E_in + 1) * 4);
    int* eu = (int*)take((size_t)(E_in + 1) * 4);
    int* ccols = (int*)take((size_t)(E_in + 1) * 4);
    int* ntasks = (int*)take((size_t)(N + 2) * 4);
    int* task_off = (int*)take((size_t)(N + 2) * 4);
    int* is_long = (int*)take((size_t)(N + 2) * 4);
    int* long_id = (int*)take((size_t)(N + 2) * 4);
    int* nslots = (int*)take((size_t)(N + 2) * 4);
    int* slot_off = (int*)take((size_t)(N + 2) * 4);
    int4* tasks_unsorted = (int4*)take((size_t)cap_tasks * 16);

    cudaMemsetAsync(counts, 0, 8 * sizeof(int), st);
    const int TB = 256;
    auto bits_for = [](uint64_t maxv) { int b = 1; while (b < 64 && (maxv >> b)) ++b; return b; };

    // 1. sort + dedupe the (u,i) keys -> CSR order
    if (E_in > 0) {
        make_keys_kernel<<<ceil_div(E_in, TB), TB, 0, st>>>(coo_u, coo_i, E_in, U, I, sb.ka, counts);
        SortBufs s1 = sb;
        s1.va = s1.vb = nullptr;
        int where = radix_sort_u64(s1, E_in, bits_for((uint64_t)U * (uint64_t)I - 1), st);
        const uint64_t* sorted = where ? sb.kb : sb.ka;
        uint64_t* other = where ? sb.ka : sb.kb;
        flag_heads_kernel<<<ceil_div(E_in + 1, TB), TB, 0, st>>>(sorted, E_in, flags);
        exclusive_scan_i32(flags, pos, E_in + 1, sb.scan_scratch, st);
        // CSC sort keys are written into the buffer not holding `sorted`
        compact_kernel<<<ceil_div(E_in + 1, TB), TB, 0, st>>>(sorted, flags, pos, E_in, U, I, eu, colidx, other, sb.va, counts);
        row_pointers_kernel<<<ceil_div(E_in > U ? E_in : U + 1, TB), TB, 0, st>>>(eu, counts, U, rowptr);
        // 2. CSC = stable sort by (i,u) with payload = CSR id.  The device-side E is data dependent, so
        //    the sort runs over all E_in slots with the unused tail padded with a key above every real one.
        pad_tail_kernel<<<ceil_div(E_in, TB), TB, 0, st>>>(other, sb.va, counts, 0, (uint64_t)U * (uint64_t)I, E_in);
        SortBufs s2 = sb;
        s2.ka = other;
        s2.kb = (uint64_t*)sorted;
        int where2 = radix_sort_u64(s2, E_in, bits_for((uint64_t)U * (uint64_t)I), st);
        const uint64_t* ck = where2 ? s2.kb : s2.ka;
        const uint32_t* cv = where2 ? s2.vb : s2.va;
        csc_unpack_kernel<<<ceil_div(E_in, TB), TB, 0, st>>>(ck, cv, counts, U, rowidx, perm, ccols);
        row_pointers_kernel<<<ceil_div(E_in > I ? E_in : I + 1, TB), TB, 0, st>>>(ccols, counts, I, colptr);
    } else {
        cudaMemsetAsync(rowptr, 0, (size_t)(U + 1) * 4, st);
        cudaMemsetAsync(colptr, 0, (size_t)(I + 1) * 4, st);
    }
    // 3. unified adjacency + task counts
    int64_t span = (E_in > N + 1 ? E_in : N + 1);
    unify_kernel<<<ceil_div(span, TB), TB, 0, st>>>(rowptr, colidx, colptr, rowidx, perm, counts, U, I, adj_ptr, adj_idx, adj_eid, ntasks);
    exclusive_scan_i32(ntasks, task_off, N + 1, sb.scan_scratch, st);
    long_flags_kernel<<<ceil_div(N + 1, TB), TB, 0, st>>>(ntasks, (int)N, is_long, nslots);
    exclusive_scan_i32(is_long, long_id, N + 1, sb.scan_scratch, st);
    exclusive_scan_i32(nslots, slot_off, N + 1, sb.scan_scratch, st);
    emit_tasks_kernel<<<ceil_div(N + 1, TB), TB, 0, st>>>(adj_ptr, task_off, long_id, slot_off, is_long, U, (int)N, tasks_unsorted,
                                                          sb.ka, sb.va, long_first_slot, long_counter, counts);
    // 4. degree bucketing: stable sort of the tasks by (side, descending length).  T is data dependent
    //    (<= cap_tasks); pad as above.
    pad_tail_kernel<<<ceil_div(cap_tasks, TB), TB, 0, st>>>(sb.ka, sb.va, counts, 1, 1ull << 17, cap_tasks);
    int where3 = radix_sort_u64(sb, cap_tasks, 18, st);
    gather_tasks_kernel<<<ceil_div(cap_tasks, TB), TB, 0, st>>>(tasks_unsorted, where3 ? sb.vb : sb.va, where3 ? sb.kb : sb.ka, counts,
                                                                (int4*)tasks, counts);
    return check_launch("graph_build");
}

