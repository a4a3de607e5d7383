// Stand-alone probe: where do 256-byte row gathers with a power-law popularity come from fastest on a B200?
//   A  baseline          every row through the read-only path (L1 allocating), as the round-1 gather kernels do
//   B  L1 policy         hot rows (rank < K) ld.global.nc.L1::evict_last, cold rows L1::no_allocate, adjacency hot-first
//   C  shared memory     persistent 1024-thread CTAs stage the K hot rows in shared memory once; cold rows L1::no_allocate
//   U  uniform           baseline kernel on uniformly random indices (the pure L2 -> SM wall)
// One 16-lane group sums the rows of one segment (a "task") and writes one 256-byte row, like aggregate_fwd without weights.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scratch_ab/probe_gather scripts/probe/probe_gather.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <algorithm>
#include <random>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

constexpr int D = 64;

__device__ __forceinline__ float4 ld_alloc(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ld_noalloc(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ float4 ld_last(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream_i32(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ unsigned group_mask() { return (threadIdx.x & 16) ? 0xFFFF0000u : 0x0000FFFFu; }

// POLICY 0: all allocating; 1: [beg,hot_end) evict_last, rest no_allocate; 2: all no_allocate
template <int POLICY>
__global__ void __launch_bounds__(256) gather_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ seg_hot, const int* __restrict__ idx,
                                                     int T, const float* __restrict__ table, float* __restrict__ out) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (t >= T) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int beg = seg_ptr[t], end = seg_ptr[t + 1];
    const int hot_end = POLICY == 1 ? beg + seg_hot[t] : beg;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 16) {
        const int m_l = base + lane16 < end ? ld_stream_i32(idx + base + lane16) : 0;
        const int cnt = min(16, end - base);
        if (POLICY == 1 && base + 16 <= hot_end) {
#pragma unroll 8
            for (int j = 0; j < cnt; ++j) {
                const int m = __shfl_sync(gm, m_l, j, 16);
                const float4 v = ld_last(table + (int64_t)m * D + lane16 * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        } else {
#pragma unroll 8
            for (int j = 0; j < cnt; ++j) {
                const int m = __shfl_sync(gm, m_l, j, 16);
                const float4 v = POLICY == 0 ? ld_alloc(table + (int64_t)m * D + lane16 * 4) : ld_noalloc(table + (int64_t)m * D + lane16 * 4);
                acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
            }
        }
    }
    *reinterpret_cast<float4*>(out + (int64_t)t * D + lane16 * 4) = acc;
}

// explicit batching: NB independent row gathers are issued before any of them is consumed (an empty asm that names every
// loaded register keeps ptxas from interleaving the adds, which in-order issue would turn into stalls); the tail of a
// segment uses predicated loads instead of a serial remainder loop
template <int NB>
__global__ void __launch_bounds__(256) gather_batched_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ idx, int T,
                                                             const float* __restrict__ table, float* __restrict__ out) {
    const int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    if (t >= T) return;
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    const int beg = seg_ptr[t], end = seg_ptr[t + 1];
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int base = beg; base < end; base += 16) {
        const int m_l = base + lane16 < end ? ld_stream_i32(idx + base + lane16) : 0;
        const int cnt = min(16, end - base);
        for (int j0 = 0; j0 < cnt; j0 += NB) {
            float4 v[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                const int m = __shfl_sync(gm, m_l, (j0 + q) & 15, 16);
                v[q] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j0 + q < cnt) v[q] = ld_alloc(table + (int64_t)m * D + lane16 * 4);
            }
            if (NB == 8)
                asm volatile("" : "+f"(v[0].x), "+f"(v[1].x), "+f"(v[2].x), "+f"(v[3].x), "+f"(v[4 % NB].x), "+f"(v[5 % NB].x), "+f"(v[6 % NB].x), "+f"(v[7 % NB].x));
            else
                asm volatile("" : "+f"(v[0].x), "+f"(v[1].x), "+f"(v[2].x), "+f"(v[3].x));
#pragma unroll
            for (int q = 0; q < NB; ++q) { acc.x += v[q].x; acc.y += v[q].y; acc.z += v[q].z; acc.w += v[q].w; }
        }
    }
    *reinterpret_cast<float4*>(out + (int64_t)t * D + lane16 * 4) = acc;
}

// persistent CTAs, K hot rows in shared memory; idx of a hot entry = slot (0..K-1), of a cold entry = row id
template <int THREADS>
__global__ void __launch_bounds__(THREADS, 1) gather_smem_kernel(const int* __restrict__ seg_ptr, const int* __restrict__ seg_hot,
                                                                 const int* __restrict__ idx_slot, int T, const float* __restrict__ table,
                                                                 const int* __restrict__ hot_rows, int K, float* __restrict__ out) {
    extern __shared__ __align__(16) float hot[];
    for (int i = threadIdx.x; i < K * 16; i += THREADS) {
        const int r = i >> 4, q = i & 15;
        reinterpret_cast<float4*>(hot)[i] = ld_noalloc(table + (int64_t)hot_rows[r] * D + q * 4);
    }
    __syncthreads();
    const int lane16 = threadIdx.x & 15;
    const unsigned gm = group_mask();
    constexpr int G = THREADS / 16;
    for (int t = blockIdx.x * G + (threadIdx.x >> 4); t < T; t += gridDim.x * G) {
        const int beg = seg_ptr[t], end = seg_ptr[t + 1];
        const int hot_end = beg + seg_hot[t];
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int base = beg; base < end; base += 16) {
            const int m_l = base + lane16 < end ? ld_stream_i32(idx_slot + base + lane16) : 0;
            const int cnt = min(16, end - base);
            if (base + 16 <= hot_end) {
#pragma unroll 8
                for (int j = 0; j < cnt; ++j) {
                    const int m = __shfl_sync(gm, m_l, j, 16);
                    const float4 v = *reinterpret_cast<const float4*>(hot + m * D + lane16 * 4);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            } else {
                const int nh = max(0, min(cnt, hot_end - base));      // mixed batch: first nh entries are slots
#pragma unroll 8
                for (int j = 0; j < cnt; ++j) {
                    const int m = __shfl_sync(gm, m_l, j, 16);
                    float4 v;
                    if (j < nh) v = *reinterpret_cast<const float4*>(hot + m * D + lane16 * 4);
                    else v = ld_noalloc(table + (int64_t)m * D + lane16 * 4);
                    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
                }
            }
        }
        *reinterpret_cast<float4*>(out + (int64_t)t * D + lane16 * 4) = acc;
    }
}

struct Workload {
    std::vector<int> seg_ptr, seg_hot, idx, idx_slot, hot_rows;
    double hot_frac;
};

// T segments of power-law lengths (mean ~deg), neighbours ~ rank^-alpha over R rows (rank -> row id through a random permutation);
// entries of a segment are reordered hot-first (rank < K), original order kept inside both parts
static Workload make_workload(int T, int R, double deg, double alpha, int K, bool uniform, uint64_t seed) {
    std::mt19937_64 rng(seed);
    std::vector<double> cdf(R);
    double s = 0;
    for (int r = 0; r < R; ++r) { s += uniform ? 1.0 : pow(r + 1.0, -alpha); cdf[r] = s; }
    std::vector<int> perm(R);
    for (int r = 0; r < R; ++r) perm[r] = r;
    std::shuffle(perm.begin(), perm.end(), rng);
    std::uniform_real_distribution<double> un(0.0, 1.0);
    Workload w;
    w.seg_ptr.push_back(0);
    w.hot_rows.resize(K);
    for (int k = 0; k < K; ++k) w.hot_rows[k] = perm[k];
    int64_t hot_total = 0;
    for (int t = 0; t < T; ++t) {
        // lengths: rank^-0.6 power law, scaled to mean deg, clipped to [1,128]
        double x = pow((t + 1.0) / T, -0.6) * deg * 0.4;
        int len = (int)std::min(128.0, std::max(1.0, x));
        std::vector<int> hotv, coldv, hots, colds;
        for (int j = 0; j < len; ++j) {
            double u = un(rng) * s;
            int r = (int)(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
            if (r >= R) r = R - 1;
            if (r < K) { hotv.push_back(perm[r]); hots.push_back(r); } else { coldv.push_back(perm[r]); colds.push_back(perm[r]); }
        }
        w.seg_hot.push_back((int)hotv.size());
        hot_total += (int64_t)hotv.size();
        for (size_t j = 0; j < hotv.size(); ++j) { w.idx.push_back(hotv[j]); w.idx_slot.push_back(hots[j]); }
        for (size_t j = 0; j < coldv.size(); ++j) { w.idx.push_back(coldv[j]); w.idx_slot.push_back(colds[j]); }
        w.seg_ptr.push_back((int)w.idx.size());
    }
    w.hot_frac = (double)hot_total / (double)w.idx.size();
    return w;
}

template <class F>
static float time_us(F f, int warm = 3, int reps = 20) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < warm; ++i) f();
    CK(cudaEventRecord(e0));
    for (int i = 0; i < reps; ++i) f();
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    CK(cudaGetLastError());
    return ms * 1000.f / reps;
}

template <class T>
static T* to_dev(const std::vector<T>& v) {
    T* p;
    CK(cudaMalloc(&p, v.size() * sizeof(T)));
    CK(cudaMemcpy(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice));
    return p;
}

int main(int argc, char** argv) {
    const int R = argc > 1 ? atoi(argv[1]) : 40981;       // rows of the gathered table
    const int T = argc > 2 ? atoi(argv[2]) : 29858;       // segments
    const double deg = argc > 3 ? atof(argv[3]) : 34.4;
    float* table;
    CK(cudaMalloc(&table, (size_t)R * D * 4));
    {
        std::vector<float> hrand((size_t)R * D);
        std::mt19937 g(7);
        std::uniform_real_distribution<float> un(-1.f, 1.f);
        for (auto& x : hrand) x = un(g);
        CK(cudaMemcpy(table, hrand.data(), hrand.size() * 4, cudaMemcpyHostToDevice));
    }
    float* out;
    CK(cudaMalloc(&out, (size_t)T * D * 4));
    // other traffic between repetitions: a 256 MB buffer written to push the table out of L1 (it stays in L2 only if it fits)
    printf("table %d rows (%.1f MB), %d segments\n", R, R * 256.0 / 1e6, T);
    for (int pass = 0; pass < 2; ++pass) {
        const double alpha = pass == 0 ? 0.8 : 0.6;
        for (int K : {384, 704, 832}) {
            Workload w = make_workload(T, R, deg, alpha, K, false, 1234);
            int *seg_ptr = to_dev(w.seg_ptr), *seg_hot = to_dev(w.seg_hot), *idx = to_dev(w.idx), *idx_slot = to_dev(w.idx_slot), *hot_rows = to_dev(w.hot_rows);
            const int64_t n = (int64_t)w.idx.size();
            const int blocks = (T * 16 + 255) / 256;
            float a = time_us([&] { gather_kernel<0><<<blocks, 256>>>(seg_ptr, seg_hot, idx, T, table, out); });
            float b = time_us([&] { gather_kernel<1><<<blocks, 256>>>(seg_ptr, seg_hot, idx, T, table, out); });
            float n2 = time_us([&] { gather_kernel<2><<<blocks, 256>>>(seg_ptr, seg_hot, idx, T, table, out); });
            float b8 = time_us([&] { gather_batched_kernel<8><<<blocks, 256>>>(seg_ptr, idx, T, table, out); });
            float b4 = time_us([&] { gather_batched_kernel<4><<<blocks, 256>>>(seg_ptr, idx, T, table, out); });
            printf("   batched8 %.1f us  batched4 %.1f us\n", b8, b4);
            const size_t smem = (size_t)K * D * 4;
            CK(cudaFuncSetAttribute(gather_smem_kernel<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CK(cudaFuncSetAttribute(gather_smem_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            float c = time_us([&] { gather_smem_kernel<1024><<<148, 1024, smem>>>(seg_ptr, seg_hot, idx_slot, T, table, hot_rows, K, out); });
            float c5 = time_us([&] { gather_smem_kernel<512><<<148, 512, smem>>>(seg_ptr, seg_hot, idx_slot, T, table, hot_rows, K, out); });
            printf("alpha %.1f K %4d: %lld visits, hot %.3f | A base %.1f us (%.2f TB/s)  B L1-policy %.1f us  N all-noalloc %.1f us  C smem1024 %.1f us  C smem512 %.1f us\n",
                   alpha, K, (long long)n, w.hot_frac, a, n * 256.0 / a / 1e6, b, n2, c, c5);
            cudaFree(seg_ptr); cudaFree(seg_hot); cudaFree(idx); cudaFree(idx_slot); cudaFree(hot_rows);
        }
    }
    {
        Workload w = make_workload(T, R, deg, 0.8, 16, true, 99);
        int *seg_ptr = to_dev(w.seg_ptr), *seg_hot = to_dev(w.seg_hot), *idx = to_dev(w.idx);
        const int64_t n = (int64_t)w.idx.size();
        const int blocks = (T * 16 + 255) / 256;
        float a = time_us([&] { gather_kernel<0><<<blocks, 256>>>(seg_ptr, seg_hot, idx, T, table, out); });
        float n2 = time_us([&] { gather_kernel<2><<<blocks, 256>>>(seg_ptr, seg_hot, idx, T, table, out); });
        printf("uniform: %lld visits | A base %.1f us (%.2f TB/s)  N all-noalloc %.1f us (%.2f TB/s)\n", (long long)n, a, n * 256.0 / a / 1e6, n2,
               n * 256.0 / n2 / 1e6);
    }
    return 0;
}
