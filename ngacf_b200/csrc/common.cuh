// Shared device helpers for the ngacf_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/ngacf_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "ngacf_b200 kernels are written for sm_100a (B200) only"
#endif

namespace ngacf {

constexpr int D = NGACF_D;
constexpr int CHUNK = NGACF_CHUNK;
constexpr float LRELU_ALPHA = 0.2f;   // SPUIGACF.py:22
constexpr int SCRATCH_STRIDE = D + 8; // per-slot partial: 64 accumulators + up to 8 per-head sums

void set_error(const char* fmt, ...);
int check_launch(const char* what);

#define NGACF_REQUIRE(cond, ...)                       \
    do {                                               \
        if (!(cond)) {                                 \
            ngacf::set_error(__VA_ARGS__);             \
            return NGACF_ERR_INVALID_ARG;              \
        }                                              \
    } while (0)

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// loads / stores with cache intent: gathered tables go through L1 (popular rows hit), streamed
// arrays do not allocate in L1.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld_gather4(const float* p) {   // read-only path, L1 allocating
    return __ldg(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int ld_stream_i32(const int* p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 ld_cg4(const float* p) {   // L2-coherent (other SMs' writes in this launch)
    return __ldcg(reinterpret_cast<const float4*>(p));
}

__device__ __forceinline__ float elu(float z) { return z > 0.f ? z : expm1f(z); }
__device__ __forceinline__ float elu_grad(float z) { return z > 0.f ? 1.f : __expf(z); }

// e = exp(-LeakyReLU(x)); the same expression is used by forward and backward so the recomputed
// weight is bit-identical to the one the forward normalised with.
__device__ __forceinline__ float edge_weight(float x) {
    float l = x > 0.f ? x : LRELU_ALPHA * x;
    return __expf(-l);
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (same constants/rounds as oracle/port.py:philox4x32_10)
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                                       uint32_t k1, uint32_t out[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0;
        uint32_t hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__host__ __device__ __forceinline__ uint32_t keep_threshold(float droprate) {
    double t = floor((1.0 - (double)droprate) * 65536.0 + 0.5);
    if (t < 0) t = 0;
    if (t > 65536.0) t = 65536.0;
    return (uint32_t)t;
}

// 16-lane group helpers: a "group" is one half of a warp; the two halves may diverge.
__device__ __forceinline__ unsigned group_mask() { return (threadIdx.x & 16) ? 0xFFFF0000u : 0x0000FFFFu; }

template <int H>
__device__ __forceinline__ float head_reduce(float v, unsigned mask) {
    // sum over the lanes that share a head: H=8 -> lane pairs, H=1 -> all 16 lanes (adjacent-pair tree)
    v += __shfl_xor_sync(mask, v, 1, 16);
    if (H == 1) {
        v += __shfl_xor_sync(mask, v, 2, 16);
        v += __shfl_xor_sync(mask, v, 4, 16);
        v += __shfl_xor_sync(mask, v, 8, 16);
    }
    return v;
}

// ---------------------------------------------------------------------------------------------
// 8-lane groups for the sparse kernels: one group per task (row or <=128-edge chunk), lane l owns
// columns 8l..8l+7 (two 16-byte vector loads per gathered row).  With H=8 a lane IS a head: the
// per-edge softmax weight is computed once per head and the backward's per-head dot products are
// lane-local; with H=1 the eight lanes share the single head.
// ---------------------------------------------------------------------------------------------
struct f8 { float4 a, b; };
__device__ __forceinline__ unsigned group8_mask() { return 0xFFu << (threadIdx.x & 24); }
__device__ __forceinline__ f8 ld_gather8(const float* p) { f8 v; v.a = ld_gather4(p); v.b = ld_gather4(p + 4); return v; }
__device__ __forceinline__ f8 ld_stream8(const float* p) { f8 v; v.a = ld_stream4(p); v.b = ld_stream4(p + 4); return v; }
__device__ __forceinline__ void st_stream8(float* p, const f8& v) { st_stream4(p, v.a); st_stream4(p + 4, v.b); }
__device__ __forceinline__ f8 zero8() { f8 v; v.a = make_float4(0.f, 0.f, 0.f, 0.f); v.b = v.a; return v; }
__device__ __forceinline__ void fma8(f8& acc, float w, const f8& x) {
    acc.a.x = fmaf(w, x.a.x, acc.a.x); acc.a.y = fmaf(w, x.a.y, acc.a.y); acc.a.z = fmaf(w, x.a.z, acc.a.z); acc.a.w = fmaf(w, x.a.w, acc.a.w);
    acc.b.x = fmaf(w, x.b.x, acc.b.x); acc.b.y = fmaf(w, x.b.y, acc.b.y); acc.b.z = fmaf(w, x.b.z, acc.b.z); acc.b.w = fmaf(w, x.b.w, acc.b.w);
}
__device__ __forceinline__ float dot8(const f8& x, const f8& y) {
    return x.a.x * y.a.x + x.a.y * y.a.y + x.a.z * y.a.z + x.a.w * y.a.w + x.b.x * y.b.x + x.b.y * y.b.y + x.b.z * y.b.z + x.b.w * y.b.w;
}
template <int H>
__device__ __forceinline__ float head_reduce8(float v, unsigned mask) {     // H=8: lane-local; H=1: sum over the 8 lanes
    if (H == 1) {
        v += __shfl_xor_sync(mask, v, 1, 8);
        v += __shfl_xor_sync(mask, v, 2, 8);
        v += __shfl_xor_sync(mask, v, 4, 8);
    }
    return v;
}
template <int H>
__device__ __forceinline__ bool head_writer8(int l8) { return H == 8 ? true : l8 == 0; }

// Partial sums of a row longer than CHUNK edges: every chunk's group parks (acc, per-head sum) in its scratch
// slot; the group that arrives LAST re-reads all slots in slot order (deterministic) and finishes the row.
template <int H>
__device__ __forceinline__ bool long_row_combine8(int lid, int chunk, const int* __restrict__ long_first_slot, int* long_counter,
                                                  float* scratch, int l8, unsigned gm, f8& acc, float& sum) {
    const int head = H == 8 ? l8 : 0;
    const int first = long_first_slot[lid];
    const int nslots = long_first_slot[lid + 1] - first;
    float* slot = scratch + (size_t)(first + chunk) * SCRATCH_STRIDE;
    *reinterpret_cast<float4*>(slot + l8 * 8) = acc.a;
    *reinterpret_cast<float4*>(slot + l8 * 8 + 4) = acc.b;
    if (head_writer8<H>(l8)) slot[D + head] = sum;
    __threadfence();
    __syncwarp(gm);                          // every lane's partial is fenced before lane 0 publishes
    int old = 0;
    if (l8 == 0) old = atomicAdd(long_counter + lid, 1);
    old = __shfl_sync(gm, old, 0, 8);
    if (old != nslots - 1) return false;
    __threadfence();
    f8 t = zero8();
    float ts = 0.f;
    for (int c = 0; c < nslots; ++c) {
        const float* sl = scratch + (size_t)(first + c) * SCRATCH_STRIDE;
        const float4 va = ld_cg4(sl + l8 * 8), vb = ld_cg4(sl + l8 * 8 + 4);
        t.a.x += va.x; t.a.y += va.y; t.a.z += va.z; t.a.w += va.w;
        t.b.x += vb.x; t.b.y += vb.y; t.b.z += vb.z; t.b.w += vb.w;
        ts += __ldcg(sl + D + head);
    }
    acc = t;
    sum = ts;
    if (l8 == 0) long_counter[lid] = 0;      // re-arm for the next launch
    return true;
}

}  // namespace ngacf
