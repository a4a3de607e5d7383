// Probe for tcgen05.mma kind::tf32 operand descriptors (SWIZZLE_NONE images, K-major vs MN-major).  Experiment tooling only.
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mkdesc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
// A: [128][64] row-major, B: [128][80] row-major; images: chunk q of row r at q*2048 + r*16
extern "C" __global__ void __launch_bounds__(128, 1) probe_kernel(const float* A, const float* B, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo,
                                                                  uint32_t b_lbo, uint32_t b_sbo, int nk, uint32_t a_step, uint32_t b_step, float* Dout) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* iA = sm;                    // 16 panels (pad to 32 panels so M=128 MN-major reads stay in bounds)
    uint8_t* iB = sm + 32 * 2048;        // 20 panels
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 52 * 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 52 * 2048 / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 128 * 16; i += 128) {
        int r = i >> 4, q = i & 15;
        *reinterpret_cast<float4*>(iA + q * 2048 + r * 16) = *reinterpret_cast<const float4*>(A + r * 64 + q * 4);
    }
    for (int i = tid; i < 128 * 20; i += 128) {
        int r = i / 20, q = i % 20;
        *reinterpret_cast<float4*>(iB + q * 2048 + r * 16) = *reinterpret_cast<const float4*>(B + r * 80 + q * 4);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (tid == 0) {
        for (int ks = 0; ks < nk; ++ks) {
            uint64_t da = mkdesc(smem_u32(iA) + ks * a_step, a_lbo, a_sbo), db = mkdesc(smem_u32(iB) + ks * b_step, b_lbo, b_sbo);
            uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < 96; c0 += 8) {
        uint32_t u[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(tm + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) Dout[(32 * warp + lane) * 96 + c0 + j] = __uint_as_float(u[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}
extern "C" int probe_launch(const float* A, const float* B, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int nk,
                            uint32_t a_step, uint32_t b_step, float* Dout) {
    size_t smem = 52 * 2048 + 64;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_kernel<<<1, 128, smem>>>(A, B, idesc, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, Dout);
    cudaError_t e = cudaDeviceSynchronize();
    return (int)e;
}

#include <cuda_bf16.h>
// bf16 variant: A: [128][64], B: [128][80] fp32 row-major; images: 16-byte chunk q (8 bf16) of row r at q*2048 + r*16
extern "C" __global__ void __launch_bounds__(128, 1) probe_bf16_kernel(const float* A, const float* B, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo,
                                                                       uint32_t b_lbo, uint32_t b_sbo, int nk, uint32_t a_step, uint32_t b_step, float* Dout) {
    extern __shared__ __align__(128) uint8_t sm[];
    uint8_t* iA = sm;
    uint8_t* iB = sm + 32 * 2048;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sm + 52 * 2048);
    uint32_t* slot = reinterpret_cast<uint32_t*>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 52 * 2048 / 4; i += 128) reinterpret_cast<float*>(sm)[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < 128 * 8; i += 128) {
        int r = i >> 3, q = i & 7;
        __nv_bfloat16 v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16_rn(A[r * 64 + q * 8 + j]);
        *reinterpret_cast<uint4*>(iA + q * 2048 + r * 16) = *reinterpret_cast<uint4*>(v);
    }
    for (int i = tid; i < 128 * 10; i += 128) {
        int r = i / 10, q = i % 10;
        __nv_bfloat16 v[8];
        for (int j = 0; j < 8; ++j) v[j] = __float2bfloat16_rn(B[r * 80 + q * 8 + j]);
        *reinterpret_cast<uint4*>(iB + q * 2048 + r * 16) = *reinterpret_cast<uint4*>(v);
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(128u) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tm = *slot;
    if (tid == 0) {
        for (int ks = 0; ks < nk; ++ks) {
            uint64_t da = mkdesc(smem_u32(iA) + ks * a_step, a_lbo, a_sbo), db = mkdesc(smem_u32(iB) + ks * b_step, b_lbo, b_sbo);
            uint32_t acc = ks ? 1u : 0u;
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tm), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}"
                 ::"r"(smem_u32(bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int warp = tid >> 5, lane = tid & 31;
    for (int c0 = 0; c0 < 96; c0 += 8) {
        uint32_t u[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7])
                     : "r"(tm + ((uint32_t)(32 * warp) << 16) + (uint32_t)c0));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int j = 0; j < 8; ++j) Dout[(32 * warp + lane) * 96 + c0 + j] = __uint_as_float(u[j]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (tid < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tm), "r"(128u) : "memory");
}
extern "C" int probe_bf16_launch(const float* A, const float* B, uint32_t idesc, uint32_t a_lbo, uint32_t a_sbo, uint32_t b_lbo, uint32_t b_sbo, int nk,
                                 uint32_t a_step, uint32_t b_step, float* Dout) {
    size_t smem = 52 * 2048 + 64;
    cudaFuncSetAttribute(probe_bf16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe_bf16_kernel<<<1, 128, smem>>>(A, B, idesc, a_lbo, a_sbo, b_lbo, b_sbo, nk, a_step, b_step, Dout);
    return (int)cudaDeviceSynchronize();
}
