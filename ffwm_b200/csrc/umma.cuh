// tcgen05 / mbarrier building blocks shared by the tensor-core convolution kernels
// (conv3x3_tc.cu, conv_gen_tc.cu: forward + data gradient; conv_gen_wgrad_tc.cu: weight gradient; corr_max.cu).
// Every encoding here was validated on a B200 through conv3x3_tc.cu's parity tests.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

#include "common.cuh"

namespace ffwm {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start address,
// leading byte offset (between the two 16-byte k-chunks), stride byte offset (between 8-row
// groups), all in 16-byte units; version 1 (Blackwell); layout type 0 (SWIZZLE_NONE).
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

// Instruction descriptor (cute::UMMA::InstrDescriptor): D = f32, A = B = tf32, both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int m, int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// D = f32, A = B = bf16 (kind::f16: format code 1), both K-major.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}

// Bounded wait: a protocol error traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_u32(bar);
    for (uint32_t spin = 0;; ++spin) {
        uint32_t done;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(done)
            : "r"(a), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > (1u << 26)) __trap();
    }
}

__device__ __forceinline__ void umma_tf32_acc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc)
        : "memory");
}

__device__ __forceinline__ void split_store(unsigned char* d, int part_bytes, const float* v) {
    float4 hi, lo;
    hi.x = __uint_as_float(__float_as_uint(v[0]) & 0xffffe000u); lo.x = v[0] - hi.x;
    hi.y = __uint_as_float(__float_as_uint(v[1]) & 0xffffe000u); lo.y = v[1] - hi.y;
    hi.z = __uint_as_float(__float_as_uint(v[2]) & 0xffffe000u); lo.z = v[2] - hi.z;
    hi.w = __uint_as_float(__float_as_uint(v[3]) & 0xffffe000u); lo.w = v[3] - hi.w;
    *reinterpret_cast<float4*>(d) = hi;
    *reinterpret_cast<float4*>(d + part_bytes) = lo;
}

__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}

// one MMA, both operands from shared memory, of the kind the operand math uses
template <bool BF>
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
    if (BF) umma_bf16(d_tmem, a_desc, b_desc, idesc, accumulate);
    else umma_tf32(d_tmem, a_desc, b_desc, idesc, accumulate);
}

// 8 fp32 -> one 16-byte slot of b1 = bf16_rn(v) and one of b2 = bf16_rn(v - b1); element j at byte 2j.
__device__ __forceinline__ void split_store_bf(unsigned char* d, int part_bytes, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __nv_bfloat162 hb = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);       // .x (low half) = v[2i]
        const __nv_bfloat162 lb = __floats2bfloat162_rn(v[2 * i] - __low2float(hb), v[2 * i + 1] - __high2float(hb));
        h[i] = *reinterpret_cast<const uint32_t*>(&hb);
        l[i] = *reinterpret_cast<const uint32_t*>(&lb);
    }
    *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(d + part_bytes) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <bool BF>
__device__ __forceinline__ void split_store_m(unsigned char* d, int part_bytes, const float* v) {
    if (BF) split_store_bf(d, part_bytes, v);
    else split_store(d, part_bytes, v);
}

}  // namespace ffwm
