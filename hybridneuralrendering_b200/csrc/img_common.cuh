// "Split images": how the training path keeps activations and gated gradients in HBM between the fused kernels.
//
// A matrix of `rows` x C fp32 values (C % 16 == 0) is stored as TWO bf16 planes, hi = bf16(v) and lo = bf16(v - hi)
// (16 mantissa bits together, fp32's exponent range: gradients need no data-dependent scale), i.e. the same 4 bytes per
// element as fp32, already in the shared-memory layout the tensor cores read:
//
//   slab s = rows 32 s .. 32 s + 31 (rows padded to a multiple of 128), C * 128 bytes:
//       [ hi plane : C/8 column groups x (32 rows x 16 B) ][ lo plane : same ]
//   byte offset of (part, row r, column c) = (r / 32) * C * 128 + part * C * 64 + (c / 8) * 512 + (r % 32) * 16 + (c % 8) * 2
//
// A slab plane is at once
//   * a canonical no-swizzle MN-major UMMA operand with the ROW index as the reduction dimension (K), the column as the
//     M/N dimension: SBO (8-column group stride) = 512 B, LBO (8-row group stride) = 128 B -- the weight-gradient kernel
//     (wgrad_img.cu) bulk-copies whole slabs and issues MMAs on them without touching a single element;
//   * 16-byte pieces that "thread = row" code reads and writes coalesced (32 consecutive rows = 512 contiguous bytes): the
//     fused forward (nbr_mlp_f16.cu) and the fused data-gradient chain (nbr_bwd_f16.cu) store their epilogue registers
//     straight into it.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace img {

constexpr int SLAB = 32;

__host__ __device__ __forceinline__ int64_t slab_bytes(int C) { return (int64_t)C * 128; }
__host__ __device__ __forceinline__ int64_t plane_bytes(int C) { return (int64_t)C * 64; }
// offset of the 16-byte piece (row, column group g) of the hi plane
__host__ __device__ __forceinline__ int64_t piece_off(int64_t row, int g, int C) {
    return (row >> 5) * slab_bytes(C) + (int64_t)g * 512 + (row & 31) * 16;
}

// two fp32 -> packed bf16x2 {low half = a, high half = b}, round to nearest even
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
    return r;
}
__device__ __forceinline__ float bf16_lo_f(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf16_hi_f(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }

// 8 fp32 values -> 8 bf16 hi (one 16-byte piece) and 8 bf16 lo = bf16(v - hi)
__device__ __forceinline__ void split8_bf16(const float* v, uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float a = v[2 * i], b = v[2 * i + 1];
        h[i] = pack_bf16(a, b);
        l[i] = pack_bf16(a - bf16_lo_f(h[i]), b - bf16_hi_f(h[i]));
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
// hi + lo pieces -> 8 fp32 values
__device__ __forceinline__ void join8_bf16(const uint4& hi, const uint4& lo, float* v) {
    const uint32_t h[4] = {hi.x, hi.y, hi.z, hi.w}, l[4] = {lo.x, lo.y, lo.z, lo.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = bf16_lo_f(h[i]) + bf16_lo_f(l[i]);
        v[2 * i + 1] = bf16_hi_f(h[i]) + bf16_hi_f(l[i]);
    }
}

}  // namespace img
