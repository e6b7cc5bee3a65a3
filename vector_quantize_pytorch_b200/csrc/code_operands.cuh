// Tensor-core operands of one codebook row (shared by vqb_codebook_prepare and vqb_ema_apply).
#pragma once
#include "vqb_common.cuh"

namespace vqb {

// ---------------------------------------------------------------------------------------------
// codebook operands.  Shared by vqb_codebook_prepare and vqb_ema_apply (vq_ema.cu).
// One warp owns one (padded) code row.  `vals(i)` yields c[i] in fp32.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void write_code_operands(const float* crow /*K x D row or nullptr for padding*/, int k, int K, int Kpad, int D,
                                    int metric, uint16_t* planes, uint16_t* bext, float* bias, float* cnorm2, float* cmax, int lane) {
  uint16_t* hi = planes + static_cast<int64_t>(k) * D;
  uint16_t* lo = planes + (static_cast<int64_t>(Kpad) + k) * D;
  if (crow == nullptr) {  // padding row: never wins (bias = +inf), contributes zeros to the MMA
    for (int i = lane; i < D; i += 32) { hi[i] = 0; lo[i] = 0; }
    if (lane == 0) bias[k] = INFINITY;
    if (lane < 16) bext[k * 16 + lane] = (lane == 0) ? float_to_bf16_bits(-3.0e38f) : 0;  // score = -huge: never wins
    return;
  }
  double n2 = 0.0;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 c = *reinterpret_cast<const float4*>(crow + i);
    const float v[4] = {c.x, c.y, c.z, c.w};
    uint16_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      h[e] = float_to_bf16_bits(v[e]);
      l[e] = float_to_bf16_bits(v[e] - bf16_bits_to_float(h[e]));
      n2 += static_cast<double>(v[e]) * static_cast<double>(v[e]);
    }
    *reinterpret_cast<uint2*>(hi + i) = make_uint2(h[0] | (uint32_t(h[1]) << 16), h[2] | (uint32_t(h[3]) << 16));
    *reinterpret_cast<uint2*>(lo + i) = make_uint2(l[0] | (uint32_t(l[1]) << 16), l[2] | (uint32_t(l[3]) << 16));
  }
  n2 = warp_sum(n2);
  if (lane == 0) {
    const float n2f = static_cast<float>(n2);
    cnorm2[k] = n2f;
    const float b = (metric == VQB_METRIC_EUCLID) ? 0.5f * n2f : 0.f;
    bias[k] = b;
    // -bias as three bf16 terms (8+8+8 mantissa bits = the exact fp32 value): the K=16 "bias MMA" of the
    // search kernel multiplies them by [1 1 1 0...] and so seeds the accumulator with -0.5||c||^2.
    const uint16_t b1 = float_to_bf16_bits(b);
    const float r1 = b - bf16_bits_to_float(b1);
    const uint16_t b2 = float_to_bf16_bits(r1);
    const uint16_t b3 = float_to_bf16_bits(r1 - bf16_bits_to_float(b2));
    uint16_t* row = bext + k * 16;
    row[0] = b1 ^ 0x8000; row[1] = b2 ^ 0x8000; row[2] = b3 ^ 0x8000;  // sign flip = negate
#pragma unroll
    for (int j = 3; j < 16; ++j) row[j] = 0;
    atomicMax(reinterpret_cast<unsigned int*>(cmax), __float_as_uint(sqrtf(n2f)));  // valid: values are >= 0
  }
}


}  // namespace vqb
