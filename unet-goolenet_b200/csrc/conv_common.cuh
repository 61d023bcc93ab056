// Device helpers shared by the implicit-GEMM kernels (conv_gemm.cu, conv_multi.cu, stem_conv.cu).
#pragma once
#include "common.cuh"
#include "engine.h"

namespace ug {

// Warp roles of the 192-thread GEMM kernels: warps 0-3 epilogue (TMEM lane quarter = warp id), warp 4 TMA producer,
// warp 5 TMEM allocator + MMA issuer.  The scheduler prefers the highest eligible warp id, so the latency-critical
// issue warps sit above the epilogue warps (see conv_multi.cu).
static constexpr int kThreads = 192;
static constexpr int kProducerWarp = 4;
static constexpr int kMmaWarp = 5;
static constexpr int kABytesPerStage = 128 * 128;  // 128 rows x 64 bf16

// kAct is a template parameter on purpose: with a run-time activation switch the compiler if-converts the erf
// polynomial of GELU into predicated code inside the unrolled per-element loop, and every ReLU epilogue then
// issues ~40 dead instructions per element (measured: ~1500 cycles per 16-column chunk).
template <int kAct>
__device__ __forceinline__ float apply_act(float t) {
  if constexpr (kAct == UG_ACT_RELU) return fmaxf(t, 0.0f);
  else if constexpr (kAct == UG_ACT_GELU) return gelu_erf(t);
  else return t;
}

// folded BN / bias + activation of 16 consecutive channels starting at `col` (a multiple of 16); scale and bias
// live in shared memory and are read as float4
template <int kAct>
__device__ __forceinline__ void epi_math16(const uint32_t (&v)[16], float (&f)[16], const float* sScale,
                                           const float* sBias, int col) {
  const float4* s4 = reinterpret_cast<const float4*>(sScale + col);
  const float4* b4 = reinterpret_cast<const float4*>(sBias + col);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float4 sc = s4[j], bi = b4[j];
    f[4 * j + 0] = apply_act<kAct>(__uint_as_float(v[4 * j + 0]) * sc.x + bi.x);
    f[4 * j + 1] = apply_act<kAct>(__uint_as_float(v[4 * j + 1]) * sc.y + bi.y);
    f[4 * j + 2] = apply_act<kAct>(__uint_as_float(v[4 * j + 2]) * sc.z + bi.z);
    f[4 * j + 3] = apply_act<kAct>(__uint_as_float(v[4 * j + 3]) * sc.w + bi.w);
  }
}

// Deferred ReLU: kernels instantiated for ReLU compute only scale*acc + bias in epi_math16_linear and apply the
// max(.,0) where the value is consumed — folded into the bf16 conversion for plain stores (pack_bf16x2_relu),
// explicitly before a residual add / gate / outc dot.
template <int kAct>
__device__ __forceinline__ void epi_math16_linear(const uint32_t (&v)[16], float (&f)[16], const float* sScale,
                                                  const float* sBias, int col) {
  if constexpr (kAct == UG_ACT_RELU) epi_math16<UG_ACT_NONE>(v, f, sScale, sBias, col);
  else epi_math16<kAct>(v, f, sScale, sBias, col);
}
template <int kAct>
__device__ __forceinline__ void epi_relu16(float (&f)[16]) {
  if constexpr (kAct == UG_ACT_RELU) {
#pragma unroll
    for (int j = 0; j < 16; ++j) f[j] = fmaxf(f[j], 0.0f);
  }
}
template <int kAct>
__device__ __forceinline__ uint32_t epi_pack2(float lo, float hi) {
  if constexpr (kAct == UG_ACT_RELU) return pack_bf16x2_relu(lo, hi);
  else return pack_bf16x2(lo, hi);
}

// residual add of 8 bf16 values held in a register quad
__device__ __forceinline__ void epi_add8(float* f, const uint4& a) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] += bf16_lo(aw[j]);
    f[2 * j + 1] += bf16_hi(aw[j]);
  }
}
// CoordAtt3 combine e1 + d*(1+g) (basicUnet.py:229) with g1 = 1 + g of 8 channels
__device__ __forceinline__ void epi_gate8(float* f, const uint4& a, const float4& g1a, const float4& g1b) {
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
  const float g1[8] = {g1a.x, g1a.y, g1a.z, g1a.w, g1b.x, g1b.y, g1b.z, g1b.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    f[2 * j] = bf16_lo(aw[j]) + f[2 * j] * g1[2 * j];
    f[2 * j + 1] = bf16_hi(aw[j]) + f[2 * j + 1] * g1[2 * j + 1];
  }
}

__device__ __forceinline__ void epi_add_gate8(const ConvKParams& p, float* f, const __nv_bfloat16* add_ptr,
                                              const float* gate_ptr) {
  const uint4 a = *reinterpret_cast<const uint4*>(add_ptr);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float e0 = bf16_lo(aw[j]), e1 = bf16_hi(aw[j]);
    if (p.mode == UG_EPI_ADD) {
      f[2 * j] += e0;
      f[2 * j + 1] += e1;
    } else {
      f[2 * j] = e0 + f[2 * j] * (1.0f + __ldg(gate_ptr + 2 * j));
      f[2 * j + 1] = e1 + f[2 * j + 1] * (1.0f + __ldg(gate_ptr + 2 * j + 1));
    }
  }
}


}  // namespace ug
