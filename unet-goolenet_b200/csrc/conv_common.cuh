// Device helpers shared by the implicit-GEMM kernels (conv_gemm.cu, conv3x3_halo.cu).
#pragma once
#include "common.cuh"
#include "engine.h"

namespace ug {

static constexpr int kThreads = 192;               // warp 0 TMA, warp 1 MMA, warps 2-5 epilogue
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

template <int kAct>
__device__ __forceinline__ void epi_math16(const uint32_t (&v)[16], float (&f)[16], const float* sScale,
                                           const float* sBias, int col) {
#pragma unroll
  for (int j = 0; j < 16; ++j) f[j] = apply_act<kAct>(__uint_as_float(v[j]) * sScale[col + j] + sBias[col + j]);
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
