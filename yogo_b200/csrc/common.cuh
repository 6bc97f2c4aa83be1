// Shared helpers for the yogo_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/yogo_b200.h"

namespace yg {

void set_error(const char* fmt, ...);
extern unsigned long long g_launches;  // kernels launched by this library (yg_launch_count)

#define YG_CHECK_ARG(cond, ...)                 \
  do {                                          \
    if (!(cond)) {                              \
      yg::set_error(__VA_ARGS__);               \
      return YG_ERR_INVALID;                    \
    }                                           \
  } while (0)

#define YG_CUDA(expr)                                                             \
  do {                                                                            \
    cudaError_t _e = (expr);                                                      \
    if (_e != cudaSuccess) {                                                      \
      yg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),       \
                    __FILE__, __LINE__);                                          \
      return YG_ERR_CUDA;                                                         \
    }                                                                             \
  } while (0)

#define YG_LAUNCH_CHECK(name)                                                     \
  do {                                                                            \
    cudaError_t _e = cudaGetLastError();                                          \
    if (_e != cudaSuccess) {                                                      \
      yg::set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));     \
      return YG_ERR_CUDA;                                                         \
    }                                                                             \
    ++yg::g_launches;                                                             \
  } while (0)

typedef __nv_bfloat16 bf16;

template <typename T> __device__ __forceinline__ float to_f(T v);
template <> __device__ __forceinline__ float to_f<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f<uint8_t>(uint8_t v) { return (float)v; }

template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ bf16 from_f<bf16>(float v) { return __float2bfloat16_rn(v); }

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

__device__ __forceinline__ float act_fwd(float v, int act) {
  if (act == YG_ACT_LRELU) return v > 0.f ? v : 0.01f * v;
  if (act == YG_ACT_SILU) return v * sigmoidf_(v);
  return v;
}
// derivative of the activation at pre-activation `pre`
__device__ __forceinline__ float act_grad(float pre, int act) {
  if (act == YG_ACT_LRELU) return pre > 0.f ? 1.f : 0.01f;
  if (act == YG_ACT_SILU) {
    float s = sigmoidf_(pre);
    return s * (1.f + pre * (1.f - s));
  }
  return 1.f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// torch.clamp(grad, -c, c) of the reference's parameter hook (model.py:76-77) propagates NaN; fminf / fmaxf would
// return the finite operand and turn a NaN gradient into a silent -c
__device__ __forceinline__ float clampf(float v, float c) {
  return (c > 0.f && v == v) ? fminf(fmaxf(v, -c), c) : v;
}

// The shared backward epilogue (yg_bwd_epilogue semantics): g = d(loss)/d(layer output)
// -> gradient wrt the layer's conv output (or wrt BN output when the layer has BN).
// Returns the transformed gradient; xhat is set when the layer has BN.
struct BwdEpi {
  const void* saved;
  int act;
  const float* dropscale;
  const float* bn_scale;
  const float* bn_shift;
  const float* bn_mean;
  const float* bn_invstd;
  double* bn_sums;
  const void* actmask;
  int io_f32;   // internal: `saved` is read and the output written as fp32 (x3.cu)
};
inline BwdEpi make_bwd_epi(const yg_bwd_epilogue* be) {
  BwdEpi e{};
  if (be) {
    e.saved = be->saved; e.act = be->act; e.dropscale = be->dropscale;
    e.bn_scale = be->bn_scale; e.bn_shift = be->bn_shift; e.bn_mean = be->bn_mean;
    e.bn_invstd = be->bn_invstd; e.bn_sums = be->bn_sums; e.actmask = be->actmask;
  }
  return e;
}
template <typename T>
__device__ __forceinline__ float bwd_epi_apply(const BwdEpi& e, float g, long long idx, int n, int c, int C,
                                               float& xhat) {
  xhat = 0.f;
  if (e.dropscale) g *= e.dropscale[(long long)n * C + c];
  if (e.saved) {
    float sv = to_f<T>(((const T*)e.saved)[idx]);
    float pre = sv;
    if (e.bn_scale) {
      pre = sv * e.bn_scale[c] + e.bn_shift[c];
      xhat = (sv - e.bn_mean[c]) * e.bn_invstd[c];
    }
    g *= act_grad(pre, e.act);
  }
  return g;
}

struct FwdEpi {
  const float* scale;
  const float* shift;
  int act;
  const float* dropscale;
  double* stats;
  void* preact;
  void* actmask;
  int io_f32;   // internal: the output / pre-activation copy are written as fp32 (x3.cu)
};
inline FwdEpi make_fwd_epi(const yg_fwd_epilogue* ep) {
  FwdEpi e{};
  if (ep) {
    e.scale = ep->scale; e.shift = ep->shift; e.act = ep->act; e.dropscale = ep->dropscale;
    e.stats = ep->stats; e.preact = ep->preact; e.actmask = ep->actmask;
  }
  return e;
}

inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

}  // namespace yg
