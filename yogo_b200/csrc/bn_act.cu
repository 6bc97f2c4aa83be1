// BatchNorm2d pieces and the elementwise normalise+activate pass.
// Replaces nn.BatchNorm2d (train: batch statistics + running-stat update, eval: affine with
// running stats) and the LeakyReLU/SiLU/Dropout2d kernels that follow it in the reference
// (/root/reference/yogo/model_defns.py:35-36, 55-56, 60-61).
#include "common.cuh"

namespace yg {

__global__ void bn_finalize_kernel(const double* __restrict__ stats, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* running_mean, float* running_var,
                                   float momentum, float eps, float* mean, float* invstd, float* scale,
                                   float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double m = stats[c] / count;
  double var = stats[C + c] / count - m * m;
  if (var < 0) var = 0;
  const float is = (float)(1.0 / sqrt(var + (double)eps));
  const float g = gamma ? gamma[c] : 1.f, b = beta ? beta[c] : 0.f;
  mean[c] = (float)m;
  invstd[c] = is;
  scale[c] = g * is;
  shift[c] = b - (float)m * g * is;
  if (running_mean) {
    // torch: running = (1-momentum)*running + momentum*batch, variance unbiased
    const double unb = count > 1 ? var * count / (count - 1.0) : var;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * (float)m;
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * (float)unb;
  }
}

__global__ void bn_fold_eval_kernel(const float* gamma, const float* beta, const float* rm, const float* rv,
                                    const float* conv_bias, float eps, float* scale, float* shift, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float is = 1.f / sqrtf(rv[c] + eps);
  const float s = (gamma ? gamma[c] : 1.f) * is;
  scale[c] = s;
  shift[c] = (beta ? beta[c] : 0.f) - rm[c] * s + (conv_bias ? conv_bias[c] * s : 0.f);
}

// elementwise over NHWC with C innermost; 8 elements per thread, 16-byte accesses when C % 8 == 0
template <typename T>
__global__ void bn_act_apply_kernel(const T* __restrict__ y, T* __restrict__ a, long long total, int HW, int C,
                                    const float* __restrict__ scale, const float* __restrict__ shift, int act,
                                    const float* __restrict__ dropscale) {
  const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (i0 >= total) return;
  if ((C & 7) == 0) {
    const int c0 = (int)(i0 % C);
    const int n = (int)(i0 / ((long long)HW * C));
    __align__(16) T in[8];
    __align__(16) T out[8];
    constexpr int V = (int)(sizeof(T) * 8 / 16);
#pragma unroll
    for (int v = 0; v < V; ++v)
      reinterpret_cast<uint4*>(in)[v] = reinterpret_cast<const uint4*>(y + i0)[v];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const int c = c0 + k;
      float v = to_f<T>(in[k]) * scale[c] + shift[c];
      v = act_fwd(v, act);
      if (dropscale) v *= dropscale[(long long)n * C + c];
      out[k] = from_f<T>(v);
    }
#pragma unroll
    for (int v = 0; v < V; ++v)
      reinterpret_cast<uint4*>(a + i0)[v] = reinterpret_cast<uint4*>(out)[v];
  } else {
    for (long long i = i0; i < i0 + 8 && i < total; ++i) {
      const int c = (int)(i % C);
      const int n = (int)(i / ((long long)HW * C));
      float v = to_f<T>(y[i]) * scale[c] + shift[c];
      v = act_fwd(v, act);
      if (dropscale) v *= dropscale[(long long)n * C + c];
      a[i] = from_f<T>(v);
    }
  }
}

// dz = gamma*invstd*(g - m1 - xhat*m2) = g*A[c] + y*B[c] + K[c] with per-channel constants staged in smem
// (the fp64 divisions happen once per channel per block, not once per element)
template <typename T>
__global__ void bn_bwd_apply_kernel(T* __restrict__ g, const T* __restrict__ y, long long total, int C, double M,
                                    const double* __restrict__ sums, const float* __restrict__ gamma,
                                    const float* __restrict__ mean, const float* __restrict__ invstd,
                                    int batch_stats) {
  extern __shared__ float kc[];  // [3][C]: A, B, K
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m1 = batch_stats ? (float)(sums[c] / M) : 0.f, m2 = batch_stats ? (float)(sums[C + c] / M) : 0.f;
    const float gi = (gamma ? gamma[c] : 1.f) * invstd[c];
    // xhat = (y - mean)*invstd  ->  dz = gi*g - gi*m2*invstd*y + gi*(m2*invstd*mean - m1)
    kc[c] = gi;
    kc[C + c] = -gi * m2 * invstd[c];
    kc[2 * C + c] = gi * (m2 * invstd[c] * mean[c] - m1);
  }
  __syncthreads();
  for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < total;
       i0 += (long long)gridDim.x * blockDim.x * 8) {
    if ((C & 7) == 0) {
      const int c0 = (int)(i0 % C);
      __align__(16) T gin[8];
      __align__(16) T yin[8];
      constexpr int V = (int)(sizeof(T) * 8 / 16);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        reinterpret_cast<uint4*>(gin)[v] = reinterpret_cast<const uint4*>(g + i0)[v];
        reinterpret_cast<uint4*>(yin)[v] = reinterpret_cast<const uint4*>(y + i0)[v];
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        gin[k] = from_f<T>(to_f<T>(gin[k]) * kc[c] + to_f<T>(yin[k]) * kc[C + c] + kc[2 * C + c]);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(g + i0)[v] = reinterpret_cast<uint4*>(gin)[v];
    } else {
      for (long long i = i0; i < i0 + 8 && i < total; ++i) {
        const int c = (int)(i % C);
        g[i] = from_f<T>(to_f<T>(g[i]) * kc[c] + to_f<T>(y[i]) * kc[C + c] + kc[2 * C + c]);
      }
    }
  }
}

__global__ void bn_param_grad_kernel(const double* __restrict__ sums, float* dgamma, float* dbeta, int C, float clip) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  if (dbeta) dbeta[c] = clampf((float)sums[c], clip);
  if (dgamma) dgamma[c] = clampf((float)sums[C + c], clip);
}

}  // namespace yg
using namespace yg;

extern "C" int yg_bn_finalize(const double* stats, double count, const float* gamma, const float* beta,
                              float* running_mean, float* running_var, float momentum, float eps,
                              float* mean, float* invstd, float* scale, float* shift, int C, void* stream) {
  YG_CHECK_ARG(stats && mean && invstd && scale && shift && C > 0 && count > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(stats, count, gamma, beta, running_mean,
                                                                      running_var, momentum, eps, mean, invstd,
                                                                      scale, shift, C);
  YG_LAUNCH_CHECK("bn_finalize");
  return YG_OK;
}

extern "C" int yg_bn_fold_eval(const float* gamma, const float* beta, const float* running_mean,
                               const float* running_var, const float* conv_bias, float eps,
                               float* scale, float* shift, int C, void* stream) {
  YG_CHECK_ARG(running_mean && running_var && scale && shift && C > 0, "bn_fold_eval: bad arguments");
  bn_fold_eval_kernel<<<cdiv(C, 128), 128, 0, (cudaStream_t)stream>>>(gamma, beta, running_mean, running_var,
                                                                       conv_bias, eps, scale, shift, C);
  YG_LAUNCH_CHECK("bn_fold_eval");
  return YG_OK;
}

extern "C" int yg_bn_act_apply(const void* y, void* a, int dtype, int N, int HW, int C,
                               const float* scale, const float* shift, int act, const float* dropscale,
                               void* stream) {
  YG_CHECK_ARG(y && a && scale && shift, "bn_act_apply: null pointer");
  const long long total = (long long)N * HW * C;
  if (total == 0) return YG_OK;
  const int blocks = cdiv(cdiv(total, 8), 256);
  if (dtype == YG_BF16)
    bn_act_apply_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, (bf16*)a, total, HW, C, scale, shift, act, dropscale);
  else
    bn_act_apply_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)y, (float*)a, total, HW, C, scale, shift, act, dropscale);
  YG_LAUNCH_CHECK("bn_act_apply");
  return YG_OK;
}

extern "C" int yg_bn_bwd_apply(void* g, const void* y, int dtype, int N, int HW, int C,
                               const double* sums, const float* gamma, const float* mean, const float* invstd,
                               float* dgamma, float* dbeta, float clip, int batch_stats, void* stream) {
  YG_CHECK_ARG(sums != nullptr, "bn_bwd_apply: sums is null");
  YG_CHECK_ARG(!g || (mean && invstd), "bn_bwd_apply: mean/invstd are null");
  const long long total = (long long)N * HW * C;
  cudaStream_t st = (cudaStream_t)stream;
  if (total && g) {
    YG_CHECK_ARG(y != nullptr, "bn_bwd_apply: y is null");
    int blocks = cdiv(cdiv(total, 8), 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    const double M = (double)N * HW;
    const size_t sm = (size_t)3 * C * sizeof(float);
    if (dtype == YG_BF16)
      bn_bwd_apply_kernel<bf16><<<blocks, 256, sm, st>>>((bf16*)g, (const bf16*)y, total, C, M, sums, gamma, mean, invstd, batch_stats);
    else
      bn_bwd_apply_kernel<float><<<blocks, 256, sm, st>>>((float*)g, (const float*)y, total, C, M, sums, gamma, mean, invstd, batch_stats);
    YG_LAUNCH_CHECK("bn_bwd_apply");
  }
  if (dgamma || dbeta) {
    bn_param_grad_kernel<<<cdiv(C, 128), 128, 0, st>>>(sums, dgamma, dbeta, C, clip);
    YG_LAUNCH_CHECK("bn_param_grad");
  }
  return YG_OK;
}
