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
                                    const float* __restrict__ dropscale, unsigned char* __restrict__ actmask) {
  constexpr int U = 4;   // 8-element vectors per thread: all loads are issued before the first use (bytes in flight)
  const long long base = ((long long)blockIdx.x * blockDim.x * U + threadIdx.x) * 8;
  if ((C & 7) == 0) {
    constexpr int V = (int)(sizeof(T) * 8 / 16);
    __align__(16) T in[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i0 = base + (long long)u * blockDim.x * 8;
      if (i0 < total) {
#pragma unroll
        for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(in[u])[v] = __ldg(reinterpret_cast<const uint4*>(y + i0) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long i0 = base + (long long)u * blockDim.x * 8;
      if (i0 >= total) continue;
      const int c0 = (int)(i0 % C);
      const int n = (int)(i0 / ((long long)HW * C));
      __align__(16) T out[8];
      unsigned bits = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int c = c0 + k;
        float v = to_f<T>(in[u][k]) * scale[c] + shift[c];
        bits |= (v > 0.f ? 1u : 0u) << k;
        v = act_fwd(v, act);
        if (dropscale) v *= dropscale[(long long)n * C + c];
        out[k] = from_f<T>(v);
      }
#pragma unroll
      for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(a + i0)[v] = reinterpret_cast<uint4*>(out)[v];
      if (actmask) actmask[i0 >> 3] = (unsigned char)bits;   // sign bits of the activation input (yg_fwd_epilogue.actmask)
    }
  } else {
    for (int u = 0; u < U; ++u) {
      const long long i0 = base + (long long)u * blockDim.x * 8;
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
  const long long stride = (long long)gridDim.x * blockDim.x * 8;
  if ((C & 7) == 0) {
    constexpr int U = 2;   // two independent 8-element vectors per iteration: four 16-byte loads in flight per thread
    constexpr int V = (int)(sizeof(T) * 8 / 16);
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < total; i0 += U * stride) {
      __align__(16) T gin[U][8];
      __align__(16) T yin[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        if (i < total) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            reinterpret_cast<uint4*>(gin[u])[v] = reinterpret_cast<const uint4*>(g + i)[v];
            reinterpret_cast<uint4*>(yin[u])[v] = __ldg(reinterpret_cast<const uint4*>(y + i) + v);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long i = i0 + u * stride;
        if (i >= total) continue;
        const int c0 = (int)(i % C);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int c = c0 + k;
          gin[u][k] = from_f<T>(to_f<T>(gin[u][k]) * kc[c] + to_f<T>(yin[u][k]) * kc[C + c] + kc[2 * C + c]);
        }
#pragma unroll
        for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(g + i)[v] = reinterpret_cast<uint4*>(gin[u])[v];
      }
    }
  } else {
    for (long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8; i0 < total; i0 += stride)
      for (long long i = i0; i < i0 + 8 && i < total; ++i) {
        const int c = (int)(i % C);
        g[i] = from_f<T>(to_f<T>(g[i]) * kc[c] + to_f<T>(y[i]) * kc[C + c] + kc[2 * C + c]);
      }
  }
}

// Channel-group variants of the two elementwise passes (C % 8 == 0): thread = (8-channel group, pixel lane) like
// bn_colsums_kernel, so the per-channel constants are loaded ONCE into registers and the inner loop is
// 16-byte load(s) -> 8 FMAs -> 16-byte store.  (The flat-index kernels above fetch every constant per element with a
// lane stride of 8 floats: 8 L1 wavefronts / 8-way bank conflicts per load, which capped them at ~1.8 TB/s.)
template <typename T>
__global__ void __launch_bounds__(256) bn_act_apply_cg_kernel(const T* __restrict__ y, T* __restrict__ a, long long npix, int HW,
                                                              int C, const float* __restrict__ scale,
                                                              const float* __restrict__ shift, int act,
                                                              const float* __restrict__ dropscale,
                                                              unsigned char* __restrict__ actmask) {
  const int groups = C >> 3, lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  if (pl >= lanes) return;
  float sc[8], sh[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { sc[k] = scale[cg * 8 + k]; sh[k] = shift[cg * 8 + k]; }
  constexpr int U = 4;
  constexpr int V = (int)(sizeof(T) * 8 / 16);
  const long long pstride = (long long)gridDim.x * lanes;
  for (long long px0 = (long long)blockIdx.x * lanes + pl; px0 < npix; px0 += U * pstride) {
    __align__(16) T in[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long px = px0 + u * pstride;
      if (px < npix) {
#pragma unroll
        for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(in[u])[v] = __ldg(reinterpret_cast<const uint4*>(y + px * C + cg * 8) + v);
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long px = px0 + u * pstride;
      if (px >= npix) break;
      const float* dsr = dropscale ? dropscale + (px / HW) * C + cg * 8 : nullptr;
      __align__(16) T out[8];
      unsigned bits = 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float v = to_f<T>(in[u][k]) * sc[k] + sh[k];
        bits |= (v > 0.f ? 1u : 0u) << k;
        v = act_fwd(v, act);
        if (dsr) v *= dsr[k];
        out[k] = from_f<T>(v);
      }
      const long long i0 = px * C + cg * 8;
#pragma unroll
      for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(a + i0)[v] = reinterpret_cast<uint4*>(out)[v];
      if (actmask) actmask[i0 >> 3] = (unsigned char)bits;
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256) bn_bwd_apply_cg_kernel(T* __restrict__ g, const T* __restrict__ y, long long npix, int C,
                                                              double M, const double* __restrict__ sums,
                                                              const float* __restrict__ gamma, const float* __restrict__ mean,
                                                              const float* __restrict__ invstd, int batch_stats,
                                                              float* __restrict__ dgamma, float* __restrict__ dbeta, float clip) {
  extern __shared__ float kc[];  // [3][C]: A, B, K (same constants as bn_bwd_apply_kernel)
  if (blockIdx.x == 0) {   // the parameter gradients are the two sums themselves (bn_param_grad_kernel in the same launch)
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
      if (dbeta) dbeta[c] = clampf((float)sums[c], clip);
      if (dgamma) dgamma[c] = clampf((float)sums[C + c], clip);
    }
  }
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const float m1 = batch_stats ? (float)(sums[c] / M) : 0.f, m2 = batch_stats ? (float)(sums[C + c] / M) : 0.f;
    const float gi = (gamma ? gamma[c] : 1.f) * invstd[c];
    kc[c] = gi;
    kc[C + c] = -gi * m2 * invstd[c];
    kc[2 * C + c] = gi * (m2 * invstd[c] * mean[c] - m1);
  }
  __syncthreads();
  const int groups = C >> 3, lanes = blockDim.x / groups;
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  if (pl >= lanes) return;
  float ka[8], kb[8], kk[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) { ka[k] = kc[cg * 8 + k]; kb[k] = kc[C + cg * 8 + k]; kk[k] = kc[2 * C + cg * 8 + k]; }
  constexpr int U = 4;
  constexpr int V = (int)(sizeof(T) * 8 / 16);
  const long long pstride = (long long)gridDim.x * lanes;
  for (long long px0 = (long long)blockIdx.x * lanes + pl; px0 < npix; px0 += U * pstride) {
    __align__(16) T gin[U][8];
    __align__(16) T yin[U][8];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long px = px0 + u * pstride;
      if (px < npix) {
#pragma unroll
        for (int v = 0; v < V; ++v) {
          reinterpret_cast<uint4*>(gin[u])[v] = reinterpret_cast<const uint4*>(g + px * C + cg * 8)[v];
          reinterpret_cast<uint4*>(yin[u])[v] = __ldg(reinterpret_cast<const uint4*>(y + px * C + cg * 8) + v);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const long long px = px0 + u * pstride;
      if (px >= npix) break;
#pragma unroll
      for (int k = 0; k < 8; ++k)
        gin[u][k] = from_f<T>(to_f<T>(gin[u][k]) * ka[k] + to_f<T>(yin[u][k]) * kb[k] + kk[k]);
#pragma unroll
      for (int v = 0; v < V; ++v) reinterpret_cast<uint4*>(g + px * C + cg * 8)[v] = reinterpret_cast<uint4*>(gin[u])[v];
    }
  }
}

// Per-channel sums over an NHWC tensor with 16-byte loads: MODE 0: sum(y), sum(y*y) (BatchNorm batch statistics);
// MODE 1: sum(g), sum(g * xhat) with xhat = (y - mean) * invstd (BatchNorm backward).  One streaming pass at HBM speed
// (cheaper than a 31-shuffle transpose-reduce per 16-column chunk in the epilogue of the producing convolution).
// Thread t owns the 8-channel group t % (C/8) and pixels t / (C/8) + k * stride.
template <typename T, int MODE, int U = 4>
__global__ void __launch_bounds__(256) bn_colsums_kernel(const T* __restrict__ a, const T* __restrict__ y, long long npix, int C,
                                                         const float* __restrict__ mean, const float* __restrict__ invstd,
                                                         double* __restrict__ sums) {
  extern __shared__ float red[];   // [lanes][2][C]: one row per pixel lane, summed in a fixed order in fp64
  const int groups = C / 8;
  const int lanes = blockDim.x / groups;           // pixel lanes per block
  const int cg = threadIdx.x % groups, pl = threadIdx.x / groups;
  float s1[8], s2[8], mu[8], is[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    s1[k] = 0.f; s2[k] = 0.f;
    mu[k] = (MODE == 1) ? mean[cg * 8 + k] : 0.f;
    is[k] = (MODE == 1) ? invstd[cg * 8 + k] : 1.f;
  }
  if (pl < lanes) {
    constexpr int V = (int)(sizeof(T) * 8 / 16);
    // U independent 16-byte loads in flight per thread; the per-thread summation order is unchanged
    const long long pstride = (long long)gridDim.x * lanes;
    for (long long px0 = (long long)blockIdx.x * lanes + pl; px0 < npix; px0 += U * pstride) {
      __align__(16) T av[U][8];
      __align__(16) T yv[U][8];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const long long off = (px0 + u * pstride) * C + cg * 8;
        if (px0 + u * pstride < npix) {
#pragma unroll
          for (int v = 0; v < V; ++v) {
            reinterpret_cast<uint4*>(av[u])[v] = __ldg(reinterpret_cast<const uint4*>(a + off) + v);
            if (MODE == 1) reinterpret_cast<uint4*>(yv[u])[v] = __ldg(reinterpret_cast<const uint4*>(y + off) + v);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (px0 + u * pstride >= npix) break;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float x = to_f<T>(av[u][k]);
          s1[k] += x;
          s2[k] += (MODE == 0) ? x * x : x * ((to_f<T>(yv[u][k]) - mu[k]) * is[k]);
        }
      }
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      red[(size_t)pl * 2 * C + cg * 8 + k] = s1[k];
      red[(size_t)pl * 2 * C + C + cg * 8 + k] = s2[k];
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) {
    double acc = 0.0;
    for (int l = 0; l < lanes; ++l) acc += (double)red[(size_t)l * 2 * C + i];
    atomicAdd(&sums[i], acc);
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
                               void* actmask, void* stream) {
  YG_CHECK_ARG(y && a && scale && shift, "bn_act_apply: null pointer");
  YG_CHECK_ARG(!actmask || (C & 7) == 0, "bn_act_apply: actmask needs C % 8 == 0");
  const long long total = (long long)N * HW * C;
  if (total == 0) return YG_OK;
  if ((C & 7) == 0 && C / 8 <= 256) {
    const long long npix = (long long)N * HW;
    const int lanes = 256 / (C / 8);
    long long want = cdiv(cdiv(npix, (long long)lanes), 4LL);
    const int blocks_cg = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
    if (dtype == YG_BF16)
      bn_act_apply_cg_kernel<bf16><<<blocks_cg, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, (bf16*)a, npix, HW, C, scale, shift, act, dropscale, (unsigned char*)actmask);
    else
      bn_act_apply_cg_kernel<float><<<blocks_cg, 256, 0, (cudaStream_t)stream>>>((const float*)y, (float*)a, npix, HW, C, scale, shift, act, dropscale, (unsigned char*)actmask);
    YG_LAUNCH_CHECK("bn_act_apply");
    return YG_OK;
  }
  const int blocks = cdiv(cdiv(total, 8), 256 * 4);   // 4 vectors of 8 elements per thread
  if (dtype == YG_BF16)
    bn_act_apply_kernel<bf16><<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)y, (bf16*)a, total, HW, C, scale, shift, act, dropscale, (unsigned char*)actmask);
  else
    bn_act_apply_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>((const float*)y, (float*)a, total, HW, C, scale, shift, act, dropscale, (unsigned char*)actmask);
  YG_LAUNCH_CHECK("bn_act_apply");
  return YG_OK;
}

static int bn_colsums(const void* a, const void* y, int dtype, long long npix, int C, const float* mean, const float* invstd,
                      double* sums, int mode, cudaStream_t st) {
  YG_CHECK_ARG(C % 8 == 0 && C / 8 <= 256, "bn sums: C must be a multiple of 8 and <= 2048, got %d", C);
  if (npix == 0) return YG_OK;
  const int lanes = 256 / (C / 8);
  long long want = (npix + lanes - 1) / lanes;
  // two blocks per SM: measured best on 205 MB tensors (fewer blocks = fewer fp64 atomics per address at the end; 8 per SM: 47.5 / 74.4 us,
  // 2 per SM with 8 loads in flight for the statistics pass: 39.4 / 68.3 us = 5.2 / 6.0 TB/s)
  const int blocks = (int)(want < 148 * 2 ? want : 148 * 2);
  const size_t sm = (size_t)lanes * 2 * C * sizeof(float);   // <= 256/(C/8) * 2C * 4 = 16 KB
  if (dtype == YG_BF16) {
    if (mode == 0) bn_colsums_kernel<bf16, 0, 8><<<blocks, 256, sm, st>>>((const bf16*)a, nullptr, npix, C, nullptr, nullptr, sums);
    else bn_colsums_kernel<bf16, 1><<<blocks, 256, sm, st>>>((const bf16*)a, (const bf16*)y, npix, C, mean, invstd, sums);
  } else {
    if (mode == 0) bn_colsums_kernel<float, 0><<<blocks, 256, sm, st>>>((const float*)a, nullptr, npix, C, nullptr, nullptr, sums);
    else bn_colsums_kernel<float, 1><<<blocks, 256, sm, st>>>((const float*)a, (const float*)y, npix, C, mean, invstd, sums);
  }
  YG_LAUNCH_CHECK("bn_colsums");
  return YG_OK;
}

extern "C" int yg_bn_stats(const void* y, int dtype, int N, int HW, int C, double* stats, void* stream) {
  YG_CHECK_ARG(y && stats, "bn_stats: null pointer");
  YG_CHECK_ARG(dtype == YG_F32 || dtype == YG_BF16, "bn_stats: dtype %d", dtype);
  return bn_colsums(y, nullptr, dtype, (long long)N * HW, C, nullptr, nullptr, stats, 0, (cudaStream_t)stream);
}

extern "C" int yg_bn_bwd_sums(const void* g, const void* y, int dtype, int N, int HW, int C, const float* mean,
                              const float* invstd, double* sums, void* stream) {
  YG_CHECK_ARG(g && y && mean && invstd && sums, "bn_bwd_sums: null pointer");
  YG_CHECK_ARG(dtype == YG_F32 || dtype == YG_BF16, "bn_bwd_sums: dtype %d", dtype);
  return bn_colsums(g, y, dtype, (long long)N * HW, C, mean, invstd, sums, 1, (cudaStream_t)stream);
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
    if ((C & 7) == 0 && C / 8 <= 256) {
      const long long npix = (long long)N * HW;
      const int lanes = 256 / (C / 8);
      long long want = cdiv(cdiv(npix, (long long)lanes), 4LL);
      const int blocks_cg = (int)(want < 148 * 8 ? (want < 1 ? 1 : want) : 148 * 8);
      if (dtype == YG_BF16)
        bn_bwd_apply_cg_kernel<bf16><<<blocks_cg, 256, sm, st>>>((bf16*)g, (const bf16*)y, npix, C, M, sums, gamma, mean, invstd, batch_stats, dgamma, dbeta, clip);
      else
        bn_bwd_apply_cg_kernel<float><<<blocks_cg, 256, sm, st>>>((float*)g, (const float*)y, npix, C, M, sums, gamma, mean, invstd, batch_stats, dgamma, dbeta, clip);
      YG_LAUNCH_CHECK("bn_bwd_apply");
      return YG_OK;
    } else
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
