// Fused AdamW over one flat fp32 parameter buffer.
// Replaces torch.optim.AdamW.step's foreach kernels (/root/reference/yogo/train.py:213-217, 324)
// with a single HBM-bound pass: 4 reads + 3 writes of 4 bytes per parameter.
#include "common.cuh"
#include <math.h>

namespace yg {
__global__ void adamw_flat_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                  float* __restrict__ v, long long n, float lr, float b1, float b2, float eps,
                                  float wd, float bc1, float bc2_sqrt, float gscale) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i] * gscale;
  float pi = p[i] * (1.f - lr * wd);
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  pi -= (lr / bc1) * (mi / denom);
  p[i] = pi; m[i] = mi; v[i] = vi;
}
// same update with the step-dependent scalars read from device memory, so that the launch can be
// replayed from a CUDA graph: hyper = [lr, beta1, beta2, eps, weight_decay, 1-beta1^t, sqrt(1-beta2^t), grad_scale]
__global__ void adamw_flat_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                                      float* __restrict__ v, long long n, const float* __restrict__ hyper) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float lr = hyper[0], b1 = hyper[1], b2 = hyper[2], eps = hyper[3], wd = hyper[4], bc1 = hyper[5],
              bc2s = hyper[6], gscale = hyper[7];
  const float gi = g[i] * gscale;
  float pi = p[i] * (1.f - lr * wd);
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  pi -= (lr / bc1) * (mi / (sqrtf(vi) / bc2s + eps));
  p[i] = pi; m[i] = mi; v[i] = vi;
}
}  // namespace yg
using namespace yg;

extern "C" int yg_adamw_flat_dev(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                                 const float* hyper_dev, void* stream) {
  YG_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && hyper_dev && n >= 0, "adamw_flat_dev: bad arguments");
  if (n == 0) return YG_OK;
  adamw_flat_dev_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, hyper_dev);
  YG_LAUNCH_CHECK("adamw_flat_dev");
  return YG_OK;
}

extern "C" int yg_adamw_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                             float lr, float beta1, float beta2, float eps, float weight_decay,
                             long long step, float grad_scale, void* stream) {
  YG_CHECK_ARG(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "adamw_flat: bad arguments");
  if (n == 0) return YG_OK;
  const float bc1 = 1.f - (float)pow((double)beta1, (double)step);
  const float bc2s = (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adamw_flat_kernel<<<cdiv(n, 256), 256, 0, (cudaStream_t)stream>>>(param, grad, exp_avg, exp_avg_sq, n, lr, beta1,
                                                                     beta2, eps, weight_decay, bc1, bc2s, grad_scale);
  YG_LAUNCH_CHECK("adamw_flat");
  return YG_OK;
}
