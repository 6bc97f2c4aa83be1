// Generic SIMT convolution kernels (any Cin/Cout, fp32 or bf16 NHWC storage, fp32 accumulate).
// These are the always-correct CUDA path: the fp32 variant is the tight-parity path against the
// oracle, and shapes the tcgen05 engine (conv_tc.cu) does not take (Cin/Cout not multiples of 16,
// the 1- or 3-channel first layer) run here.  Replaces cuDNN fprop/dgrad/wgrad as dispatched by
// nn.Conv2d in /root/reference/yogo/model_defns.py:34-67.
#include "common.cuh"

namespace yg {

// ------------------------------------------------------------------------------------------
// forward: tile 8x8 output pixels x 32 couts per block, 128 threads, 4 px x 4 co per thread
// ------------------------------------------------------------------------------------------
constexpr int F_TH = 8, F_TW = 8, F_CO = 32, F_KC = 8, F_THREADS = 128;

template <typename T, int KS>
__global__ void __launch_bounds__(F_THREADS) conv_fwd_simt_kernel(
    const T* __restrict__ x, const float* __restrict__ w, T* __restrict__ y,
    int N, int H, int W, int Cin, int Ho, int Wo, int Cout, int stride, FwdEpi ep) {
  constexpr int PAD = KS / 2;
  constexpr int PMAX = (F_TH - 1) * 2 + KS;  // patch extent for stride 2
  __shared__ float xs[PMAX][PMAX][F_KC + 1];
  __shared__ __align__(16) float ws[KS * KS][F_KC][F_CO];
  __shared__ float ssum[F_CO], ssq[F_CO];

  const int tilesW = (Wo + F_TW - 1) / F_TW;
  const int th0 = (blockIdx.x / tilesW) * F_TH, tw0 = (blockIdx.x % tilesW) * F_TW;
  const int co0 = blockIdx.y * F_CO;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int pg = tid / 8, cg = tid % 8;
  const int prow = pg / 2, pcol0 = (pg % 2) * 4;
  const int PH = (F_TH - 1) * stride + KS, PW = (F_TW - 1) * stride + KS;
  const int ih0 = th0 * stride - PAD, iw0 = tw0 * stride - PAD;

  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;

  for (int c0 = 0; c0 < Cin; c0 += F_KC) {
    __syncthreads();
    for (int i = tid; i < PH * PW * F_KC; i += F_THREADS) {
      int c = i % F_KC, pw = (i / F_KC) % PW, ph = i / (F_KC * PW);
      int ih = ih0 + ph, iw = iw0 + pw, ci = c0 + c;
      float v = 0.f;
      if (ih >= 0 && ih < H && iw >= 0 && iw < W && ci < Cin)
        v = to_f<T>(x[(((long long)n * H + ih) * W + iw) * Cin + ci]);
      xs[ph][pw][c] = v;
    }
    for (int i = tid; i < KS * KS * F_KC * F_CO; i += F_THREADS) {
      int co = i % F_CO, c = (i / F_CO) % F_KC, tap = i / (F_CO * F_KC);
      int ci = c0 + c, cog = co0 + co;
      float v = 0.f;
      if (ci < Cin && cog < Cout) v = w[((long long)cog * Cin + ci) * (KS * KS) + tap];
      ws[tap][c][co] = v;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < KS; ++r)
#pragma unroll
      for (int s = 0; s < KS; ++s)
#pragma unroll
        for (int c = 0; c < F_KC; ++c) {
          const float4 wv = *reinterpret_cast<const float4*>(&ws[r * KS + s][c][cg * 4]);
#pragma unroll
          for (int p = 0; p < 4; ++p) {
            float xv = xs[prow * stride + r][(pcol0 + p) * stride + s][c];
            acc[p][0] += xv * wv.x; acc[p][1] += xv * wv.y;
            acc[p][2] += xv * wv.z; acc[p][3] += xv * wv.w;
          }
        }
  }

  if (ep.stats) {
    if (tid < F_CO) { ssum[tid] = 0.f; ssq[tid] = 0.f; }
    __syncthreads();
  }
  const int ho = th0 + prow;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int co = co0 + cg * 4 + c;
    float lsum = 0.f, lsq = 0.f;
    if (co < Cout) {
      const float sc = ep.scale ? ep.scale[co] : 1.f, sh = ep.shift ? ep.shift[co] : 0.f;
      const float ds = ep.dropscale ? ep.dropscale[(long long)n * Cout + co] : 1.f;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int wo = tw0 + pcol0 + p;
        if (ho < Ho && wo < Wo) {
          float v = acc[p][c] * sc + sh;
          // statistics are taken on the value as stored (rounded to T) so that the later
          // normalisation of the stored tensor is self-consistent
          if (ep.stats || ep.preact) v = to_f<T>(from_f<T>(v));
          lsum += v; lsq += v * v;
          const long long o = (((long long)n * Ho + ho) * Wo + wo) * Cout + co;
          if (ep.preact) ((T*)ep.preact)[o] = from_f<T>(v);
          if (y) y[o] = from_f<T>(act_fwd(v, ep.act) * ds);
        }
      }
    }
    if (ep.stats) { atomicAdd(&ssum[cg * 4 + c], lsum); atomicAdd(&ssq[cg * 4 + c], lsq); }
  }
  if (ep.stats) {
    __syncthreads();
    if (tid < F_CO && co0 + tid < Cout) {
      atomicAdd(&ep.stats[co0 + tid], (double)ssum[tid]);
      atomicAdd(&ep.stats[Cout + co0 + tid], (double)ssq[tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// dgrad (gather form): tile 8x8 input pixels x 32 cin per block
// ------------------------------------------------------------------------------------------
template <typename T, int KS>
__global__ void __launch_bounds__(F_THREADS) conv_dgrad_simt_kernel(
    const T* __restrict__ dz, const float* __restrict__ w, T* __restrict__ dx,
    int N, int H, int W, int Cin, int Ho, int Wo, int Cout, int stride, BwdEpi be) {
  constexpr int PAD = KS / 2;
  constexpr int PMAX = F_TH + KS - 1;  // stride 1 is the largest patch
  __shared__ float gs[PMAX][PMAX][F_KC + 1];
  __shared__ __align__(16) float ws[KS * KS][F_KC][F_CO];  // [tap][co chunk][ci tile]
  __shared__ float ssum[F_CO], ssq[F_CO];

  const int tilesW = (W + F_TW - 1) / F_TW;
  const int th0 = (blockIdx.x / tilesW) * F_TH, tw0 = (blockIdx.x % tilesW) * F_TW;
  const int ci0 = blockIdx.y * F_CO;
  const int n = blockIdx.z;
  const int tid = threadIdx.x;
  const int pg = tid / 8, cg = tid % 8;
  const int prow = pg / 2, pcol0 = (pg % 2) * 4;
  // floor division for possibly negative numerators
  auto fdiv = [](int a, int b) { return (a >= 0) ? a / b : -((-a + b - 1) / b); };
  const int po_h = fdiv(th0 + PAD - (KS - 1), stride), po_w = fdiv(tw0 + PAD - (KS - 1), stride);
  const int PH = (th0 + F_TH - 1 + PAD) / stride - po_h + 1;
  const int PW = (tw0 + F_TW - 1 + PAD) / stride - po_w + 1;

  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int c = 0; c < 4; ++c) acc[p][c] = 0.f;

  for (int c0 = 0; c0 < Cout; c0 += F_KC) {
    __syncthreads();
    for (int i = tid; i < PH * PW * F_KC; i += F_THREADS) {
      int c = i % F_KC, pw = (i / F_KC) % PW, ph = i / (F_KC * PW);
      int oh = po_h + ph, ow = po_w + pw, co = c0 + c;
      float v = 0.f;
      if (oh >= 0 && oh < Ho && ow >= 0 && ow < Wo && co < Cout)
        v = to_f<T>(dz[(((long long)n * Ho + oh) * Wo + ow) * Cout + co]);
      gs[ph][pw][c] = v;
    }
    for (int i = tid; i < KS * KS * F_KC * F_CO; i += F_THREADS) {
      int ci = i % F_CO, c = (i / F_CO) % F_KC, tap = i / (F_CO * F_KC);
      int co = c0 + c, cig = ci0 + ci;
      float v = 0.f;
      if (co < Cout && cig < Cin) v = w[((long long)co * Cin + cig) * (KS * KS) + tap];
      ws[tap][c][ci] = v;
    }
    __syncthreads();
    for (int r = 0; r < KS; ++r)
      for (int s = 0; s < KS; ++s) {
        const int thh = th0 + prow + PAD - r;
        if (thh < 0 || (thh % stride) != 0) continue;
        const int ph = thh / stride - po_h;
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int tww = tw0 + pcol0 + p + PAD - s;
          if (tww < 0 || (tww % stride) != 0) continue;
          const int pw = tww / stride - po_w;
#pragma unroll
          for (int c = 0; c < F_KC; ++c) {
            const float4 wv = *reinterpret_cast<const float4*>(&ws[r * KS + s][c][cg * 4]);
            const float gv = gs[ph][pw][c];
            acc[p][0] += gv * wv.x; acc[p][1] += gv * wv.y;
            acc[p][2] += gv * wv.z; acc[p][3] += gv * wv.w;
          }
        }
      }
  }

  if (be.bn_sums) {
    if (tid < F_CO) { ssum[tid] = 0.f; ssq[tid] = 0.f; }
    __syncthreads();
  }
  const int h = th0 + prow;
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const int ci = ci0 + cg * 4 + c;
    float lsum = 0.f, lsx = 0.f;
    if (ci < Cin) {
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const int wq = tw0 + pcol0 + p;
        if (h < H && wq < W) {
          const long long o = (((long long)n * H + h) * W + wq) * Cin + ci;
          float xhat;
          float g = bwd_epi_apply<T>(be, acc[p][c], o, n, ci, Cin, xhat);
          if (be.bn_sums) g = to_f<T>(from_f<T>(g));
          lsum += g; lsx += g * xhat;
          dx[o] = from_f<T>(g);
        }
      }
    }
    if (be.bn_sums) { atomicAdd(&ssum[cg * 4 + c], lsum); atomicAdd(&ssq[cg * 4 + c], lsx); }
  }
  if (be.bn_sums) {
    __syncthreads();
    if (tid < F_CO && ci0 + tid < Cin) {
      atomicAdd(&be.bn_sums[ci0 + tid], (double)ssum[tid]);
      atomicAdd(&be.bn_sums[Cin + ci0 + tid], (double)ssq[tid]);
    }
  }
}

// ------------------------------------------------------------------------------------------
// wgrad: block = 32 co x 16 ci x all taps over a slice of output rows; split-K partials
// ------------------------------------------------------------------------------------------
constexpr int WG_CO = 32, WG_CI = 16, WG_PX = 32, WG_THREADS = 256;

template <typename T, int KS>
__global__ void __launch_bounds__(WG_THREADS) conv_wgrad_simt_kernel(
    const T* __restrict__ x, const T* __restrict__ dz, float* __restrict__ partial,
    int N, int H, int W, int Cin, int Ho, int Wo, int Cout, int stride, int rows_per_slice) {
  constexpr int PAD = KS / 2;
  constexpr int XW = (WG_PX - 1) * 2 + KS;
  __shared__ float dzs[WG_PX][WG_CO + 1];
  __shared__ float xs[KS][XW][WG_CI + 1];

  const int co0 = blockIdx.x * WG_CO, ci0 = blockIdx.y * WG_CI;
  const int slice = blockIdx.z;
  const int tid = threadIdx.x;
  const int co = tid / 8, cig = tid % 8;  // thread: 1 co x 2 ci x taps
  float acc[KS * KS][2];
#pragma unroll
  for (int t = 0; t < KS * KS; ++t) acc[t][0] = acc[t][1] = 0.f;
  float accb = 0.f;

  const long long row_begin = (long long)slice * rows_per_slice;
  long long row_end = row_begin + rows_per_slice;
  if (row_end > (long long)N * Ho) row_end = (long long)N * Ho;
  const int XWs = (WG_PX - 1) * stride + KS;

  for (long long row = row_begin; row < row_end; ++row) {
    const int n = (int)(row / Ho), ho = (int)(row % Ho);
    for (int wo0 = 0; wo0 < Wo; wo0 += WG_PX) {
      __syncthreads();
      for (int i = tid; i < WG_PX * WG_CO; i += WG_THREADS) {
        int c = i % WG_CO, p = i / WG_CO;
        int wo = wo0 + p, cog = co0 + c;
        float v = 0.f;
        if (wo < Wo && cog < Cout) v = to_f<T>(dz[(((long long)n * Ho + ho) * Wo + wo) * Cout + cog]);
        dzs[p][c] = v;
      }
      for (int i = tid; i < KS * XWs * WG_CI; i += WG_THREADS) {
        int c = i % WG_CI, pw = (i / WG_CI) % XWs, r = i / (WG_CI * XWs);
        int ih = ho * stride - PAD + r, iw = wo0 * stride - PAD + pw, ci = ci0 + c;
        float v = 0.f;
        if (ih >= 0 && ih < H && iw >= 0 && iw < W && ci < Cin)
          v = to_f<T>(x[(((long long)n * H + ih) * W + iw) * Cin + ci]);
        xs[r][pw][c] = v;
      }
      __syncthreads();
#pragma unroll 4
      for (int p = 0; p < WG_PX; ++p) {
        const float d = dzs[p][co];
        accb += d;
#pragma unroll
        for (int r = 0; r < KS; ++r)
#pragma unroll
          for (int s = 0; s < KS; ++s) {
            acc[r * KS + s][0] += d * xs[r][p * stride + s][cig * 2];
            acc[r * KS + s][1] += d * xs[r][p * stride + s][cig * 2 + 1];
          }
      }
    }
  }
  // partial layout: [slice][Cout][Cin][taps] followed by [slice][Cout] bias partials
  const int cog = co0 + co;
  if (cog < Cout) {
    float* base = partial + (long long)slice * ((long long)Cout * Cin * KS * KS + Cout);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const int ci = ci0 + cig * 2 + j;
      if (ci < Cin)
#pragma unroll
        for (int t = 0; t < KS * KS; ++t) base[((long long)cog * Cin + ci) * (KS * KS) + t] = acc[t][j];
    }
    if (blockIdx.y == 0 && cig == 0) base[(long long)Cout * Cin * KS * KS + cog] = accb;
  }
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw,
                                    float* __restrict__ dbias, long long nw, int Cout, int slices,
                                    float clip) {
  const long long stride_slice = nw + Cout;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nw + Cout) return;
  float s = 0.f;
  for (int k = 0; k < slices; ++k) s += partial[k * stride_slice + i];
  s = clampf(s, clip);
  if (i < nw) dw[i] = s;
  else if (dbias) dbias[i - nw] = s;
}

static int wgrad_slices(int N, int Ho, int Cin, int Cout) {
  long long tiles = (long long)cdiv(Cout, WG_CO) * cdiv(Cin, WG_CI);
  long long want = (148LL * 4 + tiles - 1) / tiles;
  long long rows = (long long)N * Ho;
  if (want > rows) want = rows;
  if (want < 1) want = 1;
  if (want > 1024) want = 1024;
  return (int)want;
}

size_t simt_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int ks, int stride) {
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1;
  int slices = wgrad_slices(N, Ho, Cin, Cout);
  return (size_t)slices * ((size_t)Cout * Cin * ks * ks + Cout) * sizeof(float);
}

template <typename T>
static int conv_fwd_simt_t(const void* x, const float* w, void* y, int N, int H, int W, int Cin, int Cout,
                           int ks, int stride, const FwdEpi& ep, cudaStream_t st) {
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  dim3 grid(cdiv(Ho, F_TH) * cdiv(Wo, F_TW), cdiv(Cout, F_CO), N);
  if (ks == 3)
    conv_fwd_simt_kernel<T, 3><<<grid, F_THREADS, 0, st>>>((const T*)x, w, (T*)y, N, H, W, Cin, Ho, Wo, Cout, stride, ep);
  else
    conv_fwd_simt_kernel<T, 1><<<grid, F_THREADS, 0, st>>>((const T*)x, w, (T*)y, N, H, W, Cin, Ho, Wo, Cout, stride, ep);
  YG_LAUNCH_CHECK("conv_fwd_simt");
  return YG_OK;
}

int conv_fwd_simt(const void* x, const float* w, void* y, int dtype, int N, int H, int W, int Cin, int Cout,
                  int ks, int stride, const FwdEpi& ep, cudaStream_t st) {
  return dtype == YG_BF16 ? conv_fwd_simt_t<bf16>(x, w, y, N, H, W, Cin, Cout, ks, stride, ep, st)
                          : conv_fwd_simt_t<float>(x, w, y, N, H, W, Cin, Cout, ks, stride, ep, st);
}

template <typename T>
static int conv_dgrad_simt_t(const void* dz, const float* w, void* dx, int N, int H, int W, int Cin, int Cout,
                             int ks, int stride, const BwdEpi& be, cudaStream_t st) {
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  dim3 grid(cdiv(H, F_TH) * cdiv(W, F_TW), cdiv(Cin, F_CO), N);
  if (ks == 3)
    conv_dgrad_simt_kernel<T, 3><<<grid, F_THREADS, 0, st>>>((const T*)dz, w, (T*)dx, N, H, W, Cin, Ho, Wo, Cout, stride, be);
  else
    conv_dgrad_simt_kernel<T, 1><<<grid, F_THREADS, 0, st>>>((const T*)dz, w, (T*)dx, N, H, W, Cin, Ho, Wo, Cout, stride, be);
  YG_LAUNCH_CHECK("conv_dgrad_simt");
  return YG_OK;
}

int conv_dgrad_simt(const void* dz, const float* w, void* dx, int dtype, int N, int H, int W, int Cin, int Cout,
                    int ks, int stride, const BwdEpi& be, cudaStream_t st) {
  return dtype == YG_BF16 ? conv_dgrad_simt_t<bf16>(dz, w, dx, N, H, W, Cin, Cout, ks, stride, be, st)
                          : conv_dgrad_simt_t<float>(dz, w, dx, N, H, W, Cin, Cout, ks, stride, be, st);
}

template <typename T>
static int conv_wgrad_simt_t(const void* x, const void* dz, float* dw, float* dbias, int N, int H, int W,
                             int Cin, int Cout, int ks, int stride, float clip, void* ws, size_t ws_bytes,
                             cudaStream_t st) {
  const int pad = ks / 2;
  const int Ho = (H + 2 * pad - ks) / stride + 1, Wo = (W + 2 * pad - ks) / stride + 1;
  const int slices = wgrad_slices(N, Ho, Cin, Cout);
  const size_t need = simt_wgrad_workspace(N, H, W, Cin, Cout, ks, stride);
  if (ws_bytes < need || !ws) {
    set_error("conv_wgrad: workspace %zu < %zu", ws_bytes, need);
    return YG_ERR_WORKSPACE;
  }
  const int rows_per_slice = cdiv((long long)N * Ho, slices);
  dim3 grid(cdiv(Cout, WG_CO), cdiv(Cin, WG_CI), slices);
  // partial tiles not covered by any block (ragged ci/co tiles are covered; all entries written)
  if (ks == 3)
    conv_wgrad_simt_kernel<T, 3><<<grid, WG_THREADS, 0, st>>>((const T*)x, (const T*)dz, (float*)ws, N, H, W, Cin, Ho, Wo, Cout, stride, rows_per_slice);
  else
    conv_wgrad_simt_kernel<T, 1><<<grid, WG_THREADS, 0, st>>>((const T*)x, (const T*)dz, (float*)ws, N, H, W, Cin, Ho, Wo, Cout, stride, rows_per_slice);
  YG_LAUNCH_CHECK("conv_wgrad_simt");
  const long long nw = (long long)Cout * Cin * ks * ks;
  wgrad_reduce_kernel<<<cdiv(nw + Cout, 256), 256, 0, st>>>((const float*)ws, dw, dbias, nw, Cout, slices, clip);
  YG_LAUNCH_CHECK("wgrad_reduce");
  return YG_OK;
}

int conv_wgrad_simt(const void* x, const void* dz, float* dw, float* dbias, int dtype, int N, int H, int W,
                    int Cin, int Cout, int ks, int stride, float clip, void* ws, size_t ws_bytes, cudaStream_t st) {
  return dtype == YG_BF16
             ? conv_wgrad_simt_t<bf16>(x, dz, dw, dbias, N, H, W, Cin, Cout, ks, stride, clip, ws, ws_bytes, st)
             : conv_wgrad_simt_t<float>(x, dz, dw, dbias, N, H, W, Cin, Cout, ks, stride, clip, ws, ws_bytes, st);
}

}  // namespace yg
