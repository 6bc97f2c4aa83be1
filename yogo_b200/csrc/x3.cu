// fp32 convolutions on the bf16 tensor cores: the "x3" split.
//
// compute_dtype = float32 is the reference's default (no --half, /root/reference/yogo/train.py:315-318) and has to stay
// within 1e-3 of the fp32 result including the gradients, whose error amplification through this network is ~100x
// (SURVEY.md Appendix E): one tensor-float-32 pass (2^-11 per product) does not meet that.  Every fp32 operand is
// therefore split into two bf16 terms, x = hi + lo (hi = bf16(x), lo = bf16(x - hi), |x - hi - lo| <= 2^-17 |x|), and
//     x . w  ~=  hi_x . hi_w + lo_x . hi_w + hi_x . lo_w          (the lo . lo term is 2^-18 relative)
// runs as ONE bf16 convolution with three times the input channels - activations concatenated per pixel as
// [hi, lo, hi], weights as [hi, hi, lo] - on the unchanged tcgen05 kernels, accumulating in fp32 (TMEM) and writing fp32.
// Cost: 3x the tensor work of the bf16 path plus one streaming split pass per operand; error ~2^-16, far below tf32.
// The weight gradient needs x . g over pixels: [hi_x, lo_x, hi_x] . hi_g  (one launch, the third block is unused) plus
// hi_x . lo_g (a second launch on the plain hi tensor).
#include "common.cuh"

#include <mutex>

namespace yg {

bool tc_fwd_supported(int dtype, int W, int Cin, int Cout, int ks, int stride);
bool tc_dgrad_supported(int dtype, int W, int Cin, int Cout, int ks, int stride);
bool tc_wgrad_supported(int dtype, int W, int Cin, int Cout, int ks, int stride);
int conv_fwd_tc(const void*, const float*, void*, int, int, int, int, int, int, int, const FwdEpi&, cudaStream_t);
int conv_dgrad_tc(const void*, const float*, void*, int, int, int, int, int, int, int, const BwdEpi&, cudaStream_t);
int conv_wgrad_tc(const void*, const void*, float*, float*, int, int, int, int, int, int, int, float, void*, size_t, cudaStream_t);
size_t tc_wgrad_workspace(int, int, int, int, int, int, int);

// ---- grow-only device scratch, one buffer per purpose and device (single-stream contract of the library; sized during the
// warm-up calls, so nothing is allocated while a CUDA graph is being captured)
constexpr int X3_SLOTS = 8;
static void* g_x3_buf[16][X3_SLOTS];
static size_t g_x3_size[16][X3_SLOTS];
static std::mutex g_x3_mutex;

static int x3_scratch(int slot, size_t bytes, void** out) {
  std::lock_guard<std::mutex> lock(g_x3_mutex);
  int dev = 0;
  cudaGetDevice(&dev);
  dev &= 15;
  if (g_x3_size[dev][slot] < bytes) {
    if (g_x3_buf[dev][slot]) {
      YG_CUDA(cudaDeviceSynchronize());   // in-flight kernels may still read the old buffer
      cudaFree(g_x3_buf[dev][slot]);
      g_x3_buf[dev][slot] = nullptr;
      g_x3_size[dev][slot] = 0;
    }
    const size_t want = bytes + bytes / 8 + 256;
    YG_CUDA(cudaMalloc(&g_x3_buf[dev][slot], want));
    g_x3_size[dev][slot] = want;
  }
  *out = g_x3_buf[dev][slot];
  return YG_OK;
}

// x (npix, C) fp32 -> cat3 (npix, 3C) bf16 = [hi, lo, hi] per pixel; optionally plain hi / lo tensors (npix, C)
__global__ void x3_split_kernel(const float* __restrict__ x, bf16* __restrict__ cat3, bf16* __restrict__ hi_only,
                                bf16* __restrict__ lo_only, long long npix, int C) {
  const int c4 = C >> 2;                         // C % 4 == 0
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= npix * c4) return;
  const long long px = i / c4;
  const int c = (int)(i - px * c4) * 4;
  const float4 v = *reinterpret_cast<const float4*>(x + px * C + c);
  const float f[4] = {v.x, v.y, v.z, v.w};
  __align__(8) bf16 h[4], l[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    h[k] = __float2bfloat16_rn(f[k]);
    l[k] = __float2bfloat16_rn(f[k] - __bfloat162float(h[k]));
  }
  const uint2 hv = *reinterpret_cast<const uint2*>(h), lv = *reinterpret_cast<const uint2*>(l);
  if (cat3) {
    bf16* row = cat3 + px * 3 * C;
    *reinterpret_cast<uint2*>(row + c) = hv;
    *reinterpret_cast<uint2*>(row + C + c) = lv;
    *reinterpret_cast<uint2*>(row + 2 * C + c) = hv;
  }
  if (hi_only) *reinterpret_cast<uint2*>(hi_only + px * C + c) = hv;
  if (lo_only) *reinterpret_cast<uint2*>(lo_only + px * C + c) = lv;
}

// w (Cout, Cin, taps) fp32 -> along_cin: (Cout, 3 Cin, taps) = [hi, hi, lo] over the input channels (forward);
//                             else:      (3 Cout, Cin, taps) = [hi; hi; lo] over the output channels (dgrad, whose K is Cout)
__global__ void x3_wcat_kernel(const float* __restrict__ w, float* __restrict__ out, int Cout, int Cin, int taps, int along_cin) {
  const long long total = 3LL * Cout * Cin * taps;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int t = (int)(i % taps);
  long long r = i / taps;
  int co, ci, part;
  if (along_cin) { const int cc = (int)(r % (3 * Cin)); co = (int)(r / (3 * Cin)); part = cc / Cin; ci = cc - part * Cin; }
  else { ci = (int)(r % Cin); const int oo = (int)(r / Cin); part = oo / Cout; co = oo - part * Cout; }
  const float v = w[((long long)co * Cin + ci) * taps + t];
  const float hi = __bfloat162float(__float2bfloat16_rn(v));
  out[i] = part == 2 ? __bfloat162float(__float2bfloat16_rn(v - hi)) : hi;
}

// dw = clamp(A[:, 0:Cin] + A[:, Cin:2Cin] + B), A (Cout, 3 Cin, taps), B (Cout, Cin, taps); db = clamp(dba + dbb)
__global__ void x3_wgrad_combine_kernel(const float* __restrict__ A, const float* __restrict__ B, const float* __restrict__ dba,
                                        const float* __restrict__ dbb, float* __restrict__ dw, float* __restrict__ db, int Cout,
                                        int Cin, int taps, float clip) {
  const long long nw = (long long)Cout * Cin * taps;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nw) {
    const int t = (int)(i % taps);
    const long long r = i / taps;
    const int ci = (int)(r % Cin), co = (int)(r / Cin);
    const float* a = A + ((long long)co * 3 * Cin) * taps;
    dw[i] = clampf(a[(long long)ci * taps + t] + a[(long long)(Cin + ci) * taps + t] + B[i], clip);
  } else if (i < nw + Cout && db) {
    const int co = (int)(i - nw);
    db[co] = clampf(dba[co] + dbb[co], clip);
  }
}

bool x3_fwd_supported(int W, int Cin, int Cout, int ks, int stride) {
  return ks == 3 && Cin % 16 == 0 && Cin >= 16 && tc_fwd_supported(YG_BF16, W, 3 * Cin, Cout, ks, stride);
}
bool x3_dgrad_supported(int W, int Cin, int Cout, int ks, int stride) {
  return ks == 3 && Cout % 16 == 0 && Cin % 16 == 0 && tc_dgrad_supported(YG_BF16, W, Cin, 3 * Cout, ks, stride);
}
bool x3_wgrad_supported(int W, int Cin, int Cout, int ks, int stride) {
  return ks == 3 && Cin % 16 == 0 && Cin >= 16 && tc_wgrad_supported(YG_BF16, W, 3 * Cin, Cout, ks, stride) &&
         tc_wgrad_supported(YG_BF16, W, Cin, Cout, ks, stride);
}
size_t x3_wgrad_workspace(int N, int H, int W, int Cin, int Cout, int ks, int stride) {
  if (!x3_wgrad_supported(W, Cin, Cout, ks, stride)) return 0;
  const size_t a = tc_wgrad_workspace(N, H, W, 3 * Cin, Cout, ks, stride), b = tc_wgrad_workspace(N, H, W, Cin, Cout, ks, stride);
  return a > b ? a : b;
}

static int split(const float* x, long long npix, int C, bf16* cat3, bf16* hi, bf16* lo, cudaStream_t st) {
  const long long n = npix * (C / 4);
  x3_split_kernel<<<cdiv(n, 256), 256, 0, st>>>(x, cat3, hi, lo, npix, C);
  YG_LAUNCH_CHECK("x3_split");
  return YG_OK;
}

int conv_fwd_x3(const float* x, const float* w, float* y, int N, int H, int W, int Cin, int Cout, int ks, int stride, FwdEpi ep,
                cudaStream_t st) {
  void *xs = nullptr, *wc = nullptr;
  const long long npix = (long long)N * H * W;
  int rc = x3_scratch(0, (size_t)npix * 3 * Cin * sizeof(bf16), &xs);
  if (rc) return rc;
  rc = x3_scratch(1, (size_t)3 * Cout * Cin * ks * ks * sizeof(float), &wc);
  if (rc) return rc;
  rc = split(x, npix, Cin, (bf16*)xs, nullptr, nullptr, st);
  if (rc) return rc;
  x3_wcat_kernel<<<cdiv(3LL * Cout * Cin * ks * ks, 256), 256, 0, st>>>(w, (float*)wc, Cout, Cin, ks * ks, 1);
  YG_LAUNCH_CHECK("x3_wcat");
  ep.io_f32 = 1;
  return conv_fwd_tc(xs, (const float*)wc, y, N, H, W, 3 * Cin, Cout, ks, stride, ep, st);
}

int conv_dgrad_x3(const float* dz, const float* w, float* dx, int N, int H, int W, int Cin, int Cout, int ks, int stride, BwdEpi be,
                  cudaStream_t st) {
  const int Ho = (H + 2 * (ks / 2) - ks) / stride + 1, Wo = (W + 2 * (ks / 2) - ks) / stride + 1;
  void *gs = nullptr, *wc = nullptr;
  const long long npix = (long long)N * Ho * Wo;
  int rc = x3_scratch(2, (size_t)npix * 3 * Cout * sizeof(bf16), &gs);
  if (rc) return rc;
  rc = x3_scratch(1, (size_t)3 * Cout * Cin * ks * ks * sizeof(float), &wc);
  if (rc) return rc;
  rc = split(dz, npix, Cout, (bf16*)gs, nullptr, nullptr, st);
  if (rc) return rc;
  x3_wcat_kernel<<<cdiv(3LL * Cout * Cin * ks * ks, 256), 256, 0, st>>>(w, (float*)wc, Cout, Cin, ks * ks, 0);
  YG_LAUNCH_CHECK("x3_wcat");
  be.io_f32 = 1;
  return conv_dgrad_tc(gs, (const float*)wc, dx, N, H, W, Cin, 3 * Cout, ks, stride, be, st);
}

int conv_wgrad_x3(const float* x, const float* dz, float* dw, float* dbias, int N, int H, int W, int Cin, int Cout, int ks, int stride,
                  float clip, void* ws, size_t ws_bytes, cudaStream_t st) {
  const int Ho = (H + 2 * (ks / 2) - ks) / stride + 1, Wo = (W + 2 * (ks / 2) - ks) / stride + 1;
  const long long npx = (long long)N * H * W, npz = (long long)N * Ho * Wo;
  const int taps = ks * ks;
  void *xs3 = nullptr, *xh = nullptr, *gh = nullptr, *gl = nullptr, *acc = nullptr;
  int rc;
  if ((rc = x3_scratch(0, (size_t)npx * 3 * Cin * sizeof(bf16), &xs3))) return rc;
  if ((rc = x3_scratch(3, (size_t)npx * Cin * sizeof(bf16), &xh))) return rc;
  if ((rc = x3_scratch(4, (size_t)npz * Cout * sizeof(bf16), &gh))) return rc;
  if ((rc = x3_scratch(5, (size_t)npz * Cout * sizeof(bf16), &gl))) return rc;
  const size_t na = (size_t)Cout * 3 * Cin * taps, nb = (size_t)Cout * Cin * taps;
  if ((rc = x3_scratch(6, (na + nb + 2 * (size_t)Cout) * sizeof(float), &acc))) return rc;
  float* A = (float*)acc;
  float* B = A + na;
  float* dba = B + nb;
  float* dbb = dba + Cout;
  if ((rc = split(x, npx, Cin, (bf16*)xs3, (bf16*)xh, nullptr, st))) return rc;
  if ((rc = split(dz, npz, Cout, nullptr, (bf16*)gh, (bf16*)gl, st))) return rc;
  if ((rc = conv_wgrad_tc(xs3, gh, A, dba, N, H, W, 3 * Cin, Cout, ks, stride, 0.f, ws, ws_bytes, st))) return rc;
  if ((rc = conv_wgrad_tc(xh, gl, B, dbb, N, H, W, Cin, Cout, ks, stride, 0.f, ws, ws_bytes, st))) return rc;
  x3_wgrad_combine_kernel<<<cdiv((long long)nb + Cout, 256), 256, 0, st>>>(A, B, dba, dbb, dw, dbias, Cout, Cin, taps, clip);
  YG_LAUNCH_CHECK("x3_wgrad_combine");
  return YG_OK;
}

}  // namespace yg
