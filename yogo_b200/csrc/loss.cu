// YOGOLoss forward + backward in one pass over the Sy x Sx prediction grid.
// Replaces ~60 ATen launches, two boolean-mask gathers (host syncs) and torchvision's
// box_convert / complete_box_iou_loss in /root/reference/yogo/yogo_loss.py:38-129
// (torchvision/ops/ciou_loss.py, diou_loss.py, _utils.py::_loss_inter_union).
// Reads (5+C)+6 floats per cell, writes 5+C floats of dpred.  Persistent blocks walk chunks of 256 cells (one thread per
// cell); labelled cells queue up per warp and their CIoU runs on full warps; warp-shuffle + per-block fp64 partials, the last
// block to finish sums them in block order (deterministic) and writes the four outputs - one launch.  ncu: ~600 instructions
// per 32 cells (softmax cross-entropy over the classes, 30 strided plane accesses), i.e. issue-bound at 2-3 TB/s, not HBM-bound.
#include "common.cuh"

namespace yg {

constexpr int LS_THREADS = 256;
constexpr int LS_MAXC = 27;  // 5 + C <= 32

__device__ __forceinline__ float sel_gt(float a, float b) { return a > b ? 1.f : (a == b ? 0.5f : 0.f); }
__device__ __forceinline__ float sel_lt(float a, float b) { return a < b ? 1.f : (a == b ? 0.5f : 0.f); }

// CIoU of one labelled cell (yogo_loss.py:59-105; torchvision complete_box_iou_loss with alpha as a constant): returns the
// loss term and writes the four box gradients
__device__ __forceinline__ float box_term(const float* __restrict__ pred, const float* __restrict__ label, float* __restrict__ dpred,
                                          long long cell, int N, int C, int SS, float iou_w) {
  const int n = (int)(cell / SS), k = (int)(cell % SS);
  const int D = 5 + C;
  const float* p = pred + (long long)n * D * SS + k;
  const float* l = label + (long long)n * 6 * SS + k;
  float* g = dpred ? dpred + (long long)n * D * SS + k : nullptr;
  const float invN = 1.f / (float)N;
  const float m = l[0];
  float l_iou = 0.f;
    float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
    if (m != 0.f) {
      const float p0 = p[0], p1 = p[(long long)SS], p2 = p[2LL * SS], p3 = p[3LL * SS];
      const float X1 = l[(long long)SS], Y1 = l[2LL * SS], X2 = l[3LL * SS], Y2 = l[4LL * SS];
      const float bx1 = p0 - 0.5f * p2, by1 = p1 - 0.5f * p3, bx2 = p0 + 0.5f * p2, by2 = p1 + 0.5f * p3;
      if (bx1 != bx2 && by1 != by2) {  // tested before clamping (yogo_loss.py:84-90)
        const float x1 = fminf(fmaxf(bx1, 0.f), 1.f), y1 = fminf(fmaxf(by1, 0.f), 1.f);
        const float x2 = fminf(fmaxf(bx2, 0.f), 1.f), y2 = fminf(fmaxf(by2, 0.f), 1.f);
        const float xk1 = fmaxf(x1, X1), yk1 = fmaxf(y1, Y1), xk2 = fminf(x2, X2), yk2 = fminf(y2, Y2);
        const bool im = (yk2 > yk1) && (xk2 > xk1);
        const float wI = xk2 - xk1, hI = yk2 - yk1;
        const float inter = im ? wI * hI : 0.f;
        const float wp = x2 - x1, hp = y2 - y1, wg = X2 - X1, hg = Y2 - Y1;
        const float uni = wp * hp + wg * hg - inter;
        const float eps = 1e-7f;
        const float Ue = uni + eps;
        const float iou = inter / Ue;
        const float xc1 = fminf(x1, X1), yc1 = fminf(y1, Y1), xc2 = fmaxf(x2, X2), yc2 = fmaxf(y2, Y2);
        const float cw = xc2 - xc1, ch = yc2 - yc1;
        const float diag = cw * cw + ch * ch + eps;
        const float xp = (x2 + x1) * 0.5f, yp = (y2 + y1) * 0.5f;
        const float xg = (X1 + X2) * 0.5f, yg_ = (Y1 + Y2) * 0.5f;
        const float dxc = xp - xg, dyc = yp - yg_;
        const float cent = dxc * dxc + dyc * dyc;
        const float kk = 0.40528473456935109f;  // 4 / pi^2
        const float r = wp / hp;
        const float dat = atanf(wg / hg) - atanf(r);
        const float v = kk * dat * dat;
        const float alpha = v / (1.f - iou + v + eps);  // no_grad
        l_iou = 1.f - iou + cent / diag + alpha * v;
        if (g) {
          const float imf = im ? 1.f : 0.f;
          float dI[4] = {-hI * sel_gt(x1, X1) * imf, -wI * sel_gt(y1, Y1) * imf,
                         hI * sel_lt(x2, X2) * imf, wI * sel_lt(y2, Y2) * imf};
          const float dA[4] = {-hp, -wp, hp, wp};
          const float dD[4] = {-2.f * cw * sel_lt(x1, X1), -2.f * ch * sel_lt(y1, Y1),
                               2.f * cw * sel_gt(x2, X2), 2.f * ch * sel_gt(y2, Y2)};
          const float dC[4] = {dxc, dyc, dxc, dyc};
          const float q = kk * 2.f * (-dat) * (1.f / (1.f + r * r));
          const float dv_dw = q / hp, dv_dh = q * (-wp / (hp * hp));
          const float dv[4] = {-dv_dw, -dv_dh, dv_dw, dv_dh};
          const float raw[4] = {bx1, by1, bx2, by2};
          float gg[4];
          const float scale = iou_w * invN;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float dU = dA[i] - dI[i];
            const float diou = (dI[i] * Ue - inter * dU) / (Ue * Ue);
            float t = -diou + (dC[i] * diag - cent * dD[i]) / (diag * diag) + alpha * dv[i];
            const bool pass = raw[i] >= 0.f && raw[i] <= 1.f;  // clamp backward
            gg[i] = pass ? t * scale : 0.f;
          }
          d0 = gg[0] + gg[2];
          d1 = gg[1] + gg[3];
          d2 = 0.5f * (gg[2] - gg[0]);
          d3 = 0.5f * (gg[3] - gg[1]);
        }
      }
    }
    if (g) { g[0] = d0; g[(long long)SS] = d1; g[2LL * SS] = d2; g[3LL * SS] = d3; }
  return l_iou;
}

// MAXC = compile-time bound of the class loops (4, 8, 16 or 27): the loops are fully unrolled and predicated on c < C, so a
// bound of 27 for the usual 7 classes would spend most of the issue slots on predicated-off iterations
template <int MAXC>
__global__ void __launch_bounds__(LS_THREADS) yogo_loss_kernel(
    const float* __restrict__ pred, const float* __restrict__ label, float* __restrict__ dpred,
    double* __restrict__ partial, int N, int C, int SS, float no_obj_w, float iou_w, float cls_w,
    float smoothing, float* __restrict__ out4, unsigned* __restrict__ ticket) {
  const long long total = (long long)N * SS;
  const int nchunks = (int)((total + LS_THREADS - 1) / LS_THREADS);
  float l_iou = 0.f, l_obj = 0.f, l_cls = 0.f;
  // labelled cells (<= 2.4 % of the grid) queue up per warp and their CIoU runs 32 at a time on full warps - evaluated in
  // place it made 54 % of the warps walk the whole branch for one or two lanes
  __shared__ int s_q[LS_THREADS / 32][64];
  const int wq = threadIdx.x >> 5, lq = threadIdx.x & 31;
  int qn = 0;   // warp-uniform queue length
  for (int chunk = blockIdx.x; chunk < nchunks; chunk += gridDim.x) {
  const int cell = chunk * LS_THREADS + (int)threadIdx.x;   // (N * SS < 2^31, checked by the host)
  bool labelled = false;
  if (cell < (int)total) {
    const int n = cell / SS, k = cell - n * SS;
    const int D = 5 + C;
    const float* p = pred + (long long)n * D * SS + k;
    const float* l = label + (long long)n * 6 * SS + k;
    float* g = dpred ? dpred + (long long)n * D * SS + k : nullptr;
    const float invN = 1.f / (float)N;
    const float m = l[0];
    // ---- objectness: (p4 - m)^2 * (m(1-lambda) + lambda)   (yogo_loss.py:116-119)
    {
      const float w = m * (1.f - no_obj_w) + no_obj_w;
      const float d = p[4LL * SS] - m;
      l_obj += d * d * w;
      if (g) g[4LL * SS] = 2.f * d * w * invN;
    }
    // ---- classification: label-smoothed CE on raw logits, masked (yogo_loss.py:107-114)
    {
      float lg[MAXC];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) { lg[c] = p[(long long)(5 + c) * SS]; mx = fmaxf(mx, lg[c]); }
      float se = 0.f, sl = 0.f;
      float ex[MAXC];   // exp(logit - max), reused for the softmax of the gradient
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c < C) { ex[c] = __expf(lg[c] - mx); se += ex[c]; sl += lg[c]; }   // (arguments <= 0: 2 ulp)
      const float lse = mx + __logf(se);                                        // (se in [1, C]: abs error < 1e-6)
      const float inv_se = 1.f / se;
      int tgt = (int)l[5LL * SS];  // .long() truncation
      tgt = tgt < 0 ? 0 : (tgt >= C ? C - 1 : tgt);
      float lt = 0.f;
#pragma unroll
      for (int c = 0; c < MAXC; ++c)
        if (c == tgt) lt = lg[c];
      const float nll = lse - lt;
      const float invC = 1.f / (float)C;
      const float smooth = lse - sl * invC;
      l_cls += m * ((1.f - smoothing) * nll + smoothing * smooth);
      if (g) {
        const float f = m * cls_w * invN, tdo = smoothing * invC, tdt = (1.f - smoothing) + tdo;
#pragma unroll
        for (int c = 0; c < MAXC; ++c)
          if (c < C) {
            const float sm = ex[c] * inv_se;
            g[(long long)(5 + c) * SS] = f * (sm - (c == tgt ? tdt : tdo));
          }
      }
    }
    labelled = m != 0.f;
    if (g && !labelled) { g[0] = 0.f; g[(long long)SS] = 0.f; g[2LL * SS] = 0.f; g[3LL * SS] = 0.f; }   // (box_term writes the others)
  }
  {
    const unsigned bal = __ballot_sync(0xffffffffu, labelled);
    if (labelled) s_q[wq][qn + __popc(bal & ((1u << lq) - 1u))] = cell;
    qn += __popc(bal);
    __syncwarp();
    if (qn >= 32) {
      l_iou += box_term(pred, label, dpred, (long long)s_q[wq][lq], N, C, SS, iou_w);
      const int rest = qn - 32;
      const int mv = lq < rest ? s_q[wq][32 + lq] : 0;
      __syncwarp();
      if (lq < rest) s_q[wq][lq] = mv;
      qn = rest;
      __syncwarp();
    }
  }
  }   // chunk loop
  if (lq < qn) l_iou += box_term(pred, label, dpred, (long long)s_q[wq][lq], N, C, SS, iou_w);
  // block reduction (double partials, deterministic order)
  __shared__ double red[3][LS_THREADS / 32];
  double a = warp_sum_d((double)l_iou), b = warp_sum_d((double)l_obj), c = warp_sum_d((double)l_cls);
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) { red[0][wid] = a; red[1][wid] = b; red[2][wid] = c; }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0;
    for (int i = 0; i < LS_THREADS / 32; ++i) s += red[threadIdx.x][i];
    partial[(long long)blockIdx.x * 3 + threadIdx.x] = s;
  }
  // the last block to finish sums the per-block partials in block order (deterministic) and writes the four outputs:
  // no second launch
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(ticket, 1u);
    s_last = done == gridDim.x - 1;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    __shared__ double fin[3][LS_THREADS / 32];
    double t3[3] = {0, 0, 0};
    for (int i = threadIdx.x; i < (int)gridDim.x; i += LS_THREADS) {
      t3[0] += __ldcg(partial + (long long)i * 3);
      t3[1] += __ldcg(partial + (long long)i * 3 + 1);
      t3[2] += __ldcg(partial + (long long)i * 3 + 2);
    }
    for (int q = 0; q < 3; ++q) t3[q] = warp_sum_d(t3[q]);
    if (lane == 0) { fin[0][wid] = t3[0]; fin[1][wid] = t3[1]; fin[2][wid] = t3[2]; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double t[3] = {0, 0, 0};
      for (int w = 0; w < LS_THREADS / 32; ++w) { t[0] += fin[0][w]; t[1] += fin[1][w]; t[2] += fin[2][w]; }
      const float iou = (float)(iou_w * t[0] / N);
      const float obj = (float)(t[1] / N);
      const float cls = (float)(cls_w * t[2] / N);
      out4[0] = obj + iou + cls;  // yogo_loss.py:121
      out4[1] = iou;
      out4[2] = obj;
      out4[3] = cls;
    }
  }
}

}  // namespace yg
using namespace yg;

extern "C" size_t yg_yogo_loss_workspace(int N, int Sy, int Sx) {
  return (size_t)cdiv((long long)N * Sy * Sx, LS_THREADS) * 3 * sizeof(double) + 64;   // partials + the completion ticket
}

extern "C" int yg_yogo_loss_fwd_bwd(const float* pred, const float* label, float* out4, float* dpred,
                                    int N, int num_classes, int Sy, int Sx,
                                    float no_obj_weight, float iou_weight, float classify_weight,
                                    float label_smoothing, void* workspace, size_t workspace_bytes, void* stream) {
  YG_CHECK_ARG(pred && label && out4, "yogo_loss: null pointer");
  YG_CHECK_ARG(num_classes >= 1 && num_classes <= LS_MAXC, "yogo_loss: num_classes %d not in [1,%d]", num_classes, LS_MAXC);
  YG_CHECK_ARG(N >= 1 && Sy >= 1 && Sx >= 1, "yogo_loss: empty batch or grid");
  YG_CHECK_ARG((long long)N * Sy * Sx < (1LL << 31) - 1024, "yogo_loss: N * Sy * Sx = %lld cells exceed 2^31", (long long)N * Sy * Sx);
  const size_t need = yg_yogo_loss_workspace(N, Sy, Sx);
  if (!workspace || workspace_bytes < need) {
    set_error("yogo_loss: workspace %zu < %zu", workspace_bytes, need);
    return YG_ERR_WORKSPACE;
  }
  const int SS = Sy * Sx;
  const int nchunks_h = cdiv((long long)N * SS, LS_THREADS);
  const int blocks = nchunks_h < 148 * 8 ? nchunks_h : 148 * 8;   // persistent: a block walks chunks of 256 cells
  cudaStream_t st = (cudaStream_t)stream;
  unsigned* ticket = reinterpret_cast<unsigned*>(reinterpret_cast<unsigned char*>(workspace) + (size_t)blocks * 3 * sizeof(double));
  YG_CUDA(cudaMemsetAsync(ticket, 0, sizeof(unsigned), st));   // (the kernel leaves it at zero; the workspace may be fresh memory)
#define YG_LOSS_LAUNCH(MC) yogo_loss_kernel<MC><<<blocks, LS_THREADS, 0, st>>>(pred, label, dpred, (double*)workspace, N, num_classes, SS, no_obj_weight, iou_weight, classify_weight, label_smoothing, out4, ticket)
  if (num_classes <= 4) YG_LOSS_LAUNCH(4);
  else if (num_classes <= 8) YG_LOSS_LAUNCH(8);
  else if (num_classes <= 16) YG_LOSS_LAUNCH(16);
  else YG_LOSS_LAUNCH(LS_MAXC);
#undef YG_LOSS_LAUNCH
  YG_LAUNCH_CHECK("yogo_loss");
  return YG_OK;
}
