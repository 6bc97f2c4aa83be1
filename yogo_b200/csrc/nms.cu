// Objectness threshold + box conversion + greedy NMS + class-confidence filter + per-class
// counts for a whole batch in ONE launch (one 1024-thread CTA per image).
// Replaces the per-image Python loop over format_preds
// (/root/reference/yogo/utils/prediction_formatting.py:23-93: boolean-mask gather, box_convert,
// torchvision.ops.nms, fancy-index) and get_prediction_class_counts /
// count_cells_for_formatted_preds (/root/reference/yogo/infer.py:60-124).
//
// Bit-exactness contract: every floating-point operation that decides
// the result is a single IEEE fp32 operation issued through __f*_rn intrinsics, so nvcc can
// never contract them into FMAs; the score order is a stable descending sort (ties -> lower
// candidate index first); suppression is `iou > thr` (strict; NaN never suppresses) with the
// fp32 ratio compared against the threshold as a double, exactly as torchvision's CPU kernel.
#include "common.cuh"

namespace yg {

constexpr int NMS_THREADS = 1024;
constexpr int NMS_MAX_CELLS = 16384;  // Sy*Sx limit (reference grid: 97*129 = 12513)
constexpr int NMS_CHUNK = 64;

struct NmsWs {
  int* cand_cell;     // [B][cells]  grid cell of candidate (original order)
  float4* cand_box;   // [B][cells]  xyxy of candidate
  int* order;         // [B][cells]  candidate ordinals in descending-score order
  float4* kept_box;   // [B][cells]
  int* kept_ord;      // [B][cells]  candidate ordinal of each kept box, NMS order
};

// exclusive scan of one flag per thread over the block; returns rank, total via reference
__device__ __forceinline__ int block_rank(bool flag, int* warp_tot /*[32]*/, int& total) {
  const unsigned b = __ballot_sync(0xffffffffu, flag);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int r = __popc(b & ((1u << lane) - 1u));
  __syncthreads();  // protect warp_tot reuse
  if (lane == 0) warp_tot[wid] = __popc(b);
  __syncthreads();
  int base = 0, tot = 0;
  for (int w = 0; w < NMS_THREADS / 32; ++w) {
    const int t = warp_tot[w];
    if (w < wid) base += t;
    tot += t;
  }
  total = tot;
  return base + r;
}

__device__ __forceinline__ bool iou_gt(const float4 a, const float area_a, const float4 b, const float area_b,
                                       const double thr) {
  // torchvision/csrc/ops/cpu/nms_kernel.cpp semantics
  const float xx1 = fmaxf(a.x, b.x), yy1 = fmaxf(a.y, b.y);
  const float xx2 = fminf(a.z, b.z), yy2 = fminf(a.w, b.w);
  const float w = __fsub_rn(xx2, xx1);
  const float h = __fsub_rn(yy2, yy1);
  // disjoint boxes (or NaN coordinates): the clamped intersection is 0 and 0 / x > thr is false for the thr > 0 NMS runs
  // with (0 / 0 = NaN never suppresses either) - most pairs of a sparse image end here, before the IEEE division
  if (!(w > 0.f && h > 0.f)) return false;
  const float inter = __fmul_rn(w, h);
  const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
  return (double)ovr > thr;
}
__device__ __forceinline__ float box_area(const float4 b) {
  return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
}

__global__ void __launch_bounds__(NMS_THREADS, 1) format_preds_kernel(
    const float* __restrict__ preds, int C, int cells, float obj_thresh, double iou_thresh, int do_nms,
    int xyxy, float min_cls, int* __restrict__ keep_count, float* __restrict__ rows,
    int* __restrict__ keep_index, unsigned long long* __restrict__ class_counts, NmsWs ws, int P /*pow2 >= cells*/) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned long long* keys = reinterpret_cast<unsigned long long*>(smem_raw);  // [P]
  __shared__ int warp_tot[NMS_THREADS / 32];
  __shared__ int s_kept;
  __shared__ unsigned long long s_sup, s_keepbits;
  __shared__ unsigned long long s_mask[NMS_CHUNK];
  __shared__ float4 s_cb[NMS_CHUNK];
  __shared__ float s_ca[NMS_CHUNK];
  __shared__ int s_hist[32];

  const int b = blockIdx.x, tid = threadIdx.x;
  const int D = 5 + C;
  const float* pb = preds + (long long)b * D * cells;
  int* cand_cell = ws.cand_cell + (long long)b * cells;
  float4* cand_box = ws.cand_box + (long long)b * cells;
  int* order = ws.order + (long long)b * cells;
  float4* kept_box = ws.kept_box + (long long)b * cells;
  int* kept_ord = ws.kept_ord + (long long)b * cells;

  if (tid < 32) s_hist[tid] = 0;

  // ---- phase 1: threshold + ordered compaction + boxes + sort keys ----
  int n = 0;
  for (int k0 = 0; k0 < cells; k0 += NMS_THREADS) {
    const int k = k0 + tid;
    float obj = 0.f;
    bool flag = false;
    if (k < cells) { obj = pb[4LL * cells + k]; flag = obj > obj_thresh; }
    int tot;
    const int pos = n + block_rank(flag, warp_tot, tot);
    if (flag) {
      const float cx = pb[k], cy = pb[(long long)cells + k], w = pb[2LL * cells + k], h = pb[3LL * cells + k];
      // torchvision _box_cxcywh_to_xyxy: x1 = cx - 0.5*w ...
      const float hw = __fmul_rn(0.5f, w), hh = __fmul_rn(0.5f, h);
      const float4 bx = make_float4(__fsub_rn(cx, hw), __fsub_rn(cy, hh), __fadd_rn(cx, hw), __fadd_rn(cy, hh));
      float mx = pb[5LL * cells + k];
      for (int c = 1; c < C; ++c) mx = fmaxf(mx, pb[(long long)(5 + c) * cells + k]);
      float score = __fmul_rn(mx, obj);
      if (score == 0.f) score = 0.f;  // -0 -> +0 (equal under torch's sort)
      unsigned u = __float_as_uint(score);
      u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // ascending-orderable
      cand_cell[pos] = k;
      cand_box[pos] = bx;
      if (do_nms) keys[pos] = ((unsigned long long)(~u) << 32) | (unsigned)pos;  // descending score, then index
    }
    n += tot;
  }
  __syncthreads();

  int kept = 0;
  if (!do_nms) {
    // iou_thresh == 0: NMS disabled, rows stay in grid order (prediction_formatting.py:80)
    for (int i = tid; i < n; i += NMS_THREADS) kept_ord[i] = i;
    kept = n;
    __syncthreads();
  } else {
    // ---- phase 2: bitonic sort of (score desc, index asc) keys in shared memory ----
    int Pn = 1;
    while (Pn < n) Pn <<= 1;
    for (int i = n + tid; i < Pn; i += NMS_THREADS) keys[i] = ~0ull;
    __syncthreads();
    for (int size = 2; size <= Pn; size <<= 1) {
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        for (int t = tid; t < (Pn >> 1); t += NMS_THREADS) {
          const int lo = 2 * t - (t & (stride - 1));
          const int hi = lo + stride;
          const bool up = ((lo & size) == 0);
          const unsigned long long a = keys[lo], c = keys[hi];
          if ((a > c) == up) { keys[lo] = c; keys[hi] = a; }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < n; i += NMS_THREADS) order[i] = (int)(keys[i] & 0xffffffffu);
    if (tid == 0) s_kept = 0;
    __syncthreads();

    // ---- phase 3: greedy suppression, 64 candidates per step ----
    for (int c0 = 0; c0 < n; c0 += NMS_CHUNK) {
      const int m = min(NMS_CHUNK, n - c0);
      const int kept_before = s_kept;
      if (tid < NMS_CHUNK) {
        s_mask[tid] = 0ull;
        if (tid < m) {
          const float4 bx = cand_box[order[c0 + tid]];
          s_cb[tid] = bx;
          s_ca[tid] = box_area(bx);
        }
      }
      if (tid == 0) s_sup = 0ull;
      __syncthreads();
      // (a) against every box kept so far
      {
        unsigned long long sup = 0ull;
        for (int k = tid; k < kept_before; k += NMS_THREADS) {
          const float4 kb = kept_box[k];
          const float ka = box_area(kb);
          for (int j = 0; j < m; ++j)
            if (iou_gt(kb, ka, s_cb[j], s_ca[j], iou_thresh)) sup |= (1ull << j);
        }
        unsigned lo = __reduce_or_sync(0xffffffffu, (unsigned)sup);
        unsigned hi = __reduce_or_sync(0xffffffffu, (unsigned)(sup >> 32));
        if ((tid & 31) == 0 && (lo | hi)) atomicOr(&s_sup, ((unsigned long long)hi << 32) | lo);
      }
      // (b) pairs inside the chunk: thread -> (i, 4 consecutive j)
      {
        const int i = tid >> 4, j0 = (tid & 15) << 2;
        if (i < m) {
          unsigned long long bits = 0ull;
          const float4 bi = s_cb[i];
          const float ai = s_ca[i];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = j0 + q;
            if (j > i && j < m && iou_gt(bi, ai, s_cb[j], s_ca[j], iou_thresh)) bits |= (1ull << j);
          }
          if (bits) atomicOr(&s_mask[i], bits);
        }
      }
      __syncthreads();
      // (c) serial resolve of the 64-candidate window
      if (tid == 0) {
        unsigned long long alive = (m == 64 ? ~0ull : ((1ull << m) - 1ull)) & ~s_sup;
        unsigned long long keepbits = 0ull;
        for (int i = 0; i < m; ++i)
          if ((alive >> i) & 1ull) { keepbits |= (1ull << i); alive &= ~s_mask[i]; }
        s_keepbits = keepbits;
        s_kept = kept_before + __popcll(keepbits);
      }
      __syncthreads();
      if (tid < m) {
        const unsigned long long kb = s_keepbits;
        if ((kb >> tid) & 1ull) {
          const int pos = kept_before + __popcll(kb & ((1ull << tid) - 1ull));
          kept_box[pos] = s_cb[tid];
          kept_ord[pos] = order[c0 + tid];
        }
      }
      __syncthreads();  // kept_box visible to the whole CTA for the next step (global, same CTA)
    }
    kept = s_kept;
  }

  // ---- phase 4: gather rows in output order, class-confidence filter, counts ----
  int out_n = 0;
  float* rb = rows + (long long)b * cells * D;
  int* kb = keep_index + (long long)b * cells;
  for (int r0 = 0; r0 < kept; r0 += NMS_THREADS) {
    const int r = r0 + tid;
    bool flag = false;
    int cell = 0, ord = 0, amax = 0;
    float mx = 0.f;
    if (r < kept) {
      ord = kept_ord[r];
      cell = cand_cell[ord];
      mx = pb[5LL * cells + cell];
      for (int c = 1; c < C; ++c) {
        const float v = pb[(long long)(5 + c) * cells + cell];
        if (v > mx) { mx = v; amax = c; }  // first max wins (torch.max / argmax)
      }
      flag = (min_cls > 0.f) ? (mx > min_cls) : true;
    }
    int tot;
    const int pos = out_n + block_rank(flag, warp_tot, tot);
    if (flag) {
      float* o = rb + (long long)pos * D;
      if (xyxy) {
        const float4 bx = cand_box[ord];
        o[0] = bx.x; o[1] = bx.y; o[2] = bx.z; o[3] = bx.w;
      } else {
#pragma unroll
        for (int d = 0; d < 4; ++d) o[d] = pb[(long long)d * cells + cell];
      }
      for (int d = 4; d < D; ++d) o[d] = pb[(long long)d * cells + cell];
      kb[pos] = cell;
      // count_cells_for_formatted_preds: rows with max > 0 (default threshold) are counted
      if (mx > 0.f) atomicAdd(&s_hist[amax], 1);
    }
    out_n += tot;
  }
  __syncthreads();
  if (tid == 0) keep_count[b] = out_n;
  if (tid < C && s_hist[tid]) atomicAdd(&class_counts[tid], (unsigned long long)s_hist[tid]);
}

static size_t nms_ws_bytes(int B, int cells) {
  return (size_t)B * cells * (sizeof(int) * 3 + sizeof(float4) * 2) + 256;
}

}  // namespace yg
using namespace yg;

extern "C" size_t yg_format_preds_workspace(int B, int num_classes, int Sy, int Sx) {
  return nms_ws_bytes(B, Sy * Sx);
}

extern "C" int yg_format_preds_batch(const float* preds, int B, int num_classes, int Sy, int Sx,
                                     float obj_thresh, double iou_thresh, int xyxy, float min_class_conf,
                                     int* keep_count, float* rows, int* keep_index, long long* class_counts,
                                     void* workspace, size_t workspace_bytes, void* stream) {
  YG_CHECK_ARG(preds && keep_count && rows && keep_index && class_counts, "format_preds: null pointer");
  const int cells = Sy * Sx;
  YG_CHECK_ARG(cells >= 1 && cells <= NMS_MAX_CELLS, "format_preds: Sy*Sx = %d exceeds %d", cells, NMS_MAX_CELLS);
  YG_CHECK_ARG(num_classes >= 1 && num_classes <= 27, "format_preds: num_classes %d", num_classes);
  cudaStream_t st = (cudaStream_t)stream;
  YG_CUDA(cudaMemsetAsync(class_counts, 0, sizeof(long long) * num_classes, st));
  if (B == 0) return YG_OK;
  const size_t need = nms_ws_bytes(B, cells);
  if (!workspace || workspace_bytes < need) {
    set_error("format_preds: workspace %zu < %zu", workspace_bytes, need);
    return YG_ERR_WORKSPACE;
  }
  // carve the workspace (16-byte aligned float4 arrays first)
  unsigned char* p = (unsigned char*)workspace;
  p = (unsigned char*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  NmsWs ws;
  ws.cand_box = (float4*)p; p += (size_t)B * cells * sizeof(float4);
  ws.kept_box = (float4*)p; p += (size_t)B * cells * sizeof(float4);
  ws.cand_cell = (int*)p; p += (size_t)B * cells * sizeof(int);
  ws.order = (int*)p; p += (size_t)B * cells * sizeof(int);
  ws.kept_ord = (int*)p;
  int P = 1;
  while (P < cells) P <<= 1;
  const size_t smem = (size_t)P * sizeof(unsigned long long);
  YG_CUDA(cudaFuncSetAttribute(format_preds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  format_preds_kernel<<<B, NMS_THREADS, smem, st>>>(preds, num_classes, cells, obj_thresh, iou_thresh,
                                                    iou_thresh > 0 ? 1 : 0, xyxy, min_class_conf, keep_count, rows,
                                                    keep_index, (unsigned long long*)class_counts, ws, P);
  YG_LAUNCH_CHECK("format_preds");
  return YG_OK;
}


// ------------------------------------------------------------------------------------------------
// Pairwise matching cost between labels and formatted predictions:
// cost[i][j] = 1 - box_iou(label_i, pred_j), the matrix format_preds_and_labels_v2 hands to the Hungarian solver
// (/root/reference/yogo/utils/prediction_formatting.py:296-298; torchvision.ops.box_iou = _box_inter_union, boxes.py).
// Every operation is a single correctly rounded fp32 operation in torchvision's order (never contracted to FMA), so the
// matrix is bit-identical to the reference's and the assignment computed from it is the same.
// ------------------------------------------------------------------------------------------------
namespace yg {
__global__ void box_iou_cost_kernel(const float* __restrict__ a, int a_stride, int na, const float* __restrict__ b, int b_stride,
                                    int nb, float* __restrict__ cost) {
  __shared__ float sb[256][5];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i0 = blockIdx.y * 16;
  // this thread's prediction box
  float bx1 = 0.f, by1 = 0.f, bx2 = 0.f, by2 = 0.f, barea = 0.f;
  if (j < nb) {
    const float* r = b + (long long)j * b_stride;
    bx1 = r[0]; by1 = r[1]; bx2 = r[2]; by2 = r[3];
    barea = __fmul_rn(__fsub_rn(bx2, bx1), __fsub_rn(by2, by1));
  }
  // 16 label boxes of this block row through shared memory
  if (threadIdx.x < 16 && i0 + threadIdx.x < na) {
    const float* r = a + (long long)(i0 + threadIdx.x) * a_stride;
    sb[threadIdx.x][0] = r[0]; sb[threadIdx.x][1] = r[1]; sb[threadIdx.x][2] = r[2]; sb[threadIdx.x][3] = r[3];
    sb[threadIdx.x][4] = __fmul_rn(__fsub_rn(r[2], r[0]), __fsub_rn(r[3], r[1]));
  }
  __syncthreads();
  if (j >= nb) return;
  for (int k = 0; k < 16 && i0 + k < na; ++k) {
    const float ax1 = sb[k][0], ay1 = sb[k][1], ax2 = sb[k][2], ay2 = sb[k][3], aarea = sb[k][4];
    const float ltx = ax1 > bx1 ? ax1 : bx1, lty = ay1 > by1 ? ay1 : by1;
    const float rbx = ax2 < bx2 ? ax2 : bx2, rby = ay2 < by2 ? ay2 : by2;
    float w = __fsub_rn(rbx, ltx), h = __fsub_rn(rby, lty);
    w = w < 0.f ? 0.f : w;   // clamp(min=0) keeps NaN
    h = h < 0.f ? 0.f : h;
    const float inter = __fmul_rn(w, h);
    const float uni = __fsub_rn(__fadd_rn(aarea, barea), inter);
    cost[(long long)(i0 + k) * nb + j] = __fsub_rn(1.f, __fdiv_rn(inter, uni));
  }
}
}  // namespace yg

extern "C" int yg_box_iou_cost(const float* labels, int label_stride, int n_labels, const float* preds, int pred_stride,
                               int n_preds, float* cost, void* stream) {
  YG_CHECK_ARG(n_labels >= 0 && n_preds >= 0 && label_stride >= 4 && pred_stride >= 4, "box_iou_cost: bad sizes");
  if (n_labels == 0 || n_preds == 0) return YG_OK;
  YG_CHECK_ARG(labels && preds && cost, "box_iou_cost: null pointer");
  dim3 grid(cdiv(n_preds, 256), cdiv(n_labels, 16), 1);
  yg::box_iou_cost_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, label_stride, n_labels, preds, pred_stride, n_preds, cost);
  YG_LAUNCH_CHECK("box_iou_cost");
  return YG_OK;
}
