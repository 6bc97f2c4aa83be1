// Input side of the training step: the bounding-box aware flips applied to every batch
// (/root/reference/yogo/data/data_transforms.py:51-98) and the rasterisation of ragged label lists into the
// (6, Sy, Sx) label tensor (/root/reference/yogo/data/yogo_dataset.py:24-46), whole batch per launch.
// HBM-bound byte/element shuffles: 8- or 16-byte accesses on both sides, nothing is re-read.
#include "common.cuh"

namespace yg {

// one block = FLIP_ROWS consecutive output rows, one thread = one V-element vector at a time;
// out[y][x] = in[vflip ? H-1-y : y][hflip ? W-1-x : x].  32-bit index arithmetic; the only 64-bit division is one per block.
constexpr int FLIP_ROWS = 16;
template <typename T, int V>
__global__ void __launch_bounds__(256) flip_images_kernel(const T* __restrict__ in, T* __restrict__ out, long long rows, int H,
                                                          int W, int hflip, int vflip) {
  const int vpr = W / V;   // vectors per row
  struct alignas(sizeof(T) * V) Vec { T e[V]; };
  const long long row0 = (long long)blockIdx.x * FLIP_ROWS;
  const long long plane0 = row0 / H;
  const int y0 = (int)(row0 - plane0 * H);
  const int nrows = (int)(rows - row0 < FLIP_ROWS ? rows - row0 : FLIP_ROWS);
  for (int idx = threadIdx.x; idx < nrows * vpr; idx += blockDim.x) {
    const int r = idx / vpr, vx = idx - r * vpr;
    int y = y0 + r;
    long long plane = plane0;
    while (y >= H) { y -= H; ++plane; }
    const int sy = vflip ? H - 1 - y : y;
    const int sv = hflip ? vpr - 1 - vx : vx;
    const Vec v = *reinterpret_cast<const Vec*>(in + (plane * H + sy) * (long long)W + sv * V);
    Vec o;
    if (hflip) {
#pragma unroll
      for (int k = 0; k < V; ++k) o.e[k] = v.e[V - 1 - k];
    } else {
      o = v;
    }
    *reinterpret_cast<Vec*>(out + (row0 + r) * (long long)W + vx * V) = o;
  }
}

// labels (N, 6, Sy, Sx) = [mask, x1, y1, x2, y2, class]: cells mirrored, x1' = 1 - x2, x2' = 1 - x1 (hflip) and
// y1' = 1 - y2, y2' = 1 - y1 (vflip) on EVERY cell, labelled or not, exactly as the reference's tensor expression does
__global__ void __launch_bounds__(256) flip_labels_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int Sy,
                                                          int Sx, int hflip, int vflip) {
  // blockIdx.y = n * 6 + c (one channel plane), blockIdx.x * 256 + threadIdx.x = cell of the plane: 32-bit arithmetic only
  const int cell = blockIdx.x * 256 + threadIdx.x;
  if (cell >= Sy * Sx) return;
  const int plane = blockIdx.y;
  const int n = plane / 6, c = plane - n * 6;
  const int y = cell / Sx, x = cell - y * Sx;
  const int sx = hflip ? Sx - 1 - x : x, sy = vflip ? Sy - 1 - y : y;
  int sc = c;
  bool comp = false;
  if (hflip && (c == 1 || c == 3)) { sc = 4 - c; comp = true; }
  if (vflip && (c == 2 || c == 4)) { sc = 6 - c; comp = true; }
  const float v = in[((long long)(n * 6 + sc) * Sy + sy) * Sx + sx];
  out[(long long)plane * Sy * Sx + cell] = comp ? __fsub_rn(1.f, v) : v;
}

// pass 1: owner[b][cell] = the LAST label of image b whose centre falls into the cell (the reference's Python loop
// overwrites in order); cell indices as torch computes them: ((x1 + x2) * Sx) // 2 in fp32, truncated to int, negative
// indices wrap like Python indexing, anything else out of range raises IndexError in the reference -> error flag
__global__ void label_owner_kernel(const float* __restrict__ labels, const int* __restrict__ offsets, int B, int Sy, int Sx,
                                   int* __restrict__ owner, int* __restrict__ err) {
  const int b = blockIdx.y;
  const int lo = offsets[b], hi = offsets[b + 1];
  for (int l = lo + blockIdx.x * blockDim.x + threadIdx.x; l < hi; l += gridDim.x * blockDim.x) {
    const float* r = labels + (long long)l * 5;
    const float fi = floorf(__fmul_rn(__fadd_rn(r[1], r[3]), (float)Sx) / 2.f);
    const float fj = floorf(__fmul_rn(__fadd_rn(r[2], r[4]), (float)Sy) / 2.f);
    int i = (int)fi, j = (int)fj;
    if (i < 0) i += Sx;
    if (j < 0) j += Sy;
    if (!(fi == fi) || !(fj == fj) || i < 0 || i >= Sx || j < 0 || j >= Sy) { atomicExch(err, 1); continue; }
    atomicMax(&owner[((long long)b * Sy + j) * Sx + i], l - lo);
  }
}

// pass 2: (B, 6, Sy, Sx) = [1, x1, y1, x2, y2, class] of the owning label, zeros elsewhere
__global__ void label_fill_kernel(const float* __restrict__ labels, const int* __restrict__ offsets, int B, int Sy, int Sx,
                                  const int* __restrict__ owner, float* __restrict__ out) {
  const long long cells = (long long)B * Sy * Sx;
  const long long plane = (long long)Sy * Sx;
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c < cells; c += (long long)gridDim.x * blockDim.x) {
    const long long b = c / plane, cell = c - b * plane;
    const int o = owner[c];
    float v[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (o >= 0) {
      const float* r = labels + ((long long)offsets[b] + o) * 5;
      v[0] = 1.f; v[1] = r[1]; v[2] = r[2]; v[3] = r[3]; v[4] = r[4]; v[5] = r[0];
    }
#pragma unroll
    for (int k = 0; k < 6; ++k) out[(b * 6 + k) * plane + cell] = v[k];
  }
}

}  // namespace yg
using namespace yg;

static unsigned flip_grid(long long rows) { return (unsigned)((rows + FLIP_ROWS - 1) / FLIP_ROWS); }
static int grid_for(long long work_items) {
  long long b = (work_items + 255) / 256;
  return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b));
}

extern "C" int yg_flip_images(const void* in, void* out, int dtype, int N, int C, int H, int W, int hflip, int vflip,
                              void* stream) {
  YG_CHECK_ARG(dtype == YG_U8 || dtype == YG_F32, "flip_images: dtype %d", dtype);
  const long long rows = (long long)N * C * H;
  if (rows == 0 || W == 0) return YG_OK;
  YG_CHECK_ARG(in && out && in != out, "flip_images: null or aliased pointers (the flip is out of place)");
  YG_CHECK_ARG(rows / FLIP_ROWS < (1LL << 31), "flip_images: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  const bool al = (((uintptr_t)in | (uintptr_t)out) & 15) == 0;
  if (dtype == YG_U8) {
    if (al && W % 16 == 0) flip_images_kernel<uint8_t, 16><<<flip_grid(rows), 256, 0, st>>>((const uint8_t*)in, (uint8_t*)out, rows, H, W, hflip, vflip);
    else if (al && W % 8 == 0) flip_images_kernel<uint8_t, 8><<<flip_grid(rows), 256, 0, st>>>((const uint8_t*)in, (uint8_t*)out, rows, H, W, hflip, vflip);
    else flip_images_kernel<uint8_t, 1><<<flip_grid(rows), 256, 0, st>>>((const uint8_t*)in, (uint8_t*)out, rows, H, W, hflip, vflip);
  } else {
    if (al && W % 4 == 0) flip_images_kernel<float, 4><<<flip_grid(rows), 256, 0, st>>>((const float*)in, (float*)out, rows, H, W, hflip, vflip);
    else flip_images_kernel<float, 1><<<flip_grid(rows), 256, 0, st>>>((const float*)in, (float*)out, rows, H, W, hflip, vflip);
  }
  YG_LAUNCH_CHECK("flip_images");
  return YG_OK;
}

extern "C" int yg_flip_labels(const float* in, float* out, int N, int Sy, int Sx, int hflip, int vflip, void* stream) {
  const long long total = (long long)N * 6 * Sy * Sx;
  if (total == 0) return YG_OK;
  YG_CHECK_ARG(in && out && in != out, "flip_labels: null or aliased pointers (the flip is out of place)");
  YG_CHECK_ARG((long long)N * 6 <= 65535, "flip_labels: batch %d too large for one launch", N);
  dim3 grid(cdiv((long long)Sy * Sx, 256), N * 6, 1);
  flip_labels_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(in, out, N, Sy, Sx, hflip, vflip);
  YG_LAUNCH_CHECK("flip_labels");
  return YG_OK;
}

extern "C" int yg_format_labels_batch(const float* labels, const int* offsets, int B, int max_labels, int Sy, int Sx,
                                      int* owner, float* out, int* err, void* stream) {
  if (B == 0) return YG_OK;
  YG_CHECK_ARG(offsets && owner && out && err && Sy > 0 && Sx > 0, "format_labels_batch: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  const long long cells = (long long)B * Sy * Sx;
  cudaMemsetAsync(owner, 0xFF, cells * sizeof(int), st);   // -1 = no label
  cudaMemsetAsync(err, 0, sizeof(int), st);
  if (max_labels > 0) {
    YG_CHECK_ARG(labels != nullptr, "format_labels_batch: labels is null");
    dim3 grid(cdiv(max_labels, 256), B, 1);
    label_owner_kernel<<<grid, 256, 0, st>>>(labels, offsets, B, Sy, Sx, owner, err);
    YG_LAUNCH_CHECK("format_labels owner");
  }
  label_fill_kernel<<<grid_for(cells), 256, 0, st>>>(labels, offsets, B, Sy, Sx, owner, out);
  YG_LAUNCH_CHECK("format_labels fill");
  return YG_OK;
}


// ------------------------------------------------------------------------------------------------
// Dropout2d keep-scales of ALL blocks of a step in one launch (nn.Dropout2d(p), model_defns.py:44-57: whole (n, c)
// planes are zeroed with probability p, the rest scaled by 1 / (1 - p)), plus the `num_batches_tracked += 1` of the
// BatchNorm layers.  Replaces ~5 torch elementwise launches per block (rand, >=, float, div, contiguous) and one per counter.
// Philox4x32-10 keyed by (seed, call counter, element): the counter lives in device memory and is advanced by the kernel
// itself, so that a CUDA-graph replay draws fresh masks.  state = [seed, counter, ticket].
// ------------------------------------------------------------------------------------------------
namespace yg {
__device__ __forceinline__ uint32_t mulhilo32(uint32_t a, uint32_t b, uint32_t* hi) {
  const unsigned long long p = (unsigned long long)a * b;
  *hi = (uint32_t)(p >> 32);
  return (uint32_t)p;
}
__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, hi1;
    const uint32_t lo0 = mulhilo32(0xD2511F53u, ctr.x, &hi0), lo1 = mulhilo32(0xCD9E8D57u, ctr.z, &hi1);
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u; key.y += 0xBB67AE85u;
  }
  return ctr;
}

// table: per block (offset, count, p as float bits, unused), nblk <= 16
__global__ void dropout_scales_kernel(float* __restrict__ out, const int4* __restrict__ table, int nblk, int total,
                                      unsigned long long* __restrict__ state, long long* const* __restrict__ counters, int ncounters) {
  const unsigned long long seed = state[0], call = state[1];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) {
    float p = 0.f;
    for (int b = 0; b < nblk; ++b) {
      const int4 t = table[b];
      if (i >= t.x && i < t.x + t.y) p = __int_as_float(t.z);
    }
    const uint4 r = philox4x32_10(make_uint4((uint32_t)i, 0u, (uint32_t)call, (uint32_t)(call >> 32)),
                                  make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)));
    const float u = (float)(r.x >> 8) * (1.0f / 16777216.0f);     // uniform in [0, 1)
    out[i] = (u >= p) ? 1.f / (1.f - p) : 0.f;
  }
  // the last block to finish advances the call counter and the BatchNorm counters (every block has read `call` by then)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(&state[2], 1ull);
    if (done == gridDim.x - 1) {
      state[2] = 0ull;
      state[1] = call + 1ull;
      for (int c = 0; c < ncounters; ++c) *counters[c] += 1;
    }
  }
}
}  // namespace yg

extern "C" int yg_dropout_scales(float* out, const void* table, int nblocks, int total, void* state, const void* counters,
                                 int ncounters, void* stream) {
  YG_CHECK_ARG(state && nblocks >= 0 && nblocks <= 16 && total >= 0 && ncounters >= 0, "dropout_scales: bad arguments");
  YG_CHECK_ARG(total == 0 || (out && table), "dropout_scales: null pointer");
  const int grid = total > 0 ? yg::cdiv(total, 256) : 1;
  yg::dropout_scales_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(out, (const int4*)table, nblocks, total,
                                                                    (unsigned long long*)state, (long long* const*)counters, ncounters);
  YG_LAUNCH_CHECK("dropout_scales");
  return YG_OK;
}
