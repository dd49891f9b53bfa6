// softpool.cu — softmax-weighted segmented reduction over ragged sequences, forward and backward.
//
// PARITY UNPINNED BY CONSTRUCTION.  The reference has no attention and no softmax pooling: its temporal
// pooling is np.max over axis 0 (mixed_utils/mixed_features.py:54, 61, 105) and its pooling over context clips
// a masked mean (mlp/model.py:301-304).  BASELINE.json's north_star names "masked temporal attention pooling as a
// segmented softmax-weighted reduction"; SURVEY.md §0.1 resolves that as ONE reduction family over the same
// offset tables, {max, mean, softmax-weighted}, of which only max and mean have a reference oracle
// (lirec_seg_reduce_f32).  This is the third member:
//
//     y[s, c] = sum_{r in seg s} w[r, c] * x[r, c],      w[., c] = softmax_r( beta * score[r, c] )
//
// with score = x itself (per element; `scores == NULL`), a per-element score tensor [total, dim], or ONE score
// per row [total] (attention pooling: score_ld == 0).  It is pinned against a float64 numpy statement and
// against its two limits, which ARE the reference's poolings: beta = 0 (or all-zero scores) is the masked mean,
// beta -> inf with score = x is the max (tests/test_pooling_gpu.py).  An empty segment yields zeros, like the
// other two modes.
//
// Layout: one CTA per (segment, 512-column group), one thread per four columns, the segment's rows walked
// sequentially with an online softmax (running max, running normaliser, running weighted sum), so every global
// access is a coalesced 16-byte load and the result is deterministic.  HBM-bound: forward reads x (+ scores)
// once; backward reads x, scores, dy and the forward's log-normaliser once and writes dx (+ dscores) once.
#include "rows.cuh"

namespace lirec {
namespace softpool {

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

struct Online {          // online softmax-weighted sum of one column
  float m, z, acc;
  __device__ __forceinline__ void init() { m = -INFINITY; z = 0.f; acc = 0.f; }
  __device__ __forceinline__ void add(float s, float x) {
    if (s > m) {
      const float k = __expf(m - s);           // exp(-inf) = 0 on the first row
      z = z * k + 1.f;
      acc = acc * k + x;
      m = s;
    } else {
      const float e = __expf(s - m);
      z += e;
      acc += e * x;
    }
  }
};

// score_mode: 0 = score is x, 1 = per-element tensor [total, dim], 2 = one score per row [total]
__global__ void __launch_bounds__(128)
fwd_kernel(const float* __restrict__ x, const float* __restrict__ scores, int score_mode,
           const int32_t* __restrict__ seg_off, int dim, float beta, float* __restrict__ out, int64_t out_ld,
           float* __restrict__ lse, int64_t lse_ld) {
  const int seg = blockIdx.x;
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (col >= dim) return;
  const int beg = seg_off[seg], end = seg_off[seg + 1];
  Online o[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) o[k].init();
  // rows four at a time: the loads of a batch are issued together (the online update is a dependent chain, and
  // with one row per iteration a thread had a single 16-byte load in flight: 35-53 % of the HBM peak)
  int r = beg;
  for (; r + 3 < end; r += 4) {
    float4 a[4], s[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) a[k] = ld4(x + static_cast<int64_t>(r + k) * dim + col);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (score_mode == 0) s[k] = a[k];
      else if (score_mode == 1) s[k] = ld4(scores + static_cast<int64_t>(r + k) * dim + col);
      else { const float t = scores[r + k]; s[k] = make_float4(t, t, t, t); }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      o[0].add(beta * s[k].x, a[k].x); o[1].add(beta * s[k].y, a[k].y);
      o[2].add(beta * s[k].z, a[k].z); o[3].add(beta * s[k].w, a[k].w);
    }
  }
  for (; r < end; ++r) {
    const float4 a = ld4(x + static_cast<int64_t>(r) * dim + col);
    float4 s;
    if (score_mode == 0) s = a;
    else if (score_mode == 1) s = ld4(scores + static_cast<int64_t>(r) * dim + col);
    else { const float t = scores[r]; s = make_float4(t, t, t, t); }
    o[0].add(beta * s.x, a.x); o[1].add(beta * s.y, a.y); o[2].add(beta * s.z, a.z); o[3].add(beta * s.w, a.w);
  }
  float4 y = make_float4(0.f, 0.f, 0.f, 0.f), l = make_float4(0.f, 0.f, 0.f, 0.f);
  if (end > beg) {
    y = make_float4(o[0].acc / o[0].z, o[1].acc / o[1].z, o[2].acc / o[2].z, o[3].acc / o[3].z);
    l = make_float4(o[0].m + __logf(o[0].z), o[1].m + __logf(o[1].z), o[2].m + __logf(o[2].z), o[3].m + __logf(o[3].z));
  }
  st4(out + static_cast<int64_t>(seg) * out_ld + col, y);
  if (lse) {
    if (score_mode == 2) {
      if (col == 0) lse[seg] = l.x;                      // one normaliser per segment
    } else {
      st4(lse + static_cast<int64_t>(seg) * lse_ld + col, l);
    }
  }
}

// Backward, per-element scores (score_mode 0 / 1):
//   w = exp(beta * s - lse);  dx = w * dy (+ ds when score is x);  ds = beta * w * (x - y) * dy
__global__ void __launch_bounds__(128)
bwd_elem_kernel(const float* __restrict__ x, const float* __restrict__ scores, int score_mode,
                const int32_t* __restrict__ seg_off, int dim, float beta, const float* __restrict__ y, int64_t y_ld,
                const float* __restrict__ lse, int64_t lse_ld, const float* __restrict__ dy, int64_t dy_ld,
                float* __restrict__ dx, float* __restrict__ dscores) {
  const int seg = blockIdx.x;
  const int col = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
  if (col >= dim) return;
  const int beg = seg_off[seg], end = seg_off[seg + 1];
  if (end <= beg) return;
  const float4 yy = ld4(y + static_cast<int64_t>(seg) * y_ld + col);
  const float4 ll = ld4(lse + static_cast<int64_t>(seg) * lse_ld + col);
  const float4 g = ld4(dy + static_cast<int64_t>(seg) * dy_ld + col);
#pragma unroll 4
  for (int r = beg; r < end; ++r) {
    const int64_t at = static_cast<int64_t>(r) * dim + col;
    const float4 a = ld4(x + at);
    const float4 s = score_mode == 0 ? a : ld4(scores + at);
    const float4 w = make_float4(__expf(beta * s.x - ll.x), __expf(beta * s.y - ll.y), __expf(beta * s.z - ll.z),
                                 __expf(beta * s.w - ll.w));
    float4 gx = make_float4(w.x * g.x, w.y * g.y, w.z * g.z, w.w * g.w);
    const float4 gs = make_float4(beta * gx.x * (a.x - yy.x), beta * gx.y * (a.y - yy.y), beta * gx.z * (a.z - yy.z),
                                  beta * gx.w * (a.w - yy.w));
    if (score_mode == 0) { gx.x += gs.x; gx.y += gs.y; gx.z += gs.z; gx.w += gs.w; }
    else if (dscores) st4(dscores + at, gs);
    st4(dx + at, gx);
  }
}

// Backward, one score per row (score_mode 2): one CTA per segment walks all columns;
//   w_r = exp(beta * s_r - lse);  dx[r, :] = w_r * dy;  ds_r = beta * w_r * sum_c (x[r, c] - y[c]) * dy[c]
__global__ void __launch_bounds__(256)
bwd_row_kernel(const float* __restrict__ x, const float* __restrict__ scores, const int32_t* __restrict__ seg_off,
               int dim, float beta, const float* __restrict__ y, int64_t y_ld, const float* __restrict__ lse,
               const float* __restrict__ dy, int64_t dy_ld, float* __restrict__ dx, float* __restrict__ dscores) {
  const int seg = blockIdx.x;
  const int beg = seg_off[seg], end = seg_off[seg + 1];
  if (end <= beg) return;
  __shared__ float part[8];
  const float l = lse[seg];
  const float* yrow = y + static_cast<int64_t>(seg) * y_ld;
  const float* grow = dy + static_cast<int64_t>(seg) * dy_ld;
  for (int r = beg; r < end; ++r) {
    const float w = __expf(beta * scores[r] - l);
    float dot = 0.f;
    for (int c = threadIdx.x * 4; c < dim; c += blockDim.x * 4) {
      const float4 a = ld4(x + static_cast<int64_t>(r) * dim + c), yy = ld4(yrow + c), g = ld4(grow + c);
      st4(dx + static_cast<int64_t>(r) * dim + c, make_float4(w * g.x, w * g.y, w * g.z, w * g.w));
      dot += (a.x - yy.x) * g.x + (a.y - yy.y) * g.y + (a.z - yy.z) * g.z + (a.w - yy.w) * g.w;
    }
    if (dscores) {                                          // fixed-order block reduction: deterministic
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
      if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = dot;
      __syncthreads();
      if (threadIdx.x == 0) {
        float t = 0.f;
        for (int k = 0; k < (blockDim.x >> 5); ++k) t += part[k];
        dscores[r] = beta * w * t;
      }
      __syncthreads();
    }
  }
}

static int check(const float* x, const int32_t* seg_off, int dim, int score_mode, const float* scores) {
  LIREC_REQUIRE(x && seg_off, "seg_softmax_pool: null argument");
  LIREC_REQUIRE(dim > 0 && dim % 4 == 0, "seg_softmax_pool: dim=%d must be a positive multiple of 4", dim);
  LIREC_REQUIRE(score_mode >= 0 && score_mode <= 2, "seg_softmax_pool: score_mode=%d", score_mode);
  LIREC_REQUIRE((score_mode == 0) == (scores == nullptr), "seg_softmax_pool: scores must be NULL exactly for score_mode 0");
  LIREC_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                    (score_mode != 1 || (reinterpret_cast<uintptr_t>(scores) & 15) == 0),
                "seg_softmax_pool: x / scores not 16-byte aligned");
  return LIREC_OK;
}

}  // namespace softpool
}  // namespace lirec

using namespace lirec;

extern "C" int lirec_seg_softmax_pool_fwd(const float* x, const float* scores, int32_t score_mode,
                                          const int32_t* seg_off, int32_t nseg, int32_t dim, float beta,
                                          float* out, int64_t out_ld, float* lse, int64_t lse_ld, void* stream) {
  LIREC_ENTER();
  int rc = softpool::check(x, seg_off, dim, score_mode, scores);
  if (rc != LIREC_OK) return rc;
  LIREC_REQUIRE(out && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && out_ld % 4 == 0,
                "seg_softmax_pool_fwd: output not 16-byte aligned");
  LIREC_REQUIRE(!lse || score_mode == 2 || ((reinterpret_cast<uintptr_t>(lse) & 15) == 0 && lse_ld % 4 == 0),
                "seg_softmax_pool_fwd: lse not 16-byte aligned");
  if (nseg <= 0) return LIREC_OK;
  dim3 grid(nseg, (dim / 4 + 127) / 128);
  softpool::fwd_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(x, scores, score_mode, seg_off, dim, beta,
                                                                            out, out_ld, lse, lse_ld);
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  return LIREC_OK;
}

extern "C" int lirec_seg_softmax_pool_bwd(const float* x, const float* scores, int32_t score_mode,
                                          const int32_t* seg_off, int32_t nseg, int32_t dim, float beta,
                                          const float* out, int64_t out_ld, const float* lse, int64_t lse_ld,
                                          const float* d_out, int64_t d_out_ld, float* d_x, float* d_scores,
                                          void* stream) {
  LIREC_ENTER();
  int rc = softpool::check(x, seg_off, dim, score_mode, scores);
  if (rc != LIREC_OK) return rc;
  LIREC_REQUIRE(out && lse && d_out && d_x, "seg_softmax_pool_bwd: null argument");
  LIREC_REQUIRE(((reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d_out) | reinterpret_cast<uintptr_t>(d_x)) & 15) == 0 &&
                    out_ld % 4 == 0 && d_out_ld % 4 == 0,
                "seg_softmax_pool_bwd: buffers not 16-byte aligned");
  if (nseg <= 0) return LIREC_OK;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (score_mode == 2) {
    softpool::bwd_row_kernel<<<nseg, 256, 0, s>>>(x, scores, seg_off, dim, beta, out, out_ld, lse, d_out, d_out_ld, d_x,
                                                  d_scores);
  } else {
    LIREC_REQUIRE((reinterpret_cast<uintptr_t>(lse) & 15) == 0 && lse_ld % 4 == 0 &&
                      (!d_scores || (reinterpret_cast<uintptr_t>(d_scores) & 15) == 0),
                  "seg_softmax_pool_bwd: lse / d_scores not 16-byte aligned");
    dim3 grid(nseg, (dim / 4 + 127) / 128);
    softpool::bwd_elem_kernel<<<grid, 128, 0, s>>>(x, scores, score_mode, seg_off, dim, beta, out, out_ld, lse, lse_ld,
                                                   d_out, d_out_ld, d_x, d_scores);
  }
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  return LIREC_OK;
}
