// loss.cu — fused forward + gradient kernels of the reference's max-margin losses, over
// ragged candidate tables, and the flat Adam step.
//
// The reference builds every loss from ~25-40 tiny ATen kernels plus host-side
// np.ones masks (mlp/model.py:381-575).  Here one CTA handles one clip: it masks,
// applies the sigmoid, scores every candidate (track-pair) slot, takes the arg-max
// assignment (lowest index on ties, like torch.argmax), sums the hinge terms and
// writes d(loss)/d(logits) in the same pass.  Empty slots are never materialised: the
// reference sets them to -inf (sigma = 0, zero gradient), which is what skipping does.
#include <algorithm>
#include <cmath>

#include "common.cuh"

namespace lirec {
namespace loss {

__device__ __forceinline__ float sigmoidf(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float t = 0.f;
  for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];  // fixed order: deterministic
  return t;
}

constexpr int MAX_SLOTS = 128;

// MarginLoss (mlp/model.py:444-494) / MarginTrackRelsLoss (mlp/model.py:497-575).
__global__ void __launch_bounds__(128)
track_loss_kernel(const float* __restrict__ ints, const float* __restrict__ rels,
                  const int32_t* __restrict__ cand_off, int B, const int32_t* __restrict__ labels,
                  const int32_t* __restrict__ rels_label, const int32_t* __restrict__ gt_tracks,
                  const uint8_t* __restrict__ multilab, lirec_track_loss_cfg cfg,
                  float* __restrict__ loss_per_clip, int32_t* __restrict__ assign,
                  float* __restrict__ d_ints, float* __restrict__ d_rels) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_score[MAX_SLOTS];
  __shared__ float s_red[8];
  __shared__ int s_tstar;
  const int b = blockIdx.x;
  const int beg = cand_off[b], n = cand_off[b + 1] - beg;
  const int C = cfg.n_classes, R = cfg.n_rels;
  const bool has_rels = R > 0;
  const int y = labels[b];
  const int gt0 = gt_tracks[2 * b], gt1 = gt_tracks[2 * b + 1];
  // relationship class of the ground-truth pair(s); slots past the valid prefix carry the
  // reference's pad label 0 (classification_dataloader.py:430)
  int r0 = 0, r1 = 0;
  if (has_rels) {
    r0 = (gt0 < n) ? rels_label[beg + gt0] : 0;
    r1 = (gt1 < n) ? rels_label[beg + gt1] : 0;
  }
  const float* xi = ints + static_cast<int64_t>(beg) * C;
  const float* xr = has_rels ? rels + static_cast<int64_t>(beg) * R : nullptr;
  float* gi = d_ints + static_cast<int64_t>(beg) * C;
  float* gr = has_rels ? d_rels + static_cast<int64_t>(beg) * R : nullptr;
  const float invB = 1.0f / static_cast<float>(B);

  // ---- per-slot assignment score: sigma(ints[t,y]) (+ sigma(rels[t,r0])) ----
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    float sc = sigmoidf(xi[static_cast<int64_t>(t) * C + y]);
    if (has_rels) {
      const bool rel_on = rels_label[beg + t] != R;  // None-labelled slots are all -inf
      const float pr = (rel_on && r0 < R) ? sigmoidf(xr[static_cast<int64_t>(t) * R + r0]) : 0.f;
      sc = sc + pr;
    }
    s_score[t] = sc;
  }
  // ---- tr_cat_distr: the assignment is SAMPLED from (softmax_t(ints[t,y]) + softmax_t(rels[t,r0])) / 2
  // (model.py:468-471, 538-543; a relationship column that is -inf everywhere gives NaN -> 0 there).
  // The uniform draw comes from the counter hash of (seed, clip), mirrored by oracle/dropout.py.
  if (cfg.cat_distr && !cfg.tr_correct) {
    __shared__ float s_p[MAX_SLOTS];
    if (threadIdx.x == 0) {
      float mx = -INFINITY, mxr = -INFINITY;
      for (int t = 0; t < n; ++t) {
        mx = fmaxf(mx, xi[static_cast<int64_t>(t) * C + y]);
        if (has_rels && r0 < R && rels_label[beg + t] != R) mxr = fmaxf(mxr, xr[static_cast<int64_t>(t) * R + r0]);
      }
      float zi = 0.f, zr = 0.f;
      for (int t = 0; t < n; ++t) {
        zi += expf(xi[static_cast<int64_t>(t) * C + y] - mx);
        if (has_rels && r0 < R && rels_label[beg + t] != R) zr += expf(xr[static_cast<int64_t>(t) * R + r0] - mxr);
      }
      float total = 0.f;
      for (int t = 0; t < n; ++t) {
        float pt = expf(xi[static_cast<int64_t>(t) * C + y] - mx) / zi;
        if (has_rels) {
          float pr = 0.f;
          if (r0 < R && rels_label[beg + t] != R && zr > 0.f) pr = expf(xr[static_cast<int64_t>(t) * R + r0] - mxr) / zr;
          pt = 0.5f * (pt + pr);
        }
        s_p[t] = pt;
        total += pt;
      }
      const uint32_t h = fmix32(cfg.seed ^ fmix32(static_cast<uint32_t>(b) + 0x51ED270Bu));
      const float u = static_cast<float>(h >> 8) * (1.0f / 16777216.0f) * total;
      float acc = 0.f;
      int pick = n - 1;
      for (int t = 0; t < n; ++t) {
        acc += s_p[t];
        if (u < acc) { pick = t; break; }
      }
      s_tstar = pick;
      assign[b] = pick;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0 && !(cfg.cat_distr && !cfg.tr_correct)) {
    int best = 0;
    if (!cfg.tr_correct) {
      float bv = s_score[0];
      for (int t = 1; t < n; ++t)
        if (s_score[t] > bv) { bv = s_score[t]; best = t; }
      // empty slots score exactly 0 and sit after the valid prefix, so they can only win a
      // tie against a valid slot scoring 0, which the lowest-index rule gives to the valid one
    }
    s_tstar = best;
    assign[b] = best;
  }
  __syncthreads();
  const int ts = s_tstar;
  const float pos = sigmoidf(xi[static_cast<int64_t>(ts) * C + y]);
  float pos_r = 0.f;
  bool pos_r_live = false;
  if (has_rels) {
    pos_r_live = (rels_label[beg + ts] != R) && (r0 < R);
    pos_r = pos_r_live ? sigmoidf(xr[static_cast<int64_t>(ts) * R + r0]) : 0.f;
  }
  const float mi = cfg.margin - pos, mr = cfg.margin - pos_r;
  const float wi = cfg.lymbda * invB, wr = invB;

  float loss_i = 0.f, cnt_i = 0.f, loss_r = 0.f, cnt_r = 0.f;
  if (!cfg.max_neg) {
    // ---- sum over every unmasked (slot, class) negative ----
    for (int e = threadIdx.x; e < n * C; e += blockDim.x) {
      const int t = e / C, c = e - t * C;
      bool neg = multilab[static_cast<int64_t>(b) * C + c] != 0;
      if (cfg.tr_correct) neg = neg && !(c == y && (t == gt0 || t == gt1));
      else neg = neg && (c != y);
      float g = 0.f;
      if (neg) {
        const float s = sigmoidf(xi[e]);
        const float h = mi + s;
        if (h > 0.f) { loss_i += h; cnt_i += 1.f; g = wi * s * (1.f - s); }
      }
      gi[e] = g;
    }
    if (has_rels) {
      for (int e = threadIdx.x; e < n * R; e += blockDim.x) {
        const int t = e / R, r = e - t * R;
        const int lab = rels_label[beg + t];
        bool neg = lab != R;
        if (cfg.tr_correct) neg = neg && (r != lab);
        else neg = neg && (r != r0) && (r != r1);
        float g = 0.f;
        if (neg) {
          const float s = sigmoidf(xr[e]);
          const float h = mr + s;
          if (h > 0.f) { loss_r += h; cnt_r += 1.f; g = wr * s * (1.f - s); }
        }
        gr[e] = g;
      }
    }
  } else {
    // ---- max-negative variant: one hinge per slot on its hardest negative; the
    // reference also adds relu(m - pos) for every EMPTY slot (model.py:485-486,558-562)
    for (int e = threadIdx.x; e < n * C; e += blockDim.x) gi[e] = 0.f;
    if (has_rels)
      for (int e = threadIdx.x; e < n * R; e += blockDim.x) gr[e] = 0.f;
    __syncthreads();
    for (int t = threadIdx.x; t < cfg.max_slots; t += blockDim.x) {
      float best = 0.f;
      int bc = -1;
      if (t < n) {
        for (int c = 0; c < C; ++c) {
          bool neg = multilab[static_cast<int64_t>(b) * C + c] != 0;
          if (cfg.tr_correct) neg = neg && !(c == y && (t == gt0 || t == gt1));
          else neg = neg && (c != y);
          if (!neg) continue;
          const float s = sigmoidf(xi[static_cast<int64_t>(t) * C + c]);
          if (s > best) { best = s; bc = c; }
        }
      }
      const float h = mi + best;
      if (h > 0.f) {
        loss_i += h; cnt_i += 1.f;
        if (bc >= 0) gi[static_cast<int64_t>(t) * C + bc] = wi * best * (1.f - best);
      }
      if (has_rels) {
        float bestr = 0.f;
        int br = -1;
        if (t < n) {
          const int lab = rels_label[beg + t];
          if (lab != R) {
            for (int r = 0; r < R; ++r) {
              const bool neg = cfg.tr_correct ? (r != lab) : (r != r0 && r != r1);
              if (!neg) continue;
              const float s = sigmoidf(xr[static_cast<int64_t>(t) * R + r]);
              if (s > bestr) { bestr = s; br = r; }
            }
          }
        }
        const float hr = mr + bestr;
        if (hr > 0.f) {
          loss_r += hr; cnt_r += 1.f;
          if (br >= 0) gr[static_cast<int64_t>(t) * R + br] = wr * bestr * (1.f - bestr);
        }
      }
    }
  }
  loss_i = block_sum(loss_i, s_red);
  cnt_i = block_sum(cnt_i, s_red);
  if (has_rels) {
    loss_r = block_sum(loss_r, s_red);
    cnt_r = block_sum(cnt_r, s_red);
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    // the positive receives minus the number of active hinges
    gi[static_cast<int64_t>(ts) * C + y] += -cnt_i * wi * pos * (1.f - pos);
    if (has_rels && pos_r_live) gr[static_cast<int64_t>(ts) * R + r0] += -cnt_r * wr * pos_r * (1.f - pos_r);
    loss_per_clip[b] = cfg.lymbda * loss_i * invB + loss_r * invB;
  }
}

// Row-wise hinge: MaxMarginCrossEntropyLoss (model.py:422-441) and both terms of
// MultiTaskMaxMargin (model.py:381-419).
__global__ void __launch_bounds__(128)
rowmargin_kernel(const float* __restrict__ logits, int64_t ld, int rows, int C,
                 const int32_t* __restrict__ labels, const uint8_t* __restrict__ weights, float margin,
                 float scale, float* __restrict__ loss_per_row, float* __restrict__ d_logits, int64_t d_ld) {
  pdl_wait();
  pdl_trigger();
  __shared__ float s_red[8];
  const int b = blockIdx.x;
  const int y = labels[b];
  float* g = d_logits + static_cast<int64_t>(b) * d_ld;
  if (y < 0) {  // row not selected (relationship label None)
    for (int c = threadIdx.x; c < C; c += blockDim.x) g[c] = 0.f;
    if (threadIdx.x == 0) loss_per_row[b] = 0.f;
    return;
  }
  const float* x = logits + static_cast<int64_t>(b) * ld;
  const float pos = sigmoidf(x[y]);
  float l = 0.f, cnt = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const bool neg = (c != y) && (weights == nullptr || weights[static_cast<int64_t>(b) * C + c] != 0);
    float gv = 0.f;
    if (neg) {
      const float s = sigmoidf(x[c]);
      const float h = margin - pos + s;
      if (h > 0.f) { l += h; cnt += 1.f; gv = scale * s * (1.f - s); }
    }
    if (c != y) g[c] = gv;
  }
  l = block_sum(l, s_red);
  cnt = block_sum(cnt, s_red);
  if (threadIdx.x == 0) {
    g[y] = -cnt * scale * pos * (1.f - pos);
    loss_per_row[b] = l * scale;
  }
}


// Row-wise softmax cross-entropy, forward + gradient: both terms of MultiTaskCrossEntropyLoss
// (model.py:357-378; F.cross_entropy with optional class weights, mean reduction).  The caller passes
// scale = 1 / sum_i w[y_i] over the selected rows; rows with label < 0 are skipped.
//   loss_row = scale * w[y] * (logsumexp(x) - x[y]);  d/dx_c = scale * w[y] * (softmax_c - [c == y])
__global__ void __launch_bounds__(128)
ce_kernel(const float* __restrict__ logits, int64_t ld, int C, const int32_t* __restrict__ labels,
          const float* __restrict__ class_w, float scale, float* __restrict__ loss_per_row,
          float* __restrict__ d_logits, int64_t d_ld) {
  __shared__ float s_red[8];
  const int b = blockIdx.x;
  const int y = labels[b];
  float* g = d_logits + static_cast<int64_t>(b) * d_ld;
  if (y < 0) {
    for (int c = threadIdx.x; c < C; c += blockDim.x) g[c] = 0.f;
    if (threadIdx.x == 0) loss_per_row[b] = 0.f;
    return;
  }
  const float* x = logits + static_cast<int64_t>(b) * ld;
  float mx = -INFINITY;
  for (int c = threadIdx.x; c < C; c += blockDim.x) mx = fmaxf(mx, x[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_red[warp] = mx;
  __syncthreads();
  mx = fmaxf(fmaxf(s_red[0], s_red[1]), fmaxf(s_red[2], s_red[3]));
  float z = 0.f;
  for (int c = threadIdx.x; c < C; c += blockDim.x) z += expf(x[c] - mx);
  z = block_sum(z, s_red);
  const float w = (class_w ? class_w[y] : 1.0f) * scale;
  for (int c = threadIdx.x; c < C; c += blockDim.x) g[c] = w * (expf(x[c] - mx) / z - (c == y ? 1.f : 0.f));
  if (threadIdx.x == 0) loss_per_row[b] = w * (logf(z) + mx - x[y]);
}

// Prediction arg-maxes of the evaluation loop (reference utils/evaluation.py:114-175, 179-271), one CTA
// per clip over its valid candidate slots; empty slots are -inf there, i.e. never win.
//   out[b] = { pr_track, joint_t, joint_c, joint_r, cls_at_gt0, cls_at_gt1, rel_at_gt0, rel_at_gt1 }
//   pr_track : argmax_t sigma(ints[t,y]) (+ sigma(rels (+) 0)[t, r_gt])           (:137, :221-222)
//   joint    : first arg-max of sigma(ints)[t,c] (+ sigma(rels (+) 0)[t,r]) over the flattened
//              (t,c[,r]) index, the sum taken in double like numpy's float32 + float64  (:144-147, :229-235)
//   cls/rel at gt_i : argmax over classes / relationship classes of the raw logits of slot gt_tracks[b,i]
//              (-1 when that slot is empty)                                              (:152, :241-243)
// Ties resolve to the lowest flattened index, like np.argmax.
__global__ void __launch_bounds__(128)
predict_kernel(const float* __restrict__ ints, const float* __restrict__ rels,
               const int32_t* __restrict__ cand_off, const int32_t* __restrict__ labels,
               const int32_t* __restrict__ rels_label, const int32_t* __restrict__ gt_tracks, int C, int R,
               int32_t* __restrict__ out) {
  __shared__ double s_val[128];
  __shared__ long long s_idx[128];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int beg = cand_off[b], n = cand_off[b + 1] - beg;
  const bool has_rels = R > 0;
  const int R1 = has_rels ? R + 1 : 1;
  const int y = labels[b];
  const int gt0 = gt_tracks[2 * b], gt1 = gt_tracks[2 * b + 1];
  const float* xi = ints + static_cast<int64_t>(beg) * C;
  const float* xr = has_rels ? rels + static_cast<int64_t>(beg) * R : nullptr;
  // relationship label of the ground-truth slot: evaluation.py:208 takes gt_rels[:, 0]
  const int rg = has_rels ? rels_label[beg] : 0;

  auto reduce_first_max = [&](double v, long long i) -> long long {
    s_val[tid] = v; s_idx[tid] = i;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (tid < o) {
        const double v2 = s_val[tid + o]; const long long i2 = s_idx[tid + o];
        if (v2 > s_val[tid] || (v2 == s_val[tid] && i2 < s_idx[tid])) { s_val[tid] = v2; s_idx[tid] = i2; }
      }
      __syncthreads();
    }
    const long long r = s_idx[0];
    __syncthreads();
    return r;
  };
  const double NEG = -1e300;
  // ---- track assignment for the ground-truth class (and relationship) ----
  {
    double v = NEG; long long idx = 0x7fffffffffffLL;
    for (int t = tid; t < n; t += blockDim.x) {
      double sc = static_cast<double>(sigmoidf(xi[static_cast<int64_t>(t) * C + y]));
      if (has_rels) sc += (rg < R) ? static_cast<double>(sigmoidf(xr[static_cast<int64_t>(t) * R + rg])) : 0.0;
      if (sc > v) { v = sc; idx = t; }
    }
    const long long r = reduce_first_max(v, idx);
    if (tid == 0) out[8 * b + 0] = static_cast<int32_t>(r);
  }
  // ---- joint (t, c[, r]) arg-max ----
  {
    double v = NEG; long long idx = 0x7fffffffffffLL;
    const long long total = static_cast<long long>(n) * C * R1;
    for (long long e = tid; e < total; e += blockDim.x) {
      const int t = static_cast<int>(e / (C * R1));
      const int rem = static_cast<int>(e - static_cast<long long>(t) * C * R1);
      const int c = rem / R1, r = rem - c * R1;
      double sc = static_cast<double>(sigmoidf(xi[static_cast<int64_t>(t) * C + c]));
      if (has_rels) sc += (r < R) ? static_cast<double>(sigmoidf(xr[static_cast<int64_t>(t) * R + r])) : 0.0;
      if (sc > v) { v = sc; idx = e; }
    }
    const long long r = reduce_first_max(v, idx);
    if (tid == 0) {
      const int t = static_cast<int>(r / (C * R1));
      const int rem = static_cast<int>(r - static_cast<long long>(t) * C * R1);
      out[8 * b + 1] = t; out[8 * b + 2] = rem / R1; out[8 * b + 3] = has_rels ? rem % R1 : -1;
    }
  }
  // ---- class / relationship arg-max at the ground-truth slots ----
  for (int i = 0; i < 2; ++i) {
    const int g = i ? gt1 : gt0;
    double v = NEG; long long idx = 0x7fffffffffffLL;
    if (g < n)
      for (int c = tid; c < C; c += blockDim.x) {
        const double sc = xi[static_cast<int64_t>(g) * C + c];
        if (sc > v) { v = sc; idx = c; }
      }
    long long r = reduce_first_max(v, idx);
    if (tid == 0) out[8 * b + 4 + i] = (g < n) ? static_cast<int32_t>(r) : -1;
    v = NEG; idx = 0x7fffffffffffLL;
    if (has_rels && g < n)
      for (int c = tid; c < R; c += blockDim.x) {
        const double sc = xr[static_cast<int64_t>(g) * R + c];
        if (sc > v) { v = sc; idx = c; }
      }
    r = reduce_first_max(v, idx);
    if (tid == 0) out[8 * b + 6 + i] = (has_rels && g < n) ? static_cast<int32_t>(r) : -1;
  }
}

// torch.optim.Adam (coupled L2) over the flat parameter buffer + bf16 shadow refresh.
__global__ void __launch_bounds__(256)
adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
            float* __restrict__ v, __nv_bfloat16* __restrict__ pb, int64_t n, float lr, float beta1,
            float beta2, float eps, float wd, float bc1, float bc2_sqrt, float grad_scale) {
  pdl_wait();
  pdl_trigger();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t tid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const float step = lr / bc1;
  // 128-bit path over the 16-byte aligned bulk (the flat buffers are: segments are 256-byte aligned), same
  // per-element arithmetic as the scalar tail below
  const bool vec = ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                     reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(pb) & 7) == 0;
  const int64_t n4 = vec ? n / 4 : 0;
  for (int64_t i = tid; i < n4; i += stride) {
    float4 pv = reinterpret_cast<const float4*>(p)[i];
    const float4 gr = reinterpret_cast<const float4*>(g)[i];
    float4 mv = reinterpret_cast<const float4*>(m)[i];
    float4 vv = reinterpret_cast<const float4*>(v)[i];
    float* pp = &pv.x; const float* gg = &gr.x; float* mm = &mv.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float gv = gg[k] * grad_scale + wd * pp[k];
      mm[k] = beta1 * mm[k] + (1.f - beta1) * gv;
      vp[k] = beta2 * vp[k] + (1.f - beta2) * gv * gv;
      pp[k] -= step * mm[k] / (sqrtf(vp[k]) / bc2_sqrt + eps);
    }
    reinterpret_cast<float4*>(m)[i] = mv;
    reinterpret_cast<float4*>(v)[i] = vv;
    reinterpret_cast<float4*>(p)[i] = pv;
    if (pb) {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(pv.x, pv.y), hi = __floats2bfloat162_rn(pv.z, pv.w);
      uint2 w;
      w.x = *reinterpret_cast<const uint32_t*>(&lo);
      w.y = *reinterpret_cast<const uint32_t*>(&hi);
      reinterpret_cast<uint2*>(pb)[i] = w;
    }
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += stride) {
    float pv = p[i];
    const float gv = g[i] * grad_scale + wd * pv;
    const float mv = beta1 * m[i] + (1.f - beta1) * gv;
    const float vv = beta2 * v[i] + (1.f - beta2) * gv * gv;
    m[i] = mv;
    v[i] = vv;
    pv -= step * mv / (sqrtf(vv) / bc2_sqrt + eps);
    p[i] = pv;
    if (pb) pb[i] = __float2bfloat16_rn(pv);
  }
}

}  // namespace loss
}  // namespace lirec

using namespace lirec;

extern "C" int lirec_loss_track_fwd_bwd(const float* ints, const float* rels, const int32_t* cand_off,
                                        int32_t B, const int32_t* labels, const int32_t* rels_label,
                                        const int32_t* gt_tracks, const uint8_t* multilab,
                                        lirec_track_loss_cfg cfg, float* loss_per_clip, int32_t* assign,
                                        float* d_ints, float* d_rels, void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(ints && cand_off && labels && gt_tracks && multilab && loss_per_clip && assign && d_ints,
                "track loss: null pointer");
  LIREC_REQUIRE(cfg.n_classes > 0 && cfg.n_rels >= 0, "track loss: n_classes=%d n_rels=%d", cfg.n_classes,
                cfg.n_rels);
  LIREC_REQUIRE(cfg.n_rels == 0 || (rels && rels_label && d_rels), "track loss: rel tensors missing");
  LIREC_REQUIRE(cfg.max_slots > 0 && cfg.max_slots <= loss::MAX_SLOTS, "track loss: max_slots=%d (limit %d)",
                cfg.max_slots, loss::MAX_SLOTS);
  if (B <= 0) return LIREC_OK;
  LIREC_CUDA_OK(launch_pdl(loss::track_loss_kernel, dim3(B), dim3(128), 0, static_cast<cudaStream_t>(stream), ints, rels,
                           cand_off, B, labels, rels_label, gt_tracks, multilab, cfg, loss_per_clip, assign, d_ints,
                           d_rels));
  note_launch();
  return LIREC_OK;
}

extern "C" int lirec_loss_rowmargin_fwd_bwd(const float* logits, int64_t ld, int32_t rows, int32_t C,
                                            const int32_t* labels, const uint8_t* weights, float margin,
                                            float scale, float* loss_per_row, float* d_logits, int64_t d_ld,
                                            void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(logits && labels && loss_per_row && d_logits && C > 0, "rowmargin loss: bad arguments");
  if (rows <= 0) return LIREC_OK;
  LIREC_CUDA_OK(launch_pdl(loss::rowmargin_kernel, dim3(rows), dim3(128), 0, static_cast<cudaStream_t>(stream), logits,
                           ld, rows, C, labels, weights, margin, scale, loss_per_row, d_logits, d_ld));
  note_launch();
  return LIREC_OK;
}

extern "C" int lirec_loss_ce_fwd_bwd(const float* logits, int64_t ld, int32_t rows, int32_t C, const int32_t* labels,
                                     const float* class_weights, float scale, float* loss_per_row,
                                     float* d_logits, int64_t d_ld, void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(logits && labels && loss_per_row && d_logits && C > 0, "ce loss: bad arguments");
  if (rows <= 0) return LIREC_OK;
  loss::ce_kernel<<<rows, 128, 0, static_cast<cudaStream_t>(stream)>>>(logits, ld, C, labels, class_weights, scale,
                                                                      loss_per_row, d_logits, d_ld);
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  return LIREC_OK;
}

extern "C" int lirec_predict_tracks(const float* ints, const float* rels, const int32_t* cand_off, int32_t B,
                                    const int32_t* labels, const int32_t* rels_label, const int32_t* gt_tracks,
                                    int32_t n_classes, int32_t n_rels, int32_t* out, void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(ints && cand_off && labels && gt_tracks && out && n_classes > 0, "predict: bad arguments");
  LIREC_REQUIRE(n_rels == 0 || (rels && rels_label), "predict: relationship tensors missing");
  if (B <= 0) return LIREC_OK;
  loss::predict_kernel<<<B, 128, 0, static_cast<cudaStream_t>(stream)>>>(ints, rels, cand_off, labels, rels_label,
                                                                        gt_tracks, n_classes, n_rels, out);
  LIREC_CUDA_OK(cudaGetLastError());
  note_launch();
  return LIREC_OK;
}

extern "C" int lirec_adam_flat(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                               void* param_bf16, int64_t n, float lr, float beta1, float beta2, float eps,
                               float weight_decay, int32_t step, float grad_scale, void* stream) {
  return lirec_adam_flat_ex(param, grad, exp_avg, exp_avg_sq, param_bf16, n, lr, beta1, beta2, eps, weight_decay,
                            step, grad_scale, 0, stream);
}

extern "C" int lirec_adam_flat_ex(float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                                  void* param_bf16, int64_t n, float lr, float beta1, float beta2, float eps,
                                  float weight_decay, int32_t step, float grad_scale, int32_t coresident,
                                  void* stream) {
  LIREC_ENTER();
  LIREC_REQUIRE(param && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "adam: bad arguments");
  if (n == 0) return LIREC_OK;
  // bias corrections in double like torch's Python-side arithmetic
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), static_cast<double>(step)));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(beta2), static_cast<double>(step)));
  // coresident: CTAs that fit next to a resident GEMM CTA (common.cuh), so a pass launched on a side stream really
  // runs during backward instead of after it; same kernel, same arithmetic, same element -> thread order
  static const int co_threads = coresident_threads(reinterpret_cast<const void*>(loss::adam_kernel), 256);
  const int threads = coresident ? co_threads : 256;
  const int grid = static_cast<int>(std::min<int64_t>((n / 4 + threads - 1) / threads + 1, 148 * 16));
  LIREC_CUDA_OK(launch_pdl(loss::adam_kernel, dim3(grid), dim3(threads), 0, static_cast<cudaStream_t>(stream), param, grad,
                           exp_avg, exp_avg_sq, reinterpret_cast<__nv_bfloat16*>(param_bf16), n, lr, beta1, beta2, eps,
                           weight_decay, bc1, sqrtf(bc2), grad_scale));
  note_launch();
  return LIREC_OK;
}
