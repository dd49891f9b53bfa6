// collate.cu — host-side batch assembly behind lirec_collate_tables (include/lirec_b200.h).
//
// The reference assembles a batch by np.tile / hstack / vstack of cached 6912-d rows in DataLoader
// workers (mixed_utils/classification_dataloader.py:329-334, 393-416, 474-497; default collate
// mlp/train.py:33-37).  Here a record carries only index triples into the dataset's two feature banks;
// this function turns the concatenated triples of a batch into every integer table of a packed batch
// (lirec_batch) in ONE pass of counting sorts, laid out in one int32 arena that is copied to the device
// with one transfer.  Pure host code: no CUDA call, usable in DataLoader worker processes.
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace {

// Unique ids of a set of table columns, in the order "ids the candidate rows use, ascending, then the
// ids only context rows use, ascending"; ids are shifted by B so the private zero rows -B..-1 come first.
struct Remap {
  std::vector<int32_t> lut;    // shifted id -> batch bank row
  std::vector<int32_t> order;  // batch bank row -> shifted id
  int32_t n_ints = 0;
};

void build_remap(Remap& r, int64_t size, const int32_t* ints, int64_t n_ints_vals, const int32_t* ctx,
                 int64_t n_ctx_vals) {
  std::vector<uint8_t> seen(size, 0);
  for (int64_t i = 0; i < n_ints_vals; ++i) seen[ints[i]] = 1;
  for (int64_t i = 0; i < n_ctx_vals; ++i) seen[ctx[i]] |= 2;
  r.lut.assign(size, -1);
  r.order.clear();
  for (int64_t v = 0; v < size; ++v)
    if (seen[v] & 1) {
      r.lut[v] = (int32_t)r.order.size();
      r.order.push_back((int32_t)v);
    }
  r.n_ints = (int32_t)r.order.size();
  for (int64_t v = 0; v < size; ++v)
    if (seen[v] == 2) {
      r.lut[v] = (int32_t)r.order.size();
      r.order.push_back((int32_t)v);
    }
}

// CSR inverse of one column of a [n, 3] table: off [n_unique + 1], idx [n] (positions ascending per id).
void csr_inverse(const int32_t* tbl, int64_t n, int col, int32_t n_unique, int32_t* off, int32_t* idx) {
  std::memset(off, 0, sizeof(int32_t) * (size_t)(n_unique + 1));
  for (int64_t i = 0; i < n; ++i) ++off[tbl[3 * i + col] + 1];
  for (int32_t u = 0; u < n_unique; ++u) off[u + 1] += off[u];
  std::vector<int32_t> cur(off, off + n_unique);
  for (int64_t i = 0; i < n; ++i) idx[cur[tbl[3 * i + col]]++] = (int32_t)i;
}

}  // namespace

extern "C" int64_t lirec_collate_arena_bound(int64_t B, int64_t n_cand, int64_t n_ctx, int32_t has_ctx) {
  const int64_t Ni = n_cand, Nx = has_ctx ? n_ctx : 0;
  int64_t n = (B + 1) + 3 * Ni + B + 2 * B + 2 * Ni;   // cand_off, cand_rows, labels, gt_tracks, cand_clip/slot
  n += (Ni + 1) + 2 * (2 * Ni + 1) + 3 * Ni;           // inverse CSRs of the candidate table
  if (has_ctx) {
    n += (Ni + 1) + 3 * Nx + Nx + Ni;                  // ctx_off, ctx_rows, ctx_owner, rels_label
    n += (Ni + Nx + 1) + 2 * (2 * (Ni + Nx) + 1) + 3 * Nx;
  }
  n += (Ni + Nx) + 2 * (Ni + Nx);                      // clip_src, track_src
  return n;
}

extern "C" int lirec_collate_tables(const int32_t* cand_host, const int32_t* cand_counts_host, int32_t B,
                                    const int32_t* ctx_host, const int32_t* ctx_counts_host,
                                    int32_t zero_clip, int32_t n_clip_rows, int32_t n_track_rows,
                                    int32_t max_slots, int32_t* arena_host, int64_t arena_cap,
                                    int64_t* layout_host, int32_t* sizes_host) {
  using lirec::fail;
  lirec::reset_launch_count();   // host-only: launches nothing
  if (!cand_host || !cand_counts_host || !arena_host || !layout_host || !sizes_host || B <= 0)
    return fail(LIREC_ERR_ARG, "collate: null argument or empty batch");
  const bool has_ctx = ctx_counts_host != nullptr;
  int64_t Ni = 0;
  for (int32_t b = 0; b < B; ++b) {
    const int32_t c = cand_counts_host[b];
    if (c < 1 || c > max_slots)
      return fail(LIREC_ERR_ARG, "collate: clip %d has %d candidates (every clip needs 1..%d)", b, c, max_slots);
    Ni += c;
  }
  int64_t Nx = 0;
  if (has_ctx)
    for (int64_t i = 0; i < Ni; ++i) {
      if (ctx_counts_host[i] < 0) return fail(LIREC_ERR_ARG, "collate: negative context count");
      Nx += ctx_counts_host[i];
    }
  if (Nx > 0 && !ctx_host) return fail(LIREC_ERR_ARG, "collate: context counts without context rows");
  if (arena_cap < lirec_collate_arena_bound(B, Ni, Nx, has_ctx))
    return fail(LIREC_ERR_ARG, "collate: arena too small (use lirec_collate_arena_bound)");
  if (Ni + Nx > (int64_t)1 << 29) return fail(LIREC_ERR_LIMIT, "collate: batch too large for int32 tables");

  // shifted ids: private zero row of clip b = B - 1 - b  (-1 - b, plus B), bank row r = r + B
  std::vector<int32_t> cand_t(3 * Ni), ctx_t(3 * Nx), cand_clip_of(Ni);
  {
    int64_t i = 0;
    for (int32_t b = 0; b < B; ++b)
      for (int32_t s = 0; s < cand_counts_host[b]; ++s, ++i) {
        cand_clip_of[i] = b;
        const int32_t* r = cand_host + 3 * i;
        if (r[0] < 0 || r[0] >= n_clip_rows || r[1] < 0 || r[1] >= n_track_rows || r[2] < 0 || r[2] >= n_track_rows)
          return fail(LIREC_ERR_ARG, "collate: candidate row %lld references a bank row out of range", (long long)i);
        const int32_t z = B - 1 - b;
        cand_t[3 * i + 0] = r[0] == zero_clip ? z : r[0] + B;
        cand_t[3 * i + 1] = r[1] == 0 ? z : r[1] + B;
        cand_t[3 * i + 2] = r[2] == 0 ? z : r[2] + B;
      }
  }
  if (has_ctx) {
    int64_t j = 0;
    for (int64_t i = 0; i < Ni; ++i) {
      const int32_t z = B - 1 - cand_clip_of[i];
      for (int32_t k = 0; k < ctx_counts_host[i]; ++k, ++j) {
        const int32_t* r = ctx_host + 3 * j;
        if (r[0] < 0 || r[0] >= n_clip_rows || r[1] < 0 || r[1] >= n_track_rows || r[2] < 0 || r[2] >= n_track_rows)
          return fail(LIREC_ERR_ARG, "collate: context row %lld references a bank row out of range", (long long)j);
        ctx_t[3 * j + 0] = r[0] == zero_clip ? z : r[0] + B;
        ctx_t[3 * j + 1] = r[1] == 0 ? z : r[1] + B;
        ctx_t[3 * j + 2] = r[2] == 0 ? z : r[2] + B;
      }
    }
  }

  // unique bank rows: clip column on its own, the two track columns together
  Remap clip, track;
  {
    std::vector<int32_t> a(Ni), c(Nx);
    for (int64_t i = 0; i < Ni; ++i) a[i] = cand_t[3 * i];
    for (int64_t j = 0; j < Nx; ++j) c[j] = ctx_t[3 * j];
    build_remap(clip, (int64_t)n_clip_rows + B, a.data(), Ni, c.data(), Nx);
    a.resize(2 * Ni);
    c.resize(2 * Nx);
    for (int64_t i = 0; i < Ni; ++i) a[2 * i] = cand_t[3 * i + 1], a[2 * i + 1] = cand_t[3 * i + 2];
    for (int64_t j = 0; j < Nx; ++j) c[2 * j] = ctx_t[3 * j + 1], c[2 * j + 1] = ctx_t[3 * j + 2];
    build_remap(track, (int64_t)n_track_rows + B, a.data(), 2 * Ni, c.data(), 2 * Nx);
  }
  const int32_t n_clip = (int32_t)clip.order.size(), n_track = (int32_t)track.order.size();
  sizes_host[0] = n_clip, sizes_host[1] = clip.n_ints, sizes_host[2] = n_track, sizes_host[3] = track.n_ints;

  // arena layout, in the order lirec_b200/packing.py:_INT_TABLES stages the tables
  int64_t pos = 0;
  int t = 0;
  auto put = [&](int64_t n) {
    layout_host[2 * t] = n < 0 ? 0 : pos, layout_host[2 * t + 1] = n;
    ++t;
    int32_t* p = arena_host + pos;
    if (n > 0) pos += n;
    return p;
  };
  const int64_t NO = -1;
  int32_t* cand_off = put(B + 1);
  int32_t* cand_rows = put(3 * Ni);
  int32_t* ctx_off = put(has_ctx ? Ni + 1 : NO);
  int32_t* ctx_rows = put(has_ctx ? 3 * Nx : NO);
  int32_t* ctx_owner = put(has_ctx ? Nx : NO);
  put(B);                      // labels: filled by the caller
  put(has_ctx ? Ni : NO);      // rels_label: filled by the caller
  put(2 * B);                  // gt_tracks: filled by the caller
  int32_t *inv_off[6], *inv_idx[6];
  for (int s = 0; s < 3; ++s) {
    inv_off[s] = put((s == 0 ? clip.n_ints : track.n_ints) + 1);
    inv_idx[s] = put(Ni);
  }
  for (int s = 0; s < 3; ++s) {
    inv_off[3 + s] = put(has_ctx ? (s == 0 ? n_clip : n_track) + 1 : NO);
    inv_idx[3 + s] = put(has_ctx ? Nx : NO);
  }
  int32_t* cand_clip = put(Ni);
  int32_t* cand_slot = put(Ni);
  int32_t* clip_src = put(n_clip);
  int32_t* track_src = put(n_track);

  cand_off[0] = 0;
  {
    int64_t i = 0;
    for (int32_t b = 0; b < B; ++b) {
      for (int32_t s = 0; s < cand_counts_host[b]; ++s, ++i) cand_clip[i] = b, cand_slot[i] = s;
      cand_off[b + 1] = (int32_t)i;
    }
  }
  for (int64_t i = 0; i < Ni; ++i) {
    cand_rows[3 * i + 0] = clip.lut[cand_t[3 * i + 0]];
    cand_rows[3 * i + 1] = track.lut[cand_t[3 * i + 1]];
    cand_rows[3 * i + 2] = track.lut[cand_t[3 * i + 2]];
  }
  for (int s = 0; s < 3; ++s)
    csr_inverse(cand_rows, Ni, s, s == 0 ? clip.n_ints : track.n_ints, inv_off[s], inv_idx[s]);
  if (has_ctx) {
    ctx_off[0] = 0;
    int64_t j = 0;
    for (int64_t i = 0; i < Ni; ++i) {
      for (int32_t k = 0; k < ctx_counts_host[i]; ++k, ++j) ctx_owner[j] = (int32_t)i;
      ctx_off[i + 1] = (int32_t)j;
    }
    for (j = 0; j < Nx; ++j) {
      ctx_rows[3 * j + 0] = clip.lut[ctx_t[3 * j + 0]];
      ctx_rows[3 * j + 1] = track.lut[ctx_t[3 * j + 1]];
      ctx_rows[3 * j + 2] = track.lut[ctx_t[3 * j + 2]];
    }
    for (int s = 0; s < 3; ++s)
      csr_inverse(ctx_rows, Nx, s, s == 0 ? n_clip : n_track, inv_off[3 + s], inv_idx[3 + s]);
  }
  // dataset-bank rows behind the batch banks (private zero rows read the dataset's zero rows)
  for (int32_t u = 0; u < n_clip; ++u) clip_src[u] = clip.order[u] < B ? zero_clip : clip.order[u] - B;
  for (int32_t u = 0; u < n_track; ++u) track_src[u] = track.order[u] < B ? 0 : track.order[u] - B;
  return LIREC_OK;
}

// Ragged gather of the records of `B` clips out of dataset-level tables (CSR by clip for the candidate triples,
// CSR by candidate for the context triples) into the contiguous arrays lirec_collate_tables takes.  Replaces ~10
// numpy calls per clip (or the vectorised fancy-indexing equivalent) in the DataLoader worker.  Host code.
// Returns the number of candidate rows through n_cand_out and of context rows through n_ctx_out; the output
// buffers must hold max_cand / max_ctx rows (the call fails, writing nothing beyond them, if they do not).
extern "C" int lirec_collate_gather(const int64_t* ds_cand_off, const int32_t* ds_cand, const int64_t* ds_ctx_off,
                                    const int32_t* ds_ctx_cnt, const int32_t* ds_ctx, const int64_t* idx, int32_t B,
                                    int32_t* cand_out, int32_t* counts_out, int64_t* cand_pos_out, int64_t max_cand,
                                    int32_t* ctx_out, int32_t* ctx_counts_out, int64_t max_ctx, int64_t* n_cand_out,
                                    int64_t* n_ctx_out) {
  using lirec::fail;
  lirec::reset_launch_count();
  if (!ds_cand_off || !ds_cand || !idx || !cand_out || !counts_out || !n_cand_out || !n_ctx_out || B <= 0)
    return fail(LIREC_ERR_ARG, "collate_gather: null argument or empty batch");
  const bool has_ctx = ds_ctx_off != nullptr;
  if (has_ctx && (!ds_ctx_cnt || !ds_ctx || !ctx_out || !ctx_counts_out))
    return fail(LIREC_ERR_ARG, "collate_gather: context tables missing");
  int64_t ni = 0, nx = 0;
  for (int32_t b = 0; b < B; ++b) {
    const int64_t c0 = ds_cand_off[idx[b]], c1 = ds_cand_off[idx[b] + 1];
    if (ni + (c1 - c0) > max_cand) return fail(LIREC_ERR_ARG, "collate_gather: candidate buffer too small");
    counts_out[b] = static_cast<int32_t>(c1 - c0);
    std::memcpy(cand_out + 3 * ni, ds_cand + 3 * c0, sizeof(int32_t) * 3 * static_cast<size_t>(c1 - c0));
    for (int64_t c = c0; c < c1; ++c, ++ni) {
      if (cand_pos_out) cand_pos_out[ni] = c;
      if (has_ctx) {
        const int32_t k = ds_ctx_cnt[c];
        if (nx + k > max_ctx) return fail(LIREC_ERR_ARG, "collate_gather: context buffer too small");
        ctx_counts_out[ni] = k;
        std::memcpy(ctx_out + 3 * nx, ds_ctx + 3 * ds_ctx_off[c], sizeof(int32_t) * 3 * static_cast<size_t>(k));
        nx += k;
      }
    }
  }
  *n_cand_out = ni;
  *n_ctx_out = nx;
  return LIREC_OK;
}
