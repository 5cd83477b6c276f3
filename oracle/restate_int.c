/* Plain-C restatement of the INTEGER / INDEX arithmetic of the MinGraph-UNet graph block.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): only tests/ load this library, as a second,
 * independently written checker next to the numpy forms in oracle/restate.py.  Paths below are
 * relative to the reference root (MinGraph-UNet/).
 *
 * Parity status: the reference has no golden vectors for this path; these functions are pinned by
 * the fixtures generated from the untouched reference (tests/golden/kat3_edge_index.npz: SHA-256 of
 * edge_index for 12 grid shapes; block_images.npz: hard labels, region edge lists, un-pooled maps).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* E = 2*(Hp*(Wp-1) + Wp*(Hp-1)); a single patch has no edges
 * (preprocessing/graph_construction/patch_graph_construction.py:78-97). */
int64_t oracle_grid_num_edges(int Hp, int Wp) { return 2 * ((int64_t)Hp * (Wp - 1) + (int64_t)Wp * (Hp - 1)); }

/* edge_index (2,E) int64, row 0 = source, row 1 = target, in the reference's emission order:
 * row-major walk over the patches; for each, the pair with the right neighbour (both directions),
 * then the pair with the neighbour below (patch_graph_construction.py:80-95). */
void oracle_grid_edge_index(int Hp, int Wp, int64_t* ei) {
  const int64_t E = oracle_grid_num_edges(Hp, Wp);
  int64_t* src = ei;
  int64_t* tgt = ei + E;
  int64_t k = 0;
  for (int r = 0; r < Hp; ++r) {
    for (int c = 0; c < Wp; ++c) {
      const int64_t n = (int64_t)r * Wp + c;
      if (c + 1 < Wp) {
        src[k] = n;     tgt[k] = n + 1; ++k;
        src[k] = n + 1; tgt[k] = n;     ++k;
      }
      if (r + 1 < Hp) {
        src[k] = n;      tgt[k] = n + Wp; ++k;
        src[k] = n + Wp; tgt[k] = n;      ++k;
      }
    }
  }
}

/* Complete digraph on K regions: triu_indices(K,K,1) pairs (s,t) row-major, then the reversed pairs
 * (scripts/train_end_to_end.py:376-378).  ei is (2, K*(K-1)). */
void oracle_complete_edge_index(int K, int64_t* ei) {
  const int64_t half = (int64_t)K * (K - 1) / 2, E = 2 * half;
  int64_t k = 0;
  for (int s = 0; s < K; ++s)
    for (int t = s + 1; t < K; ++t) {
      ei[k] = s;        ei[E + k] = t;
      ei[half + k] = t; ei[E + half + k] = s;
      ++k;
    }
}

/* Stable CSR of a COO edge list by target (by_target != 0) or by source: neighbours of a node appear
 * in ascending COO edge id, the order torch's CPU scatter_add_ visits them (model/gat/graph_attention.py:91,112).
 * rowptr has N+1 entries, col and eid E entries (eid nullable).  Returns the number of out-of-range indices. */
int64_t oracle_csr_from_coo(const int64_t* ei, int64_t E, int N, int by_target, int32_t* rowptr, int32_t* col, int32_t* eid) {
  const int64_t* key = by_target ? ei + E : ei;
  const int64_t* val = by_target ? ei : ei + E;
  int64_t bad = 0;
  memset(rowptr, 0, sizeof(int32_t) * (size_t)(N + 1));
  for (int64_t k = 0; k < E; ++k) {
    if (key[k] < 0 || key[k] >= N || val[k] < 0 || val[k] >= N) { ++bad; continue; }
    ++rowptr[key[k] + 1];
  }
  for (int i = 0; i < N; ++i) rowptr[i + 1] += rowptr[i];
  int32_t* fill = (int32_t*)malloc(sizeof(int32_t) * (size_t)(N > 0 ? N : 1));
  memcpy(fill, rowptr, sizeof(int32_t) * (size_t)N);
  for (int64_t k = 0; k < E; ++k) {
    if (key[k] < 0 || key[k] >= N || val[k] < 0 || val[k] >= N) continue;
    const int32_t slot = fill[key[k]]++;
    col[slot] = (int32_t)val[k];
    if (eid) eid[slot] = (int32_t)k;
  }
  free(fill);
  return bad;
}

/* hard = argmax(S, dim=1) (scripts/train_end_to_end.py:356): first maximum wins, NaN counts as maximal
 * (torch.argmax semantics). */
void oracle_argmax_rows(const float* S, int N, int K, int32_t* labels) {
  for (int n = 0; n < N; ++n) {
    const float* row = S + (size_t)n * K;
    int best = 0;
    float bv = row[0];
    for (int c = 1; c < K; ++c) {
      const float v = row[c];
      if (!(bv != bv) && (v > bv || v != v)) { bv = v; best = c; }
    }
    labels[n] = best;
  }
}

/* Source index of F.interpolate(mode='nearest') for one axis (scripts/train_end_to_end.py:417-421):
 * identity when the sizes agree, dst>>1 for exact doubling, otherwise
 * min(floor(dst * (float)in/(float)out), in-1) evaluated in float32 like ATen's nearest_idx. */
void oracle_nearest_index(int out_size, int in_size, int32_t* idx) {
  if (out_size == in_size) {
    for (int d = 0; d < out_size; ++d) idx[d] = d;
    return;
  }
  if (out_size == 2 * in_size) {
    for (int d = 0; d < out_size; ++d) idx[d] = d >> 1;
    return;
  }
  const float scale = (float)in_size / (float)out_size;
  for (int d = 0; d < out_size; ++d) {
    int s = (int)floorf((float)d * scale);
    idx[d] = s < in_size - 1 ? s : in_size - 1;
  }
}

/* Un-pool (scripts/train_end_to_end.py:404-421): out[d][y][x] = table[labels[iy*Wp+ix]][d] with the
 * nearest indices above; labels may be NULL (table then has one row per patch). */
void oracle_unpool_nearest(const float* table, const int32_t* labels, int D, int Hp, int Wp, int H, int W, float* out) {
  int32_t* iy = (int32_t*)malloc(sizeof(int32_t) * (size_t)H);
  int32_t* ix = (int32_t*)malloc(sizeof(int32_t) * (size_t)W);
  oracle_nearest_index(H, Hp, iy);
  oracle_nearest_index(W, Wp, ix);
  for (int d = 0; d < D; ++d)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x) {
        const int n = iy[y] * Wp + ix[x];
        const int row = labels ? labels[n] : n;
        out[((size_t)d * H + y) * W + x] = table[(size_t)row * D + d];
      }
  free(iy);
  free(ix);
}
