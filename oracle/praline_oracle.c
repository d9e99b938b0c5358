/*
 * praline_oracle.c -- CPU restatement of PRALINE's pairwise DP alignment core.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity oracle for the CUDA path
 * in praline_b200/csrc.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path never
 * calls into it and has no CPU fallback.
 *
 * Parity is PINNED: tests/test_oracle_pinned.py checks every function below
 *   (1) cell-for-cell against the reference's own compiled C extension
 *       (oracle/_ref/cext*.so, built by oracle/Makefile from
 *       /root/reference/praline/util/cext.c with setup.py's flags), and
 *   (2) against golden score/path vectors produced by the reference's
 *       PairwiseAligner component (tests/golden/, made by
 *       tests/golden/make_golden.py).
 *
 * Citations are relative to /root/reference/.  Nothing here is copied: the
 * reference works on strided numpy objects through the CPython API; this is a
 * plain-C statement of the same arithmetic on dense row-major buffers.
 *
 * Layouts (all dense, row-major):
 *   m  [L1][L2]          f32  match scores
 *   g1 [L1][2], g2[L2][2] f32 gap {open, extend} per position
 *   o  [L1+1][L2+1][3]   f32  DP planes: 0 = M, 1 = U (insert up), 2 = L (insert left)
 *   t  [L1+1][L2+1][3]   u8   tie flags (bit values below)
 *   z  [L1+1][L2+1]      u8   1 = masked cell (skipped, stays 0)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* praline/util/cext.c:9-15, praline/util/align.py:15-21 */
#define TB_MM (1 << 1)
#define TB_MU (1 << 2)
#define TB_ML (1 << 3)
#define TB_UO (1 << 4)
#define TB_UE (1 << 5)
#define TB_LO (1 << 6)
#define TB_LE (1 << 7)

/* praline/util/cext.c:27-31 */
enum { MODE_GLOBAL = 0, MODE_LOCAL = 1, MODE_SG_BOTH = 2, MODE_SG_ONE = 3, MODE_SG_TWO = 4 };

#define O(y, x, k) o[(((size_t)(y)) * W + (size_t)(x)) * 3 + (k)]
#define T(y, x, k) t[(((size_t)(y)) * W + (size_t)(x)) * 3 + (k)]

/* ------------------------------------------------------------------------
 * Match-score matrix.  praline/util/cext.c:308-455 (cext_build_scores) and
 * :33-97 (score_match_prof_prof).
 *
 * m[y][x] = sum over track sets n of
 *             sum_{i in nz(P1n[y])} sum_{j in nz(P2n[x])} P1n[y][i]*P2n[x][j]*Sn[i][j]
 * The reference walks the nonzero column ids of each profile row in ascending
 * order (component/align.py:449-458 builds those lists with ndarray.nonzero()),
 * accumulates one f32 per set starting from 0, then adds the per-set sums in
 * set order (cext.c:388,413).  The source writes p1*p2*score (cext.c:89); the
 * reference build (gcc -O3 -ffast-math, setup.py:28-30) evaluates it as
 * (p2*score)*p1 -- written out here so that this file needs no -ffast-math.
 * ---------------------------------------------------------------------- */
void orc_build_scores(int n_sets, const float *const *P1, const float *const *P2,
                      const float *const *S, const int *A, int L1, int L2, float *m)
{
    for (int y = 0; y < L1; y++) {
        for (int x = 0; x < L2; x++) {
            float score = 0.0f;
            for (int n = 0; n < n_sets; n++) {
                const int a = A[n];
                const float *r1 = P1[n] + (size_t)y * a;
                const float *r2 = P2[n] + (size_t)x * a;
                const float *s = S[n];
                volatile float acc = 0.0f; /* volatile: forbid re-association / FMA contraction */
                for (int i = 0; i < a; i++) {
                    if (r1[i] == 0.0f) continue;
                    for (int j = 0; j < a; j++) {
                        if (r2[j] == 0.0f) continue;
                        volatile float prod = r2[j] * s[(size_t)i * a + j];
                        prod = prod * r1[i];
                        acc = acc + prod;
                    }
                }
                score = score + acc;
            }
            m[(size_t)y * L2 + x] = score;
        }
    }
}

/* ------------------------------------------------------------------------
 * Border initialisation.  praline/component/align.py:357-385.
 * o and t must be zero-filled by the caller first (align.py:358-359).
 * ---------------------------------------------------------------------- */
void orc_init_borders(int mode, const float *g1, const float *g2, int L1, int L2,
                      float *o, uint8_t *t)
{
    const size_t W = (size_t)L2 + 1;
    const float NINF = -INFINITY;
    for (int y = 0; y <= L1; y++) for (int k = 0; k < 3; k++) O(y, 0, k) = NINF; /* :367 */
    for (int x = 0; x <= L2; x++) for (int k = 0; k < 3; k++) O(0, x, k) = NINF; /* :368 */
    O(0, 0, 0) = 0.0f;                                                           /* :369 */

    if (mode == MODE_SG_BOTH || mode == MODE_SG_ONE) {                           /* :371-372 */
        for (int y = 0; y <= L1; y++) O(y, 0, 1) = 0.0f;
    } else {                                                                     /* :373-377 */
        if (L1 > 0) O(0, 0, 1) = g1[0] - g1[1];
        for (int y = 1; y <= L1; y++) {
            /* np.arange(L1) * g1[:,1] + g1[0,0]: float64 product/sum cast to f32 on store */
            double v = (double)(y - 1) * (double)g1[(size_t)(y - 1) * 2 + 1] + (double)g1[0];
            O(y, 0, 1) = (float)v;
            T(y, 0, 1) = TB_UE;
        }
    }
    if (mode == MODE_SG_BOTH || mode == MODE_SG_TWO) {                           /* :379-380 */
        for (int x = 0; x <= L2; x++) O(0, x, 2) = 0.0f;
    } else {                                                                     /* :381-385 */
        if (L2 > 0) O(0, 0, 2) = g2[0] - g2[1];
        for (int x = 1; x <= L2; x++) {
            double v = (double)(x - 1) * (double)g2[(size_t)(x - 1) * 2 + 1] + (double)g2[0];
            O(0, x, 2) = (float)v;
            T(0, x, 2) = TB_LE;
        }
    }
}

/* ------------------------------------------------------------------------
 * Three-state fill with tie flags.  praline/util/cext.c:99-306 (cext_align);
 * twin: praline/util/align.py:41-141.  `mode` only matters for LOCAL
 * (cext.c:208-212): the M state is floored at 0.
 * ---------------------------------------------------------------------- */
void orc_fill(int mode, const float *m, const float *g1, const float *g2,
              int L1, int L2, float *o, uint8_t *t, const uint8_t *z)
{
    const size_t W = (size_t)L2 + 1;
    for (int y = 1; y <= L1; y++) {                                   /* cext.c:133 */
        for (int x = 1; x <= L2; x++) {                               /* :136 */
            if (z && z[(size_t)y * W + x]) continue;                  /* :141-149 */

            const float up_open = O(y - 1, x, 0) + g1[(size_t)(y - 1) * 2 + 0]; /* :160-162 */
            const float up_ext  = O(y - 1, x, 1) + g1[(size_t)(y - 1) * 2 + 1]; /* :163-165 */
            const float lf_open = O(y, x - 1, 0) + g2[(size_t)(x - 1) * 2 + 0]; /* :177-179 */
            const float lf_ext  = O(y, x - 1, 2) + g2[(size_t)(x - 1) * 2 + 1]; /* :180-182 */
            const float ms = m[(size_t)(y - 1) * L2 + (x - 1)];                 /* :189-190 */
            const float mm = O(y - 1, x - 1, 0) + ms;                           /* :192-194 */
            const float mu = O(y - 1, x - 1, 1) + ms;                           /* :195-197 */
            const float ml = O(y - 1, x - 1, 2) + ms;                           /* :198-200 */

            float best = (mode == MODE_LOCAL) ? 0.0f : -INFINITY;               /* :207-212 */
            if (mm > best) best = mm;                                           /* :214-222 */
            if (mu > best) best = mu;
            if (ml > best) best = ml;
            uint8_t f = 0;                                                      /* :224-233 */
            if (mm == best) f |= TB_MM;
            if (mu == best) f |= TB_MU;
            if (ml == best) f |= TB_ML;
            T(y, x, 0) = f;
            O(y, x, 0) = best;

            float ub = -INFINITY;                                               /* :247-262 */
            if (up_open > ub) ub = up_open;
            if (up_ext > ub) ub = up_ext;
            f = 0;
            if (up_open == ub) f |= TB_UO;
            if (up_ext == ub) f |= TB_UE;
            T(y, x, 1) = f;
            O(y, x, 1) = ub;

            float lb = -INFINITY;                                               /* :276-291 */
            if (lf_open > lb) lb = lf_open;
            if (lf_ext > lb) lb = lf_ext;
            f = 0;
            if (lf_open == lb) f |= TB_LO;
            if (lf_ext == lb) f |= TB_LE;
            T(y, x, 2) = f;
            O(y, x, 2) = lb;
        }
    }
}

/* ------------------------------------------------------------------------
 * End-cell selection.  praline/component/align.py:401-431.
 *   local      : first argmax of the whole o array in (y, x, state) order (:402)
 *   semiglobal : max(last row) vs max(last column); strict '>' and only when
 *                tracing from the row is allowed (both / two); scan from the far
 *                end backwards, states 0,1,2, first equal (:405-422)
 *   global     : first argmax over the three states at (L1, L2) (:427-430)
 * ---------------------------------------------------------------------- */
void orc_end_cell(int mode, const float *o, int L1, int L2, int *cy, int *cx, int *ck)
{
    const size_t W = (size_t)L2 + 1;
    if (mode == MODE_LOCAL) {
        size_t n = ((size_t)L1 + 1) * W * 3, best = 0;
        for (size_t i = 1; i < n; i++) if (o[i] > o[best]) best = i;
        *ck = (int)(best % 3); *cx = (int)((best / 3) % W); *cy = (int)(best / 3 / W);
    } else if (mode == MODE_GLOBAL) {
        int k = 0;
        for (int j = 1; j < 3; j++) if (O(L1, L2, j) > O(L1, L2, k)) k = j;
        *cy = L1; *cx = L2; *ck = k;
    } else {
        float row_max = -INFINITY, col_max = -INFINITY;
        for (int x = 0; x <= L2; x++) for (int k = 0; k < 3; k++) if (O(L1, x, k) > row_max) row_max = O(L1, x, k);
        for (int y = 0; y <= L1; y++) for (int k = 0; k < 3; k++) if (O(y, L2, k) > col_max) col_max = O(y, L2, k);
        const int from_row = (mode == MODE_SG_BOTH || mode == MODE_SG_TWO);
        if (row_max > col_max && from_row) {
            for (int x = L2; x >= 0; x--) for (int k = 0; k < 3; k++)
                if (O(L1, x, k) == row_max) { *cy = L1; *cx = x; *ck = k; return; }
        } else {
            for (int y = L1; y >= 0; y--) for (int k = 0; k < 3; k++)
                if (O(y, L2, k) == col_max) { *cy = y; *cx = L2; *ck = k; return; }
        }
    }
}

/* ------------------------------------------------------------------------
 * Traceback walk.  praline/util/align.py:144-185 (get_paths).
 * Priority MM > MU > ML > U-open > U-extend > L-open > L-extend (:161-174);
 * stops at the first cell with none of them set (:175-176); path is reversed
 * and the state stripped.  path_out holds (y, x) pairs, capacity L1+L2+2 rows.
 * Returns the number of rows.
 * ---------------------------------------------------------------------- */
int orc_get_path(const uint8_t *t, int L1, int L2, int cy, int cx, int ck, int32_t *path_out)
{
    const size_t W = (size_t)L2 + 1;
    int cap = L1 + L2 + 2, n = 0;
    int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)cap);
    int y = cy, x = cx, k = ck;
    tmp[0] = y; tmp[1] = x; n = 1;
    for (;;) {
        const uint8_t v = T(y, x, k);
        if (v & TB_MM)      { y--; x--; k = 0; }
        else if (v & TB_MU) { y--; x--; k = 1; }
        else if (v & TB_ML) { y--; x--; k = 2; }
        else if (v & TB_UO) { y--; k = 0; }
        else if (v & TB_UE) { y--; k = 1; }
        else if (v & TB_LO) { x--; k = 0; }
        else if (v & TB_LE) { x--; k = 2; }
        else break;
        if (n >= cap) break; /* cannot happen: every move decreases y + x */
        tmp[2 * n] = y; tmp[2 * n + 1] = x; n++;
    }
    for (int i = 0; i < n; i++) {
        path_out[2 * i] = tmp[2 * (n - 1 - i)];
        path_out[2 * i + 1] = tmp[2 * (n - 1 - i) + 1];
    }
    free(tmp);
    return n;
}

/* ------------------------------------------------------------------------
 * Semiglobal end-gap padding.  praline/util/align.py:268-297.
 * In place on path (capacity L1+L2+2 rows); returns the new row count.
 * The if/elif order of the reference is kept: rows first, then columns.
 * ---------------------------------------------------------------------- */
int orc_extend_semiglobal(int32_t *path, int n, int L1, int L2)
{
    int pre = 0, pre_col = 0;
    if (path[0] != 0) { pre = path[0]; pre_col = 0; }            /* :270-274 */
    else if (path[1] != 0) { pre = path[1]; pre_col = 1; }       /* :275-279 */
    if (pre) {
        memmove(path + 2 * pre, path, sizeof(int32_t) * 2 * (size_t)n);
        for (int i = 0; i < pre; i++) { path[2 * i] = 0; path[2 * i + 1] = 0; path[2 * i + pre_col] = i; }
        n += pre;
    }
    const int ly = path[2 * (n - 1)], lx = path[2 * (n - 1) + 1];
    if (ly != L1) {                                              /* :284-289 */
        for (int v = ly + 1; v <= L1; v++) { path[2 * n] = v; path[2 * n + 1] = lx; n++; }
    } else if (lx != L2) {                                       /* :290-295 */
        for (int v = lx + 1; v <= L2; v++) { path[2 * n] = ly; path[2 * n + 1] = v; n++; }
    }
    return n;
}

/* ------------------------------------------------------------------------
 * One RawPairwiseAligner call.  praline/component/align.py:302-447.
 * zero_idx: n_zero (y, x) pairs to mask (align.py:361-363), may be NULL.
 * o_out/t_out may be NULL (then scratch is allocated and freed here).
 * Returns the path length (rows), writes *score.
 * ---------------------------------------------------------------------- */
int orc_align_raw(int mode, const float *m, const float *g1, const float *g2, int L1, int L2,
                  const int32_t *zero_idx, int n_zero, float *score, int32_t *path_out,
                  float *o_out, uint8_t *t_out)
{
    const size_t W = (size_t)L2 + 1, cells = ((size_t)L1 + 1) * W;
    float *o = o_out ? o_out : (float *)malloc(cells * 3 * sizeof(float));
    uint8_t *t = t_out ? t_out : (uint8_t *)malloc(cells * 3);
    uint8_t *z = NULL;
    memset(o, 0, cells * 3 * sizeof(float));
    memset(t, 0, cells * 3);
    if (zero_idx && n_zero > 0) {
        z = (uint8_t *)calloc(cells, 1);
        for (int i = 0; i < n_zero; i++) z[(size_t)zero_idx[2 * i] * W + zero_idx[2 * i + 1]] = 1;
    }
    orc_init_borders(mode, g1, g2, L1, L2, o, t);
    orc_fill(mode, m, g1, g2, L1, L2, o, t, z);
    int cy, cx, ck;
    orc_end_cell(mode, o, L1, L2, &cy, &cx, &ck);
    *score = O(cy, cx, ck);
    int n = orc_get_path(t, L1, L2, cy, cx, ck, path_out);
    if (mode == MODE_SG_BOTH || mode == MODE_SG_ONE || mode == MODE_SG_TWO)
        n = orc_extend_semiglobal(path_out, n, L1, L2);            /* align.py:425-426 */
    if (!o_out) free(o);
    if (!t_out) free(t);
    free(z);
    return n;
}

/* ------------------------------------------------------------------------
 * One PairwiseAligner call on two index sequences (PlainTrack x PlainTrack,
 * one track set).  praline/component/align.py:163-221: one-hot profiles make
 * cext_build_scores return exactly S[a_y][b_x]; gap arrays are constant.
 * gap_open/gap_extend are the (negative) series values; linear gaps pass the
 * same value twice (align.py:182-183).
 * ---------------------------------------------------------------------- */
int orc_align_seqs(int mode, const int32_t *a, int L1, const int32_t *b, int L2,
                   const float *S, int A, float gap_open, float gap_extend,
                   float *score, int32_t *path_out)
{
    float *m = (float *)malloc(sizeof(float) * (size_t)(L1 > 0 ? L1 : 1) * (size_t)(L2 > 0 ? L2 : 1));
    float *g1 = (float *)malloc(sizeof(float) * 2 * (size_t)(L1 > 0 ? L1 : 1));
    float *g2 = (float *)malloc(sizeof(float) * 2 * (size_t)(L2 > 0 ? L2 : 1));
    for (int y = 0; y < L1; y++) for (int x = 0; x < L2; x++) m[(size_t)y * L2 + x] = S[(size_t)a[y] * A + b[x]];
    for (int y = 0; y < L1; y++) { g1[2 * y] = gap_open; g1[2 * y + 1] = gap_extend; }
    for (int x = 0; x < L2; x++) { g2[2 * x] = gap_open; g2[2 * x + 1] = gap_extend; }
    int n = orc_align_raw(mode, m, g1, g2, L1, L2, NULL, 0, score, path_out, NULL, NULL);
    free(m); free(g1); free(g2);
    return n;
}

/* Batch of sequence pairs, score (+ optional path) per pair; the scalar CPU
 * baseline that bench.py times ("port").  seqs is a flat int32 buffer, offs has
 * n_seqs+1 entries.  paths may be NULL; else path p starts at paths + 2*path_offs[p]. */
void orc_align_batch(int mode, int64_t n_pairs, const int32_t *seqs, const int64_t *offs,
                     const int32_t *pi, const int32_t *pj, const float *S, int A,
                     float gap_open, float gap_extend, float *scores,
                     int32_t *paths, const int64_t *path_offs, int32_t *path_len)
{
    for (int64_t p = 0; p < n_pairs; p++) {
        const int32_t *a = seqs + offs[pi[p]], *b = seqs + offs[pj[p]];
        const int L1 = (int)(offs[pi[p] + 1] - offs[pi[p]]), L2 = (int)(offs[pj[p] + 1] - offs[pj[p]]);
        int32_t *tmp = (int32_t *)malloc(sizeof(int32_t) * 2 * (size_t)(L1 + L2 + 2));
        int n = orc_align_seqs(mode, a, L1, b, L2, S, A, gap_open, gap_extend, &scores[p], tmp);
        if (paths) { memcpy(paths + 2 * path_offs[p], tmp, sizeof(int32_t) * 2 * (size_t)n); path_len[p] = n; }
        free(tmp);
    }
}
