/*
 * praline_b200.h -- C ABI of libpraline_b200.so: the B200 (sm_100a) drop-in for PRALINE's
 * pairwise dynamic-programming alignment core.
 *
 * What it replaces (paths relative to the reference tree, ibivu/PRALINE):
 *   praline/util/cext.c:506-520      the six functions of the `praline.util.cext` CPython module
 *   praline/component/align.py:357-431  RawPairwiseAligner's border init and end-cell choice
 *   praline/util/align.py:144-185, 268-297  get_paths and extend_path_semiglobal
 * The reference binds its native code through the CPython API; this library is bound with
 * ctypes (see INTEGRATION.md for the stub a PRALINE maintainer adds).
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; pgpu_last_error() then holds
 *     a message (thread local).  1 = bad argument, 2 = CUDA runtime error, 3 = no usable GPU.
 *   - pointers named *_dev are DEVICE pointers owned by the caller (PyTorch tensors'
 *     data_ptr() in the Python host layer); nothing is allocated behind the caller's back
 *     except in pgpu_fill_debug, which takes HOST buffers like the reference's functions do.
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued asynchronously and the
 *     caller synchronises.
 *   - modes are numbered as in praline/util/cext.c:27-31.
 *   - sequences are uint8 symbol indices (PRALINE alphabets have <= 27 symbols,
 *     praline/container/alphabet.py:96-109), concatenated, with int64 offsets [n+1].
 */
#ifndef PRALINE_B200_H
#define PRALINE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGPU_ABI_VERSION 1

#define PGPU_MODE_GLOBAL 0
#define PGPU_MODE_LOCAL 1
#define PGPU_MODE_SEMIGLOBAL_BOTH 2
#define PGPU_MODE_SEMIGLOBAL_ONE 3
#define PGPU_MODE_SEMIGLOBAL_TWO 4

/* One unit of work of the inter-task kernel: a resident sequence against a run of streamed
 * sequences; results go to consecutive output slots starting at out_base. */
typedef struct pgpu_tile {
    int32_t resident;
    int32_t stream_begin;
    int32_t stream_end;
    int32_t resident2;      /* paired-resident int16 launches: second resident or -1; else unused */
    int64_t out_base;
    int64_t out_base2;      /* paired-resident launches: first output slot of resident2 */
    int32_t b_skip;         /* leading stream elements that have no pair with resident2 */
    int32_t reserved;
} pgpu_tile;

int pgpu_abi_version(void);
int pgpu_init(int device);
void pgpu_shutdown(void);
const char* pgpu_last_error(void);

/* Columns-per-lane values the inter-task kernel is instantiated for (a resident sequence of
 * length L needs K >= ceil(L/32)), and the number of warps that share one tile. */
int pgpu_supported_k(int k);
int pgpu_warps_per_tile(void);

/*
 * The two self-contained entry points.  Everything the tile-level functions below
 * expect from their caller -- grouping pairs by resident sequence, tiles, traceback waves, border
 * arrays, workspaces -- is done inside the library (C++ host code, stream-ordered temporaries).
 *
 * pgpu_align_batch: n_pairs sequence-sequence alignments, pair k = (sequence_one = pair_i[k],
 * sequence_two = pair_j[k]), one PairwiseAligner.execute each in the reference (component/
 * align.py:88-251: cext_build_scores + border init + cext_align_<mode> + end cell + get_paths +
 * extend_path_semiglobal).  All five modes; linear gaps: gap_extend = gap_open.  scores_out [n_pairs].
 * want_paths: integer-valued scores only (else code 4: use pgpu_align_profiles pair by pair);
 * path_buf is [sum_k (L1_k + L2_k + 2)][2] int32, the region of pair k starts at the prefix sum of
 * those capacities and its path ((y, x) rows, reference orientation) occupies the LAST
 * path_len_out[k] rows of the region.  Sequence two must be <= 1024 long (code 4 otherwise).
 * Synchronises the stream once while planning (pair ids and offsets are read back).
 *
 * pgpu_align_profiles: one profile-profile alignment over n_sets track sets, m = sum_sets
 * P1 . S . P2^T in the reference's evaluation order (cext.c:308-455) followed by
 * RawPairwiseAligner (component/align.py:302-447) with per-position gap arrays g1 [L1][2],
 * g2 [L2][2] and an optional mask zmask [(L1+1)][(L2+1)] (zero_idxs).  P1/P2/S are HOST arrays of
 * DEVICE pointers.  score_out [1]; path_out [L1+L2+2][2] (may be NULL together with path_len_out),
 * the path occupies the LAST path_len_out[0] rows.  Fully asynchronous on the stream.
 */
int pgpu_align_batch(int mode, int64_t n_pairs, const uint8_t* seqs_dev, const int64_t* seq_offsets_dev,
                     const int32_t* pair_i_dev, const int32_t* pair_j_dev, const float* S_dev, int A,
                     float gap_open, float gap_extend, int want_paths, float* scores_out_dev,
                     int32_t* path_buf_dev, int32_t* path_len_out_dev, void* stream);
int pgpu_align_profiles(int mode, int n_sets, const float* const* P1_dev, const float* const* P2_dev,
                        const float* const* S_dev, const int* A, int L1, int L2, const float* g1_dev,
                        const float* g2_dev, const uint8_t* zmask_dev_or_null, float* score_out_dev,
                        int32_t* path_out_dev, int32_t* path_len_out_dev, void* stream);

/*
 * Inter-task batch (K2): sequence-sequence alignments with constant gap penalties.
 * Replaces, per pair, cext_build_scores + cext_align_<mode> + the end-cell choice
 * (cext.c:308-455, :99-306; component/align.py:401-431), i.e. one PairwiseAligner.execute
 * (component/align.py:88-251) for PlainTrack inputs.
 *
 *   transposed   0: resident sequences are `sequence_two` (DP columns, as in the reference)
 *                1: resident sequences are `sequence_one`
 *   stream_ids   sequence id per stream element, or NULL when element s is sequence s
 *   topD/leftD   DP border values max(M,U,L) along row 0 / column 0 in the KERNEL orientation
 *                (component/align.py:367-385), border_len >= 32*K+1 and > every stream length;
 *                left0/left1: the same column-0 border in closed form, D(y,0) = left0 + (y-1)*left1
 *   scores       [n_slots] f32
 *   keys         [2*n_slots] u64 scratch, local and semiglobal modes (running maxima); only
 *                slots produced by this call have their score written
 *   tb .. pair_tb  NULL for score-only runs; else packed traceback words, per-(tile,warp) word
 *                offsets, and per-slot outputs consumed by pgpu_traceback_tiles
 *   mwave, mrow_base  NULL for sequence batches.  Profile x profile batches (score only): the
 *                match scores of the wave made by pgpu_build_rows, one row of 32*K floats per
 *                stream position, and the first row per (tile, warp); seqs_dev is then unused and
 *                offs_dev holds profile ROW offsets per sequence;
 *                the buffer must be 16-byte aligned and carry 16 bytes of slack behind its last row (the
 *                kernel reads whole aligned float4s around a lane's K scores)
 */
int pgpu_align_tiles(int mode, int K, int transposed, const uint8_t* seqs_dev, const int64_t* offs_dev,
                     const int32_t* stream_ids_dev, const void* tiles_dev, int n_tiles, int64_t n_slots,
                     const float* S_dev, int A, float gap_open, float gap_extend, const float* topD_dev,
                     const float* leftD_dev, float left0, float left1, int border_len, float* scores_dev,
                     uint64_t* keys_dev,
                     uint32_t* tb_dev, const int64_t* tb_base_dev, int32_t* emit_t_dev,
                     int64_t* pair_tb_dev, const float* mwave_dev, const int64_t* mrow_base_dev, void* stream);

/*
 * Inter-task batch in packed 16-bit integers (DPX, two streamed sequences per warp): global
 * mode, score per pair, INTEGER matrix and gaps whose DP values fit int16 (the caller checks,
 * Engine.fits_s16).  Same tiles, sequences and outputs as pgpu_align_tiles; `neg` is the -inf
 * sentinel, left0/left1 the column-0 border D(y,0) = left0 + (y-1)*left1 as integers.
 * paired = 0: the two halves carry two STREAMED sequences of the tile (any pair list).
 * paired = 1: the two halves carry two RESIDENT sequences (tile.resident, tile.resident2) that
 *             share one stream -- the all-vs-all shape; results of resident2 go to out_base2.
 */
int pgpu_align_tiles16(int K, int paired, int transposed, const uint8_t* seqs_dev, const int64_t* offs_dev,
                       const int32_t* stream_ids_dev, const void* tiles_dev, int n_tiles, const float* S_dev,
                       int A, int gap_open, int gap_extend, int neg, const float* topD_dev, int left0,
                       int left1, int border_len, float* scores_dev, void* stream);

/* Traced form of pgpu_align_tiles16 (two streamed sequences per warp, any pair list): global mode,
 * integer scores, every DP value within +-16000.  Writes the score per slot and the packed
 * traceback words (4 rows x 2 register halves per 32-bit word, 0.5 B per cell), to be walked by
 * pgpu_traceback_tiles with tb_fmt = 1 (tb_fmt = 0 is the f32 kernel's layout); emit_t carries the
 * register half of a slot in bit 30.  Replaces the same reference chain as pgpu_align_tiles with
 * traceback (cext.c:308-455, :99-306; component/align.py:357-431). */
int pgpu_align_tiles16_traced(int K, int transposed, const uint8_t* seqs_dev, const int64_t* offs_dev,
                              const int32_t* stream_ids_dev, const void* tiles_dev, int n_tiles, const float* S_dev,
                              int A, int gap_open, int gap_extend, int neg, const float* topD_dev, int left0, int left1,
                              int border_len, float* scores_dev, uint32_t* tb_dev, const int64_t* tb_base_dev,
                              int32_t* emit_t_dev, int64_t* pair_tb_dev, void* stream);

/*
 * Paired-resident traced fill: tiles as for pgpu_align_tiles16 with paired = 1 (residents (i, i+1) in the two
 * register halves, one shared stream); every slot is an UNORDERED pair filled ONCE.  With a symmetric
 * substitution matrix and constant gap penalties the DP values of (sequence_one = s, sequence_two = r) are
 * the transpose of those of (r, s) (cext.c:155-200), so one nibble per cell serves the walks of both
 * alignments (tb_fmt 2 of the walk: pgpu_traceback_dual).  The kernel records every slot's (resident,
 * streamed) sequence ids.  Replaces the two PairwiseAligner executions per unordered pair of a global
 * preprofile (component/preprofile.py:127-154, component/align.py:357-431, cext.c:99-306).  The caller
 * checks the symmetry of S and the +-16000 value range. */
int pgpu_align_tiles16_paired_traced(int K, const uint8_t* seqs_dev, const int64_t* offs_dev, const void* tiles_dev,
                                     int n_tiles, const float* S_dev, int A, int gap_open, int gap_extend, int neg,
                                     const float* topD_dev, int left0, int left1, int border_len, float* scores_dev,
                                     uint32_t* tb_dev, const int64_t* tb_base_dev, int32_t* emit_t_dev,
                                     int64_t* pair_tb_dev, int32_t* slot_res_dev, int32_t* slot_str_dev, void* stream);

/*
 * Both walks of every slot of pgpu_align_tiles16_paired_traced: walk 2k takes slot k with the resident as
 * sequence one, walk 2k + 1 with the streamed sequence as sequence one (get_paths, util/align.py:144-185,
 * tie order :161-174).  counts + seq_cnt_off (offset of every sequence's [L x A] table, < 0: not a master):
 * preprofile mode as in pgpu_traceback_tiles; path_* are indexed by walk (2 * slot + orientation). */
int pgpu_traceback_dual(int K, const int64_t* offs_dev, const int32_t* slot_res_dev, const int32_t* slot_str_dev,
                        int64_t n_slots, const uint32_t* tb_dev, const int32_t* emit_t_dev, const int64_t* pair_tb_dev,
                        int code00, int top_ramp, int left_ramp, const uint8_t* seqs_dev, int32_t* counts_dev,
                        const int64_t* seq_cnt_off_dev, int A, const float* scores_dev, int use_thr, float thr,
                        const int64_t* path_off_dev, int32_t* path_buf_dev, int32_t* path_start_dev,
                        int32_t* path_len_dev, void* stream);

/*
 * Traceback of an inter-task batch (K4).  Replaces get_paths (util/align.py:144-185), the
 * semiglobal end-cell scan (component/align.py:405-426) and extend_path_semiglobal
 * (util/align.py:268-297).  Paths are (y, x) rows in the reference orientation, written
 * right-aligned into each slot's region of capacity len(resident)+len(stream)+2 rows:
 * rows [path_off[s] + path_start[s], +path_len[s]).  path_buf may be NULL when counts is given.
 *
 * Preprofile mode (counts_dev != NULL, global mode): every pair is (master = sequence one,
 * slave = sequence two); the walk adds the slave's aligned residues into the master's count
 * table counts_dev + cnt_off[slot] ([len(master)][A] int32), which is what compress_path +
 * Alignment.merge + get_frequencies produce on the host in the reference (util/align.py:187-232,
 * container/align.py:30-61; preprofile.py:127-154).  Pairs with score < threshold are skipped.
 */
int pgpu_traceback_tiles(int mode, int K, int transposed, const int64_t* offs_dev,
                         const int32_t* slot_resident_dev, const int32_t* slot_stream_dev, int64_t n_slots,
                         const uint64_t* keys_dev, const uint32_t* tb_dev, const int32_t* emit_t_dev,
                         const int64_t* pair_tb_dev, int code00, int top_ramp, int left_ramp,
                         const int64_t* path_off_dev, int32_t* path_buf_dev, int32_t* path_start_dev,
                         int32_t* path_len_dev, const uint8_t* seqs_dev, int32_t* counts_dev,
                         const int64_t* cnt_off_dev, int A, const float* scores_dev, int use_threshold,
                         float threshold, int tb_fmt, void* stream);

/*
 * Local traced batch and its walk: the inner loop of LocalMasterSlaveAligner (component/
 * preprofile.py:227-267) -- PairwiseAligner in mode "local" (cext.c:203-243 with MODE_LOCAL, end
 * cell = first row-major maximum, component/align.py:401-403) with zero_idxs = the bounding boxes of
 * the alignments found in earlier Waterman-Eggert iterations (preprofile.py:252-259; masked cells
 * keep M = U = L = 0 and no flags, cext.c:143-148).  Reference orientation only (resident =
 * sequence two), gap penalties <= 0, integer-valued scores.
 *
 *   boxes      NULL, or [n_slots][PGPU_NBOX][4] int32 (ylo, yhi, xlo, xhi), inclusive, y over sequence
 *              one; (1, 0, 1, 0) = empty
 *   box_out    [n_slots][PGPU_NBOX][4]: the walk writes the bounding box of its path into box number
 *              box_slot (may alias boxes: iteration n reads boxes 0..n-2 and writes box n-1)
 *   counts     local preprofile mode: compress_path + extend_path_local + Alignment.merge +
 *              get_frequencies (util/align.py:187-266) as atomic adds into the master's table
 * Paths are NOT extended (the reference extends after compress_path, preprofile.py:262-263).
 */
#define PGPU_NBOX 3
int pgpu_align_tiles_local(int K, const uint8_t* seqs_dev, const int64_t* offs_dev, const int32_t* stream_ids_dev,
                           const void* tiles_dev, int n_tiles, int64_t n_slots, const float* S_dev, int A,
                           float gap_open, float gap_extend, const float* topD_dev, float left0, float left1,
                           int border_len, float* scores_dev, uint64_t* keys_dev, uint32_t* tb_dev,
                           const int64_t* tb_base_dev, int32_t* emit_t_dev, int64_t* pair_tb_dev,
                           const int32_t* boxes_dev, void* stream);
int pgpu_traceback_tiles_local(int K, const int64_t* offs_dev, const int32_t* slot_resident_dev,
                               const int32_t* slot_stream_dev, int64_t n_slots, const uint64_t* keys_dev,
                               const uint32_t* tb_dev, const int32_t* emit_t_dev, const int64_t* pair_tb_dev,
                               int code00, const int64_t* path_off_dev, int32_t* path_buf_dev,
                               int32_t* path_start_dev, int32_t* path_len_dev, const uint8_t* seqs_dev,
                               int32_t* counts_dev, const int64_t* cnt_off_dev, int A, const float* scores_dev,
                               int use_threshold, float threshold, const int32_t* boxes_dev, int32_t* box_out_dev,
                               int box_slot, void* stream);

/*
 * Match-score matrix (K1).  Replaces cext_build_scores (cext.c:308-455): m[y][x] = sum over
 * track sets of P1[y] . S . P2[x]^T, evaluated in the reference's order.  P1/P2/S are HOST
 * arrays of DEVICE pointers (one per set), A the alphabet size per set.
 */
int pgpu_build_scores(int n_sets, const float* const* P1_dev, const float* const* P2_dev,
                      const float* const* S_dev, const int* A, int L1, int L2, float* m_dev, int m_pitch,
                      void* stream);
/* <= 32 consecutive matrix rows of one streamed sequence in a wave of a profile batch */
typedef struct pgpu_row_block {
    int64_t row0;     /* first matrix row */
    int64_t src0;     /* profile row feeding matrix row row0 */
    int32_t rows;
    int32_t res;      /* resident sequence id */
    int32_t dummy;    /* row0 is the region's dummy row (no profile row) */
    int32_t reserved;
} pgpu_row_block;

/*
 * Match scores of a wave of profile x profile pairs in stream order (feeds pgpu_align_tiles).
 * Same evaluation order as cext_build_scores (cext.c:63-95) for ONE track set.  prof [rows][A]
 * holds all profiles, rowoff [n_seqs+1] their row offsets; the wave is described by row blocks.
 * dense_syms: 0, or -- for batches of DENSE profiles with sequence one resident (transposed = 1) -- the number of
 * alphabet symbols that have a nonzero entry anywhere in prof: selects the packed f32x2 kernel (score_rows_x2.cu;
 * same bits, the symbol count sizes its tables; a count that is too small gives NaN rows, not wrong numbers).
 */
int pgpu_build_rows(const float* prof_dev, const int64_t* rowoff_dev, int A, const float* S_dev,
                    const void* blocks_dev, int n_blocks, int width, int transposed, int local_mode,
                    int dense_syms, float* mwave_dev, void* stream);
/*
 * Tolerance-mode (<= 1e-5 relative, not the reference's evaluation order) variant for score-only
 * profile batches: W = P . S^T (transposed = 0) or P . S (transposed = 1) per profile row from
 * pgpu_profile_times_matrix, then A fused multiply-adds per cell.
 */
int pgpu_profile_times_matrix(const float* prof_dev, const float* S_dev, int A, int64_t n_rows, int transposed,
                              float* out_dev, void* stream);
int pgpu_build_rows_fast(const float* prof_dev, const float* wres_dev, const int64_t* rowoff_dev, int A,
                         const void* blocks_dev, int n_blocks, int width, int local_mode, float* mwave_dev,
                         void* stream);
/* Tensor-core form of pgpu_build_rows_fast (tcgen05.mma kind::tf32 with an FP32-accurate hi/lo split,
 * accumulators in TMEM): same output.  One pgpu_quad per 128-row tile: <= 4 row blocks (<= 32 matrix rows
 * each) that share a resident, with the resident's first profile row and length.  A <= 32, width % 32 == 0. */
typedef struct pgpu_quad {
    int64_t q0;          /* first profile row of the resident */
    int32_t Lr;          /* resident length */
    int32_t nblk;        /* row blocks in this tile (1..4) */
    int64_t row0[4];     /* first matrix row per block */
    int64_t src0[4];     /* profile row feeding matrix row row0 (row0 + r is fed by src0 + r) */
    int32_t rows[4];
    int32_t dummy[4];    /* matrix row row0 is the region's dummy row */
    int64_t bcan;        /* first row of the resident in the pre-split store of pgpu_split_residents */
    int64_t reserved;
    int64_t can0[4];     /* per block: its first row in the pre-split store of the streamed side if the block
                            starts on an 8-row group there (rows src0 .. src0+31 = 4096 contiguous bytes), else -1 */
} pgpu_quad;
/* whi_dev / wlo_dev: NULL, or the resident side (W) pre-split by pgpu_split_residents -- the kernel then fetches its
 * B tiles with TMA bulk copies (cp.async.bulk) instead of gathering rows.  phi_dev / plo_dev: NULL, or the streamed
 * side (the profiles themselves) pre-split with the same row numbering, 32 rows of slack behind the last one: row
 * blocks with can0 >= 0 then arrive by TMA as well. */
int pgpu_build_rows_tc(const float* prof_dev, const float* wres_dev, int A, const void* quads_dev, int n_quads,
                       int width, int local_mode, float* mwave_dev, const void* whi_dev, const void* wlo_dev,
                       const void* phi_dev, const void* plo_dev, void* stream);
/* W rows [rows x A] -> tf32 hi / lo parts in the canonical K-major core-matrix layout (128 B per row, 8-row groups
 * of 1024 B); sequence s occupies rows padoff[s] .. padoff[s+1] (multiples of 32, zero rows beyond its length).
 * whi_dev / wlo_dev: padoff[n_seqs] * 128 bytes each. */
int pgpu_split_residents(const float* wres_dev, const int64_t* rowoff_dev, const int64_t* padoff_dev, int n_seqs, int A,
                         void* whi_dev, void* wlo_dev, void* stream);
/*
 * HOST-side planner of one wave of a profile batch (no device work; what Engine.align_profile_pairs feeds the
 * score-row kernels above).  Tiles [tile_begin, tile_end) are runs of stream elements, contiguous in stream order;
 * each is split over nw warps; every (tile, warp) region holds one dummy row followed by the profile rows of its
 * streamed sequences.  cs: prefix sums of the streamed lengths lens_s; str_s / res_s: streamed / resident sequence
 * per element; offs: profile row offsets per sequence.  Writes the first matrix row per region (mrow_base
 * [n_tiles * nw]), the row blocks of <= rows_per_block rows, the number of matrix rows and -- want_quads -- the
 * 128-row quads of pgpu_build_rows_tc (padoff: NULL, or the pre-split row offsets of pgpu_split_residents).
 * Returns the number of row blocks, -1 on error (capacity too small: pgpu_last_error).
 */
long long pgpu_plan_profile_wave(int n_tiles, const int64_t* tile_begin, const int64_t* tile_end, int nw,
                                 const int64_t* cs, const int64_t* lens_s, const int64_t* str_s, const int64_t* res_s,
                                 const int64_t* offs, int rows_per_block, int64_t* mrow_base, pgpu_row_block* blocks,
                                 long long blocks_cap, int64_t* n_rows_out, int want_quads, const int64_t* padoff,
                                 struct pgpu_quad* quads, long long quads_cap, long long* n_quads_out);
int pgpu_build_scores_seq(const uint8_t* a_dev, const uint8_t* b_dev, const float* S_dev, int A, int L1,
                          int L2, float* m_dev, int m_pitch, void* stream);

/*
 * Progressive-merge glue on device-resident count tables (the loop of
 * TreeMultipleSequenceAligner, component/msa.py:124-237).
 * pgpu_counts_to_profile: ProfileTrack.profile (container/sequence.py:191-203) -- f32 totals, float64 division,
 * f32 result -- for n_rows positions of [n_rows][A] int32 counts.
 * pgpu_merge_counts: ProfileTrack.merge (container/sequence.py:205-239) along the path that pgpu_align_general
 * left on the device: align_out_dev is that call's contiguous int32 output block laid out as
 * [score, cell y, x, state, path_start, path_len, 0, 0, path rows (y, x) ...]; merged_dev receives path_len - 1 rows
 * (at most max_rows) of A counts.
 */
int pgpu_counts_to_profile(const int32_t* counts_dev, int64_t n_rows, int A, float* prof_dev, void* stream);
int pgpu_merge_counts(const int32_t* counts_one_dev, const int32_t* counts_two_dev, int A, const int32_t* align_out_dev,
                      int32_t* merged_dev, int max_rows, void* stream);

/*
 * General single alignment (K3 wavefront + traceback).  Replaces RawPairwiseAligner.execute
 * (component/align.py:302-447): arbitrary match scores m [L1][m_pitch], per-position gap
 * arrays g1 [L1][2], g2 [L2][2] (var_gaps = 0 promises that every row of g1 equals g1[0] and
 * every row of g2 equals g2[0], which is what PairwiseAligner builds, component/align.py:212-217),
 * optional mask z [(L1+1)][z_pitch] (zero_idxs), any mode.  A match-score matrix whose pitch is a
 * multiple of 4 floats and covers ceil(L2/128)*128 columns runs the lean wavefront kernel.
 * Outputs: *score_out, cell_out[3] = end cell (y, x, state), path rows right-aligned in
 * path_buf (capacity L1+L2+2 rows) at [path_start[0], +path_len[0]).  path_buf may be NULL
 * (score only).  o_full / t_full, when non-NULL, receive the reference's complete o and t
 * arrays [(L1+1)][(L2+1)][3] (debug level > 1, component/align.py:390-399).
 */
int64_t pgpu_general_workspace_bytes(int L1, int L2);
int pgpu_align_general(int mode, int L1, int L2, const float* m_dev, int m_pitch, const float* g1_dev,
                       const float* g2_dev, int var_gaps, const uint8_t* z_dev, int z_pitch, void* workspace_dev,
                       float* score_out_dev, int32_t* cell_out_dev, int32_t* path_buf_dev,
                       int32_t* path_start_dev, int32_t* path_len_dev, float* o_full_dev,
                       uint8_t* t_full_dev, void* stream);
/*
 * One LONG profile x profile alignment in one call (BASELINE config 5: 20 kb x 20 kb): cext_build_scores for one
 * track set (cext.c:308-455) into m_dev [L1][m_pitch] + pgpu_align_general on it (no mask, no debug arrays).  When
 * the row-blocked wavefront serves the fill (global / semiglobal modes, var_gaps = 0, padded pitch) and the matrix
 * has 2^21 cells or more, the score matrix is built BESIDE the fill: K1 publishes a flag per finished 128 x 128
 * block, the fill (on a high-priority stream inside the library, joined to `stream` before the call returns to
 * stream order) acquires a block's flag before it reads rows of it.  Same results as the two calls in sequence;
 * workspace_dev as for pgpu_align_general.  PGPU_NO_K1_OVERLAP=1 runs the two in sequence.
 */
int pgpu_align_profile_long(int mode, const float* P1_dev, const float* P2_dev, const float* S_dev, int A, int L1, int L2,
                            float* m_dev, int m_pitch, const float* g1_dev, const float* g2_dev, int var_gaps,
                            void* workspace_dev, float* score_out_dev, int32_t* cell_out_dev, int32_t* path_buf_dev,
                            int32_t* path_start_dev, int32_t* path_len_dev, void* stream);

/*
 * Parity shim with the exact shape of the reference's cext_align_<mode>(m, g1, g2, o, t, z)
 * (cext.c:103-107): HOST buffers, dense row-major, o [(L1+1)][(L2+1)][3] f32 and t u8 are
 * overwritten completely (borders included), z may be NULL.  Synchronous.
 */
int pgpu_fill_debug(int mode, const float* m, const float* g1, const float* g2, float* o, uint8_t* t,
                    const uint8_t* z, int L1, int L2);

/*
 * Guide-tree clustering.  Replaces HierarchicalClusteringAlgorithm.merge_order
 * (praline/util/cluster.py:27-57) and its linkages (:60-111): dist_dev is the [n][n] f32
 * distance matrix GuideTreeBuilder forms (component/tree.py:142-147), linkage 0 = single,
 * 1 = complete, 2 = average.  merges_dev receives the n-1 (merge_one_id, merge_two_id) pairs in
 * the reference's order (first minimum in row-major order over the clusters in ascending id;
 * `two` is merged into `one`).  workspace_dev: pgpu_cluster_workspace_bytes(n) bytes.
 */
#define PGPU_LINKAGE_SINGLE 0
#define PGPU_LINKAGE_COMPLETE 1
#define PGPU_LINKAGE_AVERAGE 2
int64_t pgpu_cluster_workspace_bytes(int n);
int pgpu_cluster_merge_order(int n, int linkage, const float* dist_dev, void* workspace_dev, int32_t* merges_dev,
                             void* stream);

/*
 * Distance matrix of GuideTreeBuilder (praline/component/tree.py:92-147) from the condensed score
 * vector of the all-vs-all stage (np.triu_indices order): d[i][j] = d[j][i] = score, d[i][i] = 0.0
 * (tree.py:132-133), dist = (-d) + d.max() in f32 (tree.py:142-147).  dist_dev: [n][n] f32;
 * scratch_dev: 4 bytes.  The vector may be in rank slices (the padded all-gather layout): slot s with
 * cuts_dev[r] <= s < cuts_dev[r+1] is read at cond_dev[s + shift_dev[r]]; n_cuts = 0 for a plain vector.
 */
int pgpu_tree_distance(int n, const float* cond_dev, int n_cuts, const int64_t* cuts_dev, const int64_t* shift_dev,
                       float* dist_dev, void* scratch_dev, void* stream);

/*
 * Pipe-rate micro-benchmarks used for the DP roofline denominator: returns in out[0..n) the
 * measured warp-instructions per NANOSECOND per SM (wall clock, CUDA events) for (0) FADD, (1) FMNMX, (2) FMNMX3,
 * (3) the 4 FADD : 3 FMNMX mix of the score-only recurrence, (4) VIADDMNMX.S32,
 * (5) VIADDMNMX.S16x2, (6) SHFL, (7) LDS.128, out[8] = SM clock in MHz held during a burst,
 * (9) SHF, (10) IMAD, (11) LOP3, (12) IADD3, (13) the packed s16x2 recurrence of the int16 kernels
 * (2 VIADD.16x2, 2 VIADDMNMX.S16x2, 1 VIMNMX3.S16x2 per two cells).  n must be >= 14.
 */
int pgpu_microbench(double* out, int n);

#ifdef __cplusplus
}
#endif
#endif
