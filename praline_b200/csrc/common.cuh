// common.cuh -- shared declarations of the sm_100a alignment kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

// Alignment modes, numbered as in the reference (praline/util/cext.c:27-31).
enum { PG_GLOBAL = 0, PG_LOCAL = 1, PG_SG_BOTH = 2, PG_SG_ONE = 3, PG_SG_TWO = 4 };

// Tie flags of the reference traceback (praline/util/cext.c:9-15).  All seven fit one byte,
// so the debug fill keeps one byte per cell instead of the reference's three.
enum { TB_MM = 1 << 1, TB_MU = 1 << 2, TB_ML = 1 << 3, TB_UO = 1 << 4, TB_UE = 1 << 5,
       TB_LO = 1 << 6, TB_LE = 1 << 7 };

// Waterman-Eggert boxes per pair (local master-slave alignments, reference preprofile.py:227-267):
// iteration n masks the bounding boxes of the n - 1 alignments before it; PG_NBOX + 1 iterations
// are served by the batched path.  A box is (ylo, yhi, xlo, xhi), inclusive, reference orientation;
// (1, 0, 1, 0) is empty.
#define PG_NBOX 3

// One unit of work of the inter-task kernel: a resident sequence (laid across the lanes of
// every warp of the CTA) against a contiguous run of streamed sequences.
struct PgTile {
    int32_t resident;       // sequence id of the resident sequence
    int32_t stream_begin;   // [begin, end) into stream_ids (or sequence ids when stream_ids == NULL)
    int32_t stream_end;
    int32_t resident2;      // paired-resident launches: second resident (high half), -1 = none
    int64_t out_base;       // output slot of the first streamed sequence; slots are consecutive
    int64_t out_base2;      // paired-resident launches: first slot of the second resident
    int32_t b_skip;         // leading stream elements that have no pair with resident2
    int32_t slot_stride;    // output slot of stream element e: out_base + e * slot_stride (0 = 1); the traced paired kernel
                            // interleaves the slots of its two residents (stride 2) so that their walkers sit side by side
};

struct PgBorder {
    // Values of the DP borders in the kernel's orientation (columns = resident sequence,
    // rows = streamed sequence), see reference component/align.py:367-385.
    float d00;              // max3 of the three states at (0,0)
    float top0, top1;       // D(0,x) = top0 + x*top1  (x >= 1): the state-L ramp, or 0
    float left0, left1;     // D(y,0) = left0 + y*left1 (y >= 1): the state-U ramp, or 0
};

// Arguments of the inter-task streaming kernel (gotoh_stream.cu).
struct StreamArgs {
    const uint8_t* seqs;
    const int64_t* offs;
    const int32_t* stream_ids;   // NULL: stream element s is sequence id s
    const PgTile* tiles;
    const float* S;
    int A;
    int transposed;              // resident is the reference's sequence one
    float go, ge;
    const float* topD;           // D(0, x), x = 0 .. border_len-1 (kernel orientation)
    const float* leftD;          // D(y, 0) (kept for callers; the kernel uses left0/left1)
    float left0, left1;          // D(y, 0) = left0 + (y-1)*left1, y >= 1
    int border_len;
    float* scores;               // global / local: one per output slot
    unsigned long long* rowkey;  // semiglobal: (ordered f32 << 32 | x) of the last row
    unsigned long long* colkey;  //             (ordered f32 << 32 | y) of the last column
    uint32_t* tb;                // packed traceback words
    const int64_t* tb_base;      // word offset per (tile, warp)
    int32_t* emit_t;             // per slot: step at which the owning lane saw the last row
    int64_t* pair_tb;            // per slot: word offset of its warp's traceback region
    int go16, ge16, neg16, left0_16, left1_16;   // packed s16x2 variant (gotoh_stream16.cu)
    int all_ones;                // -1, opaque to ptxas: ~x = x * all_ones + all_ones stays an IMAD (FMA pipe)
    const int4* boxes;           // local traced + masks: [slot][PG_NBOX] boxes
    const float* mwave;          // profile batches: match scores [row][32*K], rows in stream order
    const int64_t* mrow_base;    // first matrix row per (tile, warp); row 0 of a region is the dummy row
    int32_t* slot_res;           // paired-resident traced launches: the kernel records (resident, streamed)
    int32_t* slot_str;           //   sequence ids of every slot for the walk
};

// Arguments of the per-pair traceback walk (traceback.cu).
struct TraceArgs {
    int64_t n_slots;
    int mode, K, transposed;
    const int64_t* offs;
    const int32_t* slot_resident;   // sequence id laid across the lanes
    const int32_t* slot_stream;     // sequence id streamed through
    const uint32_t* tb;
    const int32_t* emit_t;
    const int64_t* pair_tb;
    const unsigned long long* rowkey;
    const unsigned long long* colkey;
    int code00, top_ramp, left_ramp;
    const int64_t* path_off;        // per slot: first row of its region in path_buf (capacity Lr+Ls+2 rows)
    int32_t* path_buf;              // [rows][2] = (y, x) in the REFERENCE orientation
    int32_t* path_start;            // per slot: first used row within its region
    int32_t* path_len;              // per slot: rows used
    // preprofile mode (counts != NULL): per-master symbol counts instead of / besides paths
    const uint8_t* seqs;            // symbols (with offs)
    int32_t* counts;                // all masters' [L x A] tables
    const int64_t* cnt_off;         // per slot: offset of its master's table
    int A;
    const float* scores;            // per slot, for the score threshold
    int use_thr;
    float thr;
    const int32_t* boxes;           // local mode: [slot][PG_NBOX][4] masked boxes (may be NULL)
    int32_t* box_out;               // local mode: [slot][PG_NBOX][4], the walk writes the bounding box of its path
    int box_slot;                   //             into box number box_slot (may alias `boxes`)
    int tb_fmt;                     // 0: f32 kernel's nibbles, 8 rows per word; 1: packed int16 kernel's words (4 rows x 2 halves);
                                    // 2: paired-resident traced kernel (gotoh_stream16r.cuh): both orientations in one nibble
    int dual;                       // tb_fmt 2: two walks per slot (thread 2s: resident = sequence one, 2s+1: streamed = sequence one)
    const int64_t* seq_cnt_off;     // dual preprofile walks: offset of every sequence's count table, < 0 = not a master
};

// Arguments of the general single-alignment path (general.cu).
struct GenArgs {
    int mode, L1, L2;
    const float* m;  int m_pitch;        // [L1][m_pitch]
    const float* g1; const float* g2;    // [L1][2], [L2][2]
    const uint8_t* z; int z_pitch;       // [(L1+1)][z_pitch] or NULL
    uint8_t* flags;  int f_pitch;        // [(L1+1)][f_pitch], one byte per cell
    float* o_full;   uint8_t* t_full;    // optional [(L1+1)][(L2+1)][3] (debug / B3 shim)
    float* edge;                         // [n_strips+1][L1+1][4]: M, U, L, pad (16-byte records)
    int* progress;                       // [n_strips+1]
    float* top;                          // [3][L2+1]  border row 0
    float* lastrow;                      // [3][L2+1]
    float* lastcol;                      // [3][L1+1]
    unsigned long long* best;            // local mode: (ordered value << 32 | ~linear index)
    int* err;                            // device error word: a strip hand-off timed out (the fill is then invalid)
    int n_strips;
    int flag_fmt;                        // 0: the reference's seven flag bits, 1: compact sign bits, 2: lean words
    uint32_t* flagw;                     // lean kernel: [n_strips][L1+31][32] flag words (4 cells x 5 bits + mask bits)
    int var_gaps;                        // gap arrays vary per position (else g1[0..1], g2[0..1] are THE gap pairs)
    int flag_skew, flag_rows;            // lean kernels: word row of (y, lane) = y - 1 + flag_skew * lane; rows per strip
    const int* m_ready;                  // k_wave4 only: [ceil(L1 / 128)][m_ready_nx] flags set by k_build_scores_cols as it
    int m_ready_nx;                      //   finishes a 128 x 128 block of m (K1 running beside the fill); NULL: m is complete
    // finalize / traceback outputs
    float* score_out;                    // [1]
    int32_t* cell_out;                   // [3] y, x, state
    int32_t* path_buf;                   // [L1+L2+2][2]
    int32_t* path_start;                 // [1]
    int32_t* path_len;                   // [1]
};

struct ScoreSet { const float* P1; const float* P2; const float* S; int A; };
struct ScoreSets { ScoreSet s[8]; int n; };   // passed by value as a kernel argument

void pg_set_error(const char* fmt, ...);
int pg_launch_general(GenArgs a, int kg, cudaStream_t st);
int pg_launch_build_scores(const ScoreSets& sets, int L1, int L2, float* m, int m_pitch, cudaStream_t st, int* ready = nullptr);
bool pg_build_scores_flags_blocks(int n_sets, int L1, int L2);     // the launch above publishes per-block ready flags
bool pg_general_uses_wave4(const GenArgs& a);                      // pg_launch_general will run k_wave4 (which can poll them)
// <= 32 consecutive matrix rows of one streamed sequence in a wave of a profile batch
struct PgRowBlock {
    int64_t row0;     // first matrix row
    int64_t src0;     // profile row feeding matrix row row0 (row0 + r is fed by src0 + r)
    int32_t rows;
    int32_t res;      // resident sequence id
    int32_t dummy;    // matrix row row0 is the region's dummy row (no profile row)
    int32_t _pad;
};
int pg_launch_build_rows(const float* prof, const int64_t* rowoff, int A, const float* S, const PgRowBlock* blocks,
                         int n_blocks, int width, int transposed, float padv, int dense_syms, float* mwave, cudaStream_t st);
int pg_launch_build_rows_x2(const float* prof, const int64_t* rowoff, int A, const float* S, const PgRowBlock* blocks,
                            int n_blocks, int width, float padv, int ucap, float* mwave, cudaStream_t st);
int pg_launch_build_rows_fast(const float* prof, const float* wres, const int64_t* rowoff, int A, const PgRowBlock* blocks,
                              int n_blocks, int width, float padv, float* mwave, cudaStream_t st);
int pg_launch_build_rows_tc(const float* prof, const float* wres, int A, const void* quads, int n_quads, int width,
                            float padv, float* mwave, const void* whi, const void* wlo, const void* phi, const void* plo,
                            cudaStream_t st);
int pg_launch_split_residents(const float* wres, const int64_t* rowoff, const int64_t* padoff, int n_seqs, int A,
                              void* whi, void* wlo, cudaStream_t st);
int pg_launch_profile_times_matrix(const float* prof, const float* S, int A, int64_t n_rows, int transposed, float* out,
                                   cudaStream_t st);
int pg_launch_build_scores_seq(const uint8_t* a, const uint8_t* b, const float* S, int A, int L1, int L2,
                               float* m, int m_pitch, cudaStream_t st);
int pg_stream_supported_k(int k);
int pg_launch_cluster(int n, int linkage, const float* dist, void* work, int32_t* merges, cudaStream_t st);
size_t pg_cluster_workspace_bytes(int n);
int pg_launch_tree_distance(int n, const float* cond, int n_cuts, const int64_t* cuts, const int64_t* shift,
                            float* dist, unsigned* scratch, cudaStream_t st);
int pg_launch_stream(const StreamArgs& a, int n_tiles, int K, int mode, bool tb, cudaStream_t st);
int pg_launch_stream_ms(const StreamArgs& a, int n_tiles, int K, int km, cudaStream_t st);
int pg_launch_stream_local(const StreamArgs& a, int n_tiles, int K, bool masked, cudaStream_t st);
int pg_launch_stream16(const StreamArgs& a, int n_tiles, int K, int paired, cudaStream_t st);
int pg_launch_stream16rt(const StreamArgs& a, int n_tiles, int K, cudaStream_t st);
int pg_launch_semi_scores(int64_t n, const unsigned long long* rowkey, const unsigned long long* colkey,
                          int mode, int transposed, float* scores, cudaStream_t st);
int pg_launch_traceback(const TraceArgs& a, cudaStream_t st);
int pg_launch_counts_to_profile(const int32_t* counts, int64_t n_rows, int A, float* prof, cudaStream_t st);
int pg_launch_merge_counts(const int32_t* c1, const int32_t* c2, int A, const int32_t* hdr, int32_t* out, int max_rows,
                           cudaStream_t st);

#define PG_CUDA_OK(expr)                                                                  \
    do {                                                                                  \
        cudaError_t _e = (expr);                                                          \
        if (_e != cudaSuccess) {                                                          \
            pg_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                         __LINE__);                                                       \
            return 2;                                                                     \
        }                                                                                 \
    } while (0)
