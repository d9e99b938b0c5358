// capi.cu -- the C ABI of libpraline_b200.so (declared in include/praline_b200.h).
// Plain pointers and sizes only; every entry returns 0 on success or a non-zero code with the
// text available from pgpu_last_error().  Nothing here throws across the boundary.
#include "common.cuh"
#include <mutex>
#include "../../include/praline_b200.h"

#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <vector>

static thread_local char g_err[512] = "";

void pg_set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" {

const char* pgpu_last_error(void) { return g_err; }
int pgpu_abi_version(void) { return PGPU_ABI_VERSION; }

int pgpu_init(int device)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        pg_set_error("no CUDA device: %s", e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
        return 3;
    }
    if (device < 0 || device >= n) { pg_set_error("device %d out of range (0..%d)", device, n - 1); return 1; }
    PG_CUDA_OK(cudaSetDevice(device));
    cudaDeviceProp p;
    PG_CUDA_OK(cudaGetDeviceProperties(&p, device));
    if (p.major < 10) { pg_set_error("built for sm_100a, device is sm_%d%d", p.major, p.minor); return 3; }
    PG_CUDA_OK(cudaFree(0));
    return 0;
}

void pgpu_shutdown(void) {}

int pgpu_supported_k(int k) { return pg_stream_supported_k(k); }
int pgpu_warps_per_tile(void) { return 8; }

int pgpu_align_tiles(int mode, int K, int transposed, const uint8_t* seqs, const int64_t* offs,
                     const int32_t* stream_ids, const void* tiles, int n_tiles, int64_t n_slots,
                     const float* S, int A, float gap_open, float gap_extend, const float* topD,
                     const float* leftD, float left0, float left1, int border_len, float* scores, uint64_t* keys, uint32_t* tb,
                     const int64_t* tb_base, int32_t* emit_t, int64_t* pair_tb, const float* mwave,
                     const int64_t* mrow_base, void* stream)
{
    if (mode < 0 || mode > 4) { pg_set_error("unknown alignment mode %d", mode); return 1; }
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (border_len < 32 * K + 1) { pg_set_error("border arrays too short for K=%d", K); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const bool semi = mode != PG_GLOBAL;   // local and semiglobal results are reduced through keys
    if (semi && !keys) { pg_set_error("local / semiglobal modes need the keys scratch buffer"); return 1; }
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    a.seqs = seqs; a.offs = offs; a.stream_ids = stream_ids; a.tiles = (const PgTile*)tiles;
    a.S = S; a.A = A; a.transposed = transposed; a.go = gap_open; a.ge = gap_extend;
    a.topD = topD; a.leftD = leftD; a.border_len = border_len;
    a.left0 = left0; a.left1 = left1;
    a.scores = scores;
    a.rowkey = (unsigned long long*)keys;
    a.colkey = keys ? (unsigned long long*)keys + n_slots : nullptr;
    a.tb = tb; a.tb_base = tb_base; a.emit_t = emit_t; a.pair_tb = pair_tb;
    a.mwave = mwave; a.mrow_base = mrow_base;
    // kernel-orientation mode: a transposed launch swaps the roles of sequence one and two
    int kmode = mode;
    if (transposed && (mode == PG_SG_ONE || mode == PG_SG_TWO)) kmode = (mode == PG_SG_ONE) ? PG_SG_TWO : PG_SG_ONE;
    (void)kmode;  // borders arrive pre-oriented in topD / leftD; only the score rule needs `mode`
    if (semi) PG_CUDA_OK(cudaMemsetAsync(keys, 0, sizeof(uint64_t) * 2 * (size_t)n_slots, st));
    int rc = pg_launch_stream(a, n_tiles, K, mode, tb != nullptr, st);
    if (rc) return rc;
    if (semi) rc = pg_launch_semi_scores(n_slots, a.rowkey, a.colkey, mode, transposed, scores, st);
    return rc;
}

int pgpu_align_tiles16(int K, int paired, int transposed, const uint8_t* seqs, const int64_t* offs, const int32_t* stream_ids,
                       const void* tiles, int n_tiles, const float* S, int A, int gap_open, int gap_extend,
                       int neg, const float* topD, int left0, int left1, int border_len, float* scores,
                       void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (border_len < 32 * K + 1) { pg_set_error("border array too short for K=%d", K); return 1; }
    if (neg > -1 || neg < -32000) { pg_set_error("sentinel %d outside the int16 working range", neg); return 1; }
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    a.seqs = seqs; a.offs = offs; a.stream_ids = stream_ids; a.tiles = (const PgTile*)tiles;
    a.S = S; a.A = A; a.transposed = transposed; a.topD = topD; a.border_len = border_len; a.scores = scores;
    a.go16 = gap_open; a.ge16 = gap_extend; a.neg16 = neg; a.left0_16 = left0; a.left1_16 = left1;
    return pg_launch_stream16(a, n_tiles, K, paired, (cudaStream_t)stream);
}

// Traced form of pgpu_align_tiles16 (two streamed sequences per warp): global mode, integer scores,
// every value within +-16000.  Traceback words in the packed layout (tb_fmt = 1 of
// pgpu_traceback_tiles); emit_t carries the register half of the slot in bit 30.
int pgpu_align_tiles16_traced(int K, int transposed, const uint8_t* seqs, const int64_t* offs, const int32_t* stream_ids,
                              const void* tiles, int n_tiles, const float* S, int A, int gap_open, int gap_extend,
                              int neg, const float* topD, int left0, int left1, int border_len, float* scores,
                              uint32_t* tb, const int64_t* tb_base, int32_t* emit_t, int64_t* pair_tb, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (border_len < 32 * K + 1) { pg_set_error("border array too short for K=%d", K); return 1; }
    if (neg > -1 || neg < -16000) { pg_set_error("sentinel %d outside the traced int16 working range", neg); return 1; }
    if (!tb || !tb_base || !emit_t || !pair_tb) { pg_set_error("traced launch needs the traceback buffers"); return 1; }
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    a.seqs = seqs; a.offs = offs; a.stream_ids = stream_ids; a.tiles = (const PgTile*)tiles;
    a.S = S; a.A = A; a.transposed = transposed; a.topD = topD; a.border_len = border_len; a.scores = scores;
    a.go16 = gap_open; a.ge16 = gap_extend; a.neg16 = neg; a.left0_16 = left0; a.left1_16 = left1;
    a.tb = tb; a.tb_base = tb_base; a.emit_t = emit_t; a.pair_tb = pair_tb;
    a.all_ones = -1;
    return pg_launch_stream16(a, n_tiles, K, 0, (cudaStream_t)stream);
}

int pgpu_traceback_tiles(int mode, int K, int transposed, const int64_t* offs, const int32_t* slot_resident,
                         const int32_t* slot_stream, int64_t n_slots, const uint64_t* keys,
                         const uint32_t* tb, const int32_t* emit_t, const int64_t* pair_tb, int code00,
                         int top_ramp, int left_ramp, const int64_t* path_off, int32_t* path_buf,
                         int32_t* path_start, int32_t* path_len, const uint8_t* seqs, int32_t* counts,
                         const int64_t* cnt_off, int A, const float* scores, int use_thr, float thr, int tb_fmt,
                         void* stream)
{
    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.tb_fmt = tb_fmt;
    a.n_slots = n_slots; a.mode = mode; a.K = K; a.transposed = transposed; a.offs = offs;
    a.slot_resident = slot_resident; a.slot_stream = slot_stream; a.tb = tb; a.emit_t = emit_t;
    a.pair_tb = pair_tb;
    a.rowkey = (const unsigned long long*)keys;
    a.colkey = keys ? (const unsigned long long*)keys + n_slots : nullptr;
    a.code00 = code00; a.top_ramp = top_ramp; a.left_ramp = left_ramp;
    a.path_off = path_off; a.path_buf = path_buf; a.path_start = path_start; a.path_len = path_len;
    a.seqs = seqs; a.counts = counts; a.cnt_off = cnt_off; a.A = A; a.scores = scores; a.use_thr = use_thr; a.thr = thr;
    if (counts && (!seqs || !cnt_off || (use_thr && !scores))) { pg_set_error("preprofile mode needs seqs, cnt_off and scores"); return 1; }
    if (counts && mode != PG_GLOBAL) { pg_set_error("preprofile counts are defined for global master-slave alignments"); return 1; }
    if (!counts && !path_buf) { pg_set_error("nothing to produce: neither paths nor counts requested"); return 1; }
    return pg_launch_traceback(a, (cudaStream_t)stream);
}

// Paired-resident traced fill (gotoh_stream16r.cuh, TB): tiles as for pgpu_align_tiles16 with paired = 1; every
// slot is an UNORDERED pair whose words serve both orientations (tb_fmt = 2, pgpu_traceback_dual).  The kernel
// records the (resident, streamed) sequence ids of every slot.  Needs a symmetric S (the caller checks).
int pgpu_align_tiles16_paired_traced(int K, const uint8_t* seqs, const int64_t* offs, const void* tiles, int n_tiles,
                                     const float* S, int A, int gap_open, int gap_extend, int neg, const float* topD,
                                     int left0, int left1, int border_len, float* scores, uint32_t* tb,
                                     const int64_t* tb_base, int32_t* emit_t, int64_t* pair_tb, int32_t* slot_res,
                                     int32_t* slot_str, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (border_len < 32 * K + 1) { pg_set_error("border array too short for K=%d", K); return 1; }
    if (neg > -1 || neg < -16000) { pg_set_error("sentinel %d outside the traced int16 working range", neg); return 1; }
    if (!tb || !tb_base || !emit_t || !pair_tb || !slot_res || !slot_str) { pg_set_error("traced launch needs the traceback buffers"); return 1; }
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    a.seqs = seqs; a.offs = offs; a.stream_ids = nullptr; a.tiles = (const PgTile*)tiles;
    a.S = S; a.A = A; a.transposed = 1; a.topD = topD; a.border_len = border_len; a.scores = scores;
    a.go16 = gap_open; a.ge16 = gap_extend; a.neg16 = neg; a.left0_16 = left0; a.left1_16 = left1;
    a.tb = tb; a.tb_base = tb_base; a.emit_t = emit_t; a.pair_tb = pair_tb; a.slot_res = slot_res; a.slot_str = slot_str;
    return pg_launch_stream16rt(a, n_tiles, K, (cudaStream_t)stream);
}

// Both walks of every slot of a paired-resident traced fill: thread 2k = (sequence_one = resident), 2k + 1 =
// (sequence_one = streamed).  counts + seq_cnt_off (offset of every sequence's table, < 0: not a master):
// preprofile mode; path_* (indexed by 2 * slot + orientation): reference-format paths.
int pgpu_traceback_dual(int K, const int64_t* offs, const int32_t* slot_res, const int32_t* slot_str, int64_t n_slots,
                        const uint32_t* tb, const int32_t* emit_t, const int64_t* pair_tb, int code00, int top_ramp,
                        int left_ramp, const uint8_t* seqs, int32_t* counts, const int64_t* seq_cnt_off, int A,
                        const float* scores, int use_thr, float thr, const int64_t* path_off, int32_t* path_buf,
                        int32_t* path_start, int32_t* path_len, void* stream)
{
    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.tb_fmt = 2; a.dual = 1;
    a.n_slots = n_slots; a.mode = PG_GLOBAL; a.K = K; a.offs = offs;
    a.slot_resident = slot_res; a.slot_stream = slot_str; a.tb = tb; a.emit_t = emit_t; a.pair_tb = pair_tb;
    a.code00 = code00; a.top_ramp = top_ramp; a.left_ramp = left_ramp;
    a.path_off = path_off; a.path_buf = path_buf; a.path_start = path_start; a.path_len = path_len;
    a.seqs = seqs; a.counts = counts; a.seq_cnt_off = seq_cnt_off; a.A = A; a.scores = scores; a.use_thr = use_thr; a.thr = thr;
    if (counts && (!seqs || !seq_cnt_off || (use_thr && !scores))) { pg_set_error("preprofile mode needs seqs, seq_cnt_off and scores"); return 1; }
    if (!counts && !path_buf) { pg_set_error("nothing to produce: neither paths nor counts requested"); return 1; }
    return pg_launch_traceback(a, (cudaStream_t)stream);
}

// Local traced batch (K2 local instantiations) and its walk (K4 local mode): the inner loop of
// LocalMasterSlaveAligner (preprofile.py:227-267), boxes = Waterman-Eggert masks per slot.
int pgpu_align_tiles_local(int K, const uint8_t* seqs, const int64_t* offs, const int32_t* stream_ids,
                           const void* tiles, int n_tiles, int64_t n_slots, const float* S, int A,
                           float gap_open, float gap_extend, const float* topD, float left0, float left1,
                           int border_len, float* scores, uint64_t* keys, uint32_t* tb, const int64_t* tb_base,
                           int32_t* emit_t, int64_t* pair_tb, const int32_t* boxes, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (border_len < 32 * K + 1) { pg_set_error("border arrays too short for K=%d", K); return 1; }
    if (!keys || !tb || !tb_base || !emit_t || !pair_tb) { pg_set_error("local traced launch needs keys and the traceback buffers"); return 1; }
    if (gap_open > 0.f || gap_extend > 0.f) { pg_set_error("local traced batches need gap penalties <= 0"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    StreamArgs a;
    memset(&a, 0, sizeof(a));
    a.seqs = seqs; a.offs = offs; a.stream_ids = stream_ids; a.tiles = (const PgTile*)tiles;
    a.S = S; a.A = A; a.transposed = 0; a.go = gap_open; a.ge = gap_extend;
    a.topD = topD; a.border_len = border_len; a.left0 = left0; a.left1 = left1;
    a.scores = scores;
    a.rowkey = (unsigned long long*)keys;
    a.colkey = (unsigned long long*)keys + n_slots;
    a.tb = tb; a.tb_base = tb_base; a.emit_t = emit_t; a.pair_tb = pair_tb;
    a.boxes = (const int4*)boxes;
    PG_CUDA_OK(cudaMemsetAsync(keys, 0, sizeof(uint64_t) * 2 * (size_t)n_slots, st));
    int rc = pg_launch_stream_local(a, n_tiles, K, boxes != nullptr, st);
    if (rc) return rc;
    return pg_launch_semi_scores(n_slots, a.rowkey, a.colkey, PG_LOCAL, 0, scores, st);
}

int pgpu_traceback_tiles_local(int K, const int64_t* offs, const int32_t* slot_resident, const int32_t* slot_stream,
                               int64_t n_slots, const uint64_t* keys, const uint32_t* tb, const int32_t* emit_t,
                               const int64_t* pair_tb, int code00, const int64_t* path_off, int32_t* path_buf,
                               int32_t* path_start, int32_t* path_len, const uint8_t* seqs, int32_t* counts,
                               const int64_t* cnt_off, int A, const float* scores, int use_thr, float thr,
                               const int32_t* boxes, int32_t* box_out, int box_slot, void* stream)
{
    TraceArgs a;
    memset(&a, 0, sizeof(a));
    a.n_slots = n_slots; a.mode = PG_LOCAL; a.K = K; a.transposed = 0; a.offs = offs;
    a.slot_resident = slot_resident; a.slot_stream = slot_stream; a.tb = tb; a.emit_t = emit_t; a.pair_tb = pair_tb;
    a.rowkey = (const unsigned long long*)keys;
    a.colkey = keys ? (const unsigned long long*)keys + n_slots : nullptr;
    a.code00 = code00; a.top_ramp = 1; a.left_ramp = 1;    // local borders are the gap ramps (align.py:370-385)
    a.path_off = path_off; a.path_buf = path_buf; a.path_start = path_start; a.path_len = path_len;
    a.seqs = seqs; a.counts = counts; a.cnt_off = cnt_off; a.A = A; a.scores = scores; a.use_thr = use_thr; a.thr = thr;
    a.boxes = boxes; a.box_out = box_out; a.box_slot = box_slot;
    if (!keys) { pg_set_error("local walks start from the keys of pgpu_align_tiles_local"); return 1; }
    if (counts && (!seqs || !cnt_off || (use_thr && !scores))) { pg_set_error("preprofile mode needs seqs, cnt_off and scores"); return 1; }
    if (!counts && !path_buf && !box_out) { pg_set_error("nothing to produce: neither paths, counts nor boxes requested"); return 1; }
    if (box_out && (box_slot < 0 || box_slot >= PG_NBOX)) { pg_set_error("box_slot %d outside 0..%d", box_slot, PG_NBOX - 1); return 1; }
    return pg_launch_traceback(a, (cudaStream_t)stream);
}

int pgpu_build_scores(int n_sets, const float* const* P1, const float* const* P2, const float* const* S,
                      const int* A, int L1, int L2, float* m, int m_pitch, void* stream)
{
    if (n_sets < 1 || n_sets > 8) { pg_set_error("n_sets %d outside 1..8", n_sets); return 1; }
    ScoreSets h;
    h.n = n_sets;
    for (int i = 0; i < n_sets; i++) { h.s[i].P1 = P1[i]; h.s[i].P2 = P2[i]; h.s[i].S = S[i]; h.s[i].A = A[i]; }
    return pg_launch_build_scores(h, L1, L2, m, m_pitch, (cudaStream_t)stream);
}

int pgpu_build_rows(const float* prof, const int64_t* rowoff, int A, const float* S, const void* blocks, int n_blocks,
                    int width, int transposed, int local_mode, int dense_syms, float* mwave, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    return pg_launch_build_rows(prof, rowoff, A, S, (const PgRowBlock*)blocks, n_blocks, width, transposed,
                                local_mode ? -INFINITY : 0.f, dense_syms, mwave, (cudaStream_t)stream);
}

int pgpu_build_rows_fast(const float* prof, const float* wres, const int64_t* rowoff, int A, const void* blocks,
                         int n_blocks, int width, int local_mode, float* mwave, void* stream)
{
    return pg_launch_build_rows_fast(prof, wres, rowoff, A, (const PgRowBlock*)blocks, n_blocks, width,
                                     local_mode ? -INFINITY : 0.f, mwave, (cudaStream_t)stream);
}

int pgpu_build_rows_tc(const float* prof, const float* wres, int A, const void* quads, int n_quads, int width,
                       int local_mode, float* mwave, const void* whi, const void* wlo, const void* phi, const void* plo,
                       void* stream)
{
    return pg_launch_build_rows_tc(prof, wres, A, quads, n_quads, width, local_mode ? -INFINITY : 0.f, mwave, whi, wlo,
                                   phi, plo, (cudaStream_t)stream);
}

int pgpu_split_residents(const float* wres, const int64_t* rowoff, const int64_t* padoff, int n_seqs, int A, void* whi,
                         void* wlo, void* stream)
{
    return pg_launch_split_residents(wres, rowoff, padoff, n_seqs, A, whi, wlo, (cudaStream_t)stream);
}

int pgpu_profile_times_matrix(const float* prof, const float* S, int A, int64_t n_rows, int transposed, float* out,
                              void* stream)
{
    return pg_launch_profile_times_matrix(prof, S, A, n_rows, transposed, out, (cudaStream_t)stream);
}

int pgpu_build_scores_seq(const uint8_t* a, const uint8_t* b, const float* S, int A, int L1, int L2,
                          float* m, int m_pitch, void* stream)
{
    return pg_launch_build_scores_seq(a, b, S, A, L1, L2, m, m_pitch, (cudaStream_t)stream);
}

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

// Columns per lane of the general wavefront kernel.  `full` (debug dumps of o and t) runs the
// reference-flag kernel, instantiated for 2 and 8.  The production kernel uses ONE strip (no
// inter-strip handshake at all) while the columns fit 32 lanes x 16, and thin strips beyond
// that so that many warps ride the anti-diagonal.  PGPU_KG overrides for experiments.
static int general_kg(int L2, bool full)
{
    if (full) return L2 <= 2048 ? 2 : 8;
    const char* e = getenv("PGPU_KG");
    if (e && atoi(e) > 0) return atoi(e);
    if (L2 <= 512) { int kg = 1; while (32 * kg < L2) kg *= 2; return kg; }
    return L2 <= 8192 ? 4 : 2;
}
static int general_strips(int L2, int kg) { int ns = (L2 + 32 * kg - 1) / (32 * kg); return ns < 1 ? 1 : ns; }

int64_t pgpu_general_workspace_bytes(int L1, int L2)
{
    const int nsa = general_strips(L2, general_kg(L2, false)), nsb = general_strips(L2, general_kg(L2, true));
    const int nsl = (L2 + 127) / 128;
    const int ns = (nsa > nsb ? nsa : nsb) > nsl ? (nsa > nsb ? nsa : nsb) : nsl;
    const size_t fp = up256((size_t)L2 + 1);
    size_t b = 0;
    b += up256(sizeof(uint32_t) * 32 * (size_t)nsl * (L1 + 128));   // lean kernels' flag words (row-blocked: 4 * (ceil(L1 / 4) + 31) rows)
    b += up256((size_t)(L1 + 1) * fp);                              // flags
    b += up256(sizeof(float) * 8 * (size_t)(ns + 1) * (L1 + 1));    // edge (16-byte records; 32-byte tagged ones in the lean kernel)
    b += up256(sizeof(int) * (size_t)(ns + 1));                     // progress
    b += 2 * up256(sizeof(float) * 3 * (size_t)(L2 + 1));           // top, lastrow
    b += up256(sizeof(float) * 3 * (size_t)(L1 + 1));               // lastcol
    b += 256;                                                       // best, err
    b += up256(sizeof(int) * (size_t)((L1 + 127) / 128) * ((L2 + 127) / 128));   // K1 block flags (pgpu_align_profile_long)
    return (int64_t)b;
}

static int general_args(GenArgs& a, int& kg, int mode, int L1, int L2, const float* m, int m_pitch, const float* g1,
                        const float* g2, int var_gaps, const uint8_t* z, int z_pitch, void* workspace, float* score_out,
                        int32_t* cell_out, int32_t* path_buf, int32_t* path_start, int32_t* path_len,
                        float* o_full, uint8_t* t_full, int** ready_out)
{
    if (mode < 0 || mode > 4) { pg_set_error("unknown alignment mode %d", mode); return 1; }
    if (L1 < 1 || L2 < 1) { pg_set_error("empty sequence (L1=%d, L2=%d)", L1, L2); return 1; }
    kg = general_kg(L2, o_full != nullptr);
    const int nsa = general_strips(L2, general_kg(L2, false)), nsb = general_strips(L2, general_kg(L2, true));
    const int nsl = (L2 + 127) / 128;
    const int ns = (nsa > nsb ? nsa : nsb) > nsl ? (nsa > nsb ? nsa : nsb) : nsl;   // workspace fits every layout
    memset(&a, 0, sizeof(a));
    a.mode = mode; a.L1 = L1; a.L2 = L2; a.m = m; a.m_pitch = m_pitch; a.g1 = g1; a.g2 = g2;
    a.z = z; a.z_pitch = z_pitch; a.o_full = o_full; a.t_full = t_full;
    unsigned char* w = (unsigned char*)workspace;
    const size_t fp = up256((size_t)L2 + 1);
    a.flagw = (uint32_t*)w; w += up256(sizeof(uint32_t) * 32 * (size_t)nsl * (L1 + 128));
    a.var_gaps = var_gaps;
    a.flags = w; a.f_pitch = (int)fp; w += up256((size_t)(L1 + 1) * fp);
    a.edge = (float*)w; w += up256(sizeof(float) * 8 * (size_t)(ns + 1) * (L1 + 1));
    a.progress = (int*)w; w += up256(sizeof(int) * (size_t)(ns + 1));
    a.top = (float*)w; w += up256(sizeof(float) * 3 * (size_t)(L2 + 1));
    a.lastrow = (float*)w; w += up256(sizeof(float) * 3 * (size_t)(L2 + 1));
    a.lastcol = (float*)w; w += up256(sizeof(float) * 3 * (size_t)(L1 + 1));
    a.best = (unsigned long long*)w;
    a.err = (int*)(w + 16);
    w += 256;
    if (ready_out) *ready_out = (int*)w;        // [ceil(L1 / 128)][ceil(L2 / 128)] block flags of K1 (pgpu_align_profile_long)
    a.score_out = score_out; a.cell_out = cell_out;
    a.path_buf = path_buf; a.path_start = path_start; a.path_len = path_len;
    return 0;
}

int pgpu_align_general(int mode, int L1, int L2, const float* m, int m_pitch, const float* g1,
                       const float* g2, int var_gaps, const uint8_t* z, int z_pitch, void* workspace, float* score_out,
                       int32_t* cell_out, int32_t* path_buf, int32_t* path_start, int32_t* path_len,
                       float* o_full, uint8_t* t_full, void* stream)
{
    GenArgs a;
    int kg = 0;
    const int rc = general_args(a, kg, mode, L1, L2, m, m_pitch, g1, g2, var_gaps, z, z_pitch, workspace, score_out, cell_out,
                                path_buf, path_start, path_len, o_full, t_full, nullptr);
    if (rc) return rc;
    return pg_launch_general(a, kg, (cudaStream_t)stream);
}

// One long profile x profile alignment, K1 and K3 in one call (BASELINE config 5): cext_build_scores for one track set +
// RawPairwiseAligner (cext.c:308-455, component/align.py:302-447).  When the row-blocked wavefront serves the fill
// (global / semiglobal, constant gaps) and the column kernel builds m, the two run SIDE BY SIDE: K1 on the caller's
// stream publishes a flag per finished 128 x 128 block of m, the fill -- on a high-priority stream of its own, so that
// its few CTAs are placed as soon as K1's first blocks retire -- acquires the flag of a block before it requests rows
// of it.  The fill is a latency-bound wavefront on 79 SMs x 2 warps and K1 needs a fifth of its time: K1 disappears
// behind it.  Launch order K1 first: under a serialising tool (ncu, CUDA_LAUNCH_BLOCKING) the fill simply finds every
// flag set.  Anything else runs K1, then the fill, on the caller's stream.
int pgpu_align_profile_long(int mode, const float* P1, const float* P2, const float* S, int A, int L1, int L2, float* m,
                            int m_pitch, const float* g1, const float* g2, int var_gaps, void* workspace, float* score_out,
                            int32_t* cell_out, int32_t* path_buf, int32_t* path_start, int32_t* path_len, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    GenArgs a;
    int kg = 0;
    int* ready = nullptr;
    int rc = general_args(a, kg, mode, L1, L2, m, m_pitch, g1, g2, var_gaps, nullptr, 0, workspace, score_out, cell_out,
                          path_buf, path_start, path_len, nullptr, nullptr, &ready);
    if (rc) return rc;
    ScoreSets h;
    h.n = 1;
    h.s[0].P1 = P1; h.s[0].P2 = P2; h.s[0].S = S; h.s[0].A = A;
    cudaStream_t st = (cudaStream_t)stream;
    const bool side_by_side = pg_general_uses_wave4(a) && pg_build_scores_flags_blocks(1, L1, L2) &&
                              getenv("PGPU_NO_K1_OVERLAP") == nullptr;
    if (!side_by_side) {
        rc = pg_launch_build_scores(h, L1, L2, m, m_pitch, st);
        return rc ? rc : pg_launch_general(a, kg, st);
    }
    // one high-priority stream and two events per device, made on first use
    static cudaStream_t hp[16] = {};
    static cudaEvent_t ev[16][2] = {};
    int dev = 0;
    PG_CUDA_OK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 16) { pg_set_error("device ordinal %d outside 0..15", dev); return 1; }
    static std::mutex once;
    std::lock_guard<std::mutex> guard(once);                 // creation, and one overlapped alignment per process at a time:
                                                             // the events are shared
    if (hp[dev] == nullptr) {
        int lo = 0, hi = 0;
        PG_CUDA_OK(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        PG_CUDA_OK(cudaStreamCreateWithPriority(&hp[dev], cudaStreamNonBlocking, hi));
        PG_CUDA_OK(cudaEventCreateWithFlags(&ev[dev][0], cudaEventDisableTiming));
        PG_CUDA_OK(cudaEventCreateWithFlags(&ev[dev][1], cudaEventDisableTiming));
    }
    const int nx = (L2 + 127) / 128, ny = (L1 + 127) / 128;
    PG_CUDA_OK(cudaMemsetAsync(ready, 0, sizeof(int) * (size_t)nx * ny, st));
    PG_CUDA_OK(cudaEventRecord(ev[dev][0], st));            // the caller's uploads and the cleared flags
    PG_CUDA_OK(cudaStreamWaitEvent(hp[dev], ev[dev][0], 0));
    rc = pg_launch_build_scores(h, L1, L2, m, m_pitch, st, ready);
    if (rc) return rc;
    a.m_ready = ready;
    a.m_ready_nx = nx;
    rc = pg_launch_general(a, kg, hp[dev]);
    // the caller's stream continues when both are done (success or not: nothing may outlive the call's buffers)
    if (cudaEventRecord(ev[dev][1], hp[dev]) != cudaSuccess || cudaStreamWaitEvent(st, ev[dev][1], 0) != cudaSuccess) {
        if (!rc) { pg_set_error("joining the fill's stream failed"); rc = 2; }
    }
    return rc;
}

// B3 parity shim: the reference's cext_align_<mode>(m, g1, g2, o, t, z) with host buffers
// (praline/util/cext.c:103-107).  o and t are fully overwritten (borders included).
int pgpu_fill_debug(int mode, const float* m, const float* g1, const float* g2, float* o, uint8_t* t,
                    const uint8_t* z, int L1, int L2)
{
    if (L1 < 1 || L2 < 1) { pg_set_error("empty sequence (L1=%d, L2=%d)", L1, L2); return 1; }
    const size_t cells = (size_t)(L1 + 1) * (L2 + 1);
    const int64_t wsb = pgpu_general_workspace_bytes(L1, L2);
    unsigned char* d = nullptr;
    const size_t off_m = 0, off_g1 = up256(sizeof(float) * (size_t)L1 * L2), off_g2 = off_g1 + up256(sizeof(float) * 2 * L1),
                 off_z = off_g2 + up256(sizeof(float) * 2 * L2), off_o = off_z + up256(cells),
                 off_t = off_o + up256(sizeof(float) * 3 * cells), off_ws = off_t + up256(3 * cells),
                 off_out = off_ws + (size_t)wsb, total = off_out + 256;
    PG_CUDA_OK(cudaMalloc((void**)&d, total));
    int rc = 0;
    do {
        cudaError_t e;
#define CK(x) if ((e = (x)) != cudaSuccess) { pg_set_error("%s: %s", #x, cudaGetErrorString(e)); rc = 2; break; }
        CK(cudaMemcpy(d + off_m, m, sizeof(float) * (size_t)L1 * L2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d + off_g1, g1, sizeof(float) * 2 * L1, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d + off_g2, g2, sizeof(float) * 2 * L2, cudaMemcpyHostToDevice));
        if (z) CK(cudaMemcpy(d + off_z, z, cells, cudaMemcpyHostToDevice));
        CK(cudaMemset(d + off_o, 0, sizeof(float) * 3 * cells));
        CK(cudaMemset(d + off_t, 0, 3 * cells));
        rc = pgpu_align_general(mode, L1, L2, (const float*)(d + off_m), L2, (const float*)(d + off_g1),
                                (const float*)(d + off_g2), 1, z ? d + off_z : nullptr, L2 + 1, d + off_ws,
                                (float*)(d + off_out), (int32_t*)(d + off_out + 16), nullptr, nullptr, nullptr,
                                (float*)(d + off_o), d + off_t, nullptr);
        if (rc) break;
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(o, d + off_o, sizeof(float) * 3 * cells, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(t, d + off_t, 3 * cells, cudaMemcpyDeviceToHost));
#undef CK
    } while (0);
    cudaFree(d);
    return rc;
}

// Progressive-merge glue on device-resident count tables (merge.cu).
int pgpu_counts_to_profile(const int32_t* counts_dev, int64_t n_rows, int A, float* prof_dev, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    return pg_launch_counts_to_profile(counts_dev, n_rows, A, prof_dev, (cudaStream_t)stream);
}

int pgpu_merge_counts(const int32_t* counts_one_dev, const int32_t* counts_two_dev, int A, const int32_t* align_out_dev,
                      int32_t* merged_dev, int max_rows, void* stream)
{
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (!counts_one_dev || !counts_two_dev || !align_out_dev || !merged_dev) { pg_set_error("merge_counts: null buffer"); return 1; }
    return pg_launch_merge_counts(counts_one_dev, counts_two_dev, A, align_out_dev, merged_dev, max_rows, (cudaStream_t)stream);
}

// Guide-tree clustering (cluster.cu): the merge order of praline/util/cluster.py:27-57.
int64_t pgpu_cluster_workspace_bytes(int n) { return n < 1 ? 0 : (int64_t)pg_cluster_workspace_bytes(n); }

int pgpu_cluster_merge_order(int n, int linkage, const float* dist_dev, void* workspace_dev, int32_t* merges_dev,
                             void* stream)
{
    if (!dist_dev || !workspace_dev || (n > 1 && !merges_dev)) { pg_set_error("cluster: null buffer"); return 1; }
    return pg_launch_cluster(n, linkage, dist_dev, workspace_dev, merges_dev, (cudaStream_t)stream);
}

// Distance matrix of GuideTreeBuilder from the condensed score vector (cluster.cu).
int pgpu_tree_distance(int n, const float* cond_dev, int n_cuts, const int64_t* cuts_dev, const int64_t* shift_dev,
                       float* dist_dev, void* scratch_dev, void* stream)
{
    if (n < 1) return 0;
    if (!dist_dev || !scratch_dev || (n > 1 && !cond_dev) || (n_cuts > 0 && (!cuts_dev || !shift_dev))) {
        pg_set_error("tree_distance: null buffer");
        return 1;
    }
    return pg_launch_tree_distance(n, cond_dev, n_cuts, cuts_dev, shift_dev, dist_dev, (unsigned*)scratch_dev,
                                   (cudaStream_t)stream);
}

}  // extern "C"
