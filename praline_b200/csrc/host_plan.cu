// host_plan.cu -- host-side planning of a wave of a profile batch (no device code).
//
// Engine.align_profile_pairs lays every wave of profile x profile pairs out as matrix rows in stream order and hands
// the score-row kernels (pgpu_build_rows*, general.cu / score_rows_tc.cu / score_rows_x2.cu) row blocks and 128-row
// quads.  The guide-tree stage of BASELINE config 4 plans 2.0e6 pairs = 2.6e7 row blocks in 156 waves: in numpy
// (engine.plan_profile_wave, engine.row_block_quads -- kept as the executable specification, tests/test_host_cpu.py
// compares the two) that was 1.9 s of a 5.2 s workflow; here it is one pass over the stream elements.
#include "common.cuh"
#include "../../include/praline_b200.h"

extern "C" long long pgpu_plan_profile_wave(int n_tiles, const int64_t* tile_begin, const int64_t* tile_end, int nw,
                                            const int64_t* cs, const int64_t* lens_s, const int64_t* str_s,
                                            const int64_t* res_s, const int64_t* offs, int rows_per_block,
                                            int64_t* mrow_base, pgpu_row_block* blocks, long long blocks_cap,
                                            int64_t* n_rows_out, int want_quads, const int64_t* padoff,
                                            pgpu_quad* quads, long long quads_cap, long long* n_quads_out)
{
    if (n_tiles <= 0 || nw <= 0 || rows_per_block <= 0 || rows_per_block > 32) {
        pg_set_error("plan_profile_wave: bad arguments");
        return -1;
    }
    long long nb = 0, nq = 0;
    int64_t row = 0;                        // first matrix row of the current region
    int64_t run_res = -1;                   // resident of the current run of blocks
    int in_quad = 0;                        // blocks in the open quad
    for (int t = 0; t < n_tiles; t++) {
        const int64_t tb = tile_begin[t], te = tile_end[t];
        const int64_t per = (te - tb + nw - 1) / nw;
        for (int w = 0; w < nw; w++) {
            int64_t sb = tb + (int64_t)w * per;
            int64_t se = sb + per < te ? sb + per : te;
            if (sb > se) sb = se;
            mrow_base[(size_t)t * nw + w] = row;
            if (se <= sb) continue;
            for (int64_t e = sb; e < se; e++) {
                const int first = e == sb;
                // the element's rows in the matrix, the dummy row included for region openers
                const int64_t e_row0 = row + 1 + (cs[e] - cs[sb]) - first;
                const int64_t e_rows = lens_s[e] + first;
                const int64_t seq = str_s[e];
                const int64_t e_src0 = offs[seq] - first;
                const int64_t r = res_s[e];
                for (int64_t within = 0; within < e_rows; within += rows_per_block) {
                    if (nb >= blocks_cap) { pg_set_error("plan_profile_wave: block capacity %lld too small", blocks_cap); return -1; }
                    pgpu_row_block& b = blocks[nb];
                    b.row0 = e_row0 + within;
                    b.src0 = e_src0 + within;
                    b.rows = (int32_t)(e_rows - within < rows_per_block ? e_rows - within : rows_per_block);
                    b.res = (int32_t)r;
                    b.dummy = first && within == 0;
                    b.reserved = 0;
                    if (want_quads) {
                        // groups of <= 4 consecutive blocks with one resident
                        if (r != run_res || in_quad == 4) {
                            if (nq >= quads_cap) { pg_set_error("plan_profile_wave: quad capacity %lld too small", quads_cap); return -1; }
                            pgpu_quad& q = quads[nq++];
                            q.q0 = offs[r];
                            q.Lr = (int32_t)(offs[r + 1] - offs[r]);
                            q.nblk = 0;
                            for (int k = 0; k < 4; k++) { q.row0[k] = 0; q.src0[k] = 0; q.rows[k] = 0; q.dummy[k] = 0; q.can0[k] = -1; }
                            q.bcan = padoff ? padoff[r] : 0;
                            q.reserved = 0;
                            run_res = r;
                            in_quad = 0;
                        }
                        pgpu_quad& q = quads[nq - 1];
                        q.row0[in_quad] = b.row0;
                        q.src0[in_quad] = b.src0;
                        q.rows[in_quad] = b.rows;
                        q.dummy[in_quad] = b.dummy;
                        // streamed side: a block whose first profile row sits on an 8-row group of its sequence's
                        // pre-split rows (every block but the ones with a dummy row in front) can be fetched by TMA
                        const int64_t rel = b.src0 - offs[seq];
                        q.can0[in_quad] = (padoff && !b.dummy && rel >= 0 && rel % 8 == 0) ? padoff[seq] + rel : -1;
                        q.nblk = ++in_quad;
                    }
                    nb++;
                }
            }
            row += cs[se] - cs[sb] + 1;     // + 1: the dummy row
        }
    }
    *n_rows_out = row;
    if (n_quads_out) *n_quads_out = nq;
    return nb;
}
