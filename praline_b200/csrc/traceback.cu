// traceback.cu -- K4, per-pair traceback walk over the packed nibbles written by K2.
//
// Replaces the reference's end-cell choice (praline/component/align.py:401-431), its pointer
// chase get_paths (praline/util/align.py:144-185) and extend_path_semiglobal
// (praline/util/align.py:268-297).  One thread walks one pair.
//
// K2 stores, per interior cell, 4 sign bits in kernel orientation (columns = resident
// sequence; states 0 = M, 1 = reached from above, 2 = reached from the left):
//   bit 0     max3 of (M, U, L) AT this cell is not M (M lost strictly)
//   bit 1     among the two gap states the lower-priority one won strictly
//             bits 0-1 together give the first-priority argmax of (M, U, L) at this cell -- the
//             state the reference's walker enters when it arrives here diagonally, because the
//             reference's MM > MU > ML flag priority at (y,x) picks the first maximal state of
//             (y-1,x-1) (exact for integer-valued scores, see gotoh_stream.cu);
//   bit 2     the from-above state was EXTENDED (open beats extend on ties);
//   bit 3     the from-left state was EXTENDED.
// Border cells carry no nibble; their flags follow from the border rules of
// component/align.py:367-385 (ramp borders chain back along the edge, zero borders stop).
// The path is emitted in the REFERENCE orientation, rows (y, x), written back to front into
// the pair's region so that no reversal pass is needed.
//
// Preprofile mode (a.counts != NULL).  The walk also does, on the fly, what the reference does
// on the host for every master-slave pair: compress_path(path, 0) (util/align.py:215-232),
// Alignment.merge (container/align.py:30-61) and get_frequencies (util/align.py:187-213).  Their
// net effect on the master's count table is: for each master position y, the slave residue at
// x_y is counted iff x_y > x_{y-1}, where x_y is the slave index of the FIRST path row whose
// master index is y.  Counts are order-independent sums, so pairs add with atomics straight
// into the master's [L x A] table and no path ever leaves the device.  Pairs below the score
// threshold are skipped (preprofile.py:144-145).
//
// Local mode (K2's local traced launches, reference orientation only).  The walk starts in state M
// at the first row-major maximum (key of K2; (0,0) when nothing is positive, np.argmax over the
// zero-initialised `o`, align.py:401-403) and stops (a) where an M cell carries the stop code
// (bit 0 clear, bit 1 set: its three sums are all negative, no flag in the reference), (b) at any
// interior cell inside a Waterman-Eggert box (masked cells keep t = 0, cext.c:143-148), (c) at
// (0,0) after the border chain.  The path is not extended.  Every walk writes the bounding box of
// its path (preprofile.py:252-259) for the next iteration.  Local preprofile counts: the same
// "first row per master index" rule, plus what extend_path_local's -1 padding does to
// get_frequencies (util/align.py:205-211, :234-266): the step from -1 to x0 at master row y0 >= 1
// counts slave[x0 - 1] (Python indexing: x0 = 0 is the LAST residue) against master position y0.
//
// Dual walks (tb_fmt 2, a.dual).  The paired-resident traced kernel (gotoh_stream16r.cuh) fills an unordered
// pair (r, s) ONCE; with a symmetric substitution matrix and constant gaps the DP values of the alignment
// (sequence_one = r, sequence_two = s) are the transpose of those of (s, r), so both master-slave alignments
// of a global preprofile (preprofile.py:127-154) are walks over the same words: thread 2k takes slot k with
// the resident as sequence one (transposed tie order), thread 2k + 1 with the streamed sequence as sequence
// one.  Nibble: bit 0 = U (from above) is a maximum and M is not, bit 1 = the same for L (from the left),
// bit 2 = U of this cell was opened, bit 3 = L of the NEXT column was opened (column 1 always opens: its
// left neighbours are the -inf border).  Each orientation picks its first-priority gap state from bits 0-1.
#include "common.cuh"

__device__ __forceinline__ float tkey_value(unsigned long long k)
{
    uint32_t b = (uint32_t)(k >> 32);
    b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
    return __uint_as_float(b);
}

#ifndef PG_TB_WINDOW
#define PG_TB_WINDOW 3   // measured on B200: 3 -> 11.8 ms, 8 -> 14.0 ms, 16 -> 16.3 ms per 120k traced pairs (re-priming wastes sectors)
#endif
constexpr int NWIN = PG_TB_WINDOW;

// 64 registers: a 128-thread block of walkers fits the 8,192 registers two CTAs of the traced fill leave free on an SM
__global__ void __maxnreg__(64) k_traceback(const TraceArgs a)
{
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t slot = a.dual ? (tid >> 1) : tid;
    if (slot >= a.n_slots) return;
    const bool local = a.mode == PG_LOCAL;
    const bool below = a.counts && a.use_thr && !(a.scores[slot] >= a.thr);
    if (below && !(local && a.box_out)) return;   // a Waterman-Eggert box is due even below the threshold
    if (a.dual && a.emit_t[slot] < 0) return;          // a hole of the interleaved slot numbering (no pair there)
    const int rid = a.slot_resident[slot], sid = a.slot_stream[slot];
    const int Lr = (int)(a.offs[rid + 1] - a.offs[rid]);   // kernel columns
    const int Ls = (int)(a.offs[sid + 1] - a.offs[sid]);   // kernel rows
    const int K = a.K, lr = (Lr - 1) / K;
    const int emit = a.emit_t[slot] & 0x3fffffff;
    const int half = (a.emit_t[slot] >> 30) & 1;            // packed kernel: register half that carried this pair
    const uint32_t* tbw = a.tb + a.pair_tb[slot];
    const bool TR = a.dual ? ((tid & 1) == 0) : (a.transposed != 0);
    const int L1 = TR ? Lr : Ls, L2 = TR ? Ls : Lr;        // reference lengths
    const int fmt = a.tb_fmt;

    // word holding the flags of interior cell (yk, xk), and the nibble inside it
    auto word_index = [&](int yk, int xk) -> int64_t {
        const int lane = (xk - 1) / K, k = (xk - 1) - lane * K;
        const int step = emit - (Ls - yk) - (lr - lane);
        if (fmt == 2)       // paired-resident kernel: [step / 4][lane][k], k padded to odd -- a diagonal step is the next word
            return (int64_t)(step >> 2) * ((K | 1) * 32) + lane * (K | 1) + k;
        return (int64_t)(step >> (fmt != 0 ? 2 : 3)) * (K * 32) + k * 32 + lane;
    };
    auto nib_of = [&](uint32_t w, int yk, int xk) -> uint32_t {
        const int lane = (xk - 1) / K;
        const int step = emit - (Ls - yk) - (lr - lane);
        if (fmt == 2) {
            // paired-resident kernel: 4 rows x 4 bits per register half; -> the common nibble, except that bit 3
            // of THIS cell's nibble speaks about the next column (the L branch of the walk reads it from x - 1)
            const uint32_t n = (w >> (16 * half + 4 * (3 - (step & 3)))) & 15u;
            const uint32_t gap = (n & 3u) ? 1u : 0u;
            const uint32_t second = TR ? ((n & 2u) ? 0u : 1u) : ((n & 1u) ? 0u : 1u);
            return gap | ((gap & second) << 1) | (((n >> 2) & 1u) ^ 1u) << 2 | (((n >> 3) & 1u) ^ 1u) << 3;
        }
        if (fmt == 1) {
            // packed kernel (gotoh_stream16.cu): 4 rows per word, row r at bits 2*(3-r) of every byte;
            // byte 2h+1 = (M is max, U opened), byte 2h = (first gap state is max, L opened); 1 = yes
            const int sh = 2 * (3 - (step & 3));
            const uint32_t e13 = (w >> (8 * (2 * half + 1) + sh)) & 3u, e24 = (w >> (8 * (2 * half) + sh)) & 3u;
            return ((e13 >> 1) ^ 1u) | (((e24 >> 1) ^ 1u) << 1) | (((e13 & 1u) ^ 1u) << 2) | (((e24 & 1u) ^ 1u) << 3);
        }
        return (w >> (4 * (7 - (step & 7)))) & 15u;
    };
    // The walk is a chain of dependent loads.  Alignments of related sequences move diagonally most
    // of the time and the addresses of diagonal predecessors need no data, so the words of the next
    // NWIN diagonal cells are kept in flight (pw[j] belongs to (py - j, px - j)); any other request
    // primes the window again.
    int py = -1, px = -1;
    uint32_t pw[NWIN];
#pragma unroll
    for (int j = 0; j < NWIN; j++) pw[j] = 0u;
    auto fetch = [&](int yk, int xk) -> uint32_t { return (yk >= 1 && xk >= 1) ? __ldg(tbw + word_index(yk, xk)) : 0u; };
    auto nib_at = [&](int yk, int xk) -> uint32_t {
        if (yk == py - 1 && xk == px - 1) {             // one step down the diagonal: shift the window
#pragma unroll
            for (int j = 0; j + 1 < NWIN; j++) pw[j] = pw[j + 1];
            py = yk; px = xk;
            pw[NWIN - 1] = fetch(yk - (NWIN - 1), xk - (NWIN - 1));
        } else if (!(yk == py && xk == px)) {            // off the diagonal: prime the window here
            py = yk; px = xk;
#pragma unroll
            for (int j = 0; j < NWIN; j++) pw[j] = fetch(yk - j, xk - j);
        }
        return nib_of(pw[0], yk, xk);
    };
    auto code_at = [&](int yk, int xk) -> int {
        if (yk == 0 && xk == 0) return (a.dual && !TR && a.code00) ? 3 - a.code00 : a.code00;   // a.code00: resident = sequence one
        if (yk == 0) return 2;
        if (xk == 0) return 1;
        const uint32_t nb = nib_at(yk, xk);
        if (!(nb & 1u)) return 0;
        const bool second = (nb & 2u) != 0;
        return TR ? (second ? 1 : 2) : (second ? 2 : 1);   // reference priority U before L
    };

    // ---- start cell (kernel coordinates) ----------------------------------------------------
    int yk = Ls, xk = Lr;
    int bx[PG_NBOX][4];
#pragma unroll
    for (int b = 0; b < PG_NBOX; b++) {
        const bool have = local && a.boxes;
#pragma unroll
        for (int c = 0; c < 4; c++) bx[b][c] = have ? a.boxes[slot * (PG_NBOX * 4) + 4 * b + c] : ((c & 1) ? 0 : 1);
    }
    if (local) {
        const unsigned long long rk = a.rowkey[slot];
        const uint32_t idx = ~(uint32_t)rk;
        if (tkey_value(rk) > 0.f) { yk = (int)(idx >> 11); xk = (int)(idx & 2047u); }
        else { yk = 0; xk = 0; }
    } else if (a.mode != PG_GLOBAL) {
        const unsigned long long rk = a.rowkey[slot], ck = a.colkey[slot];
        const float rv = tkey_value(rk), cv = tkey_value(ck);
        const int rx = (int)(uint32_t)rk, cy = (int)(uint32_t)ck;
        const bool from_row = (a.mode == PG_SG_BOTH || a.mode == PG_SG_TWO);
        // reference: last row (y = L1) wins only on strict '>' and when allowed; else last column
        bool take_kernel_row;   // end cell lies on the kernel's last row (yk = Ls)
        if (!TR) take_kernel_row = (rv > cv) && from_row;
        else     take_kernel_row = !((cv > rv) && from_row);
        if (take_kernel_row) { yk = Ls; xk = rx; } else { yk = cy; xk = Lr; }
    }
    int s = local ? 0 : code_at(yk, xk);
    const int end_y = yk, end_x = xk;

    const int64_t pidx = a.dual ? tid : slot;      // dual walks: two path regions per slot
    const int64_t base = a.path_buf ? a.path_off[pidx] : 0;
    const int cap = Lr + Ls + 2;
    int w = cap;
    const bool want_path = a.path_buf != nullptr;
    auto push = [&](int yr, int xr) {
        --w;
        if (want_path) {
            a.path_buf[2 * (base + w)] = yr;
            a.path_buf[2 * (base + w) + 1] = xr;
        }
    };
    // preprofile mode: sequence one is the master, sequence two the slave
    int* cnt = nullptr;
    if (a.counts && !below) {
        if (a.seq_cnt_off) {            // dual walks: the master is sequence one of this thread's orientation
            const int64_t o = a.seq_cnt_off[TR ? rid : sid];
            if (o < 0) return;          // not a master: nothing to add
            cnt = a.counts + o;
        } else {
            cnt = a.counts + a.cnt_off[slot];
        }
    }
    const uint8_t* slave = a.seqs ? a.seqs + a.offs[TR ? sid : rid] : nullptr;
    int pend_y = -1, pend_x = 0;
    auto kept = [&](int yr, int xr) {   // (yr, xr) is the first path row with master index yr
        if (cnt && pend_y == yr + 1 && pend_x > xr) atomicAdd(cnt + (int64_t)(pend_y - 1) * a.A + slave[pend_x - 1], 1);
        pend_y = yr;
        pend_x = xr;
    };

    if (a.mode != PG_GLOBAL && !local) {  // extend_path_semiglobal, trailing part (util/align.py:283-295)
        const int ye = TR ? xk : yk, xe = TR ? yk : xk;
        if (ye != L1) { for (int v = L1; v > ye; v--) push(v, xe); }
        else if (xe != L2) { for (int v = L2; v > xe; v--) push(ye, v); }
    }

    for (;;) {
        const int cy = TR ? xk : yk, cx = TR ? yk : xk;   // reference coordinates of this cell
        push(cy, cx);
        int ny = yk, nx = xk;
        bool moved = true;
        bool halt = false;
        if (local && yk >= 1 && xk >= 1) {
#pragma unroll
            for (int b = 0; b < PG_NBOX; b++) halt |= (yk >= bx[b][0] && yk <= bx[b][1] && xk >= bx[b][2] && xk <= bx[b][3]);
            if (!halt && s == 0) halt = (nib_at(yk, xk) & 3u) == 2u;
        }
        if (halt || (yk == 0 && xk == 0)) moved = false;
        else if (xk == 0) {
            if (s == 1 && a.left_ramp) ny--; else moved = false;
        } else if (yk == 0) {
            if (s == 2 && a.top_ramp) nx--; else moved = false;
        } else if (s == 0) {
            ny--; nx--;
            s = code_at(ny, nx);
        } else {
            if (s == 1) { s = ((nib_at(yk, xk) >> 2) & 1u) ? 1 : 0; ny--; }
            else if (fmt == 2) { s = (xk >= 2 && ((nib_at(yk, xk - 1) >> 3) & 1u)) ? 2 : 0; nx--; }
            else        { s = ((nib_at(yk, xk) >> 3) & 1u) ? 2 : 0; nx--; }
        }
        if (!moved) { kept(cy, cx); break; }              // row 0 of the path is always kept
        if ((TR ? nx : ny) == cy - 1) kept(cy, cx);       // the master index steps here
        yk = ny; xk = nx;
    }

    if (local) {
        if (cnt && yk >= 1) atomicAdd(cnt + (int64_t)(yk - 1) * a.A + slave[xk >= 1 ? xk - 1 : Lr - 1], 1);
        if (a.box_out) {
            int32_t* bo = a.box_out + slot * (PG_NBOX * 4) + 4 * a.box_slot;
            bo[0] = yk; bo[1] = end_y; bo[2] = xk; bo[3] = end_x;
        }
    } else if (a.mode != PG_GLOBAL) {  // leading part (util/align.py:270-279): rows first, then columns
        const int y0 = TR ? xk : yk, x0 = TR ? yk : xk;
        if (y0 != 0) { for (int v = y0 - 1; v >= 0; v--) push(v, 0); }
        else if (x0 != 0) { for (int v = x0 - 1; v >= 0; v--) push(0, v); }
    }
    if (want_path) {
        a.path_start[pidx] = w;
        a.path_len[pidx] = cap - w;
    }
}

int pg_launch_traceback(const TraceArgs& a, cudaStream_t st)
{
    if (a.n_slots <= 0) return 0;
    if (a.mode == PG_LOCAL && (a.transposed || a.tb_fmt != 0)) {
        pg_set_error("local walks need the f32 kernel's layout in the reference orientation");
        return 1;
    }
    if (a.dual && (a.tb_fmt != 2 || a.mode != PG_GLOBAL)) {
        pg_set_error("dual walks read the paired-resident kernel's words (global mode)");
        return 1;
    }
    const int64_t n_threads = a.dual ? 2 * a.n_slots : a.n_slots;
    k_traceback<<<(unsigned)((n_threads + 127) / 128), 128, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
