// gotoh_stream16r.cuh -- paired-RESIDENT packed int16 streaming kernel (the all-vs-all shape), score-only
// (TB = false, instantiated in gotoh_stream16.cu) and traced (TB = true, gotoh_stream16t.cu).
//
// The low and high halves of every register carry two RESIDENT sequences (tile.resident, tile.resident2)
// and ONE stream runs through both: the substitution profile is stored pre-packed,
//     prof2[a][x] = { S[a][resA[x]] , S[a][resB[x]] },
// so a single shared-memory row feeds both halves without a per-column PRMT, there is one ring, one set
// of row flags, and the border re-arm is a plain register reload.  In an all-vs-all, residents i and i+1
// share every streamed j > i+1; the pair (i, i+1) itself rides along as the first stream element with its
// high half ignored (tile.b_skip).  Reference recurrence: praline/util/cext.c:99-306.
//
// Step loop.  The DPX instructions (2 VIADDMNMX + 1 VIMNMX3 per column step) run on the half-rate ALU
// pipe, which is the binding unit; everything else is kept OFF that pipe and out of the issue slots:
//  * 8-step blocks come in two forms.  The ring knows 32 steps ahead where sequences end, and a lane
//    is l steps behind lane 0, so "some lane meets a LAST row in this block" is a warp-uniform
//    function of two ballot masks.  Blocks without one (about 90 % at 400 rows per sequence) run
//    straight-line code: no flag test, no border re-arm, no emit, no register shuffling at a join.
//    The other blocks run the general step in a rolled loop.
//  * the strip edge is TWO shuffles: the sender finishes the left-gap recurrence of the receiver's
//    first column (l_out = max(L + ext, M + open) of its last column), so (l_out, D_last) is all that
//    moves; the seed of the diagonal rides in Dleft_prev.
//  * the column-0 substitution in lane 0 (M = L = -inf, D = the U border) is arithmetic on per-lane
//    constants, x * c1 + c2 with c1 = (lane != 0), which ptxas keeps as IMADs on the FMA pipe instead of
//    SELs on the ALU pipe; the border ramp is a packed add of a per-lane constant (0 off lane 0).
//  * the eight ring words of a block arrive as two LDS.128: the ring is kept in four copies shifted by
//    one word each, so that every lane's run of 8 slots is 16-byte aligned in the copy (lane & 3).
//    Ring words are NEGATED row offsets: address = word * all_ones + lane base is one IMAD.
//
// Traced form (TB): ONE fill serves BOTH orientations of a pair.  With a symmetric substitution matrix
// and constant gap penalties the DP values of (sequence_one = s, sequence_two = r) are the transpose of
// those of (r, s) with U and L swapped; what differs is the walker's tie order (MM > MU > ML,
// util/align.py:161-166), i.e. which gap state wins when U and L tie above M.  A cell therefore stores
//     bit 0  U (from above, kernel orientation) is a maximum and M is not      u == max(m + 1, d)
//     bit 1  L (from the left)                 is a maximum and M is not      l == max(m + 1, d)
//     bit 2  U of this cell was OPENED  (M_up + open >= U_up + ext)           u == Mo_up
//     bit 3  L of the NEXT column was opened                                   l_next == M + open
// and each orientation's walk reads its own priority out of bits 0-1 (traceback.cu, tb_fmt 2).  Every test
// compares a value with a maximum it took part in (a <= b), so a + ~b = a - b - 1 is -1 exactly when
// they are equal and <= -2 otherwise: VIADDMNMX(a, ~b, -2) is -1 / -2 in both halves -- one DPX
// instruction per test; t * 2^n + (2^n - 1) (an IMAD, FMA pipe) moves the flag to bit n of both halves
// with ones elsewhere, the four results are ANDed (two LOP3s) and shifted into the column's word by one
// more IMAD: four rows of both halves per 32-bit word, stored [step / 4][k][lane] (coalesced 128-byte
// lines, 0.5 B per cell).  20 instructions per column step (two cells, both orientations: 10 on the ALU
// pipe, 10 on the FMA pipe) against 17 per orientation in k_stream16<TB>.  Complements are
// x * (-1) + (-1) with an opaque -1: IMADs on the FMA pipe.
#pragma once
#include "common.cuh"

#ifndef FLAG_LAST
#define FLAG_LAST 0x80000000u
#define FLAG_EMIT 0x40000000u
#define FULL 0xffffffffu
#endif

__device__ __forceinline__ uint32_t pack2r(int v) { return ((uint32_t)(v & 0xffff)) | ((uint32_t)v << 16); }

template <int K>
__device__ __forceinline__ uint32_t pick_ur(const uint32_t (&v)[K], int k)
{
    uint32_t r = v[0];
#pragma unroll
    for (int i = 1; i < K; i++) r = (k == i) ? v[i] : r;
    return r;
}

struct K16rConst {
    uint32_t go2, ge2, NEG2, c1, c2, left1z, lanebase, ones;
    uint32_t one2, cm2, x2, x4, x8;         // traced: +1 and the clamp -2, packed; 2, 4, 8 opaque to ptxas (IMADs stay on the FMA pipe)
    uint32_t x16;                           // 16, opaque to ptxas: acc * 16 + q stays ONE IMAD (FMA pipe) instead of SHL + LOP3
};

// l_src / D_src: the edge values this lane's left neighbour produced for the row this lane works on now
// (its previous step with a lane skew of 1, two steps back with a skew of 2).
// TB: acc[k] collects the flag nibbles of column k; `first` starts a new word (the first of four rows).
template <int K, bool TB>
__device__ __forceinline__ void k16r_step(uint32_t w, const K16rConst& c, uint32_t (&Mo)[K], uint32_t (&U)[K], uint32_t (&D)[K],
                                          uint32_t l_src, uint32_t D_src, uint32_t& l_out, uint32_t& D_last,
                                          uint32_t& Dleft_prev, uint32_t& bordz, uint32_t (&acc)[TB ? K : 1], bool first)
{
    constexpr int NCH = (K + 3) / 4;
    const uint32_t pa = w * c.ones + c.lanebase;
    uint32_t sc[NCH * 4];
#pragma unroll
    for (int j = 0; j < NCH; j++) {
        if (4 * j + 1 >= K)         // a chunk that holds one column only
            asm("ld.shared.u32 %0, [%1];" : "=r"(sc[4 * j]) : "r"(pa + (uint32_t)j * 512u));
        else
            asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                : "=r"(sc[4 * j]), "=r"(sc[4 * j + 1]), "=r"(sc[4 * j + 2]), "=r"(sc[4 * j + 3])
                : "r"(pa + (uint32_t)j * 512u));
    }
    uint32_t l = __shfl_up_sync(FULL, l_src, 1) * c.c1 + c.c2;       // lane 0: L(y, 1) of an empty left side
    const uint32_t Dn = __shfl_up_sync(FULL, D_src, 1) * c.c1 + bordz;   // lane 0: the U border of this row
    bordz = __vadd2(bordz, c.left1z);
    uint32_t diag = Dleft_prev;
    Dleft_prev = Dn;
#pragma unroll
    for (int k = 0; k < K; k++) {
        const uint32_t m = __vadd2(diag, sc[k]);
        const uint32_t u = __viaddmax_s16x2(U[k], c.ge2, Mo[k]);
        diag = D[k];
        const uint32_t d = __vimax3_s16x2(m, u, l);
        const uint32_t mo = __vadd2(m, c.go2);
        const uint32_t ln = __viaddmax_s16x2(l, c.ge2, mo);       // L of the next column (the receiver's first one after k = K-1)
        if (TB) {
            const uint32_t dp = __viaddmax_s16x2(m, c.one2, d);    // max(m + 1, d): equal to a gap state only where M lost
            const uint32_t ndp = dp * c.ones + c.ones, nu = u * c.ones + c.ones, nln = ln * c.ones + c.ones;
            // -1 (equal) or -2 in both halves; t * 2^n + (2^n - 1) moves the flag to bit n of both halves and keeps
            // ones below it (32-bit IMAD: the carry out of the low half is exactly the high half's low ones)
            const uint32_t tU = __viaddmax_s16x2(u, ndp, c.cm2);
            const uint32_t tL = __viaddmax_s16x2(l, ndp, c.cm2) * c.x2 + 1u;
            const uint32_t tO = __viaddmax_s16x2(Mo[k], nu, c.cm2) * c.x4 + 3u;
            const uint32_t tN = __viaddmax_s16x2(mo, nln, c.cm2) * c.x8 + 7u;
            const uint32_t q = (tU & tL & tO) & tN & 0x000f000fu;
            acc[k] = first ? q : acc[k] * c.x16 + q;
        }
        l = ln;
        Mo[k] = mo;
        U[k] = u;
        D[k] = d;
    }
    l_out = l;
    D_last = D[K - 1];
}

// CTAs per SM the register allocation aims at.  Score-only: 24 warps per SM (80 registers) up to 13 columns
// per lane -- measured 6.94 -> 7.21 TCUPS at K = 13 against 16 warps with 121 registers --, 16 warps up to
// K = 24, and one CTA with the full register file for K = 32 (96 state registers).  Traced: one flag word
// per column more, 16 warps (128 registers) up to K = 14, else one CTA.
#ifndef K16R_MINB
#define K16R_MINB(K) ((K) <= 13 ? 3 : ((K) <= 24 ? 2 : 1))
#endif
#ifndef K16RT_MINB
#define K16RT_MINB(K) ((K) <= 14 ? 2 : 1)
#endif
// Lane skew: lane l works on stream row t - SKEW * l at step t.  With a skew of 2 the edge a lane needs
// was produced by its neighbour TWO steps earlier, so the left-gap chains of consecutive steps of a warp do
// not depend on each other (the shuffle of step t+1 can be issued before step t ends): two chains in flight
// per warp instead of one, at the price of a 62-step pipeline fill per warp and a longer ring.
#ifndef K16R_SKEW
#define K16R_SKEW 1
#endif
template <int K, int NW, bool TB>
__device__ __forceinline__ void k16r_body(const StreamArgs& a)
{
    constexpr int NCH = (K + 3) / 4;        // 16-byte chunks of 4 packed columns per lane
    constexpr int ROWB = NCH * 512;
    constexpr int KP = (K + 3) & ~3;
    constexpr int SKEW = TB ? 1 : K16R_SKEW;
    constexpr int RN = 64 * SKEW;           // ring slots: a window of 32 new slots + 31 * SKEW slots of lookback
    constexpr int RL = RN + 8;              // words per ring copy: the slots + the first 8 again
    constexpr int RW = 4 * RL + RN;         // per warp: four shifted copies of the offsets + the flag ring
    extern __shared__ __align__(16) unsigned char smem[];
    uint32_t* prof = reinterpret_cast<uint32_t*>(smem);
    uint32_t* top2 = reinterpret_cast<uint32_t*>(smem + (size_t)a.A * ROWB);            // [32][KP + 4]
    uint32_t* ring = top2 + 32 * (KP + 4) + (threadIdx.x >> 5) * RW;
    uint32_t* fring = ring + 4 * RL;
    // traced: per-warp staging block of one traceback word row, [lane][KS] (KS odd: conflict-free both ways)
    constexpr int KS = K | 1;
    uint32_t* tstage = top2 + 32 * (KP + 4) + NW * RW + (threadIdx.x >> 5) * (32 * KS);

    const PgTile tile = a.tiles[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t roffA = a.offs[tile.resident];
    const int LrA = (int)(a.offs[tile.resident + 1] - roffA);
    const bool hasB = tile.resident2 >= 0;
    const int64_t roffB = hasB ? a.offs[tile.resident2] : 0;
    const int LrB = hasB ? (int)(a.offs[tile.resident2 + 1] - roffB) : 1;

    {
        const int n = a.A * NCH * 128;
        for (int idx = threadIdx.x; idx < n; idx += NW * 32) {
            const int c = idx & 3, l = (idx >> 2) & 31, j = (idx >> 7) % NCH, sym = idx / (NCH * 128);
            const int k = 4 * j + c, x = l * K + k;
            int va = 0, vb = 0;
            if (k < K && x < LrA) {
                const int b = a.seqs[roffA + x];
                va = (int)(a.transposed ? a.S[b * a.A + sym] : a.S[sym * a.A + b]);
            }
            if (hasB && k < K && x < LrB) {
                const int b = a.seqs[roffB + x];
                vb = (int)(a.transposed ? a.S[b * a.A + sym] : a.S[sym * a.A + b]);
            }
            prof[idx] = ((uint32_t)(va & 0xffff)) | ((uint32_t)vb << 16);
        }
        for (int idx = threadIdx.x; idx < 32 * (KP + 4); idx += NW * 32) {
            const int l = idx / (KP + 4), k = idx % (KP + 4);
            int v = 0;
            if (k < K) v = (int)a.topD[l * K + k + 1];
            else if (k == KP) v = (int)a.topD[l * K];
            top2[idx] = pack2r(v);
        }
    }
    const uint32_t prof_s = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = lane; i < RW; i += 32) ring[i] = 0u;
    __syncthreads();
    if (prof_s & 15u) __trap();

    const int n_str = tile.stream_end - tile.stream_begin;
    const int per = (n_str + NW - 1) / NW;
    const int sb = tile.stream_begin + warp * per;
    const int se = min(sb + per, tile.stream_end);
    if (sb >= se) return;

    auto seq_id = [&](int s) -> int { return a.stream_ids ? a.stream_ids[s] : s; };
    auto seq_len = [&](int s) -> int {
        if (s < sb) return 1;
        const int id = seq_id(s);
        return (int)(a.offs[id + 1] - a.offs[id]);
    };
    int total = 0;
    for (int s = sb + lane; s < se; s += 32) total += seq_len(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
    const int T = (total + 1 + 31 * SKEW + 31) & ~31;

    const int lrA = (LrA - 1) / K, klA = (LrA - 1) % K, lrB = (LrB - 1) / K, klB = (LrB - 1) % K;
    K16rConst c;
    c.ones = (uint32_t)a.all_ones;                       // 0xffffffff, opaque to the compiler
    c.go2 = pack2r(a.go16);
    c.ge2 = pack2r(a.ge16);
    c.NEG2 = pack2r(a.neg16);
    c.c1 = (lane != 0) ? (0u - c.ones) : 0u;             // 1 / 0
    c.c2 = (lane != 0) ? 0u : __viaddmax_s16x2(c.NEG2, c.ge2, c.NEG2);
    c.left1z = (lane != 0) ? 0u : pack2r(a.left1_16);
    c.lanebase = prof_s + ((uint32_t)lane << 4);
    c.one2 = 0x00010001u;
    c.cm2 = pack2r(-2);
    c.x2 = (uint32_t)a.all_ones & 2u;
    c.x4 = (uint32_t)a.all_ones & 4u;
    c.x8 = (uint32_t)a.all_ones & 8u;
    c.x16 = (uint32_t)a.all_ones & 16u;
    const uint32_t left0z = (lane != 0) ? 0u : pack2r(a.left0_16);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    // slot (t - SKEW * lane) of an aligned block start t, in copy (SKEW * lane) & 3: position (t - (SKEW * lane & ~3)) & (RN - 1)
    const int back = SKEW * lane;
    const uint32_t ring_mine = ring_s + (uint32_t)(back & 3) * (RL * 4u);
    const uint32_t* mytop = top2 + lane * (KP + 4);

    uint32_t Mo[K], U[K], D[K];
    uint32_t acc[TB ? K : 1];
#pragma unroll
    for (int k = 0; k < K; k++) { Mo[k] = 0u; U[k] = 0u; D[k] = 0u; }
#pragma unroll
    for (int k = 0; k < (TB ? K : 1); k++) acc[k] = 0u;
    const int64_t tbw0 = TB ? a.tb_base[(int64_t)blockIdx.x * NW + warp] : 0;
    uint32_t l_out = 0u, D_last = 0u, Dleft_prev = 0u, bordz = 0u;
    uint32_t l_out1 = 0u, D_last1 = 0u;      // SKEW 2: the edge of the step before the last one
    int q = sb;
    int ps = sb - 1, pp = 0;
    uint32_t last_prev = 0u, last_prev2 = 0u;   // LAST slots of the previous 32-slot windows (bit j = slot t0 - 32 + j, t0 - 64 + j)

    for (int t0 = 0; t0 < T; t0 += 32) {
        uint32_t last_cur;
        {
            int s = ps, p = pp + lane;
            int len = (s < se) ? seq_len(s) : 0;
            while (s < se && p >= len) {
                p -= len;
                s++;
                len = (s < se) ? seq_len(s) : 0;
            }
            uint32_t off = 0u, flags = 0u;
            if (s < se) {
                if (s < sb) {
                    flags = FLAG_LAST;
                } else {
                    const int sym = a.seqs[a.offs[seq_id(s)] + p];
                    off = (uint32_t)(sym * ROWB);
                    flags = (p == len - 1) ? (FLAG_LAST | FLAG_EMIT) : 0u;
                }
            }
            last_cur = __ballot_sync(FULL, flags != 0u);
            __syncwarp();
            const uint32_t word = 0u - off;
#pragma unroll
            for (int cpy = 0; cpy < 4; cpy++) {
                const int pos = (t0 + lane + cpy) & (RN - 1);
                ring[cpy * RL + pos] = word;
                if (pos < 8) ring[cpy * RL + RN + pos] = word;
            }
            fring[(t0 + lane) & (RN - 1)] = flags;
            __syncwarp();
            int s31 = __shfl_sync(FULL, s, 31), p31 = __shfl_sync(FULL, p, 31) + 1;
            const int len31 = __shfl_sync(FULL, len, 31);
            if (s31 < se && p31 >= len31) { p31 = 0; s31++; }
            ps = s31;
            pp = p31;
        }

        // block length: 8 steps score-only; 4 traced (one traceback word per column and block; the straight-line
        // code of 8 traced steps is 33 KB and ran with 25 % of its stall samples on instruction fetch)
        constexpr int BLK = TB ? 4 : 8;
#pragma unroll 1
        for (int g = 0; g < 32; g += BLK) {
            // a LAST slot s is met by lane l at step s + SKEW * l: the block [t0+g, t0+g+BLK) is free of them
            // iff no LAST slot lies in [t0+g - 31*SKEW, t0+g+BLK-1]
            const uint32_t busy = (last_cur & (0xffffffffu >> (32 - BLK - g))) |
                                  (SKEW == 1 ? (last_prev >> (g + 1)) : (last_prev | (last_prev2 >> (g + 2))));
            const uint32_t rp = ring_mine + ((uint32_t)((t0 + g - (back & ~3)) & (RN - 1)) << 2);
            uint32_t* tbdst = TB ? a.tb + tbw0 + (int64_t)((t0 + g) >> 2) * (KS * 32) + lane : nullptr;
            if (busy == 0u) {
                uint32_t w[8];
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(rp) : "memory");
                if (BLK == 8)
                    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                 : "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(rp + 16u) : "memory");
#pragma unroll
                for (int i = 0; i < BLK; i++) {
                    if (SKEW == 1) {
                        k16r_step<K, TB>(w[i], c, Mo, U, D, l_out, D_last, l_out, D_last, Dleft_prev, bordz, acc, (i & 3) == 0);
                    } else {
                        const uint32_t ls = l_out1, ds = D_last1;
                        l_out1 = l_out;
                        D_last1 = D_last;
                        k16r_step<K, TB>(w[i], c, Mo, U, D, ls, ds, l_out, D_last, Dleft_prev, bordz, acc, (i & 3) == 0);
                    }
                    if (TB && (i & 3) == 3) {
                        // The walk moves along diagonals: (y - 1, x - 1) is the NEXT COLUMN of the same lane.  Words are
                        // therefore stored [step / 4][lane][k] -- neighbouring columns in one 32-byte sector -- instead
                        // of [k][lane] (a new 128-byte line per walker step: the walk ran at the random-access limit of
                        // HBM and slowed the concurrent fill down).  The warp transposes the word row through its
                        // staging block: K conflict-free STS, then KS coalesced 128-byte lines.
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < K; k++) tstage[lane * KS + k] = acc[k];
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < KS; j++) tbdst[(i >> 2) * (KS * 32) + j * 32] = tstage[j * 32 + lane];
                    }
                }
            } else {
#pragma unroll 1
                for (int i = 0; i < BLK; i++) {
                    uint32_t w;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(rp + (uint32_t)i * 4u) : "memory");
                    const uint32_t f = fring[(t0 + g + i - back) & (RN - 1)];
                    if (SKEW == 1) {
                        k16r_step<K, TB>(w, c, Mo, U, D, l_out, D_last, l_out, D_last, Dleft_prev, bordz, acc, (i & 3) == 0);
                    } else {
                        const uint32_t ls = l_out1, ds = D_last1;
                        l_out1 = l_out;
                        D_last1 = D_last;
                        k16r_step<K, TB>(w, c, Mo, U, D, ls, ds, l_out, D_last, Dleft_prev, bordz, acc, (i & 3) == 0);
                    }
                    if (TB && (i & 3) == 3) {
                        // The walk moves along diagonals: (y - 1, x - 1) is the NEXT COLUMN of the same lane.  Words are
                        // therefore stored [step / 4][lane][k] -- neighbouring columns in one 32-byte sector -- instead
                        // of [k][lane] (a new 128-byte line per walker step: the walk ran at the random-access limit of
                        // HBM and slowed the concurrent fill down).  The warp transposes the word row through its
                        // staging block: K conflict-free STS, then KS coalesced 128-byte lines.
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < K; k++) tstage[lane * KS + k] = acc[k];
                        __syncwarp();
#pragma unroll
                        for (int j = 0; j < KS; j++) tbdst[(i >> 2) * (KS * 32) + j * 32] = tstage[j * 32 + lane];
                    }
                    if (f & FLAG_LAST) {
                        if (f & FLAG_EMIT) {
                            const int e = q - tile.stream_begin;
                            const int sstr = tile.slot_stride ? tile.slot_stride : 1;      // 2: the slots of A and B interleave
                            if (lane == lrA) {
                                const int64_t slot = tile.out_base + (int64_t)e * sstr;
                                a.scores[slot] = (float)(int)(int16_t)(pick_ur<K>(D, klA) & 0xffffu);
                                if (TB) {
                                    a.emit_t[slot] = t0 + g + i;
                                    a.pair_tb[slot] = tbw0;
                                    a.slot_res[slot] = tile.resident;
                                    a.slot_str[slot] = seq_id(q);
                                }
                            }
                            if (hasB && lane == lrB && e >= tile.b_skip) {
                                const int64_t slot = tile.out_base2 + (int64_t)(e - tile.b_skip) * sstr;
                                a.scores[slot] = (float)(int)(int16_t)(pick_ur<K>(D, klB) >> 16);
                                if (TB) {
                                    a.emit_t[slot] = (t0 + g + i) | (1 << 30);      // bit 30: the high register half
                                    a.pair_tb[slot] = tbw0;
                                    a.slot_res[slot] = tile.resident2;
                                    a.slot_str[slot] = seq_id(q);
                                }
                            }
                            q++;
                        }
#pragma unroll
                        for (int k = 0; k < K; k++) {
                            Mo[k] = c.NEG2;
                            U[k] = c.NEG2;
                            D[k] = mytop[k];
                        }
                        Dleft_prev = mytop[KP];
                        bordz = left0z;
                    }
                }
            }
        }
        last_prev2 = last_prev;
        last_prev = last_cur;
    }
}

// Score-only kernel: launch bounds as measured in round 2.
template <int K, int NW, bool TB>
__global__ void __launch_bounds__(NW * 32, TB ? K16RT_MINB(K) : K16R_MINB(K)) k_stream16r(const StreamArgs a)
{
    k16r_body<K, NW, TB>(a);
}

// Traced kernel: capped at 112 registers up to K = 13 (no spills), so that two CTAs leave 8,192 registers of an SM
// free -- exactly one 128-thread block of the walk (k_traceback, 64 registers): the walk of wave w then runs in the
// leftover slots WHILE wave w + 1 fills, instead of taking turns with it (a __launch_bounds__ kernel and
// __maxnreg__ cannot be combined, hence the second entry point).
#ifndef K16RT_MAXREG
#define K16RT_MAXREG(K) ((K) <= 13 ? 112 : ((K) <= 14 ? 128 : 255))
#endif
template <int K, int NW>
__global__ void __maxnreg__(K16RT_MAXREG(K)) k_stream16rt(const StreamArgs a)
{
    k16r_body<K, NW, true>(a);
}
