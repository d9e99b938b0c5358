// gotoh_stream.cuh -- the k_stream kernel template of K2 (see gotoh_stream.cu for the design notes);
// instantiated by gotoh_stream.cu (score-only, global / semiglobal traced, matrix-fed) and
// gotoh_stream_local.cu (local traced, with and without Waterman-Eggert boxes).
#pragma once
#include "common.cuh"

#define FLAG_LAST 0x80000000u
#define FLAG_EMIT 0x40000000u
#define FULL 0xffffffffu


__device__ __forceinline__ unsigned long long pack_key(float v, int idx)
{
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (uint32_t)idx;
}

template <int K>
__device__ __forceinline__ float pick(const float (&v)[K], int k)
{
    float r = v[0];
#pragma unroll
    for (int i = 1; i < K; i++) r = (k == i) ? v[i] : r;
    return r;
}

// KM: 0 = global, 1 = local (score only), 2 = semiglobal.  TB: write packed traceback.
// TR: traceback tie order for a transposed (resident = sequence one) launch.
// MS: match scores come from a materialised matrix in HBM (profile x profile batches, rows in
//     stream order, 32*K floats per row) instead of the shared-memory substitution profile.
// MK: Waterman-Eggert boxes (preprofile.py:227-267): up to PG_NBOX rectangles of masked cells per
//     pair; a masked cell keeps M = U = L = 0 like the reference's `continue` (cext.c:143-148).
//
// Local traced launches (KM == 1 && TB; reference orientation only).  M = max(0, max3 + s); the
// walker must know where the alignment STOPS: a cell whose three sums are all negative has no flag
// (cext.c:214-243).  With gap penalties <= 0 an M cell reached through a gap opening is positive,
// so the stop test is only ever read where M is the first-priority maximum of its cell, and there
// the "second gap state" bit is free: bit 1 = sign(max3 + s) when bit 0 says M.  The end cell is
// the FIRST row-major maximum of the whole matrix (np.argmax, align.py:401-402): every lane keeps
// (best, y, x) with strict '>' row by row and the lanes meet in one 64-bit atomicMax key
// (value, ~(y << 11 | x)).
template <int K, int KM, bool TB, bool TR, bool MS, bool MK, int NW>
__global__ void __launch_bounds__(NW * 32) k_stream(const StreamArgs a)
{
    constexpr bool LT = (KM == 1) && TB;
    constexpr int UNR = 8;            // steps unrolled per inner iteration (8 = one traceback word)
    constexpr int NCH = (K + 3) / 4;
    constexpr int ROWB = NCH * 512;
    extern __shared__ __align__(16) unsigned char smem[];
    float* prof = reinterpret_cast<float*>(smem);
    uint32_t* ring = reinterpret_cast<uint32_t*>(smem + (MS ? 0 : (size_t)a.A * ROWB)) + (threadIdx.x >> 5) * 128;

    const PgTile tile = a.tiles[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t roff = a.offs[tile.resident];
    const int Lr = (int)(a.offs[tile.resident + 1] - roff);

    // ---- substitution profile of the resident, [a][chunk][lane][4] ------------------------
    if (!MS) {
        const float padv = (KM == 1) ? -INFINITY : 0.f;
        const int n = a.A * NCH * 128;
        for (int idx = threadIdx.x; idx < n; idx += NW * 32) {
            const int c = idx & 3, l = (idx >> 2) & 31, j = (idx >> 7) % NCH, sym = idx / (NCH * 128);
            const int k = 4 * j + c, x = l * K + k;
            float v = padv;
            if (k < K && x < Lr) {
                const int b = a.seqs[roff + x];
                v = a.transposed ? a.S[b * a.A + sym] : a.S[sym * a.A + b];
            }
            prof[idx] = v;
        }
    }
    for (int i = lane; i < 128; i += 32) ring[i] = MS ? 0u : (uint32_t)__cvta_generic_to_shared(smem);
    __syncthreads();

    // ---- this warp's slice of the tile's stream --------------------------------------------
    const int n_str = tile.stream_end - tile.stream_begin;
    const int per = (n_str + NW - 1) / NW;
    const int sb = tile.stream_begin + warp * per;
    const int se = min(sb + per, tile.stream_end);
    if (sb >= se) return;

    auto seq_id = [&](int s) -> int { return a.stream_ids ? a.stream_ids[s] : s; };
    auto seq_len = [&](int s) -> int {
        if (s < sb) return 1;  // the dummy row that arms the first reset
        const int id = seq_id(s);
        return (int)(a.offs[id + 1] - a.offs[id]);
    };

    int total = 0;
    for (int s = sb + lane; s < se; s += 32) total += seq_len(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(FULL, total, o);
    const int T = (total + 1 + 31 + 31) & ~31;   // dummy row + stream + drain, whole 32-step blocks

    const int lr = (Lr - 1) / K, klast = (Lr - 1) % K;
    const float go = a.go, ge = a.ge;
    const float left0 = a.left0, left1 = a.left1;
    const int64_t tbw0 = TB ? a.tb_base[(int64_t)blockIdx.x * NW + warp] : 0;
    const float* mwarp = MS ? a.mwave + a.mrow_base[(int64_t)blockIdx.x * NW + warp] * (32 * K) : nullptr;
    const uint32_t prof_s = (uint32_t)__cvta_generic_to_shared(smem);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const uint32_t lane16 = (uint32_t)lane << 4;
    if (!MS && (prof_s & 511u)) __trap();   // row addresses are OR-ed with the lane offset below

    // matrix-fed launches: the aligned 128-bit pieces that hold a lane's K scores of one step, fetched one step ahead
    constexpr int NVX = MS ? ((K % 4 == 0) ? K / 4 : (K + 3 + 3) / 4) : 1;
    // fetched one step ahead where the extra registers do not cost a resident CTA (measured: K = 13 1.64 -> 1.48 ms
    // per wave; K = 10 would drop from three CTAs per SM to two, 1.22 -> 1.33 ms, and keeps the direct loads)
    constexpr bool PF = MS && K >= 12;
    float4 pre[NVX];
#pragma unroll
    for (int j = 0; j < NVX; j++) pre[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto ms_fetch = [&](uint32_t wv) {
        const float* mr = mwarp + (size_t)(wv & 0x00ffffffu) * (32 * K) + lane * K;
        const int mis = (K % 4 == 0) ? 0 : ((lane * K) & 3);
        const float4* vb = reinterpret_cast<const float4*>(mr - mis);
#pragma unroll
        for (int j = 0; j < NVX; j++) pre[j] = __ldg(vb + j);
    };
    float Mo[K], U[K], D[K];
    uint32_t acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) { Mo[k] = 0.f; U[k] = 0.f; D[k] = 0.f; acc[k] = 0u; }
    float Mo_last = 0.f, L_last = 0.f, D_last = 0.f, Dleft_prev = 0.f;
    float best = 0.f, colbest = 0.f, yf = 0.f;   // yf = rows of the current sequence already done
    int colbest_y = 0, q = sb;
    int best_y = 0, best_x = 0;                  // local traced: first row-major cell holding `best`
    float bylo[PG_NBOX], byhi[PG_NBOX];          // MK: row range of every box, column bits of this lane
    uint32_t cmask[PG_NBOX];
#pragma unroll
    for (int b = 0; b < PG_NBOX; b++) { bylo[b] = 1.f; byhi[b] = 0.f; cmask[b] = 0u; }
    int ps = sb - 1, pp = 0;  // producer cursor: stream element / offset of stream position t0

    for (int t0 = 0; t0 < T; t0 += 32) {
        // ---- decode the next 32 stream positions into the ring (kept twice, 64 apart, so that
        //      the 32 reads below never wrap) -------------------------------------------------
        {
            int s = ps, p = pp + lane;
            int len = (s < se) ? seq_len(s) : 0;
            while (s < se && p >= len) {
                p -= len;
                s++;
                len = (s < se) ? seq_len(s) : 0;
            }
            uint32_t word = MS ? 0u : prof_s;
            if (s < se) {
                if (s < sb) {
                    word |= FLAG_LAST;
                } else if (MS) {   // the matrix row of stream position t0 + lane is that position itself
                    word = (uint32_t)(t0 + lane) | ((p == len - 1) ? (FLAG_LAST | FLAG_EMIT) : 0u);
                } else {
                    const int sym = a.seqs[a.offs[seq_id(s)] + p];
                    word = (prof_s + (uint32_t)(sym * ROWB)) | ((p == len - 1) ? (FLAG_LAST | FLAG_EMIT) : 0u);
                }
            }
            __syncwarp();
            ring[(t0 + lane) & 63] = word;
            ring[((t0 + lane) & 63) + 64] = word;
            __syncwarp();
            int s31 = __shfl_sync(FULL, s, 31), p31 = __shfl_sync(FULL, p, 31) + 1;
            const int len31 = __shfl_sync(FULL, len, 31);
            if (s31 < se && p31 >= len31) { p31 = 0; s31++; }
            ps = s31;
            pp = p31;
        }
        const uint32_t rp0 = ring_s + ((uint32_t)((t0 - lane) & 63) << 2);

#pragma unroll 1
        for (int g = 0; g < 32; g += UNR) {
            const uint32_t rp = rp0 + (uint32_t)g * 4u;
            if (PF && g == 0) {     // first step of a 32-step block: its ring entry has only just been decoded
                uint32_t w0;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(rp) : "memory");
                ms_fetch(w0);
            }
#pragma unroll
            for (int i = 0; i < UNR; i++) {
                const int t = t0 + g + i;
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(rp + (uint32_t)i * 4u) : "memory");
                float sc[NCH * 4];
                if (MS) {
                    // this step's vectors were requested one step ago; request the next step's now (the ring
                    // already holds its row) so that the HBM / L2 latency overlaps this step's cells
                    if (!PF) ms_fetch(w);
                    float4 cur[NVX];
#pragma unroll
                    for (int j = 0; j < NVX; j++) cur[j] = pre[j];
                    if (PF && (i + 1 < UNR || g + UNR < 32)) {
                        uint32_t wn;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wn) : "r"(rp + (uint32_t)(i + 1) * 4u) : "memory");
                        ms_fetch(wn);
                    }
                    if (K % 4 == 0) {
#pragma unroll
                        for (int j = 0; j < K / 4; j++) {
                            sc[4 * j] = cur[j].x; sc[4 * j + 1] = cur[j].y; sc[4 * j + 2] = cur[j].z; sc[4 * j + 3] = cur[j].w;
                        }
                    } else {
                        // K floats at a 4-byte-aligned offset: aligned 128-bit loads that cover them for any
                        // misalignment (0..3 floats, constant per lane), then a two-stage shift by selects --
                        // NVX requests per step instead of K scalar ones (the row pitch is a multiple of 16 B;
                        // the buffer carries 16 B of slack behind its last row)
                        const int mis = (lane * K) & 3;
                        float buf[NVX * 4 + 3];
#pragma unroll
                        for (int j = 0; j < NVX; j++) {
                            buf[4 * j] = cur[j].x; buf[4 * j + 1] = cur[j].y; buf[4 * j + 2] = cur[j].z; buf[4 * j + 3] = cur[j].w;
                        }
                        buf[NVX * 4] = buf[NVX * 4 + 1] = buf[NVX * 4 + 2] = 0.f;
#pragma unroll
                        for (int k = 0; k < NVX * 4 - 1; k++) buf[k] = (mis & 1) ? buf[k + 1] : buf[k];
#pragma unroll
                        for (int k = 0; k < NVX * 4 - 2; k++) buf[k] = (mis & 2) ? buf[k + 2] : buf[k];
#pragma unroll
                        for (int k = 0; k < K; k++) sc[k] = buf[k];
                    }
                } else {
                    const uint32_t pa = (w & 0x00ffffffu) | lane16;
#pragma unroll
                    for (int j = 0; j < NCH; j++)
                        asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                            : "=f"(sc[4 * j]), "=f"(sc[4 * j + 1]), "=f"(sc[4 * j + 2]), "=f"(sc[4 * j + 3])
                            : "r"(pa + (uint32_t)j * 512u));
                }

                // ---- strip edge from the left lane (its results of the previous step = my row)
                float Ml = __shfl_up_sync(FULL, Mo_last, 1);
                float Ll = __shfl_up_sync(FULL, L_last, 1);
                float Dn = __shfl_up_sync(FULL, D_last, 1);
                if (lane == 0) {   // column 0: M = L = -inf, max3 = the U border (align.py:371-377)
                    Ml = -INFINITY;
                    Ll = -INFINITY;
                    Dn = fmaf(yf, left1, left0);
                }
                yf += 1.f;
                float diag = Dleft_prev;
                Dleft_prev = Dn;
                float rb = 0.f;
                uint32_t rowmask = 0u;
                if (MK) {
#pragma unroll
                    for (int b = 0; b < PG_NBOX; b++) rowmask |= (yf >= bylo[b] && yf <= byhi[b]) ? cmask[b] : 0u;
                }

#pragma unroll
                for (int k = 0; k < K; k++) {
                    float m = diag + sc[k];
                    const float mraw = m;
                    if (KM == 1) m = fmaxf(m, 0.f);
                    const float ue = U[k] + ge;
                    const float le = Ll + ge;
                    float u = fmaxf(Mo[k], ue);
                    float l = fmaxf(Ml, le);
                    if (MK) {
                        const bool z = (rowmask >> k) & 1u;
                        m = z ? 0.f : m;
                        u = z ? 0.f : u;
                        l = z ? 0.f : l;
                    }
                    if (KM == 1) { if (LT) rb = fmaxf(rb, m); else best = fmaxf(best, m); }
                    diag = D[k];
                    float d;
                    if (LT) {
                        const float ul = fmaxf(u, l);
                        const float nm = m - ul;
                        d = fmaxf(m, ul);
                        uint32_t w4 = __funnelshift_l(__float_as_uint(Ml - le), acc[k], 1);      // L: extend
                        w4 = __funnelshift_l(__float_as_uint(Mo[k] - ue), w4, 1);               // U: extend
                        // second gap state where a gap state wins, else the stop test of M
                        const uint32_t sel = (__float_as_int(nm) < 0) ? __float_as_uint(u - l) : __float_as_uint(mraw);
                        w4 = __funnelshift_l(sel, w4, 1);
                        acc[k] = __funnelshift_l(__float_as_uint(nm), w4, 1);                   // not M
                    } else if (TB) {
                        // four sign bits per cell, shifted straight into the column's word: 1 = the
                        // SECOND operand won (strictly), so ties keep the reference's priority
                        // (open before extend, M before U before L; util/align.py:161-174)
                        const float ul = fmaxf(u, l);
                        d = fmaxf(m, ul);
                        uint32_t w4 = __funnelshift_l(__float_as_uint(Ml - le), acc[k], 1);      // L: extend
                        w4 = __funnelshift_l(__float_as_uint(Mo[k] - ue), w4, 1);               // U: extend
                        w4 = __funnelshift_l(__float_as_uint(TR ? (l - u) : (u - l)), w4, 1);   // second gap state
                        acc[k] = __funnelshift_l(__float_as_uint(m - ul), w4, 1);               // not M
                    } else {
                        d = fmaxf(fmaxf(m, u), l);
                    }
                    const float mo = m + go;
                    Mo[k] = mo;
                    U[k] = u;
                    D[k] = d;
                    Ml = mo;
                    Ll = l;
                }
                Mo_last = Ml;
                L_last = Ll;
                D_last = D[K - 1];
                if (LT && rb > best) {   // strict: the earliest row keeps a value; first column of the row
                    best = rb;
                    const float tgt = rb + go;   // M + open of the cell that holds rb (exact: integer scores)
                    int kk = 0;
#pragma unroll
                    for (int k = K - 1; k >= 0; k--) kk = (Mo[k] == tgt) ? k : kk;
                    best_y = (int)yf;
                    best_x = lane * K + kk + 1;
                }

                if (TB && (i & 7) == 7) {
                    uint32_t* dst = a.tb + tbw0 + (int64_t)(t >> 3) * (K * 32) + lane;
#pragma unroll
                    for (int k = 0; k < K; k++) dst[k * 32] = acc[k];
                }
                if (KM == 2 && lane == lr) {
                    const float dl = pick<K>(D, klast);
                    if (dl >= colbest) { colbest = dl; colbest_y = (int)yf; }
                }

                if ((int)w < 0) {   // FLAG_LAST
                    if (w & FLAG_EMIT) {
                        const int64_t slot = tile.out_base + (q - tile.stream_begin);
                        if (KM == 0) {
                            if (lane == lr) a.scores[slot] = pick<K>(D, klast);
                        } else if (KM == 1) {
                            atomicMax(a.rowkey + slot, pack_key(best, LT ? (int)~(((uint32_t)best_y << 11) | (uint32_t)best_x) : 0));
                            if (lane == lr) a.colkey[slot] = 1ull;   // marks the slot as produced
                        } else {
                            float bv = -INFINITY;
                            int bx = -1;
#pragma unroll
                            for (int k = 0; k < K; k++) {
                                const int x = lane * K + k + 1;
                                if (x <= Lr && D[k] >= bv) { bv = D[k]; bx = x; }
                            }
                            if (lane == 0) {
                                const float v = fmaf(yf - 1.f, left1, left0);   // D(y, 0) of this last row
                                if (v > bv) { bv = v; bx = 0; }
                            }
                            if (bx >= 0) atomicMax(a.rowkey + slot, pack_key(bv, bx));
                            if (lane == lr) a.colkey[slot] = pack_key(colbest, colbest_y);
                        }
                        if (TB && lane == lr) {
                            a.emit_t[slot] = t;
                            a.pair_tb[slot] = tbw0;
                        }
                        q++;
                    }
                    // re-arm the top border for the next streamed sequence
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        Mo[k] = -INFINITY;
                        U[k] = -INFINITY;
                        D[k] = a.topD[lane * K + k + 1];
                    }
                    Dleft_prev = a.topD[lane * K];
                    yf = 0.f;
                    best = 0.f;
                    best_y = 0;
                    best_x = 0;
                    if (MK) {   // boxes of the pair that starts now, columns as bits of this lane's K
                        const bool live = q < se;
                        const int4* bx = a.boxes + (tile.out_base + (q - tile.stream_begin)) * PG_NBOX;
#pragma unroll
                        for (int b = 0; b < PG_NBOX; b++) {
                            const int4 r = live ? __ldg(bx + b) : make_int4(1, 0, 1, 0);   // ylo, yhi, xlo, xhi
                            bylo[b] = (float)r.x;
                            byhi[b] = (float)r.y;
                            const int lo = max(r.z - (lane * K + 1), 0), hi = min(r.w - (lane * K + 1), K - 1);
                            cmask[b] = (lo <= hi) ? (((2u << hi) - 1u) & ~((1u << lo) - 1u)) : 0u;
                        }
                    }
                    colbest = a.topD[Lr];
                    colbest_y = 0;
                }
            }
        }
    }
}
