// gotoh_stream.cu -- K2, the inter-task three-state (Gotoh-style) DP kernel for sm_100a.
//
// Replaces, for batches of sequence-sequence pairs with constant gap penalties, the
// reference's per-pair chain  cext_build_scores -> border init -> cext_align_<mode> ->
// end-cell choice  (praline/util/cext.c:308-455, :99-306; praline/component/align.py:357-431).
//
// Shape.  One CTA owns one RESIDENT sequence: its substitution profile
//     prof[a][x] = S[a][resident[x]]   (or S[resident[x]][a] when the resident is sequence one)
// is built once in shared memory, lane-interleaved so that a lane's K consecutive columns
// arrive as conflict-free 128-bit shared loads.  Every warp lays the resident across its 32
// lanes (K columns per lane, DP state of the previous row in registers) and pushes a STREAM
// of row sequences through it as a systolic pipeline: at step t lane l works on stream
// position t - l and hands the right edge of its strip (M+open, L, max3) to lane l + 1 by
// warp shuffle.  Row sequences are concatenated back to back, so the pipeline fills once per
// warp, not once per pair; a flagged last row emits the pair's result and re-arms the top
// border.  The match-score matrix, the three DP planes and the traceback planes of the
// reference (20 B/cell) never exist in HBM: score-only runs write 4 B per PAIR, traced runs
// add one packed nibble per cell (coalesced 128 B stores, [step/8][k][lane] words).
//
// Arithmetic.  f32 like the reference.  Per cell (reference cext.c:155-200, 214-221, 249-254,
// 278-283):  U = max(M_up + open, U_up + ext), L = max(M_left + open, L_left + ext),
// M = max(M_d, U_d, L_d) + s.  The kernel carries M + open and max3(M, U, L) instead of M, U,
// L of the diagonal: max3(..) + s equals the reference's max of three sums bit for bit
// (rounding is monotone) and is 7 instead of 11 flops.  The tie FLAGS of the reference
// compare the three rounded sums; deriving them from the unrounded operands is exact when
// every score is integer valued (sequence-sequence alignment with integer matrices and
// gaps), which the host checks before it routes a traced batch here.
//
// The numpy lane-model tests/model_stream.py is the executable specification of this file.
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
template <int K, int KM, bool TB, bool TR, bool MS, int NW>
__global__ void __launch_bounds__(NW * 32) k_stream(const StreamArgs a)
{
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

    float Mo[K], U[K], D[K];
    uint32_t acc[K];
#pragma unroll
    for (int k = 0; k < K; k++) { Mo[k] = 0.f; U[k] = 0.f; D[k] = 0.f; acc[k] = 0u; }
    float Mo_last = 0.f, L_last = 0.f, D_last = 0.f, Dleft_prev = 0.f;
    float best = 0.f, colbest = 0.f, yf = 0.f;   // yf = rows of the current sequence already done
    int colbest_y = 0, q = sb;
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
#pragma unroll
            for (int i = 0; i < UNR; i++) {
                const int t = t0 + g + i;
                uint32_t w;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w) : "r"(rp + (uint32_t)i * 4u) : "memory");
                float sc[NCH * 4];
                if (MS) {
                    const float* mr = mwarp + (size_t)(w & 0x00ffffffu) * (32 * K) + lane * K;
                    if (K % 4 == 0) {
#pragma unroll
                        for (int j = 0; j < K / 4; j++) {
                            const float4 v = __ldg(reinterpret_cast<const float4*>(mr) + j);
                            sc[4 * j] = v.x; sc[4 * j + 1] = v.y; sc[4 * j + 2] = v.z; sc[4 * j + 3] = v.w;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < K; k++) sc[k] = __ldg(mr + k);
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

#pragma unroll
                for (int k = 0; k < K; k++) {
                    float m = diag + sc[k];
                    if (KM == 1) { m = fmaxf(m, 0.f); best = fmaxf(best, m); }
                    const float ue = U[k] + ge;
                    const float le = Ll + ge;
                    const float u = fmaxf(Mo[k], ue);
                    const float l = fmaxf(Ml, le);
                    diag = D[k];
                    float d;
                    if (TB) {
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
                            atomicMax(a.rowkey + slot, pack_key(best, 0));
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
                    colbest = a.topD[Lr];
                    colbest_y = 0;
                }
            }
        }
    }
}

// ---- semiglobal / local score from the keys (reference component/align.py:401-424) ------------
__device__ __forceinline__ float key_value(unsigned long long k)
{
    uint32_t b = (uint32_t)(k >> 32);
    b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
    return __uint_as_float(b);
}

__global__ void k_semi_scores(int64_t n, const unsigned long long* rowkey, const unsigned long long* colkey,
                              int mode, int transposed, float* scores)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (colkey[i] == 0ull) return;   // slot not produced by this launch (other K class / shard)
    if (mode == PG_LOCAL) { scores[i] = key_value(rowkey[i]); return; }
    const float kr = key_value(rowkey[i]), kc = key_value(colkey[i]);
    const float ref_row = transposed ? kc : kr, ref_col = transposed ? kr : kc;
    const bool from_row = (mode == PG_SG_BOTH || mode == PG_SG_TWO);
    scores[i] = (ref_row > ref_col && from_row) ? ref_row : ref_col;
}

// ---- launch ------------------------------------------------------------------------------------
constexpr int kNW = 8;

template <int K, int KM, bool TB, bool TR, bool MS = false>
static int launch_one(const StreamArgs& a, int n_tiles, cudaStream_t st)
{
    constexpr int NCH = (K + 3) / 4;
    const size_t smem = (MS ? 0 : (size_t)a.A * NCH * 512) + kNW * 128 * sizeof(uint32_t);
    auto kern = k_stream<K, KM, TB, TR, MS, kNW>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, kNW * 32, smem, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int K>
static int launch_k(const StreamArgs& a, int n_tiles, int km, bool tb, cudaStream_t st)
{
    if (a.mwave) {   // profile batches: score only, scores read from the materialised matrix
        if (tb) { pg_set_error("matrix-fed batches are score-only; traced profile alignments use the general kernel"); return 1; }
        if (km == 0) return launch_one<K, 0, false, false, true>(a, n_tiles, st);
        if (km == 1) return launch_one<K, 1, false, false, true>(a, n_tiles, st);
        return launch_one<K, 2, false, false, true>(a, n_tiles, st);
    }
    if (!tb) {
        if (km == 0) return launch_one<K, 0, false, false>(a, n_tiles, st);
        if (km == 1) return launch_one<K, 1, false, false>(a, n_tiles, st);
        return launch_one<K, 2, false, false>(a, n_tiles, st);
    }
    if (km == 1) { pg_set_error("local mode has no batched traceback; use the general kernel"); return 1; }
    if (a.transposed) {
        if (km == 0) return launch_one<K, 0, true, true>(a, n_tiles, st);
        return launch_one<K, 2, true, true>(a, n_tiles, st);
    }
    if (km == 0) return launch_one<K, 0, true, false>(a, n_tiles, st);
    return launch_one<K, 2, true, false>(a, n_tiles, st);
}

int pg_stream_supported_k(int k)
{
    switch (k) {
        case 1: case 2: case 3: case 4: case 6: case 8: case 10: case 12: case 13: case 14:
        case 16: case 20: case 24: case 32: return 1;
        default: return 0;
    }
}

int pg_launch_stream(const StreamArgs& a, int n_tiles, int K, int mode, bool tb, cudaStream_t st)
{
    const int km = (mode == PG_GLOBAL) ? 0 : (mode == PG_LOCAL ? 1 : 2);
    if (n_tiles <= 0) return 0;
    switch (K) {
        case 1: return launch_k<1>(a, n_tiles, km, tb, st);
        case 2: return launch_k<2>(a, n_tiles, km, tb, st);
        case 3: return launch_k<3>(a, n_tiles, km, tb, st);
        case 4: return launch_k<4>(a, n_tiles, km, tb, st);
        case 6: return launch_k<6>(a, n_tiles, km, tb, st);
        case 8: return launch_k<8>(a, n_tiles, km, tb, st);
        case 10: return launch_k<10>(a, n_tiles, km, tb, st);
        case 12: return launch_k<12>(a, n_tiles, km, tb, st);
        case 13: return launch_k<13>(a, n_tiles, km, tb, st);
        case 14: return launch_k<14>(a, n_tiles, km, tb, st);
        case 16: return launch_k<16>(a, n_tiles, km, tb, st);
        case 20: return launch_k<20>(a, n_tiles, km, tb, st);
        case 24: return launch_k<24>(a, n_tiles, km, tb, st);
        case 32: return launch_k<32>(a, n_tiles, km, tb, st);
        default: pg_set_error("unsupported columns-per-lane K=%d", K); return 1;
    }
}

int pg_launch_semi_scores(int64_t n, const unsigned long long* rowkey, const unsigned long long* colkey,
                          int mode, int transposed, float* scores, cudaStream_t st)
{
    if (n <= 0) return 0;
    k_semi_scores<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, rowkey, colkey, mode, transposed, scores);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
