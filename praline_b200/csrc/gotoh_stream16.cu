// gotoh_stream16.cu -- K2 in packed 16-bit integers (DPX): two streamed sequences per warp.
//
// Same systolic streaming design as gotoh_stream.cu (read that header first), for the case the
// headline workload is: global mode, score per pair, INTEGER substitution matrix and gaps.
// Every 32-bit register holds two signed 16-bit DP values -- the low half belongs to a row of
// stream A, the high half to a row of stream B, two different streamed sequences that run
// through the same resident in lock step.  Blackwell's DPX instructions then do the recurrence
// of two cells at once:
//     U = VIADDMNMX.S16x2(U_up, ext, M_up+open)      L = VIADDMNMX.S16x2(L_left, ext, M_left+open)
//     D = VIMNMX3.S16x2(M, U, L)                      M, M+open = VIADD.16x2
// 5 instructions per TWO cells against 7 per cell in f32, and the per-step overhead (ring,
// shuffles, loop) is shared by the two streams as well.  The substitution profile is stored as
// int16 in shared memory; a PRMT per column interleaves stream A's and stream B's score rows.
//
// Exactness.  With integer scores every f32 value of the reference is an integer, so integer
// arithmetic reproduces it bit for bit as long as nothing leaves the int16 range.  -inf is the
// sentinel NEG, chosen by the host below every reachable value minus the largest score, and
// the host only routes a launch here when |go| + (|ge| + max|S|) * (L1 + L2) and the sentinel
// arithmetic fit (Engine.fits_s16).  Otherwise the f32 kernel runs.
//
// Traced variant (TB).  The four tie tests of a cell compare a value with a maximum it took part
// in (a <= b always), so "a == b" is the sign of  b + ~a = b - a - 1  per 16-bit half: packed adds
// of complements, no packed subtract or compare exists.  Two PRMTs with sign replication turn the
// eight sign bits of a column step (4 tests x 2 halves) into byte masks, two LOP3s keep one bit
// per test, and one IMAD shifts them into the column's word: four rows of both halves per 32-bit
// word, stored [step/4][k][lane] (coalesced 128-byte lines, 0.5 B per cell like the f32 kernel).
// 18 instructions per TWO cells against 16 per cell in f32.  The host routes a traced global
// batch here only when all values stay within +-16000 (the tests need b - a - 1 inside int16).
#include "common.cuh"

#define FLAG_LAST 0x80000000u
#define FLAG_EMIT 0x40000000u
#define FULL 0xffffffffu

__device__ __forceinline__ uint32_t pack2(int v) { return ((uint32_t)(v & 0xffff)) | ((uint32_t)v << 16); }
// generic PRMT: a selector nibble with bit 3 set replicates the sign of the selected byte
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int K>
__device__ __forceinline__ uint32_t pick_u(const uint32_t (&v)[K], int k)
{
    uint32_t r = v[0];
#pragma unroll
    for (int i = 1; i < K; i++) r = (k == i) ? v[i] : r;
    return r;
}

// TB: write packed traceback words.  TR: tie order of a transposed (resident = sequence one) launch.
template <int K, int NW, bool TB, bool TR>
__global__ void __launch_bounds__(NW * 32) k_stream16(const StreamArgs a)
{
    constexpr int UNR = TB ? 4 : 8;         // traced: the four rows of one traceback word
    constexpr int NCH = (K + 7) / 8;        // 16-byte chunks of 8 int16 columns per lane
    constexpr int ROWB = NCH * 512;
    constexpr int KP = (K + 3) & ~3;        // top-border table, words per lane
    extern __shared__ __align__(16) unsigned char smem[];
    int16_t* prof = reinterpret_cast<int16_t*>(smem);
    uint32_t* top2 = reinterpret_cast<uint32_t*>(smem + (size_t)a.A * ROWB);            // [32][KP + 4]
    uint32_t* rings = top2 + 32 * (KP + 4) + (threadIdx.x >> 5) * 256;                   // A: [0,128)  B: [128,256)

    const PgTile tile = a.tiles[blockIdx.x];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t roff = a.offs[tile.resident];
    const int Lr = (int)(a.offs[tile.resident + 1] - roff);

    // ---- int16 substitution profile of the resident, [a][chunk][lane][8]; top border table -----
    {
        const int n = a.A * NCH * 256;
        for (int idx = threadIdx.x; idx < n; idx += NW * 32) {
            const int c = idx & 7, l = (idx >> 3) & 31, j = (idx >> 8) % NCH, sym = idx / (NCH * 256);
            const int k = 8 * j + c, x = l * K + k;
            int v = 0;
            if (k < K && x < Lr) {
                const int b = a.seqs[roff + x];
                v = (int)(a.transposed ? a.S[b * a.A + sym] : a.S[sym * a.A + b]);
            }
            prof[idx] = (int16_t)v;
        }
        for (int idx = threadIdx.x; idx < 32 * (KP + 4); idx += NW * 32) {
            const int l = idx / (KP + 4), k = idx % (KP + 4);
            int v = 0;
            if (k < K) v = (int)a.topD[l * K + k + 1];          // D(0, x) of my columns
            else if (k == KP) v = (int)a.topD[l * K];           // D(0, x0 - 1): first diagonal
            top2[idx] = pack2(v);
        }
    }
    const uint32_t prof_s = (uint32_t)__cvta_generic_to_shared(smem);
    for (int i = lane; i < 256; i += 32) rings[i] = prof_s;
    __syncthreads();
    if (prof_s & 511u) __trap();

    // ---- this warp's slice of the tile's stream, cut into stream A and stream B ----------------
    const int n_str = tile.stream_end - tile.stream_begin;
    const int per = (n_str + NW - 1) / NW;
    const int sb = tile.stream_begin + warp * per;
    const int se = min(sb + per, tile.stream_end);
    if (sb >= se) return;
    const int mid = sb + (se - sb + 1) / 2;     // A = [sb, mid), B = [mid, se) (B may be empty)

    auto seq_id = [&](int s) -> int { return a.stream_ids ? a.stream_ids[s] : s; };
    auto raw_len = [&](int s) -> int { const int id = seq_id(s); return (int)(a.offs[id + 1] - a.offs[id]); };

    int totA = 0, totB = 0;
    for (int s = sb + lane; s < mid; s += 32) totA += raw_len(s);
    for (int s = mid + lane; s < se; s += 32) totB += raw_len(s);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        totA += __shfl_xor_sync(FULL, totA, o);
        totB += __shfl_xor_sync(FULL, totB, o);
    }
    const int T = (max(totA, totB) + 1 + 31 + 31) & ~31;

    const int lr = (Lr - 1) / K, klast = (Lr - 1) % K;
    const uint32_t go2 = pack2(a.go16), ge2 = pack2(a.ge16), NEG2 = pack2(a.neg16);
    const uint32_t left0_2 = pack2(a.left0_16), left1_2 = pack2(a.left1_16);
    const uint32_t ringA_s = (uint32_t)__cvta_generic_to_shared(rings);
    const uint32_t lane16 = (uint32_t)lane << 4;
    const uint32_t* mytop = top2 + lane * (KP + 4);

    uint32_t Mo[K], U[K], D[K];
    uint32_t nMo[TB ? K : 1], acc[TB ? K : 1];     // traced: ~Mo of the previous row, flag word per column
#pragma unroll
    for (int k = 0; k < K; k++) { Mo[k] = 0u; U[k] = 0u; D[k] = 0u; }
    if (TB) {
#pragma unroll
        for (int k = 0; k < K; k++) { nMo[k] = ~0u; acc[k] = 0u; }
    }
    const int64_t tbw0 = TB ? a.tb_base[(int64_t)blockIdx.x * NW + warp] : 0;
    // the ALU pipe (DPX, PRMT, LOP3) is the traced kernel's binding unit: complements are taken as
    // x * (-1) + (-1) with the -1 from a kernel argument, which keeps them on the FMA pipe as IMADs
    const uint32_t ones = (uint32_t)a.all_ones;
    auto inv = [&](uint32_t x) -> uint32_t { return x * ones + ones; };
    uint32_t Mo_last = 0u, L_last = 0u, D_last = 0u, Dleft_prev = 0u, bord = 0u;
    int qA = sb, qB = mid;
    int psA = sb - 1, ppA = 0, psB = mid - 1, ppB = 0;   // producer cursors (element, offset)

    // decode 32 stream positions of one stream into its ring (kept twice, 64 apart)
    auto refill = [&](int t0, int s_begin, int s_end, int& ps, int& pp, uint32_t* ring) {
        auto slen = [&](int s) -> int { return (s < s_begin) ? 1 : raw_len(s); };   // dummy row arms the first reset
        int s = ps, p = pp + lane;
        int len = (s < s_end) ? slen(s) : 0;
        while (s < s_end && p >= len) {
            p -= len;
            s++;
            len = (s < s_end) ? slen(s) : 0;
        }
        uint32_t word = prof_s;
        if (s < s_end) {
            if (s < s_begin) {
                word |= FLAG_LAST;
            } else {
                const int sym = a.seqs[a.offs[seq_id(s)] + p];
                word = (prof_s + (uint32_t)(sym * ROWB)) | ((p == len - 1) ? (FLAG_LAST | FLAG_EMIT) : 0u);
            }
        }
        ring[(t0 + lane) & 63] = word;
        ring[((t0 + lane) & 63) + 64] = word;
        int s31 = __shfl_sync(FULL, s, 31), p31 = __shfl_sync(FULL, p, 31) + 1;
        const int len31 = __shfl_sync(FULL, len, 31);
        if (s31 < s_end && p31 >= len31) { p31 = 0; s31++; }
        ps = s31;
        pp = p31;
    };

    for (int t0 = 0; t0 < T; t0 += 32) {
        __syncwarp();
        refill(t0, sb, mid, psA, ppA, rings);
        refill(t0, mid, se, psB, ppB, rings + 128);
        __syncwarp();
        const uint32_t rp0 = ringA_s + ((uint32_t)((t0 - lane) & 63) << 2);

#pragma unroll 1
        for (int g = 0; g < 32; g += UNR) {
            const uint32_t rp = rp0 + (uint32_t)g * 4u;
#pragma unroll
            for (int i = 0; i < UNR; i++) {
                uint32_t wA, wB;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wA) : "r"(rp + (uint32_t)i * 4u) : "memory");
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(wB) : "r"(rp + 512u + (uint32_t)i * 4u) : "memory");
                const uint32_t paA = (wA & 0x00ffffffu) | lane16, paB = (wB & 0x00ffffffu) | lane16;
                uint32_t ra[NCH * 4], rb[NCH * 4];
#pragma unroll
                for (int j = 0; j < NCH; j++) {
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                        : "=r"(ra[4 * j]), "=r"(ra[4 * j + 1]), "=r"(ra[4 * j + 2]), "=r"(ra[4 * j + 3])
                        : "r"(paA + (uint32_t)j * 512u));
                    asm("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                        : "=r"(rb[4 * j]), "=r"(rb[4 * j + 1]), "=r"(rb[4 * j + 2]), "=r"(rb[4 * j + 3])
                        : "r"(paB + (uint32_t)j * 512u));
                }

                uint32_t Ml = __shfl_up_sync(FULL, Mo_last, 1);
                uint32_t Ll = __shfl_up_sync(FULL, L_last, 1);
                uint32_t Dn = __shfl_up_sync(FULL, D_last, 1);
                if (lane == 0) {   // column 0: M = L = -inf, max3 = the U border of each stream's row
                    Ml = NEG2;
                    Ll = NEG2;
                    Dn = bord;
                }
                bord = __vadd2(bord, left1_2);
                uint32_t diag = Dleft_prev;
                Dleft_prev = Dn;
                uint32_t nMl = TB ? inv(Ml) : 0u;

#pragma unroll
                for (int k = 0; k < K; k++) {
                    // stream A's score in the low half, stream B's in the high half
                    const uint32_t s2 = (k & 1) ? __byte_perm(ra[k >> 1], rb[k >> 1], 0x7632)
                                                : __byte_perm(ra[k >> 1], rb[k >> 1], 0x5410);
                    const uint32_t m = __vadd2(diag, s2);
                    const uint32_t u = __viaddmax_s16x2(U[k], ge2, Mo[k]);
                    const uint32_t l = __viaddmax_s16x2(Ll, ge2, Ml);
                    diag = D[k];
                    const uint32_t d = __vimax3_s16x2(m, u, l);
                    const uint32_t mo = __vadd2(m, go2);
                    if (TB) {
                        // sign per half = "equal": M is the maximum | the first-priority gap state is |
                        // U was opened | L was opened (ties keep the reference's priority)
                        const uint32_t nmo = inv(mo);
                        const uint32_t t1 = __vadd2(d, inv(m));
                        const uint32_t t2 = __vadd2(d, inv(TR ? l : u));
                        const uint32_t t3 = __vadd2(u, nMo[k]);
                        const uint32_t t4 = __vadd2(l, nMl);
                        const uint32_t p12 = prmt(t1, t2, 0xBF9Du);     // bytes: t1.B t2.B t1.A t2.A as 0x00 / 0xff
                        const uint32_t p34 = prmt(t3, t4, 0xBF9Du);
                        const uint32_t q = (p12 & 0x02020202u) | (p34 & 0x01010101u);
                        acc[k] = (i == 0) ? q : acc[k] * 4u + q;
                        nMo[k] = nmo;
                        nMl = nmo;
                    }
                    Mo[k] = mo;
                    U[k] = u;
                    D[k] = d;
                    Ml = mo;
                    Ll = l;
                }
                Mo_last = Ml;
                L_last = Ll;
                D_last = D[K - 1];
                if (TB && i == UNR - 1) {
                    uint32_t* dst = a.tb + tbw0 + (int64_t)((t0 + g) >> 2) * (K * 32) + lane;
#pragma unroll
                    for (int k = 0; k < K; k++) dst[k * 32] = acc[k];
                }

                if ((int)(wA | wB) < 0) {   // a LAST row in one of the two streams
                    const bool lastA = (int)wA < 0, lastB = (int)wB < 0;
                    if (lane == lr) {
                        const uint32_t dv = pick_u<K>(D, klast);
                        if (lastA && (wA & FLAG_EMIT)) {
                            const int64_t slot = tile.out_base + (qA - tile.stream_begin);
                            a.scores[slot] = (float)(int)(int16_t)(dv & 0xffffu);
                            if (TB) { a.emit_t[slot] = t0 + g + i; a.pair_tb[slot] = tbw0; }
                        }
                        if (lastB && (wB & FLAG_EMIT)) {
                            const int64_t slot = tile.out_base + (qB - tile.stream_begin);
                            a.scores[slot] = (float)(int)(int16_t)(dv >> 16);
                            if (TB) { a.emit_t[slot] = (t0 + g + i) | (1 << 30); a.pair_tb[slot] = tbw0; }   // bit 30: high half
                        }
                    }
                    if (lastA && (wA & FLAG_EMIT)) qA++;
                    if (lastB && (wB & FLAG_EMIT)) qB++;
                    // re-arm the top border in the half (or halves) that finished a sequence:
                    // PRMT keeps the other stream's half of every state register
                    const uint32_t sel = lastA ? (lastB ? 0x7654u : 0x3254u) : 0x7610u;
#pragma unroll
                    for (int k = 0; k < K; k++) {
                        Mo[k] = __byte_perm(Mo[k], NEG2, sel);
                        U[k] = __byte_perm(U[k], NEG2, sel);
                        D[k] = __byte_perm(D[k], mytop[k], sel);
                    }
                    Dleft_prev = __byte_perm(Dleft_prev, mytop[KP], sel);
                    bord = __byte_perm(bord, left0_2, sel);
                    if (TB) {
#pragma unroll
                        for (int k = 0; k < K; k++) nMo[k] = ~Mo[k];
                    }
                }
            }
        }
    }
}

// ---- paired RESIDENTS (the all-vs-all shape): gotoh_stream16r.cuh ------------------------------------
#include "gotoh_stream16r.cuh"


// ---- launch ------------------------------------------------------------------------------------
constexpr int kNW16 = 8;

template <int K>
static int launch16(const StreamArgs& a, int n_tiles, int paired, cudaStream_t st)
{
    constexpr int KP = (K + 3) & ~3;
    if (paired) {
        constexpr int NCH = (K + 3) / 4;
        const size_t smem = (size_t)a.A * NCH * 512 + 32 * (KP + 4) * 4 + kNW16 * (4 * (64 * K16R_SKEW + 8) + 64 * K16R_SKEW) * sizeof(uint32_t);
        auto kern = k_stream16r<K, kNW16, false>;
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        StreamArgs b = a;
        b.all_ones = -1;
        kern<<<n_tiles, kNW16 * 32, smem, st>>>(b);
    } else {
        constexpr int NCH = (K + 7) / 8;
        const size_t smem = (size_t)a.A * NCH * 512 + 32 * (KP + 4) * 4 + kNW16 * 256 * sizeof(uint32_t);
#define PG_S16(TBV, TRV)                                                                              \
    do {                                                                                              \
        auto kern = k_stream16<K, kNW16, TBV, TRV>;                                                   \
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<n_tiles, kNW16 * 32, smem, st>>>(a);                                                   \
    } while (0)
        if (!a.tb) PG_S16(false, false);
        else if (a.transposed) PG_S16(true, true);
        else PG_S16(true, false);
#undef PG_S16
    }
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_stream16(const StreamArgs& a, int n_tiles, int K, int paired, cudaStream_t st)
{
    if (n_tiles <= 0) return 0;
    switch (K) {
        case 1: return launch16<1>(a, n_tiles, paired, st);
        case 2: return launch16<2>(a, n_tiles, paired, st);
        case 3: return launch16<3>(a, n_tiles, paired, st);
        case 4: return launch16<4>(a, n_tiles, paired, st);
        case 6: return launch16<6>(a, n_tiles, paired, st);
        case 8: return launch16<8>(a, n_tiles, paired, st);
        case 10: return launch16<10>(a, n_tiles, paired, st);
        case 12: return launch16<12>(a, n_tiles, paired, st);
        case 13: return launch16<13>(a, n_tiles, paired, st);
        case 14: return launch16<14>(a, n_tiles, paired, st);
        case 16: return launch16<16>(a, n_tiles, paired, st);
        case 20: return launch16<20>(a, n_tiles, paired, st);
        case 24: return launch16<24>(a, n_tiles, paired, st);
        case 32: return launch16<32>(a, n_tiles, paired, st);
        default: pg_set_error("unsupported columns-per-lane K=%d", K); return 1;
    }
}
