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
#include "gotoh_stream.cuh"

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

template <int K, int KM, bool TB, bool TR, bool MS = false, bool MK = false>
static int launch_one(const StreamArgs& a, int n_tiles, cudaStream_t st)
{
    constexpr int NCH = (K + 3) / 4;
    const size_t smem = (MS ? 0 : (size_t)a.A * NCH * 512) + kNW * 128 * sizeof(uint32_t);
    auto kern = k_stream<K, KM, TB, TR, MS, MK, kNW>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, kNW * 32, smem, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int K>
static int launch_k(const StreamArgs& a, int n_tiles, int km, bool tb, cudaStream_t st)
{
    if (!tb) {
        if (km == 0) return launch_one<K, 0, false, false>(a, n_tiles, st);
        if (km == 1) return launch_one<K, 1, false, false>(a, n_tiles, st);
        return launch_one<K, 2, false, false>(a, n_tiles, st);
    }
    if (km == 1) { pg_set_error("local traced batches are launched through pg_launch_stream_local"); return 1; }
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
    if (a.mwave) {   // profile batches: score only, scores read from the materialised matrix (gotoh_stream_ms.cu)
        if (tb) { pg_set_error("matrix-fed batches are score-only; traced profile alignments use the general kernel"); return 1; }
        return pg_launch_stream_ms(a, n_tiles, K, km, st);
    }
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
