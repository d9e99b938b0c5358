// general.cu -- K1 (match-score matrix) and K3 (intra-task wavefront fill + traceback).
//
// The general path of the drop-in: one alignment with an arbitrary match-score matrix,
// per-position gap arrays and an optional cell mask, i.e. exactly what the reference's
// RawPairwiseAligner hands to cext_align_<mode> (praline/component/align.py:343-388,
// praline/util/cext.c:99-306), and the profile x profile score matrix of cext_build_scores
// (praline/util/cext.c:308-455, :33-97).
//
// K3 shape.  The DP matrix is cut into column strips of 32*KG columns.  A warp lays a strip
// across its lanes (KG columns per lane, previous row's M/U/L in registers) and walks down
// the rows as a systolic pipeline (lane l is one row behind lane l-1, strip edge by warp
// shuffle).  Strip s consumes the right edge of strip s-1 row by row through an edge buffer
// in global memory (L2 resident) guarded by a progress counter, so consecutive strips run as
// an anti-diagonal wavefront across warps and across CTAs.  Strips are dealt round-robin to
// all resident warps in increasing order, which makes the dependency chain deadlock free.
// Arithmetic is the reference's own: three separate sums M+s, U+s, L+s compared for the tie
// flags, -inf borders, masked cells left at 0 -- bit-identical o and t for any f32 input.
// One byte per cell holds all seven tie flags (their bit positions are disjoint,
// cext.c:9-15), instead of the reference's three traceback planes.
#include "common.cuh"

#define FULL 0xffffffffu


__device__ __forceinline__ unsigned long long gkey(float v, uint32_t lin)
{
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | (uint32_t)(~lin);
}

// ---- borders (reference component/align.py:367-385), written where the fill expects them ----
__global__ void k_gen_init(const GenArgs a)
{
    const int L1 = a.L1, L2 = a.L2;
    const float NINF = -INFINITY;
    const bool u_zero = (a.mode == PG_SG_BOTH || a.mode == PG_SG_ONE);
    const bool l_zero = (a.mode == PG_SG_BOTH || a.mode == PG_SG_TWO);
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    for (int y = tid; y <= L1; y += nt) {   // column 0 -> edge[0][y] and lastcol when L2 == 0
        float M = NINF, U, L = NINF;
        if (y == 0) M = 0.f;
        if (u_zero) U = 0.f;
        else if (y == 0) U = a.g1[0] - a.g1[1];
        else U = (float)((double)(y - 1) * (double)a.g1[(size_t)(y - 1) * 2 + 1] + (double)a.g1[0]);
        if (y == 0) L = l_zero ? 0.f : (a.g2[0] - a.g2[1]);
        float* e = a.edge + (size_t)y * 3;
        e[0] = M; e[1] = U; e[2] = L;
        if (a.o_full) {
            float* o = a.o_full + ((size_t)y * (L2 + 1)) * 3;
            o[0] = M; o[1] = U; o[2] = L;
            uint8_t* t = a.t_full + ((size_t)y * (L2 + 1)) * 3;
            t[0] = 0; t[1] = (!u_zero && y >= 1) ? TB_UE : 0; t[2] = 0;
        }
    }
    for (int x = tid; x <= L2; x += nt) {   // row 0 -> top[k][x], lastcol/lastrow seeds
        float M = NINF, U = NINF, L;
        if (x == 0) M = 0.f;
        if (l_zero) L = 0.f;
        else if (x == 0) L = a.g2[0] - a.g2[1];
        else L = (float)((double)(x - 1) * (double)a.g2[(size_t)(x - 1) * 2 + 1] + (double)a.g2[0]);
        if (x == 0) U = u_zero ? 0.f : (a.g1[0] - a.g1[1]);
        a.top[x] = M; a.top[(L2 + 1) + x] = U; a.top[2 * (L2 + 1) + x] = L;
        if (a.o_full && x >= 1) {
            float* o = a.o_full + (size_t)x * 3;
            o[0] = M; o[1] = U; o[2] = L;
            uint8_t* t = a.t_full + (size_t)x * 3;
            t[0] = 0; t[1] = 0; t[2] = (!l_zero) ? TB_LE : 0;
        }
    }
    if (tid == 0) {
        a.progress[0] = L1;
        for (int s = 1; s <= a.n_strips; s++) a.progress[s] = 0;
        *a.best = 0ull;
    }
}

template <int KG, bool LOCAL>
__global__ void __launch_bounds__(256) k_gen_fill(const GenArgs a)
{
    const int L1 = a.L1, L2 = a.L2, W = L2 + 1;
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const float NINF = -INFINITY;

    for (int strip = gw; strip < a.n_strips; strip += nw) {
        const int x0 = strip * (32 * KG) + lane * KG + 1;   // my first column (1-based)
        const float* ein = a.edge + (size_t)strip * (L1 + 1) * 3;
        float* eout = a.edge + (size_t)(strip + 1) * (L1 + 1) * 3;
        volatile int* pin = a.progress + strip;
        volatile int* pout = a.progress + strip + 1;

        float Mp[KG], Up[KG], Lp[KG], go2[KG], ge2[KG];
#pragma unroll
        for (int k = 0; k < KG; k++) {
            const int x = x0 + k;
            const bool v = x <= L2;
            Mp[k] = v ? a.top[x] : NINF;
            Up[k] = v ? a.top[W + x] : NINF;
            Lp[k] = v ? a.top[2 * W + x] : NINF;
            go2[k] = v ? a.g2[(size_t)(x - 1) * 2] : 0.f;
            ge2[k] = v ? a.g2[(size_t)(x - 1) * 2 + 1] : 0.f;
        }
        // diagonal seed for my first column: cell (0, x0-1)
        float Md = NINF, Ud = NINF, Ld = NINF;
        if (x0 <= L2) { Md = a.top[x0 - 1]; Ud = a.top[W + x0 - 1]; Ld = a.top[2 * W + x0 - 1]; }
        float Me = 0.f, Ue = 0.f, Le = 0.f;          // my right edge of the row just finished
        int avail = 0;
        float bv = NINF; uint32_t bl = 0xffffffffu;  // local: best value / smallest linear index
        const int T = L1 + 31;

        // software pipeline: everything a row needs from memory (match scores, the row's gap
        // pair, mask bits, lane 0's left edge) is requested one step ahead, so that no L2 round
        // trip sits on the per-step critical path of the wavefront
        float cm[KG], nm[KG];
        float cg1o = 0.f, cg1e = 0.f, ng1o = 0.f, ng1e = 0.f;
        uint32_t cz = 0, nz = 0;
        float pMe = 0.f, pUe = 0.f, pLe = 0.f;       // lane 0: prefetched left edge
        bool have_edge = false;
        auto fetch_row = [&](int yy, float (&mm)[KG], float& go1, float& ge1, uint32_t& zz) {
            zz = 0;
            if (yy >= 1 && yy <= L1 && x0 <= L2) {
                go1 = a.g1[(size_t)(yy - 1) * 2];
                ge1 = a.g1[(size_t)(yy - 1) * 2 + 1];
                const float* mrow = a.m + (size_t)(yy - 1) * a.m_pitch + (x0 - 1);
#pragma unroll
                for (int k = 0; k < KG; k++) {
                    mm[k] = (x0 + k <= L2) ? __ldg(mrow + k) : 0.f;
                    if (a.z && x0 + k <= L2 && a.z[(size_t)yy * a.z_pitch + x0 + k]) zz |= 1u << k;
                }
            }
        };
        fetch_row(1 - lane, cm, cg1o, cg1e, cz);

        for (int t = 0; t < T; t++) {
            const int y = t - lane + 1;               // my row this step (1-based)
            const bool active = (y >= 1) && (y <= L1) && (x0 <= L2);
            fetch_row(y + 1, nm, ng1o, ng1e, nz);     // in flight while this row is computed
            // lane 0 needs rows up to y of the left strip; wait for the producer
            int need = min(t + 1, L1);
            if (avail < need) {
                if (lane == 0) {
                    // bounded spin: a protocol bug must fail a test, not hang the GPU
                    for (long long spins = 0; (avail = *pin) < need && spins < (1ll << 26); spins++) __nanosleep(32);
                }
                avail = __shfl_sync(FULL, avail, 0);
                __threadfence();
            }
            float Ml = __shfl_up_sync(FULL, Me, 1);
            float Ul = __shfl_up_sync(FULL, Ue, 1);
            float Ll = __shfl_up_sync(FULL, Le, 1);
            if (lane == 0 && y <= L1) {
                if (have_edge) { Ml = pMe; Ul = pUe; Ll = pLe; }
                else { const float* e = ein + (size_t)y * 3; Ml = __ldcg(e); Ul = __ldcg(e + 1); Ll = __ldcg(e + 2); }
                have_edge = (y + 1 <= L1) && (y + 1 <= avail);
                if (have_edge) { const float* e = ein + (size_t)(y + 1) * 3; pMe = __ldcg(e); pUe = __ldcg(e + 1); pLe = __ldcg(e + 2); }
            }
            if (active) {
                const float g1o = cg1o, g1e = cg1e;
                const float dM = Ml, dU = Ul, dL = Ll;   // become next row's diagonal seed
                float cMl = Ml, cLl = Ll;                // left neighbour M / L in this row
                uint8_t fl[KG];
#pragma unroll
                for (int k = 0; k < KG; k++) {
                    const int x = x0 + k;
                    float M = 0.f, U = 0.f, L = 0.f;
                    uint8_t f = 0;
                    const bool valid = x <= L2;
                    const bool masked = (cz >> k) & 1u;
                    if (valid && !masked) {
                        const float s = cm[k];
                        const float up_open = Mp[k] + g1o, up_ext = Up[k] + g1e;
                        const float lf_open = cMl + go2[k], lf_ext = cLl + ge2[k];
                        const float mm = Md + s, mu = Ud + s, ml = Ld + s;
                        float best = LOCAL ? 0.f : NINF;
                        best = fmaxf(best, mm); best = fmaxf(best, mu); best = fmaxf(best, ml);
                        if (mm == best) f |= TB_MM;
                        if (mu == best) f |= TB_MU;
                        if (ml == best) f |= TB_ML;
                        M = best;
                        U = fmaxf(up_open, up_ext);
                        if (up_open == U) f |= TB_UO;
                        if (up_ext == U) f |= TB_UE;
                        L = fmaxf(lf_open, lf_ext);
                        if (lf_open == L) f |= TB_LO;
                        if (lf_ext == L) f |= TB_LE;
                    }
                    Md = Mp[k]; Ud = Up[k]; Ld = Lp[k];
                    if (valid) {
                        Mp[k] = M; Up[k] = U; Lp[k] = L;
                        cMl = M; cLl = L;
                        if (a.o_full) {
                            float* o = a.o_full + ((size_t)y * W + x) * 3;
                            o[0] = M; o[1] = U; o[2] = L;
                            uint8_t* tt = a.t_full + ((size_t)y * W + x) * 3;
                            tt[0] = f & (TB_MM | TB_MU | TB_ML);
                            tt[1] = f & (TB_UO | TB_UE);
                            tt[2] = f & (TB_LO | TB_LE);
                        }
                        if (LOCAL) {
                            const float v3 = fmaxf(fmaxf(M, U), L);
                            if (v3 >= bv) {
                                const uint32_t lin = (uint32_t)(((size_t)y * W + x) * 3);
                                const float vs[3] = {M, U, L};
#pragma unroll
                                for (int j = 0; j < 3; j++)
                                    if (vs[j] > bv || (vs[j] == bv && lin + j < bl)) { bv = vs[j]; bl = lin + j; }
                            }
                        }
                        if (y == L1) {
                            a.lastrow[x] = M; a.lastrow[W + x] = U; a.lastrow[2 * W + x] = L;
                        }
                        if (x == L2) {
                            a.lastcol[y] = M; a.lastcol[(L1 + 1) + y] = U; a.lastcol[2 * (L1 + 1) + y] = L;
                        }
                    }
                    fl[k] = f;
                }
                // next row's diagonal seed is this row's left neighbour
                Md = dM; Ud = dU; Ld = dL;
                {
                    uint8_t* frow = a.flags + (size_t)y * a.f_pitch + x0;
#pragma unroll
                    for (int k = 0; k < KG; k++) if (x0 + k <= L2) frow[k] = fl[k];
                }
                // my right edge = last valid column of my strip part (pad lanes forward nothing useful)
                Me = Mp[KG - 1]; Ue = Up[KG - 1]; Le = Lp[KG - 1];
                if (lane == 31) { float* e = eout + (size_t)y * 3; e[0] = Me; e[1] = Ue; e[2] = Le; }
            }
#pragma unroll
            for (int k = 0; k < KG; k++) cm[k] = nm[k];
            cg1o = ng1o; cg1e = ng1e; cz = nz;
            // publish progress of the right edge every 8 rows and at the end
            const int done = t - 31 + 1;   // rows finished by lane 31 after this step
            if (done >= 1 && ((done & 7) == 0 || done == L1)) {
                __threadfence();
                __syncwarp();
                if (lane == 31) *pout = done;
            }
        }
        if (LOCAL && bl != 0xffffffffu) atomicMax(a.best, gkey(bv, bl));
    }
}

// ---- end cell (reference component/align.py:401-431) -----------------------------------------
__global__ void k_gen_finalize(const GenArgs a)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int L1 = a.L1, L2 = a.L2, W = L2 + 1, H = L1 + 1;
    int cy = L1, cx = L2, ck = 0;
    float score;
    // seed the border entries of lastrow / lastcol
    auto row_at = [&](int x, int k) -> float { return (L1 == 0) ? a.top[k * W + x] : (x == 0 ? a.edge[(size_t)L1 * 3 + k] : a.lastrow[k * W + x]); };
    auto col_at = [&](int y, int k) -> float { return (y == 0) ? a.top[k * W + L2] : (L2 == 0 ? a.edge[(size_t)y * 3 + k] : a.lastcol[k * H + y]); };
    if (a.mode == PG_GLOBAL) {
        ck = 0;
        for (int k = 1; k < 3; k++) if (row_at(L2, k) > row_at(L2, ck)) ck = k;
        score = row_at(L2, ck);
    } else if (a.mode == PG_LOCAL) {
        // interior best from the fill; borders (row 0, column 0) scanned here in linear order
        float bv = -INFINITY; uint32_t bl = 0xffffffffu;
        if (*a.best) {
            uint32_t b = (uint32_t)(*a.best >> 32);
            b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
            bv = __uint_as_float(b); bl = ~(uint32_t)(*a.best);
        }
        for (int x = 0; x <= L2; x++) for (int k = 0; k < 3; k++) {
            const float v = a.top[k * W + x]; const uint32_t lin = (uint32_t)(x * 3 + k);
            if (v > bv || (v == bv && lin < bl)) { bv = v; bl = lin; }
        }
        for (int y = 1; y <= L1; y++) for (int k = 0; k < 3; k++) {
            const float v = a.edge[(size_t)y * 3 + k]; const uint32_t lin = (uint32_t)(((size_t)y * W) * 3 + k);
            if (v > bv || (v == bv && lin < bl)) { bv = v; bl = lin; }
        }
        ck = bl % 3; cx = (bl / 3) % W; cy = bl / 3 / W; score = bv;
    } else {
        float rmax = -INFINITY, cmax = -INFINITY;
        for (int x = 0; x <= L2; x++) for (int k = 0; k < 3; k++) rmax = fmaxf(rmax, row_at(x, k));
        for (int y = 0; y <= L1; y++) for (int k = 0; k < 3; k++) cmax = fmaxf(cmax, col_at(y, k));
        const bool from_row = (a.mode == PG_SG_BOTH || a.mode == PG_SG_TWO);
        bool found = false;
        if (rmax > cmax && from_row) {
            for (int x = L2; x >= 0 && !found; x--) for (int k = 0; k < 3; k++)
                if (row_at(x, k) == rmax) { cy = L1; cx = x; ck = k; found = true; break; }
            score = rmax;
        } else {
            for (int y = L1; y >= 0 && !found; y--) for (int k = 0; k < 3; k++)
                if (col_at(y, k) == cmax) { cy = y; cx = L2; ck = k; found = true; break; }
            score = cmax;
        }
    }
    *a.score_out = score;
    a.cell_out[0] = cy; a.cell_out[1] = cx; a.cell_out[2] = ck;
}

// ---- traceback (reference util/align.py:144-185, :268-297) -----------------------------------
// One warp.  The walk is a chain of dependent byte loads; instead of paying an L2 round trip per
// step, the warp stages a 16 x 64 tile of flag bytes ending at the current cell in shared memory
// (two coalesced 16-byte loads per lane) and lane 0 walks inside it until it leaves the tile.
#define TB_TR 16
#define TB_TC 64
__global__ void __launch_bounds__(32) k_gen_traceback(const GenArgs a)
{
    __shared__ __align__(16) uint8_t tile[TB_TR][TB_TC];
    const int lane = threadIdx.x;
    const int L1 = a.L1, L2 = a.L2;
    const bool u_ramp = !(a.mode == PG_SG_BOTH || a.mode == PG_SG_ONE);
    const bool l_ramp = !(a.mode == PG_SG_BOTH || a.mode == PG_SG_TWO);
    int y = a.cell_out[0], x = a.cell_out[1], k = a.cell_out[2];
    const int cap = L1 + L2 + 2;
    int w = cap;
    auto push = [&](int yy, int xx) { --w; a.path_buf[2 * w] = yy; a.path_buf[2 * w + 1] = xx; };
    const bool semi = (a.mode >= PG_SG_BOTH);
    if (lane == 0 && semi) {
        if (y != L1) { for (int v = L1; v > y; v--) push(v, x); }
        else if (x != L2) { for (int v = L2; v > x; v--) push(y, v); }
    }
    int done = 0;
    while (!done) {
        // stage the tile whose bottom-right region holds (y, x); border cells need no tile
        int ty0 = max(0, y - TB_TR + 1);
        int tx0 = max(0, (x - (TB_TC - 16)) & ~15);
        tx0 = min(tx0, a.f_pitch - TB_TC);
        if (y >= 1 && x >= 1) {
            const int r = lane >> 1, h = lane & 1;
            const int row = min(ty0 + r, L1);
            const uint4* src = reinterpret_cast<const uint4*>(a.flags + (size_t)row * a.f_pitch + tx0 + h * 32);
            uint4* dst = reinterpret_cast<uint4*>(&tile[r][h * 32]);
            dst[0] = src[0];
            dst[1] = src[1];
        }
        __syncwarp();
        if (lane == 0) {
            for (;;) {
                push(y, x);
                uint8_t f;
                if (y == 0 && x == 0) f = 0;
                else if (x == 0) f = (u_ramp && k == 1) ? TB_UE : 0;
                else if (y == 0) f = (l_ramp && k == 2) ? TB_LE : 0;
                else {
                    f = tile[y - ty0][x - tx0];
                    f &= (k == 0) ? (TB_MM | TB_MU | TB_ML) : (k == 1 ? (TB_UO | TB_UE) : (TB_LO | TB_LE));
                }
                if (f & TB_MM) { y--; x--; k = 0; }
                else if (f & TB_MU) { y--; x--; k = 1; }
                else if (f & TB_ML) { y--; x--; k = 2; }
                else if (f & TB_UO) { y--; k = 0; }
                else if (f & TB_UE) { y--; k = 1; }
                else if (f & TB_LO) { x--; k = 0; }
                else if (f & TB_LE) { x--; k = 2; }
                else { done = 1; break; }
                // still inside the staged tile (or on a border, which needs none)?
                if (y >= 1 && x >= 1 && (y < ty0 || x < tx0)) break;
            }
        }
        __syncwarp();
        done = __shfl_sync(FULL, done, 0);
        y = __shfl_sync(FULL, y, 0);
        x = __shfl_sync(FULL, x, 0);
    }
    if (lane == 0) {
        if (semi) {
            if (y != 0) { for (int v = y - 1; v >= 0; v--) push(v, 0); }
            else if (x != 0) { for (int v = x - 1; v >= 0; v--) push(0, v); }
        }
        *a.path_start = w;
        *a.path_len = cap - w;
    }
}

// ---- K1: match-score matrix in the reference's evaluation order ------------------------------
// m[y][x] = sum_sets sum_{i: P1[y][i] != 0} sum_{j: P2[x][j] != 0} (P2[x][j] * S[i][j]) * P1[y][i]
// accumulated sequentially per set, sets added in order (cext.c:63-95, :388-421).  Explicit
// _rn intrinsics keep ptxas from contracting the multiply-add.

__global__ void k_build_scores(const ScoreSets sets, int L1, int L2, float* m, int m_pitch)
{
    const int n_sets = sets.n;
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= L2 || y >= L1) return;
    float score = 0.f;
    for (int n = 0; n < n_sets; n++) {
        const ScoreSet st = sets.s[n];
        const float* r1 = st.P1 + (size_t)y * st.A;
        const float* r2 = st.P2 + (size_t)x * st.A;
        float acc = 0.f;
        for (int i = 0; i < st.A; i++) {
            const float p1 = r1[i];
            if (p1 == 0.f) continue;
            const float* srow = st.S + (size_t)i * st.A;
            for (int j = 0; j < st.A; j++) {
                const float p2 = r2[j];
                if (p2 == 0.f) continue;
                acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(p2, srow[j]), p1));
            }
        }
        score = __fadd_rn(score, acc);
    }
    m[(size_t)y * m_pitch + x] = score;
}

// Batched form for the matrix-fed streaming kernel: one matrix row per stream position of a
// wave (rows in stream order, `width` = 32*K floats each).  rowsrc[r] is the global profile row
// of the STREAMED sequence at that position (-1: dummy row), rowres[r] the resident sequence.
// Same evaluation order as above; `transposed` says the resident is sequence one.
__global__ void __launch_bounds__(128) k_build_rows(const float* __restrict__ prof, const int64_t* __restrict__ rowoff,
                                                    int A, const float* __restrict__ S,
                                                    const int32_t* __restrict__ rowsrc,
                                                    const int32_t* __restrict__ rowres, int64_t n_rows, int width,
                                                    int transposed, float padv, float* __restrict__ mwave)
{
    extern __shared__ float sh[];   // S [A*A] then the streamed profile row [A]
    const int xblocks = (width + 127) / 128;
    const int64_t r = blockIdx.x / xblocks;
    const int x = (int)(blockIdx.x % xblocks) * 128 + threadIdx.x;
    if (r >= n_rows) return;
    const int src = rowsrc[r];
    for (int i = threadIdx.x; i < A * A; i += 128) sh[i] = S[i];
    if (src >= 0) for (int i = threadIdx.x; i < A; i += 128) sh[A * A + i] = prof[(size_t)src * A + i];
    __syncthreads();
    if (x >= width) return;
    float v = padv;
    if (src >= 0) {
        const int res = rowres[r];
        const int64_t r0 = rowoff[res];
        const int Lr = (int)(rowoff[res + 1] - r0);
        if (x < Lr) {
            const float* rr = prof + (size_t)(r0 + x) * A;   // resident profile row
            const float* sr = sh + A * A;                    // streamed profile row
            float acc = 0.f;
            for (int i = 0; i < A; i++) {
                const float p1 = transposed ? rr[i] : sr[i];
                if (p1 == 0.f) continue;
                for (int j = 0; j < A; j++) {
                    const float p2 = transposed ? sr[j] : rr[j];
                    if (p2 == 0.f) continue;
                    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(p2, sh[i * A + j]), p1));
                }
            }
            v = __fadd_rn(0.f, acc);
        }
    } else {
        v = 0.f;
    }
    mwave[(size_t)r * width + x] = v;
}

// Sequence x sequence: one-hot profiles make the sum collapse to exactly S[a_y][b_x].
__global__ void k_build_scores_seq(const uint8_t* a, const uint8_t* b, const float* S, int A, int L1, int L2,
                                   float* m, int m_pitch)
{
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    const int y = blockIdx.y * blockDim.y + threadIdx.y;
    if (x >= L2 || y >= L1) return;
    m[(size_t)y * m_pitch + x] = S[(size_t)a[y] * A + b[x]];
}

// ---- host side ----------------------------------------------------------------------------------
int pg_launch_general(GenArgs a, int kg, cudaStream_t st)
{
    a.n_strips = (a.L2 + 32 * kg - 1) / (32 * kg);
    if (a.n_strips < 1) a.n_strips = 1;
    k_gen_init<<<32, 256, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    if (a.L1 > 0 && a.L2 > 0) {
        int dev = 0, sms = 148;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const int wpc = 8;
        int ctas = (a.n_strips + wpc - 1) / wpc;
        if (ctas > sms) ctas = sms;    // all warps co-resident: strips wait on lower strips only
        const bool local = a.mode == PG_LOCAL;
        if (kg == 2) {
            if (local) k_gen_fill<2, true><<<ctas, wpc * 32, 0, st>>>(a);
            else k_gen_fill<2, false><<<ctas, wpc * 32, 0, st>>>(a);
        } else if (kg == 8) {
            if (local) k_gen_fill<8, true><<<ctas, wpc * 32, 0, st>>>(a);
            else k_gen_fill<8, false><<<ctas, wpc * 32, 0, st>>>(a);
        } else {
            pg_set_error("general kernel: unsupported strip width kg=%d", kg);
            return 1;
        }
        PG_CUDA_OK(cudaGetLastError());
    }
    k_gen_finalize<<<1, 32, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    if (a.path_buf) {
        k_gen_traceback<<<1, 32, 0, st>>>(a);
        PG_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

int pg_launch_build_scores(const ScoreSets& sets, int L1, int L2, float* m, int m_pitch, cudaStream_t st)
{
    if (L1 <= 0 || L2 <= 0) return 0;
    dim3 b(32, 8), g((L2 + 31) / 32, (L1 + 7) / 8);
    k_build_scores<<<g, b, 0, st>>>(sets, L1, L2, m, m_pitch);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_build_rows(const float* prof, const int64_t* rowoff, int A, const float* S, const int32_t* rowsrc,
                         const int32_t* rowres, int64_t n_rows, int width, int transposed, float padv, float* mwave,
                         cudaStream_t st)
{
    if (n_rows <= 0) return 0;
    const int64_t blocks = n_rows * ((width + 127) / 128);
    if (blocks > 0x7fffffffll) { pg_set_error("wave too large for one launch (%lld blocks)", (long long)blocks); return 1; }
    k_build_rows<<<(unsigned)blocks, 128, sizeof(float) * (A * A + A), st>>>(prof, rowoff, A, S, rowsrc, rowres, n_rows,
                                                                           width, transposed, padv, mwave);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_build_scores_seq(const uint8_t* a, const uint8_t* b, const float* S, int A, int L1, int L2,
                               float* m, int m_pitch, cudaStream_t st)
{
    if (L1 <= 0 || L2 <= 0) return 0;
    dim3 blk(32, 8), g((L2 + 31) / 32, (L1 + 7) / 8);
    k_build_scores_seq<<<g, blk, 0, st>>>(a, b, S, A, L1, L2, m, m_pitch);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
