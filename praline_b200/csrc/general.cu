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
#include <stdlib.h>

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
        if (a.flag_fmt == 2) {          // lean kernel: tagged 32-byte records (see k_wave)
            float* e = a.edge + (size_t)y * 8;
            const float tg = __uint_as_float((uint32_t)y);
            e[0] = M; e[1] = tg; e[2] = U; e[3] = tg; e[4] = L; e[5] = tg; e[6] = 0.f; e[7] = 0.f;
        } else {
            float* e = a.edge + (size_t)y * 4;
            e[0] = M; e[1] = U; e[2] = L; e[3] = 0.f;
        }
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
        *a.err = 0;
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
        const float* ein = a.edge + (size_t)strip * (L1 + 1) * 4;
        float* eout = a.edge + (size_t)(strip + 1) * (L1 + 1) * 4;
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
                    if (avail < need) atomicOr(a.err, 1);     // timed out: flag it (k_gen_finalize -> NaN score)
                }
                avail = __shfl_sync(FULL, avail, 0);
                __threadfence();
            }
            float Ml = __shfl_up_sync(FULL, Me, 1);
            float Ul = __shfl_up_sync(FULL, Ue, 1);
            float Ll = __shfl_up_sync(FULL, Le, 1);
            if (lane == 0 && y <= L1) {
                if (have_edge) { Ml = pMe; Ul = pUe; Ll = pLe; }
                else { const float* e = ein + (size_t)y * 4; Ml = __ldcg(e); Ul = __ldcg(e + 1); Ll = __ldcg(e + 2); }
                have_edge = (y + 1 <= L1) && (y + 1 <= avail);
                if (have_edge) { const float* e = ein + (size_t)(y + 1) * 4; pMe = __ldcg(e); pUe = __ldcg(e + 1); pLe = __ldcg(e + 2); }
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
                if (lane == 31) { float* e = eout + (size_t)y * 4; e[0] = Me; e[1] = Ue; e[2] = Le; }
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


// ---- the production fill: same wavefront, compact flags ---------------------------------------
// The walker only ever follows the FIRST set flag in the reference's priority order
// (util/align.py:161-174), so one byte per cell with five sign bits is enough:
//   bit 0  M+s lost strictly against the maximum        (not MM)
//   bit 1  U+s lost strictly                            (not MU)
//   bit 2  the up state was extended, open lost strictly (open beats extend on ties)
//   bit 3  the left state was extended
//   bit 4  local mode only: L+s lost strictly too       (with bits 0,1: no flag -> path stops)
//   bit 7  masked cell (reference leaves all three flag bytes 0, cext.c:141-149)
// "x == max" for x <= max is "sign(x - max) == 0": distinct finite floats never subtract to
// zero, -inf minus -inf gives the positive default NaN (equal, as in C), so the bits are exact
// for ANY f32 input, like the three-sum comparison of the reference they encode.
#define GEN_R 8            // depth of the per-lane prefetch ring (rows in flight per lane)

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async4_cg(uint32_t dst, const void* src)   // through L2: data of other SMs
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src));
}

template <int KG, bool LOCAL, bool MASK>
__global__ void __launch_bounds__(256) k_gen_fill_fast(const GenArgs a)
{
    // per warp: GEN_R slots of [32 lanes][KG match scores + 2 gap values], filled by cp.async
    // GEN_R - 1 rows ahead of use, so that no global-memory latency sits on the wavefront's
    // per-step critical path (each lane only ever reads back what it requested itself)
    extern __shared__ __align__(16) float gsm[];
    constexpr int EOFF = (KG + 2 + 3) & ~3;        // lane 0: the left-edge record (M, U, L, pad) of the row
    constexpr int SLOTP = EOFF + 4;                // floats per lane per slot, 16 B aligned
    const int L1 = a.L1, L2 = a.L2, W = L2 + 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const float NINF = -INFINITY;
    float* ring = gsm + (size_t)wib * (GEN_R * 32 * SLOTP) + lane * SLOTP;     // + slot * 32 * SLOTP
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const bool vec = (KG % 4 == 0) && (a.m_pitch % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.m) & 15) == 0);

    for (int strip = gw; strip < a.n_strips; strip += nw) {
        const int x0 = strip * (32 * KG) + lane * KG + 1;   // my first column (1-based)
        const float* ein = a.edge + (size_t)strip * (L1 + 1) * 4;
        float* eout = a.edge + (size_t)(strip + 1) * (L1 + 1) * 4;
        volatile int* pin = a.progress + strip;
        volatile int* pout = a.progress + strip + 1;
        const bool last_strip = strip == a.n_strips - 1;
        const bool lane_on = x0 <= L2;

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
        float Md = NINF, Ud = NINF, Ld = NINF;
        if (lane_on) { Md = a.top[x0 - 1]; Ud = a.top[W + x0 - 1]; Ld = a.top[2 * W + x0 - 1]; }
        float Me = 0.f, Ue = 0.f, Le = 0.f;
        int avail = 0;
        float bv = NINF; uint32_t bl = 0xffffffffu;
        const int T = L1 + 31;

        // request everything row yy needs into its ring slot; always commits a group
        auto request = [&](int yy) {
            if (yy >= 1 && yy <= L1 && lane_on) {
                const uint32_t dst = ring_s + (uint32_t)((yy & (GEN_R - 1)) * 32 * SLOTP) * 4u;
                const float* mrow = a.m + (size_t)(yy - 1) * a.m_pitch + (x0 - 1);
                if (vec && x0 + KG - 1 <= L2) {
#pragma unroll
                    for (int k = 0; k < KG; k += 4) cp_async16(dst + k * 4, mrow + k);
                } else {
#pragma unroll
                    for (int k = 0; k < KG; k++) if (x0 + k <= L2) cp_async4(dst + k * 4, mrow + k);
                }
                cp_async4(dst + KG * 4, a.g1 + (size_t)(yy - 1) * 2);
                cp_async4(dst + KG * 4 + 4, a.g1 + (size_t)(yy - 1) * 2 + 1);
                // lane 0: the left strip's edge record, through L2 (.cg): it was written by another SM
                if (lane == 0) cp_async16(dst + EOFF * 4, ein + (size_t)yy * 4);
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // the consumer keeps GEN_R - 1 rows of slack to its producer so that every edge record
        // can be requested ahead of use
        auto wait_rows = [&](int need) {
            if (avail < need) {
                if (lane == 0) {
                    // bounded spin: a protocol bug must fail a test, not hang the GPU.  ld.acquire
                    // pairs with the producer's st.release; lane 0 is also the only reader of the edge
                    for (long long spins = 0; spins < (1ll << 26); spins++) {
                        asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(avail) : "l"(const_cast<int*>(pin)) : "memory");
                        if (avail >= need) break;
                        __nanosleep(32);
                    }
                    if (avail < need) atomicOr(a.err, 1);     // timed out: flag it (k_gen_finalize -> NaN score)
                }
                avail = __shfl_sync(FULL, avail, 0);
            }
        };
        __syncwarp();
        wait_rows(min(GEN_R - 1, L1));
#pragma unroll 1
        for (int d = 0; d < GEN_R - 1; d++) request(1 - lane + d);
        uint32_t cz = 0, nz = 0;
        auto fetch_mask = [&](int yy) -> uint32_t {
            uint32_t zz = 0;
            if (MASK && yy >= 1 && yy <= L1 && lane_on) {
#pragma unroll
                for (int k = 0; k < KG; k++)
                    if (x0 + k <= L2 && a.z[(size_t)yy * a.z_pitch + x0 + k]) zz |= 1u << k;
            }
            return zz;
        };
        cz = fetch_mask(1 - lane);

        for (int t = 0; t < T; t++) {
            __syncwarp();                               // lanes took different branches last step
            const int y = t - lane + 1;
            const bool active = (y >= 1) && (y <= L1) && lane_on;
            wait_rows(min(t + GEN_R, L1));              // lane 0 is about to request row t + GEN_R
            request(y + GEN_R - 1);
            if (MASK) nz = fetch_mask(y + 1);
            float Ml = __shfl_up_sync(FULL, Me, 1);
            float Ul = __shfl_up_sync(FULL, Ue, 1);
            float Ll = __shfl_up_sync(FULL, Le, 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(GEN_R - 1) : "memory");   // row y has landed
            const float* slot = ring + (y & (GEN_R - 1)) * 32 * SLOTP;
            if (lane == 0 && y <= L1) { Ml = slot[EOFF]; Ul = slot[EOFF + 1]; Ll = slot[EOFF + 2]; }
            if (active) {
                const float cg1o = slot[KG], cg1e = slot[KG + 1];
                const float dM = Ml, dU = Ul, dL = Ll;
                float cMl = Ml, cLl = Ll;
                uint32_t fl[KG];
#pragma unroll
                for (int k = 0; k < KG; k++) {
                    const float s = slot[k];
                    const float mm = Md + s, mu = Ud + s, ml = Ld + s;
                    float M = fmaxf(fmaxf(mm, mu), ml);
                    if (LOCAL) M = fmaxf(M, 0.f);
                    const float uo = Mp[k] + cg1o, ue = Up[k] + cg1e;
                    const float lo = cMl + go2[k], le = cLl + ge2[k];
                    float U = fmaxf(uo, ue), L = fmaxf(lo, le);
                    uint32_t f = 0;
                    if (LOCAL) f = __funnelshift_l(__float_as_uint(ml - M), f, 1);
                    f = __funnelshift_l(__float_as_uint(lo - L), f, 1);
                    f = __funnelshift_l(__float_as_uint(uo - U), f, 1);
                    f = __funnelshift_l(__float_as_uint(mu - M), f, 1);
                    f = __funnelshift_l(__float_as_uint(mm - M), f, 1);
                    if (MASK && ((cz >> k) & 1u)) { M = 0.f; U = 0.f; L = 0.f; f = 0x80u; }
                    Md = Mp[k]; Ud = Up[k]; Ld = Lp[k];
                    if (x0 + k <= L2) {
                        Mp[k] = M; Up[k] = U; Lp[k] = L;
                        cMl = M; cLl = L;
                        if (LOCAL) {
                            const float v3 = fmaxf(fmaxf(M, U), L);
                            if (v3 >= bv) {
                                const uint32_t lin = (uint32_t)(((size_t)y * W + x0 + k) * 3);
                                const float vs[3] = {M, U, L};
#pragma unroll
                                for (int j = 0; j < 3; j++)
                                    if (vs[j] > bv || (vs[j] == bv && lin + j < bl)) { bv = vs[j]; bl = lin + j; }
                            }
                        }
                    }
                    fl[k] = f;
                }
                Md = dM; Ud = dU; Ld = dL;
                {
                    uint8_t* frow = a.flags + (size_t)y * a.f_pitch + x0;
#pragma unroll
                    for (int k = 0; k < KG; k++) if (x0 + k <= L2) frow[k] = (uint8_t)fl[k];
                }
                if (y == L1) {
#pragma unroll
                    for (int k = 0; k < KG; k++)
                        if (x0 + k <= L2) { a.lastrow[x0 + k] = Mp[k]; a.lastrow[W + x0 + k] = Up[k]; a.lastrow[2 * W + x0 + k] = Lp[k]; }
                }
                if (last_strip) {
#pragma unroll
                    for (int k = 0; k < KG; k++)
                        if (x0 + k == L2) { a.lastcol[y] = Mp[k]; a.lastcol[(L1 + 1) + y] = Up[k]; a.lastcol[2 * (L1 + 1) + y] = Lp[k]; }
                }
                Me = Mp[KG - 1]; Ue = Up[KG - 1]; Le = Lp[KG - 1];
                if (lane == 31 && !last_strip) { float* e = eout + (size_t)y * 4; e[0] = Me; e[1] = Ue; e[2] = Le; }
            }
            if (MASK) cz = nz;
            // lane 31 wrote the edge records; its release store orders them before the counter
            const int done = t - 31 + 1;
            if (!last_strip && lane == 31 && done >= 1 && ((done & 7) == 0 || done == L1))
                asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(const_cast<int*>(pout)), "r"(done) : "memory");
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (LOCAL && bl != 0xffffffffu) atomicMax(a.best, gkey(bv, bl));
    }
}


// ---- K3, lean form ------------------------------------------------------------------------------
// The production wavefront when the match-score matrix has a padded pitch (the engine allocates
// it so): strips of 128 columns (4 per lane), a step loop in the style of K2 -- every pointer
// hoisted, no per-column validity tests (pad columns compute harmlessly inside the padded pitch),
// constant gap pairs as scalars when the gap models are constant (VARG = false), one uniform
// ring slot per step, the five sign bits of a lane's four cells packed into ONE 32-bit word that
// is stored in the skewed [strip][step][lane] layout (a coalesced 128-byte line per warp-step).
//
// Strip-to-strip handoff without fences: the right edge of a strip is a stream of 32-byte
// records {M, tag, U, tag | L, tag, 0, 0} with tag = row number (the buffer is zeroed before the
// launch, rows start at 1).  Every 8-byte {value, tag} pair is written and read as part of one
// aligned vector access, so a record is valid exactly when its three tags match -- no release
// store, no membar, no progress counter (the first version spent a third of its cycles in the
// ERRBAR of a predicated st.release, which drained the whole prefetch ring every step).  The
// consumer prefetches records WV_R-1 rows ahead with cp.async.cg and checks the tags at use; on
// a miss it waits until its producer is a full ring ahead and reloads the ring's records directly.
#define WV_R 8
// every {value, tag} pair is ONE 64-bit scalar access (single-copy atomic); ptxas is free to split
// a vector access into its elements, so the pairs are never loaded or stored as f32 vectors
__device__ __forceinline__ void ld_edge(const float* p, float4& e0, float4& e1)
{
    unsigned long long a, b, c;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(a) : "l"(p) : "memory");
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(b) : "l"(p + 2) : "memory");
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(c) : "l"(p + 4) : "memory");
    e0.x = __uint_as_float((uint32_t)a); e0.y = __uint_as_float((uint32_t)(a >> 32));
    e0.z = __uint_as_float((uint32_t)b); e0.w = __uint_as_float((uint32_t)(b >> 32));
    e1.x = __uint_as_float((uint32_t)c); e1.y = __uint_as_float((uint32_t)(c >> 32));
    e1.z = 0.f; e1.w = 0.f;
}
__device__ __forceinline__ void st_edge(float* p, float M, float U, float L, int y)
{
    const unsigned long long tg = (unsigned long long)(uint32_t)y << 32;
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(tg | __float_as_uint(M)) : "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p + 2), "l"(tg | __float_as_uint(U)) : "memory");
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p + 4), "l"(tg | __float_as_uint(L)) : "memory");
}
__device__ __forceinline__ bool edge_ok(const float4& e0, const float4& e1, int y)
{
    const uint32_t tg = (uint32_t)y;
    return __float_as_uint(e0.y) == tg && __float_as_uint(e0.w) == tg && __float_as_uint(e1.y) == tg;
}
// lane 0, record of row y not there yet: let the producer get a full ring ahead, then put the
// records of rows y .. y+WV_R-1 (all requested already, possibly too early) into their slots
__device__ __forceinline__ void wave_refill(const float* ein, float* ring_lane0, int slotp, int y, int t, int L1, int* err)
{
    asm volatile("cp.async.wait_group 0;" ::: "memory");      // nothing in flight may land on top of the fix-up
    const int target = min(y + WV_R - 1, L1);
    float4 e0, e1;
    bool ok = false;
    for (long long spins = 0; spins < (1ll << 24); spins++) {  // bounded: never hang ...
        ld_edge(ein + (size_t)target * 8, e0, e1);
        if ((ok = edge_ok(e0, e1, target))) break;
        __nanosleep(64);
    }
    if (!ok) atomicOr(err, 1);      // ... but never continue silently either: k_gen_finalize turns this into a NaN score
    for (int r = y; r <= target; r++) {
        ok = false;
        for (int spins = 0; spins < (1 << 20); spins++) {
            ld_edge(ein + (size_t)r * 8, e0, e1);
            if ((ok = edge_ok(e0, e1, r))) break;
        }
        if (!ok) atomicOr(err, 1);
        float* s = ring_lane0 + ((t + (r - y)) & (WV_R - 1)) * 32 * slotp;
        *reinterpret_cast<float4*>(s + 4) = e0;
        *reinterpret_cast<float4*>(s + 8) = e1;
    }
}

template <bool LOCAL, bool MASK, bool VARG>
__global__ void __launch_bounds__(256) k_wave(const GenArgs a)
{
    extern __shared__ __align__(16) float wsm[];
    constexpr int SLOTP = VARG ? 16 : 12;   // floats per lane per slot: [m x4][edge M tag U tag][edge L tag 0 0][gap open, extend, pad x2]
    const int L1 = a.L1, L2 = a.L2, W = L2 + 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const float NINF = -INFINITY;
    float* ring = wsm + (size_t)wib * (WV_R * 32 * SLOTP) + lane * SLOTP;
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const int TT = L1 + 31;                 // steps per strip == rows of its flag-word block

    for (int strip = gw; strip < a.n_strips; strip += nw) {
        const int x0 = strip * 128 + lane * 4 + 1;
        const float* ein = a.edge + (size_t)strip * (L1 + 1) * 8;
        float* eout = a.edge + (size_t)(strip + 1) * (L1 + 1) * 8;
        const bool last_strip = strip == a.n_strips - 1;
        const bool lane_on = x0 <= L2;
        uint32_t* fout = a.flagw + (size_t)strip * TT * 32 + lane;

        float Mp[4], Up[4], Lp[4], go2[4], ge2[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = x0 + k;
            const bool v = x <= L2;
            Mp[k] = v ? a.top[x] : NINF;
            Up[k] = v ? a.top[W + x] : NINF;
            Lp[k] = v ? a.top[2 * W + x] : NINF;
            go2[k] = (VARG && v) ? a.g2[(size_t)(x - 1) * 2] : a.g2[0];
            ge2[k] = (VARG && v) ? a.g2[(size_t)(x - 1) * 2 + 1] : a.g2[1];
        }
        const float cgo1 = a.g1[0], cge1 = a.g1[1];
        float Md = NINF, Ud = NINF, Ld = NINF;
        if (lane_on) { Md = a.top[x0 - 1]; Ud = a.top[W + x0 - 1]; Ld = a.top[2 * W + x0 - 1]; }
        float Me = 0.f, Ue = 0.f, Le = 0.f;
        float bv = NINF; uint32_t bl = 0xffffffffu;
        const int lcol = (L2 - 1 - strip * 128) >> 2, kcol = (L2 - 1) & 3;   // lane / cell of column L2 (last strip)

        // row yy -> ring slot of the step that consumes it (uniform slot index per step)
        auto request = [&](int yy, int use_step) {
            if (yy >= 1 && yy <= L1 && lane_on) {
                const uint32_t dst = ring_s + (uint32_t)((use_step & (WV_R - 1)) * 32 * SLOTP) * 4u;
                cp_async16(dst, a.m + (size_t)(yy - 1) * a.m_pitch + (x0 - 1));
                if (lane == 0) {
                    cp_async16(dst + 16, ein + (size_t)yy * 8);
                    cp_async16(dst + 32, ein + (size_t)yy * 8 + 4);
                }
                if (VARG) {
                    cp_async4(dst + 48, a.g1 + (size_t)(yy - 1) * 2);
                    cp_async4(dst + 52, a.g1 + (size_t)(yy - 1) * 2 + 1);
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        auto fetch_mask = [&](int yy) -> uint32_t {
            uint32_t zz = 0;
            if (MASK && yy >= 1 && yy <= L1 && lane_on) {
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (x0 + k <= L2 && a.z[(size_t)yy * a.z_pitch + x0 + k]) zz |= 1u << k;
            }
            return zz;
        };
        __syncwarp();
#pragma unroll 1
        for (int d = 0; d < WV_R - 1; d++) request(1 - lane + d, d);
        uint32_t cz = fetch_mask(1 - lane), nz = 0;

        for (int t = 0; t < TT; t++) {
            __syncwarp();
            const int y = t - lane + 1;
            request(y + WV_R - 1, t + WV_R - 1);
            if (MASK) nz = fetch_mask(y + 1);
            float Ml = __shfl_up_sync(FULL, Me, 1);
            float Ul = __shfl_up_sync(FULL, Ue, 1);
            float Ll = __shfl_up_sync(FULL, Le, 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(WV_R - 1) : "memory");
            float* slot = ring + (t & (WV_R - 1)) * 32 * SLOTP;
            // lane 0: the left strip's record of this row.  The miss path is entered by the whole
            // warp (vote), so the step loop itself never runs diverged.
            float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
            bool miss = false;
            if (lane == 0 && y <= L1) {
                e0 = *reinterpret_cast<const float4*>(slot + 4);
                e1 = *reinterpret_cast<const float4*>(slot + 8);
                miss = !edge_ok(e0, e1, y);
            }
            if (__any_sync(FULL, miss)) {
                if (lane == 0) {
#ifdef WAVE_DEBUG
                    atomicAdd(a.progress + strip, 1);
#endif
                    wave_refill(ein, ring, SLOTP, y, t, L1, a.err);
                    asm volatile("" ::: "memory");
                    e0 = *reinterpret_cast<const float4*>(slot + 4);
                    e1 = *reinterpret_cast<const float4*>(slot + 8);
                }
                __syncwarp();
            }
            if (lane == 0) { Ml = e0.x; Ul = e0.z; Ll = e1.x; }
            if (y >= 1 && y <= L1 && lane_on) {
                const float4 sv = *reinterpret_cast<const float4*>(slot);
                const float sc[4] = {sv.x, sv.y, sv.z, sv.w};
                const float g1o = VARG ? slot[12] : cgo1, g1e = VARG ? slot[13] : cge1;
                const float dM = Ml, dU = Ul, dL = Ll;
                float cMl = Ml, cLl = Ll;
                uint32_t fw = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const float mm = Md + sc[k], mu = Ud + sc[k], ml = Ld + sc[k];
                    float M = fmaxf(fmaxf(mm, mu), ml);
                    if (LOCAL) M = fmaxf(M, 0.f);
                    const float uo = Mp[k] + g1o, ue = Up[k] + g1e;
                    const float lo = cMl + go2[k], le = cLl + ge2[k];
                    float U = fmaxf(uo, ue), L = fmaxf(lo, le);
                    uint32_t f = 0;
                    if (LOCAL) f = __funnelshift_l(__float_as_uint(ml - M), f, 1);
                    f = __funnelshift_l(__float_as_uint(lo - L), f, 1);
                    f = __funnelshift_l(__float_as_uint(uo - U), f, 1);
                    f = __funnelshift_l(__float_as_uint(mu - M), f, 1);
                    f = __funnelshift_l(__float_as_uint(mm - M), f, 1);
                    if (MASK && ((cz >> k) & 1u)) { M = 0.f; U = 0.f; L = 0.f; f = 0; fw |= 1u << (24 + k); }
                    fw |= f << (5 * k);
                    Md = Mp[k]; Ud = Up[k]; Ld = Lp[k];
                    Mp[k] = M; Up[k] = U; Lp[k] = L;
                    cMl = M; cLl = L;
                    if (LOCAL && x0 + k <= L2) {
                        const float v3 = fmaxf(fmaxf(M, U), L);
                        if (v3 >= bv) {
                            const uint32_t lin = (uint32_t)(((size_t)y * W + x0 + k) * 3);
                            const float vs[3] = {M, U, L};
#pragma unroll
                            for (int j = 0; j < 3; j++)
                                if (vs[j] > bv || (vs[j] == bv && lin + j < bl)) { bv = vs[j]; bl = lin + j; }
                        }
                    }
                }
                Md = dM; Ud = dU; Ld = dL;
                fout[(size_t)t * 32] = fw;
                Me = Mp[3]; Ue = Up[3]; Le = Lp[3];
                if (y == L1) {
#pragma unroll
                    for (int k = 0; k < 4; k++)
                        if (x0 + k <= L2) { a.lastrow[x0 + k] = Mp[k]; a.lastrow[W + x0 + k] = Up[k]; a.lastrow[2 * W + x0 + k] = Lp[k]; }
                }
                if (last_strip) {
                    if (lane == lcol) {
                        float m_ = Mp[0], u_ = Up[0], l_ = Lp[0];
#pragma unroll
                        for (int k = 1; k < 4; k++) if (kcol == k) { m_ = Mp[k]; u_ = Up[k]; l_ = Lp[k]; }
                        a.lastcol[y] = m_; a.lastcol[(L1 + 1) + y] = u_; a.lastcol[2 * (L1 + 1) + y] = l_;
                    }
                } else if (lane == 31) {
                    st_edge(eout + (size_t)y * 8, Me, Ue, Le, y);
                }
            }
            if (MASK) cz = nz;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        if (LOCAL && bl != 0xffffffffu) atomicMax(a.best, gkey(bv, bl));
#ifdef WAVE_DEBUG
        if (lane == 0) printf("strip %d refills %d\n", strip, a.progress[strip]);
#endif
    }
}

// ---- K3, row-blocked form: 4 rows x 4 columns per lane and step ----------------------------------------
// k_wave advances ONE row per step: three shuffles, a cp.async wait, a tag check and a vote sit between every
// two rows of a strip, and a 20k x 20k matrix is 20,000 + 32 * 157 such steps of ~600 ns (one warp per
// scheduler, IPC 0.2: a dependent chain, profiles/r01_kwave_v3_*).  Here a lane owns a 4 x 4 block per step:
// lane l works on rows 4(t - l) + 1 .. + 4 of its four columns at step t, the twelve edge values of a block
// cross to lane l + 1 by shuffles issued back to back, lane 0 reads four tagged records of the left strip, and
// the sixteen cells of a block hold an anti-diagonal of independent work (critical path 7 cells of 16).  Steps
// per strip: L1 / 4 + 31; the critical path of the whole matrix L1 / 4 + 32 * strips steps.  Same arithmetic as
// k_wave (three separate sums, sign-bit flags, cext.c:155-290), same tagged 32-byte records, same flag words
// (one per row and lane, row y of lane l at word row y - 1 + 4 l).  Global and semiglobal modes, constant gap
// pairs, no mask; everything else runs k_wave.
#define W4_R 4                         // ring depth in steps: scores and records are requested 3 steps ahead
#define W4_SLOTF (32 * 16 + 32)        // floats per ring slot: scores [row][lane][4] + four records [8]
__global__ void __launch_bounds__(256, 1) k_wave4(const GenArgs a)
{
    extern __shared__ __align__(16) float wsm[];
    const int L1 = a.L1, L2 = a.L2, W = L2 + 1;
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    const float NINF = -INFINITY;
    float* ring = wsm + (size_t)wib * (W4_R * W4_SLOTF);
    const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
    const int NB = (L1 + 3) >> 2;            // row blocks
    const int TT = NB + 31;                  // steps per strip
    const float g1o = a.g1[0], g1e = a.g1[1], g2o = a.g2[0], g2e = a.g2[1];

    for (int strip = gw; strip < a.n_strips; strip += nw) {
        const int x0 = strip * 128 + lane * 4 + 1;
        const float* ein = a.edge + (size_t)strip * (L1 + 1) * 8;
        float* eout = a.edge + (size_t)(strip + 1) * (L1 + 1) * 8;
        const bool last_strip = strip == a.n_strips - 1;
        const bool lane_on = x0 <= L2;
        uint32_t* fout = a.flagw + (size_t)strip * (TT * 4) * 32 + lane;
        const int lcol = (L2 - 1 - strip * 128) >> 2, kcol = (L2 - 1) & 3;   // lane / cell of column L2 (last strip)

        float Mp[4], Up[4], Lp[4];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            const int x = x0 + k;
            const bool v = x <= L2;
            Mp[k] = v ? a.top[x] : NINF;
            Up[k] = v ? a.top[W + x] : NINF;
            Lp[k] = v ? a.top[2 * W + x] : NINF;
        }
        // the left neighbour's last column in the row above the block (row 0: the top border)
        float e0M = NINF, e0U = NINF, e0L = NINF;
        if (lane_on) { e0M = a.top[x0 - 1]; e0U = a.top[W + x0 - 1]; e0L = a.top[2 * W + x0 - 1]; }
        float rM[4], rU[4], rL[4];              // my last column, rows of the block just finished
#pragma unroll
        for (int r = 0; r < 4; r++) { rM[r] = 0.f; rU[r] = 0.f; rL[r] = 0.f; }

        // block bb -> ring slot of the step that consumes it.  Branch-free: rows are clamped into the matrix (blocks
        // outside it load a row nobody uses; the padded pitch covers every lane's columns).  The four 32-byte records
        // lane 0 will need (block bb0 of the left strip's edge) are eight 16-byte pieces: lane p & 7 copies piece p --
        // one cp.async instruction for the warp (lanes 8..31 repeat the pieces of lanes 0..7: same bytes, same place).
        const int pc = lane & 7;                  // my piece: row pc >> 1 of the block, half pc & 1 of its record
        // byte pointers and byte strides: one IMAD.WIDE per address (a float index costs a multiply-add plus a two-instruction
        // scale-and-add)
        const char* mcolb = reinterpret_cast<const char*>(a.m + (x0 - 1));
        const char* einb = reinterpret_cast<const char*>(ein) + (pc & 1) * 16;
        const int pitchb = a.m_pitch * 4;
        const uint32_t dst_lane = ring_s + (uint32_t)(lane * 16);
        auto request = [&](int bb, int bb0, int use_step) {
            const uint32_t slot_b = (uint32_t)((use_step & (W4_R - 1)) * (W4_SLOTF * 4));
            const uint32_t dst = dst_lane + slot_b;
            const int yb = 4 * bb;
#pragma unroll
            for (int r = 0; r < 4; r++) {
                const int y0 = min(max(yb + r, 0), L1 - 1);
                cp_async16(dst + (uint32_t)(r * 512), mcolb + (int64_t)y0 * pitchb);
            }
            const int yr = min(max(4 * bb0 + 1 + (pc >> 1), 1), L1);
            cp_async16(ring_s + slot_b + (uint32_t)(2048 + pc * 16), einb + (int64_t)yr * 32);
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        // K1 beside the fill (GenArgs.m_ready): the 128 x 128 block of m that holds the rows lane 0 is about to request --
        // the furthest of the warp -- must be complete.  One acquire load per 32 steps once K1 is ahead, which it is after
        // the first blocks (it needs a fifth of the fill's time for the whole matrix).
        const int* mrd = a.m_ready ? a.m_ready + strip : nullptr;
        int yb_ok = -1;
        auto wait_rows = [&](int bb) {
            const int ybn = min(max(bb, 0), NB - 1) >> 5;
            if (mrd == nullptr || ybn <= yb_ok) return;
            const int* f = mrd + (size_t)ybn * a.m_ready_nx;
            int v = 0;
#pragma unroll 1
            for (int spins = 0; spins < (1 << 22); spins++) {       // bounded: never hang ...
                asm volatile("ld.acquire.gpu.global.b32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                if (v) break;
                __nanosleep(200);
            }
            if (!v && lane == 0) atomicOr(a.err, 1);                // ... and never continue silently
            yb_ok = ybn;
            __syncwarp();
        };
        __syncwarp();
        wait_rows(W4_R - 2);
#pragma unroll 1
        for (int d = 0; d < W4_R - 1; d++) request(d - lane, d, d);

        for (int t = 0; t < TT; t++) {
            __syncwarp();
            const int b = t - lane;
            if (mrd != nullptr && ((t + W4_R - 1) & 31) == 0) wait_rows(t + W4_R - 1);     // a new 128-row block of m
            request(b + W4_R - 1, t + W4_R - 1, t + W4_R - 1);
            float nM[4], nU[4], nL[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                nM[r] = __shfl_up_sync(FULL, rM[r], 1);
                nU[r] = __shfl_up_sync(FULL, rU[r], 1);
                nL[r] = __shfl_up_sync(FULL, rL[r], 1);
            }
            asm volatile("cp.async.wait_group %0;" ::"n"(W4_R - 1) : "memory");
            float* slot = ring + (t & (W4_R - 1)) * W4_SLOTF;
            // lane 0: the left strip's four records of this block.  Every lane loads them (one broadcast each)
            // and the miss path is entered by the whole warp through a vote, so that lane 0 never runs the step
            // loop as a warp of its own (independent thread scheduling keeps a lane that took a long private
            // branch apart from the others: the first version of this kernel ran every step twice).
            const int nrow = min(4, L1 - 4 * t);      // valid rows of lane 0's block (b = t there)
            // validity: lane p < 8 tests the tags of ITS piece (first half of a record: the M and U tags, second half:
            // the L tag); the values go to lane 0 (every lane loads them: broadcasts, no divergence)
            bool miss;
            {
                const float4 pv = *reinterpret_cast<const float4*>(slot + 512 + pc * 4);
                const uint32_t tg = (uint32_t)(4 * t + 1 + (pc >> 1));
                const bool okp = (pc & 1) ? (__float_as_uint(pv.y) == tg)
                                          : (__float_as_uint(pv.y) == tg && __float_as_uint(pv.w) == tg);
                miss = (lane < 8) && ((pc >> 1) < nrow) && !okp;
            }
            float4 e0[4], e1[4];
#pragma unroll
            for (int r = 0; r < 4; r++) {
                e0[r] = *reinterpret_cast<const float4*>(slot + 512 + r * 8);
                e1[r] = *reinterpret_cast<const float4*>(slot + 512 + r * 8 + 4);
            }
            if (__any_sync(FULL, miss)) {
                // The records were requested before the left strip had written them.  Fetch the four of THIS block
                // directly (all loads in flight together), as often as it takes; records requested for later steps
                // are checked at their own steps.  While a strip runs in this mode its steps carry an L2 round trip
                // more than its producer's, so the producer pulls ahead until the prefetches hit again: no waiting
                // for a full ring (the first version did, and every strip started 60-80 steps behind its neighbour
                // instead of 35).
                if (lane == 0) {
                    bool ok = false;
                    for (int spins = 0; spins < (1 << 22) && !ok; spins++) {     // bounded: never hang ...
                        ok = true;
#pragma unroll
                        for (int r = 0; r < 4; r++) ld_edge(ein + (size_t)min(4 * t + 1 + r, L1) * 8, e0[r], e1[r]);
#pragma unroll
                        for (int r = 0; r < 4; r++) ok = ok && (r >= nrow || edge_ok(e0[r], e1[r], 4 * t + 1 + r));
                    }
                    if (!ok) atomicOr(a.err, 1);        // ... and never continue silently: k_gen_finalize -> NaN score
                }
                __syncwarp();
            }
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < 4; r++) { nM[r] = e0[r].x; nU[r] = e0[r].z; nL[r] = e1[r].x; }
            }
            __syncwarp();
            if (t < NB - 1) {
                // ---- interior steps: every active lane holds a full block that is not its last one.  Straight-line
                // code, cells in ANTI-DIAGONAL order (the issue order is the program order: a row-major block runs as
                // four dependent chains one after the other), nothing conditional between the cells.
                if (b >= 0 && lane_on) {
                    float sc[4][4];
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        const float4 sv = *reinterpret_cast<const float4*>(slot + r * 128 + lane * 4);
                        sc[r][0] = sv.x; sc[r][1] = sv.y; sc[r][2] = sv.z; sc[r][3] = sv.w;
                    }
                    float Mc[4][4], Uc[4][4], Lc[4][4];
                    uint32_t fw[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int d = 0; d < 7; d++) {
#pragma unroll
                        for (int r = 0; r < 4; r++) {
                            const int k = d - r;
                            if (k < 0 || k > 3) continue;
                            // neighbours: above (r-1, k), left (r, k-1), diagonal (r-1, k-1)
                            const float aM = r ? Mc[r - 1][k] : Mp[k], aU = r ? Uc[r - 1][k] : Up[k];
                            const float lM = k ? Mc[r][k - 1] : nM[r], lL = k ? Lc[r][k - 1] : nL[r];
                            float dM_, dU_, dL_;
                            if (r && k) { dM_ = Mc[r - 1][k - 1]; dU_ = Uc[r - 1][k - 1]; dL_ = Lc[r - 1][k - 1]; }
                            else if (r) { dM_ = nM[r - 1]; dU_ = nU[r - 1]; dL_ = nL[r - 1]; }
                            else if (k) { dM_ = Mp[k - 1]; dU_ = Up[k - 1]; dL_ = Lp[k - 1]; }
                            else { dM_ = e0M; dU_ = e0U; dL_ = e0L; }
                            const float s_ = sc[r][k];
                            const float mm = dM_ + s_, mu = dU_ + s_, ml = dL_ + s_;
                            const float M = fmaxf(fmaxf(mm, mu), ml);
                            const float uo = aM + g1o, ue = aU + g1e;
                            const float lo = lM + g2o, le = lL + g2e;
                            const float U = fmaxf(uo, ue), L = fmaxf(lo, le);
                            // four sign bits per cell pushed straight into the row's word (cells of a row are issued in
                            // column order): cell k ends up in bits 4 (3 - k) .. + 3, no shift / OR per cell
                            fw[r] = __funnelshift_l(__float_as_uint(lo - L), fw[r], 1);
                            fw[r] = __funnelshift_l(__float_as_uint(uo - U), fw[r], 1);
                            fw[r] = __funnelshift_l(__float_as_uint(mu - M), fw[r], 1);
                            fw[r] = __funnelshift_l(__float_as_uint(mm - M), fw[r], 1);
                            Mc[r][k] = M; Uc[r][k] = U; Lc[r][k] = L;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; r++) {
                        fout[(size_t)(4 * t + r) * 32] = fw[r];
                        rM[r] = Mc[r][3]; rU[r] = Uc[r][3]; rL[r] = Lc[r][3];
                    }
#pragma unroll
                    for (int k = 0; k < 4; k++) { Mp[k] = Mc[3][k]; Up[k] = Uc[3][k]; Lp[k] = Lc[3][k]; }
                    e0M = nM[3]; e0U = nU[3]; e0L = nL[3];
                    if (last_strip) {
                        if (lane == lcol) {
#pragma unroll
                            for (int r = 0; r < 4; r++) {
                                float m_ = Mc[r][0], u_ = Uc[r][0], l_ = Lc[r][0];
#pragma unroll
                                for (int k = 1; k < 4; k++) if (kcol == k) { m_ = Mc[r][k]; u_ = Uc[r][k]; l_ = Lc[r][k]; }
                                const int y = 4 * b + 1 + r;
                                a.lastcol[y] = m_; a.lastcol[(L1 + 1) + y] = u_; a.lastcol[2 * (L1 + 1) + y] = l_;
                            }
                        }
                    } else if (lane == 31) {
#pragma unroll
                        for (int r = 0; r < 4; r++) st_edge(eout + (size_t)(4 * b + 1 + r) * 8, rM[r], rU[r], rL[r], 4 * b + 1 + r);
                    }
                }
            } else
            if (b >= 0 && b < NB && lane_on) {
                float dM = e0M, dU = e0U, dL = e0L;
#pragma unroll
                for (int r = 0; r < 4; r++) {
                    const int y = 4 * b + 1 + r;
                    const float4 sv = *reinterpret_cast<const float4*>(slot + r * 128 + lane * 4);
                    const float sc[4] = {sv.x, sv.y, sv.z, sv.w};
                    float cMl = nM[r], cLl = nL[r];
                    float Md = dM, Ud = dU, Ld = dL;
                    uint32_t fw = 0;
#pragma unroll
                    for (int k = 0; k < 4; k++) {
                        const float mm = Md + sc[k], mu = Ud + sc[k], ml = Ld + sc[k];
                        const float M = fmaxf(fmaxf(mm, mu), ml);
                        const float uo = Mp[k] + g1o, ue = Up[k] + g1e;
                        const float lo = cMl + g2o, le = cLl + g2e;
                        const float U = fmaxf(uo, ue), L = fmaxf(lo, le);
                        fw = __funnelshift_l(__float_as_uint(lo - L), fw, 1);
                        fw = __funnelshift_l(__float_as_uint(uo - U), fw, 1);
                        fw = __funnelshift_l(__float_as_uint(mu - M), fw, 1);
                        fw = __funnelshift_l(__float_as_uint(mm - M), fw, 1);
                        Md = Mp[k]; Ud = Up[k]; Ld = Lp[k];
                        Mp[k] = M; Up[k] = U; Lp[k] = L;
                        cMl = M; cLl = L;
                    }
                    dM = nM[r]; dU = nU[r]; dL = nL[r];        // the diagonal of the next row's first cell
                    fout[(size_t)(4 * t + r) * 32] = fw;
                    rM[r] = Mp[3]; rU[r] = Up[3]; rL[r] = Lp[3];
                    if (y == L1) {
#pragma unroll
                        for (int k = 0; k < 4; k++)
                            if (x0 + k <= L2) { a.lastrow[x0 + k] = Mp[k]; a.lastrow[W + x0 + k] = Up[k]; a.lastrow[2 * W + x0 + k] = Lp[k]; }
                    }
                    if (y <= L1) {
                        if (last_strip) {
                            if (lane == lcol) {
                                float m_ = Mp[0], u_ = Up[0], l_ = Lp[0];
#pragma unroll
                                for (int k = 1; k < 4; k++) if (kcol == k) { m_ = Mp[k]; u_ = Up[k]; l_ = Lp[k]; }
                                a.lastcol[y] = m_; a.lastcol[(L1 + 1) + y] = u_; a.lastcol[2 * (L1 + 1) + y] = l_;
                            }
                        } else if (lane == 31) {
                            st_edge(eout + (size_t)y * 8, Mp[3], Up[3], Lp[3], y);
                        }
                    }
                }
                e0M = nM[3]; e0U = nU[3]; e0L = nL[3];
            }
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
    }
}

// Traceback over the lean kernel's flag words: one warp stages WT_WIN steps x 32 lanes of the
// current strip (16 KB, every row a coalesced 128-byte line, all requests in flight at once) and
// lane 0 walks inside that window -- one memory round trip per ~100 path cells.
#define WT_WIN 256      // word rows staged per window (32 KB): a diagonal path crosses 1 + flag_skew / 4 word rows per step
__global__ void __launch_bounds__(32) k_wave_traceback(const GenArgs a)
{
    __shared__ uint32_t tile[WT_WIN][32];
    const int lane = threadIdx.x;
    const int L1 = a.L1, L2 = a.L2, TT = a.flag_rows, SK = a.flag_skew;   // word row of (y, lane) = y - 1 + SK * lane
    const bool nib4 = SK == 4;      // k_wave4's words: four bits per cell, cell 0 in the highest nibble; k_wave's: 5 bits, cell 0 lowest
    const bool u_ramp = !(a.mode == PG_SG_BOTH || a.mode == PG_SG_ONE);
    const bool l_ramp = !(a.mode == PG_SG_BOTH || a.mode == PG_SG_TWO);
    int y = a.cell_out[0], x = a.cell_out[1], k = a.cell_out[2];
    const int cap = L1 + L2 + 2;
    int w = cap;
    int2* pb = reinterpret_cast<int2*>(a.path_buf);
    auto push = [&](int yy, int xx) { pb[--w] = make_int2(yy, xx); };
    const bool semi = (a.mode >= PG_SG_BOTH);
    if (lane == 0 && semi) {
        if (y != L1) { for (int v = L1; v > y; v--) push(v, x); }
        else if (x != L2) { for (int v = L2; v > x; v--) push(y, v); }
    }
    const uint32_t tile_s = (uint32_t)__cvta_generic_to_shared(&tile[0][0]);
    int done = 0;
    while (!done) {
        int strip = 0, tlo = 0;
        if (y >= 1 && x >= 1) {
            strip = (x - 1) >> 7;
            const int tcur = y - 1 + SK * (((x - 1) & 127) >> 2);
            tlo = max(0, tcur - (WT_WIN - 1));
            // 16-byte copies: lane l moves piece l & 7 of row 4 i + (l >> 3), four whole 128-byte rows per instruction
            // (64 instructions per window instead of 256 four-byte ones)
            const uint32_t* src = a.flagw + ((size_t)strip * TT + tlo) * 32 + (lane & 7) * 4;
            const int nrow = min(WT_WIN, TT - tlo);
#pragma unroll 8
            for (int r = lane >> 3; r < nrow; r += 4)
                cp_async16(tile_s + (uint32_t)(r * 32 + (lane & 7) * 4) * 4u, src + (size_t)r * 32);
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
            // interior cells first, as a select chain without data-dependent branches (the walk is one dependent
            // chain: every resolved branch costs more than the arithmetic it skips); borders, masked cells and
            // the local stop code leave this loop and take the general step below
            const int xlo = strip * 128;
            for (;;) {
                bool handled = false;
                // The move out of a cell depends on the STATE alone (diagonal, up, left), the cell's flags only choose the
                // next state: the word of the NEXT cell is requested before the current one is decoded, so that the
                // shared-memory latency runs beside the decode instead of in front of it (the walk is one dependent chain).
                bool ok = y >= 1 && x > xlo;
                int kk = (x - 1) & 3;
                uint32_t word = 0u;
                if (ok) {
                    const int ln = ((x - 1) & 127) >> 2, tn = y - 1 + SK * ln;
                    ok = tn >= tlo;
                    if (ok) word = tile[tn - tlo][ln];
                }
                while (ok) {
                    const int y2 = y - (k != 2), x2 = x - (k != 1);
                    const int ln2 = ((x2 - 1) & 127) >> 2, kk2 = (x2 - 1) & 3, tn2 = y2 - 1 + SK * ln2;
                    const bool ok2 = (y2 >= 1) & (x2 > xlo) & (tn2 >= tlo);
                    const uint32_t word2 = tile[min(max(tn2 - tlo, 0), WT_WIN - 1)][ln2];   // clamped into the tile; used only when ok2
                    const uint32_t c = nib4 ? ((word >> (4 * (3 - kk))) & 15u) : ((word >> (5 * kk)) & 31u);
                    if ((!nib4 & ((word >> (24 + kk)) & 1u)) | ((k == 0) & ((c & 19u) == 19u))) break;   // masked / local stop: general step
                    push(y, x);
                    const int nk0 = !(c & 1u) ? 0 : (!(c & 2u) ? 1 : 2);
                    const int nk1 = (c & 4u) ? 1 : 0, nk2 = (c & 8u) ? 2 : 0;
                    k = (k == 0) ? nk0 : ((k == 1) ? nk1 : nk2);
                    y = y2;
                    x = x2;
                    kk = kk2;
                    word = word2;
                    ok = ok2;
                    handled = true;
                }
                if (handled && y >= 1 && x >= 1) {
                    // left the window or the strip: stage again (unless the stop conditions above broke the loop)
                    const int tn = y - 1 + SK * (((x - 1) & 127) >> 2);
                    if (((x - 1) >> 7) != strip || tn < tlo) break;
                }
                push(y, x);
                int ny = y, nx = x, nk = k;
                bool stop = false;
                if (y >= 1 && x >= 1) {
                    const int ln = ((x - 1) & 127) >> 2, kk = (x - 1) & 3;
                    const uint32_t word = tile[y - 1 + SK * ln - tlo][ln];
                    const uint32_t c = nib4 ? ((word >> (4 * (3 - kk))) & 15u) : ((word >> (5 * kk)) & 31u);
                    if (!nib4 && ((word >> (24 + kk)) & 1u)) stop = true;   // masked cell: no flags
                    else if (k == 0) {
                        ny = y - 1; nx = x - 1;
                        if (!(c & 1)) nk = 0; else if (!(c & 2)) nk = 1; else if (c & 16) stop = true; else nk = 2;
                    } else if (k == 1) { ny = y - 1; nk = (c & 4) ? 1 : 0; }
                    else { nx = x - 1; nk = (c & 8) ? 2 : 0; }
                } else if (y == 0 && x == 0) stop = true;
                else if (x == 0) { if (u_ramp && k == 1) ny = y - 1; else stop = true; }
                else { if (l_ramp && k == 2) nx = x - 1; else stop = true; }
                if (stop) { done = 1; break; }
                y = ny; x = nx; k = nk;
                if (y >= 1 && x >= 1) {   // still inside the staged window of this strip?
                    const int tn = y - 1 + SK * (((x - 1) & 127) >> 2);
                    if (((x - 1) >> 7) != strip || tn < tlo) break;
                }
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

// ---- end cell (reference component/align.py:401-431) -----------------------------------------
// One CTA.  global: first argmax of the three states at (L1, L2).  semiglobal: max of the last
// row vs max of the last column (strict '>' and only when tracing from the row is allowed), then
// the FIRST hit scanning from the far end backwards, states 0,1,2 -- as a parallel reduction:
// every candidate gets the key (ordered value, position from the far end descending, state
// ascending) and the maximum key wins.  local: first argmax of the whole o array in
// (y, x, state) order; the interior part was reduced by the fill, the borders are added here.
__device__ __forceinline__ unsigned long long fkey(float v, uint32_t pos, int k)
{
    uint32_t b = __float_as_uint(v);
    b = (b & 0x80000000u) ? ~b : (b | 0x80000000u);
    return ((unsigned long long)b << 32) | ((unsigned long long)pos << 2) | (unsigned)(3 - k);
}

__global__ void __launch_bounds__(256) k_gen_finalize(const GenArgs a)
{
    __shared__ unsigned long long red[2][256];
    const int L1 = a.L1, L2 = a.L2, W = L2 + 1, H = L1 + 1;
    const int tid = threadIdx.x;
    auto row_at = [&](int x, int k) -> float { return (L1 == 0) ? a.top[k * W + x] : (x == 0 ? a.edge[a.flag_fmt == 2 ? (size_t)L1 * 8 + 2 * k : (size_t)L1 * 4 + k] : a.lastrow[k * W + x]); };
    auto col_at = [&](int y, int k) -> float { return (y == 0) ? a.top[k * W + L2] : (L2 == 0 ? a.edge[a.flag_fmt == 2 ? (size_t)y * 8 + 2 * k : (size_t)y * 4 + k] : a.lastcol[k * H + y]); };
    unsigned long long kr = 0ull, kc = 0ull;
    if (a.mode == PG_LOCAL) {
        // borders of o in linear (y, x, state) order: smaller linear index wins ties
        auto lkey = [&](float v, uint32_t lin) { return gkey(v, lin); };
        for (int x = tid; x <= L2; x += 256)
            for (int k = 0; k < 3; k++) { const unsigned long long c = lkey(a.top[k * W + x], (uint32_t)(x * 3 + k)); if (c > kr) kr = c; }
        for (int y = 1 + tid; y <= L1; y += 256)
            for (int k = 0; k < 3; k++) { const unsigned long long c = lkey(a.edge[a.flag_fmt == 2 ? (size_t)y * 8 + 2 * k : (size_t)y * 4 + k], (uint32_t)(((size_t)y * W) * 3 + k)); if (c > kr) kr = c; }
    } else if (a.mode != PG_GLOBAL) {
        for (int x = tid; x <= L2; x += 256)
            for (int k = 0; k < 3; k++) { const unsigned long long c = fkey(row_at(x, k), (uint32_t)x, k); if (c > kr) kr = c; }
        for (int y = tid; y <= L1; y += 256)
            for (int k = 0; k < 3; k++) { const unsigned long long c = fkey(col_at(y, k), (uint32_t)y, k); if (c > kc) kc = c; }
    }
    red[0][tid] = kr;
    red[1][tid] = kc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (tid < o) {
            if (red[0][tid + o] > red[0][tid]) red[0][tid] = red[0][tid + o];
            if (red[1][tid + o] > red[1][tid]) red[1][tid] = red[1][tid + o];
        }
        __syncthreads();
    }
    if (tid != 0) return;
    int cy = L1, cx = L2, ck = 0;
    float score;
    auto kval = [](unsigned long long k) -> float {
        uint32_t b = (uint32_t)(k >> 32);
        b = (b & 0x80000000u) ? (b & 0x7fffffffu) : ~b;
        return __uint_as_float(b);
    };
    if (a.mode == PG_GLOBAL) {
        for (int k = 1; k < 3; k++) if (row_at(L2, k) > row_at(L2, ck)) ck = k;
        score = row_at(L2, ck);
    } else if (a.mode == PG_LOCAL) {
        unsigned long long best = red[0][0];
        if (*a.best > best) best = *a.best;
        const uint32_t bl = ~(uint32_t)best;
        ck = bl % 3; cx = (bl / 3) % W; cy = bl / 3 / W; score = kval(best);
    } else {
        const unsigned long long br = red[0][0], bc = red[1][0];
        const float rmax = kval(br), cmax = kval(bc);
        const bool from_row = (a.mode == PG_SG_BOTH || a.mode == PG_SG_TWO);
        if (rmax > cmax && from_row) { cy = L1; cx = (int)((br & 0xffffffffull) >> 2); ck = 3 - (int)(br & 3ull); score = rmax; }
        else { cy = (int)((bc & 0xffffffffull) >> 2); cx = L2; ck = 3 - (int)(bc & 3ull); score = cmax; }
    }
    if (*a.err) score = __int_as_float(0x7fc00000);   // a strip hand-off timed out: the host raises on NaN
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
                    if (a.flag_fmt) {   // compact sign bits of k_gen_fill_fast -> the first reference flag
                        const uint8_t c = f;
                        if (c & 0x80) f = 0;
                        else if (k == 0) f = !(c & 1) ? TB_MM : (!(c & 2) ? TB_MU : ((c & 16) ? 0 : TB_ML));
                        else if (k == 1) f = (c & 4) ? TB_UE : TB_UO;
                        else f = (c & 8) ? TB_LE : TB_LO;
                    } else {
                        f &= (k == 0) ? (TB_MM | TB_MU | TB_ML) : (k == 1 ? (TB_UO | TB_UE) : (TB_LO | TB_LE));
                    }
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

// Block = (32*CG) columns x (8*RG) rows of m, 256 threads.  Per track set the block stages its
// rows of P1, its rows of P2 and S in shared memory and compacts the NONZERO entries of every
// staged profile row (ascending index, as build_nonzero_matrix does on the host,
// component/align.py:449-458), so a cell costs nnz1 x nnz2 terms like the reference instead of
// A x A.  Large matrices use CG x RG = 2 x 2 (a thread owns 4 cells and the staging is amortised over
// 1024 cells); small ones 1 x 1 so that the grid still covers the machine.
#define BS_MAXA 64
template <int CG, int RG>
__global__ void __launch_bounds__(256) k_build_scores(const ScoreSets sets, int L1, int L2, float* m, int m_pitch)
{
    extern __shared__ float bsm[];
    constexpr int NR = 8 * RG, NC = 32 * CG;
    const int tx = threadIdx.x, ty = threadIdx.y, tid = ty * 32 + tx;
    const int x0 = blockIdx.x * NC, y0 = blockIdx.y * NR;
    float score[RG][CG];
#pragma unroll
    for (int r = 0; r < RG; r++)
#pragma unroll
        for (int c = 0; c < CG; c++) score[r][c] = 0.f;
    for (int n = 0; n < sets.n; n++) {
        const ScoreSet st = sets.s[n];
        const int A = st.A;
        float* sS = bsm;                         // [A][A]
        float* v1 = sS + A * A;                  // [NR][A]  nonzero values of P1 rows, compacted
        float* v2 = v1 + NR * A;                 // [NC][A]
        int* c1 = reinterpret_cast<int*>(v2 + NC * A);             // [NR] counts
        int* c2 = c1 + NR;                                          // [NC]
        uint8_t* i1 = reinterpret_cast<uint8_t*>(c2 + NC);          // [NR][A] their symbol indices
        uint8_t* i2 = i1 + NR * A;                                  // [NC][A]
        float* raw = reinterpret_cast<float*>(i2 + NC * A + ((4 - ((NR + NC) * A & 3)) & 3));   // [NR + NC][A] staged rows
        __syncthreads();
        for (int i = tid; i < A * A; i += 256) sS[i] = st.S[i];
        {   // all threads stage the raw profile rows (coalesced, one memory round trip) ...
            const int n1r = min(NR, L1 - y0), n2r = min(NC, L2 - x0);
            const float* s1 = st.P1 + (size_t)y0 * A;
            const float* s2 = st.P2 + (size_t)x0 * A;
            for (int i = tid; i < n1r * A; i += 256) raw[i] = s1[i];
            for (int i = tid; i < n2r * A; i += 256) raw[NR * A + i] = s2[i];
        }
        __syncthreads();
        for (int r = tid; r < NR + NC; r += 256) {   // ... then one thread compacts one row from shared memory
            const float* src = raw + r * A;
            if (r < NR) {
                int c = 0;
                if (y0 + r < L1) for (int i = 0; i < A; i++) { const float p = src[i]; if (p != 0.f) { v1[r * A + c] = p; i1[r * A + c] = (uint8_t)i; c++; } }
                c1[r] = c;
            } else {
                const int q = r - NR;
                int c = 0;
                if (x0 + q < L2) for (int j = 0; j < A; j++) { const float p = src[j]; if (p != 0.f) { v2[q * A + c] = p; i2[q * A + c] = (uint8_t)j; c++; } }
                c2[q] = c;
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < CG; c++) {
            const int q = c * 32 + tx;
            const int n2 = c2[q];
            const float* pv2 = v2 + q * A;
            const uint8_t* pi2 = i2 + q * A;
#pragma unroll
            for (int r = 0; r < RG; r++) {
                const int rr = r * 8 + ty;
                const int n1 = c1[rr];
                float acc = 0.f;
                for (int a = 0; a < n1; a++) {
                    const float p1 = v1[rr * A + a];
                    const float* srow = sS + (int)i1[rr * A + a] * A;
                    for (int b = 0; b < n2; b++)
                        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(pv2[b], srow[pi2[b]]), p1));
                }
                score[r][c] = __fadd_rn(score[r][c], acc);
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RG; r++)
#pragma unroll
        for (int c = 0; c < CG; c++) {
            const int x = x0 + c * 32 + tx, y = y0 + r * 8 + ty;
            if (x < L2 && y < L1) m[(size_t)y * m_pitch + x] = score[r][c];
        }
}

// ---- K1 for LARGE single matrices (BASELINE config 5: 20 kb x 20 kb) --------------------------------
// k_build_scores stages and compacts (8*RG + 32*CG) profile rows per 1024 cells and runs two nested
// data-dependent loops per cell whose trip counts differ from lane to lane: 5.7 ms for 4e8 cells of
// depth-8 DNA profiles, 12 % of its own instruction bound.  Here a block owns 128 COLUMNS (one per
// thread) x BC_ROWS rows:
//  * the first rounded product of a term, fl(P2[x][j] * S[i][j]) (cext.c:89, evaluation order in the
//    header above), depends on the column and the symbol pair only: every thread tabulates it once per
//    block for its own column, T[i][q][x][4] = four consecutive nonzero entries j of column x (ascending,
//    zero padded), so that a thread's quad is one conflict-free LDS.128;
//  * the block's rows of P1 are compacted once (value, table offset of symbol i; ascending i) and every
//    thread walks the SAME row at the same time: the outer trip count is uniform, the inner one is the
//    block's maximum quad count (a padded entry adds fl(0 * p1) = 0 and leaves every partial sum as it is);
//  * a cell costs nnz1 x nq x {LDS.128, 4 FMUL, 4 FADD} + one broadcast LDS.64 per nnz1, in the
//    reference's order with its two roundings per term; the row of 128 results is one coalesced store.
// Blocks whose columns need more quads than the table has room for take the direct path (no table).
#define BC_ROWS 128
__global__ void __launch_bounds__(128) k_build_scores_cols(const ScoreSet st, int L1, int L2, float* __restrict__ m, int m_pitch, int nq_cap,
                                                           int* __restrict__ ready, uint64_t nz)
{
    extern __shared__ __align__(16) float csm[];
    const int A = st.A;
    const int tid = threadIdx.x;
    const int x0 = blockIdx.x * 128, y0 = blockIdx.y * BC_ROWS;
    const int nrows = min(BC_ROWS, L1 - y0), ncols = min(128, L2 - x0);
    float* T = csm;                                              // [A][nq_cap][128][4]
    float* sS = T + (size_t)A * nq_cap * 512;                    // [A][A]
    float2* rows = reinterpret_cast<float2*>(sS + ((A * A + 1) & ~1));     // [BC_ROWS][A] (value, byte offset of T[i])
    float* raw = reinterpret_cast<float*>(rows + BC_ROWS * A);   // [128][A] staged profile rows (P2, then P1)
    float* cval = raw + 128 * A;                                 // [A][128] compacted column entries: value
    int* cidx = reinterpret_cast<int*>(cval + A * 128);          // [A][128]                           symbol j
    __shared__ int rcnt[BC_ROWS];
    __shared__ int nqmax;
    __shared__ unsigned umask;                                   // symbols the block's rows of P1 use (A <= 32)
    if (tid == 0) { nqmax = 0; umask = 0u; }
    __syncthreads();
    if (A <= 32) {
        unsigned mk = 0u;
        if (tid < nrows) {
            const float* r1 = st.P1 + (size_t)(y0 + tid) * A;
            for (int i = 0; i < A; i++) if (r1[i] != 0.f) mk |= 1u << i;
        }
        mk = __reduce_or_sync(0xffffffffu, mk);
        if ((tid & 31) == 0 && mk) atomicOr(&umask, mk);
    }
    for (int i = tid; i < A * A; i += 128) sS[i] = st.S[i];
    {
        const float* s2 = st.P2 + (size_t)x0 * A;
        for (int i = tid; i < ncols * A; i += 128) raw[i] = s2[i];
    }
    __syncthreads();
    int n2 = 0;
    if (tid < ncols) {
        for (int j = 0; j < A; j++) {
            const float p = raw[tid * A + j];
            if (p != 0.f) { cval[n2 * 128 + tid] = p; cidx[n2 * 128 + tid] = j; n2++; }
        }
        atomicMax(&nqmax, (n2 + 3) >> 2);
    }
    __syncthreads();
    const int nq = nqmax;
    // SMALL symbol sets (config 5: DNA profiles): at most four symbols in the block's rows and four entries per column --
    // the column's whole table lives in registers, see the register path below
    const unsigned um = umask;
    const bool regp = A <= 32 && nq <= 1 && __popc(um) <= 4 && nz != 0ull;
    const bool table = !regp && nq <= nq_cap;
    if (table && tid < 128) {
        for (int i = 0; i < A; i++)
            for (int q = 0; q < nq; q++) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                float* pv = reinterpret_cast<float*>(&v);
#pragma unroll
                for (int e = 0; e < 4; e++) {
                    const int b = 4 * q + e;
                    if (tid < ncols && b < n2) pv[e] = __fmul_rn(cval[b * 128 + tid], sS[i * A + cidx[b * 128 + tid]]);
                }
                *reinterpret_cast<float4*>(T + ((size_t)(i * nq + q) * 128 + tid) * 4) = v;
            }
    }
    __syncthreads();            // raw is free again
    {
        const float* s1 = st.P1 + (size_t)y0 * A;
        for (int i = tid; i < nrows * A; i += 128) raw[i] = s1[i];
    }
    __syncthreads();
    if (tid < nrows) {
        int c = 0;
        for (int i = 0; i < A; i++) {
            const float p = raw[tid * A + i];
            if (p != 0.f) { rows[tid * A + c] = make_float2(p, __int_as_float(table ? i * nq * 2048 : i * A)); c++; }
        }
        rcnt[tid] = c;
        // (0, first table row) behind the entries: rows walked together run to the longest count, and
        // fl(t * 0) = 0 leaves a partial sum as it is
        for (int k = c; k < A; k++) rows[tid * A + k] = make_float2(0.f, __int_as_float(0));
    }
    __syncthreads();
    if (regp) {
        // ---- register path.  T[u][b] = fl(P2[x][j_b] * S[i_u][j_b]) for the block's (at most four) symbols i_u and the
        // column's (at most four) entries: 16 values per thread.  The rows become dense over the four symbols (p1 = 0
        // where a row lacks one: fl(t * 0) = 0 leaves a partial sum as it is) and two rows share every instruction in the
        // packed f32x2 forms (score_rows_x2.cu: FFMA2 with the -0 addend nz = the rounded product, FADD2); the p1 pairs of
        // four rows arrive as four broadcast LDS.128.  Per two cells 16 FFMA2 + 16 FADD2, in the reference's order.
        int us[4];
        {
            unsigned r = um;
#pragma unroll
            for (int u = 0; u < 4; u++) { us[u] = r ? __ffs(r) - 1 : 0; r &= r - 1; }
        }
        const int U = __popc(um);
        float* pdn = T;                                          // [64 row pairs][4 symbols][2 rows]
#pragma unroll
        for (int u = 0; u < 4; u++)
            pdn[((tid >> 1) * 4 + u) * 2 + (tid & 1)] = (tid < nrows && u < U) ? raw[tid * A + us[u]] : 0.f;
        float Tr[4][4];
#pragma unroll
        for (int u = 0; u < 4; u++)
#pragma unroll
            for (int b = 0; b < 4; b++)
                Tr[u][b] = (tid < ncols && u < U && b < n2) ? __fmul_rn(cval[b * 128 + tid], sS[us[u] * A + cidx[b * 128 + tid]]) : 0.f;
        __syncthreads();
        if (tid < ncols) {
            float* out = m + (size_t)y0 * m_pitch + x0 + tid;
            const uint32_t pdn_s = (uint32_t)__cvta_generic_to_shared(pdn);
            for (int r = 0; r < nrows; r += 4) {
                uint64_t pa[4], pb[4];                           // (row r, r + 1) and (row r + 2, r + 3) per symbol
                const uint32_t ad = pdn_s + (uint32_t)(r >> 1) * 32u;
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(pa[0]), "=l"(pa[1]) : "r"(ad));
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(pa[2]), "=l"(pa[3]) : "r"(ad + 16u));
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(pb[0]), "=l"(pb[1]) : "r"(ad + 32u));
                asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(pb[2]), "=l"(pb[3]) : "r"(ad + 48u));
                uint64_t acca = 0ull, accb = 0ull;
#pragma unroll
                for (int u = 0; u < 4; u++)
#pragma unroll
                    for (int b = 0; b < 4; b++) {
                        uint64_t tt, pr;
                        asm("mov.b64 %0, {%1, %1};" : "=l"(tt) : "f"(Tr[u][b]));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pr) : "l"(pa[u]), "l"(tt), "l"(nz));
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(acca) : "l"(acca), "l"(pr));
                        asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(pr) : "l"(pb[u]), "l"(tt), "l"(nz));
                        asm("add.rn.f32x2 %0, %1, %2;" : "=l"(accb) : "l"(accb), "l"(pr));
                    }
                float v0, v1, v2, v3;
                asm("mov.b64 {%0, %1}, %2;" : "=f"(v0), "=f"(v1) : "l"(acca));
                asm("mov.b64 {%0, %1}, %2;" : "=f"(v2), "=f"(v3) : "l"(accb));
                out[(size_t)r * m_pitch] = __fadd_rn(0.f, v0);
                if (r + 1 < nrows) out[(size_t)(r + 1) * m_pitch] = __fadd_rn(0.f, v1);
                if (r + 2 < nrows) out[(size_t)(r + 2) * m_pitch] = __fadd_rn(0.f, v2);
                if (r + 3 < nrows) out[(size_t)(r + 3) * m_pitch] = __fadd_rn(0.f, v3);
            }
        }
    } else
    if (tid < ncols) {
    float* out = m + (size_t)y0 * m_pitch + x0 + tid;
    if (table) {
        // explicit shared-space addresses: through generic pointers every load paid an address-space
        // conversion (S2UR CgaCtaId ...) and the loop ran 42 instructions per entry instead of 13
        const uint32_t Ts = (uint32_t)__cvta_generic_to_shared(T) + (uint32_t)tid * 16u;
        const uint32_t rows_s = (uint32_t)__cvta_generic_to_shared(rows);
        auto lds2 = [](uint32_t addr, float& x, uint32_t& y) {
            asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(x), "=r"(y) : "r"(addr));
        };
        auto lds4 = [](uint32_t addr) -> float4 {
            float4 v;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
            return v;
        };
        if (nq == 1) {
            // four rows per pass: the cell's dependent FADD chain (4 per entry) is what a thread waits on; four
            // independent chains per thread hide it at 12 warps per SM (two: 2.4 ms at 20k x 20k).  The rows run to
            // the longest of their counts over zero padded entries.
            int r = 0;
            for (; r + 3 < nrows; r += 4) {
                const int nmax = max(max(rcnt[r], rcnt[r + 1]), max(rcnt[r + 2], rcnt[r + 3]));
                uint32_t ra = rows_s + (uint32_t)(r * A) * 8u;
                const uint32_t rstep = (uint32_t)A * 8u;
                float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
                for (int a = 0; a < nmax; a++, ra += 8u) {
                    float p0, p1, p2, p3; uint32_t o0, o1, o2, o3;
                    lds2(ra, p0, o0);
                    lds2(ra + rstep, p1, o1);
                    lds2(ra + 2u * rstep, p2, o2);
                    lds2(ra + 3u * rstep, p3, o3);
                    const float4 t0 = lds4(Ts + o0), t1 = lds4(Ts + o1), t2 = lds4(Ts + o2), t3 = lds4(Ts + o3);
                    acc0 = __fadd_rn(acc0, __fmul_rn(t0.x, p0)); acc1 = __fadd_rn(acc1, __fmul_rn(t1.x, p1));
                    acc2 = __fadd_rn(acc2, __fmul_rn(t2.x, p2)); acc3 = __fadd_rn(acc3, __fmul_rn(t3.x, p3));
                    acc0 = __fadd_rn(acc0, __fmul_rn(t0.y, p0)); acc1 = __fadd_rn(acc1, __fmul_rn(t1.y, p1));
                    acc2 = __fadd_rn(acc2, __fmul_rn(t2.y, p2)); acc3 = __fadd_rn(acc3, __fmul_rn(t3.y, p3));
                    acc0 = __fadd_rn(acc0, __fmul_rn(t0.z, p0)); acc1 = __fadd_rn(acc1, __fmul_rn(t1.z, p1));
                    acc2 = __fadd_rn(acc2, __fmul_rn(t2.z, p2)); acc3 = __fadd_rn(acc3, __fmul_rn(t3.z, p3));
                    acc0 = __fadd_rn(acc0, __fmul_rn(t0.w, p0)); acc1 = __fadd_rn(acc1, __fmul_rn(t1.w, p1));
                    acc2 = __fadd_rn(acc2, __fmul_rn(t2.w, p2)); acc3 = __fadd_rn(acc3, __fmul_rn(t3.w, p3));
                }
                out[(size_t)r * m_pitch] = __fadd_rn(0.f, acc0);
                out[(size_t)(r + 1) * m_pitch] = __fadd_rn(0.f, acc1);
                out[(size_t)(r + 2) * m_pitch] = __fadd_rn(0.f, acc2);
                out[(size_t)(r + 3) * m_pitch] = __fadd_rn(0.f, acc3);
            }
            for (; r < nrows; r++) {
                const int n1 = rcnt[r];
                uint32_t ra = rows_s + (uint32_t)(r * A) * 8u;
                float acc = 0.f;
                for (int a = 0; a < n1; a++, ra += 8u) {
                    float p1; uint32_t off;
                    lds2(ra, p1, off);
                    const float4 t4 = lds4(Ts + off);
                    acc = __fadd_rn(acc, __fmul_rn(t4.x, p1));
                    acc = __fadd_rn(acc, __fmul_rn(t4.y, p1));
                    acc = __fadd_rn(acc, __fmul_rn(t4.z, p1));
                    acc = __fadd_rn(acc, __fmul_rn(t4.w, p1));
                }
                out[(size_t)r * m_pitch] = __fadd_rn(0.f, acc);
            }
        } else {
            for (int r = 0; r < nrows; r++) {
                const int n1 = rcnt[r];
                uint32_t ra = rows_s + (uint32_t)(r * A) * 8u;
                float acc = 0.f;
                for (int a = 0; a < n1; a++, ra += 8u) {
                    float p1; uint32_t off;
                    lds2(ra, p1, off);
                    for (int q = 0; q < nq; q++) {
                        const float4 t4 = lds4(Ts + off + (uint32_t)q * 2048u);
                        acc = __fadd_rn(acc, __fmul_rn(t4.x, p1));
                        acc = __fadd_rn(acc, __fmul_rn(t4.y, p1));
                        acc = __fadd_rn(acc, __fmul_rn(t4.z, p1));
                        acc = __fadd_rn(acc, __fmul_rn(t4.w, p1));
                    }
                }
                out[(size_t)r * m_pitch] = __fadd_rn(0.f, acc);
            }
        }
    } else {
        for (int r = 0; r < nrows; r++) {
            const int n1 = rcnt[r];
            const float2* rr = rows + r * A;
            float acc = 0.f;
            for (int a = 0; a < n1; a++) {
                const float2 e = rr[a];
                const float* srow = sS + __float_as_int(e.y);
                for (int b = 0; b < n2; b++)
                    acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(cval[b * 128 + tid], srow[cidx[b * 128 + tid]]), e.x));
            }
            out[(size_t)r * m_pitch] = __fadd_rn(0.f, acc);
        }
    }
    }
    if (ready != nullptr) {
        // K1 beside the fill (pgpu_align_profile_long): this 128 x 128 block of m is complete.  Every thread's stores,
        // then a device-scope release by one thread; k_wave4 acquires the flag before it requests rows of the block.
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            asm volatile("st.release.gpu.global.b32 [%0], %1;" ::"l"(ready + (size_t)blockIdx.y * gridDim.x + blockIdx.x), "r"(1) : "memory");
        }
    }
}

// Batched form for the matrix-fed streaming kernel: one matrix row per stream position of a
// wave (rows in stream order, `width` = 32*K floats each).  Work arrives as BLOCKS of <= 32
// consecutive matrix rows that belong to one streamed sequence (so their profile rows are
// consecutive too): PgRowBlock = {first matrix row, profile row of that matrix row, rows,
// resident sequence, first row is the region's dummy row}.  Same evaluation order as
// k_build_scores; `transposed` says the resident is sequence one.
//
// Both operands are compacted to their NONZERO entries first (ascending symbol, the order of
// build_nonzero_matrix, component/align.py:449-458): the 32 streamed rows once per block by
// ballot, the thread's own resident row into a [entry][thread] table.  A cell then costs
// nnz1 x nnz2 terms of {LDS.64 (value, S offset), LDS S, FMUL, FMUL, FADD}; S is padded to an odd
// row stride so that threads on different symbols hit different banks.
__global__ void __launch_bounds__(128) k_build_rows(const float* __restrict__ prof, const int64_t* __restrict__ rowoff,
                                                    int A, const float* __restrict__ S,
                                                    const PgRowBlock* __restrict__ blocks, int width,
                                                    int transposed, float padv, float* __restrict__ mwave)
{
    extern __shared__ __align__(16) float sh[];
    const int AS = (A <= 32) ? 33 : (A | 1);
    float* sS = sh;                                             // [A][AS]
    float2* sstr = reinterpret_cast<float2*>(sS + ((A * AS + 1) & ~1));   // [32][A] (value, S offset) of streamed rows
    float* rval = reinterpret_cast<float*>(sstr + 32 * A);      // [A][128] resident values
    int* roff = reinterpret_cast<int*>(rval + A * 128);         // [A][128] resident S offsets
    __shared__ int scnt[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xblocks = (width + 127) / 128;
    const PgRowBlock blk = blocks[blockIdx.x / xblocks];
    const int x = (int)(blockIdx.x % xblocks) * 128 + tid;
    for (int i = tid; i < A * A; i += 128) sS[(i / A) * AS + (i % A)] = S[i];
    // streamed rows: the streamed sequence is sequence one when the resident is sequence two
    // (row of S = its symbol, offset i*AS) and sequence two otherwise (column, offset j)
    for (int r = warp; r < blk.rows; r += 4) {
        int c = 0;
        if (!(blk.dummy && r == 0)) {
            const float* src = prof + (size_t)(blk.src0 + r) * A;
            for (int base = 0; base < A; base += 32) {
                const int i = base + lane;
                const float p = i < A ? src[i] : 0.f;
                const unsigned m = __ballot_sync(0xffffffffu, p != 0.f);
                if (p != 0.f) sstr[r * A + c + __popc(m & ((1u << lane) - 1u))] = make_float2(p, __int_as_float(transposed ? i : i * AS));
                c += __popc(m);
            }
        }
        if (lane == 0) scnt[r] = c;
    }
    const int64_t q0 = rowoff[blk.res];
    const int Lr = (int)(rowoff[blk.res + 1] - q0);
    int nres = 0;
    if (x < width && x < Lr) {
        const float* rr_ = prof + (size_t)(q0 + x) * A;   // resident profile row
        for (int i = 0; i < A; i++) {
            const float p = rr_[i];
            if (p != 0.f) { rval[nres * 128 + tid] = p; roff[nres * 128 + tid] = transposed ? i * AS : i; nres++; }
        }
    }
    __syncthreads();
    if (x >= width) return;
    for (int r = 0; r < blk.rows; r++) {
        float v = padv;
        if (blk.dummy && r == 0) v = 0.f;
        else if (x < Lr) {
            const float2* sr = sstr + r * A;
            const int ns = scnt[r];
            float acc = 0.f;
            if (transposed) {       // resident = sequence one: outer loop over MY entries, inner over the streamed row's
                for (int a = 0; a < nres; a++) {
                    const float p1 = rval[a * 128 + tid];
                    const float* srow = sS + roff[a * 128 + tid];
#pragma unroll 4
                    for (int b = 0; b < ns; b++) {
                        const float2 e = sr[b];
                        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(e.x, srow[__float_as_int(e.y)]), p1));
                    }
                }
            } else {                // resident = sequence two: outer loop over the streamed row, inner over MY entries
                for (int a = 0; a < ns; a++) {
                    const float2 e = sr[a];
                    const float p1 = e.x;
                    const float* srow = sS + __float_as_int(e.y);
#pragma unroll 4
                    for (int b = 0; b < nres; b++)
                        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(rval[b * 128 + tid], srow[roff[b * 128 + tid]]), p1));
                }
            }
            v = __fadd_rn(0.f, acc);
        }
        mwave[(size_t)(blk.row0 + r) * width + x] = v;
    }
}

// k_build_rows for a resident that is SEQUENCE ONE (transposed launches), the orientation the
// engine prefers for exact profile batches.  The first rounded product of a term,
// fl(p2[x][j] * S[i][j]), depends only on the streamed row x and the symbol pair (i, j): the block
// tabulates it once per streamed row, T[r][i][b] for the row's nonzero entries b (zero padded to a
// multiple of four, rows of T at an odd multiple of 16 bytes so that threads on different symbols
// i load from different banks).  A cell is then, per nonzero entry (i, p1) of the thread's own
// resident row, ceil(nnz2 / 4) x {LDS.128, 4 FMUL, 4 FADD} in the reference's order -- 2.25
// instructions per term instead of ~7, with the same two roundings per term.  A padded entry adds
// fl(0 * p1) = 0, which leaves every partial sum unchanged.
// one cell: for every nonzero entry (value p1, byte offset of row i of T) of the thread's resident row,
// N4 quads of tabulated first products, in order (N4 is uniform for the streamed row -> no tail loop)
template <int N4>
__device__ __forceinline__ float rows_t_cell(uint32_t tr_s, uint32_t ent_s, int nres)
{
    float acc = 0.f;
    for (int a = 0; a < nres; a++) {
        float p1;
        uint32_t off;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(p1), "=r"(off) : "r"(ent_s + (uint32_t)a * 1024u));
        const uint32_t row = tr_s + off;
        float4 t[N4];
#pragma unroll
        for (int b = 0; b < N4; b++)
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(t[b].x), "=f"(t[b].y), "=f"(t[b].z), "=f"(t[b].w) : "r"(row + (uint32_t)b * 16u));
#pragma unroll
        for (int b = 0; b < N4; b++) {
            acc = __fadd_rn(acc, __fmul_rn(t[b].x, p1));
            acc = __fadd_rn(acc, __fmul_rn(t[b].y, p1));
            acc = __fadd_rn(acc, __fmul_rn(t[b].z, p1));
            acc = __fadd_rn(acc, __fmul_rn(t[b].w, p1));
        }
    }
    return acc;
}

// two streamed rows at once for one resident row: the entry loads and the loop are shared and the two cells'
// FADD chains are independent (a cell is one dependent chain of nnz1 x 4 N4 adds: a lone chain per thread leaves the
// FP32 pipe a third busy).  N4 covers the longer of the two rows; the shorter one's table is zero padded.
template <int N4>
__device__ __forceinline__ void rows_t_cell2(uint32_t tra_s, uint32_t trb_s, uint32_t ent_s, int nres, float& ra, float& rb)
{
    float acca = 0.f, accb = 0.f;
    for (int a = 0; a < nres; a++) {
        float p1;
        uint32_t off;
        asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(p1), "=r"(off) : "r"(ent_s + (uint32_t)a * 1024u));
        float4 ta[N4], tb[N4];
#pragma unroll
        for (int b = 0; b < N4; b++) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(ta[b].x), "=f"(ta[b].y), "=f"(ta[b].z), "=f"(ta[b].w) : "r"(tra_s + off + (uint32_t)b * 16u));
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(tb[b].x), "=f"(tb[b].y), "=f"(tb[b].z), "=f"(tb[b].w) : "r"(trb_s + off + (uint32_t)b * 16u));
        }
#pragma unroll
        for (int b = 0; b < N4; b++) {
            acca = __fadd_rn(acca, __fmul_rn(ta[b].x, p1)); accb = __fadd_rn(accb, __fmul_rn(tb[b].x, p1));
            acca = __fadd_rn(acca, __fmul_rn(ta[b].y, p1)); accb = __fadd_rn(accb, __fmul_rn(tb[b].y, p1));
            acca = __fadd_rn(acca, __fmul_rn(ta[b].z, p1)); accb = __fadd_rn(accb, __fmul_rn(tb[b].z, p1));
            acca = __fadd_rn(acca, __fmul_rn(ta[b].w, p1)); accb = __fadd_rn(accb, __fmul_rn(tb[b].w, p1));
        }
    }
    ra = acca;
    rb = accb;
}

template <int RB>
__global__ void __launch_bounds__(128) k_build_rows_t(const float* __restrict__ prof, const int64_t* __restrict__ rowoff,
                                                      int A, int ST, const float* __restrict__ S,
                                                      const PgRowBlock* __restrict__ blocks, int width,
                                                      float padv, float* __restrict__ mwave)
{
    extern __shared__ __align__(16) float sh[];
    float* T = sh;                                              // [RB][A][ST]
    float* sS = T + RB * A * ST;                                // [A][A]
    float2* rent = reinterpret_cast<float2*>(sS + ((A * A + 3) & ~3));   // [A][128] my resident row: (value, byte offset i * ST * 4)
    float* sval = reinterpret_cast<float*>(rent + A * 128);     // [RB][A] streamed rows: values
    int* sj = reinterpret_cast<int*>(sval + RB * A);            // [RB][A] ... and symbols j
    __shared__ int scnt[RB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int xblocks = (width + 127) / 128;
    const PgRowBlock blk = blocks[blockIdx.x / xblocks];
    const int x = (int)(blockIdx.x % xblocks) * 128 + tid;
    for (int i = tid; i < A * A; i += 128) sS[i] = S[i];
    const int64_t q0 = rowoff[blk.res];
    const int Lr = (int)(rowoff[blk.res + 1] - q0);
    const bool mine = x < width && x < Lr;
    int nres = 0;
    if (mine) {
        const float* rr_ = prof + (size_t)(q0 + x) * A;
        for (int i = 0; i < A; i++) {
            const float p = rr_[i];
            if (p != 0.f) { rent[nres * 128 + tid] = make_float2(p, __int_as_float(i * ST * 4)); nres++; }
        }
    }
    const uint32_t ent_s = (uint32_t)__cvta_generic_to_shared(rent + tid);
    const uint32_t T_s = (uint32_t)__cvta_generic_to_shared(T);
    for (int r0 = 0; r0 < blk.rows; r0 += RB) {
        const int nr = min(RB, blk.rows - r0);
        __syncthreads();                                        // the previous pass is done with T
        for (int r = warp; r < nr; r += 4) {                    // compact the streamed rows (ascending symbol)
            int c = 0;
            if (!(blk.dummy && r0 + r == 0)) {
                const float* src = prof + (size_t)(blk.src0 + r0 + r) * A;
                for (int base = 0; base < A; base += 32) {
                    const int j = base + lane;
                    const float p = j < A ? src[j] : 0.f;
                    const unsigned m = __ballot_sync(0xffffffffu, p != 0.f);
                    if (p != 0.f) { const int at = r * A + c + __popc(m & ((1u << lane) - 1u)); sval[at] = p; sj[at] = j; }
                    c += __popc(m);
                }
            }
            if (lane == 0) scnt[r] = c;
        }
        __syncthreads();
        // T[r][i][b] = fl(p2_b * S[i][j_b]), zero padded: a warp per streamed row, a lane per entry b
        // (no index divisions: they cost more than the table's multiplies)
        for (int r = warp; r < nr; r += 4) {
            const int ns = scnt[r];
            for (int b = lane; b < ST; b += 32) {
                const bool on = b < ns;
                const float p2 = on ? sval[r * A + b] : 0.f;
                const float* scol = sS + (on ? sj[r * A + b] : 0);
                float* dst = T + (size_t)r * A * ST + b;
                for (int i = 0; i < A; i++) dst[i * ST] = on ? __fmul_rn(p2, scol[i * A]) : 0.f;
            }
        }
        __syncthreads();
        if (x < width) {
            int r = 0;
            if (mine) {
                // pairs of streamed rows with the common quad counts (the dummy row and odd tails go one by one below)
                const int rfirst = (blk.dummy && r0 == 0) ? 1 : 0;
                if (rfirst) mwave[(size_t)(blk.row0 + r0) * width + x] = 0.f;
                for (r = rfirst; r + 1 < nr; r += 2) {
                    const int n4 = (max(scnt[r], scnt[r + 1]) + 3) >> 2;
                    if (n4 < 1 || n4 > 7 || 4 * n4 > ST) break;
                    const uint32_t tra = T_s + (uint32_t)(r * A * ST) * 4u, trb = tra + (uint32_t)(A * ST) * 4u;
                    float va, vb;
                    switch (n4) {
                        case 1: rows_t_cell2<1>(tra, trb, ent_s, nres, va, vb); break;
                        case 2: rows_t_cell2<2>(tra, trb, ent_s, nres, va, vb); break;
                        case 3: rows_t_cell2<3>(tra, trb, ent_s, nres, va, vb); break;
                        case 4: rows_t_cell2<4>(tra, trb, ent_s, nres, va, vb); break;
                        case 5: rows_t_cell2<5>(tra, trb, ent_s, nres, va, vb); break;
                        case 6: rows_t_cell2<6>(tra, trb, ent_s, nres, va, vb); break;
                        default: rows_t_cell2<7>(tra, trb, ent_s, nres, va, vb); break;
                    }
                    mwave[(size_t)(blk.row0 + r0 + r) * width + x] = __fadd_rn(0.f, va);
                    mwave[(size_t)(blk.row0 + r0 + r + 1) * width + x] = __fadd_rn(0.f, vb);
                }
            }
            for (; r < nr; r++) {
                float v = padv;
                if (blk.dummy && r0 + r == 0) v = 0.f;
                else if (mine) {
                    const uint32_t tr_s = T_s + (uint32_t)(r * A * ST) * 4u;
                    const int n4 = (scnt[r] + 3) >> 2;
                    float acc;
                    switch (n4) {       // uniform per streamed row
                        case 0: acc = 0.f; break;
                        case 1: acc = rows_t_cell<1>(tr_s, ent_s, nres); break;
                        case 2: acc = rows_t_cell<2>(tr_s, ent_s, nres); break;
                        case 3: acc = rows_t_cell<3>(tr_s, ent_s, nres); break;
                        case 4: acc = rows_t_cell<4>(tr_s, ent_s, nres); break;
                        case 5: acc = rows_t_cell<5>(tr_s, ent_s, nres); break;
                        case 6: acc = rows_t_cell<6>(tr_s, ent_s, nres); break;
                        case 7: acc = rows_t_cell<7>(tr_s, ent_s, nres); break;
                        default: {      // alphabets beyond 28 symbols per row
                            acc = 0.f;
                            for (int a = 0; a < nres; a++) {
                                const float2 e = rent[a * 128 + tid];
                                const float4* row = reinterpret_cast<const float4*>(reinterpret_cast<const char*>(T + r * A * ST) + __float_as_int(e.y));
                                for (int b = 0; b < n4; b++) {
                                    const float4 t = row[b];
                                    acc = __fadd_rn(acc, __fmul_rn(t.x, e.x));
                                    acc = __fadd_rn(acc, __fmul_rn(t.y, e.x));
                                    acc = __fadd_rn(acc, __fmul_rn(t.z, e.x));
                                    acc = __fadd_rn(acc, __fmul_rn(t.w, e.x));
                                }
                            }
                        }
                    }
                    v = __fadd_rn(0.f, acc);
                }
                mwave[(size_t)(blk.row0 + r0 + r) * width + x] = v;
            }
        }
    }
}

// Tolerance-mode variant of k_build_rows for score-only profile batches (guide tree on deep
// preprofiles): the contraction P1 . S . P2^T is factored through W = P_resident . S^T (or . S),
// precomputed per sequence, so a cell costs A fused multiply-adds instead of nnz1 x nnz2
// mul-mul-add triples.  Not the reference's evaluation order: scores agree to ~1e-6 relative
// (the stated tolerance is 1e-5), so this path is opt-in and never used for traced alignments.
// Block = 32 matrix rows x 128 columns of ONE region (all its rows share a resident): the thread
// keeps its column's W row in registers and sweeps the 32 staged profile rows.
template <int AP>
__global__ void __launch_bounds__(128) k_build_rows_fast(const float* __restrict__ prof, const float* __restrict__ wres,
                                                         const int64_t* __restrict__ rowoff, int A,
                                                         const PgRowBlock* __restrict__ blocks, int width, float padv,
                                                         float* __restrict__ mwave)
{
    __shared__ float srow[32][AP];
    const int xblocks = (width + 127) / 128;
    const PgRowBlock blk = blocks[blockIdx.x / xblocks];
    const int x = (int)(blockIdx.x % xblocks) * 128 + threadIdx.x;
    for (int i = threadIdx.x; i < 32 * AP; i += 128) {
        const int rr = i / AP, c = i % AP;
        float v = 0.f;
        if (rr < blk.rows && c < A && !(blk.dummy && rr == 0)) v = prof[(size_t)(blk.src0 + rr) * A + c];
        srow[rr][c] = v;
    }
    __syncthreads();
    if (x >= width) return;
    const int64_t q0 = rowoff[blk.res];
    const int Lr = (int)(rowoff[blk.res + 1] - q0);
    float w[AP];
#pragma unroll
    for (int c = 0; c < AP; c++) w[c] = (x < Lr && c < A) ? wres[(size_t)(q0 + x) * A + c] : 0.f;
    for (int rr = 0; rr < blk.rows; rr++) {
        float acc = 0.f;
#pragma unroll
        for (int c = 0; c < AP; c++) acc = fmaf(srow[rr][c], w[c], acc);
        const bool dummy = blk.dummy && rr == 0;
        mwave[(size_t)(blk.row0 + rr) * width + x] = dummy ? 0.f : (x < Lr ? acc : padv);
    }
}

// W[row][i] = sum_j S[i][j] * P[row][j]  (transposed = 0)   or   sum_j P[row][j] * S[j][i]  (transposed = 1)
__global__ void k_profile_times_matrix(const float* __restrict__ prof, const float* __restrict__ S, int A, int64_t n_rows,
                                       int transposed, float* __restrict__ out)
{
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_rows * A) return;
    const int64_t r = idx / A;
    const int i = (int)(idx % A);
    float acc = 0.f;
    for (int j = 0; j < A; j++) acc = fmaf(prof[r * A + j], transposed ? S[j * A + i] : S[i * A + j], acc);
    out[idx] = acc;
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
// the lean kernels need the padded matrix pitch the engine allocates; anything else (and the debug dumps) runs the
// general-layout kernels
static bool general_lean(const GenArgs& a)
{
    const int lean_strips = (a.L2 + 127) / 128;
    return !a.o_full && a.flagw && a.m_pitch % 4 == 0 && a.m_pitch >= lean_strips * 128 &&
           ((reinterpret_cast<uintptr_t>(a.m) & 15) == 0) && getenv("PGPU_NO_LEAN") == nullptr;
}
bool pg_general_uses_wave4(const GenArgs& a)
{
    return general_lean(a) && a.mode != PG_LOCAL && a.z == nullptr && a.var_gaps == 0 && a.L1 > 0 && a.L2 > 0 &&
           getenv("PGPU_NO_WAVE4") == nullptr;
}
int pg_launch_general(GenArgs a, int kg, cudaStream_t st)
{
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // the lean kernel needs the padded matrix pitch the engine allocates; anything else (and the
    // debug dumps) runs the general-layout kernels
    const int lean_strips = (a.L2 + 127) / 128;
    const bool lean = general_lean(a);
    a.n_strips = lean ? lean_strips : (a.L2 + 32 * kg - 1) / (32 * kg);
    if (a.n_strips < 1) a.n_strips = 1;
    if (lean) {
        a.flag_fmt = 2;
        // tags of the strip-to-strip edge records must start invalid (row numbers start at 1)
        PG_CUDA_OK(cudaMemsetAsync(a.edge + (size_t)(a.L1 + 1) * 8, 0, sizeof(float) * 8 * (size_t)a.n_strips * (a.L1 + 1), st));
    }
    k_gen_init<<<32, 256, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    const bool local = a.mode == PG_LOCAL;
    const bool mask = a.z != nullptr;
    if (lean) {
        // all strips co-resident, spread over the SMs: a lone warp per scheduler runs its
        // dependent step loop fastest
        int wpc = (a.n_strips + sms - 1) / sms;
        if (wpc > 8) wpc = 8;
        int ctas = (a.n_strips + wpc - 1) / wpc;
        if (ctas > sms) ctas = sms;
#define PG_WAVE(LO, MA, VG)                                                                          \
    do {                                                                                             \
        auto kern = k_wave<LO, MA, VG>;                                                              \
        const size_t sm = (size_t)wpc * WV_R * 32 * (VG ? 16 : 12) * sizeof(float);                   \
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        void* kargs[] = {(void*)&a};                                                                 \
        /* strips spin on their left neighbours: every CTA must be resident, which a cooperative launch */ \
        /* guarantees (it fails instead of deadlocking when the grid cannot be co-scheduled) */          \
        PG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas), dim3(wpc * 32), kargs, sm, st)); \
    } while (0)
        const bool vg = a.var_gaps != 0;
        a.flag_skew = 1;
        a.flag_rows = a.L1 + 31;
        if (pg_general_uses_wave4(a)) {
            a.flag_skew = 4;
            a.flag_rows = (((a.L1 + 3) >> 2) + 31) * 4;
            auto kern = k_wave4;
            const size_t sm = (size_t)wpc * W4_R * W4_SLOTF * sizeof(float);
            PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
            void* kargs[] = {(void*)&a};
            PG_CUDA_OK(cudaLaunchCooperativeKernel((const void*)kern, dim3(ctas), dim3(wpc * 32), kargs, sm, st));
        } else
        if (local) { if (mask) { if (vg) PG_WAVE(true, true, true); else PG_WAVE(true, true, false); }
                     else { if (vg) PG_WAVE(true, false, true); else PG_WAVE(true, false, false); } }
        else { if (mask) { if (vg) PG_WAVE(false, true, true); else PG_WAVE(false, true, false); }
               else { if (vg) PG_WAVE(false, false, true); else PG_WAVE(false, false, false); } }
#undef PG_WAVE
        PG_CUDA_OK(cudaGetLastError());
    } else if (a.L1 > 0 && a.L2 > 0) {
        const int wpc = 8;
        int ctas = (a.n_strips + wpc - 1) / wpc;
        if (ctas > sms) ctas = sms;    // all warps co-resident: strips wait on lower strips only
        if (a.o_full) {               // debug / B3 shim: the reference's complete flag bytes and o
            a.flag_fmt = 0;
            if (kg == 2) {
                if (local) k_gen_fill<2, true><<<ctas, wpc * 32, 0, st>>>(a);
                else k_gen_fill<2, false><<<ctas, wpc * 32, 0, st>>>(a);
            } else {
                if (local) k_gen_fill<8, true><<<ctas, wpc * 32, 0, st>>>(a);
                else k_gen_fill<8, false><<<ctas, wpc * 32, 0, st>>>(a);
            }
        } else {
            a.flag_fmt = 1;
#define PG_FAST1(KGV, LO, MA)                                                                      \
    do {                                                                                           \
        auto kern = k_gen_fill_fast<KGV, LO, MA>;                                                  \
        const int w2 = a.n_strips < wpc ? a.n_strips : wpc;                                        \
        const int c2 = (a.n_strips + w2 - 1) / w2 > sms ? sms : (a.n_strips + w2 - 1) / w2;        \
        const size_t sm = (size_t)w2 * GEN_R * 32 * ((((KGV + 2) + 3) & ~3) + 4) * sizeof(float);  \
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm)); \
        kern<<<c2, w2 * 32, sm, st>>>(a);                                                          \
    } while (0)
#define PG_FAST(KGV)                                                                               \
    do {                                                                                           \
        if (local) { if (mask) PG_FAST1(KGV, true, true); else PG_FAST1(KGV, true, false); }       \
        else { if (mask) PG_FAST1(KGV, false, true); else PG_FAST1(KGV, false, false); }           \
    } while (0)
            if (kg == 1) PG_FAST(1);
            else if (kg == 2) PG_FAST(2);
            else if (kg == 4) PG_FAST(4);
            else if (kg == 8) PG_FAST(8);
            else if (kg == 16) PG_FAST(16);
            else { pg_set_error("general kernel: unsupported strip width kg=%d", kg); return 1; }
#undef PG_FAST
#undef PG_FAST1
        }
        PG_CUDA_OK(cudaGetLastError());
    }
    k_gen_finalize<<<1, 256, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    if (a.path_buf) {
        if (lean) k_wave_traceback<<<1, 32, 0, st>>>(a);
        else k_gen_traceback<<<1, 32, 0, st>>>(a);
        PG_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

bool pg_build_scores_flags_blocks(int n_sets, int L1, int L2)
{
    return n_sets == 1 && (size_t)L1 * L2 >= (size_t)1 << 21 && getenv("PGPU_K1_COLS_OFF") == nullptr;
}
int pg_launch_build_scores(const ScoreSets& sets, int L1, int L2, float* m, int m_pitch, cudaStream_t st, int* ready)
{
    if (L1 <= 0 || L2 <= 0) return 0;
    int amax = 1;
    for (int i = 0; i < sets.n; i++) {
        if (sets.s[i].A > BS_MAXA) { pg_set_error("alphabet size %d above %d", sets.s[i].A, BS_MAXA); return 1; }
        if (sets.s[i].A > amax) amax = sets.s[i].A;
    }
    if (sets.n == 1 && (size_t)L1 * L2 >= (size_t)1 << 21 && getenv("PGPU_K1_COLS_OFF") == nullptr) {
        // large single matrix: column-per-thread kernel with tabulated first products
        const int A = sets.s[0].A;
        int nq_cap = (32 * 1024) / (A * 2048);     // 30 KB of tables at A = 15: three blocks per SM
        if (nq_cap < 1) nq_cap = 1;
        if (nq_cap > (A + 3) / 4) nq_cap = (A + 3) / 4;
        const size_t smc = sizeof(float) * ((size_t)A * nq_cap * 512 + ((A * A + 1) & ~1) + 2 * (size_t)BC_ROWS * A + 128 * A + 2 * A * 128) + 64;
        auto kern = k_build_scores_cols;
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smc));
        dim3 g((L2 + 127) / 128, (L1 + BC_ROWS - 1) / BC_ROWS);
        // (beside the fill, `ready`: fewer resident blocks per SM only starve the wavefront -- measured 11.5 ms with the
        // usual occupancy, 13.5 ms with two blocks per SM, 15.3 ms with one)
        kern<<<g, 128, smc, st>>>(sets.s[0], L1, L2, m, m_pitch, nq_cap, ready,
                                  getenv("PGPU_K1_REG_OFF") ? 0ull : 0x8000000080000000ull);
        PG_CUDA_OK(cudaGetLastError());
        return 0;
    }
    if (ready != nullptr) { pg_set_error("build_scores: ready flags need the column kernel"); return 1; }
    const char* ev = getenv("PGPU_K1_BIG");
    const bool big = ev ? atoi(ev) != 0 : (size_t)L1 * L2 >= (size_t)1 << 21;   // measured: 7.7 -> 5.7 ms at 20k x 20k
    const int nr = big ? 16 : 8, nc = big ? 64 : 32;
    const size_t sm = sizeof(float) * (amax * amax + 2 * (nr + nc) * amax) + sizeof(int) * (nr + nc) + (size_t)(nr + nc) * amax + 32;
    dim3 b(32, 8), g((L2 + nc - 1) / nc, (L1 + nr - 1) / nr);
    if (big) {
        auto kern = k_build_scores<2, 2>;
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
        kern<<<g, b, sm, st>>>(sets, L1, L2, m, m_pitch);
    } else {
        k_build_scores<1, 1><<<g, b, sm, st>>>(sets, L1, L2, m, m_pitch);
    }
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
int pg_launch_build_rows(const float* prof, const int64_t* rowoff, int A, const float* S, const PgRowBlock* blocks,
                         int n_blocks, int width, int transposed, float padv, int dense_syms, float* mwave, cudaStream_t st)
{
    if (n_blocks <= 0) return 0;
    // dense profiles (the caller counted the symbols in use): packed f32x2 rows, two resident columns per thread
    if (transposed && dense_syms > 0 && A <= 32 && getenv("PGPU_NO_ROWS_X2") == nullptr)
        return pg_launch_build_rows_x2(prof, rowoff, A, S, blocks, n_blocks, width, padv, dense_syms, mwave, st);
    const int64_t nb = (int64_t)n_blocks * ((width + 127) / 128);
    if (nb > 0x7fffffffll) { pg_set_error("wave too large for one launch (%lld blocks)", (long long)nb); return 1; }
    if (transposed && getenv("PGPU_NO_ROWS_T") == nullptr) {
        // resident = sequence one: tabulated first products (k_build_rows_t) when the tables fit
        int st4 = (A + 3) / 4;
        if (!(st4 & 1)) st4++;
        const int ST = 4 * st4;
        auto need = [&](int rb) {
            return sizeof(float) * ((size_t)rb * A * ST + ((A * A + 3) & ~3) + 2 * (size_t)128 * A + 2 * (size_t)rb * A) + 16;
        };
#define PG_ROWS_T(RBV)                                                                                   \
    do {                                                                                                 \
        auto kern = k_build_rows_t<RBV>;                                                                 \
        PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)need(RBV))); \
        kern<<<(unsigned)nb, 128, need(RBV), st>>>(prof, rowoff, A, ST, S, blocks, width, padv, mwave);  \
        PG_CUDA_OK(cudaGetLastError());                                                                  \
        return 0;                                                                                        \
    } while (0)
        if (need(8) <= 72 * 1024) PG_ROWS_T(8);
        if (need(4) <= 72 * 1024) PG_ROWS_T(4);
        if (need(2) <= 100 * 1024) PG_ROWS_T(2);
#undef PG_ROWS_T
    }
    const int AS = (A <= 32) ? 33 : (A | 1);
    const size_t sm = sizeof(float) * (size_t)(((A * AS + 1) & ~1) + 2 * 32 * A + 2 * 128 * A) + 16;
    PG_CUDA_OK(cudaFuncSetAttribute(k_build_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_build_rows<<<(unsigned)nb, 128, sm, st>>>(prof, rowoff, A, S, blocks, width, transposed, padv, mwave);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_build_rows_fast(const float* prof, const float* wres, const int64_t* rowoff, int A, const PgRowBlock* blocks,
                              int n_blocks, int width, float padv, float* mwave, cudaStream_t st)
{
    if (n_blocks <= 0) return 0;
    const int64_t nb = (int64_t)n_blocks * ((width + 127) / 128);
    if (nb > 0x7fffffffll) { pg_set_error("wave too large for one launch (%lld blocks)", (long long)nb); return 1; }
    if (A <= 16) k_build_rows_fast<16><<<(unsigned)nb, 128, 0, st>>>(prof, wres, rowoff, A, blocks, width, padv, mwave);
    else if (A <= 32) k_build_rows_fast<32><<<(unsigned)nb, 128, 0, st>>>(prof, wres, rowoff, A, blocks, width, padv, mwave);
    else { pg_set_error("fast profile rows: alphabet size %d above 32", A); return 1; }
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_profile_times_matrix(const float* prof, const float* S, int A, int64_t n_rows, int transposed, float* out,
                                   cudaStream_t st)
{
    if (n_rows <= 0) return 0;
    const int64_t n = n_rows * A;
    k_profile_times_matrix<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(prof, S, A, n_rows, transposed, out);
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
