// score_rows_x2.cu -- exact-order profile score rows for DENSE profiles on the packed f32x2 pipe of sm_100a.
//
// What it computes (reference: score_match_prof_prof, praline/util/cext.c:63-95, one track set): for a resident
// profile row y (sequence one) and a streamed profile row x (sequence two)
//     m[y][x] = sum over the nonzero (i, p1) of P1[y], ascending i,
//               sum over the nonzero (j, p2) of P2[x], ascending j:   fl( fl(p2 * S[i][j]) * p1 )
// accumulated in that order with one rounding per multiply and per add (the x86-64 gcc build of the reference,
// DESIGN.md section 2).  Two roundings per term is the algorithm: the FP32 pipe is the roofline (2 lane-operations
// per term), and k_build_rows_t (general.cu) spends an issue slot per lane-operation plus its loads -- the ncu capture
// profiles/r02_kbuildrows_t2_* has the FMA pipe 50 % busy at 70 % issue utilisation.
//
// Here a thread owns TWO resident columns and walks the streamed rows two at a time, the two rows packed in one
// 64-bit register: the first products of a PAIR of streamed rows are tabulated interleaved,
//     T2[pair][u][b] = ( fl(p2A_b * S[i_u][jA_b]), fl(p2B_b * S[i_u][jB_b]) ),
// so that one LDS.128 brings two terms of both rows and every FP instruction is a packed one: FFMA2 with a -0 addend
// (= the rounded product; the addend comes in as a kernel argument because ptxas contracts mul.rn.f32x2 +
// add.rn.f32x2 into one FFMA2, which would drop a rounding) and FADD2.  A packed instruction holds the pipe for two
// cycles but takes ONE issue slot: per resident symbol and 20-entry rows that is 80 packed instructions, 10 LDS.128
// and 1 LDS.64 in 160 pipe cycles -- loads and loop run in the pipe's shadow.
//
// The resident side is made uniform: the block takes the UNION of the symbols that occur in its resident columns
// (dense preprofiles: all 20 residues) and keeps p1 per (union symbol, column), zero where a column lacks the symbol.
// A zero p1 adds fl(t * 0) = 0, which leaves every partial sum as it is, so the result is bit-identical to the
// compacted walk -- and the table address no longer depends on the lane (one broadcast wavefront per LDS.128, none
// of the bank conflicts of per-lane symbol lists).  Sparse profiles stay on k_build_rows_t: the host passes the
// number of symbols in use only when the batch is dense (Engine.align_profile_pairs).
#include "common.cuh"

namespace {

constexpr int X2_RB = 8;                // streamed rows per pass (4 packed pairs)
constexpr int X2_MAXW = 8;              // warps per block: 64 resident columns each

__device__ __forceinline__ uint64_t x2_pack(float lo, float hi)
{
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void x2_unpack(uint64_t v, float& lo, float& hi)
{
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
// fl(t * p) in both halves: FFMA2 with the -0 addend nz (x * y + -0 rounds exactly like x * y, signed zeros included)
__device__ __forceinline__ uint64_t x2_mul(uint64_t t, uint64_t p, uint64_t nz)
{
    uint64_t r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(t), "l"(p), "l"(nz));
    return r;
}
__device__ __forceinline__ uint64_t x2_add(uint64_t a, uint64_t b)
{
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// one pair of streamed rows against the thread's two resident columns (TWO = false: a warp whose second column group
// lies beyond the resident walks column a only -- the pipe time goes to the other warps); N2 LDS.128 per union symbol
template <int N2, bool TWO>
__device__ __forceinline__ void x2_cells(uint32_t t_s, uint32_t t_stride, uint32_t pd_s, uint32_t pd_stride, int U,
                                         uint64_t nz, uint64_t& ra, uint64_t& rb)
{
    uint64_t acca = 0ull, accb = 0ull;
    for (int u = 0; u < U; u++) {
        float pa, pb;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pa), "=f"(pb) : "r"(pd_s));
        const uint64_t pa2 = x2_pack(pa, pa), pb2 = x2_pack(pb, pb);
        uint64_t t[2 * N2];
#pragma unroll
        for (int b = 0; b < N2; b++)
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t[2 * b]), "=l"(t[2 * b + 1]) : "r"(t_s + (uint32_t)b * 16u));
#pragma unroll
        for (int k = 0; k < 2 * N2; k++) {
            acca = x2_add(acca, x2_mul(t[k], pa2, nz));
            if (TWO) accb = x2_add(accb, x2_mul(t[k], pb2, nz));
        }
        pd_s += pd_stride;
        t_s += t_stride;
    }
    ra = acca;
    rb = accb;
}

// rows with more entries than the instantiated N2 (alphabets above 32 entries never get here)
__device__ __forceinline__ void x2_cells_any(int n2, uint32_t t_s, uint32_t t_stride, uint32_t pd_s, uint32_t pd_stride, int U,
                                             uint64_t nz, uint64_t& ra, uint64_t& rb)
{
    uint64_t acca = 0ull, accb = 0ull;
    for (int u = 0; u < U; u++) {
        float pa, pb;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(pa), "=f"(pb) : "r"(pd_s));
        const uint64_t pa2 = x2_pack(pa, pa), pb2 = x2_pack(pb, pb);
        for (int b = 0; b < n2; b++) {
            uint64_t t0, t1;
            asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(t0), "=l"(t1) : "r"(t_s + (uint32_t)b * 16u));
            acca = x2_add(acca, x2_mul(t0, pa2, nz));
            accb = x2_add(accb, x2_mul(t0, pb2, nz));
            acca = x2_add(acca, x2_mul(t1, pa2, nz));
            accb = x2_add(accb, x2_mul(t1, pb2, nz));
        }
        pd_s += pd_stride;
        t_s += t_stride;
    }
    ra = acca;
    rb = accb;
}

__global__ void __launch_bounds__(32 * X2_MAXW) k_build_rows_x2(const float* __restrict__ prof, const int64_t* __restrict__ rowoff,
                                                                int A, int UC, const float* __restrict__ S,
                                                                const PgRowBlock* __restrict__ blocks, int width, int xblocks,
                                                                float padv, uint64_t nz, float* __restrict__ mwave)
{
    extern __shared__ __align__(16) float sh[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = blockDim.x >> 5;
    const int STB = (UC + 1) & ~1;                                  // entries per table row (pairs of floats)
    float* T2 = sh;                                                 // [RB/2][UC][STB][2]
    float* pd = T2 + (X2_RB / 2) * UC * STB * 2;                    // [UC][NW][32][2]: p1 of (union symbol, my two columns)
    float* sS = pd + UC * NW * 64;                                  // [A][A]
    float* sval = sS + ((A * A + 3) & ~3);                          // [RB][A] streamed rows, compacted: values
    int* sj = reinterpret_cast<int*>(sval + X2_RB * A);             // [RB][A] ... and symbols
    __shared__ int scnt[X2_RB];
    __shared__ int usym[32];
    __shared__ unsigned umask;

    const PgRowBlock blk = blocks[blockIdx.x / xblocks];
    const int xa = (int)(blockIdx.x % xblocks) * 64 * NW + warp * 64 + lane, xb = xa + 32;
    const int64_t q0 = rowoff[blk.res];
    const int Lr = (int)(rowoff[blk.res + 1] - q0);
    const bool minea = xa < width && xa < Lr, mineb = xb < width && xb < Lr;
    if (tid == 0) umask = 0u;
    for (int i = tid; i < A * A; i += blockDim.x) sS[i] = S[i];
    __syncthreads();
    const float* rowa = prof + (size_t)(q0 + (minea ? xa : 0)) * A;
    const float* rowb = prof + (size_t)(q0 + (mineb ? xb : 0)) * A;
    {
        unsigned mk = 0u;
        for (int i = 0; i < A; i++) {
            if (minea && rowa[i] != 0.f) mk |= 1u << i;
            if (mineb && rowb[i] != 0.f) mk |= 1u << i;
        }
        mk = __reduce_or_sync(0xffffffffu, mk);
        if (lane == 0 && mk) atomicOr(&umask, mk);
    }
    __syncthreads();
    const unsigned um = umask;
    const int U = __popc(um);
    if (tid < 32 && ((um >> tid) & 1u)) usym[__popc(um & ((1u << tid) - 1u))] = tid;
    __syncthreads();
    const bool two = __any_sync(0xffffffffu, mineb);                // warp-uniform
    const bool poisoned = U > UC;                                   // the host's symbol count was wrong: fail loudly (NaN rows)
    if (!poisoned) {
        for (int k = 0; k < U; k++) {
            const int i = usym[k];
            reinterpret_cast<float2*>(pd)[(k * NW + warp) * 32 + lane] = make_float2(minea ? rowa[i] : 0.f, mineb ? rowb[i] : 0.f);
        }
    }
    const uint32_t pd_s = (uint32_t)__cvta_generic_to_shared(pd) + (uint32_t)(warp * 32 + lane) * 8u;
    const uint32_t pd_stride = (uint32_t)NW * 256u;
    const uint32_t T_s = (uint32_t)__cvta_generic_to_shared(T2);
    const uint32_t t_stride = (uint32_t)STB * 8u;                   // one union symbol's row of pairs
    const uint32_t pair_stride = (uint32_t)UC * t_stride;

    for (int r0 = 0; r0 < blk.rows; r0 += X2_RB) {
        const int nr = min(X2_RB, blk.rows - r0);
        const int nrp = (nr + 1) & ~1;
        __syncthreads();                                            // the previous pass is done with T2 / sval
        for (int r = warp; r < nrp; r += NW) {                      // compact the streamed rows (ascending symbol)
            int c = 0;
            if (r < nr && !(blk.dummy && r0 + r == 0)) {
                const float* src = prof + (size_t)(blk.src0 + r0 + r) * A;
                const float p = lane < A ? src[lane] : 0.f;
                const unsigned m = __ballot_sync(0xffffffffu, p != 0.f);
                if (p != 0.f) { const int at = r * A + __popc(m & ((1u << lane) - 1u)); sval[at] = p; sj[at] = lane; }
                c = __popc(m);
            }
            if (lane == 0) scnt[r] = c;
        }
        __syncthreads();
        if (!poisoned) {
            // T2[pair][u][b][half] = fl(p2_b * S[i_u][j_b]) of row 2 * pair + half, zero beyond the row's entries
            for (int item = tid; item < nrp * STB; item += blockDim.x) {
                const int r = item / STB, b = item - r * STB;
                const bool on = b < min(scnt[r], STB);
                const float p2 = on ? sval[r * A + b] : 0.f;
                const float* scol = sS + (on ? sj[r * A + b] : 0);
                float* dst = T2 + ((size_t)(r >> 1) * UC * STB + b) * 2 + (r & 1);
                for (int k = 0; k < U; k++) dst[(size_t)k * STB * 2] = on ? __fmul_rn(p2, scol[usym[k] * A]) : 0.f;
            }
        }
        __syncthreads();
        if (xa < width) {
            for (int r = 0; r < nr; r += 2) {
                const int nmax = max(scnt[r], scnt[r + 1]);
                const int n2 = (nmax + 1) >> 1;
                const uint32_t t_s = T_s + (uint32_t)(r >> 1) * pair_stride;
                uint64_t va = 0ull, vb = 0ull;
                if (nmax > STB || poisoned) {
                    va = vb = 0x7fc000007fc00000ull;
                } else {
                    switch (n2) {                                   // uniform over the block
                        case 0: break;
                        case 1: if (two) x2_cells<1, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<1, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 2: if (two) x2_cells<2, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<2, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 3: if (two) x2_cells<3, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<3, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 4: if (two) x2_cells<4, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<4, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 5: if (two) x2_cells<5, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<5, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 6: if (two) x2_cells<6, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<6, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 7: if (two) x2_cells<7, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<7, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 8: if (two) x2_cells<8, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<8, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 9: if (two) x2_cells<9, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<9, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 10: if (two) x2_cells<10, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<10, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 11: if (two) x2_cells<11, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<11, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        case 12: if (two) x2_cells<12, true>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); else x2_cells<12, false>(t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                        default: x2_cells_any(n2, t_s, t_stride, pd_s, pd_stride, U, nz, va, vb); break;
                    }
                }
                float a0, a1, b0, b1;                               // (row r, row r + 1) of column a / column b
                x2_unpack(va, a0, a1);
                x2_unpack(vb, b0, b1);
                const bool d0 = blk.dummy && r0 + r == 0;           // the region's dummy row is all zeros
                float* o0 = mwave + (size_t)(blk.row0 + r0 + r) * width;
                o0[xa] = d0 ? 0.f : minea ? __fadd_rn(0.f, a0) : padv;
                if (xb < width) o0[xb] = d0 ? 0.f : mineb ? __fadd_rn(0.f, b0) : padv;
                if (r + 1 < nr) {
                    float* o1 = o0 + width;
                    o1[xa] = minea ? __fadd_rn(0.f, a1) : padv;
                    if (xb < width) o1[xb] = mineb ? __fadd_rn(0.f, b1) : padv;
                }
            }
        }
    }
}

}  // namespace

// ucap: an upper bound of the number of distinct symbols with a nonzero entry anywhere in `prof` (sizes the tables;
// a block that finds more writes NaN rows).  Resident = sequence one only (the engine's choice for exact batches).
int pg_launch_build_rows_x2(const float* prof, const int64_t* rowoff, int A, const float* S, const PgRowBlock* blocks,
                            int n_blocks, int width, float padv, int ucap, float* mwave, cudaStream_t st)
{
    if (n_blocks <= 0) return 0;
    if (A > 32 || ucap < 1) { pg_set_error("packed score rows: alphabet size %d above 32 or no symbol count", A); return 1; }
    if (ucap > A) ucap = A;
    int nw = (width + 63) / 64;
    if (nw > X2_MAXW) nw = X2_MAXW;
    if (nw < 1) nw = 1;
    const int xblocks = (width + 64 * nw - 1) / (64 * nw);
    const int64_t nb = (int64_t)n_blocks * xblocks;
    if (nb > 0x7fffffffll) { pg_set_error("wave too large for one launch (%lld blocks)", (long long)nb); return 1; }
    const int stb = (ucap + 1) & ~1;
    const size_t sm = sizeof(float) * ((size_t)(X2_RB / 2) * ucap * stb * 2 + (size_t)ucap * nw * 64 + ((A * A + 3) & ~3) +
                                       2 * (size_t)X2_RB * A) + 16;
    PG_CUDA_OK(cudaFuncSetAttribute(k_build_rows_x2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm));
    k_build_rows_x2<<<(unsigned)nb, 32 * nw, sm, st>>>(prof, rowoff, A, ucap, S, blocks, width, xblocks, padv,
                                                        0x8000000080000000ull, mwave);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
