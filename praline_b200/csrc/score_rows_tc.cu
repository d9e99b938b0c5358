// score_rows_tc.cu -- K1t, the tolerance-mode profile score rows on the 5th-generation tensor cores.
//
// What it replaces: cext_build_scores (praline/util/cext.c:308-455) for score-only batches of
// profile x profile pairs, m[y][x] = P1[y] . S . P2[x]^T, in the factored form of the fast path
// (k_build_rows_fast, general.cu): W = P_res . S^T per resident profile once, then a dense
// contraction  m[row][x] = sum_a P[row][a] * W[x][a]  over the alphabet (A <= 32).  Not the
// reference's evaluation order: scores agree to ~1e-6 relative (stated bound 1e-5), opt-in.
//
// Shape.  One CTA = one 128-row x NC-column tile: 4 row blocks of <= 32 matrix rows that share a
// resident (the wave's row blocks, engine.plan_profile_wave) x one column chunk of the resident
// (NC <= 256, a multiple of 16).  The contraction runs as tcgen05.mma kind::tf32 with an FP32-accurate
// split: every f32 operand is hi + lo with hi = tf32(x), lo = x - hi, and
//     m = P_hi.W_hi + P_lo.W_hi + P_hi.W_lo           (the lo.lo term is below 2^-22 relative)
// accumulated in one TMEM accumulator: 3 terms x 4 MMAs of K = 8 (alphabet padded to 32).
//   * operands are staged by the CTA's own threads straight into the canonical no-swizzle K-major
//     layout (8 x 16 B core matrices; SBO = 1024 B between 8-row groups, LBO = 128 B between the two
//     16-byte K chunks of an instruction) -- the rows are gathered from the profile store, so TMA has
//     nothing contiguous to fetch; fence.proxy.async hands them to the tensor core;
//   * one elected thread issues the 12 MMAs and commits them to an mbarrier;
//   * warps w and w + 4 own TMEM lanes 32w .. 32w+31 = the rows of row block w (even / odd 32-column
//     chunks): tcgen05.ld 32x32b.x32 hands a thread 32 columns of ITS row, so the chunk is transposed
//     through shared memory (the operand tiles are dead once the MMAs have committed; 16-byte pieces
//     XOR-swizzled by row, conflict-free both ways) and leaves as 128-bit stores in which a warp
//     writes four whole 128-byte row segments per instruction; pad columns, dummy rows and short
//     blocks are handled here.
// Bound: HBM writes, 4 B per cell (the same matrix the matrix-fed K2 streams back in); the tensor
// pipe needs 192 tf32 flop per cell and idles.  Two CTAs per SM (<= 96 KB smem, <= 256 TMEM columns
// each) so that one tile's stores overlap the other's staging + MMA.
#include "common.cuh"
#include <stdlib.h>

namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// canonical K-major, no swizzle: byte offset of element (row r, k) of an [R x 32] tf32 operand
__device__ __forceinline__ uint32_t core_off(int r, int k)
{
    return (uint32_t)((r >> 3) * 1024 + (k >> 2) * 128 + (r & 7) * 16 + (k & 3) * 4);
}

__device__ __forceinline__ uint64_t make_desc(uint32_t saddr)
{
    // cute::UMMA::SmemDescriptor: start >> 4 [0,14), LBO >> 4 [16,30), SBO >> 4 [32,46), version 1 [46,48),
    // base offset 0, lbo mode 0, layout type 0 (no swizzle) [61,64)
    return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)(128u >> 4) << 16) | ((uint64_t)(1024u >> 4) << 32) |
           ((uint64_t)1 << 46);
}

__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo)
{
    uint32_t h;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    lo = v - hi;
}

}  // namespace

// One 128-row tile: <= 4 row blocks of one resident, everything the CTA needs in ONE 128-byte line
// (the row-block list -> resident -> row offset chain would be three dependent round trips).
struct PgQuad {
    int64_t q0;                 // first profile row of the resident
    int32_t Lr;                 // resident length
    int32_t nblk;               // row blocks in this tile
    int64_t row0[4];            // first matrix row per block
    int64_t src0[4];            // profile row feeding it
    int32_t rows[4];
    int32_t dummy[4];
    int64_t bcan;               // first row of the resident in the pre-split canonical store (multiple of 8)
    int64_t _pad;
    int64_t can0[4];            // per block: its first row in the pre-split store of the STREAMED side when the
                                // block starts on an 8-row group there (then it is 4096 contiguous bytes), else -1
};
static_assert(sizeof(PgQuad) == 160, "PgQuad is ten 16-byte pieces");

struct RowsTcArgs {
    const float* prof;          // [rows x A] profile store (streamed side)
    const float* wres;          // [rows x A] W = P . S^T (or P . S) of the same store (resident side)
    const PgQuad* quads;
    const unsigned char* whi;   // residents pre-split into tf32 hi / lo in the canonical K-major layout (128 B per
    const unsigned char* wlo;   // row, 8-row groups of 1024 B), rows beyond a resident's length zero; NULL: gather
    const unsigned char* phi;   // the streamed side pre-split the same way (same row numbering as whi / wlo)
    const unsigned char* plo;
    int A, width, n_chunks, chunk;
    float padv;
    float* mwave;
};

constexpr int kTcThreads = 256;

__global__ void __launch_bounds__(kTcThreads) k_build_rows_tc(const RowsTcArgs a)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    __shared__ __align__(8) unsigned long long mbar;
    __shared__ __align__(8) unsigned long long mbar_b;      // completion of the B tiles' bulk copies (TMA)
    __shared__ uint32_t tmem_slot;
    // operands: A_hi | A_lo (16 KB each), B_hi | B_lo (chunk x 128 B each)
    unsigned char* sm = (unsigned char*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    unsigned char* A_hi = sm;
    unsigned char* A_lo = sm + 16384;
    unsigned char* B_hi = sm + 32768;
    unsigned char* B_lo = B_hi + (size_t)a.chunk * 128;

    __shared__ __align__(16) PgQuad quad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid < 10) reinterpret_cast<int4*>(&quad)[tid] = __ldg(reinterpret_cast<const int4*>(a.quads + blockIdx.x / a.n_chunks) + tid);
    const int c0 = (int)(blockIdx.x % a.n_chunks) * a.chunk;
    const int NC = min(a.chunk, a.width - c0);                 // columns of this tile, a multiple of 16

    // ---- TMEM accumulator: NC fp32 columns (power of two >= 32) ---------------------------------
    uint32_t ncols = 32;
    while ((int)ncols < NC) ncols <<= 1;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_slot)), "r"(ncols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar_b)));
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    const int64_t q0 = quad.q0;
    const int Lr = quad.Lr;
    const bool tma_b = a.whi != nullptr;
    if (tma_b && tid == 0) {
        // the resident's column chunk is ONE contiguous block of the pre-split store: two bulk copies
        const uint32_t bytes = (uint32_t)NC * 128u;
        const unsigned char* gh = a.whi + (size_t)(quad.bcan + c0) * 128;
        const unsigned char* gl = a.wlo + (size_t)(quad.bcan + c0) * 128;
        const uint32_t mb = smem_u32(&mbar_b);
        uint32_t a_bytes = 0;
        if (a.phi != nullptr)
            for (int b = 0; b < quad.nblk; b++) a_bytes += quad.can0[b] >= 0 ? 8192u : 0u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(2u * bytes + a_bytes) : "memory");
        if (a.phi != nullptr) {
            // a row block that starts on an 8-row group of the pre-split store is 32 rows x 128 B = 4096
            // contiguous bytes there AND in the A tile (rows 32b .. 32b+31): one bulk copy per part
            for (int b = 0; b < quad.nblk; b++) {
                if (quad.can0[b] < 0) continue;
                const size_t off = (size_t)quad.can0[b] * 128;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(A_hi + b * 4096)), "l"(a.phi + off), "r"(4096u), "r"(mb) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(smem_u32(A_lo + b * 4096)), "l"(a.plo + off), "r"(4096u), "r"(mb) : "memory");
            }
        }
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(B_hi)), "l"(gh), "r"(bytes), "r"(mb) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(B_lo)), "l"(gl), "r"(bytes), "r"(mb) : "memory");
    }

    // ---- stage the operands, split into tf32 hi / lo ---------------------------------------------
    // One operand row per thread: its <= 32 alphabet entries are independent loads (all in flight at
    // once), and the row's eight 16-byte K chunks go out as 128-bit shared stores -- a quarter warp
    // writes 8 consecutive rows of one core matrix = 128 contiguous bytes, conflict-free.
    auto stage_row = [&](const float* src, int r, unsigned char* hi_base, unsigned char* lo_base) {
        float v[32];
#pragma unroll
        for (int k = 0; k < 32; k++) v[k] = (src != nullptr && k < a.A) ? __ldg(src + k) : 0.f;
        const uint32_t o = core_off(r, 0);
#pragma unroll
        for (int j = 0; j < 8; j++) {
            float4 h, l;
            split_tf32(v[4 * j], h.x, l.x);
            split_tf32(v[4 * j + 1], h.y, l.y);
            split_tf32(v[4 * j + 2], h.z, l.z);
            split_tf32(v[4 * j + 3], h.w, l.w);
            *reinterpret_cast<float4*>(hi_base + o + j * 128) = h;
            *reinterpret_cast<float4*>(lo_base + o + j * 128) = l;
        }
    };
    for (int row = tid; row < (tma_b ? 128 : 128 + NC); row += kTcThreads) {
        if (row < 128) {
            const int b = row >> 5, rr = row & 31;
            const float* src = nullptr;
            if (a.phi != nullptr && b < quad.nblk && quad.can0[b] >= 0) continue;     // this block arrives by TMA
            if (b < quad.nblk && rr < quad.rows[b] && !(quad.dummy[b] && rr == 0))
                src = a.prof + (size_t)(quad.src0[b] + rr) * a.A;
            stage_row(src, row, A_hi, A_lo);
        } else {
            const int n = row - 128, x = c0 + n;
            stage_row(x < Lr ? a.wres + (size_t)(q0 + x) * a.A : nullptr, n, B_hi, B_lo);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // generic-proxy writes -> tensor core reads
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;

    // ---- 3 terms x 4 K-steps of tcgen05.mma kind::tf32, M = 128, N = NC -----------------------------
    if (warp == 0 && lane == 0) {
        if (tma_b) {     // the B tiles have landed (async proxy writes, ordered by the barrier)
            uint32_t done = 0;
            const uint32_t addr = smem_u32(&mbar_b);
            while (!done) {
                asm volatile(
                    "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                    : "=r"(done) : "r"(addr), "r"(0u) : "memory");
            }
        }
        // cute::UMMA::InstrDescriptor: D = f32 (1 << 4), A = B = tf32 (2 << 7, 2 << 10), both K-major, N >> 3 at 17, M >> 4 at 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NC >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t sa[3] = {smem_u32(A_hi), smem_u32(A_lo), smem_u32(A_hi)};
        const uint32_t sb[3] = {smem_u32(B_hi), smem_u32(B_hi), smem_u32(B_lo)};
#pragma unroll
        for (int t = 0; t < 3; t++) {
#pragma unroll
            for (int s = 0; s < 4; s++) {
                const uint64_t da = make_desc(sa[t] + s * 256), db = make_desc(sb[t] + s * 256);
                const uint32_t acc = (t | s) ? 1u : 0u;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                    ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&mbar)) : "memory");
    }
    {   // everybody waits for the accumulator (phase 0 of the barrier)
        uint32_t done = 0;
        const uint32_t addr = smem_u32(&mbar);
        while (!done) {
            asm volatile(
                "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                : "=r"(done) : "r"(addr), "r"(0u) : "memory");
        }
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // ---- epilogue: TMEM lane = matrix row; warps q and q + 4 write the rows of row block q -----------
    const int q = warp & 3, half = warp >> 2;
    if (q < quad.nblk) {
        const int brows = quad.rows[q];
        const bool bdummy = quad.dummy[q] != 0;
        const int64_t brow0 = quad.row0[q];
        unsigned char* tbuf = sm + warp * 4096;          // 32 rows x 128 B, inside the dead A tiles
        const uint32_t tb_s = smem_u32(tbuf);
        for (int c = half * 32; c < NC; c += 64) {
            uint32_t v[32];
            const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 8; j++)      // my row, 16-byte piece j, swizzled by the row
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(tb_s + lane * 128 + ((j ^ (lane & 7)) << 4)),
                             "r"(v[4 * j]), "r"(v[4 * j + 1]), "r"(v[4 * j + 2]), "r"(v[4 * j + 3]) : "memory");
            __syncwarp();
            const int j = lane & 7, x = c0 + c + 4 * j;
#pragma unroll
            for (int i = 0; i < 8; i++) {    // four rows per instruction, 128 contiguous bytes each
                const int rr = 4 * i + (lane >> 3);
                float4 o;
                asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(o.x), "=f"(o.y), "=f"(o.z), "=f"(o.w)
                             : "r"(tb_s + rr * 128 + ((j ^ (rr & 7)) << 4)) : "memory");
                if (rr < brows && c + 4 * j < NC) {
                    const bool dummy = bdummy && rr == 0;
                    float* of = reinterpret_cast<float*>(&o);
#pragma unroll
                    for (int e = 0; e < 4; e++) of[e] = dummy ? 0.f : (x + e < Lr ? of[e] : a.padv);
                    *reinterpret_cast<float4*>(a.mwave + (size_t)(brow0 + rr) * a.width + x) = o;
                }
            }
            __syncwarp();
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(ncols));
}

// Pre-split the resident side once per batch: W rows -> tf32 hi / lo in the canonical K-major layout, every
// sequence padded with zero rows to padoff[s + 1] - padoff[s] rows (a multiple of 32 covering its K class).
__global__ void k_split_residents(const float* __restrict__ wres, const int64_t* __restrict__ rowoff,
                                  const int64_t* __restrict__ padoff, int A, unsigned char* __restrict__ whi,
                                  unsigned char* __restrict__ wlo)
{
    const int s = blockIdx.x;
    const int64_t q0 = rowoff[s], p0 = padoff[s];
    const int L = (int)(rowoff[s + 1] - q0), P = (int)(padoff[s + 1] - p0);
    for (int idx = threadIdx.x; idx < P * 8; idx += blockDim.x) {
        const int x = idx >> 3, j = idx & 7;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; e++) v[e] = (x < L && 4 * j + e < A) ? wres[(size_t)(q0 + x) * A + 4 * j + e] : 0.f;
        float4 h, l;
        split_tf32(v[0], h.x, l.x);
        split_tf32(v[1], h.y, l.y);
        split_tf32(v[2], h.z, l.z);
        split_tf32(v[3], h.w, l.w);
        const int64_t r = p0 + x;
        const size_t o = (size_t)(r >> 3) * 1024 + (size_t)(r & 7) * 16 + (size_t)j * 128;
        *reinterpret_cast<float4*>(whi + o) = h;
        *reinterpret_cast<float4*>(wlo + o) = l;
    }
}

int pg_launch_split_residents(const float* wres, const int64_t* rowoff, const int64_t* padoff, int n_seqs, int A,
                              void* whi, void* wlo, cudaStream_t st)
{
    if (n_seqs <= 0) return 0;
    if (A < 1 || A > 32) { pg_set_error("tensor-core score rows: alphabet size %d above 32", A); return 1; }
    k_split_residents<<<n_seqs, 256, 0, st>>>(wres, rowoff, padoff, A, (unsigned char*)whi, (unsigned char*)wlo);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_build_rows_tc(const float* prof, const float* wres, int A, const void* quads, int n_quads, int width,
                            float padv, float* mwave, const void* whi, const void* wlo, const void* phi, const void* plo,
                            cudaStream_t st)
{
    if (n_quads <= 0) return 0;
    if (A < 1 || A > 32) { pg_set_error("tensor-core score rows: alphabet size %d above 32", A); return 1; }
    if (width < 32 || width % 32) { pg_set_error("tensor-core score rows: width %d is not a multiple of 32", width); return 1; }
    if (reinterpret_cast<uintptr_t>(mwave) & 15) { pg_set_error("tensor-core score rows: matrix not 16-byte aligned"); return 1; }
    RowsTcArgs a;
    a.prof = prof; a.wres = wres; a.quads = (const PgQuad*)quads;
    a.whi = (const unsigned char*)whi; a.wlo = (const unsigned char*)wlo;
    a.phi = (const unsigned char*)phi; a.plo = (const unsigned char*)plo;
    if ((phi == nullptr) != (plo == nullptr) || (phi != nullptr && whi == nullptr)) {
        pg_set_error("tensor-core score rows: phi / plo go together and need whi / wlo");
        return 1;
    }
    if ((whi == nullptr) != (wlo == nullptr)) { pg_set_error("tensor-core score rows: whi and wlo go together"); return 1; }
    a.A = A; a.width = width; a.padv = padv; a.mwave = mwave;
    // column chunks of <= 256 (two CTAs per SM).  Measured per wave: 256 -> 1.67 ms, 160 -> 1.70, 128 -> 1.77 (three CTAs
    // per SM, but the A tile is staged once per chunk), 64 -> 2.18; PGPU_TC_CHUNK overrides for experiments
    int max_chunk = 256;
    if (const char* e = getenv("PGPU_TC_CHUNK")) { const int v = atoi(e); if (v >= 16 && v <= 256) max_chunk = v / 16 * 16; }
    a.n_chunks = (width + max_chunk - 1) / max_chunk;
    a.chunk = ((width + a.n_chunks - 1) / a.n_chunks + 15) / 16 * 16;
    const int64_t nb = (int64_t)n_quads * a.n_chunks;
    if (nb > 0x7fffffffll) { pg_set_error("wave too large for one launch (%lld blocks)", (long long)nb); return 1; }
    const size_t smem = 1024 + 32768 + 2 * (size_t)a.chunk * 128;
    PG_CUDA_OK(cudaFuncSetAttribute(k_build_rows_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_build_rows_tc<<<(unsigned)nb, kTcThreads, smem, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
