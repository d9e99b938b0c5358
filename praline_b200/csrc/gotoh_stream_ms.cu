// gotoh_stream_ms.cu -- the matrix-fed instantiations of K2 (kernel template: gotoh_stream.cuh): score-only
// batches of profile x profile pairs whose match scores come from a materialised wave in HBM (made by
// pgpu_build_rows / pgpu_build_rows_fast / pgpu_build_rows_tc) instead of the shared-memory substitution
// profile.  Replaces, per pair, cext_align_<mode> + the end-cell choice (praline/util/cext.c:99-306;
// praline/component/align.py:401-431) on a match-score matrix built by cext_build_scores.
// A translation unit of its own so that the library builds in parallel.
#include "gotoh_stream.cuh"

constexpr int kNWM = 8;

template <int K, int KM>
static int launch_ms(const StreamArgs& a, int n_tiles, cudaStream_t st)
{
    const size_t smem = kNWM * 128 * sizeof(uint32_t);
    auto kern = k_stream<K, KM, false, false, true, false, kNWM>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, kNWM * 32, smem, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int K>
static int launch_ms_k(const StreamArgs& a, int n_tiles, int km, cudaStream_t st)
{
    if (km == 0) return launch_ms<K, 0>(a, n_tiles, st);
    if (km == 1) return launch_ms<K, 1>(a, n_tiles, st);
    return launch_ms<K, 2>(a, n_tiles, st);
}

int pg_launch_stream_ms(const StreamArgs& a, int n_tiles, int K, int km, cudaStream_t st)
{
    switch (K) {
        case 1: return launch_ms_k<1>(a, n_tiles, km, st);
        case 2: return launch_ms_k<2>(a, n_tiles, km, st);
        case 3: return launch_ms_k<3>(a, n_tiles, km, st);
        case 4: return launch_ms_k<4>(a, n_tiles, km, st);
        case 6: return launch_ms_k<6>(a, n_tiles, km, st);
        case 8: return launch_ms_k<8>(a, n_tiles, km, st);
        case 10: return launch_ms_k<10>(a, n_tiles, km, st);
        case 12: return launch_ms_k<12>(a, n_tiles, km, st);
        case 13: return launch_ms_k<13>(a, n_tiles, km, st);
        case 14: return launch_ms_k<14>(a, n_tiles, km, st);
        case 16: return launch_ms_k<16>(a, n_tiles, km, st);
        case 20: return launch_ms_k<20>(a, n_tiles, km, st);
        case 24: return launch_ms_k<24>(a, n_tiles, km, st);
        case 32: return launch_ms_k<32>(a, n_tiles, km, st);
        default: pg_set_error("unsupported columns-per-lane K=%d", K); return 1;
    }
}
