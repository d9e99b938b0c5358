// gotoh_stream16t.cu -- traced instantiations of the paired-resident packed kernel (gotoh_stream16r.cuh):
// one fill of an unordered pair (r, s) feeds the walks of BOTH master-slave alignments (r, s) and (s, r)
// of a global preprofile (reference: praline/component/preprofile.py:127-154 runs them as two separate
// PairwiseAligner executions).  Own translation unit: the instantiations compile in parallel with the rest.
#include "gotoh_stream16r.cuh"

constexpr int kNW16t = 8;

template <int K>
static int launch16rt(const StreamArgs& a, int n_tiles, cudaStream_t st)
{
    constexpr int KP = (K + 3) & ~3;
    constexpr int NCH = (K + 3) / 4;
    const size_t smem = (size_t)a.A * NCH * 512 + 32 * (KP + 4) * 4 + kNW16t * (4 * (64 + 8) + 64) * sizeof(uint32_t) +
                        (size_t)kNW16t * 32 * (K | 1) * sizeof(uint32_t);      // + the per-warp traceback staging blocks
    auto kern = k_stream16rt<K, kNW16t>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    StreamArgs b = a;
    b.all_ones = -1;
    kern<<<n_tiles, kNW16t * 32, smem, st>>>(b);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_stream16rt(const StreamArgs& a, int n_tiles, int K, cudaStream_t st)
{
    if (n_tiles <= 0) return 0;
    switch (K) {
        case 1: return launch16rt<1>(a, n_tiles, st);
        case 2: return launch16rt<2>(a, n_tiles, st);
        case 3: return launch16rt<3>(a, n_tiles, st);
        case 4: return launch16rt<4>(a, n_tiles, st);
        case 6: return launch16rt<6>(a, n_tiles, st);
        case 8: return launch16rt<8>(a, n_tiles, st);
        case 10: return launch16rt<10>(a, n_tiles, st);
        case 12: return launch16rt<12>(a, n_tiles, st);
        case 13: return launch16rt<13>(a, n_tiles, st);
        case 14: return launch16rt<14>(a, n_tiles, st);
        case 16: return launch16rt<16>(a, n_tiles, st);
        case 20: return launch16rt<20>(a, n_tiles, st);
        case 24: return launch16rt<24>(a, n_tiles, st);
        case 32: return launch16rt<32>(a, n_tiles, st);
        default: pg_set_error("unsupported columns-per-lane K=%d", K); return 1;
    }
}
