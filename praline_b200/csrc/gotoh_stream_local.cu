// gotoh_stream_local.cu -- the local traced instantiations of K2 (kernel template: gotoh_stream.cuh).
//
// Batched form of the reference's LocalMasterSlaveAligner inner loop (praline/component/
// preprofile.py:227-267): PairwiseAligner in mode "local" with zero_idxs = the bounding boxes of
// the alignments found in earlier Waterman-Eggert iterations.  Reference orientation only (resident
// = sequence two), because the end cell is the first ROW-MAJOR maximum of the reference's matrix.
#include "gotoh_stream.cuh"

constexpr int kNWL = 8;

template <int K, bool MK>
static int launch_local(const StreamArgs& a, int n_tiles, cudaStream_t st)
{
    constexpr int NCH = (K + 3) / 4;
    const size_t smem = (size_t)a.A * NCH * 512 + kNWL * 128 * sizeof(uint32_t);
    auto kern = k_stream<K, 1, true, false, false, MK, kNWL>;
    PG_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<n_tiles, kNWL * 32, smem, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

template <int K>
static int launch_local_k(const StreamArgs& a, int n_tiles, bool masked, cudaStream_t st)
{
    return masked ? launch_local<K, true>(a, n_tiles, st) : launch_local<K, false>(a, n_tiles, st);
}

int pg_launch_stream_local(const StreamArgs& a, int n_tiles, int K, bool masked, cudaStream_t st)
{
    if (n_tiles <= 0) return 0;
    if (a.transposed) { pg_set_error("local traced batches run in the reference orientation only"); return 1; }
    if (masked && !a.boxes) { pg_set_error("masked launch without boxes"); return 1; }
    switch (K) {
        case 1: return launch_local_k<1>(a, n_tiles, masked, st);
        case 2: return launch_local_k<2>(a, n_tiles, masked, st);
        case 3: return launch_local_k<3>(a, n_tiles, masked, st);
        case 4: return launch_local_k<4>(a, n_tiles, masked, st);
        case 6: return launch_local_k<6>(a, n_tiles, masked, st);
        case 8: return launch_local_k<8>(a, n_tiles, masked, st);
        case 10: return launch_local_k<10>(a, n_tiles, masked, st);
        case 12: return launch_local_k<12>(a, n_tiles, masked, st);
        case 13: return launch_local_k<13>(a, n_tiles, masked, st);
        case 14: return launch_local_k<14>(a, n_tiles, masked, st);
        case 16: return launch_local_k<16>(a, n_tiles, masked, st);
        case 20: return launch_local_k<20>(a, n_tiles, masked, st);
        case 24: return launch_local_k<24>(a, n_tiles, masked, st);
        case 32: return launch_local_k<32>(a, n_tiles, masked, st);
        default: pg_set_error("unsupported columns-per-lane K=%d", K); return 1;
    }
}
