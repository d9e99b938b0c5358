// merge.cu -- device-resident glue of the progressive merge.
//
// TreeMultipleSequenceAligner (praline/component/msa.py:124-237) aligns the count profiles of two clusters,
// merges them along the path (ProfileTrack.merge, container/sequence.py:205-239) and goes on with the merged
// profile.  The reference does this on the host between two aligner calls; here the count tables stay on the
// device between merges: the probability profile the aligner sees is derived from the counts by
// k_counts_to_profile, and the merged table is gathered along the traced path -- which never left the device
// -- by k_merge_counts.
#include "common.cuh"

// ProfileTrack.profile (container/sequence.py:191-203): totals = f32(sum of the int counts of a position),
// profile = f32(counts / totals) with the division carried out in float64 (int64 / float32 promotes to float64).
__global__ void k_counts_to_profile(const int32_t* __restrict__ counts, int64_t n_rows, int A, float* __restrict__ prof)
{
    const int64_t row = (int64_t)blockIdx.x * blockDim.y + threadIdx.y;
    if (row >= n_rows) return;
    const int32_t* c = counts + row * A;
    long long tot = 0;
    for (int i = threadIdx.x; i < A; i += 32) tot += c[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    const double total = (double)(float)tot;
    for (int i = threadIdx.x; i < A; i += 32) prof[row * A + i] = (float)((double)c[i] / total);
}

// Column r of the merged table = the count rows the path step r -> r + 1 consumes: row path[r+1].y - 1 of table one
// if y advances, row path[r+1].x - 1 of table two if x advances (container/sequence.py:224-237; the reference
// accumulates in f32 and ProfileTrack casts back to int, exact for counts).  hdr = the output header of
// pgpu_align_general: hdr[4] = first used row of the path region, hdr[5] = rows; rows are (y, x) pairs at hdr + 8.
__global__ void k_merge_counts(const int32_t* __restrict__ c1, const int32_t* __restrict__ c2, int A,
                               const int32_t* __restrict__ hdr, int32_t* __restrict__ out, int max_rows)
{
    const int start = hdr[4], len = hdr[5];
    const int2* path = reinterpret_cast<const int2*>(hdr + 8) + start;
    const int r = blockIdx.x * blockDim.y + threadIdx.y;
    if (r >= len - 1 || r >= max_rows) return;
    const int2 p0 = path[r], p1 = path[r + 1];
    const bool a1 = p1.x > p0.x, a2 = p1.y > p0.y;           // int2.x = y (sequence one), int2.y = x (sequence two)
    for (int i = threadIdx.x; i < A; i += 32) {
        int v = 0;
        if (a1) v += c1[(int64_t)(p1.x - 1) * A + i];
        if (a2) v += c2[(int64_t)(p1.y - 1) * A + i];
        out[(int64_t)r * A + i] = v;
    }
}

int pg_launch_counts_to_profile(const int32_t* counts, int64_t n_rows, int A, float* prof, cudaStream_t st)
{
    if (n_rows <= 0) return 0;
    dim3 b(32, 8);
    k_counts_to_profile<<<(unsigned)((n_rows + 7) / 8), b, 0, st>>>(counts, n_rows, A, prof);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pg_launch_merge_counts(const int32_t* c1, const int32_t* c2, int A, const int32_t* hdr, int32_t* out, int max_rows,
                           cudaStream_t st)
{
    if (max_rows <= 0) return 0;
    dim3 b(32, 8);
    k_merge_counts<<<(unsigned)((max_rows + 7) / 8), b, 0, st>>>(c1, c2, A, hdr, out, max_rows);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}
