// cluster.cu -- guide-tree clustering on the device.
//
// Replaces praline/util/cluster.py:27-57 (HierarchicalClusteringAlgorithm.merge_order) with
// its three linkages (:60-111).  The reference rebuilds the whole inter-cluster matrix
// `a` (float64) from the f32 distance matrix on every merge and takes `a.argmin()`; the
// observable contract is the ORDER of (merge_one_id, merge_two_id) pairs:
//   * clusters are enumerated in ascending original id (dict insertion order, :34; only
//     deletions happen), the diagonal is 2**32 (:12, :40);
//   * argmin is the FIRST minimum in row-major order (:49) -> smallest row id, then smallest
//     column id; cluster `two` is merged INTO `one` and deleted (:55-56);
//   * single / complete = min / max of the member distances, average = float64 mean of the
//     f32 member distances (:96-111), i.e. sum / (n1*n2).
// Here: S[i][j] holds the float64 sum (average) or min / max of the member distances and is
// updated in place on a merge (S[one][k] = S[one][k] (+|min|max) S[two][k]); a per-row cache
// of (minimum, smallest column achieving it) makes a merge O(n) plus the few rows whose cached
// neighbour was merged away.  For integer-valued distances (sequence scores) every sum is exact,
// so the merge order is identical to the reference's; for non-integer distances the float64 sums
// differ from numpy's pairwise summation in the last bits, which can only matter for ties closer
// than that (documented in DESIGN.md).
//
// Execution shape: the n-1 merges are a dependent chain whose cost per merge is latency, not
// bandwidth, so one COOPERATIVE kernel runs all of them with one grid-wide barrier per merge.
// Rows of S are dealt round-robin to the CTAs (row k belongs to CTA k mod G) and a row is only
// ever written by its owner: on a merge the owner of k folds S[k][two] into S[k][one] while the
// owner of `one` folds row `two` into row `one` -- the same sums from the symmetric copies, so S
// stays bitwise symmetric without any CTA writing another CTA's row.  Each CTA keeps the cached
// minima of its rows and a full copy of the cluster sizes in shared memory, publishes its best
// (value, row, column) before the barrier and picks the global first minimum from the G
// published entries after it.
#include "common.cuh"
#include <cooperative_groups.h>
namespace cg = cooperative_groups;

#define CL_T 1024
#define CL_DIAG 4294967296.0     // INFINITY of the reference (cluster.py:12)

struct ClusterArgs {
    int n, linkage;              // 0 single, 1 complete, 2 average
    const float* dist;           // [n][n] f32
    double* S;                   // [n][n]
    double* rmin;                // [n] cached row minimum of a[i][*]
    int* ridx;                   // [n] smallest column achieving it
    int* cnt;                    // [n] members
    int* alive;                  // [n]
    int* todo;                   // [n] rows to rescan this merge
    int32_t* merges;             // [n-1][2]
};

__device__ __forceinline__ double cl_value(const ClusterArgs& a, double s, int ci, int cj)
{
    return a.linkage == 2 ? s / (double)((long long)ci * (long long)cj) : s;
}

// (value, index) minimum with the smaller index winning ties
__device__ __forceinline__ void cl_min(double& v, int& i, double ov, int oi)
{
    if (ov < v || (ov == v && oi < i)) { v = ov; i = oi; }
}

__device__ __forceinline__ void cl_warp_min(double& v, int& i)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, i, o);
        cl_min(v, i, ov, oi);
    }
}

// row minimum of a[row][*] over alive columns != row (one warp)
__device__ __forceinline__ void cl_scan_row(const ClusterArgs& a, int row, int lane)
{
    const int n = a.n, ci = a.cnt[row];
    const double* srow = a.S + (size_t)row * n;
    double v = CL_DIAG;
    int idx = row;                      // an all-diagonal row can never win against a real pair
    for (int j = lane; j < n; j += 32) {
        if (j == row || !a.alive[j]) continue;
        cl_min(v, idx, cl_value(a, srow[j], ci, a.cnt[j]), j);
    }
    cl_warp_min(v, idx);
    if (lane == 0) { a.rmin[row] = v; a.ridx[row] = idx; }
}

__global__ void __launch_bounds__(256) k_cluster_init(const ClusterArgs a)
{
    const int n = a.n;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nw = (gridDim.x * blockDim.x) >> 5;
    for (int row = warp; row < n; row += nw) {
        const float* d = a.dist + (size_t)row * n;
        double* s = a.S + (size_t)row * n;
        double v = CL_DIAG;
        int idx = row;
        for (int j = lane; j < n; j += 32) {
            const double x = (double)d[j];
            s[j] = x;
            if (j != row) cl_min(v, idx, x, j);      // singletons: mean == min == max == the distance
        }
        cl_warp_min(v, idx);
        if (lane == 0) { a.rmin[row] = v; a.ridx[row] = idx; a.cnt[row] = 1; a.alive[row] = 1; }
    }
}

struct ClPub { double v; int row, col; };
#define CL_MAX_CTAS 256          // CTAs of the cooperative merge kernel (one per SM at most); sizes the pub records

// block-wide (value, index) minimum, smaller index wins ties; result valid in every thread
__device__ __forceinline__ void cl_block_min(double& v, int& i, double* sv, int* si)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    cl_warp_min(v, i);
    __syncthreads();                    // previous users of sv / si are done
    if (lane == 0) { sv[warp] = v; si[warp] = i; }
    __syncthreads();
    v = sv[lane]; i = si[lane];
    cl_warp_min(v, i);
}

__global__ void __launch_bounds__(CL_T) k_cluster_merge(const ClusterArgs a, ClPub* pub)
{
    extern __shared__ __align__(16) unsigned char csm[];
    __shared__ double sv[32];
    __shared__ int si[32];
    __shared__ int s_ntodo;
    cg::grid_group grid = cg::this_grid();
    const int n = a.n, tid = threadIdx.x;
    const int G = gridDim.x, c = blockIdx.x;
    const int nown = (n - c + G - 1) / G;             // rows c, c+G, c+2G, ...
    double* rmin = reinterpret_cast<double*>(csm);                    // [nown]
    int* cnt = reinterpret_cast<int*>(rmin + ((n + G - 1) / G + 1));  // [n]   0 = merged away
    int* ridx = cnt + n;                                              // [nown]
    int* todo = ridx + ((n + G - 1) / G + 1);                         // [nown]
    for (int i = tid; i < n; i += CL_T) cnt[i] = 1;
    for (int li = tid; li < nown; li += CL_T) { rmin[li] = a.rmin[c + li * G]; ridx[li] = a.ridx[c + li * G]; }
    __syncthreads();

    for (int it = 0; it < n - 1; it++) {
        // 1. this CTA's first minimum over its rows, published; after the barrier every CTA picks
        //    the global one: smallest value, then smallest row (its cached column is the smallest
        //    column of that row) -- the first minimum of the reference's row-major argmin
        double v = 2.0 * CL_DIAG;
        int row = 0x7fffffff;
        for (int li = tid; li < nown; li += CL_T) {
            const int k = c + li * G;
            if (cnt[k]) cl_min(v, row, rmin[li], k);
        }
        cl_block_min(v, row, sv, si);
        if (tid == 0) {
            ClPub p; p.v = v; p.row = row; p.col = row == 0x7fffffff ? 0 : ridx[(row - c) / G];
            pub[(size_t)(it & 1) * G + c] = p;
        }
        grid.sync();
        v = 2.0 * CL_DIAG; row = 0x7fffffff;
        int col = 0;
        for (int g = tid; g < G; g += CL_T) {
            const ClPub* q = pub + (size_t)(it & 1) * G + g;
            const double qv = __ldcg(&q->v);
            const int qr = __ldcg(&q->row);
            if (qv < v || (qv == v && qr < row)) { v = qv; row = qr; col = __ldcg(&q->col); }
        }
        {   // carry the column along with the (value, row) reduction
            const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int orow = __shfl_xor_sync(0xffffffffu, row, o), ocol = __shfl_xor_sync(0xffffffffu, col, o);
                if (ov < v || (ov == v && orow < row)) { v = ov; row = orow; col = ocol; }
            }
            __shared__ int sc[32];
            __syncthreads();
            if (lane == 0) { sv[warp] = v; si[warp] = row; sc[warp] = col; }
            __syncthreads();
            v = sv[lane]; row = si[lane]; col = sc[lane];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, v, o);
                const int orow = __shfl_xor_sync(0xffffffffu, row, o), ocol = __shfl_xor_sync(0xffffffffu, col, o);
                if (ov < v || (ov == v && orow < row)) { v = ov; row = orow; col = ocol; }
            }
        }
        const int one = row, two = col;
        if (c == 0 && tid == 0) { a.merges[2 * it] = one; a.merges[2 * it + 1] = two; }
        const int c1 = cnt[one] + cnt[two];
        if (tid == 0) s_ntodo = 0;
        __syncthreads();                 // everyone has read cnt[one], cnt[two]
        if (tid == 0) { cnt[one] = c1; cnt[two] = 0; }
        __syncthreads();

        // 2. own rows k != one: fold S[k][two] into S[k][one], then the cached minimum of row k.
        //    Only ONE entry of the row changed (column `two` is gone).  If the new a[k][one] is <= the
        //    cached minimum it is the new minimum, at the smallest column: on a tie the cached column
        //    is either unchanged and compared, or it was `one` / `two`, and then every other tied
        //    column is > two > one.  A row whose cached neighbour was `one` or `two` and whose new
        //    entry is larger is rescanned.
        for (int li = tid; li < nown; li += CL_T) {
            const int k = c + li * G;
            if (!cnt[k] || k == one) continue;
            double* rk = a.S + (size_t)k * n;
            const double x0 = rk[one], y0 = rk[two];
            const double s = a.linkage == 2 ? x0 + y0 : (a.linkage == 0 ? fmin(x0, y0) : fmax(x0, y0));
            rk[one] = s;
            const double x = cl_value(a, s, c1, cnt[k]);
            const int ri = ridx[li];
            double cv = rmin[li];
            if (ri == one || ri == two) {
                if (x <= cv) { rmin[li] = x; ridx[li] = one; }
                else todo[atomicAdd(&s_ntodo, 1)] = li;
                continue;
            }
            int cidx = ri;
            cl_min(cv, cidx, x, one);
            rmin[li] = cv; ridx[li] = cidx;
        }
        // 3. the owner of `one` folds row `two` into row `one` and takes the row minimum on the way
        if (one % G == c) {
            double* r1 = a.S + (size_t)one * n;
            const double* r2 = a.S + (size_t)two * n;
            double mv = CL_DIAG;
            int mi = one;
            for (int k = tid; k < n; k += CL_T) {
                if (!cnt[k] || k == one) continue;
                const double x0 = r1[k], y0 = __ldcg(r2 + k);
                const double s = a.linkage == 2 ? x0 + y0 : (a.linkage == 0 ? fmin(x0, y0) : fmax(x0, y0));
                r1[k] = s;
                cl_min(mv, mi, cl_value(a, s, c1, cnt[k]), k);
            }
            cl_block_min(mv, mi, sv, si);
            if (tid == 0) { rmin[(one - c) / G] = mv; ridx[(one - c) / G] = mi; }
        }
        __syncthreads();
        // 4. rescans, one row at a time with the whole CTA (a row is 8n bytes: many loads in flight)
        const int ntodo = s_ntodo;
        for (int q = 0; q < ntodo; q++) {
            const int li = todo[q], k = c + li * G;
            const double* rk = a.S + (size_t)k * n;
            const int ck = cnt[k];
            double mv = CL_DIAG;
            int mi = k;
            for (int j = tid; j < n; j += CL_T) {
                const double sj = rk[j];
                const int cj = cnt[j];
                if (cj && j != k) cl_min(mv, mi, cl_value(a, sj, ck, cj), j);
            }
            cl_block_min(mv, mi, sv, si);
            if (tid == 0) { rmin[li] = mv; ridx[li] = mi; }
        }
        __syncthreads();
    }
}

int pg_launch_cluster(int n, int linkage, const float* dist, void* work, int32_t* merges, cudaStream_t st)
{
    if (n < 1) { pg_set_error("clustering needs at least one object (n=%d)", n); return 1; }
    if (linkage < 0 || linkage > 2) { pg_set_error("unknown linkage %d", linkage); return 1; }
    if (n == 1) return 0;
    ClusterArgs a;
    a.n = n; a.linkage = linkage; a.dist = dist; a.merges = merges;
    unsigned char* w = (unsigned char*)work;
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    a.S = (double*)w; w += up(sizeof(double) * (size_t)n * n);
    a.rmin = (double*)w; w += up(sizeof(double) * n);
    a.ridx = (int*)w; w += up(sizeof(int) * n);
    a.cnt = (int*)w; w += up(sizeof(int) * n);
    a.alive = (int*)w; w += up(sizeof(int) * n);
    a.todo = (int*)w; w += up(sizeof(int) * n);
    ClPub* pub = (ClPub*)w;
    int dev = 0, sms = 148;
    PG_CUDA_OK(cudaGetDevice(&dev));
    PG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    int blocks = (n + 7) / 8;
    if (blocks > sms * 8) blocks = sms * 8;
    k_cluster_init<<<blocks, 256, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    // one CTA per SM at most (cooperative launch: all CTAs are co-resident), fewer for small n
    int G = n / 16;
    if (G < 1) G = 1;
    if (G > sms) G = sms;
    if (G > CL_MAX_CTAS) G = CL_MAX_CTAS;      // pub holds 2 * CL_MAX_CTAS records (pg_cluster_workspace_bytes)
    const size_t per = (size_t)(n + G - 1) / G + 1;
    const size_t smem = sizeof(double) * per + sizeof(int) * ((size_t)n + 2 * per) + 16;
    if (smem > 200 * 1024) { pg_set_error("clustering: %d objects exceed the shared-memory state of the kernel", n); return 1; }
    PG_CUDA_OK(cudaFuncSetAttribute(k_cluster_merge, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    void* args[] = {(void*)&a, (void*)&pub};
    PG_CUDA_OK(cudaLaunchCooperativeKernel((void*)k_cluster_merge, dim3(G), dim3(CL_T), args, smem, st));
    return 0;
}

// ---- distance matrix of GuideTreeBuilder (component/tree.py:92-147) ----------------------------------
// d[i][j] = d[j][i] = score of pair (i, j), d[i][i] = 0.0 (tree.py:132-133); dist = (-d) + d.max(), all f32.
// The condensed vector may be laid out in rank slices (the padded all-gather buffer of parallel.py):
// slot s of the np.triu_indices order lives at cond[s + shift[r]] for cuts[r] <= s < cuts[r+1]
// (n_cuts = 0: plain vector).
struct TreeDistArgs {
    int n, n_cuts;
    const float* cond;
    const int64_t* cuts;      // [n_cuts + 1] first slot of every rank slice
    const int64_t* shift;     // [n_cuts] element offset added to a slot of that slice
    float* dist;              // [n][n]
    unsigned* maxkey;         // ordered key of d.max(), seeded with the diagonal's 0.0
};

__device__ __forceinline__ unsigned td_key(float v)
{
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float td_unkey(unsigned k)
{
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ int64_t td_where(const TreeDistArgs& a, int64_t slot)
{
    for (int r = 0; r < a.n_cuts; r++)
        if (slot < a.cuts[r + 1]) return slot + a.shift[r];
    return slot;
}

__global__ void k_tree_max(const TreeDistArgs a, int64_t n_pairs)
{
    float m = 0.0f;   // the diagonal takes part in d.max()
    for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_pairs; s += (int64_t)gridDim.x * blockDim.x)
        m = fmaxf(m, a.cond[td_where(a, s)]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(a.maxkey, td_key(m));
}

// one 32 x 32 tile of the upper triangle per CTA: coalesced read of the condensed rows, coalesced
// writes of the tile and of its mirror image (transposed through shared memory)
__global__ void __launch_bounds__(256) k_tree_fill(const TreeDistArgs a)
{
    __shared__ float tile[32][33];
    const int bj = blockIdx.x, bi = blockIdx.y;
    if (bi > bj) return;
    const float mx = td_unkey(*a.maxkey);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int64_t n = a.n;
    for (int r = ty; r < 32; r += 8) {
        const int64_t i = (int64_t)bi * 32 + r, j = (int64_t)bj * 32 + tx;
        float d = 0.0f;
        if (i < n && j < n && i != j) {
            const int64_t lo = i < j ? i : j, hi = i < j ? j : i;
            d = a.cond[td_where(a, lo * n - lo * (lo + 1) / 2 + (hi - lo - 1))];
        }
        const float v = (-d) + mx;
        tile[r][tx] = v;
        if (i < n && j < n) a.dist[i * n + j] = v;
    }
    if (bi == bj) return;      // the diagonal tile is symmetric in itself
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int64_t j = (int64_t)bj * 32 + r, i = (int64_t)bi * 32 + tx;
        if (i < n && j < n) a.dist[j * n + i] = tile[tx][r];
    }
}

int pg_launch_tree_distance(int n, const float* cond, int n_cuts, const int64_t* cuts, const int64_t* shift,
                            float* dist, unsigned* scratch, cudaStream_t st)
{
    if (n < 1) return 0;
    TreeDistArgs a;
    a.n = n; a.n_cuts = n_cuts; a.cond = cond; a.cuts = cuts; a.shift = shift; a.dist = dist; a.maxkey = scratch;
    const unsigned zero_key = 0x80000000u;      // td_key(0.0f)
    PG_CUDA_OK(cudaMemcpyAsync(scratch, &zero_key, sizeof(unsigned), cudaMemcpyHostToDevice, st));
    const int64_t n_pairs = (int64_t)n * (n - 1) / 2;
    if (n_pairs > 0) {
        int64_t blocks = (n_pairs + 1023) / 1024;
        if (blocks > 148 * 8) blocks = 148 * 8;
        k_tree_max<<<(unsigned)blocks, 256, 0, st>>>(a, n_pairs);
        PG_CUDA_OK(cudaGetLastError());
    }
    const unsigned t = (unsigned)((n + 31) / 32);
    k_tree_fill<<<dim3(t, t), 256, 0, st>>>(a);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

size_t pg_cluster_workspace_bytes(int n)
{
    auto up = [](size_t b) { return (b + 255) & ~(size_t)255; };
    return up(sizeof(double) * (size_t)n * n) + up(sizeof(double) * n) + 4 * up(sizeof(int) * n) + up(2 * CL_MAX_CTAS * sizeof(ClPub));
}
