// capi_batch.cu -- the two self-contained entry points of the C ABI:
//
//   pgpu_align_batch     many sequence-sequence pairs: planning (grouping by resident, tiles, traceback
//                        waves, border arrays) happens HERE, in C++, so that a C / C++ / ctypes caller
//                        needs nothing but device buffers of sequences and pair ids.  One call replaces
//                        n_pairs x PairwiseAligner.execute (praline/component/align.py:88-251).
//   pgpu_align_profiles  one profile-profile alignment: cext_build_scores + RawPairwiseAligner
//                        (praline/util/cext.c:308-455; praline/component/align.py:302-447).
//
// Both are thin hosts over the tile-level entry points of capi.cu (the Python engine plans the same
// way for its own batches; praline_b200/engine.py).  Temporary device memory comes from the stream-
// ordered allocator (cudaMallocAsync) and is released on the same stream.
#include "common.cuh"
#include "../../include/praline_b200.h"

#include <algorithm>
#include <cmath>
#include <numeric>
#include <string.h>
#include <vector>

namespace {

struct AsyncPool {   // device temporaries of one call, freed on the stream when the call returns
    cudaStream_t st;
    std::vector<void*> ptrs;
    explicit AsyncPool(cudaStream_t s) : st(s) {}
    ~AsyncPool() { for (void* p : ptrs) cudaFreeAsync(p, st); }
    template <typename T> T* get(size_t n)
    {
        void* p = nullptr;
        if (cudaMallocAsync(&p, std::max<size_t>(n, 1) * sizeof(T), st) != cudaSuccess) return nullptr;
        ptrs.push_back(p);
        return (T*)p;
    }
    template <typename T> T* upload(const std::vector<T>& v)
    {
        T* p = get<T>(v.size());
        if (p && !v.empty() && cudaMemcpyAsync(p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, st) != cudaSuccess)
            return nullptr;
        return p;
    }
};

__global__ void k_scatter_f32(int64_t n, const int64_t* order, const float* src, float* dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[order[i]] = src[i];
}
__global__ void k_scatter_i32(int64_t n, const int64_t* order, const int32_t* src, int32_t* dst)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[order[i]] = src[i];
}

// max(M, U, L) along row 0 / column 0 (component/align.py:367-385), same arithmetic as numpy there:
// int64 arange times an f32 array is f64, plus the f32 scalar, stored as f32.
struct Borders {
    std::vector<float> top, left;
    int code00, top_ramp, left_ramp;
    float left0, left1;
};
Borders make_borders(int mode, float go, float ge, int maxlen)
{
    Borders b;
    const bool u_zero = mode == PG_SG_BOTH || mode == PG_SG_ONE, l_zero = mode == PG_SG_BOTH || mode == PG_SG_TWO;
    std::vector<float> ramp(maxlen + 1, 0.f);
    for (int i = 0; i < maxlen; i++) ramp[i + 1] = (float)((double)i * (double)ge + (double)go);
    const float vals[3] = {0.f, u_zero ? 0.f : go - ge, l_zero ? 0.f : go - ge};
    int arg = 0;
    for (int k = 1; k < 3; k++) if (vals[k] > vals[arg]) arg = k;
    b.left = u_zero ? std::vector<float>(maxlen + 1, 0.f) : ramp;
    b.top = l_zero ? std::vector<float>(maxlen + 1, 0.f) : ramp;
    b.left[0] = b.top[0] = vals[arg];
    b.code00 = arg;
    b.top_ramp = !l_zero;
    b.left_ramp = !u_zero;
    b.left0 = u_zero ? 0.f : go;
    b.left1 = u_zero ? 0.f : ge;
    return b;
}

int k_class(int64_t len)
{
    for (int k = 1; k <= 32; k++) if (pg_stream_supported_k(k) && 32 * k >= len) return k;
    return -1;
}

}  // namespace

extern "C" {

int pgpu_align_batch(int mode, int64_t n_pairs, const uint8_t* seqs_dev, const int64_t* seq_offsets_dev,
                     const int32_t* pair_i_dev, const int32_t* pair_j_dev, const float* S_dev, int A,
                     float gap_open, float gap_extend, int want_paths, float* scores_out_dev,
                     int32_t* path_buf_dev, int32_t* path_len_out_dev, void* stream)
{
    if (mode < 0 || mode > 4) { pg_set_error("unknown alignment mode %d", mode); return 1; }
    if (A < 1 || A > 64) { pg_set_error("alphabet size %d outside 1..64", A); return 1; }
    if (n_pairs <= 0) return 0;
    if (!seqs_dev || !seq_offsets_dev || !pair_i_dev || !pair_j_dev || !S_dev || !scores_out_dev) { pg_set_error("null buffer"); return 1; }
    if (want_paths && (!path_buf_dev || !path_len_out_dev)) { pg_set_error("want_paths needs path_buf and path_len_out"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t n = n_pairs;

    // ---- the pair list and the sequence lengths, on the host ------------------------------------
    std::vector<int32_t> pi(n), pj(n);
    PG_CUDA_OK(cudaMemcpyAsync(pi.data(), pair_i_dev, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaMemcpyAsync(pj.data(), pair_j_dev, n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    std::vector<float> S((size_t)A * A);
    PG_CUDA_OK(cudaMemcpyAsync(S.data(), S_dev, S.size() * sizeof(float), cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    int32_t max_id = 0;
    for (int64_t k = 0; k < n; k++) {
        if (pi[k] < 0 || pj[k] < 0) { pg_set_error("negative sequence id in pair %lld", (long long)k); return 1; }
        max_id = std::max(max_id, std::max(pi[k], pj[k]));
    }
    const int n_seqs = max_id + 1;
    std::vector<int64_t> offs(n_seqs + 1);
    PG_CUDA_OK(cudaMemcpyAsync(offs.data(), seq_offsets_dev, offs.size() * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
    PG_CUDA_OK(cudaStreamSynchronize(st));
    std::vector<int64_t> len(n_seqs);
    int64_t maxlen = 0;
    for (int i = 0; i < n_seqs; i++) {
        len[i] = offs[i + 1] - offs[i];
        if (len[i] < 1) { pg_set_error("empty sequences cannot be aligned (sequence %d)", i); return 1; }
        maxlen = std::max(maxlen, len[i]);
    }
    float smax = 0.f;
    bool integral = gap_open == std::round(gap_open) && gap_extend == std::round(gap_extend);
    for (float v : S) { smax = std::max(smax, std::fabs(v)); integral = integral && v == std::round(v); }
    const bool exact = integral && (double)smax * maxlen + std::fabs(gap_open) + std::fabs(gap_extend) * 2 * maxlen < 8388608.0;
    if (want_paths && !exact) {
        pg_set_error("traced batches need integer-valued scores (tie flags from unrounded operands); use pgpu_align_profiles");
        return 4;
    }
    if (want_paths && mode == PG_LOCAL && (gap_open > 0.f || gap_extend > 0.f || maxlen >= (1 << 20))) {
        pg_set_error("local traced batches need gap penalties <= 0 and sequences below 2^20");
        return 4;
    }

    // ---- order: K class of the resident (sequence two), resident, pair number -------------------
    std::vector<int> kc(n_seqs);
    for (int i = 0; i < n_seqs; i++) kc[i] = k_class(len[i]);
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), (int64_t)0);
    for (int64_t k = 0; k < n; k++)
        if (kc[pj[k]] < 0) { pg_set_error("sequence two of pair %lld is longer than 1024: use pgpu_align_profiles", (long long)k); return 4; }
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) {
        if (kc[pj[a]] != kc[pj[b]]) return kc[pj[a]] < kc[pj[b]];
        return pj[a] < pj[b];
    });
    std::vector<int32_t> res_s(n), str_s(n);
    std::vector<int64_t> cs(n + 1, 0), caps(n), poff(n, 0);
    for (int64_t s = 0; s < n; s++) {
        res_s[s] = pj[order[s]];
        str_s[s] = pi[order[s]];
        cs[s + 1] = cs[s] + len[str_s[s]];
    }
    if (want_paths) {   // regions of the caller's path buffer, in PAIR order
        std::vector<int64_t> reg(n + 1, 0);
        for (int64_t k = 0; k < n; k++) reg[k + 1] = reg[k] + len[pi[k]] + len[pj[k]] + 2;
        for (int64_t s = 0; s < n; s++) poff[s] = reg[order[s]];
    }

    AsyncPool pool(st);
    const int NW = 8;
    int spw = 16;
    while (spw > 2 && n / (NW * spw) < 148 * 6) spw /= 2;
    const int tile = NW * spw;
    const int border_len = (int)std::max<int64_t>(1024, maxlen) + 3;
    Borders B = make_borders(mode, gap_open, gap_extend, border_len - 1);
    float* top_d = pool.upload(B.top);
    float* left_d = pool.upload(B.left);
    int32_t* str_d = pool.upload(str_s);
    int32_t* res_d = want_paths ? pool.upload(res_s) : nullptr;
    int64_t* order_d = pool.upload(order);
    int64_t* poff_d = want_paths ? pool.upload(poff) : nullptr;
    float* sc_d = pool.get<float>(n);
    int32_t* plen_d = want_paths ? pool.get<int32_t>(n) : nullptr;
    int32_t* pstart_d = want_paths ? pool.get<int32_t>(n) : nullptr;
    if (!top_d || !left_d || !str_d || !order_d || !sc_d || (want_paths && (!res_d || !poff_d || !plen_d || !pstart_d))) {
        pg_set_error("device allocation failed"); return 2;
    }
    const bool semi = mode != PG_GLOBAL;
    const int64_t tb_budget = (int64_t)1 << 30;   // 4 GiB of traceback words per wave

    int64_t a0 = 0;
    while (a0 < n) {   // one K class at a time
        const int K = kc[res_s[a0]];
        int64_t a1 = a0;
        while (a1 < n && kc[res_s[a1]] == K) a1++;
        std::vector<PgTile> tiles;
        for (int64_t s = a0; s < a1;) {
            int64_t e = s;
            while (e < a1 && e - s < tile && res_s[e] == res_s[s]) e++;
            PgTile t;
            memset(&t, 0, sizeof(t));
            t.resident = res_s[s]; t.stream_begin = (int32_t)s; t.stream_end = (int32_t)e; t.resident2 = -1; t.out_base = s;
            tiles.push_back(t);
            s = e;
        }
        if (!want_paths) {
            PgTile* tiles_d = pool.upload(tiles);
            uint64_t* keys = semi ? pool.get<uint64_t>(2 * n) : nullptr;
            if (!tiles_d || (semi && !keys)) { pg_set_error("device allocation failed"); return 2; }
            int rc = pgpu_align_tiles(mode, K, 0, seqs_dev, seq_offsets_dev, str_d, tiles_d, (int)tiles.size(), n, S_dev, A,
                                      gap_open, gap_extend, top_d, left_d, B.left0, B.left1, border_len, sc_d, keys,
                                      nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, stream);
            if (rc) return rc;
        } else {
            // traceback words per (tile, warp): dummy row + stream + drain in whole 32-step blocks, 8 rows per word
            std::vector<int64_t> words(tiles.size() * NW);
            for (size_t t = 0; t < tiles.size(); t++) {
                const int64_t tb = tiles[t].stream_begin, te = tiles[t].stream_end, per = (te - tb + NW - 1) / NW;
                for (int w = 0; w < NW; w++) {
                    const int64_t sb = std::min(tb + w * per, te), se = std::min(sb + per, te);
                    const int64_t rows = cs[se] - cs[sb];
                    words[t * NW + w] = se > sb ? ((rows + 1 + 31 + 31) / 32 * 32 / 8) * (K * 32) : 0;
                }
            }
            size_t lo = 0;
            while (lo < tiles.size()) {
                size_t hi = lo;
                int64_t acc = 0;
                while (hi < tiles.size()) {
                    int64_t tw = 0;
                    for (int w = 0; w < NW; w++) tw += words[hi * NW + w];
                    if (hi > lo && acc + tw > tb_budget) break;
                    acc += tw;
                    hi++;
                }
                const int64_t s_lo = tiles[lo].stream_begin, s_hi = tiles[hi - 1].stream_end, ns = s_hi - s_lo;
                std::vector<PgTile> wt(tiles.begin() + lo, tiles.begin() + hi);
                for (PgTile& t : wt) t.out_base -= s_lo;
                std::vector<int64_t> wbase((hi - lo) * NW, 0);
                for (size_t i = 1; i < wbase.size(); i++) wbase[i] = wbase[i - 1] + words[lo * NW + i - 1];
                PgTile* tiles_d = pool.upload(wt);
                int64_t* wbase_d = pool.upload(wbase);
                uint32_t* tb = pool.get<uint32_t>((size_t)acc);
                int32_t* emit_t = pool.get<int32_t>(ns);
                int64_t* pair_tb = pool.get<int64_t>(ns);
                uint64_t* keys = semi ? pool.get<uint64_t>(2 * ns) : nullptr;
                if (!tiles_d || !wbase_d || !tb || !emit_t || !pair_tb || (semi && !keys)) { pg_set_error("device allocation failed"); return 2; }
                int rc;
                if (mode == PG_LOCAL) {
                    rc = pgpu_align_tiles_local(K, seqs_dev, seq_offsets_dev, str_d, tiles_d, (int)wt.size(), ns, S_dev, A,
                                                gap_open, gap_extend, top_d, B.left0, B.left1, border_len, sc_d + s_lo, keys, tb,
                                                wbase_d, emit_t, pair_tb, nullptr, stream);
                    if (rc) return rc;
                    rc = pgpu_traceback_tiles_local(K, seq_offsets_dev, res_d + s_lo, str_d + s_lo, ns, keys, tb, emit_t, pair_tb,
                                                    B.code00, poff_d + s_lo, path_buf_dev, pstart_d + s_lo, plen_d + s_lo,
                                                    nullptr, nullptr, nullptr, A, nullptr, 0, 0.f, nullptr, nullptr, 0, stream);
                } else {
                    rc = pgpu_align_tiles(mode, K, 0, seqs_dev, seq_offsets_dev, str_d, tiles_d, (int)wt.size(), ns, S_dev, A,
                                          gap_open, gap_extend, top_d, left_d, B.left0, B.left1, border_len, sc_d + s_lo, keys,
                                          tb, wbase_d, emit_t, pair_tb, nullptr, nullptr, stream);
                    if (rc) return rc;
                    rc = pgpu_traceback_tiles(mode, K, 0, seq_offsets_dev, res_d + s_lo, str_d + s_lo, ns, keys, tb, emit_t,
                                              pair_tb, B.code00, B.top_ramp, B.left_ramp, poff_d + s_lo, path_buf_dev,
                                              pstart_d + s_lo, plen_d + s_lo, nullptr, nullptr, nullptr, A, nullptr, 0, 0.f, 0,
                                              stream);
                }
                if (rc) return rc;
                lo = hi;
            }
        }
        a0 = a1;
    }
    const unsigned blocks = (unsigned)((n + 255) / 256);
    k_scatter_f32<<<blocks, 256, 0, st>>>(n, order_d, sc_d, scores_out_dev);
    if (want_paths) k_scatter_i32<<<blocks, 256, 0, st>>>(n, order_d, plen_d, path_len_out_dev);
    PG_CUDA_OK(cudaGetLastError());
    return 0;
}

int pgpu_align_profiles(int mode, int n_sets, const float* const* P1_dev, const float* const* P2_dev,
                        const float* const* S_dev, const int* A, int L1, int L2, const float* g1_dev,
                        const float* g2_dev, const uint8_t* zmask_dev, float* score_out_dev, int32_t* path_out_dev,
                        int32_t* path_len_out_dev, void* stream)
{
    if (mode < 0 || mode > 4) { pg_set_error("unknown alignment mode %d", mode); return 1; }
    if (L1 < 1 || L2 < 1) { pg_set_error("empty sequence (L1=%d, L2=%d)", L1, L2); return 1; }
    if (!g1_dev || !g2_dev || !score_out_dev) { pg_set_error("null buffer"); return 1; }
    if ((path_out_dev == nullptr) != (path_len_out_dev == nullptr)) { pg_set_error("path_out and path_len_out go together"); return 1; }
    cudaStream_t st = (cudaStream_t)stream;
    AsyncPool pool(st);
    const int pitch = (L2 + 127) / 128 * 128;     // whole 128-column strips: the layout the wavefront kernel wants
    float* m = pool.get<float>((size_t)L1 * pitch);
    unsigned char* ws = pool.get<unsigned char>((size_t)pgpu_general_workspace_bytes(L1, L2));
    int32_t* cell = pool.get<int32_t>(4);
    int32_t* pstart = pool.get<int32_t>(1);
    if (!m || !ws || !cell || !pstart) { pg_set_error("device allocation failed"); return 2; }
    int rc = pgpu_build_scores(n_sets, P1_dev, P2_dev, S_dev, A, L1, L2, m, pitch, stream);
    if (rc) return rc;
    return pgpu_align_general(mode, L1, L2, m, pitch, g1_dev, g2_dev, 1, zmask_dev, L2 + 1, ws, score_out_dev, cell,
                              path_out_dev, path_out_dev ? pstart : nullptr, path_len_out_dev, nullptr, nullptr, stream);
}

}  // extern "C"
