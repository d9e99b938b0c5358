// microbench.cu -- issue-rate micro-benchmarks behind the DP roofline denominator.
//
// the DP fill is bound by CUDA-core instruction issue, and
// MEASURED_PEAKS.json has no such figure, so the box is probed directly.  Each kernel runs
// 32 warps per SM, every thread carrying 8 independent dependency chains of one instruction
// kind; rates are wall-clock (CUDA events) warp-instructions per nanosecond per SM.
#include "common.cuh"
#include "../../include/praline_b200.h"

#include <vector>

#define CH 8
#define ITERS 2048          // clock64-timed burst (SM clock probe)
#define LONG_ITERS 32768    // event-timed runs behind the reported rates

template <int OP>
__global__ void __launch_bounds__(1024) k_rate(const float* fin, const int* iin, float* fout, long long* cyc,
                                               unsigned long long* ns = nullptr, int iters = ITERS)
{
    float f[CH];
    int v[CH];
    const float c1 = fin[0], c2 = fin[1];
    const int i1 = iin[0], i2 = iin[1];
#pragma unroll
    for (int i = 0; i < CH; i++) { f[i] = fin[2 + i] + threadIdx.x; v[i] = iin[2 + i] + threadIdx.x; }
    __shared__ float4 sh[1024];
    sh[threadIdx.x] = make_float4(c1, c2, c1, c2);
    __syncthreads();
    unsigned long long g0 = 0, g1 = 0;
    if (ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
            if (OP == 1) {   // alternate max / min so that ptxas cannot fuse pairs into FMNMX3
                if (it & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
                else asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2));
            }
            if (OP == 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
            if (OP == 3) {  // the score-only cell: 4 FADD, 2 FMNMX, 1 FMNMX3
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
                asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2));
                asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
                asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2));
            }
            if (OP == 4) v[i] = __viaddmax_s32(v[i], i1, i2);
            if (OP == 5) v[i] = (int)__viaddmax_s16x2((unsigned)v[i], (unsigned)i1, (unsigned)i2);
            if (OP == 6) f[i] = __shfl_up_sync(0xffffffffu, f[i], 1);
            if (OP == 9) v[i] = (int)__funnelshift_l((unsigned)i1, (unsigned)v[i], 1);
            if (OP == 10) v[i] = v[i] * i1 + i2;
            if (OP == 11) v[i] = (v[i] & i1) ^ i2;
            if (OP == 12) v[i] = v[i] + i1 + i2;
            if (OP == 13) {  // the packed score-only recurrence of two cells: 2 VIADD.16x2, 2 VIADDMNMX.S16x2, 1 VIMNMX3.S16x2
                unsigned x = (unsigned)v[i];
                const unsigned m = __vadd2(x, (unsigned)i1);
                const unsigned u = __viaddmax_s16x2(x, (unsigned)i2, m);
                const unsigned l = __viaddmax_s16x2(m, (unsigned)i2, u);
                const unsigned d = __vimax3_s16x2(m, u, l);
                v[i] = (int)__vadd2(d, (unsigned)i2);
            }
            if (OP == 7) {
                const float4 q = sh[(threadIdx.x + i + it) & 1023];
                f[i] += q.x;
            }
        }
    }
    const long long t1 = clock64();
    if (ns) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    if (ns && threadIdx.x == 0) ns[blockIdx.x] = g1 - g0;
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) acc += f[i] + (float)v[i];
    fout[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// Rate of one instruction kind in warp-instructions per NANOSECOND per SM, wall-clock timed with
// CUDA events over a ~ms run (independent of any on-chip counter); divide by the SM clock in
// GHz for a per-clock figure.
template <int OP>
static int run_rate(int sms, const float* fin, const int* iin, float* fout, long long* cyc, double* rate)
{
    k_rate<OP><<<sms, 1024>>>(fin, iin, fout, cyc, nullptr, ITERS);   // warm-up
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k_rate<OP><<<sms, 1024>>>(fin, iin, fout, cyc, nullptr, LONG_ITERS);
    cudaEventRecord(e1);
    const cudaError_t launch_rc = cudaGetLastError(), sync_rc = cudaDeviceSynchronize();
    float ms = 0.f;
    if (launch_rc == cudaSuccess && sync_rc == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    PG_CUDA_OK(launch_rc);
    PG_CUDA_OK(sync_rc);
    const double per_it = (OP == 3) ? 7.0 : (OP == 13 ? 5.0 : 1.0);
    *rate = 32.0 * CH * (double)LONG_ITERS * per_it / ((double)ms * 1e6);
    return 0;
}

extern "C" int pgpu_microbench(double* out, int n)
{
    if (n < 14) { pg_set_error("pgpu_microbench needs room for 14 doubles"); return 1; }
    int dev = 0, sms = 0, khz = 0;
    PG_CUDA_OK(cudaGetDevice(&dev));
    PG_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    PG_CUDA_OK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev));
    float hf[2 + CH] = {0.25f, -0.5f, 1, 2, 3, 4, 5, 6, 7, 8};
    int hi[2 + CH] = {3, -7, 1, 2, 3, 4, 5, 6, 7, 8};
    // every device buffer of the probe is owned by this holder, so an early PG_CUDA_OK return frees them too
    struct Bufs {
        float *fin = nullptr, *fout = nullptr;
        int* iin = nullptr;
        long long* cyc = nullptr;
        unsigned long long* ns = nullptr;
        ~Bufs() { cudaFree(fin); cudaFree(fout); cudaFree(iin); cudaFree(cyc); cudaFree(ns); }
    } b;
    PG_CUDA_OK(cudaMalloc((void**)&b.fin, sizeof(hf)));
    PG_CUDA_OK(cudaMalloc((void**)&b.iin, sizeof(hi)));
    PG_CUDA_OK(cudaMalloc((void**)&b.fout, sizeof(float) * sms * 1024));
    PG_CUDA_OK(cudaMalloc((void**)&b.cyc, sizeof(long long) * sms));
    PG_CUDA_OK(cudaMalloc((void**)&b.ns, sizeof(unsigned long long) * sms));
    float *fin = b.fin, *fout = b.fout;
    int* iin = b.iin;
    long long* cyc = b.cyc;
    unsigned long long* ns = b.ns;
    PG_CUDA_OK(cudaMemcpy(fin, hf, sizeof(hf), cudaMemcpyHostToDevice));
    PG_CUDA_OK(cudaMemcpy(iin, hi, sizeof(hi), cudaMemcpyHostToDevice));
    int rc = 0;
    rc |= run_rate<0>(sms, fin, iin, fout, cyc, &out[0]);
    rc |= run_rate<1>(sms, fin, iin, fout, cyc, &out[1]);
    rc |= run_rate<2>(sms, fin, iin, fout, cyc, &out[2]);
    rc |= run_rate<3>(sms, fin, iin, fout, cyc, &out[3]);
    rc |= run_rate<4>(sms, fin, iin, fout, cyc, &out[4]);
    rc |= run_rate<5>(sms, fin, iin, fout, cyc, &out[5]);
    rc |= run_rate<6>(sms, fin, iin, fout, cyc, &out[6]);
    rc |= run_rate<7>(sms, fin, iin, fout, cyc, &out[7]);
    rc |= run_rate<9>(sms, fin, iin, fout, cyc, &out[9]);
    rc |= run_rate<10>(sms, fin, iin, fout, cyc, &out[10]);
    rc |= run_rate<11>(sms, fin, iin, fout, cyc, &out[11]);
    rc |= run_rate<12>(sms, fin, iin, fout, cyc, &out[12]);
    rc |= run_rate<13>(sms, fin, iin, fout, cyc, &out[13]);
    // SM clock held during a burst of the cell mix: SM cycles (clock64) over %globaltimer ns
    for (int rep = 0; rep < 20; rep++) k_rate<3><<<sms, 1024>>>(fin, iin, fout, cyc, ns, ITERS);
    PG_CUDA_OK(cudaDeviceSynchronize());
    long long c0 = 0;
    unsigned long long n0 = 0;
    PG_CUDA_OK(cudaMemcpy(&c0, cyc, sizeof(long long), cudaMemcpyDeviceToHost));
    PG_CUDA_OK(cudaMemcpy(&n0, ns, sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    out[8] = (n0 > 0) ? (double)c0 / (double)n0 * 1e3 : (double)khz / 1e3;   // MHz
    return rc;
}
