// f32x2_probe.cu -- rate of the packed f32x2 add/mul of sm_100 against the scalar ones, alone and next to
// LDS.128 (the instruction mix of the exact profile score rows).  Wall-clock timed, long runs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CH 8
template <int OP>
__global__ void __launch_bounds__(512) k(const float* fin, float* fout, int iters)
{
    __shared__ __align__(16) float sh[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sh[i] = fin[i & 7] * 1e-3f;
    __syncthreads();
    float f[CH], g[CH];
    const float c1 = fin[0], c2 = fin[1];
#pragma unroll
    for (int i = 0; i < CH; i++) { f[i] = fin[2 + i] + threadIdx.x; g[i] = fin[2 + i] - threadIdx.x; }
    uint64_t cc;
    asm volatile("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c1), "f"(c2));
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sh) + (threadIdx.x & 3) * 16;
    for (int it = 0; it < iters; it++) {
        if (OP == 4 || OP == 5) {
            // one LDS.128 per 8 FP instructions (scalar) / per 4 packed ones
            float4 t;
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(t.x), "=f"(t.y), "=f"(t.z), "=f"(t.w) : "r"(sbase + ((it & 15) << 6)));
            if (OP == 4) {
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float tv = i == 0 ? t.x : i == 1 ? t.y : i == 2 ? t.z : t.w;
                    float p; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p) : "f"(tv), "f"(c1));
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i & 1]) : "f"(p));
                }
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const float tv = i == 0 ? t.x : i == 1 ? t.y : i == 2 ? t.z : t.w;
                    float p; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p) : "f"(tv), "f"(c2));
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(g[i & 1]) : "f"(p));
                }
            } else {
                uint64_t a, b, p, q, fa, ga;
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(t.x), "f"(t.y));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(b) : "f"(t.z), "f"(t.w));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(fa) : "f"(f[0]), "f"(f[1]));
                asm volatile("mov.b64 %0, {%1, %2};" : "=l"(ga) : "f"(g[0]), "f"(g[1]));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(a), "l"(cc));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(b), "l"(cc));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(fa) : "l"(p));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(ga) : "l"(q));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(b), "l"(cc));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(q) : "l"(a), "l"(cc));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(fa) : "l"(p));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(ga) : "l"(q));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[0]), "=f"(f[1]) : "l"(fa));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(g[0]), "=f"(g[1]) : "l"(ga));
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
            if (OP == 1) { float p; asm volatile("mul.rn.f32 %0, %1, %2;" : "=f"(p) : "f"(g[i]), "f"(c1)); asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(p)); }
            if (OP == 2) {
                uint64_t a; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f[i]), "f"(g[i]));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(cc));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(g[i]) : "l"(a));
            }
            if (OP == 3) {
                uint64_t a, p; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(a) : "f"(f[i]), "f"(g[i]));
                asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(p) : "l"(a), "l"(cc));
                asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a) : "l"(p));
                asm volatile("mov.b64 {%0, %1}, %2;" : "=f"(f[i]), "=f"(g[i]) : "l"(a));
            }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) acc += f[i] + g[i];
    fout[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP>
void run(const char* name, double flops_per_iter_thread, double inst_per_iter, int sms, const float* fin, float* fout)
{
    const int iters = 1 << 17;
    k<OP><<<sms * 2, 512>>>(fin, fout, 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP><<<sms * 2, 512>>>(fin, fout, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double winst = 32.0 * iters * inst_per_iter;              // FP warp-instructions per SM (32 warps)
    const double fl = 1024.0 * iters * flops_per_iter_thread;       // f32 results per SM
    printf("%-44s %8.3f ms  %6.3f FP warp-instr/ns/SM  %7.2f f32 results/ns/SM\n", name, ms, winst / (ms * 1e6), fl / (ms * 1e6));
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float hf[2 + CH] = {0.25f, -0.5f, 1, 2, 3, 4, 5, 6, 7, 8};
    float *fin, *fout;
    cudaMalloc(&fin, sizeof(hf)); cudaMalloc(&fout, 4 * sms * 1024);
    cudaMemcpy(fin, hf, sizeof(hf), cudaMemcpyHostToDevice);
    run<0>("FADD", CH, CH, sms, fin, fout);
    run<1>("FMUL+FADD", 2 * CH, 2 * CH, sms, fin, fout);
    run<2>("add.rn.f32x2", 2 * CH, CH, sms, fin, fout);
    run<3>("mul.rn.f32x2 + add.rn.f32x2", 4 * CH, 2 * CH, sms, fin, fout);
    run<4>("LDS.128 + 8 FMUL + 8 FADD", 16, 16, sms, fin, fout);
    run<5>("LDS.128 + 4 mul.f32x2 + 4 add.f32x2", 16, 8, sms, fin, fout);
    return 0;
}
