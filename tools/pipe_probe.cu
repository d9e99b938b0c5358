// pipe_probe.cu -- standalone issue-rate probe (wall-clock timed, long runs) used to pin the
// per-pipe rates quoted in DESIGN.md.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CH 8
template <int OP, int UN>
__global__ void __launch_bounds__(1024) k(const float* fin, const int* iin, float* fout, int iters)
{
    float f[CH]; int v[CH];
    const float c1 = fin[0], c2 = fin[1];
    const int i1 = iin[0], i2 = iin[1];
#pragma unroll
    for (int i = 0; i < CH; i++) { f[i] = fin[2 + i] + threadIdx.x; v[i] = iin[2 + i] + threadIdx.x; }
#pragma unroll UN
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (OP == 0) asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1));
            if (OP == 1) { if (it & 1) asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1)); else asm volatile("min.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2)); }
            if (OP == 2) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i]) : "f"(c1), "f"(c2));
            if (OP == 4) asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(v[i]) : "r"(i1));
            if (OP == 5) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(v[i]) : "r"(i1), "r"(i2));
            if (OP == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(v[i]) : "r"(i1), "r"(i2));
            if (OP == 7) asm volatile("add.s32 %0, %0, %1;" : "+r"(v[i]) : "r"(i1));
            if (OP == 8) { asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1)); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2)); }
            if (OP == 9) { asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1)); asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(v[i]) : "r"(i1)); }
            if (OP == 10) { asm volatile("{ .reg .pred p; setp.ge.f32 p, %0, %1; selp.f32 %0, %0, %2, p; }" : "+f"(f[i]) : "f"(c1), "f"(c2)); }
            if (OP == 11) { asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c1)); asm volatile("add.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2)); asm volatile("max.f32 %0, %0, %1;" : "+f"(f[i]) : "f"(c2)); }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < CH; i++) acc += f[i] + (float)v[i];
    fout[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int OP, int UN>
void run(const char* name, int per, int sms, const float* fin, const int* iin, float* fout)
{
    const int iters = 1 << 16;
    k<OP, UN><<<sms, 1024>>>(fin, iin, fout, 1024);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<OP, UN><<<sms, 1024>>>(fin, iin, fout, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double winst = 32.0 * CH * iters * per;            // warp-instructions per SM
    printf("%-28s unroll %d: %8.3f ms  %6.3f warp-instr/ns/SM  (= per clk at 1 GHz; divide by SM GHz)\n", name, UN, ms, winst / (ms * 1e6));
}
int main()
{
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float hf[2 + CH] = {0.25f, -0.5f, 1, 2, 3, 4, 5, 6, 7, 8}; int hi[2 + CH] = {3, -7, 1, 2, 3, 4, 5, 6, 7, 8};
    float *fin, *fout; int* iin;
    cudaMalloc(&fin, sizeof(hf)); cudaMalloc(&iin, sizeof(hi)); cudaMalloc(&fout, 4 * sms * 1024);
    cudaMemcpy(fin, hf, sizeof(hf), cudaMemcpyHostToDevice); cudaMemcpy(iin, hi, sizeof(hi), cudaMemcpyHostToDevice);
    run<0, 1>("FADD", 1, sms, fin, iin, fout); run<0, 4>("FADD", 1, sms, fin, iin, fout);
    run<3, 1>("FFMA", 1, sms, fin, iin, fout); run<3, 4>("FFMA", 1, sms, fin, iin, fout);
    run<1, 2>("FMNMX (min/max alternating)", 1, sms, fin, iin, fout);
    run<2, 1>("FMNMX3", 1, sms, fin, iin, fout);
    run<4, 1>("SHF", 1, sms, fin, iin, fout); run<5, 1>("IMAD", 1, sms, fin, iin, fout);
    run<6, 1>("LOP3", 1, sms, fin, iin, fout); run<7, 1>("IADD", 1, sms, fin, iin, fout); run<7, 4>("IADD", 1, sms, fin, iin, fout);
    run<8, 1>("FADD+FMNMX pairs", 2, sms, fin, iin, fout); run<9, 1>("FADD+SHF pairs", 2, sms, fin, iin, fout);
    run<10, 1>("FSETP+FSEL pairs", 2, sms, fin, iin, fout); run<11, 1>("2 FADD + 1 FMNMX", 3, sms, fin, iin, fout);
    return 0;
}
