/* capi_demo.c -- libpraline_b200.so from plain C: no Python, no torch.  Generates N sequences with a
 * small LCG, aligns all pairs (global, affine -11/-1, with paths) through pgpu_align_batch and prints
 *   <i> <j> <score> <path rows>
 * per pair; tests/test_capi_gpu.py builds and runs it and checks every line against the oracle.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/capi_demo.c -o capi_demo \
 *       -L praline_b200 -lpraline_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/praline_b200
 *   ./capi_demo matrix.f32 27 12 40        (matrix file: A*A little-endian f32, N sequences of length ~L)
 */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include "praline_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 2; } } while (0)

static uint32_t lcg(uint32_t* s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

int main(int argc, char** argv)
{
    if (argc < 5) { fprintf(stderr, "usage: %s matrix.f32 A n_seqs length\n", argv[0]); return 1; }
    const int A = atoi(argv[2]), n = atoi(argv[3]), L = atoi(argv[4]);
    float* S = (float*)malloc(sizeof(float) * A * A);
    FILE* f = fopen(argv[1], "rb");
    if (!f || fread(S, sizeof(float), (size_t)A * A, f) != (size_t)A * A) { fprintf(stderr, "cannot read %s\n", argv[1]); return 1; }
    fclose(f);
    if (pgpu_init(0)) { fprintf(stderr, "%s\n", pgpu_last_error()); return 3; }

    /* sequences: lengths L-3 .. L+3, symbols 0..19 */
    int64_t* offs = (int64_t*)malloc(sizeof(int64_t) * (n + 1));
    uint32_t seed = 12345u;
    offs[0] = 0;
    for (int i = 0; i < n; i++) offs[i + 1] = offs[i] + L - 3 + (int)(lcg(&seed) % 7u);
    uint8_t* seqs = (uint8_t*)malloc((size_t)offs[n]);
    for (int64_t k = 0; k < offs[n]; k++) seqs[k] = (uint8_t)(lcg(&seed) % 20u);
    const int64_t np = (int64_t)n * (n - 1) / 2;
    int32_t* pi = (int32_t*)malloc(sizeof(int32_t) * np);
    int32_t* pj = (int32_t*)malloc(sizeof(int32_t) * np);
    int64_t* reg = (int64_t*)malloc(sizeof(int64_t) * (np + 1));
    int64_t k = 0;
    reg[0] = 0;
    for (int i = 0; i < n; i++)
        for (int j = i + 1; j < n; j++, k++) {
            pi[k] = i; pj[k] = j;
            reg[k + 1] = reg[k] + (offs[i + 1] - offs[i]) + (offs[j + 1] - offs[j]) + 2;
        }

    uint8_t* d_seqs; int64_t* d_offs; int32_t *d_pi, *d_pj, *d_path, *d_plen; float *d_S, *d_scores;
    CK(cudaMalloc((void**)&d_seqs, (size_t)offs[n]));
    CK(cudaMalloc((void**)&d_offs, sizeof(int64_t) * (n + 1)));
    CK(cudaMalloc((void**)&d_pi, sizeof(int32_t) * np));
    CK(cudaMalloc((void**)&d_pj, sizeof(int32_t) * np));
    CK(cudaMalloc((void**)&d_S, sizeof(float) * A * A));
    CK(cudaMalloc((void**)&d_scores, sizeof(float) * np));
    CK(cudaMalloc((void**)&d_path, sizeof(int32_t) * 2 * (size_t)reg[np]));
    CK(cudaMalloc((void**)&d_plen, sizeof(int32_t) * np));
    CK(cudaMemcpy(d_seqs, seqs, (size_t)offs[n], cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_offs, offs, sizeof(int64_t) * (n + 1), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pi, pi, sizeof(int32_t) * np, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_pj, pj, sizeof(int32_t) * np, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_S, S, sizeof(float) * A * A, cudaMemcpyHostToDevice));

    if (pgpu_align_batch(PGPU_MODE_GLOBAL, np, d_seqs, d_offs, d_pi, d_pj, d_S, A, -11.0f, -1.0f, 1, d_scores, d_path,
                         d_plen, NULL)) {
        fprintf(stderr, "pgpu_align_batch: %s\n", pgpu_last_error());
        return 4;
    }
    CK(cudaDeviceSynchronize());
    float* scores = (float*)malloc(sizeof(float) * np);
    int32_t* plen = (int32_t*)malloc(sizeof(int32_t) * np);
    int32_t* path = (int32_t*)malloc(sizeof(int32_t) * 2 * (size_t)reg[np]);
    CK(cudaMemcpy(scores, d_scores, sizeof(float) * np, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(plen, d_plen, sizeof(int32_t) * np, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(path, d_path, sizeof(int32_t) * 2 * (size_t)reg[np], cudaMemcpyDeviceToHost));
    for (k = 0; k < np; k++) {
        printf("%d %d %.1f", pi[k], pj[k], scores[k]);
        const int32_t* p = path + 2 * (reg[k + 1] - plen[k]);       /* the path is right-aligned in its region */
        for (int r = 0; r < plen[k]; r++) printf(" %d,%d", p[2 * r], p[2 * r + 1]);
        printf("\n");
    }
    pgpu_shutdown();
    return 0;
}
