// pipe-rate microbenchmark: FADD / FMNMX / FADD2 / FMNMX3 and mixes (sm_100a); every instruction is an asm volatile
#include <cstdio>
#include <cuda_runtime.h>
#define NACC 32
#define ITERS 4096
typedef unsigned long long u64;
#define ADD(d, a, b) asm volatile("add.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define MAX(d, a, b) asm volatile("max.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define MAX3(d, a, b, c) asm volatile("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define ADD2(d, a, b) asm volatile("add.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b))
#define FMA(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
template <int MODE>
__global__ void __launch_bounds__(256) k(const float *x, float *y, int iters) {
    float acc[NACC], t[NACC];
    u64 p[NACC / 2], q[NACC / 2];
    float l = x[threadIdx.x + 512], l2 = x[threadIdx.x + 768];
    u64 ll = ((const u64 *)x)[threadIdx.x], ll2 = ((const u64 *)x)[threadIdx.x + 256];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = x[i], t[i] = x[i + 64];
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) p[i] = ((const u64 *)x)[i + 100], q[i] = p[i];
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (MODE == 0) {          // FADD only
#pragma unroll
            for (int i = 0; i < NACC; i++) ADD(acc[i], acc[i], l);
        } else if (MODE == 1) {   // FMNMX only
#pragma unroll
            for (int i = 0; i < NACC; i++) MAX(acc[i], acc[i], t[i]);
        } else if (MODE == 2) {   // FADD + FMNMX per pair (today's inner loop)
#pragma unroll
            for (int i = 0; i < NACC; i++) { float s; ADD(s, l, t[i]); MAX(acc[i], acc[i], s); }
        } else if (MODE == 3) {   // FADD2 only
#pragma unroll
            for (int i = 0; i < NACC / 2; i++) ADD2(p[i], p[i], ll);
        } else if (MODE == 4) {   // FMNMX3 only
#pragma unroll
            for (int i = 0; i < NACC; i++) MAX3(acc[i], acc[i], t[i], t[(i + 1) % NACC]);
        } else if (MODE == 5) {   // 2 FADD2 + 2 FMNMX3 per 4 pairs
#pragma unroll
            for (int i = 0; i < NACC / 2; i++) {
                u64 a, b;
                ADD2(a, ll, p[i]); ADD2(b, ll2, q[i]);
                float ax = __uint_as_float((unsigned)a), ay = __uint_as_float((unsigned)(a >> 32)), bx = __uint_as_float((unsigned)b), by = __uint_as_float((unsigned)(b >> 32));
                MAX3(acc[2 * i], acc[2 * i], ax, bx);
                MAX3(acc[2 * i + 1], acc[2 * i + 1], ay, by);
            }
        } else if (MODE == 6) {   // 2 FADD + 1 FMNMX3 per 2 pairs
#pragma unroll
            for (int i = 0; i < NACC; i++) { float s, u; ADD(s, l, t[i]); ADD(u, l2, t[(i + 1) % NACC]); MAX3(acc[i], acc[i], s, u); }
        } else if (MODE == 7) {   // FFMA 3-reg only
#pragma unroll
            for (int i = 0; i < NACC; i++) FMA(acc[i], acc[i], l, t[i]);
        } else if (MODE == 8) {   // 1 FADD2 + 2 FMNMX per 2 pairs
#pragma unroll
            for (int i = 0; i < NACC / 2; i++) {
                u64 a;
                ADD2(a, ll, p[i]);
                float ax = __uint_as_float((unsigned)a), ay = __uint_as_float((unsigned)(a >> 32));
                MAX(acc[2 * i], acc[2 * i], ax);
                MAX(acc[2 * i + 1], acc[2 * i + 1], ay);
            }
        }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i] + t[i];
#pragma unroll
    for (int i = 0; i < NACC / 2; i++) s += (float)(p[i] + q[i]);
    y[blockIdx.x * blockDim.x + threadIdx.x] = s + l + l2;
}
template <int MODE> void run(const char *name, double pairs_per_iter, double inst_per_iter, float *x, float *y, int wps) {
    int blocks = 148 * wps / 2;         // 256 threads = 8 warps = 2 per SMSP
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(x, y, ITERS);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(x, y, ITERS);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warps = blocks * 8.0, cyc = ms * 1e-3 * 1.965e9 * 148 * 4;       // SMSP-cycles available
    double per = cyc / (warps * ITERS);
    printf("%-38s warps/SMSP %d  %.3f ms  %.2f cycles per inst (%g per iteration), %.2f per pair\n", name, wps, ms, per / inst_per_iter, inst_per_iter,
           pairs_per_iter ? per / pairs_per_iter : 0.0);
}
int main() {
    float *x, *y; cudaMalloc(&x, 4096 * 4); cudaMalloc(&y, 148 * 8 * 256 * 4); cudaMemset(x, 0, 4096 * 4);
    for (int wps = 2; wps <= 8; wps *= 2) {
        run<0>("FADD", 0, 32, x, y, wps);
        run<7>("FFMA (3 reg)", 0, 32, x, y, wps);
        run<1>("FMNMX", 0, 32, x, y, wps);
        run<2>("FADD + FMNMX per pair", 32, 64, x, y, wps);
        run<3>("FADD2", 0, 16, x, y, wps);
        run<4>("FMNMX3", 0, 32, x, y, wps);
        run<5>("2 FADD2 + 2 FMNMX3 per 4 pairs", 64, 64, x, y, wps);
        run<6>("2 FADD + FMNMX3 per 2 pairs", 64, 96, x, y, wps);
        run<8>("FADD2 + 2 FMNMX per 2 pairs", 32, 48, x, y, wps);
    }
    return 0;
}
