// Issue / pipe micro-benchmark for the instruction mix of K1 (sm_100a): how many SMSP cycles do packed
// FFMA2, scalar FFMA and ALU-pipe (FMNMX / FSEL) instructions cost, alone and interleaved, at 1 / 2 / 3 / 4
// warps per scheduler?  Answers whether FFMA2 occupies the dispatch port for two cycles (then K1's bound is
// issue slots + FP2 count) or only the FMA pipe (then an ALU op can issue in its shadow).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o ubench_pipes ubench_pipes.cu && ./ubench_pipes
#include <cstdio>
#include <cuda_runtime.h>

#define CHAINS 8
#define UNROLL 16
#define ITERS 2000

#define F2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(b2), "l"(c2));
#define F1(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a1[i]) : "f"(b1), "f"(c1));
#define AL(i)                                                                  \
    if (u < CHAINS) asm volatile("max.f32 %0, %0, %1;" : "+f"(m1[i]) : "f"(c1)); \
    else asm volatile("min.f32 %0, %0, %1;" : "+f"(m1[i]) : "f"(b1));
#define XR(i)                                                                      \
    if (u < CHAINS) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x1[i]) : "r"(y1[i])); \
    else asm volatile("add.s32 %0, %0, %1;" : "+r"(x1[i]) : "r"(y1[i]));
#define A2(i) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(a2[i]) : "l"(d2[i]));
#define F3(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a2[i]) : "l"(d2[i]), "l"(e2[i]));
#define AR(i)                                                                      \
    if (u < CHAINS) asm volatile("max.f32 %0, %0, %1;" : "+f"(m1[i]) : "f"(n1[i])); \
    else asm volatile("min.f32 %0, %0, %1;" : "+f"(m1[i]) : "f"(n1[i]));
#define SR(i) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %1, %2, p;}" : "+f"(m1[i]) : "f"(n1[i]), "f"(a1[i]));
#define SE(i) asm volatile("{.reg .pred p; setp.gt.f32 p, %0, %1; selp.f32 %0, %1, %2, p;}" : "+f"(m1[i]) : "f"(c1), "f"(b1));

template <int MODE>
__global__ void k(unsigned long long* out, float seed, long long* cyc) {
    unsigned long long a2[CHAINS];
    float a1[CHAINS], m1[CHAINS], n1[CHAINS];
    int x1[CHAINS], y1[CHAINS];
    unsigned long long d2[CHAINS], e2[CHAINS];
    unsigned long long b2, c2;
    float b1 = seed, c1 = seed * 0.5f;
    asm("mov.b64 %0, {%1, %1};" : "=l"(b2) : "f"(b1));
    asm("mov.b64 %0, {%1, %1};" : "=l"(c2) : "f"(c1));
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a1[i] = seed + i;
        m1[i] = seed - i;
        asm("mov.b64 %0, {%1, %1};" : "=l"(a2[i]) : "f"(a1[i]));
        n1[i] = seed * (i + 2);
        x1[i] = __float_as_int(seed) + i;
        y1[i] = __float_as_int(seed) * (i + 3);
        asm("mov.b64 %0, {%1, %2};" : "=l"(d2[i]) : "f"(a1[i] * 0.5f), "f"(n1[i]));
        asm("mov.b64 %0, {%1, %2};" : "=l"(e2[i]) : "f"(a1[i] * 0.25f), "f"(n1[i] * 3.f));
    }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int u = 0; u < UNROLL / CHAINS * CHAINS; ++u) {
            const int i = u % CHAINS;
            if (MODE == 0) { F2(i) }
            if (MODE == 1) { F1(i) }
            if (MODE == 2) { AL(i) }
            if (MODE == 3) { F2(i) AL(i) }
            if (MODE == 4) { F1(i) AL(i) }
            if (MODE == 5) { F2(i) AL(i) XR(i) }
            if (MODE == 6) { F2(i) F1(i) }
            if (MODE == 7) { SE(i) }
            if (MODE == 8) { F2(i) SE(i) }
            if (MODE == 9) { A2(i) }
            if (MODE == 10) { F3(i) }
            if (MODE == 11) { A2(i) AR(i) }
            if (MODE == 12) { A2(i) AL(i) }
            if (MODE == 13) { F3(i) AR(i) }
            if (MODE == 14) { A2(i) AR(i) XR(i) }
            if (MODE == 18) { XR(i) }
            if (MODE == 19) { A2(i) XR(i) }
            if (MODE == 15) { AR(i) }
            if (MODE == 16) { A2(i) SR(i) }
            if (MODE == 17) { SR(i) }
        }
    }
    const long long t1 = clock64();
    unsigned long long s = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) s += a2[i] + __float_as_uint(a1[i]) + __float_as_uint(m1[i]) + d2[i] + e2[i] + __float_as_uint(n1[i]) + x1[i] + y1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int per_iter, unsigned long long* out, long long* cyc) {
    for (int wps = 1; wps <= 4; ++wps) {
        const int threads = 128 * wps;  // one CTA per SM, wps warps per scheduler
        k<MODE><<<148, threads>>>(out, 1.0001f, cyc);
        cudaDeviceSynchronize();
        k<MODE><<<148, threads>>>(out, 1.0001f, cyc);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        const double instr = (double)ITERS * UNROLL * per_iter * wps;  // warp instructions per scheduler
        printf("%-28s warps/SMSP %d: %.3f cycles per warp-instruction per SMSP (IPC %.3f)\n", name, wps, avg / instr,
               instr / avg);
    }
}

int main() {
    unsigned long long* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 512 * 8);
    cudaMalloc(&cyc, 148 * 8);
    run<0>("FFMA2", 1, out, cyc);
    run<1>("FFMA", 1, out, cyc);
    run<2>("FMNMX (alu)", 1, out, cyc);
    run<7>("FSETP+FSEL (alu)", 2, out, cyc);
    run<3>("FFMA2 + FMNMX 1:1", 2, out, cyc);
    run<4>("FFMA + FMNMX 1:1", 2, out, cyc);
    run<5>("FFMA2 + FMNMX + LOP3/IADD", 3, out, cyc);
    run<6>("FFMA2 + FFMA", 2, out, cyc);
    run<8>("FFMA2 + FSETP+FSEL", 3, out, cyc);
    run<9>("FADD2 rr (2 distinct 64b)", 1, out, cyc);
    run<10>("FFMA2 rrr (3 distinct 64b)", 1, out, cyc);
    run<15>("FMNMX rr", 1, out, cyc);
    run<11>("FADD2 rr + FMNMX rr", 2, out, cyc);
    run<12>("FADD2 rr + FMNMX r,const", 2, out, cyc);
    run<13>("FFMA2 rrr + FMNMX rr", 2, out, cyc);
    run<14>("FADD2 rr + FMNMX rr + LOP3/IADD rr", 3, out, cyc);
    run<18>("LOP3/IADD rr", 1, out, cyc);
    run<19>("FADD2 rr + LOP3/IADD rr", 2, out, cyc);
    run<17>("FSETP rr + FSEL rr", 2, out, cyc);
    run<16>("FADD2 rr + FSETP rr + FSEL rr", 3, out, cyc);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
