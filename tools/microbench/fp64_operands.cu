// fp64_operands.cu - what the sm_100a FP64 pipe sustains for different operand mixes.
// DFMA with one, two or three DISTINCT register operands per instruction (the rest from the
// constant bank / uniform registers), DMUL, DADD and a DFMA/DMUL/DADD blend, each as 8
// independent chains per thread on a full grid (8 warps per SM sub-partition).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_operands fp64_operands.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kChains = 8;

template <int kKind>
__global__ void __launch_bounds__(256) probe(double *out, const double *in, int iters, double m, double c) {
    double a[kChains], b[kChains], d[kChains];
    unsigned x[kChains];   // kinds 8 - 10: independent non-FP64 work issued between the DFMAs
    float fl[kChains];
    for (int i = 0; i < kChains; i++) {
        x[i] = threadIdx.x * 2654435761u + i;
        fl[i] = (float)(threadIdx.x + i);
        a[i] = in[(threadIdx.x + i) & 255];
        b[i] = in[(threadIdx.x + 2 * i + 1) & 255] * 1e-3 + 0.999;
        d[i] = in[(threadIdx.x + 3 * i + 2) & 255] * 1e-9;
    }
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < kChains; i++) {
            if (kKind == 0) a[i] = fma(a[i], m, c);              // 1 register operand (+ 2 constants)
            if (kKind == 1) a[i] = fma(a[i], b[i], c);           // 2 register operands
            if (kKind == 2) a[i] = fma(a[i], b[i], d[i]);        // 3 distinct register operands
            if (kKind == 3) a[i] = fma(b[i], d[(i + 1) % kChains], a[i]);  // 3 regs, accumulate form
            if (kKind == 4) a[i] = a[i] * b[i];                  // DMUL 2 regs
            if (kKind == 5) a[i] = a[i] + d[i];                  // DADD 2 regs
            if (kKind == 6) a[i] = fma(b[0], d[0], a[i]);        // 3 regs, two shared by all chains (.reuse)
            if (kKind == 8 || kKind == 9 || kKind == 10 || kKind == 11) a[i] = fma(a[i], b[i], c);   // 2 regs
            if (kKind == 8 || kKind == 9) x[i] = x[i] * 1664525u + 1013904223u;                        // + 1 IMAD
            if (kKind == 9) x[i] ^= x[i] >> 7;                                                       // + SHF / LOP3
            if (kKind == 10) fl[i] = fmaf(fl[i], 0.999f, 0.25f);                                     // + 1 FFMA
            if (kKind == 11) { fl[i] = fmaf(fl[i], 0.999f, 0.25f); x[i] = x[i] * 1664525u + 1013904223u; x[i] ^= x[i] >> 7; }
            if (kKind == 7) {                                    // blend ~ kernel mix: 58% DFMA(3 reg) 29% DMUL 13% DADD
                a[i] = fma(a[i], b[i], d[i]);
                if ((i & 1) == 0) a[i] = a[i] * b[(i + 1) % kChains];
                if ((i & 3) == 0) a[i] = a[i] + d[(i + 2) % kChains];
            }
        }
    }
    double s = 0;
    for (int i = 0; i < kChains; i++) s += a[i] + (double)x[i] + (double)fl[i];
    if (s == 123.456) out[0] = s;
}

template <int kKind>
static void run(const char *name, double ops_per_iter, double *out, double *in, int sms) {
    const int iters = 1 << 14;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<kKind><<<sms * 8, 256>>>(out, in, 64, 0.999999999, 1e-12);
    cudaEventRecord(e0);
    probe<kKind><<<sms * 8, 256>>>(out, in, iters, 0.999999999, 1e-12);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double inst = (double)sms * 8 * 256 * iters * ops_per_iter;   // thread instructions
    const double per_sm_clk = inst / 32.0 / sms / (ms * 1e-3 * 1.965e9);   // warp instr / SM / clk at 1965 MHz
    printf("%-44s %8.3f ms  %7.2f Tinst/s  %5.3f warp-inst/clk/SM (peak 2.0)\n", name, ms, inst / ms / 1e9, per_sm_clk);
}

int main() {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    double *out, *in;
    cudaMalloc(&out, 8);
    cudaMalloc(&in, 256 * 8);
    double h[256];
    for (int i = 0; i < 256; i++) h[i] = 1.0 + i * 1e-3;
    cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
    run<0>("DFMA a = a*const + const (1 reg)", 8, out, in, sms);
    run<1>("DFMA a = a*b + const (2 regs)", 8, out, in, sms);
    run<2>("DFMA a = a*b + d (3 regs)", 8, out, in, sms);
    run<3>("DFMA a = b*d' + a (3 regs, accumulate)", 8, out, in, sms);
    run<6>("DFMA a = b0*d0 + a (3 regs, 2 shared)", 8, out, in, sms);
    run<4>("DMUL a = a*b (2 regs)", 8, out, in, sms);
    run<5>("DADD a = a+d (2 regs)", 8, out, in, sms);
    run<7>("blend 8 DFMA(3 reg) + 4 DMUL + 2 DADD", 14, out, in, sms);
    // does a non-FP64 instruction issued between two DFMAs cost FP64 throughput?  (DFMA count only)
    run<8>("DFMA (2 regs) + 1 IMAD each: DFMA rate", 8, out, in, sms);
    run<9>("DFMA (2 regs) + IMAD + SHF + LOP3 each: DFMA rate", 8, out, in, sms);
    run<10>("DFMA (2 regs) + 1 FFMA each: DFMA rate", 8, out, in, sms);
    run<11>("DFMA (2 regs) + FFMA + IMAD + SHF + LOP3 each: DFMA rate", 8, out, in, sms);
    return 0;
}
