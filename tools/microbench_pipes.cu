// Scratch micro-benchmark (not part of the library): issue / pipe throughput of the instruction mixes the vote kernel's
// inner loop can be built from, on one B200.  Prints warp-instructions per cycle per SM sub-partition.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/microbench_pipes tools/microbench_pipes.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned long long u64;
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define FMA1(d, a, b, c) asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define SHF(d, a, c) asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(d) : "r"(a), "r"(c))
#define MIN3(d, a, b, c) asm volatile("min.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MIN3A(d, a, b, c) asm volatile("{.reg .f32 t0, t1; abs.f32 t0, %2; abs.f32 t1, %3; min.f32 %0, %1, t0, t1;}" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define MADHI(d, a, c) asm volatile("mad.hi.u32 %0, %1, %3, %2;" : "=r"(d) : "r"(a), "r"(c), "r"(two))
#define MADLO(d, a, b, c) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define IADD(d, a, b) asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b))
#define LOP(d, a, b, c) asm volatile("lop3.b32 %0, %1, %2, %3, 0xea;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
#define FADD1(d, a, b) asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b))
#define FMASAT(d, a, b, c) asm volatile("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c))
#define POPC(d, a) asm volatile("popc.b32 %0, %1;" : "=r"(d) : "r"(a))
#define PRMT(d, a, b) asm volatile("prmt.b32 %0, %1, %2, 0xffbb;" : "=r"(d) : "r"(a), "r"(b))
#define HFMA2(d, a, b, c) asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c))
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

template <int MODE>
__global__ void __launch_bounds__(256) kern(float *out, int iters, float seed) {
    float f[16];
    u64 p[16];
    unsigned r[8];
#pragma unroll
    for (int k = 0; k < 16; ++k) { f[k] = seed * (float)(threadIdx.x + k + 1); p[k] = pk(f[k], f[k] * 1.5f); }
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = threadIdx.x * 977u + k;
    // per-thread values: operands live in ordinary registers, as in the vote kernel (not in uniform registers)
    const float m1 = 0.999f + (float)threadIdx.x * 1e-9f, c1 = 1e-4f + (float)threadIdx.x * 1e-11f;
    const u64 m2 = pk(m1, m1 + 2e-3f), c2 = pk(c1, 2.f * c1);
    const unsigned two = 2u + (threadIdx.x >> 20);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // FFMA x16
#pragma unroll
            for (int k = 0; k < 16; ++k) FMA1(f[k], f[k], m1, c1);
        } else if (MODE == 1) {   // FFMA2 x16
#pragma unroll
            for (int k = 0; k < 16; ++k) FMA2(p[k], p[k], m2, c2);
        } else if (MODE == 2) {   // SHF x16 (8 chains)
#pragma unroll
            for (int k = 0; k < 16; ++k) SHF(r[k & 7], r[(k + 1) & 7], r[k & 7]);
        } else if (MODE == 3) {   // FMNMX3 x16
#pragma unroll
            for (int k = 0; k < 16; ++k) MIN3A(f[k & 7], f[k & 7], f[8 + (k & 7)], f[(k + 3) & 7]);
        } else if (MODE == 4) {   // new vote mix: per 2 votes 5 FFMA2 + 2 SHF + 1 FMNMX3; x4
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W, s;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                FMA2(s, U, m2, W);
                float sa, sb;
                upk(s, sa, sb);
                SHF(r[k], __float_as_uint(sa), r[k]);
                SHF(r[k], __float_as_uint(sb), r[k]);
                MIN3A(f[k], f[k], sa, sb);
                p[k] = s;
            }
        } else if (MODE == 5) {   // r01 mix: 8 packed + 4 SHF per 2 votes; x4
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 a, b, U, W, s1, s2;
                FMA2(a, p[k], m2, p[8 + k]);
                FMA2(b, p[k + 4], m2, p[12 + k]);
                FMA2(U, a, c2, c2);
                FMA2(W, a, m2, c2);
                FMA2(U, b, m2, U);
                FMA2(W, b, c2, W);
                FMA2(s1, U, m2, W);
                FMA2(s2, U, c2, W);
                float x, y, z, w;
                upk(s1, x, y);
                upk(s2, z, w);
                SHF(r[k], __float_as_uint(x), r[k]);
                SHF(r[k], __float_as_uint(z), r[k]);
                SHF(r[k], __float_as_uint(y), r[k]);
                SHF(r[k], __float_as_uint(w), r[k]);
                p[k] = s1; p[k + 4] = s2;
            }
        } else if (MODE == 6) {   // two-threshold mix: 6 FFMA2 + 4 SHF per 2 votes; x4
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W, s1, s2;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                FMA2(s1, U, m2, W);
                FMA2(s2, U, c2, W);
                float x, y, z, w;
                upk(s1, x, y);
                upk(s2, z, w);
                SHF(r[k], __float_as_uint(x), r[k]);
                SHF(r[k + 4], __float_as_uint(z), r[k + 4]);
                SHF(r[k], __float_as_uint(y), r[k]);
                SHF(r[k + 4], __float_as_uint(w), r[k + 4]);
                p[k] = s1; p[k + 4] = s2;
            }
        } else if (MODE == 7) {   // scalar new mix: per vote 5 FFMA + SHF (+ FMNMX3 per 2); x8 votes
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float t, U, W, sa, sb;
                FMA1(t, f[k], m1, f[8 + k]);
                FMA1(U, f[k + 4], c1, t);
                FMA1(t, f[k], c1, f[12 + k]);
                FMA1(W, f[k + 4], m1, t);
                FMA1(sa, U, m1, W);
                FMA1(t, f[k], m1, f[12 + k]);
                FMA1(U, f[k + 4], c1, t);
                FMA1(t, f[k], c1, f[8 + k]);
                FMA1(W, f[k + 4], m1, t);
                FMA1(sb, U, m1, W);
                SHF(r[k], __float_as_uint(sa), r[k]);
                SHF(r[k], __float_as_uint(sb), r[k]);
                MIN3A(f[k], f[k], sa, sb);
                f[k + 4] = sa; f[k + 8] = sb;
            }
        } else if (MODE == 8) {   // 5 FFMA2 + 2 SHF (no FMNMX3)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W, s;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                FMA2(s, U, m2, W);
                float sa, sb;
                upk(s, sa, sb);
                SHF(r[k], __float_as_uint(sa), r[k]);
                SHF(r[k], __float_as_uint(sb), r[k]);
                p[k] = s;
            }
        } else if (MODE == 9) {   // 5 FFMA2 only
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W, s;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                FMA2(s, U, m2, W);
                p[k] = s;
            }
        } else if (MODE == 10) {  // FFMA2 + FFMA alternating 1:1
#pragma unroll
            for (int k = 0; k < 8; ++k) { FMA2(p[k], p[k], m2, c2); FMA1(f[k], f[k], m1, c1); }
        } else if (MODE == 11) {  // FFMA2 : SHF 2:1
#pragma unroll
            for (int k = 0; k < 8; ++k) { FMA2(p[k], p[k], m2, c2); FMA2(p[k + 8], p[k + 8], m2, c2); SHF(r[k], r[(k + 1) & 7], r[k]); }
        } else if (MODE == 12) {  // FFMA : SHF 1:1
#pragma unroll
            for (int k = 0; k < 8; ++k) { FMA1(f[k], f[k], m1, c1); SHF(r[k], r[(k + 1) & 7], r[k]); }
        } else if (MODE == 14) {  // IMAD.HI x16
#pragma unroll
            for (int k = 0; k < 16; ++k) MADHI(r[k & 7], r[(k + 1) & 7], r[k & 7]);
        } else if (MODE == 15) {  // IMAD.LO x16
#pragma unroll
            for (int k = 0; k < 16; ++k) MADLO(r[k & 7], r[(k + 1) & 7], r[(k + 2) & 7], r[k & 7]);
        } else if (MODE == 16) {  // IADD x16
#pragma unroll
            for (int k = 0; k < 16; ++k) IADD(r[k & 7], r[(k + 1) & 7], r[k & 7]);
        } else if (MODE == 17) {  // LOP3 x16
#pragma unroll
            for (int k = 0; k < 16; ++k) LOP(r[k & 7], r[(k + 1) & 7], r[(k + 2) & 7], r[k & 7]);
        } else if (MODE == 18) {  // FADD x16
#pragma unroll
            for (int k = 0; k < 16; ++k) FADD1(f[k], f[k], c1);
        } else if (MODE == 19) {  // FFMA.SAT x16
#pragma unroll
            for (int k = 0; k < 16; ++k) FMASAT(f[k], f[k], m1, c1);
        } else if (MODE == 20) {  // POPC x16
#pragma unroll
            for (int k = 0; k < 16; ++k) POPC(r[k & 7], r[(k + 1) & 7]);
        } else if (MODE == 21) {  // PRMT x16
#pragma unroll
            for (int k = 0; k < 16; ++k) PRMT(r[k & 7], r[(k + 1) & 7], r[k & 7]);
        } else if (MODE == 22) {  // HFMA2 x16
#pragma unroll
            for (int k = 0; k < 16; ++k) HFMA2(r[k & 7], r[k & 7], r[(k + 1) & 7], r[(k + 2) & 7]);
        } else if (MODE == 23) {  // 5 FFMA2 + 2 IMAD.HI + 1 FMNMX3 per 2 votes
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W, s;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                FMA2(s, U, m2, W);
                float sa, sb;
                upk(s, sa, sb);
                MADHI(r[k], __float_as_uint(sa), r[k]);
                MADHI(r[k], __float_as_uint(sb), r[k]);
                MIN3A(f[k], f[k], sa, sb);
                p[k] = s;
            }
        } else if (MODE == 24) {  // 4 FFMA2 + 2 FFMA.SAT + 2 FADD + 2 FFMA per 2 votes (saturating indicator, G1/G2 sums)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                float ua, ub, wa, wb, ga, gb;
                upk(U, ua, ub);
                upk(W, wa, wb);
                FMASAT(ga, ua, m1, wa);
                FMASAT(gb, ub, m1, wb);
                FADD1(f[k], f[k], ga);
                FADD1(f[k], f[k], gb);
                FMA1(f[k + 4], ga, ga, f[k + 4]);
                FMA1(f[k + 4], gb, gb, f[k + 4]);
                p[k] = W;
            }
        } else if (MODE == 25) {  // 5 FFMA2 + 1 PRMT + 0.5 IADD3(3-input) + 1 FMNMX3 per 2 votes
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                u64 t, U, W, s;
                FMA2(t, p[k], m2, p[8 + k]);
                FMA2(U, p[k + 4], c2, t);
                FMA2(t, p[k], c2, p[12 + k]);
                FMA2(W, p[k + 4], m2, t);
                FMA2(s, U, m2, W);
                float sa, sb;
                upk(s, sa, sb);
                unsigned w;
                PRMT(w, __float_as_uint(sa), __float_as_uint(sb));
                IADD(r[k], r[k], w);
                MIN3A(f[k], f[k], sa, sb);
                p[k] = s;
            }
        } else if (MODE == 13) {  // FFMA : SHF 2:1
#pragma unroll
            for (int k = 0; k < 8; ++k) { FMA1(f[k], f[k], m1, c1); FMA1(f[k + 8], f[k + 8], m1, c1); SHF(r[k], r[(k + 1) & 7], r[k]); }
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 16; ++k) { float a, b; upk(p[k], a, b); acc += f[k] + a + b; }
#pragma unroll
    for (int k = 0; k < 8; ++k) acc += (float)r[k];
    if (acc == 12345.678f) out[0] = acc;
}

template <int MODE>
static void run(const char *name, double instr_per_iter, double fma_cycles_per_iter, int warps_per_smsp, float *d, int sms, double clk_hz) {
    const int iters = 1 << 13;
    const int blocks = sms * (warps_per_smsp * 4 / 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    kern<MODE><<<blocks, 256>>>(d, 64, 1e-3f);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        kern<MODE><<<blocks, 256>>>(d, iters, 1e-3f);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double cycles = best * 1e-3 * clk_hz;                        // per SMSP
    const double warp_iters_per_smsp = (double)warps_per_smsp * iters;
    printf("%-44s warps/SMSP=%d  cycles/iter/warp-slot=%.2f  issue/cycle=%.3f  fma-pipe-util=%.3f\n", name, warps_per_smsp,
           cycles / warp_iters_per_smsp, instr_per_iter * warp_iters_per_smsp / cycles, fma_cycles_per_iter * warp_iters_per_smsp / cycles);
}

int main() {
    int sms = 148, khz = 1965000;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double clk = khz * 1e3;
    printf("SMs %d, clock %.0f MHz (nominal max; utilisation figures assume it)\n", sms, clk / 1e6);
    float *d; cudaMalloc(&d, 1024);
    for (int w : {6}) {
        run<0>("FFMA x16", 16, 16, w, d, sms, clk);
        run<1>("FFMA2 x16", 16, 32, w, d, sms, clk);
        run<2>("SHF x16", 16, 0, w, d, sms, clk);
        run<3>("FMNMX3(abs) x16", 16, 0, w, d, sms, clk);
        run<9>("5 FFMA2 (chain as in vote) x4", 20, 40, w, d, sms, clk);
        run<8>("5 FFMA2 + 2 SHF x4", 28, 40, w, d, sms, clk);
        run<4>("new vote mix 5 FFMA2+2 SHF+1 FMNMX3 x4", 32, 40, w, d, sms, clk);
        run<6>("two-threshold 6 FFMA2+4 SHF x4", 40, 48, w, d, sms, clk);
        run<5>("r01 mix 8 FFMA2+4 SHF x4", 48, 64, w, d, sms, clk);
        run<7>("scalar 10 FFMA+2 SHF+1 FMNMX3 x4", 52, 40, w, d, sms, clk);
        run<10>("FFMA2:FFMA 1:1 x8", 16, 24, w, d, sms, clk);
        run<11>("FFMA2:SHF 2:1 x8", 24, 32, w, d, sms, clk);
        run<12>("FFMA:SHF 1:1 x8", 16, 8, w, d, sms, clk);
        run<13>("FFMA:SHF 2:1 x8", 24, 16, w, d, sms, clk);
        run<14>("IMAD.HI x16", 16, 0, w, d, sms, clk);
        run<15>("IMAD.LO x16", 16, 0, w, d, sms, clk);
        run<16>("IADD x16", 16, 0, w, d, sms, clk);
        run<17>("LOP3 x16", 16, 0, w, d, sms, clk);
        run<18>("FADD x16", 16, 16, w, d, sms, clk);
        run<19>("FFMA.SAT x16", 16, 16, w, d, sms, clk);
        run<20>("POPC x16", 16, 0, w, d, sms, clk);
        run<21>("PRMT x16", 16, 0, w, d, sms, clk);
        run<22>("HFMA2 x16", 16, 16, w, d, sms, clk);
        run<23>("5 FFMA2+2 IMAD.HI+1 FMNMX3 x4", 32, 40, w, d, sms, clk);
        run<24>("4 FFMA2+2 FFMA.SAT+2 FADD+2 FFMA x4", 40, 56, w, d, sms, clk);
        run<25>("5 FFMA2+1 PRMT+1 IADD+1 FMNMX3 x4", 32, 40, w, d, sms, clk);
        printf("\n");
    }
    return 0;
}
