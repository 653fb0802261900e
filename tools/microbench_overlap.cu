// Scratch micro-benchmark: can an FP32-issue-bound kernel (the vote kernel's instruction mix) and an HBM-bound streaming
// kernel share the SMs of a B200 productively?  Runs each alone, then both at once on two streams, for several residencies.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_overlap tools/microbench_overlap.cu
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))
#define SHF(d, a, c) asm volatile("shf.l.wrap.b32 %0, %1, %2, 1;" : "=r"(d) : "r"(a), "r"(c))
#define MIN3A(d, a, b, c) asm volatile("{.reg .f32 t0, t1; abs.f32 t0, %2; abs.f32 t1, %3; min.f32 %0, %1, t0, t1;}" : "=f"(d) : "f"(a), "f"(b), "f"(c))
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(u64 v, float &a, float &b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }

__global__ void __launch_bounds__(256) mix(float *out, int iters, float seed) {
    float f[8]; u64 p[16]; unsigned r[4];
    for (int k = 0; k < 16; ++k) p[k] = pk(seed * (threadIdx.x + k + 1), seed * (threadIdx.x + 2 * k + 1));
    for (int k = 0; k < 8; ++k) f[k] = seed * k;
    for (int k = 0; k < 4; ++k) r[k] = threadIdx.x + k;
    const float m1 = 0.999f + threadIdx.x * 1e-9f, c1 = 1e-4f + threadIdx.x * 1e-11f;
    const u64 m2 = pk(m1, m1 + 2e-3f), c2 = pk(c1, 2.f * c1);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            u64 t, U, W, s;
            FMA2(t, p[k], m2, p[8 + k]); FMA2(U, p[k + 4], c2, t); FMA2(t, p[k], c2, p[12 + k]); FMA2(W, p[k + 4], m2, t); FMA2(s, U, m2, W);
            float sa, sb; upk(s, sa, sb);
            SHF(r[k], __float_as_uint(sa), r[k]); SHF(r[k], __float_as_uint(sb), r[k]);
            MIN3A(f[k], f[k], sa, sb);
            p[k] = s;
        }
    }
    float acc = 0.f;
    for (int k = 0; k < 16; ++k) { float a, b; upk(p[k], a, b); acc += a + b; }
    for (int k = 0; k < 4; ++k) acc += f[k] + (float)r[k];
    if (acc == 12345.678f) out[0] = acc;
}
// streaming read: UNROLL independent 16-byte loads in flight per thread, grid-stride
template <int UNROLL>
__global__ void __launch_bounds__(256) stream_read(const float4 *__restrict__ src, size_t n, float *out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (UNROLL - 1) * stride < n; i += UNROLL * stride) {
        float4 v[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main() {
    int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const size_t bytes = 2ull << 30, n = bytes / 16;
    float4 *src; cudaMalloc(&src, bytes); cudaMemset(src, 0, bytes);
    float *out; cudaMalloc(&out, 1024);
    cudaStream_t s1, s2; cudaStreamCreate(&s1); cudaStreamCreate(&s2);
    cudaEvent_t a0, a1, b0, b1; cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
    const int iters = 40000;
    printf("mix = 5 FFMA2 + 2 SHF + 1 FMNMX3 per 2 votes (%d iterations); read = streaming 16-byte loads over %.1f GB\n", iters, bytes / 1e9);
    for (int mix_bps : {1, 2, 3}) for (int rd_bps : {1, 2, 4}) for (int unroll : {4, 8}) {
        auto launch_read = [&](cudaStream_t st) {
            if (unroll == 4) stream_read<4><<<sms * rd_bps, 256, 0, st>>>(src, n, out);
            else stream_read<8><<<sms * rd_bps, 256, 0, st>>>(src, n, out);
        };
        float t_mix, t_rd, t_mix2, t_rd2;
        mix<<<sms * mix_bps, 256, 0, s1>>>(out, 100, 1e-3f); launch_read(s2); cudaDeviceSynchronize();
        cudaEventRecord(a0, s1); mix<<<sms * mix_bps, 256, 0, s1>>>(out, iters, 1e-3f); cudaEventRecord(a1, s1); cudaDeviceSynchronize();
        cudaEventElapsedTime(&t_mix, a0, a1);
        cudaEventRecord(b0, s2); launch_read(s2); cudaEventRecord(b1, s2); cudaDeviceSynchronize();
        cudaEventElapsedTime(&t_rd, b0, b1);
        // both: the read kernel is repeated so that it covers the mix kernel's duration
        const int reps = (int)(t_mix / t_rd) + 1;
        cudaEventRecord(a0, s1); mix<<<sms * mix_bps, 256, 0, s1>>>(out, iters, 1e-3f); cudaEventRecord(a1, s1);
        cudaEventRecord(b0, s2); for (int r = 0; r < reps; ++r) launch_read(s2); cudaEventRecord(b1, s2);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&t_mix2, a0, a1); cudaEventElapsedTime(&t_rd2, b0, b1);
        printf("mix %d blk/SM, read %d blk/SM x %d loads in flight: alone mix %.2f ms, read %.0f GB/s | together mix %.2f ms (x%.2f), read %.0f GB/s (x%.2f)\n",
               mix_bps, rd_bps, unroll, t_mix, bytes / t_rd / 1e6, t_mix2, t_mix2 / t_mix, bytes * reps / t_rd2 / 1e6, (t_rd2 / reps) / t_rd);
    }
    return 0;
}
