// Scratch micro-benchmark: what does HBM3e deliver for k_gather's ADDRESS STREAM, with nothing but the loads?
// cfg2: 32 frames of 640x480, 18 discs of radius 35 per frame (3853 px each), the predicted class's 10 channel planes (4 q + 3 scales
// + 2 xy + 1 z out of 60) read on foreground pixels only: ~71 raster runs per disc, 1..71 pixels (avg 54 = 217 bytes) per plane.
// Variants (all warp-per-run, grid-stride, sums only, no stores):
//   lane      lane = pixel, 10 scalar loads per 32-pixel iteration (the shape of k_gather without its class-byte dependency)
//   lane_cls  the same behind a dependent 1-byte class load per pixel (k_gather as it is)
//   all       every load of the run issued before the first use (<= 3 x 10 scalar loads in flight per lane)
//   quad      lane = aligned 16-byte quad of the run, 10 LDG.128 in flight per lane
//   bulk      10 cp.async.bulk per run (16-byte aligned superset) into a shared-memory ring, RING runs in flight per warp
//   dense     the same number of bytes read as one dense stream (the HBM rate for reference)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/microbench_gather tools/microbench_gather.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cmath>
#include <vector>
#include <algorithm>

struct Run { int img, y, x0, len, cls; };
constexpr int B = 32, H = 480, W = 640, K = 6, HW = H * W;
constexpr unsigned FULL = 0xffffffffu;

struct Src { const float *q, *s, *v, *z; const uint8_t *cls; };
__device__ __forceinline__ void plane_bases(const Src &S, const Run &r, const float *(&pl)[10]) {
    const size_t pix = (size_t)r.y * W + r.x0, c = r.cls - 1;
    const float *q = S.q + ((size_t)r.img * 4 * K + 4 * c) * HW + pix, *s = S.s + ((size_t)r.img * 3 * K + 3 * c) * HW + pix;
    const float *v = S.v + ((size_t)r.img * 2 * K + 2 * c) * HW + pix, *z = S.z + ((size_t)r.img * K + c) * HW + pix;
    pl[0] = q; pl[1] = q + HW; pl[2] = q + 2 * HW; pl[3] = q + 3 * HW; pl[4] = s; pl[5] = s + HW; pl[6] = s + 2 * HW;
    pl[7] = v; pl[8] = v + HW; pl[9] = z;
}

template <bool CLS>
__global__ void __launch_bounds__(256) g_lane(Src S, const Run *runs, int n, float *out) {
    const int lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw) {
        Run rn = runs[r];
        for (int kb = 0; kb < rn.len; kb += 32) {
            const int kx = kb + lane;
            if (kx < rn.len) {
                if (CLS) rn.cls = S.cls[(size_t)rn.img * HW + (size_t)rn.y * W + rn.x0 + kx];
                const float *pl[10];
                plane_bases(S, rn, pl);
#pragma unroll
                for (int c = 0; c < 10; ++c) acc += __ldcs(pl[c] + kx);
            }
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void __launch_bounds__(256) g_all(Src S, const Run *runs, int n, float *out) {
    const int lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw) {
        const Run rn = runs[r];
        const float *pl[10];
        plane_bases(S, rn, pl);
        float v[3][10];
#pragma unroll
        for (int it = 0; it < 3; ++it)
#pragma unroll
            for (int c = 0; c < 10; ++c) v[it][c] = (it * 32 + lane < rn.len) ? __ldcs(pl[c] + it * 32 + lane) : 0.f;
#pragma unroll
        for (int it = 0; it < 3; ++it)
#pragma unroll
            for (int c = 0; c < 10; ++c) acc += v[it][c];
    }
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void __launch_bounds__(256) g_quad(Src S, const Run *runs, int n, float *out) {
    const int lane = threadIdx.x & 31, nw = (gridDim.x * blockDim.x) >> 5;
    float acc = 0.f;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nw) {
        Run rn = runs[r];
        const int lead = rn.x0 & 3;                 // W % 4 == 0: plane rows are 16-byte aligned
        rn.x0 -= lead;
        const int nq = (lead + rn.len + 3) >> 2;    // <= 19 quads for len <= 71
        const float *pl[10];
        plane_bases(S, rn, pl);
        if (lane < nq) {
            float4 v[10];
#pragma unroll
            for (int c = 0; c < 10; ++c) v[c] = __ldcs(reinterpret_cast<const float4 *>(pl[c]) + lane);
#pragma unroll
            for (int c = 0; c < 10; ++c) acc += v[c].x + v[c].y + v[c].z + v[c].w;
        }
    }
    if (acc == 12345.678f) out[0] = acc;
}

// ---- bulk-copy fed: per warp a ring of RING slots of 10 x 320 bytes; lane c < 10 issues the copy of plane c ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
constexpr int SEG = 320;   // bytes per plane segment slot (>= 16-byte aligned superset of 71 floats = 304)
template <int RING>
__global__ void __launch_bounds__(256) g_bulk(Src S, const Run *runs, int n, float *out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ unsigned long long bars[8 * RING];
    const int lane = threadIdx.x & 31, wv = threadIdx.x >> 5, nw = (gridDim.x * blockDim.x) >> 5;
    unsigned char *ring = smem + (size_t)wv * RING * 10 * SEG;
    unsigned long long *bar = bars + wv * RING;
    if (lane == 0)
        for (int k = 0; k < RING; ++k) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar[k])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    const int w0 = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    auto issue = [&](int r, int slot) {
        Run rn = runs[r];
        const int lead = rn.x0 & 3;
        rn.x0 -= lead;
        const uint32_t bytes = (uint32_t)((lead + rn.len + 3) >> 2) * 16u;
        const float *pl[10];
        plane_bases(S, rn, pl);
        if (lane == 0)
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[slot])), "r"(bytes * 10u) : "memory");
        __syncwarp();
        const float *src = pl[0];
#pragma unroll
        for (int c = 1; c < 10; ++c) if (lane == c) src = pl[c];
        if (lane < 10)
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                             smem_u32(ring + ((size_t)slot * 10 + lane) * SEG)), "l"(src), "r"(bytes), "r"(smem_u32(&bar[slot])) : "memory");
        return (int)bytes;
    };
    float acc = 0.f;
    int nbytes[RING];
    int head = 0;
    for (int k = 0; k < RING; ++k) nbytes[k] = 0;
#pragma unroll
    for (int k = 0; k < RING - 1; ++k) if (w0 + k * nw < n) nbytes[k] = issue(w0 + k * nw, k);
    uint32_t phase = 0;   // bit k = parity slot k waits for next
    for (int r = w0; r < n; r += nw, ++head) {
        const int slot = head % RING, ahead = r + (RING - 1) * nw, aslot = (head + RING - 1) % RING;
        int ab = 0;
        if (ahead < n) ab = issue(ahead, aslot);
#pragma unroll
        for (int k = 0; k < RING; ++k) if (k == aslot) nbytes[k] = ab;
        uint32_t ok;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(smem_u32(&bar[slot])), "r"((phase >> slot) & 1u) : "memory");
        } while (!ok);
        phase ^= 1u << slot;
        int nb = 0;
#pragma unroll
        for (int k = 0; k < RING; ++k) if (k == slot) nb = nbytes[k];
        const int nq = nb >> 4;
        if (lane < nq) {
#pragma unroll
            for (int c = 0; c < 10; ++c) {
                const float4 v = *reinterpret_cast<const float4 *>(ring + ((size_t)slot * 10 + c) * SEG + lane * 16);
                acc += v.x + v.y + v.z + v.w;
            }
        }
        __syncwarp();
    }
    if (acc == 12345.678f) out[0] = acc;
}
__global__ void __launch_bounds__(256) g_dense(const float4 *__restrict__ src, size_t n, float *out) {
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + 7 * stride < n; i += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = __ldcs(src + i + u * stride);
#pragma unroll
        for (int u = 0; u < 8; ++u) acc += v[u].x + v[u].y + v[u].z + v[u].w;
    }
    if (acc == 12345.678f) out[0] = acc;
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    std::vector<Run> runs;
    std::vector<uint8_t> cls((size_t)B * HW, 0);
    long long fg = 0;
    for (int b = 0; b < B; ++b)
        for (int d = 0; d < 18; ++d) {
            const int cx = 53 + 106 * (d % 6) + (b % 7), cy = 80 + 160 * (d / 6) + (b % 5), c = d % K + 1;
            for (int dy = -35; dy <= 35; ++dy) {
                const int half = (int)floor(sqrt(35.0 * 35.0 - (double)dy * dy));
                runs.push_back({b, cy + dy, cx - half, 2 * half + 1, c});
                for (int x = cx - half; x <= cx + half; ++x) cls[(size_t)b * HW + (size_t)(cy + dy) * W + x] = (uint8_t)c;
                fg += 2 * half + 1;
            }
        }
    // raster order per image, like the run list of the path
    std::stable_sort(runs.begin(), runs.end(), [](const Run &a, const Run &b) {
        return a.img != b.img ? a.img < b.img : a.y != b.y ? a.y < b.y : a.x0 < b.x0; });
    const int n = (int)runs.size();
    const double bytes = 40.0 * fg;
    float *q, *s, *v, *z, *out;
    uint8_t *dcls;
    Run *druns;
    cudaMalloc(&q, (size_t)B * 4 * K * HW * 4); cudaMalloc(&s, (size_t)B * 3 * K * HW * 4);
    cudaMalloc(&v, (size_t)B * 2 * K * HW * 4); cudaMalloc(&z, (size_t)B * K * HW * 4);
    cudaMemset(q, 0, (size_t)B * 4 * K * HW * 4); cudaMemset(s, 0, (size_t)B * 3 * K * HW * 4);
    cudaMemset(v, 0, (size_t)B * 2 * K * HW * 4); cudaMemset(z, 0, (size_t)B * K * HW * 4);
    cudaMalloc(&dcls, cls.size()); cudaMemcpy(dcls, cls.data(), cls.size(), cudaMemcpyHostToDevice);
    cudaMalloc(&druns, n * sizeof(Run)); cudaMemcpy(druns, runs.data(), n * sizeof(Run), cudaMemcpyHostToDevice);
    cudaMalloc(&out, 1024);
    const Src S{q, s, v, z, dcls};
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("%d runs, %lld foreground px, %.1f MB algorithmic (40 B/px), %d SMs\n", n, fg, bytes / 1e6, sms);
    cudaFuncSetAttribute(g_bulk<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4 * 10 * SEG);
    cudaFuncSetAttribute(g_bulk<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 8 * 10 * SEG);
    auto timeit = [&](const char *name, int bps, auto launch) {
        float best = 1e9f;
        for (int rep = 0; rep < 6; ++rep) {
            // flush L2 (126 MB) with a 512 MB memset of a plane that the next launch does not need first
            cudaMemsetAsync(q, 0, 512u << 20);
            cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (rep >= 1) best = std::min(best, ms);
        }
        const cudaError_t err = cudaGetLastError();
        printf("%-10s %2d blk/SM: %7.1f us  %6.0f GB/s on algorithmic bytes%s\n", name, bps, best * 1e3, bytes / best / 1e6,
               err == cudaSuccess ? "" : cudaGetErrorString(err));
    };
    for (int bps : {2, 4, 6, 8}) {
        timeit("lane", bps, [&] { g_lane<false><<<sms * bps, 256>>>(S, druns, n, out); });
        timeit("lane_cls", bps, [&] { g_lane<true><<<sms * bps, 256>>>(S, druns, n, out); });
        timeit("all", bps, [&] { g_all<<<sms * bps, 256>>>(S, druns, n, out); });
        timeit("quad", bps, [&] { g_quad<<<sms * bps, 256>>>(S, druns, n, out); });
    }
    for (int bps : {1, 2, 4}) timeit("bulk r4", bps, [&] { g_bulk<4><<<sms * bps, 256, 8 * 4 * 10 * SEG>>>(S, druns, n, out); });
    for (int bps : {1, 2}) timeit("bulk r8", bps, [&] { g_bulk<8><<<sms * bps, 256, 8 * 8 * 10 * SEG>>>(S, druns, n, out); });
    const size_t n16 = (size_t)(bytes / 16);
    for (int bps : {4, 8}) timeit("dense", bps, [&] { g_dense<<<sms * bps, 256>>>(reinterpret_cast<const float4 *>(s), n16, out); });
    return 0;
}
