// Shared device/host helpers of libfpc_b200.so (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <climits>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "fpc_b200.h"

namespace fpc {

// ---- error plumbing (thread-local message, negative return codes; never exit) ----------
char *err_buf();
int fail(int code, const char *fmt, ...);

#define FPC_CUDA_TRY(expr)                                                                   \
    do {                                                                                     \
        cudaError_t e_ = (expr);                                                             \
        if (e_ != cudaSuccess) return ::fpc::fail(FPC_ECUDA, "%s: %s", #expr, cudaGetErrorString(e_)); \
    } while (0)

// Optional per-kernel timing: the caller may hand fpc_pose_recover an array of cudaEvent_t; event 0 is
// recorded before the first kernel and event k after the k-th launch (thread-local cursor).
void stage_begin(void **events, int n, cudaStream_t st);
void stage_stamps(unsigned long long *stamps);   // call before stage_begin: device slots for %globaltimer stamps (graph-capturable)
void stage_mark();
void stage_end();

#define FPC_LAUNCH_CHECK(name)                                                               \
    do {                                                                                     \
        cudaError_t e_ = cudaGetLastError();                                                 \
        if (e_ != cudaSuccess) return ::fpc::fail(FPC_ECUDA, "launch of %s: %s", name, cudaGetErrorString(e_)); \
        ::fpc::stage_mark();                                                                 \
    } while (0)

constexpr unsigned FULL = 0xffffffffu;

__host__ __device__ inline long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }
inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (cached); grids of persistent kernels are multiples of it.
int sm_count();

// ---- nn.UpsamplingBilinear2d (align_corners=True), the x S step of smp's SegmentationHead ------------------
// Source coordinate and weights exactly as ATen computes them (ATen/native/UpSample.h: area_pixel_compute_scale,
// area_pixel_compute_source_index, guard_index_and_lambda): scale = float(in-1) / float(out-1) (host side),
// src = scale * dst, i0 = min(floor(src), in-1), w1 = clamp(src - i0, 0, 1), w0 = 1 - w1, i1 = i0 + (i0 < in-1).
struct UpParams {
    int s;            // 0/1 = inputs are full resolution
    int hl, wl;       // low-resolution plane size
    float sy, sx;     // float(hl-1)/float(h-1), float(wl-1)/float(w-1)
};
struct LerpCoord {
    int i0, i1;
    float w0, w1;
};
__device__ __forceinline__ LerpCoord lerp_coord(int dst, float scale, int in_size) {
    const float src = __fmul_rn(scale, (float)dst);
    LerpCoord c;
    c.i0 = min((int)floorf(src), in_size - 1);
    c.i1 = c.i0 + (c.i0 < in_size - 1 ? 1 : 0);
    c.w1 = fminf(fmaxf(__fsub_rn(src, (float)c.i0), 0.f), 1.f);
    c.w0 = __fsub_rn(1.f, c.w1);
    return c;
}
// Horizontal pair first, then the two rows; each "a*wa + b*wb" is fma(a, wa, b*wb) -- the contraction both ATen's
// vectorised CPU kernel and nvcc's build of the CUDA kernel end up with (pinned in tests/test_head_epilogue_*.py).
__device__ __forceinline__ float bilerp(float v00, float v01, float v10, float v11, float wx0, float wx1, float wy0, float wy1) {
    const float t0 = __fmaf_rn(v00, wx0, __fmul_rn(v01, wx1));
    const float t1 = __fmaf_rn(v10, wx0, __fmul_rn(v11, wx1));
    return __fmaf_rn(t0, wy0, __fmul_rn(t1, wy1));
}

// ---- reference arithmetic (lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu) -----
// a*b + c*d in the two arithmetic modes of fpc_b200.h.
template <int ARITH>
__device__ __forceinline__ float sum_prod(float a, float b, float c, float d) {
    if (ARITH == FPC_ARITH_IEEE) return __fadd_rn(__fmul_rn(a, b), __fmul_rn(c, d));
    return __fmaf_rn(a, b, __fmul_rn(c, d));
}

// The reference compares float magnitudes against the DOUBLE literal 1e-6 (.cu:42-43,119).
// (float)1e-6 = 0x358637BD lies just below 1e-6, so for any float x:
//   (double)x < 1e-6   <=>   x <= 1e-6f
__device__ __forceinline__ bool below_1e6(float x) { return x <= 1e-6f; }

// ransac_voting_kernel.cu:112-125 -- the cosine test of one (hypothesis, pixel) pair.
template <int ARITH>
__device__ __forceinline__ bool vote_exact(float cx, float cy, float nx, float ny, float hx, float hy, float thresh) {
    float dx = __fsub_rn(hx, cx);
    float dy = __fsub_rn(hy, cy);
    float norm1 = __fsqrt_rn(sum_prod<ARITH>(nx, nx, ny, ny));
    float norm2 = __fsqrt_rn(sum_prod<ARITH>(dx, dx, dy, dy));
    if (below_1e6(norm1) || below_1e6(norm2)) return false;
    float ang = __fdiv_rn(sum_prod<ARITH>(dx, nx, dy, ny), __fmul_rn(norm1, norm2));
    return ang > thresh;
}

// ransac_voting_kernel.cu:30-48 -- intersection of the two lines through (c0, dir d0) and
// (c1, dir d1).  Returns false (hypothesis stays (0,0)) for a degenerate pair.
template <int ARITH>
__device__ __forceinline__ bool hypothesis_exact(float d0x, float d0y, float c0x, float c0y, float d1x, float d1y,
                                                 float c1x, float c1y, float &x, float &y) {
    float nx0 = d0y, ny0 = -d0x, nx1 = d1y, ny1 = -d1x;
    float det_y = __fsub_rn(__fmul_rn(nx1, ny0), __fmul_rn(nx0, ny1));
    float det_x = __fsub_rn(__fmul_rn(ny1, nx0), __fmul_rn(ny0, nx1));
    if (below_1e6(fabsf(det_y)) || below_1e6(fabsf(det_x))) return false;
    float p0 = sum_prod<ARITH>(nx0, c0x, ny0, c0y);
    float p1 = sum_prod<ARITH>(nx1, c1x, ny1, c1y);
    float num_y, num_x;
    if (ARITH == FPC_ARITH_IEEE) {
        num_y = __fsub_rn(__fmul_rn(nx1, p0), __fmul_rn(nx0, p1));
        num_x = __fsub_rn(__fmul_rn(ny1, p0), __fmul_rn(ny0, p1));
    } else {
        // contraction pattern of nvcc 12.9 / sm_100a for the same source (see oracle/ransac_voting_ref.c)
        num_y = __fmaf_rn(nx1, p0, -__fmul_rn(nx0, p1));
        num_x = __fmaf_rn(-ny0, p1, __fmul_rn(ny1, p0));
    }
    y = __fdiv_rn(num_y, det_y);
    x = __fdiv_rn(num_x, det_x);
    return true;
}

// ---- gtf.normalize (lib/gpu_tensor_funcs.py:37-50) as a GPU torch build computes it -------------------------------
// torch.norm(dim) on CUDA accumulates  acc + a*a  with every product and every sum rounded separately (no FMA contraction),
// in index order, then takes an IEEE square root; the division is IEEE as well (measured on the B200:
// tests/test_dropin_gpu.py::test_direction_normalisation_equals_torch_cuda_bit_for_bit).  These two helpers ARE that formula.
__device__ __forceinline__ float torch_norm2(float x, float y) { return __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y))); }
__device__ __forceinline__ float torch_norm4(float a, float b, float c, float d) {
    return __fsqrt_rn(__fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(a, a), __fmul_rn(b, b)), __fmul_rn(c, c)), __fmul_rn(d, d)));
}

// ---- small device utilities --------------------------------------------------------------
__device__ __forceinline__ uint32_t mix32(uint32_t x) {  // murmur3 finaliser
    x ^= x >> 16; x *= 0x85ebca6bu; x ^= x >> 13; x *= 0xc2b2ae35u; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t hash3(uint64_t seed, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t h = mix32((uint32_t)seed ^ 0x9e3779b9u);
    h = mix32(h ^ (uint32_t)(seed >> 32));
    h = mix32(h ^ a);
    h = mix32(h + 0x7f4a7c15u + b);
    h = mix32(h ^ (c * 0x27d4eb2fu));
    return h;
}

__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int o = __shfl_up_sync(FULL, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// Exclusive scan of data[0..n) in place by ONE thread block (blockDim a multiple of 32, <= 1024).
// Each thread owns a contiguous chunk (serial scan), the chunk totals are scanned once across the
// block: two passes over the data and three barriers regardless of n.  Returns the total in every thread.
__device__ inline int block_exclusive_scan_inplace(int *data, int n) {
    __shared__ int s_warp[32];
    __shared__ int s_total;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const int per = (n + blockDim.x - 1) / blockDim.x;
    const int beg = min(n, (int)threadIdx.x * per), end = min(n, beg + per);
    int sum = 0;
    for (int i = beg; i < end; ++i) sum += data[i];
    const int inc = warp_incl_scan(sum, lane);
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        const int w = (lane < nw) ? s_warp[lane] : 0;
        const int winc = warp_incl_scan(w, lane);
        s_warp[lane] = winc - w;
        if (lane == 31) s_total = winc;
    }
    __syncthreads();
    int run = s_warp[wid] + inc - sum;
    for (int i = beg; i < end; ++i) {
        const int v = data[i];
        data[i] = run;
        run += v;
    }
    __syncthreads();
    return s_total;
}

}  // namespace fpc
