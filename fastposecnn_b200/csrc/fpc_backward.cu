// Training support (second half): the backward of the two differentiable steps at the head of
// the path, so that the reference's aggregated losses (lib/loss.py:155-545: QLoss, ScalesLoss, ZLoss, ... on matched
// AggData) can train through the drop-ins the way they train through the reference's torch ops.
//   * class_compress  (lib/gpu_tensor_funcs.py:52-99): per pixel the predicted class's channels, q / xy L2-normalised;
//   * AggregationLayer (lib/aggregation_layer.py:125-156): masked means per instance (exp for z, normalise for q), masked xy.
// Both are per-pixel scatters: HBM-bound, every gradient element written exactly once.
#include "fpc_internal.cuh"

namespace fpc {
namespace {

// d(v / |v|) applied to g:  (g - v_hat (v_hat . g)) / |v|;  the zero-norm guard of normalize() divides by 1 instead.
template <int D>
__device__ __forceinline__ void normalize_backward(const float *v, const float *g, float *out) {
    float n2 = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) n2 += v[k] * v[k];
    const float n = sqrtf(n2);
    if (n == 0.f) {
#pragma unroll
        for (int k = 0; k < D; ++k) out[k] = g[k];
        return;
    }
    const float inv = 1.f / n;
    float dot = 0.f;
#pragma unroll
    for (int k = 0; k < D; ++k) dot += v[k] * inv * g[k];
#pragma unroll
    for (int k = 0; k < D; ++k) out[k] = (g[k] - v[k] * inv * dot) * inv;
}

// Gradients of the class-compressed fields -> gradients of the raw head maps (pre-zeroed): only the predicted class's
// channels of foreground pixels receive anything.
__global__ void __launch_bounds__(256) k_class_compress_backward(const long long *__restrict__ cat, const float *__restrict__ quat,
                                                                 const float *__restrict__ xy, const float *__restrict__ g_q,
                                                                 const float *__restrict__ g_s, const float *__restrict__ g_xy,
                                                                 const float *__restrict__ g_z, float *__restrict__ d_quat,
                                                                 float *__restrict__ d_scales, float *__restrict__ d_xy,
                                                                 float *__restrict__ d_z, int C, int hw, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const long long c = cat[p];
    if (c <= 0 || c >= C) return;
    const int bi = p / hw, k = (int)c - 1, K = C - 1;
    const size_t HW = (size_t)hw, pix = (size_t)(p - bi * hw);
    if (g_q) {
        const float *src = quat + ((size_t)bi * 4 * K + 4 * k) * HW + pix;
        const float *gs = g_q + (size_t)bi * 4 * HW + pix;
        float v[4], g[4], o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { v[j] = src[j * HW]; g[j] = gs[j * HW]; }
        normalize_backward<4>(v, g, o);
        float *dst = d_quat + ((size_t)bi * 4 * K + 4 * k) * HW + pix;
#pragma unroll
        for (int j = 0; j < 4; ++j) dst[j * HW] = o[j];
    }
    if (g_xy) {
        const float *src = xy + ((size_t)bi * 2 * K + 2 * k) * HW + pix;
        const float *gs = g_xy + (size_t)bi * 2 * HW + pix;
        float v[2] = {src[0], src[HW]}, g[2] = {gs[0], gs[HW]}, o[2];
        normalize_backward<2>(v, g, o);
        float *dst = d_xy + ((size_t)bi * 2 * K + 2 * k) * HW + pix;
        dst[0] = o[0];
        dst[HW] = o[1];
    }
    if (g_s) {
        const float *gs = g_s + (size_t)bi * 3 * HW + pix;
        float *dst = d_scales + ((size_t)bi * 3 * K + 3 * k) * HW + pix;
#pragma unroll
        for (int j = 0; j < 3; ++j) dst[j * HW] = gs[j * HW];
    }
    if (g_z) d_z[((size_t)bi * K + k) * HW + pix] = g_z[(size_t)bi * HW + pix];
}

// Per-instance gradient vectors G [n,8] (already divided by the pixel count and pushed through exp / normalise by the
// caller: 4 for q, 3 for scales, 1 for z) and the dense gradient of the masked xy output -> gradients of the
// class-compressed fields.  Every element is written (zeros on background), so no memset.
__global__ void __launch_bounds__(256) k_aggregate_backward(const int *__restrict__ labels, const float *__restrict__ G,
                                                            const float *__restrict__ g_xy_dense, int n, float *__restrict__ d_q,
                                                            float *__restrict__ d_s, float *__restrict__ d_xy, float *__restrict__ d_z,
                                                            int hw, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int bi = p / hw;
    const size_t HW = (size_t)hw, pix = (size_t)(p - bi * hw);
    const int i = labels[p] - 1;
    float g[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float gx = 0.f, gy = 0.f;
    if (i >= 0 && i < n) {
        const float4 a = __ldg(reinterpret_cast<const float4 *>(G + (size_t)i * 8));
        const float4 b = __ldg(reinterpret_cast<const float4 *>(G + (size_t)i * 8 + 4));
        g[0] = a.x; g[1] = a.y; g[2] = a.z; g[3] = a.w; g[4] = b.x; g[5] = b.y; g[6] = b.z; g[7] = b.w;
        if (g_xy_dense) {
            gx = __ldcs(g_xy_dense + ((size_t)i * 2) * HW + pix);
            gy = __ldcs(g_xy_dense + ((size_t)i * 2 + 1) * HW + pix);
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) d_q[((size_t)bi * 4 + j) * HW + pix] = g[j];
#pragma unroll
    for (int j = 0; j < 3; ++j) d_s[((size_t)bi * 3 + j) * HW + pix] = g[4 + j];
    d_z[(size_t)bi * HW + pix] = g[7];
    d_xy[((size_t)bi * 2) * HW + pix] = gx;
    d_xy[((size_t)bi * 2 + 1) * HW + pix] = gy;
}

}  // namespace
}  // namespace fpc

using namespace fpc;

extern "C" {

int fpc_class_compress_backward(const int64_t *cat_mask, const float *quaternion, const float *xy, const float *g_q, const float *g_s,
                                const float *g_xy, const float *g_z, float *d_quaternion, float *d_scales, float *d_xy, float *d_z,
                                int b, int num_classes, int h, int w, void *stream) {
    if (b < 0 || h < 0 || w < 0) return fail(FPC_EINVAL, "negative size");
    if (num_classes < 2 || num_classes > 255) return fail(FPC_EINVAL, "num_classes must be in [2,255]");
    const long long P = (long long)b * h * w;
    if (P == 0) return FPC_OK;
    if (P >= (1ll << 31)) return fail(FPC_EINVAL, "b*h*w must be < 2^31");
    if (!cat_mask || !d_quaternion || !d_scales || !d_xy || !d_z) return fail(FPC_EINVAL, "NULL pointer");
    if ((g_q && !quaternion) || (g_xy && !xy)) return fail(FPC_EINVAL, "the raw head map is needed to differentiate its normalisation");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t K = (size_t)num_classes - 1, HW = (size_t)h * w;
    FPC_CUDA_TRY(cudaMemsetAsync(d_quaternion, 0, (size_t)b * 4 * K * HW * sizeof(float), st));
    FPC_CUDA_TRY(cudaMemsetAsync(d_scales, 0, (size_t)b * 3 * K * HW * sizeof(float), st));
    FPC_CUDA_TRY(cudaMemsetAsync(d_xy, 0, (size_t)b * 2 * K * HW * sizeof(float), st));
    FPC_CUDA_TRY(cudaMemsetAsync(d_z, 0, (size_t)b * K * HW * sizeof(float), st));
    k_class_compress_backward<<<ceil_div(P, 256), 256, 0, st>>>(reinterpret_cast<const long long *>(cat_mask), quaternion, xy, g_q, g_s,
                                                                g_xy, g_z, d_quaternion, d_scales, d_xy, d_z, num_classes, h * w, (int)P);
    FPC_LAUNCH_CHECK("k_class_compress_backward");
    return FPC_OK;
}

int fpc_aggregate_backward(const int32_t *labels, const float *inst_grads, const float *g_xy_dense, int n, float *d_q, float *d_s,
                           float *d_xy, float *d_z, int b, int h, int w, void *stream) {
    if (b < 0 || h < 0 || w < 0 || n < 0) return fail(FPC_EINVAL, "negative size");
    const long long P = (long long)b * h * w;
    if (P == 0) return FPC_OK;
    if (P >= (1ll << 31)) return fail(FPC_EINVAL, "b*h*w must be < 2^31");
    if (!labels || (n > 0 && !inst_grads) || !d_q || !d_s || !d_xy || !d_z) return fail(FPC_EINVAL, "NULL pointer");
    k_aggregate_backward<<<ceil_div(P, 256), 256, 0, (cudaStream_t)stream>>>(labels, inst_grads, g_xy_dense, n, d_q, d_s, d_xy, d_z,
                                                                           h * w, (int)P);
    FPC_LAUNCH_CHECK("k_aggregate_backward");
    return FPC_OK;
}

int fpc_vote_refine_backward(const float *fmask, const float *vertex, long long sN, long long sH, long long sW, long long s2,
                             const float *win_pts, const float *refined, const float *g_x, const int32_t *live, float inlier_thresh,
                             int n, int h, int w, int arith, float *d_vertex, void *stream) {
    if (n < 0 || h <= 0 || w <= 0) return fail(FPC_EINVAL, "bad size");
    if (n == 0) return FPC_OK;
    if (!fmask || !vertex || !win_pts || !refined || !g_x || !live || !d_vertex) return fail(FPC_EINVAL, "NULL pointer");
    if (arith != FPC_ARITH_IEEE && arith != FPC_ARITH_NVCC_FMA) return fail(FPC_EINVAL, "bad arith mode %d", arith);
    return launch_vote_refine_backward(fmask, vertex, sN, sH, sW, s2, win_pts, refined, g_x, live, inlier_thresh, n, h, w, arith,
                                       d_vertex, (cudaStream_t)stream);
}

}  // extern "C"
