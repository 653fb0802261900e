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

// Backward of batchwise_get_RT (lib/gpu_tensor_funcs.py:204-235 with quats_2_rotation_matrix :306-326), thread per instance.
// Forward in closed form: qn = q/|q| = (a,b,c,d), R = M(a,b,c,d)^T, t = K^-1 (x zz, y zz, zz) with zz = z/1000,
// RT = [[R, -R t], [0 0 0 1]].
__global__ void __launch_bounds__(128) k_get_rt_backward(const float *__restrict__ q_in, const float *__restrict__ xy, const float *__restrict__ z_in,
                                                         const float *__restrict__ inv_k, const float *__restrict__ gR,
                                                         const float *__restrict__ gT, const float *__restrict__ gRT, int n,
                                                         float *__restrict__ d_q, float *__restrict__ d_xy, float *__restrict__ d_z) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double q0 = q_in[4 * i], q1 = q_in[4 * i + 1], q2 = q_in[4 * i + 2], q3 = q_in[4 * i + 3];
    const double nq = sqrt(q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3), sn = nq > 0.0 ? nq : 1.0;
    const double a = q0 / sn, b = q1 / sn, c = q2 / sn, d = q3 / sn;
    const double x = xy[2 * i], y = xy[2 * i + 1], zz = (double)z_in[i] / 1000.0;
    double k[9];
    for (int j = 0; j < 9; ++j) k[j] = inv_k[j];
    const double h[3] = {x * zz, y * zz, zz};
    double t[3];
    for (int r = 0; r < 3; ++r) t[r] = k[3 * r] * h[0] + k[3 * r + 1] * h[1] + k[3 * r + 2] * h[2];
    double m[3][3];
    m[0][0] = a * a - b * b - c * c + d * d; m[0][1] = 2 * (a * b + c * d); m[0][2] = 2 * (a * c - b * d);
    m[1][0] = 2 * (a * b - c * d); m[1][1] = -a * a + b * b - c * c + d * d; m[1][2] = 2 * (b * c + a * d);
    m[2][0] = 2 * (a * c + b * d); m[2][1] = 2 * (b * c - a * d); m[2][2] = -a * a - b * b + c * c + d * d;
    // upstream gradients gathered on R (= M^T) and t
    double GR[3][3], gcol[3], gt[3];
    for (int r = 0; r < 3; ++r) {
        gcol[r] = gRT ? (double)gRT[16 * i + 4 * r + 3] : 0.0;
        for (int cc = 0; cc < 3; ++cc)
            GR[r][cc] = (gR ? (double)gR[9 * i + 3 * r + cc] : 0.0) + (gRT ? (double)gRT[16 * i + 4 * r + cc] : 0.0) - gcol[r] * t[cc];
    }
    for (int cc = 0; cc < 3; ++cc) {
        double acc = gT ? (double)gT[3 * i + cc] : 0.0;
        for (int r = 0; r < 3; ++r) acc -= m[cc][r] * gcol[r];          // -(R^T gcol)_cc, R^T = M
        gt[cc] = acc;
    }
    double gh[3];
    for (int cc = 0; cc < 3; ++cc) gh[cc] = k[cc] * gt[0] + k[3 + cc] * gt[1] + k[6 + cc] * gt[2];   // K^-T gt
    d_xy[2 * i] = (float)(gh[0] * zz);
    d_xy[2 * i + 1] = (float)(gh[1] * zz);
    d_z[i] = (float)((gh[0] * x + gh[1] * y + gh[2]) / 1000.0);
    double G[3][3];                                                        // gradient on M: G = GR^T
    for (int r = 0; r < 3; ++r)
        for (int cc = 0; cc < 3; ++cc) G[r][cc] = GR[cc][r];
    const double ga = 2 * (a * G[0][0] + b * G[0][1] + c * G[0][2] + b * G[1][0] - a * G[1][1] + d * G[1][2] + c * G[2][0] - d * G[2][1] - a * G[2][2]);
    const double gb = 2 * (-b * G[0][0] + a * G[0][1] - d * G[0][2] + a * G[1][0] + b * G[1][1] + c * G[1][2] + d * G[2][0] + c * G[2][1] - b * G[2][2]);
    const double gc = 2 * (-c * G[0][0] + d * G[0][1] + a * G[0][2] - d * G[1][0] - c * G[1][1] + b * G[1][2] + a * G[2][0] + b * G[2][1] + c * G[2][2]);
    const double gd = 2 * (d * G[0][0] + c * G[0][1] - b * G[0][2] - c * G[1][0] + d * G[1][1] + a * G[1][2] + b * G[2][0] - a * G[2][1] + d * G[2][2]);
    if (nq > 0.0) {
        const double dot = a * ga + b * gb + c * gc + d * gd;
        d_q[4 * i] = (float)((ga - a * dot) / nq);
        d_q[4 * i + 1] = (float)((gb - b * dot) / nq);
        d_q[4 * i + 2] = (float)((gc - c * dot) / nq);
        d_q[4 * i + 3] = (float)((gd - d * dot) / nq);
    } else {
        d_q[4 * i] = (float)ga; d_q[4 * i + 1] = (float)gb; d_q[4 * i + 2] = (float)gc; d_q[4 * i + 3] = (float)gd;
    }
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

int fpc_get_rt_backward(const float *q, const float *xy, const float *z, const float *inv_k, const float *g_R, const float *g_T,
                        const float *g_RT, int n, float *d_q, float *d_xy, float *d_z, void *stream) {
    if (n < 0) return fail(FPC_EINVAL, "negative size");
    if (n == 0) return FPC_OK;
    if (!q || !xy || !z || !inv_k || !d_q || !d_xy || !d_z) return fail(FPC_EINVAL, "NULL pointer");
    k_get_rt_backward<<<ceil_div(n, 128), 128, 0, (cudaStream_t)stream>>>(q, xy, z, inv_k, g_R, g_T, g_RT, n, d_q, d_xy, d_z);
    FPC_LAUNCH_CHECK("k_get_rt_backward");
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

int fpc_pose_recover_xy_backward(const int32_t *labels, const uint8_t *cat_mask_u8, const float *xy_head, const int32_t *frame_of,
                                 const float *win_pts, const float *refined, const float *g_x, const int32_t *live, float inlier_thresh,
                                 int n, int b, int num_classes, int h, int w, int arith, float *d_xy_head, void *stream) {
    if (n < 0 || b <= 0 || h <= 0 || w <= 0 || num_classes < 2) return fail(FPC_EINVAL, "bad size");
    if (!labels || !cat_mask_u8 || !xy_head || !d_xy_head) return fail(FPC_EINVAL, "NULL pointer");
    if (arith != FPC_ARITH_IEEE && arith != FPC_ARITH_NVCC_FMA) return fail(FPC_EINVAL, "bad arith mode %d", arith);
    cudaStream_t st = (cudaStream_t)stream;
    FPC_CUDA_TRY(cudaMemsetAsync(d_xy_head, 0, (size_t)b * 2 * (num_classes - 1) * h * w * sizeof(float), st));
    if (n == 0) return FPC_OK;
    if (!frame_of || !win_pts || !refined || !g_x || !live) return fail(FPC_EINVAL, "NULL pointer");
    return launch_vote_refine_backward_labels(labels, cat_mask_u8, xy_head, frame_of, win_pts, refined, g_x, live, inlier_thresh, n,
                                              num_classes - 1, h, w, arith, d_xy_head, st);
}

}  // extern "C"
