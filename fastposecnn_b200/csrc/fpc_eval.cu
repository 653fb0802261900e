// Evaluation maths on matched (ground truth, prediction) pairs -- SURVEY.md section 8f rank 4:
//   lib/gpu_tensor_funcs.py:434-455 get_raw_quat_distance, :457-476 + :752-799 get_symmetric_quat_distance (360 rotations
//   about y, composed in float64), :503-547 get_3d_ious (a Python loop with two 4x4 inversions and ~20 small kernels per
//   pair), :565-567 from_Ts_get_offset_error, :611-652 calculate_aps.
// One warp per pair: the 360 symmetry rotations are spread over the lanes (12 each, FP64 like the reference), lane 0 does
// the 3D IoU.  Nothing here is bandwidth- or FLOP-relevant (hundreds of pairs); the point is ONE launch instead of a
// Python loop with host syncs, on the same values.
#include <cmath>

#include "fpc_common.cuh"

namespace fpc {
namespace {

constexpr double RAD2DEG = 57.29577951308232;      // torch.rad2deg's constant (180 / pi)

// min(|a - b|, |a + b|) * 180/pi for two quaternions (the reference's "degree distance")
__device__ __forceinline__ double quat_distance(const double *a, const double *b) {
    double dm = 0.0, dp = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const double m = a[k] - b[k], p = a[k] + b[k];
        dm += m * m;
        dp += p * p;
    }
    return fmin(sqrt(dm), sqrt(dp)) * RAD2DEG;
}

// inverse of a general 4x4 (row-major float in, double out) by Gauss-Jordan with partial pivoting
__device__ void inverse4(const float *__restrict__ m, double inv[4][4]) {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) {
            a[r][c] = (double)m[4 * r + c];
            a[r][4 + c] = r == c ? 1.0 : 0.0;
        }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        for (int r = col + 1; r < 4; ++r)
            if (fabs(a[r][col]) > fabs(a[piv][col])) piv = r;
        if (piv != col)
            for (int c = 0; c < 8; ++c) { const double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        const double d = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= d;
        for (int r = 0; r < 4; ++r) {
            if (r == col) continue;
            const double f = a[r][col];
            for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) inv[r][c] = a[r][4 + c];
}

// per corner k of the scaled unit cube: max and min over the three world coordinates (the reference reduces over dim 0 of
// its [3,8] corner matrix, gpu_tensor_funcs.py:513-516 -- kept as written)
__device__ void corner_extents(const float *__restrict__ rt, const float *__restrict__ scales, double mx[8], double mn[8]) {
    double inv[4][4];
    inverse4(rt, inv);
    for (int k = 0; k < 8; ++k) {
        const double x = ((k & 2) ? -0.5 : 0.5) * (double)scales[0];      // corner order of get_3d_bbox (:359-368)
        const double y = ((k & 4) ? -0.5 : 0.5) * (double)scales[1];
        const double z = ((k & 1) ? -0.5 : 0.5) * (double)scales[2];
        double w[4];
        for (int r = 0; r < 4; ++r) w[r] = inv[r][0] * x + inv[r][1] * y + inv[r][2] * z + inv[r][3];
        const double cx = w[0] / w[3], cy = w[1] / w[3], cz = w[2] / w[3];
        mx[k] = fmax(cx, fmax(cy, cz));
        mn[k] = fmin(cx, fmin(cy, cz));
    }
}

__global__ void __launch_bounds__(128) k_pose_errors(const float *__restrict__ q0, const float *__restrict__ q1,
                                                     const long long *__restrict__ sym, const double *__restrict__ rot,
                                                     const float *__restrict__ rt0, const float *__restrict__ rt1,
                                                     const float *__restrict__ s0, const float *__restrict__ s1,
                                                     const float *__restrict__ t0, const float *__restrict__ t1, int m,
                                                     float *__restrict__ raw_deg, double *__restrict__ sym_deg,
                                                     float *__restrict__ iou3d, float *__restrict__ offset,
                                                     float *__restrict__ centers) {
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= m) return;
    if (q0 && q1) {
        double a[4], b[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) { a[k] = (double)q0[4 * i + k]; b[k] = (double)q1[4 * i + k]; }
        if (raw_deg && lane == 0) {
            // float32 like the reference's non-symmetric branch
            float dm = 0.f, dp = 0.f;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float mm = q0[4 * i + k] - q1[4 * i + k], pp = q0[4 * i + k] + q1[4 * i + k];
                dm += mm * mm;
                dp += pp * pp;
            }
            raw_deg[i] = fminf(sqrtf(dm), sqrtf(dp)) * 57.29577951308232f;
        }
        if (sym_deg) {
            double best = INFINITY;
            if (!sym || sym[i] != 0) {
                for (int r = lane; r < 360; r += 32) {
                    // q1 (x) (w, 0, y, 0), each product and sum rounded separately (torch evaluates it op by op)
                    const double bw = rot[4 * r], by = rot[4 * r + 2];
                    double o[4];
                    o[0] = __dsub_rn(__dmul_rn(b[0], bw), __dmul_rn(b[2], by));
                    o[1] = __dsub_rn(__dmul_rn(b[1], bw), __dmul_rn(b[3], by));
                    o[2] = __dadd_rn(__dmul_rn(b[0], by), __dmul_rn(b[2], bw));
                    o[3] = __dadd_rn(__dmul_rn(b[1], by), __dmul_rn(b[3], bw));
                    // normalize(): the norm is rounded to float32 before the division (gpu_tensor_funcs.py:43-48)
                    const double n = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
                    const double nf = n != 0.0 ? (double)(float)n : 1.0;
#pragma unroll
                    for (int k = 0; k < 4; ++k) o[k] /= nf;
                    best = fmin(best, quat_distance(a, o));
                }
#pragma unroll
                for (int s = 16; s > 0; s >>= 1) best = fmin(best, __shfl_xor_sync(FULL, best, s));
            } else {
                best = NAN;
            }
            if (lane == 0) sym_deg[i] = best;
        }
    }
    if (lane != 0) return;
    if (iou3d) {
        double mx1[8], mn1[8], mx2[8], mn2[8];
        corner_extents(rt0 + 16 * i, s0 + 3 * i, mx1, mn1);
        corner_extents(rt1 + 16 * i, s1 + 3 * i, mx2, mn2);
        double inter = 1.0, v1 = 1.0, v2 = 1.0, lo = INFINITY;
        for (int k = 0; k < 8; ++k) {
            const double e = fmin(mx1[k], mx2[k]) - fmax(mn1[k], mn2[k]);
            lo = fmin(lo, e);
            inter *= e;
            v1 *= mx1[k] - mn1[k];
            v2 *= mx2[k] - mn2[k];
        }
        if (lo < 0.0) inter = 0.0;
        iou3d[i] = (float)(inter / (v1 + v2 - inter));
    }
    if (centers) {
        // world position of the camera-frame origin under each RT: inverse(RT) @ (0,0,0,1), de-homogenised
        // (from_RTs_get_T_offset_errors, lib/gpu_tensor_funcs.py:569-609)
        double inv[4][4];
        inverse4(rt0 + 16 * i, inv);
        for (int k = 0; k < 3; ++k) centers[6 * i + k] = (float)(inv[k][3] / inv[3][3]);
        inverse4(rt1 + 16 * i, inv);
        for (int k = 0; k < 3; ++k) centers[6 * i + 3 + k] = (float)(inv[k][3] / inv[3][3]);
    }
    if (offset) {
        const float dx = t0[3 * i] - t1[3 * i], dy = t0[3 * i + 1] - t1[3 * i + 1], dz = t0[3 * i + 2] - t1[3 * i + 2];
        offset[i] = sqrtf(dx * dx + dy * dy + dz * dz) * 10.f;
    }
}

// calculate_aps: block per threshold; fraction of non-NaN values that pass `value OP threshold` (OP: 0 = less, 1 = greater)
__global__ void __launch_bounds__(256) k_threshold_fraction(const double *__restrict__ values, int n, const double *__restrict__ thresholds,
                                                            int op, float *__restrict__ out) {
    __shared__ int s_hit, s_valid;
    if (threadIdx.x == 0) { s_hit = 0; s_valid = 0; }
    __syncthreads();
    const double t = thresholds[blockIdx.x];
    int hit = 0, valid = 0;
    for (int k = threadIdx.x; k < n; k += blockDim.x) {
        const double v = values[k];
        if (v != v) continue;
        ++valid;
        hit += op == 0 ? (v < t) : (v > t);
    }
    hit = __reduce_add_sync(FULL, hit);
    valid = __reduce_add_sync(FULL, valid);
    if ((threadIdx.x & 31) == 0) { atomicAdd(&s_hit, hit); atomicAdd(&s_valid, valid); }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = __fdiv_rn((float)s_hit, (float)s_valid);      // torch: int64 / int -> float32
}

}  // namespace
}  // namespace fpc

using namespace fpc;

extern "C" {

int fpc_pose_errors(const float *q_gt, const float *q_pred, const int64_t *symmetric_ids, const double *sym_rotations,
                    const float *rt_gt, const float *rt_pred, const float *scales_gt, const float *scales_pred, const float *t_gt,
                    const float *t_pred, int m, float *raw_degrees, double *sym_degrees, float *iou_3d, float *offset_error,
                    float *world_centers, void *stream) {
    if (m < 0) return fail(FPC_EINVAL, "negative size");
    if (m == 0) return FPC_OK;
    if ((raw_degrees || sym_degrees) && (!q_gt || !q_pred)) return fail(FPC_EINVAL, "quaternions missing");
    if (sym_degrees && !sym_rotations) return fail(FPC_EINVAL, "symmetry rotation table missing");
    if (iou_3d && (!rt_gt || !rt_pred || !scales_gt || !scales_pred)) return fail(FPC_EINVAL, "RT / scales missing");
    if (world_centers && (!rt_gt || !rt_pred)) return fail(FPC_EINVAL, "RT missing");
    if (offset_error && (!t_gt || !t_pred)) return fail(FPC_EINVAL, "translations missing");
    const bool quats = raw_degrees || sym_degrees;
    k_pose_errors<<<ceil_div(m, 4), 128, 0, (cudaStream_t)stream>>>(quats ? q_gt : nullptr, quats ? q_pred : nullptr,
                                                                   (const long long *)symmetric_ids, sym_rotations, rt_gt, rt_pred,
                                                                   scales_gt, scales_pred, t_gt, t_pred, m, raw_degrees, sym_degrees,
                                                                   iou_3d, offset_error, world_centers);
    FPC_LAUNCH_CHECK("k_pose_errors");
    return FPC_OK;
}

int fpc_threshold_fraction(const double *values, int n, const double *thresholds, int num_thresholds, int op, float *out, void *stream) {
    if (n < 0 || num_thresholds < 0 || (op != 0 && op != 1)) return fail(FPC_EINVAL, "bad argument");
    if (num_thresholds == 0) return FPC_OK;
    if ((n > 0 && !values) || !thresholds || !out) return fail(FPC_EINVAL, "NULL pointer");
    k_threshold_fraction<<<num_thresholds, 256, 0, (cudaStream_t)stream>>>(values, n, thresholds, op, out);
    FPC_LAUNCH_CHECK("k_threshold_fraction");
    return FPC_OK;
}

}  // extern "C"
