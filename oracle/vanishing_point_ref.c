/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of the two vanishing-point kernels of PVNet's native module, the other half of the pybind API
 * (lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:104-105):
 *   K3  generate_hypothesis_vanishing_point_kernel     src/ransac_voting_kernel.cu:170-228
 *   K4  voting_for_hypothesis_vanishing_point_kernel   src/ransac_voting_kernel.cu:268-308
 * A hypothesis is a homogeneous point (x, y, z): the cross product of the two pixel rays' lines.
 *
 * Two arithmetic flavours, as for K1/K2 (ransac_voting_ref.c): the plain functions evaluate the source expressions in
 * IEEE binary32 with no fused operations (this file is compiled with -ffp-contract=off); the *_fma functions apply the
 * contraction nvcc 12.9 -O3 produces for sm_100a, read from the SASS of the reference's own source built into
 * oracle/_ref/:  a*b - c*d -> fma(a, b, -(c*d));  u - z*c -> fma(-c, z, u);  a*a + b*b -> fma(a, a, b*b);  the dot
 * product of K4 stays un-contracted because both products are reused for the sign tests.
 */
#include <math.h>
#include <stddef.h>

#define VP_BODY(FMA)                                                                                                  \
    for (int hi = 0; hi < hn; ++hi) {                                                                                 \
        for (int vi = 0; vi < vn; ++vi) {                                                                             \
            const int id0 = idxs[hi * vn * 2 + vi * 2], id1 = idxs[hi * vn * 2 + vi * 2 + 1];                         \
            const float dx0 = direct[id0 * vn * 2 + vi * 2], dy0 = direct[id0 * vn * 2 + vi * 2 + 1];                 \
            const float cx0 = coords[id0 * 2], cy0 = coords[id0 * 2 + 1];                                             \
            const float dx1 = direct[id1 * vn * 2 + vi * 2], dy1 = direct[id1 * vn * 2 + vi * 2 + 1];                 \
            const float cx1 = coords[id1 * 2], cy1 = coords[id1 * 2 + 1];                                             \
            float x, y, z, ex0, ex1, ey0, ey1;                                                                        \
            if (FMA) {                                                                                                \
                const float lz0 = fmaf(dx0, cy0, -(dy0 * cx0)), lz1 = fmaf(dx1, cy1, -(dy1 * cx1));                   \
                z = fmaf(dx0, dy1, -(dy0 * dx1));                                                                     \
                x = fmaf(dx1, lz0, -(dx0 * lz1));                                                                     \
                y = fmaf(dy1, lz0, -(dy0 * lz1));                                                                     \
                ex0 = fmaf(-cx0, z, x); ex1 = fmaf(-cx1, z, x); ey0 = fmaf(-cy0, z, y); ey1 = fmaf(-cy1, z, y);       \
            } else {                                                                                                  \
                const float lx0 = dy0, ly0 = -dx0, lz0 = cy0 * dx0 - cx0 * dy0;                                       \
                const float lx1 = dy1, ly1 = -dx1, lz1 = cy1 * dx1 - cx1 * dy1;                                       \
                x = ly0 * lz1 - lz0 * ly1;                                                                            \
                y = lz0 * lx1 - lx0 * lz1;                                                                            \
                z = lx0 * ly1 - ly0 * lx1;                                                                            \
                ex0 = x - z * cx0; ex1 = x - z * cx1; ey0 = y - z * cy0; ey1 = y - z * cy1;                           \
            }                                                                                                         \
            const float val_x0 = dx0 * ex0, val_x1 = dx1 * ex1, val_y0 = dy0 * ey0, val_y1 = dy1 * ey1;               \
            if (val_x0 < 0 && val_x1 < 0 && val_y0 < 0 && val_y1 < 0) { z = -z; x = -x; y = -y; }                     \
            if (val_x0 * val_x1 < 0 || val_y0 * val_y1 < 0) { x = 0.f; y = 0.f; z = 0.f; }                            \
            float *o = hypo_pts + ((size_t)hi * vn + vi) * 3;                                                         \
            o[0] = x; o[1] = y; o[2] = z;                                                                             \
        }                                                                                                             \
    }

void fpc_ref_generate_hypothesis_vp(const float *direct, const float *coords, const int *idxs, float *hypo_pts /* [hn,vn,3] */,
                                    int tn, int vn, int hn)
{
    (void)tn;
    VP_BODY(0)
}

void fpc_ref_generate_hypothesis_vp_fma(const float *direct, const float *coords, const int *idxs, float *hypo_pts, int tn, int vn,
                                        int hn)
{
    (void)tn;
    VP_BODY(1)
}

#define VP_VOTE_BODY(FMA)                                                                                             \
    for (int hi = 0; hi < hn; ++hi) {                                                                                 \
        for (int vi = 0; vi < vn; ++vi) {                                                                             \
            const float *h = hypo_pts + ((size_t)hi * vn + vi) * 3;                                                   \
            const float hx = h[0], hy = h[1], hz = h[2];                                                              \
            unsigned char *row = inliers + ((size_t)hi * vn + vi) * (size_t)tn;                                       \
            for (int ti = 0; ti < tn; ++ti) {                                                                         \
                const float cx = coords[ti * 2], cy = coords[ti * 2 + 1];                                             \
                const float dx = direct[ti * vn * 2 + vi * 2], dy = direct[ti * vn * 2 + vi * 2 + 1];                 \
                float diff_x, diff_y, norm1, norm2;                                                                   \
                if (FMA) {                                                                                            \
                    diff_x = fmaf(-cx, hz, hx); diff_y = fmaf(-cy, hz, hy);                                           \
                    norm1 = sqrtf(fmaf(dx, dx, dy * dy)); norm2 = sqrtf(fmaf(diff_x, diff_x, diff_y * diff_y));       \
                } else {                                                                                              \
                    diff_x = hx - cx * hz; diff_y = hy - cy * hz;                                                     \
                    norm1 = sqrtf(dx * dx + dy * dy); norm2 = sqrtf(diff_x * diff_x + diff_y * diff_y);               \
                }                                                                                                     \
                if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) continue;                                           \
                const float val_x = diff_x * dx, val_y = diff_y * dy;                                                 \
                const float angle_dist = (val_x + val_y) / (norm1 * norm2);                                           \
                if (val_x < 0 || val_y < 0) continue;                                                                 \
                if (fabsf(angle_dist) > inlier_thresh) row[ti] = 1;                                                   \
            }                                                                                                         \
        }                                                                                                             \
    }

void fpc_ref_voting_for_hypothesis_vp(const float *direct, const float *coords, const float *hypo_pts /* [hn,vn,3] */,
                                      unsigned char *inliers /* [hn,vn,tn], pre-zeroed */, int tn, int vn, int hn, float inlier_thresh)
{
    VP_VOTE_BODY(0)
}

void fpc_ref_voting_for_hypothesis_vp_fma(const float *direct, const float *coords, const float *hypo_pts, unsigned char *inliers,
                                          int tn, int vn, int hn, float inlier_thresh)
{
    VP_VOTE_BODY(1)
}
