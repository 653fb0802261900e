/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Never imported, linked or executed by the
 * product path (fastposecnn_b200/).  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * CPU restatement of the two PVNet voting kernels FastPoseCNN uses, in IEEE
 * binary32 arithmetic with NO fused multiply-add (build with -ffp-contract=off),
 * evaluated in the source order of the reference expressions:
 *
 *   fpc_ref_generate_hypothesis    <- lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu:11-49
 *   fpc_ref_voting_for_hypothesis  <- lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu:88-126
 *
 * plus a variant of each (`*_fma`) that reproduces the contraction pattern nvcc
 * 12.9 applies to the same expressions for sm_100a (a*b + c*d -> fma(a, b, c*d);
 * see DESIGN.md "arithmetic modes"), used to cross-check the product's
 * FPC_ARITH_NVCC_FMA mode.
 *
 * Parity pin: the reference ships no golden vectors for this path (SURVEY.md
 * section 4).  This file is pinned instead against the reference's own Python driver
 * (ransac_voting_layer_v3) executed on top of it, see oracle/ref_import.py and
 * tests/golden/.
 */
#include <math.h>
#include <stddef.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* K1: two-line intersection from the normals of two sampled pixels.
 * hypo_pts must be pre-zeroed by the caller (the reference returns at::zeros
 * and leaves degenerate pairs untouched, .cu:42-43,75). */
void fpc_ref_generate_hypothesis(const float *direct, /* [tn,vn,2] */
                                 const float *coords, /* [tn,2]    */
                                 const int *idxs,     /* [hn,vn,2] */
                                 float *hypo_pts,     /* [hn,vn,2] */
                                 int tn, int vn, int hn)
{
    (void)tn;
    for (int hvi = 0; hvi < hn * vn; ++hvi) {
        int hi = hvi / vn;
        int vi = hvi - hi * vn;
        int t0 = idxs[hi * vn * 2 + vi * 2];
        int t1 = idxs[hi * vn * 2 + vi * 2 + 1];

        float nx0 = direct[t0 * vn * 2 + vi * 2 + 1];
        float ny0 = -direct[t0 * vn * 2 + vi * 2];
        float cx0 = coords[t0 * 2];
        float cy0 = coords[t0 * 2 + 1];

        float nx1 = direct[t1 * vn * 2 + vi * 2 + 1];
        float ny1 = -direct[t1 * vn * 2 + vi * 2];
        float cx1 = coords[t1 * 2];
        float cy1 = coords[t1 * 2 + 1];

        float det_y = nx1 * ny0 - nx0 * ny1;
        float det_x = ny1 * nx0 - ny0 * nx1;
        /* float |det| compared against the DOUBLE literal 1e-6, as in the .cu */
        if ((double)fabsf(det_y) < 1e-6) continue;
        if ((double)fabsf(det_x) < 1e-6) continue;
        float p0 = nx0 * cx0 + ny0 * cy0;
        float p1 = nx1 * cx1 + ny1 * cy1;
        float y = (nx1 * p0 - nx0 * p1) / det_y;
        float x = (ny1 * p0 - ny0 * p1) / det_x;
        hypo_pts[hi * vn * 2 + vi * 2] = x;
        hypo_pts[hi * vn * 2 + vi * 2 + 1] = y;
    }
}

/* K2: cosine test of every (hypothesis, pixel) pair.  inliers must be
 * pre-zeroed by the caller (ransac_voting_gpu.py:562). */
void fpc_ref_voting_for_hypothesis(const float *direct,    /* [tn,vn,2]  */
                                   const float *coords,    /* [tn,2]     */
                                   const float *hypo_pts,  /* [hn,vn,2]  */
                                   unsigned char *inliers, /* [hn,vn,tn] */
                                   int tn, int vn, int hn, float inlier_thresh)
{
#pragma omp parallel for schedule(static)
    for (int hi = 0; hi < hn; ++hi) {
        for (int vi = 0; vi < vn; ++vi) {
            float hx = hypo_pts[hi * vn * 2 + vi * 2];
            float hy = hypo_pts[hi * vn * 2 + vi * 2 + 1];
            unsigned char *row = inliers + ((size_t)hi * vn + vi) * (size_t)tn;
            for (int ti = 0; ti < tn; ++ti) {
                float cx = coords[ti * 2];
                float cy = coords[ti * 2 + 1];
                float nx = direct[ti * vn * 2 + vi * 2];
                float ny = direct[ti * vn * 2 + vi * 2 + 1];
                float dx = hx - cx;
                float dy = hy - cy;
                float norm1 = sqrtf(nx * nx + ny * ny);
                float norm2 = sqrtf(dx * dx + dy * dy);
                if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) continue;
                float angle_dist = (dx * nx + dy * ny) / (norm1 * norm2);
                if (angle_dist > inlier_thresh) row[ti] = 1;
            }
        }
    }
}

/* ---- nvcc-contracted variants (a*b + c*d  ->  fmaf(a, b, c*d)) ------------- */

void fpc_ref_generate_hypothesis_fma(const float *direct, const float *coords,
                                     const int *idxs, float *hypo_pts,
                                     int tn, int vn, int hn)
{
    (void)tn;
    for (int hvi = 0; hvi < hn * vn; ++hvi) {
        int hi = hvi / vn;
        int vi = hvi - hi * vn;
        int t0 = idxs[hi * vn * 2 + vi * 2];
        int t1 = idxs[hi * vn * 2 + vi * 2 + 1];
        float nx0 = direct[t0 * vn * 2 + vi * 2 + 1];
        float ny0 = -direct[t0 * vn * 2 + vi * 2];
        float cx0 = coords[t0 * 2];
        float cy0 = coords[t0 * 2 + 1];
        float nx1 = direct[t1 * vn * 2 + vi * 2 + 1];
        float ny1 = -direct[t1 * vn * 2 + vi * 2];
        float cx1 = coords[t1 * 2];
        float cy1 = coords[t1 * 2 + 1];
        /* Pattern read off the SASS of the reference kernel built with nvcc 12.9
         * for sm_100a: the two determinants stay un-contracted (mul, mul, sub);
         * the projections and the y numerator contract as fma(a, b, +-(c*d));
         * for the x numerator the compiler first cancels the two negations
         * (ny = -direct_x) and then rounds the ny1*p0 product separately. */
        float det_y = nx1 * ny0 - nx0 * ny1;
        float det_x = ny1 * nx0 - ny0 * nx1;
        if ((double)fabsf(det_y) < 1e-6) continue;
        if ((double)fabsf(det_x) < 1e-6) continue;
        float p0 = fmaf(nx0, cx0, ny0 * cy0);
        float p1 = fmaf(nx1, cx1, ny1 * cy1);
        float y = fmaf(nx1, p0, -(nx0 * p1)) / det_y;
        float x = fmaf(-ny0, p1, ny1 * p0) / det_x;
        hypo_pts[hi * vn * 2 + vi * 2] = x;
        hypo_pts[hi * vn * 2 + vi * 2 + 1] = y;
    }
}

void fpc_ref_voting_for_hypothesis_fma(const float *direct, const float *coords,
                                       const float *hypo_pts, unsigned char *inliers,
                                       int tn, int vn, int hn, float inlier_thresh)
{
#pragma omp parallel for schedule(static)
    for (int hi = 0; hi < hn; ++hi) {
        for (int vi = 0; vi < vn; ++vi) {
            float hx = hypo_pts[hi * vn * 2 + vi * 2];
            float hy = hypo_pts[hi * vn * 2 + vi * 2 + 1];
            unsigned char *row = inliers + ((size_t)hi * vn + vi) * (size_t)tn;
            for (int ti = 0; ti < tn; ++ti) {
                float cx = coords[ti * 2];
                float cy = coords[ti * 2 + 1];
                float nx = direct[ti * vn * 2 + vi * 2];
                float ny = direct[ti * vn * 2 + vi * 2 + 1];
                float dx = hx - cx;
                float dy = hy - cy;
                float norm1 = sqrtf(fmaf(nx, nx, ny * ny));
                float norm2 = sqrtf(fmaf(dx, dx, dy * dy));
                if ((double)norm1 < 1e-6 || (double)norm2 < 1e-6) continue;
                float angle_dist = fmaf(dx, nx, dy * ny) / (norm1 * norm2);
                if (angle_dist > inlier_thresh) row[ti] = 1;
            }
        }
    }
}

int fpc_ref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
