// Voting half of the path: hypothesis generation from pre-sampled pixel pairs, inlier vote
// counting, winner selection, inlier refinement, and the per-instance pose finalisation.
//
// Reference behaviour being reproduced (paths relative to /root/reference/source_code/FastPoseCNN/):
//   lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu:11-49    K1 generate_hypothesis
//   lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu:88-126   K2 voting_for_hypothesis
//   lib/ransac_voting_gpu_layer/ransac_voting_gpu.py:518-607         ransac_voting_layer_v3 driver
//   lib/gpu_tensor_funcs.py:204-253, 306-326                         translation / rotation / RT
#include "fpc_internal.cuh"

#include <algorithm>
#include <cmath>
#include <cstdlib>

namespace fpc {

// =============================================================================================
// 1:1 mirrors of the reference's native module (general vn)
// =============================================================================================
template <int ARITH>
__global__ void __launch_bounds__(256) k_generate_hypothesis(const float *__restrict__ direct,
                                                             const float *__restrict__ coords,
                                                             const int *__restrict__ idxs, float *__restrict__ hypo,
                                                             int tn, int vn, int hn) {
    const int hvi = blockIdx.x * blockDim.x + threadIdx.x;
    if (hvi >= hn * vn) return;
    const int hi = hvi / vn, vi = hvi - hi * vn;
    const int t0 = idxs[2 * hvi], t1 = idxs[2 * hvi + 1];
    float x = 0.f, y = 0.f;
    if (t0 >= 0 && t0 < tn && t1 >= 0 && t1 < tn) {
        const float *d0 = direct + ((size_t)t0 * vn + vi) * 2, *d1 = direct + ((size_t)t1 * vn + vi) * 2;
        float hx, hy;
        if (hypothesis_exact<ARITH>(d0[0], d0[1], coords[2 * t0], coords[2 * t0 + 1], d1[0], d1[1], coords[2 * t1],
                                    coords[2 * t1 + 1], hx, hy)) {
            x = hx;
            y = hy;
        }
    }
    hypo[2 * hvi] = x;
    hypo[2 * hvi + 1] = y;
}

constexpr int K2_HYPS_PER_BLOCK = 32;
template <int ARITH>
__global__ void __launch_bounds__(256) k_voting_for_hypothesis(const float *__restrict__ direct,
                                                               const float *__restrict__ coords,
                                                               const float *__restrict__ hypo, uint8_t *__restrict__ inliers,
                                                               int tn, int vn, int hn, float thresh) {
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    const int vi = blockIdx.y;
    if (ti >= tn) return;
    const float cx = coords[2 * ti], cy = coords[2 * ti + 1];
    const float nx = direct[((size_t)ti * vn + vi) * 2], ny = direct[((size_t)ti * vn + vi) * 2 + 1];
    const int h0 = blockIdx.z * K2_HYPS_PER_BLOCK, h1 = min(hn, h0 + K2_HYPS_PER_BLOCK);
    for (int hi = h0; hi < h1; ++hi) {
        const float hx = hypo[((size_t)hi * vn + vi) * 2], hy = hypo[((size_t)hi * vn + vi) * 2 + 1];
        if (vote_exact<ARITH>(cx, cy, nx, ny, hx, hy, thresh)) inliers[((size_t)hi * vn + vi) * tn + ti] = 1;
    }
}

// =============================================================================================
// Vanishing-point twins of K1/K2 (src/ransac_voting_kernel.cu:170-228, 268-308; pybind names
// generate_hypothesis_vanishing_point / voting_for_hypothesis_vanishing_point, ransac_voting.cpp:104-105): a hypothesis
// is the homogeneous point (x, y, z) where the lines of two pixel rays cross.  FPC_ARITH_IEEE evaluates the source
// expressions un-contracted; FPC_ARITH_NVCC_FMA applies the contraction nvcc makes of the reference source (read from
// its SASS): a*b - c*d -> fma(a, b, -(c*d)), u - z*c -> fma(-c, z, u), a*a + b*b -> fma(a, a, b*b).
// =============================================================================================
template <int ARITH>
__device__ __forceinline__ float diff_prod(float a, float b, float c, float d) {   // a*b - c*d
    if (ARITH == FPC_ARITH_IEEE) return __fsub_rn(__fmul_rn(a, b), __fmul_rn(c, d));
    return __fmaf_rn(a, b, -__fmul_rn(c, d));
}
template <int ARITH>
__device__ __forceinline__ float sub_prod(float u, float z, float c) {              // u - z*c
    if (ARITH == FPC_ARITH_IEEE) return __fsub_rn(u, __fmul_rn(z, c));
    return __fmaf_rn(-c, z, u);
}

template <int ARITH>
__global__ void __launch_bounds__(256) k_generate_hypothesis_vp(const float *__restrict__ direct, const float *__restrict__ coords,
                                                                const int *__restrict__ idxs, float *__restrict__ hypo, int tn,
                                                                int vn, int hn) {
    const int hvi = blockIdx.x * blockDim.x + threadIdx.x;
    if (hvi >= hn * vn) return;
    const int vi = hvi % vn;
    const int t0 = idxs[2 * hvi], t1 = idxs[2 * hvi + 1];
    float x = 0.f, y = 0.f, z = 0.f;
    if (t0 >= 0 && t0 < tn && t1 >= 0 && t1 < tn) {
        const float dx0 = direct[((size_t)t0 * vn + vi) * 2], dy0 = direct[((size_t)t0 * vn + vi) * 2 + 1];
        const float dx1 = direct[((size_t)t1 * vn + vi) * 2], dy1 = direct[((size_t)t1 * vn + vi) * 2 + 1];
        const float cx0 = coords[2 * t0], cy0 = coords[2 * t0 + 1], cx1 = coords[2 * t1], cy1 = coords[2 * t1 + 1];
        // lines l = (dy, -dx, cy*dx - cx*dy); point = l0 x l1 (signs folded: ly = -dx)
        const float lz0 = diff_prod<ARITH>(dx0, cy0, dy0, cx0), lz1 = diff_prod<ARITH>(dx1, cy1, dy1, cx1);
        z = diff_prod<ARITH>(dx0, dy1, dy0, dx1);
        x = diff_prod<ARITH>(dx1, lz0, dx0, lz1);
        y = diff_prod<ARITH>(dy1, lz0, dy0, lz1);
        const float val_x0 = __fmul_rn(dx0, sub_prod<ARITH>(x, z, cx0)), val_x1 = __fmul_rn(dx1, sub_prod<ARITH>(x, z, cx1));
        const float val_y0 = __fmul_rn(dy0, sub_prod<ARITH>(y, z, cy0)), val_y1 = __fmul_rn(dy1, sub_prod<ARITH>(y, z, cy1));
        if (val_x0 < 0.f && val_x1 < 0.f && val_y0 < 0.f && val_y1 < 0.f) { x = -x; y = -y; z = -z; }
        if (__fmul_rn(val_x0, val_x1) < 0.f || __fmul_rn(val_y0, val_y1) < 0.f) { x = 0.f; y = 0.f; z = 0.f; }   // rays do not meet
    }
    hypo[3 * hvi] = x;
    hypo[3 * hvi + 1] = y;
    hypo[3 * hvi + 2] = z;
}

template <int ARITH>
__global__ void __launch_bounds__(256) k_voting_for_hypothesis_vp(const float *__restrict__ direct, const float *__restrict__ coords,
                                                                  const float *__restrict__ hypo, uint8_t *__restrict__ inliers,
                                                                  int tn, int vn, int hn, float thresh) {
    const int ti = blockIdx.x * blockDim.x + threadIdx.x;
    const int vi = blockIdx.y;
    if (ti >= tn) return;
    const float cx = coords[2 * ti], cy = coords[2 * ti + 1];
    const float dx = direct[((size_t)ti * vn + vi) * 2], dy = direct[((size_t)ti * vn + vi) * 2 + 1];
    const float norm1 = __fsqrt_rn(sum_prod<ARITH>(dx, dx, dy, dy));
    if (below_1e6(norm1)) return;
    const int h0 = blockIdx.z * K2_HYPS_PER_BLOCK, h1 = min(hn, h0 + K2_HYPS_PER_BLOCK);
    for (int hi = h0; hi < h1; ++hi) {
        const float *h = hypo + ((size_t)hi * vn + vi) * 3;
        const float diff_x = sub_prod<ARITH>(h[0], h[2], cx), diff_y = sub_prod<ARITH>(h[1], h[2], cy);
        const float norm2 = __fsqrt_rn(sum_prod<ARITH>(diff_x, diff_x, diff_y, diff_y));
        if (below_1e6(norm2)) continue;
        const float val_x = __fmul_rn(diff_x, dx), val_y = __fmul_rn(diff_y, dy);
        if (val_x < 0.f || val_y < 0.f) continue;                          // the ray points away from the hypothesis
        const float angle = __fdiv_rn(__fadd_rn(val_x, val_y), __fmul_rn(norm1, norm2));
        if (fabsf(angle) > thresh) inliers[((size_t)hi * vn + vi) * tn + ti] = 1;
    }
}

// =============================================================================================
// Vote counting: constants, geometry of the instance-local frame, the per-vote arithmetic
// =============================================================================================
// Work item = (instance, chunk of <= VOTE_CHUNK voting records, batch of <= VHB hypotheses), handed out through
// an atomic ticket so every resident block stays busy until the work runs out.  Staging is double buffered and
// done by the copy engine: one thread arms an mbarrier and issues five bulk copies (cp.async.bulk: the four
// SoA record planes x, y, dir_x, dir_y of the chunk and the prepared hypotheses of the batch) for the NEXT item while
// all warps vote on the current one.  Every lane keeps VQ hypotheses in registers and walks the pixels with
// broadcast LDS.128 loads (4 pixels per load), two pixels per packed f32x2 instruction (FFMA2).
// No inlier matrix is ever written (the reference materialises hn*tn bytes and re-reads them,
// ransac_voting_gpu.py:562-566).
//
// Exactness.  The reference decides   cos = dot(d,n) / (|n| |d|) > t,  d = fl(h - c),  in rounded binary32 ops; its
// rounded cosine is within (7 + 1/t + T(t)) u of the cosine of the exact d = h - c (u = 2^-24; the T(t) u is fl(h - c)).
// The fast test works in tangent form, far better conditioned next to cos = 1:  with U = d.n and W = d x n,
//   cos > c  <=>  U > 0 and |W| < T(c) U,   T(c) = sqrt(1 - c^2) / c.
// FIVE fused multiply-adds per vote.  Every block first re-expresses the chunk it staged in instance-local
// coordinates (origin o = centre of the instance's bounding box, h' = h - o, c' = c - o exact for pixel coordinates):
// the direction n is scaled by a power of two (exact) so that |n| <= 1, and the pixel planes (x, y) are overwritten by
//   pu = -(c'.n),  pw = -(c' x n)        so that        U = fma(h'x, nx, fma(h'y, ny, pu)),
//                                                        W = fma(h'x, ny, fma(-h'y, nx, pw)),   s = fma(U, -tau, |W|)
// with tau = T(t) (middle of the reference's uncertainty interval).  sign(s) is the fast answer, added to the count
// with one LEA.HI per vote.  The absolute error of U and W is <= 8u (|h'x| + |h'y| + Rx + Ry) =: E (Rx, Ry = half
// extents of the bounding box), so the answer is certain as soon as
//   |s| >= delta(h) = 1.01 [ max(T_lo - tau, tau - T_hi) (|d|max + E) + (1 + T_lo) E ],   |d|max = |h'| + |(Rx, Ry)|,
// T_hi = T(t (1 + eps)), T_lo = T(t (1 - eps)), eps = 1.02 (7 + 1/t + T(t)) u  (vote_consts; derivation in DESIGN.md).  The hot loop only
// tracks  m = min |s|  per hypothesis over VHALF = 8 pixels (one 3-input FMNMX per two votes) and leaves a one-word
// note per (lane, 8 pixels); notes with m < delta (about one vote in 10^3 lands there) are handed to k_vote_settle,
// which recomputes the bit-identical s of their 8 pixels -- one thread per (note, pixel) -- and settles the uncertain
// ones with the reference expression itself (explicitly rounded intrinsics, original operands re-read from the record
// planes), correcting the fast count.  Hypotheses within 1e-3 of a pixel-lattice point (where |d| may fall under the
// reference's 1e-6 guard), non-finite or absurdly far hypotheses have ALL their votes settled that way; pixels whose
// direction the reference's |n| < 1e-6 guard skips are given pw = 1e30 (never an inlier), pixels with an absurd |n|
// get n = 0, pw = 0 (s = 0: always uncertain).
constexpr int VT = 256;            // threads per block
constexpr int VQ = 4;              // hypotheses per lane
constexpr int VHB = 512;           // hypotheses per batch (4 groups of 32 lanes x VQ)
constexpr int VROUND = 16;         // pixels per round (unit of the work split between warps)
#ifndef FPC_VOTE_NOTE_PX
#define FPC_VOTE_NOTE_PX 8
#endif
constexpr int VHALF = FPC_VOTE_NOTE_PX;   // pixels per note (unit of the re-examination by k_vote_settle): 16 halves the notes, 8 halves the settle work
constexpr int VNH = VROUND / VHALF;       // notes per round
constexpr int VSEGCAP = 4096;      // flagged notes one block can hand to k_vote_settle; beyond that items are re-examined in place
constexpr int VNOTES = 4096 * (VNH > 2 ? 2 : VNH);  // raw notes per work item: rounds x notes per round x hypothesis groups (padded to 2^k) x 32 lanes (32 KB at most)
constexpr float V_FAR = 1e12f;     // |h'x| + |h'y| beyond this (or NaN): the hypothesis is voted exactly
constexpr float V_NEVER = 1e30f;   // pw of a pixel that can never be an inlier

struct VoteConsts {
    float ntau;               // -tau
    float half_w;             // max(T_lo - tau, tau - T_hi), rounded up
    float one_plus_tlo;       // 1 + T_lo, rounded up
    int all_exact;            // thresh outside the fast test's domain: settle every vote exactly
    int nb;                   // hypothesis batches per (instance, chunk)
    int debug_skip;           // FPC_VOTE_DEBUG_SKIP (timing experiments only): 1 re-examination, 2 exact list, 4 hot loop, 8 prepare
};

// Instance-local frame: origin = centre of the bounding box (integer valued), half extents bound |c'x|, |c'y|.
struct VoteFrame {
    float ox, oy, rsum, rdiag;
};
__device__ __forceinline__ VoteFrame vote_frame(const InstTables &T, int i) {
    const int x0 = T.xmin[i], x1 = T.xmax[i], y0 = T.ymin[i], y1 = T.ymax[i];
    VoteFrame f;
    f.ox = (float)((x0 + x1 + 1) >> 1);
    f.oy = (float)((y0 + y1 + 1) >> 1);
    const float rx = 0.5f * (float)(x1 - x0) + 1.f, ry = 0.5f * (float)(y1 - y0) + 1.f;   // >= |c'x|, |c'y|
    f.rsum = rx + ry;
    // >= |c'| for every voting pixel: the farthest pixel from the origin as measured by k_gather, never more than the box diagonal
    f.rdiag = 1.0001f * fminf(sqrtf(rx * rx + ry * ry), sqrtf(__int_as_float(T.rmax2[i])) + 1e-3f);
    return f;
}
__device__ __forceinline__ bool near_lattice(float x, float y) {
    return fabsf(x - rintf(x)) < 1e-3f && fabsf(y - rintf(y)) < 1e-3f;
}
// |s| >= band_delta(h') : the sign of s is certainly the reference's answer (see the comment above)
__device__ __forceinline__ float band_delta(float hx, float hy, float rsum, float rdiag, const VoteConsts &vc) {
    const float a = fabsf(hx) + fabsf(hy);
    const float E = 4.8e-7f * (a + rsum);               // 8 u (|h'x| + |h'y| + Rx + Ry), rounded up
    return 1.01f * (vc.half_w * (1.0002f * (a + rdiag) + E) + vc.one_plus_tlo * E);
}

// =============================================================================================
// Hypotheses (K1) for every live instance: one thread per (instance, hypothesis)
// =============================================================================================
// Besides the hypothesis itself (absolute coordinates, the reference's value) every thread writes what the vote kernel
// needs of it, so that no block recomputes it per work item: hloc = (h'x, h'y, band_delta, 0) in the instance-local
// frame; band_delta = 0 marks a hypothesis that stays off the fast path (the fast loop then sees a far-away dummy).
template <int ARITH>
__global__ void __launch_bounds__(256) k_hypotheses(InstTables T, const int *__restrict__ counters, PathParams pp, RecPlanes rec,
                                                    float2 *__restrict__ hyp_g, float4 *__restrict__ hloc, int4 *__restrict__ work,
                                                    float4 *__restrict__ workf, VoteConsts vc) {
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    const int hn = pp.hn, nb = vc.nb;
    const long long total = (long long)N * hn;
    // threads [total, total + N): work descriptors of the vote kernel (instance, first record, pixels | hypotheses << 16,
    // first hypothesis) and the padding records of one instance each; threads [0, total): one hypothesis each
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total + N; idx += (long long)gridDim.x * blockDim.x) {
        if (idx >= total) {
            const int i = (int)(idx - total);
            const int tn = T.tn[i], w0 = T.workoff[i], px = T.pxoff[i];
            const int chunk = vote_item_px(i, N, pp.vote_chunk, pp.vote_tail);
            const int chunks = (tn + chunk - 1) / chunk;
            for (int k = tn; k < ((tn + 15) & ~15); ++k) {            // padding records: can never be inliers
                const size_t o = (size_t)px + k;
                rec.x[o] = 1e18f; rec.y[o] = 1e18f; rec.nx[o] = 0.f; rec.ny[o] = 0.f;
            }
            const VoteFrame f = vote_frame(T, i);
            for (int c = 0; c < chunks; ++c)
                for (int b = 0; b < nb; ++b) {
                    const int npx = min(chunk, tn - c * chunk), nh = min(VHB, hn - b * VHB);
                    work[w0 + c * nb + b] = make_int4(i, px + c * chunk, npx | (nh << 16), b * VHB);
                    workf[w0 + c * nb + b] = make_float4(f.ox, f.oy, f.rsum, f.rdiag);
                }
            continue;
        }
        const int i = (int)(idx / hn), h = (int)(idx - (long long)i * hn);
        const int tn = T.tn[i];
        float2 hp = make_float2(0.f, 0.f);
        float4 hl = make_float4(3e15f, 1e15f, 0.f, 0.f);
        if (tn > 0) {
            int t0, t1;
            if (pp.idxs) {
                t0 = min(max(pp.idxs[idx * 2], 0), tn - 1);
                t1 = min(max(pp.idxs[idx * 2 + 1], 0), tn - 1);
            } else {
                t0 = (int)(hash3(pp.seed, (uint32_t)i, (uint32_t)h, 0u) % (uint32_t)tn);
                t1 = (int)(hash3(pp.seed, (uint32_t)i, (uint32_t)h, 1u) % (uint32_t)tn);
            }
            const size_t b0 = (size_t)T.pxoff[i] + t0, b1 = (size_t)T.pxoff[i] + t1;
            float x, y;
            if (hypothesis_exact<ARITH>(rec.nx[b0], rec.ny[b0], rec.x[b0], rec.y[b0], rec.nx[b1], rec.ny[b1], rec.x[b1],
                                        rec.y[b1], x, y))
                hp = make_float2(x, y);
            const VoteFrame f = vote_frame(T, i);
            const float lx = hp.x - f.ox, ly = hp.y - f.oy;
            if (!vc.all_exact && !near_lattice(hp.x, hp.y) && fabsf(lx) + fabsf(ly) < V_FAR)
                hl = make_float4(lx, ly, band_delta(lx, ly, f.rsum, f.rdiag, vc), 0.f);
        }
        hyp_g[idx] = hp;
        hloc[idx] = hl;
    }
}

// =============================================================================================
// Vote counting kernel
// =============================================================================================
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
__device__ __forceinline__ float min3(float a, float b, float c) {   // FMNMX3 (sm_100)
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---- mbarrier / bulk-copy primitives (PTX ISA 8.x, sm_90+) ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(u64 *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(u64 *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, u64 *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(u64 *bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!ok);
}

// After the prepare pass: pu[] = -(c'.n), pw[] = -(c' x n), nx[] / ny[] = the power-of-two scaled direction.
struct __align__(128) VoteBuf {
    float pu[VOTE_CHUNK], pw[VOTE_CHUNK], nx[VOTE_CHUNK], ny[VOTE_CHUNK];
    float4 hloc[VHB];      // (h'x, h'y, band_delta, -) of the batch, written by k_hypotheses
};
struct VoteItem {
    int wi, i, hb, nh, npx, src;
    VoteFrame f;
};
struct __align__(128) VoteSmem {
    VoteBuf buf[2];
    unsigned notes[VNOTES];        // [round][group][lane]: sign bytes of (min |s| - band_delta) of the lane's VQ hypotheses
    unsigned short exlist[VHB];
    u64 bar[2];
    int4 dw[2];                    // descriptor of the item that goes into buffer b next (cp.async by thread 0, one item ahead)
    float4 df[2];
    int dwi[2];                    // its ticket
    VoteItem item[2];
    int nex[2];
    int segn;                      // entries in this block's segment of the global note list
};

// s of one vote, scalar twin of vote_pair below: the SAME five roundings in the same order.
__device__ __forceinline__ float vote_s(float hx, float hy, float nx, float ny, float pu, float pw, float ntau) {
    const float U = __fmaf_rn(hx, nx, __fmaf_rn(hy, ny, pu));
    const float W = __fmaf_rn(hx, ny, __fmaf_rn(-hy, nx, pw));
    return __fmaf_rn(U, ntau, fabsf(W));
}
// one hypothesis against two pixels (packed f32x2 lanes): adds the two sign bits to cnt, folds |s| into m
template <bool PACKED>
__device__ __forceinline__ void vote_pair(float hx, float hy, float nxa, float nxb, float nya, float nyb, float pua, float pub,
                                          float pwa, float pwb, float ntau, unsigned &cnt, float &m) {
    float sa, sb;
    if (PACKED) {
        const u64 hx2 = pk2(hx, hx), hy2 = pk2(hy, hy), nhy2 = pk2(-hy, -hy);
        const u64 nx2 = pk2(nxa, nxb), ny2 = pk2(nya, nyb);
        const u64 U = fma2(hx2, nx2, fma2(hy2, ny2, pk2(pua, pub)));
        const u64 W = fma2(hx2, ny2, fma2(nhy2, nx2, pk2(pwa, pwb)));
        float wa, wb;
        unpk2(W, wa, wb);
        unpk2(fma2(U, pk2(ntau, ntau), pk2(fabsf(wa), fabsf(wb))), sa, sb);
    } else {
        sa = vote_s(hx, hy, nxa, nya, pua, pwa, ntau);
        sb = vote_s(hx, hy, nxb, nyb, pub, pwb, ntau);
    }
    cnt += __float_as_uint(sa) >> 31;      // LEA.HI: one instruction per sign
    cnt += __float_as_uint(sb) >> 31;
    m = min3(m, fabsf(sa), fabsf(sb));
}

// ---- prepare: one voting record (x, y, dir) -> (pu, pw, scaled dir) in the instance-local frame.  Used by the vote kernel
//      (on the staged chunk) and by the settle kernel (on single records): the SAME roundings, bit-identical results.
// Common case, branch-free: a (nearly) unit direction of a real pixel; in: cx, cy = pixel, out: cx, cy = pu, pw.
// Returns false if the pixel needs prepare_pixel_rare (which is then given the ORIGINAL pixel coordinates).
template <int ARITH>
__device__ __forceinline__ bool prepare_pixel_fast(float &cx, float &cy, const float nx, const float ny, bool real, const VoteFrame &f) {
    const float nn = sum_prod<ARITH>(nx, nx, ny, ny);
    const float ccx = cx - f.ox, ccy = cy - f.oy;
    cx = -__fmaf_rn(ccx, nx, __fmul_rn(ccy, ny));
    cy = -__fmaf_rn(ccx, ny, -__fmul_rn(ccy, nx));
    return real && nn <= 1.0002f && nn >= 0.25f;
}
template <int ARITH>
__device__ __forceinline__ void prepare_pixel_rare(float ocx, float ocy, float &pu, float &pw, float &nx, float &ny, bool real,
                                                   const VoteFrame &f) {
    const float nn = sum_prod<ARITH>(nx, nx, ny, ny);
    if (real && nn <= 1.0002f && nn >= 0.25f) return;         // handled by prepare_pixel_fast
    float X = 0.f, Y = V_NEVER, NX = 0.f, NY = 0.f;           // padding pixel: never an inlier
    if (real) {
        // the reference skips |n| = sqrt(nn) < 1e-6 (.cu:119): certain without the square root unless nn is within 10 % of
        // 1e-12 (sqrt(0.9e-12) < 1e-6 - 5e-8, sqrt(1.1e-12) > 1e-6 + 4e-8)
        bool live = nn > 1.1e-12f;
        if (!live && nn >= 0.9e-12f) live = !below_1e6(__fsqrt_rn(nn));
        if (!live) {
            // |n| under the reference's guard, or NaN: never an inlier
        } else if (!(nn < 1e36f)) {
            Y = 0.f;                                          // absurd direction: s = 0, settled exactly
        } else {
            // scale by a power of two so that nn lands in [0.25, 1): nn = f * 2^e2, f in [1, 2)
            const int e2 = (int)((__float_as_uint(nn) >> 23) & 0xffu) - 127;
            const int sh = (e2 + 2) >> 1;                     // ceil((e2 + 1) / 2), also for e2 < 0
            const float sc = __uint_as_float((unsigned)(127 - sh) << 23);
            NX = nx * sc;
            NY = ny * sc;
            const float ccx = ocx - f.ox, ccy = ocy - f.oy;
            X = -__fmaf_rn(ccx, NX, __fmul_rn(ccy, NY));
            Y = -__fmaf_rn(ccx, NY, -__fmul_rn(ccy, NX));
        }
    }
    pu = X; pw = Y; nx = NX; ny = NY;
}

// Register cap: 64 per thread, so that the kernel's two blocks per SM hold exactly half of the register file and two blocks of
// the bandwidth-bound kernels of the other batches in flight (arg-max: 2 x 14k registers, gather: 2 x 16k) fit beside them.  At 80
// registers (what the compiler takes without a cap) only one of those blocks fits and the kernels of different batches take turns
// instead of sharing the SMs: the hot loop is the same 294 instructions either way, the kernel alone is 2 % slower at 64, the
// pipelined step 4 % (cfg2) to 7 % (cfg4) faster (profiles/r02_ab_vote_registers.txt).
#ifndef FPC_VOTE_MAXREG
#define FPC_VOTE_MAXREG 64
#endif

template <int ARITH, bool PACKED>
__global__ void __maxnreg__(FPC_VOTE_MAXREG) k_vote(InstTables T, int *__restrict__ counters, PathParams pp, RecPlanes rec,
                                                            const float2 *__restrict__ hyp_g, const float4 *__restrict__ hloc_g,
                                                            int *__restrict__ votes, const int4 *__restrict__ work,
                                                            const float4 *__restrict__ workf, uint4 *__restrict__ segs,
                                                            int *__restrict__ segcnt, VoteConsts vc) {
    extern __shared__ __align__(128) unsigned char vote_smem_raw[];
    VoteSmem &sm = *reinterpret_cast<VoteSmem *>(vote_smem_raw);
    if (counters[FPC_CNT_FLAGS]) return;
    const int W = counters[FPC_CNT_WORK];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int hn = pp.hn;
    const float thresh = pp.inlier_thresh;

    // Thread 0 feeds the pipeline and must never wait for global memory (its warp votes too, and the other seven wait
    // for it at the next barrier): tickets are drawn three items ahead, the descriptor of item k+2 travels to shared
    // memory by cp.async while item k is voted on, and staging item k+1 only touches shared memory.
    auto request = [&](int slot, int wi) {      // descriptor of ticket wi -> slot (asynchronous)
        sm.dwi[slot] = wi;
        if (wi < W) {
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sm.dw[slot])), "l"(work + wi) : "memory");
            asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&sm.df[slot])), "l"(workf + wi) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto stage = [&](int b) {                   // descriptor in slot b -> arm the buffer's barrier, start its bulk copies
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        VoteItem it;
        it.wi = sm.dwi[b];
        it.i = it.hb = it.nh = it.npx = it.src = 0;
        it.f.ox = it.f.oy = it.f.rsum = it.f.rdiag = 0.f;
        if (it.wi < W) {
            const int4 d = sm.dw[b];
            const float4 f = sm.df[b];
            it.i = d.x;
            it.src = d.y;
            it.npx = d.z & 0xffff;
            it.nh = d.z >> 16;
            it.hb = d.w;
            it.f.ox = f.x; it.f.oy = f.y; it.f.rsum = f.z; it.f.rdiag = f.w;
            const uint32_t pbytes = (uint32_t)((it.npx + 3) & ~3) * 4u;
            const uint32_t hbytes = (uint32_t)it.nh * 16u;
            const size_t src = (size_t)d.y;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic accesses of this buffer are done
            mbar_expect_tx(&sm.bar[b], 4u * pbytes + hbytes);
            bulk_g2s(sm.buf[b].pu, rec.x + src, pbytes, &sm.bar[b]);
            bulk_g2s(sm.buf[b].pw, rec.y + src, pbytes, &sm.bar[b]);
            bulk_g2s(sm.buf[b].nx, rec.nx + src, pbytes, &sm.bar[b]);
            bulk_g2s(sm.buf[b].ny, rec.ny + src, pbytes, &sm.bar[b]);
            bulk_g2s(sm.buf[b].hloc, hloc_g + (size_t)it.i * hn + it.hb, hbytes, &sm.bar[b]);
        }
        sm.item[b] = it;
        sm.nex[b] = 0;
    };

    uint4 *seg = segs + (size_t)blockIdx.x * VSEGCAP;
    if (tid == 0) {
        mbar_init(&sm.bar[0], 1);
        mbar_init(&sm.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sm.segn = 0;
    }
    __syncthreads();
    int ticket_ahead = 0;   // thread 0 only: ticket of the item two after the current one
    if (tid == 0) {
        const int t0 = atomicAdd(&counters[FPC_CNT_TICKET], 1);
        const int t1 = atomicAdd(&counters[FPC_CNT_TICKET], 1);
        ticket_ahead = atomicAdd(&counters[FPC_CNT_TICKET], 1);
        request(0, t0);
        stage(0);
        request(1, t1);
    }
    __syncthreads();
    int cur = 0;
    uint32_t phase[2] = {0u, 0u};
    while (true) {
        const VoteItem it = sm.item[cur];
        if (it.wi >= W) break;
        if (tid == 0) {                        // stage the next item while this one is voted on; ask for the one after
            stage(cur ^ 1);
            request(cur, ticket_ahead);
            ticket_ahead = atomicAdd(&counters[FPC_CNT_TICKET], 1);
        }
        mbar_wait(&sm.bar[cur], phase[cur]);
        phase[cur] ^= 1u;
        VoteBuf &B = sm.buf[cur];
        const int i = it.i, npx = it.npx, nh = it.nh, hb = it.hb;
        const size_t gsrc = (size_t)it.src;
        const float2 *hyp_i = hyp_g + (size_t)i * hn + hb;             // absolute hypotheses of the batch (exact paths only)
        int *votes_i = votes + (size_t)i * hn + hb;
        const int nrounds = (npx + VROUND - 1) / VROUND;
        // ---- prepare pass: (x, y, dir) -> (pu, pw, scaled dir) in the instance-local frame ----
        if (!(vc.debug_skip & 8) && tid * 4 < nrounds * VROUND) {
            float4 cx = *reinterpret_cast<float4 *>(&B.pu[tid * 4]), cy = *reinterpret_cast<float4 *>(&B.pw[tid * 4]);
            float4 nx = *reinterpret_cast<float4 *>(&B.nx[tid * 4]), ny = *reinterpret_cast<float4 *>(&B.ny[tid * 4]);
            float *pcx = &cx.x, *pcy = &cy.x, *pnx = &nx.x, *pny = &ny.x;
            bool rare = false;
            float ocx[4], ocy[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                ocx[j] = pcx[j]; ocy[j] = pcy[j];
                rare |= !prepare_pixel_fast<ARITH>(pcx[j], pcy[j], pnx[j], pny[j], tid * 4 + j < npx, it.f);
            }
            if (rare) {
                // padding, a direction the reference skips, or one that needs scaling: redo these pixels the long way
#pragma unroll
                for (int j = 0; j < 4; ++j) prepare_pixel_rare<ARITH>(ocx[j], ocy[j], pcx[j], pcy[j], pnx[j], pny[j], tid * 4 + j < npx, it.f);
            }
            *reinterpret_cast<float4 *>(&B.pu[tid * 4]) = cx; *reinterpret_cast<float4 *>(&B.pw[tid * 4]) = cy;
            *reinterpret_cast<float4 *>(&B.nx[tid * 4]) = nx; *reinterpret_cast<float4 *>(&B.ny[tid * 4]) = ny;
        }
        // hypotheses off the fast path (band_delta = 0) -> exact list; dummies behind the last hypothesis of a partial group
        const int G = (nh + 127) >> 7;  // groups of 128 hypotheses (32 lanes x VQ), <= 4
        const int gsh = G > 2 ? 2 : G - 1, Gp = 1 << gsh;                  // groups padded to a power of two (note index)
        for (int k = tid; k < G * 128; k += VT) {
            if (k >= nh) B.hloc[k] = make_float4(3e15f, 1e15f, 0.f, 0.f);
            else if (B.hloc[k].z == 0.f) sm.exlist[atomicAdd(&sm.nex[cur], 1)] = (unsigned short)k;
        }
        __syncthreads();
        // ---- fast voting: warp -> (hypothesis group g, every parts-th round) ----
        const int parts = 8 / Gp;
        if (!(vc.debug_skip & 4) && wv < G * parts) {
            const int g = wv % G, part = wv / G;
            float hx[VQ], hy[VQ];   // instance-local hypotheses
            unsigned dqi[VQ];       // bit pattern of band_delta of each (0: not on the fast path)
            unsigned cnt[VQ];
#pragma unroll
            for (int q = 0; q < VQ; ++q) {
                const float4 v = B.hloc[g * 128 + q * 32 + lane];
                hx[q] = v.x; hy[q] = v.y;
                dqi[q] = __float_as_uint(v.z);
                cnt[q] = 0u;
            }
            // one note word per (lane, VHALF pixels of a round): index ((rd VNH + hf) Gp + g) 32 + lane
            unsigned *note = &sm.notes[(part * VNH * Gp + g) * 32 + lane];
            const int note_step = parts * VNH * Gp * 32;
            for (int rd = part; rd < nrounds; rd += parts, note += note_step) {
                const int kb = rd * VROUND;
#pragma unroll
              for (int hf = 0; hf < VNH; ++hf) {
                float m[VQ];
#pragma unroll
                for (int q = 0; q < VQ; ++q) m[q] = 3e38f;
#pragma unroll
                for (int j4 = hf * VHALF; j4 < (hf + 1) * VHALF; j4 += 4) {
                    const float4 pu = *reinterpret_cast<const float4 *>(&B.pu[kb + j4]);
                    const float4 pw = *reinterpret_cast<const float4 *>(&B.pw[kb + j4]);
                    const float4 nx = *reinterpret_cast<const float4 *>(&B.nx[kb + j4]);
                    const float4 ny = *reinterpret_cast<const float4 *>(&B.ny[kb + j4]);
#pragma unroll
                    for (int q = 0; q < VQ; ++q) {
                        vote_pair<PACKED>(hx[q], hy[q], nx.x, nx.y, ny.x, ny.y, pu.x, pu.y, pw.x, pw.y, vc.ntau, cnt[q], m[q]);
                        vote_pair<PACKED>(hx[q], hy[q], nx.z, nx.w, ny.z, ny.w, pu.z, pu.w, pw.z, pw.w, vc.ntau, cnt[q], m[q]);
                    }
                }
                // byte q of the note: sign set <=> some vote of hypothesis q in these VHALF pixels may be uncertain (min |s| < band_delta,
                // compared as integers: both are non-negative floats).  Stored unconditionally -- no branch, no atomic; the
                // signs counted above stay in the count, k_vote_settle corrects them.
                unsigned t[VQ];
#pragma unroll
                for (int q = 0; q < VQ; ++q) t[q] = __float_as_uint(m[q]) - dqi[q];
                note[hf * Gp * 32] = __byte_perm(__byte_perm(t[0], t[1], 0x0073), __byte_perm(t[2], t[3], 0x7300), 0x7610);
              }
            }
#pragma unroll
            for (int q = 0; q < VQ; ++q)
                if (dqi[q] && cnt[q]) atomicAdd(&votes_i[g * 128 + q * 32 + lane], (int)cnt[q]);
        }
        __syncthreads();
        // ---- flagged notes (sign bit of a byte set) -> this block's segment of the global list; k_vote_settle re-examines
        //      them after the kernel (bit-identical s, the reference expression where |s| < band_delta) and corrects the counts
        const int nnotes = (vc.debug_skip & 1) ? 0 : nrounds * VNH * Gp * 32;
        for (int e4 = tid * 4; e4 < nnotes; e4 += VT * 4) {
            // four notes per 16-byte load; almost all of them carry no flag
            const uint4 w4 = *reinterpret_cast<const uint4 *>(&sm.notes[e4]);
            if (!((w4.x | w4.y | w4.z | w4.w) & 0x80808080u)) continue;
            const unsigned ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const unsigned w = ws[k] & 0x80808080u;
                const int e = e4 + k, g = (e >> 5) & (Gp - 1);
                if (!w || g >= G) continue;
                unsigned qm = ((w >> 7) & 1u) | ((w >> 14) & 2u) | ((w >> 21) & 4u) | ((w >> 28) & 8u);
                const int ln = e & 31, hr = e >> (5 + gsh);             // lane, index of the VHALF-pixel unit
                const int slot = atomicAdd(&sm.segn, 1);
                if (slot < VSEGCAP) {
                    // self-contained entry: first record of the round, first hypothesis of the lane, frame origin, q mask | live pixels
                    seg[slot] = make_uint4((unsigned)(gsrc + hr * VHALF), (unsigned)((size_t)i * hn + hb + g * 128 + ln),
                                           (unsigned)it.f.ox | ((unsigned)it.f.oy << 16), qm | ((unsigned)max(0, min(VHALF, npx - hr * VHALF)) << 4));
                } else {
                    // the segment is full: re-examine this note right here
                    while (qm) {
                        const int q = __ffs(qm) - 1;
                        qm &= qm - 1;
                        const int hidx = g * 128 + q * 32 + ln;
                        const float4 v = B.hloc[hidx];
                        for (int j = 0; j < VHALF; ++j) {
                            const int k2 = hr * VHALF + j;
                            const float s = vote_s(v.x, v.y, B.nx[k2], B.ny[k2], B.pu[k2], B.pw[k2], vc.ntau);
                            if (fabsf(s) < v.z) {
                                const int fast = (int)(__float_as_uint(s) >> 31);
                                const float2 hp = hyp_i[hidx];
                                const int exact = vote_exact<ARITH>(rec.x[gsrc + k2], rec.y[gsrc + k2], rec.nx[gsrc + k2], rec.ny[gsrc + k2], hp.x, hp.y, thresh) ? 1 : 0;
                                if (exact != fast) atomicAdd(&votes_i[hidx], exact - fast);
                            }
                        }
                    }
                }
            }
        }
        // ---- hypotheses on (or within 1e-3 of) the pixel lattice, far or non-finite: every vote with the reference expression ----
        const int nex = (vc.debug_skip & 2) ? 0 : sm.nex[cur];
        for (int e = 0; e < nex; ++e) {
            const int hidx = sm.exlist[e];
            const float2 hp = hyp_i[hidx];
            int c = 0;
            for (int k = tid; k < npx; k += VT)
                c += vote_exact<ARITH>(rec.x[gsrc + k], rec.y[gsrc + k], rec.nx[gsrc + k], rec.ny[gsrc + k], hp.x, hp.y, thresh) ? 1 : 0;
            c = __reduce_add_sync(FULL, c);
            if (lane == 0 && c) atomicAdd(&votes_i[hidx], c);
        }
        __syncthreads();   // every read of buffer `cur` is done: it may be refilled by the next prefetch
        cur ^= 1;
    }
    if (tid == 0) segcnt[blockIdx.x] = min(sm.segn, VSEGCAP);
}

// Re-examination of the rounds the vote kernel flagged: thread = (note, pixel of its round).  The record is prepared
// and s recomputed with the vote kernel's own functions (bit-identical); where |s| < band_delta the reference expression on
// the original operands decides, and a fast answer that differs is corrected in the vote counts.
constexpr int SETTLE_SPLIT = 16;
template <int ARITH>
__global__ void __launch_bounds__(256) k_vote_settle(const int *__restrict__ counters, PathParams pp, RecPlanes rec,
                                                     const float2 *__restrict__ hyp_g, const float4 *__restrict__ hloc_g,
                                                     int *__restrict__ votes, const uint4 *__restrict__ segs,
                                                     const int *__restrict__ segcnt, int nseg, VoteConsts vc) {
    if (counters[FPC_CNT_FLAGS]) return;
    // every thread's iteration is a chain of dependent loads (note -> record, hypothesis): the kernel is latency-bound, so
    // each segment is spread over SETTLE_SPLIT blocks
    for (int job = blockIdx.x; job < nseg * SETTLE_SPLIT; job += gridDim.x) {
        const int sgi = job / SETTLE_SPLIT, part = job - sgi * SETTLE_SPLIT;
        const int n = min(segcnt[sgi], VSEGCAP);
        const uint4 *seg = segs + (size_t)sgi * VSEGCAP;
        for (int idx = part * blockDim.x + threadIdx.x; idx < n * VHALF; idx += SETTLE_SPLIT * blockDim.x) {
            const uint4 ent = seg[idx / VHALF];
            const int j = idx % VHALF;
            const VoteFrame f{(float)(ent.z & 0xffffu), (float)(ent.z >> 16), 0.f, 0.f};
            const bool real = j < (int)(ent.w >> 4);
            unsigned qm = ent.w & 0xfu;
            const size_t r = (size_t)ent.x + j;                    // padded ranges: always readable
            const float x = rec.x[r], y = rec.y[r], dx = rec.nx[r], dy = rec.ny[r];
            float pu = x, pw = y, nx = dx, ny = dy;
            if (!prepare_pixel_fast<ARITH>(pu, pw, nx, ny, real, f)) prepare_pixel_rare<ARITH>(x, y, pu, pw, nx, ny, real, f);
            while (qm) {
                const int q = __ffs(qm) - 1;
                qm &= qm - 1;
                const size_t h = (size_t)ent.y + q * 32;
                const float4 v = hloc_g[h];
                const float s = vote_s(v.x, v.y, nx, ny, pu, pw, vc.ntau);
                if (fabsf(s) < v.z) {
                    const int fast = (int)(__float_as_uint(s) >> 31);
                    const float2 hp = hyp_g[h];
                    const int exact = vote_exact<ARITH>(x, y, dx, dy, hp.x, hp.y, pp.inlier_thresh) ? 1 : 0;
                    if (exact != fast) atomicAdd(&votes[h], exact - fast);
                }
            }
        }
    }
}

// =============================================================================================
// Winner, refinement, masked means, pose
// =============================================================================================
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(FULL, v, d);
    return v;
}

// Moore-Penrose inverse of the symmetric PSD 2x2 [[a,b],[b,c]] applied to (r0, r1)
// (ransac_voting_gpu.py:598 with b_inv -> torch.pinverse on torch >= 2, :503-516).
__device__ inline void solve_sym2_pinv(double a, double b, double c, double r0, double r1, double &x, double &y) {
    const double tr = a + c;
    const double df = a - c;
    const double rt = sqrt(df * df + 4.0 * b * b);
    const double l1 = 0.5 * (tr + rt);   // >= l2
    const double l2 = 0.5 * (tr - rt);
    x = 0.0;
    y = 0.0;
    if (!(l1 > 0.0)) return;             // zero matrix (no inliers): pinverse = 0
    const double det = a * c - b * b;
    if (l2 > 1e-7 * l1) {                // full rank at binary32 resolution: ordinary inverse
        x = (c * r0 - b * r1) / det;
        y = (a * r1 - b * r0) / det;
        return;
    }
    // rank 1: project on the dominant eigenvector v1, x = v1 (v1 . r) / l1
    double vx, vy;
    if (fabs(b) > 0.0) { vx = l1 - c; vy = b; } else if (a >= c) { vx = 1.0; vy = 0.0; } else { vx = 0.0; vy = 1.0; }
    const double nn = vx * vx + vy * vy;
    const double pr = (vx * r0 + vy * r1) / (nn * l1);
    x = vx * pr;
    y = vy * pr;
}

constexpr int FT = 256;   // threads per instance in k_finalize
template <int ARITH>
__global__ void __launch_bounds__(FT) k_finalize(InstTables T, RowTables R, const int *__restrict__ counters,
                                                  PathParams pp, RecPlanes rec,
                                                  const float2 *__restrict__ hyp_g, const int *__restrict__ votes,
                                                  const float *__restrict__ inv_k, float *__restrict__ table) {
    __shared__ double s_d[FT / 32][13];
    __shared__ int s_i[FT / 32][3];
    __shared__ float s_win[2];
    __shared__ float s_ref[3];
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int hn = pp.hn;
    for (int i = blockIdx.x; i < N; i += gridDim.x) {
        const int tn = T.tn[i];
        const int cnt = T.count[i];
        // ---- winner: first maximum of the vote counts (torch.max over dim 0, ransac_voting_gpu.py:567)
        int bv = -1, bi = INT_MAX;
        if (tn > 0)
            for (int h = tid; h < hn; h += FT) {
                const int v = votes[(size_t)i * hn + h];
                if (v > bv) { bv = v; bi = h; }
            }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            const int ov = __shfl_xor_sync(FULL, bv, d), oi = __shfl_xor_sync(FULL, bi, d);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        __syncthreads();  // previous instance's shared values consumed
        if (lane == 0) { s_i[wv][0] = bv; s_i[wv][1] = bi; }
        __syncthreads();
        if (tid == 0) {
            for (int k = 1; k < FT / 32; ++k)
                if (s_i[k][0] > bv || (s_i[k][0] == bv && s_i[k][1] < bi)) { bv = s_i[k][0]; bi = s_i[k][1]; }
            float2 wpt = make_float2(0.f, 0.f);
            // the running best only moves when the ratio strictly improves on 0 (ransac_voting_gpu.py:572-574)
            if (tn > 0 && bv > 0) wpt = hyp_g[(size_t)i * hn + bi];
            s_win[0] = wpt.x;
            s_win[1] = wpt.y;
            s_i[0][0] = bv;
            s_i[0][1] = bi;
        }
        __syncthreads();
        const float wx = s_win[0], wy = s_win[1];
        const int win_votes = s_i[0][0], win_idx = s_i[0][1];
        // ---- refinement vote + normal equations over the inliers (ransac_voting_gpu.py:584-598)
        double a00 = 0, a01 = 0, a11 = 0, b0 = 0, b1 = 0;
        int ninl = 0;
        const size_t rb = (size_t)T.pxoff[i];     // multiple of 16 records; the range is padded with never-inlier records
        for (int k4 = tid * 4; k4 < tn; k4 += FT * 4) {
            const float4 X = *reinterpret_cast<const float4 *>(rec.x + rb + k4);
            const float4 Y = *reinterpret_cast<const float4 *>(rec.y + rb + k4);
            const float4 NX = *reinterpret_cast<const float4 *>(rec.nx + rb + k4);
            const float4 NY = *reinterpret_cast<const float4 *>(rec.ny + rb + k4);
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w};
            const float nxs[4] = {NX.x, NX.y, NX.z, NX.w}, nys[4] = {NY.x, NY.y, NY.z, NY.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (k4 + j < tn && vote_exact<ARITH>(xs[j], ys[j], nxs[j], nys[j], wx, wy, pp.inlier_thresh)) {
                    const double nx = nys[j], ny = -(double)nxs[j];    // normal = (dir_y, -dir_x)
                    const double bb = nx * xs[j] + ny * ys[j];
                    a00 += nx * nx; a01 += nx * ny; a11 += ny * ny;
                    b0 += nx * bb; b1 += ny * bb;
                    ++ninl;
                }
            }
        }
        // ---- masked sums: add up the (instance,row) partials in a fixed order
        double sm[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[k] = 0.0;
        for (int r = T.rowoff[i] + tid; r < T.rowoff[i + 1]; r += FT) {
            const float4 u = *reinterpret_cast<const float4 *>(R.sum + (size_t)r * 8);
            const float4 v = *reinterpret_cast<const float4 *>(R.sum + (size_t)r * 8 + 4);
            sm[0] += u.x; sm[1] += u.y; sm[2] += u.z; sm[3] += u.w;
            sm[4] += v.x; sm[5] += v.y; sm[6] += v.z; sm[7] += v.w;
        }
        double red[13] = {a00, a01, a11, b0, b1, sm[0], sm[1], sm[2], sm[3], sm[4], sm[5], sm[6], sm[7]};
#pragma unroll
        for (int k = 0; k < 13; ++k) red[k] = warp_sum_d(red[k]);
        ninl = __reduce_add_sync(FULL, ninl);
        if (lane == 0) {
#pragma unroll
            for (int k = 0; k < 13; ++k) s_d[wv][k] = red[k];
            s_i[wv][2] = ninl;
        }
        __syncthreads();
        if (tid == 0) {
            double t[13];
            int inl = 0;
            for (int k = 0; k < 13; ++k) t[k] = 0.0;
            for (int wq = 0; wq < FT / 32; ++wq) {          // fixed order: deterministic sums
                for (int k = 0; k < 13; ++k) t[k] += s_d[wq][k];
                inl += s_i[wq][2];
            }
            double rx = wx, ry = wy;   // v1 (ransac_voting_gpu.py:11-98) returns the winning hypothesis itself
            if (pp.refine) {
                rx = 0.0; ry = 0.0;
                if (tn > 0) solve_sym2_pinv(t[0], t[1], t[2], t[3], t[4], rx, ry);
            }
            const float x = (float)rx, y = (float)ry;
            // means (aggregation_layer.py:138-149)
            const double inv_cnt = 1.0 / (double)cnt;
            float q[4], sc[3];
            for (int k = 0; k < 4; ++k) q[k] = (float)(t[5 + k] * inv_cnt);
            for (int k = 0; k < 3; ++k) sc[k] = (float)(t[9 + k] * inv_cnt);
            const float z = expf((float)(t[12] * inv_cnt));
            float qn = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
            if (qn != 0.f) for (int k = 0; k < 4; ++k) q[k] /= qn;
            float *row = table + (size_t)i * FPC_POSE_ROW;
            row[FPC_ROW_CLASS] = __int_as_float(T.mincls[i]);
            row[FPC_ROW_SAMPLE] = __int_as_float(T.root[i] / pp.hw);
            row[FPC_ROW_COUNT] = __int_as_float(cnt);
            for (int k = 0; k < 4; ++k) row[FPC_ROW_Q + k] = q[k];
            for (int k = 0; k < 3; ++k) row[FPC_ROW_SCALES + k] = sc[k];
            row[FPC_ROW_XY] = x;
            row[FPC_ROW_XY + 1] = y;
            row[FPC_ROW_Z] = z;
            if (inv_k) pose_from_qxyz(q, x, y, z, inv_k, row + FPC_ROW_R, row + FPC_ROW_T, row + FPC_ROW_RT);
            row[FPC_ROW_HYP] = wx;
            row[FPC_ROW_HYP + 1] = wy;
            row[FPC_ROW_WIN_IDX] = __int_as_float(tn > 0 ? win_idx : -1);
            row[FPC_ROW_WIN_COUNT] = __int_as_float(tn > 0 ? win_votes : 0);
            row[FPC_ROW_TN] = __int_as_float(tn);
            row[FPC_ROW_REFINE_INL] = __int_as_float(inl);
            row[FPC_ROW_BBOX] = __int_as_float((T.ymin[i] << 16) | (T.xmin[i] & 0xffff));
            s_ref[0] = x;
            s_ref[1] = y;
            s_ref[2] = qn;
        }
        if (pp.extra) {
            // ---- PVNet v4 / v5 extras: one more pass over the records with the refined point
            __syncthreads();
            const float rx = s_ref[0], ry = s_ref[1];
            double ss = 0.0;
            int nin = 0, nconf = 0;
            for (int k4 = tid * 4; k4 < tn; k4 += FT * 4) {
                const float4 X = *reinterpret_cast<const float4 *>(rec.x + rb + k4);
                const float4 Y = *reinterpret_cast<const float4 *>(rec.y + rb + k4);
                const float4 NX = *reinterpret_cast<const float4 *>(rec.nx + rb + k4);
                const float4 NY = *reinterpret_cast<const float4 *>(rec.ny + rb + k4);
                const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w};
                const float nxs[4] = {NX.x, NX.y, NX.z, NX.w}, nys[4] = {NY.x, NY.y, NY.z, NY.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (k4 + j >= tn) continue;
                    if (vote_exact<ARITH>(xs[j], ys[j], nxs[j], nys[j], wx, wy, pp.inlier_thresh)) {
                        // residual of the inlier's line equation at the refined point (ransac_voting_gpu.py:757-758)
                        const double nx = nys[j], ny = -(double)nxs[j];
                        const double res = nx * ((double)rx - xs[j]) + ny * ((double)ry - ys[j]);
                        ss += res * res;
                        ++nin;
                    }
                    nconf += vote_exact<ARITH>(xs[j], ys[j], nxs[j], nys[j], rx, ry, 0.999f);   // :855-856
                }
            }
            ss = warp_sum_d(ss);
            nin = __reduce_add_sync(FULL, nin);
            nconf = __reduce_add_sync(FULL, nconf);
            __syncthreads();                       // s_d / s_i of the main pass were consumed by thread 0 above
            if (lane == 0) { s_d[wv][0] = ss; s_i[wv][0] = nin; s_i[wv][1] = nconf; }
            __syncthreads();
            if (tid == 0) {
                double tot = 0.0;
                int a = 0, c = 0;
                for (int wq = 0; wq < FT / 32; ++wq) { tot += s_d[wq][0]; a += s_i[wq][0]; c += s_i[wq][1]; }
                pp.extra[4 * (size_t)i] = tn > 0 ? (float)(tot / (double)a) : 1.f;         // v4 skip value: ones (:696)
                pp.extra[4 * (size_t)i + 1] = tn > 0 ? (float)c / (float)tn : 0.f;           // v5 skip value: zeros (:793)
                pp.extra[4 * (size_t)i + 2] = s_ref[2];                                      // |mean quaternion| before normalisation
                pp.extra[4 * (size_t)i + 3] = 0.f;
            }
        }
    }
}

// gpu_tensor_funcs.batchwise_get_RT for n instances (drop-in samplewise_get_RT)
__global__ void __launch_bounds__(128) k_get_rt(const float *__restrict__ q, const float *__restrict__ xy,
                                                const float *__restrict__ z, const float *__restrict__ inv_k,
                                                float *__restrict__ R, float *__restrict__ Tt, float *__restrict__ RT,
                                                int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float qq[4] = {q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]};
    pose_from_qxyz(qq, xy[2 * i], xy[2 * i + 1], z[i], inv_k, R + (size_t)i * 9, Tt + (size_t)i * 3, RT + (size_t)i * 16);
}

// =============================================================================================
// Training support: backward of the refinement solve (ransac_voting_gpu.py:584-598) w.r.t. the direction field.
// x* = A^-1 b with A = sum n n^T, b = sum n (n . c) over the winner's inliers (n = (dir_y, -dir_x)); the inlier set is a
// constant of the differentiation, exactly as in the reference where it enters as a 0/1 weight.  For an upstream
// gradient g on x*:  u = A^+ g,  r = c - x*,  dL/dn = (n . r) u + r (n . u),  dL/d dir = (-dn_y, dn_x).
// Block per instance, two passes over its [h,w] plane; every element of d_vertex is written (zeros off the inlier set).
// =============================================================================================
template <int ARITH>
__global__ void __launch_bounds__(256) k_vote_refine_backward(const float *__restrict__ fmask, const float *__restrict__ vertex,
                                                              long long sN, long long sH, long long sW, long long s2,
                                                              const float *__restrict__ win_pts, const float *__restrict__ refined,
                                                              const float *__restrict__ g_x, const int *__restrict__ live, float thresh,
                                                              int h, int w, float *__restrict__ d_vertex) {
    __shared__ double s_a[8][3];
    __shared__ double s_u[2];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int hw = h * w;
    const float wx = win_pts[2 * i], wy = win_pts[2 * i + 1];
    const bool on = live[i] != 0;
    const float *mk = fmask + (size_t)i * hw;
    const float *vt = vertex + (long long)i * sN;
    double a00 = 0, a01 = 0, a11 = 0;
    if (on)
        for (int p = tid; p < hw; p += 256) {
            if (mk[p] == 0.f) continue;
            const int y = p / w, x = p - y * w;
            const float dx = vt[(long long)y * sH + (long long)x * sW], dy = vt[(long long)y * sH + (long long)x * sW + s2];
            if (!vote_exact<ARITH>((float)x, (float)y, dx, dy, wx, wy, thresh)) continue;
            const double nx = dy, ny = -(double)dx;
            a00 += nx * nx; a01 += nx * ny; a11 += ny * ny;
        }
    a00 = warp_sum_d(a00); a01 = warp_sum_d(a01); a11 = warp_sum_d(a11);
    if (lane == 0) { s_a[wv][0] = a00; s_a[wv][1] = a01; s_a[wv][2] = a11; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0, t1 = 0, t2 = 0, ux = 0, uy = 0;
        for (int k = 0; k < 8; ++k) { t0 += s_a[k][0]; t1 += s_a[k][1]; t2 += s_a[k][2]; }
        if (on) solve_sym2_pinv(t0, t1, t2, (double)g_x[2 * i], (double)g_x[2 * i + 1], ux, uy);
        s_u[0] = ux;
        s_u[1] = uy;
    }
    __syncthreads();
    const double ux = s_u[0], uy = s_u[1];
    const double rx0 = refined[2 * i], ry0 = refined[2 * i + 1];
    float2 *out = reinterpret_cast<float2 *>(d_vertex) + (size_t)i * hw;
    for (int p = tid; p < hw; p += 256) {
        float2 g = make_float2(0.f, 0.f);
        if (on && mk[p] != 0.f) {
            const int y = p / w, x = p - y * w;
            const float dx = vt[(long long)y * sH + (long long)x * sW], dy = vt[(long long)y * sH + (long long)x * sW + s2];
            if (vote_exact<ARITH>((float)x, (float)y, dx, dy, wx, wy, thresh)) {
                const double nx = dy, ny = -(double)dx;
                const double rx = (double)x - rx0, ry = (double)y - ry0;
                const double nr = nx * rx + ny * ry, nu = nx * ux + ny * uy;
                const double gnx = nr * ux + rx * nu, gny = nr * uy + ry * nu;
                g = make_float2((float)(-gny), (float)gnx);
            }
        }
        out[p] = g;
    }
}

// The same for the fused path: membership from the label volume, direction = the predicted class's raw xy channels
// L2-normalised exactly as the gather kernel does (IEEE sqrt / divide), gradient pushed back through that normalisation
// into the raw head map (pre-zeroed).  Block per instance, two passes over its frame.
template <int ARITH>
__global__ void __launch_bounds__(256) k_vote_refine_backward_labels(const int *__restrict__ labels, const uint8_t *__restrict__ cls,
                                                                     const float *__restrict__ xy_head, const int *__restrict__ frame_of,
                                                                     const float *__restrict__ win_pts, const float *__restrict__ refined,
                                                                     const float *__restrict__ g_x, const int *__restrict__ live,
                                                                     float thresh, int K, int h, int w, float *__restrict__ d_xy_head) {
    __shared__ double s_a[8][3];
    __shared__ double s_u[2];
    const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int hw = h * w, want = i + 1;
    const int img = frame_of[i];
    const bool on = live[i] != 0;
    const float wx = win_pts[2 * i], wy = win_pts[2 * i + 1];
    const int *lab = labels + (size_t)img * hw;
    const uint8_t *cl = cls + (size_t)img * hw;
    const float *head = xy_head + (size_t)img * 2 * K * hw;
    float *dhead = d_xy_head + (size_t)img * 2 * K * hw;
    auto direction = [&](int p, float &rx, float &ry, float &nrm, float &dx, float &dy, size_t &off) {
        off = (size_t)2 * ((int)cl[p] - 1) * hw + p;
        rx = head[off];
        ry = head[off + hw];
        nrm = torch_norm2(rx, ry);
        dx = rx; dy = ry;
        if (nrm != 0.f) { dx = __fdiv_rn(rx, nrm); dy = __fdiv_rn(ry, nrm); }
    };
    double a00 = 0, a01 = 0, a11 = 0;
    if (on)
        for (int p = tid; p < hw; p += 256) {
            if (lab[p] != want) continue;
            float rx, ry, nrm, dx, dy;
            size_t off;
            direction(p, rx, ry, nrm, dx, dy, off);
            const int y = p / w, x = p - y * w;
            if (!vote_exact<ARITH>((float)x, (float)y, dx, dy, wx, wy, thresh)) continue;
            const double nx = dy, ny = -(double)dx;
            a00 += nx * nx; a01 += nx * ny; a11 += ny * ny;
        }
    a00 = warp_sum_d(a00); a01 = warp_sum_d(a01); a11 = warp_sum_d(a11);
    if (lane == 0) { s_a[wv][0] = a00; s_a[wv][1] = a01; s_a[wv][2] = a11; }
    __syncthreads();
    if (tid == 0) {
        double t0 = 0, t1 = 0, t2 = 0, ux = 0, uy = 0;
        for (int k = 0; k < 8; ++k) { t0 += s_a[k][0]; t1 += s_a[k][1]; t2 += s_a[k][2]; }
        if (on) solve_sym2_pinv(t0, t1, t2, (double)g_x[2 * i], (double)g_x[2 * i + 1], ux, uy);
        s_u[0] = ux;
        s_u[1] = uy;
    }
    __syncthreads();
    if (!on) return;
    const double ux = s_u[0], uy = s_u[1];
    const double rx0 = refined[2 * i], ry0 = refined[2 * i + 1];
    for (int p = tid; p < hw; p += 256) {
        if (lab[p] != want) continue;
        float rx, ry, nrm, dx, dy;
        size_t off;
        direction(p, rx, ry, nrm, dx, dy, off);
        const int y = p / w, x = p - y * w;
        if (!vote_exact<ARITH>((float)x, (float)y, dx, dy, wx, wy, thresh)) continue;
        const double nx = dy, ny = -(double)dx;
        const double qx = (double)x - rx0, qy = (double)y - ry0;
        const double nr = nx * qx + ny * qy, nu = nx * ux + ny * uy;
        const double gnx = nr * ux + qx * nu, gny = nr * uy + qy * nu;
        double gdx = -gny, gdy = gnx;                                   // gradient on the unit direction
        if (nrm != 0.f) {                                               // ... back through dir = raw / |raw|
            const double dot = dx * gdx + dy * gdy;
            gdx = (gdx - dx * dot) / nrm;
            gdy = (gdy - dy * dot) / nrm;
        }
        dhead[off] = (float)gdx;
        dhead[off + hw] = (float)gdy;
    }
}

int launch_vote_refine_backward_labels(const int *labels, const uint8_t *cls, const float *xy_head, const int *frame_of,
                                       const float *win_pts, const float *refined, const float *g_x, const int *live, float thresh,
                                       int n, int K, int h, int w, int arith, float *d_xy_head, cudaStream_t st) {
    if (n == 0) return FPC_OK;
    if (arith == FPC_ARITH_IEEE)
        k_vote_refine_backward_labels<FPC_ARITH_IEEE><<<n, 256, 0, st>>>(labels, cls, xy_head, frame_of, win_pts, refined, g_x, live, thresh, K, h, w, d_xy_head);
    else
        k_vote_refine_backward_labels<FPC_ARITH_NVCC_FMA><<<n, 256, 0, st>>>(labels, cls, xy_head, frame_of, win_pts, refined, g_x, live, thresh, K, h, w, d_xy_head);
    FPC_LAUNCH_CHECK("k_vote_refine_backward_labels");
    return FPC_OK;
}

int launch_vote_refine_backward(const float *fmask, const float *vertex, long long sN, long long sH, long long sW, long long s2,
                                const float *win_pts, const float *refined, const float *g_x, const int *live, float thresh, int n,
                                int h, int w, int arith, float *d_vertex, cudaStream_t st) {
    if (n == 0) return FPC_OK;
    if (arith == FPC_ARITH_IEEE)
        k_vote_refine_backward<FPC_ARITH_IEEE><<<n, 256, 0, st>>>(fmask, vertex, sN, sH, sW, s2, win_pts, refined, g_x, live, thresh, h, w, d_vertex);
    else
        k_vote_refine_backward<FPC_ARITH_NVCC_FMA><<<n, 256, 0, st>>>(fmask, vertex, sN, sH, sW, s2, win_pts, refined, g_x, live, thresh, h, w, d_vertex);
    FPC_LAUNCH_CHECK("k_vote_refine_backward");
    return FPC_OK;
}

// =============================================================================================
// host-side launchers
// =============================================================================================
int launch_generate_hypothesis(const float *direct, const float *coords, const int *idxs, float *hypo, int tn, int vn,
                               int hn, int arith, cudaStream_t st) {
    const int n = hn * vn;
    if (n == 0) return FPC_OK;
    if (arith == FPC_ARITH_IEEE)
        k_generate_hypothesis<FPC_ARITH_IEEE><<<ceil_div(n, 256), 256, 0, st>>>(direct, coords, idxs, hypo, tn, vn, hn);
    else
        k_generate_hypothesis<FPC_ARITH_NVCC_FMA><<<ceil_div(n, 256), 256, 0, st>>>(direct, coords, idxs, hypo, tn, vn, hn);
    FPC_LAUNCH_CHECK("k_generate_hypothesis");
    return FPC_OK;
}

int launch_voting_for_hypothesis(const float *direct, const float *coords, const float *hypo, uint8_t *inliers, int tn,
                                 int vn, int hn, float thresh, int arith, cudaStream_t st) {
    if (tn == 0 || vn == 0 || hn == 0) return FPC_OK;
    dim3 grid(ceil_div(tn, 256), vn, ceil_div(hn, K2_HYPS_PER_BLOCK));
    if (arith == FPC_ARITH_IEEE)
        k_voting_for_hypothesis<FPC_ARITH_IEEE><<<grid, 256, 0, st>>>(direct, coords, hypo, inliers, tn, vn, hn, thresh);
    else
        k_voting_for_hypothesis<FPC_ARITH_NVCC_FMA><<<grid, 256, 0, st>>>(direct, coords, hypo, inliers, tn, vn, hn, thresh);
    FPC_LAUNCH_CHECK("k_voting_for_hypothesis");
    return FPC_OK;
}

int launch_generate_hypothesis_vp(const float *direct, const float *coords, const int *idxs, float *hypo, int tn, int vn, int hn,
                                  int arith, cudaStream_t st) {
    const int n = hn * vn;
    if (n == 0) return FPC_OK;
    if (arith == FPC_ARITH_IEEE)
        k_generate_hypothesis_vp<FPC_ARITH_IEEE><<<ceil_div(n, 256), 256, 0, st>>>(direct, coords, idxs, hypo, tn, vn, hn);
    else
        k_generate_hypothesis_vp<FPC_ARITH_NVCC_FMA><<<ceil_div(n, 256), 256, 0, st>>>(direct, coords, idxs, hypo, tn, vn, hn);
    FPC_LAUNCH_CHECK("k_generate_hypothesis_vp");
    return FPC_OK;
}

int launch_voting_for_hypothesis_vp(const float *direct, const float *coords, const float *hypo, uint8_t *inliers, int tn, int vn,
                                    int hn, float thresh, int arith, cudaStream_t st) {
    if (tn == 0 || vn == 0 || hn == 0) return FPC_OK;
    dim3 grid(ceil_div(tn, 256), vn, ceil_div(hn, K2_HYPS_PER_BLOCK));
    if (arith == FPC_ARITH_IEEE)
        k_voting_for_hypothesis_vp<FPC_ARITH_IEEE><<<grid, 256, 0, st>>>(direct, coords, hypo, inliers, tn, vn, hn, thresh);
    else
        k_voting_for_hypothesis_vp<FPC_ARITH_NVCC_FMA><<<grid, 256, 0, st>>>(direct, coords, hypo, inliers, tn, vn, hn, thresh);
    FPC_LAUNCH_CHECK("k_voting_for_hypothesis_vp");
    return FPC_OK;
}

// Inner-loop variant: 1 = packed f32x2 arithmetic (FFMA2/FMUL2/FADD2, Blackwell only), 0 = scalar.
// FPC_VOTE_PACKED=0 in the environment selects the scalar loop (kept for A/B measurements).
static int g_vote_packed = -1;
void set_vote_packed(int v) { g_vote_packed = v; }
static int vote_packed() {
    if (g_vote_packed < 0) {
        const char *e = getenv("FPC_VOTE_PACKED");
        g_vote_packed = (e && e[0] == '0') ? 0 : 1;
    }
    return g_vote_packed;
}

// Constants of the tangent-form fast test (see the comment above k_vote and DESIGN.md), computed in double.
static VoteConsts vote_consts(const PathParams &pp) {
    VoteConsts vc;
    vc.nb = (pp.hn + VHB - 1) / VHB;
    vc.all_exact = 1;
    { const char *e = getenv("FPC_VOTE_DEBUG_SKIP"); vc.debug_skip = e ? atoi(e) : 0; }
    vc.ntau = 0.f;
    vc.half_w = 0.f;
    vc.one_plus_tlo = 1.f;
    const double t = (double)pp.inlier_thresh;
    const double u = ldexp(1.0, -24);
    if (!(t > 1e-3) || !(t < 1.0)) return vc;                 // outside the fast test's domain: settle every vote exactly
    // reference's rounded cosine vs the cosine of the exact d = h - c, first order in u: dot (1 + 1/t) u (two products,
    // one sum; |dx nx| + |dy ny| <= |d||n|), each norm 2u, their product u, the quotient u, and fl(h - c) turns d by at
    // most u rad = T(t) u relative in the cosine; 2 % for the second-order terms
    const double eps_r = 1.02 * (7.0 + 1.0 / t + sqrt(1.0 - t * t) / t) * u;
    const double c_hi = t * (1.0 + eps_r), c_lo = t * (1.0 - eps_r);
    if (!(c_hi < 1.0)) return vc;
    auto T = [](double c) { return sqrt(1.0 - c * c) / c; };
    const double t_hi = T(c_hi), t_lo = T(c_lo);              // t_hi < t_lo
    const float tau = (float)(0.5 * (t_hi + t_lo));
    if (!((double)tau > t_hi && (double)tau < t_lo)) return vc;   // interval narrower than binary32 resolves
    float hw = (float)std::max(t_lo - (double)tau, (double)tau - t_hi);
    if ((double)hw < std::max(t_lo - (double)tau, (double)tau - t_hi)) hw = nextafterf(hw, INFINITY);
    float opt = (float)(1.0 + t_lo);
    if ((double)opt < 1.0 + t_lo) opt = nextafterf(opt, INFINITY);
    vc.ntau = -tau;
    vc.half_w = hw;
    vc.one_plus_tlo = opt;
    vc.all_exact = 0;
    return vc;
}

template <int ARITH>
static int launch_hypotheses_t(const Workspace &ws, const PathParams &pp, float2 *hyp_out, cudaStream_t st) {
    const VoteConsts vc = vote_consts(pp);
    k_hypotheses<ARITH><<<sm_count() * 8, 256, 0, st>>>(ws.T, ws.counters, pp, ws.rec, hyp_out, ws.hloc, ws.work, ws.workf, vc);
    FPC_LAUNCH_CHECK("k_hypotheses");
    return FPC_OK;
}

template <int ARITH, bool PACKED>
static int launch_vote_t(const Workspace &ws, const PathParams &pp, const float2 *hyp, int *votes, cudaStream_t st) {
    static thread_local int blocks_per_sm = 0;
    const int smem = (int)sizeof(VoteSmem);
    if (blocks_per_sm == 0) {
        FPC_CUDA_TRY(cudaFuncSetAttribute(k_vote<ARITH, PACKED>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vote<ARITH, PACKED>, VT, smem) != cudaSuccess || n < 1) n = 2;
        // Two blocks per SM (that is also what the shared-memory buffers allow): the hot loop saturates the issue ports with two
        // warps per scheduler, and the rest of the SM is worth more to the kernels of the other batches in flight.
        // FPC_VOTE_BLOCKS_PER_SM=1 overrides (experiments).
        const char *e = getenv("FPC_VOTE_BLOCKS_PER_SM");
        n = std::min(n, (e && atoi(e) >= 1) ? atoi(e) : 2);
        blocks_per_sm = n;
    }
    const VoteConsts vc = vote_consts(pp);
    k_vote<ARITH, PACKED><<<sm_count() * blocks_per_sm, VT, smem, st>>>(ws.T, ws.counters, pp, ws.rec, hyp, ws.hloc, votes, ws.work, ws.workf, ws.segs, ws.segcnt, vc);
    FPC_LAUNCH_CHECK("k_vote");
    k_vote_settle<ARITH><<<sm_count() * blocks_per_sm * SETTLE_SPLIT, 256, 0, st>>>(ws.counters, pp, ws.rec, hyp, ws.hloc, votes, ws.segs, ws.segcnt,
                                                          sm_count() * blocks_per_sm, vc);
    FPC_LAUNCH_CHECK("k_vote_settle");
    return FPC_OK;
}

int launch_vote(const Workspace &ws, const PathParams &pp, float2 *hyp_out, int *votes, cudaStream_t st) {
    int rc = pp.arith == FPC_ARITH_IEEE ? launch_hypotheses_t<FPC_ARITH_IEEE>(ws, pp, hyp_out, st)
                                        : launch_hypotheses_t<FPC_ARITH_NVCC_FMA>(ws, pp, hyp_out, st);
    if (rc != FPC_OK) return rc;
    if (pp.arith == FPC_ARITH_IEEE)
        return vote_packed() ? launch_vote_t<FPC_ARITH_IEEE, true>(ws, pp, hyp_out, votes, st)
                             : launch_vote_t<FPC_ARITH_IEEE, false>(ws, pp, hyp_out, votes, st);
    return vote_packed() ? launch_vote_t<FPC_ARITH_NVCC_FMA, true>(ws, pp, hyp_out, votes, st)
                         : launch_vote_t<FPC_ARITH_NVCC_FMA, false>(ws, pp, hyp_out, votes, st);
}

int vote_batches(int hn) { return (hn + VHB - 1) / VHB; }
int vote_seg_cap() { return VSEGCAP; }

int vote_tail_div() {
    static int v = 0;
    if (v == 0) {
        const char *e = getenv("FPC_VOTE_TAIL_DIV");
        v = e ? std::max(1, std::min(8, atoi(e))) : 1;
    }
    return v;
}

// Pixels per vote work item: P/2048 rounded down to a power of two, clamped to [128, FPC_VOTE_ITEM_PX].  The upper clamp
// (default 1024 = the buffer size, FPC_VOTE_ITEM_PX in the environment, at most VOTE_CHUNK = the shared-memory buffer) trades per-item overhead
// against the tail of the ticket queue: with ~5 items per resident block a block that draws one item more than its
// neighbours finishes 20 % later; but every item costs ~350 instructions per warp of set-up, and with several batches in
// flight the idle tail of one batch is filled by the kernels of the next, so fewer, larger items win (measured).
int vote_chunk_for(long long P, int hn) {
    static int cap = 0;
    if (cap == 0) {
        const char *e = getenv("FPC_VOTE_ITEM_PX");
        int v = e ? atoi(e) : 1024;
        int c = 128;
        while (c * 2 <= v && c * 2 <= VOTE_CHUNK) c *= 2;
        cap = c;
    }
    int c = cap;
    while (c > 128 && (long long)c * 2048 > P) c >>= 1;
    // the note table holds (VHALF-pixel units) x hypothesis groups (padded to 1, 2 or 4) <= VNOTES / 32 per item: 1024 px for
    // hn <= 256, 512 px beyond
    const int G = (std::min(hn, VHB) + 127) / 128, Gp = G > 2 ? 4 : G;
    while (c > 128 && Gp * (c / VHALF) > VNOTES / 32) c >>= 1;
    return c;
}

int launch_finalize(const Workspace &ws, const PathParams &pp, const float2 *hyp, const int *votes, const float *inv_k,
                    float *pose_table, cudaStream_t st) {
    const int grid = sm_count() * 8;
    if (pp.arith == FPC_ARITH_IEEE)
        k_finalize<FPC_ARITH_IEEE><<<grid, FT, 0, st>>>(ws.T, ws.R, ws.counters, pp, ws.rec, hyp, votes, inv_k, pose_table);
    else
        k_finalize<FPC_ARITH_NVCC_FMA><<<grid, FT, 0, st>>>(ws.T, ws.R, ws.counters, pp, ws.rec, hyp, votes, inv_k, pose_table);
    FPC_LAUNCH_CHECK("k_finalize");
    return FPC_OK;
}

int launch_get_rt(const float *q, const float *xy, const float *z, const float *inv_k, float *R, float *T, float *RT,
                  int n, cudaStream_t st) {
    if (n == 0) return FPC_OK;
    k_get_rt<<<ceil_div(n, 128), 128, 0, st>>>(q, xy, z, inv_k, R, T, RT, n);
    FPC_LAUNCH_CHECK("k_get_rt");
    return FPC_OK;
}

}  // namespace fpc
