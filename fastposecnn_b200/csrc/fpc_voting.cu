// Voting half of the path: hypothesis generation from pre-sampled pixel pairs, inlier vote
// counting, winner selection, inlier refinement, and the per-instance pose finalisation.
//
// Reference behaviour being reproduced (paths relative to /root/reference/source_code/FastPoseCNN/):
//   lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu:11-49    K1 generate_hypothesis
//   lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu:88-126   K2 voting_for_hypothesis
//   lib/ransac_voting_gpu_layer/ransac_voting_gpu.py:518-607         ransac_voting_layer_v3 driver
//   lib/gpu_tensor_funcs.py:204-253, 306-326                         translation / rotation / RT
#include "fpc_internal.cuh"

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
// Fused hypothesis generation + vote counting
// =============================================================================================
// Work item = (instance, chunk of <= VOTE_CHUNK voting records), handed out through an atomic ticket
// so every resident block stays busy until the work runs out.  The chunk's pixels are staged in
// shared memory once as six SoA planes (-x, -y, dir_x, dir_y, k_hi, k_lo); every lane keeps VQ
// hypotheses in registers and walks the pixels with broadcast LDS.128 loads (4 pixels per load).
// No inlier matrix is ever written (the reference materialises hn*tn bytes and re-reads them,
// ransac_voting_gpu.py:562-566).
//
// Exactness.  The reference decides   dot(d,n) / (|n| |d|) > thresh   in rounded binary32 ops.
// The fast test compares  s = (d.n)|d.n|  with  |d|^2 * (|n| thresh)^2 * (1 +- e),  e = 2^-18 = 64 u:
//   hi = s - k_hi |d|^2 >= 0  -> certainly an inlier,   lo = s - k_lo |d|^2 < 0 -> certainly not.
// Worst-case relative rounding error of the fast ratio is 12 u and of the reference's squared cosine
// 16 u, so the two decisions can only differ inside the band (28 u < e; derivation in DESIGN.md).
// Votes that land inside the band (a few per 10^4) are queued in shared memory and settled afterwards
// by all lanes in parallel with the reference expression itself (explicitly rounded intrinsics), as
// are all votes of hypotheses within 1e-3 of a pixel-lattice point (where |d| may fall under the
// reference's 1e-6 guard).  Both sign bits of every vote are collected with funnel shifts, so the hot
// loop has no branch.
constexpr int VT = 256;            // threads per block
constexpr int VQ = 4;              // hypotheses per lane
constexpr int VHB = 1024;          // hypotheses per shared-memory batch (8 groups of 32 lanes x VQ)
constexpr int VROUND = 16;         // pixels per sign-collection round (2 bits per vote in a 32-bit word)
constexpr int VQCAP = 2048;        // deferred exact-vote queue entries
constexpr float BAND_E = 3.814697265625e-06f;  // 2^-18

typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
    u64 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpk2(u64 v, float &lo, float &hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 add2(u64 a, u64 b) {
    u64 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 mul2(u64 a, u64 b) {
    u64 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
    u64 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

__device__ __forceinline__ bool near_lattice(float x, float y) {
    return fabsf(x - rintf(x)) < 1e-3f && fabsf(y - rintf(y)) < 1e-3f;
}

struct VoteSmem {
    float ncx[VOTE_CHUNK], ncy[VOTE_CHUNK], nx[VOTE_CHUNK], ny[VOTE_CHUNK], khi[VOTE_CHUNK], klo[VOTE_CHUNK];
    float2 hyp[VHB];
    unsigned queue[VQCAP];
    unsigned short exlist[VHB];
    int qn, nex, wi;
};

// one vote on the fast path: appends sign(hi), sign(lo) to acc
__device__ __forceinline__ void vote_fast(float hx, float hy, float ncx, float ncy, float nx, float ny, float khi,
                                          float klo, unsigned &acc) {
    const float dx = hx + ncx, dy = hy + ncy;
    const float d2 = fmaf(dy, dy, dx * dx);
    const float dt = fmaf(dy, ny, dx * nx);
    const float s = dt * fabsf(dt);
    const float hi = fmaf(d2, khi, s);
    const float lo = fmaf(d2, klo, s);
    acc = __funnelshift_l(__float_as_uint(hi), acc, 1);
    acc = __funnelshift_l(__float_as_uint(lo), acc, 1);
}
// two pixels (a, b) against one hypothesis, packed f32x2 arithmetic
__device__ __forceinline__ void vote_fast2(u64 hx2, u64 hy2, u64 ncx2, u64 ncy2, u64 nx2, u64 ny2, u64 khi2, u64 klo2,
                                           unsigned &acc) {
    const u64 dx = add2(hx2, ncx2), dy = add2(hy2, ncy2);
    const u64 d2 = fma2(dy, dy, mul2(dx, dx));
    const u64 dt = fma2(dy, ny2, mul2(dx, nx2));
    float dta, dtb;
    unpk2(dt, dta, dtb);
    const u64 s = pk2(dta * fabsf(dta), dtb * fabsf(dtb));
    const u64 hi = fma2(d2, khi2, s), lo = fma2(d2, klo2, s);
    float hia, hib, loa, lob;
    unpk2(hi, hia, hib);
    unpk2(lo, loa, lob);
    acc = __funnelshift_l(__float_as_uint(hia), acc, 1);
    acc = __funnelshift_l(__float_as_uint(loa), acc, 1);
    acc = __funnelshift_l(__float_as_uint(hib), acc, 1);
    acc = __funnelshift_l(__float_as_uint(lob), acc, 1);
}

template <int ARITH, bool PACKED>
__global__ void __launch_bounds__(VT, 3) k_vote(InstTables T, int *__restrict__ counters, PathParams pp,
                                                const float4 *__restrict__ rec, float2 *__restrict__ hyp_g,
                                                int *__restrict__ votes) {
    __shared__ __align__(16) VoteSmem sm;
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    const int W = counters[FPC_CNT_WORK];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    const int hn = pp.hn;
    const float thresh = pp.inlier_thresh;
    const bool all_exact = !(thresh > 0.f);
    const float t2 = thresh * thresh;

    while (true) {
        __syncthreads();  // previous work item is done with shared memory
        if (tid == 0) {
            sm.wi = atomicAdd(&counters[FPC_CNT_TICKET], 1);
            sm.qn = 0;
            sm.nex = 0;
        }
        __syncthreads();
        const int wi = sm.wi;
        if (wi >= W) break;
        const int i = upper_index(T.workoff, N, wi);
        const int chunk = wi - T.workoff[i];
        const int tn = T.tn[i];
        const int px0 = chunk * VOTE_CHUNK;
        const int npx = min(VOTE_CHUNK, tn - px0);
        const int nrounds = (npx + VROUND - 1) / VROUND;
        const float4 *rec_i = rec + T.pxoff[i];
        // ---- stage the chunk's pixels ----
        for (int k = tid; k < nrounds * VROUND; k += VT) {
            float ncx = -1e18f, ncy = -1e18f, nx = 0.f, ny = 0.f, khi = -1.f, klo = -1.f;  // padding: never an inlier
            if (k < npx) {
                const float4 r = rec_i[px0 + k];
                ncx = -r.x; ncy = -r.y; nx = r.z; ny = r.w;
                const float n1sq = sum_prod<ARITH>(nx, nx, ny, ny);
                if (below_1e6(__fsqrt_rn(n1sq))) {
                    khi = -1e30f; klo = -1e30f;      // the reference skips pixels with |n| < 1e-6 (.cu:119)
                } else {
                    const float m = n1sq * t2;
                    khi = -(m * (1.0f + BAND_E));
                    klo = -(m * (1.0f - BAND_E));
                }
            }
            sm.ncx[k] = ncx; sm.ncy[k] = ncy; sm.nx[k] = nx; sm.ny[k] = ny; sm.khi[k] = khi; sm.klo[k] = klo;
        }
        for (int hb = 0; hb < hn; hb += VHB) {
            const int nh = min(VHB, hn - hb);
            const int G = (nh + 127) >> 7;  // groups of 128 hypotheses (32 lanes x VQ)
            // ---- hypotheses of this batch (ransac_voting_kernel.cu:11-49) ----
            for (int k = tid; k < G * 128; k += VT) {
                float2 hp = make_float2(0.f, 0.f);
                if (k < nh) {
                    const int h = hb + k;
                    int t0, t1;
                    if (pp.idxs) {
                        t0 = pp.idxs[((size_t)i * hn + h) * 2];
                        t1 = pp.idxs[((size_t)i * hn + h) * 2 + 1];
                        t0 = min(max(t0, 0), tn - 1);
                        t1 = min(max(t1, 0), tn - 1);
                    } else {
                        t0 = (int)(hash3(pp.seed, (uint32_t)i, (uint32_t)h, 0u) % (uint32_t)tn);
                        t1 = (int)(hash3(pp.seed, (uint32_t)i, (uint32_t)h, 1u) % (uint32_t)tn);
                    }
                    const float4 r0 = rec_i[t0], r1 = rec_i[t1];
                    float x, y;
                    if (hypothesis_exact<ARITH>(r0.z, r0.w, r0.x, r0.y, r1.z, r1.w, r1.x, r1.y, x, y)) hp = make_float2(x, y);
                    if (chunk == 0) hyp_g[(size_t)i * hn + h] = hp;
                    if (all_exact || near_lattice(hp.x, hp.y)) sm.exlist[atomicAdd(&sm.nex, 1)] = (unsigned short)k;
                }
                sm.hyp[k] = hp;
            }
            __syncthreads();
            // ---- fast voting: warp -> (hypothesis group g, every parts-th round) ----
            const int parts = 8 / G;  // G <= 8 because VHB = 8 * 128
            if (wv < G * parts) {
                const int g = wv % G, part = wv / G;
                float hx[VQ], hy[VQ];
                int cnt[VQ];
                bool ex[VQ];
#pragma unroll
                for (int q = 0; q < VQ; ++q) {
                    const int idx = g * 128 + q * 32 + lane;
                    const float2 hp = sm.hyp[idx];
                    hx[q] = hp.x; hy[q] = hp.y;
                    cnt[q] = 0;
                    ex[q] = (idx >= nh) || all_exact || near_lattice(hp.x, hp.y);   // not counted on the fast path
                }
                for (int rd = part; rd < nrounds; rd += parts) {
                    unsigned acc[VQ];
#pragma unroll
                    for (int q = 0; q < VQ; ++q) acc[q] = 0u;
                    const int kb = rd * VROUND;
#pragma unroll
                    for (int j4 = 0; j4 < VROUND; j4 += 4) {
                        const float4 ncx = *reinterpret_cast<const float4 *>(&sm.ncx[kb + j4]);
                        const float4 ncy = *reinterpret_cast<const float4 *>(&sm.ncy[kb + j4]);
                        const float4 nx = *reinterpret_cast<const float4 *>(&sm.nx[kb + j4]);
                        const float4 ny = *reinterpret_cast<const float4 *>(&sm.ny[kb + j4]);
                        const float4 khi = *reinterpret_cast<const float4 *>(&sm.khi[kb + j4]);
                        const float4 klo = *reinterpret_cast<const float4 *>(&sm.klo[kb + j4]);
                        if (PACKED) {
                            const u64 ncxa = pk2(ncx.x, ncx.y), ncxb = pk2(ncx.z, ncx.w);
                            const u64 ncya = pk2(ncy.x, ncy.y), ncyb = pk2(ncy.z, ncy.w);
                            const u64 nxa = pk2(nx.x, nx.y), nxb = pk2(nx.z, nx.w);
                            const u64 nya = pk2(ny.x, ny.y), nyb = pk2(ny.z, ny.w);
                            const u64 kha = pk2(khi.x, khi.y), khb = pk2(khi.z, khi.w);
                            const u64 kla = pk2(klo.x, klo.y), klb = pk2(klo.z, klo.w);
#pragma unroll
                            for (int q = 0; q < VQ; ++q) {
                                const u64 hx2 = pk2(hx[q], hx[q]), hy2 = pk2(hy[q], hy[q]);
                                vote_fast2(hx2, hy2, ncxa, ncya, nxa, nya, kha, kla, acc[q]);
                                vote_fast2(hx2, hy2, ncxb, ncyb, nxb, nyb, khb, klb, acc[q]);
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < VQ; ++q) {
                                vote_fast(hx[q], hy[q], ncx.x, ncy.x, nx.x, ny.x, khi.x, klo.x, acc[q]);
                                vote_fast(hx[q], hy[q], ncx.y, ncy.y, nx.y, ny.y, khi.y, klo.y, acc[q]);
                                vote_fast(hx[q], hy[q], ncx.z, ncy.z, nx.z, ny.z, khi.z, klo.z, acc[q]);
                                vote_fast(hx[q], hy[q], ncx.w, ncy.w, nx.w, ny.w, khi.w, klo.w, acc[q]);
                            }
                        }
                    }
                    // pixel j of the round sits at bits (2*(15-j)+1: sign(hi), 2*(15-j): sign(lo))
#pragma unroll
                    for (int q = 0; q < VQ; ++q) {
                        const unsigned a = acc[q];
                        cnt[q] += __popc(~a & 0xAAAAAAAAu);                 // hi >= 0
                        unsigned b = (a >> 1) & ~a & 0x55555555u;          // hi < 0 and lo >= 0: inside the band
                        if (b && !ex[q]) {
                            const unsigned hidx = (unsigned)(g * 128 + q * 32 + lane);
                            while (b) {
                                const int pos = __ffs(b) - 1;
                                b &= b - 1;
                                const int k = kb + 15 - (pos >> 1);
                                if (k < npx) {
                                    const int slot = atomicAdd(&sm.qn, 1);
                                    if (slot < VQCAP) {
                                        sm.queue[slot] = ((unsigned)k << 16) | hidx;
                                    } else if (vote_exact<ARITH>(-sm.ncx[k], -sm.ncy[k], sm.nx[k], sm.ny[k], hx[q], hy[q], thresh)) {
                                        atomicAdd(&votes[(size_t)i * hn + hb + hidx], 1);   // queue full: settle it right here
                                    }
                                }
                            }
                        }
                    }
                }
#pragma unroll
                for (int q = 0; q < VQ; ++q) {
                    const int idx = g * 128 + q * 32 + lane;
                    if (!ex[q] && cnt[q]) atomicAdd(&votes[(size_t)i * hn + hb + idx], cnt[q]);
                }
            }
            __syncthreads();
            // ---- band votes: the reference expression, one queued vote per thread ----
            const int nq = min(sm.qn, VQCAP);
            for (int e = tid; e < nq; e += VT) {
                const unsigned ent = sm.queue[e];
                const int k = (int)(ent >> 16), hidx = (int)(ent & 0xffffu);
                const float2 hp = sm.hyp[hidx];
                if (vote_exact<ARITH>(-sm.ncx[k], -sm.ncy[k], sm.nx[k], sm.ny[k], hp.x, hp.y, thresh))
                    atomicAdd(&votes[(size_t)i * hn + hb + hidx], 1);
            }
            // ---- hypotheses on (or within 1e-3 of) the pixel lattice: every vote with the reference expression ----
            const int nex = sm.nex;
            for (int e = 0; e < nex; ++e) {
                const int hidx = sm.exlist[e];
                const float2 hp = sm.hyp[hidx];
                int c = 0;
                for (int k = tid; k < npx; k += VT)
                    c += vote_exact<ARITH>(-sm.ncx[k], -sm.ncy[k], sm.nx[k], sm.ny[k], hp.x, hp.y, thresh) ? 1 : 0;
                c = __reduce_add_sync(FULL, c);
                if (lane == 0 && c) atomicAdd(&votes[(size_t)i * hn + hb + hidx], c);
            }
            __syncthreads();
            if (tid == 0) { sm.qn = 0; sm.nex = 0; }
            __syncthreads();
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

template <int ARITH>
__global__ void __launch_bounds__(128) k_finalize(InstTables T, RowTables R, const int *__restrict__ counters,
                                                  PathParams pp, const float4 *__restrict__ rec,
                                                  const float2 *__restrict__ hyp_g, const int *__restrict__ votes,
                                                  const float *__restrict__ inv_k, float *__restrict__ table) {
    __shared__ double s_d[4][13];
    __shared__ int s_i[4][3];
    __shared__ float s_win[2];
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
            for (int h = tid; h < hn; h += 128) {
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
            for (int k = 1; k < 4; ++k)
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
        const float4 *rec_i = rec + T.pxoff[i];
        for (int k = tid; k < tn; k += 128) {
            const float4 r = rec_i[k];
            if (vote_exact<ARITH>(r.x, r.y, r.z, r.w, wx, wy, pp.inlier_thresh)) {
                const double nx = r.w, ny = -(double)r.z;    // normal = (dir_y, -dir_x)
                const double bb = nx * r.x + ny * r.y;
                a00 += nx * nx; a01 += nx * ny; a11 += ny * ny;
                b0 += nx * bb; b1 += ny * bb;
                ++ninl;
            }
        }
        // ---- masked sums: add up the (instance,row) partials in a fixed order
        double sm[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) sm[k] = 0.0;
        for (int r = T.rowoff[i] + tid; r < T.rowoff[i + 1]; r += 128) {
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
            for (int k = 0; k < 13; ++k) t[k] = s_d[0][k] + s_d[1][k] + s_d[2][k] + s_d[3][k];
            const int inl = s_i[0][2] + s_i[1][2] + s_i[2][2] + s_i[3][2];
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

template <int ARITH, bool PACKED>
static int launch_vote_t(const Workspace &ws, const PathParams &pp, float2 *hyp_out, int *votes, cudaStream_t st) {
    static thread_local int blocks_per_sm = 0;
    if (blocks_per_sm == 0) {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_vote<ARITH, PACKED>, VT, 0) != cudaSuccess || n < 1) n = 2;
        blocks_per_sm = n;
    }
    k_vote<ARITH, PACKED><<<sm_count() * blocks_per_sm, VT, 0, st>>>(ws.T, ws.counters, pp, ws.rec, hyp_out, votes);
    FPC_LAUNCH_CHECK("k_vote");
    return FPC_OK;
}

int launch_vote(const Workspace &ws, const PathParams &pp, float2 *hyp_out, int *votes, cudaStream_t st) {
    if (pp.arith == FPC_ARITH_IEEE)
        return vote_packed() ? launch_vote_t<FPC_ARITH_IEEE, true>(ws, pp, hyp_out, votes, st)
                             : launch_vote_t<FPC_ARITH_IEEE, false>(ws, pp, hyp_out, votes, st);
    return vote_packed() ? launch_vote_t<FPC_ARITH_NVCC_FMA, true>(ws, pp, hyp_out, votes, st)
                         : launch_vote_t<FPC_ARITH_NVCC_FMA, false>(ws, pp, hyp_out, votes, st);
}

int launch_finalize(const Workspace &ws, const PathParams &pp, const float2 *hyp, const int *votes, const float *inv_k,
                    float *pose_table, cudaStream_t st) {
    const int grid = sm_count() * 8;
    if (pp.arith == FPC_ARITH_IEEE)
        k_finalize<FPC_ARITH_IEEE><<<grid, 128, 0, st>>>(ws.T, ws.R, ws.counters, pp, ws.rec, hyp, votes, inv_k, pose_table);
    else
        k_finalize<FPC_ARITH_NVCC_FMA><<<grid, 128, 0, st>>>(ws.T, ws.R, ws.counters, pp, ws.rec, hyp, votes, inv_k, pose_table);
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
