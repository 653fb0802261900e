// Ground-truth <-> prediction matching: lib/matching.py:226-325 batchwise_find_matches and
// lib/gpu_tensor_funcs.py:386-409 batchwise_get_2d_iou.
//
// The reference expands both mask sets to [n1,n2,h,w] and sums logical_and / logical_or: O(n1*n2*h*w) element work
// and traffic.  Here every mask is read ONCE and packed to one bit per pixel together with its pixel count and bounding
// box (k_pack_masks: HBM-bound, 4 B/px in, 1/8 B/px out; or k_pack_labels straight from the label volume of the fused
// path, which never builds dense masks).  A pair's intersection is then popc(a & b) over the overlap of the two
// bounding boxes (disjoint boxes cost nothing), the union follows by inclusion-exclusion, and
// IoU = float(inter) / float(union) is the same correctly-rounded fp32 quotient torch's int64 true-division produces
// -- so IoU values, the first-maximum pairing and the match order are bit-exact, not approximately equal.
#include <algorithm>

#include "fpc_common.cuh"

namespace fpc {
namespace {

constexpr int META = FPC_MASK_META;   // int32 words per mask: count, ymin, ymax, wmin, wmax (word columns), 3 spare

__global__ void __launch_bounds__(256) k_mask_meta_init(int *__restrict__ meta, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int *m = meta + (size_t)i * META;
    m[0] = 0; m[1] = INT_MAX; m[2] = -1; m[3] = INT_MAX; m[4] = -1; m[5] = 0; m[6] = 0; m[7] = 0;
}

__device__ __forceinline__ void meta_add_row(int *__restrict__ m, int count, int y, int wlo, int whi) {
    atomicAdd(m + 0, count);
    atomicMin(m + 1, y);
    atomicMax(m + 2, y);
    atomicMin(m + 3, wlo);
    atomicMax(m + 4, whi);
}

// One warp per image row of one mask; U coalesced loads in flight per lane, one ballot per 32 pixels.
template <typename T, int U>
__global__ void __launch_bounds__(256) k_pack_masks(const T *__restrict__ masks, long long rows, int h, int w, int wpr,
                                                    unsigned *__restrict__ bits, int *__restrict__ meta) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
        const T *__restrict__ src = masks + r * w;
        unsigned *__restrict__ dst = bits + r * wpr;
        int count = 0, wlo = INT_MAX, whi = -1;
        for (int base = 0; base < w; base += 32 * U) {
            T v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int x = base + 32 * k + lane;
                v[k] = x < w ? __ldcs(src + x) : T(0);
            }
            unsigned mine = 0;
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const unsigned word = __ballot_sync(FULL, v[k] != T(0));   // non-zero = set (NaN included), as logical_and does
                if (word) {
                    count += __popc(word);
                    wlo = min(wlo, (base >> 5) + k);
                    whi = (base >> 5) + k;
                }
                if (lane == k) mine = word;
            }
            const int wi = (base >> 5) + lane;
            if (lane < U && wi < wpr) dst[wi] = mine;
        }
        if (lane == 0 && count) meta_add_row(meta + (r / h) * META, count, (int)(r % h), wlo, whi);
    }
}

// The fast variant for w % 4 == 0 and 16-byte aligned planes: 4 pixels per load (float4 / uchar4), up to U loads
// (U x 512 pixels) in flight per warp -- the whole 640-px row at once.  A lane's 4 pixels make a nibble; 8 lanes make a word.
template <typename T> struct Vec4;
template <> struct Vec4<float> { using type = float4; };
template <> struct Vec4<unsigned char> { using type = uchar4; };

template <typename T, int U>
__global__ void __launch_bounds__(256) k_pack_masks_v4(const T *__restrict__ masks, long long rows, int h, int w, int wpr,
                                                       unsigned *__restrict__ bits, int *__restrict__ meta) {
    using V = typename Vec4<T>::type;
    const int lane = threadIdx.x & 31;
    const int shift = 4 * (lane & 7);
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
        const V *__restrict__ src = reinterpret_cast<const V *>(masks + r * w);
        unsigned *__restrict__ dst = bits + r * wpr;
        int count = 0, wlo = INT_MAX, whi = -1;
        for (int base = 0; base < w; base += 128 * U) {
            V v[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int x = base + 128 * k + 4 * lane;
                v[k] = x < w ? __ldcs(src + (x >> 2)) : V{};
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                unsigned word = ((v[k].x != T(0)) | ((v[k].y != T(0)) << 1) | ((v[k].z != T(0)) << 2) | ((v[k].w != T(0)) << 3)) << shift;
                word |= __shfl_xor_sync(FULL, word, 1);
                word |= __shfl_xor_sync(FULL, word, 2);
                word |= __shfl_xor_sync(FULL, word, 4);
                const int wi = ((base + 128 * k) >> 5) + (lane >> 3);
                if ((lane & 7) == 0 && wi < wpr) {
                    dst[wi] = word;
                    if (word) {
                        count += __popc(word);
                        wlo = min(wlo, wi);
                        whi = max(whi, wi);
                    }
                }
            }
        }
        count = __reduce_add_sync(FULL, count);
        if (count) {                                  // warp-uniform
            wlo = __reduce_min_sync(FULL, wlo);
            whi = __reduce_max_sync(FULL, whi);
            if (lane == 0) meta_add_row(meta + (r / h) * META, count, (int)(r % h), wlo, whi);
        }
    }
}

// Label volume [b,h,w] (0 = background, id = 1 + instance index, as fpc_pose_recover writes it) -> the same bit planes.
// `bits` is zero-filled beforehand; a (mask, row, word) cell belongs to exactly one warp iteration, so plain stores.
__global__ void __launch_bounds__(256) k_pack_labels(const int *__restrict__ labels, long long rows, int h, int w, int wpr, int n,
                                                     unsigned *__restrict__ bits, int *__restrict__ meta) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
        const int *__restrict__ src = labels + r * w;
        const int y = (int)(r % h);
        constexpr int U = 5;                                   // loads in flight per lane (a 640-px row = 4 batches)
        for (int base = 0; base < wpr; base += U) {
            int L[U];
#pragma unroll
            for (int k = 0; k < U; ++k) {
                const int x = (base + k) * 32 + lane;
                L[k] = (base + k < wpr && x < w) ? __ldg(src + x) : 0;
            }
#pragma unroll
            for (int k = 0; k < U; ++k) {
                if (!__any_sync(FULL, L[k] != 0)) continue;
                const unsigned group = __match_any_sync(FULL, L[k]);
                if (L[k] > 0 && L[k] <= n && (__ffs(group) - 1) == lane) {
                    const int wi = base + k;
                    bits[((size_t)(L[k] - 1) * h + y) * wpr + wi] = group;
                    meta_add_row(meta + (size_t)(L[k] - 1) * META, __popc(group), y, wi, wi);
                }
            }
        }
    }
}

struct MaskSet {
    const unsigned *bits;
    const int *meta;
};

// Cheap per-lane test before any bit plane is touched: both masks non-empty and their bounding boxes overlap.
__device__ __forceinline__ bool may_intersect(const int *__restrict__ ma, const int *__restrict__ mb) {
    return ma[0] != 0 && mb[0] != 0 && max(ma[1], mb[1]) <= min(ma[2], mb[2]) && max(ma[3], mb[3]) <= min(ma[4], mb[4]);
}

// |A_i and B_j| by the whole warp: popc(a & b) over the overlap of the two bounding boxes (in rows x 32-px words).
__device__ __forceinline__ int warp_intersection(const MaskSet A, int i, const MaskSet B, int j, int h, int wpr, int lane) {
    const int *ma = A.meta + (size_t)i * META, *mb = B.meta + (size_t)j * META;
    const int y0 = max(ma[1], mb[1]), y1 = min(ma[2], mb[2]);
    const int x0 = max(ma[3], mb[3]), x1 = min(ma[4], mb[4]);
    const int nwx = x1 - x0 + 1, total = (y1 - y0 + 1) * nwx;
    const unsigned *pa = A.bits + ((size_t)i * h + y0) * wpr + x0;
    const unsigned *pb = B.bits + ((size_t)j * h + y0) * wpr + x0;
    int inter = 0;
    for (int t = lane; t < total; t += 32) {
        const int dy = t / nwx, dx = t - dy * nwx;
        inter += __popc(__ldg(pa + dy * wpr + dx) & __ldg(pb + dy * wpr + dx));
    }
    return __reduce_add_sync(FULL, inter);
}

// torch: int64 / int64 -> both sides to float32, IEEE divide (0/0 = NaN).
__device__ __forceinline__ float iou_value(int inter, int ca, int cb) {
    return __fdiv_rn(__int2float_rn(inter), __int2float_rn(ca + cb - inter));
}

// 32 consecutive (i,j) pairs per warp: every lane screens its own pair (most are disjoint and are written at once as
// 0, or NaN for two empty masks); the warp then works through the surviving pairs together.
__global__ void __launch_bounds__(256) k_mask_iou(MaskSet A, int na, MaskSet B, int nb, int h, int wpr, float *__restrict__ iou) {
    const int lane = threadIdx.x & 31;
    const long long pairs = (long long)na * nb;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long p0 = (((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5) * 32; p0 < pairs; p0 += nwarps * 32) {
        const long long p = p0 + lane;
        int i = 0, j = 0, ca = 0, cb = 0, inter = 0;
        bool cand = false;
        if (p < pairs) {
            i = (int)(p / nb);
            j = (int)(p - (long long)i * nb);
            const int *ma = A.meta + (size_t)i * META, *mb = B.meta + (size_t)j * META;
            ca = ma[0];
            cb = mb[0];
            cand = may_intersect(ma, mb);
        }
        unsigned todo = __ballot_sync(FULL, cand);
        while (todo) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1;
            const int v = warp_intersection(A, __shfl_sync(FULL, i, src), B, __shfl_sync(FULL, j, src), h, wpr, lane);
            if (lane == src) inter = v;
        }
        if (p < pairs) iou[p] = iou_value(inter, ca, cb);
    }
}

// lib/matching.py:253-296 for every class at once: block per ground-truth mask.  Each warp screens 32 predictions at a
// time (same class -- ANY frame: the reference never compares sample ids -- and overlapping boxes), then intersects
// the survivors in ascending order; first maximum wins, kept only if > 0.
constexpr int MATCH_WARPS = 4;
__global__ void __launch_bounds__(MATCH_WARPS * 32) k_match_best(MaskSet G, const long long *__restrict__ class_g, MaskSet P,
                                                                 const long long *__restrict__ class_p, int np, int h, int wpr,
                                                                 int *__restrict__ best_pred, float *__restrict__ best_iou) {
    __shared__ float s_v[MATCH_WARPS];
    __shared__ int s_j[MATCH_WARPS];
    const int i = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const long long c = class_g[i];
    const int *mg = G.meta + (size_t)i * META;
    const int cg = mg[0];
    float bv = 0.f;
    int bj = -1;
    for (int j0 = wid * 32; j0 < np; j0 += MATCH_WARPS * 32) {
        const int j = j0 + lane;
        const bool cand = j < np && class_p[j] == c && may_intersect(mg, P.meta + (size_t)j * META);
        unsigned todo = __ballot_sync(FULL, cand);
        while (todo) {                                  // ascending j: strict > keeps the first maximum
            const int jj = j0 + __ffs(todo) - 1;
            todo &= todo - 1;
            const int inter = warp_intersection(G, i, P, jj, h, wpr, lane);
            if (inter == 0) continue;                   // IoU 0 never passes the "> 0" rule
            const float v = iou_value(inter, cg, P.meta[(size_t)jj * META]);
            if (v > bv) { bv = v; bj = jj; }
        }
    }
    if (lane == 0) { s_v[wid] = bv; s_j[wid] = bj; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < MATCH_WARPS; ++k)
            if (s_j[k] >= 0 && (s_v[k] > bv || (s_v[k] == bv && s_j[k] < bj))) { bv = s_v[k]; bj = s_j[k]; }
        best_pred[i] = bj;
        best_iou[i] = bv;
    }
}

// Output order of the reference: classes ascending (torch.unique, :253), ground-truth index ascending inside a class.
__global__ void __launch_bounds__(256) k_match_order(const long long *__restrict__ class_g, const int *__restrict__ best_pred, int ng,
                                                     int *__restrict__ pairs, int *__restrict__ n_matches) {
    // warp per matched ground truth: its rank = number of matched ground truths that sort before it
    const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (i >= ng || best_pred[i] < 0) return;               // warp-uniform
    const long long c = class_g[i];
    int rank = 0;
    for (int k = lane; k < ng; k += 32) {
        const long long ck = class_g[k];
        rank += (best_pred[k] >= 0) && (ck < c || (ck == c && k < i));
    }
    rank = __reduce_add_sync(FULL, rank);
    if (lane == 0) {
        pairs[2 * rank] = i;
        pairs[2 * rank + 1] = best_pred[i];
        atomicAdd(n_matches, 1);
    }
}

// Dense [m,h,w] f32 0/1 masks of the listed instances (row k: instance inst_of[k] of frame frame_of[k]) painted from the
// label volume -- the matched predictions' masks, the only dense masks a label-volume prediction set ever needs.
__global__ void __launch_bounds__(256) k_paint_instances(const int *__restrict__ labels, int b, long long hw, const long long *__restrict__ frame_of,
                                                         const long long *__restrict__ inst_of, int m, float *__restrict__ out) {
    const int hw4 = (int)(hw >> 2);
    for (int k = blockIdx.y; k < m; k += gridDim.y) {                     // one output plane per blockIdx.y: no per-thread division
        const long long f = frame_of[k];
        const int want = (int)inst_of[k] + 1;
        const bool live = f >= 0 && f < b;
        const int4 *__restrict__ src = reinterpret_cast<const int4 *>(labels + (live ? f : 0) * hw);
        float4 *__restrict__ dst = reinterpret_cast<float4 *>(out + (size_t)k * hw);
        for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < hw4; q += gridDim.x * blockDim.x) {
            float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
            if (live) {
                const int4 L = __ldg(src + q);
                o = make_float4(L.x == want ? 1.f : 0.f, L.y == want ? 1.f : 0.f, L.z == want ? 1.f : 0.f, L.w == want ? 1.f : 0.f);
            }
            __stcs(dst + q, o);
        }
    }
}
__global__ void __launch_bounds__(256) k_paint_instances_scalar(const int *__restrict__ labels, int b, long long hw, const long long *__restrict__ frame_of,
                                                                const long long *__restrict__ inst_of, long long total, float *__restrict__ out) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long k = t / hw, q = t - k * hw;
        const long long f = frame_of[k];
        out[t] = (f >= 0 && f < b && labels[f * hw + q] == (int)inst_of[k] + 1) ? 1.f : 0.f;
    }
}

int grid_for_warps(long long warps) {
    return (int)std::max<long long>(1, std::min<long long>(ceil_div_ll(warps, 8), (long long)sm_count() * 8));
}

int check_plane(int n, int h, int w) {
    if (n < 0 || h <= 0 || w <= 0) return fail(FPC_EINVAL, "bad mask size n=%d h=%d w=%d", n, h, w);
    return FPC_OK;
}

}  // namespace
}  // namespace fpc

using namespace fpc;

extern "C" {

int fpc_pack_masks(const void *masks, int elem, int n, int h, int w, uint32_t *bits, int32_t *meta, void *stream) {
    if (check_plane(n, h, w) != FPC_OK) return FPC_EINVAL;
    if (n == 0) return FPC_OK;
    if (!masks || !bits || !meta) return fail(FPC_EINVAL, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int wpr = (w + 31) / 32;
    const long long rows = (long long)n * h;
    k_mask_meta_init<<<ceil_div(n, 256), 256, 0, st>>>(meta, n);
    FPC_LAUNCH_CHECK("k_mask_meta_init");
    if (elem != FPC_MASK_F32 && elem != FPC_MASK_U8) return fail(FPC_EINVAL, "unknown mask element kind %d", elem);
    const size_t esz = elem == FPC_MASK_F32 ? 4 : 1;
    const bool vec = (w % 4 == 0) && ((uintptr_t)masks % (4 * esz) == 0);
    const int grid = grid_for_warps(rows);
    if (elem == FPC_MASK_F32) {
        if (vec) k_pack_masks_v4<float, 5><<<grid, 256, 0, st>>>((const float *)masks, rows, h, w, wpr, bits, meta);
        else k_pack_masks<float, 8><<<grid, 256, 0, st>>>((const float *)masks, rows, h, w, wpr, bits, meta);
    } else {
        if (vec) k_pack_masks_v4<unsigned char, 5><<<grid, 256, 0, st>>>((const unsigned char *)masks, rows, h, w, wpr, bits, meta);
        else k_pack_masks<unsigned char, 8><<<grid, 256, 0, st>>>((const unsigned char *)masks, rows, h, w, wpr, bits, meta);
    }
    FPC_LAUNCH_CHECK("k_pack_masks");
    return FPC_OK;
}

int fpc_pack_labels(const int32_t *labels, int b, int h, int w, int n, uint32_t *bits, int32_t *meta, void *stream) {
    if (check_plane(n, h, w) != FPC_OK || b < 0) return fail(FPC_EINVAL, "bad size");
    if (n == 0) return FPC_OK;
    if (!labels || !bits || !meta) return fail(FPC_EINVAL, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int wpr = (w + 31) / 32;
    FPC_CUDA_TRY(cudaMemsetAsync(bits, 0, (size_t)n * h * wpr * sizeof(uint32_t), st));
    k_mask_meta_init<<<ceil_div(n, 256), 256, 0, st>>>(meta, n);
    FPC_LAUNCH_CHECK("k_mask_meta_init");
    if (b == 0) return FPC_OK;
    const long long rows = (long long)b * h;
    k_pack_labels<<<grid_for_warps(rows), 256, 0, st>>>(labels, rows, h, w, wpr, n, bits, meta);
    FPC_LAUNCH_CHECK("k_pack_labels");
    return FPC_OK;
}

int fpc_mask_iou(const uint32_t *bits_a, const int32_t *meta_a, int na, const uint32_t *bits_b, const int32_t *meta_b, int nb,
                 int h, int w, float *iou, void *stream) {
    if (check_plane(na, h, w) != FPC_OK || nb < 0) return fail(FPC_EINVAL, "bad size");
    if (na == 0 || nb == 0) return FPC_OK;
    if (!bits_a || !meta_a || !bits_b || !meta_b || !iou) return fail(FPC_EINVAL, "NULL pointer");
    const long long pairs = (long long)na * nb;
    k_mask_iou<<<grid_for_warps(pairs), 256, 0, (cudaStream_t)stream>>>(MaskSet{bits_a, meta_a}, na, MaskSet{bits_b, meta_b}, nb, h,
                                                                          (w + 31) / 32, iou);
    FPC_LAUNCH_CHECK("k_mask_iou");
    return FPC_OK;
}

int fpc_match_instances(const uint32_t *bits_g, const int32_t *meta_g, const int64_t *class_g, int ng, const uint32_t *bits_p,
                        const int32_t *meta_p, const int64_t *class_p, int np, int h, int w, int32_t *best_pred, float *best_iou,
                        int32_t *pairs, int32_t *n_matches, void *stream) {
    if (check_plane(ng, h, w) != FPC_OK || np < 0) return fail(FPC_EINVAL, "bad size");
    if (!n_matches) return fail(FPC_EINVAL, "NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    FPC_CUDA_TRY(cudaMemsetAsync(n_matches, 0, sizeof(int32_t), st));
    if (ng == 0) return FPC_OK;
    if (!bits_g || !meta_g || !class_g || !best_pred || !best_iou || !pairs || (np > 0 && (!bits_p || !meta_p || !class_p)))
        return fail(FPC_EINVAL, "NULL pointer");
    k_match_best<<<ng, MATCH_WARPS * 32, 0, st>>>(MaskSet{bits_g, meta_g}, (const long long *)class_g, MaskSet{bits_p, meta_p},
                                                  (const long long *)class_p, np, h, (w + 31) / 32, best_pred, best_iou);
    FPC_LAUNCH_CHECK("k_match_best");
    k_match_order<<<ceil_div(ng, 8), 256, 0, st>>>((const long long *)class_g, best_pred, ng, pairs, n_matches);
    FPC_LAUNCH_CHECK("k_match_order");
    return FPC_OK;
}

int fpc_paint_instances(const int32_t *labels, int b, int h, int w, const int64_t *frame_of, const int64_t *inst_of, int m,
                        float *out, void *stream) {
    if (check_plane(m, h, w) != FPC_OK || b < 0) return fail(FPC_EINVAL, "bad size");
    if (m == 0) return FPC_OK;
    if (!labels || !frame_of || !inst_of || !out) return fail(FPC_EINVAL, "NULL pointer");
    const long long hw = (long long)h * w, total = hw * m;
    const bool vec = hw % 4 == 0 && (uintptr_t)labels % 16 == 0 && (uintptr_t)out % 16 == 0;
    const long long items = vec ? total / 4 : total;
    const int grid = (int)std::min<long long>(ceil_div_ll(items, 256), (long long)sm_count() * 16);
    if (vec) {
        const dim3 g2((unsigned)std::min<long long>(ceil_div_ll(hw / 4, 256), 64), (unsigned)std::min(m, 65535));
        k_paint_instances<<<g2, 256, 0, (cudaStream_t)stream>>>(labels, b, hw, (const long long *)frame_of, (const long long *)inst_of, m, out);
    } else
        k_paint_instances_scalar<<<grid, 256, 0, (cudaStream_t)stream>>>(labels, b, hw, (const long long *)frame_of, (const long long *)inst_of, items, out);
    FPC_LAUNCH_CHECK("k_paint_instances");
    return FPC_OK;
}

}  // extern "C"
