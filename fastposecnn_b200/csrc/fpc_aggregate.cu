// Aggregation half of the path: mask arg-max, 4-connected instance labelling, per-instance
// statistics, (instance,row) tables, masked sums and voting records.
//
// Reference behaviour being reproduced (paths relative to /root/reference/source_code/FastPoseCNN/):
//   lib/pose_regressor.py:449            cat_mask = argmax(log_softmax(mask logits))
//   lib/gpu_tensor_funcs.py:52-99        class_compress (select the predicted class's channels, normalise q / xy)
//   lib/aggregation_layer.py:43-59,160-183  connected components of cat_mask != 0, 4-connectivity, per image,
//                                        labels in raster order of each component's first pixel
//   lib/aggregation_layer.py:87-156      class id = min non-zero class in the component, masked means
//   lib/ransac_voting_gpu_layer/ransac_voting_gpu.py:532-550  per-instance pixel list in raster order
#include "fpc_internal.cuh"

#include <algorithm>

namespace fpc {

// =============================================================================================
// A. arg-max over the mask logits + run-aware label initialisation
// =============================================================================================
// One thread owns 4 consecutive pixels (one 16-byte load per class plane).  HBM-bound: 4*C bytes
// read, 5 bytes written per pixel.  label[p] is initialised to the first pixel of p's horizontal
// foreground run *inside the warp's 128-pixel span* (and inside its image row), -1 for background,
// so that only span-crossing and vertical adjacencies are left for the union-find merge.
__global__ void __launch_bounds__(256) k_argmax_init_v4(const float *__restrict__ mask, uint8_t *__restrict__ cls,
                                                        int *__restrict__ label, int C, int hw, int w, int P4) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int p = t * 4;
    int nib = 0, x0 = 0;
    int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (t < P4) {
        const int bi = p / hw;
        const int pix = p - bi * hw;
        x0 = pix % w;
        const float4 *src = reinterpret_cast<const float4 *>(mask + (size_t)bi * C * hw + pix);
        const int plane4 = hw >> 2;
        float4 best = __ldcs(src);
        for (int c = 1; c < C; ++c) {
            float4 v = __ldcs(src + (size_t)c * plane4);
            // strict '>' keeps the first maximum, like torch.argmax
            if (v.x > best.x) { best.x = v.x; a0 = c; }
            if (v.y > best.y) { best.y = v.y; a1 = c; }
            if (v.z > best.z) { best.z = v.z; a2 = c; }
            if (v.w > best.w) { best.w = v.w; a3 = c; }
        }
        nib = (a0 != 0) | ((a1 != 0) << 1) | ((a2 != 0) << 2) | ((a3 != 0) << 3);
    }
    // --- run start inside the warp span -----------------------------------------------------
    int prev_last = __shfl_up_sync(FULL, (nib >> 3) & 1, 1);
    if (lane == 0) prev_last = 0;
    const bool cont = (nib & 1) && prev_last && (x0 != 0);   // my first pixel continues the previous lane's run
    const unsigned transparent = __ballot_sync(FULL, (nib == 0xF) && cont);
    const unsigned below = ~transparent & ((1u << lane) - 1u);
    const int s = below ? (31 - __clz(below)) : 0;           // nearest lane below me whose pixels break/start the run
    const int nib_s = __shfl_sync(FULL, nib, s);
    const int tail_ones = __clz(~((unsigned)nib_s << 28));   // trailing foreground pixels of lane s (0..4)
    const int chain_start = p - 4 * (lane - s) + (4 - tail_ones);
    if (t < P4) {
        int lab[4];
        int cur = -1;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if ((nib >> j) & 1) {
                if (cur < 0) cur = (j == 0 && cont) ? chain_start : p + j;
                lab[j] = cur;
            } else {
                lab[j] = -1;
                cur = -1;
            }
        }
        // labels are only read where cls != 0: all-background threads (most of the image) skip the 16-byte store
        if (nib) *reinterpret_cast<int4 *>(label + p) = make_int4(lab[0], lab[1], lab[2], lab[3]);
        *reinterpret_cast<uchar4 *>(cls + p) = make_uchar4((unsigned char)a0, (unsigned char)a1, (unsigned char)a2, (unsigned char)a3);
    }
}

// Scalar variant for widths that are not a multiple of 4 (or misaligned inputs): span = 1 pixel.
__global__ void __launch_bounds__(256) k_argmax_init_v1(const float *__restrict__ mask, uint8_t *__restrict__ cls,
                                                        int *__restrict__ label, int C, int hw, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const int bi = p / hw, pix = p - bi * hw;
    const float *src = mask + (size_t)bi * C * hw + pix;
    float best = __ldcs(src);
    int arg = 0;
    for (int c = 1; c < C; ++c) {
        float v = __ldcs(src + (size_t)c * hw);
        if (v > best) { best = v; arg = c; }
    }
    cls[p] = (uint8_t)arg;
    label[p] = arg ? p : -1;
}

// Label initialisation from an already categorical mask (AggregationLayer drop-in: cat_mask int64).
__global__ void __launch_bounds__(256) k_init_from_catmask(const long long *__restrict__ cat, uint8_t *__restrict__ cls,
                                                           int *__restrict__ label, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const long long c = cat[p];
    cls[p] = (uint8_t)(c < 0 ? 0 : (c > 255 ? 255 : c));
    label[p] = c != 0 ? p : -1;
}

// =============================================================================================
// B/C. union-find merge and flatten (roots = smallest linear index of the component)
// =============================================================================================
__device__ __forceinline__ int uf_find(const int *L, int x) {
    while (true) {
        int p = L[x];
        if (p == x) return x;
        x = p;
    }
}
__device__ __forceinline__ void uf_unite(int *L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { int t = a; a = b; b = t; }      // link the larger root under the smaller one
        int old = atomicMin(&L[a], b);
        if (old == a) return;                         // a was still a root: done
        a = old;                                      // somebody re-parented a meanwhile: retry from there
    }
}

// `span`: width of the pixel spans inside which k_argmax_init_* already linked horizontal runs
// (128 for the v4 kernel, 1 otherwise).  Spans start at multiples of `span` in linear index.
__global__ void __launch_bounds__(256) k_ccl_merge(const uint8_t *__restrict__ cls, int *label, int w, int hw, int P,
                                                   int span) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    if (!cls[p]) return;
    const int pix = p % hw;
    const int y = pix / w, x = pix - y * w;
    const bool left = x > 0 && cls[p - 1];
    if (left && (p % span) == 0) uf_unite(label, p, p - 1);
    if (y > 0 && cls[p - w]) {
        // if left and up-left are both foreground, the left pixel already carries this adjacency
        if (!(left && cls[p - w - 1])) uf_unite(label, p, p - w);
    }
}

constexpr int TILE = 1024;  // pixels per block in the flatten / id-assignment kernels (256 threads x 4)

__global__ void __launch_bounds__(256) k_ccl_flatten(int *label, int *__restrict__ tile_roots, int P) {
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int n = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = blockIdx.x * TILE + j * 256 + threadIdx.x;
        if (p < P) {
            const int l = label[p];
            if (l >= 0) {
                const int r = uf_find(label, l);
                if (r != l) label[p] = r;
                n += (r == p);
            }
        }
    }
    n = __reduce_add_sync(FULL, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) tile_roots[blockIdx.x] = s_n;
}

// D. exclusive scan of the per-tile root counts (single block) -> N
__global__ void __launch_bounds__(1024) k_scan_tiles(int *tile_roots, int ntiles, int *counters, int max_instances) {
    const int total = block_exclusive_scan_inplace(tile_roots, ntiles);
    if (threadIdx.x == 0) {
        counters[FPC_CNT_INSTANCES] = total;
        counters[FPC_CNT_FLAGS] = total > max_instances ? FPC_FLAG_INSTANCES : 0;
        counters[FPC_CNT_TICKET] = 0;
    }
}

// E. instance id = rank of the root pixel in raster order over the whole batch volume
//    (== scipy.ndimage.label's numbering, aggregation_layer.py:178)
__global__ void __launch_bounds__(256) k_assign_ids(const int *__restrict__ label, const int *__restrict__ tile_base,
                                                    int *__restrict__ idmap, InstTables T, int P, int max_instances) {
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int base = tile_base[blockIdx.x];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int p = blockIdx.x * TILE + j * 256 + threadIdx.x;
        const bool root = (p < P) && (label[p] == p);
        const unsigned bal = __ballot_sync(FULL, root);
        if (lane == 0) s_w[wid] = __popc(bal);
        __syncthreads();
        int woff = 0, tot = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int c = s_w[k];
            woff += (k < wid) ? c : 0;
            tot += c;
        }
        if (root) {
            const int id = base + woff + __popc(bal & ((1u << lane) - 1u));
            idmap[p] = id;
            if (id < max_instances) {
                T.root[id] = p;
                T.count[id] = 0;
                T.ymin[id] = INT_MAX;
                T.ymax[id] = -1;
                T.xmin[id] = INT_MAX;
                T.xmax[id] = -1;
                T.mincls[id] = INT_MAX;
                T.tiny[id] = 0;
            }
        }
        base += tot;
        __syncthreads();
    }
}

// F. per-instance pixel count, bounding box and minimum class id; rewrites label[p] from
//    "root pixel index" to "instance id + 1" (0 = background), i.e. the scipy label volume.
__global__ void __launch_bounds__(256) k_instance_stats(int *label, const int *__restrict__ idmap,
                                                        const uint8_t *__restrict__ cls, InstTables T, int w, int hw, int P,
                                                        int max_instances) {
    const int lane = threadIdx.x & 31;
    const int gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int p = gw * 128 + it * 32 + lane;
        int id = -1, x = 0, y = 0, c = 0;
        if (p < P) {
            const int l = label[p];
            if (l >= 0) {
                id = idmap[l];
                const int pix = p % hw;
                y = pix / w;
                x = pix - y * w;
                c = cls[p];
            }
            label[p] = id + 1;
        }
        unsigned rem = __ballot_sync(FULL, id >= 0);
        while (rem) {
            const int cur = __shfl_sync(FULL, id, __ffs(rem) - 1);
            const bool mine = (id == cur);
            const unsigned gm = __ballot_sync(FULL, mine);
            rem &= ~gm;
            const int xmn = __reduce_min_sync(FULL, mine ? x : INT_MAX);
            const int xmx = __reduce_max_sync(FULL, mine ? x : -1);
            const int ymn = __reduce_min_sync(FULL, mine ? y : INT_MAX);
            const int ymx = __reduce_max_sync(FULL, mine ? y : -1);
            const int cmn = __reduce_min_sync(FULL, mine ? c : INT_MAX);
            if (lane == 0 && cur < max_instances) {
                atomicAdd(&T.count[cur], __popc(gm));
                atomicMin(&T.xmin[cur], xmn);
                atomicMax(&T.xmax[cur], xmx);
                atomicMin(&T.ymin[cur], ymn);
                atomicMax(&T.ymax[cur], ymx);
                atomicMin(&T.mincls[cur], cmn);
            }
        }
    }
}

// =============================================================================================
// Vectorised variants (4 consecutive pixels per thread; need w % 4 == 0 and 16-byte aligned buffers).
// They read the 1-byte class map first and leave at once when all 4 pixels are background, so the
// background (78 % of cfg2) costs 1 B/px per pass.  `label` is only defined on foreground pixels
// until k_instance_stats_v4 rewrites the whole volume.
// =============================================================================================
__global__ void __launch_bounds__(256) k_ccl_merge_v4(const uint8_t *__restrict__ cls, int *label, int w, int hw, int P4,
                                                      int span) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= P4) return;
    const int p = t * 4;
    const uchar4 c = *reinterpret_cast<const uchar4 *>(cls + p);
    if (!(c.x | c.y | c.z | c.w)) return;
    const int pix = p % hw;
    const int y = pix / w, x0 = pix - y * w;
    const bool fg[4] = {c.x != 0, c.y != 0, c.z != 0, c.w != 0};
    const bool left0 = x0 > 0 && cls[p - 1];
    bool up[4] = {false, false, false, false};
    bool upleft0 = false;
    if (y > 0) {
        const uchar4 u = *reinterpret_cast<const uchar4 *>(cls + p - w);
        up[0] = u.x != 0; up[1] = u.y != 0; up[2] = u.z != 0; up[3] = u.w != 0;
        upleft0 = x0 > 0 && cls[p - w - 1];
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!fg[j]) continue;
        const bool left = j ? fg[j - 1] : left0;
        // horizontal adjacency not yet linked by the label initialisation (span = 128: only at span starts;
        // span = 1: every pixel)
        if (left && ((p + j) % span) == 0) uf_unite(label, p + j, p + j - 1);
        if (!up[j]) continue;
        const bool upleft = j ? up[j - 1] : upleft0;
        if (!(left && upleft)) uf_unite(label, p + j, p + j - w);
    }
}

// tile = 1024 consecutive pixels = 256 threads x 4
__global__ void __launch_bounds__(256) k_ccl_flatten_v4(const uint8_t *__restrict__ cls, int *label,
                                                        int *__restrict__ tile_roots, int P4) {
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int t = blockIdx.x * 256 + threadIdx.x;
    int n = 0;
    if (t < P4) {
        const int p = t * 4;
        const uchar4 c = *reinterpret_cast<const uchar4 *>(cls + p);
        if (c.x | c.y | c.z | c.w) {
            const int4 l = *reinterpret_cast<const int4 *>(label + p);
            const int lv[4] = {l.x, l.y, l.z, l.w};
            const bool fg[4] = {c.x != 0, c.y != 0, c.z != 0, c.w != 0};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (!fg[j]) continue;
                const int r = uf_find(label, lv[j]);
                if (r != lv[j]) label[p + j] = r;
                n += (r == p + j);
            }
        }
    }
    n = __reduce_add_sync(FULL, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) tile_roots[blockIdx.x] = s_n;
}

__global__ void __launch_bounds__(256) k_assign_ids_v4(const uint8_t *__restrict__ cls, const int *__restrict__ label,
                                                       const int *__restrict__ tile_base, int *__restrict__ idmap,
                                                       InstTables T, int P4, int max_instances) {
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int t = blockIdx.x * 256 + threadIdx.x;
    const int p = t * 4;
    int rootbits = 0;
    if (t < P4) {
        const uchar4 c = *reinterpret_cast<const uchar4 *>(cls + p);
        if (c.x | c.y | c.z | c.w) {
            const int4 l = *reinterpret_cast<const int4 *>(label + p);
            rootbits = ((c.x && l.x == p) ? 1 : 0) | ((c.y && l.y == p + 1) ? 2 : 0) | ((c.z && l.z == p + 2) ? 4 : 0) |
                       ((c.w && l.w == p + 3) ? 8 : 0);
        }
    }
    const int mine = __popc(rootbits);
    const int inc = warp_incl_scan(mine, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (!rootbits) return;
    int woff = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) woff += (k < wid) ? s_w[k] : 0;
    int id = tile_base[blockIdx.x] + woff + inc - mine;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (!((rootbits >> j) & 1)) continue;
        idmap[p + j] = id;
        if (id < max_instances) {
            T.root[id] = p + j;
            T.count[id] = 0;
            T.ymin[id] = INT_MAX; T.ymax[id] = -1; T.xmin[id] = INT_MAX; T.xmax[id] = -1;
            T.mincls[id] = INT_MAX;
                T.tiny[id] = 0;
        }
        ++id;
    }
}

// Per-block shared-memory accumulators: a 1024-pixel tile touches one to three instances, so the
// count/bbox/min-class atomics of a whole tile collapse into six global atomics per (tile, instance)
// instead of six per (warp, instance) -- the global atomics all land on a handful of cache lines and
// serialise in L2 otherwise.
struct StatSlots {
    int id[8], cnt[8], xmn[8], xmx[8], ymn[8], ymx[8], cmn[8];
};
__device__ __forceinline__ void stat_update(StatSlots &S, InstTables &T, int cur, int cnt, int xmn, int xmx, int ymn,
                                            int ymx, int cmn, int max_instances) {
    if (cur >= max_instances) return;
#pragma unroll 1
    for (int s = 0; s < 8; ++s) {
        int o = S.id[s];
        if (o == -1) o = atomicCAS(&S.id[s], -1, cur);
        if (o == -1 || o == cur) {
            atomicAdd(&S.cnt[s], cnt);
            atomicMin(&S.xmn[s], xmn); atomicMax(&S.xmx[s], xmx);
            atomicMin(&S.ymn[s], ymn); atomicMax(&S.ymx[s], ymx);
            atomicMin(&S.cmn[s], cmn);
            return;
        }
    }
    atomicAdd(&T.count[cur], cnt);          // more than 8 instances in one tile: straight to global
    atomicMin(&T.xmin[cur], xmn); atomicMax(&T.xmax[cur], xmx);
    atomicMin(&T.ymin[cur], ymn); atomicMax(&T.ymax[cur], ymx);
    atomicMin(&T.mincls[cur], cmn);
}

__global__ void __launch_bounds__(256) k_instance_stats_v4(int *label, const int *__restrict__ idmap,
                                                           const uint8_t *__restrict__ cls, InstTables T, int w, int hw,
                                                           int P4, int max_instances) {
    __shared__ StatSlots S;
    const int lane = threadIdx.x & 31;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int p = t * 4;
    uchar4 c = make_uchar4(0, 0, 0, 0);
    if (t < P4) c = *reinterpret_cast<const uchar4 *>(cls + p);
    if (threadIdx.x < 8) {
        S.id[threadIdx.x] = -1; S.cnt[threadIdx.x] = 0;
        S.xmn[threadIdx.x] = INT_MAX; S.xmx[threadIdx.x] = -1; S.ymn[threadIdx.x] = INT_MAX; S.ymx[threadIdx.x] = -1;
        S.cmn[threadIdx.x] = INT_MAX;
    }
    // all-background tile (most of the image): write the zero labels and leave after this one barrier
    if (!__syncthreads_or(c.x | c.y | c.z | c.w)) {
        if (t < P4) *reinterpret_cast<int4 *>(label + p) = make_int4(0, 0, 0, 0);
        return;
    }
    int id[4] = {-1, -1, -1, -1};
    int cc[4] = {0, 0, 0, 0};
    int x0 = 0, y = 0;
    if (t < P4) {
        if (c.x | c.y | c.z | c.w) {
            const int4 l = *reinterpret_cast<const int4 *>(label + p);
            cc[0] = c.x; cc[1] = c.y; cc[2] = c.z; cc[3] = c.w;
            if (c.x) id[0] = idmap[l.x];
            if (c.y) id[1] = idmap[l.y];
            if (c.z) id[2] = idmap[l.z];
            if (c.w) id[3] = idmap[l.w];
            const int pix = p % hw;
            y = pix / w;
            x0 = pix - y * w;
        }
        *reinterpret_cast<int4 *>(label + p) = make_int4(id[0] + 1, id[1] + 1, id[2] + 1, id[3] + 1);
    }
    const bool any = (id[0] & id[1] & id[2] & id[3]) != -1;   // some pixel is foreground (ids are >= 0 or -1)
    if (__any_sync(FULL, any)) {
        // threads whose four pixels share one id: aggregate across the warp, one update per distinct id
        const bool uniform = id[0] >= 0 && id[0] == id[1] && id[1] == id[2] && id[2] == id[3];
        unsigned rem = __ballot_sync(FULL, uniform);
        while (rem) {
            const int cur = __shfl_sync(FULL, id[0], __ffs(rem) - 1);
            const bool mine = uniform && id[0] == cur;
            const unsigned gm = __ballot_sync(FULL, mine);
            rem &= ~gm;
            const int xmn = __reduce_min_sync(FULL, mine ? x0 : INT_MAX);
            const int xmx = __reduce_max_sync(FULL, mine ? x0 + 3 : -1);
            const int ymn = __reduce_min_sync(FULL, mine ? y : INT_MAX);
            const int ymx = __reduce_max_sync(FULL, mine ? y : -1);
            const int cmn = __reduce_min_sync(FULL, mine ? min(min(cc[0], cc[1]), min(cc[2], cc[3])) : INT_MAX);
            if (lane == 0) stat_update(S, T, cur, 4 * __popc(gm), xmn, xmx, ymn, ymx, cmn, max_instances);
        }
        // threads straddling an instance border: per pixel
        if (any && !uniform) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (id[j] >= 0) stat_update(S, T, id[j], 1, x0 + j, x0 + j, y, y, cc[j], max_instances);
        }
    }
    __syncthreads();
    if (threadIdx.x < 8 && S.id[threadIdx.x] >= 0) {
        const int s = threadIdx.x, cur = S.id[s];
        atomicAdd(&T.count[cur], S.cnt[s]);
        atomicMin(&T.xmin[cur], S.xmn[s]); atomicMax(&T.xmax[cur], S.xmx[s]);
        atomicMin(&T.ymin[cur], S.ymn[s]); atomicMax(&T.ymax[cur], S.ymx[s]);
        atomicMin(&T.mincls[cur], S.cmn[s]);
    }
}

// G. rows of every instance's bounding box -> rowoff (exclusive scan over instances, single block)
__global__ void __launch_bounds__(1024) k_scan_rows_per_instance(InstTables T, int *counters, int max_instances,
                                                                  long long max_rows) {
    const int N = min(counters[FPC_CNT_INSTANCES], max_instances);
    for (int i = threadIdx.x; i < N; i += blockDim.x) T.rowoff[i] = T.ymax[i] - T.ymin[i] + 1;
    __syncthreads();
    const int total = block_exclusive_scan_inplace(T.rowoff, N);
    if (threadIdx.x == 0) {
        T.rowoff[N] = total;
        counters[FPC_CNT_ROWS] = total;
        if ((long long)total > max_rows) atomicOr(&counters[FPC_CNT_FLAGS], FPC_FLAG_ROWS);
    }
}

__device__ __forceinline__ float select_uniform(const PathParams &pp, int p) {
    if (pp.select_u) return pp.select_u[p];
    return (float)(hash3(pp.seed, (uint32_t)p, 0x5e1ec7u, 0u) >> 8) * (1.0f / 16777216.0f);
}

// H. one block per instance: voting pixels of every row of its bounding box, their exclusive prefix inside
//    the instance (-> position of the row's first voting record), the row -> instance map, tn, and the
//    zeroing of the instance's vote counters.
//    ransac_voting_gpu.py:536-545: fewer than min_num pixels -> the instance does not vote;
//    more than max_num -> Bernoulli(max_num / count) sub-sampling.
constexpr int ROWS_PER_PASS = 1024;
constexpr int ROW_CONTIG = 1 << 30;   // the row's members are one contiguous run: no label test needed
constexpr int ROW_SUB = 1 << 29;      // instance larger than max_num: Bernoulli sub-sampling of the voters
constexpr int ROW_VOTES = 1 << 28;    // instance has at least min_num pixels
constexpr int ROW_LEN_MASK = (1 << 28) - 1;
__global__ void __launch_bounds__(1024) k_rows(const int *__restrict__ label, InstTables T, RowTables R,
                                               const int *__restrict__ counters, PathParams pp, int *__restrict__ votes) {
    __shared__ int s_cnt[ROWS_PER_PASS];
    __shared__ int s_w[32];
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    for (int i = blockIdx.x; i < N; i += gridDim.x) {
        const int r0 = T.rowoff[i], nrows = T.rowoff[i + 1] - r0;
        const int cnt = T.count[i], ymin = T.ymin[i], x0 = T.xmin[i], x1 = T.xmax[i];
        const int img = T.root[i] / pp.hw;
        const bool votes_at_all = cnt >= pp.min_num;
        const bool sub = cnt > pp.max_num;
        const float thr = (float)pp.max_num / (float)cnt;
        for (int k = tid; k < pp.hn; k += 1024) votes[(size_t)i * pp.hn + k] = 0;
        int carry = 0;
        for (int rb = 0; rb < nrows; rb += ROWS_PER_PASS) {
            const int nr = min(ROWS_PER_PASS, nrows - rb);
            for (int rr = wv; rr < nr; rr += 32) {
                // members of this row: count, extent [xs, xe], and how many of them vote
                int n = 0, members = 0, xs = INT_MAX, xe = -1;
                const int rowbase = img * pp.hw + (ymin + rb + rr) * pp.w;
                for (int xb = x0; xb <= x1; xb += 32) {
                    const int x = xb + lane;
                    bool m = (x <= x1) && (label[rowbase + x] == i + 1);
                    const unsigned mb = __ballot_sync(FULL, m);
                    if (mb) {
                        members += __popc(mb);
                        xs = min(xs, xb + __ffs(mb) - 1);
                        xe = max(xe, xb + 31 - __clz(mb));
                    }
                    if (m && sub) m = select_uniform(pp, rowbase + x) < thr;
                    if (votes_at_all) n += __popc(__ballot_sync(FULL, m));
                }
                if (lane == 0) {
                    s_cnt[rr] = n;
                    if (xe < xs) { xs = x0; xe = x0 - 1; }                 // empty row of a bounding box (dense problems)
                    const int len = xe - xs + 1;
                    const int flags = (members == len ? ROW_CONTIG : 0) | (sub ? ROW_SUB : 0) | (votes_at_all ? ROW_VOTES : 0);
                    R.desc[r0 + rb + rr] = make_int4(i, rowbase + xs, len | flags, 0);   // .w (prefix) is filled in below
                }
            }
            __syncthreads();
            // exclusive scan of s_cnt[0..nr): one row per thread
            const int v = tid < nr ? s_cnt[tid] : 0;
            const int inc = warp_incl_scan(v, lane);
            if (lane == 31) s_w[wv] = inc;
            __syncthreads();
            if (wv == 0) {
                const int w = s_w[lane];
                const int winc = warp_incl_scan(w, lane);
                s_w[lane] = winc - w;
                if (lane == 31) s_cnt[0] = winc;      // total of the pass (s_cnt is free again)
            }
            __syncthreads();
            if (tid < nr) R.desc[r0 + rb + tid].w = carry + s_w[wv] + inc - v;
            carry += s_cnt[0];
            __syncthreads();
        }
        if (tid == 0) T.tn[i] = carry;
    }
}

// I2. record offsets and vote work items per instance (single block)
__global__ void __launch_bounds__(1024) k_scan_records(InstTables T, int *counters, long long max_records, int chunk,
                                                       int nbatch) {
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int tn = T.tn[i];
        T.pxoff[i] = (tn + 15) & ~15;                        // ranges padded to whole 16-pixel voting rounds
        T.workoff[i] = ((tn + chunk - 1) / chunk) * nbatch;
    }
    __syncthreads();
    const int total = block_exclusive_scan_inplace(T.pxoff, N);
    __syncthreads();
    const int work = block_exclusive_scan_inplace(T.workoff, N);
    if (threadIdx.x == 0) {
        T.pxoff[N] = total;
        T.workoff[N] = work;
        counters[FPC_CNT_RECORDS] = total;
        counters[FPC_CNT_WORK] = work;
        if ((long long)total > max_records) atomicOr(&counters[FPC_CNT_FLAGS], FPC_FLAG_RECORDS);
    }
}

// J. one warp per (instance,row): masked sums of the predicted class's quaternion / scales / z
//    (aggregation_layer.py:125-149, on the class-compressed + per-pixel-normalised fields of
//    gpu_tensor_funcs.py:78-94) and the raster-ordered voting records (x, y, dir_x, dir_y)
//    (ransac_voting_gpu.py:547-550).  Head maps are read exactly once, foreground pixels only.
// MODE 0: raw head maps, class selected per pixel, q/xy normalised here (fused path)
// MODE 1: class-compressed CategoricalData [b,4|3|2,h,w] (AggregationLayer drop-in)
// MODE 2: voting records only, directions from a strided `vertex[N,h,w,vn,2]` view (ransac_voting_layer* drop-in)
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_gather(const int *__restrict__ label, const uint8_t *__restrict__ cls,
                                                InstTables T, RowTables R, const int *__restrict__ counters,
                                                PathParams pp, FieldSrc F, RecPlanes rec, bool want_rec) {
    if (counters[FPC_CNT_FLAGS]) return;
    const int rows = counters[FPC_CNT_ROWS];
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int K = pp.num_classes - 1;
    const size_t hw = (size_t)pp.hw;
    for (int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < rows; r += nwarps) {
        const int4 d = R.desc[r];                 // (instance, first member pixel, length | flags, record prefix)
        const int i = d.x, p0 = d.y, len = d.z & ROW_LEN_MASK;
        const bool contig = d.z & ROW_CONTIG, sub = d.z & ROW_SUB, votes = d.z & ROW_VOTES;
        const int img = p0 / pp.hw;
        const int pix0 = p0 - img * pp.hw;
        const int y = pix0 / pp.w, x0 = pix0 - y * pp.w;
        const float thr = sub ? (float)pp.max_num / (float)T.count[i] : 2.f;
        const int rec0 = (want_rec && votes) ? T.pxoff[i] + d.w : 0;
        int running = 0;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int kb = 0; kb < len; kb += 32) {
            const int kx = kb + lane;
            const int p = p0 + kx;
            const bool mem = (kx < len) && (contig || label[p] == i + 1);
            float vx = 0.f, vy = 0.f;
            if (mem) {
                const size_t pix = (size_t)(pix0 + kx);
                float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, zz = 0.f;
                if (MODE == 2) {
                    const float *v = F.xy + (long long)(img / F.div) * F.sN + (long long)y * F.sH + (long long)(x0 + kx) * F.sW;
                    vx = v[0];
                    vy = v[F.s2];
                } else if (MODE == 0) {
                    const size_t koff = (size_t)((int)cls[p] - 1) * hw;   // predicted class of THIS pixel (class_compress is per pixel)
                    const float *q = F.quaternion + (size_t)img * 4 * K * hw + 4 * koff + pix;
                    const float *s = F.scales + (size_t)img * 3 * K * hw + 3 * koff + pix;
                    const float *v = F.xy + (size_t)img * 2 * K * hw + 2 * koff + pix;
                    q0 = __ldcs(q); q1 = __ldcs(q + hw); q2 = __ldcs(q + 2 * hw); q3 = __ldcs(q + 3 * hw);
                    s0 = __ldcs(s); s1 = __ldcs(s + hw); s2 = __ldcs(s + 2 * hw);
                    zz = __ldcs(F.z + (size_t)img * K * hw + koff + pix);
                    vx = __ldcs(v); vy = __ldcs(v + hw);
                    // quaternion: only its masked mean is used (1e-4 budget) -> one reciprocal, four multiplies;
                    // direction: feeds the votes -> keep the reference's value / norm with IEEE sqrt and divide
                    const float qq = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
                    if (qq != 0.f) { const float rq = rsqrtf(qq); q0 *= rq; q1 *= rq; q2 *= rq; q3 *= rq; }
                    const float vn = __fsqrt_rn(vx * vx + vy * vy);
                    if (vn != 0.f) { vx = __fdiv_rn(vx, vn); vy = __fdiv_rn(vy, vn); }
                } else {
                    // already class-compressed CategoricalData (lib/type_hinting.py:12-17): [b,4|3|2,h,w], z [b,h,w]
                    const float *q = F.quaternion + ((size_t)img * 4) * hw + pix;
                    const float *s = F.scales + ((size_t)img * 3) * hw + pix;
                    const float *v = F.xy + ((size_t)img * 2) * hw + pix;
                    q0 = q[0]; q1 = q[hw]; q2 = q[2 * hw]; q3 = q[3 * hw];
                    s0 = s[0]; s1 = s[hw]; s2 = s[2 * hw];
                    zz = F.z[(size_t)img * hw + pix];
                    vx = v[0]; vy = v[hw];
                }
                acc[0] += q0; acc[1] += q1; acc[2] += q2; acc[3] += q3;
                acc[4] += s0; acc[5] += s1; acc[6] += s2; acc[7] += zz;
            }
            bool sel = mem && votes;
            if (sel && sub) sel = select_uniform(pp, p) < thr;
            const unsigned bal = __ballot_sync(FULL, sel);
            if (sel && want_rec) {
                // a direction the reference's |n| < 1e-6 guard would skip but that is not exactly zero: the fast
                // vote test cannot see that, so the whole instance is voted with the reference expression
                if ((vx != 0.f || vy != 0.f) && fmaf(vx, vx, vy * vy) < 1.1e-12f) T.tiny[i] = 1;
                const int idx = rec0 + running + __popc(bal & ((1u << lane) - 1u));
                rec.x[idx] = (float)(x0 + kx); rec.y[idx] = (float)y; rec.nx[idx] = vx; rec.ny[idx] = vy;
            }
            running += __popc(bal);
        }
        if (MODE != 2) {
            // warp sum of 8 values in 9 shuffles: halve the value set at every exchange (lane bit 4 keeps q or s/z, ...)
            float a4[4], a2[2], a1;
            {
                const bool up = lane & 16;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float send = up ? acc[k] : acc[k + 4];
                    const float keep = up ? acc[k + 4] : acc[k];
                    a4[k] = keep + __shfl_xor_sync(FULL, send, 16);
                }
            }
            {
                const bool up = lane & 8;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float send = up ? a4[k] : a4[k + 2];
                    const float keep = up ? a4[k + 2] : a4[k];
                    a2[k] = keep + __shfl_xor_sync(FULL, send, 8);
                }
            }
            {
                const bool up = lane & 4;
                const float send = up ? a2[0] : a2[1];
                const float keep = up ? a2[1] : a2[0];
                a1 = keep + __shfl_xor_sync(FULL, send, 4);
            }
            a1 += __shfl_xor_sync(FULL, a1, 2);
            a1 += __shfl_xor_sync(FULL, a1, 1);
            // lane L (L % 4 == 0) now holds value index ((L>>4)&1)*4 + ((L>>3)&1)*2 + ((L>>2)&1)
            if ((lane & 3) == 0) {
                const int k = ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1);
                R.sum[(size_t)r * 8 + k] = a1;
            }
        }
    }
}


// =============================================================================================
// Dense "problems" (drop-in voting entry points): one plane per problem
// =============================================================================================
// problem j votes with the pixels of plane src_plane[j] where  fmask != 0  (ransac_voting_layer_v3:
// mask[N,h,w], any float)  or  imask == match[j]  (ransac_voting_layer v1: class-id mask [b,h,w]).
// Writes the same label volume / tables the connected-component path produces (label = j+1 inside the
// problem's plane j), so every later kernel is shared.
__global__ void __launch_bounds__(256) k_dense_problems(const float *__restrict__ fmask, const int *__restrict__ imask,
                                                        int nplanes_per_src, int match_base, int *__restrict__ label,
                                                        InstTables T, int w, int hw, int nprob) {
    const int lane = threadIdx.x & 31;
    const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long warps_per_plane = (hw + 127) / 128;
    const int j = (int)(gw / warps_per_plane);
    if (j >= nprob) return;
    const int base = (int)(gw - (long long)j * warps_per_plane) * 128;
    int cnt = 0, xmn = INT_MAX, xmx = -1, ymn = INT_MAX, ymx = -1;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
        const int pix = base + it * 32 + lane;
        if (pix < hw) {
            bool m;
            if (fmask) {
                m = fmask[(size_t)j * hw + pix] != 0.f;
            } else {
                const int src = j / nplanes_per_src, k = j - src * nplanes_per_src;
                m = imask[(size_t)src * hw + pix] == match_base + k;
            }
            label[(size_t)j * hw + pix] = m ? j + 1 : 0;
            if (m) {
                const int y = pix / w, x = pix - y * w;
                ++cnt;
                xmn = min(xmn, x); xmx = max(xmx, x); ymn = min(ymn, y); ymx = max(ymx, y);
            }
        }
    }
    cnt = __reduce_add_sync(FULL, cnt);
    if (cnt) {
        xmn = __reduce_min_sync(FULL, xmn); xmx = __reduce_max_sync(FULL, xmx);
        ymn = __reduce_min_sync(FULL, ymn); ymx = __reduce_max_sync(FULL, ymx);
        if (lane == 0) {
            atomicAdd(&T.count[j], cnt);
            atomicMin(&T.xmin[j], xmn); atomicMax(&T.xmax[j], xmx);
            atomicMin(&T.ymin[j], ymn); atomicMax(&T.ymax[j], ymx);
        }
    }
}

__global__ void __launch_bounds__(256) k_dense_init(InstTables T, int *counters, int hw, int nprob) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) {
        counters[FPC_CNT_INSTANCES] = nprob;
        counters[FPC_CNT_FLAGS] = 0;
        counters[FPC_CNT_TICKET] = 0;
    }
    if (j >= nprob) return;
    T.root[j] = j * hw;   // plane index = root / hw
    T.count[j] = 0;
    T.xmin[j] = INT_MAX; T.xmax[j] = -1; T.ymin[j] = INT_MAX; T.ymax[j] = -1;
    T.mincls[j] = 0;
    T.tiny[j] = 0;
}

// empty problems get a one-row, zero-width box so that the row tables stay well formed
__global__ void __launch_bounds__(256) k_dense_fix_empty(InstTables T, int nprob) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nprob) return;
    if (T.count[j] == 0) { T.xmin[j] = 0; T.xmax[j] = -1; T.ymin[j] = 0; T.ymax[j] = 0; }
}

int launch_dense_problems(const Workspace &ws, const PathParams &pp, const float *fmask, const int *imask,
                          int nplanes_per_src, int match_base, int nprob, cudaStream_t st) {
    k_dense_init<<<ceil_div(std::max(nprob, 1), 256), 256, 0, st>>>(ws.T, ws.counters, pp.hw, nprob);
    FPC_LAUNCH_CHECK("k_dense_init");
    if (nprob > 0) {
        const long long warps = (long long)nprob * ((pp.hw + 127) / 128);
        k_dense_problems<<<(unsigned)ceil_div_ll(warps, 8), 256, 0, st>>>(fmask, imask, nplanes_per_src, match_base, ws.label,
                                                                          ws.T, pp.w, pp.hw, nprob);
        FPC_LAUNCH_CHECK("k_dense_problems");
        k_dense_fix_empty<<<ceil_div(nprob, 256), 256, 0, st>>>(ws.T, nprob);
        FPC_LAUNCH_CHECK("k_dense_fix_empty");
    }
    k_scan_rows_per_instance<<<1, 1024, 0, st>>>(ws.T, ws.counters, pp.max_instances, pp.max_rows);
    FPC_LAUNCH_CHECK("k_scan_rows_per_instance");
    return FPC_OK;
}

// Dense reference-layout outputs of AggregationLayer.forward (aggregation_layer.py:101-105,152-153):
// instance_masks [N,h,w] f32 0/1 and the masked direction field xy [N,2,h,w].
__global__ void __launch_bounds__(256) k_materialize(const int *__restrict__ label, const float *__restrict__ table,
                                                     const float *__restrict__ xy_cat, float *__restrict__ masks,
                                                     float *__restrict__ xy_mask, int hw, long long total) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / hw);
        const int pix = (int)(t - (long long)i * hw);
        const int img = __float_as_int(table[(size_t)i * FPC_POSE_ROW + FPC_ROW_SAMPLE]);
        const bool m = label[(size_t)img * hw + pix] == i + 1;
        if (masks) masks[t] = m ? 1.f : 0.f;
        if (xy_mask) {
            // the reference multiplies: mask * value (so a masked-out NaN stays NaN; we write 0 * v as well)
            const float vx = xy_cat[((size_t)img * 2) * hw + pix], vy = xy_cat[((size_t)img * 2 + 1) * hw + pix];
            const float mm = m ? 1.f : 0.f;
            xy_mask[((size_t)i * 2) * hw + pix] = mm * vx;
            xy_mask[((size_t)i * 2 + 1) * hw + pix] = mm * vy;
        }
    }
}

int launch_materialize(const int *label, const float *table, const float *xy_cat, float *masks, float *xy_mask, int n,
                       int hw, cudaStream_t st) {
    const long long total = (long long)n * hw;
    if (total == 0) return FPC_OK;
    const int grid = (int)std::min<long long>(ceil_div_ll(total, 256), (long long)sm_count() * 32);
    k_materialize<<<grid, 256, 0, st>>>(label, table, xy_cat, masks, xy_mask, hw, total);
    FPC_LAUNCH_CHECK("k_materialize");
    return FPC_OK;
}

// =============================================================================================
// host-side launch sequence of the aggregation half
// =============================================================================================
int launch_label_and_tables(const Workspace &ws, const PathParams &pp, const float *mask_logits,
                            const long long *cat_mask_i64, cudaStream_t st) {
    const int P = pp.P;
    const int ntiles = ceil_div(P, TILE);
    int span = 1;
    if (mask_logits) {
        const bool vec_ok = (pp.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(mask_logits) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(ws.label) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(ws.cls) & 3) == 0);
        if (vec_ok) {
            const int P4 = P / 4;
            k_argmax_init_v4<<<ceil_div(P4, 256), 256, 0, st>>>(mask_logits, ws.cls, ws.label, pp.num_classes, pp.hw,
                                                                pp.w, P4);
            span = 128;
        } else {
            k_argmax_init_v1<<<ceil_div(P, 256), 256, 0, st>>>(mask_logits, ws.cls, ws.label, pp.num_classes, pp.hw, P);
        }
        FPC_LAUNCH_CHECK("k_argmax_init");
    } else {
        k_init_from_catmask<<<ceil_div(P, 256), 256, 0, st>>>(cat_mask_i64, ws.cls, ws.label, P);
        FPC_LAUNCH_CHECK("k_init_from_catmask");
    }
    const bool v4 = (pp.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(ws.label) & 15) == 0) &&
                    ((reinterpret_cast<uintptr_t>(ws.cls) & 3) == 0);
    const int P4 = P / 4;
    if (v4) {
        k_ccl_merge_v4<<<ceil_div(P4, 256), 256, 0, st>>>(ws.cls, ws.label, pp.w, pp.hw, P4, span);
        FPC_LAUNCH_CHECK("k_ccl_merge");
        k_ccl_flatten_v4<<<ntiles, 256, 0, st>>>(ws.cls, ws.label, ws.tile_roots, P4);
        FPC_LAUNCH_CHECK("k_ccl_flatten");
    } else {
        k_ccl_merge<<<ceil_div(P, 256), 256, 0, st>>>(ws.cls, ws.label, pp.w, pp.hw, P, span);
        FPC_LAUNCH_CHECK("k_ccl_merge");
        k_ccl_flatten<<<ntiles, 256, 0, st>>>(ws.label, ws.tile_roots, P);
        FPC_LAUNCH_CHECK("k_ccl_flatten");
    }
    k_scan_tiles<<<1, 1024, 0, st>>>(ws.tile_roots, ntiles, ws.counters, pp.max_instances);
    FPC_LAUNCH_CHECK("k_scan_tiles");
    if (v4) {
        k_assign_ids_v4<<<ntiles, 256, 0, st>>>(ws.cls, ws.label, ws.tile_roots, ws.idmap, ws.T, P4, pp.max_instances);
        FPC_LAUNCH_CHECK("k_assign_ids");
        k_instance_stats_v4<<<ceil_div(P4, 256), 256, 0, st>>>(ws.label, ws.idmap, ws.cls, ws.T, pp.w, pp.hw, P4,
                                                               pp.max_instances);
        FPC_LAUNCH_CHECK("k_instance_stats");
    } else {
        k_assign_ids<<<ntiles, 256, 0, st>>>(ws.label, ws.tile_roots, ws.idmap, ws.T, P, pp.max_instances);
        FPC_LAUNCH_CHECK("k_assign_ids");
        k_instance_stats<<<ceil_div(P, 128 * 8), 256, 0, st>>>(ws.label, ws.idmap, ws.cls, ws.T, pp.w, pp.hw, P,
                                                               pp.max_instances);
        FPC_LAUNCH_CHECK("k_instance_stats");
    }
    k_scan_rows_per_instance<<<1, 1024, 0, st>>>(ws.T, ws.counters, pp.max_instances, pp.max_rows);
    FPC_LAUNCH_CHECK("k_scan_rows_per_instance");
    return FPC_OK;
}

int launch_rows_and_records(const Workspace &ws, const PathParams &pp, const FieldSrc &F, int gather_mode,
                            bool want_records, int vote_chunk, cudaStream_t st) {
    const int grid = sm_count() * 32;   // one warp per (instance,row) item, grid-stride: plenty of loads in flight
    k_rows<<<sm_count() * 2, 1024, 0, st>>>(ws.label, ws.T, ws.R, ws.counters, pp, ws.votes);
    FPC_LAUNCH_CHECK("k_rows");
    k_scan_records<<<1, 1024, 0, st>>>(ws.T, ws.counters, pp.max_records, vote_chunk, vote_batches(pp.hn));
    FPC_LAUNCH_CHECK("k_scan_records");
    if (gather_mode == 0)
        k_gather<0><<<grid, 256, 0, st>>>(ws.label, ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    else if (gather_mode == 1)
        k_gather<1><<<grid, 256, 0, st>>>(ws.label, ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    else
        k_gather<2><<<grid, 256, 0, st>>>(ws.label, ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    FPC_LAUNCH_CHECK("k_gather");
    return FPC_OK;
}

}  // namespace fpc
