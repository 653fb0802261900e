// Instance labelling on RUNS (maximal horizontal foreground segments) instead of pixels.
//
// Only two kernels touch every pixel: the arg-max over the mask logits (HBM-bound, 28 B/px) and the run
// emission (reads the 1-byte class map).  Connected components, instance numbering, per-instance statistics
// and the per-instance run tables all work on the run list (about one run per instance per image row:
// ~41k runs for 9.8M pixels at cfg2).
//
// Reference behaviour being reproduced (paths relative to /root/reference/source_code/FastPoseCNN/):
//   lib/pose_regressor.py:449               cat_mask = argmax(log_softmax(mask logits))
//   lib/aggregation_layer.py:43-59,160-183  connected components of cat_mask != 0, 4-connectivity, per image,
//                                           labels in raster order of each component's first pixel
//   lib/aggregation_layer.py:87-118         instance count, class id = min non-zero class in the component
#include "fpc_internal.cuh"

#include <algorithm>
#include <cstdlib>

namespace fpc {

constexpr int TILE = 1024;  // pixels (or runs) per block in the counting / ranking kernels

// =============================================================================================
// 1. class map + number of run starts per 1024-pixel tile
// =============================================================================================
// A run is a maximal horizontal foreground segment inside one image row and inside one `span` of the linear
// pixel index (spans start at multiples of `span`; the arg-max kernels cut runs at span borders so that no thread
// needs its left neighbour's class from another warp; cut pieces are re-joined by the run merge).
__device__ __forceinline__ void tile_count(int n, int *tile_runs) {
    __shared__ int s_n;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    n = __reduce_add_sync(FULL, n);
    if ((threadIdx.x & 31) == 0 && n) atomicAdd(&s_n, n);
    __syncthreads();
    if (threadIdx.x == 0) tile_runs[blockIdx.x] = s_n;
}

// One thread owns 4 consecutive pixels (one 16-byte streaming load per class plane).  HBM-bound: 4*C bytes
// read, 1 byte written per pixel.  span = 128 pixels (one warp).
__global__ void __launch_bounds__(256) k_argmax_runs_v4(const float *__restrict__ mask, uint8_t *__restrict__ cls,
                                                        int *__restrict__ tile_runs, int C, int hw, int w, int P4) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int p = t * 4;
    int nib = 0, x0 = 0;
    if (t < P4) {
        const int bi = p / hw;
        const int pix = p - bi * hw;
        x0 = pix % w;
        const float4 *src = reinterpret_cast<const float4 *>(mask + (size_t)bi * C * hw + pix);
        const int plane4 = hw >> 2;
        float4 best = __ldcs(src);
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        for (int c = 1; c < C; ++c) {
            const float4 v = __ldcs(src + (size_t)c * plane4);
            // strict '>' keeps the first maximum, like torch.argmax
            if (v.x > best.x) { best.x = v.x; a0 = c; }
            if (v.y > best.y) { best.y = v.y; a1 = c; }
            if (v.z > best.z) { best.z = v.z; a2 = c; }
            if (v.w > best.w) { best.w = v.w; a3 = c; }
        }
        nib = (a0 != 0) | ((a1 != 0) << 1) | ((a2 != 0) << 2) | ((a3 != 0) << 3);
        *reinterpret_cast<uchar4 *>(cls + p) = make_uchar4((unsigned char)a0, (unsigned char)a1, (unsigned char)a2, (unsigned char)a3);
    }
    int prev_last = __shfl_up_sync(FULL, (nib >> 3) & 1, 1);
    if (lane == 0 || x0 == 0) prev_last = 0;                 // span start or row start: a new run begins
    const int starts = nib & ~((nib << 1) | prev_last);      // foreground whose left neighbour is not
    tile_count(__popc(starts & 0xF), tile_runs);
}


// Persistent, software-pipelined variant of k_argmax_runs_v4 (P % 1024 == 0): a block walks over tiles and issues the C
// 16-byte loads of its NEXT tile before it picks the arg-max of the current one, so every thread keeps 2 C loads in flight
// and the bandwidth no longer depends on how many short-lived blocks are resident (plain kernel: 56 us at 8 blocks/SM, 88 us
// at 3, 118 us at 2).  Two blocks per SM reach the HBM rate and fit next to the vote kernel of another batch
// (tools/microbench_overlap.cu: an issue-bound kernel and a streaming kernel with deep loads in flight slow each other by
// 5-15 % only).  A copy-engine fed variant (cp.async.bulk of whole tiles into a 3-stage shared-memory ring, one block per
// SM) was measured too: 84 us, the 4 KB bulk copies of one block do not keep enough bytes in flight.  Same outputs as
// k_argmax_runs_v4.
template <int CT>
__global__ void __launch_bounds__(256) k_argmax_runs_p4(const float *__restrict__ mask, uint8_t *__restrict__ cls,
                                                        int *__restrict__ tile_runs, int hw, int w, int ntiles) {
    __shared__ int s_n[2];
    const int tid = threadIdx.x, lane = tid & 31;
    const int plane4 = hw >> 2;
    // (image, pixel in image, column) of this thread's first pixel, advanced incrementally from tile to tile: no division in
    // the loop (a 64-bit p / hw per tile used to be a third of this kernel's instructions)
    const int step = gridDim.x * TILE;
    const int step_img = step / hw, step_pix = step - step_img * hw, step_x = step % w;
    int tile = blockIdx.x;
    int p = tile * TILE + tid * 4;
    int bi = p / hw, pix = p - bi * hw, x0 = p % w;                 // hw is a multiple of w: (p % hw) % w == p % w
    // the NEXT tile's coordinates
    int bi_n = bi, pix_n = pix;
    auto src_of = [&](int img, int px) { return reinterpret_cast<const float4 *>(mask + (size_t)img * CT * hw + px); };
    float4 nxt[CT];
    if (tile < ntiles) {
        const float4 *src = src_of(bi, pix);
#pragma unroll
        for (int c = 0; c < CT; ++c) nxt[c] = __ldcs(src + (size_t)c * plane4);
    }
    for (int it = 0; tile < ntiles; ++it) {
        float4 cur[CT];
#pragma unroll
        for (int c = 0; c < CT; ++c) cur[c] = nxt[c];
        const int tn = tile + gridDim.x;
        bi_n = bi + step_img; pix_n = pix + step_pix;
        if (pix_n >= hw) { pix_n -= hw; ++bi_n; }
        if (tn < ntiles) {
            const float4 *src = src_of(bi_n, pix_n);
#pragma unroll
            for (int c = 0; c < CT; ++c) nxt[c] = __ldcs(src + (size_t)c * plane4);
        }
        float4 best = cur[0];
        int a0 = 0, a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
        for (int c = 1; c < CT; ++c) {
            // strict '>' keeps the first maximum, like torch.argmax
            if (cur[c].x > best.x) { best.x = cur[c].x; a0 = c; }
            if (cur[c].y > best.y) { best.y = cur[c].y; a1 = c; }
            if (cur[c].z > best.z) { best.z = cur[c].z; a2 = c; }
            if (cur[c].w > best.w) { best.w = cur[c].w; a3 = c; }
        }
        const int nib = (a0 != 0) | ((a1 != 0) << 1) | ((a2 != 0) << 2) | ((a3 != 0) << 3);
        *reinterpret_cast<uchar4 *>(cls + p) = make_uchar4((unsigned char)a0, (unsigned char)a1, (unsigned char)a2, (unsigned char)a3);
        int prev_last = __shfl_up_sync(FULL, (nib >> 3) & 1, 1);
        if (lane == 0 || x0 == 0) prev_last = 0;                 // span start or row start: a new run begins
        const int starts = nib & ~((nib << 1) | prev_last);      // foreground whose left neighbour is not
        // run starts of the tile: two alternating counters, one barrier per tile
        int *cnt = &s_n[it & 1];
        if (tid == 0) s_n[(it + 1) & 1] = 0;                     // next tile's counter (its last reader passed a barrier ago)
        const int n = __reduce_add_sync(FULL, __popc(starts & 0xF));
        if (it == 0) {                                           // first tile: the counter has not been cleared yet
            if (tid == 0) *cnt = 0;
            __syncthreads();
        }
        if (lane == 0 && n) atomicAdd(cnt, n);
        __syncthreads();
        if (tid == 0) tile_runs[tile] = *cnt;
        tile = tn;
        p += step; bi = bi_n; pix = pix_n;
        x0 += step_x;
        if (x0 >= w) x0 -= w;
    }
}

// Scalar variant (any width / alignment): one pixel per thread, span = 32 pixels, tile = 1024-thread block.
__global__ void __launch_bounds__(1024) k_argmax_runs_v1(const float *__restrict__ mask, uint8_t *__restrict__ cls,
                                                         int *__restrict__ tile_runs, int C, int hw, int w, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int fg = 0, x = 0;
    if (p < P) {
        const int bi = p / hw, pix = p - bi * hw;
        x = pix % w;
        const float *src = mask + (size_t)bi * C * hw + pix;
        float best = __ldcs(src);
        int arg = 0;
        for (int c = 1; c < C; ++c) {
            const float v = __ldcs(src + (size_t)c * hw);
            if (v > best) { best = v; arg = c; }
        }
        cls[p] = (uint8_t)arg;
        fg = arg != 0;
    }
    int prev = __shfl_up_sync(FULL, fg, 1);
    if (lane == 0 || x == 0) prev = 0;
    tile_count(fg && !prev, tile_runs);
}

// Head-epilogue fusion: the same class map and run-start counts, but the C mask logits of
// a pixel are the x S bilinear up-sampling (smp SegmentationHead, lib/pose_regressor.py:633-639) of the LOW-RESOLUTION
// head output, evaluated here -- the [b,C,h,w] logits never exist.  One thread = 4 consecutive output pixels; for
// S >= 3 they touch at most 3 low-res columns, so a class plane costs 6 cached loads per thread (1/16 of the bytes
// of the full-resolution kernel at S = 4) and 24 FP32 operations.
//
// CT > 0 (class count known at compile time): the 6 x CT taps are loaded up front (all in flight together), and a thread
// whose 6 taps are all background-dominant (logit 0 >= every other logit) skips the interpolation: bilinear weights are
// non-negative and every rounding step is monotone, so background then also wins -- or ties, and arg-max keeps the
// first maximum -- at all 4 output pixels.  Most of an image is such background.
template <int CT>
__global__ void __launch_bounds__(256) k_argmax_runs_up4(const float *__restrict__ mask_lr, uint8_t *__restrict__ cls,
                                                         int *__restrict__ tile_runs, int C, int hw, int w, int P4, UpParams up) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int p = t * 4;
    int nib = 0, x0 = 0;
    if (t < P4) {
        const int bi = p / hw;
        const int pix = p - bi * hw;
        const int y = pix / w;
        x0 = pix - y * w;
        const LerpCoord Y = lerp_coord(y, up.sy, up.hl);
        const LerpCoord X0 = lerp_coord(x0, up.sx, up.wl);
        const int cA = X0.i0, cB = min(cA + 1, up.wl - 1), cC = min(cA + 2, up.wl - 1);
        const size_t lhw = (size_t)up.hl * up.wl;
        const float *r0 = mask_lr + (size_t)bi * C * lhw + (size_t)Y.i0 * up.wl;
        const float *r1 = mask_lr + (size_t)bi * C * lhw + (size_t)Y.i1 * up.wl;
        int arg[4] = {0, 0, 0, 0};
        if (CT > 0) {
            float v[CT > 0 ? CT : 1][6];
#pragma unroll
            for (int c = 0; c < CT; ++c) {
                v[c][0] = __ldg(r0 + c * lhw + cA); v[c][1] = __ldg(r0 + c * lhw + cB); v[c][2] = __ldg(r0 + c * lhw + cC);
                v[c][3] = __ldg(r1 + c * lhw + cA); v[c][4] = __ldg(r1 + c * lhw + cB); v[c][5] = __ldg(r1 + c * lhw + cC);
            }
            bool bg = true;
#pragma unroll
            for (int k = 0; k < 6; ++k) {
                float m = v[CT > 1 ? 1 : 0][k];
#pragma unroll
                for (int c = 2; c < CT; ++c) m = fmaxf(m, v[c][k]);
                bg = bg && (v[0][k] >= m);                                 // false for NaN: falls through to the full evaluation
            }
            if (!bg) {
                LerpCoord X[4];
                X[0] = X0;
#pragma unroll
                for (int j = 1; j < 4; ++j) X[j] = lerp_coord(x0 + j, up.sx, up.wl);
                float best[4];
#pragma unroll
                for (int c = 0; c < CT; ++c) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool sh = X[j].i0 != cA;                     // pixel j starts one low-res column further right
                        const float val = bilerp(sh ? v[c][1] : v[c][0], sh ? v[c][2] : v[c][1], sh ? v[c][4] : v[c][3],
                                                 sh ? v[c][5] : v[c][4], X[j].w0, X[j].w1, Y.w0, Y.w1);
                        if (c == 0) best[j] = val;
                        else if (val > best[j]) { best[j] = val; arg[j] = c; }   // strict '>' keeps the first maximum, like torch.argmax
                    }
                }
            }
        } else {
            LerpCoord X[4];
            bool sh[4];
            X[0] = X0;
#pragma unroll
            for (int j = 1; j < 4; ++j) X[j] = lerp_coord(x0 + j, up.sx, up.wl);
#pragma unroll
            for (int j = 0; j < 4; ++j) sh[j] = X[j].i0 != cA;
            float best[4];
            for (int c = 0; c < C; ++c) {
                const float a0 = __ldg(r0 + cA), b0 = __ldg(r0 + cB), c0 = __ldg(r0 + cC);
                const float a1 = __ldg(r1 + cA), b1 = __ldg(r1 + cB), c1 = __ldg(r1 + cC);
                r0 += lhw;
                r1 += lhw;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float val = bilerp(sh[j] ? b0 : a0, sh[j] ? c0 : b0, sh[j] ? b1 : a1, sh[j] ? c1 : b1, X[j].w0, X[j].w1, Y.w0, Y.w1);
                    if (c == 0) best[j] = val;
                    else if (val > best[j]) { best[j] = val; arg[j] = c; }
                }
            }
        }
        nib = (arg[0] != 0) | ((arg[1] != 0) << 1) | ((arg[2] != 0) << 2) | ((arg[3] != 0) << 3);
        *reinterpret_cast<uchar4 *>(cls + p) = make_uchar4((unsigned char)arg[0], (unsigned char)arg[1], (unsigned char)arg[2], (unsigned char)arg[3]);
    }
    int prev_last = __shfl_up_sync(FULL, (nib >> 3) & 1, 1);
    if (lane == 0 || x0 == 0) prev_last = 0;
    const int starts = nib & ~((nib << 1) | prev_last);
    tile_count(__popc(starts & 0xF), tile_runs);
}

// Scalar variant of the above (any width, any S >= 2): one pixel per thread, 4 taps per class plane.
__global__ void __launch_bounds__(1024) k_argmax_runs_up1(const float *__restrict__ mask_lr, uint8_t *__restrict__ cls,
                                                          int *__restrict__ tile_runs, int C, int hw, int w, int P, UpParams up) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    int fg = 0, x = 0;
    if (p < P) {
        const int bi = p / hw, pix = p - bi * hw;
        const int y = pix / w;
        x = pix - y * w;
        const LerpCoord Y = lerp_coord(y, up.sy, up.hl), X = lerp_coord(x, up.sx, up.wl);
        const size_t lhw = (size_t)up.hl * up.wl;
        const float *r0 = mask_lr + (size_t)bi * C * lhw + (size_t)Y.i0 * up.wl;
        const float *r1 = mask_lr + (size_t)bi * C * lhw + (size_t)Y.i1 * up.wl;
        float best = 0.f;
        int arg = 0;
        for (int c = 0; c < C; ++c) {
            const float v = bilerp(__ldg(r0 + X.i0), __ldg(r0 + X.i1), __ldg(r1 + X.i0), __ldg(r1 + X.i1), X.w0, X.w1, Y.w0, Y.w1);
            r0 += lhw;
            r1 += lhw;
            if (c == 0) best = v;
            else if (v > best) { best = v; arg = c; }
        }
        cls[p] = (uint8_t)arg;
        fg = arg != 0;
    }
    int prev = __shfl_up_sync(FULL, fg, 1);
    if (lane == 0 || x == 0) prev = 0;
    tile_count(fg && !prev, tile_runs);
}

// Class map from an already categorical mask (AggregationLayer drop-in: cat_mask int64) or from dense problem
// planes (voting drop-ins: problem j owns plane j; member = fmask != 0, or imask[j / per_src] == match_base + j % per_src).
// Neighbours can be read directly, so runs are not cut: span = "infinite".
__global__ void __launch_bounds__(1024) k_cls_runs(const long long *__restrict__ cat, const float *__restrict__ fmask,
                                                   const int *__restrict__ imask, int per_src, int match_base,
                                                   uint8_t *__restrict__ cls, int *__restrict__ tile_runs, int hw, int w, int P) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    auto member = [&](int q) -> int {
        if (cat) { const long long c = cat[q]; return (int)(c < 0 ? 0 : (c > 255 ? 255 : c)); }
        const int j = q / hw;
        if (fmask) return fmask[q] != 0.f ? 1 : 0;
        const int src = j / per_src, k = j - src * per_src;
        return imask[(size_t)src * hw + (q - j * hw)] == match_base + k ? 1 : 0;
    };
    int start = 0;
    if (p < P) {
        const int c = member(p);
        cls[p] = (uint8_t)c;
        const int x = (p % hw) % w;
        start = c != 0 && (x == 0 || member(p - 1) == 0);
    }
    tile_count(start, tile_runs);
}

// 2/6. exclusive scan of per-tile counts (single block); total -> counters[which] (+ capacity flag)
__global__ void __launch_bounds__(1024) k_scan_tiles(int *tile_counts, int ntiles, int *counters, int which, long long cap,
                                                     int flag, int reset_flags) {
    const int total = block_exclusive_scan_inplace(tile_counts, ntiles);
    if (threadIdx.x == 0) {
        tile_counts[ntiles] = total;
        counters[which] = total;
        if (reset_flags) {
            counters[FPC_CNT_FLAGS] = 0;
            counters[FPC_CNT_TICKET] = 0;
        }
        if ((long long)total > cap) atomicOr(&counters[FPC_CNT_FLAGS], flag);
    }
}

// =============================================================================================
// 3. run emission: start / end of every run in raster order, run id per foreground pixel, first run of every row
// =============================================================================================
__global__ void __launch_bounds__(256) k_emit_runs(const uint8_t *__restrict__ cls, const int *__restrict__ tile_base,
                                                   RunTables RT, int w, int P, int span, long long cap) {
    __shared__ int s_w[8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = (blockIdx.x * 256 + threadIdx.x) * 4;
    const int smask = span - 1;                     // span is a power of two
    int fgb = 0, stb = 0, enb = 0, x0 = 0;
    if (p < P) {
        x0 = p - (p / w) * w;                       // image planes are whole rows: p % w is the column
        int c[6] = {0, 0, 0, 0, 0, 0};              // classes of pixels p-1 .. p+4
        if (p + 4 <= P) {
            const uchar4 v = *reinterpret_cast<const uchar4 *>(cls + p);
            c[1] = v.x; c[2] = v.y; c[3] = v.z; c[4] = v.w;
        } else {
            for (int j = 0; j < 4; ++j) c[j + 1] = (p + j < P) ? cls[p + j] : 0;
        }
        if (c[1] | c[2] | c[3] | c[4]) {            // neighbours only matter next to foreground
            c[0] = p > 0 ? cls[p - 1] : 0;
            c[5] = p + 4 < P ? cls[p + 4] : 0;
            int x = x0;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c[j + 1]) {
                    fgb |= 1 << j;
                    if (x == 0 || ((p + j) & smask) == 0 || !c[j]) stb |= 1 << j;
                    if (x == w - 1 || ((p + j + 1) & smask) == 0 || !c[j + 2]) enb |= 1 << j;
                }
                if (++x == w) x = 0;
            }
        }
    }
    const bool row_start = (x0 == 0) || (x0 + 3 >= w);
    // all-background tile (most of the image): no run starts or ends here; only the row index needs this tile's base
    if (!__syncthreads_or(fgb)) {
        if (p < P) {
            const int base = tile_base[blockIdx.x];
            if (row_start) {
                int x = x0;
                for (int j = 0; j < 4; ++j) {
                    if (p + j < P && x == 0) RT.rowrun[(p + j) / w] = base;
                    if (++x == w) x = 0;
                }
            }
            if (p + 4 >= P) RT.rowrun[P / w] = tile_base[gridDim.x];
        }
        return;
    }
    const int mine = __popc(stb);
    const int inc = warp_incl_scan(mine, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (p >= P) return;
    // only threads that start/end a run or own the first pixel of an image row have anything to write
    if (!(stb | enb) && !row_start && p + 4 < P) return;
    int before = tile_base[blockIdx.x] + inc - mine;           // run starts before my first pixel
#pragma unroll
    for (int k = 0; k < 8; ++k) before += (k < wid) ? s_w[k] : 0;
    int x = x0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        if (p + j < P) {
            if (x == 0) RT.rowrun[(p + j) / w] = before;       // first run at or after the start of this image row
            if ((stb >> j) & 1) {
                if (before < cap) { RT.start[before] = p + j; RT.parent[before] = before; }
                ++before;
            }
            if (((enb >> j) & 1) && before - 1 < cap) RT.end[before - 1] = p + j;
        }
        if (++x == w) x = 0;
    }
    if (p + 4 >= P) RT.rowrun[P / w] = tile_base[gridDim.x];   // sentinel: total number of runs
}

// The same for w % 16 == 0 (and spans that are multiples of 16): one thread owns 16 pixels = one 16-byte load; its
// foreground / run-start / run-end masks come from byte compares and shifts instead of a per-pixel loop, and only set bits
// are visited when writing.  64 threads per 1024-pixel tile; ~13x fewer instructions per pixel than the 4-pixel kernel.
__device__ __forceinline__ unsigned nonzero_bytes4(unsigned q) {
    // 0xFF per non-zero byte -> one bit per byte (bits land on 24..27: no carries between the partial products)
    return ((__vcmpne4(q, 0u) & 0x01010101u) * 0x01020408u) >> 24;
}
__global__ void __launch_bounds__(64) k_emit_runs16(const uint8_t *__restrict__ cls, const int *__restrict__ tile_base, RunTables RT,
                                                    int w, int P, int span, long long cap) {
    __shared__ int s_w[2];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int p = (blockIdx.x * 64 + threadIdx.x) * 16;          // P is a multiple of 16
    const int smask = span - 1;
    unsigned fg = 0, st = 0, en = 0;
    if (p < P) {
        const uint4 v = *reinterpret_cast<const uint4 *>(cls + p);
        fg = nonzero_bytes4(v.x) | (nonzero_bytes4(v.y) << 4) | (nonzero_bytes4(v.z) << 8) | (nonzero_bytes4(v.w) << 12);
    }
    const int x0 = p < P ? p - (p / w) * w : 1;                  // the 16 pixels lie in one image row (w % 16 == 0)
    if (fg) {
        const bool left_cut = x0 == 0 || (p & smask) == 0;                      // row start or span start: a new run begins
        const bool right_cut = x0 + 16 == w || ((p + 16) & smask) == 0;
        const unsigned prev = ((fg & 1u) && !left_cut && cls[p - 1]) ? 1u : 0u;
        const unsigned next = ((fg & 0x8000u) && !right_cut && cls[p + 16]) ? 0x8000u : 0u;
        st = fg & ~((fg << 1) | prev) & 0xFFFFu;
        en = fg & ~((fg >> 1) | next);
    }
    if (!__syncthreads_or(fg)) {                                 // all-background tile: only the row index needs the base
        if (p < P) {
            if (x0 == 0) RT.rowrun[p / w] = tile_base[blockIdx.x];
            if (p + 16 >= P) RT.rowrun[P / w] = tile_base[gridDim.x];
        }
        return;
    }
    const int mine = __popc(st);
    const int inc = warp_incl_scan(mine, lane);
    if (lane == 31) s_w[wid] = inc;
    __syncthreads();
    if (p >= P) return;
    int before = tile_base[blockIdx.x] + inc - mine + (wid ? s_w[0] : 0);     // run starts before my first pixel
    if (x0 == 0) RT.rowrun[p / w] = before;                      // first run at or after the start of this image row
    unsigned both = st | en;
    while (both) {
        const int j = __ffs(both) - 1;
        both &= both - 1;
        if ((st >> j) & 1u) {
            if (before < cap) { RT.start[before] = p + j; RT.parent[before] = before; }
            ++before;
        }
        if (((en >> j) & 1u) && before - 1 < cap) RT.end[before - 1] = p + j;
    }
    if (p + 16 >= P) RT.rowrun[P / w] = tile_base[gridDim.x];    // sentinel: total number of runs
}

// =============================================================================================
// 4. union-find over runs (roots = smallest run id = the run holding the component's first pixel)
// =============================================================================================
// find with path halving: every visited node is re-pointed to its grandparent (always an ancestor with a smaller
// index, so concurrent finds / atomicMin links stay consistent and pointers only ever decrease)
__device__ __forceinline__ int uf_find(int *L, int x) {
    while (true) {
        const int p = L[x];
        if (p == x) return x;
        const int gp = L[p];
        if (gp != p) L[x] = gp;
        x = gp;
    }
}
// read-only find for the flatten pass: there every run must end up pointing at its ROOT, so no other thread may
// re-point it to a mere ancestor while it is being flattened
__device__ __forceinline__ int uf_find_ro(const int *L, int x) {
    while (true) {
        const int p = L[x];
        if (p == x) return x;
        x = p;
    }
}
__device__ __forceinline__ void uf_unite(int *L, int a, int b) {
    while (true) {
        a = uf_find(L, a);
        b = uf_find(L, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }   // link the larger root under the smaller one
        const int old = atomicMin(&L[a], b);
        if (old == a) return;                            // a was still a root: done
        a = old;                                         // somebody re-parented a meanwhile: retry from there
    }
}

// one thread per run: joins it with the runs of the row above whose column intervals overlap its own (found by
// binary search in that row's slice of the run list), and with its left neighbour when the run was cut at a span border
__global__ void __launch_bounds__(256) k_run_merge(RunTables RT, const int *__restrict__ counters, int w, int h, long long cap) {
    if (counters[FPC_CNT_FLAGS]) return;     // more runs than the tables hold: reported through the flags
    const int M = (int)min((long long)counters[FPC_CNT_ROWS], cap);
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const int s = RT.start[m], e = RT.end[m];
        const int row = s / w;                     // global row index = image * h + y
        const int x0 = s - row * w, x1 = e - row * w;
        if (m > 0 && x0 > 0 && RT.end[m - 1] == s - 1) uf_unite(RT.parent, m, m - 1);
        if (row % h == 0) continue;
        const int ub = (row - 1) * w;              // linear index of the first pixel of the row above
        int lo = RT.rowrun[row - 1], hi = RT.rowrun[row];
        const int m1 = hi;
        while (lo < hi) {                          // first run of the row above that ends at or after x0
            const int mid = (lo + hi) >> 1;
            if (RT.end[mid] - ub < x0) lo = mid + 1; else hi = mid;
        }
        for (int u = lo; u < m1 && RT.start[u] - ub <= x1; ++u) uf_unite(RT.parent, m, u);
    }
}

// 5. path compression + number of roots per 1024-run tile
__global__ void __launch_bounds__(1024) k_run_flatten(RunTables RT, const int *__restrict__ counters, int *__restrict__ tile_roots,
                                                      long long cap) {
    const int M = (counters[FPC_CNT_FLAGS] & FPC_FLAG_ROWS) ? 0 : (int)min((long long)counters[FPC_CNT_ROWS], cap);
    const int m = blockIdx.x * TILE + threadIdx.x;
    int root = 0;
    if (m < M) {
        const int r = uf_find_ro(RT.parent, m);
        RT.parent[m] = r;
        root = (r == m);
    }
    tile_count(root, tile_roots);
}

// 7. instance id = rank of the root run in raster order (== scipy.ndimage.label's numbering, aggregation_layer.py:178)
__global__ void __launch_bounds__(1024) k_run_assign(RunTables RT, const int *__restrict__ counters, const int *__restrict__ tile_base,
                                                     InstTables T, int max_instances, long long cap) {
    __shared__ int s_w[32];
    const int M = (counters[FPC_CNT_FLAGS] & FPC_FLAG_ROWS) ? 0 : (int)min((long long)counters[FPC_CNT_ROWS], cap);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int m = blockIdx.x * TILE + threadIdx.x;
    const bool root = (m < M) && (RT.parent[m] == m);
    const unsigned bal = __ballot_sync(FULL, root);
    if (lane == 0) s_w[wid] = __popc(bal);
    __syncthreads();
    if (!root) return;
    int id = tile_base[blockIdx.x] + __popc(bal & ((1u << lane) - 1u));
    for (int k = 0; k < wid; ++k) id += s_w[k];
    RT.inst[m] = id;
    if (id < max_instances) {
        T.root[id] = RT.start[m];
        T.count[id] = 0;
        T.nruns[id] = 0;
        T.ymin[id] = INT_MAX; T.ymax[id] = -1; T.xmin[id] = INT_MAX; T.xmax[id] = -1;
        T.mincls[id] = INT_MAX;
        T.rmax2[id] = 0;
    }
}

// 8. per-instance pixel count, run count, bounding box and minimum class id; every run learns its instance.
//    dense != 0: the instance of a run is the plane it lies in (voting drop-ins: no connected components).
__global__ void __launch_bounds__(256) k_run_stats(RunTables RT, const int *__restrict__ counters, InstTables T, int w, int hw,
                                                   int max_instances, long long cap, int dense) {
    if (counters[FPC_CNT_FLAGS]) return;
    const int M = (int)min((long long)counters[FPC_CNT_ROWS], cap);
    for (int m = blockIdx.x * blockDim.x + threadIdx.x; m < M; m += gridDim.x * blockDim.x) {
        const int s = RT.start[m], e = RT.end[m];
        const int id = dense ? s / hw : RT.inst[RT.parent[m]];
        RT.inst[m] = id;
        if ((unsigned)id >= (unsigned)max_instances) continue;
        const int pix = s % hw;
        const int y = pix / w, x0 = pix - y * w, x1 = x0 + (e - s);
        atomicAdd(&T.count[id], e - s + 1);
        atomicAdd(&T.nruns[id], 1);
        atomicMin(&T.xmin[id], x0); atomicMax(&T.xmax[id], x1);
        atomicMin(&T.ymin[id], y); atomicMax(&T.ymax[id], y);
    }
}

// 9. run slots of every instance -> rowoff (exclusive scan over instances, single block)
__global__ void __launch_bounds__(1024) k_scan_slots(InstTables T, int *counters, int max_instances) {
    const int N = min(counters[FPC_CNT_INSTANCES], max_instances);
    for (int i = threadIdx.x; i < N; i += blockDim.x) T.rowoff[i] = T.nruns[i];
    __syncthreads();
    const int total = block_exclusive_scan_inplace(T.rowoff, N);
    if (threadIdx.x == 0) T.rowoff[N] = total;
}

// 10. one block per instance: its runs in raster order (found through the per-row run index), the number of voting
//     pixels before each of them (-> position of the run's first voting record), tn, and the zeroing of the instance's
//     vote counters.  ransac_voting_gpu.py:536-545: fewer than min_num pixels -> the instance does not vote; more than
//     max_num -> Bernoulli(max_num / count) sub-sampling.
__global__ void __launch_bounds__(256) k_run_slots(RunTables RT, InstTables T, RowTables R, const int *__restrict__ counters,
                                                   PathParams pp, int *__restrict__ votes, const uint8_t *__restrict__ cls) {
    __shared__ int s_w[8];
    __shared__ int s_carry[2];
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    const int tid = threadIdx.x, lane = tid & 31, wv = tid >> 5;
    for (int i = blockIdx.x; i < N; i += gridDim.x) {
        const int slot0 = T.rowoff[i];
        const int cnt = T.count[i], ymin = T.ymin[i], ymax = T.ymax[i];
        const int img = T.root[i] / pp.hw;
        const bool votes_at_all = cnt >= pp.min_num;
        const bool sub = cnt > pp.max_num;
        const float thr = (float)pp.max_num / (float)cnt;
        const int flags = ROW_CONTIG | (sub ? ROW_SUB : 0) | (votes_at_all ? ROW_VOTES : 0);
        for (int k = tid; k < pp.hn; k += 256) votes[(size_t)i * pp.hn + k] = 0;
        if (tid == 0) { s_carry[0] = 0; s_carry[1] = 0; }      // (slots, voting pixels) of the rows handled so far
        // farthest pixel from the centre of the bounding box (an end of some run): the vote kernel bounds |hypothesis - pixel|
        // with it (tighter than the box diagonal: 36 instead of 51 pixels for a disc of radius 35)
        const float ox = (float)((T.xmin[i] + T.xmax[i] + 1) >> 1), oy = (float)((ymin + ymax + 1) >> 1);
        float r2max = 0.f;
        __syncthreads();
        for (int yb = ymin; yb <= ymax; yb += 256) {
            // one thread per image row: own runs of this row and their voting pixels
            const int y = yb + tid;
            int nr = 0, nv = 0, m0 = 0, m1 = 0;
            if (y <= ymax) {
                m0 = RT.rowrun[img * pp.h + y];
                m1 = RT.rowrun[img * pp.h + y + 1];
                for (int m = m0; m < m1; ++m)
                    if (RT.inst[m] == i) {
                        ++nr;
                        {
                            const int xa = (RT.start[m] - img * pp.hw) - y * pp.w, xb = xa + (RT.end[m] - RT.start[m]);
                            const float dy = (float)y - oy, da = (float)xa - ox, db = (float)xb - ox;
                            r2max = fmaxf(r2max, fmaf(dy, dy, fmaxf(da * da, db * db)));
                        }
                        if (votes_at_all) {
                            if (!sub) nv += RT.end[m] - RT.start[m] + 1;
                            else for (int q = RT.start[m]; q <= RT.end[m]; ++q) nv += select_uniform(pp, q) < thr ? 1 : 0;
                        }
                    }
            }
            // exclusive scans of (nr, nv) over the 256 rows of this pass
            const int inr = warp_incl_scan(nr, lane), inv = warp_incl_scan(nv, lane);
            if (lane == 31) s_w[wv] = inr;
            __syncthreads();
            int slot = s_carry[0] + inr - nr;
            for (int k = 0; k < wv; ++k) slot += s_w[k];
            const int tot_r = s_w[0] + s_w[1] + s_w[2] + s_w[3] + s_w[4] + s_w[5] + s_w[6] + s_w[7];
            __syncthreads();
            if (lane == 31) s_w[wv] = inv;
            __syncthreads();
            int pref = s_carry[1] + inv - nv;
            for (int k = 0; k < wv; ++k) pref += s_w[k];
            const int tot_v = s_w[0] + s_w[1] + s_w[2] + s_w[3] + s_w[4] + s_w[5] + s_w[6] + s_w[7];
            __syncthreads();
            if (tid == 0) { s_carry[0] += tot_r; s_carry[1] += tot_v; }
            if (y <= ymax) {
                for (int m = m0; m < m1; ++m)
                    if (RT.inst[m] == i) {
                        const int s = RT.start[m], len = RT.end[m] - s + 1;
                        R.desc[slot0 + slot] = make_int4(i, s, len | ((int)cls[s] << ROW_CLS_SHIFT) | flags, pref);
                        ++slot;
                        if (votes_at_all) {
                            if (!sub) pref += len;
                            else for (int q = s; q < s + len; ++q) pref += select_uniform(pp, q) < thr ? 1 : 0;
                        }
                    }
            }
            __syncthreads();
        }
        if (tid == 0) T.tn[i] = s_carry[1];
        r2max = __int_as_float(__reduce_max_sync(FULL, __float_as_int(r2max)));   // non-negative floats order like their bit patterns
        if (lane == 0 && r2max > 0.f) atomicMax(&T.rmax2[i], __float_as_int(r2max));
        __syncthreads();
    }
}

// instance ids per pixel (the scipy label volume, 0 = background) painted from the run list; only on request
__global__ void __launch_bounds__(256) k_relabel(RunTables RT, const int *__restrict__ counters, int *__restrict__ out, long long cap) {
    if (counters[FPC_CNT_FLAGS]) return;
    const int M = (int)min((long long)counters[FPC_CNT_ROWS], cap);
    const int lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int m = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; m < M; m += nwarps) {
        const int s = RT.start[m], e = RT.end[m], v = RT.inst[m] + 1;
        for (int q = s + lane; q <= e; q += 32) out[q] = v;
    }
}

__global__ void __launch_bounds__(256) k_dense_init(InstTables T, int *counters, int hw, int nprob) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j == 0) counters[FPC_CNT_INSTANCES] = nprob;
    if (j >= nprob) return;
    T.root[j] = j * hw;   // plane index = root / hw
    T.count[j] = 0;
    T.nruns[j] = 0;
    T.xmin[j] = INT_MAX; T.xmax[j] = -1; T.ymin[j] = INT_MAX; T.ymax[j] = -1;
    T.mincls[j] = 0;
    T.rmax2[j] = 0;
}
// empty problems: a zero-height row range so that the slot kernel has nothing to walk
__global__ void __launch_bounds__(256) k_dense_fix_empty(InstTables T, int nprob) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= nprob) return;
    if (T.count[j] == 0) { T.xmin[j] = 0; T.xmax[j] = -1; T.ymin[j] = 0; T.ymax[j] = -1; }
}

// =============================================================================================
// host-side launch sequences
// =============================================================================================
static int launch_run_tables(const Workspace &ws, const PathParams &pp, int span, bool dense, int nprob, cudaStream_t st) {
    const int P = pp.P;
    const int ntiles = ceil_div(P, TILE);
    const long long cap = pp.max_rows;
    k_scan_tiles<<<1, 1024, 0, st>>>(ws.tile_roots, ntiles, ws.counters, FPC_CNT_ROWS, cap, FPC_FLAG_ROWS, 1);
    FPC_LAUNCH_CHECK("k_scan_tiles");
    if (pp.w % 16 == 0 && span % 16 == 0 && (reinterpret_cast<uintptr_t>(ws.cls) & 15) == 0)
        k_emit_runs16<<<ntiles, 64, 0, st>>>(ws.cls, ws.tile_roots, ws.RT, pp.w, P, span, cap);
    else
        k_emit_runs<<<ntiles, 256, 0, st>>>(ws.cls, ws.tile_roots, ws.RT, pp.w, P, span, cap);
    FPC_LAUNCH_CHECK("k_emit_runs");
    const int rgrid = sm_count() * 8;
    const int rtiles = ceil_div(std::min<long long>(cap, (long long)P), TILE);
    if (!dense) {
        k_run_merge<<<rgrid, 256, 0, st>>>(ws.RT, ws.counters, pp.w, pp.h, cap);
        FPC_LAUNCH_CHECK("k_run_merge");
        k_run_flatten<<<rtiles, 1024, 0, st>>>(ws.RT, ws.counters, ws.run_tiles, cap);
        FPC_LAUNCH_CHECK("k_run_flatten");
        k_scan_tiles<<<1, 1024, 0, st>>>(ws.run_tiles, rtiles, ws.counters, FPC_CNT_INSTANCES, pp.max_instances,
                                         FPC_FLAG_INSTANCES, 0);
        FPC_LAUNCH_CHECK("k_scan_roots");
        k_run_assign<<<rtiles, 1024, 0, st>>>(ws.RT, ws.counters, ws.run_tiles, ws.T, pp.max_instances, cap);
        FPC_LAUNCH_CHECK("k_run_assign");
    } else {
        k_dense_init<<<ceil_div(std::max(nprob, 1), 256), 256, 0, st>>>(ws.T, ws.counters, pp.hw, nprob);
        FPC_LAUNCH_CHECK("k_dense_init");
    }
    k_run_stats<<<rgrid, 256, 0, st>>>(ws.RT, ws.counters, ws.T, pp.w, pp.hw, pp.max_instances, cap, dense ? 1 : 0);
    FPC_LAUNCH_CHECK("k_run_stats");
    if (dense && nprob > 0) {
        k_dense_fix_empty<<<ceil_div(nprob, 256), 256, 0, st>>>(ws.T, nprob);
        FPC_LAUNCH_CHECK("k_dense_fix_empty");
    }
    k_scan_slots<<<1, 1024, 0, st>>>(ws.T, ws.counters, pp.max_instances);
    FPC_LAUNCH_CHECK("k_scan_slots");
    return FPC_OK;
}

int launch_label_and_tables(const Workspace &ws, const PathParams &pp, const float *mask_logits,
                            const long long *cat_mask_i64, cudaStream_t st) {
    const int P = pp.P;
    const int ntiles = ceil_div(P, TILE);
    int span;
    if (mask_logits && pp.up.s > 1) {
        if (pp.w % 4 == 0 && pp.up.s >= 3 && (reinterpret_cast<uintptr_t>(ws.cls) & 3) == 0) {
            if (pp.num_classes == 7)   // the reference's class count (6 objects + background): unrolled, background skip
                k_argmax_runs_up4<7><<<ntiles, 256, 0, st>>>(mask_logits, ws.cls, ws.tile_roots, 7, pp.hw, pp.w, P / 4, pp.up);
            else
                k_argmax_runs_up4<0><<<ntiles, 256, 0, st>>>(mask_logits, ws.cls, ws.tile_roots, pp.num_classes, pp.hw, pp.w, P / 4, pp.up);
            span = 128;
        } else {
            k_argmax_runs_up1<<<ntiles, 1024, 0, st>>>(mask_logits, ws.cls, ws.tile_roots, pp.num_classes, pp.hw, pp.w, P, pp.up);
            span = 32;
        }
        FPC_LAUNCH_CHECK("k_argmax_runs");
    } else if (mask_logits) {
        const bool vec_ok = (pp.w % 4 == 0) && ((reinterpret_cast<uintptr_t>(mask_logits) & 15) == 0) &&
                            ((reinterpret_cast<uintptr_t>(ws.cls) & 3) == 0);
        if (vec_ok) {
            // persistent pipelined kernel (default); FPC_ARGMAX_MODE=0 selects the one-tile-per-block kernel (A/B measurements)
            static int mode = -1, p4_blocks = 2;
            if (mode < 0) {
                const char *e = getenv("FPC_ARGMAX_MODE");
                mode = e ? atoi(e) : 2;
                const char *bpsm = getenv("FPC_ARGMAX_BLOCKS_PER_SM");
                if (bpsm && atoi(bpsm) >= 1) p4_blocks = atoi(bpsm);
            }
            if (mode != 0 && pp.num_classes == 7 && P % TILE == 0)
                k_argmax_runs_p4<7><<<std::min(ntiles, sm_count() * p4_blocks), 256, 0, st>>>(mask_logits, ws.cls, ws.tile_roots, pp.hw,
                                                                                             pp.w, ntiles);
            else
                k_argmax_runs_v4<<<ntiles, 256, 0, st>>>(mask_logits, ws.cls, ws.tile_roots, pp.num_classes, pp.hw, pp.w, P / 4);
            span = 128;
        } else {
            k_argmax_runs_v1<<<ntiles, 1024, 0, st>>>(mask_logits, ws.cls, ws.tile_roots, pp.num_classes, pp.hw, pp.w, P);
            span = 32;
        }
        FPC_LAUNCH_CHECK("k_argmax_runs");
    } else {
        k_cls_runs<<<ntiles, 1024, 0, st>>>(cat_mask_i64, nullptr, nullptr, 1, 0, ws.cls, ws.tile_roots, pp.hw, pp.w, P);
        FPC_LAUNCH_CHECK("k_cls_runs");
        span = 1 << 30;
    }
    return launch_run_tables(ws, pp, span, /*dense=*/false, 0, st);
}

int launch_dense_problems(const Workspace &ws, const PathParams &pp, const float *fmask, const int *imask,
                          int nplanes_per_src, int match_base, int nprob, cudaStream_t st) {
    const int P = pp.P;
    const int ntiles = ceil_div(P, TILE);
    k_cls_runs<<<ntiles, 1024, 0, st>>>(nullptr, fmask, imask, nplanes_per_src, match_base, ws.cls, ws.tile_roots, pp.hw, pp.w, P);
    FPC_LAUNCH_CHECK("k_cls_runs");
    return launch_run_tables(ws, pp, 1 << 30, /*dense=*/true, nprob, st);
}

int launch_slots(const Workspace &ws, const PathParams &pp, cudaStream_t st) {
    k_run_slots<<<sm_count() * 8, 256, 0, st>>>(ws.RT, ws.T, ws.R, ws.counters, pp, ws.votes, ws.cls);
    FPC_LAUNCH_CHECK("k_run_slots");
    return FPC_OK;
}

int launch_relabel(const Workspace &ws, const PathParams &pp, int *labels_out, cudaStream_t st) {
    FPC_CUDA_TRY(cudaMemsetAsync(labels_out, 0, (size_t)pp.P * sizeof(int), st));
    k_relabel<<<sm_count() * 8, 256, 0, st>>>(ws.RT, ws.counters, labels_out, pp.max_rows);
    FPC_LAUNCH_CHECK("k_relabel");
    return FPC_OK;
}

}  // namespace fpc
