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
#include <cstdlib>

namespace fpc {

// I2. record offsets and vote work items per instance (single block)
__global__ void __launch_bounds__(1024) k_scan_records(InstTables T, int *counters, long long max_records, int chunk,
                                                       int nbatch, int tail_div) {
    if (counters[FPC_CNT_FLAGS]) return;
    const int N = counters[FPC_CNT_INSTANCES];
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const int tn = T.tn[i];
        T.pxoff[i] = (tn + 15) & ~15;                        // ranges padded to whole 16-pixel voting rounds
        const int ipx = vote_item_px(i, N, chunk, tail_div);
        T.workoff[i] = ((tn + ipx - 1) / ipx) * nbatch;
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
// MODE 3: like 0, but the head maps are LOW RESOLUTION and every value is their x S bilinear up-sampling evaluated
//         here (head-epilogue fusion: 4 cached taps + 6 FP32 operations per channel, foreground pixels only)
template <int MODE>
__global__ void __launch_bounds__(256, 4) k_gather(const uint8_t *__restrict__ cls,
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
        const bool sub = d.z & ROW_SUB, votes = d.z & ROW_VOTES;
        const int img = p0 / pp.hw;
        const int pix0 = p0 - img * pp.hw;
        const int y = pix0 / pp.w, x0 = pix0 - y * pp.w;
        const float thr = sub ? (float)pp.max_num / (float)T.count[i] : 2.f;
        const int rec0 = (want_rec && votes) ? T.pxoff[i] + d.w : 0;
        LerpCoord LY{0, 0, 0.f, 0.f};
        if (MODE == 3) LY = lerp_coord(y, pp.up.sy, pp.up.hl);
        // MODE 0: bases of the run's first pixel in the four head maps (class 1, channel 0)
        const float *qrow = nullptr, *srow = nullptr, *vrow = nullptr, *zrow = nullptr;
        if (MODE == 0) {
            qrow = F.quaternion + (size_t)img * 4 * K * hw + pix0;
            srow = F.scales + (size_t)img * 3 * K * hw + pix0;
            vrow = F.xy + (size_t)img * 2 * K * hw + pix0;
            zrow = F.z + (size_t)img * K * hw + pix0;
        }
        int running = 0, cmin = INT_MAX;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
        for (int kb = 0; kb < len; kb += 32) {
            const int kx = kb + lane;
            const int p = p0 + kx;
            const bool mem = kx < len;   // slots are contiguous runs
            float vx = 0.f, vy = 0.f;
            // MODE 3 (low-resolution heads): the predicted class's 10 values of this pixel (q0..q3, s0..s2, z, dir_x, dir_y),
            // each the x S bilinear interpolation of its low-res plane.  The 32 pixels of a warp iteration span <= 9 low-res
            // columns for S >= 4, so the 2 x 16 taps of a plane are loaded ONCE by the warp (lane = row*16 + column) and
            // handed to the pixels by shuffles: 1 load + 4 shuffles per channel instead of 4 loads and their address
            // arithmetic.  That needs one class for the whole iteration; pixels of touching blobs of different classes (and
            // S < 4) take the per-lane path.  Both paths feed identical operands to bilerp(): identical bits.
            float lr[10];
            int lr_cp = 0;
            if (MODE == 3) {
                lr_cp = mem ? (int)cls[p] : 0;
                const LerpCoord LX = lerp_coord(min(x0 + kx, pp.w - 1), pp.up.sx, pp.up.wl);
                const int wl = pp.up.wl;
                const size_t lhw = (size_t)pp.up.hl * wl;
                const unsigned members = __ballot_sync(FULL, mem);           // lane 0 is always a member
                const int c_first = __shfl_sync(FULL, LX.i0, 0);
                const int c_last = __shfl_sync(FULL, LX.i1, 31 - __clz(members));
                const int cp0 = __shfl_sync(FULL, lr_cp, 0);
                const bool uniform = __all_sync(FULL, !mem || lr_cp == cp0);
                if (uniform && c_last - c_first < 16) {
                    const int toff = ((lane >> 4) ? LY.i1 : LY.i0) * wl + min(c_first + (lane & 15), wl - 1);
                    const int s00 = LX.i0 - c_first, s01 = LX.i1 - c_first;  // source lanes of this pixel's row-0 taps; +16: row 1
                    const size_t kc = (size_t)img * K + (cp0 - 1);           // (image, predicted class) -> channel group
                    const float *const planes[4] = {F.quaternion + 4 * kc * lhw, F.scales + 3 * kc * lhw, F.z + kc * lhw,
                                                    F.xy + 2 * kc * lhw};
                    int o = 0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int nch = g == 0 ? 4 : g == 1 ? 3 : g == 2 ? 1 : 2;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            if (c < nch) {
                                const float t = __ldg(planes[g] + (size_t)c * lhw + toff);
                                const float v00 = __shfl_sync(FULL, t, s00), v01 = __shfl_sync(FULL, t, s01);
                                const float v10 = __shfl_sync(FULL, t, s00 + 16), v11 = __shfl_sync(FULL, t, s01 + 16);
                                lr[o++] = bilerp(v00, v01, v10, v11, LX.w0, LX.w1, LY.w0, LY.w1);
                            }
                        }
                    }
                } else if (mem) {
                    const int o00 = LY.i0 * wl + LX.i0, o01 = LY.i0 * wl + LX.i1, o10 = LY.i1 * wl + LX.i0, o11 = LY.i1 * wl + LX.i1;
                    auto tap = [&](const float *__restrict__ plane) {
                        return bilerp(__ldg(plane + o00), __ldg(plane + o01), __ldg(plane + o10), __ldg(plane + o11),
                                      LX.w0, LX.w1, LY.w0, LY.w1);
                    };
                    const size_t kc = (size_t)img * K + (lr_cp - 1);
                    const float *q = F.quaternion + 4 * kc * lhw, *sp = F.scales + 3 * kc * lhw, *v = F.xy + 2 * kc * lhw;
                    lr[0] = tap(q); lr[1] = tap(q + lhw); lr[2] = tap(q + 2 * lhw); lr[3] = tap(q + 3 * lhw);
                    lr[4] = tap(sp); lr[5] = tap(sp + lhw); lr[6] = tap(sp + 2 * lhw);
                    lr[7] = tap(F.z + kc * lhw);
                    lr[8] = tap(v); lr[9] = tap(v + lhw);
                }
            }
            if (mem) {
                const size_t pix = (size_t)(pix0 + kx);
                float q0 = 0.f, q1 = 0.f, q2 = 0.f, q3 = 0.f, s0 = 0.f, s1 = 0.f, s2 = 0.f, zz = 0.f;
                if (MODE == 2) {
                    const float *v = F.xy + (long long)(img / F.div) * F.sN + (long long)y * F.sH + (long long)(x0 + kx) * F.sW;
                    vx = v[0];
                    vy = v[F.s2];
                } else if (MODE == 0 || MODE == 3) {
                    const int cp = MODE == 3 ? lr_cp : (int)cls[p];
                    cmin = min(cmin, cp);
                    if (MODE == 0) {
                        // predicted class of THIS pixel (class_compress is per pixel); 32-bit offsets from per-run bases
                        const unsigned o = (unsigned)(cp - 1) * (unsigned)pp.hw;
                        const float *q = qrow + (4u * o + (unsigned)kx), *s = srow + (3u * o + (unsigned)kx);
                        const float *v = vrow + (2u * o + (unsigned)kx);
                        q0 = __ldcs(q); q1 = __ldcs(q + hw); q2 = __ldcs(q + 2 * hw); q3 = __ldcs(q + 3 * hw);
                        s0 = __ldcs(s); s1 = __ldcs(s + hw); s2 = __ldcs(s + 2 * hw);
                        zz = __ldcs(zrow + (o + (unsigned)kx));
                        vx = __ldcs(v); vy = __ldcs(v + hw);
                    } else {
                        q0 = lr[0]; q1 = lr[1]; q2 = lr[2]; q3 = lr[3];
                        s0 = lr[4]; s1 = lr[5]; s2 = lr[6]; zz = lr[7];
                        vx = lr[8]; vy = lr[9];
                    }
                    // quaternion: only its masked mean is used (1e-4 budget) -> one reciprocal, four multiplies;
                    // direction: feeds the votes -> keep the reference's value / norm with IEEE sqrt and divide
                    const float qq = q0 * q0 + q1 * q1 + q2 * q2 + q3 * q3;
                    if (qq != 0.f) { const float rq = rsqrtf(qq); q0 *= rq; q1 *= rq; q2 *= rq; q3 *= rq; }
                    const float vn = torch_norm2(vx, vy);
                    if (vn != 0.f) { vx = __fdiv_rn(vx, vn); vy = __fdiv_rn(vy, vn); }
                } else {
                    // already class-compressed CategoricalData (lib/type_hinting.py:12-17): [b,4|3|2,h,w], z [b,h,w]
                    cmin = min(cmin, (int)cls[p]);
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
                const int idx = rec0 + running + __popc(bal & ((1u << lane) - 1u));
                rec.x[idx] = (float)(x0 + kx); rec.y[idx] = (float)y; rec.nx[idx] = vx; rec.ny[idx] = vy;
            }
            running += __popc(bal);
        }
        if (MODE != 2) {
            // class id of the instance = min class over its pixels (aggregation_layer.py:113): one atomic per run
            const int cm = __reduce_min_sync(FULL, cmin);
            if (lane == 0 && cm != INT_MAX) atomicMin(&T.mincls[i], cm);
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
int launch_rows_and_records(const Workspace &ws, const PathParams &pp, const FieldSrc &F, int gather_mode,
                            bool want_records, int vote_chunk, cudaStream_t st) {
    const int grid = sm_count() * 32;   // one warp per (instance,row) item, grid-stride: plenty of loads in flight
    int rc = launch_slots(ws, pp, st);
    if (rc != FPC_OK) return rc;
    k_scan_records<<<1, 1024, 0, st>>>(ws.T, ws.counters, pp.max_records, vote_chunk, vote_batches(pp.hn), pp.vote_tail);
    FPC_LAUNCH_CHECK("k_scan_records");
    if (gather_mode == 0)
        k_gather<0><<<grid, 256, 0, st>>>(ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    else if (gather_mode == 1)
        k_gather<1><<<grid, 256, 0, st>>>(ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    else if (gather_mode == 2)
        k_gather<2><<<grid, 256, 0, st>>>(ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    else
        k_gather<3><<<grid, 256, 0, st>>>(ws.cls, ws.T, ws.R, ws.counters, pp, F, ws.rec, want_records);
    FPC_LAUNCH_CHECK("k_gather");
    return FPC_OK;
}

}  // namespace fpc
