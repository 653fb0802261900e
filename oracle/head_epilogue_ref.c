/* ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).
 *
 * CPU restatement of the x S bilinear up-sampling of smp's SegmentationHead, the step the reference runs between its
 * heads' 1x1 convolutions and the pose-recovery path (lib/pose_regressor.py:633-666: SegmentationHead(...,
 * kernel_size=1, upsampling=4) = Conv2d -> nn.UpsamplingBilinear2d(scale_factor=4) -> identity; align_corners=True).
 *
 * The arithmetic is ATen's (third party, torch 2.11 in this image; the reference pins pytorch 1.8.0 / 1.7.1,
 * environment_linux.yaml:48,114 -- the coordinate rule below is unchanged since 1.5):
 *   ATen/native/UpSample.h   area_pixel_compute_scale:        scale = float(in - 1) / (out - 1)   (0 if out == 1)
 *                            area_pixel_compute_source_index: src   = scale * dst
 *                            guard_index_and_lambda:          i0 = min(floor(src), in - 1), w1 = clamp(src - i0, 0, 1)
 *                            compute_source_index_and_lambda: i1 = i0 + (i0 < in - 1), w0 = 1 - w1
 *   ATen/native/cpu/UpSampleKernel.cpp  Interpolate<2>::eval:  out = row(i0y) * wy0 + row(i1y) * wy1,
 *                                                              row(r) = in[r][i0x] * wx0 + in[r][i1x] * wx1
 * with every "a*wa + b*wb" evaluated as fma(a, wa, b*wb): that is what the AVX2/AVX-512 build of the CPU kernel does
 * (tests/test_head_epilogue_oracle.py pins this function bit-for-bit against torch.nn.UpsamplingBilinear2d on the
 * CPU) and what nvcc's contraction makes of the CUDA kernel's
 *   h0lambda * (w0lambda * v00 + w1lambda * v01) + h1lambda * (w0lambda * v10 + w1lambda * v11)
 * (ATen/native/cuda/UpSampleBilinear2d.cu; pinned on the GPU box in tests/test_head_epilogue_gpu.py).
 * This file is compiled with -ffp-contract=off, so the only fused operations are the explicit fmaf calls.
 */
#include <math.h>
#include <stddef.h>

typedef struct {
    int i0, i1;
    float w0, w1;
} coord_t;

static coord_t source_coord(int dst, float scale, int in_size)
{
    coord_t c;
    const float src = scale * (float)dst;
    int i0 = (int)floorf(src);
    if (i0 > in_size - 1) i0 = in_size - 1;
    float w1 = src - (float)i0;
    if (w1 < 0.f) w1 = 0.f;
    if (w1 > 1.f) w1 = 1.f;
    c.i0 = i0;
    c.i1 = i0 + (i0 < in_size - 1 ? 1 : 0);
    c.w1 = w1;
    c.w0 = 1.f - w1;
    return c;
}

void fpc_ref_upsample_bilinear(const float *in, /* [planes, hl, wl] */
                               long long planes, int hl, int wl, int scale,
                               float *out /* [planes, hl*scale, wl*scale] */)
{
    const int h = hl * scale, w = wl * scale;
    const float sy = h > 1 ? (float)(hl - 1) / (float)(h - 1) : 0.f;
    const float sx = w > 1 ? (float)(wl - 1) / (float)(w - 1) : 0.f;
#pragma omp parallel for schedule(static)
    for (long long row = 0; row < planes * h; ++row) {
        const long long plane = row / h;
        const int y = (int)(row - plane * h);
        const coord_t cy = source_coord(y, sy, hl);
        const float *r0 = in + ((size_t)plane * hl + cy.i0) * wl;
        const float *r1 = in + ((size_t)plane * hl + cy.i1) * wl;
        float *dst = out + (size_t)row * w;
        for (int x = 0; x < w; ++x) {
            const coord_t cx = source_coord(x, sx, wl);
            const float t0 = fmaf(r0[cx.i0], cx.w0, r0[cx.i1] * cx.w1);
            const float t1 = fmaf(r1[cx.i0], cx.w0, r1[cx.i1] * cx.w1);
            dst[x] = fmaf(t0, cy.w0, t1 * cy.w1);
        }
    }
}
