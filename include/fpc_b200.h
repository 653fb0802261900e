/*
 * fpc_b200.h -- C ABI of libfpc_b200.so: FastPoseCNN's post-network pose-recovery
 * path as hand-written sm_100a CUDA kernels.
 *
 * Conventions:
 *   - every pointer is a DEVICE pointer unless its name ends in `_host`;
 *   - the library allocates nothing, keeps nothing: the caller owns inputs, outputs
 *     and the workspace (query its size with the matching *_workspace_bytes call);
 *   - every launch goes to the `stream` argument (a cudaStream_t passed as void*);
 *     no call synchronises, so every call can be captured in a CUDA graph;
 *   - every function returns FPC_OK or a negative FPC_E* code and never exits or
 *     throws (the reference's gpuErrchk calls exit(), src/cuda_common.h:19-26);
 *     fpc_last_error() returns the message of the calling thread's last failure;
 *   - there is no CPU fallback.
 *
 * File:line citations below are relative to
 * /root/reference/source_code/FastPoseCNN/ and name the reference interface each
 * entry point replaces.
 */
#ifndef FPC_B200_H_
#define FPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPC_VERSION 100 /* 0.1.0 */

#if defined(__GNUC__)
#define FPC_API __attribute__((visibility("default")))
#else
#define FPC_API
#endif

enum {
    FPC_OK = 0,
    FPC_EINVAL = -1,    /* bad argument (null pointer, non-positive size, misalignment) */
    FPC_ECUDA = -2,     /* a CUDA runtime call or launch failed                        */
    FPC_ECAPACITY = -3, /* workspace / table capacity too small for this call          */
};

/* Arithmetic of the hypothesis and voting kernels.  Both are IEEE binary32.
 *   FPC_ARITH_IEEE      every product and sum rounded separately, in source order
 *                       (what a CPU build of ransac_voting_kernel.cu computes; the oracle)
 *   FPC_ARITH_NVCC_FMA  the contraction pattern nvcc applies to the same source for
 *                       sm_100a (a*b + c*d -> fma(a, b, c*d)); matches a GPU build of the
 *                       reference kernels */
enum { FPC_ARITH_IEEE = 0, FPC_ARITH_NVCC_FMA = 1 };

FPC_API int fpc_version(void);
FPC_API const char *fpc_last_error(void);

/* ------------------------------------------------------------------------------------
 * 1:1 mirrors of the reference's native module `ransac_voting`
 * ---------------------------------------------------------------------------------- */

/* lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:20-31 (kernel: ransac_voting_kernel.cu:11-49).
 * direct [tn,vn,2] f32, coords [tn,2] f32, idxs [hn,vn,2] i32 -> hypo_pts [hn,vn,2] f32.
 * Degenerate pairs (|det| < 1e-6) yield (0,0), as the reference's at::zeros output does. */
FPC_API int fpc_generate_hypothesis(const float *direct, const float *coords, const int32_t *idxs,
                            float *hypo_pts, int tn, int vn, int hn, int arith, void *stream);

/* lib/ransac_voting_gpu_layer/src/ransac_voting.cpp:41-55 (kernel: ransac_voting_kernel.cu:88-126).
 * Sets inliers[hi,vi,ti] = 1 where the cosine test passes; like the reference it never
 * writes zeros (the caller pre-zeroes, ransac_voting_gpu.py:562). */
FPC_API int fpc_voting_for_hypothesis(const float *direct, const float *coords, const float *hypo_pts,
                              uint8_t *inliers, int tn, int vn, int hn, float inlier_thresh,
                              int arith, void *stream);

/* The vanishing-point twins of the two functions above -- the other half of the reference's pybind module
 * (src/ransac_voting.cpp:62-73 generate_hypothesis_vanishing_point, :83-97 voting_for_hypothesis_vanishing_point,
 * kernels src/ransac_voting_kernel.cu:170-228, 268-308).  hypo_pts is [hn,vn,3]: homogeneous (x, y, z), all zero
 * when the two rays do not meet; a pixel votes when its ray points towards the hypothesis and |cos| > thresh. */
FPC_API int fpc_generate_hypothesis_vanishing_point(const float *direct, const float *coords, const int32_t *idxs, float *hypo_pts,
                                                    int tn, int vn, int hn, int arith, void *stream);
FPC_API int fpc_voting_for_hypothesis_vanishing_point(const float *direct, const float *coords, const float *hypo_pts,
                                                      uint8_t *inliers, int tn, int vn, int hn, float inlier_thresh, int arith,
                                                      void *stream);

/* ------------------------------------------------------------------------------------
 * Stage entry points behind the reference's Python operators
 * ---------------------------------------------------------------------------------- */

/* gpu_tensor_funcs.normalize (lib/gpu_tensor_funcs.py:37-50) for a contiguous
 * [outer, c, inner] view normalised over the middle axis. */
FPC_API int fpc_normalize(const float *in, float *out, long long outer, int c, long long inner, void *stream);

/* Model.class_compression (lib/pose_regressor.py:445-457) = arg-max over the mask logits
 * + gpu_tensor_funcs.class_compress (lib/gpu_tensor_funcs.py:52-99).
 *   mask_logits [b,C,h,w] (may be NULL when cat_mask_in is given)
 *   cat_mask_in [b,h,w] i64 or NULL (NULL: computed from mask_logits and written to cat_mask_out)
 *   heads: quaternion [b,4(C-1),h,w], scales [b,3(C-1),h,w], xy [b,2(C-1),h,w], z [b,C-1,h,w]
 *   outputs: cat_mask_out [b,h,w] i64 (may be NULL), q_out [b,4,h,w], s_out [b,3,h,w],
 *            xy_out [b,2,h,w], z_out [b,h,w]; q and xy are L2-normalised per pixel. */
FPC_API int fpc_class_compress(const float *mask_logits, const int64_t *cat_mask_in,
                       const float *quaternion, const float *scales, const float *xy, const float *z,
                       int64_t *cat_mask_out, float *q_out, float *s_out, float *xy_out, float *z_out,
                       int b, int num_classes, int h, int w, void *stream);

/* gpu_tensor_funcs.batchwise_get_RT / samplewise_get_RT (lib/gpu_tensor_funcs.py:204-253)
 * with quats_2_rotation_matrix (:306-326).  q [n,4], xy [n,2], z [n,1], inv_k [3,3]
 * -> R [n,3,3], T [n,3], RT [n,4,4]. */
FPC_API int fpc_get_rt(const float *q, const float *xy, const float *z, const float *inv_k,
               float *R, float *T, float *RT, int n, void *stream);

/* ------------------------------------------------------------------------------------
 * The fused path: head maps -> per-instance pose table
 * (Model.class_compression + aggregate + hough_voting + perform_RT_calculation,
 *  lib/pose_regressor.py:445-504)
 * ---------------------------------------------------------------------------------- */

#define FPC_POSE_ROW 48 /* 32-bit words per pose-table row */
/* word offsets inside a pose-table row (ints are stored as int32 bit patterns) */
enum {
    FPC_ROW_CLASS = 0,   /* i32  class id = min non-zero class in the component (aggregation_layer.py:113) */
    FPC_ROW_SAMPLE = 1,  /* i32  frame index (sample_ids, :91-98)                                           */
    FPC_ROW_COUNT = 2,   /* i32  pixels in the instance mask                                                */
    FPC_ROW_Q = 3,       /* f32x4 normalised mean quaternion                                               */
    FPC_ROW_SCALES = 7,  /* f32x3 mean scales                                                              */
    FPC_ROW_XY = 10,     /* f32x2 refined centre (x = column, y = row)                                     */
    FPC_ROW_Z = 12,      /* f32   exp(mean z)                                                              */
    FPC_ROW_T = 13,      /* f32x3                                                                          */
    FPC_ROW_R = 16,      /* f32x9 row-major                                                                */
    FPC_ROW_RT = 25,     /* f32x16 row-major                                                               */
    FPC_ROW_HYP = 41,    /* f32x2 winning (unrefined) hypothesis                                           */
    FPC_ROW_WIN_IDX = 43,   /* i32 index of the winning hypothesis (first max)                             */
    FPC_ROW_WIN_COUNT = 44, /* i32 its inlier count                                                        */
    FPC_ROW_TN = 45,        /* i32 pixels that voted (after the max_num sub-sampling)                      */
    FPC_ROW_REFINE_INL = 46,/* i32 inliers of the refinement vote                                          */
    FPC_ROW_BBOX = 47,      /* i32 (ymin << 16) | xmin                                                     */
};

/* counters[] words the host reads back (one D2H copy of FPC_NUM_COUNTERS int32) */
enum {
    FPC_CNT_INSTANCES = 0, /* N found (may exceed max_instances -> FPC_FLAG_INSTANCES)  */
    FPC_CNT_ROWS = 1,
    FPC_CNT_RECORDS = 2,
    FPC_CNT_WORK = 3,
    FPC_CNT_FLAGS = 4,
    FPC_CNT_TICKET = 5,
    FPC_NUM_COUNTERS = 16,
};
enum { FPC_FLAG_INSTANCES = 1, FPC_FLAG_ROWS = 2, FPC_FLAG_RECORDS = 4 };

typedef struct fpc_recover_args {
    /* sizes */
    int32_t b, h, w, num_classes; /* num_classes includes background */
    int32_t hn;                   /* hypotheses per instance (HPARAM.HV_NUM_OF_HYPOTHESES) */
    int32_t max_instances;        /* capacity of every per-instance table                  */
    int64_t max_records;          /* capacity of the voting-record array (<= b*h*w)        */
    int64_t max_rows;             /* capacity of the (instance,row) table (<= b*h*w)       */
    /* voting parameters (ransac_voting_gpu.py:518-519) */
    float inlier_thresh;
    int32_t min_num, max_num;
    int32_t arith;                /* FPC_ARITH_*                                           */
    uint64_t seed;                /* device-side sampling when idxs / select_u are NULL    */
    /* inputs */
    const float *mask_logits;     /* [b,C,h,w]          */
    const float *quaternion;      /* [b,4(C-1),h,w]     */
    const float *scales;          /* [b,3(C-1),h,w]     */
    const float *xy;              /* [b,2(C-1),h,w]     */
    const float *z;               /* [b,C-1,h,w]        */
    const float *inv_intrinsics;  /* [3,3]              */
    const int32_t *idxs;          /* [max_instances,hn,2] fixed pre-sampled pixel pairs, or NULL */
    const float *select_u;        /* [b,h,w] uniforms for the max_num sub-sampling, or NULL      */
    /* outputs */
    float *pose_table;            /* [max_instances, FPC_POSE_ROW]                         */
    int32_t *counters;            /* [FPC_NUM_COUNTERS]                                    */
    uint8_t *cat_mask_u8;         /* [b,h,w] predicted class per pixel (may be NULL -> in workspace) */
    int32_t *labels;              /* [b,h,w] instance id + 1, 0 = background (may be NULL -> workspace) */
    float *hyp_out;               /* [max_instances,hn,2] or NULL                          */
    int32_t *vote_counts_out;     /* [max_instances,hn]  or NULL (then kept in workspace)  */
    /* scratch */
    void *workspace;
    size_t workspace_bytes;
    void *stream;
    /* optional per-kernel timing: cudaEvent_t handles, event 0 recorded before the first kernel and
     * event k after the k-th launch (k = 1..fpc_pose_recover_num_launches()); NULL entries are skipped */
    void **stage_events;
    int32_t num_stage_events;
    /* Head-epilogue fusion, fpc_pose_recover only.  0 or 1: the five head maps are full
     * resolution [b,.,h,w] (above).  S > 1: they are the LOW-RESOLUTION outputs of the heads' 1x1 convolutions,
     * [b,.,h/S,w/S] (h, w multiples of S), and the x S bilinear up-sampling of smp's SegmentationHead
     * (nn.UpsamplingBilinear2d(scale_factor=S), align_corners=True; lib/pose_regressor.py:633-666) is evaluated on the
     * fly inside the arg-max and gather kernels -- the [b,67,h,w] head maps are never written or read. */
    int32_t upsample;
    /* Optional [max_instances,4] f32: per instance (residual variance of the refinement inliers about the refined point,
     * PVNet v4 ransac_voting_gpu.py:757-759; fraction of the voters that vote for the refined point at threshold 0.999,
     * PVNet v5 :855-857; norm of the mean quaternion before its normalisation, needed by the backward of the aggregation;
     * 0).  One extra pass over the instance's records. */
    float *extra_out;
    /* Optional device timeline (fpc_pose_recover only): [fpc_pose_recover_num_launches() + 1] uint64 in device memory.  A
     * one-thread kernel writes %globaltimer (ns) into slot 0 before the first kernel and into slot k after the k-th: unlike
     * CUDA events these stamps can be captured in a CUDA graph (tools/timeline.py).  Adds one tiny launch per kernel. */
    unsigned long long *stage_stamps;
} fpc_recover_args;

/* sizeof(fpc_recover_args) as this library was compiled: lets a binding check its own struct definition. */
FPC_API size_t fpc_recover_args_size(void);

/* Workspace size for fpc_pose_recover with these sizes (only the size fields are read). */
FPC_API size_t fpc_pose_recover_workspace_bytes(const fpc_recover_args *args);

/* Runs the whole path.  Launches only; read `counters` (and the first
 * counters[FPC_CNT_INSTANCES] rows of pose_table) after synchronising the stream. */
FPC_API int fpc_pose_recover(const fpc_recover_args *args);

/* AggregationLayer.forward (lib/aggregation_layer.py:61-158) on already class-compressed
 * CategoricalData: cat_mask [b,h,w] i64, and in `args` quaternion [b,4,h,w], scales [b,3,h,w],
 * xy [b,2,h,w], z [b,h,w] (mask_logits / inv_intrinsics / idxs are ignored).  Fills class, sample,
 * count, quaternion, scales and z of every pose-table row and the label volume (args->labels or the
 * workspace); nothing votes.  Same workspace size as fpc_pose_recover. */
FPC_API int fpc_aggregate(const fpc_recover_args *args, const int64_t *cat_mask);

/* Labelling only: AggregationLayer.batchwise_break_segmentation_mask (lib/aggregation_layer.py:160-183, scipy / cupyx
 * ndimage.label with the no-cross-image 4-connected structure of :43-59).  cat_mask [b,h,w] i64 (non-zero = foreground)
 * -> args->labels [b,h,w] i32 (0 = background, k = k-th component in raster order of first pixels, image 0 first) and
 * counters[FPC_CNT_INSTANCES] = number of components.  Uses args->{b,h,w,max_instances,max_rows,workspace,stream}; the
 * head-map pointers are not read.  Same workspace size as fpc_pose_recover. */
FPC_API int fpc_label_instances(const fpc_recover_args *args, const int64_t *cat_mask);

/* ransac_voting_layer_v3 / ransac_voting_layer (lib/ransac_voting_gpu_layer/ransac_voting_gpu.py:518-607,
 * :11-98) on dense masks.  args->b is the number of voting problems (v3: instances; v1: images x
 * (class_num-1)), args->h/w the plane size.  Problem j votes with the pixels of
 *   fmask[j] != 0                                   (v3: mask [N,h,w] f32), or
 *   imask[j / nplanes_per_src] == match_base + j % nplanes_per_src   (v1: class-id mask [b,h,w] i32),
 * with directions vertex[(j / nplanes_per_src)*sN + y*sH + x*sW + {0, s2}] (element strides, so a
 * non-contiguous [.,h,w,vn,2] view of one keypoint works).  refine=1: inlier refinement (v3);
 * refine=0: the winning hypothesis (v1).  Results: FPC_ROW_XY / HYP / WIN_* / TN of each table row.
 * args->select_u, if given, is [nproblems,h,w].  Same workspace size as fpc_pose_recover. */
FPC_API int fpc_vote_dense(const fpc_recover_args *args, const float *fmask, const int32_t *imask, int nplanes_per_src,
                           int match_base, const float *vertex, long long sN, long long sH, long long sW, long long s2,
                           int refine);

/* Dense reference-layout outputs of AggregationLayer.forward (lib/aggregation_layer.py:101-105,152-153):
 * instance_masks [n,h,w] f32 0/1 and xy_mask [n,2,h,w] = mask * xy_cat[sample]; either may be NULL.
 * labels [b,h,w] = instance id + 1; pose_table rows supply each instance's frame. */
FPC_API int fpc_materialize_instances(const int32_t *labels, const float *pose_table, const float *xy_cat,
                                      float *instance_masks, float *xy_mask, int n, int h, int w, void *stream);

/* nn.UpsamplingBilinear2d(scale_factor=scale) (align_corners=True) of `planes` planes [hl,wl] -> [hl*scale, wl*scale]:
 * the up-sampling step of smp's SegmentationHead (lib/pose_regressor.py:633-666) as a stand-alone operator.  Same
 * arithmetic as ATen: src = dst * float(in-1)/float(out-1); value = fma(fma(v00,wx0,v01*wx1), wy0, fma(v10,wx0,v11*wx1)*wy1)
 * -- bit-identical to torch's CPU kernel (planes of >= 3200 output pixels) and to its CUDA kernel. */
FPC_API int fpc_upsample_bilinear(const float *in, long long planes, int hl, int wl, int scale, float *out, void *stream);

/* ---- ground-truth <-> prediction matching ---------------------------------------
 * Replaces lib/gpu_tensor_funcs.py:386-409 batchwise_get_2d_iou (expand both mask sets to [n1,n2,h,w], sum
 * logical_and / logical_or) and the per-class loop of lib/matching.py:253-296 batchwise_find_matches.
 *
 * A mask set is held as bit planes: bits [n, h, ceil(w/32)] u32 (bit x%32 of word x/32 = pixel x set, padding bits 0)
 * plus meta [n, FPC_MASK_META] i32 = {pixel count, ymin, ymax, first word column, last word column, 0, 0, 0}
 * (empty mask: count 0, ymin > ymax).  Caller-allocated, like everything else. */
enum { FPC_MASK_META = 8 };
enum { FPC_MASK_F32 = 0, FPC_MASK_U8 = 1 };

/* Dense masks [n,h,w] (elem = FPC_MASK_F32 or FPC_MASK_U8/bool; non-zero = set, the reference's logical_and rule)
 * -> bit planes + meta.  Reads every mask element exactly once. */
FPC_API int fpc_pack_masks(const void *masks, int elem, int n, int h, int w, uint32_t *bits, int32_t *meta, void *stream);

/* Label volume [b,h,w] i32 (0 = background, k = instance k-1; what fpc_pose_recover writes to args->labels)
 * -> bit planes + meta of instances 0..n-1, without ever building dense instance masks. */
FPC_API int fpc_pack_labels(const int32_t *labels, int b, int h, int w, int n, uint32_t *bits, int32_t *meta, void *stream);

/* iou [na,nb] f32 = |A_i and B_j| / |A_i or B_j| as float(int) / float(int), 0/0 = NaN: bit-identical to the
 * reference's int64 true division (gpu_tensor_funcs.py:396-407). */
FPC_API int fpc_mask_iou(const uint32_t *bits_a, const int32_t *meta_a, int na, const uint32_t *bits_b, const int32_t *meta_b,
                         int nb, int h, int w, float *iou, void *stream);

/* The pairing of lib/matching.py:253-296 for all classes in one call.  For every ground-truth mask i:
 * best_pred[i] = the FIRST prediction of the same class (any frame) with the largest IoU, or -1 if that IoU is not > 0;
 * best_iou[i] = that IoU (0 if none).  pairs [ng,2] i32 receives (gt index, pred index) of the *n_matches matched
 * ground truths in the reference's output order: class id ascending, then gt index ascending.  class ids are int64
 * (torch long), as the reference stores them. */
FPC_API int fpc_match_instances(const uint32_t *bits_g, const int32_t *meta_g, const int64_t *class_g, int ng,
                                const uint32_t *bits_p, const int32_t *meta_p, const int64_t *class_p, int np, int h, int w,
                                int32_t *best_pred, float *best_iou, int32_t *pairs, int32_t *n_matches, void *stream);

/* out [m,h,w] f32 0/1: row k = the mask of instance inst_of[k] in frame frame_of[k] of the label volume [b,h,w]
 * (labels == inst_of[k] + 1).  The stacked `instance_masks` of matched predictions (lib/matching.py:40-58) without ever
 * holding all N dense masks. */
FPC_API int fpc_paint_instances(const int32_t *labels, int b, int h, int w, const int64_t *frame_of, const int64_t *inst_of, int m,
                                float *out, void *stream);

/* ---- evaluation maths on m matched (ground truth, prediction) pairs ---------------
 * One launch for lib/gpu_tensor_funcs.py:434-455 get_raw_quat_distance (raw_degrees [m] f32), :457-476 + :752-799
 * get_symmetric_quat_distance (sym_degrees [m] f64: min over the 360 rotations sym_rotations [360,4] f64 about y; NaN for
 * pairs whose symmetric_ids entry is 0; symmetric_ids NULL = every pair), :503-547 get_3d_ious (iou_3d [m] f32, with the
 * reference's reduction over the coordinate axis kept) and :565-567 from_Ts_get_offset_error (offset_error [m] f32).
 * world_centers [m,6] f32 = inverse(RT) @ (0,0,0,1) de-homogenised, for the ground truth then the prediction: the two
 * points from_RTs_get_T_offset_errors (:569-609) compares.
 * Any output may be NULL (then its inputs may be NULL too).  q [m,4], RT [m,4,4], scales [m,3], T [m,3], row-major. */
FPC_API int fpc_pose_errors(const float *q_gt, const float *q_pred, const int64_t *symmetric_ids, const double *sym_rotations,
                            const float *rt_gt, const float *rt_pred, const float *scales_gt, const float *scales_pred,
                            const float *t_gt, const float *t_pred, int m, float *raw_degrees, double *sym_degrees, float *iou_3d,
                            float *offset_error, float *world_centers, void *stream);

/* lib/gpu_tensor_funcs.py:611-652 calculate_aps for one (metric, class): out[t] = fraction of the non-NaN values with
 * value < thresholds[t] (op 0, torch.less) or value > thresholds[t] (op 1, torch.greater). */
FPC_API int fpc_threshold_fraction(const double *values, int n, const double *thresholds, int num_thresholds, int op, float *out,
                                   void *stream);

/* ---- training support: backward of the two differentiable steps at the head of the path --------------------------
 * (the reference trains its aggregated losses, lib/loss.py:155-545, through these torch ops)
 *
 * class_compress (lib/gpu_tensor_funcs.py:52-99).  g_* = gradients of the class-compressed fields ([b,4|3|2,h,w], z
 * [b,h,w]; any may be NULL); d_* = gradients of the raw head maps ([b,4K|3K|2K|K,h,w], zero-filled here, then the
 * predicted class's channels of foreground pixels are written; q / xy go through the Jacobian of their L2 normalisation,
 * for which the raw `quaternion` / `xy` head maps are read). */
FPC_API int fpc_class_compress_backward(const int64_t *cat_mask, const float *quaternion, const float *xy, const float *g_q,
                                        const float *g_s, const float *g_xy, const float *g_z, float *d_quaternion,
                                        float *d_scales, float *d_xy, float *d_z, int b, int num_classes, int h, int w,
                                        void *stream);

/* AggregationLayer.forward (lib/aggregation_layer.py:125-156).  labels [b,h,w] i32 from the forward call; inst_grads [n,8] =
 * per-instance gradient of (q0..q3, s0..s2, z) w.r.t. ONE member pixel (the caller folds in 1/count, exp and the
 * normalisation Jacobian: n small vectors); g_xy_dense [n,2,h,w] = gradient of the masked xy output or NULL.  Writes every
 * element of d_q [b,4,h,w], d_s [b,3,h,w], d_xy [b,2,h,w], d_z [b,h,w] (zeros on background). */
FPC_API int fpc_aggregate_backward(const int32_t *labels, const float *inst_grads, const float *g_xy_dense, int n, float *d_q,
                                   float *d_s, float *d_xy, float *d_z, int b, int h, int w, void *stream);

/* Backward of fpc_get_rt (lib/gpu_tensor_funcs.py:204-235, 306-326): gradients g_R [n,3,3], g_T [n,3], g_RT [n,4,4] (any may
 * be NULL) -> d_q [n,4], d_xy [n,2], d_z [n]. */
FPC_API int fpc_get_rt_backward(const float *q, const float *xy, const float *z, const float *inv_k, const float *g_R,
                                const float *g_T, const float *g_RT, int n, float *d_q, float *d_xy, float *d_z, void *stream);

/* Backward of the refinement solve of ransac_voting_layer_v3 (ransac_voting_gpu.py:584-598) w.r.t. the direction field.
 * fmask [n,h,w] f32 and vertex (element strides sN, sH, sW, s2 as in fpc_vote_dense) are the forward inputs; win_pts [n,2]
 * the winning hypotheses (they define the inlier set, a constant of the differentiation as in the reference), refined [n,2]
 * the forward result, g_x [n,2] its gradient, live [n] i32 = 0 for instances the forward skipped (< min_num pixels).
 * d_vertex [n,h,w,2] contiguous: every element written (zeros off the inlier set). */
FPC_API int fpc_vote_refine_backward(const float *fmask, const float *vertex, long long sN, long long sH, long long sW,
                                     long long s2, const float *win_pts, const float *refined, const float *g_x,
                                     const int32_t *live, float inlier_thresh, int n, int h, int w, int arith, float *d_vertex,
                                     void *stream);

/* The same for the fused path (fpc_pose_recover): membership from the label volume [b,h,w], direction = the predicted
 * class's raw xy channels of xy_head [b,2K,h,w] normalised as the forward does; frame_of [n] = frame of each instance.
 * d_xy_head [b,2K,h,w] is zero-filled here, then the inlier pixels receive the gradient pushed through that normalisation. */
FPC_API int fpc_pose_recover_xy_backward(const int32_t *labels, const uint8_t *cat_mask_u8, const float *xy_head,
                                         const int32_t *frame_of, const float *win_pts, const float *refined, const float *g_x,
                                         const int32_t *live, float inlier_thresh, int n, int b, int num_classes, int h, int w,
                                         int arith, float *d_xy_head, void *stream);

/* Number of kernels fpc_pose_recover launches per call (for launch accounting). */
FPC_API int fpc_pose_recover_num_launches(void);

/* Name of the k-th kernel (0-based) fpc_pose_recover launches, for profiles and bench output. */
FPC_API const char *fpc_pose_recover_kernel_name(int k);

/* Measurement helper (not on the path): a pure FFMA loop, `blocks` x 256 threads, 16 independent
 * accumulators x `iters` iterations per thread = blocks*256*iters*16 FMAs (x2 flop).  bench.py
 * times it with CUDA events to get this chip's FP32 peak, the voting kernel's roofline denominator
 * (MEASURED_PEAKS.json carries none). */
FPC_API int fpc_bench_fp32_fma(float *sink, int blocks, int iters, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* FPC_B200_H_ */
