// Internal structures shared by the translation units of libfpc_b200.so.
#pragma once

#include "fpc_common.cuh"

namespace fpc {

// Per-instance tables (capacity = max_instances [+1 for the offset arrays]).
struct InstTables {
    int *root;     // linear index (over the [b,h,w] volume) of the component's first pixel
    int *count;    // pixels in the instance mask
    int *ymin, *ymax, *xmin, *xmax;
    int *mincls;   // min non-zero class id inside the component (aggregation_layer.py:113)
    int *nruns;    // runs (horizontal segments) of the instance
    int *rmax2;    // bit pattern of max |pixel - centre of the bounding box|^2 over the instance (k_gather; the vote kernel's |d| bound)
    int *rowoff;   // [N+1] first (instance,row) item of the instance
    int *tn;       // pixels that vote (0 if count < min_num; ~max_num if sub-sampled)
    int *pxoff;    // [N+1] first voting record of the instance
    int *workoff;  // [N+1] first vote work item (chunk of pixels) of the instance
};

// Run tables (capacity max_rows): maximal horizontal foreground segments in raster order.
struct RunTables {
    int *start, *end;   // first / last pixel (linear index over the [b,h,w] volume)
    int *parent;        // union-find over runs; after flattening: the component's first run
    int *inst;          // instance id
    int *rowrun;        // [b*h+1] first run at or after the start of every image row
};

constexpr int ROW_CONTIG = 1 << 30;   // the slot's members are one contiguous run: no label test needed
constexpr int ROW_SUB = 1 << 29;      // instance larger than max_num: Bernoulli sub-sampling of the voters
constexpr int ROW_VOTES = 1 << 28;    // instance has at least min_num pixels
constexpr int ROW_LEN_MASK = (1 << 20) - 1;  // run length (<= image width < 65536)
constexpr int ROW_CLS_SHIFT = 20;            // bits 20..27: class of the run's FIRST pixel (k_gather_p prefetches its channels)

// Per-(instance,run) slot tables, instance-major, raster order inside an instance.
struct RowTables {
    int4 *desc;    // (instance, first member pixel, length | ROW_* flags, exclusive prefix of voting pixels in the instance)
    float *sum;    // [rows,8] partial sums: q0..q3, s0..s2, z
};

// Where the per-pixel regression fields come from.
struct FieldSrc {
    const float *quaternion, *scales, *xy, *z;
    // gather mode 2: `xy` is a strided vertex view; element strides of (problem, row, column, component)
    long long sN, sH, sW, s2;
    int div;   // problems per source plane (v1: classes per image); vertex plane = problem / div
};

struct PathParams {
    int b, h, w, hw, P;          // P = b*h*w  (< 2^31)
    int num_classes, hn;
    int max_instances;
    long long max_records, max_rows;
    float inlier_thresh;
    int min_num, max_num;
    int arith;
    unsigned long long seed;
    const int *idxs;             // [max_instances,hn,2] or nullptr
    const float *select_u;       // [b,h,w] or nullptr
    int refine;                  // 1: inlier refinement (v3); 0: winning hypothesis as is (v1)
    int vote_chunk;              // pixels per vote work item for this problem size (vote_chunk_for)
    int vote_tail;               // divisor of the item size at the tail of the queue (vote_item_px)
    UpParams up;                 // head-epilogue fusion: head maps are low resolution, x up.s bilinear on the fly
    float *extra;                // [max_instances,2] (v4 residual variance, v5 confidence at 0.999) or nullptr
};

// Voting records as four SoA planes, instance-major, raster order inside an instance; every instance's range
// starts at a multiple of 4 records so that chunks can be bulk-copied (16-byte aligned) into shared memory.
struct RecPlanes {
    float *x, *y, *nx, *ny;
};

struct Workspace {
    uint8_t *cls;      // [P]
    int *tile_roots;   // [ntiles+1] run starts per 1024-pixel tile (scanned in place)
    int *run_tiles;    // [max_rows/1024+2] component roots per 1024-run tile (scanned in place)
    RunTables RT;
    int *counters;     // [FPC_NUM_COUNTERS]
    InstTables T;
    RowTables R;
    RecPlanes rec;     // 4 x [max_records]
    int4 *work;        // vote work descriptors
    float4 *workf;     // ... and the instance-local frame of each (origin, extents)
    uint4 *segs;       // [vote blocks, vote_seg_cap()] flagged rounds handed from k_vote to k_vote_settle: (work item, note)
    int *segcnt;       // [vote blocks]
    float2 *hyp;       // [max_instances, hn]
    float4 *hloc;      // [max_instances, hn] hypotheses prepared for the vote kernel: (h'x, h'y, band_delta, -)
    int *votes;        // [max_instances, hn]
};

#ifndef FPC_VOTE_CHUNK
#define FPC_VOTE_CHUNK 1024
#endif
constexpr int VOTE_CHUNK = FPC_VOTE_CHUNK;  // pixels per vote work item (upper bound: shared-memory buffers have this size)
// Small problems (a single frame, a few instances) would give the persistent vote kernel only a handful of work items;
// they are cut into smaller chunks so that the votes still spread over the machine (fpc_voting.cu).
int vote_chunk_for(long long P, int hn);
// Pixels per work item of instance i of N.  The vote kernel's blocks pull items in instance order; with only ~5 items per
// block a block that draws one item more than its neighbours finishes up to 20 % later, so the items at the END of the
// queue (the last fifth of the instances) are a quarter of the size: the tail shrinks to a quarter item.
__host__ __device__ inline int vote_item_px(int i, int N, int chunk, int tail_div) {
    return (chunk >= 128 * tail_div && i >= N - N / 5) ? chunk / tail_div : chunk;
}
int vote_tail_div();   // 1 by default (measured: smaller tail items cost more instructions than the tail they save), FPC_VOTE_TAIL_DIV in the environment (1 = all items the same size)

// fpc_aggregate.cu
int launch_label_and_tables(const Workspace &ws, const PathParams &pp, const float *mask_logits,
                            const long long *cat_mask_i64, cudaStream_t st);
int launch_rows_and_records(const Workspace &ws, const PathParams &pp, const FieldSrc &F, int gather_mode,
                            bool want_records, int vote_chunk, cudaStream_t st);
int launch_slots(const Workspace &ws, const PathParams &pp, cudaStream_t st);
int launch_relabel(const Workspace &ws, const PathParams &pp, int *labels_out, cudaStream_t st);
int launch_dense_problems(const Workspace &ws, const PathParams &pp, const float *fmask, const int *imask,
                          int nplanes_per_src, int match_base, int nprob, cudaStream_t st);
int launch_materialize(const int *label, const float *table, const float *xy_cat, float *masks, float *xy_mask, int n,
                       int hw, cudaStream_t st);

// fpc_voting.cu
int launch_generate_hypothesis(const float *direct, const float *coords, const int *idxs, float *hypo, int tn, int vn,
                               int hn, int arith, cudaStream_t st);
int launch_voting_for_hypothesis(const float *direct, const float *coords, const float *hypo, uint8_t *inliers, int tn,
                                 int vn, int hn, float thresh, int arith, cudaStream_t st);
int launch_generate_hypothesis_vp(const float *direct, const float *coords, const int *idxs, float *hypo, int tn, int vn, int hn,
                                  int arith, cudaStream_t st);
int launch_voting_for_hypothesis_vp(const float *direct, const float *coords, const float *hypo, uint8_t *inliers, int tn, int vn,
                                    int hn, float thresh, int arith, cudaStream_t st);
int launch_vote_refine_backward(const float *fmask, const float *vertex, long long sN, long long sH, long long sW, long long s2,
                                const float *win_pts, const float *refined, const float *g_x, const int *live, float thresh, int n,
                                int h, int w, int arith, float *d_vertex, cudaStream_t st);
int launch_vote_refine_backward_labels(const int *labels, const uint8_t *cls, const float *xy_head, const int *frame_of,
                                       const float *win_pts, const float *refined, const float *g_x, const int *live, float thresh,
                                       int n, int K, int h, int w, int arith, float *d_xy_head, cudaStream_t st);
void set_vote_packed(int v);
int vote_batches(int hn);
int vote_seg_cap();
int launch_vote(const Workspace &ws, const PathParams &pp, float2 *hyp_out, int *votes, cudaStream_t st);
int launch_finalize(const Workspace &ws, const PathParams &pp, const float2 *hyp, const int *votes, const float *inv_k,
                    float *pose_table, cudaStream_t st);
int launch_get_rt(const float *q, const float *xy, const float *z, const float *inv_k, float *R, float *T, float *RT,
                  int n, cudaStream_t st);

// uniform in [0,1) deciding whether pixel p of a > max_num instance votes (ransac_voting_gpu.py:542-545)
__device__ __forceinline__ float select_uniform(const PathParams &pp, int p) {
    if (pp.select_u) return pp.select_u[p];
    return (float)(hash3(pp.seed, (uint32_t)p, 0x5e1ec7u, 0u) >> 8) * (1.0f / 16777216.0f);
}

// Largest i in [0,n) with a[i] <= v  (a ascending, a[0] == 0 <= v).
__device__ __forceinline__ int upper_index(const int *__restrict__ a, int n, int v) {
    int lo = 0, hi = n;  // invariant: a[lo] <= v, (hi == n or a[hi] > v)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (a[mid] <= v) lo = mid; else hi = mid;
    }
    return lo;
}

// gpu_tensor_funcs.batchwise_get_RT (lib/gpu_tensor_funcs.py:204-235) for one instance, with
// quats_2_rotation_matrix (:306-326, including its final transpose).  The reference builds
// inv_RT = [[R^-1, T],[0,0,0,1]] and inverts it; in closed form that is RT = [[R, -R T],[0,0,0,1]].
__device__ inline void pose_from_qxyz(const float *q, float x, float y, float z, const float *__restrict__ inv_k,
                                      float *R, float *T, float *RT) {
    const float zz = z / 1000.f;
    const float hx = x * zz, hy = y * zz;
    float t[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) t[r] = inv_k[3 * r] * hx + inv_k[3 * r + 1] * hy + inv_k[3 * r + 2] * zz;
    const float n = sqrtf(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    const float sn = n > 0.f ? n : 1.f;
    const float a = q[0] / sn, b = q[1] / sn, c = q[2] / sn, d = q[3] / sn;
    const float aa = a * a, bb = b * b, cc = c * c, dd = d * d;
    float m[3][3];
    m[0][0] = aa - bb - cc + dd;  m[0][1] = 2.f * (a * b + c * d); m[0][2] = 2.f * (a * c - b * d);
    m[1][0] = 2.f * (a * b - c * d); m[1][1] = -aa + bb - cc + dd; m[1][2] = 2.f * (b * c + a * d);
    m[2][0] = 2.f * (a * c + b * d); m[2][1] = 2.f * (b * c - a * d); m[2][2] = -aa - bb + cc + dd;
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        float rt = 0.f;
#pragma unroll
        for (int cidx = 0; cidx < 3; ++cidx) {
            const float v = m[cidx][r];   // transpose
            R[3 * r + cidx] = v;
            RT[4 * r + cidx] = v;
            rt += v * t[cidx];
        }
        RT[4 * r + 3] = -rt;
        T[r] = t[r];
    }
    RT[12] = 0.f; RT[13] = 0.f; RT[14] = 0.f; RT[15] = 1.f;
}

}  // namespace fpc
