// ORACLE -- TEST INFRASTRUCTURE ONLY.
// Minimal pybind11 binding (ours) around the reference's OWN, unmodified CUDA source
// lib/ransac_voting_gpu_layer/src/ransac_voting_kernel.cu, which is compiled from where it lies under
// /root/reference by oracle/build_ref_cuda.py into oracle/_ref/ (git-ignored).  The reference's own binding
// (src/ransac_voting.cpp) cannot be used: its line 5 `extern THCState* state;` does not compile against
// torch >= 1.11.  Exposes the same four functions the reference module does (ransac_voting.cpp:102-107).
#include <torch/extension.h>

at::Tensor generate_hypothesis_launcher(at::Tensor direct, at::Tensor coords, at::Tensor idxs);
void voting_for_hypothesis_launcher(at::Tensor direct, at::Tensor coords, at::Tensor hypo_pts, at::Tensor inliers,
                                    float inlier_thresh);

at::Tensor generate_hypothesis_vanishing_point_launcher(at::Tensor direct, at::Tensor coords, at::Tensor idxs);
void voting_for_hypothesis_vanishing_point_launcher(at::Tensor direct, at::Tensor coords, at::Tensor hypo_pts, at::Tensor inliers,
                                                    float inlier_thresh);

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.def("generate_hypothesis", &generate_hypothesis_launcher, "reference K1 (ransac_voting_kernel.cu:51-86)");
    m.def("voting_for_hypothesis", &voting_for_hypothesis_launcher, "reference K2 (ransac_voting_kernel.cu:129-167)");
    m.def("generate_hypothesis_vanishing_point", &generate_hypothesis_vanishing_point_launcher, "reference K3 (.cu:230-266)");
    m.def("voting_for_hypothesis_vanishing_point", &voting_for_hypothesis_vanishing_point_launcher, "reference K4 (.cu:311-350)");
}
