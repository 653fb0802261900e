"""fastposecnn_b200 -- B200-native (sm_100a) implementation of FastPoseCNN's post-network pose-recovery
path: per-pixel head outputs -> per-instance 6D pose and size.

Drop-in names (same call signatures and tensor layouts as the reference):
    AggregationLayer, HoughVotingLayer, ransac_voting_layer_v3, ransac_voting_layer,
    class_compress, class_compression, normalize, samplewise_get_RT, batchwise_get_RT, Model,
    matching.batchwise_find_matches / batchwise_find_matches2, batchwise_get_2d_iou, torch_get_2d_iou
Fused fast entry:
    pose_recover(logits, inv_intrinsics, hn, idxs=None) / PoseRecoveryEngine

Every operator calls hand-written CUDA kernels in ``libfpc_b200.so`` through a C ABI
(``include/fpc_b200.h``).  There is no CPU path and no PyTorch fallback: a missing library or a
non-CUDA tensor raises.
"""
from . import synthetic  # noqa: F401  (pure torch, importable without the native library)
from .aggregation_layer import AggregationLayer  # noqa: F401
from . import head_epilogue, matching  # noqa: F401
from .gpu_tensor_funcs import (batchwise_get_2d_iou, batchwise_get_RT, calculate_aps, class_compress,  # noqa: F401
                               class_compression, from_Ts_get_offset_error, get_3d_iou, get_3d_ious, get_quat_distance,
                               get_raw_quat_distance, get_symmetric_quat_distance, normalize, quats_2_rotation_matrix,
                               samplewise_get_RT, torch_get_2d_iou)
from .matching import batchwise_find_matches, batchwise_find_matches2  # noqa: F401
from .head_epilogue import lowres_logits, split_xyz, upsample_bilinear  # noqa: F401
from .hough_voting import HoughVotingLayer  # noqa: F401
from .model import Model, PoseRecovery  # noqa: F401
from .pose_recovery import PoseRecoveryEngine, pose_recover  # noqa: F401
from .ransac_voting_gpu_layer.ransac_voting_gpu import b_inv, ransac_voting_layer, ransac_voting_layer_v3  # noqa: F401

__version__ = "0.1.0"
