"""Dict contracts of the path (same keys as the reference's lib/type_hinting.py:5-32)."""
import typing

import torch


class LogitData(typing.TypedDict, total=False):
    mask: torch.Tensor          # [b,C,h,w]
    quaternion: torch.Tensor    # [b,4(C-1),h,w]
    scales: torch.Tensor        # [b,3(C-1),h,w]
    z: torch.Tensor             # [b,C-1,h,w]
    xy: torch.Tensor            # [b,2(C-1),h,w]


class CategoricalData(typing.TypedDict, total=False):
    mask: torch.Tensor          # [b,h,w] int64
    quaternion: torch.Tensor    # [b,4,h,w]
    scales: torch.Tensor        # [b,3,h,w]
    z: torch.Tensor             # [b,h,w]
    xy: torch.Tensor            # [b,2,h,w]


class AggData(typing.TypedDict, total=False):
    class_ids: torch.Tensor     # [N] int64
    sample_ids: torch.Tensor    # [N] int64
    instance_masks: torch.Tensor  # [N,h,w] float32 0/1
    quaternion: torch.Tensor    # [N,4]
    scales: torch.Tensor        # [N,3]
    z: torch.Tensor             # [N,1]
    xy: torch.Tensor            # [N,2,h,w] after aggregation, [N,2] (col,row) after voting
    R: torch.Tensor             # [N,3,3]
    T: torch.Tensor             # [N,3]
    RT: torch.Tensor            # [N,4,4]
