"""Head-epilogue fusion: the step right BEFORE the pose-recovery path.

In the reference every head is smp's ``SegmentationHead`` = ``Conv2d(in, out, kernel_size=1)`` ->
``nn.UpsamplingBilinear2d(scale_factor=4)`` -> identity (lib/pose_regressor.py:633-666), and ``pure_model_forward``
(:706-741) hands the path 67 full-resolution channels: 82 MB per 640x480 frame written by the up-sampling kernels and
read back by the path.  Bilinear up-sampling is a fixed 4-tap stencil, so the path can evaluate it where it consumes
the value: the arg-max kernel interpolates the 7 mask logits of every pixel, the gather kernel the predicted class's
10 channels of foreground pixels only, both from the low-resolution conv outputs (1/16 of the bytes, L2 resident).

* ``lowres_logits(heads, decoder_outputs)``  -- runs only the 1x1 convolutions of the reference's own head modules
  (cuDNN through torch: a plain library GEMM) and splits xyz -> xy, z like :729-732, at low resolution;
* ``pose_recover(lowres, inv_K, hn, upsample=4)`` (``pose_recovery.py``) -- the fused path on those;
* ``upsample_bilinear(x, scale)`` -- the up-sampling as a stand-alone operator with the same arithmetic (bit-identical
  to ``nn.UpsamplingBilinear2d`` on CUDA and to ATen's vectorised CPU kernel), for callers that still want the maps.
"""
from __future__ import annotations

from typing import Dict, Mapping

import torch

from . import _lib


def upsample_bilinear(x: torch.Tensor, scale: int) -> torch.Tensor:
    """``nn.UpsamplingBilinear2d(scale_factor=scale)(x)`` (align_corners=True) for ``[..., hl, wl]`` float32."""
    x = _lib.require_cuda(x, "x", torch.float32)
    if x.dim() < 2 or int(scale) != scale or scale < 1:
        raise RuntimeError("upsample_bilinear: expected [..., h, w] and an integer scale >= 1")
    hl, wl = x.shape[-2:]
    out = torch.empty(tuple(x.shape[:-2]) + (hl * scale, wl * scale), dtype=torch.float32, device=x.device)
    planes = x.numel() // max(hl * wl, 1)
    with torch.cuda.device(x.device):
        _lib.check(_lib.lib().fpc_upsample_bilinear(x.data_ptr(), planes, hl, wl, int(scale), out.data_ptr(),
                                                    _lib.current_stream(x.device)))
    return out


def split_xyz(xyz_logits: torch.Tensor) -> Dict[str, torch.Tensor]:
    """lib/pose_regressor.py:729-732: channels (3k, 3k+1) of class k are its xy direction, channel 3k+2 its z."""
    b, c, h, w = xyz_logits.shape
    if c % 3:
        raise RuntimeError("split_xyz: the translation head has 3 channels per class")
    v = xyz_logits.reshape(b, c // 3, 3, h, w)          # no index tensors: capturable in a CUDA graph
    return {"xy": v[:, :, :2].reshape(b, 2 * (c // 3), h, w).contiguous(), "z": v[:, :, 2].contiguous()}


def _conv_of(head) -> torch.nn.Module:
    """The 1x1 convolution of a SegmentationHead-like ``nn.Sequential(conv, upsampling, activation)``."""
    if isinstance(head, torch.nn.Conv2d):
        return head
    conv = head[0]
    if not isinstance(conv, torch.nn.Conv2d):
        raise TypeError("expected a Conv2d or a Sequential whose first module is the head's Conv2d")
    return conv


def lowres_logits(heads: Mapping[str, torch.nn.Module], decoder_outputs: Mapping[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """Low-resolution LogitData from the reference's own head modules.  ``heads`` / ``decoder_outputs`` are keyed
    ``mask, rotation, translation, scales`` (``segmentation_head, rotation_head, translation_head, scales_head`` and the
    four decoder outputs of lib/pose_regressor.py:713-722).  Only the convolutions run; feed the result to
    ``pose_recover(..., upsample=S)``."""
    out = {"mask": _conv_of(heads["mask"])(decoder_outputs["mask"]).contiguous(),
           "quaternion": _conv_of(heads["rotation"])(decoder_outputs["rotation"]).contiguous(),
           "scales": _conv_of(heads["scales"])(decoder_outputs["scales"]).contiguous()}
    out.update(split_xyz(_conv_of(heads["translation"])(decoder_outputs["translation"])))
    return out
