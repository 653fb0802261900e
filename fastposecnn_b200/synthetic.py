"""Synthetic NOCS-shaped head outputs.

There is no dataset and no checkpoint on the build or GPU boxes, so every test
and benchmark feeds the path with head maps rendered here: non-touching discs on
a grid, class = grid index mod (C-1) + 1, centre-direction unit vectors with a
little noise inside each disc.  Layouts follow BASELINE.json ``configs``.

Layout of the heads is the one ``PoseRegressor.pure_model_forward`` produces
(lib/pose_regressor.py:709-743): NCHW float32, class-major channel groups
(``torch.chunk`` order, lib/gpu_tensor_funcs.py:68).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Sequence, Tuple

import torch

Disc = Tuple[float, float, float, int]          # (centre x [col], centre y [row], radius, class id 1..C-1)

# NOCS CAMERA intrinsics (tools/project.py:78)
CAMERA_INTRINSICS = [[577.5, 0.0, 319.5], [0.0, 577.5, 239.5], [0.0, 0.0, 1.0]]


def camera_intrinsics(device="cpu") -> torch.Tensor:
    return torch.tensor(CAMERA_INTRINSICS, dtype=torch.float32, device=device)


@dataclass(frozen=True)
class Workload:
    name: str
    batch: int
    h: int
    w: int
    grid_cols: int
    grid_rows: int
    radius: int
    hyps: int
    num_classes: int = 7

    def discs(self) -> List[Disc]:
        """One frame's discs, raster order of the grid (row-major)."""
        sx, sy = self.w / self.grid_cols, self.h / self.grid_rows
        out = []
        for r in range(self.grid_rows):
            for c in range(self.grid_cols):
                k = r * self.grid_cols + c
                out.append((float(int(sx / 2 + sx * c)), float(int(sy / 2 + sy * r)), float(self.radius),
                            k % (self.num_classes - 1) + 1))
        return out

    @property
    def instances_per_frame(self) -> int:
        return self.grid_cols * self.grid_rows


# BASELINE.json configs[0..3]
WORKLOADS: Dict[str, Workload] = {
    "cfg1": Workload("cfg1_b1_640x480_6inst_hn128", 1, 480, 640, 3, 2, 30, 128),
    "cfg2": Workload("cfg2_b32_640x480_18inst_hn128", 32, 480, 640, 6, 3, 35, 128),
    "cfg3": Workload("cfg3_b256_640x480_18inst_hn512", 256, 480, 640, 6, 3, 35, 512),
    "cfg4": Workload("cfg4_1280x960_20inst_hn1024", 1, 960, 1280, 5, 4, 97, 1024),
}


def render_heads(frames: Sequence[Sequence[Disc]], h: int, w: int, num_classes: int = 7, seed: int = 0,
                 device="cpu", quantize_mask: bool = True) -> Dict[str, torch.Tensor]:
    """Renders ``len(frames)`` frames.  Returns the LogitData dict
    (lib/type_hinting.py:5-10): mask [b,C,h,w], quaternion [b,4(C-1),h,w],
    scales [b,3(C-1),h,w], xy [b,2(C-1),h,w], z [b,C-1,h,w]."""
    dev = torch.device(device)
    g = torch.Generator(device=dev).manual_seed(seed)
    b, k = len(frames), num_classes - 1

    def randn(*shape, std=1.0):
        return torch.randn(*shape, generator=g, device=dev, dtype=torch.float32) * std

    mask = randn(b, num_classes, h, w, std=0.1)
    if quantize_mask:
        # multiples of 2^-10: distinct logits differ by >= ~1e-3, so argmax(log_softmax(x)) == argmax(x)
        mask = torch.round(mask * 1024.0) / 1024.0
    mask[:, 0] += 1.0
    quat = randn(b, 4 * k, h, w)
    scales = torch.rand(b, 3 * k, h, w, generator=g, device=dev, dtype=torch.float32)
    xy = randn(b, 2 * k, h, w, std=0.02)
    z = 6.9 + randn(b, k, h, w, std=0.01)

    ys = torch.arange(h, device=dev, dtype=torch.float32).view(h, 1)
    xs = torch.arange(w, device=dev, dtype=torch.float32).view(1, w)
    for bi, discs in enumerate(frames):
        for (cx, cy, r, cls) in discs:
            y0, y1 = max(int(cy - r) - 1, 0), min(int(cy + r) + 2, h)
            x0, x1 = max(int(cx - r) - 1, 0), min(int(cx + r) + 2, w)
            if y0 >= y1 or x0 >= x1:
                continue
            dx = cx - xs[:, x0:x1]
            dy = cy - ys[y0:y1, :]
            d2 = dx * dx + dy * dy
            inside = d2 <= r * r
            nrm = torch.sqrt(d2).clamp_min(1e-12)
            mask[bi, cls, y0:y1, x0:x1] += 5.0 * inside
            c0 = 2 * (cls - 1)
            xy[bi, c0, y0:y1, x0:x1] += (dx / nrm) * inside
            xy[bi, c0 + 1, y0:y1, x0:x1] += (dy / nrm) * inside
    return {"mask": mask, "quaternion": quat, "scales": scales, "xy": xy, "z": z}


def render_workload(wl: Workload, batch: int = None, seed: int = 0, device="cpu") -> Dict[str, torch.Tensor]:
    b = wl.batch if batch is None else batch
    return render_heads([wl.discs()] * b, wl.h, wl.w, wl.num_classes, seed=seed, device=device)


def disc_pixel_count(cx: float, cy: float, r: float, h: int, w: int) -> int:
    ys = torch.arange(h, dtype=torch.float32).view(h, 1)
    xs = torch.arange(w, dtype=torch.float32).view(1, w)
    return int((((xs - cx) ** 2 + (ys - cy) ** 2) <= r * r).sum())


def presampled_idxs(tns: Sequence[int], hn: int, vn: int = 1, seed: int = 1234, min_num: int = 5) -> torch.Tensor:
    """Fixed pre-sampled hypothesis pixel pairs, ``[N,hn,vn,2]`` int32 on the CPU.

    One generator, drawn in instance order and only for instances with
    ``tn >= min_num`` -- the same stream ``oracle.port.seeded_idx_source`` hands to
    the reference's ``random_(0, tn)`` call site (ransac_voting_gpu.py:552).
    Rows of skipped instances are zero."""
    g = torch.Generator().manual_seed(seed)
    out = torch.zeros((len(tns), hn, vn, 2), dtype=torch.int32)
    for i, tn in enumerate(tns):
        if tn >= min_num:
            out[i] = torch.randint(0, int(tn), (hn, vn, 2), generator=g, dtype=torch.int32)
    return out


def render_lowres_heads(frames: Sequence[Sequence[Disc]], h: int, w: int, scale: int = 4, num_classes: int = 7, seed: int = 0,
                        device="cpu") -> Dict[str, torch.Tensor]:
    """The same scenes as ``render_heads`` but as the heads' LOW-RESOLUTION outputs ``[b,.,h/scale,w/scale]`` (what the
    1x1 convolutions of smp's SegmentationHead emit before ``nn.UpsamplingBilinear2d(scale_factor=scale)``,
    lib/pose_regressor.py:633-666).  Discs are given in full-resolution pixels and mapped through the align_corners
    grid, so after up-sampling they sit where ``render_heads`` would put them."""
    if h % scale or w % scale:
        raise ValueError("h and w must be multiples of scale")
    hl, wl = h // scale, w // scale
    sy, sx = (hl - 1) / max(h - 1, 1), (wl - 1) / max(w - 1, 1)
    low = [[(cx * sx, cy * sy, r * 0.5 * (sx + sy), cls) for (cx, cy, r, cls) in discs] for discs in frames]
    return render_heads(low, hl, wl, num_classes=num_classes, seed=seed, device=device)
