"""Runs the head-epilogue-fused path (low-resolution inputs, cfg2 scenes) a few times; target of ncu captures."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastposecnn_b200 import synthetic as syn  # noqa: E402
from fastposecnn_b200.pose_recovery import PoseRecoveryEngine  # noqa: E402

dev = torch.device("cuda:0")
wl = syn.WORKLOADS[os.environ.get("WL", "cfg2")]
b = int(os.environ.get("B", wl.batch))
low = syn.render_lowres_heads([wl.discs()] * b, wl.h, wl.w, 4, wl.num_classes, seed=1000, device=dev)
inv_k = torch.inverse(syn.camera_intrinsics()).to(dev)
eng = PoseRecoveryEngine(b, wl.h, wl.w, wl.num_classes, wl.hyps, dev, max_instances=max(1024, 2 * b * len(wl.discs())), upsample=4)
for _ in range(int(os.environ.get("ITERS", 3))):
    eng.launch(low, inv_k)
    print("instances", eng.fetch_count())
