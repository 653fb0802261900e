"""BASELINE.json configs[4]: a random-init encoder/decoder IN PLAIN TORCH that only FEEDS the pose-recovery path.

Not part of the product (the network stays in torch, SURVEY.md section 2.1: `lib/pose_regressor.py` rest = out of
scope).  It restates the *shape* of the reference's PoseRegressor (lib/pose_regressor.py:582-743: one encoder, four
FPN decoders, four 1x1 heads, xyz split into xy / z per class, :729-732) with torchvision's ResNet-18, because
`segmentation_models_pytorch` is not installed here.

    python examples/network_feed.py [--batch 64] [--steps 5]      # frames/s of network + path on one GPU
"""
from __future__ import annotations

import argparse
import os
import sys

import torch
import torch.nn as nn
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class _FPNDecoder(nn.Module):
    def __init__(self, chans=(64, 128, 256, 512), width=64):
        super().__init__()
        self.lateral = nn.ModuleList(nn.Conv2d(c, width, 1) for c in chans)
        self.smooth = nn.Conv2d(width, width, 3, padding=1)

    def forward(self, feats):
        x = self.lateral[-1](feats[-1])
        for lat, f in zip(reversed(self.lateral[:-1]), reversed(feats[:-1])):
            x = F.interpolate(x, size=f.shape[-2:], mode="nearest") + lat(f)
        return F.relu(self.smooth(x))            # 1/4 resolution


class TorchFeeder(nn.Module):
    """images [b,3,h,w] -> LogitData (mask [b,C,h,w], quaternion [b,4(C-1),h,w], scales, xy, z)."""

    def __init__(self, num_classes: int = 7):
        super().__init__()
        import torchvision
        r = torchvision.models.resnet18(weights=None)
        self.stem = nn.Sequential(r.conv1, r.bn1, r.relu, r.maxpool)
        self.layers = nn.ModuleList([r.layer1, r.layer2, r.layer3, r.layer4])
        k = num_classes - 1
        self.k = k
        self.decoders = nn.ModuleList(_FPNDecoder() for _ in range(4))
        self.heads = nn.ModuleList(nn.Conv2d(64, c, 1) for c in (num_classes, 4 * k, 3 * k, 3 * k))

    def forward(self, x):
        h, w = x.shape[-2:]
        f = self.stem(x)
        feats = []
        for layer in self.layers:
            f = layer(f)
            feats.append(f)
        outs = [F.interpolate(head(dec(feats)), size=(h, w), mode="bilinear", align_corners=False)
                for dec, head in zip(self.decoders, self.heads)]
        mask, quat, xyz, scales = outs
        idx = torch.arange(3 * self.k, device=x.device)
        xy = xyz[:, idx[idx % 3 != 2]]            # channels 3k, 3k+1 (lib/pose_regressor.py:729-732)
        z = xyz[:, idx[idx % 3 == 2]]             # channel 3k+2
        return {"mask": mask.contiguous(), "quaternion": quat.contiguous(), "scales": scales.contiguous(),
                "xy": xy.contiguous(), "z": z.contiguous()}


def main():
    from fastposecnn_b200 import synthetic as syn
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--hn", type=int, default=128)
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    net = TorchFeeder().to(dev).eval()
    imgs = torch.randn(args.batch, 3, 480, 640, device=dev)
    inv_k = torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()
    eng = PoseRecoveryEngine(args.batch, 480, 640, 7, args.hn, dev, max_instances=65536)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    tn = tp = 0.0
    with torch.no_grad():
        for it in range(args.steps + 2):
            ev[0].record()
            logits = net(imgs)
            ev[1].record()
            eng.launch(logits, inv_k)
            n = eng.fetch_count()
            ev[2].record()
            torch.cuda.synchronize()
            if it >= 2:
                tn += ev[0].elapsed_time(ev[1])
                tp += ev[1].elapsed_time(ev[2])
    print(f"batch {args.batch}: network {tn / args.steps:.2f} ms, pose recovery {tp / args.steps:.3f} ms ({n} instances), "
          f"{args.batch / ((tn + tp) / args.steps * 1e-3):.0f} frames/s end to end")


if __name__ == "__main__":
    main()
