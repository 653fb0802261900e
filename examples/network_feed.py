"""BASELINE.json configs[4]: a random-init encoder/decoder IN PLAIN TORCH that only FEEDS the pose-recovery path.

Not part of the product (the network stays in torch, SURVEY.md section 2.1: `lib/pose_regressor.py` rest = out of
scope).  It restates the *shape* of the reference's PoseRegressor (lib/pose_regressor.py:582-743: one encoder, four
FPN decoders, four 1x1 heads, xyz split into xy / z per class, :729-732) with torchvision's ResNet-18, because
`segmentation_models_pytorch` is not installed here.

    python examples/network_feed.py [--batch 64] [--steps 5]      # frames/s of network + path on one GPU

Two flows are timed: the reference's (heads up-sample all 67 channels x4 with nn.UpsamplingBilinear2d, then the
full-resolution path) and the fused one (only the heads' 1x1 convolutions run in torch; the path interpolates inside its
kernels).
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
        # smp's SegmentationHead (lib/pose_regressor.py:633-666): Conv2d(k=1) -> UpsamplingBilinear2d(x4) -> identity
        self.heads = nn.ModuleList(nn.Sequential(nn.Conv2d(64, c, 1), nn.UpsamplingBilinear2d(scale_factor=4), nn.Identity())
                                   for c in (num_classes, 4 * k, 3 * k, 3 * k))

    def decode(self, x):
        """The four decoder outputs at 1/4 resolution (mask, rotation, translation, scales)."""
        f = self.stem(x)
        feats = []
        for layer in self.layers:
            f = layer(f)
            feats.append(f)
        return [dec(feats) for dec in self.decoders]

    def lowres(self, x):
        """Head convolutions only: the low-resolution LogitData for ``pose_recover(..., upsample=4)``."""
        from fastposecnn_b200 import lowres_logits
        d = self.decode(x)
        names = ("mask", "rotation", "translation", "scales")
        return lowres_logits(dict(zip(names, self.heads)), dict(zip(names, d)))

    def forward(self, x):
        outs = [head(d) for head, d in zip(self.heads, self.decode(x))]
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
    for fused in (False, True):
        eng = PoseRecoveryEngine(args.batch, 480, 640, 7, args.hn, dev, max_instances=65536, upsample=4 if fused else 1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        tn = tp = 0.0
        with torch.no_grad():
            for it in range(args.steps + 2):
                ev[0].record()
                logits = net.lowres(imgs) if fused else net(imgs)
                ev[1].record()
                eng.launch(logits, inv_k)
                n = eng.fetch_count()
                ev[2].record()
                torch.cuda.synchronize()
                if it >= 2:
                    tn += ev[0].elapsed_time(ev[1])
                    tp += ev[1].elapsed_time(ev[2])
        flow = "fused head epilogue (convs only + low-res path)" if fused else "reference flow (x4 up-sampled heads + path)     "
        print(f"batch {args.batch} {flow}: network {tn / args.steps:.2f} ms, pose recovery {tp / args.steps:.3f} ms "
              f"({n} instances), {args.batch / ((tn + tp) / args.steps * 1e-3):.0f} frames/s end to end")
        del eng


if __name__ == "__main__":
    main()
