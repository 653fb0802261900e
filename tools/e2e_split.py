"""Per-kernel time split of the zero-copy e2e path (pinned host head maps read in place over PCIe)."""
import sys, torch
sys.path.insert(0, '.')
from fastposecnn_b200 import synthetic as syn, _lib
from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
dev = torch.device('cuda:0'); wl = syn.WORKLOADS['cfg2']; bpg = 32; hn = wl.hyps
logits = syn.render_workload(wl, batch=bpg, seed=1000, device=dev)
host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in logits.items()}
for k, v in logits.items(): host[k].copy_(v)
inv_k = torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()
discs = wl.discs(); tn = [syn.disc_pixel_count(cx, cy, r, wl.h, wl.w) for (cx, cy, r, _c) in discs]; n_exp = bpg * len(discs)
eng = PoseRecoveryEngine(bpg, wl.h, wl.w, 7, hn, dev, max_instances=max(1024, 2 * n_exp))
idxs = torch.zeros((eng.max_instances, hn, 2), dtype=torch.int32); idxs[:n_exp] = syn.presampled_idxs(tn * bpg, hn).reshape(n_exp, hn, 2); idxs = idxs.to(dev)
nk = eng.num_launches
names = [_lib.lib().fpc_pose_recover_kernel_name(k).decode() for k in range(nk)]
evs = [torch.cuda.Event(enable_timing=True) for _ in range(nk + 1)]
for e in evs: e.record()
for _ in range(3):
    eng.launch(host, inv_k, idxs=idxs, stage_events=evs); assert eng.fetch_count() == n_exp
acc = [0.0] * nk
for _ in range(5):
    eng.launch(host, inv_k, idxs=idxs, stage_events=evs); eng.fetch_count()
    for k in range(nk): acc[k] += evs[k].elapsed_time(evs[k + 1]) / 5
for n, t in zip(names, acc): print(f"{n:20s} {t:8.3f} ms")
print("total", sum(acc), " argmax GB/s", 275.25e6 / (acc[0] * 1e-3) / 1e9, " gather GB/s (algorithmic 88.8 MB)", 88.77e6 / (acc[names.index('k_gather')] * 1e-3) / 1e9)
