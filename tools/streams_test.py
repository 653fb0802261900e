"""Throughput with NS pipeline slots, each replaying its CUDA graph on its OWN stream (kernels of different batches may
then overlap: the HBM-bound arg-max / gather of one batch with the FP32-bound vote of another)."""
import sys, time, torch
sys.path.insert(0, '.')
from fastposecnn_b200 import synthetic as syn
from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device('cuda:0'); wl = syn.WORKLOADS['cfg2']; bpg = 32; hn = wl.hyps
logits = syn.render_workload(wl, batch=bpg, seed=1000, device=dev)
inv_k = torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()
discs = wl.discs(); tn = [syn.disc_pixel_count(cx, cy, r, wl.h, wl.w) for (cx, cy, r, _c) in discs]; n_exp = bpg * len(discs)
engs = [PoseRecoveryEngine(bpg, wl.h, wl.w, 7, hn, dev, max_instances=max(1024, 2 * n_exp)) for _ in range(NS)]
idxs = torch.zeros((engs[0].max_instances, hn, 2), dtype=torch.int32); idxs[:n_exp] = syn.presampled_idxs(tn * bpg, hn).reshape(n_exp, hn, 2); idxs = idxs.to(dev)
streams = [torch.cuda.Stream() for _ in range(NS)]
for e in engs:
    e.capture(logits, inv_k, idxs=idxs)
torch.cuda.synchronize()
def submit(k):
    s = k % NS
    with torch.cuda.stream(streams[s]):
        engs[s].replay(); engs[s].enqueue_fetch()
def finish(k):
    assert engs[k % NS].wait_count() == n_exp
for k in range(2 * NS): submit(k)
for k in range(2 * NS): finish(k)
torch.cuda.synchronize()
K = 60
t0 = time.perf_counter()
for k in range(K):
    if k >= NS: finish(k - NS)
    submit(k)
for k in range(K - NS, K): finish(k)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print(f"streams={NS}: {dt*1e3:.4f} ms/step  {bpg/dt:.0f} frames/s")
