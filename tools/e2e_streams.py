"""Does overlapping batch k's sparse gather with batch k+1's mask arg-max (both zero-copy over PCIe) help?"""
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
NS = int(sys.argv[1]) if len(sys.argv) > 1 else 2
engs = [PoseRecoveryEngine(bpg, wl.h, wl.w, 7, hn, dev, max_instances=max(1024, 2 * n_exp)) for _ in range(NS)]
streams = [torch.cuda.Stream() for _ in range(NS)]
idxs = torch.zeros((engs[0].max_instances, hn, 2), dtype=torch.int32); idxs[:n_exp] = syn.presampled_idxs(tn * bpg, hn).reshape(n_exp, hn, 2); idxs = idxs.to(dev)
tables = [torch.empty((engs[0].max_instances, _lib.POSE_ROW), dtype=torch.float32).pin_memory() for _ in range(NS)]
def submit(k):
    s = k % NS
    with torch.cuda.stream(streams[s]):
        engs[s].launch(host, inv_k, idxs=idxs)
        engs[s].enqueue_fetch()
def finish(k):
    s = k % NS
    n = engs[s].wait_count()
    with torch.cuda.stream(streams[s]):
        tables[s][:n].copy_(engs[s].pose_table[:n], non_blocking=True)
    return n
torch.cuda.synchronize()
for k in range(4): submit(k); 
for k in range(4): finish(k)
torch.cuda.synchronize()
import time
K = 12
t0 = time.perf_counter()
for k in range(K):
    submit(k)
    if k >= NS - 1: assert finish(k - (NS - 1)) == n_exp
for k in range(K - (NS - 1), K): finish(k)
torch.cuda.synchronize()
dt = (time.perf_counter() - t0) / K
print(f"streams={NS}: {dt*1e3:.3f} ms/step  {bpg/dt:.0f} frames/s")
