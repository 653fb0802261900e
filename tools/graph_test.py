import sys, time, torch
sys.path.insert(0,'/root/repo')
from fastposecnn_b200 import synthetic as syn, _lib
from fastposecnn_b200.pose_recovery import PoseRecoveryEngine
dev=torch.device('cuda:0'); wl=syn.WORKLOADS['cfg2']; bpg=32; hn=wl.hyps
logits=syn.render_workload(wl,batch=bpg,seed=1000,device=dev)
inv_k=torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()
discs=wl.discs(); tn=[syn.disc_pixel_count(cx,cy,r,wl.h,wl.w) for (cx,cy,r,_c) in discs]; n_exp=bpg*len(discs)
engs=[PoseRecoveryEngine(bpg,wl.h,wl.w,7,hn,dev,max_instances=max(1024,2*n_exp)) for _ in range(2)]
idxs=torch.zeros((engs[0].max_instances,hn,2),dtype=torch.int32); idxs[:n_exp]=syn.presampled_idxs(tn*bpg,hn).reshape(n_exp,hn,2); idxs=idxs.to(dev)
for e in engs:
    for _ in range(3): e.launch(logits,inv_k,idxs=idxs); assert e.fetch_count()==n_exp
# eager timing
def run(fn,K=40):
    torch.cuda.synchronize(); s=torch.cuda.Event(enable_timing=True); t=torch.cuda.Event(enable_timing=True); s.record()
    for k in range(K): fn(k)
    t.record(); torch.cuda.synchronize(); return s.elapsed_time(t)/K
def eager(k):
    e=engs[k%2]; e.launch(logits,inv_k,idxs=idxs); e.enqueue_fetch()
    if k>0: engs[(k-1)%2].wait_count()
print('eager pipelined ms/step', run(eager))
graphs=[]
side=torch.cuda.Stream()
for e in engs:
    g=torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        e.launch(logits,inv_k,idxs=idxs); side.synchronize()
        with torch.cuda.graph(g, stream=side):
            e.launch(logits,inv_k,idxs=idxs)
    graphs.append(g)
torch.cuda.synchronize()
def graphed(k):
    e=engs[k%2]; graphs[k%2].replay(); e.enqueue_fetch()
    if k>0: engs[(k-1)%2].wait_count()
print('graph pipelined ms/step', run(graphed))
print('N', engs[0].wait_count(), engs[1].wait_count())
