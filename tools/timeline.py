#!/usr/bin/env python
"""Device timeline of the pipelined path (no nsys in this image): the library records a CUDA event between every two kernels
of every batch (fpc_recover_args.stage_events); with `depth` batches in flight on their own streams the (start, end) pairs
show which kernels of different batches ran at the same time.

    python tools/timeline.py [--depth 4] [--steps 12] [--out gpurun_out/r02_timeline.json]

Prints, per kernel, its mean duration alone-in-its-stream and the fraction of its run time during which at least one kernel
of ANOTHER batch was also running, plus the steady-state step time (launches are eager here, so the host can limit it)."""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from fastposecnn_b200 import _lib
    from fastposecnn_b200 import synthetic as syn
    from fastposecnn_b200.pose_recovery import PoseRecoveryPipeline
    ap = argparse.ArgumentParser()
    ap.add_argument("--depth", type=int, default=4)
    ap.add_argument("--steps", type=int, default=12)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--eager", action="store_true", help="eager launches + host-recorded events (host-bound; default: CUDA graphs with "
                                                         "event-record nodes, the bench's steady state)")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "r02_timeline.json"))
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    wl = syn.WORKLOADS[args.workload]
    bpg = 32 if args.workload in ("cfg2", "cfg3") else (4 if args.workload == "cfg4" else wl.batch)
    logits = syn.render_workload(wl, batch=bpg, seed=1000, device=dev)
    inv_k = torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()
    tn = [syn.disc_pixel_count(cx, cy, r, wl.h, wl.w) for (cx, cy, r, _c) in wl.discs()]
    n_exp = bpg * len(tn)
    pipe = PoseRecoveryPipeline(args.depth, bpg, wl.h, wl.w, wl.num_classes, wl.hyps, dev, max_instances=max(1024, 2 * n_exp), seed=1234)
    idxs = torch.zeros((pipe.engines[0].max_instances, wl.hyps, 2), dtype=torch.int32)
    idxs[:n_exp] = syn.presampled_idxs(tn * bpg, wl.hyps).reshape(n_exp, wl.hyps, 2)
    idxs = idxs.to(dev)
    L = _lib.lib()
    nk = L.fpc_pose_recover_num_launches()
    names = [L.fpc_pose_recover_kernel_name(k).decode() for k in range(nk)]
    for _ in range(2 * args.depth):
        pipe.submit(logits, inv_k, idxs=idxs)
    pipe.drain()
    torch.cuda.synchronize()
    def new_events():
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(nk + 1)]
        for e in ev:
            e.record()
        return ev
    spans = []     # (batch, kernel, start_ms, end_ms)
    if args.eager:
        evs = [new_events() for _ in range(args.steps)]
        torch.cuda.synchronize()
        origin = torch.cuda.Event(enable_timing=True)
        origin.record()
        for k in range(args.steps):
            pipe.submit(logits, inv_k, idxs=idxs, stage_events=evs[k])
        pipe.drain()
        pipe.join()
        torch.cuda.synchronize()
        for b, ev in enumerate(evs):
            t = [origin.elapsed_time(e) for e in ev]
            spans += [(b, names[k], t[k], t[k + 1]) for k in range(nk)]
        lo, hi = args.depth, args.steps - args.depth
    else:
        # every engine's graph carries its own %globaltimer stamps (one-thread kernels between the path's kernels); after the run
        # they hold the times of the engine's LAST replay, i.e. of the last `depth` batches
        stamps = [torch.zeros(nk + 1, dtype=torch.int64, device=dev) for _ in range(args.depth)]
        for e, st in zip(pipe.engines, stamps):
            e.capture(logits, inv_k, idxs=idxs, stage_stamps=st)
        torch.cuda.synchronize()
        for k in range(args.steps * args.depth):
            pipe.submit(replay=True)
        pipe.drain()
        pipe.join()
        torch.cuda.synchronize()
        t0 = min(int(st[0]) for st in stamps)
        for b, st in enumerate(stamps):
            t = [(int(v) - t0) * 1e-6 for v in st.tolist()]
            spans += [(b, names[k], t[k], t[k + 1]) for k in range(nk)]
        lo, hi = 0, args.depth
    steady = [s for s in spans if lo <= s[0] < hi]
    starts = sorted(min(s[2] for s in steady if s[0] == b) for b in range(lo, hi))
    step_ms = (starts[-1] - starts[0]) / max(len(starts) - 1, 1)          # mean distance between consecutive batch starts
    stats = {}
    for b, name, s0, s1 in steady:
        others = sorted((max(s0, o0), min(s1, o1)) for ob, _n, o0, o1 in spans if ob != b and o1 > s0 and o0 < s1)
        cov, cur = 0.0, s0
        for a0, a1 in others:
            if a1 > cur:
                cov += a1 - max(a0, cur)
                cur = a1
        st = stats.setdefault(name, [0, 0.0, 0.0])
        st[0] += 1
        st[1] += s1 - s0
        st[2] += cov
    mode = "eager launches, host-recorded events" if args.eager else "CUDA-graph replay, %globaltimer stamps between kernels"
    print(f"{args.workload} b{bpg}, {args.depth} batches in flight, {mode}: {step_ms * 1e3:.1f} us between batch starts")
    print(f"{'kernel':18s} {'mean us':>9s} {'overlapped with another batch':>30s}")
    for name in names:
        n, dur, cov = stats[name]
        print(f"{name:18s} {dur / n * 1e3:9.1f} {100 * cov / max(dur, 1e-12):29.0f}%")
    # text timeline: one row per batch, one column per 10 us; letters = the big kernels, '.' = the small ones
    big = {"k_argmax_runs": "A", "k_gather": "G", "k_vote": "V", "k_vote_settle": "s", "k_hypotheses": "h", "k_finalize": "f"}
    t_lo, t_hi = min(s[2] for s in steady), max(s[3] for s in steady)
    cols = int((t_hi - t_lo) / 0.01) + 1
    print(f"timeline, 10 us per column (A arg-max, G gather, h hypotheses, V vote, s settle, f finalize, . other kernels):")
    for b in range(lo, hi):
        row = [" "] * cols
        for bb, name, s0, s1 in steady:
            if bb != b:
                continue
            for c in range(int((s0 - t_lo) / 0.01), min(cols, int((s1 - t_lo) / 0.01) + 1)):
                if row[c] == " " or row[c] == ".":
                    row[c] = big.get(name, ".")
        print(f"batch {b}: " + "".join(row))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump({"workload": args.workload, "frames": bpg, "depth": args.depth, "step_us": step_ms * 1e3, "kernels": names,
               "spans_ms": [[b, n, round(s0, 5), round(s1, 5)] for b, n, s0, s1 in spans]}, open(args.out, "w"))


if __name__ == "__main__":
    main()
