#!/usr/bin/env python
"""Benchmark of the pose-recovery hot path (BASELINE.json metric: aggregation+voting frames/s @640x480 b32).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path (head maps -> per-instance pose table) over one batch of
synthetic NOCS-shaped head outputs: BASELINE.json configs[1] (b=32 frames of 640x480, 6 classes x 3
instances, 128 fixed pre-sampled hypotheses per instance) PER GPU (weak scaling: frames are sharded by
image, the only collective is the all-gather of the per-instance pose tables).

Prints ONE JSON line on rank 0 (see the task contract); `value` = frames/s with inputs resident in HBM,
`e2e` = frames/s through the public API with HOST (pinned) head maps, H2D/D2H inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "aggregation+voting frames/s @640x480 b32"
UNIT = "frames/s"
# dram__bytes_read.sum + dram__bytes_write.sum per launch, extracted from the committed `ncu --set full` capture of this very
# command by tools/ncu_traffic.py: {"<workload>_b<frames per GPU>": {"<kernel>": bytes}}
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")


def load_ncu_traffic(workload: str, bpg: int):
    try:
        with open(NCU_TRAFFIC_FILE) as f:
            return json.load(f).get(f"{workload}_b{bpg}", {})
    except (OSError, ValueError):
        return {}


# the reference's four timed stages (lib/pose_regressor.py:43-48, 563-570; tools/timer.py) and the kernels that do their work
STAGES = {
    "class_compression": ("k_argmax_runs",),
    "aggregate": ("k_scan_tiles", "k_emit_runs", "k_run_merge", "k_run_flatten", "k_scan_roots", "k_run_assign", "k_run_stats",
                  "k_scan_slots", "k_run_slots", "k_scan_records", "k_gather"),
    "hough_voting": ("k_hypotheses", "k_vote", "k_vote_settle"),
    "perform_RT_calculation": ("k_finalize",),
}


# --------------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------------
def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(local_rank: int):
    """Pins this process (and therefore the pages of the pinned buffers it allocates afterwards: first touch) to the CPUs of
    the NUMA node its GPU hangs off, read from sysfs.  Returns a description for the bench line (no numactl in this image)."""
    info = {"gpu": local_rank, "numa_node": None, "cpus": None, "bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(int(os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",")[local_rank]) if os.environ.get("CUDA_VISIBLE_DEVICES") else local_rank)
        busid = pynvml.nvmlDeviceGetPciInfo(h).busId
        busid = busid.decode() if isinstance(busid, bytes) else busid
        busid = busid.lower()[-12:]                       # 0000:3b:00.0
        with open(f"/sys/bus/pci/devices/{busid}/numa_node") as f:
            node = int(f.read().strip())
        info["numa_node"] = node
        if node >= 0:
            with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
                cpulist = f.read().strip()
            cpus = set()
            for part in cpulist.split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
            allowed = cpus & os.sched_getaffinity(0)
            if allowed:
                os.sched_setaffinity(0, allowed)
                info.update({"cpus": cpulist, "bound": True})
    except Exception as e:                                  # not fatal: the numbers are then simply not NUMA-local
        info["error"] = str(e)[:120]
    return info


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons of one GPU while the timed region runs."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 8:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_frames_per_s(wl, frames_per_step: int, steps: int, warmup: int):
    """The reference's CPU path on a bounded sample of the workload, all host threads: the reference's OWN modules
    (lib/gpu_tensor_funcs.py, aggregation_layer.py, hough_voting.py, ransac_voting_gpu.py, unmodified, installed under
    baseline/_ref by __graft_entry__.build()) on CPU tensors -- torch-CPU aggregation + scipy CCL + the C restatement of
    the two voting kernels behind the pybind module's names (kind "reference"); the oracle port if they are not installed
    (kind "port").  Returns (frames/s, s per step, cores, OpenMP threads, kind)."""
    from fastposecnn_b200 import synthetic as syn
    from oracle import native, port, ref_import
    ncores = os.cpu_count() or 1
    torch.set_num_threads(ncores)
    logits = syn.render_workload(wl, batch=frames_per_step, seed=0, device="cpu")
    inv_k = torch.inverse(syn.camera_intrinsics())
    kind = "reference" if ref_import.available() else "port"
    times = []
    import warnings
    for it in range(warmup + steps):
        t0 = time.perf_counter()
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if kind == "reference":
                ref_import.reference_pose_recover(logits, inv_k, wl.hyps, port.seeded_idx_source(1234))
            else:
                port.pose_recover(logits, inv_k, wl.hyps, idx_source=port.seeded_idx_source(1234))
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    total = sum(times)
    return frames_per_step * len(times) / total, total / len(times), ncores, native.num_threads(), kind


# --------------------------------------------------------------------------------------------
# reference arm
# --------------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from fastposecnn_b200 import synthetic as syn
    wl = syn.WORKLOADS.get(args.workload, syn.WORKLOADS["cfg2"])     # cfg5: the path's share, on cfg2-shaped head maps
    frames_per_step = args.ref_frames
    fps, s_per_step, ncores, omp, kind = cpu_reference_frames_per_s(wl, frames_per_step, args.steps, args.warmup)
    what = ("the reference's own lib/ modules (baseline/_ref, unmodified) on CPU tensors" if kind == "reference"
            else "oracle/port.py on CPU tensors")
    sample = (f"{frames_per_step} frames of {wl.name} per step ({args.steps} timed steps after {args.warmup} warm-up), "
              f"{what}, torch threads={ncores}, OpenMP threads={omp}")
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": s_per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl.name, "frames_per_step": frames_per_step, "hypotheses": wl.hyps,
                   "instances_per_frame": wl.instances_per_frame},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": ncores, "kind": kind, "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    from fastposecnn_b200 import _lib
    from fastposecnn_b200 import synthetic as syn
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine, PoseRecoveryPipeline
    from fastposecnn_b200.sharding import OverlappedGather

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device: the path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    numa = bind_to_gpu_numa_node(local_rank) if not args.no_numa_bind else {"bound": False}
    if args.l2_fetch_granularity:
        import ctypes
        torch.zeros(1, device=dev)
        rt = ctypes.CDLL("libcudart.so.12")
        rc = rt.cudaDeviceSetLimit(5, ctypes.c_size_t(args.l2_fetch_granularity))      # cudaLimitMaxL2FetchGranularity = 0x05
        got = ctypes.c_size_t(0)
        rt.cudaDeviceGetLimit(ctypes.byref(got), 5)
        print(f"cudaLimitMaxL2FetchGranularity: set rc={rc}, now {got.value}", file=sys.stderr)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = syn.WORKLOADS[args.workload]
    # frames per GPU: cfg2 32 (weak scaling); cfg3 = 256 frames image-sharded over the ranks (32 per GPU at 8; one GPU alone
    # runs one such 32-frame shard); cfg4 = 4 frames of 1280x960; cfg1 = 1
    default_bpg = {"cfg3": max(32, wl.batch // max(world, 8)) if world == 1 else wl.batch // world, "cfg4": 4}.get(args.workload, wl.batch)
    bpg = args.batch_per_gpu or default_bpg
    hn = wl.hyps
    hbm_peak, peak_src = load_peaks()

    # ---- synthetic inputs, resident in HBM (2.6 GB per 32 frames: far larger than the 126 MB L2) ----
    logits = syn.render_workload(wl, batch=bpg, seed=1000 + rank, device=dev)
    inv_k = torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()
    discs = wl.discs()
    tn_disc = [syn.disc_pixel_count(cx, cy, r, wl.h, wl.w) for (cx, cy, r, _c) in discs]
    n_expected = bpg * len(discs)
    depth = max(1, args.pipeline_depth)
    pipe = PoseRecoveryPipeline(depth, bpg, wl.h, wl.w, wl.num_classes, hn, dev, max_instances=max(1024, 2 * n_expected),
                                multi_stream=not args.single_stream)
    eng = pipe.engines[0]
    idxs = torch.zeros((eng.max_instances, hn, 2), dtype=torch.int32)
    idxs[:n_expected] = syn.presampled_idxs(tn_disc * bpg, hn, seed=1234).reshape(n_expected, hn, 2)
    idxs = idxs.to(dev)
    # the one collective of the path: all-gather of the pose tables, on its own stream (overlaps the next batch)
    gatherer = OverlappedGather(pipe.engines, world, dev) if (world > 1 and not args.no_gather) else None

    nk = eng.num_launches
    kernel_names = [_lib.lib().fpc_pose_recover_kernel_name(k).decode() for k in range(nk)]

    def make_events(n):
        evs = [torch.cuda.Event(enable_timing=True) for _ in range(n)]
        for e in evs:
            e.record()           # instantiates the cudaEvent_t
        return evs

    def after_launch(e):
        if gatherer is not None:
            gatherer.after_launch(e)

    def before_launch(e):
        if gatherer is not None:
            gatherer.before_reuse(e)

    def check(res):
        if res is not None and res[1] != n_expected:
            raise RuntimeError(f"synthetic workload produced {res[1]} instances, expected {n_expected}")

    use_graph = not args.no_graph

    def step(stage_events=None, replay=False):
        # one pass of the path over one batch; the host waits for the count of the batch `depth-1` steps back
        check(pipe.submit(logits, inv_k, idxs=idxs, stage_events=stage_events, after_launch=after_launch, replay=replay,
                          before_launch=before_launch))

    # ---- warm-up ----
    for _ in range(max(args.warmup, 3)):
        step()
    for res in pipe.drain():
        check(res)
    if use_graph:
        pipe.capture(logits, inv_k, idxs=idxs)          # the launches of one step become one graph launch
        for _ in range(2 * depth):
            step(replay=True)
        for res in pipe.drain():
            check(res)
    n = n_expected
    step_events = [make_events(nk + 1) for _ in range(args.steps)]
    torch.cuda.synchronize()

    # ---- FP32 peak of this chip (pure FFMA loop), denominator of the voting roofline ----
    sink = torch.zeros(4, dtype=torch.float32, device=dev)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    fma_blocks, fma_iters = sms * 16, 1 << 14
    best_ms = None
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.check(_lib.lib().fpc_bench_fp32_fma(sink.data_ptr(), fma_blocks, fma_iters, _lib.current_stream(dev)))
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best_ms = ms if best_ms is None else min(best_ms, ms)
    fp32_peak_tflops = fma_blocks * 256 * fma_iters * 16 * 2 / (best_ms * 1e-3) / 1e12

    # ---- timed region: K steps, device-timed, barrier + synchronize on both sides ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(args.steps):
        step(replay=use_graph)
    for res in pipe.drain():
        check(res)
    pipe.join()                                                      # every slot stream is inside the timed region
    if gatherer is not None:
        torch.cuda.current_stream().wait_stream(gatherer.stream)     # ... and so are the last all-gathers
    t_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    elapsed_ms = t_start.elapsed_time(t_end)
    # ---- same K steps again, instrumented: serial (one stream, eager launches), the library records a CUDA event
    #      between every two kernels.  The 16 event records per step and the missing cross-batch overlap make this pass
    #      slower than the headline loop; it only attributes time to kernels, each measured running alone.
    pipe_i = PoseRecoveryPipeline(2, bpg, wl.h, wl.w, wl.num_classes, hn, dev, max_instances=max(1024, 2 * n_expected),
                                  multi_stream=False)
    for _ in range(3):
        check(pipe_i.submit(logits, inv_k, idxs=idxs))
    for res in pipe_i.drain():
        check(res)
    torch.cuda.synchronize()
    i_start, i_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    i_start.record()
    for k in range(args.steps):
        check(pipe_i.submit(logits, inv_k, idxs=idxs, stage_events=step_events[k]))
    for res in pipe_i.drain():
        check(res)
    i_end.record()
    torch.cuda.synchronize()
    instrumented_ms_per_step = i_start.elapsed_time(i_end) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    per_rank_ms = [elapsed_ms / args.steps]
    if world > 1:
        # every rank's own device time: the reported step is the MAX (contract); the spread tells rank skew from a real
        # scaling cost (median close to the 1-GPU step = the extra is the slowest GPU of the box, not the collective)
        allt = torch.zeros(world, dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(allt, torch.tensor([elapsed_ms], dtype=torch.float64, device=dev))
        per_rank_ms = [v / args.steps for v in allt.tolist()]
        elapsed_ms = max(allt.tolist())
    ms_per_step = elapsed_ms / args.steps
    value = world * bpg / (ms_per_step * 1e-3)

    # per-kernel durations inside the timed region (events recorded by the library between launches)
    kernel_ms = [0.0] * nk
    for evs in step_events:
        for k in range(nk):
            kernel_ms[k] += evs[k].elapsed_time(evs[k + 1])
    kernel_ms = [v / args.steps for v in kernel_ms]

    # ---- algorithmic work ----
    hw = wl.h * wl.w
    fg = sum(tn_disc)
    bytes_argmax = 4 * wl.num_classes * hw * bpg                      # 28 B per pixel: the 7 mask logits
    bytes_gather = 40 * fg * bpg                                      # predicted class's 4+3+2+1 floats, fg pixels only
    bytes_agg = (4 * wl.num_classes * hw + 40 * fg + 184 * len(discs)) * bpg
    flop_vote = sum(12.0 * tn * (hn + 1) + 4.0 * tn for tn in tn_disc) * bpg
    i_arg, i_gather, i_vote = kernel_names.index("k_argmax_runs"), kernel_names.index("k_gather"), kernel_names.index("k_vote")
    vote_side = {kernel_names.index(k) for k in ("k_hypotheses", "k_vote", "k_vote_settle") if k in kernel_names}
    agg_ms = sum(kernel_ms[k] for k in range(nk) if k not in vote_side)
    ncu_traffic = load_ncu_traffic(args.workload, bpg)

    def hbm(bytes_, ms, kernel=None):
        a = bytes_ / (ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                "traffic": ncu_traffic.get(kernel)}
    roof_argmax = dict(hbm(bytes_argmax, kernel_ms[i_arg], "k_argmax_runs"), kernel="k_argmax_runs", ms=kernel_ms[i_arg],
                       algorithmic_bytes=bytes_argmax, peak_source=peak_src)
    roof_gather = dict(hbm(bytes_gather, kernel_ms[i_gather], "k_gather"), kernel="k_gather", ms=kernel_ms[i_gather],
                       algorithmic_bytes=bytes_gather, peak_source=peak_src)
    roof_agg = dict(hbm(bytes_agg, agg_ms), kernel="all aggregation kernels (everything but the voting kernels)", ms=agg_ms,
                    algorithmic_bytes=bytes_agg, peak_source=peak_src)
    a_v = flop_vote / (kernel_ms[i_vote] * 1e-3) / 1e12
    roof_vote = {"bound": "fp32", "achieved": a_v, "peak": fp32_peak_tflops, "unit": "TFLOP/s", "frac": a_v / fp32_peak_tflops,
                 "traffic": ncu_traffic.get("k_vote"), "kernel": "k_vote", "ms": kernel_ms[i_vote], "algorithmic_flop": flop_vote,
                 "votes_per_s": flop_vote / 12.0 / (kernel_ms[i_vote] * 1e-3),
                 "peak_source": "measured in this run: fpc_bench_fp32_fma (pure FFMA loop), 2 flop per FMA; the voting kernel is "
                                "bound by the FP32 pipe, not by HBM or the tensor cores (BASELINE.json north_star)"}
    dominant = max(range(nk), key=lambda k: kernel_ms[k])
    per_kernel = {"k_argmax_runs": roof_argmax, "k_gather": roof_gather, "k_vote": roof_vote}
    stage_ms = {st: round(sum(kernel_ms[kernel_names.index(k)] for k in ks if k in kernel_names), 5) for st, ks in STAGES.items()}

    # ---- end to end through the public API with HOST buffers ----
    e2e = None
    if not args.no_e2e:
        host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in logits.items()}
        for k, v in logits.items():
            host[k].copy_(v)
        table_host = torch.empty((eng.max_instances, _lib.POSE_ROW), dtype=torch.float32).pin_memory()
        if args.e2e_mode == "copy":
            # plain staging: H2D copy of all 67 channels, then the device-resident path
            dev_in = {k: torch.empty_like(v) for k, v in logits.items()}
            h2d = sum(v.numel() * v.element_size() for v in host.values())
        else:
            # zero-copy ingestion: the kernels read the pinned host head maps in place over PCIe; only the 7 mask
            # logits of every pixel and the predicted class's 10 channels of foreground pixels cross the bus
            dev_in = host
            h2d = bytes_argmax + bytes_gather

        # `--e2e-depth` batches in flight (default 1; with 2, one stream each): while batch k's gather kernel pulls its scattered
        # foreground rows over PCIe, batch k+1's arg-max streams its mask logits -- the bus stays busy.  Every step still
        # pays its own host->device traffic and its own device->host read of N and of the pose table.
        e2e_depth = max(1, args.e2e_depth) if args.e2e_mode == "zerocopy" else 1     # staged copies share one input buffer
        pipe_e = PoseRecoveryPipeline(e2e_depth, bpg, wl.h, wl.w, wl.num_classes, hn, dev, max_instances=max(1024, 2 * n_expected),
                                      multi_stream=e2e_depth > 1)
        e2e_gather = OverlappedGather(pipe_e.engines, world, dev) if (world > 1 and not args.no_gather) else None

        def e2e_after(e):
            if e2e_gather is not None:
                e2e_gather.after_launch(e)

        def e2e_before(e):
            if e2e_gather is not None:
                e2e_gather.before_reuse(e)

        def e2e_collect(res):
            eng_o, n_o = res
            table_host[:n_o].copy_(eng_o.pose_table[:n_o], non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return n_o

        def e2e_step():
            if args.e2e_mode == "copy":
                for k in host:
                    dev_in[k].copy_(host[k], non_blocking=True)
            res = pipe_e.submit(dev_in, inv_k, idxs=idxs, after_launch=e2e_after, before_launch=e2e_before)
            return e2e_collect(res) if res is not None else None

        def e2e_drain():
            n_last = None
            for res in pipe_e.drain():
                n_last = e2e_collect(res)
            if e2e_gather is not None:
                e2e_gather.finish()
            return n_last
        n_ = None
        for _ in range(2 * e2e_depth):
            n_ = e2e_step() or n_
        n_ = e2e_drain() or n_
        d2h = n_ * _lib.POSE_ROW * 4 + _lib.NUM_COUNTERS * 4
        ksteps = max(3, min(args.steps, 10))
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s0.record()
        for _ in range(ksteps):
            e2e_step()
        e2e_drain()
        s1.record()
        torch.cuda.synchronize()
        ems = s0.elapsed_time(s1)
        if world > 1:
            t = torch.tensor([ems], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ems = float(t.item())
        # ceiling of the host side: plain pinned -> device copies of the mask logits on ALL ranks at once
        link_dst = torch.empty_like(logits["mask"])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0.record()
        for _ in range(3):
            link_dst.copy_(host["mask"], non_blocking=True)
        l1.record()
        torch.cuda.synchronize()
        link_ms = l0.elapsed_time(l1) / 3
        if world > 1:
            t = torch.tensor([link_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            link_ms = float(t.item())
        link_gbs = world * host["mask"].numel() * 4 / (link_ms * 1e-3) / 1e9
        del link_dst
        e2e = {"value": world * bpg / (ems / ksteps * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "host_to_device_copy_gbs_all_ranks": link_gbs,
               "achieved_host_read_gbs_all_ranks": world * h2d / (ems / ksteps * 1e-3) / 1e9,
               "ms_per_step": ems / ksteps, "steps": ksteps,
               "mode": args.e2e_mode, "batches_in_flight": e2e_depth,
               "note": ("pinned host head maps read in place by the kernels (zero-copy over PCIe: h2d bytes = algorithmic "
                        "28 B/px + 40 B/fg px) -> D2H of N and the pose table, every step" if args.e2e_mode == "zerocopy" else
                        "pinned host head maps -> H2D copy of all 67 channels -> fpc_pose_recover -> D2H of N and the pose table")}

    # ---- CPU baseline beside it (rank 0, N=1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        fps, s_per, ncores, omp, kind = cpu_reference_frames_per_s(wl, args.ref_frames, 3, 1)
        cpu = {"value": fps, "unit": UNIT, "cores": ncores, "kind": kind,
               "sample": f"{args.ref_frames} frames of {wl.name}, 1 warm-up + 3 timed passes of "
                         + ("the reference's own lib/ modules (baseline/_ref)" if kind == "reference" else "oracle/port.py")
                         + f" (torch-CPU aggregation + scipy CCL + C voting kernels, OpenMP threads={omp})"}

    matching = head_epilogue = None
    if rank == 0 and world == 1 and not args.no_matching:
        matching = matching_section(dev, wl, logits, inv_k, hbm_peak, peak_src)
    if rank == 0 and world == 1 and not args.no_head_epilogue and wl.h % 4 == 0 and wl.w % 4 == 0:
        head_epilogue = head_epilogue_section(dev, wl, bpg, hn, inv_k, args.steps, depth, nk, kernel_names)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl.name, "frames_per_gpu": bpg, "global_batch": world * bpg, "hypotheses": hn,
                       "instances_per_frame": len(discs), "parallelism": f"image-sharded x{world}",
                       "l2": "inputs (2.6 GB of head maps per GPU) are larger than the 126 MB L2; no flush needed",
                       "arith": "IEEE (un-contracted) voting arithmetic", "timed": f"{nk} kernels (one CUDA graph) + D2H read of N per step; {depth} steps in flight"
                                + ("" if args.single_stream else ", one stream each (kernels of different batches overlap)")
                                + (" + NCCL all-gather of pose tables" if world > 1 else "")},
            # the dominant kernel = the one with the largest share of the step (live CUDA-event times below)
            "roofline": per_kernel.get(kernel_names[dominant], roof_vote),
            "roofline_per_kernel": [roof_vote, roof_argmax, roof_gather],
            "roofline_fp32_voting": roof_vote,
            "roofline_gather": roof_gather,
            "roofline_aggregation_total": roof_agg,
            "dominant_kernel": kernel_names[dominant],
            "stage_ms": stage_ms,
            "stage_ms_note": "the reference's four timed stages (lib/pose_regressor.py:43-48) = sums of the kernels that do their "
                             "work: " + "; ".join(f"{st} = {' + '.join(ks)}" for st, ks in STAGES.items())
                             + " (the class selection / normalisation of class_compress runs inside k_gather, the refinement "
                               "solve of hough_voting inside k_finalize)",
            "kernel_ms": {kernel_names[k]: round(kernel_ms[k], 5) for k in range(nk)},
            "kernel_ms_note": "per-kernel CUDA-event times from an instrumented, serial repeat of the same K steps "
                              f"({instrumented_ms_per_step:.4f} ms/step: one stream, eager launches, 16 event records per "
                              "step; the headline loop replays graphs on one stream per in-flight batch)",
            "cuda_graph": use_graph,
            "fp32_peak_tflops_measured": fp32_peak_tflops,
            "cpu_baseline": cpu,
            "e2e": e2e,
            "matching": matching,
            "head_epilogue": head_epilogue,
            "gpu_launches": nk * args.steps,
            "clocks": clocks,
            "instances": n,
            "numa": numa,
            "all_gather": (world > 1 and not args.no_gather),
            "ms_per_step_per_rank": {"min": min(per_rank_ms), "median": statistics.median(per_rank_ms), "max": max(per_rank_ms)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------------
# cfg5: BASELINE.json configs[4] -- random-init encoder/decoder in torch feeding the path
# --------------------------------------------------------------------------------------------
def run_cfg5(args):
    """One step = 64 frames per GPU (weak scaling): images -> encoder / 4 FPN decoders / 4 head convolutions (plain torch,
    random init: examples/network_feed.TorchFeeder restates the SHAPE of lib/pose_regressor.py:582-743) -> low-resolution head
    maps -> fpc_pose_recover(upsample=4) -> pose table; network and path are captured in ONE CUDA graph
    (lib/pose_regressor.py:706-770, inference.py:84-95).  Next to it the reference flow: the heads up-sample all 67
    channels x4 in torch, then the full-resolution path."""
    import torch.distributed as dist
    from examples.network_feed import TorchFeeder
    from fastposecnn_b200 import _lib
    from fastposecnn_b200 import synthetic as syn
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device: the path has no CPU fallback")
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import datetime
        # a collective that cannot complete fails after two minutes instead of holding the GPUs until the job is killed
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
    b, h, w, C, hn = args.batch_per_gpu or 64, 480, 640, 7, 128
    # the SAME random network and images on every rank: the instance count, hence the grown table capacity, hence the size of the
    # all-gather buffers must agree across ranks (with per-rank seeds the 8-GPU run hung; unequal buffers are the likely cause)
    torch.manual_seed(1234)
    net = TorchFeeder(C).to(dev).eval()
    host_imgs = torch.randn(b, 3, h, w).pin_memory()
    imgs = torch.empty((b, 3, h, w), device=dev)
    imgs.copy_(host_imgs)
    inv_k = torch.inverse(syn.camera_intrinsics()).to(dev).contiguous()

    def make_engine(upsample, caps):
        return PoseRecoveryEngine(b, h, w, C, hn, dev, upsample=upsample, seed=1234, **caps)

    def settle_capacity(upsample, feed):
        """Random-init heads give thousands of speckle instances: grow the tables until one batch fits."""
        caps = {"max_instances": 16384}
        for _ in range(6):
            eng = make_engine(upsample, caps)
            with torch.no_grad():
                eng.launch(feed(imgs), inv_k)
            try:
                return eng, eng.fetch_count()
            except _lib.CapacityError as e:
                mi, mr, mrec = e.grown(eng.max_instances, eng.max_rows, eng.max_records, b * h * w, h)
                caps = {"max_instances": mi, "max_rows": mr, "max_records": mrec}
                del eng
        raise RuntimeError("cfg5: capacity did not settle")

    results = {}
    with torch.no_grad():
        for name, upsample, feed in (("fused", 4, net.lowres), ("reference_flow", 1, net.forward)):
            eng, n_inst = settle_capacity(upsample, feed)
            # eager, instrumented: network vs path
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            tn = tp = 0.0
            reps = max(3, min(args.steps, 5))
            for it in range(reps + 2):
                ev[0].record()
                logits = feed(imgs)
                ev[1].record()
                eng.launch(logits, inv_k)
                eng.enqueue_fetch()
                ev[2].record()
                eng.wait_count()
                torch.cuda.synchronize()
                if it >= 2:
                    tn += ev[0].elapsed_time(ev[1])
                    tp += ev[1].elapsed_time(ev[2])
            results[name] = {"network_ms": tn / reps, "path_ms": tp / reps, "instances": n_inst,
                             "frames_per_s_eager": b / ((tn + tp) / reps * 1e-3)}
            if name == "reference_flow":
                del eng
                continue
            # one CUDA graph: H2D of the images, network, path
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=side):
                    imgs.copy_(host_imgs, non_blocking=True)
                    eng.launch(net.lowres(imgs), inv_k)
                graph_dev = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph_dev, stream=side):
                    eng.launch(net.lowres(imgs), inv_k)
            torch.cuda.current_stream(dev).wait_stream(side)
            table_host = torch.empty((eng.max_instances, _lib.POSE_ROW), dtype=torch.float32).pin_memory()
            gathered = torch.empty((world,) + tuple(eng.table_full.shape), dtype=torch.float32, device=dev) if world > 1 else None
            if world > 1:
                cap = torch.tensor([eng.max_instances, -eng.max_instances], dtype=torch.int64, device=dev)
                dist.all_reduce(cap, op=dist.ReduceOp.MAX)
                if int(cap[0]) != -int(cap[1]):
                    raise RuntimeError(f"cfg5: table capacities differ across ranks ({-int(cap[1])}..{int(cap[0])}); the all-gather needs equal sizes")

            def step(g, read_table):
                g.replay()
                if gathered is not None:
                    dist.all_gather_into_tensor(gathered.view(-1), eng.table_full.view(-1))
                n = eng.fetch_count()
                if read_table:
                    table_host[:n].copy_(eng.pose_table[:n], non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                return n
            sampler = ClockSampler(local_rank)
            timed = {}
            for key, g, read_table in (("resident", graph_dev, False), ("e2e", graph, True)):
                for _ in range(max(args.warmup, 3)):
                    step(g, read_table)
                if rank == 0 and key == "resident":
                    sampler.start()
                    time.sleep(0.3)
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(args.steps):
                    n_last = step(g, read_table)
                t1.record()
                torch.cuda.synchronize()
                ms = t0.elapsed_time(t1)
                if world > 1:
                    t = torch.tensor([ms], dtype=torch.float64, device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                timed[key] = ms / args.steps
            clocks = sampler.stop() if rank == 0 else None
            results[name].update({"graph_ms_per_step": timed["resident"], "e2e_ms_per_step": timed["e2e"], "instances": n_last,
                                  "d2h": n_last * _lib.POSE_ROW * 4 + _lib.NUM_COUNTERS * 4})
    if rank == 0:
        fused = results["fused"]
        line = {
            "metric": "end-to-end inference frames/s @640x480 b64 (torch encoder/decoder + pose recovery)", "unit": UNIT,
            "value": world * b / (fused["graph_ms_per_step"] * 1e-3), "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": fused["graph_ms_per_step"], "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "cfg5_b64_640x480_resnet18fpn_random_init_hn128", "frames_per_gpu": b, "global_batch": world * b,
                       "hypotheses": hn, "parallelism": f"image-sharded x{world}",
                       "network": "torchvision ResNet-18 + 4 FPN decoders + 4 1x1 heads, random init, fp32 (cuDNN; out of scope, it only "
                                  "feeds the path)",
                       "l2": "every step streams 236 MB of images and ~1 GB of activations: far larger than the 126 MB L2",
                       "timed": "one CUDA graph per step (network + 16 path kernels) + D2H read of N"
                                + (" + NCCL all-gather of pose tables" if world > 1 else "")},
            "roofline": None,
            "cfg5": results,
            "e2e": {"value": world * b / (fused["e2e_ms_per_step"] * 1e-3), "unit": UNIT, "h2d_bytes_per_step": host_imgs.numel() * 4,
                    "d2h_bytes_per_step": fused["d2h"], "ms_per_step": fused["e2e_ms_per_step"],
                    "note": "pinned host images -> H2D copy (inside the graph) -> network -> path -> D2H of N and the pose table"},
            "cpu_baseline": None,
            "gpu_launches": 16 * args.steps,
            "clocks": clocks,
            "instances": fused["instances"],
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def head_epilogue_section(dev, wl, bpg, hn, inv_k, steps, depth, nk, kernel_names):
    """SURVEY.md section 8f rank 2 beside the headline: the same scenes handed over as the heads' LOW-RESOLUTION outputs
    ([b,67,h/4,w/4], what the 1x1 convolutions of smp's SegmentationHead emit), the x4 bilinear up-sampling evaluated
    inside the arg-max and gather kernels.  Next to it: what the reference flow spends on producing the full-resolution
    maps with torch's own CUDA up-sampling kernel, which the fused path never runs."""
    from fastposecnn_b200 import _lib
    from fastposecnn_b200 import synthetic as syn
    from fastposecnn_b200.pose_recovery import PoseRecoveryEngine, PoseRecoveryPipeline
    S = 4
    low = syn.render_lowres_heads([wl.discs()] * bpg, wl.h, wl.w, S, wl.num_classes, seed=1000, device=dev)
    cap = max(1024, 2 * bpg * len(wl.discs()))
    pipe = PoseRecoveryPipeline(depth, bpg, wl.h, wl.w, wl.num_classes, hn, dev, max_instances=cap, upsample=S)
    n = None
    for _ in range(3):
        pipe.submit(low, inv_k)
    for _e, n in pipe.drain():
        pass
    pipe.capture(low, inv_k)
    for _ in range(2 * depth):
        pipe.submit(replay=True)
    pipe.drain()
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(steps):
        pipe.submit(replay=True)
    pipe.drain()
    pipe.join()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / steps
    # per-kernel times: serial instrumented pass
    pipe_i = PoseRecoveryPipeline(2, bpg, wl.h, wl.w, wl.num_classes, hn, dev, max_instances=cap, upsample=S, multi_stream=False)
    evs = []
    for k in range(3 + steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(nk + 1)]
        for e in ev:
            e.record()
        pipe_i.submit(low, inv_k, stage_events=ev)
        if k >= 3:
            evs.append(ev)
    pipe_i.drain()
    torch.cuda.synchronize()
    kms = [sum(ev[k].elapsed_time(ev[k + 1]) for ev in evs) / len(evs) for k in range(nk)]
    # the reference flow's up-sampling (torch CUDA kernel) of the same 67 channels, and the full-resolution path on its output
    up = torch.nn.UpsamplingBilinear2d(scale_factor=S)
    full = {k: up(v) for k, v in low.items()}
    torch.cuda.synchronize()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record()
    for _ in range(5):
        for v in low.values():
            up(v)
    u1.record()
    torch.cuda.synchronize()
    up_ms = u0.elapsed_time(u1) / 5
    eng_full = PoseRecoveryEngine(bpg, wl.h, wl.w, wl.num_classes, hn, dev, max_instances=cap)
    for _ in range(3):
        eng_full.launch(full, inv_k)
    n_full = eng_full.fetch_count()
    eng_full.capture(full, inv_k)
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    f0.record()
    for _ in range(steps):
        eng_full.replay()
    f1.record()
    torch.cuda.synchronize()
    full_ms = f0.elapsed_time(f1) / steps
    eng_low = pipe.engines[0]
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    g0.record()
    for _ in range(steps):
        eng_low.replay()
    g1.record()
    torch.cuda.synchronize()
    low_serial_ms = g0.elapsed_time(g1) / steps
    same = bool(torch.equal(eng_low.cat_mask_u8, eng_full.cat_mask_u8))
    # end to end from pinned host buffers: H2D copy of the low-resolution maps (1/16 of the bytes), the path, D2H of the table
    host = {k: torch.empty(v.shape, dtype=v.dtype).pin_memory() for k, v in low.items()}
    for k, v in low.items():
        host[k].copy_(v)
    stage = {k: torch.empty_like(v) for k, v in low.items()}
    table_host = torch.empty((eng_low.max_instances, _lib.POSE_ROW), dtype=torch.float32).pin_memory()

    def e2e_step():
        for k in host:
            stage[k].copy_(host[k], non_blocking=True)
        eng_low.launch(stage, inv_k)
        n_ = eng_low.fetch_count()
        table_host[:n_].copy_(eng_low.pose_table[:n_], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return n_
    for _ in range(2):
        e2e_step()
    ksteps = max(3, min(steps, 10))
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    h0.record()
    for _ in range(ksteps):
        e2e_step()
    h1.record()
    torch.cuda.synchronize()
    e2e_ms = h0.elapsed_time(h1) / ksteps
    h2d = sum(v.numel() * v.element_size() for v in host.values())
    return {
        "what": "same scenes as low-resolution head outputs [b,67,%d,%d]; x%d bilinear up-sampling fused into k_argmax_runs / k_gather"
                % (wl.h // S, wl.w // S, S),
        "value": bpg / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "instances": n,
        "kernel_ms": {kernel_names[k]: round(kms[k], 5) for k in range(nk)},
        "one_stream_graph_replay_ms": {"lowres_fused": low_serial_ms, "full_resolution_path_on_torch_upsampled_maps": full_ms,
                                       "torch_cuda_upsampling_of_67_channels": up_ms},
        "class_map_identical_to_full_resolution_path": same, "instances_full_resolution_path": n_full,
        "input_bytes_per_step": {"lowres": h2d, "full_resolution": h2d * S * S},
        "e2e": {"value": bpg / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": n * _lib.POSE_ROW * 4 + _lib.NUM_COUNTERS * 4,
                "note": "pinned host low-resolution head outputs -> H2D copy -> fused path -> D2H of N and the pose table"},
    }


def matching_section(dev, wl, logits, inv_k, hbm_peak, peak_src, iters=10):
    """SURVEY.md section 8f rank 1 beside the headline: the ground-truth <-> prediction matching step that follows the
    path (lib/matching.py:226-325), on this batch's own predictions against ground truths made by shifting them by
    (3,-2) px.  Kernels are timed alone with CUDA events through the C ABI on preallocated buffers; the whole
    ``batchwise_find_matches`` call is timed through the drop-in API; and the reference's algorithm (per class, expand
    both mask sets to [n1,n2,h,w]) is timed written in torch on the same GPU."""
    import fastposecnn_b200 as fp
    from fastposecnn_b200 import _lib, matching

    def timed(fn, n=iters, warm=3):
        for _ in range(warm):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    preds = fp.pose_recover(logits, inv_k, wl.hyps, materialize_dense=True)
    n = int(preds["class_ids"].shape[0])
    # the drop-in voting drivers on the reference's dense per-instance layout (what HoughVotingLayer hands them), SURVEY 8f rank 3
    from fastposecnn_b200.ransac_voting_gpu_layer import ransac_voting_gpu as rvg
    vertex = preds["xy_mask"].permute(0, 2, 3, 1).unsqueeze(3)              # the reference's non-contiguous [N,h,w,1,2] view
    drivers = {
        "ransac_voting_layer_v3 (dense [N,h,w] masks + [N,h,w,1,2] vertex view)":
            timed(lambda: rvg.ransac_voting_layer_v3(preds["instance_masks"], vertex, wl.hyps), n=5, warm=2),
        "ransac_voting_layer_v4 (+ residual variance)":
            timed(lambda: rvg.ransac_voting_layer_v4(preds["instance_masks"], vertex, wl.hyps), n=5, warm=2),
        "ransac_voting_layer_v5 (+ confidence, max_num=30000)":
            timed(lambda: rvg.ransac_voting_layer_v5(preds["instance_masks"], vertex, wl.hyps, max_num=30000), n=5, warm=2),
    }
    del vertex
    gts = {k: v.clone() for k, v in preds.items() if k not in ("labels", "cat_mask", "xy_mask")}
    gts["instance_masks"] = torch.roll(gts["instance_masks"], shifts=(3, -2), dims=(1, 2)).contiguous()
    gts["symmetric_ids"] = gts["class_ids"] % 2
    preds = {k: v for k, v in preds.items() if k != "xy_mask"}
    h, w = gts["instance_masks"].shape[1:]
    L, st = _lib.lib(), _lib.current_stream(dev)
    gs, ps = matching.MaskSet(n, h, w, dev), matching.MaskSet(n, h, w, dev)
    iou = torch.empty((n, n), dtype=torch.float32, device=dev)
    best, biou = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
    pairs, nm = torch.empty((n, 2), dtype=torch.int32, device=dev), torch.empty(1, dtype=torch.int32, device=dev)
    gm, lab, gc, pc = gts["instance_masks"], preds["labels"], gts["class_ids"].contiguous(), preds["class_ids"].contiguous()
    b = lab.shape[0]
    t_pack = timed(lambda: _lib.check(L.fpc_pack_masks(gm.data_ptr(), _lib.MASK_F32, n, h, w, gs.bits.data_ptr(), gs.meta.data_ptr(), st)))
    t_lab = timed(lambda: _lib.check(L.fpc_pack_labels(lab.data_ptr(), b, h, w, n, ps.bits.data_ptr(), ps.meta.data_ptr(), st)))
    t_iou = timed(lambda: _lib.check(L.fpc_mask_iou(gs.bits.data_ptr(), gs.meta.data_ptr(), n, ps.bits.data_ptr(), ps.meta.data_ptr(), n,
                                                    h, w, iou.data_ptr(), st)))
    t_pair = timed(lambda: _lib.check(L.fpc_match_instances(gs.bits.data_ptr(), gs.meta.data_ptr(), gc.data_ptr(), n, ps.bits.data_ptr(),
                                                            ps.meta.data_ptr(), pc.data_ptr(), n, h, w, best.data_ptr(), biou.data_ptr(),
                                                            pairs.data_ptr(), nm.data_ptr(), st)))
    res = {}
    sparse = {k: v for k, v in preds.items() if k != "instance_masks"}
    t_dense = timed(lambda: res.__setitem__("d", fp.batchwise_find_matches(preds, gts)), n=5, warm=2)
    t_sparse = timed(lambda: res.__setitem__("s", fp.batchwise_find_matches(sparse, gts)), n=5, warm=2)
    m = int(res["d"]["class_ids"].shape[0])
    same = all(torch.equal(res["d"][k], res["s"][k]) for k in res["d"])

    def reference_style():
        k = 0
        for c in torch.unique(gts["class_ids"]):
            gi, pi = torch.where(gts["class_ids"] == c)[0], torch.where(preds["class_ids"] == c)[0]
            if gi.shape[0] == 0 or pi.shape[0] == 0:
                continue
            m1, m2 = gts["instance_masks"][gi], preds["instance_masks"][pi]
            e1 = m1.unsqueeze(1).expand((m1.shape[0], m2.shape[0], h, w))
            e2 = m2.expand((m1.shape[0], m2.shape[0], h, w))
            v, _ = torch.max(torch.logical_and(e1, e2).sum(dim=(2, 3)) / torch.logical_or(e1, e2).sum(dim=(2, 3)), dim=1)
            k += int((v > 0).sum())
        return k
    t_ref = timed(lambda: res.__setitem__("r", reference_style()), n=2, warm=1)
    # evaluation maths on the matched pairs: one launch each
    mt = res["d"]
    evals = {
        "get_quat_distance (360-rotation symmetry for odd classes)":
            timed(lambda: fp.get_quat_distance(mt["quaternion"][0], mt["quaternion"][1], mt["symmetric_ids"])),
        "get_3d_ious": timed(lambda: fp.get_3d_ious(mt["RT"][0], mt["RT"][1], mt["scales"][0], mt["scales"][1])),
        "from_Ts_get_offset_error": timed(lambda: fp.from_Ts_get_offset_error(mt["T"][0], mt["T"][1])),
    }
    # training support (second half): forward + backward of class_compression -> AggregationLayer on this batch
    import types
    layer = fp.AggregationLayer(types.SimpleNamespace(HV_NUM_OF_HYPOTHESES=wl.hyps), wl.num_classes, max_instances=max(1024, 2 * n))
    train_in = {k: (v.detach().clone().requires_grad_(True) if k != "mask" else v) for k, v in logits.items()}

    def train_step():
        for k, v in train_in.items():
            if k != "mask":
                v.grad = None
        agg_t = layer(fp.class_compression(train_in, wl.num_classes))
        (agg_t["quaternion"].sum() + agg_t["scales"].sum() + agg_t["z"].sum()).backward()
    t_train = timed(train_step, n=3, warm=1)
    del layer, train_in
    pack_bytes = n * h * w * 4 + n * h * ((w + 31) // 32) * 4
    a = pack_bytes / (t_pack * 1e-3) / 1e9
    stack_bytes = 4 * n * h * w * 4          # read + write of the gt rows and of the matched prediction rows
    return {
        "what": "batchwise_find_matches (lib/matching.py:226-325) on this batch: n_gt = n_pred = %d masks of %dx%d" % (n, w, h),
        "matches": m, "pairs_identical_dense_vs_label_volume": bool(same), "reference_algorithm_matches": int(res["r"]),
        "find_matches_ms": {"dense_predictions": t_dense, "label_volume_predictions": t_sparse,
                            "reference_algorithm_torch_same_gpu": t_ref},
        "speedup_vs_reference_algorithm_same_gpu": t_ref / t_dense,
        "kernel_ms": {"k_pack_masks_v4 (+meta init)": t_pack, "k_pack_labels (+memset, meta init)": t_lab, "k_mask_iou": t_iou,
                      "k_match_best + k_match_order (+memset)": t_pair},
        "voting_drivers_ms": drivers, "evaluation_ms": evals,
        "training_forward_backward_ms": {"class_compression + AggregationLayer, losses on q / scales / z": t_train},
        "roofline": {"bound": "hbm", "kernel": "k_pack_masks_v4", "achieved": a, "peak": hbm_peak, "unit": "GB/s", "frac": a / hbm_peak,
                     "algorithmic_bytes": pack_bytes, "traffic": None, "peak_source": peak_src,
                     "note": "4 B/px read once + 1 bit/px written; the stacked [2,M,h,w] instance_masks output of the call "
                             "(%.2f GB moved) is what the remaining time of find_matches goes to" % (stack_bytes / 1e9)},
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[0..4]; cfg5 = random-init encoder/decoder in torch feeding the path (b=64)")
    ap.add_argument("--batch-per-gpu", type=int, default=0, help="frames per GPU (default: 32 for cfg2/cfg3, 4 for cfg4, 64 for cfg5)")
    ap.add_argument("--ref-frames", type=int, default=2, help="frames per CPU-reference pass (bounded sample)")
    ap.add_argument("--pipeline-depth", type=int, default=4, help="batches in flight (1 = wait for N after every step)")
    ap.add_argument("--single-stream", action="store_true", help="all pipeline slots on one stream (no cross-batch overlap)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches in the timed loop instead of CUDA-graph replay")
    ap.add_argument("--e2e-mode", default="zerocopy", choices=["zerocopy", "copy"])
    ap.add_argument("--e2e-depth", type=int, default=2,
                    help="batches in flight in the end-to-end loop, one stream each (measured, profiles/r02_e2e_depth.txt: 2 keeps the PCIe "
                         "link busy across the batch boundary, +7 %% over 1; 3 and 4 add nothing)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-gather", action="store_true", help="N > 1: skip the all-gather of the pose tables (attribution experiment only)")
    ap.add_argument("--no-numa-bind", action="store_true", help="do not pin the process to the CPUs of its GPU's NUMA node")
    ap.add_argument("--no-head-epilogue", action="store_true", help="skip the low-resolution-input (SURVEY 8f rank 2) measurements")
    ap.add_argument("--no-matching", action="store_true", help="skip the matching (SURVEY 8f rank 1) measurements")
    ap.add_argument("--l2-fetch-granularity", type=int, default=0, choices=[0, 32, 64, 128],
                    help="experiment: cudaLimitMaxL2FetchGranularity (bytes) before the run; 0 = leave the driver default")
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "b200" and world != args.gpus:
        if args.gpus > 1:
            raise SystemExit(f"--gpus {args.gpus} needs torchrun with {args.gpus} ranks (WORLD_SIZE={world})")
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "cfg5":
        run_cfg5(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
