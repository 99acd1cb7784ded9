"""Headline benchmark: pose estimates / second at num_envs=1024 (BASELINE.json), one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--num-envs 1024] [--precision fp16x2|bf16x3|bf16]

A step = one pass of the hot path (preprocess -> backbone x2 views -> plane-sweep volume -> 3-D U-Net -> decode ->
fit) over ALL num_envs environments of synthetic input.  With N > 1 (torchrun) the environments are sharded
contiguously over the ranks (no data-path collective) and the per-env poses are all-gathered over NCCL; the total
stays num_envs, i.e. strong scaling, as the metric is defined at num_envs=1024.

`value`  : device-resident inputs, CUDA-event timing of exactly K steps, max over ranks.
`e2e`    : the same metric through AdaPoseEstimator_v5.estimate() with pinned HOST inputs, host->device copies and the
           device->host read of the boxes inside the timed region.
`--impl reference`: the CPU oracle (a port of the reference's per-env loop; the reference itself does not travel to the
           GPU box) timed on the host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from rgbmanip_b200 import synth, weights  # noqa: E402

GF_BACKBONE_PER_FRAME = 57.0177e9       # SURVEY.md A.3 (2 * MACs of every conv of the PSPNet backbone)
GF_BACKBONE_TC_PER_FRAME = 57.0177e9 - 0.2360e9 - 0.1156e9 - 0.0128e9 - 0.0066e9   # minus conv1, layer2.0 strided convs, psp
GF_COSTREG_PER_VIEW = 24.4506e9
DECODE_BYTES_PER_VIEW = 1708092
# ncu capture of the 40 backbone tc_conv_kernel launches of one chunk (fp16x2: 148 frames; bf16x3: 128 frames): dram__bytes_read + write summed over the launch
# group, and sm__pipe_tensor_cycles_active time-weighted over it.  fp16x2: profiles/r01_chunk_by_kernel.csv;
# bf16x3: profiles/r01_bf16x3_backbone_tc_summary.csv
NCU_TC = {"fp16x2": (11141.8e6 / 148, 0.720, "profiles/r01_chunk_by_kernel.csv"),
          "bf16x3": (18958.8e6 / 128, 0.712, "profiles/r01_bf16x3_backbone_tc_summary.csv")}
CFG = {"name": "adapose_v5", "task_name": "one_drawer_cabinet", "load": False, "img_size": 224, "use_depth": True,
       "n_pts": 1024, "direct_regression": True, "real_world": False}


def workload_name(n):
    return (f"adapose_v5 estimate(), num_envs={n}, 2 views/env, 480x640 fp32 RGB + u8 mask -> [N,8,3] world boxes "
            "(BASELINE configs[3]; all four adapose_* yamls share this architecture)")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"], "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx = max(mx, float(r[1]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dist_setup(n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        import torch.distributed as dist
        if torch.cuda.is_available():
            torch.cuda.set_device(local)
        dist.init_process_group("nccl" if torch.cuda.is_available() else "gloo")
    return rank, world, local


def cpu_oracle_rate(n_envs, seed=0):
    """The oracle's per-env loop (= the reference's algorithm) on the host cores -> (estimates/s, threads, seconds)."""
    from oracle import adapose_oracle as O
    torch.set_num_threads(os.cpu_count())
    sd = weights.init_state_dict(0)
    batch = synth.make_batch(n_envs + 1, seed=seed, special=False)
    np.random.seed(0)
    O.estimate(sd, CFG, *batch.slice(0, 1).args())          # warm-up env
    t0 = time.perf_counter()
    O.estimate(sd, CFG, *batch.slice(1, n_envs + 1).args())
    dt = time.perf_counter() - t0
    return n_envs / dt, torch.get_num_threads(), dt


def run_reference(args, rank, world):
    if rank != 0:
        return
    sample = 6
    vals = []
    for i in range(args.warmup + args.steps):
        rate, threads, dt = cpu_oracle_rate(sample, seed=i)
        if i >= args.warmup:
            vals.append((rate, dt))
    rate = sample * len(vals) / sum(d for _, d in vals)
    line = {"metric": "pose estimates/sec at num_envs=1024", "value": rate, "unit": "estimates/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(d for _, d in vals) / len(vals),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "impl": "reference",
            "config": {"workload": workload_name(args.num_envs),
                       "note": "reference's per-env CPU loop (oracle port, torch CPU fp32 eval mode); cost is linear in num_envs"},
            "cpu_baseline": {"value": rate, "unit": "estimates/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} envs per step of the {args.num_envs}-env workload (per-env loop, linear)"},
            "e2e": {"value": rate, "unit": "estimates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def instrumented_pass(eng, n_env, E1, E2):
    """One extra, untimed-for-the-headline pass with a CUDA-event pair around every launch group."""
    ev = lambda: torch.cuda.Event(enable_timing=True)
    rec = []

    def timed(name, kind, fn):
        a, b = ev(), ev()
        a.record(); fn(); b.record()
        rec.append((name, kind, a, b))

    for name, op in eng.backbone_ops:
        timed(name, getattr(op, "kind", "aux"), lambda op=op: op(2 * n_env))
    lib, L = eng.lib, sys.modules["rgbmanip_b200._lib"]
    timed("stereo_all", "stereo", lambda: eng.stereo(n_env, E1, E2))
    for name, op in eng.cr_ops:
        timed(name, getattr(op, "kind", "aux") + "3d", lambda op=op: op(n_env))
    torch.cuda.synchronize()
    out = {}
    for name, kind, a, b in rec:
        d = out.setdefault(kind, {"ms": 0.0, "launches": 0})
        d["ms"] += a.elapsed_time(b); d["launches"] += 1
    return out, {name: a.elapsed_time(b) for name, kind, a, b in rec}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--num-envs", type=int, default=1024)
    ap.add_argument("--chunk", type=int, default=74,
                    help="envs per chunk; 74 = num_SMs / 2: 148 frames per backbone launch = whole waves of 128-row tiles on 148 SMs")
    ap.add_argument("--precision", default="fp16x2")
    ap.add_argument("--unique", type=int, default=32)
    ap.add_argument("--cpu-sample", type=int, default=32)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (profiling runs)")
    ap.add_argument("--ring", action="store_true", help="also time the device view ring (one new view per controller step, "
                                                        "cached features for the other; SURVEY 8(f)-1) and add it as `view_ring`")
    args = ap.parse_args()
    rank, world, local = dist_setup(args.gpus)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the product path has no CPU fallback)"
    import torch.distributed as dist
    from rgbmanip_b200 import _lib
    from rgbmanip_b200.estimator import AdaPoseEstimator_v5
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    lib = _lib.load()
    N = args.num_envs
    from rgbmanip_b200.dist import shard_range
    lo, hi, per = shard_range(N, rank, world)
    n_loc = hi - lo
    # synthetic inputs: `unique` distinct envs tiled over this rank's shard, resident in HBM and mirrored in pinned host memory
    base = synth.make_batch(args.unique, seed=100 + rank, special=True)
    idx = np.arange(n_loc) % args.unique
    host = {}
    for name, arr in zip(("K", "rgb1", "mask1", "E1", "rgb2", "mask2", "E2"), base.args()):
        t = torch.from_numpy(np.ascontiguousarray(arr))
        if name.startswith("mask"):
            t = t.to(torch.uint8)
        host[name] = t[idx].contiguous().pin_memory()
    devt = {k: v.to(dev) for k, v in host.items()}
    est = AdaPoseEstimator_v5(None, dict(CFG), None, state_dict=weights.init_state_dict(0), device=dev, max_envs=args.chunk,
                              precision=args.precision)
    eng = est.estimator
    order = ("K", "rgb1", "mask1", "E1", "rgb2", "mask2", "E2")
    gathered = torch.zeros((world * per, 8, 3), dtype=torch.float64, device=dev)

    def step_device():
        out = est.estimate(*[devt[k] for k in order], return_tensor=True)
        if world > 1:
            pad = out if n_loc == per else torch.cat([out, out.new_zeros((per - n_loc, 8, 3))])
            dist.all_gather_into_tensor(gathered, pad)
            return gathered
        return out

    def step_host():
        out = est.estimate(*[host[k] for k in order], return_tensor=True)
        if world > 1:
            pad = out if n_loc == per else torch.cat([out, out.new_zeros((per - n_loc, 8, 3))])
            dist.all_gather_into_tensor(gathered, pad)
            out = gathered
        return out.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        l0 = lib.adp_launch_count()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        barrier()
        wall = time.perf_counter() - t0
        ms = a.elapsed_time(b)
        launches = lib.adp_launch_count() - l0
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall, launches

    sampler = ClockSampler(local)
    sampler.start()
    ms, wall, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop()
    value = N * args.steps / (ms / 1e3)
    eng.check_error_flag()

    e2e = None
    if not args.no_e2e:
        ms_h, wall_h, _ = timed(step_host, max(1, args.steps), 1)
        h2d = sum(host[k].numel() * host[k].element_size() for k in order)
        e2e = {"value": N * max(1, args.steps) / wall_h, "unit": "estimates/s", "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(n_loc * 24 * 8), "timed": "host wall clock around estimate(); pinned host inputs"}

    ring_info = None
    if args.ring:
        from rgbmanip_b200.view_ring import ViewRing
        ring = ViewRing(est, n_loc, 5)
        views = [{"camera0": {"Color": devt["rgb1"], "Mask": devt["mask1"], "Intrinsic": devt["K"], "Extrinsic": devt["E1"]}},
                 {"camera0": {"Color": devt["rgb2"], "Mask": devt["mask2"], "Intrinsic": devt["K"], "Extrinsic": devt["E2"]}}]
        pose = torch.zeros((n_loc, 7), dtype=torch.float64, device=dev)
        state = {"t": 0}

        def step_ring():
            ring.add_view(views[state["t"] % 2], pose)
            ring.accumulate_steps += 1
            state["t"] += 1
            return ring.get_estimation(return_tensor=True)

        ms_r, _, _ = timed(step_ring, max(2, args.steps), 3)
        ring_info = {"value": N * max(2, args.steps) / (ms_r / 1e3), "unit": "estimates/s", "ms_per_step": ms_r / max(2, args.steps),
                     "what": "controller step = add_view (1 new frame per env: preprocess + backbone) + get_estimation (stereo head on "
                             "cached features), frames resident in HBM", "cache_gb": ring.feat.numel() * 6 / 1e9}
        del ring
        torch.cuda.empty_cache()

    # per-kernel-class timing of one chunk, live, on the launching stream
    n_chunk = min(eng.E, n_loc)
    E1c = devt["E1"][:n_chunk].contiguous(); E2c = devt["E2"][:n_chunk].contiguous()
    est.estimate(*[devt[k][:n_chunk] for k in order], return_tensor=True)
    # three instrumented passes, per-op median: a single pass right after the e2e leg catches the GPU mid clock ramp
    passes = [instrumented_pass(eng, n_chunk, E1c, E2c) for _ in range(3)]
    per_op = {k: float(np.median([p[1][k] for p in passes])) for k in passes[0][1]}
    classes = {k: {"ms": float(np.median([p[0][k]["ms"] for p in passes])), "launches": passes[0][0][k]["launches"]} for k in passes[0][0]}
    pk = peaks()
    frames = 2 * n_chunk
    npass = eng.npass
    tc_ms = classes.get("tc", {}).get("ms", 0.0)
    tc_flops = GF_BACKBONE_TC_PER_FRAME * frames
    roof = None
    if tc_ms > 0:
        ach = tc_flops / (tc_ms / 1e3) / 1e12
        roof = {"bound": "tensor", "kernel": "tc_conv_kernel (tcgen05 implicit-GEMM, backbone 2-D convs)", "achieved": ach,
                "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                "traffic": NCU_TC[eng.precision][0] * frames if eng.precision in NCU_TC else None,
                "traffic_note": "dram__bytes_read+write summed over the launch group, ncu capture of one chunk scaled to this chunk "
                                f"({NCU_TC[eng.precision][2] if eng.precision in NCU_TC else 'no capture'})",
                "tensor_pipe_active_ncu": NCU_TC[eng.precision][1] if eng.precision in NCU_TC else None,
                "peak_source": pk["src"] + " (sustained: timed inside a long step)",
                "algorithmic_flops_per_launch_group": tc_flops, "launches": classes["tc"]["launches"],
                "tensor_pipe_work_frac": ach * npass / pk["tf_sustained"],
                "note": f"algorithmic FLOPs (SURVEY A.3) / summed CUDA-event time of the {classes['tc']['launches']} launches of one chunk "
                        f"({frames} frames); precision {eng.precision} issues {npass} MMA pass(es) per algorithmic FLOP"}
    dec_ms = per_op.get("stereo_all", 0.0) - sum(v for k, v in per_op.items() if k.startswith("cr."))
    kernels = {k: {"ms_per_chunk": v["ms"], "launches": v["launches"]} for k, v in classes.items()}
    kernels["chunk_envs"] = n_chunk
    kernels["costreg_tflops"] = (GF_COSTREG_PER_VIEW * n_chunk / (sum(v for k, v in per_op.items() if k.startswith("cr.")) / 1e3) / 1e12
                                 if any(k.startswith("cr.") for k in per_op) else None)
    kernels["volume_decode_fit_ms"] = dec_ms

    if rank == 0:
        cpu_rate, threads, cpu_dt = (0.0, 0, 0.0) if args.no_cpu else cpu_oracle_rate(args.cpu_sample)
        line = {"metric": "pose estimates/sec at num_envs=1024", "value": value, "unit": "estimates/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "fp16" if eng.precision == "fp16x2" else "bf16", "data": f"synthetic ({args.unique} seeded envs tiled to {N})",
                "config": {"workload": workload_name(N),
                           "precision": eng.precision, "chunk_envs": eng.E, "sharding": f"env-sharded dp{world}, NCCL all-gather of poses",
                           "l2": "inputs (>= 8 GB per step) exceed the 126 MB L2; no explicit flush needed",
                           "sampling": "device hash sampler for the 1024-pixel subset"},
                "roofline": roof, "cpu_baseline": {"value": cpu_rate, "unit": "estimates/s", "cores": threads, "kind": "port",
                                                   "sample": f"{args.cpu_sample} envs of the workload through the oracle's per-env loop ({cpu_dt:.1f} s)"},
                "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "kernels": kernels,
                "wall_s_timed_region": wall}
        if ring_info is not None:
            line["view_ring"] = ring_info
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
