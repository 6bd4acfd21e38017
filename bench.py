#!/usr/bin/env python
"""
bench.py — CaVE loss+grad throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload tsp50]

A "step" is one pass of the hot path (scan/pack of A + projection + push-inside + cosine loss +
reduction + analytic backward) over one synthetic batch: TSP-50 DFJ binding constraints
(d = 1225, m = 1325 + k tight cuts), CaVE+ inner_ratio 0.2, batch 4096 per GPU
(BASELINE.json configs[2], the configuration the metric is quoted on).  Instances shard across
GPUs by instance with no data-path collective (weak scaling: 4096 per GPU).

Prints ONE JSON line (rank 0).  `value` is device-resident throughput (inputs already in HBM, the
dense [B, m_max, d] float32 layout the reference API hands over, pack rebuilt every step = cold);
`e2e` is the same metric through the module call with HOST tensors (pinned), including the
host->device copy of pred_cost and tight_ctrs and the device->host read of loss and gradient.
`--impl reference` times the reference's CPU path (scipy nnls per instance + the torch epilogue,
restated in oracle/cave_oracle.py; the Python reference itself cannot travel to the GPU box) on
all host cores on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "cave_loss_grad_instances_per_sec"
UNIT = "instances/s"
WORKLOADS = {  # name -> (synth kind, per-GPU batch, mode, inner_ratio)
    "tsp50": ("tsp50", 4096, 1, 0.2),
    "tsp20": ("tsp20", 4096, 1, 0.2),
    "vrp20": ("vrp20", 4096, 1, 0.2),
    "sp5": ("sp5", 4096, 0, 0.0),
}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tsp50", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override (0 = workload default)")
    ap.add_argument("--precision", default="fp64", choices=["fp64", "fp32"])
    ap.add_argument("--regime", default="uniform", choices=["uniform", "near"])
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="instances in the CPU baseline sample (0 = 2 x cores)")
    return ap.parse_args()


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_one(args):
    """One instance through the reference's CPU path (oracle restatement): projection by
    scipy.optimize.nnls + target + loss + gradient."""
    from oracle import cave_oracle as O
    pred, ctr, mode, ratio = args
    out = O.forward_backward(pred[None], ctr[None], minimize=True, mode=mode, inner_ratio=ratio, reduction="none")
    return float(out["loss_i"][0])


def _cpu_pool(cores):
    import multiprocessing as mp
    for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[k] = "1"
    return mp.get_context("fork").Pool(cores)


def _cpu_sample(kind, n, mode, regime, seed):
    from cave_b200 import synth
    insts = synth.make_batch(kind, n, seed=seed)
    m_max = max(i.m for i in insts)
    pred = synth.predictions(insts, seed, regime)
    return [(pred[b], insts[b].dense(m_max), mode, 0.2) for b in range(n)]


def run_reference(args):
    """Reference arm: rank 0 alone; each step = `cores` instances of the workload on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, batch, mode, ratio = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    per_step = args.cpu_sample or cores
    pool = _cpu_pool(cores)
    work = _cpu_sample(kind, per_step, mode, args.regime, seed=1)
    for _ in range(max(args.warmup, 0)):
        pool.map(_cpu_one, work[:cores])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pool.map(_cpu_one, work)
    dt = time.perf_counter() - t0
    pool.close()
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.workload} (SURVEY App. B synthetic), CaVE+ inner_ratio 0.2, solver='nnls' CPU path",
                   "sample_instances_per_step": per_step, "regime": args.regime},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} instances/step x {args.steps} steps, multiprocessing.Pool({cores}) over "
                                   "oracle.forward_backward (scipy.optimize.nnls + epilogue), BLAS threads = 1"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self.thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {0x2: "applications_clocks", 0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x10: "sync_boost",
                 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake", 0x100: "display"}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, n in names.items():
                    if r & bit:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(0.02)

    def start(self):
        if self.nv:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join()
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ----------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    import torch.distributed as dist

    from cave_b200 import EPO, _lib, cave_forward_backward, innerConeAlignedCosine, exactConeAlignedCosine, pack_constraints, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs CUDA devices (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    kind, batch, mode, ratio = WORKLOADS[args.workload]
    batch = args.batch or batch
    # every rank owns its own shard of the global batch (no data-path collective)
    insts = synth.make_batch(kind, batch, seed=1000 + rank)
    A = synth.densify(insts, device=dev)
    B, m_max, d = A.shape
    pred_np = synth.predictions(insts, 1000 + rank, args.regime)
    pred = torch.tensor(pred_np, device=dev)
    alg_bytes = float(sum(4 * i.m * d + 8 * d + 4 for i in insts))
    gen_rows = [int((np.bincount(i.rows, minlength=i.m) > 1).sum()) for i in insts]
    del insts
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak, peak_src = (peaks["hbm_gbs"], "measured") if "hbm_gbs" in peaks else (6650.0, "fallback")

    def step():
        return cave_forward_backward(pred, A, -1.0, mode, ratio, "mean", precision=args.precision)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        """n calls of fn bracketed by barrier + synchronize; CUDA events on the launching stream."""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms / n

    for _ in range(max(args.warmup, 3)):
        out = step()
    sampler = ClockSampler(local)
    sampler.start()
    ms_step = timed(step, args.steps)
    clocks = sampler.stop()
    value = B * world / (ms_step * 1e-3)

    # ---- per-kernel timings for the roofline (same stream, CUDA events, same inputs)
    pack = pack_constraints(A)
    ms_scan = timed(lambda: pack_constraints(A), max(3, args.steps // 2))
    warm = lambda: cave_forward_backward(pred, A, -1.0, mode, ratio, "mean", precision=args.precision, pack=pack)  # noqa: E731
    warm()
    ms_solve = timed(warm, max(3, args.steps // 2))
    st = cave_forward_backward(pred, A, -1.0, mode, ratio, "mean", precision=args.precision, want_status=True)
    status = st["status"].cpu().numpy()
    iters = st["iters"].cpu().numpy()
    loss_val = float(st["loss"])
    scan_bytes = float(B) * m_max * d * 4
    solve_bytes = float(sum(gen_rows)) * d * 4 + 2.0 * B * d * 4
    kernels = [
        {"name": "scan_kernel", "ms": ms_scan, "bytes": scan_bytes, "gbs": scan_bytes / ms_scan / 1e6,
         "frac_of_hbm_peak": scan_bytes / ms_scan / 1e6 / hbm_peak},
        {"name": "solve_kernel+finalize", "ms": ms_solve, "bytes": solve_bytes, "gbs": solve_bytes / ms_solve / 1e6,
         "note": "shared-memory / latency bound; reads only the general rows of A"},
    ]
    dom = max(kernels, key=lambda k: k["ms"])
    # step-level roofline: algorithmic bytes of the whole path over the whole step, against HBM
    achieved = alg_bytes / ms_step / 1e6
    # DRAM bytes per step from the ncu --set full capture of this command (profiles/r1_ncu_summary.txt):
    # scan 26.884 + 0.223 GB, plan 0.015 GB, solve 0.240 + 0.027 GB; only valid for the default workload
    traffic = 27.39e9 if (args.workload == "tsp50" and B == 4096) else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "traffic": traffic, "peak_source": peak_src, "dominant_kernel": dom["name"],
                "algorithmic_bytes_per_step": alg_bytes,
                "note": "achieved = sum_i (4 m_i d + 8 d + 4) bytes / step time; per-kernel split in `kernels`"}

    # ---- end to end through the module call with host tensors
    e2e = None
    e2e_resident = None
    if not args.no_e2e:
        class Model:
            modelSense = EPO.MINIMIZE
        mod = (innerConeAlignedCosine(Model(), solver="cuda", inner_ratio=ratio, seed=0,
                                      solver_kwargs={"precision": args.precision})
               if mode == 1 else exactConeAlignedCosine(Model(), solver="cuda", solver_kwargs={"precision": args.precision}))
        # pinned host copy of the dense constraints, bounded so that N ranks on one box never pin more than
        # ~40 % of the host's free memory between them (the leg is PCIe-bound, so the rate does not depend on Be)
        Be = B
        try:
            import psutil
            budget = 0.4 * psutil.virtual_memory().available / max(world, 1)
            Be = int(max(64, min(B, budget // (m_max * d * 4))))
        except ImportError:
            Be = B if world == 1 else min(B, 1024)
        A_host = torch.empty((Be,) + tuple(A.shape[1:]), dtype=A.dtype, pin_memory=True)
        A_host.copy_(A[:Be])
        pred_host_e = torch.empty((Be, d), dtype=pred.dtype, pin_memory=True)
        pred_host_e.copy_(pred[:Be])
        pred_host = torch.empty(pred.shape, dtype=pred.dtype, pin_memory=True)
        pred_host.copy_(pred)

        def e2e_step():
            p = pred_host_e.requires_grad_(True)
            p.grad = None
            loss = mod(p, A_host)        # H2D of pred_cost and tight_ctrs inside; loss and grad come back to host
            loss.backward()
            return loss.item()

        e2e_step()
        n_e2e = max(2, min(args.steps, 5))
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / n_e2e
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": Be * world / dt, "unit": UNIT, "h2d_bytes_per_step": int(A_host.numel() * 4 + pred_host_e.numel() * 4),
               "d2h_bytes_per_step": int(pred_host_e.numel() * 4 + 4), "ms_per_step": dt * 1e3, "steps": n_e2e,
               "batch_per_gpu": Be,
               "note": "module call with pinned host pred_cost and tight_ctrs; PCIe copy of the dense constraints dominates"
                       + ("" if Be == B else f"; batch bounded to {Be}/GPU by host memory")}
        # the same module call with the constraints kept resident on the device (CavePack over the dataset +
        # per-step instance index): what a multi-epoch trainer does, since A_i never changes between epochs
        perm_host = torch.randperm(B, dtype=torch.int32).pin_memory()

        def e2e_resident_step():
            p = pred_host.requires_grad_(True)
            p.grad = None
            loss = mod(p, pack, index=perm_host)       # H2D: pred_cost + index; D2H: loss + gradient
            loss.backward()
            return loss.item()

        e2e_resident_step()
        n_res = max(3, args.steps)
        barrier()
        t0 = time.perf_counter()
        for _ in range(n_res):
            e2e_resident_step()
        barrier()
        dtr = (time.perf_counter() - t0) / n_res
        if world > 1:
            t = torch.tensor([dtr], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dtr = float(t.item())
        e2e_resident = {"value": B * world / dtr, "unit": UNIT, "h2d_bytes_per_step": int(pred_host.numel() * 4 + B * 4),
                        "d2h_bytes_per_step": int(pred_host.numel() * 4 + 4), "ms_per_step": dtr * 1e3, "steps": n_res,
                        "note": "device-resident packed dataset (pack built once, outside the timed region) + instance "
                                "index per step; host pred_cost in, host loss and gradient out"}
        del A_host

    plan_info = pack.launch_plan(args.precision)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64" if args.precision == "fp64" else "f64 state + f32 factor", "data": "synthetic",
        "config": {"workload": f"{args.workload} DFJ synthetic (SURVEY App. B), CaVE+ inner_ratio {ratio}, batch {B}/GPU, "
                               f"pred regime {args.regime}, dense float32 [B,{m_max},{d}] resident in HBM, cold pack",
                   "batch_per_gpu": B, "m_max": m_max, "d": d, "l2_policy": "inputs larger than L2 (A = %.1f GB)" % (scan_bytes / 1e9)},
        "clocks": clocks, "e2e": e2e, "e2e_resident_dataset": e2e_resident, "gpu_launches": 8 * args.steps,
        "roofline": roofline, "kernels": kernels, "solve_launch_plan": plan_info,
        "solver": {"status_counts": {str(k): int(v) for k, v in zip(*np.unique(status, return_counts=True))},
                   "iters_mean": float(iters.mean()), "iters_max": int(iters.max()), "loss": loss_val},
    }

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = args.cpu_sample or 2 * cores
        pool = _cpu_pool(cores)
        work = _cpu_sample(kind, n, mode, args.regime, seed=1)
        pool.map(_cpu_one, work[:cores])
        t0 = time.perf_counter()
        pool.map(_cpu_one, work)
        dt = time.perf_counter() - t0
        pool.close()
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{n} instances of the same workload, multiprocessing.Pool({cores}) over "
                                          "oracle.forward_backward (scipy.optimize.nnls + epilogue), BLAS threads = 1"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
