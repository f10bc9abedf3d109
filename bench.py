#!/usr/bin/env python
"""bench.py -- RNN-T loss+grad throughput on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c3|c4|c5mb]

A "step" is one pass of the hot path over one batch of synthetic input: loss + dense gradient
w.r.t. log_probs for all lattices of the batch (what the reference's forward call produces,
pytorch_binding/benchmark.py:36-43), log_softmax excluded.

  value      lattices/s, inputs resident in HBM, CUDA events over exactly K steps, max over ranks.
             Buffers rotate over R input sets and R live outputs so that consecutive steps never
             touch the same lines (aggregate footprint >> 126 MB L2).
  e2e        same metric through the public Python API (warp_rnnt_b200.rnnt_loss + backward) with
             HOST (pinned) inputs: H2D copy of the step's inputs and a D2H read of the loss inside
             the timed region.
  roofline   algorithmic bytes per launch (4*N*T*U*V grad write + 8*N*T*U log-prob
             reads + 4*N*(U-1) + 12*N) / measured step time, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the CPU oracle port (oracle/rnnt_oracle.c, f32 flavour, OpenMP over lattices) timed on
             this box's host cores on a bounded sample of the same workload (rank 0, N=1 only).

Multi-GPU (launched by torchrun): every rank runs the same per-GPU workload on its own shard
(weak scaling, no data-path collective) and the scalar loss is all-reduced over NCCL each step.

--impl reference: the UNMODIFIED reference kernels (oracle/_ref/warp_rnnt_ref_C.so, built from
/root/reference by oracle/build_ref.py) through the reference's own operator API on the GPU, same
harness; falls back to the CPU oracle port when that extension or a GPU is not available.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np
import torch

WORKLOADS = {
    # name: (N per GPU, T, U, V, description)
    "c2": (128, 150, 40, 28, "N=128 T=150 U=40 V=28 gather=False loss+grad (BASELINE configs[1])"),
    "c3": (32, 150, 20, 5000, "N=32 T=150 U=20 V=5000 large-vocab (BASELINE configs[2])"),
    "c4": (64, 1500, 300, 50, "N=64 T=1500 U=300 V=50 long-utterance (BASELINE configs[3], dense layout)"),
    "c5mb": (32, 600, 150, 1024, "N=32 T=600 U=150 V=1024 = one micro-batch of BASELINE configs[4] (256/GPU)"),
}


def b_alg(N, T, U, V):
    return 4 * N * T * U * V + 8 * N * T * U + 4 * N * (U - 1) + 12 * N


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def make_host_inputs(N, T, U, V, seed):
    """The reference's recipe (pytorch_binding/benchmark.py:11-27): randn -> log_softmax, labels in
    [1,V), full lengths.  Pinned host tensors."""
    g = torch.Generator().manual_seed(seed)
    xs = torch.log_softmax(torch.randn((N, T, U, V), dtype=torch.float32, generator=g), dim=-1)
    ys = torch.randint(1, V, (N, U - 1), dtype=torch.int, generator=g)
    xn = torch.full((N,), T, dtype=torch.int)
    yn = torch.full((N,), U - 1, dtype=torch.int)
    pin = torch.cuda.is_available()
    return tuple(t.pin_memory() if pin else t for t in (xs, ys, xn, yn))


class ClockSampler:
    """SM clock / throttle reasons via NVML while the GPU is under load."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._t = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.nv is not None:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()

    def stop(self):
        self._stop.set()
        if self._t is not None:
            self._t.join(timeout=1.0)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def cpu_oracle_rate(N, T, U, V, budget_s=8.0, threads=None):
    """lattices/s of the CPU oracle port (f32 flavour = the reference's arithmetic) on this box."""
    from oracle import oracle
    if threads:
        oracle.set_threads(threads)
    xs, ys, xn, yn = [t.numpy() for t in make_host_inputs(N, T, U, V, seed=N)]
    oracle.dense(xs[:2], ys[:2], xn[:2], yn[:2], dtype="f32")           # warm (loads the library)
    reps, t0 = 0, time.perf_counter()
    while True:
        oracle.dense(xs, ys, xn, yn, dtype="f32")
        reps += 1
        el = time.perf_counter() - t0
        if el >= budget_s or reps >= 200:
            break
    # BASELINE configs[0]: one lattice through the numpy (awni ref_transduce style) restatement
    t1 = time.perf_counter()
    oracle.ref_transduce_np(xs[0].astype(np.float64), ys[0])
    np_ms = (time.perf_counter() - t1) * 1e3
    return {"value": N * reps / el, "unit": "lattices/s", "cores": oracle.num_threads(), "kind": "port",
            "sample": "%d x the full %s batch through oracle/rnnt_oracle.c (f32, OpenMP over lattices), %.1f s"
                      % (reps, "N=%d T=%d U=%d V=%d" % (N, T, U, V), el),
            "numpy_ref_transduce_ms_per_lattice": np_ms, "host_cpus": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--lse", default="auto", choices=["auto", "fast", "exact"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = min(steps, 20)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    N, T, U, V, desc = WORKLOADS[args.workload]
    have_cuda = torch.cuda.is_available()

    if args.impl == "reference":
        return run_reference(args, rank, world, N, T, U, V, desc, have_cuda)
    if not have_cuda:
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU oracle)")

    import torch.distributed as dist
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    import warp_rnnt_b200 as w
    w.set_lse_mode(args.lse)

    # R rotating input sets, R live outputs: no step re-touches lines of the previous ones
    per_set = 2 * 4 * N * T * U * V
    R = int(max(2, min(6, (700e6 // per_set) + 1))) if per_set < 4e9 else 1
    host = [make_host_inputs(N, T, U, V, seed=1000 * rank + N + i) for i in range(R)]
    sets = [tuple(t.to(dev, non_blocking=True) for t in h) for h in host]
    keep = [None] * R
    torch.cuda.synchronize()

    pending = []                                        # in-flight scalar all-reduces (world > 1)

    def step(i):
        s = sets[i % R]
        costs, grads = w._C.rnnt_loss(s[0], s[1], s[2], s[3])
        keep[i % R] = (costs, grads)
        if world > 1:
            # the one collective of the sharded path: 1 float, issued asynchronously so that step i's
            # all-reduce (NCCL stream) overlaps step i+1's kernel; the value is only consumed one step
            # later (logging / optimizer in a real loop), which is when we wait on it
            loss = costs.sum()
            pending.append((dist.all_reduce(loss, async_op=True), loss))
            if len(pending) > 1:
                pending.pop(0)[0].wait()
        return costs

    def drain():
        while pending:
            pending.pop(0)[0].wait()

    sampler = ClockSampler(local_rank)
    sampler.start()
    # prime the caching allocator: R live outputs + the one being produced must all exist before
    # the timed region, or a cudaMalloc (~1 ms) lands inside it
    for i in range(R + 2):
        step(i)
    for i in range(args.warmup):
        step(i)
    drain()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches0 = w._C.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(args.warmup + i)
    drain()                                             # every all-reduce of the K steps has completed
    e1.record()
    torch.cuda.synchronize()
    launches = w._C.launch_count() - launches0
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    ms_per_step = ms / args.steps
    value = N * world * args.steps / (ms * 1e-3)

    # the opt-in short LSE chain, same protocol (extra information; the headline is the default mode)
    fast_ms = None
    if world == 1 and args.lse == "auto":
        w.set_lse_mode("fast")
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for i in range(args.steps):
            step(3 + i)
        f1.record()
        torch.cuda.synchronize()
        fast_ms = f0.elapsed_time(f1) / args.steps
        w.set_lse_mode("auto")

    # ---- end to end through the public API, host buffers
    ke = args.e2e_steps or min(args.steps, 20)
    hb = sum(t.numel() * t.element_size() for t in host[0])

    def e2e_step(i):
        h = host[i % R]
        x = h[0].to(dev, non_blocking=True).requires_grad_(True)
        ys, xn, yn = (t.to(dev, non_blocking=True) for t in h[1:])
        if world > 1:
            from warp_rnnt_b200.parallel import rnnt_loss_sharded
            loss = rnnt_loss_sharded(x, ys, xn, yn, reduction="sum")
        else:
            loss = w.rnnt_loss(x, ys, xn, yn, reduction="sum")
        loss.backward()
        keep[i % R] = x.grad
        return float(loss.item())                       # D2H read of the step's result

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(ke):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    clocks = sampler.stop()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    balg = b_alg(N, T, U, V)
    achieved = balg / (ms_per_step * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "r1_summary.json")
    if os.path.exists(prof):
        try:
            traffic = json.load(open(prof)).get(args.workload, {}).get("dram_bytes_per_launch")
        except Exception:
            traffic = None
    out = {
        "metric": "RNN-T loss+grad lattices/sec", "value": value, "unit": "lattices/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "%s: %s" % (args.workload, desc), "lattices_per_gpu": N, "T": T, "U": U, "V": V,
                   "global_batch": N * world, "parallelism": "batch-sharded x%d, scalar-loss all-reduce" % world,
                   "lse_mode": args.lse + (" (= exact: results bit-identical to the reference kernels)" if args.lse == "auto" else ""),
                   "l2_protocol": "%d rotating input sets + %d live outputs (%.0f MB) > L2" % (R, R, R * per_set / 1e6),
                   "timed_call": "_C.rnnt_loss (loss + dense grads, one fused kernel)"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "algorithmic_bytes_per_launch": balg, "peak_source": peak_src,
                     "kernel": "k_fused" if args.workload in ("c2", "c3") else "k_gather+k_wavefront+k_expand"},
        "e2e": {"value": N * world * ke / e2e_s, "unit": "lattices/s", "h2d_bytes_per_step": hb,
                "d2h_bytes_per_step": 4, "steps": ke, "ms_per_step": e2e_s / ke * 1e3,
                "call": "warp_rnnt_b200.rnnt_loss(reduction='sum') + backward, pinned host inputs"},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if fast_ms is not None:
        out["lse_fast"] = {"ms_per_step": fast_ms, "value": N / (fast_ms * 1e-3), "unit": "lattices/s",
                           "roofline_frac": balg / (fast_ms * 1e-3) / 1e9 / peak,
                           "note": "opt-in RNNT_LSE_FAST (fp32-noise-level deviation from the reference, <= ~1e-4 on gradients); "
                                   "the headline value uses the default mode, which is bit-identical to the reference"}
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_oracle_rate(N, T, U, V)
    print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


class _RefRNNTLoss(torch.autograd.Function):
    """The reference's own autograd wrapper, re-stated because its module needs installed dist
    metadata (pytorch_binding/warp_rnnt/__init__.py:4-6, :9-24): grads in forward, mul_ in backward."""

    @staticmethod
    def forward(ctx, core, log_probs, labels, frames_lengths, labels_lengths):
        costs, ctx.grads = core.rnnt_loss(xs=log_probs, ys=labels, xn=frames_lengths, yn=labels_lengths,
                                          blank=0, fastemit_lambda=0.0)
        return costs

    @staticmethod
    def backward(ctx, grads_output):
        grads_output = grads_output.view(-1, 1, 1, 1).to(ctx.grads)
        return None, ctx.grads.mul_(grads_output), None, None, None


def run_reference(args, rank, world, N, T, U, V, desc, have_cuda):
    if rank != 0:
        return                                          # rank 0 alone runs the reference arm
    ref = None
    if have_cuda:
        try:
            from oracle import build_ref
            ref = build_ref.load()
        except Exception:
            ref = None
    cpu = None if args.no_cpu_baseline and ref is not None else cpu_oracle_rate(N, T, U, V)
    base = {"impl": "reference", "metric": "RNN-T loss+grad lattices/sec", "unit": "lattices/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "%s: %s" % (args.workload, desc), "lattices_per_gpu": N, "T": T, "U": U, "V": V,
                       "global_batch": N, "parallelism": "none (the reference is single-GPU; rank 0 only)"}}
    if cpu is not None:
        base["cpu_baseline"] = cpu
    if ref is None:
        # CPU arm: the oracle port on the host cores (no GPU or the reference extension did not travel)
        base.update({"value": cpu["value"], "ms_per_step": 1e3 * N / cpu["value"], "gpu_launches": 0,
                     "e2e": {"value": cpu["value"], "unit": "lattices/s", "h2d_bytes_per_step": 0,
                             "d2h_bytes_per_step": 0},
                     "reference_kind": "cpu oracle port (oracle/_ref not loadable)"})
        print(json.dumps(base), flush=True)
        return
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    per_set = 2 * 4 * N * T * U * V
    R = int(max(2, min(6, (700e6 // per_set) + 1))) if per_set < 4e9 else 1
    host = [make_host_inputs(N, T, U, V, seed=N + i) for i in range(R)]
    sets = [tuple(t.to(dev) for t in h) for h in host]
    keep = [None] * R
    steps = args.steps

    def step(i):
        s = sets[i % R]
        keep[i % R] = ref.rnnt_loss(s[0], s[1], s[2], s[3])

    sampler = ClockSampler(dev.index or 0)
    sampler.start()
    for i in range(R + 2):                              # prime the caching allocator (see main())
        step(i)
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(args.warmup + i)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    ke = args.e2e_steps or min(steps, 20)

    def e2e_step(i):
        h = host[i % R]
        x = h[0].to(dev, non_blocking=True).requires_grad_(True)
        ys, xn, yn = (t.to(dev, non_blocking=True) for t in h[1:])
        loss = _RefRNNTLoss.apply(ref, x, ys, xn, yn).sum()
        loss.backward()
        keep[i % R] = x.grad
        return float(loss.item())

    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(ke):
        e2e_step(i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    hb = sum(t.numel() * t.element_size() for t in host[0])
    peak, peak_src = peaks()
    balg = b_alg(N, T, U, V)
    ms_per_step = ms / steps
    base.update({"value": N * steps / (ms * 1e-3), "ms_per_step": ms_per_step,
                 "roofline": {"bound": "hbm", "achieved": balg / (ms_per_step * 1e-3) / 1e9, "peak": peak,
                              "unit": "GB/s", "frac": balg / (ms_per_step * 1e-3) / 1e9 / peak, "traffic": None,
                              "peak_source": peak_src},
                 "e2e": {"value": N * ke / e2e_s, "unit": "lattices/s", "h2d_bytes_per_step": hb,
                         "d2h_bytes_per_step": 4, "steps": ke, "ms_per_step": e2e_s / ke * 1e3},
                 "gpu_launches": None, "clocks": sampler.stop(),
                 "reference_kind": "unmodified reference kernels (oracle/_ref/warp_rnnt_ref_C.so) on the same B200"})
    print(json.dumps(base), flush=True)


if __name__ == "__main__":
    main()
