#!/usr/bin/env python3
"""Benchmark of the RWM / PT-RWM sampling hot path (contract: see the task statement; metric: BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c4|c4u|c5|c5f]

One "step" = one pass of the hot path over one batch: a single launch of the persistent fused kernel that advances
every chain of the workload by `T` Metropolis steps (plus swap sweeps).  Default workload = BASELINE config 3, the
one the north-star target is quoted on: PT-RWM, RoughCarpet d=20 (modes +-5, w .5/.3/.2), var 0.9, geometric ladder of
8 temperatures, swap_every 10, burn-in 2000, 1024 independent ladders (8192 chains) per GPU; ladders shard across
ranks with no data-path collective (weak scaling: 1024 ladders per GPU), NCCL only all-reduces the accumulators.

value      chain-steps/s, whole job, inputs resident in HBM, device-timed with CUDA events (max over ranks)
e2e        same metric through the C-ABI host-buffer entry rwmpt_run_host: pinned HOST buffers in, H2D + kernel + D2H
roofline   SFU (transcendental) issue roofline of SURVEY.md section 8(d): S = 121 transcendentals per chain-step (C3)
cpu_baseline  the NumPy oracle port timed on this box's host cores on a bounded sample (reported, not the target)
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# algorithmic work per chain-step: F fp32 flops, S transcendentals, stored bytes
WORKLOADS = {
    "c3": dict(desc="C3 PT-RWM RoughCarpet d=20 K=8 swap_every=10 Normal var=0.9, 1024 ladders/GPU, accumulators only",
               kind="pt", target="rough_carpet", dim=20, K=8, units=1024, T=500_000, burn_in=2000, swap_every=10,
               var=0.9, F=644, S=121, bytes=0),
    "c2": dict(desc="C2 RWM EvenRosenbrock d=20 Normal var=0.297436^2/20, 4096 chains/GPU, accumulators only",
               kind="rwm", target="even_rosenbrock", dim=20, K=1, units=4096, T=1_000_000, burn_in=1000, swap_every=1,
               var=0.297436 ** 2 / 20, F=294, S=41, bytes=0),
    "c4": dict(desc="C4 PT-RWM ThreeMixture d=50 (+-15) K=8 Laplace var_i=2.38^2/50, 512 ladders/GPU, all chains stored",
               kind="pt", target="three_mixture", dim=50, K=8, units=512, T=2_000, burn_in=0, swap_every=10,
               var=2.38 ** 2 / 50, F=1066, S=55, bytes=204, proposal="laplace", store="all"),
    "c4u": dict(desc="C4 PT-RWM ThreeMixture d=50 (+-15) K=8 UniformRadius r=1, 512 ladders/GPU, all chains stored",
                kind="pt", target="three_mixture", dim=50, K=8, units=512, T=2_000, burn_in=0, swap_every=10,
                var=1.0, F=1122, S=109, bytes=204, proposal="uniform", store="all"),
    "c5": dict(desc="C5 RWM FullRosenbrock d=100, 64 variances x 256 chains (16384 chains/GPU), accumulators only",
               kind="rwm", target="full_rosenbrock", dim=100, K=1, units=16384, T=20_000, burn_in=1000, swap_every=1,
               var=None, F=1895, S=201, bytes=0),
    "c5f": dict(desc="C5 RWM NealFunnel d=100, 64 variances x 256 chains (16384 chains/GPU), accumulators only",
                kind="rwm", target="neal_funnel", dim=100, K=1, units=16384, T=20_000, burn_in=1000, swap_every=1,
                var=None, F=1315, S=202, bytes=0),
}
# dram__bytes_read.sum + dram__bytes_write.sum of the hot kernel for one launch of the workload at its default run length,
# from the committed `ncu --set full` captures (profiles/): (bytes, summary file)
NCU_TRAFFIC = {
    "c3": (1_166_592 + 42_752, "profiles/r1j_c3_mcmc_kernel_ncu_full_summary.txt"),
    "c4": (3_691_008 + 1_618_394_000, "profiles/r1j_c4_mcmc_kernel_ncu_full_summary.txt"),
    "c5": (7_150_592 + 0, "profiles/r1g_c5_mcmc_kernel_ncu_full_summary.txt"),
}
NOMINAL_SFU_GOPS = 148 * 16 * 1.965     # 16 SFU lanes / SM / clk
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e-3


def make_target(name, dim):
    import rwm_pt_pytorch_b200.target_distributions as td
    if name == "rough_carpet":
        return td.RoughCarpetDistributionTorch(dim, device="cpu")
    if name == "even_rosenbrock":
        return td.EvenRosenbrockTorch(dim, device="cpu")
    if name == "full_rosenbrock":
        return td.FullRosenbrockTorch(dim, device="cpu")
    if name == "neal_funnel":
        return td.NealFunnelTorch(dim, device="cpu")
    if name == "three_mixture":
        return td.ThreeMixtureDistributionTorch(dim, device="cpu", mode_centers=[[-15.0] + [0.0] * (dim - 1), [0.0] * dim,
                                                                                [15.0] + [0.0] * (dim - 1)])
    raise ValueError(name)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"]
            if self.index is not None:
                cmd += ["-i", str(self.index)]
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=10)   # NVML polling stalls cudaMalloc / cudaFree of later phases: make sure it is gone
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows if len(r) >= 9 for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference's path on the host cores
# ---------------------------------------------------------------------------------------------------------
def _cpu_worker(job):
    wl_name, n_units, T, seed = job
    from oracle import rwmpt_oracle as O
    wl = WORKLOADS[wl_name]
    d, K = wl["dim"], wl["K"]
    spec = make_target(wl["target"], d).spec()
    rs = np.random.RandomState(seed)
    t0 = time.perf_counter()
    if wl["kind"] == "pt":
        betas = O.geometric_ladder()
        std = np.sqrt(np.asarray([np.float32(wl["var"] / b) for b in betas], np.float32))
        inc = (rs.randn(T, n_units, K, d).astype(np.float32) * std[None, None, :, None]).astype(np.float32)
        R = T // wl["swap_every"]
        O.pt_run(spec, np.zeros((n_units, K, d), np.float32), betas, inc, rs.rand(T, n_units, K), rs.rand(R + 1, n_units, K - 1),
                 wl["swap_every"], burn_in=0, keep_states=False)
    else:
        var = wl["var"] or 0.34 ** 2 / d
        inc = O.normal_increments(rs.randn(T, n_units, d), var, 1.0)
        O.rwm_run(spec, 1e-8 * rs.randn(n_units, d).astype(np.float32), 1.0, inc, rs.rand(T, n_units), burn_in=0, keep_states=False)
    return n_units * K * T, time.perf_counter() - t0


def cpu_arm(wl_name, procs, units_per_proc, T):
    """chain-steps/s of the oracle port: `procs` independent processes (chains are embarrassingly parallel)."""
    import multiprocessing as mp
    jobs = [(wl_name, units_per_proc, T, 100 + i) for i in range(procs)]
    t0 = time.perf_counter()
    if procs == 1:
        res = [_cpu_worker(jobs[0])]
    else:
        with mp.get_context("spawn").Pool(procs) as pool:
            res = pool.map(_cpu_worker, jobs)
    wall = time.perf_counter() - t0
    total = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return total / busy, wall


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    units, T = (32, 2000) if wl["kind"] == "pt" else (1024, 2000)
    for _ in range(max(args.warmup, 0) and 1):
        cpu_arm(args.workload, procs, units, max(T // 10, 50))
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_arm(args.workload, procs, units, T)
        vals.append(v)
    wall = time.perf_counter() - t_all
    value = float(np.mean(vals))
    sample = (f"NumPy oracle port (oracle/rwmpt_oracle.py) of the reference's path, {procs} processes x {units} "
              f"{'ladders' if wl['kind'] == 'pt' else 'chains'} x {T} steps per bench step, injected NumPy randomness")
    line = {"impl": "reference", "metric": "chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "note": "reference CPU path timed on the host cores of this box"},
            "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": procs, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def build_sampler(wl, dev, rank, store, lanes, seed=2026):
    from rwm_pt_pytorch_b200.algorithms import RandomWalkMH_GPU_Optimized as RWM, ParallelTemperingRWM_GPU_Optimized as PT
    from rwm_pt_pytorch_b200.proposal_distributions import LaplaceProposal
    import torch
    t = make_target(wl["target"], wl["dim"])
    d, K, n = wl["dim"], wl["K"], wl["units"]
    if wl["kind"] == "pt":
        prop = None
        if wl.get("proposal") == "laplace":
            prop = LaplaceProposal(d, torch.full((d,), wl["var"]), 1.0, torch.device("cpu"), torch.float32)
        elif wl.get("proposal") == "uniform":
            from rwm_pt_pytorch_b200.proposal_distributions import UniformRadiusProposal
            prop = UniformRadiusProposal(d, wl["var"], 1.0, torch.device("cpu"), torch.float32)
        algo = PT(d, wl["var"], t, geom_temp_spacing=True, swap_every=wl["swap_every"], burn_in=wl["burn_in"], device=dev,
                  num_ladders=n, store=store, seed=seed, chain_id_base=rank * n * K, lanes_per_chain=lanes,
                  proposal_distribution=prop, initial_states=np.zeros((n, 1, d), np.float32))
        batch = algo._require_batch()
    else:
        var = wl["var"]
        if var is None:   # C5: 64 variance values x 256 chains
            xs = np.linspace(0.01, 1.3, 64)
            var = np.repeat(xs ** 2 / d, n // 64)
        np.random.seed(1 + rank)
        algo = RWM(d, var, t, burn_in=wl["burn_in"], device=dev, num_chains=n, store=store, seed=seed,
                   chain_id_base=rank * n, lanes_per_chain=lanes)
        algo._ensure_batch(1)
        batch = algo._batch
    return algo, batch, t


def e2e_run(wl, t, batch, dev_index, T, reps):
    """chain-steps/s through rwmpt_run_host: pinned host buffers -> H2D -> fused kernel -> D2H, wall-clock timed."""
    import torch
    from rwm_pt_pytorch_b200 import _lib
    lib = _lib.load()
    d, K, n = wl["dim"], wl["K"], wl["units"]
    nc = n * K
    pin = lambda x: x.contiguous().pin_memory()
    params = pin(t.pack())
    beta, scale = pin(batch.beta.cpu()), pin(batch.prop_scale.cpu())
    dscale = None if batch.prop_dim_scale is None else pin(batch.prop_dim_scale.cpu())
    state0 = torch.zeros((nc, d), dtype=torch.float32)
    logp0 = pin(t.log_density(state0.to(f"cuda:{dev_index}")).cpu())
    state = pin(state0.clone()); logp = pin(logp0.clone())
    acc = pin(torch.zeros(nc, dtype=torch.int64)); sq = pin(torch.zeros(nc, dtype=torch.float64))
    sacc = pin(torch.zeros((n, max(K - 1, 1)), dtype=torch.int64)); last = pin(torch.zeros(nc, dtype=torch.int64))
    a = _lib.RunArgs()
    a.target = _lib.TargetT(t.family_id, d, params.data_ptr(), params.numel())
    a.proposal_family, a.n_temps = batch.prop_family, K
    a.prop_scale, a.beta = scale.data_ptr(), beta.data_ptr()
    a.prop_dim_scale = None if dscale is None else dscale.data_ptr()
    a.n_ladders, a.n_steps, a.burn_in, a.step_offset = n, T, wl["burn_in"], 0
    a.swap_every, a.swap_mode = wl["swap_every"], 0
    a.state, a.logp = state.data_ptr(), logp.data_ptr()
    a.seed = 7
    a.chain_id_base = batch.chain_id_base
    a.accept_count, a.sq_jump_sum = acc.data_ptr(), sq.data_ptr()
    a.swap_accepts, a.swap_last_attempt = sacc.data_ptr(), last.data_ptr()
    a.lanes_per_chain = batch.lanes_per_chain
    a.schedule = batch.schedule
    h2d, d2h = C.c_uint64(), C.c_uint64()
    times = []
    batch.run(T)                       # the set-up above left the GPU idle long enough to drop its clocks: ramp them up again
    torch.cuda.synchronize()
    for i in range(reps + 1):
        state.copy_(state0); logp.copy_(logp0); acc.zero_(); sq.zero_(); sacc.zero_(); last.zero_()
        t0 = time.perf_counter()
        _lib.check(lib.rwmpt_run_host(C.byref(a), dev_index, C.byref(h2d), C.byref(d2h)))
        dt = time.perf_counter() - t0
        if i > 0:
            times.append(dt)
        if os.environ.get("RWMPT_BENCH_DEBUG"):
            print(f"e2e call {i}: {dt * 1e3:.2f} ms", file=sys.stderr)
    accept = float(acc.sum().item()) / max(nc * (T - wl["burn_in"]), 1)
    # median of the calls: a concurrent NVML query or the allocator occasionally stalls one call for hundreds of ms; every
    # call's time is reported next to it (e2e.calls_ms)
    return float(np.median(times)), int(h2d.value), int(d2h.value), accept, [round(x * 1e3, 2) for x in times]


def aux_kernel_rates(dev, hbm_peak):
    """The path's second hand-written kernel, the ESJD reduction over stored samples (HBM-read bound: 4*n*d algorithmic
    bytes per chain), timed alone with CUDA events on a buffer several times larger than L2."""
    import torch
    from rwm_pt_pytorch_b200 import _lib
    lib = _lib.load()
    out = {}
    for name, (B, S, d) in {"esjd_reduce d=50": (4096, 1025, 50), "esjd_reduce d=20": (8192, 2049, 20)}.items():
        x = torch.randn((B, S, d), device=dev, dtype=torch.float32)
        res = torch.empty(B, device=dev, dtype=torch.float64)
        moved = torch.empty(B, device=dev, dtype=torch.int64)
        row = {}
        for form, mv in (("flat", None), ("with_moved_count", moved.data_ptr())):
            times = []
            for i in range(7):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.check(lib.rwmpt_esjd_reduce(x.data_ptr(), B, S, 0, S, d, res.data_ptr(), mv, _lib.stream_ptr(dev)))
                e1.record()
                torch.cuda.synchronize()
                if i >= 2:
                    times.append(e0.elapsed_time(e1))
            gbs = x.numel() * 4 / (float(np.mean(times)) * 1e-3) / 1e9
            row[form] = {"ms": float(np.mean(times)), "GB/s": gbs, "frac_of_hbm_peak": gbs / hbm_peak}
        row["bytes"] = x.numel() * 4
        out[name] = row
        del x
    # the stand-alone plugin kernels behind target.log_density(x) and proposal.sample(n) (the iterative
    # ladder construction evaluates up to 1e6 densities per estimate)
    import rwm_pt_pytorch_b200.target_distributions as td
    n, d = 4_000_000, 20
    t = td.RoughCarpetDistributionTorch(d, device="cpu")
    params = t.device_params(dev)
    tgt = _lib.target_struct(t.family_id, d, params)
    x = torch.randn((n, d), device=dev, dtype=torch.float32) * 4
    res = torch.empty(n, device=dev, dtype=torch.float32)

    def timed(fn):
        times = []
        for i in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            if i >= 2:
                times.append(e0.elapsed_time(e1))
        return float(np.mean(times))

    ms = timed(lambda: _lib.check(lib.rwmpt_log_density(tgt, x.data_ptr(), n, res.data_ptr(), 0, _lib.stream_ptr(dev))))
    out["log_density RoughCarpet d=20"] = {"ms": ms, "rows_per_s": n / (ms * 1e-3), "GB/s": n * (d + 1) * 4 / (ms * 1e-3) / 1e9,
                                            "frac_of_hbm_peak": n * (d + 1) * 4 / (ms * 1e-3) / 1e9 / hbm_peak}
    for fam, name in ((0, "normal"), (1, "laplace"), (2, "uniform_radius")):
        ms = timed(lambda: _lib.check(lib.rwmpt_proposal_sample(fam, d, 0.5, None, n, 7, 0, x.data_ptr(), _lib.stream_ptr(dev))))
        out[f"proposal_sample {name} d=20"] = {"ms": ms, "rows_per_s": n / (ms * 1e-3), "GB/s": n * d * 4 / (ms * 1e-3) / 1e9,
                                                "frac_of_hbm_peak": n * d * 4 / (ms * 1e-3) / 1e9 / hbm_peak}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--T", type=int, default=0, help="Metropolis steps per launch (0 = workload default)")
    ap.add_argument("--units", type=int, default=0, help="ladders / chains per GPU (0 = workload default)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per chain (0 = auto)")
    ap.add_argument("--store", default="", choices=["", "none", "cold", "all"], help="override the workload's trajectory storage")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--also", default="", help="comma list of extra workloads reported under 'also' (rank 0, N=1)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    from rwm_pt_pytorch_b200 import _lib
    from rwm_pt_pytorch_b200 import distributed as D

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the sampling path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    wl = dict(WORKLOADS[args.workload])
    if args.T:
        wl["T"] = args.T
    if args.units:
        wl["units"] = args.units
    store = args.store or wl.get("store", "none")
    if store != "none" and "store" not in wl:
        wl["bytes"] = 4 * wl["dim"] + 4 if store == "all" else (4 * wl["dim"] + 4) / wl["K"]

    def measure(wl, store, with_clocks):
        algo, batch, t = build_sampler(wl, dev, rank, store, args.lanes)
        T, nc = wl["T"], wl["units"] * wl["K"]
        if store != "none":
            batch.allocate_storage(store, T * (args.steps + args.warmup) + 2, 1, with_logp=True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
        reduced = None
        for _ in range(args.warmup):
            batch.run(T)
            if world > 1:                                                 # also warms the NCCL communicator up
                D.allreduce_statistics_tensor(D.local_statistics_tensor(algo))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        # one nvidia-smi poller per job (rank 0; it lists every GPU when the job has several): NVML queries take a driver
        # lock that CUDA calls of the same process tree can wait on
        sampler = ClockSampler(local if world == 1 else None) if (with_clocks and rank == 0) else None
        if sampler:
            sampler.start()
        evs = []
        launches = 0
        for _ in range(args.steps):
            flush.fill_(1)                                                # untimed L2 flush between timed steps
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            batch.run(T)                                                  # ONE launch of the fused kernel
            launches += 1
            if world > 1:                                                 # the path's only collective: accumulators
                reduced = D.allreduce_statistics_tensor(D.local_statistics_tensor(algo))   # device ops + NCCL, no host sync
            e1.record()
            evs.append((e0, e1))
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        step_ms = [a.elapsed_time(b) for a, b in evs]
        ms = sum(step_ms)
        if reduced is not None:
            reduced = D.statistics_from_tensor(reduced)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        ms_per_step = ms / args.steps
        rate = world * nc * T / (ms_per_step * 1e-3)
        algo._refresh_stats()
        return dict(rate=rate, ms_per_step=ms_per_step, launches=launches, clocks=clocks, algo=algo, batch=batch, t=t,
                    reduced=reduced, step_ms=[round(x, 3) for x in step_ms])

    m = measure(wl, store, True)
    algo, batch = m["algo"], m["batch"]
    e2e = None
    if not args.no_e2e and store == "none":
        # every rank runs its shard through the host-buffer C-ABI entry at the same time; the job's rate uses the slowest
        if world > 1:
            dist.barrier()
        e_time, h2d, d2h, e_acc, e_calls = e2e_run(wl, m["t"], batch, local, wl["T"], reps=max(3, min(args.steps, 5)))
        if world > 1:
            tt = torch.tensor([e_time], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_time = float(tt.item())
        e2e = {"value": world * wl["units"] * wl["K"] * wl["T"] / e_time, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "acceptance_rate": e_acc,
               "calls_ms": e_calls,
               "how": "rwmpt_run_host (C ABI, pinned host buffers): H2D + one fused launch + D2H per rank, wall clock, "
                      "median of the calls on each rank (calls_ms: rank 0), max over ranks"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp32_tf, sfu_g = C.c_double(), C.c_double()
    _lib.check(_lib.load().rwmpt_probe_peaks(C.byref(fp32_tf), C.byref(sfu_g)))
    per_gpu = m["rate"] / world
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    t_sfu = wl["S"] / (sfu_g.value * 1e9)
    t_fp = wl["F"] / (fp32_tf.value * 1e12)
    t_hbm = wl["bytes"] / (hbm_peak * 1e9)
    bound = max((t_sfu, "sfu"), (t_fp, "fp32"), (t_hbm, "hbm"))[1]
    if bound == "hbm":
        roof = {"bound": "hbm", "achieved": per_gpu * wl["bytes"] / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s"}
    elif bound == "sfu":
        roof = {"bound": "sfu", "achieved": per_gpu * wl["S"] / 1e9, "peak": sfu_g.value, "unit": "Gtranscendental/s",
                "peak_source": "measured on this GPU (rwmpt_probe_peaks, dependent-free MUFU.EX2 loop); nominal %.0f" % NOMINAL_SFU_GOPS}
    else:
        roof = {"bound": "fp32", "achieved": per_gpu * wl["F"] / 1e12, "peak": fp32_tf.value, "unit": "TFLOP/s",
                "peak_source": "measured on this GPU (rwmpt_probe_peaks, dependent-free FFMA loop); nominal %.1f" % NOMINAL_FP32_TFLOPS}
    roof["frac"] = roof["achieved"] / roof["peak"]
    default_shape = wl["T"] == WORKLOADS[args.workload]["T"] and wl["units"] == WORKLOADS[args.workload]["units"] and \
        store == WORKLOADS[args.workload].get("store", "none")
    tr = NCU_TRAFFIC.get(args.workload) if default_shape else None
    roof["traffic"] = tr[0] if tr else None          # bytes per launch (ncu), next to the algorithmic bytes per launch
    roof["traffic_source"] = tr[1] if tr else None
    roof["algorithmic_bytes_per_launch"] = wl["bytes"] * wl["units"] * wl["K"] * wl["T"]
    roof["per_chain_step"] = {"F": wl["F"], "S": wl["S"], "stored_bytes": wl["bytes"]}
    roof["frac_sfu_measured"] = per_gpu * wl["S"] / (sfu_g.value * 1e9)
    roof["frac_sfu_nominal"] = per_gpu * wl["S"] / (NOMINAL_SFU_GOPS * 1e9)
    roof["frac_fp32_measured"] = per_gpu * wl["F"] / (fp32_tf.value * 1e12)
    roof["measured_peaks"] = {"fp32_tflops": fp32_tf.value, "sfu_gops": sfu_g.value, "hbm_gbs": hbm_peak}

    post = batch.post_burn_in_steps()
    esjd = float(algo.expected_squared_jump_distance_gpu()) if post > 0 else None
    line = {
        "metric": "chain-steps/sec", "value": m["rate"], "unit": "chain-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"], "steps_per_launch": wl["T"], "chains_per_gpu": wl["units"] * wl["K"],
                   "lanes_per_chain": args.lanes or "auto", "l2": "256 MiB flush write between timed steps",
                   "rng": "in-kernel Philox4x32-10", "math": "fast"},
        "gpu_launches": m["launches"], "step_ms": m["step_ms"], "clocks": m["clocks"], "roofline": roof,
        "esjd": esjd, "esjd_per_sec": None if esjd is None else esjd * m["rate"] / wl["K"],
        "acceptance_rate": float(batch.accept_count.sum().item()) / max(post * batch.n_chains, 1),
    }
    if m["reduced"] is not None:
        line["all_ranks"] = D.pooled_summary(m["reduced"], m["reduced"]["chain_steps"] / wl["K"])
    if wl["kind"] == "pt":
        line["swap_acceptance_rate"] = algo.swap_acceptance_rate
        line["ladder_steps_per_sec"] = m["rate"] / wl["K"]
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu and world == 1:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        units, Tc = (128, 3000) if wl["kind"] == "pt" else (1024, 8000)
        v, wall = cpu_arm(args.workload, 1, units, Tc)
        line["cpu_baseline"] = {"value": v, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                                "sample": f"NumPy oracle port, 1 process, {units} {'ladders' if wl['kind'] == 'pt' else 'chains'} x {Tc} steps "
                                          f"({wall:.1f} s wall); all-core figure: bench.py --impl reference"}
    if world == 1 and args.workload == "c3" and not args.no_e2e:
        line["aux_kernels"] = aux_kernel_rates(dev, hbm_peak)
    if args.also and world == 1:
        line["also"] = {}
        for name in args.also.split(","):
            w2 = dict(WORKLOADS[name])
            m2 = measure(w2, w2.get("store", "none"), False)
            line["also"][name] = {"workload": w2["desc"], "value": m2["rate"], "unit": "chain-steps/s", "ms_per_step": m2["ms_per_step"],
                                  "frac_sfu_measured": m2["rate"] * w2["S"] / (sfu_g.value * 1e9),
                                  "frac_fp32_measured": m2["rate"] * w2["F"] / (fp32_tf.value * 1e12),
                                  "hbm_write_gbs": m2["rate"] * w2["bytes"] / 1e9}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
