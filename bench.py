#!/usr/bin/env python3
"""Benchmark of the RWM / PT-RWM sampling hot path (contract: see the task statement; metric: BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c3|c2|c2d10|c2d30|c4|c4u|c5|c5f]
                    [--scaling weak|strong] [--also default|none|<comma list>]

One "step" = one pass of the hot path over one batch: a single launch of the persistent fused kernel that advances
every chain of the workload by `T` Metropolis steps (plus swap sweeps).  Default workload = BASELINE config 3, the
one the north-star target is quoted on: PT-RWM, RoughCarpet d=20 (modes +-5, w .5/.3/.2), var 0.9, geometric ladder of
8 temperatures, swap_every 10, burn-in 2000, 1024 independent ladders (8192 chains) per GPU; ladders shard across
ranks with no data-path collective (weak scaling: 1024 ladders per GPU; `--scaling strong`: 1024 ladders in total), NCCL
only all-reduces the accumulators and gathers retained samples.

value         chain-steps/s, whole job, inputs resident in HBM, device-timed with CUDA events (max over ranks)
e2e           same metric through the C-ABI host-buffer entry rwmpt_run_host: pinned HOST buffers in, H2D + kernel + D2H
roofline      SFU (transcendental) issue roofline of SURVEY.md section 8(d) when nothing is stored (C3: S = 121 per
              chain-step), HBM-write roofline when trajectories are (C4: 204 B per stored chain-step)
also          the other BASELINE configurations (C2 d=10/20/30, C4 Laplace / UniformRadius, C5 both targets), each with
              its own roofline block (N = 1 only)
cpu_baseline / reference_numpy / reference_torch
              the UNMODIFIED reference (pip-installed copy under baseline/_ref) timed on this box in the same run: its
              NumPy samplers (algorithms/rwm.py, pt_rwm.py) on one core and on every core, and its PyTorch classes on
              device='cuda' and 'cpu'; plus the NumPy oracle port (oracle/rwmpt_oracle.py).  Reported baselines, not targets.
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
    "c2d10": dict(desc="C2 RWM EvenRosenbrock d=10 Normal var=0.161282^2/10, 4096 chains/GPU, accumulators only",
                  kind="rwm", target="even_rosenbrock", dim=10, K=1, units=4096, T=1_000_000, burn_in=1000, swap_every=1,
                  var=0.161282 ** 2 / 10, F=149, S=21, bytes=0),
    "c2d30": dict(desc="C2 RWM EvenRosenbrock d=30 Normal var=0.085641^2/30, 4096 chains/GPU, accumulators only",
                  kind="rwm", target="even_rosenbrock", dim=30, K=1, units=4096, T=1_000_000, burn_in=1000, swap_every=1,
                  var=0.085641 ** 2 / 30, F=439, S=61, bytes=0),
    "c4": dict(desc="C4 PT-RWM ThreeMixture d=50 (+-15) K=8 Laplace var_i=2.38^2/50, 512 ladders/GPU, all 8 chains stored",
               kind="pt", target="three_mixture", dim=50, K=8, units=512, T=100_000, burn_in=0, swap_every=10,
               var=2.38 ** 2 / 50, F=1066, S=55, bytes=204, proposal="laplace", store="all", e2e_T=4000),
    "c4u": dict(desc="C4 PT-RWM ThreeMixture d=50 (+-15) K=8 UniformRadius r=1, 512 ladders/GPU, all 8 chains stored",
                kind="pt", target="three_mixture", dim=50, K=8, units=512, T=100_000, burn_in=0, swap_every=10,
                var=1.0, F=1122, S=109, bytes=204, proposal="uniform", store="all", e2e_T=4000),
    "c5": dict(desc="C5 RWM FullRosenbrock d=100, 64 variances x 256 chains (16384 chains/GPU), accumulators only",
               kind="rwm", target="full_rosenbrock", dim=100, K=1, units=16384, T=20_000, burn_in=1000, swap_every=1,
               var=None, sweep=(0.01, 1.0), F=1895, S=201, bytes=0),
    "c5f": dict(desc="C5 RWM NealFunnel d=100, 64 variances x 256 chains (16384 chains/GPU), accumulators only",
                kind="rwm", target="neal_funnel", dim=100, K=1, units=16384, T=20_000, burn_in=1000, swap_every=1,
                var=None, sweep=(0.01, 2.5), F=1315, S=202, bytes=0),
}
DEFAULT_ALSO = "c2d10,c2,c2d30,c4,c4u,c5,c5f"
# dram__bytes_read.sum + dram__bytes_write.sum of the hot kernel for ONE launch at the workload's default shape, from the
# committed `ncu --set full` captures: profiles/ncu_traffic.json = {workload: {"bytes": ..., "source": "profiles/..."}}
# (a profiler counter cannot be read inside an unprofiled timed run; the file is regenerated with every capture)
NCU_TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic.json")
NOMINAL_SFU_GOPS = 148 * 16 * 1.965     # 16 SFU lanes / SM / clk
NOMINAL_FP32_TFLOPS = 148 * 128 * 2 * 1.965e-3
README_LADDER = [1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125, 0.015625, 0.01]   # the geometric ladder (pt_rwm_gpu_optimized.py:245-257)


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


def job_gpu_indices(world):
    """nvidia-smi indices of the GPUs this job runs on: LOCAL_RANK r uses CUDA device r, i.e. entry r of
    CUDA_VISIBLE_DEVICES when that is set (integer ordinals), else physical GPU r."""
    vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
    ids = [v.strip() for v in vis.split(",") if v.strip()]
    if ids and all(v.isdigit() for v in ids):
        return [int(v) for v in ids[:world]]
    return list(range(world))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of the JOB's GPUs, sampled every 200 ms while the timed region runs.  The poller is
    started BEFORE the warm-up steps and the timed region only begins once its first row has arrived: its start-up (NVML
    initialisation, up to a second on a fresh box) stalls a running kernel by 25-100 ms, which used to land in the first timed
    step now and then (step_ms [194.7, 167.5, 167.2, ...]); the steady 200 ms polling does not show.  `mark()` at the start of
    the timed region drops the rows read before it."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, indices):
        self.indices, self.rows, self.proc, self.t0 = list(indices), [], None, 0.0

    def start(self):
        try:
            cmd = ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                   "-i", ",".join(str(i) for i in self.indices)]
            self.proc = subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_ready(self, wait_s=15.0):
        t_end = time.perf_counter() + wait_s
        while self.proc is not None and not self.rows and self.proc.poll() is None and time.perf_counter() < t_end:
            time.sleep(0.02)                                              # start-up still in progress

    def mark(self):
        self.t0 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=10)   # NVML polling stalls cudaMalloc / cudaFree of later phases: make sure it is gone
        except subprocess.TimeoutExpired:
            self.proc.kill()
        rows = [r for t, r in self.rows if t >= self.t0 and len(r) >= 9 and r[0].isdigit() and int(r[0]) in self.indices]
        sm = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in rows for n, v in zip(names, r[5:9]) if v.lower().startswith("active")})
        per_gpu = {}
        for r in rows:
            if r[1].replace(".", "").isdigit():
                per_gpu.setdefault(int(r[0]), []).append(float(r[1]))
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm), "gpus": self.indices,
                "sm_mhz_per_gpu": {str(k): float(np.median(v)) for k, v in sorted(per_gpu.items())}}


# ---------------------------------------------------------------------------------------------------------
# CPU baselines 1: the NumPy oracle port of the reference's path (kind "port")
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


# ---------------------------------------------------------------------------------------------------------
# CPU / PyTorch baselines 2: the UNMODIFIED reference (kind "reference").  Every leg runs in its own process
# (`bench.py --ref-leg ...`), so the reference's top-level package names never meet this repo's modules.
# ---------------------------------------------------------------------------------------------------------
def reference_root():
    """Where the unmodified reference can be imported from: the pip-installed copy under baseline/_ref (it travels to the
    GPU box with the snapshot) or, in the authoring container, /root/reference itself."""
    for p in (os.path.join(ROOT, "baseline", "_ref"), os.environ.get("RWMPT_REFERENCE", "/root/reference")):
        if p and os.path.isdir(os.path.join(p, "algorithms")) and os.path.isdir(os.path.join(p, "target_distributions")):
            return p
    return None


def _ref_leg_main(leg, n, seed):
    """Runs inside the child process: imports the reference as-is (stub matplotlib: it is imported at module scope but
    never used on this path) and times one of its own samplers.  Prints one JSON dict."""
    import types
    root = reference_root()
    if root is None:
        print(json.dumps({"unavailable": "reference not installed (baseline/_ref missing)"}))
        return
    for name in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.path[:] = [root] + [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    import contextlib
    import io
    import torch
    torch.set_num_threads(1)
    out = {"leg": leg, "steps": n, "reference_root": os.path.relpath(root, ROOT) if root.startswith(ROOT) else root}
    quiet = contextlib.redirect_stdout(io.StringIO())
    if leg == "numpy_rwm":
        # BASELINE config 1 exactly: algorithms/rwm.py on the NumPy RoughCarpet (+-15), var 2.38^2/20, seed 1 after construction
        from algorithms.rwm import RandomWalkMH
        import target_distributions as td
        algo = RandomWalkMH(20, 2.38 ** 2 / 20, td.RoughCarpetDistribution(20))
        np.random.seed(seed)
        t0 = time.perf_counter()
        for _ in range(n):
            algo.step()
        dt = time.perf_counter() - t0
        chain = np.asarray(algo.chain)
        out.update(chain_steps=n, seconds=dt, acceptance_rate=algo.acceptance_rate,
                   esjd=float(np.mean(np.sum((chain[1:] - chain[:-1]) ** 2, axis=1))), api="algorithms.rwm.RandomWalkMH.step")
    elif leg == "numpy_pt":
        # config 3's target and ladder on the reference's NumPy PT sampler (algorithms/pt_rwm.py; its swap_every is fixed at 20)
        from algorithms.pt_rwm import ParallelTemperingRWM
        import target_distributions as td
        tgt = td.RoughCarpetDistribution(20)
        tgt.modes = [-5, 0, 5]
        algo = ParallelTemperingRWM(20, 0.9, tgt, beta_ladder=list(README_LADDER))
        np.random.seed(seed)
        t0 = time.perf_counter()
        for _ in range(n):
            algo.step()
        dt = time.perf_counter() - t0
        out.update(chain_steps=n * len(README_LADDER), seconds=dt, swap_acceptance_rate=algo.acceptance_rate,
                   api="algorithms.pt_rwm.ParallelTemperingRWM.step")
    elif leg in ("torch_pt_cpu", "torch_pt_cuda"):
        from algorithms.pt_rwm_gpu_optimized import ParallelTemperingRWM_GPU_Optimized
        import target_distributions as td
        dev = "cuda" if leg.endswith("cuda") else "cpu"
        with quiet:
            tgt = td.RoughCarpetDistributionTorch(20, device=dev)
            algo = ParallelTemperingRWM_GPU_Optimized(20, 0.9, tgt, geom_temp_spacing=True, swap_every=10, burn_in=0, device=dev,
                                                      pre_allocate_steps=n)
            torch.manual_seed(seed)
            t0 = time.perf_counter()
            algo.generate_samples(n)
            if dev == "cuda":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out.update(chain_steps=n * algo.num_chains, seconds=dt, swap_acceptance_rate=float(algo.swap_acceptance_rate),
                   api="algorithms.pt_rwm_gpu_optimized.ParallelTemperingRWM_GPU_Optimized.generate_samples", device=dev)
    elif leg in ("torch_rwm_cpu", "torch_rwm_cuda"):
        from algorithms.rwm_gpu_optimized import RandomWalkMH_GPU_Optimized
        import target_distributions as td
        dev = "cuda" if leg.endswith("cuda") else "cpu"
        with quiet:
            tgt = td.EvenRosenbrockTorch(20, device=dev)
            algo = RandomWalkMH_GPU_Optimized(20, 0.297436 ** 2 / 20, tgt, burn_in=0, device=dev, pre_allocate_steps=n)
            torch.manual_seed(seed)
            t0 = time.perf_counter()
            algo.generate_samples(n)
            if dev == "cuda":
                torch.cuda.synchronize()
            dt = time.perf_counter() - t0
        out.update(chain_steps=n, seconds=dt, acceptance_rate=float(algo.acceptance_rate),
                   api="algorithms.rwm_gpu_optimized.RandomWalkMH_GPU_Optimized.generate_samples", device=dev)
    else:
        raise SystemExit(f"unknown reference leg {leg}")
    out["chain_steps_per_s"] = out["chain_steps"] / out["seconds"]
    print(json.dumps(out))


def run_ref_legs(legs):
    """Run reference legs [(leg, n, seed), ...] concurrently, one child process each; returns their JSON dicts."""
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
    for k in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT"):
        env.pop(k, None)
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), "--ref-leg", leg, "--ref-n", str(n), "--ref-seed", str(seed)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, env=env, cwd=ROOT) for leg, n, seed in legs]
    res = []
    for (leg, n, seed), p in zip(legs, procs):
        try:
            so, se = p.communicate(timeout=900)
            line = [l for l in so.splitlines() if l.startswith("{")]
            res.append(json.loads(line[-1]) if line else {"leg": leg, "error": (se or so)[-400:]})
        except subprocess.TimeoutExpired:
            p.kill()
            res.append({"leg": leg, "error": "timeout"})
    return res


def reference_numpy_all_cores(leg, n, procs):
    """`procs` independent processes of one NumPy leg at once (chains / ladders are embarrassingly parallel: this is how
    the reference itself scales out, one Slurm array task per seed); aggregate chain-steps/s = total work / slowest."""
    res = run_ref_legs([(leg, n, 1000 + i) for i in range(procs)])
    ok = [r for r in res if "chain_steps_per_s" in r]
    if not ok:
        return None, res[:1]
    total = sum(r["chain_steps"] for r in ok)
    return total / max(r["seconds"] for r in ok), ok


def reference_baselines(kind, quick=False):
    """The unmodified reference timed on this box: NumPy sampler on 1 core and on every core,
    PyTorch class on device='cpu' and device='cuda' (one chain / one ladder per object, as designed)."""
    if reference_root() is None:
        return {"unavailable": "reference not installed: baseline/_ref missing (python __graft_entry__.py builds it where "
                               "/root/reference exists)"}
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    import torch
    has_cuda = torch.cuda.is_available()
    out = {"host_cores": cores}
    if kind == "pt":
        n1, nall, ncpu, ncuda = (1500, 600, 2000, 600) if quick else (5000, 2500, 8000, 3000)
        one = run_ref_legs([("numpy_pt", n1, 1)])[0]
        allv, _ = reference_numpy_all_cores("numpy_pt", nall, procs)
        t_cpu = run_ref_legs([("torch_pt_cpu", ncpu, 1)])[0]
        t_cuda = run_ref_legs([("torch_pt_cuda", ncuda, 1)])[0] if has_cuda else {"unavailable": "no CUDA device"}
    else:
        n1, nall, ncpu, ncuda = (8000, 4000, 8000, 4000) if quick else (100_000, 20_000, 40_000, 20_000)
        one = run_ref_legs([("numpy_rwm", n1, 1)])[0]
        allv, _ = reference_numpy_all_cores("numpy_rwm", nall, procs)
        t_cpu = run_ref_legs([("torch_rwm_cpu", ncpu, 1)])[0]
        t_cuda = run_ref_legs([("torch_rwm_cuda", ncuda, 1)])[0] if has_cuda else {"unavailable": "no CUDA device"}
    out["numpy_1core"] = one
    out["numpy_all_cores"] = {"chain_steps_per_s": allv, "processes": procs, "steps_per_process": nall}
    out["torch_cpu"] = t_cpu
    out["torch_cuda"] = t_cuda
    return out


def run_reference_arm(args):
    """`--impl reference`: the reference's own CPU implementation of the path on this box's host cores -- the unmodified
    NumPy sampler (algorithms/pt_rwm.py for the PT workloads, algorithms/rwm.py for the RWM ones) from baseline/_ref, one
    process per core; the NumPy oracle port only if the reference copy is absent."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    wl = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    procs = max(1, min(cores, 64))
    use_ref = reference_root() is not None
    leg = "numpy_pt" if wl["kind"] == "pt" else "numpy_rwm"
    n_ref = 1200 if wl["kind"] == "pt" else 8000          # ~3 s of work per process and bench step
    units, T = (32, 2000) if wl["kind"] == "pt" else (1024, 2000)

    def one_step(n_scale=1.0):
        if use_ref:
            v, _ = reference_numpy_all_cores(leg, max(int(n_ref * n_scale), 50), procs)
            return v
        return cpu_arm(args.workload, procs, units, max(int(T * n_scale), 50))[0]

    for _ in range(1 if args.warmup > 0 else 0):
        one_step(0.1)
    vals = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        vals.append(one_step())
    wall = time.perf_counter() - t_all
    vals = [v for v in vals if v]
    value = float(np.mean(vals)) if vals else 0.0
    if use_ref:
        kind = "reference"
        sample = (f"UNMODIFIED reference NumPy sampler ({'algorithms/pt_rwm.py ParallelTemperingRWM, 8-temperature ladder' if leg == 'numpy_pt' else 'algorithms/rwm.py RandomWalkMH'}"
                  f", imported from baseline/_ref), {procs} processes x 1 {'ladder' if leg == 'numpy_pt' else 'chain'} x {n_ref} steps per bench step")
    else:
        kind = "port"
        sample = (f"NumPy oracle port (oracle/rwmpt_oracle.py) of the reference's path, {procs} processes x {units} "
                  f"{'ladders' if wl['kind'] == 'pt' else 'chains'} x {T} steps per bench step, injected NumPy randomness")
    line = {"impl": "reference", "metric": "chain-steps/sec", "value": value, "unit": "chain-steps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1), "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64" if use_ref else "f32", "data": "synthetic",
            "config": {"workload": wl["desc"], "note": "reference CPU path timed on the host cores of this box"},
            "cpu_baseline": {"value": value, "unit": "chain-steps/s", "cores": procs, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "chain-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------------------
def build_sampler(wl, dev, rank, store, lanes, seed=2026, first_unit=None):
    from rwm_pt_pytorch_b200.algorithms import RandomWalkMH_GPU_Optimized as RWM, ParallelTemperingRWM_GPU_Optimized as PT
    from rwm_pt_pytorch_b200.proposal_distributions import LaplaceProposal
    import torch
    t = make_target(wl["target"], wl["dim"])
    d, K, n = wl["dim"], wl["K"], wl["units"]
    first = rank * n if first_unit is None else first_unit
    if wl["kind"] == "pt":
        prop = None
        if wl.get("proposal") == "laplace":
            prop = LaplaceProposal(d, torch.full((d,), wl["var"]), 1.0, torch.device("cpu"), torch.float32)
        elif wl.get("proposal") == "uniform":
            from rwm_pt_pytorch_b200.proposal_distributions import UniformRadiusProposal
            prop = UniformRadiusProposal(d, wl["var"], 1.0, torch.device("cpu"), torch.float32)
        algo = PT(d, wl["var"], t, geom_temp_spacing=True, swap_every=wl["swap_every"], burn_in=wl["burn_in"], device=dev,
                  num_ladders=n, store=store, seed=seed, chain_id_base=first * K, lanes_per_chain=lanes, swap_mode="reference",
                  proposal_distribution=prop, initial_states=np.zeros((n, 1, d), np.float32))
        batch = algo._require_batch()
    else:
        var = wl["var"]
        if var is None:   # C5: 64 variance values x (n / 64) chains
            xs = np.linspace(wl["sweep"][0], wl["sweep"][1], 64)
            var = np.repeat(xs ** 2 / d, max(n // 64, 1))[:n]
        np.random.seed(1 + rank)
        algo = RWM(d, var, t, burn_in=wl["burn_in"], device=dev, num_chains=n, store=store, seed=seed,
                   chain_id_base=first, lanes_per_chain=lanes)
        algo._ensure_batch(1)
        batch = algo._batch
    return algo, batch, t


def e2e_run(wl, t, batch, dev_index, T, reps, store="none"):
    """chain-steps/s through rwmpt_run_host: pinned host buffers -> H2D -> fused kernel -> D2H (state, accumulators and,
    when the workload stores trajectories, every retained row), wall-clock timed."""
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
    samples = slp = None
    if store != "none":
        n_stored = nc if store == "all" else n
        samples = torch.empty((n_stored, T, d), dtype=torch.float32).pin_memory()
        slp = torch.empty((n_stored, T), dtype=torch.float32).pin_memory()
        a.samples, a.sample_logp = samples.data_ptr(), slp.data_ptr()
        a.store_mode = _lib.STORE_MODES[store]
        a.store_start, a.thin, a.sample_stride, a.sample_rows = 0, 1, T, T
    h2d, d2h = C.c_uint64(), C.c_uint64()
    times = []
    batch.run(min(T, 20000))           # the set-up above left the GPU idle long enough to drop its clocks: ramp them up again
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
    """The path's other hand-written kernels timed alone with CUDA events on buffers several times larger than L2: the
    ESJD reduction over stored samples (HBM-read bound: 4*n*d algorithmic bytes per chain), the batched log-density and
    the proposal samplers."""
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


def roofline_block(name, wl, per_gpu_rate, peaks, default_shape):
    """t_min = max(F / P_fp32, S / P_sfu, bytes / BW_hbm) per chain-step; achieved = rate x the
    bounding resource's algorithmic work per chain-step.  default_shape: the workload runs its default number of units and
    storage mode, so the committed ncu capture (profiles/ncu_traffic.json) describes this launch."""
    fp32_tf, sfu_g, hbm_peak, have_file = peaks
    t_sfu = wl["S"] / (sfu_g * 1e9)
    t_fp = wl["F"] / (fp32_tf * 1e12)
    t_hbm = wl["bytes"] / (hbm_peak * 1e9)
    bound = max((t_sfu, "sfu"), (t_fp, "fp32"), (t_hbm, "hbm"))[1]
    if bound == "hbm":
        roof = {"bound": "hbm", "achieved": per_gpu_rate * wl["bytes"] / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "peak_source": "MEASURED_PEAKS.json hbm_gbs" if have_file else "fallback 6650 GB/s"}
    elif bound == "sfu":
        roof = {"bound": "sfu", "achieved": per_gpu_rate * wl["S"] / 1e9, "peak": sfu_g, "unit": "Gtranscendental/s",
                "peak_source": "measured on this GPU (rwmpt_probe_peaks, dependent-free MUFU.EX2 loop); nominal %.0f" % NOMINAL_SFU_GOPS}
    else:
        roof = {"bound": "fp32", "achieved": per_gpu_rate * wl["F"] / 1e12, "peak": fp32_tf, "unit": "TFLOP/s",
                "peak_source": "measured on this GPU (rwmpt_probe_peaks, dependent-free FFMA loop); nominal %.1f" % NOMINAL_FP32_TFLOPS}
    roof["frac"] = roof["achieved"] / roof["peak"]
    tr = None
    if default_shape and os.path.exists(NCU_TRAFFIC_FILE):
        try:
            tr = json.load(open(NCU_TRAFFIC_FILE)).get(name)
        except (OSError, ValueError):
            tr = None
    # ncu dram bytes per launch (read + write), next to the algorithmic bytes per launch; captures of stored-trajectory
    # workloads are taken at a shorter run and scale with the steps per launch, accumulator-only ones do not depend on it
    if tr:
        k = wl["T"] / tr["capture_steps_per_launch"] if tr.get("scales_with_steps") else 1.0
        roof["traffic"] = tr["bytes"] * k
        roof["traffic_source"] = tr["source"] + (f" (captured at {tr['capture_steps_per_launch']} steps per launch, scaled)" if k != 1.0 else "")
    else:
        roof["traffic"] = roof["traffic_source"] = None
    roof["algorithmic_bytes_per_launch"] = wl["bytes"] * wl["units"] * wl["K"] * wl["T"]
    roof["per_chain_step"] = {"F": wl["F"], "S": wl["S"], "stored_bytes": wl["bytes"]}
    roof["frac_sfu_measured"] = per_gpu_rate * wl["S"] / (sfu_g * 1e9)
    roof["frac_sfu_nominal"] = per_gpu_rate * wl["S"] / (NOMINAL_SFU_GOPS * 1e9)
    roof["frac_fp32_measured"] = per_gpu_rate * wl["F"] / (fp32_tf * 1e12)
    if wl["bytes"]:
        roof["frac_hbm"] = per_gpu_rate * wl["bytes"] / 1e9 / hbm_peak
    return roof


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: the workload's units PER GPU (default); strong: the workload's units in TOTAL, sharded over the ranks")
    ap.add_argument("--T", type=int, default=0, help="Metropolis steps per launch (0 = workload default)")
    ap.add_argument("--units", type=int, default=0, help="ladders / chains per GPU (0 = workload default)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per chain (0 = auto)")
    ap.add_argument("--store", default="", choices=["", "none", "cold", "all"], help="override the workload's trajectory storage")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU / reference baseline legs")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-aux", action="store_true")
    ap.add_argument("--quick", action="store_true", help="shorter baseline legs (development runs)")
    ap.add_argument("--also", default="default", help="'default' (every other BASELINE config when the workload is c3), 'none', or a comma list")
    ap.add_argument("--ref-leg", default="", help=argparse.SUPPRESS)
    ap.add_argument("--ref-n", type=int, default=1000, help=argparse.SUPPRESS)
    ap.add_argument("--ref-seed", type=int, default=1, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args.ref_leg:
        return _ref_leg_main(args.ref_leg, args.ref_n, args.ref_seed)
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
    first_unit = None
    total_units = wl["units"] * world
    if args.scaling == "strong":
        # the workload's units in total: contiguous blocks per rank, global Philox subsequences => identical results for any N
        total_units = wl["units"]
        first_unit, cnt = D.shard_range(total_units, rank, world)
        wl["units"] = cnt
    store = args.store or wl.get("store", "none")
    if store != "none" and "store" not in wl:
        wl["bytes"] = 4 * wl["dim"] + 4 if store == "all" else (4 * wl["dim"] + 4) / wl["K"]

    def measure(wl, store, with_clocks, steps, warmup, first_unit=None):
        algo, batch, t = build_sampler(wl, dev, rank, store, args.lanes, first_unit=first_unit)
        T, nc = wl["T"], wl["units"] * wl["K"]
        note = None
        if store != "none":
            # ONE buffer of T rows (+ row 0), re-used by every launch (LadderBatch.rewind_storage); shrink T when it does not fit
            n_stored = nc if store == "all" else wl["units"]
            free, _ = torch.cuda.mem_get_info(dev)
            fit = int(0.80 * free / (n_stored * (wl["dim"] + 1) * 4)) - 2
            if T > fit:
                note = f"steps per launch reduced from {T} to {fit}: the trajectory buffer must fit in {free / 2**30:.0f} GiB of free HBM"
                T = wl["T"] = fit
            batch.allocate_storage(store, T + 2, 1, with_logp=True)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)     # > 126 MB L2
        reduced = None
        # one nvidia-smi poller per job (rank 0), restricted to the job's own GPUs: NVML queries take a driver lock that CUDA
        # calls of the same process tree can wait on, and idle GPUs of the box would drag the median down
        sampler = ClockSampler(job_gpu_indices(world) if world > 1 else [job_gpu_indices(local + 1)[local]]) if (with_clocks and rank == 0) else None
        if sampler:
            sampler.start()                                               # start-up cost lands in the warm-up steps
        for w in range(warmup):
            if sampler and w == warmup - 1:
                # the poller's first row must have arrived before the LAST warm-up step: waiting for it leaves the GPU idle (clocks
                # drop), and the last warm-up step brings it back to speed right before the timed region
                torch.cuda.synchronize()
                sampler.wait_ready()
            batch.rewind_storage()
            batch.run(T)
            if world > 1:                                                 # also warms the NCCL communicator up
                D.allreduce_statistics_tensor(D.local_statistics_tensor(algo))
        torch.cuda.synchronize()
        if sampler:
            sampler.wait_ready()                                          # (--warmup 0)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        if sampler:
            sampler.mark()                                                # rows from here on belong to the timed region
        evs = []
        launches = 0
        for _ in range(steps):
            flush.fill_(1)                                                # untimed L2 flush between timed steps
            batch.rewind_storage()                                        # untimed: row 0 <- current state (stored workloads)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            batch.run(T)                                                  # ONE launch of the fused kernel
            launches += 1
            if world > 1:                                                 # the path's only in-loop collective: accumulators
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
        units_all = torch.tensor([float(wl["units"])], dtype=torch.float64, device=dev)
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
            dist.all_reduce(units_all, op=dist.ReduceOp.SUM)
        ms_per_step = ms / steps
        rate = float(units_all.item()) * wl["K"] * T / (ms_per_step * 1e-3)
        algo._refresh_stats()
        del flush
        return dict(rate=rate, ms_per_step=ms_per_step, launches=launches, clocks=clocks, algo=algo, batch=batch, t=t,
                    reduced=reduced, step_ms=[round(x, 3) for x in step_ms], note=note)

    def e2e_block(wl, m, store):
        if world > 1:
            dist.barrier()
        T_e = wl.get("e2e_T", wl["T"]) if store != "none" else wl["T"]
        T_e = min(T_e, wl["T"])
        e_time, h2d, d2h, e_acc, e_calls = e2e_run(wl, m["t"], m["batch"], local, T_e, reps=max(3, min(args.steps, 5)), store=store)
        units_all = float(wl["units"])
        if world > 1:
            tt = torch.tensor([e_time], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            uu = torch.tensor([units_all], dtype=torch.float64, device=dev)
            dist.all_reduce(uu, op=dist.ReduceOp.SUM)
            e_time, units_all = float(tt.item()), float(uu.item())
        out = {"value": units_all * wl["K"] * T_e / e_time, "unit": "chain-steps/s", "h2d_bytes_per_step": h2d * world,
               "d2h_bytes_per_step": d2h * world, "acceptance_rate": e_acc, "calls_ms": e_calls, "steps_per_call": T_e,
               "how": "rwmpt_run_host (C ABI, pinned host buffers): H2D + one fused launch + D2H per rank, wall clock, "
                      "median of the calls on each rank (calls_ms: rank 0), max over ranks"}
        if store != "none":
            out["d2h_gbs"] = d2h / e_time / 1e9
            out["note"] = "the D2H copy of the retained trajectories is inside the timed call (PCIe-bound)"
        return out

    m = measure(wl, store, True, args.steps, args.warmup, first_unit=first_unit)
    algo, batch = m["algo"], m["batch"]
    e2e = None if args.no_e2e else e2e_block(wl, m, store)

    # ---- the path's second collective on hardware: gather of retained samples over NCCL (cold chains of config 4) ----
    gather = None
    if world > 1 and not args.no_aux:
        w4 = dict(WORKLOADS["c4"], units=64, T=2000)
        algo4, b4, _ = build_sampler(w4, dev, rank, "cold", 0)
        b4.allocate_storage("cold", w4["T"] + 2, 1, with_logp=False)
        b4.run(w4["T"])
        local_block = b4.samples[:, : w4["T"] + 1]
        D.gather_samples(local_block, [w4["units"]] * world, dst=0)       # warm-up
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        got = D.gather_samples(local_block, [w4["units"]] * world, dst=0)
        e1.record()
        torch.cuda.synchronize()
        sums = torch.zeros(world, dtype=torch.float64, device=dev)
        sums[rank] = local_block.double().sum()
        dist.all_reduce(sums)
        if rank == 0:
            per = w4["units"]
            ok = all(abs(float(got[r * per:(r + 1) * per].double().sum()) - float(sums[r])) <= 1e-6 * max(abs(float(sums[r])), 1.0)
                     for r in range(world))
            nbytes = got.numel() * 4
            gather = {"what": "NCCL gather to rank 0 of the retained cold-chain samples (config 4 shape: 64 ladders/GPU x 2001 rows x 50)",
                      "bytes": nbytes, "ms": e0.elapsed_time(e1), "GB/s": nbytes / (e0.elapsed_time(e1) * 1e-3) / 1e9,
                      "blocks_match_rank_checksums": bool(ok)}
        del algo4, b4, got

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    fp32_tf, sfu_g = C.c_double(), C.c_double()
    _lib.check(_lib.load().rwmpt_probe_peaks(C.byref(fp32_tf), C.byref(sfu_g)))
    have_peaks = os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json"))
    peaks_json = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if have_peaks else {}
    hbm_peak = float(peaks_json.get("hbm_gbs", 6650.0))
    peaks = (fp32_tf.value, sfu_g.value, hbm_peak, have_peaks)
    base = WORKLOADS[args.workload]
    default_shape = wl["units"] == base["units"] and store == base.get("store", "none")   # (the steps per launch may differ)
    per_gpu = m["rate"] / world
    roof = roofline_block(args.workload, wl, per_gpu, peaks, default_shape)
    roof["measured_peaks"] = {"fp32_tflops": fp32_tf.value, "sfu_gops": sfu_g.value, "hbm_gbs": hbm_peak}

    post = batch.post_burn_in_steps()
    esjd = float(algo.expected_squared_jump_distance_gpu()) if post > 0 else None
    line = {
        "metric": "chain-steps/sec", "value": m["rate"], "unit": "chain-steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"] if args.scaling == "weak" else wl["desc"].replace("/GPU", " in total"),
                   "steps_per_launch": wl["T"], "chains_per_gpu": wl["units"] * wl["K"], "units_total": total_units,
                   "lanes_per_chain": args.lanes or "auto", "geometry_E_W": list(batch.geometry()),
                   "l2": "256 MiB flush write between timed steps", "rng": "in-kernel Philox4x32-10", "math": "fast"},
        "gpu_launches": m["launches"], "step_ms": m["step_ms"], "clocks": m["clocks"], "roofline": roof,
        "esjd": esjd, "esjd_per_sec": None if esjd is None else esjd * m["rate"] / wl["K"],
        "acceptance_rate": float(batch.accept_count.sum().item()) / max(post * batch.n_chains, 1),
    }
    if m["note"]:
        line["config"]["note"] = m["note"]
    if m["reduced"] is not None:
        line["all_ranks"] = D.pooled_summary(m["reduced"], m["reduced"]["chain_steps"] / wl["K"])
    if wl["kind"] == "pt":
        line["swap_acceptance_rate"] = algo.swap_acceptance_rate
        line["ladder_steps_per_sec"] = m["rate"] / wl["K"]
    if e2e is not None:
        line["e2e"] = e2e
    if gather is not None:
        line["gather"] = gather
    del algo, batch, m
    torch.cuda.empty_cache()
    if not args.no_cpu and world == 1:
        os.environ.setdefault("OMP_NUM_THREADS", "1")
        units, Tc = (128, 3000) if wl["kind"] == "pt" else (1024, 8000)
        if args.quick:
            Tc //= 4
        v, wall = cpu_arm(args.workload, 1, units, Tc)
        port = {"value": v, "unit": "chain-steps/s", "cores": 1, "kind": "port",
                "sample": f"NumPy oracle port, 1 process, {units} {'ladders' if wl['kind'] == 'pt' else 'chains'} x {Tc} steps ({wall:.1f} s wall)"}
        ref = reference_baselines(wl["kind"], quick=args.quick)
        if "unavailable" in ref:
            line["cpu_baseline"] = port
            line["reference_unavailable"] = ref["unavailable"]
        else:
            allc = ref["numpy_all_cores"]
            line["cpu_baseline"] = {
                "value": allc["chain_steps_per_s"], "unit": "chain-steps/s", "cores": allc["processes"], "kind": "reference",
                "sample": (f"UNMODIFIED reference NumPy sampler from baseline/_ref ({'algorithms/pt_rwm.py, config 3 target and 8-temperature ladder' if wl['kind'] == 'pt' else 'algorithms/rwm.py, config 1'}): "
                           f"{allc['processes']} processes x {allc['steps_per_process']} steps; single core: "
                           f"{ref['numpy_1core'].get('chain_steps_per_s', float('nan')):.0f} chain-steps/s"),
                "one_core": ref["numpy_1core"]}
            line["reference_torch"] = {"kind": "reference", "unit": "chain-steps/s",
                                       "what": "the reference's PyTorch class, one chain / ladder per object, generate_samples(n), unmodified",
                                       "cpu": ref["torch_cpu"], "cuda": ref["torch_cuda"]}
            line["oracle_port"] = port
    if world == 1 and args.workload == "c3" and not args.no_aux:
        line["aux_kernels"] = aux_kernel_rates(dev, hbm_peak)
    also = DEFAULT_ALSO if (args.also == "default" and args.workload == "c3" and args.scaling == "weak") else \
        ("" if args.also in ("default", "none") else args.also)
    if also and world == 1:
        line["also"] = {}
        # the strong-scaling regime on one GPU: config 3 with the ladders one GPU holds when 1024 are sharded over 2 / 4 / 8
        # GPUs (auto schedule = warp-specialised kernel below 3.5 ladders per SM) next to the fused kernel forced (schedule 1)
        few = {}
        for u in (512, 256, 128):
            wf = dict(WORKLOADS["c3"], units=u, T=100_000)
            row = {}
            for tag, sched in (("auto", None), ("fused_kernel_forced", "1")):
                old_env = os.environ.get("RWMPT_SCHEDULE")
                if sched is None:
                    os.environ.pop("RWMPT_SCHEDULE", None)
                else:
                    os.environ["RWMPT_SCHEDULE"] = sched
                mf = measure(wf, "none", False, 3, 3)
                row[tag] = mf["rate"]
                if old_env is None:
                    os.environ.pop("RWMPT_SCHEDULE", None)
                else:
                    os.environ["RWMPT_SCHEDULE"] = old_env
                del mf
            row["frac_of_1024_ladder_rate"] = row["auto"] / line["value"]
            few[str(u)] = row
        line["also"]["c3_few_ladders_per_gpu"] = {"unit": "chain-steps/s", "steps_per_launch": 100_000, "ladders": few,
                                                   "note": "per-GPU rate at the shard sizes of strong scaling (1024 ladders over 2 / 4 / 8 GPUs)"}
        for name in also.split(","):
            w2 = dict(WORKLOADS[name])
            st2 = w2.get("store", "none")
            m2 = measure(w2, st2, False, max(2, min(args.steps, 5)), 3)
            row = {"workload": w2["desc"], "value": m2["rate"], "unit": "chain-steps/s", "ms_per_step": m2["ms_per_step"],
                   "steps_per_launch": w2["T"], "step_ms": m2["step_ms"], "geometry_E_W": list(m2["batch"].geometry()),
                   "roofline": roofline_block(name, w2, m2["rate"], peaks, True),
                   "acceptance_rate": float(m2["batch"].accept_count.sum().item()) / max(m2["batch"].post_burn_in_steps() * m2["batch"].n_chains, 1)}
            if m2["note"]:
                row["note"] = m2["note"]
            if st2 != "none" and not args.no_e2e:
                # drop the big device buffer before the host-buffer call stages its own
                m2["batch"].allocate_storage("none", 0)
                torch.cuda.empty_cache()
                row["e2e"] = e2e_block(w2, m2, st2)
            if name == "c2" and not args.no_cpu:
                rb = reference_baselines("rwm", quick=args.quick)
                if "unavailable" not in rb:
                    row["cpu_baseline"] = {"value": rb["numpy_all_cores"]["chain_steps_per_s"], "unit": "chain-steps/s",
                                           "cores": rb["numpy_all_cores"]["processes"], "kind": "reference",
                                           "sample": "UNMODIFIED reference algorithms/rwm.py (config 1: RoughCarpet d=20, the only NumPy "
                                                     "target family the reference ships for this shape), one process per core",
                                           "one_core": rb["numpy_1core"]}
                    row["reference_torch"] = {"kind": "reference", "unit": "chain-steps/s", "cpu": rb["torch_cpu"], "cuda": rb["torch_cuda"],
                                              "what": "RandomWalkMH_GPU_Optimized on EvenRosenbrock d=20 (config 2), one chain per object"}
            line["also"][name] = row
            del m2
            torch.cuda.empty_cache()
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
