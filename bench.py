#!/usr/bin/env python
"""bench.py -- switch-agent decisions/sec of the SwitchFL lockstep hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA, one rank per GPU under torchrun)
    python bench.py --impl reference [...]                         the reference's CPU path (oracle port, all host cores)

Workload (N = 1): BASELINE.json configs[1] -- 4096 lockstep environments of the test_model.py map
(18x18, 2 trains, malfunction rate 0.01 / 5-15, gamma 1, eps .5 decay .9997, lr .1; synthetic-map stand-in
``c1_synth18`` because flatland's generators are unavailable) learning with distributed Q-learning.  One
"step" = one launch of the hot-path kernel advancing every environment by ``--ticks`` flatland ticks plus
every switch-agent decision, Q-update and episode reset in between.  N > 1: the same per GPU (weak scaling),
environments sharded by seed range, no data-path collective.

Prints ONE JSON line (rank 0).  value = device-timed (CUDA events, max over ranks) decisions/s with all
state resident in HBM; e2e = the same through the public Python API (DistrQLearning.learn_chunk) with the
per-env hyper-parameter block copied host->device and the per-env counters copied device->host every step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from __graft_entry__ import load_package  # noqa: E402

METRIC = "switch_agent_decisions_per_sec"
HP = dict(gamma=1.0, epsilon=0.5, epsilon_decay_rate=0.9997, lr=0.1, lr_decay_rate=1.0, default_q=0.0)   # test_model.py:56-63
SEED = 450565
# workload -> (fixture, envs per GPU, Q hash rows per env, description)
WORKLOADS = {
    "c2": ("c1_synth18", 4096, 1024,
           "C2: 4096 lockstep envs of the test_model.py map (c1_synth18 stand-in), distributed Q-learning, learn mode"),
    "c3": ("@c3", 20480, 32768,
           "C3: hyperparam_tuning.py default grid (eps .5, decay .9997, lr .1) x seeds {64,65,66,67,69} = 5 maps (80x80, 15 trains, "
           "25 cities, no malfunctions; synthetic stand-ins), each (map, point) replicated with distinct RNG streams: "
           "5 x 4096 envs per GPU, distributed Q-learning, learn mode"),
    "c4": ("c4_rail100_t50", 8192, 65536,
           "C4: large synthetic map (100x100, 50 trains, 354 switches, malfunctions), 65536 envs per 8 GPUs = 8192 per GPU, "
           "distributed Q-learning, learn mode"),
    "c5": ("c4_rail100_t50", 8192, 2,
           "C5 (extension, no reference counterpart): the C4 map in shared-table mode -- all 8192 envs of a GPU read one dense "
           "Q table and accumulate TD steps; every step (512 ticks) the integer accumulators are all-reduced over the GPUs "
           "(NCCL) and the mean step is folded into the table"),
}
SHARED_Q = False
FIXTURE = os.path.join(ROOT, "tests", "golden", "c1_synth18.fixture.npz")
WORKLOAD = WORKLOADS["c2"][3]


def select_workload(args):
    global FIXTURE, WORKLOAD, SHARED_Q
    name, envs, q_cap, desc = WORKLOADS[args.workload]
    SHARED_Q = args.workload == "c5"
    FIXTURE = name if name.startswith("@") else os.path.join(ROOT, "tests", "golden", name + ".fixture.npz")
    WORKLOAD = desc if not args.envs or args.envs == envs else desc + f" [envs per GPU overridden: {args.envs}]"
    args.envs = args.envs or envs
    args.q_cap = args.q_cap or q_cap


C3_SEEDS = (64, 65, 66, 67, 69)                                          # hyperparam_tuning.py:10


def workload_fixtures():
    """The maps of the selected workload: one for C2 / C4, the five seed maps of the hyper-parameter grid for C3."""
    from switchfl_b200 import mapgen
    if FIXTURE == "@c3":                                                  # hyperparam_tuning.py:17-26, synthetic stand-ins
        return [mapgen.c3_fixture(s) for s in C3_SEEDS]
    return [mapgen.load_fixture(FIXTURE)]


def bytes_per_decision(k_bar: float, P: float, A: float, A2: float) -> float:
    """Algorithmic bytes per decision, SURVEY.md section 8(d):  24*k + (26P+23) + 8A + (24P+138) + (32+8A')."""
    return 24.0 * k_bar + (26.0 * P + 23.0) + 8.0 * A + (24.0 * P + 138.0) + (32.0 + 8.0 * A2)


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi sampled every 20 ms from before the warm-up until the end of the e2e arm; the median SM clock is taken
    over the samples that fall inside the device-timed region when there are any, else over all samples under load."""
    Q = "timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index=0):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.window = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "20"], stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def mark(self, t0, t1):
        self.window = (t0, t1)

    def stop(self):
        import datetime
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, sm_in, mx, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 8:
                continue
            try:
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            sm.append(clk); mx.append(cmax)
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if self.window and self.window[0] <= ts <= self.window[1]:
                    sm_in.append(clk)
            except ValueError:
                pass
            for n, v in zip(names, parts[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            use = sm_in if sm_in else sm
            out.update(sm_mhz=float(np.median(use)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       samples_in_timed_region=len(sm_in))
        return out


# ---------------------------------------------------------------------------------------------- CPU arm
def _cpu_worker(args):
    seed, budget_s, n_ep, fixture = args
    global FIXTURE
    FIXTURE = fixture
    load_package()
    from switchfl_b200 import backend, mapgen
    from oracle.switchfl_oracle import SwitchFLOracle
    fxs = workload_fixtures()
    fx = fxs[seed % len(fxs)]
    rm = backend.RailMap(fx)
    o = SwitchFLOracle(fx, rm.tab, seed=seed, **HP)
    rng = np.random.default_rng(seed)
    o.episode = 0
    t0 = time.perf_counter()
    dec = 0
    eps = 0
    while True:
        m = o.run_episode(rng, greedy=False, learn=True)
        dec += m["decisions"]
        eps += 1
        if (n_ep and eps >= n_ep) or (not n_ep and time.perf_counter() - t0 >= budget_s):
            break
    return dec, time.perf_counter() - t0, eps


def cpu_sample(budget_s: float, cores: int, n_ep: int = 0, seed0: int = SEED):
    """The reference's parallelism model (hyperparam_tuning.py:85-91): independent OS processes, one per core,
    one seed each, no communication; each runs the oracle's learn() episodes."""
    import multiprocessing as mp
    ctx = mp.get_context("spawn")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(seed0 + i, budget_s, n_ep, FIXTURE) for i in range(cores)])
    wall = time.perf_counter() - t0
    dec = sum(r[0] for r in res)
    busy = max(r[1] for r in res)
    return dec, busy, wall, sum(r[2] for r in res)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step_s = 1.0
    for _ in range(args.warmup):
        cpu_sample(per_step_s, cores)
    dec, busy = 0, 0.0
    for _ in range(args.steps):
        d, b, _, _ = cpu_sample(per_step_s, cores)
        dec += d
        busy += b
    value = dec / busy
    sample = f"{cores} processes x {per_step_s:.0f} s of oracle learn() episodes per step on {os.path.basename(FIXTURE)}, one seed per process"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "decisions/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * busy / max(args.steps, 1), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_envs": cores, "note": "CPU oracle port of the reference path (reference itself needs flatland, absent); "
                       "its flatland substrate is lighter than real flatland, so this arm is faster than the true reference"},
            "cpu_baseline": {"value": value, "unit": "decisions/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------- our arm
def run_ours(args):
    import torch
    load_package()
    from switchfl_b200 import api, backend, mapgen
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    dev = f"cuda:{local}"
    fxs = workload_fixtures()
    B = args.envs
    parts = len(fxs)
    Bp = B // parts                                                       # environments per map
    B = Bp * parts

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def seeds_of(k):
        return np.arange(Bp, dtype=np.uint64) + np.uint64(SEED + rank * B + k * Bp)

    # ---------------- device-resident arm: Engine level, CUDA-event timed (one engine per map, one stream)
    rms = [backend.RailMap(fx) for fx in fxs]
    engs = [backend.Engine(rm, n_envs=Bp, device=dev, q_cap=args.q_cap, ep_cap=4, lanes=args.lanes or None, shared_q=SHARED_Q,
                           cta_warps=args.cta_warps or None) for rm in rms]
    lanes = engs[0].lanes
    for k, eng in enumerate(engs):
        eng.set_hparams(**HP, seeds=seeds_of(k), episodes=-1)
        eng.reset()
        eng.enable_q_init(True)
        if SHARED_Q:
            eng.init_shared_q(HP["default_q"])

    def launch(eng):
        eng.run(backend.MODE_LEARN, args.ticks)
        if SHARED_Q:
            eng.shared_q_sync(dist)                                       # all-reduce of the accumulators + apply kernel

    # Timing hygiene: the env state + Q tables are larger than L2, but on the small map the lines a launch actually touches
    # are not (ncu: 11 MB of DRAM traffic per launch), so L2 is flushed between the timed steps (256 MB written); the
    # flush sits inside the bracketed region, i.e. `value` pays for it.  Large-map workloads touch GBs per launch.
    state_bytes = sum(eng.sizes.state_bytes for eng in engs)
    flush = None
    if args.flush_l2 == "on" or (args.flush_l2 == "auto" and state_bytes < (1 << 30)):
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local) if rank == 0 else None
    for _ in range(args.warmup):
        for eng in engs:
            launch(eng)
    barrier()

    def totals():
        d = t = tt = 0
        for eng in engs:
            a, b_ = eng.total_decisions()
            d += a; t += b_; tt += int(eng.counters()["train_ticks"].sum())
        return d, t, tt

    d0, t0, tt0 = totals()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.time()
    start, stop = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    start.record()
    for a, b in evs:
        if flush is not None:
            flush.zero_()
        a.record()
        for eng in engs:
            launch(eng)
        b.record()
    stop.record()
    barrier()
    if sampler:
        sampler.mark(wall0, time.time())
    ms = start.elapsed_time(stop)
    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in evs])) / parts            # mean duration of ONE k_run launch
    d1, t1, tt1 = totals()
    episodes = aborted = q_rows_max = 0
    state_mb = 0.0
    for eng in engs:
        # On congested maps the reference itself dies in observer.py:294-307 ("No train detected at active switch");
        # the kernel abandons such an episode and resets the env.  Every other error bit is fatal here.
        eng.check_errors(allow=backend.ERR_NO_TRAIN_AT_SWITCH)
        cn = eng.counters()
        episodes += int(cn["episodes"].sum()); aborted += int(cn["aborted"].sum())
        q_rows_max = max(q_rows_max, int(cn["q_rows"].max()))
        state_mb += eng.sizes.state_bytes / 1e6
        eng.close()
    dec, ticks, train_ticks = d1 - d0, t1 - t0, tt1 - tt0
    del engs
    torch.cuda.empty_cache()

    # ---------------- end-to-end arm: public API, host buffers in the timed region
    models = []
    for k, fx in enumerate(fxs):
        env = api.ASyncSwitchEnv(api.RailEnv(fx), max_steps=100_000, n_envs=Bp, device=dev, q_cap=args.q_cap, ep_cap=4,
                                 shared_q=SHARED_Q, _engine_kwargs={"lanes": args.lanes or None, "cta_warps": args.cta_warps or None})
        models.append(api.DistrQLearning(env=env, seeds=seeds_of(k), dist=dist, **HP))
    for _ in range(args.warmup):
        for m in models:
            m.learn_chunk(args.ticks)
    barrier()
    c0 = sum(int(m.learn_chunk(0)["decisions"].sum()) for m in models)
    barrier()
    w0 = time.perf_counter()
    for _ in range(args.steps):
        cs = [m.learn_chunk(args.ticks) for m in models]
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - w0
    e2e_dec = sum(int(c["decisions"].sum()) for c in cs) - c0
    clocks = sampler.stop() if sampler else None
    h2d = sum(int(m.env.engine.sizes.hparams_bytes) for m in models)
    d2h = sum(int(m.env.engine.sizes.counters_bytes) for m in models)

    # ---------------- reduce over ranks: max time, summed work
    if dist is not None:
        t = torch.tensor([ms, e2e_s, kern_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        w = torch.tensor([dec, ticks, train_ticks, e2e_dec], device=dev, dtype=torch.int64)
        dist.all_reduce(w, op=dist.ReduceOp.SUM)
        ms, e2e_s, kern_ms = (float(x) for x in t.tolist())
        dec, ticks, train_ticks, e2e_dec = (int(x) for x in w.tolist())
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    value = dec / (ms / 1000.0)
    k_bar = train_ticks / max(dec, 1)
    P = float(np.mean(np.concatenate([rm.tab.sw_P for rm in rms]))); A = float(np.mean(np.concatenate([rm.tab.sw_A for rm in rms])))
    bpd = bytes_per_decision(k_bar, P, A, A)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    per_gpu_dec_per_launch = dec / world / args.steps / parts
    achieved = per_gpu_dec_per_launch * bpd / (kern_ms / 1000.0) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath)).get(args.workload, {})
        if tj.get("ticks") == args.ticks and tj.get("envs") == B:
            traffic = tj.get("dram_bytes_per_launch")
    line = {"metric": METRIC, "value": value, "unit": "decisions/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_envs_per_gpu": B, "maps": parts, "ticks_per_step": args.ticks, "q_cap": args.q_cap, "lanes_per_env": lanes,
                       "train_ticks_per_decision": k_bar, "ticks": ticks, "episodes_rank0": episodes,
                       "episodes_abandoned_rank0": aborted, "q_rows_max_rank0": q_rows_max,
                       "l2": (f"L2 flushed between the timed steps (256 MB written, inside the timed region); env state + Q tables "
                              f"{state_mb:.0f} MB per GPU" if flush is not None else
                              f"inputs larger than L2: {state_mb:.0f} MB of env state + Q tables per GPU vs 126 MB L2, GBs touched per launch"),
                       "sharding": ("envs by seed range; per step one integer all-reduce (sum) of the shared table's accumulators" if SHARED_Q
                                    else "envs by seed range, no data-path collective")},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "kernel": "k_run", "bytes_per_decision": bpd, "kernel_ms": kern_ms, "peak_source": peak_src,
                         "note": "per-env decision chains are serial: latency/issue-bound, not HBM-bound (SURVEY 8d honest note)"},
            "e2e": {"value": e2e_dec / e2e_s, "unit": "decisions/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps * parts * (2 if SHARED_Q else 1), "clocks": clocks}
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        d, busy, wall, eps = cpu_sample(args.cpu_seconds, cores)
        line["cpu_baseline"] = {"value": d / busy, "unit": "decisions/s", "cores": cores, "kind": "port",
                                "sample": f"{cores} processes x {args.cpu_seconds:.0f} s of oracle learn() episodes on the same map "
                                          f"({eps} episodes, {d} decisions), one seed per process as hyperparam_tuning.py:85-91"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS), help="c2 = BASELINE.json configs[1] (default), c4 = configs[3]")
    ap.add_argument("--envs", type=int, default=0, help="environments per GPU (0 = the workload's own)")
    ap.add_argument("--ticks", type=int, default=512, help="flatland ticks per env per step (launch)")
    ap.add_argument("--q-cap", type=int, default=0, help="Q hash rows per environment (0 = the workload's own)")
    ap.add_argument("--lanes", type=int, default=0, help="lanes of a warp per environment (0 = library default for the batch size)")
    ap.add_argument("--cta-warps", type=int, default=0, help="warps per CTA of the hot-path kernel (0 = library default)")
    ap.add_argument("--flush-l2", default="auto", choices=["auto", "on", "off"],
                    help="write 256 MB between the timed steps (auto: when the whole state is under 1 GiB)")
    ap.add_argument("--cpu-seconds", type=float, default=3.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    select_workload(args)
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
