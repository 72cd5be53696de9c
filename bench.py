#!/usr/bin/env python3
"""bench.py — MPC solves/sec of the hot path on N B200s (BASELINE.json metric), one JSON line on rank 0.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config c3] [--impl ours|reference]

A "step" is one pass of the hot path over one batch of synthetic requests: config C3 of BASELINE.json
(batch 65536 per GPU, control_steps = 10, footprint + costmap on a 1000x1000 grid, opt_tolerance = 1e-3) —
the configuration the >= 1e6 solves/s target is quoted on.  Weak scaling: every rank solves its own 65536
requests of the same distribution and the solved (vx, vy, omega) are all-gathered over NCCL inside the step.

  value      solves/s with the requests already resident in HBM (kernel + gather), CUDA-event time, max over ranks
  e2e        solves/s through neompc_solve_batch with pinned HOST buffers: H2D requests, solve, D2H responses
  roofline   algorithmic bytes (76 B/solve + costmap once per launch) / measured kernel time
  cpu_baseline  the reference algorithm (oracle port of srv.py:363-364, scipy SLSQP) on a bounded sample, all host cores

`--impl reference` times only the CPU reference arm (rank 0), same metric and config.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if "reference" in sys.argv:      # the CPU arm: one thread per worker process (set before numpy loads its BLAS)
    for _v in ("OMP_NUM_THREADS", "MKL_NUM_THREADS", "OPENBLAS_NUM_THREADS"):
        os.environ.setdefault(_v, "1")

import numpy as np  # noqa: E402

METRIC = "mpc_solves_per_sec"
UNIT = "solves/s"
L2_FLUSH_BYTES = 256 << 20


# ------------------------------------------------------------------------------------------------ CPU reference arm
_W = {}


PER_GPU_BATCH = {"c2": 4096, "c3": 65536, "c4": 131072, "c5": 100000}


def _cpu_init(cfg, batch, use_port):
    os.environ["OMP_NUM_THREADS"] = "1"
    import oracle
    from oracle import ref_runner
    from oracle.costmap import GridCostmap, FreeSpaceCostmap
    from neo_mpc_planner2_b200 import workloads
    wl = workloads.config(cfg, batch=batch)
    _W["wl"] = wl
    _W["p"] = oracle.MpcParams(**wl.params)
    _W["cm"] = (GridCostmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y) if wl.cells is not None
                else FreeSpaceCostmap())
    _W["oracle"] = oracle
    # the UNMODIFIED reference (oracle/_ref, byte-compiled from /root/reference by oracle/build_ref.py) under ROS stand-ins
    _W["ref"] = None if use_port or not ref_runner.available() else ref_runner.ReferenceSolver(wl.params, _W["cm"], wl.footprint)


def _cpu_solve(i):
    """One reference solve, exactly the reference's call (srv.py:363-364): cold start, SLSQP, ftol = opt_tolerance —
    on the reference's own objects when oracle/_ref is there, else on the oracle port of the same lines."""
    oracle, wl, p, cm = _W["oracle"], _W["wl"], _W["p"], _W["cm"]
    from oracle.mpc_oracle import footprint_world
    prob = oracle.Problem.from_record(wl.requests[i])
    if _W["ref"] is not None:
        res = _W["ref"].solve(prob)
    else:
        fpw = footprint_world(wl.footprint, prob.pose_x, prob.pose_y, prob.pose_yaw)
        res = oracle.slsqp_solve(p, cm, fpw, prob)
    return i, float(res.fun), res.x.astype(np.float64)


def _cpu_solve_tight(args):
    """The same problem converged tightly with the reference's solver (ftol 1e-10): best of a cold start, a start from the
    ftol = opt_tolerance point and a start from the GPU's solution — the yardstick for 'how far is a first control from
    the best known optimum' (untimed; only with --dump-ref)."""
    i, x_gpu = args
    oracle, wl, p, cm = _W["oracle"], _W["wl"], _W["p"], _W["cm"]
    from oracle.mpc_oracle import footprint_world
    prob = oracle.Problem.from_record(wl.requests[i])
    fpw = footprint_world(wl.footprint, prob.pose_x, prob.pose_y, prob.pose_yaw)
    res = oracle.slsqp_solve(p, cm, fpw, prob)
    cands = [oracle.slsqp_solve(p, cm, fpw, prob, x0=res.x, ftol=1e-10, maxiter=400),
             oracle.slsqp_solve(p, cm, fpw, prob, ftol=1e-10, maxiter=400)]
    if x_gpu is not None:
        cands.append(oracle.slsqp_solve(p, cm, fpw, prob, x0=np.asarray(x_gpu, dtype=np.float64), ftol=1e-10, maxiter=400))
    best = min(cands, key=lambda r: r.fun)
    return i, float(best.fun), best.x.astype(np.float64)


def common_config(args, n_steps, opt_tolerance, batch_per_gpu):
    """The `config` object — the same keys and values in both arms (`--impl ours` and `--impl reference`)."""
    wl_name, _ = _workload_name(args.config)
    return {"workload": wl_name, "batch_per_gpu": int(batch_per_gpu), "control_steps": int(n_steps),
            "opt_tolerance": float(opt_tolerance), "cold_start": True,
            "footprint_mode": "moving (opt-in, not the reference's objective)" if args.footprint_mode else "static (reference)",
            "costmap_mode": "bilinear (opt-in, not the reference's objective)" if args.costmap_mode else "nearest cell (reference)",
            "l2": f"flushed between timed iterations ({L2_FLUSH_BYTES >> 20} MiB write)",
            "parallelism": (f"batch sharded over {args.gpus} GPU(s), one NCCL all-gather of (vx,vy,omega) per step"
                            if args.gpus > 1 else "single GPU")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        cores = len(os.sched_getaffinity(0))
    except AttributeError:
        cores = os.cpu_count() or 1
    from oracle import ref_runner
    use_port = bool(args.port) or not ref_runner.available()
    sample = args.cpu_sample or {"c2": 512, "c3": 64, "c4": 16, "c5": 64}[args.config]
    per_gpu = args.batch or PER_GPU_BATCH[args.config]
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init, initargs=(args.config, per_gpu, use_port)) as pool:
        for _ in range(max(args.warmup, 1)):
            pool.map(_cpu_solve, range(min(cores, sample)))
        t0 = time.perf_counter()
        last = None
        for _ in range(args.steps):
            last = pool.map(_cpu_solve, range(sample), chunksize=max(1, sample // (cores * 4)))
        dt = time.perf_counter() - t0
        tight = None
        if args.dump_ref:                                           # untimed: tightly converged yardstick
            kt = min(sample, 4 * cores)
            xg = np.load(args.gpu_plan)["plan"] if args.gpu_plan else None
            tight = pool.map(_cpu_solve_tight, [(i, None if xg is None else xg[i]) for i in range(kt)])
    if args.dump_ref and last:
        np.savez(args.dump_ref, J=np.array([r[1] for r in last]), x=np.stack([r[2] for r in last]),
                 J_tight=np.array([r[1] for r in tight]), x_tight=np.stack([r[2] for r in tight]))
    value = args.steps * sample / dt
    from neo_mpc_planner2_b200 import workloads
    wl = workloads.config(args.config, batch=64)
    kind = "port" if use_port else "reference"
    how = ("oracle port of srv.py:204-269 + :363-364 (oracle/mpc_oracle.py)" if use_port else
           "the UNMODIFIED mpc_optimization_server.py (oracle/_ref, byte-compiled from /root/reference) under ROS stand-in "
           "modules: its own objective / f_constraint / bnds / cons through minimize(SLSQP) exactly as srv.py:363-364")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": common_config(args, wl.control_steps, wl.params["opt_tolerance"], per_gpu),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "per_core": value / cores, "how": how,
                         "sample": f"first {sample} problems of the workload per step, cold start, "
                                   f"multiprocessing.Pool({cores})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def _workload_name(cfg):
    from neo_mpc_planner2_b200 import workloads
    wl = workloads.config(cfg, batch=64)
    names = {"c2": "C2 batch=4096 control_steps=3 costmap 200x200",
             "c3": "C3 batch=65536/GPU control_steps=10 footprint+costmap 1000x1000",
             "c4": "C4 batch=131072/GPU control_steps=20 costmap 1000x1000",
             "c5": "C5 fleet sweep 100k poses x 8 carrots, control_steps=10, costmap 2000x2000"}
    return names.get(cfg, cfg), wl.control_steps


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.power_w = []
        self.reasons = set()
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        while not self._halt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                self.power_w.append(self.nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                r = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples),
                "power_w_max": float(max(self.power_w)) if self.power_w else None}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from neo_mpc_planner2_b200 import workloads
    from neo_mpc_planner2_b200.abi import REQUEST_DTYPE, RESPONSE_DTYPE
    from neo_mpc_planner2_b200.solver import BatchSolver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback for the solve)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    per_gpu = PER_GPU_BATCH[args.config]
    if args.batch:
        per_gpu = args.batch
    # every rank: same map, its own slice of the request distribution (seeded by rank)
    wl = workloads.config(args.config, batch=per_gpu, seed=None if rank == 0 else 1000 + rank)
    n = wl.batch
    n_steps = wl.control_steps
    knobs = dict(lanes_per_instance=args.lanes, footprint_mode=args.footprint_mode, costmap_mode=args.costmap_mode,
                 costmap_guidance=args.costmap_guidance)
    fleet = None
    if world > 1:
        # one rank of a fleet: the sharding and the single collective (NCCL all-gather of the solved twists, enqueued by
        # the library behind its solve kernel) are libneompc's own (include/neompc.h "multi-GPU")
        from neo_mpc_planner2_b200.fleet import FleetSolver
        fleet = FleetSolver(wl.params, device=local, **knobs)
        solver = fleet.solver
    else:
        solver = BatchSolver(wl.params, device=local, **knobs)
    solver.load_workload(wl)
    G, S = solver.tiling_for(n)          # what the dispatcher launches for this batch (C2's 4096 requests: the latency tiling)

    req_host = torch.from_numpy(wl.requests.view(np.uint8).reshape(n, REQUEST_DTYPE.itemsize)).pin_memory()
    resp_host = torch.empty((n, RESPONSE_DTYPE.itemsize), dtype=torch.uint8).pin_memory()
    d_reqs = req_host.to(dev)
    d_out = torch.empty((n, RESPONSE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    # the gather payload / result, double-buffered: [world * n, 3], this rank's rows are [rank * n, (rank + 1) * n)
    d_all = [torch.zeros((world * n, 3), dtype=torch.float32, device=dev) for _ in range(2)]
    all_host = torch.empty((world * n, 3), dtype=torch.float32).pin_memory()
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)              # solve kernel + all timing events
    torch.cuda.set_stream(stream)

    def launch_step(k, k_ev=None):
        """Solve kernel of step k on `stream`; with N > 1 the library also enqueues the all-gather of step k on its own
        communication stream behind the kernel."""
        if k_ev is not None:
            k_ev[0].record(stream)
        if world > 1:
            solver.solve_gather_device(d_reqs.data_ptr(), n, n, d_out.data_ptr(), d_all[k & 1].data_ptr(), stream.cuda_stream)
        else:
            solver.solve_device(d_reqs.data_ptr(), n, d_out.data_ptr(), d_all[k & 1].data_ptr(), None, stream.cuda_stream)
        if k_ev is not None:
            k_ev[1].record(stream)               # (brackets the kernel only: the gather is on the other stream)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for k in range(max(args.warmup, 3)):
        flush.fill_(1)
        launch_step(k)
        if world > 1:
            solver.gather_wait(stream.cuda_stream, 0)
    barrier()

    sampler = ClockSampler(local)
    sampler.start()
    launches0 = solver.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True),
            torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps + 1)]
    barrier()
    # A timed step = one solve kernel + one all-gather.  With N > 1 the step is software-pipelined: the gather of step
    # k-1 (which depends on kernel k-1 only) runs on the library's communication stream while kernel k computes; the
    # bracket of step k closes when kernel k AND gather k-1 are done.  A last bracket drains the gather of the final step,
    # so K kernels and K gathers are inside timed regions; the L2 flush between steps is outside them.
    for k in range(args.steps):
        flush.fill_(k & 0xFF)                                    # L2 flush between timed iterations (untimed)
        evs[k][0].record(stream)
        launch_step(k, (evs[k][2], evs[k][3]))
        if world > 1 and k > 0:
            solver.gather_wait(stream.cuda_stream, 1)            # the gather before the one just enqueued
        evs[k][1].record(stream)
    evs[args.steps][0].record(stream)
    if world > 1:
        solver.gather_wait(stream.cuda_stream, 0)
    evs[args.steps][1].record(stream)
    barrier()
    launches = solver.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b, _, _ in evs]                    # K steps + the drain bracket
    kern_ms = [c.elapsed_time(d) for _, _, c, d in evs[:args.steps]]
    total_ms = float(sum(step_ms))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    value = world * n * args.steps / (total_ms_max * 1e-3)

    # ---- what was gathered is what was solved (untimed)
    gather_check = None
    if world > 1:
        last = d_all[(args.steps - 1) & 1]
        resp_dev = np.frombuffer(d_out.cpu().numpy().tobytes(), dtype=RESPONSE_DTYPE)
        mine = np.stack([resp_dev["vx"], resp_dev["vy"], resp_dev["omega"]], axis=1)
        own_ok = last[rank * n:(rank + 1) * n].cpu().numpy().tobytes() == mine.tobytes()
        # every rank holds the same tensor: compare a position-weighted checksum of the raw bits across ranks
        bits = last.view(torch.int32).to(torch.int64).flatten()
        digest = torch.stack([bits.sum(), (bits * (torch.arange(bits.numel(), device=dev) % 8191 + 1)).sum()])
        digests = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(digests, digest)
        same = all(bool((d == digests[0]).all()) for d in digests)
        # and equal to ONE GPU solving the same requests: rank 0 re-solves the first rows of every rank's shard
        kchk = min(n, 8192 // world)
        head = [torch.zeros((kchk, REQUEST_DTYPE.itemsize), dtype=torch.uint8, device=dev) for _ in range(world)]
        dist.all_gather(head, d_reqs[:kchk].contiguous())
        flags = torch.tensor([int(own_ok), int(same), 1], device=dev)
        if rank == 0:
            chk = BatchSolver(wl.params, device=local, **dict(knobs, lanes_per_instance=G))
            chk.load_workload(wl)
            reqs_all = np.frombuffer(torch.cat(head).cpu().numpy().tobytes(), dtype=REQUEST_DTYPE)
            one = chk.solve(reqs_all)
            chk.close()
            one_tw = np.stack([one["vx"], one["vy"], one["omega"]], axis=1).reshape(world, kchk, 3)
            got = last.view(world, n, 3)[:, :kchk].cpu().numpy()
            flags[2] = int(got.tobytes() == one_tw.tobytes())
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        gather_check = {"own_rows_equal_own_responses": bool(flags[0]), "all_ranks_hold_the_same_tensor": bool(flags[1]),
                        "equal_to_one_gpu_solve": bool(flags[2]), "rows_compared_with_one_gpu": int(world * kchk)}
        assert all(bool(f) for f in flags), gather_check

    # ---- e2e: host buffers (H2D + solve [+ gather] + D2H inside the timed region)
    def e2e_step():
        if world > 1:
            d_reqs.copy_(req_host, non_blocking=True)                                   # H2D of this rank's shard
            solver.solve_gather_device(d_reqs.data_ptr(), n, n, d_out.data_ptr(), d_all[0].data_ptr(), stream.cuda_stream)
            solver.gather_wait(stream.cuda_stream, 0)
            all_host.copy_(d_all[0], non_blocking=True)                                 # D2H of ALL ranks' twists
            stream.synchronize()
        else:
            solver.solve_twists_raw(req_host.data_ptr(), n, all_host.data_ptr())        # H2D requests, solve, D2H twists
    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = args.steps
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_s = float(t.item())
    e2e_value = world * n * e2e_steps / e2e_s
    e2e_full = None
    host_path = None
    if world == 1:
        host_path = {1: "pinned mailbox", 2: "staged copies pipelined with the solve in chunks (H2D requests, D2H results)",
                     3: "zero-copy: the solve kernel reads the pinned request buffer and writes the pinned result buffer over PCIe "
                        "itself (one launch); the bytes cross the bus inside the timed region all the same"}.get(solver.last_host_path)
    if world > 1:
        resp_host.copy_(d_out)
        torch.cuda.synchronize(dev)
    else:                                        # the same with the full 32-byte responses coming back
        for _ in range(2):
            solver.solve_raw(req_host.data_ptr(), n, resp_host.data_ptr())
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            solver.solve_raw(req_host.data_ptr(), n, resp_host.data_ptr())
        torch.cuda.synchronize(dev)
        e2e_full = {"value": n * e2e_steps / (time.perf_counter() - t0), "unit": UNIT,
                    "d2h_bytes_per_step": int(n * RESPONSE_DTYPE.itemsize), "api": "neompc_solve_batch (full responses)"}
    clocks = sampler.stop()

    # ---- sustained: >= 2 s of back-to-back solve kernels (no L2 flush, no host gaps) — does the flushed 20-step figure
    # survive once the board is at temperature and power?  Reported next to `value`, which stays the flushed figure.
    k_est = max(float(np.mean(kern_ms)), 1e-3)
    n_sus = int(args.sustained_s * 1e3 / k_est) + 1
    sus_sampler = ClockSampler(local)
    sus_sampler.start()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    s0.record(stream)
    for k in range(n_sus):
        solver.solve_device(d_reqs.data_ptr(), n, d_out.data_ptr(), d_all[k & 1].data_ptr(), None, stream.cuda_stream)
    s1.record(stream)
    barrier()
    sus_ms = s0.elapsed_time(s1)
    sus_clocks = sus_sampler.stop()
    sustained = {"seconds": sus_ms / 1e3, "steps": n_sus, "ms_per_step": sus_ms / n_sus,
                 "value_per_gpu": n * n_sus / (sus_ms * 1e-3), "unit": UNIT, "l2": "not flushed (back-to-back launches)",
                 "sm_mhz_median": sus_clocks["sm_mhz"], "power_w_max": sus_clocks["power_w_max"],
                 "reasons": sus_clocks["reasons"], "clock_samples": sus_clocks["samples"]}

    # ---- latency of ONE request through the message-level entry (what the controller plugin calls every tick,
    # replacing the ROS service hop + scipy solve): neompc_solve_msgs, stateful instance 0, host buffers
    lat = None
    if rank == 0:
        from neo_mpc_planner2_b200.server import requests_to_msgs
        solver.reserve_instances(1)
        msg = requests_to_msgs(wl.requests[:1])
        msg["instance_id"] = 0
        msg["delta_t"] = 1.0 / 30.0
        for _ in range(20):
            solver.solve_msgs(msg)
        ts = []
        for _ in range(200):
            solver.reset_state()                                  # cold start every time, like the batch numbers
            t1 = time.perf_counter()
            solver.solve_msgs(msg)
            ts.append(time.perf_counter() - t1)
        lat = {"n": 1, "api": "neompc_solve_msgs (pack + solve + D2H, cold start)", "median_us": 1e6 * float(np.median(ts)),
               "p99_us": 1e6 * float(np.percentile(ts, 99))}

    # ---- full controller tick through the C++ plugin (NeoMpcPlanner::computeVelocityCommands over libneompc: costmap
    # checksum / upload when it changed, front half, solve, one synchronise), control_steps = 10, on a 60x60 local
    # costmap and on a 1000x1000 one
    plugin_lat = None
    demo = os.path.join(ROOT, "neo_mpc_planner2_b200", "plugin", "plugin_demo")
    if rank == 0 and os.path.exists(demo):
        import subprocess
        plugin_lat = []
        for w, h in ((60, 60), (1000, 1000)):
            try:
                r = subprocess.run([demo, "latency", str(w), str(h)], capture_output=True, text=True, timeout=120,
                                   env=dict(os.environ, CUDA_VISIBLE_DEVICES=os.environ.get("CUDA_VISIBLE_DEVICES", str(local))))
                plugin_lat.append(json.loads(r.stdout.strip().splitlines()[-1]))
            except Exception as exc:
                plugin_lat.append({"costmap": [w, h], "error": str(exc)})

    resp = np.frombuffer(resp_host.numpy().tobytes(), dtype=RESPONSE_DTYPE)
    iters_med = float(np.median(resp["iters"]))
    evals_mean = float(resp["evals"].mean())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "6650 GB/s (of fallback)"
        bytes_per_solve = wl.algorithmic_bytes_per_solve(n)
        # what the round-2 kernel is expected to pull from HBM per launch: the requests + the corner-packed costmap copy
        # (4 cells per 32-bit word, padded by twice a plan's reach)
        import math
        kernel_reads = n * REQUEST_DTYPE.itemsize
        if wl.cells is not None:
            reach = math.ceil(float(wl.params["max_vel_trans"]) * float(wl.params["prediction_horizon"]) / wl.resolution)
            pad = 2 * reach + 6
            kernel_reads += (wl.cells.shape[1] + 2 * pad) * (wl.cells.shape[0] + 2 * pad) * 4
        k_ms = float(np.mean(kern_ms))
        achieved = bytes_per_solve * n / (k_ms * 1e-3) / 1e9
        wl_name, _ = _workload_name(args.config)
        traffic = None            # dram bytes per launch from the committed ncu capture of the same kernel/config
        issue = None              # the ceiling that actually binds: warp-instruction issue rate (SURVEY §7 hard part 6)
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(args.config)
            if tr and tr["kernel"] == f"solve_kernel<{G},{S}>" and tr["batch"] == n and not args.footprint_mode and not args.costmap_mode:
                traffic = tr["dram_bytes_per_launch"]
                sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
                mhz = clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0
                peak_issue = sm_count * 4 * mhz * 1e6            # 4 schedulers per SM, one warp instruction per clock each
                ach_issue = tr["inst_executed"] / (k_ms * 1e-3)  # instruction count of the committed capture / live time
                issue = {"achieved_gwarp_inst_per_s": ach_issue / 1e9, "peak_gwarp_inst_per_s": peak_issue / 1e9,
                         "frac": ach_issue / peak_issue, "warp_inst_per_launch": tr["inst_executed"],
                         "source": "inst_executed from " + tr["source"].split(":")[0] + " (ncu), time measured live"}
        except Exception:
            pass
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": common_config(args, n_steps, wl.params["opt_tolerance"], n),
            "arm": {"lanes_per_instance": G, "steps_per_lane": S, "iters_median": iters_med, "evals_mean": evals_mean,
                    "costmap_guidance": "off (round-1 strategy)" if args.costmap_guidance else "on",
                    "gather": "overlapped with the next step's solve on a second stream" if world > 1 else None},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_solve * n,
                         "expected_kernel_reads_per_launch": int(kernel_reads), "kernel": f"solve_kernel<{G},{S}>", "kernel_ms": k_ms,
                         "bytes_per_solve": bytes_per_solve, "peak_source": peak_src,
                         "issue_slots": issue,
                         "note": "path is instruction/latency bound (FP32 + MUFU + shuffles), not HBM bound; see DESIGN.md"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(world * n * REQUEST_DTYPE.itemsize),
                    "d2h_bytes_per_step": int(n * 12) if world == 1 else int(world * world * n * 12),
                    "ms_per_step": 1e3 * e2e_s / e2e_steps, "with_full_responses": e2e_full, "host_path": host_path,
                    "api": "neompc_solve_batch_twists (pinned host buffers: requests in, (vx,vy,omega) out)" if world == 1 else
                           "per rank: H2D of its shard, neompc_solve_gather_device (solve + NCCL all-gather), D2H of all ranks' "
                           "(vx,vy,omega); bytes are whole-job totals"},
            "gather_check": gather_check,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "single_request_latency": lat,
            "plugin_tick_latency": plugin_lat,
            "sustained": sustained,
        }
        if world == 1 and not args.no_cpu_baseline and not args.footprint_mode and not args.costmap_mode:
            # the reference arm in a FRESH interpreter (no CUDA context / torch thread pools in the forked workers)
            import subprocess
            import tempfile
            sample = args.cpu_sample or {"c2": 1024, "c3": 128, "c4": 32, "c5": 128}[args.config]   # ~10-30 s of CPU work
            env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1")
            tmpd = tempfile.mkdtemp(prefix="neompc_ref_")
            dump, gplan = os.path.join(tmpd, "ref.npz"), os.path.join(tmpd, "gpu_plan.npz")
            sub = wl.requests[:sample]
            _, plan = solver.solve(sub, want_plan=True)
            np.savez(gplan, plan=plan)
            base = [sys.executable, os.path.abspath(__file__), "--impl", "reference", "--config", args.config,
                    "--batch", str(per_gpu)]
            res = subprocess.run(base + ["--steps", "1", "--warmup", "1", "--cpu-sample", str(sample), "--dump-ref", dump,
                                         "--gpu-plan", gplan], capture_output=True, text=True, env=env, timeout=900)
            try:
                ref = json.loads(res.stdout.strip().splitlines()[-1])
                line["cpu_baseline"] = ref["cpu_baseline"]
                line["cpu_baseline"]["sample"] += f", 1 timed pass of {ref['ms_per_step'] / 1e3:.1f} s"
                if ref["cpu_baseline"]["kind"] == "reference":       # the oracle port of the same lines, for comparison
                    rp = subprocess.run(base + ["--port", "--steps", "1", "--warmup", "1", "--cpu-sample", str(sample)],
                                        capture_output=True, text=True, env=env, timeout=900)
                    line["cpu_baseline"]["oracle_port_value"] = json.loads(rp.stdout.strip().splitlines()[-1])["value"]
                # cost residual J_gpu - J_scipy on the same problems (BASELINE.json metric), both evaluated by the
                # float64 oracle objective at the respective solutions (untimed; checker use of oracle/)
                import oracle
                from oracle.costmap import GridCostmap
                from oracle.mpc_oracle import footprint_world
                refd = np.load(dump)
                k = len(refd["J"])
                pm = oracle.MpcParams(**wl.params)
                cm = GridCostmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y) if wl.cells is not None else None
                fpl = np.array([cm.getFootprintCost(footprint_world(wl.footprint, float(r["pose_x"]), float(r["pose_y"]),
                                float(r["pose_yaw"]))) == 1.0 for r in sub]) if cm is not None else None
                Jg = oracle.objective_batch(pm, cm, sub, plan.astype(np.float64), fp_lethal=fpl)
                dJ = Jg - refd["J"]
                tol = float(wl.params["opt_tolerance"])
                du = np.abs(plan[:, :3].astype(np.float64) - refd["x"][:, :3]).max(axis=1)
                line["cost_residual"] = {
                    "definition": "J_gpu - J_scipy(ftol=opt_tolerance), float64 oracle objective, same problems",
                    "problems": int(k), "median": float(np.median(dJ)), "p99": float(np.percentile(dJ, 99)),
                    "max": float(dJ.max()), "frac_worse_than_1e-4": float((dJ > 1e-4).mean()),
                    "frac_worse_than_opt_tol": float((dJ > tol).mean()),
                    "first_control_abs_diff_median": float(np.median(du)),
                    "first_control_abs_diff_p90": float(np.percentile(du, 90))}
                if "x_tight" in refd:
                    # the same comparison against the best tightly converged reference optimum known (ftol 1e-10; cold
                    # start, from the ftol = opt_tolerance point, from the GPU's point): separates the solver's own error
                    # from scipy's early stop.  Problems where the GPU's plan is cheaper than that optimum by more than
                    # 1e-5 say nothing about velocities (the reference sits in a worse basin): counted, not compared.
                    kt = len(refd["J_tight"])
                    gap = Jg[:kt] - refd["J_tight"]
                    keep = gap >= -1e-5
                    dut = np.abs(plan[:kt, :3].astype(np.float64) - refd["x_tight"][:, :3]).max(axis=1)[keep]
                    dus = np.abs(refd["x"][:kt, :3] - refd["x_tight"][:, :3]).max(axis=1)
                    line["cost_residual"].update({
                        "tight_problems": int(kt), "gpu_plan_cheaper_than_tight_scipy": int((~keep).sum()),
                        "J_gpu_minus_J_scipy_tight_median": float(np.median(gap)),
                        "J_gpu_minus_J_scipy_tight_p99": float(np.percentile(gap, 99)),
                        "J_gpu_minus_J_scipy_tight_max": float(gap.max()),
                        "first_control_vs_tight_scipy_median": float(np.median(dut)),
                        "first_control_vs_tight_scipy_p90": float(np.percentile(dut, 90)),
                        "first_control_vs_tight_scipy_p99": float(np.percentile(dut, 99)),
                        "scipy_at_opt_tolerance_vs_tight_scipy_median": float(np.median(dus)),
                        "scipy_at_opt_tolerance_vs_tight_scipy_p90": float(np.percentile(dus, 90))})
            except Exception as exc:  # keep the GPU numbers even if the CPU arm failed
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                        "sample": f"reference arm failed: {exc}: {res.stderr[-300:]}"}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    solver.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c2", "c3", "c4", "c5"])
    ap.add_argument("--batch", type=int, default=0, help="override the per-GPU batch")
    ap.add_argument("--lanes", type=int, default=0, help="lanes per instance (0 = auto)")
    ap.add_argument("--footprint-mode", type=int, default=0, choices=[0, 1],
                    help="0 = the reference's static footprint term (default, parity mode); 1 = opt-in moving footprint "
                         "(SURVEY 8f row N1; no CPU baseline / cost residual for it)")
    ap.add_argument("--costmap-mode", type=int, default=0, choices=[0, 1],
                    help="0 = the reference's nearest-cell costmap term (default, parity mode); 1 = opt-in bilinear term "
                         "with gradient (SURVEY 8f row N4; no CPU baseline / cost residual for it)")
    ap.add_argument("--costmap-guidance", type=int, default=0, choices=[0, 1],
                    help="0 = costmap guidance on (default); 1 = off: solve on the reference's objective from the start "
                         "(the round-1 strategy; A/B measurements)")
    ap.add_argument("--sustained-s", type=float, default=2.0, help="length of the back-to-back run of the `sustained` record")
    ap.add_argument("--cpu-sample", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--dump-ref", default="", help="reference arm: save per-problem J and x of the last pass (npz)")
    ap.add_argument("--gpu-plan", default="", help="reference arm: npz with the GPU's plans, a third start of the tight solves")
    ap.add_argument("--port", action="store_true", help="reference arm: time the oracle port instead of oracle/_ref")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
