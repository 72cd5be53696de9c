"""GPU box: time neompc_build_requests_device (row N2) on a fleet, and the fused pipeline ticks -> requests -> solve."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_mpc_planner2_b200 import workloads
from neo_mpc_planner2_b200.abi import TICK_DTYPE, REQUEST_DTYPE, RESPONSE_DTYPE, CARROT_INFO_DTYPE, README_SAMPLE
from neo_mpc_planner2_b200.solver import BatchSolver

n = int(sys.argv[1]) if len(sys.argv) > 1 else 800000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
wl = workloads.config("c5", batch=64)
s_ = np.linspace(0, 1, L)
plan = np.stack([-45 + 90 * s_, 30 * np.sin(3 * np.pi * s_), np.zeros(L)], 1)
plan[:, 2] = np.arctan2(np.gradient(plan[:, 1]), np.gradient(plan[:, 0]))
rng = np.random.default_rng(0)
k = rng.integers(0, L, n)
t = np.zeros(n, TICK_DTYPE)
t["pose_x"] = plan[k, 0] + rng.uniform(-0.3, 0.3, n); t["pose_y"] = plan[k, 1] + rng.uniform(-0.3, 0.3, n)
t["pose_yaw"] = plan[k, 2] + rng.uniform(-0.5, 0.5, n); t["plan_start"] = np.maximum(0, k - 20); t["delta_t"] = 1 / 30
dev = torch.device("cuda", 0)
with BatchSolver(dict(wl.params), device=0) as s:
    s.load_workload(wl); s.set_plan(plan)
    cp = s.carrot_params(0.4, 0.4, 0.4, 30.0)
    d_t = torch.from_numpy(t.view(np.uint8).reshape(n, TICK_DTYPE.itemsize)).to(dev)
    d_r = torch.empty((n, REQUEST_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_i = torch.empty((n, CARROT_INFO_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_o = torch.empty((n, RESPONSE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    st = torch.cuda.Stream(dev); torch.cuda.set_stream(st)
    def carrots(): s.build_requests_device(cp, d_t.data_ptr(), n, d_r.data_ptr(), d_i.data_ptr(), stream=st.cuda_stream)
    def solve(): s.solve_device(d_r.data_ptr(), n, d_o.data_ptr(), None, None, st.cuda_stream)
    for _ in range(3): carrots(); solve()
    torch.cuda.synchronize()
    res = {}
    for name, fn in (("build_requests", carrots), ("build_requests+solve", lambda: (carrots(), solve()))):
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in ev:
            a.record(st); fn(); b.record(st)
        torch.cuda.synchronize()
        ms = float(np.mean([a.elapsed_time(b) for a, b in ev]))
        res[name] = {"ms": ms, "robots_per_s": n / ms * 1e3}
    bytes_alg = n * (48 + 64 + 16) + L * 24 + wl.cells.size
    res["build_requests"]["algorithmic_GBps"] = bytes_alg / res["build_requests"]["ms"] / 1e6
    print(json.dumps({"robots": n, "plan_poses": L, **res}))
