"""GPU box: latency of ONE request through neompc_solve_msgs for different lane-group sizes (control_steps 10 and 3)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_mpc_planner2_b200 import workloads
from neo_mpc_planner2_b200.server import requests_to_msgs
from neo_mpc_planner2_b200.solver import BatchSolver

for cfg in ("c3", "c2"):
    wl = workloads.config(cfg, batch=64)
    for lanes in (0, 4, 8, 16, 32):
        with BatchSolver(wl.params, lanes_per_instance=lanes) as s:
            s.load_workload(wl)
            s.reserve_instances(1)
            res = []
            for k in range(8):
                msg = requests_to_msgs(wl.requests[k:k + 1]); msg["instance_id"] = 0; msg["delta_t"] = 1 / 30
                for _ in range(5):
                    s.solve_msgs(msg)
                ts = []
                for _ in range(100):
                    s.reset_state()
                    t = time.perf_counter(); out = s.solve_msgs(msg); ts.append(time.perf_counter() - t)
                res.append(np.median(ts))
            print(cfg, "N", s.control_steps, "lanes", lanes, "tiling", s.tiling, "median latency over 8 problems %.1f us (min %.1f max %.1f) iters %d"
                  % (1e6 * np.mean(res), 1e6 * min(res), 1e6 * max(res), out["iters"][0]))
