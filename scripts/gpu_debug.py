"""Scratch diagnostics run on the GPU box (not part of the product)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from neo_mpc_planner2_b200 import workloads
from neo_mpc_planner2_b200.solver import BatchSolver
from tests.hostsim import HostSim

what = sys.argv[1]
if what == "lanes":
    wl = workloads.config("c3", batch=1024)
    hs = HostSim(wl.params, wl.cells, wl.resolution, (wl.origin_x, wl.origin_y), footprint=wl.footprint)
    oh, ph = hs.solve(wl.requests)
    print("hostsim iters med", np.median(oh["iters"]), "evals mean", oh["evals"].mean(), "status", np.bincount(oh["status"], minlength=3), "cost mean", oh["cost"].mean())
    for lanes in (2, 4, 8, 16, 32):
        with BatchSolver(wl.params, lanes_per_instance=lanes) as s:
            s.load_workload(wl)
            o, p = s.solve(wl.requests, want_plan=True)
            print("lanes", lanes, s.tiling, "iters med", np.median(o["iters"]), "evals mean", o["evals"].mean(), "status",
                  np.bincount(o["status"], minlength=3), "cost mean", o["cost"].mean(), "dcost vs hostsim p50/p99",
                  np.percentile(np.abs(o["cost"] - oh["cost"]), [50, 99]))
elif what == "n64":
    n = int(sys.argv[2])
    lanes = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    wl = workloads.config("c3", batch=256)
    wl.params["control_steps"] = n
    with BatchSolver(wl.params, lanes_per_instance=lanes) as s:
        s.load_workload(wl)
        print("tiling", s.tiling)
        U = np.zeros((256, 3 * n), np.float32)
        J, G = s.eval_objective(wl.requests, U)
        print("eval ok", J[:4])
        o = s.solve(wl.requests)
        print("solve ok", np.median(o["iters"]))
