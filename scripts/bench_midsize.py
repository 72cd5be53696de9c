"""GPU box: kernel time of mid-size batches (device under-filled) by lane tiling; CUDA events, requests resident."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_mpc_planner2_b200 import workloads
from neo_mpc_planner2_b200.abi import REQUEST_DTYPE, RESPONSE_DTYPE
from neo_mpc_planner2_b200.solver import BatchSolver
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(dev); torch.cuda.set_stream(stream)
for cfg, sizes in (("c3", (256, 1000, 4096, 8192, 16384)), ("c2", (1000, 4096, 16384))):
    for n in sizes:
        wl = workloads.config(cfg, batch=n)
        row = []
        for lanes in ((0, 4, 8, 16) if cfg == "c3" else (0, 1, 2, 4)):
            with BatchSolver(wl.params, lanes_per_instance=lanes) as s:
                s.load_workload(wl)
                d_reqs = torch.from_numpy(wl.requests.view(np.uint8).reshape(n, REQUEST_DTYPE.itemsize)).to(dev)
                d_out = torch.empty((n, RESPONSE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
                for _ in range(5):
                    s.solve_device(d_reqs.data_ptr(), n, d_out.data_ptr(), None, None, stream.cuda_stream)
                torch.cuda.synchronize()
                ms = []
                for _ in range(20):
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record(); s.solve_device(d_reqs.data_ptr(), n, d_out.data_ptr(), None, None, stream.cuda_stream); b.record()
                    torch.cuda.synchronize(); ms.append(a.elapsed_time(b))
                row.append("lanes %s %s: %.1f us" % (lanes if lanes else "auto", s.tiling if lanes else "", 1e3 * float(np.median(ms))))
        print(cfg, "N", wl.control_steps, "n", n, " | ".join(row))
