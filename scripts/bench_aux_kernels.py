"""GPU box: time the small kernels around the solve at C3 size (65536 instances, control_steps 10): pack_kernel
(neompc_pack_requests) and local_plan_kernel (neompc_local_plan_device).  CUDA events on the launch stream, L2 flushed
between iterations; prints one JSON object."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from neo_mpc_planner2_b200 import workloads
from neo_mpc_planner2_b200.abi import REQUEST_DTYPE, RESPONSE_DTYPE, MSG_DTYPE, PLAN_POSE_DTYPE
from neo_mpc_planner2_b200.server import requests_to_msgs
from neo_mpc_planner2_b200.solver import BatchSolver

n = 65536
wl = workloads.config("c3", batch=n)
N = wl.control_steps
dev = torch.device("cuda", 0)
peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")) else {}
peak = float(peaks.get("hbm_gbs", 6650.0))
out = {"n": n, "control_steps": N, "hbm_peak_gbs": peak}
with BatchSolver(wl.params) as s:
    s.load_workload(wl)
    msgs = requests_to_msgs(wl.requests)
    d_msgs = torch.from_numpy(msgs.view(np.uint8).reshape(n, MSG_DTYPE.itemsize)).to(dev)
    d_reqs = torch.empty((n, REQUEST_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_out = torch.empty((n, RESPONSE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    d_plan = torch.empty((n, 3 * N), dtype=torch.float32, device=dev)
    d_poses = torch.empty((n, N + 1, PLAN_POSE_DTYPE.itemsize), dtype=torch.uint8, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.Stream(dev)          # a real (non-default) stream: events and kernels share it
    torch.cuda.set_stream(stream)
    st = stream.cuda_stream

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ms = []
        for k in range(reps):
            flush.fill_(k & 0xFF)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ms.append(a.elapsed_time(b))
        return float(np.median(ms))

    t = timed(lambda: s.pack_requests_device(d_msgs.data_ptr(), n, d_reqs.data_ptr(), st))
    by = n * (MSG_DTYPE.itemsize + REQUEST_DTYPE.itemsize)
    out["pack_kernel"] = {"ms": t, "bytes_per_unit": MSG_DTYPE.itemsize + REQUEST_DTYPE.itemsize, "GBps": by / t / 1e6, "frac_of_hbm_peak": by / t / 1e6 / peak}
    s.solve_device(d_reqs.data_ptr(), n, d_out.data_ptr(), None, d_plan.data_ptr(), st)
    t = timed(lambda: s.local_plan_device(d_reqs.data_ptr(), d_plan.data_ptr(), n, d_poses.data_ptr(), st))
    bpu = REQUEST_DTYPE.itemsize + 12 * N + PLAN_POSE_DTYPE.itemsize * (N + 1)
    out["local_plan_kernel"] = {"ms": t, "bytes_per_unit": bpu, "GBps": n * bpu / t / 1e6, "frac_of_hbm_peak": n * bpu / t / 1e6 / peak}
print(json.dumps(out))
