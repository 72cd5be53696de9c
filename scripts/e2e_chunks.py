"""Tuning aid (GPU box): e2e time of neompc_solve_batch_twists / neompc_solve_batch on C3 for NEOMPC_CHUNKS = 1..8."""
import os, sys, time, subprocess, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if len(sys.argv) > 1:
    import numpy as np, torch
    from neo_mpc_planner2_b200 import workloads
    from neo_mpc_planner2_b200.abi import REQUEST_DTYPE, RESPONSE_DTYPE
    from neo_mpc_planner2_b200.solver import BatchSolver
    wl = workloads.config("c3")
    n = wl.batch
    s = BatchSolver(wl.params)
    s.load_workload(wl)
    req = torch.from_numpy(wl.requests.view(np.uint8).reshape(n, 64)).pin_memory()
    tw = torch.empty((n, 3), dtype=torch.float32).pin_memory()
    rs = torch.empty((n, 32), dtype=torch.uint8).pin_memory()
    out = {}
    fns = (("twists", lambda: s.solve_twists_raw(req.data_ptr(), n, tw.data_ptr())),
           ("full", lambda: s.solve_raw(req.data_ptr(), n, rs.data_ptr())))
    for _ in range(300): fns[1][1]()                      # bring the board to its working clocks
    for rep in range(2):
        for name, fn in fns:
            for _ in range(20): fn()
            t0 = time.perf_counter()
            for _ in range(100): fn()
            out[f"{name}{rep}"] = round((time.perf_counter() - t0) / 100 * 1e3, 4)
    d = torch.empty((n, 64), dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(50): d.copy_(req, non_blocking=True)
    torch.cuda.synchronize()
    out["h2d_4MiB_ms"] = round((time.perf_counter() - t0) / 50 * 1e3, 4)
    t0 = time.perf_counter()
    for _ in range(50): rs.copy_(d[:, :32].contiguous(), non_blocking=True)
    torch.cuda.synchronize()
    out["d2h_2MiB_ms"] = round((time.perf_counter() - t0) / 50 * 1e3, 4)
    print(json.dumps({"chunks": os.environ.get("NEOMPC_CHUNKS"), **out}))
else:
    # first the zero-copy path (pinned buffers: the default), then the staged-copy pipeline with different chunkings
    for label, extra in (("zero-copy (default for pinned buffers)", {}), ("staged copies, 3 chunks 1:4:6", {"NEOMPC_NO_ZEROCOPY": "1"}),
                         ("staged copies, 3 equal chunks", {"NEOMPC_NO_ZEROCOPY": "1", "NEOMPC_CHUNKS": "3", "NEOMPC_CHUNK_WEIGHTS": "1,1,1"}),
                         ("staged copies, 1 chunk", {"NEOMPC_NO_ZEROCOPY": "1", "NEOMPC_CHUNKS": "1"})):
        r = subprocess.run([sys.executable, __file__, "x"], env=dict(os.environ, **extra), capture_output=True, text=True)
        print(label, r.stdout.strip() or r.stderr[-300:])
