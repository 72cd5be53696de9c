"""Shared helpers for the parity tests."""
from __future__ import annotations

import numpy as np

import oracle
from oracle.costmap import GridCostmap
from oracle.mpc_oracle import footprint_world
from neo_mpc_planner2_b200 import workloads


def setup_workload(name, batch, control_steps=None, **param_over):
    wl = workloads.config(name, batch=batch)
    params = dict(wl.params)
    if control_steps:
        params["control_steps"] = control_steps
    params.update(param_over)
    wl.params = params
    p = oracle.MpcParams(**params)
    cm = GridCostmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y) if wl.cells is not None else None
    return wl, p, cm


def footprint_lethal_flags(wl, cm, reqs=None):
    reqs = wl.requests if reqs is None else reqs
    if cm is None:
        return np.zeros(len(reqs), dtype=bool)
    return np.array([cm.getFootprintCost(footprint_world(wl.footprint, float(r["pose_x"]), float(r["pose_y"]),
                                                         float(r["pose_yaw"]))) == 1.0 for r in reqs])


def near_cell_edge(p, cm, reqs, U, tol_cells=2e-3):
    """True for problems whose rollout passes within tol_cells of a costmap cell edge: float32 vs float64
    rounding may put such a sample into the neighbouring cell, which is a legitimate, counted difference."""
    if cm is None:
        return np.zeros(len(reqs), dtype=bool)
    from oracle.mpc_oracle import rollout_batch
    _, _, _, px, py = rollout_batch(p, reqs, U)
    return (cm.edge_distance_cells(px, py) < tol_cells).any(axis=1)


def feasibility_violation(params, plan):
    """Largest violation of the box (srv.py:127-133) and disc (srv.py:157-158) constraints."""
    n = int(params["control_steps"])
    U = np.asarray(plan, dtype=np.float64).reshape(len(plan), n, 3)
    lo = np.array([params["min_vel_x"], params["min_vel_y"], params["min_vel_theta"]])
    hi = np.array([params["max_vel_x"], params["max_vel_y"], params["max_vel_theta"]])
    box = np.maximum(lo - U, U - hi).max()
    disc = (np.sqrt(U[:, :, 0] ** 2 + U[:, :, 1] ** 2) - params["max_vel_trans"]).max()
    return max(box, disc, 0.0)


def scipy_solutions(wl, p, cm, idx, tight=False):
    """Reference solves (oracle.slsqp_solve == srv.py:363-364) for the listed problems."""
    from oracle.costmap import FreeSpaceCostmap
    cmo = cm if cm is not None else FreeSpaceCostmap()
    out = []
    for i in idx:
        prob = oracle.Problem.from_record(wl.requests[i])
        fpw = footprint_world(wl.footprint, prob.pose_x, prob.pose_y, prob.pose_yaw)
        res = oracle.slsqp_solve(p, cmo, fpw, prob)
        if tight:
            a = oracle.slsqp_solve(p, cmo, fpw, prob, x0=res.x, ftol=1e-10, maxiter=400)
            b = oracle.slsqp_solve(p, cmo, fpw, prob, ftol=1e-10, maxiter=400)
            out.append((res, a if a.fun <= b.fun else b))
        else:
            out.append((res, None))
    return out


# ---------------------------------------------------------------------------------------------------------------
# scipy reference solves in worker processes (spawned: safe next to a CUDA context in the parent)
# ---------------------------------------------------------------------------------------------------------------
_PW = {}


def _pw_init(name, batch, control_steps, param_over, nomap):
    wl, p, cm = setup_workload(name, batch, control_steps, **param_over)
    if nomap:
        wl.cells, cm = None, None
    from oracle.costmap import FreeSpaceCostmap
    _PW.update(wl=wl, p=p, cm=cm if cm is not None else FreeSpaceCostmap())


def _pw_solve(job):
    """job = (index, tight, x_extra): the reference's solve (srv.py:363-364) at ftol = opt_tolerance, and optionally the
    best of three tightly converged solves (cold start, from the ftol = opt_tolerance point, from x_extra)."""
    i, tight, x_extra = job
    wl, p, cm = _PW["wl"], _PW["p"], _PW["cm"]
    prob = oracle.Problem.from_record(wl.requests[i])
    fpw = footprint_world(wl.footprint, prob.pose_x, prob.pose_y, prob.pose_yaw)
    res = oracle.slsqp_solve(p, cm, fpw, prob)
    out = {"i": i, "fun": float(res.fun), "x": np.asarray(res.x, dtype=np.float64)}
    if tight:
        cands = [oracle.slsqp_solve(p, cm, fpw, prob, x0=res.x, ftol=1e-10, maxiter=400),
                 oracle.slsqp_solve(p, cm, fpw, prob, ftol=1e-10, maxiter=400)]
        if x_extra is not None:
            cands.append(oracle.slsqp_solve(p, cm, fpw, prob, x0=np.asarray(x_extra, dtype=np.float64), ftol=1e-10,
                                            maxiter=400))
        best = min(cands, key=lambda r: r.fun)
        out.update(fun_tight=float(best.fun), x_tight=np.asarray(best.x, dtype=np.float64))
    return out


def scipy_reference(name, batch, control_steps, idx, tight=False, plans=None, param_over=None, nomap=False, workers=None):
    """Reference solves for problems `idx` of workloads.config(name, batch), spread over worker SUBPROCESSES
    (`python -m tests.refworker`, fresh interpreters: no fork next to a CUDA context).  `plans`: [len(idx), 3N] extra
    starting points of the tight solves (the GPU's solutions).  Returns a list of dicts in the order of idx."""
    import os
    import pickle
    import subprocess
    import sys
    from concurrent.futures import ThreadPoolExecutor
    idx = list(idx)
    jobs = [(i, tight, None if plans is None else np.asarray(plans[k])) for k, i in enumerate(idx)]
    workers = workers or max(1, min(len(jobs), (os.cpu_count() or 2), 16))
    chunks = [jobs[w::workers] for w in range(workers)]
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", PYTHONPATH=root)

    def run(chunk):
        payload = pickle.dumps(((name, batch, control_steps, dict(param_over or {}), nomap), chunk))
        res = subprocess.run([sys.executable, "-m", "tests.refworker"], input=payload, capture_output=True, cwd=root,
                             env=env, timeout=3000)
        if res.returncode != 0:
            raise RuntimeError(res.stderr.decode()[-2000:])
        return pickle.loads(res.stdout)

    with ThreadPoolExecutor(workers) as ex:
        parts = list(ex.map(run, chunks))
    by_i = {}
    for part in parts:
        for out in part:
            by_i[out["i"]] = out
    return [by_i[i] for i in idx]


def expected_edge_fraction(n_steps, tol_cells=2e-3):
    """Share of random plans with a rollout point within tol_cells of a cell edge (either axis): 1 - (1 - 4 tol)^N."""
    return 1.0 - (1.0 - 4.0 * tol_cells) ** n_steps


def residual_stats(dJ, tol):
    dJ = np.asarray(dJ, dtype=np.float64)
    return dict(n=len(dJ), median=float(np.median(dJ)), p99=float(np.percentile(dJ, 99)), max=float(dJ.max()),
                worse_1e4=float((dJ > 1e-4).mean()), worse_tol=float((dJ > tol).mean()))


def first_control_distance(plans, Jg, refs):
    """|u0_gpu - u0_best| (sup norm over vx, vy, omega) against the best tightly converged reference optimum known, for the
    problems where that optimum is at least as good as the GPU's plan (gap >= -1e-5); the others say nothing about
    velocities — the reference sits in a worse basin of the costmap staircase — and are counted.  Returns (du, n_better)."""
    gap = np.array([Jg[k] - r["fun_tight"] for k, r in enumerate(refs)])
    du = np.array([np.abs(np.asarray(plans[k][:3], dtype=np.float64) - r["x_tight"][:3]).max() for k, r in enumerate(refs)])
    keep = gap >= -1e-5
    return du[keep], int((~keep).sum()), gap
