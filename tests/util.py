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
