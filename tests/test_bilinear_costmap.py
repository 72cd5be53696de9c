"""NEOMPC_COSTMAP_BILINEAR: the costmap term interpolated between cell centres, with its gradient.

Opt-in and deliberately NOT the reference's behaviour (the reference reads the cell under the predicted position:
costmap_mode = 0, which every parity test covers).  The oracle restates the mode in float64
(``oracle.objective(..., bilinear=True)``, ``GridCostmap.bilinear_at_world``); its analytic gradient is checked against
central differences, the device against the oracle: value <= 2e-5 relative, gradient <= 2e-5 relative to max(1, |g|_inf).
"""
import numpy as np
import pytest

import oracle
from oracle.mpc_oracle import footprint_world, rollout_batch
from neo_mpc_planner2_b200.abi import COSTMAP_BILINEAR
from tests.util import setup_workload, footprint_lethal_flags, feasibility_violation

SMOOTH = 1e-2


def test_oracle_bilinear_properties():
    wl, p, cm = setup_workload("c3", 64, 10)
    # at cell centres the bilinear term equals the reference's term
    mx, my = np.meshgrid(np.arange(5, 995, 37), np.arange(7, 995, 41))
    wx = wl.origin_x + (mx + 0.5) * wl.resolution
    wy = wl.origin_y + (my + 0.5) * wl.resolution
    c, l = cm.bilinear_at_world(wx, wy)[:2]
    cn = cm.cost_at_world(wx, wy)
    assert np.abs(c - cn).max() <= 1e-9 and np.abs(l - (cn == 1.0)).max() <= 1e-9
    ref_term = np.where(cn == 1.0, 1000.0, p.w_costmap) * cn ** 2
    assert np.abs(p.w_costmap * c ** 2 + (1000.0 - p.w_costmap) * l ** 2 - ref_term).max() <= 1e-5
    # analytic gradient == central differences of the bilinear objective
    rng = np.random.default_rng(4)
    U = rng.uniform(-0.7, 0.7, (8, 30))
    G = oracle.gradient_batch(p, wl.requests[:8], U, eps_control=0.0, bilinear_costmap=cm)
    h = 1e-6
    for b in range(8):
        fd = np.array([(oracle.objective_batch(p, cm, wl.requests[b:b + 1], (U[b] + h * e)[None], bilinear=True)[0]
                        - oracle.objective_batch(p, cm, wl.requests[b:b + 1], (U[b] - h * e)[None], bilinear=True)[0]) / (2 * h)
                       for e in np.eye(30)])
        assert np.abs(fd - G[b]).max() <= 1e-6 * max(1.0, np.abs(fd).max())
    # scalar restatement == vectorised restatement
    prob = oracle.Problem.from_record(wl.requests[0])
    fpw = footprint_world(wl.footprint, prob.pose_x, prob.pose_y, prob.pose_yaw)
    fpl = footprint_lethal_flags(wl, cm, wl.requests[:1])
    assert abs(oracle.objective(p, cm, fpw, prob, U[0], bilinear=True)
               - oracle.objective_batch(p, cm, wl.requests[:1], U[:1], fp_lethal=fpl, bilinear=True)[0]) <= 1e-9


def check_mode(evaluate, solve_bilinear, solve_nearest, wl, p, cm):
    rng = np.random.default_rng(31)
    n = p.control_steps
    U = rng.uniform(-0.7, 0.7, (wl.batch, 3 * n)).astype(np.float32)
    fpl = footprint_lethal_flags(wl, cm)
    J, G = evaluate(wl.requests, U)
    Jo = oracle.objective_batch(p, cm, wl.requests, U.astype(np.float64), fp_lethal=fpl, bilinear=True)
    Go = oracle.gradient_batch(p, wl.requests, U.astype(np.float64), eps_control=SMOOTH, bilinear_costmap=cm)
    assert (np.abs(J - Jo) / np.maximum(1.0, np.abs(Jo))).max() <= 2e-5     # smooth: no cell-edge exclusions needed
    # near lethal cells the (1000 - w_costmap) l^2 barrier makes gradients of O(100): compare relative to the row's scale
    assert (np.abs(G - Go).max(axis=1) <= 2e-5 * np.maximum(1.0, np.abs(Go).max(axis=1))).all()
    out, plan = solve_bilinear(wl.requests)
    assert feasibility_violation(wl.params, plan) <= 1e-6
    Jb = oracle.objective_batch(p, cm, wl.requests, plan.astype(np.float64), fp_lethal=fpl, bilinear=True)
    assert np.abs(out["cost"] - Jb).max() <= 2e-5 * max(1.0, np.abs(Jb).max())
    _, plan_n = solve_nearest(wl.requests)
    Jn = oracle.objective_batch(p, cm, wl.requests, plan_n.astype(np.float64), fp_lethal=fpl, bilinear=True)
    assert (Jb <= Jn + 1e-4).mean() >= 0.9 and Jb.mean() < Jn.mean()
    # the gradient steers plans away from inflated obstacles: lower mean cell cost along the plan
    cost_along = []
    for pl in (plan, plan_n):
        _, _, _, px, py = rollout_batch(p, wl.requests, pl.astype(np.float64))
        cost_along.append(cm.cost_at_world(px, py).mean())
    assert cost_along[0] < cost_along[1]


def test_hostsim_bilinear():
    from tests.hostsim import HostSim
    wl, p, cm = setup_workload("c3", 256, 10, w_costmap=0.5)
    env = (wl.params, wl.cells, wl.resolution, (wl.origin_x, wl.origin_y))
    hb = HostSim(*env, footprint=wl.footprint, costmap_mode=COSTMAP_BILINEAR)
    hn = HostSim(*env, footprint=wl.footprint)
    check_mode(hb.eval, hb.solve, hn.solve, wl, p, cm)


def test_hostsim_bilinear_vs_scipy():
    """The same smooth NLP given to the reference's optimizer call (SLSQP with finite differences sees the gradient too)."""
    from tests.hostsim import HostSim
    wl, p, cm = setup_workload("c3", 64, 10, w_costmap=0.5)
    hb = HostSim(wl.params, wl.cells, wl.resolution, (wl.origin_x, wl.origin_y), footprint=wl.footprint,
                 costmap_mode=COSTMAP_BILINEAR)
    out, plan = hb.solve(wl.requests[:10])
    fpl = footprint_lethal_flags(wl, cm, wl.requests[:10])
    Jb = oracle.objective_batch(p, cm, wl.requests[:10], plan.astype(np.float64), fp_lethal=fpl, bilinear=True)
    dJ = []
    for i in range(10):
        prob = oracle.Problem.from_record(wl.requests[i])
        fpw = footprint_world(wl.footprint, prob.pose_x, prob.pose_y, prob.pose_yaw)
        dJ.append(Jb[i] - float(oracle.slsqp_solve(p, cm, fpw, prob, bilinear=True).fun))
    dJ = np.array(dJ)
    assert np.median(dJ) <= 0.0 and dJ.max() <= 5 * p.opt_tolerance


@pytest.mark.gpu
@pytest.mark.parametrize("cfg,n_steps,lanes,enc", [("c3", 10, 0, 0), ("c3", 10, 8, 0), ("c2", 3, 0, 0), ("c3", 20, 0, 1)])
def test_gpu_bilinear(cfg, n_steps, lanes, enc):
    from neo_mpc_planner2_b200.solver import BatchSolver
    from oracle.costmap import GridCostmap
    wl, p, cm = setup_workload(cfg, 512, n_steps, w_costmap=0.5)
    if enc == 1:                                   # nav2 raw bytes: 0..254, lethal = 254
        raw = np.where(wl.cells == 100, 254, np.minimum(252, (wl.cells.astype(np.int32) * 253) // 100)).astype(np.uint8)
        wl.cells, wl.encoding = raw, 1
        cm = GridCostmap(raw, wl.resolution, wl.origin_x, wl.origin_y, encoding=1)
    with BatchSolver(wl.params, lanes_per_instance=lanes, costmap_mode=COSTMAP_BILINEAR) as sb, \
            BatchSolver(wl.params, lanes_per_instance=lanes) as sn:
        sb.load_workload(wl)
        sn.load_workload(wl)
        check_mode(sb.eval_objective, lambda r: sb.solve(r, want_plan=True), lambda r: sn.solve(r, want_plan=True), wl, p, cm)
