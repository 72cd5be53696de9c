"""The oracle restatement vs the golden vectors frozen from the UNMODIFIED reference
(tests/golden/make_golden.py).  Bit-exact (==) in float64: objective (srv.py:204-269),
f_constraint (:157-158), the SLSQP call (:363-364) and the optimizer() state machine (:349-403)."""
import hashlib

import numpy as np
import pytest

import oracle
from oracle.costmap import GridCostmap, FreeSpaceCostmap, bresenham_cells
from oracle.mpc_oracle import REQUEST_FIELDS
from neo_mpc_planner2_b200 import workloads


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _prob(fin):
    p = {k: fin[k] for k in REQUEST_FIELDS}
    p["pose_yaw"] = fin["pose_yaw_true"]
    return p


def _grid(case):
    cells = np.random.default_rng(case["grid_seed"]).integers(0, 101, (200, 200)).astype(np.uint8)
    if case["grid_all_lethal"]:
        cells[:, :] = 100
    assert _sha(cells) == case["grid_sha256"], "numpy Generator stream changed: regenerate the goldens"
    return GridCostmap(cells, 0.05, -5.0, -5.0)


def test_known_answers(golden):
    k = golden["kat"]
    p = oracle.MpcParams(**k["params"])
    fpw = oracle.mpc_oracle.footprint_world(k["footprint_robot"], k["problem"]["pose_x"],
                                            k["problem"]["pose_y"], k["problem"]["pose_yaw_true"])
    cm = FreeSpaceCostmap()
    assert oracle.objective(p, cm, fpw, _prob(k["problem"]), np.zeros(9)) == k["J_zero"] == 0.5010200000000001
    assert oracle.objective(p, cm, fpw, _prob(k["problem"]), np.array(k["u_probe"])) == k["J_probe"]
    res = oracle.slsqp_solve(p, cm, fpw, _prob(k["problem"]))
    assert res.x.tolist() == k["slsqp"]["x"]
    assert float(res.fun) == k["slsqp"]["fun"]
    assert (res.nit, res.nfev, res.status) == (k["slsqp"]["nit"], k["slsqp"]["nfev"], k["slsqp"]["status"])
    srv = oracle.OracleServer(p, cm, k["footprint_robot"])
    out = srv.tick(_prob(k["problem"]))
    assert list(out) == k["first_tick"]["output"]
    assert srv.initial_guess.tolist() == k["first_tick"]["next_initial_guess"]


def test_objective_and_constraint_bit_exact(golden):
    for case in golden["objective_cases"]:
        p = oracle.MpcParams(**case["params"])
        cm = _grid(case)
        u = np.array(case["u"])
        fpw = [tuple(v) for v in case["footprint_world"]]
        assert oracle.objective(p, cm, fpw, _prob(case["problem"]), u) == case["J"]
        assert [float(oracle.f_constraint(p, u, i)) for i in range(p.control_steps)] == case["constraints"]
        assert cm.getFootprintCost(fpw) == case["footprint_cost"]


def test_objective_batch_matches_scalar(golden):
    for case in golden["objective_cases"]:
        p = oracle.MpcParams(**case["params"])
        cm = _grid(case)
        prob = _prob(case["problem"])
        reqs = {k: np.array([v]) for k, v in prob.items()}
        J = oracle.objective_batch(p, cm, reqs, np.array([case["u"]]),
                                   fp_lethal=[case["footprint_cost"] == 1.0])[0]
        assert J == pytest.approx(case["J"], rel=1e-12, abs=1e-12)


def test_slsqp_cases_bit_exact(golden):
    wl = workloads.config("c2", batch=64)
    assert _sha(wl.cells) == golden["c2_grid_sha256"], "C2 costmap recipe changed: regenerate the goldens"
    cm = GridCostmap(wl.cells, wl.resolution, wl.origin_x, wl.origin_y)
    for case in golden["slsqp_cases"]:
        p = oracle.MpcParams(**case["params"])
        fpw = [tuple(v) for v in case["footprint_world"]]
        res = oracle.slsqp_solve(p, cm, fpw, _prob(case["problem"]))
        assert res.x.tolist() == case["x"]
        assert float(res.fun) == case["fun"]
        assert (res.nit, res.nfev, res.status, bool(res.success)) == (
            case["nit"], case["nfev"], case["status"], case["success"])


def _sequence_costmap(seq):
    if seq["grid"] is None:
        return FreeSpaceCostmap()
    if seq["name"].startswith("c2map"):
        wl = workloads.config("c2", batch=64)
        cells = wl.cells
    elif seq["name"].startswith("wall"):
        cells = np.zeros((200, 200), dtype=np.uint8)
        cells[:, 112:] = 99
        cells[:, 116:] = 100
    elif seq["name"].startswith("footprint"):
        cells = np.zeros((200, 200), dtype=np.uint8)
        cells[100:104, 106:110] = 100
    else:
        raise KeyError(seq["name"])
    assert _sha(cells) == seq["grid"]["sha256"]
    return GridCostmap(cells, 0.05, seq["origin"][0], seq["origin"][1])


def test_tick_sequences_bit_exact(golden):
    for seq in golden["tick_sequences"]:
        p = oracle.MpcParams(**seq["params"])
        srv = oracle.OracleServer(p, _sequence_costmap(seq), seq["footprint_robot"])
        for t in seq["ticks"]:
            out = srv.tick(_prob(t["problem"]))
            assert list(out) == t["output"], seq["name"]
            assert srv.initial_guess.tolist() == t["initial_guess_after"]
            assert (srv.collision, srv.collision_footprint) == (t["collision"], t["collision_footprint"])
            assert float(srv.waiting_time) == t["waiting_time"]


def test_gradient_batch_matches_finite_differences():
    rng = np.random.default_rng(7)
    for n in (3, 10, 20):
        wl = workloads.config("c2", batch=8)
        p = oracle.MpcParams.readme_sample(control_steps=n)
        U = rng.uniform(-0.6, 0.6, (8, 3 * n))
        G = oracle.gradient_batch(p, wl.requests, U)
        h = 1e-6
        for j in range(3 * n):
            Up, Um = U.copy(), U.copy()
            Up[:, j] += h
            Um[:, j] -= h
            fd = (oracle.objective_batch(p, None, wl.requests, Up)
                  - oracle.objective_batch(p, None, wl.requests, Um)) / (2 * h)
            assert np.allclose(G[:, j], fd, atol=2e-8, rtol=1e-6)


def test_bresenham_closed_form():
    rng = np.random.default_rng(3)
    for _ in range(200):
        x0, y0, x1, y1 = (int(v) for v in rng.integers(-30, 30, 4))
        pts = list(bresenham_cells(x0, y0, x1, y1))
        dx, dy = abs(x1 - x0), abs(y1 - y0)
        sx = 1 if x1 >= x0 else -1
        sy = 1 if y1 >= y0 else -1
        assert pts[0] == (x0, y0) and pts[-1] == (x1, y1) and len(pts) == max(dx, dy) + 1
        for k, (x, y) in enumerate(pts):
            if dx >= dy:
                assert (x, y) == (x0 + k * sx, y0 + sy * ((dx // 2 + k * dy) // dx if dx else 0))
            else:
                assert (x, y) == (x0 + sx * ((dy // 2 + k * dx) // dy), y0 + k * sy)
