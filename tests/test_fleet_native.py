"""The native multi-GPU entry points of libneompc (include/neompc.h "multi-GPU"): NCCL bound at run time, one handle per
GPU, contiguous shards, one all-gather of (vx, vy, omega).  The single-rank case runs on any GPU box; the two-GPU cases need
`gpurun --gpus 2` (skipped otherwise).  bench.py --gpus N exercises the one-process-per-GPU form (FleetSolver) under torchrun."""
import numpy as np
import pytest

from tests.util import setup_workload

pytestmark = pytest.mark.gpu


def _gpus():
    import torch
    return torch.cuda.device_count()


def _single_gpu_twists(wl, reqs, lanes):
    from neo_mpc_planner2_b200.solver import BatchSolver
    with BatchSolver(wl.params, lanes_per_instance=lanes) as s:
        s.load_workload(wl)
        out = s.solve(reqs)
    return np.stack([out["vx"], out["vy"], out["omega"]], axis=1), out


def test_single_rank_communicator():
    """A fleet of one: the NCCL path (dlopen, communicator, in-place all-gather) on one GPU; bit-equal to the plain solve."""
    from neo_mpc_planner2_b200.fleet import LocalFleet
    wl, p, cm = setup_workload("c3", 5000, 10)
    want, out1 = _single_gpu_twists(wl, wl.requests, 4)
    with LocalFleet(wl.params, [0], lanes_per_instance=4) as fleet:
        fleet.load_workload(wl)
        assert fleet.solvers[0].comm_info() == (1, 0)
        twist, out = fleet.solve(wl.requests, want_responses=True)
        again = fleet.solve(wl.requests)
    assert twist.tobytes() == want.tobytes() and again.tobytes() == want.tobytes()
    assert out.tobytes() == out1.tobytes()


@pytest.mark.parametrize("n_total", [10001, 4096, 3])
def test_local_fleet_matches_single_gpu(n_total):
    """neompc_fleet_solve over two GPUs: the gathered twists equal the one-GPU solve of the same requests bit for bit, on
    both ranks' copies; ragged and nearly empty shards included."""
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from neo_mpc_planner2_b200.fleet import LocalFleet, shard_bounds
    wl, p, cm = setup_workload("c3", n_total, 10)
    want, out1 = _single_gpu_twists(wl, wl.requests, 4)
    with LocalFleet(wl.params, [0, 1], lanes_per_instance=4) as fleet:
        fleet.load_workload(wl)
        twist, out = fleet.solve(wl.requests, want_responses=True)
        other = fleet.gathered_on(1, n_total)
    assert twist.tobytes() == want.tobytes()
    assert other.tobytes() == want.tobytes()
    assert out.tobytes() == out1.tobytes()
    lo, hi = shard_bounds(n_total, 2, 1)
    assert hi == n_total and lo == (n_total + 1) // 2


def test_local_fleet_keeps_state_on_the_owning_gpu():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    from neo_mpc_planner2_b200.fleet import LocalFleet
    wl, p, cm = setup_workload("c3", 2000, 10)
    req = wl.requests.copy()
    req["instance_id"] = np.arange(len(req), dtype=np.uint32)
    with LocalFleet(wl.params, [0, 1]) as fleet:
        fleet.load_workload(wl)
        fleet.reserve_instances(len(req))
        _, first = fleet.solve(req, want_responses=True)
        _, second = fleet.solve(req, want_responses=True)
    assert (first["flags"] & 4).all() and not (second["flags"] & 4).any()       # new-goal reset once, warm start after
    assert second["iters"].mean() < first["iters"].mean()
