"""CPU oracle for the neo_mpc_planner2 hot path (TEST INFRASTRUCTURE ONLY).

This package restates, in plain numpy / scipy, the algorithm of the reference's
``mpc_optimization_server.py`` (objective, constraint, SLSQP call, per-tick state
machine) so that the CUDA path can be checked against it on identical inputs.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it.  Nothing under ``neo_mpc_planner2_b200/``
imports it, and the product path raises if its CUDA library is missing.

Parity status (see DESIGN.md "Oracle"):
  * objective / constraint / state machine: PINNED bit-exactly against the unmodified
    reference file, imported under ROS stub modules in the build container
    (``tests/golden/make_golden.py`` -> ``tests/golden/*.json``).
  * scipy SLSQP: third-party (reference README pins scipy 1.6.3; this image has 1.18.1);
    the reference has no tests or golden vectors for it.  Pinned only against outputs of
    the reference run here with scipy 1.18.1.
  * costmap (``neo_nav2_py_costmap2D``): third-party, absent, unversioned ->
    semantics DECLARED in ``oracle/costmap.py``; parity at that boundary is unpinned.
"""
from .costmap import GridCostmap, ENC_OCCUPANCY, ENC_NAV2_RAW  # noqa: F401
from .mpc_oracle import (  # noqa: F401
    MpcParams, Problem, OracleServer, objective, f_constraint, euler_yaw, quirk_yaw,
    slsqp_solve, objective_batch, gradient_batch, collision_check, initial_guess_update,
    quat_from_yaw, local_plan, footprint_at, moving_footprint_lethal,
)
