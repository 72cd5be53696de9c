"""neo_mpc_planner2_b200 — B200-native batched MPC solve behind the reference's Optimizer schema.

The package holds only what the hot path needs: ``csrc/`` (CUDA kernels + the C-ABI library
``libneompc.so``), a ctypes binding, and the host-side mirror of the reference's
``MpcOptimizationServer.optimizer`` interface.  There is no CPU fallback: importing
``neo_mpc_planner2_b200.solver`` raises if ``libneompc.so`` has not been built.
"""
__version__ = "0.1.0"
