"""ctypes driver of the host emulation of the device solver core (TEST TOOLING ONLY, see hostsim.cpp)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

from neo_mpc_planner2_b200.abi import REQUEST_DTYPE, RESPONSE_DTYPE, params_record

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
SO = os.path.join(HERE, "_build", "libneompc_hostsim.so")


def build(force=False):
    srcs = [os.path.join(HERE, "hostsim.cpp"),
            os.path.join(ROOT, "neo_mpc_planner2_b200", "csrc", "mpc_core.cuh"),
            os.path.join(ROOT, "neo_mpc_planner2_b200", "csrc", "mpc_setup.h"),
            os.path.join(ROOT, "include", "neompc.h")]
    if not force and os.path.exists(SO) and all(os.path.getmtime(SO) >= os.path.getmtime(s) for s in srcs):
        return SO
    os.makedirs(os.path.dirname(SO), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"),
                           "-o", SO, srcs[0]])
    return SO


class HostSim:
    def __init__(self, params, cells=None, resolution=0.05, origin=(0.0, 0.0), encoding=0, footprint=None,
                 state_rows=0, **knobs):
        self.lib = ctypes.CDLL(build())
        self.params = params_record(params, **knobs)
        self.n_steps = int(self.params["control_steps"])
        self.cells = None if cells is None else np.ascontiguousarray(cells, dtype=np.uint8)
        self.res, self.origin, self.enc = float(resolution), origin, int(encoding)
        fp = np.zeros((0, 2), np.float32) if footprint is None else np.asarray(footprint, np.float32).reshape(-1, 2)
        self.fp = np.ascontiguousarray(fp)
        self.lib.hostsim_state_stride.restype = ctypes.c_int
        self.stride = self.lib.hostsim_state_stride(self.n_steps)
        self.state = np.zeros((state_rows, self.stride), np.float32) if state_rows else None

    def _env(self):
        c = self.cells
        H, W = (c.shape if c is not None else (0, 0))
        return [self.params.ctypes.data_as(ctypes.c_void_p),
                c.ctypes.data_as(ctypes.c_void_p) if c is not None else None,
                ctypes.c_int(W), ctypes.c_int(H), ctypes.c_double(self.res),
                ctypes.c_double(self.origin[0]), ctypes.c_double(self.origin[1]), ctypes.c_int(self.enc),
                self.fp.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(self.fp))]

    def solve(self, reqs, tol_pg=-1.0, tol_f=-1.0):
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        n = len(reqs)
        out = np.zeros(n, RESPONSE_DTYPE)
        plan = np.zeros((n, 3 * self.n_steps), np.float32)
        st = self.state
        rc = self.lib.hostsim_solve(*self._env(), reqs.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n),
                                    out.ctypes.data_as(ctypes.c_void_p), plan.ctypes.data_as(ctypes.c_void_p),
                                    st.ctypes.data_as(ctypes.c_void_p) if st is not None else None,
                                    ctypes.c_uint(0 if st is None else len(st)),
                                    ctypes.c_float(tol_pg), ctypes.c_float(tol_f))
        assert rc == 0
        return out, plan

    def eval(self, reqs, U, grad=True):
        reqs = np.ascontiguousarray(reqs, dtype=REQUEST_DTYPE)
        U = np.ascontiguousarray(U, dtype=np.float32)
        n = len(reqs)
        J = np.zeros(n, np.float32)
        G = np.zeros((n, 3 * self.n_steps), np.float32) if grad else None
        rc = self.lib.hostsim_eval(*self._env(), reqs.ctypes.data_as(ctypes.c_void_p),
                                   U.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(n),
                                   J.ctypes.data_as(ctypes.c_void_p),
                                   G.ctypes.data_as(ctypes.c_void_p) if grad else None)
        assert rc == 0
        return J, G

    def project(self, V):
        V = np.ascontiguousarray(V, dtype=np.float32).copy()
        self.lib.hostsim_project(self.params.ctypes.data_as(ctypes.c_void_p), V.ctypes.data_as(ctypes.c_void_p),
                                 ctypes.c_size_t(V.size // 3))
        return V
