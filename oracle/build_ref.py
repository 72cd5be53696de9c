#!/usr/bin/env python3
"""Recipe for ``oracle/_ref``: byte-compiles the UNMODIFIED reference server module from the sources where they lie
(``/root/reference/neo_mpc_planner2/mpc_optimization_server.py``) into ``oracle/_ref/mpc_optimization_server.pyc.bin``.

The reference's hot path is Python, so "compiling the reference" is ``py_compile``; the .pyc is a build output (listed in
.gitignore, shipped to the GPU box like the built .so files) — no reference source is copied into the repository.
``oracle/ros_stubs.py: load_reference_compiled`` imports it under the ROS stand-in modules; ``bench.py --impl reference``
and the ``cpu_baseline`` leg then time the reference's own ``objective`` / ``f_constraint`` / ``bnds`` / ``cons`` through
``minimize`` exactly as srv.py:363-364 does (``oracle/ref_runner.py``).  Runs in the build container only
(``__graft_entry__.build()`` calls it when /root/reference is present).
"""
from __future__ import annotations

import hashlib
import json
import os
import py_compile
import sys

SRC = "/root/reference/neo_mpc_planner2/mpc_optimization_server.py"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
OUT = os.path.join(OUT_DIR, "mpc_optimization_server.pyc.bin")


def build(src: str = SRC) -> str | None:
    if not os.path.exists(src):
        return None
    os.makedirs(OUT_DIR, exist_ok=True)
    py_compile.compile(src, cfile=OUT, doraise=True, invalidation_mode=py_compile.PycInvalidationMode.UNCHECKED_HASH)
    meta = {"source": src, "sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
            "python": sys.version.split()[0]}
    with open(os.path.join(OUT_DIR, "BUILD_INFO.json"), "w") as f:
        json.dump(meta, f, indent=1)
    return OUT


if __name__ == "__main__":
    print(build() or "no /root/reference here: nothing built")
