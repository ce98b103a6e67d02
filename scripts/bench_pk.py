#!/usr/bin/env python
"""Stage kernel of the Pk basis alone (P1-P3, HLLC, periodic 512x512 box): thread-per-cell kernel
(cell_stage.cuh, default) against the tile kernel (DFLO_B200_PK=tile).  One process per setting:
    python scripts/bench_pk.py; DFLO_B200_PK=tile python scripts/bench_pk.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dflo_b200 import abi  # noqa: E402

PERIODIC = {1: ("periodic", 3), 3: ("periodic", 1), 2: ("periodic", 4), 4: ("periodic", 2)}


def main():
    n = 512
    for k in (1, 2, 3):
        params, pair = abi.make_params(basis="Pk", degree=k, flux="hllc", bc=PERIODIC, cfl=0.5, compat="mpi")
        mesh = abi.Mesh("rectangle", [n, n, 0.0, 1.0, 0.0, 1.0, 4, 2, 1, 3])
        eng = abi.Engine(mesh.flatten(params, pair), params, device=0)
        D = eng.D
        ns = D // 4
        u = np.zeros((n * n, 4, ns))
        rng = np.random.default_rng(0)
        u[:, :, 0] = np.array([0.3, 0.1, 1.0, 2.6])
        u[:, :, 1:] = 0.01 * rng.standard_normal((n * n, 4, ns - 1))
        eng.set_solution(np.ascontiguousarray(u.reshape(-1)))
        ms = eng.time_stage_kernel(rk=1, reps=10, flush_bytes=256 << 20)
        print(json.dumps({"kernel": os.environ.get("DFLO_B200_PK", "cell"), "basis": "P%d" % k, "cells": n * n, "dofs": n * n * D,
                          "stage_kernel_us": 1e3 * ms, "gdof_per_s": n * n * D / ms / 1e6}))
        eng.close()


if __name__ == "__main__":
    main()
