#!/usr/bin/env python
"""Small runs of every kernel family for compute-sanitizer (memcheck / racecheck / synccheck), one GPU:

    compute-sanitizer --tool memcheck  python scripts/sanitize_cases.py
    compute-sanitizer --tool racecheck python scripts/sanitize_cases.py

cfg1-size meshes: the row stage kernel (Q1..Q4, periodic + physical boundaries, multi-block forward step), the Pk
thread-per-cell stage kernel (P1, P2), the generic tile kernel (P3, degree 0), the TVB / positivity / minmax limiter
kernels, the KXRCF indicator, the mapped stage kernel (skewed cells, hanging nodes), boundary-expression and
external-force evaluation, set/get layout kernels, dt.  Results
are also checked against the oracle so a sanitizer run doubles as a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import (DMR_BC, PERIODIC_BOX, SOD_BC, STEP_BC, Case, ic_dmr, ic_smooth, ic_sod, ic_sod_moving,  # noqa: E402
                     ic_sod_moving_wavy, ic_step, ic_vortex)

RUNS = [
    ("row Q1 lxf periodic", ("isentropic_vortex", [16]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=1, flux="lxf", cfl=0.9), None),
    ("row Q2 hllc periodic", ("isentropic_vortex", [12]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=2, flux="hllc", cfl=0.9), None),
    ("row Q3 roe periodic", ("isentropic_vortex", [16]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=3, flux="roe", cfl=0.9), None),
    ("row Q4 kep periodic", ("isentropic_vortex", [8]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=4, flux="kep", cfl=0.9, compat="mpi"), None),
    ("row Q3 kfvs step tvb+pos", ("forward_step", [0.1]), STEP_BC, ic_step,
     dict(basis="Qk", degree=3, flux="kfvs", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.5), (4.2, 0.0, 1.4, 8.8)),
    ("row Q2 hllc dmr tvb", ("double_mach", [8]), DMR_BC, ic_dmr,
     dict(basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True, beta=1.0, M=100.0, cfl=0.9), (57.1576766498, -33.0, 8.0, 563.5)),
    ("row Q2 sw gravity boundaries", ("forward_step", [0.2]), {1: "inflow", 2: "slip", 3: "pressure"}, ic_smooth,
     dict(basis="Qk", degree=2, flux="sw", cfl=0.5, gravity=0.7), (1.0, 0.2, 1.4, 8.8)),
    ("pk cell P1 roe", ("isentropic_vortex", [12]), PERIODIC_BOX, ic_vortex, dict(basis="Pk", degree=1, flux="roe", cfl=0.9), None),
    ("pk cell P2 hllc sod tvb+pos", ("sod_tube", [40, 4]), SOD_BC, ic_sod,
     dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.9), (0.0, 0.0, 1.0, 2.5)),
    ("tile P3 lxf", ("isentropic_vortex", [8]), PERIODIC_BOX, ic_vortex, dict(basis="Pk", degree=3, flux="lxf", cfl=0.9), None),
    ("tile Q0 lxf", ("isentropic_vortex", [16]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=0, flux="lxf", cfl=0.9), None),
    ("kxrcf Q2 density", ("sod_tube", [40, 4]), {0: "outflow", 1: "outflow", 2: "inflow"}, ic_sod_moving,
     dict(basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True, beta=2.0, M=0.0, cfl=0.5, shock_indicator="density"), (0.3, 0.1, 1.0, 2.55)),
    ("minmax Q2 + pos", ("sod_tube", [40, 4]), SOD_BC, ic_sod_moving_wavy,
     dict(basis="Qk", degree=2, flux="hllc", limiter="minmax", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.4), (0.3, 0.1, 1.0, 2.5)),
    # the mapped stage kernel: skewed quadrilaterals with mixed cell orientations, and a refined patch (faces with hanging nodes)
    ("mapped Q2 hllc skewed", ("rectangle_skew", [12, 12, -5, 5, -5, 5, 4, 2, 1, 3, 0.15, 1]), PERIODIC_BOX, ic_vortex,
     dict(basis="Qk", degree=2, flux="hllc", cfl=0.3, mapping="q1", compat="mpi"), None),
    ("mapped Q3 roe hanging nodes", ("rectangle_refined", [12, 12, -5, 5, -5, 5, 4, 2, 1, 3, 3, 9, 2, 10]), PERIODIC_BOX, ic_vortex,
     dict(basis="Qk", degree=3, flux="roe", cfl=0.3, compat="mpi"), None),
    ("pk cell P2 kfvs two blocks", ("isentropic_vortex", [14]), PERIODIC_BOX, ic_vortex, dict(basis="Pk", degree=2, flux="kfvs", cfl=0.9), None),
]


def main():
    only = sys.argv[1].split(",") if sys.argv[1:] else []    # comma-separated substrings of the case names
    worst = 0.0
    for name, mesh, bc, ic, prm, g in RUNS:
        if only and not any(o in name for o in only):
            continue
        c = Case(mesh, bc, ic, backend="cuda", **prm)
        if g is not None:
            c.set_boundary(values=g)
        if prm.get("limiter", "none") != "none":
            c.limit_initial()
        r_o, r_e = c.rhs_pair()
        e_rhs = np.abs(r_o - r_e).max() / max(1.0, np.abs(r_o).max())
        for _ in range(2):
            c.step()
        err = c.rel_err()
        t, _ = c.engine.advance(2, elapsed=c.t)     # the captured-graph path as well
        c.engine.poll_error()
        print("%-32s rhs %.2e  2 steps %.2e  launches %d" % (name, e_rhs, err, c.engine.launch_count()), flush=True)
        worst = max(worst, e_rhs, err)
        c.close()
    assert worst < 1e-9, worst
    print("sanitize_cases: ok, worst relative error %.2e" % worst)


if __name__ == "__main__":
    main()
