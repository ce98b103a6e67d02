"""Case lists shared by the CPU-emulation tier and the GPU tier (same inputs, same tolerances)."""
from helpers import (DMR_BC, PERIODIC_BOX, SOD_BC, STEP_BC, ic_dmr, ic_pulse, ic_smooth, ic_sod, ic_step, ic_vortex)

ALL_FLUXES = ["lxf", "sw", "kfvs", "roe", "hllc", "kep"]   # kep: src_mpi only
BASES = [("Qk", 0), ("Qk", 1), ("Qk", 2), ("Qk", 3), ("Qk", 4), ("Pk", 1), ("Pk", 2), ("Pk", 3)]

# (id, mesh, bc, ic, params, n_steps): the five BASELINE configurations at oracle-friendly sizes
BASELINE_SMALL = [
    ("cfg1_vortex_Q1_lxf", ("isentropic_vortex", [32]), PERIODIC_BOX, ic_vortex,
     dict(basis="Qk", degree=1, flux="lxf", cfl=0.9), 3),
    ("cfg2_vortex_Q3_roe", ("isentropic_vortex", [16]), PERIODIC_BOX, ic_vortex,
     dict(basis="Qk", degree=3, flux="roe", cfl=0.9), 3),
    ("cfg3_sod_P2_hllc_tvb_pos", ("sod_tube", [100, 10]), SOD_BC, ic_sod,
     dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.9), 3),
    ("cfg4_dmr_Q2_hllc_tvb", ("double_mach", [16]), DMR_BC, ic_dmr,
     dict(basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True, beta=1.0, M=100.0, cfl=0.9), 3),
    ("cfg5_step_Q3_kfvs_tvb_pos", ("forward_step", [0.05]), STEP_BC, ic_step,
     dict(basis="Qk", degree=3, flux="kfvs", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.5), 3),
]

# tolerances: fp64, relative to max(1, |.|_inf)
TOL_RHS = 1e-13
TOL_STEP_SMOOTH = 1e-12
TOL_STEP_SHOCK = 1e-9   # limiter branch flips at the reference's 1e-10 "change" threshold

# 20-step horizons of SURVEY.md 8(d): (config key, generator size for the GPU tier, for the CPU-emulation tier)
BASELINE_HORIZON = [("cfg1", [32], [16]), ("cfg2", [48], [10]), ("cfg3", [200, 20], [40, 4]), ("cfg4", [24], [8]), ("cfg5", [0.04], [0.1])]
TOL_20_SMOOTH = 1e-11
