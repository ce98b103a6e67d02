"""CPU tier: the product's kernel phase code and engine orchestration (dflo_b200/csrc/kernels.cuh,
engine_core.h, partition.h), compiled for the host by tests/emu and run thread by thread, against
the CPU oracle.  Same cases and tolerances as the GPU tier (tests/test_gpu_parity.py); this tier
exists so index arithmetic, buffer rotation, limiter logic and halo lists are proven before GPU
time is spent.  It never stands in for the CUDA library: the GPU tier calls libdflo_b200.so only."""
import ctypes

from dflo_b200 import abi

import numpy as np
import pytest

from cases import ALL_FLUXES, BASELINE_HORIZON, BASELINE_SMALL, BASES, TOL_RHS, TOL_STEP_SHOCK, TOL_STEP_SMOOTH
from helpers import (check_horizons, DMR_BC, PERIODIC_BOX, SOD_BC, STEP_BC, Case, emu_lib, ic_disc_box, ic_dmr, ic_pulse, ic_pulse_box,
                     ic_smooth, ic_sod, ic_step, ic_vortex)


def _rhs_ok(c):
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    return r_o


@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("basis,k", BASES)
def test_rhs_and_step_periodic(basis, k, flux):
    c = Case(("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex, basis=basis, degree=k, flux=flux, cfl=0.5)
    _rhs_ok(c)
    _, dt_o, dt_e = c.step()
    assert abs(dt_o - dt_e) <= 1e-12 * dt_o
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("compat", ["src", "mpi"])
@pytest.mark.parametrize("basis,k,flux", [("Qk", 2, "lxf"), ("Pk", 2, "lxf"), ("Qk", 1, "roe")])
def test_all_boundary_kinds_multiblock_gravity(basis, k, flux, compat):
    """inflow / slip / pressure / farfield on the 3-block L-shaped forward-step mesh (non-lexicographic
    cell order), gravity source on; LxF exercises the src-vs-src_mpi boundary-average fork."""
    bc = {1: "inflow", 2: "slip", 3: "pressure", 0: "farfield"}
    c = Case(("forward_step", [0.2]), bc, ic_smooth, basis=basis, degree=k, flux=flux, cfl=0.5, compat=compat, gravity=0.7)
    c.set_boundary(values=(0.5, 0.1, 1.2, 3.0), wiggle=0.1)
    _rhs_ok(c)
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("basis,k", [("Pk", 1), ("Pk", 2), ("Qk", 1), ("Qk", 2), ("Qk", 3)])
def test_sod_tvb_positivity(basis, k):
    c = Case(("sod_tube", [20, 3]), SOD_BC, ic_sod, basis=basis, degree=k, flux="hllc", limiter="TVB", char_lim=True,
             pos_lim=True, M=0.0, beta=2.0, cfl=0.5)
    c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    c.limit_initial()
    flips = sum(c.step()[0] for _ in range(2))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    assert np.count_nonzero(c.oracle.limited_flags()) > 0
    c.close()


@pytest.mark.parametrize("basis,k", [("Pk", 2), ("Qk", 2)])
def test_positivity_only(basis, k):
    c = Case(("sod_tube", [20, 4]), SOD_BC, ic_pulse, basis=basis, degree=k, flux="lxf", pos_lim=True, cfl=0.15)
    c.set_boundary(values=(0.0, 0.0, 0.05, 0.05))
    flips = sum(c.step()[0] for _ in range(4))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    c.close()


@pytest.mark.parametrize("char_lim", [False, True])
def test_tvb_component_vs_characteristic_and_pk_angular_momentum(char_lim):
    c = Case(("sod_tube", [16, 3]), SOD_BC, ic_sod, basis="Pk", degree=2, flux="roe", limiter="TVB", char_lim=char_lim,
             conserve_angular_momentum=True, M=0.0, beta=1.5, cfl=0.4)
    c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    c.limit_initial()
    flips = sum(c.step()[0] for _ in range(2))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    c.close()


@pytest.mark.parametrize("name,mesh,bc,ic,prm,nsteps", BASELINE_SMALL, ids=[b[0] for b in BASELINE_SMALL])
def test_baseline_configs_small(name, mesh, bc, ic, prm, nsteps):
    """The five BASELINE.json configurations at sizes the emulation finishes in seconds."""
    small = {"cfg1": ("isentropic_vortex", [8]), "cfg2": ("isentropic_vortex", [5]), "cfg3": ("sod_tube", [24, 3]),
             "cfg4": ("double_mach", [6]), "cfg5": ("forward_step", [0.2])}[name[:4]]
    c = Case(small, bc, ic, **prm)
    if "dmr" in name:
        c.set_boundary(values=(57.1576766498, -33.0, 8.0, 563.5))
    elif "step" in name:
        c.set_boundary(values=(4.2, 0.0, 1.4, 8.8))
    elif "sod" in name:
        c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    shocked = prm.get("limiter", "none") != "none"
    if shocked:
        c.limit_initial()
    _rhs_ok(c)
    flips = sum(c.step()[0] for _ in range(2))
    assert c.rel_err() <= (TOL_STEP_SHOCK if shocked else TOL_STEP_SMOOTH)
    if "kfvs" not in name:
        assert flips == 0
    c.close()


# ---- invariants of the path: hold for the oracle AND the engine ----------------
@pytest.mark.parametrize("basis,k,flux", [("Qk", 2, "roe"), ("Pk", 2, "hllc"), ("Qk", 1, "lxf"), ("Qk", 3, "kfvs")])
def test_free_stream_preservation(basis, k, flux):
    """uniform state, periodic box: rhs == 0 to round-off on both sides"""
    ic = lambda x, y: np.stack([0.3 + 0 * x, -0.2 + 0 * x, 1.1 + 0 * x, 2.7 + 0 * x], axis=-1)
    c = Case(("isentropic_vortex", [5]), PERIODIC_BOX, ic, basis=basis, degree=k, flux=flux)
    r_o, r_e = c.rhs_pair()
    # KFVS: the A&S erf makes H(W,W,n) != F(W).n at 1e-7, identically on every face, so the
    # volume and surface terms do not cancel pointwise -- but they do in the cell integral
    tol = 1e-6 if flux == "kfvs" else 1e-12
    assert np.abs(r_o).max() < tol and np.abs(r_e).max() < tol
    c.close()


@pytest.mark.parametrize("basis,k", [("Qk", 2), ("Pk", 2)])
def test_conservation_periodic(basis, k):
    """sum over cells of the rhs tested against phi = 1 vanishes on a periodic mesh: every interior
    flux enters twice with opposite sign (bit-identical evaluations on both sides)."""
    c = Case(("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex, basis=basis, degree=k, flux="roe")
    r_o, r_e = c.rhs_pair()
    D, ns = c.oracle.D, c.oracle.D // 4
    for r in (r_o, r_e):
        R = r.reshape(-1, 4, ns)
        if basis == "Qk":
            tot = R.sum(axis=(0, 2))           # sum_i rhs_i = rhs tested against sum_i phi_i = 1
        else:
            tot = R[:, :, 0].sum(axis=0)       # phi_0 = 1 on the unit cell
        assert np.abs(tot).max() < 1e-12 * max(1.0, np.abs(r).max())
    c.close()


def test_limiter_preserves_cell_means_and_flags_match():
    c = Case(("sod_tube", [20, 3]), SOD_BC, ic_sod, basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True,
             pos_lim=True, M=0.0, beta=2.0, cfl=0.5)
    c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    before = c.engine.cell_average().copy()
    c.limit_initial()
    u = c.solution().reshape(-1, 4, 9)
    gx, gw = c.oracle.tables()
    w2 = np.outer(gw, gw).reshape(-1)
    means = (u * w2).sum(axis=2)
    assert np.abs(means - before).max() < 1e-13
    c.close()


def test_tvb_large_M_leaves_smooth_data_untouched():
    c = Case(("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex, basis="Qk", degree=2, flux="roe", limiter="TVB",
             char_lim=True, M=1.0e6, beta=1.0)
    u0 = c.solution().copy()
    c.limit_initial()
    assert np.array_equal(u0, c.solution())
    assert np.count_nonzero(c.engine.limited_flags()) == 0
    c.close()


# ---- sharded engine: N in-process ranks with an emulated halo exchange ------------------------
@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("case", ["vortex_Q2_roe", "sod_P2_tvb_pos", "step_Q1_tvb", "sod_Q2_minmax"])
def test_sharded_matches_single(case, world):
    if case == "vortex_Q2_roe":
        args = (("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex)
        prm = dict(basis="Qk", degree=2, flux="roe", cfl=0.5)
        bval = None
    elif case == "sod_P2_tvb_pos":
        args = (("sod_tube", [18, 3]), SOD_BC, ic_sod)
        prm = dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, M=0.0, beta=2.0, cfl=0.5)
        bval = (0.0, 0.0, 1.0, 2.5)
    elif case == "sod_Q2_minmax":
        from helpers import ic_sod_moving_wavy
        args = (("sod_tube", [18, 3]), SOD_BC, ic_sod_moving_wavy)
        prm = dict(basis="Qk", degree=2, flux="hllc", limiter="minmax", char_lim=True, pos_lim=True, M=0.0, beta=2.0, cfl=0.4)
        bval = (0.3, 0.1, 1.0, 2.5)
    else:
        args = (("forward_step", [0.2]), STEP_BC, ic_step)
        prm = dict(basis="Qk", degree=1, flux="lxf", limiter="TVB", char_lim=True, M=0.0, beta=2.0, cfl=0.5)
        bval = (4.2, 0.0, 1.4, 8.8)
    one = Case(*args, **prm)
    many = Case(*args, world=world, **prm)
    for c in (one, many):
        if bval:
            c.set_boundary(values=bval)
        if "limiter" in prm:
            c.limit_initial()
    r1 = one.rhs_pair()[1]
    rn = many.rhs_pair()[1]
    assert np.array_equal(r1, rn)
    for _ in range(2):
        one.step()
        many.step()
    # sharding must not change a single bit: same kernels, same inputs, ghost cells updated redundantly
    assert np.array_equal(one.solution(), many.solution())
    assert many.rel_err() <= TOL_STEP_SHOCK
    one.close()
    many.close()


@pytest.mark.parametrize("compat", ["src", "mpi"])
def test_time_dependent_boundary_expression_under_advance(compat):
    """The moving-shock top boundary of the double Mach reflection (examples/double_mach_reflection/
    input.prm:35-41) evaluated on the device at each stage's BC time (src/claw.cc:736-745 vs
    src_mpi/claw.cc:769-773) inside dflo_b200_advance."""
    from helpers import time_dependent_bc_case
    err, c = time_dependent_bc_case("emu", compat)
    assert err <= TOL_STEP_SHOCK
    c.close()


@pytest.mark.parametrize("variable", ["density", "energy"])
@pytest.mark.parametrize("basis,k", [("Qk", 1), ("Qk", 2), ("Pk", 2)])
def test_kxrcf_shock_indicator_gates_the_limiter(basis, k, variable):
    """`shock indicator = density | energy`: the KXRCF indicator (src/indicator.cc:50-198) decides
    which cells the TVB limiter may touch (src/limiter.cc:263, 406)."""
    from helpers import kxrcf_case
    err, ind_err, flips, c = kxrcf_case("emu", basis, k, variable)
    assert err <= TOL_STEP_SHOCK and ind_err <= 1e-10 and flips == 0
    flags = c.oracle.limited_flags()
    assert 0 < np.count_nonzero(flags) < flags.size          # selective: the shock region only
    c.close()


@pytest.mark.parametrize("k,char_lim,pos_lim", [(1, True, False), (2, True, True), (2, False, False), (3, True, False)])
def test_minmax_limiter(k, char_lim, pos_lim):
    """limiter type = minmax of the MPI tree (src_mpi/limiter.cc:400-553, Qk only): moving Sod problem so
    that the streamline direction of the characteristic projection is defined (with a smooth ripple: on an
    exactly constant cell the reference's `du > 0` test reads round-off noise); decisions and solution
    against the oracle, with and without the characteristic projection (without it the reference's
    range starts from 0) and with the positivity limiter behind it."""
    from helpers import ic_sod_moving_wavy
    c = Case(("sod_tube", [20, 3]), SOD_BC, ic_sod_moving_wavy, basis="Qk", degree=k, flux="hllc", limiter="minmax",
             char_lim=char_lim, pos_lim=pos_lim, M=0.0, beta=2.0, cfl=0.4)
    c.set_boundary(values=(0.3, 0.1, 1.0, 2.5))
    c.limit_initial()
    flips = sum(c.step()[0] for _ in range(2))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    assert np.count_nonzero(c.oracle.limited_flags() & 1) > 0
    c.close()


def test_minmax_limiter_is_qk_only():
    """src_mpi/parameters.cc:610-611: 'minmax limiter is implemented only for Qk'."""
    with pytest.raises(Exception):
        Case(("sod_tube", [8, 2]), SOD_BC, ic_sod, basis="Pk", degree=1, flux="lxf", limiter="minmax", M=0.0, beta=1.0, cfl=0.4)


@pytest.mark.parametrize("basis,k,flux", [("Qk", 1, "lxf"), ("Qk", 3, "roe"), ("Pk", 2, "hllc"), ("Qk", 2, "kep")])
def test_external_force_mpi(basis, k, flux):
    """External force of the MPI tree (f_0 / f_1 value, src_mpi/assemble_explicit.cc:56-58, 84;
    src_mpi/equation.h:1189-1202): forcing term gravity * (rho f, m.f) with a position-dependent f, right-hand
    side and a step against the oracle; f = (0,-1) reproduces the hard-wired forcing of src/ bit for bit."""
    bc = {1: "inflow", 2: "slip", 3: "pressure", 0: "farfield"}
    c = Case(("forward_step", [0.2]), bc, ic_smooth, basis=basis, degree=k, flux=flux, cfl=0.5, compat="mpi", gravity=0.7)
    c.set_boundary(values=(0.5, 0.1, 1.2, 3.0), wiggle=0.1)
    r_src = c.rhs_pair()[1].copy()
    c.set_external_force("0.0", "-1.0", lambda x, y: (0.0 * x, -1.0 + 0.0 * x))
    assert np.array_equal(c.rhs_pair()[1], r_src)
    c.set_external_force("0.3*sin(2*x)+y", "-1.0+0.2*x*y", lambda x, y: (0.3 * np.sin(2 * x) + y, -1.0 + 0.2 * x * y))
    _rhs_ok(c)
    assert np.abs(c.rhs_pair()[1] - r_src).max() > 1e-3
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.fixture
def pk_cell_kernel(monkeypatch):
    """Selects the thread-per-cell Pk stage kernel (dflo_b200/csrc/cell_stage.cuh, the default on the device) in the
    CPU emulation, whose default is the tile kernel."""
    monkeypatch.setenv("DFLO_EMU_PK", "cell")
    L = emu_lib()
    L.dflo_emu_cell_stage_launches.restype = ctypes.c_int64
    before = L.dflo_emu_cell_stage_launches()
    yield
    assert L.dflo_emu_cell_stage_launches() > before, "the cell stage kernel did not run"


@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("k", [1, 2])
def test_pk_cell_kernel_rhs_and_step_periodic(pk_cell_kernel, k, flux):
    c = Case(("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex, basis="Pk", degree=k, flux=flux, cfl=0.5)
    _rhs_ok(c)
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("k,flux", [(1, "roe"), (2, "hllc"), (2, "kfvs"), (2, "lxf")])
def test_pk_cell_kernel_several_blocks(pk_cell_kernel, k, flux):
    """More cells than one block of 128 (and a ragged last block): faces between blocks and periodic pairs go through the
    job list, faces inside a block are solved once and handed to the neighbour -- periodic box of 196 cells, and the
    three-block forward-step mesh (252 cells) with every boundary kind."""
    c = Case(("isentropic_vortex", [14]), PERIODIC_BOX, ic_vortex, basis="Pk", degree=k, flux=flux, cfl=0.5)
    _rhs_ok(c)
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()
    bc = {1: "inflow", 2: "slip", 3: "pressure", 0: "farfield"}
    c = Case(("forward_step", [0.1]), bc, ic_smooth, basis="Pk", degree=k, flux=flux, cfl=0.5, gravity=0.3)
    assert c.oracle.n_cells > 128
    c.set_boundary(values=(0.5, 0.1, 1.2, 3.0), wiggle=0.1)
    _rhs_ok(c)
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


def test_pk_cell_kernel_mesh_check():
    """pk_cell_mesh_ok (cell_stage.cuh): the cell kernel hands the moments of a face to the neighbour's face F ^ 1, so it
    is only admitted when every interior face is seen as F / F ^ 1 with equal flags by its two (updated) cells; otherwise
    the engine keeps the tile kernel."""
    L = emu_lib()
    i32, u8 = ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_ubyte)
    L.dflo_emu_pk_cell_mesh_ok.argtypes = [i32, u8, ctypes.c_int]

    def ok(nbr, fl, n=None):
        nbr, fl = np.ascontiguousarray(nbr, dtype=np.int32), np.ascontiguousarray(fl, dtype=np.uint8)
        return bool(L.dflo_emu_pk_cell_mesh_ok(nbr.ctypes.data_as(i32), fl.ctypes.data_as(u8), len(nbr) if n is None else n))

    # two cells side by side: cell 0's right face (1) is cell 1's left face (0); boundary faces elsewhere
    nbr = [[-1, 1, -2, -3], [0, -4, -5, -6]]
    fl = [[0, 1, 0, 0], [0, 0, 0, 0]]
    assert ok(nbr, fl)
    # the neighbour sees the face as its face 2 (a rotated cell): not admitted
    assert not ok([[-1, 1, -2, -3], [-4, -5, 0, -6]], fl)
    # a flipped or periodic pair is solved from both sides and needs no counterpart
    for flag in (2, 4):
        assert ok([[-1, 1, -2, -3], [-4, -5, 0, -6]], [[0, flag, 0, 0], [0, 0, flag, 0]])
    # flags that differ between the two sides: not admitted
    assert not ok(nbr, [[0, 1, 0, 0], [4, 0, 0, 0]])
    # only updated cells matter: with n_compute = 1 the rotated neighbour is a ghost cell outside every block
    assert ok([[-1, 1, -2, -3], [-4, -5, 0, -6]], fl, n=1)


@pytest.mark.parametrize("compat", ["src", "mpi"])
def test_pk_cell_kernel_boundaries_gravity_external_force(pk_cell_kernel, compat):
    bc = {1: "inflow", 2: "slip", 3: "pressure", 0: "farfield"}
    c = Case(("forward_step", [0.2]), bc, ic_smooth, basis="Pk", degree=2, flux="lxf", cfl=0.5, compat=compat, gravity=0.7)
    c.set_boundary(values=(0.5, 0.1, 1.2, 3.0), wiggle=0.1)
    _rhs_ok(c)
    if compat == "mpi":
        c.set_external_force("0.3*sin(2*x)+y", "-1.0+0.2*x*y", lambda x, y: (0.3 * np.sin(2 * x) + y, -1.0 + 0.2 * x * y))
        _rhs_ok(c)
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("world", [1, 2, 3])
def test_pk_cell_kernel_sod_tvb_positivity_sharded(pk_cell_kernel, world):
    """cfg3 in small: P2, HLLC, TVB + positivity; the sharded runs (2-layer halo, ghost layer 1 updated
    redundantly by the cell kernel, means only) equal the single run bit for bit."""
    prm = dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, M=0.0, beta=2.0, cfl=0.5)
    one = Case(("sod_tube", [18, 3]), SOD_BC, ic_sod, **prm)
    many = Case(("sod_tube", [18, 3]), SOD_BC, ic_sod, world=world, **prm)
    for c in (one, many):
        c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
        c.limit_initial()
    flips = 0
    for _ in range(2):
        one.step()
        flips += many.step()[0]
    assert np.array_equal(one.solution(), many.solution())
    assert many.rel_err() <= TOL_STEP_SHOCK and flips == 0
    one.close()
    many.close()


def test_fixed_time_step_when_cfl_is_not_positive():
    """claw.cc:457-461: global time stepping with cfl <= 0 takes `time step` from the input file -- through
    compute_dt and through advance (which must not take silent zero-length steps), clipped at the final time;
    cfl <= 0 without a time step is refused at creation."""
    c = Case(("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex, basis="Qk", degree=1, flux="lxf", cfl=0.0, time_step=1e-3)
    e = c.engine
    assert e.compute_dt(0.0) == 1e-3
    assert e.compute_dt(0.0, final_time=4e-4) == 4e-4
    t, dt = e.advance(5)
    assert abs(t - 5e-3) < 1e-15 and dt == 1e-3
    t, dt = e.advance(3, elapsed=t, final_time=6.5e-3)
    assert abs(t - 6.5e-3) < 1e-15
    # same steps on the oracle with the same dt
    o = c.oracle
    for dt_o in [1e-3] * 6 + [0.5e-3]:
        for rk in range(o.n_rk):
            assert o.rk_stage(rk, dt_o)[0] == 0
        o.commit_step()
    assert c.rel_err() <= 1e-12
    c.close()
    for bad in (0.0, -1.0):
        with pytest.raises(Exception):
            Case(("isentropic_vortex", [4]), PERIODIC_BOX, ic_vortex, basis="Qk", degree=1, flux="lxf", cfl=bad, time_step=-1.0)


def test_time_step_skips_cells_without_a_valid_value():
    """std::min (global_dt, dt(c)) skips a NaN dt(c) (claw.cc:508): one broken cell mean must not turn the global
    time step into NaN, 0 or a negative number."""
    c = Case(("isentropic_vortex", [6]), PERIODIC_BOX, ic_vortex, basis="Qk", degree=1, flux="lxf", cfl=0.5)
    dt_ok = c.engine.compute_dt(0.0)
    u = c.u0.copy().reshape(-1, 4, 4)
    u[7, 2, :] = -1.0          # negative density in one cell: sound speed NaN
    c.engine.set_solution(u.reshape(-1))
    dt = c.engine.compute_dt(0.0)
    assert np.isfinite(dt) and dt > 0 and abs(dt - dt_ok) < 0.2 * dt_ok
    c.close()


@pytest.mark.parametrize("key,_,size", BASELINE_HORIZON, ids=[b[0] for b in BASELINE_HORIZON])
def test_baseline_configs_20_step_horizon(key, _, size):
    """SURVEY.md 8(d) horizons (1 RHS / 1 step / 20 steps) on the CPU emulation of the kernel code."""
    check_horizons(key, size, "emu")


# ---------------------------------------------------------------------------------------------
# mapping = q1: straight-sided general quadrilaterals, Qk
# ---------------------------------------------------------------------------------------------
def _q1_case(backend, k, flux, rotate, bc, ic, n=8, **extra):
    ids = (4, 2, 1, 3)
    return Case(("rectangle_skew", [n, n, -5, 5, -5, 5, *ids, 0.15, rotate]), bc, ic, backend=backend, basis="Qk", degree=k, flux=flux,
                cfl=0.05 if flux == "kep" else 0.3, mapping="q1", **extra)   # kep on this coarse mesh: unstable beyond a few small steps, Cartesian or not


@pytest.mark.parametrize("rotate", [0, 1])
@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_q1_mapping_rhs_and_steps_periodic(k, flux, rotate):
    """MappingQ1 on smoothly skewed quadrilaterals, periodic box: right-hand side, compute_time_step_q and three
    steps against the oracle; rotate = 1 mixes the cell orientations (neighbours on arbitrary faces, reversed lines)."""
    c = _q1_case("emu", k, flux, rotate, PERIODIC_BOX, ic_vortex, compat="mpi")
    _rhs_ok(c)
    for _ in range(3):
        _, dt_o, dt_e = c.step()
        assert abs(dt_o - dt_e) <= 1e-12 * dt_o
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("compat", ["src", "mpi"])
@pytest.mark.parametrize("k,flux", [(1, "lxf"), (2, "roe"), (3, "hllc"), (2, "kfvs")])
def test_q1_mapping_all_boundary_kinds_gravity(k, flux, compat):
    bc = {1: "inflow", 2: "slip", 3: "pressure", 4: "farfield"}
    c = _q1_case("emu", k, flux, 1, bc, ic_smooth, compat=compat, gravity=0.7)
    c.set_boundary(values=(1.0, 0.2, 1.4, 8.8), wiggle=0.05)
    _rhs_ok(c)
    for _ in range(2):
        c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


def test_q1_mapping_free_stream_and_cartesian_limit():
    """A uniform state is preserved on skewed cells (metric identities of the bilinear map), and on an undistorted
    mesh mapping = q1 reproduces mapping = cartesian (only the time-step formula differs: claw.cc:518-557 vs 484-511)."""
    bc = {1: "outflow", 2: "outflow", 3: "outflow", 4: "outflow"}
    uniform = lambda x, y: np.stack([0.7 + 0 * x, -0.3 + 0 * x, 1.2 + 0 * x, 3.0 + 0 * x], axis=-1)
    c = _q1_case("emu", 3, "hllc", 1, bc, uniform)
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_e).max() < 1e-13 and np.abs(r_o).max() < 1e-13
    c.close()
    ids = (4, 2, 1, 3)
    res = []
    for mp in ("cartesian", "q1"):
        c = Case(("rectangle_skew", [6, 6, -5, 5, -5, 5, *ids, 0.0, 0]), PERIODIC_BOX, ic_vortex, backend="emu", basis="Qk", degree=2,
                 flux="roe", cfl=0.5, mapping=mp, compat="mpi")
        res.append(c.rhs_pair()[1])
        c.close()
    assert np.abs(res[0] - res[1]).max() <= 1e-13 * np.abs(res[0]).max()


def test_q1_mapping_refusals():
    """src/parameters.cc:545-549: TVB and Pk need Cartesian grids; a skewed mesh is refused under mapping = cartesian."""
    ids = (4, 2, 1, 3)
    skew = ("rectangle_skew", [4, 4, -5, 5, -5, 5, *ids, 0.15, 0])
    for kw in (dict(basis="Pk", degree=1, mapping="q1"), dict(basis="Qk", degree=1, mapping="q1", limiter="TVB"),
               dict(basis="Qk", degree=1, mapping="cartesian")):
        params, pair = abi.make_params(bc=PERIODIC_BOX, flux="lxf", **kw)
        mesh = abi.Mesh(skew[0], skew[1], lib=emu_lib())
        flat = mesh.flatten(params, pair)
        with pytest.raises(abi.DfloError):
            abi.Engine(flat, params, lib=emu_lib(), prefix="dflo_emu_")


@pytest.mark.parametrize("k,ic", [(1, ic_pulse_box), (2, ic_disc_box), (3, ic_disc_box)])
def test_q1_mapping_positivity_limiter(k, ic):
    """The positivity limiter on mapped cells (the reference admits it: parameters.cc:536-550 refuses only TVB and Pk off
    Cartesian grids): positivity.cc evaluates the solution at GLL x Gauss points of the unit cell and scales about the
    cell average taken with the mapped JxW.  Skewed quadrilaterals with mixed orientations; the limiter must act (both
    the density and the pressure stage) and its decisions must agree cell by cell."""
    ids = (4, 2, 1, 3)
    c = Case(("rectangle_skew", [10, 10, -5, 5, -5, 5, *ids, 0.15, 1]), PERIODIC_BOX, ic, basis="Qk", degree=k, flux="lxf",
             pos_lim=True, cfl=0.15, mapping="q1", compat="mpi")
    acted, flips = 0, 0
    for _ in range(4):
        flips += c.step()[0]
        acted |= int(np.bitwise_or.reduce(c.oracle.limited_flags()))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    assert acted & 2 and acted & 4, "the limiter never acted: the case does not test it"
    c.close()


def test_q1_mapping_positivity_limiter_sharded():
    """positivity on mapped cells over two ranks (one ghost layer; the limiter is local to a cell): bit for bit the
    single-rank result"""
    ids = (4, 2, 1, 3)
    args = (("rectangle_skew", [10, 10, -5, 5, -5, 5, *ids, 0.15, 1]), PERIODIC_BOX, ic_disc_box)
    prm = dict(basis="Qk", degree=2, flux="lxf", pos_lim=True, cfl=0.15, mapping="q1", compat="mpi")
    one, two = Case(*args, **prm), Case(*args, world=2, **prm)
    for _ in range(3):
        one.step()
        two.step()
    assert np.array_equal(one.solution(), two.solution())
    assert two.rel_err() <= TOL_STEP_SHOCK
    one.close()
    two.close()


@pytest.mark.parametrize("world", [2, 3])
def test_q1_mapping_sharded_matches_single(world):
    """mapping = q1 on a sharded context: cells by id over the ranks, one ghost layer, the mapped stage kernel reads its
    neighbours (ghosts included) from the exchanged buffers -- bit for bit the single-rank result."""
    bc = {1: "inflow", 2: "slip", 3: ("periodic", 1), 4: "farfield"}
    bc = {1: ("periodic", 3), 3: ("periodic", 1), 2: "slip", 4: "farfield"}
    ids = (4, 2, 1, 3)
    args = (("rectangle_skew", [7, 6, -5, 5, -5, 5, *ids, 0.15, 1]), bc, ic_smooth)
    prm = dict(basis="Qk", degree=2, flux="hllc", cfl=0.3, mapping="q1", compat="mpi")
    one = Case(*args, **prm)
    many = Case(*args, world=world, **prm)
    for c in (one, many):
        c.set_boundary(values=(1.0, 0.2, 1.4, 8.8))
    assert np.array_equal(one.rhs_pair()[1], many.rhs_pair()[1])
    for _ in range(2):
        one.step()
        many.step()
    assert np.array_equal(one.solution(), many.solution())
    assert many.rel_err() <= TOL_STEP_SMOOTH
    one.close()
    many.close()


# ---------------------------------------------------------------------------------------------
# time step type = local (src/claw.cc:444-478, 694-713)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("basis,k,flux,mapping", [("Qk", 1, "kfvs", "cartesian"), ("Qk", 3, "roe", "cartesian"), ("Pk", 2, "hllc", "cartesian"),
                                                  ("Qk", 1, "kfvs", "q1"), ("Qk", 2, "hllc", "q1")])
def test_local_time_stepping(basis, k, flux, mapping):
    """Every cell advances with its own dt(cell); the clock moves by the smallest one, which is not clipped at the final
    time.  compute_dt, the stage update and the device-side advance against the oracle."""
    from helpers import local_time_stepping_case
    local_time_stepping_case("emu", basis, k, flux, mapping)


def test_local_time_stepping_compression_corner():
    """examples/compression_corner: Q1, KFVS, mapping = q1, time step type = local on the two-block trapezoid mesh."""
    from helpers import compression_corner_case
    compression_corner_case("emu")


# ---------------------------------------------------------------------------------------------
# faces with hanging nodes (src/refine.cc; MeshWorker sub-face rule)
# ---------------------------------------------------------------------------------------------
def _refined_case(backend, k, flux, mapping, bc, ic, patch=(2, 5, 1, 4), n=(7, 6), rotate=0, **extra):
    ids = (4, 2, 1, 3)
    return Case(("rectangle_refined", [n[0], n[1], -5, 5, -5, 5, *ids, *patch, rotate]), bc, ic, backend=backend, basis="Qk", degree=k, flux=flux,
                cfl=0.05 if flux == "kep" else 0.3, mapping=mapping, **extra)


@pytest.mark.parametrize("mapping", ["cartesian", "q1"])
@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_hanging_nodes_rhs_and_steps(k, flux, mapping):
    """A patch of cells split into four inside a periodic box: every face on the rim of the patch has a hanging node and
    is integrated from the fine side on the two halves (FESubfaceValues on the coarse side).  Right-hand side, time step
    and three steps against the oracle, under both mappings."""
    c = _refined_case("emu", k, flux, mapping, PERIODIC_BOX, ic_vortex, compat="mpi")
    assert c.oracle.n_cells == 7 * 6 + 3 * 9
    _rhs_ok(c)
    for _ in range(3):
        _, dt_o, dt_e = c.step()
        assert abs(dt_o - dt_e) <= 1e-12 * dt_o
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("k,flux", [(1, "lxf"), (2, "roe"), (3, "hllc")])
def test_hanging_nodes_mixed_orientations(k, flux):
    """The same with the vertex order of the cells turned at random: coarse and fine cells meet on arbitrary local faces
    and run along the shared line in either direction (mapping = q1)."""
    c = _refined_case("emu", k, flux, "q1", PERIODIC_BOX, ic_vortex, rotate=1, compat="mpi")
    _rhs_ok(c)
    for _ in range(3):
        c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


def test_hanging_nodes_free_stream_conservation_and_boundaries():
    """Uniform flow through a refined patch that touches the boundary stays uniform; in a periodic box the integral of every
    conserved variable is kept to round-off (the two sides of a sub-face subtract the same bits); all boundary kinds."""
    bc = {1: "outflow", 2: "outflow", 3: "outflow", 4: "outflow"}
    uniform = lambda x, y: np.stack([0.7 + 0 * x, -0.3 + 0 * x, 1.2 + 0 * x, 3.0 + 0 * x], axis=-1)
    c = _refined_case("emu", 3, "hllc", "cartesian", bc, uniform, patch=(0, 3, 2, 6))
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_e).max() < 1e-13 and np.abs(r_o).max() < 1e-13
    c.close()
    c = _refined_case("emu", 2, "roe", "cartesian", PERIODIC_BOX, ic_vortex, compat="mpi")
    v, cells, _, _ = c.mesh.primitive()
    area = (v[cells[:, 1], 0] - v[cells[:, 0], 0]) * (v[cells[:, 2], 1] - v[cells[:, 0], 1])
    gx, gw = c.oracle.tables()
    w = (gw[None, :] * gw[:, None]).reshape(-1)
    total = lambda u: (u.reshape(len(cells), 4, -1) * w[None, None, :] * area[:, None, None]).sum(axis=(0, 2))
    t0 = total(c.solution())
    te, _ = c.engine.advance(3)
    assert te > 0 and np.abs(total(c.solution()) - t0).max() <= 1e-12 * np.abs(t0).max()
    c.close()
    bc = {1: "inflow", 2: "slip", 3: "pressure", 4: "farfield"}
    c = _refined_case("emu", 2, "hllc", "q1", bc, ic_smooth, patch=(0, 3, 2, 6), gravity=0.7)
    c.set_boundary(values=(1.0, 0.2, 1.4, 8.8), wiggle=0.05)
    _rhs_ok(c)
    for _ in range(2):
        c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


def test_hanging_nodes_refusals():
    ids = (4, 2, 1, 3)
    mesh = ("rectangle_refined", [6, 6, -5, 5, -5, 5, *ids, 2, 4, 2, 4])
    for kw in (dict(basis="Pk", degree=1), dict(basis="Qk", degree=1, limiter="TVB")):
        params, pair = abi.make_params(bc={}, flux="lxf", **kw)
        m = abi.Mesh(mesh[0], mesh[1], lib=emu_lib())
        flat = m.flatten(params, pair)
        assert flat.contents.n_hanging_faces == 8
        with pytest.raises(abi.DfloError):
            abi.Engine(flat, params, lib=emu_lib(), prefix="dflo_emu_")


@pytest.mark.parametrize("mapping,world", [("cartesian", 1), ("q1", 1), ("cartesian", 2)])
def test_hanging_nodes_positivity_limiter(mapping, world):
    """The positivity limiter next to hanging nodes (it is local to a cell): a refined patch whose rim cuts the edge of the
    dense disc, both limiter stages acting, the oracle's decision in every cell; sharded = single bit for bit."""
    ids = (4, 2, 1, 3)
    args = (("rectangle_refined", [10, 10, -5, 5, -5, 5, *ids, 4, 8, 3, 7]), PERIODIC_BOX, ic_disc_box)
    prm = dict(basis="Qk", degree=2, flux="lxf", pos_lim=True, cfl=0.1, mapping=mapping, compat="mpi")
    c = Case(*args, world=world, **prm)
    one = Case(*args, **prm) if world > 1 else None
    acted, flips = 0, 0
    for _ in range(4):
        flips += c.step()[0]
        acted |= int(np.bitwise_or.reduce(c.oracle.limited_flags()))
        if one:
            one.step()
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    assert acted & 2 and acted & 4, "the limiter never acted: the case does not test it"
    if one:
        assert np.array_equal(one.solution(), c.solution())
        one.close()
    c.close()


@pytest.mark.parametrize("world", [2, 3])
def test_hanging_nodes_sharded_matches_single(world):
    """Cells by id over the ranks with hanging nodes on (and next to) the cuts: a coarse cell's halo holds BOTH fine cells
    behind its face; bit for bit the single-rank result."""
    bc = {1: ("periodic", 3), 3: ("periodic", 1), 2: "slip", 4: "farfield"}
    ids = (4, 2, 1, 3)
    args = (("rectangle_refined", [7, 6, -5, 5, -5, 5, *ids, 2, 5, 1, 5]), bc, ic_smooth)
    prm = dict(basis="Qk", degree=2, flux="hllc", cfl=0.3, compat="mpi")
    one = Case(*args, **prm)
    many = Case(*args, world=world, **prm)
    for c in (one, many):
        c.set_boundary(values=(1.0, 0.2, 1.4, 8.8))
    assert np.array_equal(one.rhs_pair()[1], many.rhs_pair()[1])
    for _ in range(2):
        one.step()
        many.step()
    assert np.array_equal(one.solution(), many.solution())
    assert many.rel_err() <= TOL_STEP_SMOOTH
    one.close()
    many.close()
