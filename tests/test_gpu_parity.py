"""GPU tier: the CUDA library, called through the C ABI, against the CPU oracle."""
import os

import numpy as np
import pytest

from cases import ALL_FLUXES, BASELINE_HORIZON, BASELINE_SMALL, BASES, TOL_RHS, TOL_STEP_SHOCK, TOL_STEP_SMOOTH
from dflo_b200 import abi
from helpers import (check_horizons, DMR_BC, PERIODIC_BOX, SOD_BC, STEP_BC, Case, ic_disc_box, ic_dmr, ic_pulse, ic_pulse_box, ic_smooth,
                     ic_sod, ic_vortex)

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("basis,k", BASES)
def test_rhs_and_steps_periodic(basis, k, flux):
    c = Case(("isentropic_vortex", [9]), PERIODIC_BOX, ic_vortex, backend="cuda", basis=basis, degree=k, flux=flux, cfl=0.5)
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    for _ in range(2):
        _, dt_o, dt_e = c.step()
        assert abs(dt_o - dt_e) <= 1e-12 * dt_o
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("compat", ["src", "mpi"])
@pytest.mark.parametrize("basis,k,flux", [("Qk", 2, "lxf"), ("Pk", 2, "lxf"), ("Qk", 1, "roe"), ("Qk", 3, "hllc")])
def test_all_boundary_kinds_multiblock_gravity(basis, k, flux, compat):
    bc = {1: "inflow", 2: "slip", 3: "pressure", 0: "farfield"}
    c = Case(("forward_step", [0.1]), bc, ic_smooth, backend="cuda", basis=basis, degree=k, flux=flux, cfl=0.5,
             compat=compat, gravity=0.7)
    c.set_boundary(values=(0.5, 0.1, 1.2, 3.0), wiggle=0.1)
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    for _ in range(2):
        c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("flux", ["lxf", "hllc", "roe"])
@pytest.mark.parametrize("basis,k", [("Pk", 1), ("Pk", 2), ("Qk", 1), ("Qk", 2), ("Qk", 3)])
def test_sod_tvb_positivity(basis, k, flux):
    c = Case(("sod_tube", [20, 4]), SOD_BC, ic_sod, backend="cuda", basis=basis, degree=k, flux=flux, limiter="TVB",
             char_lim=True, pos_lim=True, M=0.0, beta=2.0, cfl=0.5)
    c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    c.limit_initial()
    flips = 0
    for _ in range(3):
        flips += c.step()[0]
    assert c.rel_err() <= TOL_STEP_SHOCK
    assert flips == 0
    assert np.count_nonzero(c.oracle.limited_flags()) > 0
    c.close()


@pytest.mark.parametrize("basis,k", [("Pk", 1), ("Pk", 2), ("Qk", 1), ("Qk", 2), ("Qk", 3)])
def test_positivity_only(basis, k):
    c = Case(("sod_tube", [20, 6]), SOD_BC, ic_pulse, backend="cuda", basis=basis, degree=k, flux="lxf", pos_lim=True, cfl=0.15)
    c.set_boundary(values=(0.0, 0.0, 0.05, 0.05))
    flips = 0
    for _ in range(6):
        flips += c.step()[0]
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    c.close()


@pytest.mark.parametrize("name,mesh,bc,ic,prm,nsteps", BASELINE_SMALL, ids=[b[0] for b in BASELINE_SMALL])
def test_baseline_configs_small(name, mesh, bc, ic, prm, nsteps):
    c = Case(mesh, bc, ic, backend="cuda", **prm)
    if "dmr" in name:
        c.set_boundary(values=(57.1576766498, -33.0, 8.0, 563.5))
    elif "step" in name:
        c.set_boundary(values=(4.2, 0.0, 1.4, 8.8))
    elif "sod" in name:
        c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    shocked = prm.get("limiter", "none") != "none"
    if shocked:
        c.limit_initial()
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    flips = 0
    for _ in range(nsteps):
        flips += c.step()[0]
    assert c.rel_err() <= (TOL_STEP_SHOCK if shocked else TOL_STEP_SMOOTH)
    if "kfvs" not in name:   # see DESIGN.md: the reference's A&S ERF jumps by 2e-9 at s = 0
        assert flips == 0
    c.close()


@pytest.mark.parametrize("key,size,_", BASELINE_HORIZON, ids=[b[0] for b in BASELINE_HORIZON])
def test_baseline_configs_20_step_horizon(key, size, _):
    """L-infinity against the oracle after 1 RHS / 1 step / 20 steps, all five BASELINE configurations
    (cfg4 with its time-dependent boundary expression), limiter decisions bit for bit."""
    check_horizons(key, size, "cuda")


def test_advance_graph_matches_stagewise():
    """dflo_b200_advance (CUDA-graph replay, dt on the device) == stage-by-stage calls."""
    prm = dict(basis="Qk", degree=3, flux="roe", cfl=0.9)
    a = Case(("isentropic_vortex", [16]), PERIODIC_BOX, ic_vortex, backend="cuda", **prm)
    b = Case(("isentropic_vortex", [16]), PERIODIC_BOX, ic_vortex, backend="cuda", **prm)
    for _ in range(5):
        a.step()
    t, _ = b.engine.advance(5)
    assert abs(t - a.t) <= 1e-12 * a.t
    assert np.abs(a.solution() - b.solution()).max() <= 1e-13
    assert b.engine.launch_count() > 0
    a.close()
    b.close()


def test_set_solution_between_advances_reuses_the_step_graphs():
    """set_solution on a single-GPU ctx keeps the captured step graphs (they depend on the starting buffer only):
    advance - set_solution - advance must equal a fresh ctx bit for bit, whichever buffer the state sits in."""
    prm = dict(basis="Qk", degree=2, flux="hllc", cfl=0.9, limiter="TVB", char_lim=True, M=0.0, beta=2.0)
    a = Case(("sod_tube", [40, 4]), SOD_BC, ic_sod, backend="cuda", **prm)
    b = Case(("sod_tube", [40, 4]), SOD_BC, ic_sod, backend="cuda", **prm)
    for c in (a, b):
        c.set_boundary(values=(0.0, 0.0, 1.0, 2.5))
    ta, _ = a.engine.advance(4)
    ref = a.solution().copy()
    assert np.all(np.isfinite(ref)) and 0 < ta < 1 and np.abs(ref - a.u0).max() > 1e-3
    for n_before in (1, 2, 3):          # leaves the state in each of the rotating buffers
        b.engine.advance(n_before)
        b.engine.set_solution(b.u0)
        tb, _ = b.engine.advance(4)
        assert tb == ta and np.array_equal(b.solution(), ref)
    a.close()
    b.close()


# ---------------------------------------------------------------------------------------------
# BASELINE.json full sizes: no oracle can follow here in seconds, so the checks are the
# size-independent properties of the scheme (conservation, free-stream preservation, translation
# invariance across tile boundaries, quiescent regions staying bit-wise quiescent).
# ---------------------------------------------------------------------------------------------
def _gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def _nodal_ic(nx, ny, x0, x1, y0, y1, k, fn):
    gx, _ = _gauss01(k + 1)
    hx, hy = (x1 - x0) / nx, (y1 - y0) / ny
    xs = x0 + hx * (np.arange(nx)[:, None] + gx[None, :])
    ys = y0 + hy * (np.arange(ny)[:, None] + gx[None, :])
    X = np.broadcast_to(xs[None, :, None, :], (ny, nx, k + 1, k + 1))
    Y = np.broadcast_to(ys[:, None, :, None], (ny, nx, k + 1, k + 1))
    f = fn(X, Y)                                                    # [j][i][b][a][c]
    return np.ascontiguousarray(np.transpose(f, (0, 1, 4, 2, 3))).reshape(-1)


def _engine(kind, args, bc, **prm):
    params, pair = abi.make_params(bc=bc, **prm)
    mesh = abi.Mesh(kind, args)
    flat = mesh.flatten(params, pair)
    return abi.Engine(flat, params), mesh


def test_full_size_cfg2_conservation_freestream_translation():
    """configs[1]: isentropic vortex, Q3, 256x256, Roe, periodic."""
    n, k = 256, 3
    prm = dict(basis="Qk", degree=k, flux="roe", cfl=0.9, compat="mpi")
    eng, _ = _engine("rectangle", [n, n, -5, 5, -5, 5, 4, 2, 1, 3], PERIODIC_BOX, **prm)
    D = eng.D
    # (1) free stream: a uniform state has zero residual and does not move
    W = np.array([0.7, -0.3, 1.3, 2.9])
    u = np.ascontiguousarray(np.broadcast_to(W[None, :, None], (n * n, 4, D // 4))).reshape(-1).copy()
    eng.set_solution(u)
    eng.assemble_rhs(0.0)
    assert np.abs(eng.get_rhs()).max() <= 1e-12
    eng.advance(3)
    assert np.abs(eng.get_solution() - u).max() <= 1e-13
    # (2) conservation on the periodic box: the sum of the cell averages is invariant to round-off
    u0 = _nodal_ic(n, n, -5, 5, -5, 5, k, ic_vortex)
    eng.set_solution(u0)
    s0 = eng.cell_average().sum(axis=0)
    t, _ = eng.advance(10)
    s1 = eng.cell_average().sum(axis=0)
    assert np.abs(s1 - s0).max() <= 1e-12 * np.abs(s0).max()
    ua = eng.get_solution().reshape(n, n, 4, k + 1, k + 1)
    # accuracy: after 10 steps the vortex (advected with u = 0.5) is still the analytic one
    exact = _nodal_ic(n, n, -5, 5, -5, 5, k, lambda x, y: ic_vortex(x - 0.5 * t, y)).reshape(ua.shape)
    assert np.abs(ua - exact).max() <= 2e-4   # h^4-level truncation error at h = 10/256
    # (3) translation invariance: the same problem shifted by 3 cells in x and 5 in y (not a multiple
    # of the 8x4 tile) gives the shifted solution: tile-interior and tile-edge faces evaluate
    # identical Riemann problems; only the periodic seam (integrated from both sides with their own
    # normals, src_mpi/assemble_explicit.cc:186-260) sees different data, a round-off effect
    sx, sy = 3, 5
    h = 10.0 / n
    ub = _nodal_ic(n, n, -5, 5, -5, 5, k, ic_vortex).reshape(n, n, 4, k + 1, k + 1)
    eng.set_solution(np.ascontiguousarray(np.roll(ub, (sy, sx), axis=(0, 1))).reshape(-1))
    eng.advance(10)
    ushift = eng.get_solution().reshape(n, n, 4, k + 1, k + 1)
    assert np.abs(np.roll(ua, (sy, sx), axis=(0, 1)) - ushift).max() <= 1e-13
    assert h > 0
    eng.close()


def test_full_size_cfg4_quiescent_regions_and_limiter():
    """configs[3]: double Mach reflection, Q2, ~1M cells, HLLC + TVB (M = 100).  Ahead of and far
    behind the shock the states are uniform: they must stay exactly uniform (zero residual, limiter
    inactive) while the limiter works on the shock; density and pressure stay positive."""
    ny = 512
    prm = dict(basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True, beta=1.0, M=100.0, cfl=0.9)
    eng, mesh = _engine("double_mach", [ny], DMR_BC, **prm)
    nc, D = mesh.n_cells, eng.D
    assert nc > 1000000
    fa = mesh.flat_arrays()
    xc = fa["origin"][:, 0] + 0.5 * fa["size"][:, 0]
    yc = fa["origin"][:, 1] + 0.5 * fa["size"][:, 1]
    gx, _ = _gauss01(3)
    X = fa["origin"][:, 0, None, None] + fa["size"][:, 0, None, None] * gx[None, None, :]
    Y = fa["origin"][:, 1, None, None] + fa["size"][:, 1, None, None] * gx[None, :, None]
    u0 = np.ascontiguousarray(np.transpose(ic_dmr(np.broadcast_to(X, (nc, 3, 3)), np.broadcast_to(Y, (nc, 3, 3))), (0, 3, 1, 2))).reshape(-1)
    eng.set_solution(u0)
    for b, comp_expr in {3: ("57.1576766498*(x<1.0/6.0+(1+20*t)/sqrt(3))", "-33.0*(x<1.0/6.0+(1+20*t)/sqrt(3))",
                             "8.0*(x<1.0/6.0+(1+20*t)/sqrt(3)) + 1.4*(x>=1.0/6.0+(1+20*t)/sqrt(3))",
                             "563.5*(x<1.0/6.0+(1+20*t)/sqrt(3)) + 2.5*(x>=1.0/6.0+(1+20*t)/sqrt(3))"),
                         4: ("57.1576766498", "-33.0", "8.0", "563.5")}.items():
        for c, e in enumerate(comp_expr):
            eng.set_boundary_expression(b, c, e)
    eng.limit_initial_condition()
    t, _ = eng.advance(4)
    eng.poll_error()
    u = eng.get_solution().reshape(nc, 4, 9)
    # distance of the cell centre to the initial shock line x = 1/6 + y/sqrt(3); it moves < 10*t*2
    d = (xc - 1.0 / 6.0 - yc / np.sqrt(3.0)) * np.cos(np.pi / 6.0)
    margin = 20.0 * t + 24.0 / ny   # shock motion + the 12 stages' numerical domain of dependence (one cell per stage)
    ahead, behind = d > margin, (d < -margin) & (xc > 0.3) & (yc > 0.2)
    assert ahead.sum() > 100000 and behind.sum() > 1000
    pre = np.array([0.0, 0.0, 1.4, 2.5])
    post = np.array([57.1576766498, -33.0, 8.0, 563.5])
    assert np.abs(u[ahead] - pre[None, :, None]).max() <= 1e-12
    assert np.abs(u[behind] - post[None, :, None]).max() <= 1e-10
    flags = eng.limited_flags()
    assert flags[ahead].max() == 0 and np.count_nonzero(flags) > 0
    avg = eng.cell_average()
    assert avg[:, 2].min() > 0 and (0.4 * (avg[:, 3] - 0.5 * (avg[:, 0] ** 2 + avg[:, 1] ** 2) / avg[:, 2])).min() > 0
    eng.close()


def test_full_size_cfg3_mass_conservation():
    """configs[2] at bench size: Sod tube, P2, HLLC, TVB + positivity, 1600x160.  Until the waves
    reach the ends nothing flows through any boundary: mass and energy are conserved to round-off;
    the solution stays y-independent (every row of cells evolves identically)."""
    nx, ny = 1600, 160
    prm = dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.9)
    eng, mesh = _engine("sod_tube", [nx, ny], SOD_BC, **prm)
    nc, D = mesh.n_cells, eng.D
    fa = mesh.flat_arrays()
    xc = fa["origin"][:, 0] + 0.5 * fa["size"][:, 0]
    u0 = np.zeros((nc, 4, D // 4))
    u0[:, 2, 0] = np.where(xc <= 0.5, 1.0, 0.125)    # mode 0 = cell mean (orthonormal Legendre, phi_0 = 1)
    u0[:, 3, 0] = np.where(xc <= 0.5, 2.5, 0.25)
    eng.set_solution(u0.reshape(-1))
    g = np.zeros((eng.n_bfaces, eng.nqf, 4))
    g[...] = (0.0, 0.0, 1.0, 2.5)
    eng.set_boundary_values(g)
    eng.limit_initial_condition()
    s0 = eng.cell_average().sum(axis=0)
    eng.advance(20)
    eng.poll_error()
    avg = eng.cell_average()
    s1 = avg.sum(axis=0)
    assert abs(s1[2] - s0[2]) <= 1e-12 * s0[2] and abs(s1[3] - s0[3]) <= 1e-12 * s0[3]
    rows = avg.reshape(ny, nx, 4)
    assert np.abs(rows[0] - rows[ny // 2]).max() <= 1e-13 and np.abs(rows[0] - rows[-1]).max() <= 1e-13
    assert np.abs(rows[:, :, 1]).max() <= 1e-12      # no y momentum appears
    assert np.count_nonzero(eng.limited_flags()) > 0
    eng.close()


@pytest.mark.parametrize("compat", ["src", "mpi"])
def test_time_dependent_boundary_expression_under_advance(compat):
    """Device-evaluated moving-shock boundary of the double Mach reflection inside the captured step,
    refreshed only when the BC time changes (src: t, then t+dt; src_mpi: t)."""
    from helpers import time_dependent_bc_case
    err, c = time_dependent_bc_case("cuda", compat)
    assert err <= TOL_STEP_SHOCK
    c.close()


@pytest.mark.parametrize("variable", ["density", "energy"])
@pytest.mark.parametrize("basis,k", [("Qk", 1), ("Qk", 2), ("Qk", 3), ("Pk", 2)])
def test_kxrcf_shock_indicator_gates_the_limiter(basis, k, variable):
    """KXRCF shock indicator (src/indicator.cc:50-198) on the device: indicator values, limiter
    decisions and solution against the oracle."""
    from helpers import kxrcf_case
    err, ind_err, flips, c = kxrcf_case("cuda", basis, k, variable)
    assert err <= TOL_STEP_SHOCK and ind_err <= 1e-10 and flips == 0
    flags = c.oracle.limited_flags()
    assert 0 < np.count_nonzero(flags) < flags.size
    c.close()


@pytest.mark.parametrize("k,char_lim,pos_lim", [(1, True, False), (2, True, True), (2, False, False), (3, True, True)])
def test_minmax_limiter(k, char_lim, pos_lim):
    """limiter type = minmax of the MPI tree (src_mpi/limiter.cc:400-553) on the device: limiter decisions
    and solution against the oracle on a rippled moving Sod problem (see tests/test_emu_parity.py)."""
    from helpers import ic_sod_moving_wavy
    c = Case(("sod_tube", [40, 4]), SOD_BC, ic_sod_moving_wavy, backend="cuda", basis="Qk", degree=k, flux="hllc",
             limiter="minmax", char_lim=char_lim, pos_lim=pos_lim, M=0.0, beta=2.0, cfl=0.4)
    c.set_boundary(values=(0.3, 0.1, 1.0, 2.5))
    c.limit_initial()
    # without the characteristic projection the reference's range starts from 0 (src_mpi/limiter.cc:438): minima of
    # density are never limited and this shock problem blows up after three steps, in the oracle as well -- one step
    flips = sum(c.step()[0] for _ in range(3 if char_lim else 1))
    assert np.isfinite(c.oracle.solution()).all()
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    flags = c.oracle.limited_flags()
    assert 0 < np.count_nonzero(flags & 1) < flags.size
    c.close()


@pytest.mark.parametrize("basis,k,flux", [("Qk", 1, "lxf"), ("Qk", 3, "roe"), ("Pk", 2, "hllc"), ("Qk", 2, "kep")])
def test_external_force_mpi(basis, k, flux):
    """External force of the MPI tree (f_0 / f_1 value, src_mpi/assemble_explicit.cc:56-58, 84;
    src_mpi/equation.h:1189-1202): forcing term gravity * (rho f, m.f) with a position-dependent f, right-hand
    side and a step against the oracle; f = (0,-1) reproduces the hard-wired forcing of src/ bit for bit."""
    bc = {1: "inflow", 2: "slip", 3: "pressure", 0: "farfield"}
    c = Case(("forward_step", [0.2]), bc, ic_smooth, backend="cuda", basis=basis, degree=k, flux=flux, cfl=0.5, compat="mpi", gravity=0.7)
    c.set_boundary(values=(0.5, 0.1, 1.2, 3.0), wiggle=0.1)
    r_src = c.rhs_pair()[1].copy()
    c.set_external_force("0.0", "-1.0", lambda x, y: (0.0 * x, -1.0 + 0.0 * x))
    assert np.array_equal(c.rhs_pair()[1], r_src)
    c.set_external_force("0.3*sin(2*x)+y", "-1.0+0.2*x*y", lambda x, y: (0.3 * np.sin(2 * x) + y, -1.0 + 0.2 * x * y))
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    assert np.abs(c.rhs_pair()[1] - r_src).max() > 1e-3
    c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


def test_run_loop_output_schedule_and_vtu_files(tmp_path):
    """ConservationLaw::run with output on (src/claw.cc:1010-1017, 1093-1099; src/output.cc:33-79): initial
    solution, then every `output: iter step` steps, numbered solution-NNN.vtu + shock.vtu, each file equal to
    the host writer applied to the engine's solution at that moment."""
    import ctypes
    import filecmp
    from test_host import PRM_DIR, _claw_api, _read_vtu
    L = _claw_api(abi.load_library())
    L.dflo_claw_set_output.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.dflo_claw_set_output.restype = None
    L.dflo_claw_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, abi.c_double_p, ctypes.POINTER(ctypes.c_int)]
    L.dflo_claw_get_solution.argtypes = [ctypes.c_void_p, abi.c_double_p, ctypes.c_size_t]
    L.dflo_claw_write_vtu.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    over = b"subsection output\n set iter step = 2\n set schlieren plot = true\nend\n"
    h = L.dflo_claw_create(os.path.join(PRM_DIR, "cfg3_sod_P2_hllc_tvb_pos.prm").encode(), b"sod_tube 40 4", over, abi.COMPAT["src"])
    assert h, L.dflo_host_last_error()
    h = ctypes.c_void_p(h)
    out = str(tmp_path) + "/"
    L.dflo_claw_set_output(h, out.encode())
    assert L.dflo_claw_setup(h, 0, 0, 1, None) == 0, L.dflo_host_last_error()
    t, done = ctypes.c_double(0.0), ctypes.c_int(0)
    assert L.dflo_claw_run(h, 5, 0, ctypes.byref(t), ctypes.byref(done)) == 0, L.dflo_host_last_error()
    assert done.value == 5 and t.value > 0
    files = sorted(os.listdir(out))
    assert files == ["shock.vtu", "solution-000.vtu", "solution-001.vtu", "solution-002.vtu"]   # it = 0, 2, 4
    nc, D = 160, 24
    f0, f2 = _read_vtu(out + "solution-000.vtu"), _read_vtu(out + "solution-002.vtu")
    assert f0["n_cells"] == nc * 4 and f0["n_points"] == nc * 9 and f0["point_names"][-1] == "schlieren_plot"
    # initial file: the Sod states (examples/sod_shock_tube/input.prm:45-50), untouched by the limiters
    x = f0["points"][:, 0]
    left = x < 0.5 - 1e-12
    np.testing.assert_allclose(f0["point"]["Density"][left], 1.0, rtol=1e-9)
    np.testing.assert_allclose(f0["point"]["Pressure"][x > 0.5 + 1e-12], 0.1, rtol=1e-9)
    assert np.abs(f2["point"]["XVelocity"]).max() > 0.1                 # the waves have started
    assert np.abs(f2["point"]["Pressure"] - f0["point"]["Pressure"]).max() > 1e-3
    # explicit numbered write now (after step 5) == host writer on the solution read back through the C ABI
    assert L.dflo_claw_write_vtu(h, out.encode()) == 0
    u = np.zeros(nc * D)
    assert L.dflo_claw_get_solution(h, abi._dp(u), u.size) == 0
    mesh = abi.Mesh(handle=L.dflo_claw_mesh(h), owned=False)
    mesh.write_solution_vtu(out + "check.vtu", u, "Pk", 2, schlieren_plot=True, time=t.value, cycle=3)
    assert filecmp.cmp(out + "solution-003.vtu", out + "check.vtu", shallow=False)
    am = ctypes.c_double(0.0)                                          # compute_angular_momentum, src/claw.cc:604-635
    L.dflo_claw_angular_momentum.argtypes = [ctypes.c_void_p, abi.c_double_p]
    assert L.dflo_claw_angular_momentum(h, ctypes.cast(ctypes.byref(am), abi.c_double_p)) == 0
    assert am.value == mesh.angular_momentum(u, "Pk", 2) and am.value < -1e-5
    sh = _read_vtu(out + "shock.vtu")
    assert sh["n_cells"] == nc and sh["point_names"] == ["mu_shock", "shock_indicator"]
    assert np.all(sh["point"]["shock_indicator"] == 1e20)               # `shock indicator = limiter`, src/indicator.cc:15-31
    L.dflo_claw_destroy(h)
    # MPI tree (src_mpi/output.cc:34-86): output/solution-NNNN.RRR.vtu with "subdomain" + master_file.visit, no shock.vtu
    out2 = out + "mpi/"
    os.makedirs(out2)
    h = ctypes.c_void_p(L.dflo_claw_create(os.path.join(PRM_DIR, "cfg1_isentropic_vortex_Q1_lxf.prm").encode(),
                                           b"isentropic_vortex 8", b"subsection output\n set iter step = 1\nend\n", abi.COMPAT["mpi"]))
    assert h, L.dflo_host_last_error()
    L.dflo_claw_set_output(h, out2.encode())
    assert L.dflo_claw_setup(h, 0, 0, 1, None) == 0, L.dflo_host_last_error()
    assert L.dflo_claw_run(h, 2, 0, ctypes.byref(t), ctypes.byref(done)) == 0, L.dflo_host_last_error()
    assert sorted(os.listdir(out2)) == ["master_file.visit", "output"]
    assert sorted(os.listdir(out2 + "output")) == ["solution-%04d.000.vtu" % i for i in range(3)]
    assert open(out2 + "master_file.visit").read().split() == ["!NBLOCKS", "1"] + ["output/solution-%04d.000.vtu" % i for i in range(3)]
    L.dflo_claw_destroy(h)
    # "output: format = tecplot" (src/output.cc:51-52, 65-66, 80-84): solution-NNN.plt + shock.plt
    out3 = out + "plt/"
    os.makedirs(out3)
    h = ctypes.c_void_p(L.dflo_claw_create(os.path.join(PRM_DIR, "cfg3_sod_P2_hllc_tvb_pos.prm").encode(), b"sod_tube 20 2",
                                           b"subsection output\n set iter step = 1\n set format = tecplot\nend\n", abi.COMPAT["src"]))
    assert h, L.dflo_host_last_error()
    L.dflo_claw_set_output(h, out3.encode())
    assert L.dflo_claw_setup(h, 0, 0, 1, None) == 0, L.dflo_host_last_error()
    assert L.dflo_claw_run(h, 1, 0, ctypes.byref(t), ctypes.byref(done)) == 0, L.dflo_host_last_error()
    assert sorted(os.listdir(out3)) == ["shock.plt", "solution-000.plt", "solution-001.plt"]
    head = open(out3 + "solution-001.plt").read(600)
    assert '"XMomentum", "YMomentum", "Density", "Energy", "XVelocity", "YVelocity", "Pressure"' in head
    assert "f=feblock, n=360, e=160, et=quadrilateral" in head
    f = _read_vtu(out2 + "output/solution-0002.000.vtu")
    assert f["n_cells"] == 64 and f["point_names"][-2:] == ["Pressure", "subdomain"] and np.all(f["point"]["subdomain"] == 0)
    L.dflo_claw_destroy(h)


# ---------------------------------------------------------------------------------------------
# mapping = q1: the mapped stage kernel on straight-sided general quadrilaterals
# ---------------------------------------------------------------------------------------------
def _q1_case(k, flux, rotate, bc, ic, n=8, **extra):
    ids = (4, 2, 1, 3)
    cfl = extra.pop("cfl", 0.05 if flux == "kep" else 0.3)
    return Case(("rectangle_skew", [n, n, -5, 5, -5, 5, *ids, 0.15, rotate]), bc, ic, backend="cuda", basis="Qk", degree=k, flux=flux,
                cfl=cfl, mapping="q1", **extra)


@pytest.mark.parametrize("rotate", [0, 1])
@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_q1_mapping_rhs_and_steps_periodic(k, flux, rotate):
    c = _q1_case(k, flux, rotate, PERIODIC_BOX, ic_vortex, compat="mpi")
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    for _ in range(3):
        _, dt_o, dt_e = c.step()
        assert abs(dt_o - dt_e) <= 1e-12 * dt_o
    assert c.rel_err() <= TOL_STEP_SMOOTH
    # whole steps on the device (graph replay, compute_time_step_q on the device) continue the same trajectory
    o = c.oracle
    t = c.t
    for _ in range(2):
        dt = o.compute_dt(t)
        for rk in range(o.n_rk):
            assert o.rk_stage(rk, dt)[0] == 0
        o.commit_step()
        t += dt
    te, _ = c.engine.advance(2, elapsed=c.t)
    assert abs(te - t) <= 1e-12 * t and c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("k,ic", [(1, ic_pulse_box), (2, ic_disc_box), (3, ic_disc_box)])
def test_q1_mapping_positivity_limiter(k, ic):
    """The positivity limiter on mapped cells (positivity.cc works on unit-cell point values and the mapped cell average;
    parameters.cc:536-550 refuses only TVB and Pk off Cartesian grids): it must act -- density and pressure stage -- with
    the oracle's decisions in every cell, through rk_stage and through whole steps on the device."""
    c = _q1_case(k, "lxf", 1, PERIODIC_BOX, ic, n=10, compat="mpi", pos_lim=True, cfl=0.15)
    acted, flips = 0, 0
    for _ in range(4):
        flips += c.step()[0]
        acted |= int(np.bitwise_or.reduce(c.oracle.limited_flags()))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    assert acted & 2 and acted & 4, "the limiter never acted: the case does not test it"
    o, t = c.oracle, c.t
    for _ in range(2):
        dt = o.compute_dt(t)
        for rk in range(o.n_rk):
            assert o.rk_stage(rk, dt)[0] == 0
        o.commit_step()
        t += dt
    te, _ = c.engine.advance(2, elapsed=c.t)
    c.engine.poll_error()
    assert abs(te - t) <= 1e-12 * t and c.rel_err() <= TOL_STEP_SHOCK
    c.close()


@pytest.mark.parametrize("compat", ["src", "mpi"])
@pytest.mark.parametrize("k,flux", [(1, "lxf"), (2, "roe"), (3, "hllc"), (2, "kfvs")])
def test_q1_mapping_all_boundary_kinds_gravity(k, flux, compat):
    bc = {1: "inflow", 2: "slip", 3: "pressure", 4: "farfield"}
    c = _q1_case(k, flux, 1, bc, ic_smooth, compat=compat, gravity=0.7)
    c.set_boundary(values=(1.0, 0.2, 1.4, 8.8), wiggle=0.05)
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    for _ in range(2):
        c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


@pytest.mark.parametrize("basis,k,flux,mapping", [("Qk", 1, "kfvs", "cartesian"), ("Qk", 3, "roe", "cartesian"), ("Qk", 2, "hllc", "cartesian"),
                                                  ("Pk", 2, "hllc", "cartesian"), ("Pk", 3, "lxf", "cartesian"), ("Qk", 1, "kfvs", "q1"),
                                                  ("Qk", 2, "hllc", "q1")])
def test_local_time_stepping(basis, k, flux, mapping):
    """time step type = local (src/claw.cc:444-478, 694-713) through the row kernel, the Pk cell kernel, the tile kernel
    and the mapped kernel."""
    from helpers import local_time_stepping_case
    local_time_stepping_case("cuda", basis, k, flux, mapping)


def test_compression_corner_q1_local():
    """examples/compression_corner (Q1, KFVS, mapping = q1, time step type = local) at the size of the shipped .geo."""
    from helpers import compression_corner_case
    compression_corner_case("cuda", size=(9, 29, 19), nsteps=20)


def test_compression_corner_driver_run(tmp_path):
    """The standalone front end on the q1 example deck: setup, 30 local-time steps, a VTU file whose points are the
    MAPPED patch vertices (all inside the channel, the ramp cells above the ramp)."""
    import ctypes
    import os
    from test_host import ROOT, _claw_api, _read_vtu
    L = _claw_api(abi.load_library())
    L.dflo_claw_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, abi.c_double_p, ctypes.POINTER(ctypes.c_int)]
    L.dflo_claw_write_vtu.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    prm = os.path.join(ROOT, "tests", "golden", "prm_q1", "compression_corner_Q1_kfvs_q1_local.prm")
    h = ctypes.c_void_p(L.dflo_claw_create(prm.encode(), b"compression_corner 9 29 19", None, abi.COMPAT["src"]))
    assert h, L.dflo_host_last_error()
    assert L.dflo_claw_setup(h, 0, 0, 1, None) == 0, L.dflo_host_last_error()
    t, done = ctypes.c_double(0.0), ctypes.c_int(0)
    assert L.dflo_claw_run(h, 30, 0, ctypes.byref(t), ctypes.byref(done)) == 0, L.dflo_host_last_error()
    assert done.value == 30 and t.value > 0
    path = str(tmp_path / "corner.vtu")
    assert L.dflo_claw_write_vtu(h, path.encode()) == 0, L.dflo_host_last_error()
    f = _read_vtu(path)
    x, y = f["points"][:, 0], f["points"][:, 1]
    ramp = np.where(x > 1.0, np.tan(np.radians(9.5)) * (x - 1.0), 0.0)
    assert x.min() >= -1e-12 and x.max() <= 5.0 + 1e-12 and np.all(y >= ramp - 1e-9) and y.max() <= 3.0 + 1e-12
    assert np.any(np.abs(y - ramp) < 1e-9) and np.all(f["point"]["Density"] > 0.5)
    assert f["point"]["Density"].max() > 1.05          # the oblique shock off the ramp is forming
    L.dflo_claw_destroy(h)


# ---------------------------------------------------------------------------------------------
# faces with hanging nodes: sub-face integration in the mapped stage kernel
# ---------------------------------------------------------------------------------------------
def _refined_case(k, flux, mapping, bc, ic, patch=(2, 5, 1, 4), n=(7, 6), rotate=0, **extra):
    ids = (4, 2, 1, 3)
    cfl = extra.pop("cfl", 0.05 if flux == "kep" else 0.3)
    return Case(("rectangle_refined", [n[0], n[1], -5, 5, -5, 5, *ids, *patch, rotate]), bc, ic, backend="cuda", basis="Qk", degree=k,
                flux=flux, cfl=cfl, mapping=mapping, **extra)


@pytest.mark.parametrize("k,mapping,rotate", [(1, "cartesian", 0), (2, "cartesian", 0), (3, "cartesian", 0), (1, "q1", 1), (2, "q1", 1)])
def test_hanging_nodes_positivity_limiter(k, mapping, rotate):
    """The positivity limiter next to hanging nodes (local to a cell; the refined patch's rim cuts the edge of the dense
    disc): both limiter stages act with the oracle's decision in every cell."""
    c = _refined_case(k, "lxf", mapping, PERIODIC_BOX, ic_pulse_box if k == 1 else ic_disc_box, patch=(4, 8, 3, 7), n=(10, 10), rotate=rotate,
                      compat="mpi", pos_lim=True, cfl=0.1)
    acted, flips = 0, 0
    for _ in range(4):
        flips += c.step()[0]
        acted |= int(np.bitwise_or.reduce(c.oracle.limited_flags()))
    assert c.rel_err() <= TOL_STEP_SHOCK and flips == 0
    assert acted & 4 and (acted & 2 or k == 1), "the limiter never acted: the case does not test it"
    c.close()


@pytest.mark.parametrize("mapping,rotate", [("cartesian", 0), ("q1", 0), ("q1", 1)])
@pytest.mark.parametrize("flux", ALL_FLUXES)
@pytest.mark.parametrize("k", [0, 1, 2, 3, 4])
def test_hanging_nodes_rhs_and_steps(k, flux, mapping, rotate):
    c = _refined_case(k, flux, mapping, PERIODIC_BOX, ic_vortex, rotate=rotate, compat="mpi")
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    for _ in range(3):
        _, dt_o, dt_e = c.step()
        assert abs(dt_o - dt_e) <= 1e-12 * dt_o
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()


def test_hanging_nodes_conservation_and_boundaries():
    c = _refined_case(2, "roe", "cartesian", PERIODIC_BOX, ic_vortex, compat="mpi")
    v, cells, _, _ = c.mesh.primitive()
    area = (v[cells[:, 1], 0] - v[cells[:, 0], 0]) * (v[cells[:, 2], 1] - v[cells[:, 0], 1])
    gx, gw = c.oracle.tables()
    w = (gw[None, :] * gw[:, None]).reshape(-1)
    total = lambda u: (u.reshape(len(cells), 4, -1) * w[None, None, :] * area[:, None, None]).sum(axis=(0, 2))
    t0 = total(c.solution())
    te, _ = c.engine.advance(5)
    assert te > 0 and np.abs(total(c.solution()) - t0).max() <= 1e-12 * np.abs(t0).max()
    c.close()
    bc = {1: "inflow", 2: "slip", 3: "pressure", 4: "farfield"}
    c = _refined_case(3, "hllc", "q1", bc, ic_smooth, patch=(0, 3, 2, 6), rotate=1, gravity=0.7)
    c.set_boundary(values=(1.0, 0.2, 1.4, 8.8), wiggle=0.05)
    r_o, r_e = c.rhs_pair()
    assert np.abs(r_o - r_e).max() <= TOL_RHS * max(1.0, np.abs(r_o).max()) * np.sqrt(c.oracle.D)
    for _ in range(2):
        c.step()
    assert c.rel_err() <= TOL_STEP_SMOOTH
    c.close()
