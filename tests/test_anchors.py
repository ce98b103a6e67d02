"""Reference-independent anchors (test matrix): results that do not depend on anybody's reading
of deal.II's conventions.

 * h-convergence of the advected isentropic vortex (exact solution: src_mpi/ic.cc:44-61 translated with u_inf = 0.5;
   src/ic.cc:44-61 is the stationary form) -- L2 error of the density at the Gauss points must fall with order k+1;
 * the Sod shock tube of examples/sod_shock_tube (states: state.m; configuration: BASELINE configs[2]) against the
   exact solution of the Riemann problem -- the L1 error of the density must fall under refinement, at the ~1st order
   a limited scheme has on a discontinuous solution.

Both run on the CPU oracle (this proves the restatement of the assembly / RK / limiter level, which no golden vector of
the reference pins) and, in the GPU tier, on the CUDA library through the C ABI.  A wrong quadrature weight, trace
table, lifting sign, RK coefficient or inverse mass shows up as a lost order, whatever the oracle and the kernels
agree on between themselves.
"""
import numpy as np
import pytest

from helpers import PERIODIC_BOX, SOD_BC, abi, gpu_available
from oracle import oracle as O

GAMMA = 1.4


# ---------------------------------------------------------------------------------------------
# exact solutions
# ---------------------------------------------------------------------------------------------
def vortex_exact(x, y, t):
    """src_mpi/ic.cc:44-61 at time t: the vortex is carried along x with the free stream (M_inf = 0.5, a_inf = 1);
    periodic images beyond [-5,5] change the density by < 1e-10."""
    dx = (x - 0.5 * t + 5.0) % 10.0 - 5.0
    return O.isentropic_vortex(dx, y, compat="mpi")


def sod_exact_density(x, t, x0=0.5, left=(1.0, 0.0, 1.0), right=(0.125, 0.0, 0.1)):
    """Exact solution of the Riemann problem (Toro, ch. 4) for examples/sod_shock_tube/state.m: left rarefaction,
    contact, right shock.  Newton iteration on the pressure function for the star state."""
    g = GAMMA
    rl, ul, pl = left
    rr, ur, pr = right
    al, ar = np.sqrt(g * pl / rl), np.sqrt(g * pr / rr)

    def f(p, r, pk, a):
        if p > pk:      # shock
            A, B = 2.0 / ((g + 1) * r), (g - 1) / (g + 1) * pk
            return (p - pk) * np.sqrt(A / (p + B)), np.sqrt(A / (p + B)) * (1.0 - 0.5 * (p - pk) / (B + p))
        return 2 * a / (g - 1) * ((p / pk) ** ((g - 1) / (2 * g)) - 1.0), 1.0 / (r * a) * (p / pk) ** (-(g + 1) / (2 * g))

    p = 0.5 * (pl + pr)
    for _ in range(50):
        fl, dfl = f(p, rl, pl, al)
        fr, dfr = f(p, rr, pr, ar)
        dp = (fl + fr + ur - ul) / (dfl + dfr)
        p -= dp
        if abs(dp) < 1e-15 * p:
            break
    us = 0.5 * (ul + ur) + 0.5 * (f(p, rr, pr, ar)[0] - f(p, rl, pl, al)[0])
    rsl = rl * (p / pl) ** (1.0 / g)                                        # behind the rarefaction
    rsr = rr * ((p / pr + (g - 1) / (g + 1)) / ((g - 1) / (g + 1) * p / pr + 1.0))   # behind the shock
    asl = al * (p / pl) ** ((g - 1) / (2 * g))
    s_head, s_tail = ul - al, us - asl
    s_shock = ur + ar * np.sqrt((g + 1) / (2 * g) * p / pr + (g - 1) / (2 * g))
    xi = (x - x0) / t
    fan = rl * (2.0 / (g + 1) + (g - 1) / ((g + 1) * al) * (ul - xi)) ** (2.0 / (g - 1))
    return np.where(xi < s_head, rl, np.where(xi < s_tail, fan, np.where(xi < us, rsl, np.where(xi < s_shock, rsr, rr))))


def test_sod_exact_solution_known_values():
    """Toro, table 4.2, test 1: p* = 0.30313, u* = 0.92745, rho*L = 0.42632, rho*R = 0.26557."""
    t = 0.2
    rho = sod_exact_density(np.array([0.0, 0.5 + 0.5 * t * 0.92745 * 0.99, 0.5 + 0.5 * t * (0.92745 + 1.75216), 1.0]), t)
    assert abs(rho[0] - 1.0) < 1e-14 and abs(rho[3] - 0.125) < 1e-14
    assert abs(rho[1] - 0.42632) < 1e-5 and abs(rho[2] - 0.26557) < 1e-5


# ---------------------------------------------------------------------------------------------
# evaluation of a solution vector at the cell Gauss points
# ---------------------------------------------------------------------------------------------
def density_at_gauss_points(u, basis, k, gx):
    """u: [n_cells][4][n_s] in the reference DoF layout.  Qk: the DoFs are the values (collocated Gauss nodes);
    Pk: modes of the orthonormal Legendre basis, y-degree outer / x-degree inner."""
    if basis == "Qk":
        return u[:, 2, :]
    from numpy.polynomial import legendre as Lg
    P = [np.sqrt(2 * i + 1) * Lg.legval(2 * gx - 1, [0] * i + [1]) for i in range(k + 1)]
    rho = np.zeros((u.shape[0], (k + 1) ** 2))
    m = 0
    for j in range(k + 1):
        for i in range(k + 1 - j):
            rho += u[:, 2, m][:, None] * (P[j][:, None] * P[i][None, :]).reshape(-1)[None, :]   # point q = a + (k+1) b
            m += 1
    return rho


class _OracleRunner:
    """advance-to-time on the CPU oracle"""

    def __init__(self, mesh, bc, prm, u_of_xy, g=None):
        okw = dict(prm)
        self.o = O.Oracle(*mesh, O.make_params(bc=bc, n_threads=8, **okw))
        xq = self.o.cell_qpoints()
        self.o.set_initial_condition(u_of_xy(xq[..., 0], xq[..., 1]))
        self.o.compute_cell_average()
        if g is not None and self.o.n_bfaces:
            gv = np.zeros((self.o.n_bfaces, self.o.nqf, 4))
            gv[...] = np.asarray(g)
            self.o.set_bc_values(gv)
        if prm.get("limiter", "none") != "none":
            self.o.apply_limiter()
            self.o.commit_step()

    def run_to(self, T):
        o, t = self.o, 0.0
        while t < T * (1.0 - 1e-14):
            dt = o.compute_dt(t, T)
            for rk in range(o.n_rk):
                assert o.rk_stage(rk, dt)[0] == 0
            o.commit_step()
            t += dt
        return o.solution()


class _EngineRunner:
    """advance-to-time on the CUDA library (dflo_b200_advance: dt on the device, clipped at the final time)"""

    def __init__(self, mesh_gen, bc, prm, u_of_xy, g=None):
        self.params, pair = abi.make_params(bc=bc, **prm)
        self.mesh = abi.Mesh(mesh_gen[0], mesh_gen[1])
        flat = self.mesh.flatten(self.params, pair)
        self.e = abi.Engine(flat, self.params)
        # initial DoFs: the oracle's interpolation / projection of the same point values
        o = O.Oracle(*self.mesh.primitive(), O.make_params(bc=bc, **prm))
        xq = o.cell_qpoints()
        o.set_initial_condition(u_of_xy(xq[..., 0], xq[..., 1]))
        self.xq, self.tables, self.n_cells = xq, o.tables(), o.n_cells
        self.e.set_solution(o.solution())
        if g is not None and o.n_bfaces:
            gv = np.zeros((o.n_bfaces, o.nqf, 4))
            gv[...] = np.asarray(g)
            self.e.set_boundary_values(gv)
        if prm.get("limiter", "none") != "none":
            self.e.limit_initial_condition()

    def run_to(self, T):
        t = 0.0
        while t < T * (1.0 - 1e-14):
            t, _ = self.e.advance(50, elapsed=t, final_time=T)
        u = self.e.get_solution()
        self.e.close()
        return u


VORTEX = [("Qk", 1, (16, 32, 64), 0.5, 0.5), ("Qk", 2, (16, 32, 64), 0.5, 0.4), ("Qk", 3, (16, 32, 64), 0.25, 0.2),
          ("Pk", 1, (16, 32, 64), 0.5, 0.5), ("Pk", 2, (16, 32, 64), 0.5, 0.4)]


def _vortex_l2(u, basis, k, n, xq, tables, T):
    gx, gw = tables
    rho = density_at_gauss_points(u.reshape(n * n, 4, -1), basis, k, gx)
    ex = vortex_exact(xq[..., 0], xq[..., 1], T)[..., 2]
    w = (gw[:, None] * gw[None, :]).reshape(-1) * (10.0 / n) ** 2
    return float(np.sqrt(((rho - ex) ** 2 * w[None, :]).sum()))


def _check_orders(errs, k, what, upwind=True):
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    # upwind-type fluxes (Roe): design order k+1 on the finest pair (>= k + 0.8), nothing pre-asymptotic below k + 0.5;
    # the Lax-Friedrichs flux only guarantees k + 1/2 (observed: Q2 2.68, odd degrees k + 1)
    if upwind:
        assert rates[-1] >= k + 0.8 and min(rates) >= k + 0.5, (what, errs, rates)
    else:
        assert rates[-1] >= k + 0.5 and min(rates) >= k + 0.4, (what, errs, rates)
    return rates


@pytest.mark.parametrize("basis,k,ns,T,cfl", VORTEX)
def test_vortex_h_convergence_oracle(basis, k, ns, T, cfl):
    prm = dict(basis=basis, degree=k, flux="roe", cfl=cfl, compat="mpi")
    errs = []
    for n in ns:
        r = _OracleRunner(O.rect_mesh(n, n, -5, 5, -5, 5), PERIODIC_BOX, prm, lambda x, y: vortex_exact(x, y, 0.0))
        u = r.run_to(T)
        errs.append(_vortex_l2(u, basis, k, n, r.o.cell_qpoints(), r.o.tables(), T))
    _check_orders(errs, k, "oracle %s%d" % (basis, k))


@pytest.mark.gpu
@pytest.mark.parametrize("flux", ["roe", "lxf"])
@pytest.mark.parametrize("basis,k,ns,T,cfl", VORTEX)
def test_vortex_h_convergence_gpu(basis, k, ns, T, cfl, flux):
    assert gpu_available()
    prm = dict(basis=basis, degree=k, flux=flux, cfl=cfl, compat="mpi")
    errs = []
    for n in tuple(ns) + ((128,) if k < 3 else ()):
        r = _EngineRunner(("isentropic_vortex", [n]), PERIODIC_BOX, prm, lambda x, y: vortex_exact(x, y, 0.0))
        u = r.run_to(T)
        errs.append(_vortex_l2(u, basis, k, n, r.xq, r.tables, T))
    _check_orders(errs, k, "gpu %s%d %s" % (basis, k, flux), upwind=flux != "lxf")


SOD_PRM = dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.9)
SOD_T = 0.1


def _sod_l1(u, k, nx, ny, xq, tables):
    gx, gw = tables
    rho = density_at_gauss_points(u.reshape(nx * ny, 4, -1), "Pk", k, gx)
    ex = sod_exact_density(xq[..., 0], SOD_T)
    w = (gw[:, None] * gw[None, :]).reshape(-1) * (1.0 / nx) * (0.1 / ny)
    return float((np.abs(rho - ex) * w[None, :]).sum() / 0.1)        # per unit height: error of the 1-D profile


def _sod_ic(x, y):
    rho = np.where(x <= 0.5, 1.0, 0.125)
    return np.stack([0 * x, 0 * x, rho, np.where(x <= 0.5, 2.5, 0.25)], axis=-1)


def _check_sod(errs, what):
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    # a limited scheme on a solution with a contact and a shock: L1 order between ~0.7 and 1
    assert all(r > 0.6 for r in rates) and errs[0] < 0.02, (what, errs, rates)


def test_sod_vs_exact_riemann_solution_oracle():
    errs = []
    for nx in (50, 100, 200):
        ny = nx // 10
        r = _OracleRunner(O.rect_mesh(nx, ny, 0.0, 1.0, 0.0, 0.1, ids=(2, 1, 0, 0)), SOD_BC, SOD_PRM, _sod_ic, g=(0.0, 0.0, 1.0, 2.5))
        u = r.run_to(SOD_T)
        errs.append(_sod_l1(u, 2, nx, ny, r.o.cell_qpoints(), r.o.tables()))
    _check_sod(errs, "oracle")


@pytest.mark.gpu
def test_sod_vs_exact_riemann_solution_gpu():
    assert gpu_available()
    errs = []
    for nx in (100, 200, 400, 800):
        ny = nx // 10
        r = _EngineRunner(("sod_tube", [nx, ny]), SOD_BC, SOD_PRM, _sod_ic, g=(0.0, 0.0, 1.0, 2.5))
        u = r.run_to(SOD_T)
        errs.append(_sod_l1(u, 2, nx, ny, r.xq, r.tables))
    _check_sod(errs, "gpu")


# ---------------------------------------------------------------------------------------------
# mapping = q1: the same vortex on smoothly skewed quadrilaterals
# ---------------------------------------------------------------------------------------------
Q1_VORTEX = [(1, (16, 32, 64), 0.5, 0.25), (2, (16, 32, 64), 0.5, 0.2), (3, (16, 32, 64), 0.25, 0.1)]   # cfl of compute_time_step_q is ~2x looser


def _skew_args(n, rotate=0):
    return [n, n, -5, 5, -5, 5, 4, 2, 1, 3, 0.15, rotate]


def _q1_l2(u, k, xq, jxw, T):
    rho = u.reshape(xq.shape[0], 4, -1)[:, 2, :]
    ex = vortex_exact(xq[..., 0], xq[..., 1], T)[..., 2]
    return float(np.sqrt(((rho - ex) ** 2 * jxw).sum()))


def _q1_jxw(verts, cells, gx, gw):
    """w_a w_b det J at the Gauss nodes of every cell (bilinear map of the four vertices)."""
    v = verts[cells]                                         # [nc][4][2]
    xi, eta = np.meshgrid(gx, gx, indexing="xy")             # node q = a + n1 b: xi = gx[a], eta = gx[b]
    xi, eta = xi.reshape(-1), eta.reshape(-1)
    w = (gw[None, :] * gw[:, None]).reshape(-1)
    d = lambda i, j, t: (v[:, i, t] - v[:, j, t])[:, None]
    xxi = d(1, 0, 0) * (1 - eta) + d(3, 2, 0) * eta
    xeta = d(2, 0, 0) * (1 - xi) + d(3, 1, 0) * xi
    yxi = d(1, 0, 1) * (1 - eta) + d(3, 2, 1) * eta
    yeta = d(2, 0, 1) * (1 - xi) + d(3, 1, 1) * xi
    return w[None, :] * (xxi * yeta - xeta * yxi)


@pytest.mark.parametrize("k,ns,T,cfl", Q1_VORTEX)
def test_q1_vortex_h_convergence_oracle(k, ns, T, cfl):
    prm = dict(basis="Qk", degree=k, flux="roe", cfl=cfl, compat="mpi", mapping="q1")
    errs = []
    for n in ns[:2] if k == 3 else ns:
        mesh = abi.Mesh("rectangle_skew", _skew_args(n))
        v, c, bl, bi = mesh.primitive()
        r = _OracleRunner((v, c, bl, bi), PERIODIC_BOX, prm, lambda x, y: vortex_exact(x, y, 0.0))
        u = r.run_to(T)
        gx, gw = r.o.tables()
        errs.append(_q1_l2(u, k, r.o.cell_qpoints(), _q1_jxw(v, c, gx, gw), T))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    assert rates[-1] >= k + (0.6 if k == 3 else 0.8), (errs, rates)     # Q3 16 -> 32 is pre-asymptotic (3.7), as on Cartesian cells


@pytest.mark.gpu
@pytest.mark.parametrize("rotate", [0, 1])
@pytest.mark.parametrize("k,ns,T,cfl", Q1_VORTEX)
def test_q1_vortex_h_convergence_gpu(k, ns, T, cfl, rotate):
    assert gpu_available()
    prm = dict(basis="Qk", degree=k, flux="roe", cfl=cfl, compat="mpi", mapping="q1")
    errs = []
    for n in ns:
        r = _EngineRunner(("rectangle_skew", _skew_args(n, rotate)), PERIODIC_BOX, prm, lambda x, y: vortex_exact(x, y, 0.0))
        v, c, _, _ = r.mesh.primitive()
        u = r.run_to(T)
        errs.append(_q1_l2(u, k, r.xq, _q1_jxw(v, c, *r.tables), T))
    _check_orders(errs, k, "gpu q1 Q%d rotate %d" % (k, rotate))


# ---------------------------------------------------------------------------------------------
# faces with hanging nodes: the vortex through a once-refined patch whose rim cuts the core
# ---------------------------------------------------------------------------------------------
def _refined_args(n):
    return [n, n, -5, 5, -5, 5, 4, 2, 1, 3, n // 4, n // 2, n // 4, 3 * n // 4]


def _rect_l2(u, v, c, gx_gw, xq, T):
    gx, gw = gx_gw
    area = (v[c[:, 1], 0] - v[c[:, 0], 0]) * (v[c[:, 2], 1] - v[c[:, 0], 1])
    w = (gw[None, :] * gw[:, None]).reshape(-1)
    rho = u.reshape(len(c), 4, -1)[:, 2, :]
    ex = vortex_exact(xq[..., 0], xq[..., 1], T)[..., 2]
    return float(np.sqrt((((rho - ex) ** 2) * w[None, :] * area[:, None]).sum()))


HANGING_VORTEX = [(1, (16, 32, 64), 0.5, 0.4), (2, (16, 32, 64), 0.5, 0.3), (3, (16, 32, 64), 0.25, 0.15)]


@pytest.mark.parametrize("k,ns,T,cfl", HANGING_VORTEX)
def test_hanging_nodes_vortex_h_convergence_oracle(k, ns, T, cfl):
    prm = dict(basis="Qk", degree=k, flux="roe", cfl=cfl, compat="mpi")
    errs = []
    for n in ns[:2] if k == 3 else ns:
        mesh = abi.Mesh("rectangle_refined", _refined_args(n))
        v, c, bl, bi = mesh.primitive()
        r = _OracleRunner((v, c, bl, bi), PERIODIC_BOX, prm, lambda x, y: vortex_exact(x, y, 0.0))
        u = r.run_to(T)
        errs.append(_rect_l2(u, v, c, r.o.tables(), r.o.cell_qpoints(), T))
    rates = [np.log2(errs[i] / errs[i + 1]) for i in range(len(errs) - 1)]
    assert rates[-1] >= k + (0.6 if k == 3 else 0.8), (errs, rates)


@pytest.mark.gpu
@pytest.mark.parametrize("k,ns,T,cfl", HANGING_VORTEX)
def test_hanging_nodes_vortex_h_convergence_gpu(k, ns, T, cfl):
    assert gpu_available()
    prm = dict(basis="Qk", degree=k, flux="roe", cfl=cfl, compat="mpi")
    errs = []
    for n in ns:
        r = _EngineRunner(("rectangle_refined", _refined_args(n)), PERIODIC_BOX, prm, lambda x, y: vortex_exact(x, y, 0.0))
        v, c, _, _ = r.mesh.primitive()
        u = r.run_to(T)
        errs.append(_rect_l2(u, v, c, r.tables, r.xq, T))
    _check_orders(errs, k, "gpu hanging Q%d" % k)
