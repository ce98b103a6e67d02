"""TEST INFRASTRUCTURE ONLY. ctypes binding of the CPU oracle (oracle/dflo_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package dflo_b200 never does.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

FLUX = {"lxf": 0, "sw": 1, "kfvs": 2, "roe": 3, "hllc": 4, "kep": 5}   # src/parameters.h:229; kep: src_mpi/parameters.cc:179
BC = {"inflow": 0, "outflow": 1, "slip": 2, "pressure": 3, "farfield": 4, "periodic": 5}
QK, PK = 0, 1


class OracleParams(ctypes.Structure):
    _fields_ = [
        ("basis", ctypes.c_int), ("degree", ctypes.c_int), ("flux_type", ctypes.c_int),
        ("limiter_type", ctypes.c_int), ("char_lim", ctypes.c_int), ("pos_lim", ctypes.c_int),
        ("conserve_angular_momentum", ctypes.c_int),
        ("M", ctypes.c_double), ("beta", ctypes.c_double), ("gravity", ctypes.c_double),
        ("cfl", ctypes.c_double),
        ("bc_kind", ctypes.c_int * 10), ("periodic_pair", ctypes.c_int * 10),
        ("compat", ctypes.c_int), ("n_threads", ctypes.c_int), ("shock_indicator", ctypes.c_int),
        ("mapping", ctypes.c_int), ("local_time_step", ctypes.c_int),
    ]


def build(ref=False):
    """Compile the oracle (and, where /root/reference exists, the reference-backed variants)."""
    targets = ["all"]
    if ref and os.path.exists("/root/reference/src/equation.h"):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", _HERE] + targets, check=True)


def lib_path(variant="restated"):
    if variant == "restated":
        return os.path.join(_HERE, "_build", "liboracle.so")
    if variant == "refphys":
        return os.path.join(_HERE, "_ref", "liboracle_refphys.so")
    if variant == "physref":
        return os.path.join(_HERE, "_ref", "libphys_reference.so")
    raise ValueError(variant)


_libs = {}
_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def _d(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_dp)


def _i(a):
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(_ip)


def load(variant="restated"):
    if variant in _libs:
        return _libs[variant]
    path = lib_path(variant)
    if not os.path.exists(path) and variant == "restated":
        build()
    L = ctypes.CDLL(path)
    for name in ("phys_pressure", "phys_sound_speed", "phys_max_eigenvalue"):
        getattr(L, name).restype = ctypes.c_double
    L.phys_impl_name.restype = ctypes.c_char_p
    if variant != "physref":
        L.oracle_create.restype = ctypes.c_void_p
        L.oracle_last_error.restype = ctypes.c_char_p
        L.oracle_ark.restype = ctypes.c_double
        L.oracle_compute_dt.restype = ctypes.c_double
        L.oracle_run_steps.restype = ctypes.c_double
    _libs[variant] = L
    return L


class Physics:
    """Point-wise physics of either implementation (variant 'restated' or 'physref')."""

    def __init__(self, variant="restated"):
        self.L = load(variant)
        self.name = self.L.phys_impl_name().decode()

    def flux(self, flux_type, n, Wp, Wm, Ap=None, Am=None):
        Ap = Wp if Ap is None else Ap
        Am = Wm if Am is None else Am
        out = np.zeros(4)
        args = [np.ascontiguousarray(a, dtype=np.float64) for a in (n, Wp, Wm, Ap, Am)]
        self.L.phys_numerical_flux(int(flux_type), *[_d(a) for a in args], _d(out))
        return out

    def flux_matrix(self, W):
        out = np.zeros(8)
        self.L.phys_flux_matrix(_d(np.ascontiguousarray(W, dtype=np.float64)), _d(out))
        return out.reshape(4, 2)

    def wminus(self, kind, n, Wp, g):
        out = np.zeros(4)
        args = [np.ascontiguousarray(a, dtype=np.float64) for a in (n, Wp, g)]
        self.L.phys_wminus(int(kind), *[_d(a) for a in args], _d(out))
        return out

    def eigen(self, W):
        m = [np.zeros(16) for _ in range(4)]
        self.L.phys_eigen(_d(np.ascontiguousarray(W, dtype=np.float64)), *[_d(x) for x in m])
        return [x.reshape(4, 4) for x in m]  # Rx, Lx, Ry, Ly

    def to_char(self, Lm, W):
        w = np.array(W, dtype=np.float64)
        self.L.phys_to_char(_d(np.ascontiguousarray(Lm.reshape(-1))), _d(w))
        return w

    def to_con(self, Rm, W):
        w = np.array(W, dtype=np.float64)
        self.L.phys_to_con(_d(np.ascontiguousarray(Rm.reshape(-1))), _d(w))
        return w

    def pressure(self, W):
        return self.L.phys_pressure(_d(np.ascontiguousarray(W, dtype=np.float64)))

    def sound_speed(self, W):
        return self.L.phys_sound_speed(_d(np.ascontiguousarray(W, dtype=np.float64)))

    def max_eigenvalue(self, W):
        return self.L.phys_max_eigenvalue(_d(np.ascontiguousarray(W, dtype=np.float64)))


def make_params(basis="Qk", degree=1, flux="lxf", limiter="none", char_lim=False, pos_lim=False,
                conserve_angular_momentum=False, M=0.0, beta=1.0, gravity=0.0, cfl=0.9,
                bc=None, compat="src", n_threads=1, shock_indicator="limiter", mapping="cartesian", local_time_step=False):
    """bc: {boundary_id: kind} or {boundary_id: ("periodic", partner_id)}; default outflow
    (src/parameters.cc:384)."""
    p = OracleParams()
    p.basis = QK if basis == "Qk" else PK
    p.degree = degree
    p.flux_type = FLUX[flux]
    p.limiter_type = {"none": 0, "TVB": 1, "tvb": 1, "minmax": 2}[limiter]
    p.char_lim, p.pos_lim = int(char_lim), int(pos_lim)
    p.conserve_angular_momentum = int(conserve_angular_momentum)
    p.M, p.beta, p.gravity, p.cfl = M, beta, gravity, cfl
    for b in range(10):
        p.bc_kind[b] = BC["outflow"]
        p.periodic_pair[b] = -1
    for b, kind in (bc or {}).items():
        if isinstance(kind, tuple):
            p.bc_kind[b] = BC[kind[0]]
            p.periodic_pair[b] = kind[1]
        else:
            p.bc_kind[b] = BC[kind]
    p.compat = 0 if compat == "src" else 1
    p.n_threads = n_threads
    p.shock_indicator = {"limiter": 0, "density": 1, "energy": 2}[shock_indicator]
    p.mapping = {"cartesian": 0, "q1": 1}[mapping]
    p.local_time_step = int(local_time_step)
    return p


class Oracle:
    def __init__(self, vertices, cells, blines, bline_id, params, variant="restated"):
        self.L = load(variant)
        self.params = params
        v = np.ascontiguousarray(vertices, dtype=np.float64)
        c = np.ascontiguousarray(cells, dtype=np.int32)
        bl = np.ascontiguousarray(blines, dtype=np.int32).reshape(-1, 2)
        bi = np.ascontiguousarray(bline_id, dtype=np.int32)
        self.h = self.L.oracle_create(len(v), _d(v), len(c), _i(c), len(bl), _i(bl), _i(bi),
                                      ctypes.byref(params))
        if not self.h:
            raise RuntimeError("oracle_create: " + self.L.oracle_last_error().decode())
        self.h = ctypes.c_void_p(self.h)
        self.n_cells = self.L.oracle_n_cells(self.h)
        self.D = self.L.oracle_dofs_per_cell(self.h)
        self.nqf = self.L.oracle_n_q_face(self.h)
        self.nq = self.L.oracle_n_q_cell(self.h)
        self.n_bfaces = self.L.oracle_n_bfaces(self.h)
        self.n_rk = self.L.oracle_n_rk(self.h)
        self.ark = [self.L.oracle_ark(self.h, r) for r in range(self.n_rk)]

    def __del__(self):
        try:
            if self.h:
                self.L.oracle_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def neighbors(self):
        out = np.zeros((self.n_cells, 4), dtype=np.int32)
        self.L.oracle_get_neighbors(self.h, _i(out))
        return out

    def bfaces(self):
        n = max(self.n_bfaces, 1)
        cell = np.zeros(n, dtype=np.int32)
        face = np.zeros(n, dtype=np.int32)
        bid = np.zeros(n, dtype=np.int32)
        xq = np.zeros((n, self.nqf, 2))
        self.L.oracle_get_bfaces(self.h, _i(cell), _i(face), _i(bid), _d(xq))
        k = self.n_bfaces
        return cell[:k], face[:k], bid[:k], xq[:k]

    def cell_qpoints(self):
        xq = np.zeros((self.n_cells, self.nq, 2))
        self.L.oracle_get_cell_qpoints(self.h, _d(xq))
        return xq

    def tables(self):
        n1 = self.nqf
        gx, gw = np.zeros(n1), np.zeros(n1)
        self.L.oracle_get_tables(self.h, _d(gx), _d(gw))
        return gx, gw

    def set_initial_condition(self, f):
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert f.shape == (self.n_cells, self.nq, 4)
        self.L.oracle_set_initial_condition(self.h, _d(f))

    def set_solution(self, u):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        assert u.size == self.n_cells * self.D
        self.L.oracle_set_solution(self.h, _d(u))

    def solution(self):
        u = np.zeros(self.n_cells * self.D)
        self.L.oracle_get_solution(self.h, _d(u))
        return u

    def commit_step(self):
        self.L.oracle_commit_step(self.h)

    def set_bc_values(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert g.shape == (self.n_bfaces, self.nqf, 4)
        if self.n_bfaces:
            self.L.oracle_set_bc_values(self.h, _d(g))

    def set_external_force(self, f):
        """f: [n_cells][nq][2] values of the MPI tree's external force at the cell quadrature points."""
        f = np.ascontiguousarray(f, dtype=np.float64)
        assert f.shape == (self.n_cells, self.nq, 2)
        self.L.oracle_set_external_force(self.h, _d(f))

    def compute_cell_average(self):
        self.L.oracle_compute_cell_average(self.h)

    def cell_average(self):
        a = np.zeros((self.n_cells, 4))
        self.L.oracle_get_cell_average(self.h, _d(a))
        return a

    def assemble(self):
        self.L.oracle_assemble(self.h)
        r = np.zeros(self.n_cells * self.D)
        self.L.oracle_get_rhs(self.h, _d(r))
        return r

    def compute_dt(self, elapsed=0.0, final_time=1e20, time_step=-1.0):
        return self.L.oracle_compute_dt(self.h, ctypes.c_double(elapsed), ctypes.c_double(final_time),
                                        ctypes.c_double(time_step))

    def shock_indicator(self):
        s = np.zeros(self.n_cells)
        self.L.oracle_get_shock_indicator(self.h, _d(s))
        return s

    def apply_limiter(self):
        self.L.oracle_apply_limiter(self.h)

    def apply_positivity(self):
        return self.L.oracle_apply_positivity(self.h)

    def limited_flags(self):
        f = np.zeros(self.n_cells, dtype=np.int32)
        self.L.oracle_get_limited_flags(self.h, _i(f))
        return f

    def rk_stage(self, rk, dt):
        res = ctypes.c_double(0.0)
        err = self.L.oracle_rk_stage(self.h, int(rk), ctypes.c_double(dt), ctypes.byref(res))
        return err, res.value

    def run_steps(self, n):
        err = ctypes.c_int(0)
        t = self.L.oracle_run_steps(self.h, int(n), ctypes.byref(err))
        return t, err.value


def rect_mesh(nx, ny, x0, x1, y0, y1, ids=(4, 2, 1, 3)):
    """Structured nx x ny mesh of [x0,x1]x[y0,y1] in gmsh-like primitive form, cells numbered x
    fastest. ids = boundary ids of the (left, right, bottom, top) sides (default: the Physical
    Line numbering of examples/isentropic_vortex/grid.geo: bottom 1, right 2, top 3, left 4)."""
    xs = np.linspace(x0, x1, nx + 1)
    ys = np.linspace(y0, y1, ny + 1)
    X, Y = np.meshgrid(xs, ys, indexing="xy")
    verts = np.stack([X.ravel(), Y.ravel()], axis=1)
    vid = lambda i, j: j * (nx + 1) + i
    cells = np.array([[vid(i, j), vid(i + 1, j), vid(i, j + 1), vid(i + 1, j + 1)]
                      for j in range(ny) for i in range(nx)], dtype=np.int32)
    bl, bi = [], []
    for j in range(ny):
        bl.append((vid(0, j), vid(0, j + 1))); bi.append(ids[0])
        bl.append((vid(nx, j), vid(nx, j + 1))); bi.append(ids[1])
    for i in range(nx):
        bl.append((vid(i, 0), vid(i + 1, 0))); bi.append(ids[2])
        bl.append((vid(i, ny), vid(i + 1, ny))); bi.append(ids[3])
    return verts, cells, np.array(bl, dtype=np.int32), np.array(bi, dtype=np.int32)


def isentropic_vortex(x, y, compat="src", beta=5.0, x0=0.0, y0=0.0, mach=0.5):
    """src/ic.cc:44-61 (stationary, p = rho^gamma) or src_mpi/ic.cc:44-61 (advected, p = rho^gamma/gamma)."""
    g = 1.4
    a1 = 0.5 * beta / np.pi
    a2 = 0.5 * (g - 1.0) * a1 * a1  # src/ic.h:44-45, src_mpi/ic.h:48-49
    r2 = (x - x0) ** 2 + (y - y0) ** 2
    rho = (1.0 - a2 * np.exp(1.0 - r2)) ** (1.0 / (g - 1.0))
    vx = -a1 * (y - y0) * np.exp(0.5 * (1.0 - r2))
    vy = a1 * (x - x0) * np.exp(0.5 * (1.0 - r2))
    if compat == "src":
        pre = rho ** g
    else:
        vx = vx + mach
        pre = rho ** g / g
    return np.stack([rho * vx, rho * vy, rho, pre / (g - 1.0) + 0.5 * rho * (vx * vx + vy * vy)], axis=-1)
