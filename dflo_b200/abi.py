"""ctypes binding of the C ABI (include/dflo_b200.h, include/dflo_host.h).

This is plumbing for tests and bench.py: every call goes through the same extern "C" entry
points a dflo (C++/deal.II) host would bind.  There is no Python or CPU implementation of the
hot path here: if libdflo_b200.so is missing, or no CUDA device is present, calls fail loudly.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# DFLO_B200_LIB: developer override to compare differently tuned builds of the same library
LIB_PATH = os.environ.get("DFLO_B200_LIB") or os.path.join(_HERE, "csrc", "libdflo_b200.so")

MAX_BOUNDARIES = 10
FLUX = {"lxf": 0, "sw": 1, "kfvs": 2, "roe": 3, "hllc": 4, "kep": 5}
BC = {"inflow": 0, "outflow": 1, "slip": 2, "pressure": 3, "farfield": 4, "periodic": 5}
BASIS = {"Qk": 0, "Pk": 1}
LIMITER = {"none": 0, "TVB": 1, "minmax": 2}   # minmax: src_mpi/parameters.h:235
COMPAT = {"src": 0, "mpi": 1}
FACE_OWNER, FACE_PERIODIC, FACE_FLIP = 1, 2, 4

E_NO_DEVICE = -8
E_NEGATIVE_STATE = -4
E_POSLIM_ROOT = -5

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int32)
c_u8_p = ctypes.POINTER(ctypes.c_uint8)
c_u32_p = ctypes.POINTER(ctypes.c_uint32)


class FlatMesh(ctypes.Structure):
    _fields_ = [
        ("n_cells", ctypes.c_int32),
        ("cell_origin", c_double_p), ("cell_size", c_double_p),
        ("neighbor", c_int_p), ("face_flags", c_u8_p),
        ("n_boundary_faces", ctypes.c_int32),
        ("bface_cell", c_int_p), ("bface_face", c_int_p), ("bface_id", c_int_p),
        ("cell_vertices", c_double_p), ("neighbor_face", c_u8_p),
        ("n_hanging_faces", ctypes.c_int32), ("hanging", c_int_p),
    ]


INDICATOR = {"limiter": 0, "density": 1, "energy": 2}   # src/parameters.cc:229-237


class Params(ctypes.Structure):
    _fields_ = [
        ("basis", ctypes.c_int32), ("degree", ctypes.c_int32), ("flux_type", ctypes.c_int32),
        ("limiter_type", ctypes.c_int32), ("char_lim", ctypes.c_int32), ("pos_lim", ctypes.c_int32),
        ("conserve_angular_momentum", ctypes.c_int32), ("compat", ctypes.c_int32),
        ("M", ctypes.c_double), ("beta", ctypes.c_double), ("gravity", ctypes.c_double),
        ("cfl", ctypes.c_double), ("time_step", ctypes.c_double),
        ("bc_kind", ctypes.c_int32 * MAX_BOUNDARIES),
        ("shock_indicator", ctypes.c_int32), ("mapping", ctypes.c_int32),
        ("local_time_step", ctypes.c_int32), ("reserved1", ctypes.c_int32),
    ]


class DfloError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("dflo_b200 error %d: %s" % (code, msg))
        self.code = code


def make_params(basis="Qk", degree=1, flux="lxf", limiter="none", char_lim=False, pos_lim=False,
                conserve_angular_momentum=False, M=0.0, beta=1.0, gravity=0.0, cfl=0.9,
                time_step=-1.0, bc=None, compat="src", shock_indicator="limiter", mapping="cartesian", local_time_step=False):
    """bc: {boundary_id: kind} or {boundary_id: ("periodic", partner)}; default outflow
    (reference src/parameters.cc:384). Returns (Params, periodic_pair[10])."""
    p = Params()
    p.shock_indicator = INDICATOR[shock_indicator]
    p.mapping = {"cartesian": 0, "q1": 1}[mapping]
    p.local_time_step = int(local_time_step)
    p.basis, p.degree, p.flux_type = BASIS[basis], degree, FLUX[flux]
    p.limiter_type = LIMITER[limiter]
    p.char_lim, p.pos_lim = int(char_lim), int(pos_lim)
    p.conserve_angular_momentum = int(conserve_angular_momentum)
    p.compat = COMPAT[compat]
    p.M, p.beta, p.gravity, p.cfl, p.time_step = M, beta, gravity, cfl, time_step
    pair = (ctypes.c_int32 * MAX_BOUNDARIES)(*([-1] * MAX_BOUNDARIES))
    for b in range(MAX_BOUNDARIES):
        p.bc_kind[b] = BC["outflow"]
    for b, kind in (bc or {}).items():
        if isinstance(kind, tuple):
            p.bc_kind[b] = BC[kind[0]]
            pair[b] = kind[1]
        else:
            p.bc_kind[b] = BC[kind]
    return p, pair


_lib = None


def load_library(path=None):
    """Load libdflo_b200.so (built in-tree by __graft_entry__.build())."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError("dflo_b200: %s not found -- run `python -c 'import __graft_entry__ as g; g.build()'`; "
                           "there is no fallback implementation" % path)
    lib = ctypes.CDLL(path, mode=ctypes.RTLD_GLOBAL)
    _declare_host(lib)
    _declare_engine(lib, "dflo_b200_")
    if path == LIB_PATH:
        _lib = lib
    return lib


def _declare_host(L):
    L.dflo_host_last_error.restype = ctypes.c_char_p
    L.dflo_mesh_create.restype = ctypes.c_void_p
    L.dflo_mesh_create.argtypes = [ctypes.c_char_p, c_double_p, ctypes.c_int]
    L.dflo_mesh_read_gmsh.restype = ctypes.c_void_p
    L.dflo_mesh_read_gmsh.argtypes = [ctypes.c_char_p]
    L.dflo_mesh_write_gmsh.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.dflo_mesh_destroy.argtypes = [ctypes.c_void_p]
    for n in ("n_vertices", "n_cells", "n_blines"):
        getattr(L, "dflo_mesh_" + n).argtypes = [ctypes.c_void_p]
    L.dflo_mesh_vertices.restype = c_double_p
    L.dflo_mesh_vertices.argtypes = [ctypes.c_void_p]
    for n in ("cells", "blines", "bline_ids"):
        f = getattr(L, "dflo_mesh_" + n)
        f.restype = c_int_p
        f.argtypes = [ctypes.c_void_p]
    L.dflo_mesh_flatten.argtypes = [ctypes.c_void_p, c_int_p, c_int_p]
    L.dflo_mesh_flat.restype = ctypes.POINTER(FlatMesh)
    L.dflo_mesh_flat.argtypes = [ctypes.c_void_p]
    L.dflo_expr_eval.argtypes = [ctypes.c_char_p, ctypes.c_int, c_double_p, c_double_p, ctypes.c_double, c_double_p]
    L.dflo_host_write_solution_vtu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_size_t,
                                               ctypes.c_int, ctypes.c_double, ctypes.c_uint, ctypes.c_char_p]
    L.dflo_host_write_solution_piece_vtu.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_size_t,
                                                     ctypes.c_int, ctypes.c_double, ctypes.c_uint, ctypes.c_int, ctypes.c_int,
                                                     ctypes.c_int, ctypes.c_char_p]
    L.dflo_host_write_solution_tecplot.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_size_t,
                                                   ctypes.c_int, ctypes.c_double, ctypes.c_char_p]
    L.dflo_host_angular_momentum.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, c_double_p, ctypes.c_size_t, c_double_p]
    L.dflo_host_write_shock_vtu.argtypes = [ctypes.c_void_p, c_double_p, c_double_p, ctypes.c_char_p]


def _declare_engine(L, prefix):
    def f(name):
        return getattr(L, prefix + name)
    vp = ctypes.c_void_p
    f("strerror").restype = ctypes.c_char_p
    f("strerror").argtypes = [ctypes.c_int]
    f("last_error").restype = ctypes.c_char_p
    f("last_error").argtypes = [vp]
    f("create").argtypes = [ctypes.POINTER(FlatMesh), ctypes.POINTER(Params), ctypes.c_int, ctypes.POINTER(vp)]
    f("create_sharded").argtypes = [ctypes.POINTER(FlatMesh), ctypes.POINTER(Params), ctypes.c_int, ctypes.c_int,
                                    ctypes.c_int, vp, ctypes.POINTER(vp)]
    f("destroy").argtypes = [vp]
    f("destroy").restype = None
    for n in ("dofs_per_cell", "n_q_face", "n_rk"):
        f(n).argtypes = [vp]
    f("ark").restype = ctypes.c_double
    f("ark").argtypes = [vp, ctypes.c_int]
    f("n_cells_owned").restype = ctypes.c_int64
    f("n_cells_owned").argtypes = [vp]
    f("cell_range").restype = ctypes.c_int64
    f("cell_range").argtypes = [vp, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]
    f("set_solution").argtypes = [vp, c_double_p, c_u32_p, ctypes.c_size_t]
    f("get_solution").argtypes = [vp, c_double_p, c_u32_p, ctypes.c_size_t]
    f("get_rhs").argtypes = [vp, c_double_p, c_u32_p, ctypes.c_size_t]
    f("get_cell_average").argtypes = [vp, c_double_p]
    f("commit_step").argtypes = [vp]
    f("set_boundary_values").argtypes = [vp, c_double_p]
    f("set_boundary_expression").argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]
    f("set_external_force").argtypes = [vp, ctypes.c_char_p, ctypes.c_char_p]
    f("assemble_rhs").argtypes = [vp, ctypes.c_double]
    f("rk_stage").argtypes = [vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, c_double_p]
    f("compute_dt").argtypes = [vp, ctypes.c_double, ctypes.c_double, c_double_p]
    f("limit_initial_condition").argtypes = [vp]
    f("advance").argtypes = [vp, ctypes.c_int, ctypes.c_double, c_double_p, c_double_p]
    f("poll_error").argtypes = [vp]
    f("get_limited_flags").argtypes = [vp, c_int_p]
    f("get_shock_indicator").argtypes = [vp, c_double_p]
    f("launch_count").restype = ctypes.c_int64
    f("launch_count").argtypes = [vp]
    f("stream").restype = vp
    f("stream").argtypes = [vp]
    f("synchronize").argtypes = [vp]
    f("last_advance_ms").argtypes = [vp, ctypes.POINTER(ctypes.c_float)]
    f("time_stage_kernel").argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(ctypes.c_float)]


def _dp(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_double_p)


class Mesh:
    """Host mesh (include/dflo_host.h dflo_mesh_*)."""

    def __init__(self, kind=None, args=(), gmsh_path=None, lib=None, handle=None, owned=True):
        self.L = lib or load_library()
        self.owned = owned
        if handle is not None:
            self.h = ctypes.c_void_p(handle)
        elif gmsh_path is not None:
            self.h = ctypes.c_void_p(self.L.dflo_mesh_read_gmsh(gmsh_path.encode()))
        else:
            a = np.ascontiguousarray(args, dtype=np.float64)
            self.h = ctypes.c_void_p(self.L.dflo_mesh_create(kind.encode(), _dp(a), len(a)))
        if not self.h:
            raise DfloError(-1, self.L.dflo_host_last_error().decode())
        self.flat = None

    def __del__(self):
        try:
            if self.owned and self.h:
                self.L.dflo_mesh_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @property
    def n_cells(self):
        return self.L.dflo_mesh_n_cells(self.h)

    def primitive(self):
        """(vertices[nv,2], cells[nc,4], blines[nb,2], bline_id[nb]) as numpy copies."""
        nv, nc, nb = (self.L.dflo_mesh_n_vertices(self.h), self.L.dflo_mesh_n_cells(self.h),
                      self.L.dflo_mesh_n_blines(self.h))
        v = np.ctypeslib.as_array(self.L.dflo_mesh_vertices(self.h), shape=(nv, 2)).copy()
        c = np.ctypeslib.as_array(self.L.dflo_mesh_cells(self.h), shape=(nc, 4)).copy()
        if nb:
            bl = np.ctypeslib.as_array(self.L.dflo_mesh_blines(self.h), shape=(nb, 2)).copy()
            bi = np.ctypeslib.as_array(self.L.dflo_mesh_bline_ids(self.h), shape=(nb,)).copy()
        else:
            bl, bi = np.zeros((0, 2), np.int32), np.zeros((0,), np.int32)
        return v, c, bl, bi

    def write_gmsh(self, path):
        rc = self.L.dflo_mesh_write_gmsh(self.h, path.encode())
        if rc:
            raise DfloError(rc, "cannot write " + path)

    def flatten(self, params, periodic_pair):
        rc = self.L.dflo_mesh_flatten(self.h, ctypes.cast(params.bc_kind, c_int_p), ctypes.cast(periodic_pair, c_int_p))
        if rc:
            raise DfloError(rc, self.L.dflo_host_last_error().decode())
        self.flat = self.L.dflo_mesh_flat(self.h)
        return self.flat

    def write_solution_vtu(self, path, u, basis, degree, schlieren_plot=False, time=0.0, cycle=0, cells=None, subdomain=-1):
        """output_results on a host copy of the solution (src/output.cc:33-68); basis: "Qk" | "Pk"; cells = (begin, end)
        writes one process' piece like src_mpi/output.cc (u stays the global vector)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        c0, c1 = (0, -1) if cells is None else cells
        rc = self.L.dflo_host_write_solution_piece_vtu(self.h, {"Qk": 0, "Pk": 1}[basis], degree, _dp(u), u.size,
                                                       int(schlieren_plot), time, cycle, c0, c1, subdomain, path.encode())
        if rc:
            raise DfloError(rc, self.L.dflo_host_last_error().decode())

    def write_solution_tecplot(self, path, u, basis, degree, schlieren_plot=False, time=0.0):
        """output_results with "output: format = tecplot" (src/output.cc:51-52, 65-66)."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        rc = self.L.dflo_host_write_solution_tecplot(self.h, {"Qk": 0, "Pk": 1}[basis], degree, _dp(u), u.size,
                                                     int(schlieren_plot), time, path.encode())
        if rc:
            raise DfloError(rc, self.L.dflo_host_last_error().decode())

    def angular_momentum(self, u, basis, degree):
        """compute_angular_momentum (src/claw.cc:604-635) of a host copy of the solution."""
        u = np.ascontiguousarray(u, dtype=np.float64)
        v = ctypes.c_double(0.0)
        rc = self.L.dflo_host_angular_momentum(self.h, {"Qk": 0, "Pk": 1}[basis], degree, _dp(u), u.size,
                                               ctypes.cast(ctypes.byref(v), c_double_p))
        if rc:
            raise DfloError(rc, self.L.dflo_host_last_error().decode())
        return v.value

    def write_shock_vtu(self, path, shock_indicator, mu_shock=None):
        """shock.vtu of src/output.cc:70-79."""
        s = np.ascontiguousarray(shock_indicator, dtype=np.float64)
        mu = None if mu_shock is None else np.ascontiguousarray(mu_shock, dtype=np.float64)
        rc = self.L.dflo_host_write_shock_vtu(self.h, None if mu is None else _dp(mu), _dp(s), path.encode())
        if rc:
            raise DfloError(rc, self.L.dflo_host_last_error().decode())

    def flat_arrays(self):
        fm = self.flat.contents
        n, nb = fm.n_cells, fm.n_boundary_faces
        out = dict(
            origin=np.ctypeslib.as_array(fm.cell_origin, shape=(n, 2)).copy(),
            size=np.ctypeslib.as_array(fm.cell_size, shape=(n, 2)).copy(),
            neighbor=np.ctypeslib.as_array(fm.neighbor, shape=(n, 4)).copy(),
            face_flags=np.ctypeslib.as_array(fm.face_flags, shape=(n, 4)).copy(),
        )
        if nb:
            out.update(bface_cell=np.ctypeslib.as_array(fm.bface_cell, shape=(nb,)).copy(),
                       bface_face=np.ctypeslib.as_array(fm.bface_face, shape=(nb,)).copy(),
                       bface_id=np.ctypeslib.as_array(fm.bface_id, shape=(nb,)).copy())
        else:
            z = np.zeros((0,), np.int32)
            out.update(bface_cell=z, bface_face=z, bface_id=z)
        return out


def expr_eval(expr, x, y, t=0.0, lib=None):
    L = lib or load_library()
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
    y = np.ascontiguousarray(y, dtype=np.float64).reshape(-1)
    out = np.zeros_like(x)
    rc = L.dflo_expr_eval(expr.encode(), len(x), _dp(x), _dp(y), float(t), _dp(out))
    if rc:
        raise DfloError(rc, L.dflo_host_last_error().decode())
    return out


class Engine:
    """The RK-stage engine behind include/dflo_b200.h.  `prefix`/`lib` exist so that the CPU test
    tier can drive the emulation build of the same ABI (tests/emu); the product always uses the
    defaults."""

    def __init__(self, flat_mesh, params, device=0, rank=0, world=1, nccl_id=None, lib=None, prefix="dflo_b200_"):
        self.L = lib or load_library()
        self.prefix = prefix
        self._keep = (flat_mesh, params)
        h = ctypes.c_void_p()
        if world == 1:
            rc = self._f("create")(flat_mesh, ctypes.byref(params), device, ctypes.byref(h))
        else:
            idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
            rc = self._f("create_sharded")(flat_mesh, ctypes.byref(params), device, rank, world, idbuf, ctypes.byref(h))
        if rc:
            raise DfloError(rc, self._f("strerror")(rc).decode() + ": " + self._f("last_error")(None).decode())
        self.h = h
        self.D = self._f("dofs_per_cell")(h)
        self.nqf = self._f("n_q_face")(h)
        self.n_rk = self._f("n_rk")(h)
        self.ark = [self._f("ark")(h, r) for r in range(self.n_rk)]
        self.n_cells = flat_mesh.contents.n_cells
        self.n_bfaces = flat_mesh.contents.n_boundary_faces
        self.rank, self.world = rank, world

    def _f(self, name):
        return getattr(self.L, self.prefix + name)

    def _check(self, rc):
        if rc:
            raise DfloError(rc, self._f("strerror")(rc).decode() + ": " + self._f("last_error")(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self._f("destroy")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def cell_range(self):
        b, e = ctypes.c_int64(), ctypes.c_int64()
        self._f("cell_range")(self.h, ctypes.byref(b), ctypes.byref(e))
        return b.value, e.value

    def set_solution(self, u, dof_map=None):
        u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1)
        dm = None if dof_map is None else np.ascontiguousarray(dof_map, dtype=np.uint32).ctypes.data_as(c_u32_p)
        self._check(self._f("set_solution")(self.h, _dp(u), dm, u.size))

    def get_solution(self, out=None, dof_map=None):
        u = np.zeros(self.n_cells * self.D) if out is None else out
        dm = None if dof_map is None else np.ascontiguousarray(dof_map, dtype=np.uint32).ctypes.data_as(c_u32_p)
        self._check(self._f("get_solution")(self.h, _dp(u), dm, u.size))
        return u

    def get_rhs(self):
        r = np.zeros(self.n_cells * self.D)
        self._check(self._f("get_rhs")(self.h, _dp(r), None, r.size))
        return r

    def cell_average(self):
        a = np.zeros((self.n_cells, 4))
        self._check(self._f("get_cell_average")(self.h, _dp(a)))
        return a

    def commit_step(self):
        self._check(self._f("commit_step")(self.h))

    def set_boundary_values(self, g):
        g = np.ascontiguousarray(g, dtype=np.float64)
        assert g.size == self.n_bfaces * self.nqf * 4
        if g.size:
            self._check(self._f("set_boundary_values")(self.h, _dp(g)))

    def set_boundary_expression(self, boundary_id, comp, expr):
        self._check(self._f("set_boundary_expression")(self.h, boundary_id, comp, expr.encode()))

    def set_external_force(self, fx_expr, fy_expr):
        """"f_0 value" / "f_1 value" of the MPI tree (src_mpi/parameters.cc:488-497)."""
        self._check(self._f("set_external_force")(self.h, fx_expr.encode(), fy_expr.encode()))

    def assemble_rhs(self, t_bc=0.0):
        self._check(self._f("assemble_rhs")(self.h, float(t_bc)))

    def rk_stage(self, rk, t_bc, dt, want_norm=False):
        res = ctypes.c_double(0.0)
        self._check(self._f("rk_stage")(self.h, rk, float(t_bc), float(dt), ctypes.byref(res) if want_norm else None))
        return res.value

    def compute_dt(self, elapsed=0.0, final_time=1e20):
        dt = ctypes.c_double(0.0)
        self._check(self._f("compute_dt")(self.h, float(elapsed), float(final_time), ctypes.byref(dt)))
        return dt.value

    def limit_initial_condition(self):
        self._check(self._f("limit_initial_condition")(self.h))

    def advance(self, n_steps, elapsed=0.0, final_time=1e20):
        t, dt = ctypes.c_double(elapsed), ctypes.c_double(0.0)
        self._check(self._f("advance")(self.h, int(n_steps), float(final_time), ctypes.byref(t), ctypes.byref(dt)))
        return t.value, dt.value

    def poll_error(self):
        self._check(self._f("poll_error")(self.h))

    def limited_flags(self):
        f = np.zeros(self.n_cells, dtype=np.int32)
        self._check(self._f("get_limited_flags")(self.h, f.ctypes.data_as(c_int_p)))
        return f

    def shock_indicator(self):
        s = np.zeros(self.n_cells)
        self._check(self._f("get_shock_indicator")(self.h, _dp(s)))
        return s

    def launch_count(self):
        return self._f("launch_count")(self.h)

    def synchronize(self):
        self._check(self._f("synchronize")(self.h))

    def time_stage_kernel(self, rk=1, reps=20, flush_bytes=0):
        ms = ctypes.c_float(0.0)
        self._check(self._f("time_stage_kernel")(self.h, rk, reps, flush_bytes, ctypes.byref(ms)))
        return ms.value

    def last_advance_ms(self):
        ms = ctypes.c_float(0.0)
        self._check(self._f("last_advance_ms")(self.h, ctypes.byref(ms)))
        return ms.value
