"""CPU tier: host front end (input.prm parser, expression VM, mesh generators / gmsh reader,
flattening, partitioner) and the C-ABI surface of libdflo_b200.so.  No GPU compute calls."""
import ctypes
import glob
import os
import re

import numpy as np
import pytest

from dflo_b200 import abi
from helpers import PERIODIC_BOX, SOD_BC, STEP_BC, emu_lib
from oracle import oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PRM_DIR = os.path.join(ROOT, "tests", "golden", "prm")


@pytest.fixture(scope="module")
def lib():
    return abi.load_library()


def _claw_api(L):
    L.dflo_claw_create.restype = ctypes.c_void_p
    L.dflo_claw_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    L.dflo_claw_destroy.argtypes = [ctypes.c_void_p]
    L.dflo_claw_params.restype = ctypes.POINTER(abi.Params)
    L.dflo_claw_params.argtypes = [ctypes.c_void_p]
    L.dflo_claw_periodic_pairs.restype = ctypes.POINTER(ctypes.c_int)
    L.dflo_claw_periodic_pairs.argtypes = [ctypes.c_void_p]
    L.dflo_claw_n_dofs.argtypes = [ctypes.c_void_p]
    L.dflo_claw_final_time.restype = ctypes.c_double
    L.dflo_claw_final_time.argtypes = [ctypes.c_void_p]
    L.dflo_claw_boundary_expression.restype = ctypes.c_char_p
    L.dflo_claw_boundary_expression.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
    L.dflo_claw_initial_condition.argtypes = [ctypes.c_void_p, abi.c_double_p, ctypes.c_size_t]
    L.dflo_claw_setup.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.dflo_claw_mesh.restype = ctypes.c_void_p
    L.dflo_claw_mesh.argtypes = [ctypes.c_void_p]
    return L


# ---------------------------------------------------------------------------------------------
# C-ABI surface
# ---------------------------------------------------------------------------------------------
def _declared_symbols(header):
    text = open(os.path.join(ROOT, "include", header)).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dflo_(?:b200|mesh|expr|claw|host)_[a-z0-9_]+)\s*\(", text)))


@pytest.mark.parametrize("header", ["dflo_b200.h", "dflo_host.h"])
def test_library_exports_every_declared_symbol(lib, header):
    names = _declared_symbols(header)
    assert len(names) >= (25 if header == "dflo_b200.h" else 20)
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_abi_version_and_strerror(lib):
    lib.dflo_b200_abi_version.restype = ctypes.c_int
    assert lib.dflo_b200_abi_version() == 4   # v4: cell_vertices / neighbor_face in the flat mesh, mapping + local_time_step in the parameters
    assert lib.dflo_b200_strerror(0) == b"ok"
    assert b"Negative states" in lib.dflo_b200_strerror(abi.E_NEGATIVE_STATE)       # positivity.cc:33-37
    assert b"positivity limiter" in lib.dflo_b200_strerror(abi.E_POSLIM_ROOT)       # positivity.cc:160-169


def test_params_struct_layout_matches_header():
    """ctypes mirror == the C struct of include/dflo_b200.h (checked against a compiled sizeof)."""
    import subprocess
    import tempfile
    src = '#include "dflo_b200.h"\n#include <stdio.h>\n#include <stddef.h>\nint main(){printf("%zu %zu %zu %zu\\n",' \
          'sizeof(dflo_params),offsetof(dflo_params,M),offsetof(dflo_params,bc_kind),sizeof(dflo_flat_mesh));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")], check=True)
        out = subprocess.run([os.path.join(d, "t")], capture_output=True, text=True, check=True).stdout.split()
    assert [int(x) for x in out] == [ctypes.sizeof(abi.Params), abi.Params.M.offset, abi.Params.bc_kind.offset,
                                     ctypes.sizeof(abi.FlatMesh)]


def test_no_cpu_fallback_without_a_device(lib):
    """On a box with no CUDA device the product refuses to run (DFLO_E_NO_DEVICE); with a device
    this test only checks that creation works and the kernels are counted."""
    import torch
    params, pair = abi.make_params(bc=PERIODIC_BOX, basis="Qk", degree=1)
    mesh = abi.Mesh("isentropic_vortex", [4], lib=lib)
    flat = mesh.flatten(params, pair)
    if torch.cuda.is_available():
        e = abi.Engine(flat, params)
        e.close()
    else:
        with pytest.raises(abi.DfloError) as ei:
            abi.Engine(flat, params)
        assert ei.value.code == abi.E_NO_DEVICE
        assert "no CPU fallback" in str(ei.value)


def test_product_never_touches_the_oracle():
    """oracle/ is test infrastructure: nothing under dflo_b200/ may import, include or link it."""
    bad = []
    for path in glob.glob(os.path.join(ROOT, "dflo_b200", "**", "*"), recursive=True):
        if os.path.isdir(path) or path.endswith((".so", ".o")) or "/_obj/" in path or os.path.basename(path) == "dflo_b200":
            continue
        try:
            text = open(path, errors="ignore").read()
        except Exception:
            continue
        if re.search(r"(from|import)\s+oracle|oracle/|liboracle|phys_restated|dflo_oracle", text):
            bad.append(path)
    assert not bad, bad
    import subprocess
    needed = subprocess.run(["readelf", "-d", abi.LIB_PATH], capture_output=True, text=True).stdout
    # cudart is linked statically; NCCL is bound with dlopen only when a sharded ctx is created
    assert "oracle" not in needed and "nccl" not in needed


# ---------------------------------------------------------------------------------------------
# input.prm
# ---------------------------------------------------------------------------------------------
EXPECT = {
    "cfg1": dict(basis=0, degree=1, flux=abi.FLUX["lxf"], limiter=0, pos=0, cfl=0.9, n_periodic=4),
    "cfg2": dict(basis=0, degree=3, flux=abi.FLUX["roe"], limiter=0, pos=0, cfl=0.9, n_periodic=4),
    "cfg3": dict(basis=1, degree=2, flux=abi.FLUX["hllc"], limiter=1, pos=1, cfl=0.9, n_periodic=0, M=0.0, beta=2.0),
    "cfg4": dict(basis=0, degree=2, flux=abi.FLUX["hllc"], limiter=1, pos=0, cfl=0.9, n_periodic=0, M=100.0, beta=1.0),
    "cfg5": dict(basis=0, degree=3, flux=abi.FLUX["kfvs"], limiter=1, pos=1, cfl=0.5, n_periodic=0, M=0.0, beta=2.0),
}
MESH_FOR = {"cfg1": "isentropic_vortex 8", "cfg2": "isentropic_vortex 8", "cfg3": "sod_tube 20 2", "cfg4": "double_mach 4",
            "cfg5": "forward_step 0.2"}


@pytest.mark.parametrize("prm", sorted(f for f in os.listdir(PRM_DIR) if f.endswith(".prm")))
def test_prm_baseline_fixtures(lib, prm):
    L = _claw_api(lib)
    key = prm[:4]
    h = L.dflo_claw_create(os.path.join(PRM_DIR, prm).encode(), MESH_FOR[key].encode(), None, abi.COMPAT["mpi"])
    assert h, L.dflo_host_last_error()
    p = L.dflo_claw_params(h).contents
    e = EXPECT[key]
    assert (p.basis, p.degree, p.flux_type, p.limiter_type, p.pos_lim) == (e["basis"], e["degree"], e["flux"], e["limiter"], e["pos"])
    assert p.cfl == e["cfl"]
    if "M" in e:
        assert (p.M, p.beta) == (e["M"], e["beta"]) and p.char_lim == 1
    pairs = [L.dflo_claw_periodic_pairs(h)[i] for i in range(10)]
    assert sum(1 for x in pairs if x >= 0) == e["n_periodic"]
    if key == "cfg4":
        assert p.bc_kind[3] == abi.BC["inflow"] and p.bc_kind[1] == abi.BC["slip"]
        ex = L.dflo_claw_boundary_expression(h, 3, 2).decode()
        x = np.array([0.1, 1.0, 2.0])
        got = abi.expr_eval(ex, x, 0 * x + 1.0, t=0.01, lib=lib)
        xs = 1.0 / 6.0 + (1 + 20 * 0.01) / np.sqrt(3.0)
        assert np.array_equal(got, np.where(x < xs, 8.0, 1.4))
    # host-side initial condition in the reference DoF layout
    n = L.dflo_claw_n_dofs(h)
    u = np.zeros(n)
    assert L.dflo_claw_initial_condition(h, u.ctypes.data_as(abi.c_double_p), n) == 0
    assert np.all(np.isfinite(u))
    D = 4 * ((e["degree"] + 1) ** 2 if e["basis"] == 0 else (e["degree"] + 1) * (e["degree"] + 2) // 2)
    rho = u.reshape(-1, 4, D // 4)[:, 2, :]
    assert rho[:, 0].min() > 0.0          # Qk: nodal densities; Pk: mode 0 = mean density
    L.dflo_claw_destroy(h)


def test_prm_rejects_undeclared_keys_and_bad_combinations(lib, tmp_path):
    L = _claw_api(lib)
    base = open(os.path.join(PRM_DIR, "cfg3_sod_P2_hllc_tvb_pos.prm")).read()
    cases = {
        "undeclared": base + "\nset no such key = 1\n",                                     # ParameterHandler: error
        "tvb_needs_cartesian": base.replace("set mapping   = cartesian", "set mapping   = q1"),   # parameters.cc:536-541
        "bad_flux": base.replace("set flux = hllc", "set flux = hllx"),
        "minmax_needs_qk": base.replace("set type = TVB", "set type = minmax"),           # src_mpi/parameters.cc:610-611
    }
    for name, text in cases.items():
        p = tmp_path / (name + ".prm")
        p.write_text(text)
        h = L.dflo_claw_create(str(p).encode(), b"sod_tube 10 2", None, 0)
        assert not h, name
        assert L.dflo_host_last_error()
    # limiter type = minmax (src_mpi/parameters.cc:205) is accepted on Qk
    p = tmp_path / "minmax_qk.prm"
    p.write_text(base.replace("set type = TVB", "set type = minmax").replace("set basis     = Pk", "set basis     = Qk"))
    h = L.dflo_claw_create(str(p).encode(), b"sod_tube 10 2", None, 0)
    assert h, L.dflo_host_last_error()
    L.dflo_claw_destroy(h)
    # the positivity limiter alone is admitted on mapped cells (parameters.cc:536-550 refuses only TVB and Pk there)
    p = tmp_path / "q1_positivity.prm"
    p.write_text(base.replace("set mapping   = cartesian", "set mapping   = q1").replace("set type = TVB", "set type = none")
                 .replace("set basis     = Pk", "set basis     = Qk"))
    h = L.dflo_claw_create(str(p).encode(), b"sod_tube 10 2", None, 0)
    assert h, L.dflo_host_last_error()
    L.dflo_claw_destroy(h)
    # src/ does not know periodic boundaries (SURVEY 8a forks): compat=src rejects them, compat=mpi accepts
    vort = os.path.join(PRM_DIR, "cfg1_isentropic_vortex_Q1_lxf.prm").encode()
    h = L.dflo_claw_create(vort, b"isentropic_vortex 4", None, abi.COMPAT["mpi"])
    assert h
    L.dflo_claw_destroy(h)


def test_prm_parses_the_reference_examples(lib):
    """Every explicit-path example deck the reference ships parses (when /root/reference is here)."""
    L = _claw_api(lib)
    decks = {"isentropic_vortex": "isentropic_vortex 4", "sod_shock_tube": "sod_tube 10 2",
             "double_mach_reflection": "double_mach 4", "forward_step": "forward_step 0.2"}
    if not os.path.isdir("/root/reference/examples"):
        pytest.skip("reference tree not present on this box")
    for ex, mesh in decks.items():
        h = L.dflo_claw_create(("/root/reference/examples/%s/input.prm" % ex).encode(), mesh.encode(), None, abi.COMPAT["mpi"])
        assert h, (ex, L.dflo_host_last_error())
        L.dflo_claw_destroy(h)


# ---------------------------------------------------------------------------------------------
# expressions (deal.II FunctionParser / muparser subset)
# ---------------------------------------------------------------------------------------------
def test_expression_vm(lib):
    x = np.linspace(-2.0, 3.0, 41)
    y = np.linspace(0.5, 1.5, 41)
    t = 0.3
    pi = np.pi
    cases = {
        "1.0 + 2*x - y/4": 1.0 + 2 * x - y / 4,
        "x^2 + y^3": x ** 2 + y ** 3,
        "-x^2": -(x ** 2),
        "2^3^2": 512.0 + 0 * x,
        "sin(_pi*x)*cos(y) + exp(-t)": np.sin(pi * x) * np.cos(y) + np.exp(-t),
        "sqrt(abs(x)) * (x >= 0) + (x < 0)*3": np.sqrt(np.abs(x)) * (x >= 0) + (x < 0) * 3,
        "(x<=0.5) + 0.125*(x>0.5)": (x <= 0.5) + 0.125 * (x > 0.5),
        "57.1576766498*(x<1.0/6.0+(1+20*t)/sqrt(3)) + 0.0": 57.1576766498 * (x < 1.0 / 6.0 + (1 + 20 * t) / np.sqrt(3.0)),
        "1e-3*x + 2.5E2": 1e-3 * x + 250.0,
        "(x > 0) && (y < 1) || (x == -2)": ((x > 0) & (y < 1)) | (x == -2),
        "log(y) + tan(y/2)": np.log(y) + np.tan(y / 2),
    }
    for ex, want in cases.items():
        got = abi.expr_eval(ex, x, y, t, lib=lib)
        assert np.allclose(got, np.asarray(want, dtype=float), rtol=1e-15, atol=1e-15), ex
    for bad in ["1 +", "foo(x)", "(x", "x y", ""]:
        with pytest.raises(abi.DfloError):
            abi.expr_eval(bad, x, y, t, lib=lib)


# ---------------------------------------------------------------------------------------------
# meshes, flattening, partitioning
# ---------------------------------------------------------------------------------------------
MESHES = [("isentropic_vortex", [6], PERIODIC_BOX), ("sod_tube", [12, 3], SOD_BC), ("double_mach", [4], {1: "slip", 3: "inflow", 4: "inflow"}),
          ("forward_step", [0.2], STEP_BC), ("rectangle", [5, 3, 0.0, 2.0, -1.0, 1.0, 4, 2, 1, 3], PERIODIC_BOX)]


@pytest.mark.parametrize("kind,args,bc", MESHES, ids=[m[0] for m in MESHES])
def test_flatten_matches_oracle_topology(lib, kind, args, bc):
    """the product's flattening (neighbour lists, boundary faces, face ownership, periodic partners)
    against the oracle's independent derivation from the same primitive mesh"""
    params, pair = abi.make_params(bc=bc)
    m = abi.Mesh(kind, args, lib=lib)
    m.flatten(params, pair)
    fa = m.flat_arrays()
    v, c, bl, bi = m.primitive()
    o = O.Oracle(v, c, bl, bi, O.make_params(bc=bc))
    assert np.array_equal(fa["neighbor"], o.neighbors())
    oc, of, oid, _ = o.bfaces()
    assert np.array_equal(fa["bface_cell"], oc) and np.array_equal(fa["bface_face"], of) and np.array_equal(fa["bface_id"], oid)
    nb, fl = fa["neighbor"], fa["face_flags"]
    n = len(nb)
    for cell in range(n):
        for f in range(4):
            j = nb[cell, f]
            if j < 0:
                assert fl[cell, f] == 0
                continue
            assert nb[j, f ^ 1] == cell                        # symmetric
            if fl[cell, f] & abi.FACE_PERIODIC:
                assert fl[j, f ^ 1] & abi.FACE_PERIODIC        # both sides integrate (src_mpi 186-260)
            else:                                              # MeshWorker: exactly one side owns the face
                assert bool(fl[cell, f] & abi.FACE_OWNER) != bool(fl[j, f ^ 1] & abi.FACE_OWNER)
                assert bool(fl[cell, f] & abi.FACE_OWNER) == (cell < j)
    assert np.all(fa["size"] > 0)


def test_generators_reproduce_the_geo_files(lib):
    """cell counts / extents / boundary ids of the transfinite .geo files"""
    m = abi.Mesh("isentropic_vortex", [32], lib=lib)
    v, c, bl, bi = m.primitive()
    assert len(c) == 1024 and v.min() == -5.0 and v.max() == 5.0 and sorted(set(bi)) == [1, 2, 3, 4]
    m = abi.Mesh("sod_tube", [100, 10], lib=lib)
    v, c, bl, bi = m.primitive()
    assert len(c) == 1000 and v[:, 0].max() == 1.0 and abs(v[:, 1].max() - 0.1) < 1e-15 and sorted(set(bi)) == [0, 1, 2]
    m = abi.Mesh("double_mach", [16], lib=lib)
    v, c, bl, bi = m.primitive()
    # grid.geo: n1 = ceil(x0/dy) = 3, n2 = ceil((4-x0)/dy) = 62 columns of squares around x0 = 1/6
    assert len(c) == 65 * 16 and abs(v[:, 0].max() - (1 / 6 + 62 / 16)) < 1e-14 and v[:, 1].max() == 1.0
    assert abs(v[:, 0].min() - (1 / 6 - 3 / 16)) < 1e-14 and sorted(set(bi)) == [0, 1, 2, 3, 4]
    m = abi.Mesh("forward_step", [0.1], lib=lib)
    v, c, bl, bi = m.primitive()
    # 3 blocks: [0,0.6]x[0,0.2], [0,0.6]x[0.2,1], [0.6,3]x[0.2,1]
    assert len(c) == 6 * 2 + 6 * 8 + 24 * 8
    assert v[:, 0].max() == 3.0 and v[:, 1].max() == 1.0 and sorted(set(bi)) == [1, 2, 3]


def test_gmsh_v2_round_trip(lib, tmp_path):
    m = abi.Mesh("forward_step", [0.2], lib=lib)
    path = str(tmp_path / "step.msh")
    m.write_gmsh(path)
    text = open(path).read()
    assert text.startswith("$MeshFormat\n2.") and "$Elements" in text
    r = abi.Mesh(gmsh_path=path, lib=lib)
    a, b = m.primitive(), r.primitive()
    assert np.array_equal(a[0], b[0])
    # same cells up to the reader's CCW -> lexicographic conversion; same boundary ids
    assert np.array_equal(np.sort(a[1], axis=1), np.sort(b[1], axis=1))
    assert np.array_equal(np.sort(a[2], axis=1), np.sort(b[2], axis=1)) and np.array_equal(a[3], b[3])
    params, pair = abi.make_params(bc=STEP_BC)
    m.flatten(params, pair)
    r.flatten(params, pair)
    fa, fb = m.flat_arrays(), r.flat_arrays()
    for k in fa:
        assert np.array_equal(fa[k], fb[k]), k
    with pytest.raises(abi.DfloError):
        abi.Mesh(gmsh_path=str(tmp_path / "missing.msh"), lib=lib)


@pytest.mark.parametrize("world", [2, 3, 8])
@pytest.mark.parametrize("layers_case", ["none", "tvb"])
def test_partition_ranges_and_halo_symmetry(world, layers_case):
    """contiguous cell-id ranges cover the mesh; what rank a sends to b is what b expects from a"""
    L = emu_lib()
    prm = dict(basis="Qk", degree=1, flux="lxf")
    if layers_case == "tvb":
        prm.update(limiter="TVB", M=0.0, beta=1.0)
    params, pair = abi.make_params(bc=STEP_BC, **prm)
    m = abi.Mesh("forward_step", [0.1], lib=L)
    flat = m.flatten(params, pair)
    engines = [abi.Engine(flat, params, rank=r, world=world, nccl_id=b"\0" * 128, lib=L, prefix="dflo_emu_") for r in range(world)]
    ranges = [e.cell_range() for e in engines]
    assert ranges[0][0] == 0 and ranges[-1][1] == m.n_cells
    assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
    sizes = [b - a for a, b in ranges]
    assert max(sizes) - min(sizes) <= 1
    u = np.arange(m.n_cells * engines[0].D, dtype=np.float64)
    for e in engines:
        e.set_solution(u)          # triggers the first halo exchange
    for e in engines:
        for i in range(L.dflo_emu_n_peers(e.h)):
            p = L.dflo_emu_peer_rank(e.h, i)
            assert L.dflo_emu_halo_send_count(e.h, p) == L.dflo_emu_halo_recv_count(engines[p].h, e.rank)
            assert L.dflo_emu_halo_send_count(e.h, p) > 0
        # ghost layers: 1 without limiter, 2 with TVB (SURVEY 8e option 2)
        assert L.dflo_emu_n_local(e.h) > e.cell_range()[1] - e.cell_range()[0]
    for e in engines:
        e.close()


ROW_MESHES = [
    ("isentropic_vortex", [16], PERIODIC_BOX), ("isentropic_vortex", [41], PERIODIC_BOX), ("isentropic_vortex", [3], PERIODIC_BOX),
    ("sod_tube", [100, 10], SOD_BC), ("forward_step", [0.05], STEP_BC),
    ("double_mach", [16], {0: "outflow", 1: "slip", 2: "outflow", 3: "inflow", 4: "inflow"}),
]


@pytest.mark.parametrize("kind,args,bc", ROW_MESHES, ids=[m[0] + str(m[1][0]) for m in ROW_MESHES])
@pytest.mark.parametrize("world,layers", [(1, 1), (2, 1), (3, 2)])
def test_row_kernel_tile_descriptors(kind, args, bc, world, layers):
    """row_desc.h: every face of every computed cell has exactly one agent in the register-blocked
    stage kernel's tile descriptors (own thread / L job), with the right neighbour, ghost-trace
    source, flip and reference "plus" side -- on lattice tiles, ragged multi-block tiles, tiny
    periodic meshes (periodic partner inside the tile) and the strip tiles of a ghost layer."""
    L = emu_lib()
    L.dflo_emu_rowdesc_check.argtypes = [ctypes.POINTER(abi.FlatMesh), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                         ctypes.POINTER(ctypes.c_int)]
    params, pair = abi.make_params(bc=bc, basis="Qk", degree=1, flux="lxf")
    m = abi.Mesh(kind, args, lib=L)
    flat = m.flatten(params, pair)
    if m.n_cells < world:
        pytest.skip("mesh smaller than world")
    for n1 in (2, 3, 4, 5):
        for rank in range(world):
            nt = ctypes.c_int(0)
            bad = L.dflo_emu_rowdesc_check(flat, n1, rank, world, layers, ctypes.byref(nt))
            assert bad == 0, (n1, rank, bad)
            assert nt.value >= 1


# ---------------------------------------------------------------------------------------------
# output path: VTU writers of src/output.cc:33-87 on a host copy of the solution
# ---------------------------------------------------------------------------------------------
def _read_vtu(path):
    import xml.etree.ElementTree as ET
    piece = ET.parse(path).getroot().find("UnstructuredGrid/Piece")
    arr = lambda e: np.array(e.text.split(), dtype=np.float64)
    kids = lambda tag: list(piece.find(tag)) if piece.find(tag) is not None else []
    out = dict(n_points=int(piece.get("NumberOfPoints")), n_cells=int(piece.get("NumberOfCells")),
               points=arr(piece.find("Points/DataArray")).reshape(-1, 3),
               cells={e.get("Name"): arr(e).astype(np.int64) for e in piece.find("Cells")},
               point_names=[e.get("Name") for e in kids("PointData")],
               cell_names=[e.get("Name") for e in kids("CellData")])
    out["point"] = {e.get("Name"): arr(e) for e in kids("PointData")}
    for e in kids("PointData"):          # vector ranges "A__B" (3 components, z = 0): also under their component names
        if e.get("NumberOfComponents") == "3":
            v = arr(e).reshape(-1, 3)
            assert np.all(v[:, 2] == 0)
            for i, n in enumerate(e.get("Name").split("__")):
                out["point"][n] = v[:, i]
    out["cell"] = {e.get("Name"): arr(e) for e in kids("CellData")}
    return out


@pytest.mark.parametrize("basis,k", [("Qk", 1), ("Qk", 3), ("Pk", 1), ("Pk", 2), ("Pk", 3), ("Qk", 0)])
def test_vtu_output_reproduces_the_dg_polynomial(lib, tmp_path, basis, k):
    """solution-NNN.vtu: degree x degree sub-quads per cell (DataOut::build_patches (mapping, fe.degree),
    src/output.cc:45), the reference's array names in the reference's order (component_names + Postprocessor,
    src/equation.cc:59-145), vertex values = the cell's own polynomial.  The field is a polynomial the element
    represents exactly, so the file must reproduce it (and |grad rho|^2) at every patch vertex to the 10
    digits written; DoFs come from the oracle's interpolation/projection."""
    from oracle import oracle as O
    nx, ny = 4, 3
    bc = {0: "slip", 1: "outflow", 2: "inflow"}
    params, pair = abi.make_params(bc=bc, basis=basis, degree=k)
    mesh = abi.Mesh("rectangle", [nx, ny, 0.0, 1.0, 0.0, 0.6, 0, 1, 2, 2], lib=lib)
    mesh.flatten(params, pair)
    v, c, bl, bi = mesh.primitive()
    o = O.Oracle(v, c, bl, bi, O.make_params(bc=bc, basis=basis, degree=k))
    a, b = (k, k) if basis == "Qk" else (max(k - 1, 0), min(k, 1))   # x^a y^b stays inside Qk / Pk

    def field(x, y):
        rho = 2.0 + 0.3 * x ** k + 0.2 * y ** k + 0.1 * x ** a * y ** b
        return np.stack([0.5 + 0.4 * x ** a * y ** b, -0.2 + 0.1 * x ** k, rho, 5.0 + y ** k + 0.5 * x ** min(k, 1)], axis=-1)

    def grad_rho(x, y):
        dx = 0.3 * k * x ** max(k - 1, 0) * (k > 0) + 0.1 * a * x ** max(a - 1, 0) * y ** b * (a > 0)
        dy = 0.2 * k * y ** max(k - 1, 0) * (k > 0) + 0.1 * b * x ** a * y ** max(b - 1, 0) * (b > 0)
        return dx, dy

    xq = o.cell_qpoints()
    o.set_initial_condition(field(xq[..., 0], xq[..., 1]))
    u = o.solution()
    path = str(tmp_path / "solution-000.vtu")
    mesh.write_solution_vtu(path, u, basis, k, schlieren_plot=True, time=0.25, cycle=7)
    assert "<!-- time 0.25 cycle 7 -->" in open(path).read(300)
    f = _read_vtu(path)
    nsub = max(k, 1)
    nc = nx * ny
    assert f["n_points"] == nc * (nsub + 1) ** 2 and f["n_cells"] == nc * nsub ** 2
    # DataOutBase::write_vtu: vector ranges first (momentum, velocity), then the scalars in the order added
    assert f["point_names"] == ["XMomentum__YMomentum", "XVelocity__YVelocity", "Density", "Energy", "Pressure", "schlieren_plot"]
    # geometry: every cell's own (nsub+1)^2 lexicographic vertices, nothing shared between cells
    fa = mesh.flat_arrays()
    t = np.arange(nsub + 1) / nsub
    px = fa["origin"][:, None, None, 0] + t[None, None, :] * fa["size"][:, None, None, 0] + 0 * t[None, :, None]
    py = fa["origin"][:, None, None, 1] + t[None, :, None] * fa["size"][:, None, None, 1] + 0 * t[None, None, :]
    np.testing.assert_allclose(f["points"][:, 0], px.reshape(-1), rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(f["points"][:, 1], py.reshape(-1), rtol=1e-9, atol=1e-12)
    assert np.all(f["points"][:, 2] == 0)
    conn = f["cells"]["connectivity"].reshape(-1, 4)
    assert np.all(f["cells"]["types"] == 9) and np.array_equal(f["cells"]["offsets"], 4 * np.arange(1, len(conn) + 1))
    assert np.array_equal(conn // (nsub + 1) ** 2, np.repeat(np.arange(nc), nsub ** 2)[:, None] * np.ones(4, np.int64))
    q = f["points"][conn]                                   # counter-clockwise unit sub-quads
    assert np.all(q[:, 1, 0] > q[:, 0, 0]) and np.all(q[:, 2, 1] > q[:, 1, 1]) and np.all(q[:, 3, 0] == q[:, 0, 0])
    # values
    w = field(px.reshape(-1), py.reshape(-1))
    if k == 0:   # piecewise constants: the cell value at every vertex
        xc = fa["origin"] + 0.5 * fa["size"]
        w = np.repeat(field(xc[:, 0], xc[:, 1]), 4, axis=0)
    for i, name in enumerate(["XMomentum", "YMomentum", "Density", "Energy"]):
        np.testing.assert_allclose(f["point"][name], w[:, i], rtol=2e-9, atol=1e-11)
    np.testing.assert_allclose(f["point"]["XVelocity"], w[:, 0] / w[:, 2], rtol=2e-9)
    np.testing.assert_allclose(f["point"]["YVelocity"], w[:, 1] / w[:, 2], rtol=2e-9, atol=1e-11)
    np.testing.assert_allclose(f["point"]["Pressure"], 0.4 * (w[:, 3] - 0.5 * (w[:, 0] ** 2 + w[:, 1] ** 2) / w[:, 2]), rtol=2e-9)
    gx, gy = grad_rho(px.reshape(-1), py.reshape(-1)) if k > 0 else (np.zeros(4 * nc), np.zeros(4 * nc))
    np.testing.assert_allclose(f["point"]["schlieren_plot"], gx ** 2 + gy ** 2, rtol=1e-8, atol=1e-10)
    # without the schlieren switch the array is absent (Postprocessor::get_names)
    mesh.write_solution_vtu(path, u, basis, k)
    assert _read_vtu(path)["point_names"][-1] == "Pressure"
    # wrong vector length / unwritable path are errors, not silent truncation
    with pytest.raises(abi.DfloError):
        mesh.write_solution_vtu(path, u[:-1], basis, k)
    with pytest.raises(abi.DfloError):
        mesh.write_solution_vtu(str(tmp_path / "no_such_dir" / "x.vtu"), u, basis, k)


def test_shock_vtu_cell_vectors(lib, tmp_path):
    """shock.vtu (src/output.cc:70-79): one quad per cell, mu_shock and shock_indicator the way DataOut writes cell
    vectors -- point data, constant on the four vertices of a cell."""
    params, pair = abi.make_params(bc={0: "slip", 1: "outflow", 2: "inflow"}, basis="Qk", degree=2)
    mesh = abi.Mesh("sod_tube", [10, 2], lib=lib)
    mesh.flatten(params, pair)
    nc = mesh.n_cells
    ind = np.linspace(0.0, 3.0, nc)
    path = str(tmp_path / "shock.vtu")
    mesh.write_shock_vtu(path, ind)
    f = _read_vtu(path)
    assert f["n_cells"] == nc and f["n_points"] == 4 * nc and f["point_names"] == ["mu_shock", "shock_indicator"]
    assert f["cell_names"] == []
    np.testing.assert_allclose(f["point"]["shock_indicator"], np.repeat(ind, 4), rtol=1e-9)
    assert np.all(f["point"]["mu_shock"] == 0)
    mesh.write_shock_vtu(path, ind, mu_shock=2 * ind)
    np.testing.assert_allclose(_read_vtu(path)["point"]["mu_shock"], np.repeat(2 * ind, 4), rtol=1e-9)


def test_vtu_pieces_of_the_mpi_tree_tile_the_whole_file(lib, tmp_path):
    """src_mpi/output.cc:34-86: every process writes the cells it owns (+ "subdomain"); the pieces of a partition put
    end to end are the single-process file."""
    bc = {0: "slip", 1: "outflow", 2: "inflow"}
    params, pair = abi.make_params(bc=bc, basis="Pk", degree=2)
    mesh = abi.Mesh("sod_tube", [12, 3], lib=lib)
    mesh.flatten(params, pair)
    nc, D = mesh.n_cells, 24
    u = np.random.default_rng(3).uniform(0.5, 2.0, nc * D)
    whole = str(tmp_path / "whole.vtu")
    mesh.write_solution_vtu(whole, u, "Pk", 2, schlieren_plot=True, time=0.5, cycle=2)
    fw = _read_vtu(whole)
    cuts = [0, 10, 11, 36]
    pts, arrays = [], {n: [] for n in fw["point_names"]}
    for r in range(3):
        p = str(tmp_path / ("solution-0002.%03d.vtu" % r))
        mesh.write_solution_vtu(p, u, "Pk", 2, schlieren_plot=True, time=0.5, cycle=2, cells=(cuts[r], cuts[r + 1]), subdomain=r)
        f = _read_vtu(p)
        n = cuts[r + 1] - cuts[r]
        assert f["n_cells"] == 4 * n and f["n_points"] == 9 * n and f["point_names"] == fw["point_names"] + ["subdomain"]
        assert np.all(f["point"]["subdomain"] == r)
        assert f["cells"]["connectivity"].max() == 9 * n - 1          # piece-local point numbering
        pts.append(f["points"])
        for name in fw["point_names"]:
            arrays[name].append(f["point"][name])
    assert np.array_equal(np.concatenate(pts), fw["points"])
    for name in fw["point_names"]:
        assert np.array_equal(np.concatenate(arrays[name]), fw["point"][name])
    with pytest.raises(abi.DfloError):
        mesh.write_solution_vtu(whole, u, "Pk", 2, cells=(5, nc + 1))


@pytest.mark.parametrize("basis,k", [("Qk", 1), ("Qk", 3), ("Pk", 1), ("Pk", 2), ("Pk", 3)])
def test_angular_momentum_of_a_polynomial_field(lib, basis, k):
    """compute_angular_momentum (src/claw.cc:604-635): int (x m_y - y m_x) with QGauss(k+1)^2 per cell -- exact
    for momentum of total degree <= k, checked against a much finer quadrature of the analytic field."""
    bc = {0: "slip", 1: "outflow", 2: "inflow"}
    params, pair = abi.make_params(bc=bc, basis=basis, degree=k)
    nx, ny, x0, x1, y0, y1 = 5, 3, -0.4, 1.1, 0.2, 0.8
    mesh = abi.Mesh("rectangle", [nx, ny, x0, x1, y0, y1, 0, 1, 2, 2], lib=lib)
    mesh.flatten(params, pair)
    v, c, bl, bi = mesh.primitive()
    o = O.Oracle(v, c, bl, bi, O.make_params(bc=bc, basis=basis, degree=k))
    mxf = lambda x, y: 0.3 + 0.7 * x ** k - 0.2 * y ** k + 0.5 * x ** (k - 1) * y
    myf = lambda x, y: -0.1 + 0.4 * y ** k + 0.9 * x * y ** (k - 1)
    xq = o.cell_qpoints()
    X, Y = xq[..., 0], xq[..., 1]
    o.set_initial_condition(np.stack([mxf(X, Y), myf(X, Y), 1.0 + 0 * X, 3.0 + 0 * X], axis=-1))
    got = mesh.angular_momentum(o.solution(), basis, k)
    g, w = np.polynomial.legendre.leggauss(12)
    gx, wx = 0.5 * (x1 - x0) * g + 0.5 * (x1 + x0), 0.5 * (x1 - x0) * w
    gy, wy = 0.5 * (y1 - y0) * g + 0.5 * (y1 + y0), 0.5 * (y1 - y0) * w
    XX, YY = np.meshgrid(gx, gy, indexing="ij")
    want = np.sum((XX * myf(XX, YY) - YY * mxf(XX, YY)) * wx[:, None] * wy[None, :])
    assert abs(got - want) <= 1e-13 * max(1.0, abs(want))
    with pytest.raises(abi.DfloError):
        mesh.angular_momentum(o.solution()[:-1], basis, k)


def test_tecplot_output_carries_the_same_patches_as_the_vtu(lib, tmp_path):
    """"output: format = tecplot" (src/output.cc:51-52, 65-66): FEBLOCK zone, x / y / variables over the same patch
    vertices as the VTU file, 1-based counter-clockwise quadrilaterals."""
    bc = {0: "slip", 1: "outflow", 2: "inflow"}
    params, pair = abi.make_params(bc=bc, basis="Qk", degree=2)
    mesh = abi.Mesh("sod_tube", [6, 2], lib=lib)
    mesh.flatten(params, pair)
    nc, D = mesh.n_cells, 36
    u = np.random.default_rng(5).uniform(0.5, 2.0, nc * D)
    vtu, plt = str(tmp_path / "s.vtu"), str(tmp_path / "s.plt")
    mesh.write_solution_vtu(vtu, u, "Qk", 2, schlieren_plot=True, time=0.125)
    mesh.write_solution_tecplot(plt, u, "Qk", 2, schlieren_plot=True, time=0.125)
    f = _read_vtu(vtu)
    lines = [ln for ln in open(plt).read().splitlines() if not ln.startswith("#")]
    names = ["XMomentum", "YMomentum", "Density", "Energy", "XVelocity", "YVelocity", "Pressure", "schlieren_plot"]
    assert lines[0] == 'Variables="x", "y", ' + ", ".join('"%s"' % n for n in names)   # tecplot: one variable per component
    assert lines[1] == 'zone t="time=0.125" f=feblock, n=%d, e=%d, et=quadrilateral' % (f["n_points"], f["n_cells"])
    blocks = "\n".join(lines[2:]).split("\n\n")
    assert len(blocks) == 2 + 8 + 1
    np.testing.assert_array_equal(np.array(blocks[0].split(), float), f["points"][:, 0])
    np.testing.assert_array_equal(np.array(blocks[1].split(), float), f["points"][:, 1])
    for i, name in enumerate(names):
        np.testing.assert_array_equal(np.array(blocks[2 + i].split(), float), f["point"][name])
    conn = np.array(blocks[-1].split(), int).reshape(-1, 4)
    np.testing.assert_array_equal(conn - 1, f["cells"]["connectivity"].reshape(-1, 4))


def test_bench_e2e_entry_contract():
    """bench.py `e2e`: the contract keys, and the batches-in-flight figure next to the serial chain."""
    import bench
    serial = bench.e2e_entry(12_000_000, 20, 0.03, None, 3, 4_194_304)
    assert {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"} <= set(serial)
    assert serial["in_flight"] == 1 and serial["h2d_bytes_per_step"] == serial["d2h_bytes_per_step"] == 4_194_304 * 8
    assert abs(serial["value"] - 12_000_000 * 20 / 0.03 / 1e6) < 1e-6
    piped = bench.e2e_entry(12_000_000, 20, 0.03, 0.05, 3, 4_194_304)
    assert piped["in_flight"] == 3 and piped["steps"] == 60 and piped["value_one_context_serial"] == serial["value"]
    assert abs(piped["value"] - 12_000_000 * 60 / 0.05 / 1e6) < 1e-6


def test_initial_condition_syntax_error_is_reported(lib, tmp_path):
    """FunctionParser::initialize throws on a malformed `initial condition/w_i value` (src/parameters.cc:524-526);
    the front end must report it instead of running on an all-zero component."""
    L = _claw_api(lib)
    base = open(os.path.join(PRM_DIR, "cfg3_sod_P2_hllc_tvb_pos.prm")).read()
    p = tmp_path / "bad_ic.prm"
    p.write_text(base.replace("set w_2 value = 1.0*(x<=0.5) + 0.125*(x>0.5)", "set w_2 value = 1.0*(x<=0.5 + 0.125*(x>0.5)"))
    h = L.dflo_claw_create(str(p).encode(), b"sod_tube 10 2", None, 0)
    if h:   # the deck parses: the error must surface when the initial condition is evaluated
        n = L.dflo_claw_n_dofs(h)
        u = np.zeros(n)
        rc = L.dflo_claw_initial_condition(h, u.ctypes.data_as(abi.c_double_p), n)
        assert rc != 0
        assert b"w_2" in L.dflo_host_last_error()
        L.dflo_claw_destroy(h)
    else:
        assert L.dflo_host_last_error()


def test_compression_corner_deck_q1_local(lib):
    """The q1 example deck (examples/compression_corner/input.prm rewritten): mapping = q1 and time step type = local
    reach the engine parameters; the mesh is general quadrilaterals; the host-side initial condition is evaluated at the
    MAPPED support points."""
    L = _claw_api(lib)
    prm = os.path.join(ROOT, "tests", "golden", "prm_q1", "compression_corner_Q1_kfvs_q1_local.prm")
    h = L.dflo_claw_create(prm.encode(), b"compression_corner 9 29 19", None, abi.COMPAT["src"])
    assert h, L.dflo_host_last_error()
    p = L.dflo_claw_params(h).contents
    assert (p.basis, p.degree, p.flux_type, p.mapping, p.local_time_step, p.limiter_type) == (0, 1, abi.FLUX["kfvs"], 1, 1, 0)
    assert p.bc_kind[1] == abi.BC["slip"] and p.bc_kind[2] == abi.BC["inflow"] and p.bc_kind[3] == abi.BC["outflow"]
    m = abi.Mesh(handle=L.dflo_claw_mesh(h), owned=False, lib=lib)
    assert m.n_cells == (9 + 29) * 19
    L.dflo_claw_destroy(h)
    # an initial condition that depends on x: the value at a support point of the last ramp cell is the value at its mapped position
    over = b"subsection initial condition\n set w_2 value = 1.0 + x + 10*y\nend\n"
    h = L.dflo_claw_create(prm.encode(), b"compression_corner 2 3 2", over, abi.COMPAT["src"])
    assert h, L.dflo_host_last_error()
    n = L.dflo_claw_n_dofs(h)
    u = np.zeros(n)
    assert L.dflo_claw_initial_condition(h, u.ctypes.data_as(abi.c_double_p), n) == 0
    mesh = abi.Mesh(handle=L.dflo_claw_mesh(h), owned=False, lib=lib)
    v, c, _, _ = mesh.primitive()
    g = 0.5 - 0.5 / np.sqrt(3.0)                     # first Gauss node of QGauss(2)
    last = v[c[-1]]                                   # [4][2] vertices of the last cell
    N = np.array([(1 - g) * (1 - g), g * (1 - g), (1 - g) * g, g * g])
    x, y = N @ last[:, 0], N @ last[:, 1]
    assert abs(u.reshape(-1, 4, 4)[-1, 2, 0] - (1.0 + x + 10 * y)) < 1e-13
    assert abs(y - (last[0, 1] + g * (last[2, 1] - last[0, 1]))) > 1e-3   # ... which is not where a rectangle would put it
    L.dflo_claw_destroy(h)


def test_dealii_adapter_compiles_and_flattens(tmp_path):
    """adapters/dealii_flatten.h against adapters/dealii_mock (a MOCK of the deal.II calls it makes): compiled with g++,
    the flat mesh equals the library's own flattening of the same rectangle, boundary expressions are captured from the
    ParameterHandler; without a GPU the library must refuse to create a context (no CPU fallback)."""
    import subprocess
    exe = str(tmp_path / "test_adapter")
    subprocess.run(["g++", "-std=c++17", "-Wall", "-I" + os.path.join(ROOT, "include"), "-I" + os.path.join(ROOT, "adapters"),
                    "-I" + os.path.join(ROOT, "adapters", "dealii_mock"), os.path.join(ROOT, "adapters", "test_adapter.cc"),
                    "-L" + os.path.join(ROOT, "dflo_b200", "csrc"), "-ldflo_b200", "-Wl,-rpath," + os.path.join(ROOT, "dflo_b200", "csrc"),
                    "-o", exe], check=True)
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0 and "all checks passed" in out.stdout, out.stdout + out.stderr
