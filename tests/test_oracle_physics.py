"""CPU tier: pins the oracle's point-wise physics (and the product's device functions, compiled
for the CPU by tests/emu) against golden vectors produced by the reference's own equation.h
object code (tests/golden/make_flux_kat.py), plus the analytic flux invariants of SURVEY.md 4."""
import ctypes
import os

import numpy as np
import pytest

from helpers import emu_lib
from oracle import oracle as O

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "flux_kat.npz"))
FLUXES = ["lxf", "sw", "kfvs", "roe", "hllc"]
_dp = ctypes.POINTER(ctypes.c_double)


def _same(a, b):
    return np.array_equal(a, b, equal_nan=True)


@pytest.fixture(scope="module")
def restated():
    return O.Physics("restated")


def test_survey_known_answers(restated):
    """The hand-recorded values of SURVEY.md 8(c) (reference equation.h, n=(0.6,0.8))."""
    n, WL, WR = (0.6, 0.8), (0.3, -0.1, 1.0, 2.5), (0.05, 0.02, 0.125, 0.25)
    kat = {
        "lxf": (0.52199004209564104, 0.3447366197940922, 0.68613714733474396, 1.8141846188607702),
        "roe": (0.48566775495783115, 0.42439728290399226, 0.47085985533785968, 1.4973618712745225),
        "hllc": (0.45559658522249147, 0.35586620540619268, 0.50319181644759203, 1.370058573283619),
        "kfvs": (0.46126118003871136, 0.40063848597968527, 0.42258017447492202, 1.3107229181031321),
        "sw": (0.47042244344292855, 0.39701299976028093, 0.45290785537573736, 1.563905004396974),
    }
    for name, want in kat.items():
        got = restated.flux(O.FLUX[name], n, WL, WR)
        assert np.array_equal(got, np.array(want)), name
    assert np.abs(restated.wminus(O.BC["slip"], n, WL, WL) - np.array([0.18, -0.26, 1.0, 2.5])).max() < 1e-15  # printed rounded
    assert restated.eigen(WL)[1][2][0] == -0.12077157840536142
    assert restated.flux(O.FLUX["hllc"], n, WL, WL)[3] == pytest.approx(0.348, abs=1e-15)


@pytest.mark.parametrize("variant", ["restated", "physref"])
def test_oracle_physics_bit_exact_vs_reference_golden(variant):
    """restated C == the reference's equation.h, bit for bit, on every golden vector; where
    oracle/_ref exists the reference object code itself is re-checked against the fixture."""
    if variant == "physref" and not os.path.exists(O.lib_path("physref")):
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    P = O.Physics(variant)
    g = GOLD
    n = len(g["WL"])
    for f in range(5):
        got = np.array([P.flux(f, g["N"][i], g["WL"][i], g["WR"][i], g["AL"][i], g["AR"][i]) for i in range(n)])
        assert _same(got, g["flux"][f]), FLUXES[f]
    assert _same(np.array([P.flux_matrix(w) for w in g["WL"]]), g["fmat"])
    for k in range(5):
        got = np.array([P.wminus(k, g["N"][i], g["WL"][i], g["G"][i]) for i in range(n)])
        assert _same(got, g["wminus"][k]), k
    eig = np.array([np.stack(P.eigen(w)) for w in g["WL"]])
    assert _same(eig, g["eig"])
    assert _same(np.array([P.to_char(g["eig"][i, 1], g["WR"][i]) for i in range(n)]), g["tochar"])
    assert _same(np.array([P.to_con(g["eig"][i, 0], g["WR"][i]) for i in range(n)]), g["tocon"])
    sc = np.array([[P.pressure(w), P.sound_speed(w), P.max_eigenvalue(w)] for w in g["WL"]])
    assert _same(sc, g["scal"])


def _emu_flux(L, f, n, wl, wr, al, ar):
    out = np.zeros(4)
    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (n, wl, wr, al, ar)]
    L.dflo_emu_numerical_flux(f, *[a.ctypes.data_as(_dp) for a in args], out.ctypes.data_as(_dp))
    return out


def test_device_physics_matches_reference_golden():
    """dflo_b200/csrc/euler.cuh (the device functions, compiled for the host by tests/emu) against
    the reference's outputs.  Not bit-exact by design (the device code contracts to FMA and hoists
    reciprocals); a few ulp relative to the magnitude of the flux vector."""
    L = emu_lib()
    g = GOLD
    n = len(g["WL"])
    P = O.Physics("restated")
    fmat_r = np.array([P.flux_matrix(w) for w in g["WR"]])
    for f in range(5):
        for i in range(n):
            want = g["flux"][f, i]
            if not np.all(np.isfinite(want)):
                continue
            got = _emu_flux(L, f, g["N"][i], g["WL"][i], g["WR"][i], g["AL"][i], g["AR"][i])
            # split fluxes cancel: measure against the size of the one-sided physical fluxes
            scale = max(1.0, np.abs(want).max(), np.abs(g["fmat"][i]).max(), np.abs(fmat_r[i]).max())
            assert np.abs(got - want).max() <= 2e-13 * scale, (FLUXES[f], i, got, want)
    for i in range(n):   # kep (src_mpi/equation.h:842-921) against its own golden vectors
        want = g["flux_kep"][i]
        got = _emu_flux(L, O.FLUX["kep"], g["N"][i], g["WL"][i], g["WR"][i], g["AL"][i], g["AR"][i])
        scale = max(1.0, np.abs(want).max(), np.abs(g["fmat"][i]).max(), np.abs(fmat_r[i]).max())
        assert np.abs(got - want).max() <= 2e-13 * scale, ("kep", i, got, want)
    for i in range(n):
        F = np.zeros(8)
        w = np.ascontiguousarray(g["WL"][i])
        L.dflo_emu_flux_matrix(w.ctypes.data_as(_dp), F.ctypes.data_as(_dp))
        assert np.abs(F.reshape(4, 2) - g["fmat"][i]).max() <= 1e-13 * max(1.0, np.abs(g["fmat"][i]).max())
        for k in range(5):
            out = np.zeros(4)
            nn, gg = np.ascontiguousarray(g["N"][i]), np.ascontiguousarray(g["G"][i])
            L.dflo_emu_wminus(k, nn.ctypes.data_as(_dp), w.ctypes.data_as(_dp), gg.ctypes.data_as(_dp), out.ctypes.data_as(_dp))
            assert np.abs(out - g["wminus"][k, i]).max() <= 1e-13 * max(1.0, np.abs(g["wminus"][k, i]).max())
        m = [np.zeros(16) for _ in range(4)]
        L.dflo_emu_eigen(w.ctypes.data_as(_dp), *[x.ctypes.data_as(_dp) for x in m])
        got = np.stack([x.reshape(4, 4) for x in m])
        assert np.abs(got - g["eig"][i]).max() <= 1e-12 * max(1.0, np.abs(g["eig"][i]).max())


@pytest.mark.parametrize("flux", FLUXES)
def test_flux_invariants(restated, flux):
    """Consistency H(W,W,n) = F(W).n, antisymmetry H(L,R,n) = -H(R,L,-n), rotation invariance."""
    g = GOLD
    f = O.FLUX[flux]
    # KFVS sits on the A&S erf fit (absolute error 1.5e-7): its identities hold to that level only
    tol_c, tol_a = (1e-6, 1e-6) if flux == "kfvs" else (1e-11, 1e-10)
    for i in range(60):
        W, R, n = g["WL"][i], g["WR"][i], g["N"][i]
        Fn = restated.flux_matrix(W) @ n
        H = restated.flux(f, n, W, W)
        assert np.abs(H - Fn).max() <= tol_c * max(1.0, np.abs(Fn).max())
        if flux in ("lxf", "roe", "hllc", "kfvs"):
            a = restated.flux(f, n, W, R, W, R)
            b = restated.flux(f, -n, R, W, R, W)
            assert np.abs(a + b).max() <= tol_a * max(1.0, np.abs(a).max())
        # rotate the frame by 90 degrees: momentum components rotate, scalars do not
        rot = lambda w: np.array([-w[1], w[0], w[2], w[3]])
        a = restated.flux(f, n, W, R, W, R)
        b = restated.flux(f, np.array([-n[1], n[0]]), rot(W), rot(R), rot(W), rot(R))
        assert np.abs(rot(a) - b).max() <= 1e-10 * max(1.0, np.abs(a).max())


def test_kfvs_uses_abramowitz_stegun_erf(restated):
    """equation.h:686-709: ERF is the 5-term A&S 7.1.26 fit (|error| ~ 1.5e-7), not std::erf.  A
    numpy KFVS built on the A&S fit reproduces the reference flux to round-off; the same formula
    on the exact erf does not."""
    import math

    def as_erf(x):
        t = 1.0 / (1.0 + 0.3275911 * abs(x))
        y = 1.0 - (((((1.061405429 * t - 1.453152027) * t) + 1.421413741) * t - 0.284496736) * t + 0.254829592) * t * math.exp(-x * x)
        return math.copysign(y, x)

    def kfvs(n, Wp, Wm, erf):
        out = np.zeros(4)
        for sign, W in ((+1.0, Wp), (-1.0, Wm)):
            rho, E = W[2], W[3]
            vn = (W[0] * n[0] + W[1] * n[1]) / rho
            p = 0.4 * (E - 0.5 * (W[0] ** 2 + W[1] ** 2) / rho)
            beta = 0.5 * rho / p
            s = vn * math.sqrt(beta)
            A = 0.5 * (1.0 + sign * erf(s))
            B = 0.5 * sign * math.exp(-s * s) / math.sqrt(math.pi * beta)
            uf = vn * A + B
            out += np.array([p * n[0] * A + W[0] * uf, p * n[1] * A + W[1] * uf, rho * uf,
                             (E + p) * vn * A + (E + 0.5 * p) * B])
        return out
    n, WL, WR = np.array([0.6, 0.8]), np.array([0.3, -0.1, 1.0, 2.5]), np.array([0.05, 0.02, 0.125, 0.25])
    ref = restated.flux(O.FLUX["kfvs"], n, WL, WR)
    assert np.abs(kfvs(n, WL, WR, as_erf) - ref).max() < 1e-14
    assert np.abs(kfvs(n, WL, WR, math.erf) - ref).max() > 1e-9


def test_minmod_device():
    """limiter.cc:15-30 TVB minmod through the device function."""
    L = emu_lib()
    mm = L.dflo_emu_minmod
    assert mm(0.5, 2.0, 3.0, 1.0) == 0.5            # |a| < M dx^2 -> a
    assert mm(2.0, 3.0, 4.0, 1.0) == 2.0            # same sign -> smallest magnitude
    assert mm(-2.0, -3.0, -1.5, 1.0) == -1.5
    assert mm(2.0, -3.0, 4.0, 1.0) == 0.0           # sign change -> 0
    assert mm(2.0, 3.0, 0.0, 1.0) == 0.0


def test_axis_fluxes_match_reference_convention():
    """euler.cuh face_flux_axis (what the register-blocked stage kernel calls: the face solved
    along +e_x / +e_y, the integrating cell on either side) against the reference's general-normal
    flux called the way MeshWorker does: n = outward normal of the integrating ("plus") cell."""
    L = emu_lib()
    g = GOLD
    P = O.Physics("restated")
    fmat_r = np.array([P.flux_matrix(w) for w in g["WR"]])
    worst = 0.0
    for f in range(6):   # the five fluxes of src/ and kep of src_mpi/
        for i in range(0, len(g["WL"]), 2):
            lo, hi, alo, ahi = g["WL"][i], g["WR"][i], g["AL"][i], g["AR"][i]
            scale = max(1.0, np.abs(g["fmat"][i]).max(), np.abs(fmat_r[i]).max())
            for d in (0, 1):
                e = np.array([1.0, 0.0]) if d == 0 else np.array([0.0, 1.0])
                for plus_low in (1, 0):
                    # reference call: plus cell first, its outward normal; flux along +e_d = sign * that
                    want = P.flux(f, e, lo, hi, alo, ahi) if plus_low else -P.flux(f, -e, hi, lo, ahi, alo)
                    if not np.all(np.isfinite(want)):
                        continue
                    got = np.zeros(4)
                    args = [np.ascontiguousarray(a, dtype=np.float64) for a in (lo, hi, alo, ahi)]
                    L.dflo_emu_face_flux_axis(f, d, plus_low, *[a.ctypes.data_as(_dp) for a in args], got.ctypes.data_as(_dp))
                    # kep's dissipation is built from the (here unrelated, random) averages: measure against the flux itself too
                    err = np.abs(got - want).max() / max(scale, np.abs(want).max())
                    worst = max(worst, err)
                    assert err <= 2e-13, (f, i, d, plus_low, got, want)
    # zero normal velocity: the reference's A&S ERF is not odd at s = 0; the mirrored evaluation must reproduce it
    W = np.array([0.0, 0.0, 1.4, 8.8])
    for d in (0, 1):
        e = np.array([1.0, 0.0]) if d == 0 else np.array([0.0, 1.0])
        want = -P.flux(O.FLUX["kfvs"], -e, W, W, W, W)
        got = np.zeros(4)
        L.dflo_emu_face_flux_axis(O.FLUX["kfvs"], d, 0, *[W.ctypes.data_as(_dp)] * 4, got.ctypes.data_as(_dp))
        assert np.abs(got - want).max() <= 1e-14 * 8.8


def test_kep_flux_restated_bit_exact_vs_src_mpi_golden(restated):
    """kep_flux (src_mpi/equation.h:842-921, with logavg 27-45 and kep_diff_matrix 749-837): the
    plain-C restatement against the golden vectors produced by that header's own object code; where
    oracle/_ref exists the object code is re-checked against the fixture."""
    g = GOLD
    n = len(g["WL"])
    ref = None
    p = os.path.join(os.path.dirname(O.lib_path("physref")), "libphys_reference_mpi.so")
    if os.path.exists(p):
        ref = ctypes.CDLL(p)
    for i in range(n):
        got = restated.flux(O.FLUX["kep"], g["N"][i], g["WL"][i], g["WR"][i], g["AL"][i], g["AR"][i])
        assert _same(got, g["flux_kep"][i]), i
        if ref is not None:
            a = [np.ascontiguousarray(x, dtype=np.float64) for x in (g["N"][i], g["WL"][i], g["WR"][i], g["AL"][i], g["AR"][i])]
            o = np.zeros(4)
            ref.phys_mpi_kep_flux(*[x.ctypes.data_as(_dp) for x in a], o.ctypes.data_as(_dp))
            assert _same(o, g["flux_kep"][i]), i
    # consistency: H(W, W, n) = F(W).n (the entropy-variable jump vanishes)
    for i in range(40):
        W, nn = g["WL"][i], g["N"][i]
        Fn = restated.flux_matrix(W) @ nn
        H = restated.flux(O.FLUX["kep"], nn, W, W, W, W)
        assert np.abs(H - Fn).max() <= 1e-11 * max(1.0, np.abs(Fn).max())


def test_mpi_extras_restated_bit_exact_and_device_functions():
    """Streamline eigenvector matrices of the minmax limiter (src_mpi/equation.h:299-335) and the external
    forcing vector (src_mpi/equation.h:1189-1202): the plain-C restatement is bit-exact to the golden vectors
    produced by that header's own object code (tests/golden/make_mpi_extras_kat.py); where oracle/_ref
    exists the object code is re-checked; the product's device functions (normalised velocity instead of
    cos/sin(atan2)) agree to round-off, and L R = I."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "mpi_extras_kat.npz"))
    Lr = O.load("restated")
    E = emu_lib()
    ref = None
    p = os.path.join(os.path.dirname(O.lib_path("physref")), "libphys_reference_mpi.so")
    if os.path.exists(p):
        ref = ctypes.CDLL(p)
    for i in range(len(g["W"])):
        w, f = np.ascontiguousarray(g["W"][i]), np.ascontiguousarray(g["F"][i])
        R, L, G = np.zeros(16), np.zeros(16), np.zeros(4)
        Lr.phys_eigen_stream_restated(w.ctypes.data_as(_dp), R.ctypes.data_as(_dp), L.ctypes.data_as(_dp))
        Lr.phys_ext_forcing_restated(w.ctypes.data_as(_dp), f.ctypes.data_as(_dp), G.ctypes.data_as(_dp))
        assert _same(R, g["R"][i]) and _same(L, g["L"][i]) and _same(G, g["G"][i]), i
        if ref is not None:
            R2, L2, G2 = np.zeros(16), np.zeros(16), np.zeros(4)
            ref.phys_mpi_eigen_stream(w.ctypes.data_as(_dp), R2.ctypes.data_as(_dp), L2.ctypes.data_as(_dp))
            ref.phys_mpi_ext_forcing(w.ctypes.data_as(_dp), f.ctypes.data_as(_dp), G2.ctypes.data_as(_dp))
            assert _same(R2, g["R"][i]) and _same(L2, g["L"][i]) and _same(G2, g["G"][i]), i
        Rd, Ld, Gd = np.zeros(16), np.zeros(16), np.zeros(4)
        E.dflo_emu_eigen_stream(w.ctypes.data_as(_dp), Rd.ctypes.data_as(_dp), Ld.ctypes.data_as(_dp))
        E.dflo_emu_forcing_ext(w.ctypes.data_as(_dp), f.ctypes.data_as(_dp), Gd.ctypes.data_as(_dp))
        assert np.abs(Rd - g["R"][i]).max() <= 1e-13 * max(1.0, np.abs(g["R"][i]).max()), i
        assert np.abs(Ld - g["L"][i]).max() <= 1e-13 * max(1.0, np.abs(g["L"][i]).max()), i
        assert np.abs(Gd - g["G"][i]).max() <= 1e-14 * max(1.0, np.abs(g["G"][i]).max()), i
        I = Ld.reshape(4, 4) @ Rd.reshape(4, 4)
        assert np.abs(I - np.eye(4)).max() <= 1e-10 * max(1.0, np.abs(Rd).max() * np.abs(Ld).max()), i
    # the forcing of src/ (equation.h:829-850) is the external force (0,-1)
    P = O.Physics("restated")
    for i in range(20):
        w, f, G = np.ascontiguousarray(g["W"][i]), np.array([0.0, -1.0]), np.zeros(4), 
        Lr.phys_ext_forcing_restated(w.ctypes.data_as(_dp), f.ctypes.data_as(_dp), G.ctypes.data_as(_dp))
        G0 = np.zeros(4)
        Lr.phys_forcing(w.ctypes.data_as(_dp), G0.ctypes.data_as(_dp))
        assert np.array_equal(G + 0.0, G0 + 0.0), i
