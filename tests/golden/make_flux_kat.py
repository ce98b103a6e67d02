"""Generates tests/golden/flux_kat.npz from the reference's OWN object code.

Run in the build container (needs /root/reference): `python tests/golden/make_flux_kat.py`.
oracle/_ref/libphys_reference.so is /root/reference/src/equation.h compiled unmodified against the
deal.II stub (oracle/Makefile `ref`); every number stored here is therefore an output of the
reference's EulerEquations<2> templates: numerical fluxes (equation.h:324-782), flux matrix
(158-193), forcing (829-850), boundary ghost states (939-1033), eigenvector matrices (225-265),
characteristic transforms (270-306), pressure / sound speed / max eigenvalue (84-152).
Inputs: numpy default_rng(0), rho in [0.1,10], p in [0.1,100], |v| <= 3c, plus
the hand-picked vectors of SURVEY.md 8(c), near-vacuum, supersonic and symmetric (W,W) states.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle as O  # noqa: E402


def random_states(rng, n):
    rho = rng.uniform(0.1, 10.0, n)
    p = rng.uniform(0.1, 100.0, n)
    c = np.sqrt(1.4 * p / rho)
    speed = rng.uniform(0.0, 3.0, n) * c
    ang = rng.uniform(0.0, 2 * np.pi, n)
    vx, vy = speed * np.cos(ang), speed * np.sin(ang)
    return np.stack([rho * vx, rho * vy, rho, p / 0.4 + 0.5 * rho * (vx * vx + vy * vy)], axis=1)


def main():
    O.build(ref=True)
    P = O.Physics("physref")
    assert P.name.startswith("reference"), P.name
    rng = np.random.default_rng(0)
    n = 600
    WL, WR = random_states(rng, n), random_states(rng, n)
    AL, AR = random_states(rng, n), random_states(rng, n)
    ang = rng.uniform(0.0, 2 * np.pi, n)
    N = np.stack([np.cos(ang), np.sin(ang)], axis=1)
    # hand-picked rows: SURVEY 8(c) vector, axis normals, identical states, strong shock, near vacuum
    WL[0], WR[0], N[0] = (0.3, -0.1, 1.0, 2.5), (0.05, 0.02, 0.125, 0.25), (0.6, 0.8)
    AL[0], AR[0] = WL[0], WR[0]
    for i, nn in enumerate([(1, 0), (-1, 0), (0, 1), (0, -1)]):
        N[1 + i] = nn
    WR[5:25] = WL[5:25]                      # consistency H(W,W,n) = F(W).n
    WL[25], WR[25] = (57.1576766498, -33.0, 8.0, 563.5), (0.0, 0.0, 1.4, 2.5)   # DMR shock
    WL[26], WR[26] = (0.0, 0.0, 1.0, 2.5), (0.0, 0.0, 0.125, 0.25)              # Sod
    WL[27], WR[27] = (1e-8, 0.0, 1e-6, 1e-6), (0.0, 0.0, 1.0, 2.5)              # near vacuum
    WL[28], WR[28] = (4.2, 0.0, 1.4, 8.8), (4.2, 0.0, 1.4, 8.8)                 # Mach 3 free stream
    N[28] = (1.0, 0.0)
    G = random_states(rng, n)
    flux = np.zeros((5, n, 4))
    for f in range(5):
        for i in range(n):
            flux[f, i] = P.flux(f, N[i], WL[i], WR[i], AL[i], AR[i])
    fmat = np.array([P.flux_matrix(w) for w in WL])
    wminus = np.zeros((5, n, 4))
    for k in range(5):
        for i in range(n):
            wminus[k, i] = P.wminus(k, N[i], WL[i], G[i])
    eig = np.array([np.stack(P.eigen(w)) for w in WL])     # [n][4 matrices][4][4] Rx, Lx, Ry, Ly
    tochar = np.array([P.to_char(eig[i, 1], WR[i]) for i in range(n)])
    tocon = np.array([P.to_con(eig[i, 0], WR[i]) for i in range(n)])
    scal = np.array([[P.pressure(w), P.sound_speed(w), P.max_eigenvalue(w)] for w in WL])
    # kep_flux exists only in the MPI tree: its header src_mpi/equation.h, compiled unmodified into a
    # library of its own (oracle/_ref/libphys_reference_mpi.so, oracle/Makefile `ref`)
    import ctypes
    M = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libphys_reference_mpi.so"))
    M.phys_mpi_impl_name.restype = ctypes.c_char_p
    assert M.phys_mpi_impl_name().decode().startswith("reference:src_mpi")
    dp = ctypes.POINTER(ctypes.c_double)
    flux_kep = np.zeros((n, 4))
    for i in range(n):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (N[i], WL[i], WR[i], AL[i], AR[i])]
        o = np.zeros(4)
        M.phys_mpi_kep_flux(*[x.ctypes.data_as(dp) for x in a], o.ctypes.data_as(dp))
        flux_kep[i] = o
    out = os.path.join(HERE, "flux_kat.npz")
    np.savez_compressed(out, WL=WL, WR=WR, AL=AL, AR=AR, N=N, G=G, flux=flux, fmat=fmat, wminus=wminus, eig=eig,
                        tochar=tochar, tocon=tocon, scal=scal, flux_kep=flux_kep)
    print("wrote", out, os.path.getsize(out), "bytes;", P.name)


if __name__ == "__main__":
    main()
