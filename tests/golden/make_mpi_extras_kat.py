"""Generates tests/golden/mpi_extras_kat.npz from the reference's OWN object code.

Run in the build container (needs /root/reference): `python tests/golden/make_mpi_extras_kat.py`.
oracle/_ref/libphys_reference_mpi.so is /root/reference/src_mpi/equation.h compiled unmodified against
the deal.II stub (oracle/Makefile `ref`); the numbers stored here are outputs of its
compute_eigen_matrix (W, R, L) -- the streamline-direction eigenvector matrices of the minmax limiter
(src_mpi/equation.h:299-335, src_mpi/limiter.cc:450) -- and of compute_forcing_vector with an external
force (src_mpi/equation.h:1189-1202).  Inputs: numpy default_rng(1), same state distribution as
make_flux_kat.py, plus states at rest and axis-aligned / negative-x velocities (atan2 branch cuts).
"""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from make_flux_kat import random_states  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    O.build(ref=True)
    M = ctypes.CDLL(os.path.join(os.path.dirname(os.path.dirname(HERE)), "oracle", "_ref", "libphys_reference_mpi.so"))
    M.phys_mpi_impl_name.restype = ctypes.c_char_p
    assert M.phys_mpi_impl_name().decode().startswith("reference:src_mpi")
    dp = ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(1)
    n = 600
    W = random_states(rng, n)
    F = rng.uniform(-3.0, 3.0, (n, 2))
    W[0] = (0.0, 0.0, 1.0, 2.5)            # at rest: atan2(0,0) = 0
    W[1] = (0.5, 0.0, 1.0, 2.5)
    W[2] = (-0.5, 0.0, 1.0, 2.5)           # theta = pi
    W[3] = (0.0, 0.5, 1.0, 2.5)
    W[4] = (0.0, -0.5, 1.0, 2.5)
    W[5] = (-0.5, -0.0, 1.0, 2.5)          # theta = -pi
    W[6] = (4.2, 0.0, 1.4, 8.8)
    F[0], F[1] = (0.0, -1.0), (0.0, 0.0)
    R, L, G = np.zeros((n, 16)), np.zeros((n, 16)), np.zeros((n, 4))
    for i in range(n):
        w = np.ascontiguousarray(W[i])
        f = np.ascontiguousarray(F[i])
        M.phys_mpi_eigen_stream(w.ctypes.data_as(dp), R[i].ctypes.data_as(dp), L[i].ctypes.data_as(dp))
        M.phys_mpi_ext_forcing(w.ctypes.data_as(dp), f.ctypes.data_as(dp), G[i].ctypes.data_as(dp))
    out = os.path.join(HERE, "mpi_extras_kat.npz")
    np.savez_compressed(out, W=W, F=F, R=R, L=L, G=G)
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
