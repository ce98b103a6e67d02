#!/usr/bin/env python
"""All five BASELINE.json configurations through the input.prm front end (dflo_claw_*, the mirror of
ConservationLaw<2>::run), one GPU: MDoF-updates/s per config with device-resident state (CUDA events
around each dflo_b200_advance), the stage kernel alone, and the HBM-roofline fraction of the whole
stage (stage kernel + limiter).  bench.py stays the contract line for configs[1]; this script fills
the per-config table of DESIGN.md.   python scripts/bench_configs.py [--small] [--steps N]
Under torchrun (one rank per GPU) the same mesh is sharded over the ranks: strong scaling, e.g. of
configs[3] (double Mach reflection, ~1M cells), the case BASELINE.json names for 1 -> 8 GPUs."""
import argparse
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dflo_b200 import abi  # noqa: E402

PRM = os.path.join(ROOT, "tests", "golden", "prm")
FULL = {
    "cfg1": ("cfg1_isentropic_vortex_Q1_lxf.prm", "isentropic_vortex 32"),
    "cfg2": ("cfg2_isentropic_vortex_Q3_roe.prm", "isentropic_vortex 256"),
    "cfg3": ("cfg3_sod_P2_hllc_tvb_pos.prm", "sod_tube 1600 160"),
    "cfg4": ("cfg4_double_mach_Q2_hllc_tvb.prm", "double_mach 512"),
    "cfg5": ("cfg5_forward_step_Q3_kfvs_tvb_pos.prm", "forward_step 0.0025"),
}
SMALL = {"cfg1": "isentropic_vortex 32", "cfg2": "isentropic_vortex 64", "cfg3": "sod_tube 100 10", "cfg4": "double_mach 64",
         "cfg5": "forward_step 0.02"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--configs", default="cfg1,cfg2,cfg3,cfg4,cfg5")
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        # strong scaling of a configuration over the GPUs of one node: launched under torchrun
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = abi.load_library()
    L.dflo_claw_create.restype = ctypes.c_void_p
    L.dflo_claw_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
    L.dflo_claw_destroy.argtypes = [ctypes.c_void_p]
    L.dflo_claw_setup.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    L.dflo_claw_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int)]
    L.dflo_claw_engine.restype = ctypes.c_void_p
    L.dflo_claw_engine.argtypes = [ctypes.c_void_p]
    L.dflo_claw_n_dofs.argtypes = [ctypes.c_void_p]
    L.dflo_claw_params.restype = ctypes.POINTER(abi.Params)
    L.dflo_claw_params.argtypes = [ctypes.c_void_p]
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    out = []
    for key in args.configs.split(","):
        prm, mesh = FULL[key]
        if args.small:
            mesh = SMALL[key]
        h = L.dflo_claw_create(os.path.join(PRM, prm).encode(), mesh.encode(), None, abi.COMPAT["mpi"])
        assert h, L.dflo_host_last_error()
        idbuf = None
        if world > 1:
            idbuf = (ctypes.c_char * 128)()
            if rank == 0:
                assert L.dflo_b200_nccl_unique_id(idbuf) == 0
            tt = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
            dist.broadcast(tt, 0)
            idbuf = ctypes.create_string_buffer(bytes(tt.cpu().tolist()), 128)
        rc = L.dflo_claw_setup(h, local, rank, world, idbuf)
        assert rc == 0, L.dflo_host_last_error()
        ctx = ctypes.c_void_p(L.dflo_claw_engine(h))
        n_dofs = L.dflo_claw_n_dofs(h)
        p = L.dflo_claw_params(h).contents
        n_rk = L.dflo_b200_n_rk(ctx)
        D = L.dflo_b200_dofs_per_cell(ctx)
        t, done, ms = ctypes.c_double(0.0), ctypes.c_int(0), ctypes.c_float(0.0)
        for _ in range(5):
            assert L.dflo_claw_run(h, 1, 0, ctypes.byref(t), ctypes.byref(done)) == 0, L.dflo_host_last_error()
        total = 0.0
        for _ in range(args.steps):
            assert L.dflo_claw_run(h, 1, 0, ctypes.byref(t), ctypes.byref(done)) == 0, L.dflo_host_last_error()
            L.dflo_b200_last_advance_ms(ctx, ctypes.byref(ms))
            total += ms.value
        kms = ctypes.c_float(0.0)
        L.dflo_b200_time_stage_kernel(ctx, n_rk - 1, 10, 0, ctypes.byref(kms))
        if world > 1:
            v = torch.tensor([total, kms.value], dtype=torch.float64, device="cuda")
            dist.all_reduce(v, op=dist.ReduceOp.MAX)
            total, kms.value = float(v[0]), float(v[1])
        limited = p.limiter_type != 0 or p.pos_lim != 0
        bpu = (32.0 + 64.0 / D) if limited else (24.0 + 32.0 / D)
        ms_stage = total / args.steps / n_rk
        line = {"config": key, "n_gpus": world, "mesh": mesh, "cells": n_dofs // D, "dofs": n_dofs, "rk_stages": n_rk, "limited": bool(limited),
                "ms_per_step": total / args.steps, "mdof_per_s": n_dofs * n_rk * args.steps / (total * 1e-3) / 1e6,
                "stage_kernel_ms": kms.value, "ms_per_stage_all_kernels": ms_stage,
                "bytes_per_dof_update": bpu, "hbm_frac_whole_stage": n_dofs * bpu / (ms_stage * 1e-3) / 1e9 / peak, "t_end": t.value}
        if rank == 0:
            print(json.dumps(line), flush=True)
        out.append(line)
        L.dflo_claw_destroy(h)
    if world > 1:
        dist.destroy_process_group()
    return out


if __name__ == "__main__":
    main()
