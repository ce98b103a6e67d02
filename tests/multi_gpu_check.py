"""Multi-GPU parity check, run on a GPU box under torchrun (not collected by pytest):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tests/multi_gpu_check.py

Every rank owns one cell-id shard of the same global mesh (dflo_b200_create_sharded, NCCL halo
exchange per RK stage, all-reduced dt).  After a few steps of dflo_b200_advance the gathered
solution must equal (a) the single-GPU run of the same mesh bit for bit -- the sharded run
evaluates exactly the same kernels on the same inputs -- and (b) the CPU oracle to tolerance.
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
sys.path.insert(0, os.path.dirname(HERE))

from dflo_b200 import abi  # noqa: E402
from helpers import DMR_BC, PERIODIC_BOX, SOD_BC, STEP_BC, ic_dmr, ic_sod, ic_sod_moving, ic_sod_moving_wavy, ic_step, ic_vortex  # noqa: E402
from oracle import oracle as O  # noqa: E402

CASES = [
    ("vortex_Q3_roe", ("isentropic_vortex", [32]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=3, flux="roe", cfl=0.9), None, 4, 1e-12),
    ("vortex_Q1_lxf", ("isentropic_vortex", [32]), PERIODIC_BOX, ic_vortex, dict(basis="Qk", degree=1, flux="lxf", cfl=0.9), None, 4, 1e-12),
    ("sod_P2_hllc_tvb_pos", ("sod_tube", [100, 10]), SOD_BC, ic_sod,
     dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.9),
     (0.0, 0.0, 1.0, 2.5), 3, 1e-9),
    ("dmr_Q2_hllc_tvb", ("double_mach", [16]), DMR_BC, ic_dmr,
     dict(basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True, beta=1.0, M=100.0, cfl=0.9),
     (57.1576766498, -33.0, 8.0, 563.5), 3, 1e-9),
    ("sod_Q2_kxrcf_density", ("sod_tube", [100, 10]), {0: "outflow", 1: "outflow", 2: "inflow"}, ic_sod_moving,
     dict(basis="Qk", degree=2, flux="hllc", limiter="TVB", char_lim=True, beta=2.0, M=0.0, cfl=0.5, shock_indicator="density"),
     (0.3, 0.1, 1.0, 2.55), 3, 1e-9),
    ("sod_Q2_minmax_pos", ("sod_tube", [100, 10]), SOD_BC, ic_sod_moving_wavy,
     dict(basis="Qk", degree=2, flux="hllc", limiter="minmax", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.4),
     (0.3, 0.1, 1.0, 2.5), 3, 1e-9),
    # BASELINE configs[4] at test size: forward step, three lattice blocks, KFVS, TVB + positivity (2-layer halo, stand-alone exchange)
    ("step_Q3_kfvs_tvb_pos", ("forward_step", [0.05]), STEP_BC, ic_step,
     dict(basis="Qk", degree=3, flux="kfvs", limiter="TVB", char_lim=True, pos_lim=True, beta=2.0, M=0.0, cfl=0.5),
     (4.2, 0.0, 1.4, 8.8), 3, 1e-9),
    # mapping = q1 on skewed quadrilaterals with mixed cell orientations, and a mesh with hanging nodes: the mapped stage kernel
    ("q1_Q2_hllc_skew", ("rectangle_skew", [24, 24, -5, 5, -5, 5, 4, 2, 1, 3, 0.15, 1]), PERIODIC_BOX, ic_vortex,
     dict(basis="Qk", degree=2, flux="hllc", cfl=0.3, mapping="q1", compat="mpi"), None, 3, 1e-12),
    ("hanging_Q3_roe", ("rectangle_refined", [24, 24, -5, 5, -5, 5, 4, 2, 1, 3, 6, 18, 5, 19]), PERIODIC_BOX, ic_vortex,
     dict(basis="Qk", degree=3, flux="roe", cfl=0.3, compat="mpi"), None, 3, 1e-12),
]


def sharded_output_check(L, rank, world, local):
    """The standalone driver on a sharded setup (src_mpi/output.cc:34-86): every rank writes
    output/solution-NNNN.RRR.vtu with the cells it owns, rank 0 master_file.visit; the pieces put end to end must be
    the file the single-GPU driver writes for the same run, byte for byte in every data array."""
    import ctypes
    import shutil
    from test_host import PRM_DIR, _claw_api, _read_vtu
    _claw_api(L)
    L.dflo_claw_set_output.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    L.dflo_claw_set_output.restype = None
    L.dflo_claw_run.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, abi.c_double_p, ctypes.POINTER(ctypes.c_int)]
    L.dflo_claw_write_vtu.argtypes = [ctypes.c_void_p, ctypes.c_char_p]
    out = "/tmp/dflo_b200_sharded_output/"
    if rank == 0:
        shutil.rmtree(out, ignore_errors=True)
        os.makedirs(out)
    idbuf = (abi.ctypes.c_char * 128)()
    if rank == 0:
        assert L.dflo_b200_nccl_unique_id(idbuf) == 0
    t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    nccl_id = bytes(t.cpu().tolist())
    over = b"subsection output\n set iter step = 2\nend\n"
    prm = os.path.join(PRM_DIR, "cfg3_sod_P2_hllc_tvb_pos.prm").encode()

    def drive(world_, rank_, where):
        h = ctypes.c_void_p(L.dflo_claw_create(prm, b"sod_tube 100 10", over, abi.COMPAT["src"]))
        assert h, L.dflo_host_last_error()
        L.dflo_claw_set_output(h, where.encode())
        assert L.dflo_claw_setup(h, local, rank_, world_, nccl_id if world_ > 1 else None) == 0, L.dflo_host_last_error()
        tt, done = ctypes.c_double(0.0), ctypes.c_int(0)
        assert L.dflo_claw_run(h, 4, 0, ctypes.byref(tt), ctypes.byref(done)) == 0, L.dflo_host_last_error()
        L.dflo_claw_destroy(h)
        return tt.value

    t_sh = drive(world, rank, out)
    dist.barrier()
    good = True
    if rank == 0:
        os.makedirs(out + "single")
        t_1 = drive(1, 0, out + "single/")
        visit = open(out + "master_file.visit").read().split()
        good = visit[:2] == ["!NBLOCKS", str(world)] and len(visit) == 2 + 3 * world and t_1 == t_sh
        for n in range(3):                                    # it = 0, 2, 4
            whole = _read_vtu(out + "single/solution-%03d.vtu" % n)
            parts = [_read_vtu(out + "output/solution-%04d.%03d.vtu" % (n, r)) for r in range(world)]
            good = good and visit[2 + n * world:2 + (n + 1) * world] == ["output/solution-%04d.%03d.vtu" % (n, r) for r in range(world)]
            good = good and all(np.all(p["point"]["subdomain"] == r) for r, p in enumerate(parts))
            good = good and np.array_equal(np.concatenate([p["points"] for p in parts]), whole["points"])
            for name in whole["point_names"]:
                good = good and np.array_equal(np.concatenate([p["point"][name] for p in parts]), whole["point"][name])
        print("%-22s world %d: pieces of 3 outputs == single-GPU files, master_file.visit lists %d files  %s"
              % ("driver_output_sod_P2", world, len(visit) - 2, "OK" if good else "FAIL"), flush=True)
    dist.barrier()
    return good


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = abi.load_library()
    ok = True
    for name, mesh_spec, bc, ic, prm, bvals, nsteps, tol in CASES:
        idbuf = (abi.ctypes.c_char * 128)()
        if rank == 0:
            assert L.dflo_b200_nccl_unique_id(idbuf) == 0
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        nccl_id = bytes(t.cpu().tolist())

        params, pair = abi.make_params(bc=bc, **prm)
        mesh = abi.Mesh(mesh_spec[0], mesh_spec[1])
        flat = mesh.flatten(params, pair)
        v, c, bl, bi = mesh.primitive()
        orc = O.Oracle(v, c, bl, bi, O.make_params(bc=bc, **prm))
        xq = orc.cell_qpoints()
        orc.set_initial_condition(ic(xq[..., 0], xq[..., 1]))
        orc.compute_cell_average()
        u0 = orc.solution().copy()
        g = None
        if bvals is not None and orc.n_bfaces:
            g = np.zeros((orc.n_bfaces, orc.nqf, 4))
            g[...] = np.asarray(bvals)
            orc.set_bc_values(g)
        limited = prm.get("limiter", "none") != "none"

        def run(engine):
            engine.set_solution(u0)
            if g is not None:
                engine.set_boundary_values(g)
            if limited:
                engine.limit_initial_condition()
            tt = 0.0
            for _ in range(nsteps):
                tt, _ = engine.advance(1, elapsed=tt)
            return tt

        sharded = abi.Engine(flat, params, device=local, rank=rank, world=world, nccl_id=nccl_id)
        t_sh = run(sharded)
        u = np.zeros(orc.n_cells * orc.D)
        sharded.get_solution(out=u)          # fills this rank's owned range only
        b, e = sharded.cell_range()
        mask = np.zeros_like(u)
        mask[b * orc.D:e * orc.D] = 1.0
        ut = torch.from_numpy(u * mask).cuda()
        dist.all_reduce(ut)
        u_sh = ut.cpu().numpy()
        sharded.close()
        if rank == 0:
            single = abi.Engine(flat, params, device=local)
            t_1 = run(single)
            u_1 = single.get_solution()
            single.close()
            if limited:
                orc.apply_limiter()
                orc.commit_step()
            tt = 0.0
            for _ in range(nsteps):
                dt = orc.compute_dt(tt)
                for rk in range(orc.n_rk):
                    orc.rk_stage(rk, dt)
                orc.commit_step()
                tt += dt
            uo = orc.solution()
            d1 = np.abs(u_sh - u_1).max()
            do = np.abs(u_sh - uo).max() / max(1.0, np.abs(uo).max())
            good = d1 == 0.0 and do <= tol and abs(t_sh - t_1) == 0.0
            ok = ok and good
            print("%-22s world %d: |sharded - single| = %.3e (t %.3e)  rel err vs oracle = %.3e  %s"
                  % (name, world, d1, abs(t_sh - t_1), do, "OK" if good else "FAIL"), flush=True)
        dist.barrier()
    ok = sharded_output_check(L, rank, world, local) and ok
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
