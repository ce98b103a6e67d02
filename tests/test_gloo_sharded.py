"""CPU tier, N > 1 path: two OS processes (one per rank, torch.distributed `gloo` on 127.0.0.1),
each owning one cell-id shard of the mesh through the sharded C ABI (create_sharded).  The ranks
run the emulation build of the engine; the halo payload each stage travels over gloo exactly
where the CUDA build posts its NCCL send/recv group, and dt is the all-reduced minimum
(src_mpi/claw.cc:579).  Result must equal the single-rank run bit for bit and the oracle to
tolerance."""
import os
import socket
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch
    import torch.distributed as dist
    from dflo_b200 import abi
    import helpers as H

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    L = H.emu_lib()
    mesh_spec, bc, ic, prm, bval, nsteps = CASES[case]
    params, pair = abi.make_params(bc=bc, **prm)
    mesh = abi.Mesh(mesh_spec[0], mesh_spec[1], lib=L)
    flat = mesh.flatten(params, pair)
    eng = abi.Engine(flat, params, rank=rank, world=world, nccl_id=b"\0" * 128, lib=L, prefix="dflo_emu_")

    def exchange():
        """the halo: one send + one recv per neighbouring rank (NCCL group in the CUDA build)"""
        peers = [L.dflo_emu_peer_rank(eng.h, i) for i in range(L.dflo_emu_n_peers(eng.h))]
        reqs, bufs = [], {}
        for p in peers:
            n = L.dflo_emu_halo_send_count(eng.h, p)
            sb = np.zeros(n)
            L.dflo_emu_halo_get_send(eng.h, p, sb.ctypes.data_as(abi.c_double_p))
            reqs.append(dist.isend(torch.from_numpy(sb), p))
            rb = torch.zeros(L.dflo_emu_halo_recv_count(eng.h, p), dtype=torch.float64)
            bufs[p] = rb
            reqs.append(dist.irecv(rb, p))
        for r in reqs:
            r.wait()
        for p, rb in bufs.items():
            a = rb.numpy()
            L.dflo_emu_halo_put_recv(eng.h, p, a.ctypes.data_as(abi.c_double_p))

    # identical initial data on every rank (host-side IC), each rank keeps its own cells
    from oracle import oracle as O
    v, c, bl, bi = mesh.primitive()
    okw = {k: x for k, x in prm.items() if k != "time_step"}
    orc = O.Oracle(v, c, bl, bi, O.make_params(bc=bc, **okw))
    xq = orc.cell_qpoints()
    orc.set_initial_condition(ic(xq[..., 0], xq[..., 1]))
    u0 = orc.solution().copy()
    eng.set_solution(u0)
    exchange()
    if bval is not None:
        _, _, _, xqb = orc.bfaces()
        g = np.zeros((orc.n_bfaces, orc.nqf, 4))
        g[...] = np.asarray(bval)
        eng.set_boundary_values(g)
    if prm.get("limiter", "none") != "none":
        eng.limit_initial_condition()
        exchange()
    t = 0.0
    for _ in range(nsteps):
        dt = torch.tensor([eng.compute_dt(t)], dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MIN)          # C5: Utilities::MPI::min(global_dt)
        for rk in range(eng.n_rk):
            eng.rk_stage(rk, t, float(dt))
            exchange()
        eng.commit_step()
        t += float(dt)
    u = np.zeros_like(u0)
    eng.get_solution(out=u)
    b, e = eng.cell_range()
    np.save(os.path.join(out_dir, "u_rank%d.npy" % rank), u[b * eng.D:e * eng.D])
    np.save(os.path.join(out_dir, "range_rank%d.npy" % rank), np.array([b, e, t]))
    eng.close()
    dist.barrier()
    dist.destroy_process_group()


def _cases():
    import helpers as H
    return {
        "vortex_Q2_roe": (("isentropic_vortex", [8]), H.PERIODIC_BOX, H.ic_vortex, dict(basis="Qk", degree=2, flux="roe", cfl=0.5), None, 2),
        "sod_P2_hllc_tvb_pos": (("sod_tube", [24, 3]), H.SOD_BC, H.ic_sod,
                                dict(basis="Pk", degree=2, flux="hllc", limiter="TVB", char_lim=True, pos_lim=True, M=0.0,
                                     beta=2.0, cfl=0.5), (0.0, 0.0, 1.0, 2.5), 2),
    }


sys.path.insert(0, HERE)
CASES = _cases()


@pytest.mark.parametrize("case", sorted(CASES))
def test_two_ranks_over_gloo_match_single_rank_and_oracle(case, tmp_path):
    import torch.multiprocessing as mp
    import helpers as H
    H.build_emu()
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), case, str(tmp_path)), nprocs=world, join=True)
    mesh_spec, bc, ic, prm, bval, nsteps = CASES[case]

    def fresh():
        c = H.Case(mesh_spec, bc, ic, **prm)
        if bval is not None:
            c.set_boundary(values=bval)
        if prm.get("limiter", "none") != "none":
            c.limit_initial()
        return c
    # single rank advanced with the ENGINE's own dt, like the workers (Case.step feeds the oracle's
    # dt to both sides, and the two may differ in the last bit)
    one = fresh()
    e1, t1 = one.engine, 0.0
    for _ in range(nsteps):
        dt = e1.compute_dt(t1)
        for rk in range(e1.n_rk):
            e1.rk_stage(rk, t1, dt)
        e1.commit_step()
        t1 += dt
    u1 = one.solution()
    orc = fresh()
    for _ in range(nsteps):
        orc.step()
    D = e1.D
    got = np.zeros_like(u1)
    covered = 0
    for r in range(world):
        b, e, t = np.load(os.path.join(tmp_path, "range_rank%d.npy" % r))
        b, e = int(b), int(e)
        got[b * D:e * D] = np.load(os.path.join(tmp_path, "u_rank%d.npy" % r))
        covered += e - b
        assert t == t1
    assert covered == one.oracle.n_cells
    assert np.array_equal(got, u1)                       # sharding changes no bit
    uo = orc.oracle.solution()
    assert np.abs(got - uo).max() / max(1.0, np.abs(uo).max()) <= 1e-9
    one.close()
    orc.close()
