#!/usr/bin/env python
"""bench.py -- million DoF-updates/s of the explicit RK stage (BASELINE.json metric), L-infinity against the CPU oracle.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     CPU reference arm (oracle on host cores)

One "step" = one full time step of the hot path (compute_time_step + n_rk RK stages: residual, M^-1, RK combine, cell
average, configured limiters) over the whole mesh; one DoF-update = one scalar unknown advanced through one RK stage.

Headline (`value`): BASELINE.json configs[1] (isentropic vortex, Q3, 256x256 Cartesian, periodic, Roe, SSP-RK3).  For
N > 1 the per-GPU work is kept (weak scaling): the periodic box grows to 256 x 256N cells sharded by cell id, one halo
exchange per stage stored straight into the peers' memory over NVLink (fused into the stage kernel).

The same JSON line also carries
  linf_vs_ref   L-infinity of the engine against the oracle on the SAME 256x256 mesh after 1 RHS / 1 step / 20 steps
                (the oracle run that also yields `cpu_baseline`), N = 1;
  configs       N = 1: the other BASELINE configurations at full size (cfg1, cfg3 1600x160, cfg4 2048x512, cfg5
                cl=0.0025) -- MDoF/s, ms/step, whole-stage HBM roofline fraction, stage kernel alone -- each with
                linf_vs_ref at the three horizons on a >= 64k-cell mesh of the same case (size stated per entry);
  strong        N > 1: cfg4 (double Mach reflection, 1 049 088 cells) and cfg5 (forward step, 403 200 cells) sharded
                over the N GPUs -- MDoF/s, the single-GPU figure of the same box measured on rank 0, efficiency,
                linf_vs_single (sharded result against the single-GPU result: must be 0);
  linf_vs_single  N > 1: the weak-scaling workload itself, sharded against single-GPU after 3 steps (must be 0).

Timing: every step is one dflo_b200_advance() (CUDA graph replay) timed with CUDA events on the ctx stream; L2
(126 MB) is flushed before every timed step by rewriting a 256 MB buffer; per-step event times summed, max over ranks.
`e2e` = the same step through the C ABI with HOST buffers: set_solution (pinned host -> device) + advance +
get_solution (device -> pinned host) inside the timed region.

The oracle (oracle/, tests/helpers.py) is imported only by the checker legs (`linf_vs_ref`, `cpu_baseline`,
--impl reference); nothing in the timed GPU regions touches it.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "million DoF-updates/sec (explicit RK stage)"
UNIT = "MDoF-updates/s"
PERIODIC = {1: ("periodic", 3), 3: ("periodic", 1), 2: ("periodic", 4), 4: ("periodic", 2)}
WORKLOADS = {
    # name: (description, basis, degree, flux, cells per side per GPU)
    "cfg2": ("isentropic_vortex, Q3, 256x256 Cartesian, Roe flux, explicit RK3, periodic", "Qk", 3, "roe", 256),
    "cfg1": ("isentropic_vortex, Q1, 32x32 Cartesian, LxF flux, explicit RK2, periodic", "Qk", 1, "lxf", 32),
}
PRM_DIR = os.path.join(ROOT, "tests", "golden", "prm")
# the other BASELINE configurations: deck, full-size mesh, mesh of the parity leg (generator arguments), description
CONFIGS = {
    "cfg1": ("cfg1_isentropic_vortex_Q1_lxf.prm", "isentropic_vortex 32", [32], "isentropic_vortex, Q1, 32x32 Cartesian, LxF, RK2 (launch-latency regime)"),
    "cfg3": ("cfg3_sod_P2_hllc_tvb_pos.prm", "sod_tube 1600 160", [800, 80], "sod_shock_tube, P2 Legendre, 1600x160, HLLC, TVB + positivity"),
    "cfg4": ("cfg4_double_mach_Q2_hllc_tvb.prm", "double_mach 512", [128], "double_mach_reflection, Q2, 2048x512 (+ inflow block), HLLC, TVB"),
    "cfg5": ("cfg5_forward_step_Q3_kfvs_tvb_pos.prm", "forward_step 0.0025", [0.00625], "forward_step, Q3, 3-block mesh cl=0.0025, KFVS, TVB + positivity"),
}
FLUSH_BYTES = 256 * 1024 * 1024


def isentropic_vortex(x, y):
    """src_mpi/ic.cc:44-61: vortex advected with M_inf = 0.5 along x (exact solution)."""
    g, beta = 1.4, 5.0
    a1 = 0.5 * beta / np.pi
    a2 = 0.5 * (g - 1.0) * a1 * a1
    r2 = x * x + y * y
    rho = (1.0 - a2 * np.exp(1.0 - r2)) ** (1.0 / (g - 1.0))
    vx = 0.5 - a1 * y * np.exp(0.5 * (1.0 - r2))
    vy = a1 * x * np.exp(0.5 * (1.0 - r2))
    pre = rho ** g / g
    return np.stack([rho * vx, rho * vy, rho, pre / (g - 1.0) + 0.5 * rho * (vx * vx + vy * vy)], axis=-1)


def gauss01(n):
    x, w = np.polynomial.legendre.leggauss(n)
    return 0.5 * (x + 1.0), 0.5 * w


def initial_dofs(nx, ny, x0, x1, y0, y1, k):
    """Qk interpolation of the vortex at the Gauss nodes, reference DoF layout [cell][comp][node]."""
    gx, _ = gauss01(k + 1)
    hx, hy = (x1 - x0) / nx, (y1 - y0) / ny
    xs = x0 + hx * (np.arange(nx)[:, None] + gx[None, :])          # [nx][a]
    ys = y0 + hy * (np.arange(ny)[:, None] + gx[None, :])          # [ny][b]
    X = np.broadcast_to(xs[None, :, None, :], (ny, nx, k + 1, k + 1))   # [j][i][b][a]
    Y = np.broadcast_to(ys[:, None, :, None], (ny, nx, k + 1, k + 1))
    f = isentropic_vortex(X, Y)                                     # [j][i][b][a][c]
    u = np.transpose(f, (0, 1, 4, 2, 3)).reshape(ny * nx, 4, (k + 1) ** 2)
    return np.ascontiguousarray(u).reshape(-1)


def sample_clocks(stop, out, ready=None):
    """SM clock + throttle reasons while the timed region runs (B200_PROFILING.md clocks line).  NVML is
    polled back to back (the timed region of this workload lasts milliseconds, too short for
    nvidia-smi's loop mode); `ready` is set once NVML is initialised so that the caller starts the
    timed region only when the sampler is live.  nvidia-smi -lms is the fallback when the NVML binding
    is missing."""
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(dev)
        mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        names = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        if ready is not None:
            ready.set()
        last = False
        while not last:
            last = stop.is_set()      # one more sample after the stop request: a slow NVML never leaves the list empty
            sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
            try:
                r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
            except Exception:
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
            out.append("%d, %d, %s" % (sm, mx, ", ".join("Active" if r & bit else "Not Active" for _, bit in names)))
            time.sleep(0.0002)
        return
    except Exception:
        pass
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "--query-gpu=" + q, "--format=csv,noheader,nounits", "-i", str(dev), "-lms", "20"],
                             stdout=subprocess.PIPE, text=True)
    except Exception:
        if ready is not None:
            ready.set()
        return
    while not stop.is_set():
        line = p.stdout.readline()
        if not line:
            break
        out.append(line.strip())
        if ready is not None:
            ready.set()
    p.terminate()


def clocks_summary(lines):
    sm, mx, reasons = [], 0, set()
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    for ln in lines:
        f = [s.strip() for s in ln.split(",")]
        try:
            sm.append(float(f[0]))
            mx = max(mx, float(f[1]))
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        except Exception:
            continue
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
            "samples": len(sm)}


def measured_peak_hbm():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p))["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(kernel):
    """dram read + write bytes of one launch of the stage kernel from the ncu --set full capture of THIS build
    (profiles/stage_kernel_traffic.json carries the source hash of the kernel files it was taken on), else None."""
    p = os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")
    try:
        d = json.load(open(p))
        if d.get("kernel") == kernel and d.get("source_sha16") == kernel_source_sha16():
            return d["dram_bytes_read"] + d["dram_bytes_write"]
    except Exception:
        pass
    return None


def kernel_source_sha16():
    import hashlib
    h = hashlib.sha256()
    for f in ("row_kernel.cuh", "kernels.cuh", "euler.cuh", "row_desc.h", "p2p_halo.cuh"):
        try:
            h.update(open(os.path.join(ROOT, "dflo_b200", "csrc", f), "rb").read())
        except OSError:
            pass
    return h.hexdigest()[:16]


def workload_config(workload, world):
    """The `config` object both arms print (same keys, same values: the reference arm times a sample of THIS workload)."""
    desc, basis, k, _, npg = WORKLOADS[workload]
    nx, ny = npg, npg * world
    D = 4 * (k + 1) ** 2
    return {"workload": desc, "cells": nx * ny, "dofs": nx * ny * D, "rk_stages": 1 if k == 0 else 2 if k == 1 else 3,
            "mesh": "%dx%d" % (nx, ny), "n_gpus": world}


# ---------------------------------------------------------------------------------------------
# checker legs: the CPU oracle (test infrastructure) beside the engine.  Never inside a timed GPU region.
# ---------------------------------------------------------------------------------------------
def _oracle_variant():
    from oracle import oracle as O
    variant = "refphys" if os.path.exists(O.lib_path("refphys")) else "restated"
    if variant == "restated":
        O.build()
    return variant


def parity_leg(key, size, nsteps, threads):
    """Engine (C ABI, cuda:0) against the oracle on the same mesh: L-infinity after 1 RHS / 1 step / nsteps steps, and
    the oracle's own throughput on that run (the CPU baseline of this configuration)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    from oracle import oracle as O
    variant = _oracle_variant()
    c, bc_fn = helpers.baseline_case(key, size, backend="cuda", oracle_variant=variant, oracle_threads=threads)
    h = helpers.parity_horizons(c, nsteps=nsteps, bc_fn=bc_fn)
    o = c.oracle
    shocked = c.prm_kw.get("limiter", "none") != "none"
    updates = o.n_cells * o.D * o.n_rk * nsteps
    out = {
        "rhs": h["rhs"], "step1": h["step1"], "step%d" % nsteps: h["step%d" % nsteps], "limiter_flips": h["limiter_flips"],
        "norm": "max |u_gpu - u_ref| / max(1, max |u_ref|), conserved variables; rhs likewise",
        "mesh": "%s %s: %d cells, %d DoF" % (helpers.BASELINE_CASES[key][0], " ".join(str(s) for s in size), o.n_cells, o.n_cells * o.D),
        "tolerance": {"rhs": 1e-13 * float(np.sqrt(o.D)), "step1": 1e-9 if shocked else 1e-12, "step%d" % nsteps: 1e-9 if shocked else 1e-11},
        "oracle": "restated assembly (oracle/dflo_oracle.cc), physics %s" % O.load(variant).phys_impl_name().decode(),
    }
    tol = out["tolerance"]
    out["within_tolerance"] = bool(out["rhs"] <= tol["rhs"] and out["step1"] <= tol["step1"] and out["step%d" % nsteps] <= tol["step%d" % nsteps])
    cpu = {"value": updates / c.oracle_seconds / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
           "physics": O.load(variant).phys_impl_name().decode(),
           "sample": "%d steps of %s, %.1f s in the oracle's dt + stage calls" % (nsteps, out["mesh"], c.oracle_seconds),
           "sample_cells": o.n_cells, "sample_steps": nsteps, "wall_s": c.oracle_seconds}
    c.close()
    return out, cpu


def cpu_baseline(workload, budget_s=15.0, threads=None):
    """The oracle (kind 'port': restated assembly; physics = the reference's own equation.h object
    code when oracle/_ref was built) timed on this box's host cores on a bounded sample."""
    from oracle import oracle as O
    desc, basis, k, flux, _ = WORKLOADS[workload]
    threads = threads or os.cpu_count() or 1
    variant = _oracle_variant()

    def make(n):
        p = O.make_params(basis=basis, degree=k, flux=flux, bc=PERIODIC, cfl=0.9, n_threads=threads)
        o = O.Oracle(*O.rect_mesh(n, n, -5, 5, -5, 5), p, variant=variant)
        xq = o.cell_qpoints()
        o.set_initial_condition(isentropic_vortex(xq[..., 0], xq[..., 1]))
        o.compute_cell_average()
        return o
    n = 32
    o = make(n)
    t0 = time.perf_counter()
    o.run_steps(1)
    t1 = time.perf_counter() - t0
    per_cell = t1 / (n * n)
    # size the sample: ~budget_s of CPU work in a handful of steps
    n = int(min(256, max(32, np.sqrt(budget_s / 4.0 / per_cell))))
    o = make(n)
    steps = max(1, int(budget_s / (per_cell * n * n)))
    t0 = time.perf_counter()
    o.run_steps(steps)
    dt = time.perf_counter() - t0
    updates = n * n * o.D * o.n_rk * steps
    return {"value": updates / dt / 1e6, "unit": UNIT, "cores": threads, "kind": "port",
            "physics": O.load(variant).phys_impl_name().decode(),
            "sample": "%d steps of %s on a %dx%d sample mesh, %.1f s wall" % (steps, desc.split(",")[0] + " " + basis[0] + str(k)
                                                                              + " " + flux, n, n, dt),
            "sample_cells": n * n, "sample_steps": steps, "wall_s": dt}


def b0_single_thread(workload="cfg1", budget_s=3.0):
    """the serial CPU port (one thread) on the reference's own CPU-runnable case at its
    shipped size -- configs[0], 32 x 32 cells."""
    from oracle import oracle as O
    desc, basis, k, flux, npg = WORKLOADS[workload]
    variant = _oracle_variant()
    p = O.make_params(basis=basis, degree=k, flux=flux, bc=PERIODIC, cfl=0.9, n_threads=1)
    o = O.Oracle(*O.rect_mesh(npg, npg, -5, 5, -5, 5), p, variant=variant)
    xq = o.cell_qpoints()
    o.set_initial_condition(isentropic_vortex(xq[..., 0], xq[..., 1]))
    o.compute_cell_average()
    t0 = time.perf_counter()
    o.run_steps(2)
    per_step = (time.perf_counter() - t0) / 2
    steps = max(2, int(budget_s / max(per_step, 1e-6)))
    t0 = time.perf_counter()
    o.run_steps(steps)
    dt = time.perf_counter() - t0
    return {"value": npg * npg * o.D * o.n_rk * steps / dt / 1e6, "unit": UNIT, "cores": 1, "kind": "port",
            "physics": O.load(variant).phys_impl_name().decode(),
            "sample": "%d steps of %s at its shipped size (%dx%d cells), %.1f s wall, one thread" % (steps, desc.split(",")[0], npg, npg, dt),
            "sample_cells": npg * npg, "sample_steps": steps, "wall_s": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    from oracle import oracle as O
    _, basis, k, flux, _ = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    variant = _oracle_variant()
    # each "step" = one time step on a bounded sample mesh sized for ~1 s
    n = 32
    p = O.make_params(basis=basis, degree=k, flux=flux, bc=PERIODIC, cfl=0.9, n_threads=threads)

    def make(n):
        o = O.Oracle(*O.rect_mesh(n, n, -5, 5, -5, 5), p, variant=variant)
        xq = o.cell_qpoints()
        o.set_initial_condition(isentropic_vortex(xq[..., 0], xq[..., 1]))
        o.compute_cell_average()
        return o
    o = make(n)
    t0 = time.perf_counter()
    o.run_steps(1)
    per_cell = (time.perf_counter() - t0) / (n * n)
    total = max(1, args.steps + args.warmup)
    n = int(min(256, max(32, np.sqrt(min(2.0, 150.0 / total) / per_cell))))
    o = make(n)
    o.run_steps(args.warmup)
    t0 = time.perf_counter()
    o.run_steps(args.steps)
    dt = time.perf_counter() - t0
    updates = n * n * o.D * o.n_rk * args.steps
    val = updates / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.workload, max(world, args.gpus)),
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                         "physics": O.load(variant).phys_impl_name().decode(),
                         "sample": "each step = one time step on %dx%d cells of the same case (bounded CPU sample), %d timed steps, "
                                   "OpenMP workers + serial copier (WorkStream-like), %d host threads" % (n, n, args.steps, threads)},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "dflo itself needs deal.II (absent): this arm times the CPU restatement of its explicit path "
                "(oracle/), physics from the reference's own equation.h when oracle/_ref is present",
    }
    print(json.dumps(line))


class near_gpu_cores:
    """Context manager: bind the calling process to the CPU cores NVML lists as local to GPU `index` (the NUMA node its PCIe
    root hangs on), so that memory first touched inside -- the pinned staging buffers -- lands there; restores the previous
    affinity on exit.  Does nothing when NVML or the affinity call is unavailable, or the mask is empty in this cpuset."""

    def __init__(self, index):
        self.index, self.old = index, None

    def __enter__(self):
        try:
            if os.environ.get("DFLO_BENCH_NUMA", "1") == "0":
                return self
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(visible.split(",")[self.index]) if visible and visible.split(",")[self.index].isdigit() else self.index
            h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
            cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
            allowed = os.sched_getaffinity(0)
            if cpus & allowed and (cpus & allowed) != allowed:
                self.old = allowed
                os.sched_setaffinity(0, cpus & allowed)
        except Exception:                                    # noqa: BLE001 -- a placement hint, never an error
            self.old = None
        return self

    def __exit__(self, *exc):
        if self.old is not None:
            try:
                os.sched_setaffinity(0, self.old)
            except OSError:
                pass
        return False


def e2e_entry(updates_per_step, steps, serial_s, pipe_s, n_ctx, owned_dof):
    serial = updates_per_step * steps / serial_s / 1e6
    what = "set_solution(pinned host) + advance(1 step) + get_solution(pinned host) per step"
    e = {"value": serial, "unit": UNIT, "h2d_bytes_per_step": int(owned_dof * 8), "d2h_bytes_per_step": int(owned_dof * 8),
         "steps": steps, "in_flight": 1, "what": what + ", one context, each step's input is the previous step's output"}
    if pipe_s:
        e.update({"value": updates_per_step * steps * n_ctx / pipe_s / 1e6, "steps": steps * n_ctx, "in_flight": n_ctx,
                  "value_one_context_serial": serial,
                  "what": what + "; %d independent batches in flight (one context, host thread and pinned buffer each), "
                                 "so one batch's copy in overlaps the other's copy out" % n_ctx})
    return e


# ---------------------------------------------------------------------------------------------
# the input.prm front end (dflo_claw_*, mirror of ConservationLaw<2>::run) for the other configurations
# ---------------------------------------------------------------------------------------------
class Claw:
    def __init__(self, L, key, mesh=None):
        from dflo_b200 import abi
        self.L, self.abi = L, abi
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
        L.dflo_claw_get_solution.argtypes = [ctypes.c_void_p, abi.c_double_p, ctypes.c_size_t]
        prm, full, _, self.desc = CONFIGS[key]
        self.mesh = mesh or full
        self.h = ctypes.c_void_p(L.dflo_claw_create(os.path.join(PRM_DIR, prm).encode(), self.mesh.encode(), None, abi.COMPAT["mpi"]))
        if not self.h:
            raise RuntimeError(L.dflo_host_last_error().decode())
        self.t, self.done, self.ms = ctypes.c_double(0.0), ctypes.c_int(0), ctypes.c_float(0.0)

    def setup(self, device, rank=0, world=1, nccl_id=None):
        idbuf = ctypes.create_string_buffer(bytes(nccl_id), 128) if nccl_id is not None else None
        if self.L.dflo_claw_setup(self.h, device, rank, world, idbuf) != 0:
            raise RuntimeError(self.L.dflo_host_last_error().decode())
        self.ctx = ctypes.c_void_p(self.L.dflo_claw_engine(self.h))
        self.n_dofs = self.L.dflo_claw_n_dofs(self.h)
        self.n_rk = self.L.dflo_b200_n_rk(self.ctx)
        self.D = self.L.dflo_b200_dofs_per_cell(self.ctx)
        p = self.L.dflo_claw_params(self.h).contents
        self.limited = p.limiter_type != 0 or p.pos_lim != 0
        self.basis, self.degree, self.flux = p.basis, p.degree, p.flux_type

    def step(self):
        """one time step; returns its device time in ms (CUDA events on the ctx stream)"""
        if self.L.dflo_claw_run(self.h, 1, 0, ctypes.byref(self.t), ctypes.byref(self.done)) != 0:
            raise RuntimeError(self.L.dflo_host_last_error().decode())
        self.L.dflo_b200_last_advance_ms(self.ctx, ctypes.byref(self.ms))
        return self.ms.value

    def stage_kernel_ms(self, reps=10, flush=FLUSH_BYTES):
        kms = ctypes.c_float(0.0)
        self.L.dflo_b200_time_stage_kernel(self.ctx, self.n_rk - 1, reps, flush, ctypes.byref(kms))
        return kms.value

    def solution(self):
        u = np.zeros(self.n_dofs)
        if self.L.dflo_claw_get_solution(self.h, u.ctypes.data_as(self.abi.c_double_p), u.size) != 0:
            raise RuntimeError(self.L.dflo_host_last_error().decode())
        return u

    def launch_count(self):
        return self.L.dflo_b200_launch_count(self.ctx)

    def close(self):
        if self.h:
            self.L.dflo_claw_destroy(self.h)
            self.h = None


def timed_claw_steps(cl, steps, warmup, flush, barrier):
    for _ in range(warmup):
        cl.step()
    barrier()
    total = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
        barrier()
        total += cl.step()
    return total


def config_entry(L, key, steps, warmup, flush, peak, torch):
    """Full-size single-GPU throughput of one BASELINE configuration through the input.prm front end."""
    cl = Claw(L, key)
    cl.setup(0)
    l0 = cl.launch_count()
    total = timed_claw_steps(cl, steps, warmup, flush, torch.cuda.synchronize)
    launches = cl.launch_count() - l0
    kms = cl.stage_kernel_ms()
    bpu = (32.0 + 64.0 / cl.D) if cl.limited else (24.0 + 32.0 / cl.D)
    ms_step = total / steps
    ach = cl.n_dofs * bpu / (ms_step / cl.n_rk * 1e-3) / 1e9
    bpu_k = 24.0 + 32.0 / cl.D
    e = {"config": key, "workload": cl.desc, "mesh": cl.mesh, "cells": cl.n_dofs // cl.D, "dofs": cl.n_dofs, "rk_stages": cl.n_rk,
         "limited": bool(cl.limited), "steps": steps, "ms_per_step": ms_step,
         "mdof_per_s": cl.n_dofs * cl.n_rk * steps / (total * 1e-3) / 1e6, "gpu_launches": int(launches),
         "l2": "flushed before every timed step (256 MB rewrite)",
         "roofline": {"bound": "hbm", "scope": "whole stage (stage kernel + limiter kernels), ms_per_step / rk_stages",
                      "bytes_per_dof_update": bpu, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak},
         "stage_kernel": {"ms": kms, "bytes_per_dof_update": bpu_k,
                          "frac": cl.n_dofs * bpu_k / (kms * 1e-3) / 1e9 / peak}}
    cl.close()
    return e


def q1_entry(steps, warmup, flush, peak, torch, parity_steps, threads, with_oracle):
    """mapping = q1: the cfg2 case on 256 x 256 smoothly skewed quadrilaterals through the mapped
    stage kernel -- throughput, roofline fraction (same 24 + 32/D bytes per DoF-update; the 64 B of vertices per cell
    are charged nothing), and parity at the three horizons on a 128 x 128 mesh of the same kind."""
    from dflo_b200 import abi
    n, k = 256, 3
    params, pair = abi.make_params(basis="Qk", degree=k, flux="roe", bc=PERIODIC, cfl=0.45, compat="mpi", mapping="q1")
    mesh = abi.Mesh("rectangle_skew", [n, n, -5, 5, -5, 5, 4, 2, 1, 3, 0.15, 0])
    flat = mesh.flatten(params, pair)
    eng = abi.Engine(flat, params)
    D = eng.D
    n_dof = n * n * D
    v, c, _, _ = mesh.primitive()
    gx, _ = gauss01(k + 1)
    xi, eta = np.meshgrid(gx, gx, indexing="xy")
    xi, eta = xi.reshape(-1), eta.reshape(-1)
    N = np.stack([(1 - xi) * (1 - eta), xi * (1 - eta), (1 - xi) * eta, xi * eta], axis=0)     # [4][nq]
    X = np.einsum("cv,vq->cq", v[c][:, :, 0], N)
    Y = np.einsum("cv,vq->cq", v[c][:, :, 1], N)
    u0 = np.ascontiguousarray(np.transpose(isentropic_vortex(X, Y), (0, 2, 1))).reshape(-1)     # [cell][comp][node]
    eng.set_solution(u0)
    t = 0.0
    for _ in range(warmup):
        t, _ = eng.advance(1, elapsed=t)
    l0 = eng.launch_count()
    total = 0.0
    for _ in range(steps):
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize()
        t, _ = eng.advance(1, elapsed=t)
        total += eng.last_advance_ms()
    launches = eng.launch_count() - l0
    kms = eng.time_stage_kernel(rk=1, reps=10, flush_bytes=FLUSH_BYTES)
    eng.poll_error()
    eng.close()
    bpu = 24.0 + 32.0 / D
    ms_step = total / steps
    e = {"config": "q1", "workload": "isentropic_vortex, Q3, 256x256 smoothly skewed quadrilaterals, mapping = q1, Roe, RK3, periodic",
         "mesh": "rectangle_skew 256 256 (amplitude 0.15)", "cells": n * n, "dofs": n_dof, "rk_stages": eng.n_rk, "limited": False,
         "steps": steps, "ms_per_step": ms_step, "mdof_per_s": n_dof * eng.n_rk * steps / (total * 1e-3) / 1e6,
         "gpu_launches": int(launches), "l2": "flushed before every timed step (256 MB rewrite)",
         "roofline": {"bound": "hbm", "scope": "whole stage, ms_per_step / rk_stages", "bytes_per_dof_update": bpu,
                      "achieved": n_dof * bpu / (ms_step / eng.n_rk * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                      "frac": n_dof * bpu / (ms_step / eng.n_rk * 1e-3) / 1e9 / peak},
         "stage_kernel": {"name": "MappedStageKernel<4,roe>", "ms": kms, "bytes_per_dof_update": bpu, "frac": n_dof * bpu / (kms * 1e-3) / 1e9 / peak}}
    if with_oracle:
        e["linf_vs_ref"], e["cpu_baseline"] = parity_leg("q1", [128, 128], parity_steps, threads)
    return e


def strong_entry(L, key, steps, warmup, flush, peak, torch, dist, rank, world, local_rank, new_nccl_id):
    """One BASELINE configuration at full size sharded over the N GPUs (strong scaling), the single-GPU run of the
    same deck on rank 0 beside it, and the sharded result against the single-GPU result (bit for bit: 0)."""
    def barrier():
        torch.cuda.synchronize()
        dist.barrier()
    cl = Claw(L, key)
    cl.setup(local_rank, rank, world, new_nccl_id())
    total = timed_claw_steps(cl, steps, warmup, flush, barrier)
    v = torch.tensor([total], dtype=torch.float64, device="cuda")
    dist.all_reduce(v, op=dist.ReduceOp.MAX)
    total = float(v[0])
    u = torch.from_numpy(cl.solution()).cuda()          # own range filled, zeros elsewhere
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    n_dofs, n_rk, D, desc, mesh, t_sh = cl.n_dofs, cl.n_rk, cl.D, cl.desc, cl.mesh, cl.t.value
    cl.close()
    e = None
    if rank == 0:
        one = Claw(L, key)
        one.setup(local_rank)
        total1 = timed_claw_steps(one, steps, warmup, flush, torch.cuda.synchronize)
        u1 = one.solution()
        linf = float(np.abs(u.cpu().numpy() - u1).max())
        v1 = n_dofs * n_rk * steps / (total1 * 1e-3) / 1e6
        vN = n_dofs * n_rk * steps / (total * 1e-3) / 1e6
        e = {"config": key, "workload": desc, "mesh": mesh, "cells": n_dofs // D, "dofs": n_dofs, "rk_stages": n_rk, "n_gpus": world,
             "scaling": "strong", "steps": steps, "ms_per_step": total / steps, "mdof_per_s": vN,
             "mdof_per_s_1gpu_same_box": v1, "ms_per_step_1gpu": total1 / steps, "speedup": vN / v1, "efficiency": vN / v1 / world,
             "linf_vs_single": linf, "linf_steps": steps + warmup, "t_end": t_sh, "t_end_1gpu": one.t.value,
             "l2": "flushed before every timed step (256 MB rewrite)",
             "parallelism": "cells sharded by id over %d GPUs, 2-layer halo, one exchange per stage over NVLink peer memory" % world}
        one.close()
    dist.barrier()
    return e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip every oracle leg (cpu_baseline, linf_vs_ref)")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--configs", default=os.environ.get("DFLO_BENCH_CONFIGS", "cfg1,cfg3,cfg4,cfg5"),
                    help="other BASELINE configurations measured beside the headline (N = 1), '' = none")
    ap.add_argument("--strong", default=os.environ.get("DFLO_BENCH_STRONG", "cfg4,cfg5"),
                    help="configurations sharded over the N GPUs at full size (N > 1), '' = none")
    ap.add_argument("--next-rows", default=os.environ.get("DFLO_BENCH_NEXT", "q1"),
                    help="SURVEY 8(f) rows measured beside the BASELINE configurations (N = 1): q1 = mapping q1 on skewed quadrilaterals")
    ap.add_argument("--config-steps", type=int, default=20)
    ap.add_argument("--parity-steps", type=int, default=20)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    from dflo_b200 import abi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    L = abi.load_library()
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def new_nccl_id():
        idbuf = (abi.ctypes.c_char * 128)()
        if rank == 0:
            assert L.dflo_b200_nccl_unique_id(idbuf) == 0
        t = torch.tensor(list(bytes(idbuf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    nccl_id = new_nccl_id() if world > 1 else None

    desc, basis, k, flux, npg = WORKLOADS[args.workload]
    # weak scaling: the box grows in y, so the contiguous cell-id range of a rank is a compact
    # npg x npg square (cells are numbered x fastest) with 2 x npg interface cells
    nx, ny = npg, npg * world
    x0, x1, y0, y1 = -5.0, 5.0, -5.0 * world, 5.0 * world
    fixed_dt = float(os.environ.get("DFLO_BENCH_FIXED_DT", "0"))     # developer experiments: no time-step reduction at all
    params, pair = abi.make_params(basis=basis, degree=k, flux=flux, bc=PERIODIC, cfl=0.0 if fixed_dt > 0 else 0.9,
                                   time_step=fixed_dt if fixed_dt > 0 else -1.0, compat="mpi")
    mesh = abi.Mesh("rectangle", [nx, ny, x0, x1, y0, y1, 4, 2, 1, 3])
    flat = mesh.flatten(params, pair)
    eng = abi.Engine(flat, params, device=local_rank, rank=rank, world=world, nccl_id=nccl_id)
    D, n_rk = eng.D, eng.n_rk
    n_dof = nx * ny * D
    # the pinned host buffer of the e2e leg on the NUMA node of this rank's GPU (first touch with the rank bound to the
    # cores NVML reports as near the device; the binding is dropped again: the CPU legs use every core)
    with near_gpu_cores(local_rank):
        u_host = torch.empty(n_dof, dtype=torch.float64, pin_memory=True)
        u_host.numpy()[:] = 0.0
    u_host.numpy()[:] = initial_dofs(nx, ny, x0, x1, y0, y1, k)
    u_np = u_host.numpy()
    eng.set_solution(u_np)

    flush = None if args.no_flush else torch.empty(FLUSH_BYTES, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_step(t):
        if flush is not None:
            flush.zero_()
            torch.cuda.synchronize()
        if world > 1:
            dist.barrier()            # untimed: the ranks launch the step within microseconds of each other
        t, _ = eng.advance(1, elapsed=t)
        return t, eng.last_advance_ms()

    t = 0.0
    for _ in range(args.warmup):
        t, _ = timed_step(t)
    clk_lines, stop, clk_ready = [], threading.Event(), threading.Event()
    th = threading.Thread(target=sample_clocks, args=(stop, clk_lines, clk_ready), daemon=True)
    if rank == 0:
        th.start()
        clk_ready.wait(10.0)
        del clk_lines[:]          # keep only samples taken inside the timed region
    barrier()
    launches0 = eng.launch_count()
    wall0 = time.perf_counter()
    ms_total = 0.0
    for _ in range(args.steps):
        t, ms = timed_step(t)
        ms_total += ms
    barrier()
    wall = time.perf_counter() - wall0
    stop.set()
    launches = eng.launch_count() - launches0
    # back-to-back (no flush, one advance call, one graph launch per step): launch-overhead view
    barrier()
    tb, _ = eng.advance(args.steps, elapsed=t)
    ms_b2b = eng.last_advance_ms()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e_steps = max(3, min(args.steps, 20))
    te = tb
    eng.set_solution(u_np)                                    # untimed: the first pass of the chain (graph capture)
    te, _ = eng.advance(1, elapsed=te)
    eng.get_solution(out=u_np)
    barrier()
    e0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.set_solution(u_np)
        te, _ = eng.advance(1, elapsed=te)
        eng.get_solution(out=u_np)
    barrier()
    e2e_s = time.perf_counter() - e0
    owned_dof = eng.cell_range()[1] * D - eng.cell_range()[0] * D
    # The chain above is serial by construction (copy in, step, copy out), so PCIe carries one direction at a time.
    # Independent batches pipeline: E2E_CTX contexts (a ctx is single-threaded by contract, include/dflo_b200.h), one
    # host thread and one pinned buffer each, run the same set_solution + advance + get_solution per step, and the copy
    # in of one batch overlaps the copy out of the other.  Single GPU only: several sharded contexts per GPU could
    # deadlock on each other's flag waits (a resident stage kernel of one context can starve the peer's other context),
    # so sharded runs report the serial chain.
    E2E_CTX = int(os.environ.get("DFLO_BENCH_E2E_CTX", "3"))
    e2e_pipe_s = None
    if world == 1 and E2E_CTX > 1:
        try:
            engs = [eng] + [abi.Engine(flat, params, device=local_rank) for _ in range(E2E_CTX - 1)]
            with near_gpu_cores(local_rank):
                bufs = [u_host] + [torch.empty(n_dof, dtype=torch.float64, pin_memory=True) for _ in range(E2E_CTX - 1)]
            for b_ in bufs[1:]:
                b_.copy_(u_host)
            gate, errs, ends = threading.Barrier(E2E_CTX + 1), [], [0.0] * E2E_CTX

            def worker(i):
                try:
                    torch.cuda.set_device(local_rank)
                    e_, a_ = engs[i], bufs[i].numpy()
                    e_.set_solution(a_)
                    tt, _ = e_.advance(1, elapsed=0.0)      # untimed: first capture of this ctx
                    e_.get_solution(out=a_)
                    gate.wait(60.0)
                    for _ in range(e2e_steps):
                        e_.set_solution(a_)
                        tt, _ = e_.advance(1, elapsed=tt)
                        e_.get_solution(out=a_)
                    ends[i] = time.perf_counter()
                except Exception as ex:                      # noqa: BLE001
                    errs.append(repr(ex))
                    gate.abort()

            ths = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(E2E_CTX)]
            for th_ in ths:
                th_.start()
            gate.wait(60.0)
            p0 = time.perf_counter()
            for th_ in ths:
                th_.join(120.0)
            if not errs and all(ends):
                e2e_pipe_s = max(ends) - p0
            for e_ in engs[1:]:
                e_.close()
        except Exception as ex:                              # noqa: BLE001
            print("bench.py: pipelined e2e skipped: %r" % (ex,), file=sys.stderr)
            e2e_pipe_s = None

    # ---- kernel-only timing of the dominant (stage) kernel for the roofline ----
    k_ms = eng.time_stage_kernel(rk=1, reps=20, flush_bytes=0 if args.no_flush else FLUSH_BYTES)
    k_ms_warm = eng.time_stage_kernel(rk=1, reps=20, flush_bytes=0)
    eng.poll_error()

    # ---- N > 1: the sharded result against the single-GPU result of the same (weak-scaling) mesh ----
    linf_single = None
    if world > 1:
        u0 = initial_dofs(nx, ny, x0, x1, y0, y1, k)
        eng.set_solution(u0)
        eng.advance(3, elapsed=0.0)
        mine = np.zeros(n_dof)
        eng.get_solution(out=mine)
        g = torch.from_numpy(mine).cuda()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        if rank == 0:
            one = abi.Engine(flat, params, device=local_rank)
            one.set_solution(u0)
            one.advance(3, elapsed=0.0)
            linf_single = {"value": float(np.abs(one.get_solution() - g.cpu().numpy()).max()), "steps": 3,
                           "what": "max |u_sharded - u_single_gpu| over all DoFs of the %dx%d mesh (absolute; bit for bit = 0)" % (nx, ny)}
            one.close()
        del g
        dist.barrier()

    vals = torch.tensor([ms_total, ms_b2b, e2e_s, k_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(vals, op=dist.ReduceOp.MAX)
    ms_total, ms_b2b, e2e_s, k_ms = [float(v) for v in vals.cpu()]
    eng.close()
    peak, peak_src = measured_peak_hbm()

    # ---- N > 1: strong scaling of the limited configurations at full size ----
    strong = []
    if world > 1:
        for key in [s for s in args.strong.split(",") if s]:
            e = strong_entry(L, key, args.config_steps, 5, flush, peak, torch, dist, rank, world, local_rank, new_nccl_id)
            if e:
                strong.append(e)

    if rank == 0:
        updates_per_step = n_dof * n_rk
        value = updates_per_step * args.steps / (ms_total * 1e-3) / 1e6
        bytes_per_update = 24.0 + 32.0 / D   # read u, read u_old, write u (+ cell mean)
        alg_bytes_launch = (n_dof // world) * bytes_per_update
        achieved = alg_bytes_launch / (k_ms * 1e-3) / 1e9
        kernel = "row_stage_kernel<%d,%s>" % (k + 1, flux) if basis == "Qk" else "PkCellStageKernel<%d,%s>" % (k + 1, flux)
        cfg = workload_config(args.workload, world)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": cfg,
            "setup": {"parallelism": ("cells sharded by id over %d GPUs, 1 halo exchange per stage over NVLink peer memory "
                                      "(fused into the stage kernel)" % world) if world > 1 else "one GPU, no exchange",
                      "l2": "flushed before every timed step (256 MB rewrite)" if flush is not None else "not flushed",
                      "timing": "CUDA events on the ctx stream around each step's graph launch, summed, max over ranks"},
            "value_back_to_back_no_flush": updates_per_step * args.steps / (ms_b2b * 1e-3) / 1e6,
            "wall_s_timed_region": wall,
            "e2e": e2e_entry(updates_per_step, e2e_steps, e2e_s, e2e_pipe_s, E2E_CTX, owned_dof),
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(kernel) if world == 1 else None, "kernel": kernel,
                         "kernel_ms": k_ms, "kernel_ms_l2_warm": k_ms_warm,
                         "algorithmic_bytes_per_launch": alg_bytes_launch,
                         "bytes_per_dof_update": bytes_per_update, "peak_source": peak_src},
            "clocks": clocks_summary(clk_lines),
        }
        if world > 1:
            line["linf_vs_single"] = linf_single
            line["strong"] = strong
        else:
            threads = os.cpu_count() or 1
            if not args.no_cpu_baseline:
                # the oracle on the SAME mesh: parity at the three horizons, and its own speed = the CPU baseline
                size = [npg]
                linf, cpu = parity_leg(args.workload, size, args.parity_steps, threads)
                line["linf_vs_ref"] = linf
                line["cpu_baseline"] = cpu
            cfgs = []
            for key in [s for s in args.configs.split(",") if s]:
                e = config_entry(L, key, args.config_steps, 5, flush, peak, torch)
                if not args.no_cpu_baseline:
                    e["linf_vs_ref"], e["cpu_baseline"] = parity_leg(key, CONFIGS[key][2], args.parity_steps, threads)
                    if key == "cfg1":
                        e["cpu_baseline_1_thread"] = b0_single_thread("cfg1")
                cfgs.append(e)
            if "q1" in args.next_rows.split(","):
                cfgs.append(q1_entry(args.config_steps, 5, flush, peak, torch, args.parity_steps, threads, not args.no_cpu_baseline))
            line["configs"] = cfgs
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
