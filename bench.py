#!/usr/bin/env python
"""bench.py -- Newton-step DOF/s and CG-MatMult HBM GB/s of the MacroC hot path on B200.

Contract (see the task statement): `python bench.py --gpus N --steps K --warmup W` prints ONE
JSON line on rank 0.  For N > 1 it is launched by torch.distributed.run, one rank per GPU.

Workload (BASELINE.json configs[3], weak scaling; at N=1 it is configs[2]): a block of
256 x 256 x (256*N) nodes, z-slab DMDA split (-da_processors_z N), bending boundary condition,
linear-elastic homogenised D (E=1e7, nu=0.25), CG + Jacobi with the reference's tolerances.
A "step" is one MacroC time step = one full Newton iteration of src/main.c:53-82: halo of u,
residual + norm, Jacobian assembly + Dirichlet rows/cols, the complete PCG solve to rtol 1e-5,
u += du, and the second residual evaluation that makes the Newton loop break.

  value      = global DOF / step time, operands resident in HBM (device events, max over ranks)
  e2e        = the same step driven through the C ABI with HOST buffers: u is uploaded from pinned
               host memory before and downloaded after every step, inside the timed region
  roofline   = the operator application of the headline operator (block-stencil SpMV), timed live
               with CUDA events around every 8th application inside the timed solves; algorithmic
               bytes 72*nb + 16*nd (full storage) or 36*(nb + Nn) + 16*nd (symmetric storage)
  cpu_baseline / --impl reference = the CPU oracle's PETSc-shaped path (scalar CSR AIJ, unfused
               CG, the reference's own 4-deep Ke loop) on the box's host cores: a genuinely timed
               full Newton step on a 96^3 grid; per-DOF phase costs are scaled to the workload's
               CG iteration count (the only extrapolated quantity).
This program never writes inside the repository (scratch goes to gpurun_out/).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

DEFAULT_OPERATOR = "sym"        # assembled operator, symmetric storage (same matrix, same CG history, half the bytes)
METRIC = "newton_step_dof_per_s"
UNIT = "DOF/s"

# CG iterations of one Newton step of the workload (a property of the problem, identical for both
# arms because they run the same algorithm).  Measured by the GPU arm in round 1 (profiles/
# cg_iterations.json, profiles/r1_bench_*.json); used by the CPU arm to scale its genuinely timed
# 96^3 step to the workload.  key = number of z-slabs of 256 planes.  --cg-its overrides.
KNOWN_CG_ITS = {1: 882, 2: 1052, 4: 1052, 8: 986}


def load_known_its(n):
    return KNOWN_CG_ITS.get(n)


def bench_config(nx, ny, nz, world, custom):
    """The `config` object -- identical in both arms (how an arm stores the operator is its own business
    and is reported beside it as `operator`)."""
    return {"workload": (f"{nx}x{ny}x{nz} nodes hex8 cantilever (custom grid, z-slab DMDA split), bending BC, one "
                         "Newton step per time step") if custom else
                        (f"{nx}x{ny}x{nz} nodes hex8 cantilever (BASELINE configs[3]: 256^3 nodes per GPU, "
                         "z-slab DMDA split; N=1 is configs[2]), bending BC, one Newton step per time step"),
            "grid": [nx, ny, nz], "ndof": 3 * nx * ny * nz, "matrix": "assembled 27-point 3x3-block stencil (PETSc MATAIJ in the reference)",
            "parallelism": f"z-slabs x{world}", "ksp": "cg+jacobi rtol 1e-5", "l2": "inputs >> L2 (operator >= 17 GB per GPU)"}


def workload(n_gpus: int, grid: int, override=None):
    nx = ny = grid
    nz = grid * n_gpus
    if override and all(override):
        nx, ny, nz = override
    nd = 3 * nx * ny * nz
    nb = (3 * nx - 2) * (3 * ny - 2) * (3 * nz - 2)
    return nx, ny, nz, nd, nb


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------
# CPU arm (the oracle = PETSc-shaped restatement of the reference; oracle/ is only used here
# as the timed CPU baseline and in tests as the checker)
# --------------------------------------------------------------------------------------

class CpuSample:
    """The CPU arm's two measurements on the host cores (`threads` OpenMP threads standing in for
    MPI ranks, z-slab row partition):

    full_step(): ONE GENUINELY TIMED Newton step of the reference's algorithm on a G^3 grid
      (default 96^3, 2.65 M DOF): halo + strains, residual + norm, Jacobian with the reference's
      own 4-deep Ke loop (assembly.c:94-99) + MatZeroRowsColumns, the complete un-fused PETSc-shaped
      CG solve to rtol 1e-5, u += du, second residual.  Nothing in it is modelled.  Its per-DOF
      phase costs are then scaled to the workload: only the CG iteration count changes
      (`its` of the 256^3-per-GPU problem instead of the count measured on G^3).
    quick_model(): the round-1 sample (residual + Jacobian loops on 48^3, 10 CG iterations on a
      96^3 AIJ matrix) -- kept as the warm-up step and to validate the per-DOF model against
      full_step() (reported as `model_over_measured`)."""

    def __init__(self, threads: int, grid: int = 96, n_cg: int = 10):
        self.threads = threads
        self.grid = grid
        self.n_cg = n_cg
        self._full = None
        self._quick = None
        self.desc = (f"per step: one genuinely timed full Newton step of the reference algorithm on a {grid}^3 grid "
                     f"(2 residuals, Jacobian with the reference's 4-deep Ke loop, complete CG+Jacobi solve to rtol 1e-5) "
                     f"on {threads} host threads; value = 1 / (2 c_res + c_jac + its c_it) with the per-DOF phase costs "
                     "measured in that step and its = CG iterations of the workload (the only scaled quantity)")

    def _oracle(self, n, faithful):
        from oracle import oracle as O
        r = min(self.threads, n)
        return O.Oracle(O.Config(NX=n, NY=n, NZ=n, bc_type=0, lx=50., ly=1., lz=50., faithful_ke=faithful,
                                 nthreads=self.threads, nranks=r, px=1, py=1, pz=r))

    def full_step(self, its_workload: int):
        if self._full is None:
            self._full = self._oracle(self.grid, 1)
            self._t = 1
        o = self._full
        T = time.perf_counter
        t0 = T()
        o.apply_bc_on_u(o.get_displacement(self._t)); self._t += 1
        o.set_strains(); o.homogenize(); o.assembly_res()
        t1 = T()
        o.assembly_jac()
        t2 = T()
        its, _ = o.solve()
        t3 = T()
        o.update_u()
        o.set_strains(); o.homogenize(); o.assembly_res()
        t4 = T()
        nd = o.ndof
        c_res = 0.5 * ((t1 - t0) + (t4 - t3)) / nd
        c_jac = (t2 - t1) / nd
        c_it = (t3 - t2) / max(its, 1) / nd
        per_dof = 2 * c_res + c_jac + its_workload * c_it
        return {"dof_per_s": 1.0 / per_dof, "measured_step_s": t4 - t0, "measured_grid": [self.grid] * 3,
                "measured_ndof": nd, "measured_cg_iterations": its, "measured_dof_per_s": nd / (t4 - t0),
                "s_per_dof_residual": c_res, "s_per_dof_jacobian": c_jac, "s_per_dof_cg_iteration": c_it}

    def quick_model(self, its: int):
        if self._quick is None:
            asm = self._oracle(48, 1)
            asm.apply_bc_on_u(-1e-3)
            slv = self._oracle(96, 0)
            slv.assembly_jac()
            self._quick = (asm, slv)
        asm, slv = self._quick
        t0 = time.perf_counter()
        asm.set_strains(); asm.homogenize(); asm.assembly_res()
        t_res = time.perf_counter() - t0
        t0 = time.perf_counter()
        asm.assembly_jac()
        t_jac = time.perf_counter() - t0
        t_cg = slv.time_cg_iterations(self.n_cg) / self.n_cg
        c_res, c_jac, c_it = t_res / asm.ndof, t_jac / asm.ndof, t_cg / slv.ndof
        return {"dof_per_s": 1.0 / (2 * c_res + c_jac + its * c_it), "s_per_dof_residual": c_res,
                "s_per_dof_jacobian": c_jac, "s_per_dof_cg_iteration": c_it}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    nx, ny, nz, nd, nb = workload(args.gpus, args.grid, (args.nx, args.ny, args.nz))
    custom = bool(args.nx and args.ny and args.nz)
    its = args.cg_its or load_known_its(args.gpus) or 1000
    cpu = CpuSample(threads, args.cpu_grid)
    model = None
    for _ in range(args.warmup):                    # warm-up: the cheap per-DOF model (pages in the library, spins up the threads)
        model = cpu.quick_model(its)
    vals, parts = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu.full_step(its)
        vals.append(r["dof_per_s"]); parts.append(r)
    wall = time.perf_counter() - t0
    v = statistics.mean(vals)
    last = parts[-1]
    cb = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu.desc,
          "measured_step": {k: last[k] for k in ("measured_grid", "measured_ndof", "measured_step_s", "measured_cg_iterations",
                                                  "measured_dof_per_s")},
          "s_per_dof": {k: statistics.mean(p[k] for p in parts) for k in parts[0] if k.startswith("s_per_dof")},
          "cg_iterations_of_workload": its}
    if model:
        # the cheap model evaluated for the SAME 96^3 step that was timed: how good is per-DOF scaling?
        pred = 1.0 / (2 * model["s_per_dof_residual"] + model["s_per_dof_jacobian"] +
                      last["measured_cg_iterations"] * model["s_per_dof_cg_iteration"])
        cb["model_over_measured"] = pred / last["measured_dof_per_s"]
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * nd / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": bench_config(nx, ny, nz, args.gpus, custom),
        "operator": "assembled (scalar AIJ CSR, 32-bit column indices)",
        "note": ("the reference's CPU path (PETSc-shaped oracle port) cannot hold the workload's 48.5 GB AIJ matrix per GPU-sized "
                 f"slab in bounded time: every timed step is a full Newton step on {args.cpu_grid}^3 nodes "
                 f"({wall / max(args.steps, 1):.1f} s each), scaled per DOF to the workload's {its} CG iterations; "
                 "ms_per_step is that scaled time for the whole workload"),
        "cpu_baseline": cb,
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# --------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------

def run_gpu_arm(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    import macroc_b200 as M

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch N>1 with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
        args.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [M.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        uid = box[0]

    def bcast_id():
        """a fresh NCCL id for one more communicator (rank 0 creates it)"""
        box = [M.get_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nx, ny, nz, nd, nb = workload(world, args.grid, (args.nx, args.ny, args.nz))
    custom = bool(args.nx and args.ny and args.nz)
    OPS = {"assembled": M.OP_ASSEMBLED, "sym": M.OP_ASSEMBLED_SYM, "matrix-free": M.OP_MATRIX_FREE}
    op_name = "matrix-free" if args.matrix_free else args.operator
    op = OPS[op_name]
    # the reference's default lengths (macroc.h:47-49), lz grows with the slab count so dz is fixed;
    # the element matrix is the unit-cube one scaled by wg, so CG counts do not
    # depend on the lengths -- they only keep |RES| above the absolute Newton tolerance 1e-1
    cfg = M.Config(NX=nx, NY=ny, NZ=nz, pz=world, lx=50.0, ly=1.0, lz=50.0 * (1 if custom else world), bc_type=M.BC_BENDING,
                   ts=args.steps + args.warmup + 1, device=local_rank, op=op)
    m = M.MacroC(cfg, rank=rank, nranks=world, unique_id=uid)
    nloc = m.local_ndof
    allreduce_path = m.allreduce_path()
    part = M.partition(cfg, rank, world)
    zs, nzl = part["corners"][2], part["corners"][5]
    zblocks = 3 * nzl - (1 if zs == 0 else 0) - (1 if zs + nzl == nz else 0)
    nb_local = (3 * nx - 2) * (3 * ny - 2) * zblocks
    # algorithmic bytes of this rank's operator application: every stored 3x3 block once,
    # p read once, w written once.  Symmetric storage keeps the diagonal block and one of each
    # off-diagonal pair: (nb + Nn) / 2 blocks.
    apply_bytes_local = {"assembled": 72 * nb_local + 16 * nloc, "sym": 36 * (nb_local + nloc // 3) + 16 * nloc,
                         "matrix-free": 16 * nloc}[op_name]

    # ---- warm-up (also sizes the operator, JIT-free) -------------------------------------
    step_idx = 1                                        # time step 0 does no work
    cg_its = []
    for _ in range(args.warmup):
        r = m.time_step(step_idx); step_idx += 1
        cg_its.append(sum(r["ksp_its"]))

    # ---- timed region 1: operands resident in HBM ----------------------------------------
    sampler = ClockSampler(local_rank)
    m.profile_enable(True, 8)
    launches0 = m.launch_count()
    barrier()
    sampler.start()
    m.event_record(0)
    t0 = time.perf_counter()
    newton = []
    for _ in range(args.steps):
        r = m.time_step(step_idx); step_idx += 1
        cg_its.append(sum(r["ksp_its"])); newton.append(r["newton_its"])
    m.event_record(1)
    barrier()
    wall = time.perf_counter() - t0
    dev_ms = m.event_elapsed_ms(0, 1)
    clocks = sampler.stop()
    launches = m.launch_count() - launches0
    apply_ms, apply_samples = m.profile_get()
    solve_ms, solve_its = m.profile_get_solve()
    m.profile_enable(False, 1)
    ms_step = max_over_ranks(max(dev_ms, 0.0)) / args.steps
    wall_ms_step = max_over_ranks(wall * 1e3) / args.steps
    apply_ms_max = max_over_ranks(apply_ms)
    cg_iteration_ms = max_over_ranks(solve_ms / solve_its if solve_its else 0.0)
    total_launches = int(sum_over_ranks(float(launches)))

    # ---- timed region 2: end to end through the C ABI with HOST buffers -------------------
    h_u = torch.empty(nloc, dtype=torch.float64).pin_memory()
    m.get_vec_ptr(M.VEC_U, h_u.data_ptr())
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        m.set_vec_ptr(M.VEC_U, h_u.data_ptr())            # H2D of the step's input displacement
        r = m.time_step(step_idx); step_idx += 1
        m.get_vec_ptr(M.VEC_U, h_u.data_ptr())            # D2H of the result
        cg_its.append(sum(r["ksp_its"]))
    barrier()
    e2e_ms_step = max_over_ranks((time.perf_counter() - t0) * 1e3) / args.steps
    h2d = int(sum_over_ranks(8.0 * nloc)); d2h = int(sum_over_ranks(8.0 * nloc + 8.0 * 2))

    # ---- the other operators offered alongside (north_star): one warm-up + one timed step each ----
    alt = {}
    if not args.no_matrix_free:
        for name in ("assembled", "sym", "matrix-free"):
            if name == op_name:
                continue
            try:
                m.set_operator(OPS[name])
                m.time_step(step_idx); step_idx += 1               # warm-up (allocates the operator)
                barrier()
                m.profile_enable(True, 8)
                m.event_record(2)
                r = m.time_step(step_idx); step_idx += 1
                m.event_record(3)
                barrier()
                a_ms = max_over_ranks(m.event_elapsed_ms(2, 3))
                s_ms, s_its = m.profile_get_solve()
                ap_ms, _ = m.profile_get()
                m.profile_enable(False, 1)
                alt[name] = {"value": nd / (a_ms * 1e-3), "unit": UNIT, "ms_per_step": a_ms, "cg_iterations": sum(r["ksp_its"]),
                             "newton_its": r["newton_its"], "operator_apply_ms": max_over_ranks(ap_ms),
                             "cg_iteration_ms": max_over_ranks(s_ms / s_its if s_its else 0.0),
                             "note": "the same Newton step with this operator"}
            except Exception as exc:                               # an optional section must not cost the headline line
                alt[name] = {"error": str(exc)}
                m.profile_enable(False, 1)
        m.set_operator(op)
        m.assembly_jac()

    # ---- isolated kernel timings (explain the headline; outside the timed regions) ---------
    kern = {}
    fp64_tflops = None
    if not args.no_kernels:
        fp64_tflops = max_over_ranks(m.fp64_probe())
        names = [("apply_matrix_free", 1), ("residual", 4), ("pcg_iteration_matrix_free", 5)]
        if op == M.OP_ASSEMBLED:
            names += [("spmv_assembled", 0), ("jacobian_fill", 3), ("pcg_iteration_assembled", 2), ("jacobian_per_element", 7)]
        if op == M.OP_ASSEMBLED_SYM:
            # the per-element Jacobian kernel (uniform D from constant memory) into the symmetric layout
            names += [("spmv_sym", 8), ("pcg_iteration_sym", 9), ("jacobian_per_element_sym", 17), ("jacobian_per_element", 7)]
        for name, what in names:
            m.time_kernel(what, 5)                          # 5 warm-ups, 20 timed launches, median
            kern[name] = max_over_ranks(statistics.median(m.time_kernel(what, 1) for _ in range(20)))
        if op != M.OP_MATRIX_FREE:
            m.assembly_jac()                                # what=7 left the element-kernel operator behind (same values to 1e-13)

    # ---- BASELINE configs[4]: 512^3 strong scaling on 8 GPUs (every rank takes part) --------
    other = {}
    m.close()
    c5g = int(os.environ.get("MACROC_BENCH_C5_GRID", "512" if world == 8 else "0"))   # (the variable: dry runs of this section)
    if c5g and not custom and not args.no_extras:
        try:
            c5 = M.MacroC(M.Config(NX=c5g, NY=c5g, NZ=c5g, pz=world, lx=50.0, ly=1.0, lz=50.0, bc_type=M.BC_BENDING, ts=8,
                                   device=local_rank, op=op), rank=rank, nranks=world,
                          unique_id=bcast_id() if world > 1 else None)
            c5.time_step(1)
            c5.profile_enable(True, 8)
            barrier()
            c5.event_record(0)
            rs = [c5.time_step(t) for t in (2, 3)]
            c5.event_record(1)
            barrier()
            ms5 = max_over_ranks(c5.event_elapsed_ms(0, 1)) / len(rs)
            s_ms, s_its = c5.profile_get_solve()
            ap5, _ = c5.profile_get()
            other["strong_scaling_512cube_8gpu" if (c5g, world) == (512, 8) else f"strong_scaling_{c5g}cube_{world}gpu"] = {
                "ndof": 3 * c5g ** 3, "steps": len(rs), "ms_per_step": ms5, "value": 3 * c5g ** 3 / (ms5 * 1e-3), "unit": UNIT,
                "cg_iterations_per_step": statistics.mean(sum(r["ksp_its"]) for r in rs),
                "newton_its_per_step": [r["newton_its"] for r in rs], "operator": op_name,
                "cg_iteration_ms": max_over_ranks(s_ms / s_its if s_its else 0.0), "operator_apply_ms": max_over_ranks(ap5),
                "note": f"BASELINE configs[4]: {c5g}^3 nodes, -da_processors_z {world}, bending BC, multi-step Newton"}
            c5.close()
        except Exception as exc:
            other["strong_scaling_512cube_8gpu"] = {"error": str(exc)}

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- BASELINE configs[1] (cantilever 128x32x32, lx=10 ly=lz=1) on one GPU, for the record ----
    if world == 1 and not args.no_extras:
        c2 = M.MacroC(M.Config(NX=128, NY=32, NZ=32, lx=10., ly=1., lz=1., bc_type=M.BC_BENDING, device=local_rank, op=op))
        for t in (1, 2):
            c2.time_step(t)
        c2.event_record(0)
        rs = [c2.time_step(t) for t in (3, 4, 5, 6, 7)]
        c2.event_record(1)
        ms = c2.event_elapsed_ms(0, 1) / len(rs)
        other["cantilever_128x32x32"] = {"ndof": 393216, "ms_per_step": ms, "value": 393216 / (ms * 1e-3), "unit": UNIT,
                                         "cg_iterations_per_step": statistics.mean(sum(r["ksp_its"]) for r in rs),
                                         "newton_its_per_step": [r["newton_its"] for r in rs], "operator": op_name}
        c2.close()
        if op != M.OP_MATRIX_FREE:
            c2 = M.MacroC(M.Config(NX=128, NY=32, NZ=32, lx=10., ly=1., lz=1., bc_type=M.BC_BENDING, device=local_rank, op=M.OP_MATRIX_FREE))
            for t in (1, 2):
                c2.time_step(t)
            c2.event_record(0)
            rs = [c2.time_step(t) for t in (3, 4, 5, 6, 7)]
            c2.event_record(1)
            ms = c2.event_elapsed_ms(0, 1) / len(rs)
            other["cantilever_128x32x32_matrix_free"] = {"ndof": 393216, "ms_per_step": ms, "value": 393216 / (ms * 1e-3), "unit": UNIT,
                                                         "cg_iterations_per_step": statistics.mean(sum(r["ksp_its"]) for r in rs),
                                                         "newton_its_per_step": [r["newton_its"] for r in rs], "operator": "matrix-free"}
            c2.close()
        # the per-element Jacobian with a tangent per Gauss point (the north_star's assembly kernel) on the
        # workload grid: its own context (38 GB of tangents), uniform values written by the device stand-in
        if not args.no_kernels and not custom:
            try:
                g = args.grid
                pg = M.MacroC(M.Config(NX=g, NY=g, NZ=g, bc_type=M.BC_BENDING, device=local_rank, material=M.MAT_PER_GP,
                                       op=op if op != M.OP_MATRIX_FREE else M.OP_ASSEMBLED))
                pg.apply_bc_on_u(-1e-3); pg.set_strains(); pg.homogenize()
                for w_el, key in ((7, "jacobian_per_element_per_gp"), (17, "jacobian_per_element_per_gp_sym")):   # full / symmetric layout
                    pg.time_kernel(w_el, 2)
                    kern[key] = statistics.median(pg.time_kernel(w_el, 1) for _ in range(5))
                kern["residual_per_gp"] = statistics.median(pg.time_kernel(4, 1) for _ in range(5))
                # north_star: "DMMA ... only if ncu shows a win over FFMA": the element contraction both ways, measured
                ab = {"dmma_tflops_measured": pg.dmma_probe(), "elements": (g - 1) ** 3,
                      "what": "Ke = sum_gp B^T C_gp B of every element from per-Gauss-point tangents (csrc/dmma_ab.cuh)"}
                for v, nm in ((0, "dfma_sparsity_aware_ms"), (1, "dmma_dense_24_per_gp_ms"), (2, "dmma_upper_tiles_18_per_gp_ms")):
                    ab[nm] = pg.contraction_ab(v, reps=3)[0]
                other["contraction_ab"] = ab
                pg.close()
            except Exception as exc:
                kern["jacobian_per_element_per_gp"] = None
                other["per_gp_error"] = str(exc)

    peak, peak_src = peaks()
    its_step = statistics.mean(cg_its[args.warmup:args.warmup + args.steps]) if cg_its else 0
    value = nd / (ms_step * 1e-3)
    if op != M.OP_MATRIX_FREE:
        achieved = apply_bytes_local / (apply_ms_max * 1e-3) / 1e9 if apply_ms_max > 0 else 0.0
        kname = "k_spmv_tma" if op == M.OP_ASSEMBLED else "k_spmv_sym"
        roof = {"bound": "hbm",
                "kernel": ("k_spmv_tma<8,4> (assembled 27-slot 3x3-block stencil SpMV + fused p.w)" if op == M.OP_ASSEMBLED else
                           "k_spmv_sym (assembled operator, symmetric storage: 14 of 27 slots, band sweep + fused p.w)"),
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "frac_of_nominal_8000_gbs": achieved / 8000.0,
                "algorithmic_bytes_per_launch": apply_bytes_local,
                "algorithmic_bytes_formula": "72*nb + 16*nd" if op == M.OP_ASSEMBLED else "36*(nb + Nn) + 16*nd",
                "launch_ms": apply_ms_max,
                "samples_in_timed_region": apply_samples, "traffic": load_traffic(kname)}
    else:
        flops = 486.0 * (nloc / 3)
        achieved = flops / (apply_ms_max * 1e-3) / 1e12 if apply_ms_max > 0 else 0.0
        pk = fp64_tflops or 37.0
        roof = {"bound": "fp64", "kernel": "k_apply_mf_march (matrix-free class-stencil apply, z-marching)", "achieved": achieved,
                "peak": pk, "unit": "TFLOP/s", "frac": achieved / pk,
                "peak_source": "measured DFMA rate (macroc_fp64_probe)" if fp64_tflops else "nominal B200 fp64 (unmeasured)",
                "launch_ms": apply_ms_max, "samples_in_timed_region": apply_samples, "traffic": None}

    threads = os.cpu_count() or 1
    cpu_obj = None
    if not args.no_cpu and world == 1:                      # reported at N=1 only (bench contract)
        cpu = CpuSample(threads, args.cpu_grid)
        its_w = int(round(its_step)) or 1
        model = cpu.quick_model(its_w)                      # warm-up + the cheap per-DOF model
        r = cpu.full_step(its_w)                            # one genuinely timed Newton step on cpu_grid^3
        pred = 1.0 / (2 * model["s_per_dof_residual"] + model["s_per_dof_jacobian"] +
                      r["measured_cg_iterations"] * model["s_per_dof_cg_iteration"])
        cpu_obj = {"value": r["dof_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu.desc,
                   "measured_step": {k: r[k] for k in ("measured_grid", "measured_ndof", "measured_step_s",
                                                       "measured_cg_iterations", "measured_dof_per_s")},
                   "s_per_dof": {k: v for k, v in r.items() if k.startswith("s_per_dof")},
                   "cg_iterations_of_workload": its_w, "model_over_measured": pred / r["measured_dof_per_s"]}

    # FP64-pipe fractions of the FP64-bound kernels against the MEASURED DFMA rate
    fp64 = None
    if fp64_tflops:
        nn = nloc / 3
        ne = (nx - 1) * (ny - 1) * (nz - 1) / world
        fp64 = {"dfma_tflops_measured": fp64_tflops, "how": "k_fp64_probe(_const): 16 independent DFMA chains per thread, 8 CTAs/SM, register and constant-bank multiplier, best of 2 x 5 launches"}
        if kern.get("apply_matrix_free"):
            fp64["apply_matrix_free_frac"] = 486.0 * nn / (kern["apply_matrix_free"] * 1e-3) / 1e12 / fp64_tflops
        # element Jacobian: EXECUTED flops of the node-centric kernels (2 x 2112 FMA per node row entry set: 19 008 FMA per
        # node in the full layout, 12 960 in the symmetric one) over the measured DFMA rate
        nodes = nn
        for key, fma_per_node in (("jacobian_per_element", 19008.0), ("jacobian_per_element_sym", 12960.0),
                                  ("jacobian_per_element_per_gp", 19008.0), ("jacobian_per_element_per_gp_sym", 12960.0)):
            if kern.get(key):
                fp64[key + "_frac"] = 2 * fma_per_node * nodes / (kern[key] * 1e-3) / 1e12 / fp64_tflops
        for key, stored in (("jacobian_per_element_per_gp", 72.0 * nb_local), ("jacobian_per_element_per_gp_sym", 36.0 * (nb_local + nn))):
            if kern.get(key):
                fp64[key + "_hbm_frac"] = (stored + 2304.0 * ne) / (kern[key] * 1e-3) / 1e9 / peak

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": bench_config(nx, ny, nz, world, custom),
        "operator": {"assembled": "assembled, full 27-slot block storage", "sym": "assembled, symmetric block storage (14 of 27 slots)",
                     "matrix-free": "matrix-free 27-point stencil"}[op_name],
        "cg_iterations_per_step": its_step, "newton_its_per_step": newton, "wall_ms_per_step": wall_ms_step,
        "cg_iteration_ms": cg_iteration_ms, "cg_allreduce": allreduce_path,
        "cg_matmult_gbps": roof.get("achieved") if op != M.OP_MATRIX_FREE else None,
        "cg_iteration_dof_per_s": nd * its_step / (ms_step * 1e-3) if its_step else None,
        "roofline": roof, "cpu_baseline": cpu_obj, "matrix_free": alt.get("matrix-free"),
        "assembled_sym": alt.get("sym"), "assembled_full": alt.get("assembled"), "other_configs": other,
        "e2e": {"value": nd / (e2e_ms_step * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_step,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": total_launches, "clocks": clocks, "kernels_ms": kern, "fp64": fp64,
    }
    emit(line)
    if world > 1:
        dist.barrier(); dist.destroy_process_group()
    return 0


def load_traffic(kernel: str):
    """dram bytes per launch from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f).get(kernel)
    except Exception:
        return None


_REAL_STDOUT = None


def emit(line: dict):
    """The ONE JSON line goes to the real stdout; everything else the process prints (NCCL
    banners, library chatter) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode()); sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--grid", type=int, default=256, help="nodes per direction per GPU")
    ap.add_argument("--nx", type=int, default=0, help="custom global grid (with --ny --nz), e.g. 512 512 512 = BASELINE configs[4]")
    ap.add_argument("--ny", type=int, default=0)
    ap.add_argument("--nz", type=int, default=0)
    ap.add_argument("--operator", default=DEFAULT_OPERATOR, choices=["assembled", "sym", "matrix-free"],
                    help="operator of the headline solve: assembled (full 27-slot storage, the reference's MATAIJ), "
                         "sym (assembled, symmetric storage) or matrix-free")
    ap.add_argument("--matrix-free", action="store_true", help="same as --operator matrix-free")
    ap.add_argument("--cg-its", type=int, default=0, help="(reference arm) CG iterations of one step of the workload")
    ap.add_argument("--cpu-grid", type=int, default=96, help="(CPU arm) nodes per direction of the genuinely timed Newton step")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-matrix-free", action="store_true", help="skip the extra matrix-free time steps")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1] cantilever record")
    ap.add_argument("--no-kernels", action="store_true", help="skip the isolated kernel timings")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
