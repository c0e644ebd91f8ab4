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
  roofline   = the assembled block-stencil SpMV (k_spmv), timed live with CUDA events around
               every 8th application inside the timed solves; algorithmic bytes 72*nb + 16*nd
  cpu_baseline / --impl reference = the CPU oracle's PETSc-shaped path (scalar CSR AIJ, unfused
               CG, the reference's own 4-deep Ke loop) on the box's host cores.
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

METRIC = "newton_step_dof_per_s"
UNIT = "DOF/s"

# CG iterations of one Newton step of the workload (a property of the problem, identical for
# both arms because they run the same algorithm): measured by the GPU arm, used by the CPU arm
# to extrapolate its bounded sample to the full step.  key = number of z-slabs.
KNOWN_CG_ITS = {1: None, 2: None, 4: None, 8: None}
KNOWN_CG_ITS_FILE = os.path.join(ROOT, "profiles", "cg_iterations.json")


def load_known_its(n):
    try:
        with open(KNOWN_CG_ITS_FILE) as f:
            return json.load(f).get(str(n))
    except Exception:
        return KNOWN_CG_ITS.get(n)


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
    """Bounded sample of the workload on the host cores: the reference's residual + Jacobian
    loops (faithful 4-deep Ke loop, assembly.c:94-99) on a 48^3 grid and `n_cg` un-fused PETSc
    CG iterations on a 96^3 AIJ matrix; per-DOF costs are scaled to one Newton step of the
    workload (2 residuals + 1 Jacobian + `its` CG iterations)."""

    def __init__(self, threads: int, n_cg: int = 10):
        from oracle import oracle as O
        self.threads = threads
        self.n_cg = n_cg
        na, ns = 48, 96
        ra = min(threads, na)
        self.asm = O.Oracle(O.Config(NX=na, NY=na, NZ=na, bc_type=0, lx=1., ly=1., lz=1., faithful_ke=1,
                                     nthreads=threads, nranks=ra, px=1, py=1, pz=ra))
        self.asm.apply_bc_on_u(-1e-3)
        rs = min(threads, ns)
        self.slv = O.Oracle(O.Config(NX=ns, NY=ns, NZ=ns, bc_type=0, lx=1., ly=1., lz=1., faithful_ke=0,
                                     nthreads=threads, nranks=rs, px=1, py=1, pz=rs))
        self.slv.assembly_jac()
        self.desc = (f"per step: reference residual+Jacobian loops on {na}^3 nodes and {n_cg} PETSc-shaped CG "
                     f"iterations on a {ns}^3 AIJ matrix, {threads} OpenMP threads standing in for MPI ranks; "
                     "per-DOF costs scaled to one Newton step (2 residuals + 1 Jacobian + its CG iterations)")

    def step(self, its: int):
        t0 = time.perf_counter()
        self.asm.set_strains(); self.asm.homogenize(); self.asm.assembly_res()
        t_res = time.perf_counter() - t0
        t0 = time.perf_counter()
        self.asm.assembly_jac()
        t_jac = time.perf_counter() - t0
        t_cg = self.slv.time_cg_iterations(self.n_cg) / self.n_cg
        c_res = t_res / self.asm.ndof
        c_jac = t_jac / self.asm.ndof
        c_it = t_cg / self.slv.ndof
        per_dof = 2 * c_res + c_jac + its * c_it
        return {"dof_per_s": 1.0 / per_dof, "s_per_dof_residual": c_res, "s_per_dof_jacobian": c_jac,
                "s_per_dof_cg_iteration": c_it}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    nx, ny, nz, nd, nb = workload(args.gpus, args.grid, (args.nx, args.ny, args.nz))
    its = args.cg_its or load_known_its(args.gpus) or 1000
    cpu = CpuSample(threads)
    for _ in range(args.warmup):
        cpu.step(its)
    vals, parts = [], []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = cpu.step(its)
        vals.append(r["dof_per_s"]); parts.append(r)
    wall = time.perf_counter() - t0
    v = statistics.mean(vals)
    line = {
        "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * nd / v, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{nx}x{ny}x{nz} nodes hex8 cantilever (bending BC), one Newton step, "
                               f"CG+Jacobi rtol 1e-5 ({its} CG iterations), AIJ scalar CSR on the host CPU",
                   "grid": [nx, ny, nz], "ndof": nd, "cg_iterations_assumed": its,
                   "note": "ms_per_step is the extrapolated time of one full Newton step on the host cores; "
                           f"the timed sample took {wall / max(args.steps, 1):.1f} s per step"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu.desc,
                         "s_per_dof": {k: statistics.mean(p[k] for p in parts) for k in parts[0] if k != "dof_per_s"}},
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
    op = M.OP_MATRIX_FREE if args.matrix_free else M.OP_ASSEMBLED
    # the reference's default lengths (macroc.h:47-49), lz grows with the slab count so dz is fixed;
    # the element matrix is the unit-cube one scaled by wg, so CG counts do not
    # depend on the lengths -- they only keep |RES| above the absolute Newton tolerance 1e-1
    cfg = M.Config(NX=nx, NY=ny, NZ=nz, pz=world, lx=50.0, ly=1.0, lz=50.0 * (1 if custom else world), bc_type=M.BC_BENDING,
                   ts=args.steps + args.warmup + 1, device=local_rank, op=op)
    m = M.MacroC(cfg, rank=rank, nranks=world, unique_id=uid)
    nloc = m.local_ndof
    part = M.partition(cfg, rank, world)
    zs, nzl = part["corners"][2], part["corners"][5]
    zblocks = 3 * nzl - (1 if zs == 0 else 0) - (1 if zs + nzl == nz else 0)
    nb_local = (3 * nx - 2) * (3 * ny - 2) * zblocks
    spmv_bytes_local = 72 * nb_local + 16 * nloc      # algorithmic bytes of this rank's SpMV launch(es)

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
    m.profile_enable(False, 1)
    ms_step = max_over_ranks(max(dev_ms, 0.0)) / args.steps
    wall_ms_step = max_over_ranks(wall * 1e3) / args.steps
    apply_ms_max = max_over_ranks(apply_ms)
    total_launches = int(sum_over_ranks(float(launches)))

    # ---- timed region 2: end to end through the C ABI with host buffers -------------------
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

    # ---- the matrix-free variant offered alongside (north_star): one more time step ----------
    mf = None
    if op == M.OP_ASSEMBLED and not args.no_matrix_free:
        m.set_operator(M.OP_MATRIX_FREE)
        m.time_step(step_idx); step_idx += 1               # warm-up
        barrier()
        m.event_record(2)
        r = m.time_step(step_idx); step_idx += 1
        m.event_record(3)
        barrier()
        mf_ms = max_over_ranks(m.event_elapsed_ms(2, 3))
        mf = {"value": nd / (mf_ms * 1e-3), "unit": UNIT, "ms_per_step": mf_ms, "cg_iterations": sum(r["ksp_its"]),
              "newton_its": r["newton_its"], "note": "same Newton step with the matrix-free 27-point operator"}
        m.set_operator(M.OP_ASSEMBLED)
        m.assembly_jac()

    # ---- the symmetric-storage operator (opt-in; one rank, uniform tangent): one more time step ----
    sym = None
    if op == M.OP_ASSEMBLED and world == 1 and not args.no_matrix_free:
        try:
            m.set_operator(M.OP_ASSEMBLED_SYM)
            m.time_step(step_idx); step_idx += 1               # warm-up
            m.event_record(2)
            r = m.time_step(step_idx); step_idx += 1
            m.event_record(3)
            sym_ms = m.event_elapsed_ms(2, 3)
            sym = {"value": nd / (sym_ms * 1e-3), "unit": UNIT, "ms_per_step": sym_ms, "cg_iterations": sum(r["ksp_its"]),
                   "newton_its": r["newton_its"],
                   "note": "same Newton step with 14 of the 27 slots stored (A is symmetric), MACROC_OP_ASSEMBLED_SYM"}
            if not args.no_kernels:
                m.time_kernel(8, 5)
                sym["spmv_ms"] = statistics.median(m.time_kernel(8, 1) for _ in range(20))
                sym["pcg_iteration_ms"] = statistics.median(m.time_kernel(9, 1) for _ in range(20))
                sym["spmv_algorithmic_gbps"] = (1008.0 + 48.0) * (nloc / 3) / (sym["spmv_ms"] * 1e-3) / 1e9
        except Exception as exc:                               # an optional section must not cost the headline line
            sym = {"error": str(exc)}
        m.set_operator(M.OP_ASSEMBLED)
        m.assembly_jac()

    # ---- isolated kernel timings (explain the headline; outside the timed regions) ---------
    kern = {}
    if not args.no_kernels:
        for name, what in (("spmv_assembled", 0), ("apply_matrix_free", 1), ("jacobian_fill", 3), ("residual", 4),
                           ("pcg_iteration_assembled", 2), ("pcg_iteration_matrix_free", 5)):
            if what in (0, 2, 3) and op == M.OP_MATRIX_FREE:
                continue
            m.time_kernel(what, 5)                          # 5 warm-ups, 20 timed launches, median
            kern[name] = max_over_ranks(statistics.median(m.time_kernel(what, 1) for _ in range(20)))

    if rank != 0:
        if world > 1:
            dist.barrier(); dist.destroy_process_group()
        return 0

    # ---- BASELINE configs[1] (cantilever 128x32x32, lx=10 ly=lz=1) on one GPU, for the record ----
    other = {}
    if world == 1 and not args.no_extras:
        m.close()
        c2 = M.MacroC(M.Config(NX=128, NY=32, NZ=32, lx=10., ly=1., lz=1., bc_type=M.BC_BENDING, device=local_rank))
        for t in (1, 2):
            c2.time_step(t)
        c2.event_record(0)
        rs = [c2.time_step(t) for t in (3, 4, 5, 6, 7)]
        c2.event_record(1)
        ms = c2.event_elapsed_ms(0, 1) / len(rs)
        other["cantilever_128x32x32"] = {"ndof": 393216, "ms_per_step": ms, "value": 393216 / (ms * 1e-3), "unit": UNIT,
                                         "cg_iterations_per_step": statistics.mean(sum(r["ksp_its"]) for r in rs),
                                         "newton_its_per_step": [r["newton_its"] for r in rs]}
        c2.close()

    peak, peak_src = peaks()
    its_step = statistics.mean(cg_its[args.warmup:args.warmup + args.steps]) if cg_its else 0
    value = nd / (ms_step * 1e-3)
    if op == M.OP_ASSEMBLED:
        achieved = spmv_bytes_local / (apply_ms_max * 1e-3) / 1e9 if apply_ms_max > 0 else 0.0
        roof = {"bound": "hbm", "kernel": "k_spmv_tma<8,4> (assembled 27-slot 3x3-block stencil SpMV + fused p.w)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                "frac_of_nominal_8000_gbs": achieved / 8000.0,
                "algorithmic_bytes_per_launch": spmv_bytes_local, "launch_ms": apply_ms_max,
                "samples_in_timed_region": apply_samples, "traffic": load_traffic("k_spmv_tma")}
    else:
        flops = 486.0 * (nloc / 3)
        achieved = flops / (apply_ms_max * 1e-3) / 1e12 if apply_ms_max > 0 else 0.0
        roof = {"bound": "fp64", "kernel": "k_apply_mf (matrix-free class-stencil apply)", "achieved": achieved,
                "peak": 37.0, "unit": "TFLOP/s", "frac": achieved / 37.0, "peak_source": "nominal B200 fp64 (unmeasured)",
                "launch_ms": apply_ms_max, "samples_in_timed_region": apply_samples, "traffic": None}

    threads = os.cpu_count() or 1
    cpu_obj = None
    if not args.no_cpu and world == 1:                      # reported at N=1 only (bench contract)
        cpu = CpuSample(threads)
        cpu.step(int(its_step) or 1)
        r = cpu.step(int(its_step) or 1)
        cpu_obj = {"value": r["dof_per_s"], "unit": UNIT, "cores": threads, "kind": "port", "sample": cpu.desc,
                   "s_per_dof": {k: v for k, v in r.items() if k != "dof_per_s"}}
        try:
            os.makedirs(os.path.dirname(KNOWN_CG_ITS_FILE), exist_ok=True)
            known = {}
            if os.path.exists(KNOWN_CG_ITS_FILE):
                known = json.load(open(KNOWN_CG_ITS_FILE))
            if not custom:
                known[str(world)] = int(round(its_step))
            json.dump(known, open(KNOWN_CG_ITS_FILE, "w"))
        except Exception:
            pass

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": (f"{nx}x{ny}x{nz} nodes hex8 cantilever (custom grid, z-slab DMDA split), bending BC, one "
                                "Newton step per time step") if custom else
                               (f"{nx}x{ny}x{nz} nodes hex8 cantilever (BASELINE configs[3]: 256^3 nodes per GPU, "
                                "z-slab DMDA split; N=1 is configs[2]), bending BC, one Newton step per time step"),
                   "grid": [nx, ny, nz], "ndof": nd, "operator": "matrix-free" if args.matrix_free else "assembled",
                   "parallelism": f"z-slabs x{world}", "ksp": "cg+jacobi rtol 1e-5", "l2": "inputs >> L2 (32 GB operator)",
                   "cg_iterations_per_step": its_step, "newton_its_per_step": newton,
                   "wall_ms_per_step": wall_ms_step},
        "cg_matmult_gbps": roof.get("achieved") if op == M.OP_ASSEMBLED else None,
        "cg_iteration_dof_per_s": nd * its_step / (ms_step * 1e-3) if its_step else None,
        "roofline": roof, "cpu_baseline": cpu_obj, "matrix_free": mf, "assembled_sym": sym, "other_configs": other,
        "e2e": {"value": nd / (e2e_ms_step * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms_step,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": total_launches, "clocks": clocks, "kernels_ms": kern,
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
    ap.add_argument("--matrix-free", action="store_true", help="solve with the matrix-free operator")
    ap.add_argument("--cg-its", type=int, default=0, help="(reference arm) CG iterations of one step")
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
