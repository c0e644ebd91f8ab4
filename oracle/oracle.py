"""ctypes binding of the CPU oracle (oracle/macroc_oracle.c) -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product (macroc_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liboracle.so")
REF_BIN = os.path.join(HERE, "_ref", "macroc_ref")


def build(quiet: bool = True) -> None:
    """Compile liboracle.so and, if /root/reference exists, oracle/_ref/macroc_ref."""
    subprocess.run(["make", "-C", HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


class _Cfg(C.Structure):
    _fields_ = [
        ("NX", C.c_int), ("NY", C.c_int), ("NZ", C.c_int),
        ("px", C.c_int), ("py", C.c_int), ("pz", C.c_int),
        ("nranks", C.c_int),
        ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double),
        ("bc_type", C.c_int),
        ("E", C.c_double), ("nu", C.c_double),
        ("rtol", C.c_double), ("abstol", C.c_double), ("dtol", C.c_double),
        ("maxits", C.c_int),
        ("newton_min_tol", C.c_double), ("newton_rel_tol", C.c_double),
        ("newton_max_its", C.c_int),
        ("dt", C.c_double), ("final_time", C.c_double),
        ("ts", C.c_int),
        ("faithful_ke", C.c_int), ("nthreads", C.c_int), ("physical_B", C.c_int),
    ]


class _StepLog(C.Structure):
    _fields_ = [
        ("newton_its", C.c_int),
        ("ksp_its", C.c_int * 8),
        ("res_norm", C.c_double * 8),
        ("ksp_rnorm", C.c_double * 8),
        ("n_res", C.c_int),
        ("U", C.c_double), ("force", C.c_double),
    ]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        build()
    L = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    L.orc_default_config.argtypes = [C.POINTER(_Cfg)]
    L.orc_create.argtypes = [C.POINTER(_Cfg)]
    L.orc_create.restype = C.c_void_p
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_calc_B.argtypes = [C.c_int, dp]
    L.orc_isotropic_D.argtypes = [C.c_double, C.c_double, dp]
    L.orc_elem_jac.argtypes = [dp, C.c_double, dp]
    L.orc_elem_res.argtypes = [dp, C.c_double, dp]
    for name in ("orc_proc_grid",):
        getattr(L, name).argtypes = [C.c_void_p, ip]
    for name in ("orc_corners", "orc_ghost_corners", "orc_elements_sizes"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_int, ip]
    L.orc_nelem.argtypes = [C.c_void_p, C.c_int]
    L.orc_elements.argtypes = [C.c_void_p, C.c_int]
    L.orc_elements.restype = ip
    L.orc_l2g.argtypes = [C.c_void_p, C.c_int]
    L.orc_l2g.restype = ip
    L.orc_bc_list.argtypes = [C.c_void_p, C.c_int, C.POINTER(ip)]
    L.orc_bc_list_positive.argtypes = [C.c_void_p, C.c_int, C.POINTER(ip)]
    L.orc_ndof.argtypes = [C.c_void_p]; L.orc_ndof.restype = C.c_int64
    L.orc_nnz.argtypes = [C.c_void_p]; L.orc_nnz.restype = C.c_int64
    L.orc_wg.argtypes = [C.c_void_p]; L.orc_wg.restype = C.c_double
    L.orc_get_displacement.argtypes = [C.c_void_p, C.c_int]
    L.orc_get_displacement.restype = C.c_double
    L.orc_apply_bc_on_u.argtypes = [C.c_void_p, C.c_double]
    L.orc_set_strains.argtypes = [C.c_void_p]
    L.orc_homogenize.argtypes = [C.c_void_p]
    L.orc_assembly_res.argtypes = [C.c_void_p, dp]
    L.orc_assembly_jac.argtypes = [C.c_void_p]
    L.orc_solve.argtypes = [C.c_void_p, ip, dp]
    L.orc_update_u.argtypes = [C.c_void_p]
    L.orc_calc_force.argtypes = [C.c_void_p]; L.orc_calc_force.restype = C.c_double
    L.orc_run.argtypes = [C.c_void_p, C.POINTER(_StepLog), C.c_char_p]
    L.orc_get_vec.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_set_vec.argtypes = [C.c_void_p, C.c_int, dp]
    L.orc_get_block_stencil.argtypes = [C.c_void_p, dp]
    L.orc_get_csr.argtypes = [C.c_void_p, C.POINTER(C.POINTER(C.c_int64)),
                              C.POINTER(C.POINTER(C.c_int32)), C.POINTER(dp)]
    L.orc_natural_to_petsc.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.orc_matmult.argtypes = [C.c_void_p, dp, dp]
    L.orc_strain.argtypes = [C.c_void_p, C.c_int]; L.orc_strain.restype = dp
    L.orc_stress.argtypes = [C.c_void_p, C.c_int]; L.orc_stress.restype = dp
    L.orc_time_cg_iterations.argtypes = [C.c_void_p, C.c_int]
    L.orc_time_cg_iterations.restype = C.c_double
    L.orc_wtime.restype = C.c_double
    _lib = L
    return L


def _dptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def calc_B(gp: int) -> np.ndarray:
    B = np.zeros((6, 24))
    lib().orc_calc_B(gp, _dptr(B))
    return B


def isotropic_D(E: float = 1.0e7, nu: float = 0.25) -> np.ndarray:
    D = np.zeros((6, 6))
    lib().orc_isotropic_D(E, nu, _dptr(D))
    return D


def elem_jac(ctan: np.ndarray, wg: float) -> np.ndarray:
    ctan = np.ascontiguousarray(ctan, dtype=np.float64).reshape(8, 36)
    Ae = np.zeros((24, 24))
    lib().orc_elem_jac(_dptr(ctan), wg, _dptr(Ae))
    return Ae


def elem_res(stress: np.ndarray, wg: float) -> np.ndarray:
    stress = np.ascontiguousarray(stress, dtype=np.float64).reshape(8, 6)
    be = np.zeros(24)
    lib().orc_elem_res(_dptr(stress), wg, _dptr(be))
    return be


@dataclass
class StepLog:
    newton_its: int
    ksp_its: list
    res_norm: list
    ksp_rnorm: list
    U: float
    force: float


@dataclass
class Config:
    NX: int = 40
    NY: int = 3
    NZ: int = 40
    px: int = 0
    py: int = 0
    pz: int = 0
    nranks: int = 1
    lx: float = 50.0
    ly: float = 1.0
    lz: float = 50.0
    bc_type: int = 1
    E: float = 1.0e7
    nu: float = 0.25
    rtol: float = 1.0e-5
    abstol: float = 1.0e-50
    dtol: float = 1.0e4
    maxits: int = 10000
    newton_min_tol: float = 1.0e-1
    newton_rel_tol: float = 1.0e-4
    newton_max_its: int = 5
    dt: float = 0.001
    final_time: float = 1.0
    ts: int = 1
    faithful_ke: int = 1
    nthreads: int = 1
    physical_B: int = 0
    extra: dict = field(default_factory=dict)

    def to_c(self) -> _Cfg:
        c = _Cfg()
        for name, _ in _Cfg._fields_:
            setattr(c, name, getattr(self, name))
        return c


class Oracle:
    """One MacroC problem on `nranks` simulated MPI ranks (all in this process)."""

    def __init__(self, cfg: Config):
        self.cfg = cfg
        self._L = lib()
        cc = cfg.to_c()
        self._h = self._L.orc_create(C.byref(cc))
        if not self._h:
            raise ValueError("orc_create failed (bad processor grid?)")
        self.ndof = int(self._L.orc_ndof(self._h))
        self.nnodes = self.ndof // 3
        self.nnz = int(self._L.orc_nnz(self._h))
        self.wg = float(self._L.orc_wg(self._h))
        pg = (C.c_int * 3)()
        self._L.orc_proc_grid(self._h, pg)
        self.proc_grid = tuple(pg)
        self.nranks = self.proc_grid[0] * self.proc_grid[1] * self.proc_grid[2]

    def close(self):
        if self._h:
            self._L.orc_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- DMDA queries ---------------------------------------------------------
    def _i6(self, fn, rank, n=6):
        out = (C.c_int * n)()
        fn(self._h, rank, out)
        return tuple(out)

    def corners(self, rank): return self._i6(self._L.orc_corners, rank)
    def ghost_corners(self, rank): return self._i6(self._L.orc_ghost_corners, rank)
    def elements_sizes(self, rank): return self._i6(self._L.orc_elements_sizes, rank, 3)

    def elements(self, rank) -> np.ndarray:
        n = self._L.orc_nelem(self._h, rank)
        p = self._L.orc_elements(self._h, rank)
        return np.ctypeslib.as_array(p, shape=(max(n, 0) * 8,)).reshape(-1, 8).copy() if n > 0 else np.zeros((0, 8), np.int32)

    def l2g(self, rank) -> np.ndarray:
        g = self.ghost_corners(rank)
        n = g[3] * g[4] * g[5] * 3
        return np.ctypeslib.as_array(self._L.orc_l2g(self._h, rank), shape=(n,)).copy()

    def bc_list(self, rank) -> np.ndarray:
        p = C.POINTER(C.c_int)()
        n = self._L.orc_bc_list(self._h, rank, C.byref(p))
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32)

    def bc_list_positive(self, rank) -> np.ndarray:
        p = C.POINTER(C.c_int)()
        n = self._L.orc_bc_list_positive(self._h, rank, C.byref(p))
        return np.ctypeslib.as_array(p, shape=(n,)).copy() if n > 0 else np.zeros(0, np.int32)

    def natural_to_petsc(self) -> np.ndarray:
        perm = np.zeros(self.nnodes, np.int32)
        self._L.orc_natural_to_petsc(self._h, perm.ctypes.data_as(C.POINTER(C.c_int32)))
        return perm

    def dirichlet_mask_natural(self) -> np.ndarray:
        """bool[ndof] in natural ordering: union of all ranks' Dirichlet lists."""
        m = np.zeros(self.ndof, bool)
        for r in range(self.nranks):
            m[self.bc_list_positive(r)] = True
        perm = self.natural_to_petsc()
        return m.reshape(-1, 3)[perm].reshape(-1)

    # --- hot path ---------------------------------------------------------------
    def get_displacement(self, time_s): return float(self._L.orc_get_displacement(self._h, time_s))
    def apply_bc_on_u(self, U): return self._L.orc_apply_bc_on_u(self._h, U)
    def set_strains(self): return self._L.orc_set_strains(self._h)
    def homogenize(self): return self._L.orc_homogenize(self._h)

    def assembly_res(self) -> float:
        n = C.c_double()
        self._L.orc_assembly_res(self._h, C.byref(n))
        return n.value

    def assembly_jac(self): return self._L.orc_assembly_jac(self._h)

    def solve(self):
        its = C.c_int(); rn = C.c_double()
        self._L.orc_solve(self._h, C.byref(its), C.byref(rn))
        return its.value, rn.value

    def update_u(self): return self._L.orc_update_u(self._h)
    def calc_force(self): return float(self._L.orc_calc_force(self._h))

    def run(self, log_path: str | None = None):
        steps = (_StepLog * max(self.cfg.ts, 1))()
        self._L.orc_run(self._h, steps, log_path.encode() if log_path else None)
        out = []
        for s in steps[: self.cfg.ts]:
            out.append(StepLog(s.newton_its, list(s.ksp_its[: s.newton_its]),
                               list(s.res_norm[: s.n_res]), list(s.ksp_rnorm[: s.newton_its]),
                               s.U, s.force))
        return out

    # --- export -------------------------------------------------------------------
    def get_vec(self, which: str) -> np.ndarray:
        out = np.zeros(self.ndof)
        self._L.orc_get_vec(self._h, {"u": 0, "du": 1, "b": 2}[which], _dptr(out))
        return out

    def set_vec(self, which: str, v: np.ndarray):
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert v.size == self.ndof
        self._L.orc_set_vec(self._h, {"u": 0, "du": 1, "b": 2}[which], _dptr(v))

    def block_stencil(self) -> np.ndarray:
        out = np.zeros((self.nnodes, 27, 3, 3))
        self._L.orc_get_block_stencil(self._h, _dptr(out))
        return out

    def csr(self):
        """(rowptr, col, val) views in PETSc global ordering."""
        rp = C.POINTER(C.c_int64)(); cp = C.POINTER(C.c_int32)(); vp = C.POINTER(C.c_double)()
        self._L.orc_get_csr(self._h, C.byref(rp), C.byref(cp), C.byref(vp))
        rowptr = np.ctypeslib.as_array(rp, shape=(self.ndof + 1,))
        col = np.ctypeslib.as_array(cp, shape=(self.nnz,))
        val = np.ctypeslib.as_array(vp, shape=(self.nnz,))
        return rowptr, col, val

    def matmult(self, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.zeros(self.ndof)
        self._L.orc_matmult(self._h, _dptr(x), _dptr(y))
        return y

    def strain(self, rank) -> np.ndarray:
        n = self._L.orc_nelem(self._h, rank)
        return np.ctypeslib.as_array(self._L.orc_strain(self._h, rank), shape=(n, 8, 6)).copy()

    def stress(self, rank) -> np.ndarray:
        n = self._L.orc_nelem(self._h, rank)
        return np.ctypeslib.as_array(self._L.orc_stress(self._h, rank), shape=(n, 8, 6)).copy()

    def time_cg_iterations(self, n: int) -> float:
        return float(self._L.orc_time_cg_iterations(self._h, n))


def run_reference(args: list[str], cwd: str, dump_prefix: str | None = None) -> str:
    """Run oracle/_ref/macroc_ref (the reference's own sources over the serial
    PETSc shim) with MacroC's command-line flags; returns its stdout."""
    env = dict(os.environ)
    if dump_prefix:
        env["MACROC_SHIM_DUMP"] = dump_prefix
    r = subprocess.run([REF_BIN] + [str(a) for a in args], cwd=cwd, env=env,
                       capture_output=True, text=True, check=True)
    return r.stdout


def read_shim_matrix(path: str):
    """CSR dumped by the shim's KSPSolve (natural ordering, one rank)."""
    with open(path, "rb") as f:
        n, nnz = np.fromfile(f, np.int64, 2)
        rowptr = np.fromfile(f, np.int64, n + 1)
        col = np.fromfile(f, np.int32, nnz)
        val = np.fromfile(f, np.float64, nnz)
    return rowptr, col, val
