"""ctypes binding of include/macroc_b200.h (libmacroc_b200.so).

This is the host-side mirror of the reference's operator interface for the hot
path (reference include/macroc.h:130-155): the same function names, argument
meaning and error convention (PetscErrorCode-style ints, raised here as
MacrocError).  All numerics run in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# MACROC_B200_LIB: an alternative build of the same library (measurement builds)
LIB_PATH = os.environ.get("MACROC_B200_LIB") or os.path.join(HERE, "lib", "libmacroc_b200.so")
CSRC = os.path.join(HERE, "csrc")

BC_BENDING, BC_CIRCLE = 0, 1
VEC_U, VEC_DU, VEC_B = 0, 1, 2
OP_ASSEMBLED, OP_MATRIX_FREE, OP_ASSEMBLED_SYM = 0, 1, 2
MAT_UNIFORM, MAT_PER_GP = 0, 1
JAC_AUTO, JAC_ELEMENT = 0, 1
ERR_NO_DEVICE = 97

# every symbol include/macroc_b200.h declares
EXPORTS = [
    "macroc_default_config", "macroc_config_from_args", "macroc_get_unique_id", "macroc_create",
    "macroc_destroy", "macroc_last_error", "macroc_partition", "macroc_bc_lists",
    "macroc_get_displacement", "macroc_apply_bc_on_u", "macroc_set_strains", "macroc_assembly_res",
    "macroc_assembly_jac", "macroc_solve_Ax", "macroc_ksp_reason", "macroc_update_u", "macroc_calc_B",
    "macroc_calc_force", "macroc_time_step", "macroc_local_ndof", "macroc_global_ndof", "macroc_set_vec",
    "macroc_get_vec", "macroc_get_matrix_blocks", "macroc_matmult", "macroc_get_strain_stress",
    "macroc_time_kernel", "macroc_launch_count", "macroc_device_synchronize", "macroc_version",
    "macroc_event_record", "macroc_event_elapsed_ms", "macroc_profile_enable", "macroc_profile_get",
    "macroc_homogenize", "macroc_gp_arrays", "macroc_set_gp_data", "macroc_set_operator", "macroc_write_pvtu",
    "macroc_loopback_id", "macroc_fp64_probe", "macroc_dmma_probe", "macroc_contraction_ab", "macroc_profile_get_solve", "macroc_allreduce_path",
]


class MacrocError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"macroc_b200 error {code}: {msg}")
        self.code = code


class CConfig(C.Structure):
    _fields_ = [
        ("NX", C.c_int32), ("NY", C.c_int32), ("NZ", C.c_int32),
        ("px", C.c_int32), ("py", C.c_int32), ("pz", C.c_int32),
        ("lx", C.c_double), ("ly", C.c_double), ("lz", C.c_double),
        ("bc_type", C.c_int32), ("ts", C.c_int32), ("vtu_freq", C.c_int32), ("pad0", C.c_int32),
        ("dt", C.c_double), ("final_time", C.c_double),
        ("newton_max_its", C.c_int32),
        ("newton_min_tol", C.c_double), ("newton_rel_tol", C.c_double),
        ("ksp_rtol", C.c_double), ("ksp_abstol", C.c_double), ("ksp_dtol", C.c_double),
        ("ksp_maxits", C.c_int32),
        ("E", C.c_double), ("nu", C.c_double),
        ("D", C.c_double * 36),
        ("use_D", C.c_int32), ("op", C.c_int32), ("device", C.c_int32),
        ("material", C.c_int32), ("jac_mode", C.c_int32), ("physical_B", C.c_int32),
        ("strict_fp", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


def build(verbose: bool = False) -> str:
    """Compile libmacroc_b200.so for sm_100a with nvcc (in-tree)."""
    subprocess.run(["make", "-C", CSRC], check=True, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    """Load the C-ABI library; fails loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  macroc_b200 has no CPU fallback.")
    # torch first: its bundled libnccl.so.2 must be the one the process binds
    try:
        import torch  # noqa: F401
    except Exception:
        pass
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    L.macroc_default_config.argtypes = [C.POINTER(CConfig)]
    L.macroc_config_from_args.argtypes = [C.POINTER(CConfig), C.c_int, C.POINTER(C.c_char_p)]
    L.macroc_get_unique_id.argtypes = [C.c_void_p]
    L.macroc_loopback_id.argtypes = [C.c_int, C.c_void_p]
    L.macroc_create.argtypes = [C.POINTER(CConfig), C.c_int, C.c_int, C.c_void_p, C.POINTER(vp)]
    L.macroc_destroy.argtypes = [vp]
    L.macroc_last_error.argtypes = [vp]
    L.macroc_last_error.restype = C.c_char_p
    L.macroc_partition.argtypes = [C.POINTER(CConfig), C.c_int, C.c_int, C.POINTER(C.c_int32)]
    L.macroc_bc_lists.argtypes = [C.POINTER(CConfig), C.c_int, C.c_int, C.POINTER(C.c_int32), dp,
                                  C.POINTER(C.c_int32)]
    L.macroc_get_displacement.argtypes = [vp, C.c_int]
    L.macroc_get_displacement.restype = C.c_double
    L.macroc_apply_bc_on_u.argtypes = [vp, C.c_double]
    L.macroc_set_strains.argtypes = [vp, C.c_int]
    L.macroc_assembly_res.argtypes = [vp, dp]
    L.macroc_homogenize.argtypes = [vp]
    L.macroc_gp_arrays.argtypes = [vp, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                   C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.macroc_set_gp_data.argtypes = [vp, C.c_void_p, C.c_void_p]
    L.macroc_assembly_jac.argtypes = [vp]
    L.macroc_solve_Ax.argtypes = [vp, ip, dp]
    L.macroc_ksp_reason.argtypes = [vp, ip]
    L.macroc_set_operator.argtypes = [vp, C.c_int]
    L.macroc_write_pvtu.argtypes = [vp, C.c_char_p]
    L.macroc_update_u.argtypes = [vp]
    L.macroc_calc_B.argtypes = [C.c_int, dp]
    L.macroc_calc_force.argtypes = [vp, dp]
    L.macroc_time_step.argtypes = [vp, C.c_int, ip, dp, ip, ip, dp]
    L.macroc_local_ndof.argtypes = [vp]; L.macroc_local_ndof.restype = C.c_int64
    L.macroc_global_ndof.argtypes = [vp]; L.macroc_global_ndof.restype = C.c_int64
    L.macroc_set_vec.argtypes = [vp, C.c_int, C.c_void_p]
    L.macroc_get_vec.argtypes = [vp, C.c_int, C.c_void_p]
    L.macroc_get_matrix_blocks.argtypes = [vp, C.c_void_p]
    L.macroc_matmult.argtypes = [vp, C.c_int, C.c_void_p, C.c_void_p]
    L.macroc_get_strain_stress.argtypes = [vp, C.c_void_p, C.c_void_p, C.POINTER(C.c_int64)]
    L.macroc_time_kernel.argtypes = [vp, C.c_int, C.c_int, C.c_int, dp]
    L.macroc_fp64_probe.argtypes = [vp, dp]
    L.macroc_dmma_probe.argtypes = [vp, dp]
    L.macroc_contraction_ab.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_void_p, dp]
    L.macroc_allreduce_path.argtypes = [vp]
    L.macroc_launch_count.argtypes = [vp]; L.macroc_launch_count.restype = C.c_uint64
    L.macroc_device_synchronize.argtypes = [vp]
    L.macroc_event_record.argtypes = [vp, C.c_int]
    L.macroc_event_elapsed_ms.argtypes = [vp, C.c_int, C.c_int, dp]
    L.macroc_profile_enable.argtypes = [vp, C.c_int, C.c_int]
    L.macroc_profile_get.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.macroc_profile_get_solve.argtypes = [vp, dp, C.POINTER(C.c_int64)]
    L.macroc_version.restype = C.c_int
    _lib = L
    return L


@dataclass
class Config:
    """MacroC's run parameters (reference src/init.c:47-83, defaults of include/macroc.h:36-51)."""
    NX: int = 40
    NY: int = 3
    NZ: int = 40
    px: int = 0
    py: int = 0
    pz: int = 0
    lx: float = 50.0
    ly: float = 1.0
    lz: float = 50.0
    bc_type: int = BC_CIRCLE
    ts: int = 1
    dt: float = 0.001
    final_time: float = 1.0
    newton_max_its: int = 5
    newton_min_tol: float = 1.0e-1
    newton_rel_tol: float = 1.0e-4
    ksp_rtol: float = 1.0e-5
    ksp_abstol: float = 1.0e-50
    ksp_dtol: float = 1.0e4
    ksp_maxits: int = 10000
    E: float = 1.0e7
    nu: float = 0.25
    op: int = OP_ASSEMBLED
    device: int = -1
    material: int = MAT_UNIFORM
    jac_mode: int = JAC_AUTO
    physical_B: int = 0
    strict_fp: int = 0
    D: np.ndarray | None = None

    def to_c(self) -> CConfig:
        c = CConfig()
        lib().macroc_default_config(C.byref(c))
        for k in ("NX", "NY", "NZ", "px", "py", "pz", "lx", "ly", "lz", "bc_type", "ts", "dt", "final_time",
                  "newton_max_its", "newton_min_tol", "newton_rel_tol", "ksp_rtol", "ksp_abstol", "ksp_dtol",
                  "ksp_maxits", "E", "nu", "op", "device", "material", "jac_mode", "physical_B", "strict_fp"):
            setattr(c, k, getattr(self, k))
        if self.D is not None:
            d = np.ascontiguousarray(self.D, dtype=np.float64).reshape(36)
            for i in range(36):
                c.D[i] = d[i]
            c.use_D = 1
        return c

    @staticmethod
    def from_args(argv: list[str]) -> "Config":
        """Parse MacroC's command line (-da_grid_x ..., -ts, -dt, -bc_type, -new_its ...)."""
        c = CConfig()
        L = lib()
        L.macroc_default_config(C.byref(c))
        arr = (C.c_char_p * len(argv))(*[str(a).encode() for a in argv])
        rc = L.macroc_config_from_args(C.byref(c), len(argv), arr)
        if rc:
            raise MacrocError(rc, "unsupported option value")
        out = Config()
        for k in ("NX", "NY", "NZ", "px", "py", "pz", "lx", "ly", "lz", "bc_type", "ts", "dt", "final_time",
                  "newton_max_its", "newton_min_tol", "newton_rel_tol", "ksp_rtol", "ksp_abstol", "ksp_dtol",
                  "ksp_maxits", "E", "nu", "op", "device"):
            setattr(out, k, getattr(c, k))
        return out


def partition(cfg: Config, rank: int, nranks: int) -> dict:
    """DMDA corners / ghost corners / element sizes of `rank` (host only, no GPU)."""
    out = (C.c_int32 * 15)()
    cc = cfg.to_c()
    rc = lib().macroc_partition(C.byref(cc), rank, nranks, out)
    if rc:
        raise MacrocError(rc, "macroc_partition")
    v = list(out)
    return {"corners": tuple(v[0:6]), "ghost_corners": tuple(v[6:12]), "elements_sizes": tuple(v[12:15])}


def bc_lists(cfg: Config, rank: int, nranks: int):
    """bc_init's index_dirichlet (global dof ids, -1 padded) and the coef with value = coef*U."""
    cc = cfg.to_c()
    n = C.c_int32()
    L = lib()
    rc = L.macroc_bc_lists(C.byref(cc), rank, nranks, None, None, C.byref(n))
    if rc:
        raise MacrocError(rc, "macroc_bc_lists")
    idx = np.zeros(n.value, np.int32)
    coef = np.zeros(n.value, np.float64)
    L.macroc_bc_lists(C.byref(cc), rank, nranks, idx.ctypes.data_as(C.POINTER(C.c_int32)),
                      coef.ctypes.data_as(C.POINTER(C.c_double)), C.byref(n))
    return idx, coef


def calc_B(gp: int) -> np.ndarray:
    B = np.zeros((6, 24))
    rc = lib().macroc_calc_B(gp, B.ctypes.data_as(C.POINTER(C.c_double)))
    if rc:
        raise MacrocError(rc, "macroc_calc_B")
    return B


def get_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = lib().macroc_get_unique_id(buf)
    if rc:
        raise MacrocError(rc, lib().macroc_last_error(None).decode())
    return buf.raw


def loopback_id(nranks: int) -> bytes:
    """Id of an in-process communicator: `nranks` MacroC objects created with it, each driven by
    its own host thread, stand in for `nranks` ranks on one GPU (decomposition tests)."""
    buf = C.create_string_buffer(128)
    rc = lib().macroc_loopback_id(nranks, buf)
    if rc:
        raise MacrocError(rc, "macroc_loopback_id")
    return buf.raw


class MacroC:
    """One rank's slab of a MacroC problem on one B200 (the reference's globals + hot-path functions)."""

    def __init__(self, cfg: Config, rank: int = 0, nranks: int = 1, unique_id: bytes | None = None):
        self.cfg = cfg
        self.rank, self.nranks = rank, nranks
        self._L = lib()
        self._h = C.c_void_p()
        cc = cfg.to_c()
        idbuf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        rc = self._L.macroc_create(C.byref(cc), rank, nranks, idbuf, C.byref(self._h))
        if rc:
            self._h = C.c_void_p()
            raise MacrocError(rc, self._L.macroc_last_error(None).decode())
        self.local_ndof = int(self._L.macroc_local_ndof(self._h))
        self.global_ndof = int(self._L.macroc_global_ndof(self._h))

    def _chk(self, rc):
        if rc:
            raise MacrocError(rc, self._L.macroc_last_error(self._h).decode())

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self._L.macroc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- the reference's functions --------------------------------------------------
    def get_displacement(self, time_s: int) -> float:
        return float(self._L.macroc_get_displacement(self._h, time_s))

    def apply_bc_on_u(self, U: float):
        self._chk(self._L.macroc_apply_bc_on_u(self._h, U))

    def set_strains(self, materialize: bool = False):
        self._chk(self._L.macroc_set_strains(self._h, int(materialize)))

    def assembly_res(self) -> float:
        n = C.c_double()
        self._chk(self._L.macroc_assembly_res(self._h, C.byref(n)))
        return n.value

    def homogenize(self):
        self._chk(self._L.macroc_homogenize(self._h))

    def gp_arrays(self):
        """Device pointers (ints) of strain, stress, ctan, the number of Gauss points and the SoA pitch."""
        a, b, c, n, pitch = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_int64(), C.c_int64()
        self._chk(self._L.macroc_gp_arrays(self._h, C.byref(a), C.byref(b), C.byref(c), C.byref(n), C.byref(pitch)))
        return a.value, b.value, c.value, n.value, pitch.value

    def set_gp_data(self, stress: np.ndarray | None = None, ctan: np.ndarray | None = None):
        s = np.ascontiguousarray(stress, dtype=np.float64) if stress is not None else None
        t = np.ascontiguousarray(ctan, dtype=np.float64) if ctan is not None else None
        self._chk(self._L.macroc_set_gp_data(self._h, s.ctypes.data if s is not None else None,
                                             t.ctypes.data if t is not None else None))

    def assembly_jac(self):
        self._chk(self._L.macroc_assembly_jac(self._h))

    def solve_Ax(self):
        its, rn = C.c_int(), C.c_double()
        self._chk(self._L.macroc_solve_Ax(self._h, C.byref(its), C.byref(rn)))
        return its.value, rn.value

    def write_pvtu(self, file_prefix: str):
        self._chk(self._L.macroc_write_pvtu(self._h, file_prefix.encode()))

    def set_operator(self, op: int):
        self._chk(self._L.macroc_set_operator(self._h, op))

    def ksp_reason(self) -> int:
        r = C.c_int()
        self._chk(self._L.macroc_ksp_reason(self._h, C.byref(r)))
        return r.value

    def update_u(self):
        self._chk(self._L.macroc_update_u(self._h))

    def calc_force(self) -> float:
        f = C.c_double()
        self._chk(self._L.macroc_calc_force(self._h, C.byref(f)))
        return f.value

    def time_step(self, time_s: int) -> dict:
        """main.c:53-82 for one time step."""
        nit, nres = C.c_int(), C.c_int()
        n = max(int(self.cfg.newton_max_its), 1) + 1           # one |RES| per iteration + the final check
        res = (C.c_double * n)(); kits = (C.c_int * n)(); krn = (C.c_double * n)()
        self._chk(self._L.macroc_time_step(self._h, time_s, C.byref(nit), res, C.byref(nres), kits, krn))
        return {"newton_its": nit.value, "res_norm": list(res[: nres.value]),
                "ksp_its": list(kits[: nit.value]), "ksp_rnorm": list(krn[: nit.value])}

    # --- vectors / matrix -------------------------------------------------------------
    def set_vec(self, which: int, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.float64)
        assert host.size == self.local_ndof
        self._chk(self._L.macroc_set_vec(self._h, which, host.ctypes.data))

    def set_vec_ptr(self, which: int, host_ptr: int):
        self._chk(self._L.macroc_set_vec(self._h, which, host_ptr))

    def get_vec(self, which: int, out: np.ndarray | None = None) -> np.ndarray:
        if out is None:
            out = np.empty(self.local_ndof)
        self._chk(self._L.macroc_get_vec(self._h, which, out.ctypes.data))
        return out

    def get_vec_ptr(self, which: int, host_ptr: int):
        self._chk(self._L.macroc_get_vec(self._h, which, host_ptr))

    def get_matrix_blocks(self) -> np.ndarray:
        out = np.empty((self.local_ndof // 3, 27, 3, 3))
        self._chk(self._L.macroc_get_matrix_blocks(self._h, out.ctypes.data))
        return out

    def matmult(self, x: np.ndarray, op: int = OP_ASSEMBLED) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        y = np.empty_like(x)
        self._chk(self._L.macroc_matmult(self._h, op, x.ctypes.data, y.ctypes.data))
        return y

    def get_strain_stress(self):
        n = C.c_int64()
        self._chk(self._L.macroc_get_strain_stress(self._h, None, None, C.byref(n)))
        e = np.empty((n.value // 8, 8, 6)); s = np.empty((n.value // 8, 8, 6))
        self._chk(self._L.macroc_get_strain_stress(self._h, e.ctypes.data, s.ctypes.data, C.byref(n)))
        return e, s

    # --- measurement ---------------------------------------------------------------------
    def time_kernel(self, what: int, reps: int = 20, flush_l2: bool = False) -> float:
        ms = C.c_double()
        self._chk(self._L.macroc_time_kernel(self._h, what, reps, int(flush_l2), C.byref(ms)))
        return ms.value

    def allreduce_path(self) -> str:
        return {0: "none (one rank)", 1: "ncclAllReduce", 2: "peer-mapped mailboxes in the reduction kernels", 3: "loopback host sum"}.get(
            int(self._L.macroc_allreduce_path(self._h)), "?")

    def fp64_probe(self) -> float:
        """Measured DFMA rate of the device in TFLOP/s (register-resident FMA chains)."""
        t = C.c_double()
        self._chk(self._L.macroc_fp64_probe(self._h, C.byref(t)))
        return t.value

    def dmma_probe(self) -> float:
        """Measured rate of mma.sync.m8n8k4.f64 (DMMA) in TFLOP/s."""
        t = C.c_double()
        self._chk(self._L.macroc_dmma_probe(self._h, C.byref(t)))
        return t.value

    def contraction_ab(self, variant: int, reps: int = 3, n_full: int = 0):
        """(ms, Ke[n_full][24][24]) of the element-contraction A/B kernels: 0 DFMA, 1 DMMA dense, 2 DMMA upper tiles."""
        ms = C.c_double()
        full = np.zeros((max(n_full, 1), 24, 24))
        self._chk(self._L.macroc_contraction_ab(self._h, variant, reps, n_full, full.ctypes.data if n_full else None, C.byref(ms)))
        return ms.value, full[:n_full]

    def event_record(self, slot: int):
        self._chk(self._L.macroc_event_record(self._h, slot))

    def event_elapsed_ms(self, a: int, b: int) -> float:
        ms = C.c_double()
        self._chk(self._L.macroc_event_elapsed_ms(self._h, a, b, C.byref(ms)))
        return ms.value

    def profile_enable(self, enable: bool = True, stride: int = 8):
        self._chk(self._L.macroc_profile_enable(self._h, int(enable), stride))

    def profile_get(self):
        ms, n = C.c_double(), C.c_int64()
        self._chk(self._L.macroc_profile_get(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def profile_get_solve(self):
        ms, n = C.c_double(), C.c_int64()
        self._chk(self._L.macroc_profile_get_solve(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def launch_count(self) -> int:
        return int(self._L.macroc_launch_count(self._h))

    def synchronize(self):
        self._chk(self._L.macroc_device_synchronize(self._h))
