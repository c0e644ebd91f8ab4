"""Generates tests/golden/*.npz from the REFERENCE ITSELF: oracle/_ref/macroc_ref
is GG1991/macroc's own src/*.c compiled (unmodified, in place) over the serial
PETSc shim in oracle/shim/ with the linear-elastic MicroPP stand-in.

Run in the build container (needs /root/reference):  python tests/golden/make_golden.py
Each fixture holds, for one command line: the |RES| and KSP lines the binary
printed, the matrix A and right-hand side b handed to the last KSPSolve, its
solution x, and the final displacement u (natural ordering).
"""
import glob
import os
import re
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O  # noqa: E402

CASES = {
    # name: (NX, NY, NZ, ts, bc_type, extra flags)
    "readme_4x4x2_circle": (4, 4, 2, 3, 1, []),
    "readme_4x4x2_bending": (4, 4, 2, 3, 0, []),
    "ctest_5x2x2_bending": (5, 2, 2, 5, 0, []),
    "ctest_4x4x4_circle": (4, 4, 4, 5, 1, []),
    "ctest_3x3x3_bending": (3, 3, 3, 5, 0, []),
    "ctest_5x3x4_bending": (5, 3, 4, 5, 0, []),
    "beam_16x6x6_bending": (16, 6, 6, 3, 0, ["-lx", "10", "-ly", "1", "-lz", "1"]),
    "plate_9x3x9_circle": (9, 3, 9, 3, 1, ["-lx", "4", "-lz", "4"]),
}


def main():
    O.build()
    out_dir = os.path.dirname(os.path.abspath(__file__))
    for name, (nx, ny, nz, ts, bc, extra) in CASES.items():
        with tempfile.TemporaryDirectory() as d:
            args = ["-da_grid_x", nx, "-da_grid_y", ny, "-da_grid_z", nz, "-ts", ts, "-bc_type", bc] + extra
            txt = O.run_reference(args, d, os.path.join(d, "dump"))
            res = [float(m) for m in re.findall(r"\|RES\| = (\S+)", txt)]
            ksp = [(float(a), int(b)) for a, b in re.findall(r"KSP : \|Ax - b\|/\|Ax\| = (\S+)\tIts = (\d+)", txt)]
            newton = [int(m) for m in re.findall(r"Newton Iteration = (\d+)", txt)]
            fa = sorted(glob.glob(os.path.join(d, "dump_A*.bin")), key=lambda p: int(re.findall(r"_A(\d+)", p)[0]))
            rowptr, col, val = O.read_shim_matrix(fa[-1])
            n = len(fa) - 1
            b = np.fromfile(os.path.join(d, f"dump_b{n}.bin"))
            x = np.fromfile(os.path.join(d, f"dump_x{n}.bin"))
            u = np.fromfile(os.path.join(d, "dump_vec0.bin"))
            info = np.loadtxt(os.path.join(d, "info.dat"), ndmin=2)
            np.savez_compressed(os.path.join(out_dir, name + ".npz"), args=np.array([str(a) for a in args]),
                                res_norms=np.array(res), ksp_rnorm=np.array([k[0] for k in ksp]),
                                ksp_its=np.array([k[1] for k in ksp]), newton_lines=np.array(newton),
                                rowptr=rowptr, col=col, val=val, b=b, x=x, u=u, force=info[:, 3], U=info[:, 2])
            print(name, "res", res[:4], "ksp", ksp[:3])
    # VTU output of the reference (src/output.c) for one case: -vtu_freq 1, keep the last step
    import shutil
    with tempfile.TemporaryDirectory() as d:
        args = ["-da_grid_x", 5, "-da_grid_y", 3, "-da_grid_z", 4, "-ts", 3, "-bc_type", 0, "-vtu_freq", 1]
        O.run_reference(args, d)
        os.makedirs(os.path.join(out_dir, "vtu"), exist_ok=True)
        for f in ("solution_2.pvtu", "solution_2-subdo-0.vtu"):
            shutil.copy(os.path.join(d, f), os.path.join(out_dir, "vtu", "ctest_5x3x4_bending_" + f))
        print("vtu fixtures written")


if __name__ == "__main__":
    main()
