"""The drop-in boundary as an executable: oracle/_ref/macroc_dropin is the reference's OWN
src/main.c, init.c, forces.c, output.c and util.c (unmodified, compiled in place by oracle/Makefile
`dropin`) linked against libmacroc_b200.so through macroc_b200/host/dropin_glue.c, which stands
where src/assembly.c and src/bcs.c stood (include/macroc.h:130-155).  Its log and info.dat must
match what the all-reference binary (oracle/_ref/macroc_ref) produced for the same command lines
(the golden fixtures of tests/golden/make_golden.py)."""
import os
import re
import subprocess

import numpy as np
import pytest

from helpers import golden_cases, load_golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DROPIN = os.path.join(ROOT, "oracle", "_ref", "macroc_dropin")
REFBIN = os.path.join(ROOT, "oracle", "_ref", "macroc_ref")


def parse(txt):
    res = [float(x) for x in re.findall(r"\|RES\| = (\S+)", txt)]
    ksp = [(float(a), int(b)) for a, b in re.findall(r"KSP : \|Ax - b\|/\|Ax\| = (\S+)\tIts = (\d+)", txt)]
    newton = [int(x) for x in re.findall(r"Newton Iteration = (\d+)", txt)]
    return res, ksp, newton


def test_dropin_binary_is_built_from_reference_main():
    """CPU-side: the binary exists (built where /root/reference is present) and binds the C ABI."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/macroc_dropin not built (no reference tree in this checkout)")
    out = subprocess.run(["nm", "-D", "--undefined-only", DROPIN], capture_output=True, text=True).stdout
    for sym in ("macroc_create", "macroc_set_strains", "macroc_assembly_res", "macroc_assembly_jac", "macroc_solve_Ax",
                "macroc_apply_bc_on_u", "macroc_bc_lists", "macroc_calc_B"):
        assert sym in out, sym
    # the reference's own Newton loop is in there (its strings), not ours
    strings = subprocess.run(["strings", DROPIN], capture_output=True, text=True).stdout
    assert "Homogenizing MicroPP" in strings and "Assemblying RHS" in strings


@pytest.mark.gpu
@pytest.mark.parametrize("material", ["uniform", "per_gp"])
@pytest.mark.parametrize("name", golden_cases())
def test_reference_main_over_the_c_abi(name, material, tmp_path):
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/macroc_dropin not built")
    z, kv = load_golden(name)
    args = [a for kvp in kv.items() for a in kvp]
    env = dict(os.environ, MACROC_DROPIN_MATERIAL=material)
    r = subprocess.run([DROPIN] + args, cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    res, ksp, newton = parse(r.stdout)
    assert newton == list(z["newton_lines"])                       # identical Newton history
    assert len(res) == len(z["res_norms"]) and len(ksp) == len(z["ksp_its"])
    assert all(abs(k[1] - int(b)) <= 1 for k, b in zip(ksp, z["ksp_its"]))
    first = 0
    for t in range(int(kv["-ts"])):
        n_lines = 1 if z["res_norms"][first] == 0 else 2
        assert res[first] == pytest.approx(float(z["res_norms"][first]), rel=1e-5, abs=1e-300)
        first += n_lines
    info = np.loadtxt(tmp_path / "info.dat", ndmin=2)              # forces.c, unmodified, fed through MicroPP's calls
    assert np.allclose(info[:, 2], z["U"]) and np.allclose(info[:, 3], z["force"], rtol=1e-4, atol=1e-9)
    if os.path.exists(REFBIN):                                      # side by side with the all-reference binary
        ref_dir = tmp_path / "ref"
        ref_dir.mkdir()
        rr = subprocess.run([REFBIN] + args, cwd=ref_dir, capture_output=True, text=True, timeout=600)
        assert rr.returncode == 0
        strip = lambda s: [re.sub(r"[-+]?\d+\.\d+(e[-+]\d+)?", "#", ln) for ln in s.splitlines()
                           if not ln.startswith("Elapsed")]
        a, b = strip(r.stdout), strip(rr.stdout)
        assert len(a) == len(b)
        # same lines in the same order once floating-point literals are masked (CG counts may differ by 1)
        diff = [(x, y) for x, y in zip(a, b) if x != y and "Its =" not in x]
        assert not diff, diff[:3]


@pytest.mark.gpu
def test_dropin_writes_reference_vtu(tmp_path):
    """output.c (unmodified) writes the VTU from the host Vec u the glue keeps current."""
    if not os.path.exists(DROPIN):
        pytest.skip("oracle/_ref/macroc_dropin not built")
    args = "-da_grid_x 5 -da_grid_y 3 -da_grid_z 4 -ts 3 -bc_type 0 -vtu_freq 1".split()
    r = subprocess.run([DROPIN] + args, cwd=tmp_path, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-1500:]
    gold = open(os.path.join(ROOT, "tests", "golden", "vtu", "ctest_5x3x4_bending_solution_2-subdo-0.vtu")).read().split()
    mine = open(tmp_path / "solution_2-subdo-0.vtu").read().split()
    assert len(gold) == len(mine)
    for g, m in zip(gold, mine):
        try:
            fg, fm = float(g), float(m)
        except ValueError:
            assert g == m
            continue
        assert fm == pytest.approx(fg, rel=1e-4, abs=1e-7 * max(1.0, abs(fg)))
