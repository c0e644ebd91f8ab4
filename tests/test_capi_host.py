"""CPU tests of the C-ABI library: it loads, exports every symbol
include/macroc_b200.h declares, its host-side logic (option parsing, DMDA
partition, Dirichlet lists) agrees with the oracle, and it refuses to compute
without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import macroc_b200 as M
from macroc_b200 import capi
from oracle import oracle as O
from conftest import has_gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "macroc_b200.h")).read()
    declared = set(re.findall(r"\b(macroc_[a-zA-Z_0-9]+)\s*\(", hdr))
    L = M.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/macroc_b200.h but not exported"
    assert declared == set(capi.EXPORTS)
    assert L.macroc_version() >= 100


def test_default_config_matches_reference_defaults():
    c = capi.CConfig()
    M.lib().macroc_default_config(C.byref(c))
    # include/macroc.h:36-51, src/init.c:31,64,147-148
    assert (c.NX, c.NY, c.NZ) == (40, 3, 40) and (c.lx, c.ly, c.lz) == (50.0, 1.0, 50.0)
    assert c.bc_type == M.BC_CIRCLE and c.ts == 1 and c.dt == 0.001 and c.final_time == 1.0
    assert (c.newton_max_its, c.newton_min_tol, c.newton_rel_tol) == (5, 1e-1, 1e-4)
    assert (c.ksp_rtol, c.ksp_abstol, c.ksp_dtol, c.ksp_maxits) == (1e-5, 1e-50, 1e4, 10000)
    assert (c.E, c.nu) == (1e7, 0.25)


def test_command_line_surface():
    cfg = M.Config.from_args("-da_grid_x 128 -da_grid_y 32 -da_grid_z 32 -lx 10 -ly 1 -lz 1 -ts 10 -dt 0.01 "
                             "-bc_type 0 -da_processors_z 4 -new_its 3 -new_tol 0.5".split())
    assert (cfg.NX, cfg.NY, cfg.NZ, cfg.lx, cfg.ly, cfg.lz) == (128, 32, 32, 10.0, 1.0, 1.0)
    assert (cfg.ts, cfg.dt, cfg.bc_type, cfg.pz) == (10, 0.01, 0, 4)
    assert (cfg.newton_max_its, cfg.newton_min_tol) == (3, 0.5)        # README aliases
    cfg = M.Config.from_args("-newton_max_its 7 -newton_min_tol 1e-3 -newton_rel_tol 1e-6 -ksp_rtol 1e-8".split())
    assert (cfg.newton_max_its, cfg.newton_min_tol, cfg.newton_rel_tol, cfg.ksp_rtol) == (7, 1e-3, 1e-6, 1e-8)
    with pytest.raises(M.MacrocError):
        M.Config.from_args("-ksp_type gmres".split())


@pytest.mark.parametrize("grid", [(5, 2, 2), (5, 3, 7), (8, 4, 9)])
@pytest.mark.parametrize("nranks", [1, 2, 3, 4])
def test_partition_matches_dmda_restatement(grid, nranks):
    NX, NY, NZ = grid
    if nranks > NZ:
        pytest.skip("more ranks than planes")
    cfg = M.Config(NX=NX, NY=NY, NZ=NZ, px=1, py=1, pz=nranks)
    o = O.Oracle(O.Config(NX=NX, NY=NY, NZ=NZ, nranks=nranks, px=1, py=1, pz=nranks))
    planes = 0
    for r in range(nranks):
        p = M.partition(cfg, r, nranks)
        assert p["corners"] == o.corners(r)
        assert p["ghost_corners"] == o.ghost_corners(r)
        assert p["elements_sizes"] == o.elements_sizes(r)
        planes += p["corners"][5]
    assert planes == NZ


@pytest.mark.parametrize("bc", [M.BC_BENDING, M.BC_CIRCLE])
@pytest.mark.parametrize("nranks", [1, 2, 3])
def test_dirichlet_lists_match_bc_init(bc, nranks):
    kw = dict(NX=7, NY=3, NZ=8, lx=6.0, lz=7.0, bc_type=bc, px=1, py=1, pz=nranks)
    o = O.Oracle(O.Config(nranks=nranks, **kw))
    for r in range(nranks):
        idx, coef = M.bc_lists(M.Config(**kw), r, nranks)
        assert np.array_equal(idx, o.bc_list(r))
        assert set(np.unique(coef)) <= {0.0, 1.0}
    if bc == M.BC_CIRCLE:
        idx, coef = M.bc_lists(M.Config(**dict(kw, pz=1)), 0, 1)
        assert coef.sum() > 0            # some node falls inside the circle on this grid


@pytest.mark.parametrize("grid", [(5, 2, 2), (6, 4, 9), (9, 7, 8)])
@pytest.mark.parametrize("nranks,pg", [(2, (0, 0, 0)), (3, (0, 0, 0)), (4, (0, 0, 0)), (8, (0, 0, 0)), (2, (2, 1, 1)),
                                       (4, (2, 2, 1)), (4, (1, 2, 2)), (8, (2, 2, 2)), (6, (3, 2, 1)), (4, (0, 0, 2)),
                                       (6, (0, 3, 0))])
@pytest.mark.parametrize("bc", [M.BC_BENDING, M.BC_CIRCLE])
def test_general_dmda_boxes_match_restatement(grid, nranks, pg, bc):
    """-da_processors_x/y/z or PETSC_DECIDE: corners, ghost corners, element sizes and the
    Dirichlet lists in PETSc's rank-contiguous global numbering."""
    NX, NY, NZ = grid
    kw = dict(NX=NX, NY=NY, NZ=NZ, bc_type=bc, lx=5.0, lz=6.0, px=pg[0], py=pg[1], pz=pg[2])
    try:
        o = O.Oracle(O.Config(nranks=nranks, **kw))
    except ValueError:
        with pytest.raises(M.MacrocError):
            M.partition(M.Config(**kw), 0, nranks)
        return
    for r in range(nranks):
        p = M.partition(M.Config(**kw), r, nranks)
        assert p["corners"] == o.corners(r) and p["ghost_corners"] == o.ghost_corners(r)
        assert p["elements_sizes"] == o.elements_sizes(r)
        idx, coef = M.bc_lists(M.Config(**kw), r, nranks)
        assert np.array_equal(idx, o.bc_list(r))


def test_impossible_decompositions_are_refused():
    with pytest.raises(M.MacrocError):
        M.partition(M.Config(NX=8, NY=8, NZ=2, pz=3), 0, 3)          # more slabs than planes
    with pytest.raises(M.MacrocError):
        M.partition(M.Config(NX=8, NY=8, NZ=8, px=2, py=2, pz=2), 0, 4)   # 2*2*2 != 4


def test_calc_B_host_helper():
    for gp in range(8):
        assert np.array_equal(M.calc_B(gp), O.calc_B(gp))


@pytest.mark.skipif(has_gpu(), reason="only meaningful without a GPU")
def test_no_cpu_fallback():
    with pytest.raises(M.MacrocError) as e:
        M.MacroC(M.Config(NX=4, NY=4, NZ=2))
    assert e.value.code == capi.ERR_NO_DEVICE


def test_product_does_not_reference_the_oracle():
    """The product path must not import, link or call anything under oracle/."""
    pkg = os.path.join(ROOT, "macroc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".c", ".cpp")) or f == "Makefile":
                txt = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"


def test_header_is_plain_c99(tmp_path):
    """The drop-in boundary must be includable from C (the reference is a C99 project)."""
    import subprocess
    src = tmp_path / "use_header.c"
    src.write_text('#include "macroc_b200.h"\n'
                   "int main(void) { macroc_config c; macroc_ctx *x = 0; (void)x;\n"
                   "  return macroc_default_config(&c) + (int)sizeof(c.D) - 288; }\n")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-c", str(src), "-o", str(tmp_path / "use_header.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_config_struct_layout_matches_ctypes_mirror(tmp_path):
    """sizeof / offsets of macroc_config as the C compiler sees them == the ctypes mirror."""
    import subprocess
    src = tmp_path / "layout.c"
    fields = ["NX", "px", "lx", "bc_type", "ts", "vtu_freq", "dt", "newton_max_its", "newton_min_tol", "ksp_rtol",
              "ksp_maxits", "E", "D", "use_D", "op", "device", "material", "jac_mode"]
    body = "".join(f'  printf("{f} %zu\\n", offsetof(macroc_config, {f}));\n' for f in fields)
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "macroc_b200.h"\nint main(void) {\n'
                   '  printf("sizeof %zu\\n", sizeof(macroc_config));\n' + body + "  return 0; }\n")
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    assert int(out["sizeof"]) == C.sizeof(capi.CConfig)
    for f in fields:
        assert int(out[f]) == getattr(capi.CConfig, f).offset, f


def test_partition_and_bc_lists_fuzz():
    """Random grids and processor grids (hypothesis): the library's DMDA restatement and Dirichlet
    lists equal the oracle's for every rank, or both refuse the decomposition."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=60, deadline=None)
    @given(st.integers(2, 11), st.integers(2, 9), st.integers(2, 11), st.integers(1, 8),
           st.sampled_from([0, 1, 2, 3]), st.sampled_from([0, 1, 2]), st.sampled_from([0, 1, 2, 4]),
           st.sampled_from([M.BC_BENDING, M.BC_CIRCLE]))
    def check(NX, NY, NZ, nranks, px, py, pz, bc):
        kw = dict(NX=NX, NY=NY, NZ=NZ, bc_type=bc, lx=4.0, lz=5.0, px=px, py=py, pz=pz)
        try:
            o = O.Oracle(O.Config(nranks=nranks, **kw))
        except ValueError:
            o = None
        if o is None or any(v <= 0 for v in o.proc_grid):
            with pytest.raises(M.MacrocError):
                M.partition(M.Config(**kw), 0, nranks)
            return
        for r in range(nranks):
            p = M.partition(M.Config(**kw), r, nranks)
            assert p["corners"] == o.corners(r) and p["ghost_corners"] == o.ghost_corners(r)
            assert p["elements_sizes"] == o.elements_sizes(r)
            idx, _ = M.bc_lists(M.Config(**kw), r, nranks)
            assert np.array_equal(idx, o.bc_list(r))

    check()
