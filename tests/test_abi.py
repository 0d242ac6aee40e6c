"""CPU checks of the C-ABI library: it loads, exports every symbol include/sphbvf.h declares, the
ctypes mirror of the config struct matches the header, host-only entry points work, and creating
a context without a CUDA device fails loudly (no CPU fallback).  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT, load_package


def header_symbols():
    text = open(os.path.join(ROOT, "include", "sphbvf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sphbvf_[a-z_0-9]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    pkg = load_package()
    L = pkg.lib()
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(L, s), "libsphbvf.so does not export " + s
    assert sorted(pkg.capi.SYMBOLS) == syms, set(pkg.capi.SYMBOLS) ^ set(syms)
    assert L.sphbvf_version() == 1


def test_field_enum_matches_header():
    pkg = load_package()
    text = open(os.path.join(ROOT, "include", "sphbvf.h")).read()
    body = re.search(r"enum sphbvf_field \{(.*?)\};", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = [n.strip().split("=")[0].strip() for n in body.split(",") if n.strip()]
    names = [n[len("SPHBVF_F_"):].lower() for n in names if n != "SPHBVF_F_COUNT"]
    mine = [n.lower() for n in pkg.capi._FIELD_NAMES]
    assert names == mine


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    pkg = load_package()
    assert pkg.lib().sphbvf_device_count() == 0
    meta = dict(dim=2, periodic=[0, 0, 1], boxlo=[0, 0, 0], boxhi=[1, 1, 1], ntypes=1, S=0, variant=0, skin=0.01,
                every=1, delay=10, check=1, dt=1e-4, types=[dict(mass=1, rho0=1, c0=1, G0=0)],
                pairs=[dict(i=1, j=1, eta=1, h=0.1, cutc=0.1, kappa=[])])
    with pytest.raises(pkg.SphbvfError):
        pkg.Engine(meta)


def test_proc_grid_and_brick_bounds():
    pkg = load_package()
    L = pkg.lib()
    prd = (C.c_double * 3)(1.0, 1.0, 1.0)
    grid = (C.c_int * 3)()
    for n, want in ((1, (1, 1, 1)), (2, None), (4, None), (8, (2, 2, 2))):
        assert L.sphbvf_proc_grid(n, 3, prd, C.byref(grid)) == 0
        assert grid[0] * grid[1] * grid[2] == n
        if want:
            assert tuple(grid) == want
    assert L.sphbvf_proc_grid(4, 2, prd, C.byref(grid)) == 0 and tuple(grid) == (2, 2, 1)
    prd2 = (C.c_double * 3)(4.0, 1.0, 1.0)
    assert L.sphbvf_proc_grid(4, 3, prd2, C.byref(grid)) == 0 and tuple(grid) == (4, 1, 1)
    # bricks tile the box exactly: neighbours share the same double for the common face
    meta = dict(dim=3, periodic=[1, 0, 0], boxlo=[-0.3, 0.0, 0.1], boxhi=[1.7, 1.0, 0.9], ntypes=1, S=0, variant=0,
                skin=0.01, every=1, delay=10, check=1, dt=1e-4)
    cfg = pkg.config_from_meta(meta, procgrid=(3, 2, 2), nranks=12)
    lo, hi = (C.c_double * 3)(), (C.c_double * 3)()
    bounds = []
    for r in range(12):
        assert L.sphbvf_brick_bounds(C.byref(cfg), r, C.byref(lo), C.byref(hi)) == 0
        bounds.append((tuple(lo), tuple(hi)))
    assert L.sphbvf_brick_bounds(C.byref(cfg), 12, C.byref(lo), C.byref(hi)) != 0
    for r in range(12):
        ix = r % 3
        if ix < 2:
            assert bounds[r][1][0] == bounds[r + 1][0][0]
        else:
            assert bounds[r][1][0] == 1.7
    vol = sum(np.prod(np.array(h) - np.array(l)) for l, h in bounds)
    assert abs(vol - 2.0 * 1.0 * 0.8) < 1e-12


def test_comm_plan_is_symmetric():
    """peer(d) of rank r must list r as its peer in direction -d with the opposite shift."""
    pkg = load_package()
    L = pkg.lib()
    for dim, periodic, grid in ((3, [0, 0, 0], (2, 2, 2)), (3, [1, 0, 1], (2, 1, 2)), (2, [1, 1, 0], (2, 2, 1)),
                                (2, [1, 0, 0], (1, 2, 1)), (3, [1, 1, 1], (4, 2, 1))):
        n = grid[0] * grid[1] * grid[2]
        meta = dict(dim=dim, periodic=periodic, boxlo=[0, 0, 0], boxhi=[2.0, 1.0, 1.5], ntypes=1, S=0, variant=0,
                    skin=0.01, every=1, delay=10, check=1, dt=1e-4)
        cfg = pkg.config_from_meta(meta, procgrid=grid, nranks=n)
        plans = []
        for r in range(n):
            peer, shift = (C.c_int * 27)(), (C.c_double * 81)()
            assert L.sphbvf_comm_plan(C.byref(cfg), r, C.byref(peer), C.byref(shift)) == 0
            plans.append((list(peer), np.array(shift).reshape(27, 3)))
        for r in range(n):
            peer, shift = plans[r]
            assert peer[13] == -1
            for d in range(27):
                if peer[d] < 0:
                    assert not shift[d].any()
                    continue
                if dim == 2:
                    assert d // 9 == 1
                back_peer, back_shift = plans[peer[d]]
                assert back_peer[26 - d] == r
                assert np.array_equal(back_shift[26 - d], -shift[d])


def test_header_is_plain_c(tmp_path):
    """include/sphbvf.h is the drop-in boundary: it must compile as C99 with nothing but the C library, declare
    every function with C linkage only, and use no C++ or CUDA types."""
    import subprocess
    src = tmp_path / "abi.c"
    src.write_text('#include "sphbvf.h"\nint main(void) { sphbvf_config c; (void)c; return sphbvf_version() < 0; }\n')
    inc = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include")
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    text = re.sub(r"/\*.*?\*/", "", open(os.path.join(inc, "sphbvf.h")).read(), flags=re.S)   # code only
    for bad in ("cuda", "torch", "std::", "class ", "template", "#include <c"):
        assert bad not in text.replace("SPHBVF_ECUDA", ""), bad
