"""The C-ABI library loads without a GPU and exports every symbol include/pcreg.h declares; without a
device the product path fails loudly (no CPU fallback)."""
import ctypes as C
import os
import re
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


@pytest.fixture(scope="module")
def built():
    from pcreg_b200.build import build_library
    return build_library()


def _declared():
    with open(os.path.join(ROOT, "include", "pcreg.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pcreg_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported(built):
    lib = C.CDLL(built)
    names = _declared()
    assert len(names) >= 19
    for n in names:
        assert hasattr(lib, n), "libpcreg_b200.so does not export %s" % n


def test_binding_table_matches_header(built):
    from pcreg_b200 import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    assert lib.pcreg_abi_version() == 2


def test_struct_layouts_match_header():
    """ctypes mirrors of the POD option structs have the C layout (sizes on x86-64 LP64)."""
    from pcreg_b200 import _lib
    assert C.sizeof(_lib.ModelOpts) == 72
    assert C.sizeof(_lib.AlignOpts) == 48
    assert C.sizeof(_lib.RansacOpts) == 24
    assert C.sizeof(_lib.IcpOpts) == 40
    assert C.sizeof(_lib.MatchOpts) == 56


def test_defaults_are_the_reference_constants(built):
    from pcreg_b200 import _lib
    lib = _lib.load()
    a = _lib.AlignOpts()
    lib.pcreg_align_opts_default(C.byref(a))
    assert (a.k_frac, a.R_w, a.r_local, a.min_local) == (0.85, 3.5, 2.0, 25)   # AlignPoints_KNN.m:20, _weighted.m:16, _c.m:13-14
    o = _lib.IcpOpts()
    lib.pcreg_icp_opts_default(C.byref(o))
    assert (o.k_frac, o.R_w, o.reflection_fix) == (0.85, 3.5, 0)
    mo = _lib.MatchOpts()
    lib.pcreg_match_opts_default(C.byref(mo))                                   # completeExperiment.m:112-122
    assert (mo.unnormalize, mo.norm_factor, mo.change_metric, mo.metric_factor) == (1, 2.0, 1, 0.6)
    assert (mo.match_threshold, mo.max_ratio, mo.metric, mo.unique) == (10.0, 0.99, 0, 1)


def test_no_cpu_fallback_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the failure path is exercised on CPU-only hosts")
    import pcreg_b200
    with pytest.raises(pcreg_b200.PcregError):
        pcreg_b200.init()
    import numpy as np
    with pytest.raises(pcreg_b200.PcregError):
        pcreg_b200.Model(np.zeros((10, 3)))


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "pcreg_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                with open(os.path.join(dirpath, fn)) as f:
                    txt = f.read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), fn


def test_header_is_plain_c99_and_links_from_c(built, tmp_path):
    """The drop-in boundary is a C ABI: include/pcreg.h must compile as C99 (no C++-isms, no torch types) and a C program
    must link against the library and call a non-compute entry point (no GPU needed for that)."""
    src = tmp_path / "c_user.c"
    src.write_text('#include <stdio.h>\n#include "pcreg.h"\n'
                   'int main(void) { pcreg_icp_opts o; pcreg_icp_opts_default(&o);\n'
                   '  printf("%d %d %.2f\\n", pcreg_abi_version(), o.iters, o.k_frac);\n'
                   '  return (pcreg_abi_version() > 0 && o.k_frac == 0.85) ? 0 : 1; }\n')
    import subprocess
    exe = tmp_path / "c_user"
    libdir = os.path.dirname(built)
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), str(src),
                        "-L", libdir, "-l:" + os.path.basename(built), "-Wl,-rpath," + libdir, "-o", str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
