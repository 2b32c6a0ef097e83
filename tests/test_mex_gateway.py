"""The MEX gateway cannot run under MATLAB here (no MATLAB / Octave / mex.h in the image).  It is type-checked against a
stand-in mex.h, every ABI symbol it calls must be declared in include/pcreg.h, and its marshaling is EXECUTED against an
in-memory fake of the mx* runtime with stubs of the C ABI (tests/fake_mex/gateway_harness.cpp): 1-based <-> 0-based
indices, column-major layouts, struct parsing, [] for degenerate results, error text, no leaked arrays."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gateway_typechecks_against_fake_mex():
    src = os.path.join(ROOT, "pcreg_b200", "csrc", "pcreg_mex.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "fake_mex"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_gateway_marshaling_runs_against_fake_runtime(tmp_path):
    exe = str(tmp_path / "gateway_harness")
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "fake_mex"),
                        os.path.join(ROOT, "tests", "fake_mex", "gateway_harness.cpp"),
                        os.path.join(ROOT, "pcreg_b200", "csrc", "pcreg_mex.cpp"), "-o", exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK:"), r.stdout + r.stderr


def test_gateway_only_calls_declared_abi():
    with open(os.path.join(ROOT, "include", "pcreg.h")) as f:
        hdr = re.sub(r"/\*.*?\*/", "", f.read(), flags=re.S)
    declared = set(re.findall(r"\b(pcreg_[a-z0-9_]+)\s*\(", hdr))
    with open(os.path.join(ROOT, "pcreg_b200", "csrc", "pcreg_mex.cpp")) as f:
        src = re.sub(r"//.*", "", f.read())
    used = set(re.findall(r"\b(pcreg_[a-z0-9_]+)\s*\(", src)) - {"pcreg_mex"}
    assert used and used <= declared, used - declared


def test_matlab_shims_cover_the_reference_signatures():
    d = os.path.join(ROOT, "pcreg_b200", "matlab")
    for name in ("AlignPoints", "AlignPoints_KNN", "AlignPoints_knn", "AlignPoints_weighted", "AlignPoints_c",
                 "AlignPoints_KNN_c", "estimateTransform", "ransac", "getLocalPoints", "getSpacialHistogramDescriptors", "getMatches", "ransac_windows", "pcreg_icp"):
        with open(os.path.join(d, name + ".m")) as f:
            first = f.readline()
        assert first.startswith("function") and name + "(" in first, first
