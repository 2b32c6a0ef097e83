"""The MEX gateway cannot run here (no MATLAB / Octave / mex.h in the image); it is at least type-checked
against a stand-in mex.h, and every ABI symbol it calls must be declared in include/pcreg.h."""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gateway_typechecks_against_fake_mex():
    src = os.path.join(ROOT, "pcreg_b200", "csrc", "pcreg_mex.cpp")
    r = subprocess.run(["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Werror", "-I", os.path.join(ROOT, "tests", "fake_mex"), src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


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
