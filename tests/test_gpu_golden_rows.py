"""The CUDA path against the COMMITTED golden fixture tests/golden/rows_small.json (inputs + oracle outputs of the
AlignPoints* family, ransac.m, getLocalPoints.m and getMatches.m on small seeded problems; generator:
tests/golden/make_golden.py, kept current by tests/test_oracle.py).  Same tolerances as the live-oracle tests."""
import importlib.util
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _load():
    with open(os.path.join(HERE, "golden", "rows_small.json")) as f:
        G = json.load(f)
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    return G, mg


def test_align_family_against_committed_fixture(pcreg):
    G, mg = _load()
    p = np.asarray(G["align"]["pts"])
    scale = np.abs(p).max()
    for name, fn in mg.ALIGN_CASES:
        want = G["align"]["cases"][name]
        got = fn(pcreg, p)
        assert want["coeff"] is not None and got[1] is not None, name
        assert np.max(np.abs(got[1] - np.asarray(want["coeff"]))) < 1e-9, name
        assert np.max(np.abs(got[0] - np.asarray(want["aligned"]))) < 1e-9 * scale, name
        if want["c"] is not None:
            assert np.max(np.abs(got[2] - np.asarray(want["c"]))) < 1e-12 * scale, name


def test_ransac_against_committed_fixture(pcreg):
    G, _ = _load()
    R = G["ransac"]
    got = pcreg.ransac_seeded(np.asarray(R["p1"]), np.asarray(R["p2"]), R["coef"], seed=R["seed"], return_triplets=True)
    assert got["triplets"][:5].tolist() == R["triplets_head"]            # the documented device-side sampler
    assert got["best"] == R["best"] and got["numSuccess"] == R["numSuccess"] and got["maxInliers"] == R["maxInliers"]
    assert got["inlierIdx"].tolist() == R["inlierIdx"]
    assert np.linalg.norm(got["T"] - np.asarray(R["T"])) < 1e-9
    # the batched-windows call on the same window
    b = pcreg.ransac_batch([np.asarray(R["p1"])], [np.asarray(R["p2"])], R["coef"], seeds=[R["seed"]])[0]
    assert b["best"] == R["best"] and b["inlierIdx"].tolist() == R["inlierIdx"] and np.array_equal(b["T"], got["T"])


def test_local_points_against_committed_fixture(pcreg):
    G, _ = _load()
    Lp = G["local_points"]
    cloud = np.asarray(Lp["cloud"])
    for c, want in zip(Lp["centres"], Lp["results"]):
        p, d = pcreg.getLocalPoints(cloud, Lp["R"], c, Lp["min_points"], Lp["max_points"])
        if want["pts"] is None:
            assert p is None and d is None
        else:
            assert np.array_equal(p, np.asarray(want["pts"])) and np.array_equal(d, np.asarray(want["dists"]))   # bit-exact


def test_get_matches_against_committed_fixture(pcreg):
    G, _ = _load()
    M = G["matches"]
    pairs, metric = pcreg.getMatches(np.asarray(M["descSurface"]), np.asarray(M["descModel"]), M["par"], return_metric=True)
    assert np.asarray(pairs).tolist() == M["pairs"]                     # every decision of the fixture has a margin > 1e-6
    assert np.allclose(metric, np.asarray(M["metric"]), rtol=1e-12, atol=0)


def test_descriptors_against_committed_fixture(pcreg):
    """getSpacialHistogramDescriptors.m: same surviving keypoints, same 980 integer bin counts."""
    from pcreg_b200 import synth
    G, _ = _load()
    D = G["descriptors"]
    model = np.asarray(synth.make_model(D["model"][0], D["model"][1]), dtype=np.float64)
    opts = dict(D["opts"], thVar=tuple(D["opts"]["thVar"]))
    m = pcreg.Model(model)
    feat, desc = pcreg.getSpacialHistogramDescriptors(m, np.asarray(D["keypoints"]), opts)
    m.destroy()
    assert np.array_equal(feat, np.asarray(D["feat"]))
    assert np.array_equal(np.asarray(desc).astype(np.int64), np.asarray(D["desc"], dtype=np.int64))
