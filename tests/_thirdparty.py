"""Inputs and comparison rule of the third-party golden fixtures (tests/golden/thirdparty_topk.json), shared by the
generator script (scripts/make_thirdparty_golden.py) and the CPU / GPU tests.  Test infrastructure."""
import json
import os

import numpy as np

from ragfin_b200.synthetic import synth_rows   # the product's generator: pure numpy integer hashing, exact in fp32


def raw_inputs(seed, n, dim, nq, scale_seed):
    """Rows / queries with norms all over the place: a COSINE collection is fed un-normalised vectors."""
    rng = np.random.default_rng(scale_seed)
    x = (synth_rows(seed, 0, n, dim) * rng.uniform(0.1, 30.0, size=(n, 1))).astype(np.float32)
    q = (synth_rows(seed + 1, 0, nq, dim) * rng.uniform(0.5, 4.0, size=(nq, 1))).astype(np.float32)
    return x, q


def cases():
    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "thirdparty_topk.json")) as f:
        return json.load(f)["cases"]


def check(case, ids, sims):
    """ids / sims [kept queries, k] of an implementation under test against one committed third-party case."""
    want_ids, want_sims = np.asarray(case["ids"]), np.asarray(case["sims"])
    assert ids.shape == want_ids.shape
    if case["strict"]:
        assert np.array_equal(ids, want_ids), case["name"]
        assert np.allclose(sims, want_sims, rtol=1e-5, atol=1e-6), case["name"]      # north_star: 1e-5 relative for fp32
        return
    for r in range(ids.shape[0]):                                  # large k: set + order wherever the gaps are clear
        assert set(ids[r].tolist()) == set(want_ids[r].tolist()), (case["name"], r)
        assert np.allclose(np.sort(sims[r])[::-1], want_sims[r], rtol=1e-5, atol=1e-6)
        gaps = -np.diff(want_sims[r])
        clear = np.concatenate([[True], gaps > 1e-6]) & np.concatenate([gaps > 1e-6, [True]])
        assert np.array_equal(ids[r][clear], want_ids[r][clear]), (case["name"], r)
