"""The committed fixtures of tests/golden/ (see make_golden.py there):
* the reference's literal known answers pin the CPU oracle,
* the oracle's stored iterates guard the oracle against accidental changes (and are what the GPU tests of
  test_gpu_parity.py::test_engine_matches_committed_golden_iterates compare the engine with)."""
import json
import os

import numpy as np
import pytest

from oracle import cosmo_oracle as O
from tests import golden_problems as G
from tests.golden import make_golden

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

with open(os.path.join(HERE, "reference_known_answers.json")) as _f:
    KNOWN = json.load(_f)


@pytest.mark.parametrize("case", KNOWN, ids=[c["id"] for c in KNOWN])
def test_oracle_reproduces_the_reference_known_answers(case):
    P, q, cons = getattr(G, case["problem"])()
    Pm, qm, A, b, cones = O.assemble(P, q, cons)
    res = O.solve(Pm, qm, A, b, cones, O.Settings(**(case.get("settings") or {})))
    assert res.status == case["status"], case["ref"]
    if case.get("obj") is not None:
        assert abs(res.obj_val - case["obj"]) < case["atol"], case["ref"]
    if case.get("x") is not None:
        assert np.max(np.abs(res.x - np.array(case["x"]))) < case["atol"], case["ref"]
    if case.get("rho_updates") is not None:
        assert list(res.info.rho_updates) == case["rho_updates"]


def test_committed_known_answers_match_the_generator():
    assert json.loads(json.dumps(make_golden.KNOWN)) == KNOWN


def test_oracle_still_produces_the_committed_iterates():
    stored = np.load(os.path.join(HERE, "oracle_iterates.npz"))
    fresh = make_golden.iterates()
    assert sorted(stored.files) == sorted(fresh.keys())
    for k in stored.files:
        a, b = stored[k], fresh[k]
        if k.endswith("iter_sg"):
            assert np.array_equal(a, b), k
        else:   # BLAS / LAPACK builds differ in the last bits; trajectories are stable to ~1e-13
            assert np.allclose(a, b, rtol=1e-8, atol=1e-10), k
