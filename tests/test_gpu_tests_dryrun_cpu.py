"""Dry run of GPU test bodies on the CPU: the CUDA engine is replaced by the oracle-backed stand-in of
tests/oracle_engine.py and the functions of tests/test_gpu_parity.py are called directly.  What this checks is the
Python side of those tests (imports, helpers, fixtures, the host glue they drive) -- a NameError in a GPU test would
otherwise only show up on the next GPU run.  Assertion failures are tolerated where the stand-in legitimately differs
from the engine (it ignores the D/E unscaling of the termination test); every other exception fails the test."""
import importlib
import inspect

import pytest

from cosmo_b200 import engine as E, model as M
from tests.oracle_engine import OracleEngine

CASES = ["test_engine_matches_committed_golden_iterates", "test_g6_chordal_sdp_through_the_clique_batch",
         "test_g4_g5_g11_literal_problems", "test_g15_g16_exp_pow_cone_problems", "test_project_exp_pow_cones",
         "test_accelerated_iterates_match_oracle", "test_accelerator_rho_adaption_limits", "test_g1_simple_qp",
         "test_g2_box_statuses", "test_g3_hs21_with_soc_and_merging", "test_g14_model_updates_and_warm_start",
         "test_project_composite_matches_oracle", "test_soc_branches", "test_complex_psd_cone_projection_and_least_eigenvalue",
         "test_project_psd_sign_function_path"]
MUST_PASS = {"test_engine_matches_committed_golden_iterates", "test_g6_chordal_sdp_through_the_clique_batch",
             "test_project_exp_pow_cones", "test_soc_branches", "test_complex_psd_cone_projection_and_least_eigenvalue"}


def _calls(fn):
    """argument tuples of a (possibly multiply) parametrized test function: first value of every parameter"""
    marks = [m for m in getattr(fn, "pytestmark", []) if m.name == "parametrize"]
    kwargs = {}
    for m in marks:
        names = [a.strip() for a in m.args[0].split(",")]
        first = m.args[1][0]
        first = first if isinstance(first, (tuple, list)) and len(names) > 1 else (first,)
        kwargs.update(dict(zip(names, first)))
    return kwargs


@pytest.mark.parametrize("name", CASES)
def test_gpu_test_body_runs_against_the_oracle_stand_in(name, monkeypatch):
    monkeypatch.setattr(M._eng, "Engine", OracleEngine)
    monkeypatch.setattr(E, "Engine", OracleEngine)
    monkeypatch.setenv("COSMO_B200_TEST_EXPERIMENTAL", "1")
    T = importlib.import_module("tests.test_gpu_parity")
    fn = getattr(T, name)
    kwargs = _calls(fn)
    if "monkeypatch" in inspect.signature(fn).parameters:
        kwargs["monkeypatch"] = monkeypatch
    try:
        fn(**kwargs)
    except AssertionError:
        if name in MUST_PASS:
            raise
