"""CPU: the plain-C oracle against the committed golden vectors (outputs of the reference's own code, tests/golden/make_golden.py).
This is what pins the oracle on machines without /root/reference (the GPU box)."""
import numpy as np
import pytest

from common import GoldenCase, assert_frames_equal, golden_cases

CASES = golden_cases()


def test_golden_files_present():
    assert len(CASES) >= 5


@pytest.mark.parametrize("path", CASES, ids=[p.stem for p in CASES])
def test_oracle_matches_reference_golden(oracle_built, path):
    case = GoldenCase(path)
    orun = case.oracle_run()
    got = orun.views
    # the oracle's 'visible' uses 0xFF for "not written by the reference"; the golden file holds the reference's bytes (0/1)
    for g, w in zip(got, case.frames):
        if "visible" in g and "visible" in w:
            for k in range(len(g["visible"])):
                wrote = g["visible"][k] != 0xFF
                assert np.array_equal(g["visible"][k][wrote], w["visible"][k][wrote]), f"{case.name}: isVisible pool {k}"
                assert not w["visible"][k][~wrote].any()
            del g["visible"]
    assert_frames_equal(got, case.frames, case.render_types, case.name)
