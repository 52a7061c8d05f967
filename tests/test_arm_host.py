"""Host-side shell of the arm API (no GPU): rrtk.NLinkArm against values the unmodified reference produced
(arm02:236-262; generated with oracle/ref_loader.load('arm02') in the build container)."""
import numpy as np

import rrtk


def test_nlinkarm_matches_reference_points():
    arm = rrtk.NLinkArm([0.5, 0.5, 0.3, 0.5, 0.1], [0.0] * 5)
    assert arm.n_links == 5 and arm.lim == sum([0.5, 0.5, 0.3, 0.5, 0.1])
    assert np.allclose(np.array(arm.points)[:, 1], 0.0) and arm.points[5][0] == 0.5 + 0.5 + 0.3 + 0.5 + 0.1
    # the script drives the 5-link arm with TWO joint angles (arm02:98): links 2..5 are collinear along a0 + a1
    arm.update_joints([0.7, -1.9])
    ref = [[0, 0], [0.38242109364224425, 0.3221088436188455], [0.563599970880581, -0.14391069936476764],
           [0.6723072972235831, -0.4235224251549355], [0.8534861744619199, -0.8895419681385486],
           [0.8897219499095872, -0.9827458767352713]]
    assert np.array_equal(np.array(arm.points, dtype=float), np.array(ref, dtype=float))
    assert np.array_equal(arm.end_effector, np.array(ref[5]))


def test_nlinkarm_rejects_mismatched_lengths():
    import pytest
    with pytest.raises(ValueError):
        rrtk.NLinkArm([1.0, 1.0], [0.0])
