"""The parity measures of harness/parity.py (used by the GPU tests and by bench.py to turn "same corner set up to response
ties within 1e-5" into numbers) must themselves discriminate: cv2's own output passes, every kind of corruption is named."""
import numpy as np
import pytest

cv2 = pytest.importorskip("cv2")

from harness import parity as hp
from harness import synth


@pytest.fixture(scope="module")
def scene():
    img = synth.base_frame(21)
    eig = cv2.cornerMinEigenVal(img, 3, ksize=3)
    xy = cv2.goodFeaturesToTrack(img, 300, 0.01, 5).reshape(-1, 2)
    return img, eig, xy


def test_reference_output_is_valid_and_identical_to_itself(scene):
    img, eig, xy = scene
    par = hp.list_parity(xy, xy, eig)
    assert par["identical"] and par["set_equal"] and par["symmetric_difference"] == 0
    assert par["first_differing_rank"] is None and par["max_swapped_gap_rel"] == 0.0 and par["swaps_within_tie"]
    ok, why = hp.gftt_valid_up_to_ties(img, xy, 300, 0.01, 5, eig=eig)
    assert ok, why


def test_order_swaps_are_measured_by_their_response_gap(scene):
    img, eig, xy = scene
    e = eig[xy[:, 1].astype(int), xy[:, 0].astype(int)].astype(np.float64)
    swapped = xy.copy()
    swapped[[10, 40]] = swapped[[40, 10]]                       # two corners far apart in response trade places
    par = hp.list_parity(swapped, xy, eig)
    assert not par["identical"] and par["set_equal"] and par["first_differing_rank"] == 10
    assert par["max_swapped_gap_rel"] == pytest.approx((e[10] - e[40]) / float(eig.max()), rel=1e-6)
    assert not par["swaps_within_tie"]
    ok, why = hp.gftt_valid_up_to_ties(img, swapped, 300, 0.01, 5, eig=eig)
    assert not ok and any("order violations" in w for w in why)


def test_corruptions_are_named(scene):
    img, eig, xy = scene
    # a corner that is no local maximum
    bad = xy.copy(); bad[5, 0] += 1
    ok, why = hp.gftt_valid_up_to_ties(img, bad, 300, 0.01, 5, eig=eig)
    assert not ok and any("not candidates" in w or "order violations" in w for w in why)
    # a missing strong corner (list not truncated: fewer than max_corners)
    ok, why = hp.gftt_valid_up_to_ties(img, np.delete(xy, 3, axis=0), 1000, 0.01, 5, eig=eig)
    assert not ok and any("missing" in w for w in why)
    # two corners closer than min_dist
    near = xy.copy(); near[7] = near[6] + np.array([1, 0])
    ok, why = hp.gftt_valid_up_to_ties(img, near, 300, 0.01, 5, eig=eig)
    assert not ok
    # a different set
    par = hp.list_parity(xy[:-1], xy[1:], eig)
    assert not par["set_equal"] and par["symmetric_difference"] == 2
