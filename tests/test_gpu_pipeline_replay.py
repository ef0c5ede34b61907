"""GPU: the reference's front-end call pattern (BASELINE config 1: 1241x376, 400 features, 255x255 ROI grid,
LK 32x32 / maxLevel 4) re-driven in Python with the OpenCV plugins and with the GPU plugins: the per-frame
feature sets the pipeline would carry (integer-truncated, OpenCVLucasKanadeFM.cpp:25) must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_front_end_replay_matches_opencv_plugins(ctx):
    pytest.importorskip("cv2")
    from oracle.replay_backend import Cv2Backend
    from harness import replay
    frames = replay.synthetic_sequence(12, stream=3)
    ref = replay.run_front_end(frames, Cv2Backend(), min_tracked=400, tol=10**6)   # tol raised: re-extraction on every frame
    got = replay.run_front_end(frames, replay.GpuBackend(ctx), min_tracked=400, tol=10**6)
    n_extract = 0
    for k, ((fr, tr, er), (fg, tg, eg)) in enumerate(zip(ref, got)):
        assert tr == tg and er == eg, f"frame {k}: tracked {tr} vs {tg}, extracted {er} vs {eg}"
        assert fr.shape == fg.shape, f"frame {k}: {fr.shape} vs {fg.shape}"
        # truncation of positions that agree to 1e-4 px can differ only when a coordinate sits on an integer
        diff = np.abs(fr - fg).max(axis=1) if len(fr) else np.zeros(0)
        assert (diff > 1).sum() == 0, f"frame {k}: {int((diff > 1).sum())} features moved by more than a pixel"
        assert (diff > 0).mean() < 0.01 if len(diff) else True
        n_extract += int(er)
    assert n_extract >= 2          # the ROI-grid extraction path ran on later frames too
    assert ref[-1][1] > 100


def test_grid_rois_match_reference_tiling():
    from harness import replay
    rois = replay.grid_rois(376, 1241)
    assert len(rois) == 10                                   # 5 x 2 tiles
    assert rois[4] == (1020, 0, 221, 255) and rois[9] == (1020, 255, 221, 121)


def _run_tracker(ctx, frames, min_tracked, tol):
    rows, cols = frames[0].shape
    tr = ctx.tracker(rows, cols, win=(32, 32), max_level=4, capacity=4096, min_tracked=min_tracked, tracked_tol=tol)
    f0 = tr.init(frames[0])
    log = [(f0, len(f0), True, None)]
    for k in range(1, len(frames)):
        xy, pi, nt, ex = tr.add_frame(frames[k])
        log.append((xy, nt, ex, pi))
    tr.close()
    return log


@pytest.mark.parametrize("tol", [150, 10**6])
def test_resident_tracker_matches_call_by_call_front_end(ctx, tol):
    """pmv_tracker (images, pyramids and tracks resident, one pyramid build per frame) == the same front end driven
    call by call through pmv_gftt / pmv_lk_track == the OpenCV plugins: identical feature lists on every frame,
    including the frames where the ROI-grid re-extraction + hasNeighbor de-duplication runs."""
    pytest.importorskip("cv2")
    from oracle.replay_backend import Cv2Backend
    from harness import replay
    frames = replay.synthetic_sequence(10, stream=5)
    ref = replay.run_front_end(frames, replay.GpuBackend(ctx), min_tracked=400, tol=tol)
    cvr = replay.run_front_end(frames, Cv2Backend(), min_tracked=400, tol=tol)
    got = _run_tracker(ctx, frames, 400, tol)
    for k, ((fr, tr, er), (fg, tg, eg, pi), (fc, tc, ec)) in enumerate(zip(ref, got, cvr)):
        assert (tr, er) == (tg, eg) == (tc, ec), f"frame {k}: tracked / extracted {tr, er} vs {tg, eg} vs {tc, ec}"
        assert np.array_equal(fr, fg), f"frame {k}: resident tracker and call-by-call GPU front end differ"
        assert fc.shape == fg.shape
        d = np.abs(fc - fg).max(axis=1) if len(fc) else np.zeros(0)
        assert (d > 1).sum() == 0 and ((d > 0).mean() < 0.01 if len(d) else True)   # truncation of positions equal to 1e-4 px
        if pi is not None:
            assert (pi[:tg] >= 0).all() and (np.diff(pi[:tg]) > 0).all() and (pi[tg:] == -1).all()
    if tol == 150:
        assert not any(e for (_, _, e) in ref[1:])        # plain tracking frames: no re-extraction
    else:
        assert all(e for (_, _, e) in ref[1:])


def test_resident_tracker_errors(ctx, pmv):
    with pytest.raises(pmv.PmvError):
        ctx.tracker(0, 100)
    tr = ctx.tracker(64, 96, win=(21, 21), max_level=2, capacity=64)
    img = np.zeros((64, 96), np.uint8)
    with pytest.raises(pmv.PmvError):
        tr.add_frame(img)                                  # init first
    assert len(tr.init(img)) == 0                          # flat image: no corners
    xy, pi, nt, ex = tr.add_frame(img)
    assert len(xy) == 0 and nt == 0 and ex                 # nothing tracked -> re-extraction attempted, nothing found
    tr.close()
