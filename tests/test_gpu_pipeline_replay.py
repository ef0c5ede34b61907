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
