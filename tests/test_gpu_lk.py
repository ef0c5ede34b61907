"""GPU parity (through the C ABI): pyramid, Scharr and pyramidal LK vs the oracle / cv2 / fixtures."""
from pathlib import Path

import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"

EDGE_PTS = np.array([[0, 0], [1240, 375], [-5, 10], [1250, 100], [600, -3], [3.5, 370.2], [1239.7, 2.1],
                     [-25, -25], [1262, 380], [620.5, 188.25]], np.float32)


def _pyr_oracle(img, levels):
    out, cur = [], img
    for _ in range(levels):
        cur = oracle.pyr_down(np.ascontiguousarray(cur))
        out.append(cur)
    return out


@pytest.mark.parametrize("shape", [(376, 1241), (47, 156), (33, 70), (129, 257), (540, 960)])
def test_pyramid_bit_exact(ctx, synth, shape):
    img = synth.base_frame(17, shape[0], shape[1])
    got = ctx.pyramid_build(img, (5, 5), 3)
    want = _pyr_oracle(img, len(got))
    assert len(got) == oracle.pyr_levels(shape[0], shape[1], 5, 5, 3)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_pyramid_strided_view(ctx, synth):
    """Non-contiguous ROI view (step = parent cols), as cv::Mat sub-views arrive in the pipeline."""
    big = synth.base_frame(18, 300, 500)
    view = big[20:275, 100:355]
    got = ctx.pyramid_build(view, (21, 21), 2)
    want = _pyr_oracle(np.ascontiguousarray(view), 2)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_pyramid_4k_checksum(ctx, synth):
    img = synth.base_frame(19, 2160, 3840)
    got = ctx.pyramid_build(img, (21, 21), 3)
    want = _pyr_oracle(img, 3)
    assert [g.shape for g in got] == [(1080, 1920), (540, 960), (270, 480)]
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_scharr_bit_exact(ctx, synth):
    for shape in [(376, 1241), (47, 156), (3, 5)]:
        img = synth.base_frame(20, shape[0], shape[1])
        assert np.array_equal(ctx.scharr(img), oracle.scharr(img))


@pytest.mark.parametrize("win,ml", [((21, 21), 3), ((32, 32), 4), ((15, 15), 2), ((9, 13), 3), ((25, 25), 3)])
def test_lk_vs_oracle_and_cv2(ctx, synth, win, ml):
    cv2 = pytest.importorskip("cv2")
    f0, f1 = synth.frame_pair(3)
    pts = np.concatenate([synth.track_points(f0, 2000, 7), EDGE_PTS])
    nx, st, err = ctx.lk_track(f0, f1, pts, win, ml)
    ox, ost, oerr = oracle.lk_track(f0, f1, pts, win, ml)
    assert np.array_equal(st, ost)                      # identical status flags
    ok = st == 1
    assert ok.sum() > 1900
    assert np.abs(nx[ok] - ox[ok]).max() < 0.01         # north_star: within 0.01 px
    assert np.abs(err[ok] - oerr[ok]).max() < 0.01
    c1, cst, cerr = cv2.calcOpticalFlowPyrLK(f0, f1, pts.reshape(-1, 1, 2), None, winSize=win, maxLevel=ml)
    assert np.array_equal(st, cst.ravel())
    assert np.abs(nx[ok] - c1.reshape(-1, 2)[ok]).max() < 0.01


def test_lk_flags(ctx, synth):
    f0, f1 = synth.frame_pair(4)
    pts = synth.track_points(f0, 300, 9)
    init = pts + np.float32([1.5, -1.0])
    nx, st, err = ctx.lk_track(f0, f1, pts, (21, 21), 3, flags=4 | 8, init=init)
    ox, ost, oerr = oracle.lk_track(f0, f1, pts, (21, 21), 3, flags=4 | 8, init=init)
    assert np.array_equal(st, ost)
    ok = st == 1
    assert np.abs(nx[ok] - ox[ok]).max() < 0.01
    assert np.allclose(err[ok], oerr[ok], rtol=1e-4, atol=1e-6)


def test_lk_flat_and_lost(ctx, synth):
    """Flat image -> minEig test fails -> status 0 everywhere; unrelated image pair -> same flags as oracle."""
    flat = np.full((120, 160), 77, np.uint8)
    pts = np.float32([[20, 20], [80, 60], [150, 110]])
    nx, st, err = ctx.lk_track(flat, flat, pts, (21, 21), 2)
    ox, ost, oerr = oracle.lk_track(flat, flat, pts, (21, 21), 2)
    assert st.sum() == 0 and np.array_equal(st, ost)
    assert np.allclose(nx, ox)
    a = synth.base_frame(31, 120, 160)
    b = synth.base_frame(32, 120, 160)
    pts = synth.track_points(a, 100, 1)
    nx, st, err = ctx.lk_track(a, b, pts, (15, 15), 2)
    ox, ost, oerr = oracle.lk_track(a, b, pts, (15, 15), 2)
    assert np.array_equal(st, ost)
    ok = st == 1
    # unrelated images: Newton can bounce, compare only where both report a converged track
    assert np.abs(nx[ok] - ox[ok]).max() < 0.05 or ok.sum() == 0


def test_lk_empty_and_errors(ctx, pmv, synth):
    f0, f1 = synth.frame_pair(5, h=100, w=120)
    nx, st, err = ctx.lk_track(f0, f1, np.zeros((0, 2), np.float32), (21, 21), 3)
    assert nx.shape == (0, 2)
    with pytest.raises(pmv.PmvError):
        ctx.lk_track(f0, f1, np.float32([[5, 5]]), (2, 2), 3)      # OpenCV asserts winSize > 2
    with pytest.raises(pmv.PmvError):
        ctx.lk_track(f0, f1, np.float32([[5, 5]]), (41, 41), 3)    # unsupported window


def test_lk_golden_fixture(ctx):
    g = np.load(GOLD / "lk_small.npz")
    nx, st, err = ctx.lk_track(g["prev"], g["next"], g["pts"], tuple(int(v) for v in g["win"]), int(g["max_level"]))
    assert np.array_equal(st, g["status"])
    ok = st == 1
    assert np.abs(nx[ok] - g["next_xy"][ok]).max() < 0.01
    assert np.abs(err[ok] - g["err"][ok]).max() < 0.01


def test_lk_batched_matches_single(ctx, synth):
    B, n = 6, 400
    prev = np.stack([synth.frame_pair(40 + b)[0] for b in range(B)])
    nxt = np.stack([synth.frame_pair(40 + b)[1] for b in range(B)])
    pts = np.stack([synth.track_points(prev[b], n, b) for b in range(B)])
    nx, st, err = ctx.lk_track_batched(prev, nxt, pts, (21, 21), 3)
    for b in range(B):
        ox, ost, oerr = oracle.lk_track(prev[b], nxt[b], pts[b], (21, 21), 3)
        assert np.array_equal(st[b], ost)
        ok = ost == 1
        assert np.abs(nx[b][ok] - ox[ok]).max() < 0.01


def test_lk_batched_dev_full_size_properties(ctx, synth):
    """BASELINE config-2 shape (reduced batch): device-resident call; property checks that do not
    need the oracle at full size -- identity pair tracks to itself with err 0."""
    torch = pytest.importorskip("torch")
    B, n = 16, 2000
    f0 = synth.base_frame(77)
    pts = synth.track_points(f0, n, 3)
    d_img = torch.from_numpy(np.stack([f0] * B)).cuda()
    d_pts = torch.from_numpy(np.stack([pts] * B)).cuda()
    d_nx = torch.zeros(B, n, 2, device="cuda")
    d_st = torch.zeros(B, n, dtype=torch.uint8, device="cuda")
    d_err = torch.zeros(B, n, device="cuda")
    stream = torch.cuda.Stream()
    stream.wait_stream(torch.cuda.current_stream())
    ctx.set_stream(stream.cuda_stream)
    ctx.lk_track_batched_dev(d_img.data_ptr(), d_img.data_ptr(), B, 376 * 1241, 376, 1241, 1241, d_pts.data_ptr(), n,
                             d_nx.data_ptr(), d_st.data_ptr(), d_err.data_ptr(), (21, 21), 3)
    torch.cuda.synchronize()
    ctx.set_stream(None)
    assert int(d_st.sum()) == B * n
    assert float((d_nx - d_pts).abs().max()) < 1e-3
    assert float(d_err.abs().max()) < 0.01   # fixed-point resampling noise only
