"""GPU parity of the bundle adjuster (through the C ABI) against the fp64 oracle.  Bars (north_star):
residuals / Jacobians within 1e-5 relative, final cost within 1e-6 relative for the same iteration
count.  (The oracle itself is unpinned against Ceres -- see tests/test_oracle_ba.py.)"""
import numpy as np
import pytest

import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["window", "general"], autouse=True)
def ba_path(request, monkeypatch):
    """Run every solve twice: the window-batched kernels (Nc <= 22) and the general multi-kernel path."""
    if request.param == "general":
        monkeypatch.setenv("PMV_BA_FORCE_GENERAL", "1")
        monkeypatch.delenv("PMV_BA_FORCE_WINDOW", raising=False)
    else:
        monkeypatch.delenv("PMV_BA_FORCE_GENERAL", raising=False)
        monkeypatch.setenv("PMV_BA_FORCE_WINDOW", "1")   # a lone window defaults to the general path
    return request.param


def _args(w):
    return (w["poses"], w["points"], w["obs"], w["cam_idx"], w["pt_idx"], w["K"])


def test_residual_jacobian_vs_oracle(ctx, synth):
    w = synth.ba_window(1, n_poses=8, n_points=300)
    r, Jc, Jp, cost = ctx.ba_eval(*_args(w))
    ro, Jco, Jpo, costo = oracle.ba_eval(*_args(w))
    assert np.abs(r - ro).max() <= 1e-5 * np.abs(ro).max()
    assert np.abs(Jc - Jco).max() <= 1e-5 * np.abs(Jco).max()
    assert np.abs(Jp - Jpo).max() <= 1e-5 * np.abs(Jpo).max()
    # per-entry relative check where entries are not tiny (analytic vs Jets agree to ~1e-12)
    big = np.abs(Jco) > 1e-3 * np.abs(Jco).max()
    assert np.abs(Jc[big] / Jco[big] - 1).max() < 1e-9
    assert abs(cost - costo) <= 1e-10 * costo


def test_small_angle_and_no_loss(ctx):
    K = np.array([[700., 0, 300], [0, 700, 200], [0, 0, 1]])
    poses = np.array([[0, 0, 0, .1, -.2, .3], [1e-9, -2e-9, 5e-10, 0, 0, 0], [.3, -.2, .1, .2, .1, 0]])
    pts = np.array([[1.0, -0.5, -12.0], [-2, 1, -30.]])
    cam = np.array([0, 1, 2, 0, 1, 2], np.int32); pt = np.array([0, 0, 0, 1, 1, 1], np.int32)
    obs = np.array([[320., 190.]] * 6)
    for delta in (1.0, 0.0):
        got = ctx.ba_eval(poses, pts, obs, cam, pt, K, delta)
        want = oracle.ba_eval(poses, pts, obs, cam, pt, K, delta)
        for g, wv in zip(got[:3], want[:3]):
            assert np.allclose(g, wv, rtol=1e-10, atol=1e-9)
        assert np.isclose(got[3], want[3], rtol=1e-12)


@pytest.mark.parametrize("npose,npts,iters", [(5, 150, 5), (6, 80, 10), (3, 600, 5), (20, 400, 5)])
def test_solve_vs_oracle(ctx, synth, npose, npts, iters):
    w = synth.ba_window(2 + npose, n_poses=npose, n_points=npts)
    p, x, s = ctx.ba_solve(*_args(w), 1.0, iters)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, iters)
    assert s["iterations"] == so["iterations"] and s["successful_steps"] == so["successful_steps"]
    assert s["termination"] == so["termination"]
    assert abs(s["initial_cost"] - so["initial_cost"]) <= 1e-10 * so["initial_cost"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]        # north_star bar
    assert np.abs(p - po).max() < 1e-5 and np.abs(x - xo).max() < 1e-4


def test_solve_pipeline_shape_sparse_visibility(ctx, synth):
    """Reference shape: <=5 poses, a few hundred points each seen in 2-5 frames, 5 iterations."""
    rng = np.random.default_rng(0)
    w = synth.ba_window(77, n_poses=5, n_points=400)
    keep = rng.random(len(w["obs"])) < 0.6
    a = (w["poses"], w["points"], w["obs"][keep], w["cam_idx"][keep], w["pt_idx"][keep], w["K"])
    p, x, s = ctx.ba_solve(*a, 1.0, 5)
    po, xo, so = oracle.ba_solve(*a, 1.0, 5)
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
    assert s["iterations"] == so["iterations"]
    unseen = np.setdiff1d(np.arange(400), w["pt_idx"][keep])
    assert np.array_equal(x[unseen], w["points"][unseen])          # blocks without residuals untouched


def test_rejected_steps_follow_oracle(ctx, synth):
    """Start far from the optimum so the LM loop rejects steps and shrinks the radius."""
    w = synth.ba_window(5, n_poses=4, n_points=60)
    rng = np.random.default_rng(1)
    poses = w["poses"] + rng.normal(0, 0.05, w["poses"].shape)
    pts = w["points"] + rng.normal(0, 2.0, w["points"].shape)
    a = (poses, pts, w["obs"], w["cam_idx"], w["pt_idx"], w["K"])
    p, x, s = ctx.ba_solve(*a, 1.0, 12)
    po, xo, so = oracle.ba_solve(*a, 1.0, 12)
    assert s["iterations"] == so["iterations"] and s["successful_steps"] == so["successful_steps"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]


def test_converged_problem_terminates_like_oracle(ctx, synth):
    w = synth.ba_window(3, n_poses=5, n_points=60, outlier_frac=0.0)
    obs = np.concatenate([synth.project(w["poses_true"][c], w["points_true"][[p]], w["K"])[0]
                          for c, p in zip(w["cam_idx"], w["pt_idx"])])
    a = (w["poses"], w["points"], obs, w["cam_idx"], w["pt_idx"], w["K"])
    p, x, s = ctx.ba_solve(*a, 1.0, 50)
    po, xo, so = oracle.ba_solve(*a, 1.0, 50)
    assert s["termination"] in (1, 2, 3) and s["final_cost"] < 1e-9
    assert abs(s["iterations"] - so["iterations"]) <= 1        # at the 1e-16 cost floor the last test may flip


def test_full_size_window_config4_shape(ctx, synth, ba_path):
    """One BASELINE config-4 window (20 poses, 2 000 points, ~35 k observations), 5 iterations."""
    if ba_path == "general":
        pytest.skip("covered by the smaller shapes; the general path is atomics-bound at this density")
    w = synth.ba_window(2)
    p, x, s = ctx.ba_solve(*_args(w), 1.0, 5)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, 5)
    assert s["iterations"] == so["iterations"] and s["successful_steps"] == so["successful_steps"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
    assert np.abs(p - po).max() < 1e-5


def test_batched_windows(ctx, synth):
    ws = [synth.ba_window(10 + i, n_poses=6, n_points=120) for i in range(5)]
    off = np.cumsum([0] + [len(w["obs"]) for w in ws])
    P, X, S = ctx.ba_solve_batched(np.stack([w["poses"] for w in ws]), np.stack([w["points"] for w in ws]),
                                   np.concatenate([w["obs"] for w in ws]), np.concatenate([w["cam_idx"] for w in ws]),
                                   np.concatenate([w["pt_idx"] for w in ws]), off, ws[0]["K"], 1.0, 5)
    for i, w in enumerate(ws):
        po, xo, so = oracle.ba_solve(*_args(w), 1.0, 5)
        assert abs(S[i]["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
        assert S[i]["iterations"] == so["iterations"]
        assert np.abs(P[i] - po).max() < 1e-5


@pytest.mark.parametrize("order", ["camera_major", "point_major", "shuffled"])
def test_many_windows_any_observation_order(ctx, synth, order):
    """Problem creation sorts an unordered list of >= 64 windows window by window (stable counting sort by point; the
    camera-major order CeresBundleAdjustment::apply produces needs no per-point sort), takes a (point, camera)-ordered
    list in place, and insertion-sorts the points of an arbitrary order: every order gives the oracle's solve."""
    nw = 70
    ws = [synth.ba_window(200 + i, n_poses=5, n_points=90) for i in range(nw)]
    rng = np.random.default_rng(11)
    obs, cam, pt = [], [], []
    for w in ws:
        n = len(w["cam_idx"])
        if order == "camera_major":
            idx = np.lexsort((w["pt_idx"], w["cam_idx"]))
        elif order == "point_major":
            idx = np.lexsort((w["cam_idx"], w["pt_idx"]))
        else:
            idx = rng.permutation(n)
        obs.append(w["obs"][idx]); cam.append(w["cam_idx"][idx]); pt.append(w["pt_idx"][idx])
    off = np.cumsum([0] + [len(c) for c in cam])
    P, X, S = ctx.ba_solve_batched(np.stack([w["poses"] for w in ws]), np.stack([w["points"] for w in ws]),
                                   np.concatenate(obs), np.concatenate(cam), np.concatenate(pt), off, ws[0]["K"], 1.0, 5)
    for i in (0, 1, 33, nw - 1):
        po, xo, so = oracle.ba_solve(*_args(ws[i]), 1.0, 5)
        assert abs(S[i]["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"], (order, i)
        assert S[i]["iterations"] == so["iterations"]
        assert np.abs(P[i] - po).max() < 1e-5


def test_large_reduced_system_uses_blocked_cholesky(ctx, synth):
    """n = 6*40 = 240 > 160: the blocked HBM Cholesky (DMMA trailing update) instead of the smem one."""
    w = synth.ba_large(3, n_poses=40, n_points=3000, views=5, span=20)
    p, x, s = ctx.ba_solve(*_args(w), 1.0, 4)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, 4)
    assert s["iterations"] == so["iterations"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
    assert np.abs(p - po).max() < 1e-5


@pytest.mark.parametrize("nposes,npts,span", [(40, 3000, 20), (100, 6000, 40), (30, 1500, 30), (67, 4000, 12), (200, 8000, 40)])
def test_banded_cholesky_cluster_kernel(ctx, synth, monkeypatch, ba_path, nposes, npts, span):
    """Banded reduced systems are factorised by one cluster launch (ba_chol_band.cu); it must agree with the
    oracle and with the multi-launch blocked Cholesky (PMV_CHOL_NO_CLUSTER=1) on the same problem."""
    if ba_path == "window":
        pytest.skip("n > 160: one path")
    w = synth.ba_large(11 + nposes, n_poses=nposes, n_points=npts, views=5, span=span)
    monkeypatch.delenv("PMV_CHOL_NO_CLUSTER", raising=False)
    p, x, s = ctx.ba_solve(*_args(w), 1.0, 4)
    monkeypatch.setenv("PMV_CHOL_NO_CLUSTER", "1")
    p2, x2, s2 = ctx.ba_solve(*_args(w), 1.0, 4)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, 4)
    assert s["iterations"] == so["iterations"] == s2["iterations"]
    assert abs(s["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
    assert abs(s["final_cost"] - s2["final_cost"]) <= 1e-9 * so["final_cost"]
    assert np.abs(p - po).max() < 1e-5
    assert np.abs(x - xo).max() < 1e-4


@pytest.mark.parametrize("nposes,npts,span", [(40, 3000, 20), (12, 900, 12)])
def test_pair_list_schur_matches_point_kernel(ctx, pmv, synth, monkeypatch, ba_path, nposes, npts, span):
    """Large single problems eliminate the points block by block of S (ba_pair_schur_kernel, per-camera-pair
    entry lists) instead of point by point with atomics; forced here on a small problem, both must agree with
    the oracle.  A resident problem is used because one-shot solves of small problems live in the arena."""
    if ba_path == "window":
        pytest.skip("general path only")
    w = synth.ba_large(5 + nposes, n_poses=nposes, n_points=npts, views=5, span=span)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, 4)
    res = {}
    for mode in ("pairs", "points"):
        if mode == "pairs":
            monkeypatch.setenv("PMV_BA_FORCE_PAIRS", "1")
        else:
            monkeypatch.delenv("PMV_BA_FORCE_PAIRS", raising=False)
        prob = ctx.ba_problem(*_args(w), 1.0)
        prob.solve(4)
        p, x, s = prob.download()
        prob.close()
        res[mode] = (p, x, s[0])
        assert s[0]["iterations"] == so["iterations"]
        assert abs(s[0]["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
        assert np.abs(p[0] - po).max() < 1e-5
    assert abs(res["pairs"][2]["final_cost"] - res["points"][2]["final_cost"]) <= 1e-9 * so["final_cost"]


@pytest.mark.parametrize("nposes,npts,span,views", [(40, 3000, 20, 5), (12, 900, 12, 5), (30, 2500, 16, 3), (24, 1500, 24, 8)])
def test_run_organised_schur_matches_oracle(ctx, pmv, synth, monkeypatch, ba_path, nposes, npts, span, views):
    """BAL-scale problems walk RUNS of points that share their camera tuple (ba_runs.cuh): residuals and Jacobians are
    recomputed per observation, the blocks of a tuple are kept in registers and S is touched once per run.  Forced
    here on small problems (tuple sizes 1..8 occur near the ends of the trajectory); compared with the oracle, with
    the point-by-point kernel, and on a problem with rejected steps."""
    if ba_path == "window":
        pytest.skip("general path only")
    w = synth.ba_large(11 + nposes, n_poses=nposes, n_points=npts, views=views, span=span)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, 4)
    res = {}
    for mode in ("runs", "points"):
        if mode == "runs":
            monkeypatch.setenv("PMV_BA_FORCE_RUNS", "1")
        else:
            monkeypatch.delenv("PMV_BA_FORCE_RUNS", raising=False)
        prob = ctx.ba_problem(*_args(w), 1.0)
        prob.solve(4)
        p, x, s = prob.download()
        prob.close()
        res[mode] = (p, x, s[0])
        assert s[0]["iterations"] == so["iterations"] and s[0]["successful_steps"] == so["successful_steps"]
        assert abs(s[0]["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]
        assert abs(s[0]["initial_cost"] - so["initial_cost"]) <= 1e-9 * so["initial_cost"]
        assert np.abs(p[0] - po).max() < 1e-5 and np.abs(x[0] - xo).max() < 1e-4
    assert abs(res["runs"][2]["final_cost"] - res["points"][2]["final_cost"]) <= 1e-9 * so["final_cost"]
    monkeypatch.delenv("PMV_BA_FORCE_RUNS", raising=False)


def test_run_organised_rejected_steps(ctx, synth, monkeypatch):
    """A badly initialised problem makes LM reject steps: the run path re-solves with the same linearisation
    (recomputed, not stored) and must follow the oracle's accept / reject sequence."""
    monkeypatch.setenv("PMV_BA_FORCE_RUNS", "1")
    w = synth.ba_large(3, n_poses=16, n_points=1200, views=4, span=12)
    rng = np.random.default_rng(0)
    pts = w["points"] + rng.normal(0, 2.0, w["points"].shape)
    args = (w["poses"], pts, w["obs"], w["cam_idx"], w["pt_idx"], w["K"])
    po, xo, so = oracle.ba_solve(*args, 1.0, 8)
    prob = ctx.ba_problem(*args, 1.0)
    prob.solve(8)
    p, x, s = prob.download()
    prob.close()
    assert s[0]["iterations"] == so["iterations"] and s[0]["successful_steps"] == so["successful_steps"]
    assert abs(s[0]["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"]


@pytest.mark.parametrize("nposes,npts,span", [(400, 20000, 20), (250, 12000, 30), (700, 30000, 24)])
def test_partitioned_and_two_sided_band_solves_match_one_sided_and_oracle(ctx, pmv, synth, monkeypatch, ba_path, nposes, npts, span):
    """Long banded reduced camera systems are not factorised as one chain of block steps: by default the band is cut once
    in the middle and the leading and the index-reversed trailing system are factorised by two clusters (BASplit); with
    PMV_CHOL_PARTS=P >= 3 (opt-in) P segments separated by P - 1 separators are eliminated concurrently (spikes for the
    left couplings, block-tridiagonal separator system: BAPart).  Same iterates as the one-sided factorisation and as
    the oracle."""
    if ba_path == "window":
        pytest.skip("general path only")
    w = synth.ba_large(31, n_poses=nposes, n_points=npts, views=5, span=span)
    po, xo, so = oracle.ba_solve(*_args(w), 1.0, 3)
    res = {}
    for mode in ("part", "part3", "split", "one_sided"):
        monkeypatch.delenv("PMV_CHOL_NO_SPLIT", raising=False)
        monkeypatch.delenv("PMV_CHOL_PARTS", raising=False)
        if mode == "part3":
            monkeypatch.setenv("PMV_CHOL_PARTS", "3")
        elif mode == "part":
            monkeypatch.setenv("PMV_CHOL_PARTS", "4")
        elif mode == "one_sided":
            monkeypatch.setenv("PMV_CHOL_NO_SPLIT", "1")
        prob = ctx.ba_problem(*_args(w), 1.0)
        prob.solve(3)
        p, x, s = prob.download()
        prob.close()
        res[mode] = (p, s[0])
        assert s[0]["iterations"] == so["iterations"] and s[0]["successful_steps"] == so["successful_steps"], mode
        assert abs(s[0]["final_cost"] - so["final_cost"]) <= 1e-6 * so["final_cost"], mode
        assert np.abs(p[0] - po).max() < 1e-5, mode
    monkeypatch.delenv("PMV_CHOL_NO_SPLIT", raising=False)
    monkeypatch.delenv("PMV_CHOL_PARTS", raising=False)
    for mode in ("part", "part3", "split"):
        assert np.abs(res[mode][0] - res["one_sided"][0]).max() < 1e-8, mode
        assert abs(res[mode][1]["final_cost"] - res["one_sided"][1]["final_cost"]) <= 1e-10 * so["final_cost"], mode


@pytest.mark.parametrize("runs", [False, True])
def test_observation_order_does_not_matter(ctx, synth, monkeypatch, ba_path, runs):
    """Problem creation takes the caller's list as it is when it is already sorted by (point, camera) -- the order
    exporters and the generators here produce -- and sorts it otherwise: a shuffled list must give the same solve."""
    if runs:
        if ba_path == "window":
            pytest.skip("general path only")
        monkeypatch.setenv("PMV_BA_FORCE_RUNS", "1")
        w = synth.ba_large(5, n_poses=30, n_points=2500, views=5, span=16)
    else:
        w = synth.ba_window(9, n_poses=6, n_points=300)
    rng = np.random.default_rng(3)
    order = rng.permutation(len(w["cam_idx"]))
    out = []
    for obs, cam, pt in ((w["obs"], w["cam_idx"], w["pt_idx"]), (w["obs"][order], w["cam_idx"][order], w["pt_idx"][order])):
        prob = ctx.ba_problem(w["poses"], w["points"], np.ascontiguousarray(obs), np.ascontiguousarray(cam), np.ascontiguousarray(pt), w["K"], 1.0)
        prob.solve(4)
        p, x, s = prob.download()
        prob.close()
        out.append((p[0], x[0], s[0]))
    (p0, x0, s0), (p1, x1, s1) = out
    assert s0["iterations"] == s1["iterations"] and s0["successful_steps"] == s1["successful_steps"]
    assert abs(s0["final_cost"] - s1["final_cost"]) <= 1e-10 * s0["final_cost"]
    assert np.abs(p0 - p1).max() < 1e-8 and np.abs(x0 - x1).max() < 1e-5   # fp64 atomics reorder the sums of weakly constrained points
    monkeypatch.delenv("PMV_BA_FORCE_RUNS", raising=False)


def test_resident_problem_reset_and_errors(ctx, pmv, synth):
    w = synth.ba_window(9, n_poses=5, n_points=100)
    prob = ctx.ba_problem(*_args(w))
    prob.solve(5)
    p1, x1, s1 = prob.download()
    prob.reset(); prob.solve(5)
    p2, x2, s2 = prob.download()
    assert np.allclose(p1, p2, atol=1e-9) and abs(s1[0]["final_cost"] - s2[0]["final_cost"]) <= 1e-9 * s1[0]["final_cost"]
    assert prob.device_bytes > 0
    prob.close()
    with pytest.raises(pmv.PmvError):
        ctx.ba_solve(w["poses"], w["points"], w["obs"], w["cam_idx"] + 100, w["pt_idx"], w["K"])
