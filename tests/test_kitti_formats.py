"""CPU: KITTI wire formats + error report (host/pmv_kitti.h through the C ABI) against the Python
restatement of OdometryPipeline.cpp:497-669, 272-300 in oracle/kitti.py, on well-formed files and on the odd input the
reference's tokenisation reacts to (double spaces, short lines, text tokens, trailing blanks, no final newline)."""
import numpy as np
import pytest

from oracle import kitti as ok


def _write_poses(path, R, t, fmt="%.6e", sep=" ", trailing=""):
    with open(path, "w") as f:
        for Rk, tk in zip(R, t):
            row = np.c_[Rk, tk].ravel()
            f.write(sep.join(fmt % v for v in row) + trailing + "\n")


def _poses(n, seed=0):
    rng = np.random.default_rng(seed)
    R = np.stack([np.linalg.qr(rng.normal(size=(3, 3)))[0] for _ in range(n)])
    t = np.cumsum(rng.normal(0, 1, (n, 3)), 0)
    return R, t


def test_parse_poses(pmv, tmp_path):
    R, t = _poses(40)
    p = tmp_path / "00.txt"
    _write_poses(p, R, t)
    Rg, tg = pmv.kitti_parse_poses(p)
    Ro, to = ok.parse_poses(p, 1 << 30)
    assert np.array_equal(Rg.reshape(-1, 9), np.array(Ro)) and np.array_equal(tg, np.array(to))
    assert np.abs(Rg - R).max() < 1e-6 and np.abs(tg - t).max() < 1e-4
    Rs, ts = pmv.kitti_parse_poses(p, stop=7)                       # `stop` frames only (:537)
    assert len(Rs) == 7 and np.array_equal(Rs, Rg[:7])
    with pytest.raises(pmv.PmvError):
        pmv.kitti_parse_poses(tmp_path / "missing.txt")


def test_parse_poses_odd_input(pmv, tmp_path):
    p = tmp_path / "odd.txt"
    p.write_text("1 0  0 5 0 1 0 6 0 0 1 7 99 98\n"                 # double space, extra tokens
                 "1 2 3\n"                                          # short line: the rest stays 0
                 "\n"                                               # blank line: a zero pose (getline returns it)
                 "a 1e-3 -2.5x 4 . 6 7 8 9 10 11 12 \n"             # text tokens parse as 0 / numeric prefix
                 "1 2 3 4 5 6 7 8 9 10 11 12")                      # no final newline
    Rg, tg = pmv.kitti_parse_poses(p)
    Ro, to = ok.parse_poses(p, 1 << 30)
    assert len(Rg) == 5 and not Rg[2].any()
    assert np.array_equal(Rg.reshape(-1, 9), np.array(Ro)) and np.array_equal(tg, np.array(to))
    assert np.array_equal(tg[0], [5, 6, 7]) and np.array_equal(Rg[1].ravel(), [1, 2, 3, 0, 0, 0, 0, 0, 0])
    assert Rg[3, 0, 0] == 0 and Rg[3, 0, 1] == 1e-3 and Rg[3, 0, 2] == -2.5


def test_parse_calibration(pmv, tmp_path):
    p = tmp_path / "calib.txt"
    P = [[718.856, 0.0, 607.1928, 0.0, 0.0, 718.856, 185.2157, 0.0, 0.0, 0.0, 1.0, 0.0],
         [718.856, 0.0, 607.1928, -386.1448, 0.0, 718.856, 185.2157, 0.0, 0.0, 0.0, 1.0, 0.0]]
    p.write_text("".join("P%d: " % i + " ".join("%.12e" % v for v in row) + "\n" for i, row in enumerate(P)))
    for num in (0, 1):
        K = pmv.kitti_parse_calibration(p, num)
        assert np.array_equal(K.ravel(), np.array(ok.parse_calibration(p, num)))
        assert np.allclose(K, np.array(P[num]).reshape(3, 4)[:, :3], rtol=1e-12)
    # a line that stops early leaves the rest of K as it was; a missing line leaves K untouched
    q = tmp_path / "short.txt"
    q.write_text("P0: 1 2 3 4 5 6\n")
    K0 = np.arange(9.0) + 100
    K = pmv.kitti_parse_calibration(q, 0, K0)
    assert np.array_equal(K.ravel(), np.array(ok.parse_calibration(q, 0, K0)))
    assert np.array_equal(K.ravel(), [1, 2, 3, 5, 104, 105, 106, 107, 108])      # "6" has no space after it
    assert np.array_equal(pmv.kitti_parse_calibration(q, 3, K0).ravel(), K0)
    with pytest.raises(pmv.PmvError):
        pmv.kitti_parse_calibration(tmp_path / "missing.txt", 0)


@pytest.mark.parametrize("init_offset", [0, 2])
def test_error_report(pmv, init_offset):
    n = 25
    gR, gt = _poses(n + init_offset, 3)
    rng = np.random.default_rng(5)
    R = gR[:n] + rng.normal(0, 1e-2, (n, 3, 3)); t = gt[:n] + rng.normal(0, 0.1, (n, 3))
    stats, text = pmv.kitti_error_report(R, t, gR, gt, init_offset, runtime=12.5)
    want = ok.error_report(R.reshape(n, 9).tolist(), t.tolist(), gR.reshape(-1, 9).tolist(), gt.tolist(), init_offset)
    for k, v in want.items():
        assert stats[k] == pytest.approx(v, rel=1e-13), k
    lines = text.strip().split("\n")
    assert lines[0] == "Runtime: 12.5" and [l.split(":")[0] for l in lines[1:]] == ["R total", "R min", "R max", "R std", "t total", "t min",
                                                                                     "t max", "t std"]
    assert float(lines[1].split(": ")[1]) == pytest.approx(want["R_total"], rel=1e-5)      # default ostream precision: 6 digits
