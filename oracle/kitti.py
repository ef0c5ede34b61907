"""oracle/kitti.py -- TEST INFRASTRUCTURE ONLY: Python restatement of the reference's KITTI text handling, the checker of
host/pmv_kitti.h / pmv_kitti_*.  Parity unpinned against a compiled reference (OdometryPipeline.cpp needs dlib's GUI and
OpenCV highgui and is unbuildable here); it follows the source line by line instead:
OdometryPipeline::split (OdometryPipeline.cpp:497-520), ::parsePoses (:525-593), ::parseCalibration (:595-653),
::standardDeviation (:657-669) and the error loop / report of ::run (:272-300)."""
import math
import re


def _to_double(tok):
    # std::stringstream >> double: the longest numeric prefix, 0 when there is none
    m = re.match(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?)", tok)
    return float(m.group(0)) if m else 0.0


def _getlines(text):
    # std::getline: a final newline does not start another (empty) line
    lines = text.split("\n")
    return lines[:-1] if lines and lines[-1] == "" else lines


def split(s, delim=" "):
    return [t for t in s.split(delim) if t]               # :497-520: empty tokens dropped


def parse_poses(path, stop):
    R, t = [], []
    with open(path) as f:
        for k, line in enumerate(_getlines(f.read())):
            if k >= stop:
                break
            Rk = [0.0] * 9; tk = [0.0] * 3
            for i, tok in enumerate(split(line)):          # :541-587: switch over the token index
                if i >= 12:
                    continue
                j = _to_double(tok)
                if i % 4 == 3:
                    tk[i // 4] = j
                else:
                    Rk[3 * (i // 4) + i % 4] = j
            R.append(Rk); t.append(tk)
    return R, t


def parse_calibration(path, num_calib, K=None):
    K = [0.0] * 9 if K is None else list(K)
    with open(path) as f:
        for i, calib in enumerate(_getlines(f.read())):
            if i != num_calib:
                continue
            k = 0
            while " " in calib:                            # :612-650: only tokens followed by a space
                pos = calib.index(" ")
                j = _to_double(calib[:pos]); calib = calib[pos + 1:]
                if 1 <= k <= 11 and k % 4 != 0:
                    K[3 * ((k - 1) // 4) + (k - 1) % 4] = j
                k += 1
    return K


def standard_deviation(v):
    avg = sum(v) / len(v)
    return math.sqrt(sum((x - avg) ** 2 for x in v) / (len(v) - 1))


def error_report(R, t, gt_R, gt_t, init_offset):
    gt_R = [list(g) for g in gt_R]; gt_t = [list(g) for g in gt_t]
    eR, et = [], []
    for i in range(1, len(t)):                             # :275-288
        gt_t[i + init_offset][2] *= -1
        gt_R[i + init_offset][6] *= -1
        gt_R[i + init_offset][2] *= -1
        et.append(math.sqrt(sum((a - b) ** 2 for a, b in zip(t[i], gt_t[i + init_offset]))))
        eR.append(math.sqrt(sum((a - b) ** 2 for a, b in zip(R[i], gt_R[i]))))      # gt_R[i]: the reference's own indexing
    return {"R_total": sum(eR), "R_min": min(eR), "R_max": max(eR), "R_std": standard_deviation(eR),
            "t_total": sum(et), "t_min": min(et), "t_max": max(et), "t_std": standard_deviation(et)}
