"""TEST INFRASTRUCTURE ONLY (like everything under oracle/): the reference's own front-end plugins as a backend
of pmv_b200.replay.run_front_end -- OpenCVGoodFeatureExtractor (reference OpenCVGoodFeatureExtractor.cpp:7) and
OpenCVLucasKanadeFM (OpenCVLucasKanadeFM.cpp:5-32).  Used by tests/test_gpu_pipeline_replay.py and by the
cpu_baseline leg of `bench.py --workload pipeline`; never imported by the product package."""
from __future__ import annotations

import numpy as np

WIN = (32, 32)
MAX_LEVEL = 4


class Cv2Backend:
    """The reference's own plugins: OpenCVGoodFeatureExtractor + OpenCVLucasKanadeFM (cv2 = same kernels)."""
    name = "cv2"

    def extract(self, img, roi, max_feats):
        import cv2
        x, y, w, h = roi
        # numpy views lose cv::Mat ROI parentage; the C++ call reads parent pixels at the ROI rim -> crop a
        # response computed on the parent (SURVEY §8c caveat) is what oracle.gftt does; cv2 on a copy is the
        # isolated variant.  Use the oracle form so both backends follow the C++ semantics.
        from . import gftt
        xy, _ = gftt(img, max_feats, 0.01, 5.0, roi=roi)
        return xy.astype(np.int32)

    def track(self, prev, nxt, pts):
        import cv2
        if len(pts) == 0:
            return np.zeros((0, 2), np.float32), np.zeros(0, np.uint8)
        nx, st, _ = cv2.calcOpticalFlowPyrLK(prev, nxt, pts.astype(np.float32).reshape(-1, 1, 2), None,
                                             winSize=WIN, maxLevel=MAX_LEVEL)
        return nx.reshape(-1, 2), st.ravel()
