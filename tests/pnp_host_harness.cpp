// CPU build of the product's pose-solver arithmetic (practical-multi-view_b200/csrc/pnp_math.cuh is __host__ __device__)
// so that tests can pin it to cv2.solvePnP(SOLVEPNP_EPNP) without a GPU.  Test infrastructure only.
#include "../practical-multi-view_b200/csrc/pnp_math.cuh"

extern "C" {

__attribute__((visibility("default"))) double pnp_host_epnp(const double *X, const double *uv, int n, double fu, double fv, double uc,
                                                          double vc, double *R, double *t)
{
    pnp::EPnP e;
    e.n = n; e.fu = fu; e.fv = fv; e.uc = uc; e.vc = vc;
    for (int i = 0; i < n; i++) {
        for (int k = 0; k < 3; k++) e.pws[i][k] = X[3 * i + k];
        e.us[i][0] = uv[2 * i]; e.us[i][1] = uv[2 * i + 1];
    }
    return e.compute_pose(R, t);
}

__attribute__((visibility("default"))) void pnp_host_epnp_debug(const double *X, const double *uv, int n, double fu, double fv, double uc, double vc, int which, int gn, double *R, double *t, double* err)
{
    pnp::EPnP e;
    e.n = n; e.fu = fu; e.fv = fv; e.uc = uc; e.vc = vc;
    for (int i = 0; i < n; i++) { for (int k = 0; k < 3; k++) e.pws[i][k] = X[3 * i + k]; e.us[i][0] = uv[2 * i]; e.us[i][1] = uv[2 * i + 1]; }
    e.dbg_which = which; e.dbg_gn = gn;
    *err = e.compute_pose(R, t);
}

__attribute__((visibility("default"))) void pnp_host_subsets(int n, int model_points, int count, int *out)
{
    pnp::CvRng rng;
    for (int s = 0; s < count; s++) {
        int *idx = out + s * model_points;
        for (int i = 0; i < model_points;) {
            int v = rng.uniform(0, n);
            bool dup = false;
            for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
            if (dup) continue;
            idx[i++] = v;
        }
    }
}

// sequential restatement of RANSACPointSetRegistrator::run with PnPRansacCallback (5-point EPnP, float reprojection error)
__attribute__((visibility("default"))) int pnp_host_ransac(const float *X, const float *uv, int n, double fu, double fv, double uc, double vc,
                                                          int max_iters, double thr, double conf, double *R, double *t, unsigned char *mask)
{
    pnp::CvRng rng;
    int niters = max_iters, maxgood = 0, found = 0;
    unsigned char *tmp = new unsigned char[n];
    for (int iter = 0; iter < niters; iter++) {
        int idx[5];
        for (int i = 0; i < 5;) {
            int v = rng.uniform(0, n);
            bool dup = false;
            for (int j = 0; j < i; j++) dup = dup || idx[j] == v;
            if (dup) continue;
            idx[i++] = v;
        }
        pnp::EPnP e;
        e.n = 5; e.fu = fu; e.fv = fv; e.uc = uc; e.vc = vc;
        for (int i = 0; i < 5; i++) {
            for (int k = 0; k < 3; k++) e.pws[i][k] = X[3 * idx[i] + k];
            e.us[i][0] = uv[2 * idx[i]]; e.us[i][1] = uv[2 * idx[i] + 1];
        }
        double Rh[9], th[3];
        e.compute_pose(Rh, th);
        int good = pnp::count_inliers(X, uv, n, Rh, th, fu, fv, uc, vc, (float)(thr * thr), tmp);
        if (good > (maxgood > 4 ? maxgood : 4)) {
            for (int i = 0; i < n; i++) mask[i] = tmp[i];
            for (int i = 0; i < 9; i++) R[i] = Rh[i];
            for (int i = 0; i < 3; i++) t[i] = th[i];
            maxgood = good; found = 1;
            niters = pnp::ransac_update_num_iters(conf, (double)(n - good) / n, 5, niters);
        }
    }
    delete[] tmp;
    return found ? maxgood : 0;
}

// serial driver of the refinement (the device kernel runs the same steps with the accumulation spread over a CTA)
__attribute__((visibility("default"))) int pnp_host_refine(const float *X, const float *uv, const unsigned char *mask, int n, double fu, double fv,
                                                          double uc, double vc, double *rvec, double *tvec)
{
    double R[9], t[3] = {tvec[0], tvec[1], tvec[2]};
    pnp::rodrigues_to_matrix(rvec, R);
    double lambda = 1e-3, cost = 0;
    int it = 0;
    for (; it < 100; it++) {
        double H[21] = {0}, g[6] = {0};
        cost = 0;
        for (int i = 0; i < n; i++) {
            if (mask && !mask[i]) continue;
            double r[2], J[12];
            pnp::reproj_jac(X + 3 * i, uv + 2 * i, R, t, fu, fv, uc, vc, r, J);
            pnp::lm_accumulate(r, J, H, g, &cost);
        }
        bool moved = false;
        double dn = 0;
        for (int tries = 0; tries < 30 && !moved; tries++) {
            double d[6], Rn[9], tn[3];
            if (!pnp::lm_solve(H, g, lambda, d)) { lambda *= 10; continue; }
            pnp::lm_apply(R, t, d, Rn, tn);
            double c2 = 0;
            for (int i = 0; i < n; i++) {
                if (mask && !mask[i]) continue;
                double r[2], J[12];
                pnp::reproj_jac(X + 3 * i, uv + 2 * i, Rn, tn, fu, fv, uc, vc, r, J);
                c2 += r[0] * r[0] + r[1] * r[1];
            }
            if (c2 <= cost) {
                for (int k = 0; k < 9; k++) R[k] = Rn[k];
                for (int k = 0; k < 3; k++) t[k] = tn[k];
                dn = 0; for (int k = 0; k < 6; k++) dn += d[k] * d[k];
                lambda = lambda > 1e-12 ? lambda * 0.1 : lambda;
                moved = true;
            } else lambda *= 10;
        }
        if (!moved || dn < 1e-24) break;
    }
    pnp::matrix_to_rodrigues(R, rvec);
    for (int k = 0; k < 3; k++) tvec[k] = t[k];
    return it;
}

__attribute__((visibility("default"))) int pnp_host_update_iters(double p, double ep, int mp, int mx) { return pnp::ransac_update_num_iters(p, ep, mp, mx); }

__attribute__((visibility("default"))) void pnp_host_rodrigues(const double *r, double *R, double *r_back)
{
    pnp::rodrigues_to_matrix(r, R);
    pnp::matrix_to_rodrigues(R, r_back);
}

}
