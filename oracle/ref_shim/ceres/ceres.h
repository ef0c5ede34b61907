// TEST INFRASTRUCTURE ONLY -- the slice of <ceres/ceres.h> the reference touches (CeresBundleAdjustment.cpp:15-61,
// ProjectionResidual.cpp:3-8), so that those files compile UNCHANGED into oracle/_ref/:
//   * Jet<T,N> forward-mode dual numbers + AutoDiffCostFunction: the reference's templated residual functor is
//     evaluated by REAL automatic differentiation of the reference's own source (this is what pins the residual and
//     its Jacobians; Jet arithmetic follows the rules documented in Ceres' jet.h);
//   * Problem / HuberLoss / Solver::Options / Summary / Solve: containers only.  ceres::Solve hands the recorded
//     residual blocks to the plain-C oracle minimiser (oracle/pmv_oracle_ba.c, a restatement of Ceres' published
//     trust-region LM + SPARSE_SCHUR -- Ceres itself is absent, so the MINIMISER stays "parity unpinned") with the
//     residual/Jacobian evaluation hooked back to the reference's CostFunction::Evaluate.
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include <algorithm>
#include <limits>
#include <map>
#include <memory>
#include <unordered_map>
#include "rotation.h"

namespace ceres {
template <typename T, int N> struct Jet {
    T a; T v[N];
    Jet() : a(0) { for (int i = 0; i < N; i++) v[i] = T(0); }
    Jet(const T& s) : a(s) { for (int i = 0; i < N; i++) v[i] = T(0); }          // implicit, like Ceres' Jet(const T&)
    Jet(const T& s, int k) : a(s) { for (int i = 0; i < N; i++) v[i] = T(0); v[k] = T(1); }
};
#define PMV_JET_BIN(op, expr_a, expr_v)                                                                     \
    template <typename T, int N> inline Jet<T, N> operator op(const Jet<T, N>& f, const Jet<T, N>& g)       \
    { Jet<T, N> h; h.a = expr_a; for (int i = 0; i < N; i++) h.v[i] = expr_v; return h; }
PMV_JET_BIN(+, f.a + g.a, f.v[i] + g.v[i])
PMV_JET_BIN(-, f.a - g.a, f.v[i] - g.v[i])
PMV_JET_BIN(*, f.a * g.a, f.a * g.v[i] + f.v[i] * g.a)
#undef PMV_JET_BIN
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, const Jet<T, N>& g)
{   // jet.h: h = f/g, dh = (df - f/g dg)/g, with one reciprocal
    Jet<T, N> h; const T g_a_inverse = T(1.0) / g.a; const T f_a_by_g_a = f.a * g_a_inverse;
    h.a = f_a_by_g_a; for (int i = 0; i < N; i++) h.v[i] = (f.v[i] - f_a_by_g_a * g.v[i]) * g_a_inverse; return h;
}
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f) { Jet<T, N> h; h.a = -f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
// mixed Jet/scalar forms (the functor multiplies Jets by `double camera[k]`)
template <typename T, int N> inline Jet<T, N> operator+(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a + s; return h; }
template <typename T, int N> inline Jet<T, N> operator+(T s, const Jet<T, N>& f) { Jet<T, N> h = f; h.a = s + f.a; return h; }
template <typename T, int N> inline Jet<T, N> operator-(const Jet<T, N>& f, T s) { Jet<T, N> h = f; h.a = f.a - s; return h; }
template <typename T, int N> inline Jet<T, N> operator-(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = s - f.a; for (int i = 0; i < N; i++) h.v[i] = -f.v[i]; return h; }
template <typename T, int N> inline Jet<T, N> operator*(const Jet<T, N>& f, T s) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; i++) h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator*(T s, const Jet<T, N>& f) { Jet<T, N> h; h.a = f.a * s; for (int i = 0; i < N; i++) h.v[i] = f.v[i] * s; return h; }
template <typename T, int N> inline Jet<T, N> operator/(const Jet<T, N>& f, T s) { const T si = T(1.0) / s; Jet<T, N> h; h.a = f.a * si; for (int i = 0; i < N; i++) h.v[i] = f.v[i] * si; return h; }
template <typename T, int N> inline Jet<T, N> operator/(T s, const Jet<T, N>& g) { const T m = -s / (g.a * g.a); Jet<T, N> h; h.a = s / g.a; for (int i = 0; i < N; i++) h.v[i] = g.v[i] * m; return h; }
template <typename T, int N> inline bool operator>(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a > g.a; }
template <typename T, int N> inline bool operator<(const Jet<T, N>& f, const Jet<T, N>& g) { return f.a < g.a; }
template <typename T, int N> inline Jet<T, N> sqrt(const Jet<T, N>& f) { Jet<T, N> h; const T t = std::sqrt(f.a); h.a = t; const T d = T(1.0) / (T(2.0) * t); for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h; }
template <typename T, int N> inline Jet<T, N> cos(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::cos(f.a); const T d = -std::sin(f.a); for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h; }
template <typename T, int N> inline Jet<T, N> sin(const Jet<T, N>& f) { Jet<T, N> h; h.a = std::sin(f.a); const T d = std::cos(f.a); for (int i = 0; i < N; i++) h.v[i] = f.v[i] * d; return h; }

class CostFunction {
public:
    virtual ~CostFunction() {}
    virtual bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const = 0;
    int num_residuals() const { return num_residuals_; }
    const std::vector<int>& parameter_block_sizes() const { return sizes_; }
protected:
    int num_residuals_ = 0; std::vector<int> sizes_;
};

// AutoDiffCostFunction<Functor, kNumResiduals, N0, N1>: two parameter blocks are all the reference uses
template <typename Functor, int kNumResiduals, int N0, int N1> class AutoDiffCostFunction : public CostFunction {
public:
    explicit AutoDiffCostFunction(Functor* f) : functor_(f) { num_residuals_ = kNumResiduals; sizes_ = {N0, N1}; }
    ~AutoDiffCostFunction() { delete functor_; }
    bool Evaluate(double const* const* parameters, double* residuals, double** jacobians) const override
    {
        if (!jacobians) return (*functor_)(parameters[0], parameters[1], residuals);
        typedef Jet<double, N0 + N1> J;
        J x0[N0], x1[N1], out[kNumResiduals];
        for (int i = 0; i < N0; i++) x0[i] = J(parameters[0][i], i);
        for (int i = 0; i < N1; i++) x1[i] = J(parameters[1][i], N0 + i);
        if (!(*functor_)(x0, x1, out)) return false;
        for (int r = 0; r < kNumResiduals; r++) {
            residuals[r] = out[r].a;
            if (jacobians[0]) for (int i = 0; i < N0; i++) jacobians[0][r * N0 + i] = out[r].v[i];            // row major
            if (jacobians[1]) for (int i = 0; i < N1; i++) jacobians[1][r * N1 + i] = out[r].v[N0 + i];
        }
        return true;
    }
private:
    Functor* functor_;
};

class LossFunction { public: virtual ~LossFunction() {} virtual void Evaluate(double sq_norm, double out[3]) const = 0; virtual double huber_delta() const { return -1.0; } };
class HuberLoss : public LossFunction {
public:
    explicit HuberLoss(double a) : a_(a), b_(a * a) {}
    void Evaluate(double s, double rho[3]) const override
    {
        if (s > b_) { const double r = std::sqrt(s); rho[0] = 2.0 * a_ * r - b_; rho[1] = std::max(std::numeric_limits<double>::min(), a_ / r); rho[2] = -rho[1] / (2.0 * s); }
        else { rho[0] = s; rho[1] = 1.0; rho[2] = 0.0; }
    }
    double huber_delta() const override { return a_; }
private:
    const double a_, b_;
};

enum LinearSolverType { DENSE_NORMAL_CHOLESKY, DENSE_QR, SPARSE_NORMAL_CHOLESKY, DENSE_SCHUR, SPARSE_SCHUR, ITERATIVE_SCHUR, CGNR };

class Problem {
public:
    struct ResidualBlock { CostFunction* cost; LossFunction* loss; double* x0; double* x1; };
    ~Problem() { for (auto& b : blocks_) { delete b.cost; delete b.loss; } }               // Problem owns them (Ceres default)
    void AddResidualBlock(CostFunction* c, LossFunction* l, double* x0, double* x1) { blocks_.push_back({c, l, x0, x1}); }
    const std::vector<ResidualBlock>& blocks() const { return blocks_; }
private:
    std::vector<ResidualBlock> blocks_;
};

struct Solver {
    struct Options { LinearSolverType linear_solver_type = SPARSE_NORMAL_CHOLESKY; bool minimizer_progress_to_stdout = false;
                     int num_threads = 1; int max_num_iterations = 50; };
    struct Summary { double initial_cost = 0, final_cost = 0; int iterations = 0, num_successful_steps = 0, termination = 0;
                     double final_radius = 0; std::string FullReport() const; std::string BriefReport() const { return FullReport(); } };
};
void Solve(const Solver::Options& options, Problem* problem, Solver::Summary* summary);   // oracle/ref_shim/shim_impl.cpp
const Solver::Summary& LastSummary();                                                      // harness: summary of the last Solve
}  // namespace ceres
