// TEST INFRASTRUCTURE — the OsqpEigen::Solver calls of IMPCProblem::solve (IMPCProblem.cpp:140-145, 225-296), recording
// the QP data (P, q, A, l, u) exactly as the reference hands them over.  solveProblem() passes them to a solve function the
// checker installs (ref_mpc_set_qp_solver; the tests install the oracle's exact active-set solver — the QP is strictly
// convex in the inputs, so its minimiser does not depend on the solver) and returns its answer as the OSQP solution.
#pragma once
#include <Eigen/Dense>
#include <memory>

namespace OsqpEigen {
enum class ErrorExitFlag { NoError = 0, DataValidationError, SettingsValidationError, LinsysSolverLoadError, LinsysSolverInitError, NonCvxError, MemAllocError, WorkspaceNotInitError };
enum class Status { DualInfeasibleInaccurate = 4, PrimalInfeasibleInaccurate = 3, SolvedInaccurate = 2, Solved = 1, MaxIterReached = -2, PrimalInfeasible = -3, DualInfeasible = -4, Sigint = -5, NonCvx = -7, Unsolved = -10 };

// int solve(n, m, P[n*n] col-major, q[n], A[m*n] col-major, l[m], u[m], z_out[n]) -> OSQP status value
typedef int (*QPSolveFn)(int, int, const double*, const double*, const double*, const double*, const double*, double*);
inline QPSolveFn& qpSolveFunction() { static QPSolveFn f = nullptr; return f; }

class Settings
{
public:
    void setWarmStart(bool v) { warmStart = v; }
    void setVerbosity(bool v) { verbose = v; }
    void setPolish(bool v) { polish = v; }
    bool warmStart = false, verbose = true, polish = false;
};
class Data
{
public:
    void setNumberOfVariables(int n) { nVar = n; }
    void setNumberOfConstraints(int m) { nCon = m; }
    bool setHessianMatrix(const Eigen::View& P_) { P = P_; return P.rows() == nVar && P.cols() == nVar; }
    bool setGradient(const Eigen::View& q_) { q = q_; return q.size() == nVar; }
    bool setLinearConstraintsMatrix(const Eigen::View& A_) { A = A_; return A.rows() == nCon && A.cols() == nVar; }
    bool setLowerBound(const Eigen::View& l_) { l = l_; return l.size() == nCon; }
    bool setUpperBound(const Eigen::View& u_) { u = u_; return u.size() == nCon; }
    int nVar = 0, nCon = 0;
    Eigen::MatrixXd P, A;
    Eigen::VectorXd q, l, u;
};
class Solver
{
public:
    Solver() : m_settings(new Settings), m_data(new Data) {}
    const std::unique_ptr<Settings>& settings() const { return m_settings; }
    const std::unique_ptr<Data>& data() const { return m_data; }
    bool isInitialized() const { return m_init; }
    bool initSolver() { m_init = true; m_z = Eigen::VectorXd::Zero(m_data->nVar); return true; }
    bool updateHessianMatrix(const Eigen::View& P) { return m_data->setHessianMatrix(P); }
    bool updateGradient(const Eigen::View& q) { return m_data->setGradient(q); }
    bool updateLinearConstraintsMatrix(const Eigen::View& A) { return m_data->setLinearConstraintsMatrix(A); }
    bool updateBounds(const Eigen::View& l, const Eigen::View& u) { return m_data->setLowerBound(l) && m_data->setUpperBound(u); }
    ErrorExitFlag solveProblem()
    {
        if (!m_init) return ErrorExitFlag::WorkspaceNotInitError;
        ++nSolves;
        if (!qpSolveFunction()) { m_status = Status::Unsolved; return ErrorExitFlag::NoError; }
        Eigen::VectorXd z = m_z;
        m_status = Status(qpSolveFunction()(m_data->nVar, m_data->nCon, m_data->P.data(), m_data->q.data(), m_data->A.data(),
                                            m_data->l.data(), m_data->u.data(), z.data()));
        m_z = z;
        return ErrorExitFlag::NoError;
    }
    Status getStatus() const { return m_status; }
    const Eigen::VectorXd& getSolution() const { return m_z; }
    int nSolves = 0;
private:
    std::unique_ptr<Settings> m_settings;
    std::unique_ptr<Data> m_data;
    bool m_init = false;
    Status m_status = Status::Unsolved;
    Eigen::VectorXd m_z;
};
} // namespace OsqpEigen
