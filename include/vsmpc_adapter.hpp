// vsmpc_adapter.hpp — header-only C++ host adapter over the C-ABI of include/vsmpc.h.
//
// vsmpc::VariableSamplingMPC keeps the method names, argument meaning and bool error convention of the
// reference class `VariableSamplingMPC : IMPCProblem`
//   (src/flight-controller/momentum-based-linear-mpc-lib/include/variableSamplingMPC/variableSamplingMPC.h:15-41,
//    .../include/IMPCProblem/IMPCProblem.h:35-118)
// for ONE MPC instance (n_instances = 1), so that the reference's C++ controller and its pybind module
// (bindings/python/MPCPyBindings.cpp:22-90) can bind this class instead.  Getters take anything with
// .data()/.size() (std::vector<double>, Eigen::VectorXd, Eigen::Ref<...>) and check the size exactly like
// the reference getters (variableSamplingMPC.cpp:114-227).
//
// vsmpc::BatchedMPC is the same call sequence for B instances on one GPU (structure-of-arrays pack).
//
// Depends on the C++ standard library only.  All compute is in libvsmpc.so (CUDA); there is no CPU path.
#ifndef VSMPC_ADAPTER_HPP
#define VSMPC_ADAPTER_HPP

#include <cstddef>
#include <cstring>
#include <iostream>
#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "vsmpc.h"

namespace vsmpc
{

// ---- parameters: group VS_MPC_CONFIG of src/config/vs_mcp_config.xml:5-44, shipped values as defaults ----
struct Params
{
    int nIter = 17, nIterSmall = 7, controlHorizon = 12;
    double periodMPC = 0.005, periodMPCLargeSteps = 0.1, periodMPCSmallSteps = 0.005;
    bool useJetDynamic = true, useEstimatedThrust = true;
    std::string jointsLambdaOption = "unfiltered";
    double weightCoMPos[3] = {500, 500, 5000};
    double weightCoMPosError[3] = {25000, 25000, 50000};
    double weightLinMom[3] = {1, 1, 1.5};
    double weightRPY[3] = {1000, 1000, 1000};
    double weightRPYError[3] = {10000, 10000, 10000};
    double weightAngMom[3] = {80, 80, 80};
    double weightDeltaJoint[VSMPC_NJ] = {65000, 65000, 65000, 65000, 65000, 65000, 65000, 65000};
    double weightThrottle = 80000, weightInitialThrottle = 80000, weightRegularizationJointPos = 20;
    double throttleMin = 0, throttleMax = 100;
    // UT/src/JetModel.cpp:13-26
    double jetCoeff[13] = {-4.64730485e-01, -8.13171858e+00, -6.19539230e+00, 6.61113140e-01, 1.67673231e+00,
                           -4.83287064e-01, 8.77996617e+00, -1.01096376e+00, -5.86442286e-01, 5.19093322e-01,
                           -4.23782666e-01, -1.45705257e+00, -7.83052261e-03};
    double jetNorm[4] = {108.309, 65.793, 47.333, 31.483};
    // trajectories (TrajectoryManager.cpp:67-140): alphaGravity[alphaLen]; 3 x trajLen arrays, sample-major
    std::vector<double> alphaGravity;
    int alphaFps = 10;
    std::vector<double> positionCoM, velocityCoM, RPY, RPYDot;
    int trajFps = 10;
    int solver = 0;
    // optional joint-limit rows (JointPositionConstraint, constraintsVSMPC.cpp:388-468; degrees like jointPos_max / jointPos_min)
    bool useJointLimits = false;
    double jointPosMin[VSMPC_NJ] = {-180, -180, -180, -180, -180, -180, -180, -180};
    double jointPosMax[VSMPC_NJ] = {180, 180, 180, 180, 180, 180, 180, 180};
};

// ---- one instance's input pack (what update(QPInput&) reads) ------------------------------
class Pack
{
public:
    double v[VSMPC_PACK_DOUBLES];
    Pack() { std::memset(v, 0, sizeof(v)); }
    double& operator[](int i) { return v[i]; }
    const double& operator[](int i) const { return v[i]; }
    // n contiguous doubles
    void set(int off, const double* src, int n) { std::memcpy(v + off, src, sizeof(double) * n); }
    // column-major r x c matrix (Eigen default) -> row-major block at off
    void setRowMajorFromColMajor(int off, const double* src, int r, int c)
    {
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < c; ++j)
                v[off + i * c + j] = src[j * r + i];
    }
    // selected columns of a column-major r x c matrix -> row-major r x sel.size() block
    void setRowMajorCols(int off, const double* src, int r, const std::vector<int>& sel)
    {
        const int n = static_cast<int>(sel.size());
        for (int i = 0; i < r; ++i)
            for (int j = 0; j < n; ++j)
                v[off + i * n + j] = src[sel[j] * r + i];
    }
    void setSelected(int off, const double* src, const std::vector<int>& sel)
    {
        for (size_t j = 0; j < sel.size(); ++j)
            v[off + j] = src[sel[j]];
    }
};

namespace detail
{
inline void fill_config(const Params& p, vsmpc_config& c)
{
    std::memset(&c, 0, sizeof(c));
    c.n_iter = p.nIter;
    c.n_iter_small = p.nIterSmall;
    c.control_horizon = p.controlHorizon;
    c.period_mpc = p.periodMPC;
    c.period_large = p.periodMPCLargeSteps;
    c.period_small = p.periodMPCSmallSteps;
    c.use_jet_dynamic = p.useJetDynamic ? 1 : 0;
    c.use_estimated_thrust = p.useEstimatedThrust ? 1 : 0;
    c.joints_lambda_option = p.jointsLambdaOption == "unfiltered" ? 0 : 1;
    for (int a = 0; a < 3; ++a)
    {
        c.weight_com_pos[a] = p.weightCoMPos[a];
        c.weight_com_pos_error[a] = p.weightCoMPosError[a];
        c.weight_lin_mom[a] = p.weightLinMom[a];
        c.weight_rpy[a] = p.weightRPY[a];
        c.weight_rpy_error[a] = p.weightRPYError[a];
        c.weight_ang_mom[a] = p.weightAngMom[a];
    }
    for (int a = 0; a < VSMPC_NJ; ++a)
        c.weight_delta_joint[a] = p.weightDeltaJoint[a];
    c.weight_throttle = p.weightThrottle;
    c.weight_initial_throttle = p.weightInitialThrottle;
    c.weight_regularization_joint_pos = p.weightRegularizationJointPos;
    c.throttle_min = p.throttleMin;
    c.throttle_max = p.throttleMax;
    std::memcpy(c.jet_coeff, p.jetCoeff, sizeof(c.jet_coeff));
    std::memcpy(c.jet_norm, p.jetNorm, sizeof(c.jet_norm));
    c.alpha_gravity = p.alphaGravity.data();
    c.alpha_len = static_cast<int>(p.alphaGravity.size());
    c.alpha_fps = p.alphaFps;
    c.position_com = p.positionCoM.data();
    c.velocity_com = p.velocityCoM.data();
    c.rpy = p.RPY.data();
    c.rpy_dot = p.RPYDot.data();
    c.traj_len = static_cast<int>(p.positionCoM.size() / 3);
    c.traj_fps = p.trajFps;
    c.solver = p.solver;
    c.use_joint_limits = p.useJointLimits ? 1 : 0;
    for (int a = 0; a < VSMPC_NJ; ++a)
    {
        c.joint_pos_min_deg[a] = p.jointPosMin[a];
        c.joint_pos_max_deg[a] = p.jointPosMax[a];
    }
}
} // namespace detail

// ---- B instances on one GPU -------------------------------------------------------------------------------
class BatchedMPC
{
public:
    BatchedMPC() = default;
    BatchedMPC(const BatchedMPC&) = delete;
    BatchedMPC& operator=(const BatchedMPC&) = delete;
    ~BatchedMPC() { destroy(); }

    bool create(const Params& p, int nInstances, int device = 0)
    {
        destroy();
        if (p.positionCoM.size() != p.velocityCoM.size() || p.positionCoM.size() != p.RPY.size()
            || p.positionCoM.size() != p.RPYDot.size() || p.positionCoM.size() % 3 != 0)
            return error("trajectory arrays differ in length");
        vsmpc_config c;
        detail::fill_config(p, c);
        const int rc = vsmpc_create(&c, nInstances, device, &m_h);
        if (rc != VSMPC_OK)
        {
            const std::string msg = m_h ? vsmpc_last_error(m_h) : "vsmpc_create failed";
            destroy();
            return error(msg);
        }
        m_B = nInstances;
        return true;
    }
    // pack: double[VSMPC_PACK_DOUBLES][B]; jointPosSel: double[8][B]; phase0: int[B] or nullptr
    bool configure(const double* pack, const double* jointPosSel, const int* phase0 = nullptr)
    {
        return check(vsmpc_configure(m_h, pack, jointPosSel, phase0));
    }
    bool update(const double* pack) { return check(vsmpc_set_state(m_h, pack)); }
    bool updateDevice(const double* packDev) { return check(vsmpc_set_state_device(m_h, packDev)); }
    bool solveMPC() { return check(vsmpc_solve(m_h)); }
    bool solveAsync() { return check(vsmpc_solve_async(m_h)); }
    bool wait() { return check(vsmpc_wait(m_h)); }
    // rows: double[B][VSMPC_OUT_DOUBLES], status: int[B]
    bool getOutput(double* rows, int* status) { return check(vsmpc_get_output(m_h, rows, status)); }
    // refs: double[B][VSMPC_REF_DOUBLES] — the QPInput fields update() writes (costsVSMPC.cpp:155-160,
    // systemDynamicsVSMPC.cpp:310)
    bool getReferences(double* refs) { return check(vsmpc_get_references(m_h, refs)); }
    // IMPCProblem::getHessian / getLinearConstraintMatrix / getGradient / getLowerBound / getUpperBound (IMPCProblem.h:88-112)
    bool getHessian(int instance, double* P) { return check(vsmpc_get_hessian(m_h, instance, P)); }
    bool getLinearConstraintMatrix(int instance, double* A) { return check(vsmpc_get_constraint_matrix(m_h, instance, A)); }
    bool getQPVectors(double* q, double* l, double* u) { return check(vsmpc_get_qp_vectors(m_h, q, l, u)); }
    bool setFullSolution(bool on) { return check(vsmpc_set_full_solution(m_h, on ? 1 : 0)); }
    bool getSolution(double* z) { return check(vsmpc_get_full_solution(m_h, z)); }
    // per-instance model parameters / joint limits (BASELINE configs[4]): double[19][B]; double[8][B] each [rad]
    bool setInstanceParams(const double* table) { return check(vsmpc_set_instance_params(m_h, table)); }
    bool setJointLimits(const double* qMin, const double* qMax) { return check(vsmpc_set_joint_limits(m_h, qMin, qMax)); }
    bool setWarmStart(bool on) { return check(vsmpc_set_warm_start(m_h, on ? 1 : 0)); }
    // robot states instead of packs (Robot::setState on the device, UT/src/Robot.cpp:212-332): the kinematic tree once, then
    // double[kinStateDoubles()][B] per tick (VSMPC_KS_* rows)
    bool setKinModel(const vsmpc_kin_model& model) { return check(vsmpc_set_kin_model(m_h, &model)); }
    int kinStateDoubles() const { return vsmpc_kin_state_doubles(m_h); }
    bool configureKinematics(const double* kinState, const int* phase0 = nullptr)
    {
        return check(vsmpc_configure_kinematics(m_h, kinState, phase0));
    }
    bool updateKinematics(const double* kinState) { return check(vsmpc_set_state_kinematics(m_h, kinState)); }
    int getNOptimizationVariables() const { return vsmpc_n_var(m_h); }
    int getNConstraints() const { return vsmpc_n_constraints(m_h); }
    int nInstances() const { return m_B; }
    vsmpc_handle* handle() { return m_h; }
    const std::string& lastError() const { return m_err; }

private:
    void destroy()
    {
        if (m_h)
            vsmpc_destroy(m_h);
        m_h = nullptr;
        m_B = 0;
    }
    bool error(const std::string& msg)
    {
        m_err = msg;
        std::cerr << "[vsmpc] " << msg << std::endl; // the reference logs through yError()
        return false;
    }
    bool check(int rc)
    {
        if (rc == VSMPC_OK)
            return true;
        return error(m_h ? vsmpc_last_error(m_h) : "null handle");
    }
    vsmpc_handle* m_h = nullptr;
    int m_B = 0;
    std::string m_err;
};

// ---- a batch of instances as structure-of-arrays host buffers (what vsmpc_configure / vsmpc_set_state take) -------
// The C++ host layer of north_star (a): every MPC instance hands over its own Pack (array-of-structures: one robot, one
// QPInput), PackBatch scatters it into column i of the SoA pack double[VSMPC_PACK_DOUBLES][B] (row = scalar, column =
// instance) so that the device reads it coalesced; jointPosSel double[8][B] likewise.
// A few persistent host threads for PackBatch::setMany: starting threads per tick costs more than the scatter itself
// (8 x std::thread ~ 0.1 ms against 0.26 ms for 1024 instances on one core).  run(n, f) calls f(k) for k = 0 .. n - 1, the
// caller takes part, returns when all are done.  One run at a time.
class PackThreads
{
public:
    explicit PackThreads(int nThreads) : m_n(nThreads < 1 ? 1 : nThreads)
    {
        for (int t = 1; t < m_n; ++t)
            m_workers.emplace_back([this] { work(); });
    }
    PackThreads(const PackThreads&) = delete;
    PackThreads& operator=(const PackThreads&) = delete;
    ~PackThreads()
    {
        {
            std::lock_guard<std::mutex> lk(m_mu);
            m_stop = true;
            ++m_gen;
        }
        m_cv.notify_all();
        for (std::thread& t : m_workers)
            t.join();
    }
    int size() const { return m_n; }
    void run(int nTasks, const std::function<void(int)>& f)
    {
        unsigned long gen;
        {
            std::lock_guard<std::mutex> lk(m_mu);
            m_f = &f;
            m_tasks = nTasks;
            m_left.store(nTasks);
            gen = ++m_gen;
            m_state.store((unsigned long long)gen << 32);        // (generation, next task): a straggler of an older run claims nothing
        }
        m_cv.notify_all();
        drain(gen, &f, nTasks);
        while (m_left.load(std::memory_order_acquire) > 0)      // the last pieces are microseconds away
            std::this_thread::yield();
    }

private:
    void drain(unsigned long gen, const std::function<void(int)>* f, int tasks)
    {
        for (;;)
        {
            unsigned long long cur = m_state.load();
            if ((unsigned long)(cur >> 32) != (gen & 0xfffffffful) || (int)(cur & 0xffffffffull) >= tasks)
                return;
            if (!m_state.compare_exchange_weak(cur, cur + 1))
                continue;
            (*f)((int)(cur & 0xffffffffull));
            m_left.fetch_sub(1, std::memory_order_release);
        }
    }
    void work()
    {
        unsigned long seen = 0;
        for (;;)
        {
            const std::function<void(int)>* f;
            int tasks;
            {
                std::unique_lock<std::mutex> lk(m_mu);
                m_cv.wait(lk, [&] { return m_gen != seen; });
                seen = m_gen;
                if (m_stop)
                    return;
                f = m_f;
                tasks = m_tasks;
            }
            drain(seen, f, tasks);
        }
    }
    int m_n;
    std::vector<std::thread> m_workers;
    std::mutex m_mu;
    std::condition_variable m_cv;
    unsigned long m_gen = 0;
    bool m_stop = false;
    const std::function<void(int)>* m_f = nullptr;
    int m_tasks = 0;
    std::atomic<unsigned long long> m_state{0};
    std::atomic<int> m_left{0};
};

class PackBatch
{
public:
    // pinned: the SoA buffers are page-locked (vsmpc_host_alloc), which makes vsmpc_set_state's copy asynchronous and lets
    // it overlap the QP kernel of the tick before; falls back to pageable memory when no CUDA device is present
    explicit PackBatch(int nInstances, bool pinned = false) : m_B(nInstances)
    {
        const size_t np = (size_t)VSMPC_PACK_DOUBLES * nInstances, nj = (size_t)VSMPC_NJ * nInstances;
        void* mem = nullptr;
        if (pinned && vsmpc_host_alloc((np + nj) * sizeof(double), &mem) == VSMPC_OK && mem)
        {
            m_pinned = static_cast<double*>(mem);
            std::memset(m_pinned, 0, (np + nj) * sizeof(double));
            m_pack = m_pinned;
            m_jpos = m_pinned + np;
        }
        else
        {
            m_store.assign(np + nj, 0.0);
            m_pack = m_store.data();
            m_jpos = m_store.data() + np;
        }
    }
    PackBatch(const PackBatch&) = delete;
    PackBatch& operator=(const PackBatch&) = delete;
    ~PackBatch()
    {
        if (m_pinned)
            vsmpc_host_free(m_pinned);
    }
    int size() const { return m_B; }
    bool isPinned() const { return m_pinned != nullptr; }
    bool set(int i, const Pack& p)
    {
        if (i < 0 || i >= m_B)
            return false;
        for (int r = 0; r < VSMPC_PACK_DOUBLES; ++r)
            m_pack[(size_t)r * m_B + i] = p.v[r];
        return true;
    }
    // n consecutive instances from their records (array of structures) — the per-tick path of a batch.  One instance at a
    // time (set) touches VSMPC_PACK_DOUBLES cache lines per instance, 8 KB apart at B = 1024; here eight instances are taken
    // together so that every row receives one full 64-byte line, and the instance range is dealt to the host threads of
    // `pool` in pieces of 64 instances that start on a line boundary (no line is shared by two threads).
    bool setMany(int first, const Pack* packs, int n, PackThreads* pool = nullptr)
    {
        if (first < 0 || n < 0 || first + n > m_B || !packs)
            return false;
        if (!pool || pool->size() <= 1 || n < 128)
        {
            scatter(first, packs, n);
            return true;
        }
        // pieces of 64 instances starting on a 64-byte boundary of the rows: a few per thread, dealt dynamically
        const int head = (8 - (first & 7)) & 7, piece = 64;
        if (head > 0)
            scatter(first, packs, head < n ? head : n);
        if (n <= head)
            return true;
        const int nTasks = (n - head + piece - 1) / piece;
        const std::function<void(int)> f = [=](int k) {
            const int t0 = head + k * piece;
            scatter(first + t0, packs + t0, n - t0 < piece ? n - t0 : piece);
        };
        pool->run(nTasks, f);
        return true;
    }
    // one field of one instance (n scalars at row offset off), e.g. only what changed since the last tick
    bool setField(int i, int off, const double* src, int n)
    {
        if (i < 0 || i >= m_B || off < 0 || off + n > VSMPC_PACK_DOUBLES)
            return false;
        for (int r = 0; r < n; ++r)
            m_pack[(size_t)(off + r) * m_B + i] = src[r];
        return true;
    }
    // Robot::getJointPos() of instance i restricted to the controlled joints (configure only)
    bool setJointPos(int i, const std::vector<double>& jointPos, const std::vector<int>& controlledJoints)
    {
        if (i < 0 || i >= m_B || controlledJoints.size() != VSMPC_NJ)
            return false;
        for (int a = 0; a < VSMPC_NJ; ++a)
        {
            if (controlledJoints[a] < 0 || controlledJoints[a] >= static_cast<int>(jointPos.size()))
                return false;
            m_jpos[(size_t)a * m_B + i] = jointPos[controlledJoints[a]];
        }
        return true;
    }
    Pack get(int i) const
    {
        Pack p;
        for (int r = 0; r < VSMPC_PACK_DOUBLES; ++r)
            p.v[r] = m_pack[(size_t)r * m_B + i];
        return p;
    }
    const double* pack() const { return m_pack; }
    double* pack() { return m_pack; }
    const double* jointPosSel() const { return m_jpos; }

private:
    void scatter(int first, const Pack* packs, int n)
    {
        int i = 0;
        for (; i < n && ((first + i) & 7) != 0; ++i)       // up to the first 64-byte boundary of the rows
            set(first + i, packs[i]);
        for (; i + 8 <= n; i += 8)
        {
            const Pack* q = packs + i;
            double* dst = m_pack + first + i;
            for (int r = 0; r < VSMPC_PACK_DOUBLES; ++r, dst += m_B)
            {
                dst[0] = q[0].v[r]; dst[1] = q[1].v[r]; dst[2] = q[2].v[r]; dst[3] = q[3].v[r];
                dst[4] = q[4].v[r]; dst[5] = q[5].v[r]; dst[6] = q[6].v[r]; dst[7] = q[7].v[r];
            }
        }
        for (; i < n; ++i)
            set(first + i, packs[i]);
    }
    int m_B;
    double* m_pack = nullptr;
    double* m_jpos = nullptr;
    double* m_pinned = nullptr;
    std::vector<double> m_store;
};

// ---- B instances sharded over several GPUs of one box, one host thread (vsmpc_multi_*) ---------------------------
class MultiGpuMPC
{
public:
    MultiGpuMPC() = default;
    MultiGpuMPC(const MultiGpuMPC&) = delete;
    MultiGpuMPC& operator=(const MultiGpuMPC&) = delete;
    ~MultiGpuMPC() { destroy(); }
    // devices: one entry per shard (empty: 0 .. nGpus-1); an index may repeat
    bool create(const Params& p, int nInstances, int nGpus, const std::vector<int>& devices = {})
    {
        destroy();
        if (!devices.empty() && static_cast<int>(devices.size()) != nGpus)
            return error("one device index per shard");
        vsmpc_config c;
        detail::fill_config(p, c);
        const int rc = vsmpc_create_multi(&c, nInstances, nGpus, devices.empty() ? nullptr : devices.data(), &m_h);
        if (rc != VSMPC_OK)
        {
            const std::string msg = m_h ? vsmpc_multi_last_error(m_h) : "vsmpc_create_multi failed";
            destroy();
            return error(msg);
        }
        m_B = nInstances;
        return true;
    }
    bool configure(const PackBatch& b, const int* phase0 = nullptr)
    {
        return b.size() == m_B && check(vsmpc_multi_configure(m_h, b.pack(), b.jointPosSel(), phase0));
    }
    bool update(const PackBatch& b) { return b.size() == m_B && check(vsmpc_multi_set_state(m_h, b.pack())); }
    bool solveMPC() { return check(vsmpc_multi_solve(m_h)); }
    bool getOutput(double* rows, int* status) { return check(vsmpc_multi_get_output(m_h, rows, status)); }
    int nInstances() const { return m_B; }
    int nShards() const { return vsmpc_multi_n_shards(m_h); }
    bool shard(int g, int& first, int& count) const { return vsmpc_multi_shard(m_h, g, &first, &count, nullptr) == VSMPC_OK; }
    const std::string& lastError() const { return m_err; }

private:
    void destroy()
    {
        if (m_h)
            vsmpc_multi_destroy(m_h);
        m_h = nullptr;
        m_B = 0;
    }
    bool error(const std::string& msg)
    {
        m_err = msg;
        std::cerr << "[vsmpc] " << msg << std::endl;
        return false;
    }
    bool check(int rc) { return rc == VSMPC_OK ? true : error(m_h ? vsmpc_multi_last_error(m_h) : "null handle"); }
    vsmpc_multi* m_h = nullptr;
    int m_B = 0;
    std::string m_err;
};

// ---- drop-in for the reference class (one instance) ---------------------------------------------------------
class VariableSamplingMPC
{
public:
    // IMPCProblem::configure(parametersHandler, qpInput) (IMPCProblem.cpp:3-148).  `pack` carries what the
    // costs/constraints read from QPInput/Robot at configure time; jointPos = Robot::getJointPos() (all
    // joints); controlledJoints = indices of the 8 controlled joints in that vector
    // (m_jointSelectorVector, variableSamplingMPC.cpp:25-37).
    bool configure(const Params& p, const Pack& pack, const std::vector<double>& jointPos,
                   const std::vector<int>& controlledJoints, int device = 0)
    {
        if (controlledJoints.size() != VSMPC_NJ)
        {
            std::cerr << "[vsmpc] The number of controlled joints defined in the systemDynamic.h file is different "
                         "from the size of the 'controlledJoints' parameter"
                      << std::endl; // variableSamplingMPC.cpp:18-23
            return false;
        }
        for (int j : controlledJoints)
            if (j < 0 || j >= static_cast<int>(jointPos.size()))
                return false;
        m_sel = controlledJoints;
        m_jointsPositionReference = jointPos; // variableSamplingMPC.cpp:60
        if (!m_impl.create(p, 1, device) || !m_impl.setFullSolution(true))
            return false;
        double jp[VSMPC_NJ];
        for (int a = 0; a < VSMPC_NJ; ++a)
            jp[a] = jointPos[m_sel[a]];
        m_nIter = p.nIter;
        m_ctrlHorizon = p.controlHorizon;
        m_nThrottleBlocks = p.controlHorizon - p.nIterSmall + 1;
        std::memset(m_out, 0, sizeof(m_out));
        for (int a = 0; a < VSMPC_NJ; ++a)
            m_out[VSMPC_OUT_JOINTS_REF + a] = jp[a];
        m_status = 0;
        return m_impl.configure(pack.v, jp) && m_impl.getReferences(m_refs);
    }
    // IMPCProblem::update(QPInput&) (IMPCProblem.cpp:150-194); afterwards the published QPInput fields are current
    bool update(const Pack& pack) { return m_impl.update(pack.v) && m_impl.getReferences(m_refs); }
    // the QPInput fields the path writes during configure / update
    double getAlphaGravity() const { return m_refs[VSMPC_REF_ALPHA_GRAVITY]; }
    template <class V> bool getPosCoMReference(V&& out) const { return copyOut(out, m_refs + VSMPC_REF_POS_COM, 3, "getPosCoMReference"); }
    template <class V> bool getRPYReference(V&& out) const { return copyOut(out, m_refs + VSMPC_REF_RPY, 3, "getRPYReference"); }
    template <class V> bool getMomentumReference(V&& out) const { return copyOut(out, m_refs + VSMPC_REF_MOMENTUM, 6, "getMomentumReference"); }
    // VariableSamplingMPC::solveMPC (variableSamplingMPC.cpp:88-112); like the reference it returns true
    // also when the solver status is not "solved" (outputs are then held)
    bool solveMPC()
    {
        if (!m_impl.solveMPC() || !m_impl.getOutput(m_out, &m_status))
            return false;
        for (int a = 0; a < VSMPC_NJ; ++a)
            m_jointsPositionReference[m_sel[a]] = m_out[VSMPC_OUT_JOINTS_REF + a];
        return true;
    }
    template <class V> bool getJointsReferencePosition(V&& out) const
    {
        return copyOut(out, m_jointsPositionReference.data(), m_jointsPositionReference.size(), "getJointsReferencePosition");
    }
    template <class V> bool getThrottleReference(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_THROTTLE, VSMPC_NT, "getThrottleReference"); }
    template <class V> bool getThrustReference(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_THRUST, VSMPC_NT, "getThrustReference"); }
    template <class V> bool getThrustDotReference(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_THRUST_DOT, VSMPC_NT, "getThrustDotReference"); }
    template <class V> bool getFinalCoMPosition(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_FINAL_STATE + 0, 3, "getFinalCoMPosition"); }
    template <class V> bool getFinalLinMom(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_FINAL_STATE + 3, 3, "getFinalLinMom"); }
    template <class V> bool getFinalRPY(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_FINAL_STATE + 6, 3, "getFinalRPY"); }
    template <class V> bool getFinalAngMom(V&& out) const { return copyOut(out, m_out + VSMPC_OUT_FINAL_STATE + 9, 3, "getFinalAngMom"); }
    // the 120 inputs [dq_0..dq_11 | v_0..v_5] (the reference's size check at variableSamplingMPC.cpp:116 is
    // inconsistent with its assignment; here the vector must have nInputs elements)
    template <class V> bool getMPCSolution(V&& out)
    {
        const int nVar = m_impl.getNOptimizationVariables();
        const int nIn = VSMPC_NJ * m_ctrlHorizon + VSMPC_NT * m_nThrottleBlocks;
        if (static_cast<int>(out.size()) != nIn)
            return false;
        std::vector<double> z(nVar);
        if (!m_impl.getSolution(z.data()))
            return false;
        std::memcpy(out.data(), z.data() + (nVar - nIn), sizeof(double) * nIn);
        return true;
    }
    template <class V> bool getSolution(V&& out)
    {
        if (static_cast<int>(out.size()) != m_impl.getNOptimizationVariables())
            return false;
        return m_impl.getSolution(out.data());
    }
    double getNStatesMPC() const { return VSMPC_NX; }
    double getNInputMPC() const { return VSMPC_NJ + VSMPC_NT; }
    int getNOptimizationVariables() const { return m_impl.getNOptimizationVariables(); }
    int getNConstraints() const { return m_impl.getNConstraints(); }
    int getQPProblemStatus() const { return m_status; }
    // dense row-major copies (the reference returns Eigen::Ref to column-major members; same entries)
    template <class V> bool getHessian(V&& out)
    {
        const size_t n = static_cast<size_t>(getNOptimizationVariables());
        return static_cast<size_t>(out.size()) == n * n && m_impl.getHessian(0, out.data());
    }
    template <class V> bool getLinearConstraintMatrix(V&& out)
    {
        const size_t n = static_cast<size_t>(getNOptimizationVariables()), m = static_cast<size_t>(getNConstraints());
        return static_cast<size_t>(out.size()) == n * m && m_impl.getLinearConstraintMatrix(0, out.data());
    }
    template <class V> bool getGradient(V&& out) { return qpVector(out, 0); }
    template <class V> bool getLowerBound(V&& out) { return qpVector(out, 1); }
    template <class V> bool getUpperBound(V&& out) { return qpVector(out, 2); }
    BatchedMPC& impl() { return m_impl; }

private:
    template <class V> static bool copyOut(V&& out, const double* src, size_t n, const char* who)
    {
        if (static_cast<size_t>(out.size()) != n)
        {
            std::cerr << "[vsmpc] VariableSamplingMPC::" << who << ": wrong size of the input vector" << std::endl;
            return false;
        }
        std::memcpy(out.data(), src, sizeof(double) * n);
        return true;
    }
    template <class V> bool qpVector(V&& out, int which)
    {
        const size_t n = static_cast<size_t>(getNOptimizationVariables()), m = static_cast<size_t>(getNConstraints());
        if (static_cast<size_t>(out.size()) != (which == 0 ? n : m))
            return false;
        std::vector<double> q(n), l(m), u(m);
        if (!m_impl.getQPVectors(q.data(), l.data(), u.data()))
            return false;
        const std::vector<double>& src = which == 0 ? q : (which == 1 ? l : u);
        std::memcpy(out.data(), src.data(), sizeof(double) * src.size());
        return true;
    }
    BatchedMPC m_impl;
    std::vector<int> m_sel;
    std::vector<double> m_jointsPositionReference;
    double m_out[VSMPC_OUT_DOUBLES];
    double m_refs[VSMPC_REF_DOUBLES] = {};
    int m_status = 0;
    int m_nIter = 0, m_ctrlHorizon = 0, m_nThrottleBlocks = 0;
};

} // namespace vsmpc
#endif // VSMPC_ADAPTER_HPP
