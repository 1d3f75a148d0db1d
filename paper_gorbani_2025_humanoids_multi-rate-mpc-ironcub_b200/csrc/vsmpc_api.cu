// C-ABI host layer (include/vsmpc.h): owns the device buffers of one batch of MPC instances on one
// GPU and launches the kernels.  Mirrors the call sequence of the reference's
// VariableSamplingMPC (configure -> update -> solveMPC -> getters).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "vsmpc_common.cuh"
#include "vsmpc_plant.cuh"

namespace vsmpc
{
cudaError_t launch_plant(const DeviceConfig* d_cfg, const PlantModel* d_pm, int B, int mode, double* ps,
                         const double* pp, const double* out_rows, const int* status, double* pack, double* rec,
                         const double* ip, const double* st, const double* thr_sub, cudaStream_t s);
cudaError_t launch_jet_nn_eval(const JetNN* d_nn, int n_groups, float dt, const float* T_in, const float* u_in, float* T_out,
                               float* Td_out, cudaStream_t s);
cudaError_t launch_jet_nn_ekf(const DeviceConfig* d_cfg, const PlantModel* d_pm, const JetNN* d_nn, int B, double* ps,
                              const double* out_rows, const double* ip, double* thr_sub, cudaStream_t s);
cudaError_t launch_linearise(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, int mode,
                             const double* pack, const double* joint_pos_sel, const int* phase0, double* st,
                             int* si, const double* alpha_traj, const double* traj_pos, const double* traj_vel,
                             const double* traj_rpy, const double* traj_rpyd, double* qd, const double* ip,
                             int* fb_count, const double* jl, cudaStream_t s);
bool fallback_supported(const DeviceConfig& cfg);
size_t fallback_slot_doubles(const DeviceConfig& cfg);
size_t fallback_pos_ints(const DeviceConfig& cfg);
void fallback_positions(const DeviceConfig& cfg, int* out);
cudaError_t launch_qp_fallback(const DeviceConfig& h_cfg, int B, int n_slots, const double* qd, const int* fb_list,
                               const int* fb_count, const int* pos, double* scratch, double* z, double* st, double* out_rows,
                               int* status, int* n_factor, int* n_solve, int* n_pivot, int want_z, double* out2, int* status2,
                               cudaStream_t s);
cudaError_t launch_expand_dynamics(const DeviceConfig* d_cfg, int B, const double* qd, double* A, double* BJ,
                                   double* BT, double* c, cudaStream_t s);
cudaError_t launch_expand_qp_vectors(const DeviceConfig* d_cfg, int B, const double* qd, double* q, double* l,
                                     double* u, cudaStream_t s);
cudaError_t launch_expand_hessian(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, double* P, cudaStream_t s);
cudaError_t launch_expand_constraint_matrix(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, const double* qd_inst,
                                            double* scratch, double* M, cudaStream_t s);
size_t generic_scratch_doubles(const DeviceConfig& cfg);
bool generic_supported(const DeviceConfig& cfg);
cudaError_t launch_qp_generic(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                              double* ws, double* scratch, double* z, double* st, double* out_rows, int* status,
                              int* n_factor, int* n_solve, cudaStream_t s);
size_t structured_scratch_doubles(const DeviceConfig& cfg);
cudaError_t launch_qp_structured(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                                 double* ws, double* scratch, double* z, double* st, double* out_rows, int* status,
                                 int* n_factor, int* n_solve, int want_z, cudaStream_t s);
bool condensed_supported(const DeviceConfig& cfg);
int condensed_phase_clocks(long long* host, int n);
int condensed_wide_phase_clocks(long long* host, int n);
size_t condensed_ws_doubles(const DeviceConfig& cfg);
cudaError_t launch_qp_condensed(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                                double* ws, double* z, double* st, double* out_rows, int* status, int* n_factor,
                                int* n_solve, int* n_pivot, int want_z, int* fb_list, int* fb_count, int fb_mode,
                                double* out2, int* status2, unsigned* jlset, cudaStream_t s);
size_t condensed_jlset_words();
bool condensed_wide_supported(const DeviceConfig& cfg);
size_t condensed_wide_ws_doubles(const DeviceConfig& cfg);
size_t condensed_wide_scratch_doubles(const DeviceConfig& cfg);
cudaError_t launch_qp_condensed_wide(const DeviceConfig& h_cfg, int B, const double* qd, double* ws, double* scratch,
                                     double* z, double* st, double* out_rows, int* status, int* n_factor, int* n_solve,
                                     int* n_pivot, int want_z, int* fb_list, int* fb_count, int fb_mode, signed char* wset,
                                     int warm, double* out2, int* status2, unsigned* jlset, cudaStream_t s);
size_t condensed_wide_wset_bytes(const DeviceConfig& cfg);
struct KinModelDev;
const char* kin_prepare(const vsmpc_kin_model& m, KinModelDev& K);
size_t kin_model_bytes();
int kin_state_rows(const KinModelDev& K);
cudaError_t launch_kinematics(const KinModelDev* d_model, int ks_rows, int B, const double* ks, double* pack, double* jpos,
                              cudaStream_t s);
} // namespace vsmpc

using namespace vsmpc;

constexpr int SOLVER_WIDE = 3;   // internal: chosen by the default solver for long horizons
static std::atomic<int> g_last_qp_solver{0}; // development (vsmpc_debug_phase_clocks, handle-less): which kernel stamped its clocks last

struct vsmpc_handle
{
    DeviceConfig cfg{};
    DeviceConfig* d_cfg = nullptr;
    int B = 0;
    int device = 0;
    int solver = 0;           // 0 condensed, 1 generic, 2 structured, SOLVER_WIDE condensed with several column warps
    bool configured = false;
    bool has_state = false;
    bool want_full = false;   // write the full primal (IMPCProblem::getSolution) every solve
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    // host -> device staging of vsmpc_set_state: its own copy stream and two pack buffers, so that the copy of tick
    // j+1 overlaps the kernels of tick j
    cudaStream_t copy_stream = nullptr;
    cudaStream_t out_stream = nullptr;   // device -> host read-back of vsmpc_get_output_async
    cudaEvent_t ev_solved = nullptr;     // last QP kernel done
    double* d_pack_in[2] = {nullptr, nullptr};
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    cudaEvent_t ev_k1[2] = {nullptr, nullptr};
    cudaEvent_t ev_out[2] = {nullptr, nullptr};
    int pack_idx = 0, out_idx = 0;
    // Overlapped linearisation (host-pack path): the linearise kernel of tick j+1 runs on its own stream WHILE the QP kernel of
    // tick j runs on the compute stream — it only needs tick j+1's pack and the tick state the previous linearise kernel left
    // (the QP kernels write one block of that state, the joint accumulator, which the linearise kernel does not use).  What the
    // two kernels would share is double-buffered: the QP data block K1 -> K2 and the fallback list / count.
    cudaStream_t k1_stream = nullptr;
    cudaEvent_t ev_lin[2] = {nullptr, nullptr};      // linearise kernel that filled QP buffer q is done (on k1_stream)
    cudaEvent_t ev_qp[2] = {nullptr, nullptr};       // QP + fallback kernels that read QP buffer q are done (on the compute stream)
    cudaEvent_t ev_lin_main = nullptr;               // last linearise kernel launched on the compute stream
    double* d_qd2[2] = {nullptr, nullptr};
    int* d_fb_list2[2] = {nullptr, nullptr};
    int* d_fb_count2[2] = {nullptr, nullptr};
    int qd_idx = 0;                                  // buffer the last linearise kernel filled = the one the next solve reads
    bool lin_on_k1 = false;                          // that kernel ran on k1_stream and the compute stream has not waited for it yet
    bool capturing = false;                          // a rollout tick is being captured: no event records for other streams
    // asynchronous read-back: the rows are snapshotted on the compute stream into one of two staging buffers and copied to
    // the host from there, so that the next QP kernel (which rewrites d_out in place: held outputs) does not wait for PCIe
    double* d_out_stage[2] = {nullptr, nullptr};
    int* d_status_stage[2] = {nullptr, nullptr};
    int stage_written = -1;   // staging buffer the last QP kernel filled by itself (condensed kernels), -1: none
    // device buffers
    double* d_pack = nullptr;
    double* d_jpos = nullptr;
    int* d_phase = nullptr;
    double* d_st = nullptr;
    int* d_si = nullptr;
    double* d_alpha = nullptr;
    double *d_tpos = nullptr, *d_tvel = nullptr, *d_trpy = nullptr, *d_trpyd = nullptr;
    double* d_qd = nullptr;
    double* d_ws = nullptr;
    double* d_scratch = nullptr;
    double* d_z = nullptr;
    double* d_out = nullptr;
    int* d_status = nullptr;
    int *d_nf = nullptr, *d_ns = nullptr, *d_np = nullptr;   // factorisations, forward passes, exchange pivots of the last solve
    cudaEvent_t ev_stream = nullptr;     // orders a new stream behind the old one in vsmpc_set_stream
    // fallback QP kernel (vsmpc_qp_fallback.cu): instances the Riccati recursion breaks down on
    int* d_fb_list = nullptr;            // int[B]
    int* d_fb_count = nullptr;           // zeroed by the linearise kernel, filled by the QP kernel
    int* d_fb_pos = nullptr;             // bordered-band positions of the KKT unknowns
    double* d_fb_scratch = nullptr;      // n_slots x fallback_slot_doubles
    int fb_slots = 0;
    int fb_mode = 0;                     // 0 off, 1 on (default where supported), 2 every instance (tests)
    // device-resident closed loop
    PlantModel* d_pm = nullptr;
    double* d_ps = nullptr;
    double* d_pp = nullptr;
    JetNN* d_nn = nullptr;    // neural jet plant + EKF constants (jet-NN mode of the rollout)
    double* d_thr_sub = nullptr;
    bool use_nn = false;
    double* d_ip = nullptr;   // per-instance jet model / throttle limits (optional)
    bool use_ip = false;
    // batched reduced kinematics (vsmpc_kinematics.cu): the pack is built on the device from the robot state
    KinModelDev* d_kin = nullptr;
    double* d_ks[2] = {nullptr, nullptr};   // robot-state staging, double[ks_rows][B] each (pipelined like the pack)
    int ks_rows = 0;
    signed char* d_wset = nullptr;   // long-horizon kernel: working set of the last solve per instance (warm start)
    size_t wset_bytes = 0;
    unsigned* d_jlset = nullptr;     // reference-horizon kernel with joint-limit rows: working set of the joint boxes per instance
    size_t jlset_bytes = 0;
    bool warm = true;
    double* d_jl = nullptr;   // per-instance joint limits [rad], SoA double[16][B]: 8 lower rows, 8 upper rows (optional)
    bool use_jl_table = false;
    bool rollout_ready = false;
    cudaGraphExec_t tick_graph = nullptr;
    std::string err;
};

namespace
{
int fail(vsmpc_handle* h, int code, const std::string& msg)
{
    if (h)
        h->err = msg;
    return code;
}

int cuda_fail(vsmpc_handle* h, cudaError_t e, const char* what)
{
    return fail(h, VSMPC_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CK(call)                                      \
    do                                                \
    {                                                 \
        cudaError_t e__ = (call);                     \
        if (e__ != cudaSuccess)                       \
            return cuda_fail(h, e__, #call);          \
    } while (0)

// Trajectory::upsample, UT/src/TrajectoryManager.cpp:23-39 (drops the last sample)
std::vector<double> upsample(const double* v, int dim, int len, int fps, int des_fps)
{
    std::vector<double> out;
    const double ratio = static_cast<double>(des_fps) / fps;
    for (int i = 0; i + 1 < len; ++i)
        for (size_t k = 0; k < ratio; ++k)
            for (int a = 0; a < dim; ++a)
                out.push_back(v[(size_t)i * dim + a] + (v[(size_t)(i + 1) * dim + a] - v[(size_t)i * dim + a]) * (k / ratio));
    return out;
}

template <typename T> cudaError_t dalloc(T** p, size_t n)
{
    return cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T));
}
} // namespace

// VariableSamplingMPC::configure (variableSamplingMPC.cpp:56-60): the held outputs start from zero and the joint
// reference from Robot::getJointPos(), so that a tick that fails before any success holds the configure-time posture
// (ps != nullptr: rollout — the commands in effect in the plant state are what a failed first tick holds)
__global__ void seed_outputs_kernel(int B, const double* __restrict__ jpos, const double* __restrict__ ps,
                                    double* __restrict__ out_rows, int* __restrict__ status)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B)
        return;
    double* o = out_rows + (size_t)i * VSMPC_OUT_DOUBLES;
    for (int e = 0; e < VSMPC_OUT_DOUBLES; ++e)
        o[e] = 0.0;
    for (int a = 0; a < NJ; ++a)
        o[VSMPC_OUT_JOINTS_REF + a] = jpos[(size_t)a * B + i];
    if (ps)
        for (int j = 0; j < NT; ++j)
        {
            o[VSMPC_OUT_THROTTLE + j] = ps[(size_t)(PS_THROTTLE + j) * B + i];
            o[VSMPC_OUT_THRUST + j] = ps[(size_t)(PS_TDES + j) * B + i];
            o[VSMPC_OUT_THRUST_DOT + j] = ps[(size_t)(PS_TDDES + j) * B + i];
        }
    status[i] = VSMPC_STATUS_SOLVED;
}

// snapshot of the output rows + status for the asynchronous read-back: one launch instead of two device-to-device copies
__global__ void snapshot_outputs_kernel(size_t n_rows_doubles, int B, const double* __restrict__ rows, const int* __restrict__ status,
                                        double* __restrict__ rows_out, int* __restrict__ status_out)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_rows_doubles / 2; e += stride)
        reinterpret_cast<double2*>(rows_out)[e] = reinterpret_cast<const double2*>(rows)[e];
    for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < (size_t)B; e += stride)
        status_out[e] = status[e];
}

static void drop_tick_graph(vsmpc_handle* h)
{ // the captured tick bakes the kernel arguments in (per-instance table, full-solution flag, jet-NN mode)
    if (h->tick_graph)
    {
        cudaGraphExecDestroy(h->tick_graph);
        h->tick_graph = nullptr;
    }
}

// host SoA [rows][stride] (column window of a larger batch) -> device SoA [rows][B]
static cudaError_t copy_soa_h2d(double* dst, const double* src, int rows, size_t B, size_t stride, cudaStream_t s)
{
    if (stride == B)
        return cudaMemcpyAsync(dst, src, (size_t)rows * B * 8, cudaMemcpyHostToDevice, s);
    return cudaMemcpy2DAsync(dst, B * 8, src, stride * 8, B * 8, rows, cudaMemcpyHostToDevice, s);
}

extern "C" {

static int solve_launch(vsmpc_handle* h);

const char* vsmpc_last_error(const vsmpc_handle* h)
{
    return h ? h->err.c_str() : "null handle";
}

int vsmpc_create(const vsmpc_config* c, int n_instances, int device, vsmpc_handle** out)
{
    if (!c || !out || n_instances <= 0)
        return VSMPC_ERR_ARG;
    *out = nullptr;
    vsmpc_handle* h = new (std::nothrow) vsmpc_handle();
    if (!h)
        return VSMPC_ERR_ARG;
    auto bail = [&](int code, const std::string& msg) {
        // keep the handle alive so that the message can be read, but mark it unusable
        h->err = msg;
        h->B = 0;
        *out = h;
        return code;
    };
    if (c->n_iter < 2 || c->n_iter > MAX_ITER || c->n_iter_small < 2 || c->control_horizon < c->n_iter_small
        || c->control_horizon > c->n_iter)
        return bail(VSMPC_ERR_ARG, "need 2 <= nIterSmall <= controlHorizon <= nIter <= VSMPC_MAX_ITER");
    if (c->joints_lambda_option != 0)
        return bail(VSMPC_ERR_UNSUPPORTED, "jointsLambdaOption 'constant' is not implemented (only 'unfiltered')");
    if (!c->alpha_gravity || c->alpha_len < 1 || !c->position_com || !c->velocity_com || !c->rpy || !c->rpy_dot
        || c->traj_len < 1)
        return bail(VSMPC_ERR_ARG, "trajectory arrays missing");
    if (!(c->period_mpc > 0) || !(c->period_large > 0) || !(c->period_small > 0))
        return bail(VSMPC_ERR_ARG, "periods must be positive");
    // the resampling loops divide by the file rates and the kernels by the 20-tick ratio: reject what the reference
    // would loop or divide by zero on (TrajectoryManager.cpp:23-39, costsVSMPC.cpp:68-69)
    if (c->alpha_fps <= 0 || c->traj_fps <= 0)
        return bail(VSMPC_ERR_ARG, "trajectory rates (alpha_fps, traj_fps) must be positive");
    if ((int)(1 / c->period_mpc) <= 0 || (int)(1 / c->period_large) <= 0)
        return bail(VSMPC_ERR_ARG, "1/periodMPC and 1/periodMPCLargeSteps must be at least 1 Hz");
    if (std::lround(c->period_large / c->period_small) < 1 || c->period_large / c->period_small > 1e6)
        return bail(VSMPC_ERR_ARG, "periodMPCLargeSteps / periodMPCSmallSteps must round to a ratio in [1, 1e6]");
    if ((double)c->alpha_len * (1 / c->period_mpc) / c->alpha_fps > 5e7
        || (double)c->traj_len * (1 / c->period_large) / c->traj_fps > 5e7)
        return bail(VSMPC_ERR_ARG, "resampled trajectory too long");

    DeviceConfig& g = h->cfg;
    g.N = c->n_iter;
    g.Ns = c->n_iter_small;
    g.Nc = c->control_horizon;
    g.NC = g.N - g.Ns + 1;
    g.nblk = g.Nc - g.Ns + 1;
    g.n_var = NX * (g.N + 1) + NJ * g.Nc + NT * g.nblk;               // variableSamplingMPC.cpp:44-45
    g.n_con = NX * g.N + NX + NT * (g.N - g.Ns + 1);                  // constraintsVSMPC.cpp:7,283; IQPUtilsMPC.cpp:59
    g.ratio = (int)std::lround(c->period_large / c->period_small);    // costsVSMPC.cpp:69
    g.use_jet_dynamic = c->use_jet_dynamic;
    g.use_estimated_thrust = c->use_estimated_thrust;
    g.use_jl = c->use_joint_limits ? 1 : 0;
    if (g.use_jl)
    { // JointPositionConstraint (optional rows): nJoints * nIter rows after the throttle rows, constraintsVSMPC.cpp:391
        g.n_con += NJ * g.N;
        for (int a = 0; a < NJ; ++a)
        {
            g.jl_min[a] = c->joint_pos_min_deg[a] * M_PI / 180.0;    // :421-423
            g.jl_max[a] = c->joint_pos_max_deg[a] * M_PI / 180.0;
            if (!(g.jl_min[a] < g.jl_max[a]))
                return bail(VSMPC_ERR_ARG, "joint limits: need jointPos_min < jointPos_max for every controlled joint");
        }
    }
    g.qd_stride = (QD_XREF + 12 * g.NC + 3) & ~3;
    g.st_rows = ST_WIN + 12 * g.NC;
    for (int i = 0; i < NX; ++i)
        g.Qd[i] = 0.0;
    for (int a = 0; a < 3; ++a)
    { // costsVSMPC.cpp:78-93
        g.Qd[IX_COM + a] = c->weight_com_pos[a];
        g.Qd[IX_LIN + a] = c->weight_lin_mom[a];
        g.Qd[IX_RPY + a] = c->weight_rpy[a];
        g.Qd[IX_ANG + a] = c->weight_ang_mom[a];
        g.Qd[IX_EP + a] = c->weight_com_pos_error[a];
        g.Qd[IX_ER + a] = c->weight_rpy_error[a];
    }
    for (int a = 0; a < NJ; ++a) // costsVSMPC.cpp:375-382 + :564-571
        g.Rqd[a] = c->weight_delta_joint[a] + c->weight_regularization_joint_pos;
    g.w_reg_q = c->weight_regularization_joint_pos;
    g.w_t = c->weight_throttle;
    g.w_i = c->weight_initial_throttle;
    g.throttle_min = c->throttle_min;
    g.throttle_max = c->throttle_max;
    std::memcpy(g.jc, c->jet_coeff, sizeof(g.jc));
    std::memcpy(g.jn, c->jet_norm, sizeof(g.jn));
    { // warp function, constraintsVSMPC.cpp:45-51,76-84,156-159
        const double beta2 = (c->period_large - g.Ns * c->period_small) / (g.Ns * (g.Ns - 1));
        const double beta1 = c->period_small - beta2;
        auto warp = [&](double t) { return beta1 * t + beta2 * t * t; };
        for (int k = 0; k < g.N; ++k)
            g.dt[k] = k < g.Ns ? warp(k + 1) - warp(k) : c->period_large;
    }
    // trajectories with the TrajectoryManager's resampling (TrajectoryManager.cpp:121-126)
    std::vector<double> alpha(c->alpha_gravity, c->alpha_gravity + c->alpha_len);
    const int des_alpha = (int)(1 / c->period_mpc); // double -> int at systemDynamicsVSMPC.cpp:272
    if (c->alpha_fps != des_alpha && c->alpha_len > 1)
        alpha = upsample(c->alpha_gravity, 1, c->alpha_len, c->alpha_fps, des_alpha);
    const int des_traj = (int)(1 / c->period_large); // costsVSMPC.cpp:68
    std::vector<double> tp(c->position_com, c->position_com + 3 * (size_t)c->traj_len);
    std::vector<double> tv(c->velocity_com, c->velocity_com + 3 * (size_t)c->traj_len);
    std::vector<double> tr(c->rpy, c->rpy + 3 * (size_t)c->traj_len);
    std::vector<double> td(c->rpy_dot, c->rpy_dot + 3 * (size_t)c->traj_len);
    if (c->traj_fps != des_traj && c->traj_len > 1)
    {
        tp = upsample(c->position_com, 3, c->traj_len, c->traj_fps, des_traj);
        tv = upsample(c->velocity_com, 3, c->traj_len, c->traj_fps, des_traj);
        tr = upsample(c->rpy, 3, c->traj_len, c->traj_fps, des_traj);
        td = upsample(c->rpy_dot, 3, c->traj_len, c->traj_fps, des_traj);
    }
    g.alpha_len = (int)alpha.size();
    g.traj_len = (int)(tp.size() / 3);
    if (g.alpha_len < 1 || g.traj_len < 1)
        return bail(VSMPC_ERR_ARG, "empty trajectory after resampling");

    h->B = n_instances;
    h->device = device;
    // solver routing: 0 = default (condensed-throttle Riccati kernel; horizons it does not cover fall back to the
    // generic dense kernel), 1 = generic dense variant, 2 = structured one-warp kernel
    if (c->solver < 0 || c->solver > 2)
        return bail(VSMPC_ERR_ARG, "solver must be 0 (default), 1 (generic) or 2 (structured)");
    h->solver = c->solver;
    if (h->solver == 0 && !condensed_supported(g))   // horizons beyond 6 throttle blocks / 32 knots: the condensed kernel
        h->solver = condensed_wide_supported(g) ? SOLVER_WIDE : 1;   // with several column warps, else the generic one
    if (h->solver == 1 && !generic_supported(g))
        return bail(VSMPC_ERR_UNSUPPORTED, "horizons with more than 48 throttle blocks are not supported");
    if (h->solver == 2 && (NT * g.nblk > 24 || g.N > 32))
        return bail(VSMPC_ERR_UNSUPPORTED, "the structured one-warp kernel covers horizons with <= 6 throttle blocks and <= 32 knots");
    // joint-limit rows: the reference-horizon kernel and the long-horizon kernel with up to two column warps carry the joint
    // boxes in their own working set; the fallback kernel is the net behind a working set that does not settle and the
    // carrier where the QP kernel has none — so the rows go as far as the fallback kernel's size limit
    if (g.use_jl && !((h->solver == 0 || h->solver == SOLVER_WIDE) && fallback_supported(g)))
        return bail(VSMPC_ERR_UNSUPPORTED, "joint-limit rows need the default solver on a horizon the fallback kernel covers");
    const int B = n_instances;
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess)
        return bail(VSMPC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
    const size_t scratch = h->solver == 1 ? generic_scratch_doubles(g)
                           : (h->solver == 2 ? structured_scratch_doubles(g)
                                             : (h->solver == SOLVER_WIDE ? condensed_wide_scratch_doubles(g) : 4));
    bool ok = true;
    auto A = [&](cudaError_t r) { ok = ok && (r == cudaSuccess); if (r != cudaSuccess) e = r; };
    A(cudaStreamCreateWithFlags(&h->own_stream, cudaStreamNonBlocking));
    h->stream = h->own_stream;
    A(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    A(cudaStreamCreateWithFlags(&h->out_stream, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&h->ev_solved, cudaEventDisableTiming));
    A(cudaEventCreateWithFlags(&h->ev_stream, cudaEventDisableTiming));
    for (int q = 0; q < 2; ++q)
    {
        A(dalloc(&h->d_pack_in[q], (size_t)VSMPC_PACK_DOUBLES * B));
        A(cudaEventCreateWithFlags(&h->ev_h2d[q], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&h->ev_k1[q], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&h->ev_out[q], cudaEventDisableTiming));
    }
    A(dalloc(&h->d_cfg, 1));
    A(dalloc(&h->d_pack, (size_t)VSMPC_PACK_DOUBLES * B));
    A(dalloc(&h->d_jpos, (size_t)NJ * B));
    A(dalloc(&h->d_phase, (size_t)B));
    A(dalloc(&h->d_st, (size_t)g.st_rows * B));
    A(dalloc(&h->d_si, (size_t)SI_COUNT * B));
    A(dalloc(&h->d_alpha, alpha.size()));
    A(dalloc(&h->d_tpos, tp.size()));
    A(dalloc(&h->d_tvel, tv.size()));
    A(dalloc(&h->d_trpy, tr.size()));
    A(dalloc(&h->d_trpyd, td.size()));
    A(dalloc(&h->d_qd2[0], (size_t)g.qd_stride * B));
    A(dalloc(&h->d_qd2[1], (size_t)g.qd_stride * B));
    h->d_qd = h->d_qd2[0];
    A(cudaStreamCreateWithFlags(&h->k1_stream, cudaStreamNonBlocking));
    A(cudaEventCreateWithFlags(&h->ev_lin_main, cudaEventDisableTiming));
    for (int q = 0; q < 2; ++q)
    {
        A(cudaEventCreateWithFlags(&h->ev_lin[q], cudaEventDisableTiming));
        A(cudaEventCreateWithFlags(&h->ev_qp[q], cudaEventDisableTiming));
    }
    {
        size_t wsd = (size_t)g.N * WS_STAGE;
        if (h->solver == 0 && condensed_ws_doubles(g) > wsd)
            wsd = condensed_ws_doubles(g);
        if (h->solver == SOLVER_WIDE)
            wsd = condensed_wide_ws_doubles(g);
        A(dalloc(&h->d_ws, wsd * B));
    }
    A(dalloc(&h->d_scratch, scratch * B));
    if (h->solver == SOLVER_WIDE)
    {
        h->wset_bytes = condensed_wide_wset_bytes(g) * B;
        A(dalloc(&h->d_wset, h->wset_bytes));
        if (ok)
            A(cudaMemset(h->d_wset, 0xFF, h->wset_bytes));     // all-lower vertex
    }
    if ((h->solver == 0 || h->solver == SOLVER_WIDE) && g.use_jl)
    {
        h->jlset_bytes = condensed_jlset_words() * sizeof(unsigned) * B;
        A(dalloc(&h->d_jlset, h->jlset_bytes / sizeof(unsigned)));
        if (ok)
            A(cudaMemset(h->d_jlset, 0, h->jlset_bytes));         // nothing clamped
    }
    A(dalloc(&h->d_z, (size_t)g.n_var * B));
    A(dalloc(&h->d_out, (size_t)VSMPC_OUT_DOUBLES * B));
    for (int q = 0; q < 2; ++q)
    {
        A(dalloc(&h->d_out_stage[q], (size_t)VSMPC_OUT_DOUBLES * B));
        A(dalloc(&h->d_status_stage[q], (size_t)B));
    }
    A(dalloc(&h->d_status, (size_t)B));
    A(dalloc(&h->d_nf, (size_t)B));
    A(dalloc(&h->d_ns, (size_t)B));
    A(dalloc(&h->d_np, (size_t)B));
    if ((h->solver == 0 || h->solver == SOLVER_WIDE) && fallback_supported(g))
    {
        // slots: concurrent fallback solves; a slot holds the dense [K | R] of one instance (9 MB at the reference horizon)
        const size_t per = fallback_slot_doubles(g) * sizeof(double);
        // one slot per 128 instances (0.2 % of a Monte Carlo sweep needs the fallback every tick: one round instead of three at
        // 65 536 instances), at most 512 slots and 6 GB of the 180 GB
        size_t slots = std::max<size_t>(2, std::min<size_t>(512, (size_t)B / 128));
        slots = std::max<size_t>(1, std::min(slots, ((size_t)6 << 30) / per));
        slots = std::min(slots, (size_t)B);
        h->fb_slots = (int)slots;
        std::vector<int> pos(fallback_pos_ints(g));
        fallback_positions(g, pos.data());
        for (int q = 0; q < 2; ++q)
        {
            A(dalloc(&h->d_fb_list2[q], (size_t)B));
            A(dalloc(&h->d_fb_count2[q], 1));
        }
        h->d_fb_list = h->d_fb_list2[0];
        h->d_fb_count = h->d_fb_count2[0];
        A(dalloc(&h->d_fb_pos, pos.size()));
        A(dalloc(&h->d_fb_scratch, fallback_slot_doubles(g) * slots));
        if (ok)
        {
            A(cudaMemcpy(h->d_fb_pos, pos.data(), pos.size() * sizeof(int), cudaMemcpyHostToDevice));
            A(cudaMemset(h->d_fb_count2[0], 0, sizeof(int)));
            A(cudaMemset(h->d_fb_count2[1], 0, sizeof(int)));
        }
        h->fb_mode = 1;
    }
    if (ok)
    {
        A(cudaMemcpy(h->d_cfg, &g, sizeof(g), cudaMemcpyHostToDevice));
        A(cudaMemcpy(h->d_alpha, alpha.data(), alpha.size() * 8, cudaMemcpyHostToDevice));
        A(cudaMemcpy(h->d_tpos, tp.data(), tp.size() * 8, cudaMemcpyHostToDevice));
        A(cudaMemcpy(h->d_tvel, tv.data(), tv.size() * 8, cudaMemcpyHostToDevice));
        A(cudaMemcpy(h->d_trpy, tr.data(), tr.size() * 8, cudaMemcpyHostToDevice));
        A(cudaMemcpy(h->d_trpyd, td.data(), td.size() * 8, cudaMemcpyHostToDevice));
        A(cudaMemset(h->d_out, 0, (size_t)VSMPC_OUT_DOUBLES * B * 8));
        A(cudaMemset(h->d_status, 0, (size_t)B * 4));
        A(cudaMemset(h->d_z, 0, (size_t)g.n_var * B * 8));
        A(cudaMemset(h->d_nf, 0, (size_t)B * 4));
        A(cudaMemset(h->d_ns, 0, (size_t)B * 4));
        A(cudaMemset(h->d_np, 0, (size_t)B * 4));
    }
    if (!ok)
    {
        std::string msg = std::string("device allocation/initialisation failed: ") + cudaGetErrorString(e);
        vsmpc_destroy(h);
        h = new (std::nothrow) vsmpc_handle();
        if (!h)
            return VSMPC_ERR_CUDA;
        return bail(VSMPC_ERR_CUDA, msg);
    }
    *out = h;
    return VSMPC_OK;
}

int vsmpc_destroy(vsmpc_handle* h)
{
    if (!h)
        return VSMPC_OK;
    if (h->B > 0)
        cudaSetDevice(h->device);
    void* ptrs[] = {h->d_cfg, h->d_pack, h->d_jpos, h->d_phase, h->d_st, h->d_si, h->d_alpha, h->d_tpos,
                    h->d_tvel, h->d_trpy, h->d_trpyd, h->d_qd2[0], h->d_qd2[1], h->d_ws, h->d_scratch, h->d_z, h->d_out,
                    h->d_status, h->d_nf, h->d_ns, h->d_np, h->d_fb_list2[0], h->d_fb_list2[1], h->d_fb_count2[0],
                    h->d_fb_count2[1], h->d_fb_pos, h->d_fb_scratch, h->d_pm, h->d_ps, h->d_pp, h->d_ip, h->d_jl, h->d_wset, h->d_jlset, h->d_kin, h->d_ks[0], h->d_ks[1],
                    h->d_out_stage[0], h->d_out_stage[1], h->d_status_stage[0], h->d_status_stage[1], h->d_pack_in[0], h->d_pack_in[1], h->d_nn, h->d_thr_sub};
    for (int q = 0; q < 2; ++q)
    {
        if (h->ev_h2d[q]) cudaEventDestroy(h->ev_h2d[q]);
        if (h->ev_k1[q]) cudaEventDestroy(h->ev_k1[q]);
        if (h->ev_out[q]) cudaEventDestroy(h->ev_out[q]);
    }
    for (int q = 0; q < 2; ++q)
    {
        if (h->ev_lin[q]) cudaEventDestroy(h->ev_lin[q]);
        if (h->ev_qp[q]) cudaEventDestroy(h->ev_qp[q]);
    }
    if (h->ev_lin_main)
        cudaEventDestroy(h->ev_lin_main);
    if (h->k1_stream)
        cudaStreamDestroy(h->k1_stream);
    if (h->copy_stream)
        cudaStreamDestroy(h->copy_stream);
    if (h->out_stream)
        cudaStreamDestroy(h->out_stream);
    if (h->ev_solved)
        cudaEventDestroy(h->ev_solved);
    if (h->ev_stream)
        cudaEventDestroy(h->ev_stream);
    drop_tick_graph(h);
    for (void* p : ptrs)
        if (p)
            cudaFree(p);
    if (h->own_stream)
        cudaStreamDestroy(h->own_stream);
    delete h;
    return VSMPC_OK;
}

int vsmpc_set_stream(vsmpc_handle* h, void* s)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    cudaStream_t ns = s ? reinterpret_cast<cudaStream_t>(s) : h->own_stream;
    if (ns != h->stream)
    { // work already queued on the old stream (a linearise kernel between set_state and solve) stays ordered before
      // whatever the new stream runs next
        CK(cudaSetDevice(h->device));
        CK(cudaEventRecord(h->ev_stream, h->stream));
        CK(cudaStreamWaitEvent(ns, h->ev_stream, 0));
    }
    h->stream = ns;
    return VSMPC_OK;
}

int vsmpc_n_var(const vsmpc_handle* h) { return h ? h->cfg.n_var : -1; }
int vsmpc_n_constraints(const vsmpc_handle* h) { return h ? h->cfg.n_con : -1; }
int vsmpc_n_instances(const vsmpc_handle* h) { return h ? h->B : -1; }

// the compute stream catches up with a linearise kernel that ran on k1_stream (before anything on it reads the QP data)
static int join_linearise(vsmpc_handle* h)
{
    if (h->lin_on_k1)
    {
        CK(cudaStreamWaitEvent(h->stream, h->ev_lin[h->qd_idx], 0));
        h->lin_on_k1 = false;
    }
    return VSMPC_OK;
}

// linearise kernel on the compute stream (device-resident packs, configure, rollouts): it fills the current QP buffer
static int run_linearise(vsmpc_handle* h, int mode)
{
    int rc = join_linearise(h);
    if (rc)
        return rc;
    if (mode == 1 && h->d_wset)     // IMPCProblem::configure: no previous solve, the guess is the all-lower vertex
        CK(cudaMemsetAsync(h->d_wset, 0xFF, h->wset_bytes, h->stream));
    if (mode == 1 && h->d_jlset)
        CK(cudaMemsetAsync(h->d_jlset, 0, h->jlset_bytes, h->stream));
    CK(launch_linearise(h->d_cfg, h->cfg, h->B, mode, h->d_pack, h->d_jpos, mode == 1 ? h->d_phase : nullptr, h->d_st,
                        h->d_si, h->d_alpha, h->d_tpos, h->d_tvel, h->d_trpy, h->d_trpyd, h->d_qd,
                        h->use_ip ? h->d_ip : nullptr, h->d_fb_count, h->use_jl_table ? h->d_jl : nullptr, h->stream));
    if (!h->capturing)
        CK(cudaEventRecord(h->ev_lin_main, h->stream));
    return VSMPC_OK;
}

// linearise kernel of a HOST pack staged in d_pack_in[q], on k1_stream, into QP buffer q: overlaps the QP kernel of the tick
// before on the compute stream.  Ordered behind: the copy of the pack (ev_h2d[q]), the QP / fallback kernels that read buffer
// q two ticks ago (ev_qp[q]), the last linearise kernel that ran on the compute stream (tick state), and — k1_stream being
// in order — the linearise kernel of the tick before.
static int run_linearise_overlapped(vsmpc_handle* h, int q, const double* pack_dev)
{
    static int no_overlap = -1;     // development: VSMPC_NO_OVERLAP=1 keeps the linearise kernel on the compute stream
    if (no_overlap < 0)
    {
        const char* e = getenv("VSMPC_NO_OVERLAP");
        no_overlap = e ? atoi(e) : 0;
    }
    if (no_overlap)
    {
        CK(cudaStreamWaitEvent(h->stream, h->ev_h2d[q], 0));
        double* saved = h->d_pack;
        h->d_pack = const_cast<double*>(pack_dev);
        const int rc = run_linearise(h, 0);
        h->d_pack = saved;
        if (rc)
            return rc;
        CK(cudaEventRecord(h->ev_k1[q], h->stream));
        return VSMPC_OK;
    }
    CK(cudaStreamWaitEvent(h->k1_stream, h->ev_h2d[q], 0));
    CK(cudaStreamWaitEvent(h->k1_stream, h->ev_qp[q], 0));
    CK(cudaStreamWaitEvent(h->k1_stream, h->ev_lin_main, 0));
    h->qd_idx = q;
    h->d_qd = h->d_qd2[q];
    if (h->d_fb_list2[q])
    {
        h->d_fb_list = h->d_fb_list2[q];
        h->d_fb_count = h->d_fb_count2[q];
    }
    drop_tick_graph(h);         // a captured rollout tick bakes the buffer pointers in
    CK(launch_linearise(h->d_cfg, h->cfg, h->B, 0, pack_dev, h->d_jpos, nullptr, h->d_st, h->d_si, h->d_alpha, h->d_tpos,
                        h->d_tvel, h->d_trpy, h->d_trpyd, h->d_qd, h->use_ip ? h->d_ip : nullptr, h->d_fb_count,
                        h->use_jl_table ? h->d_jl : nullptr, h->k1_stream));
    CK(cudaEventRecord(h->ev_lin[q], h->k1_stream));
    CK(cudaEventRecord(h->ev_k1[q], h->k1_stream));     // the pack staging buffer q is free again
    h->lin_on_k1 = true;
    return VSMPC_OK;
}

int vsmpc_configure(vsmpc_handle* h, const double* pack_host, const double* joint_pos_sel_host, const int* phase0_host)
{
    return vsmpc_configure_strided(h, pack_host, joint_pos_sel_host, phase0_host, h ? (size_t)h->B : 0);
}

int vsmpc_configure_strided(vsmpc_handle* h, const double* pack_host, const double* joint_pos_sel_host, const int* phase0_host,
                            size_t row_stride)
{
    if (!h || h->B <= 0 || !pack_host || !joint_pos_sel_host || row_stride < (size_t)h->B)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_configure: null argument or row stride smaller than the batch");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B;
    CK(copy_soa_h2d(h->d_pack, pack_host, VSMPC_PACK_DOUBLES, B, row_stride, h->stream));
    CK(copy_soa_h2d(h->d_jpos, joint_pos_sel_host, NJ, B, row_stride, h->stream));
    if (phase0_host)
    {
        for (size_t i = 0; i < B; ++i)
            if (phase0_host[i] < 0 || phase0_host[i] >= h->cfg.ratio)
                return fail(h, VSMPC_ERR_ARG, "vsmpc_configure: phase0 out of [0, ratio)");
        CK(cudaMemcpyAsync(h->d_phase, phase0_host, B * 4, cudaMemcpyHostToDevice, h->stream));
    }
    else
        CK(cudaMemsetAsync(h->d_phase, 0, B * 4, h->stream));
    int rc = run_linearise(h, 1);
    if (rc)
        return rc;
    seed_outputs_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->d_jpos, nullptr, h->d_out, h->d_status);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->configured = true;
    h->has_state = false;
    return VSMPC_OK;
}

int vsmpc_set_instance_params(vsmpc_handle* h, const double* ip_host)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    drop_tick_graph(h);
    if (!ip_host)
    {
        h->use_ip = false;
        return VSMPC_OK;
    }
    const size_t B = h->B;
    for (size_t i = 0; i < B; ++i)
    {
        const double c12 = ip_host[(size_t)(IP_JC + 12) * B + i], sd = ip_host[(size_t)(IP_JN + 3) * B + i];
        const double tmin = ip_host[(size_t)IP_TMIN * B + i], tmax = ip_host[(size_t)IP_TMAX * B + i];
        if (!(c12 != 0.0) || !(sd > 0.0) || !(ip_host[(size_t)(IP_JN + 1) * B + i] > 0.0) || !(tmin < tmax))
            return fail(h, VSMPC_ERR_ARG, "vsmpc_set_instance_params: need c12 != 0, positive standard deviations, throttle_min < throttle_max");
    }
    if (!h->d_ip)
        CK(dalloc(&h->d_ip, (size_t)IP_ROWS * B));
    CK(cudaMemcpyAsync(h->d_ip, ip_host, (size_t)IP_ROWS * B * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->use_ip = true;
    return VSMPC_OK;
}

int vsmpc_set_joint_limits(vsmpc_handle* h, const double* q_min_host, const double* q_max_host)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    if (!h->cfg.use_jl)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_set_joint_limits: the handle was created without use_joint_limits");
    CK(cudaSetDevice(h->device));
    drop_tick_graph(h);
    if (h->d_jlset)       // new boxes: the working set of the last solve is no guess for them
        CK(cudaMemsetAsync(h->d_jlset, 0, h->jlset_bytes, h->stream));
    if (!q_min_host && !q_max_host)
    {
        h->use_jl_table = false;
        return VSMPC_OK;
    }
    if (!q_min_host || !q_max_host)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_set_joint_limits: give both tables or neither");
    const size_t n = (size_t)NJ * h->B;
    for (size_t e = 0; e < n; ++e)
        if (!(q_min_host[e] < q_max_host[e]))
            return fail(h, VSMPC_ERR_ARG, "vsmpc_set_joint_limits: need q_min < q_max for every joint of every instance");
    if (!h->d_jl)
        CK(dalloc(&h->d_jl, 2 * n));
    CK(cudaMemcpyAsync(h->d_jl, q_min_host, n * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_jl + n, q_max_host, n * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    h->use_jl_table = true;
    return VSMPC_OK;
}

int vsmpc_host_alloc(size_t bytes, void** out)
{
    if (!out || bytes == 0)
        return VSMPC_ERR_ARG;
    *out = nullptr;
    return cudaHostAlloc(out, bytes, cudaHostAllocDefault) == cudaSuccess ? VSMPC_OK : VSMPC_ERR_CUDA;
}

int vsmpc_host_free(void* p)
{
    return (!p || cudaFreeHost(p) == cudaSuccess) ? VSMPC_OK : VSMPC_ERR_CUDA;
}

int vsmpc_set_warm_start(vsmpc_handle* h, int enable)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    if (h->warm != (enable != 0))
        drop_tick_graph(h);
    h->warm = enable != 0;
    return VSMPC_OK;
}

// ---- batched reduced kinematics: robot state in, pack built on the device -------------------------------------
int vsmpc_set_kin_model(vsmpc_handle* h, const vsmpc_kin_model* model)
{
    if (!h || h->B <= 0 || !model)
        return VSMPC_ERR_ARG;
    std::vector<unsigned char> buf(kin_model_bytes());
    KinModelDev& K = *reinterpret_cast<KinModelDev*>(buf.data());
    const char* why = kin_prepare(*model, K);
    if (why[0])
        return fail(h, VSMPC_ERR_ARG, std::string("vsmpc_set_kin_model: ") + why);
    CK(cudaSetDevice(h->device));
    drop_tick_graph(h);
    const int rows = kin_state_rows(K);
    if (!h->d_kin)
        CK(cudaMalloc(reinterpret_cast<void**>(&h->d_kin), buf.size()));
    if (rows != h->ks_rows)
        for (int q = 0; q < 2; ++q)
        {
            if (h->d_ks[q])
                cudaFree(h->d_ks[q]);
            h->d_ks[q] = nullptr;
            CK(dalloc(&h->d_ks[q], (size_t)rows * h->B));
        }
    h->ks_rows = rows;
    CK(cudaMemcpyAsync(h->d_kin, buf.data(), buf.size(), cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_kin_state_doubles(const vsmpc_handle* h) { return (h && h->ks_rows > 0) ? h->ks_rows : -1; }

int vsmpc_configure_kinematics(vsmpc_handle* h, const double* kin_state_host, const int* phase0_host)
{
    if (!h || h->B <= 0 || !kin_state_host)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_configure_kinematics: null argument");
    if (!h->d_kin)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_configure_kinematics: vsmpc_set_kin_model first");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B;
    CK(cudaMemcpyAsync(h->d_ks[0], kin_state_host, (size_t)h->ks_rows * B * 8, cudaMemcpyHostToDevice, h->stream));
    CK(launch_kinematics(h->d_kin, h->ks_rows, h->B, h->d_ks[0], h->d_pack, h->d_jpos, h->stream));
    if (phase0_host)
    {
        for (size_t i = 0; i < B; ++i)
            if (phase0_host[i] < 0 || phase0_host[i] >= h->cfg.ratio)
                return fail(h, VSMPC_ERR_ARG, "vsmpc_configure_kinematics: phase0 out of [0, ratio)");
        CK(cudaMemcpyAsync(h->d_phase, phase0_host, B * 4, cudaMemcpyHostToDevice, h->stream));
    }
    else
        CK(cudaMemsetAsync(h->d_phase, 0, B * 4, h->stream));
    int rc = run_linearise(h, 1);
    if (rc)
        return rc;
    seed_outputs_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->d_jpos, nullptr, h->d_out, h->d_status);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->configured = true;
    h->has_state = false;
    return VSMPC_OK;
}

int vsmpc_set_state_kinematics(vsmpc_handle* h, const double* kin_state_host)
{
    if (!h || h->B <= 0 || !kin_state_host)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_set_state_kinematics: null argument");
    if (!h->d_kin)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_set_state_kinematics: vsmpc_set_kin_model first");
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_set_state_kinematics: configure first");
    CK(cudaSetDevice(h->device));
    // same pipeline as vsmpc_set_state: the (four times smaller) robot state is staged on the copy stream, the kinematics
    // kernel builds the pack of this tick in the staging pack buffer, then the linearise kernel runs on it
    const int q = h->pack_idx ^= 1;
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_k1[q], 0));
    CK(cudaMemcpyAsync(h->d_ks[q], kin_state_host, (size_t)h->ks_rows * h->B * 8, cudaMemcpyHostToDevice, h->copy_stream));
    CK(cudaEventRecord(h->ev_h2d[q], h->copy_stream));
    // kinematics + linearise kernels on k1_stream: both overlap the QP kernel of the tick before (run_linearise_overlapped)
    CK(cudaStreamWaitEvent(h->k1_stream, h->ev_h2d[q], 0));
    CK(launch_kinematics(h->d_kin, h->ks_rows, h->B, h->d_ks[q], h->d_pack_in[q], nullptr, h->k1_stream));
    int rc = run_linearise_overlapped(h, q, h->d_pack_in[q]);
    if (rc)
        return rc;
    h->has_state = true;
    return VSMPC_OK;
}

int vsmpc_get_kinematics_pack(vsmpc_handle* h, double* pack_host)
{
    if (h && h->B > 0 && join_linearise(h) != VSMPC_OK)
        return VSMPC_ERR_CUDA;
    if (!h || h->B <= 0 || !pack_host)
        return VSMPC_ERR_ARG;
    if (!h->d_kin)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_get_kinematics_pack: vsmpc_set_kin_model first");
    CK(cudaSetDevice(h->device));
    const double* src = h->has_state ? h->d_pack_in[h->pack_idx] : h->d_pack;
    CK(cudaMemcpyAsync(pack_host, src, (size_t)VSMPC_PACK_DOUBLES * h->B * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_debug_set_working_set(vsmpc_handle* h, const signed char* wset_host)
{
    if (!h || h->B <= 0 || !wset_host)
        return VSMPC_ERR_ARG;
    if (!h->d_wset)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_debug_set_working_set: this horizon does not use the long-horizon kernel");
    CK(cudaSetDevice(h->device));
    const size_t nv = (size_t)NT * h->cfg.nblk, pitch = h->wset_bytes / h->B;
    CK(cudaMemcpy2DAsync(h->d_wset, pitch, wset_host, nv, nv, h->B, cudaMemcpyHostToDevice, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_set_state(vsmpc_handle* h, const double* pack_host)
{
    return vsmpc_set_state_strided(h, pack_host, h ? (size_t)h->B : 0);
}

int vsmpc_set_state_strided(vsmpc_handle* h, const double* pack_host, size_t row_stride)
{
    if (!h || h->B <= 0 || !pack_host || row_stride < (size_t)h->B)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_set_state: null argument or row stride smaller than the batch");
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_set_state: configure first");
    CK(cudaSetDevice(h->device));
    // stage the pack on the copy stream into the buffer the previous-but-one tick used (its linearise kernel must be
    // done with it), then make the compute stream wait for the copy: H2D of this tick overlaps the QP kernel of the last
    const int q = h->pack_idx ^= 1;
    CK(cudaStreamWaitEvent(h->copy_stream, h->ev_k1[q], 0));
    CK(copy_soa_h2d(h->d_pack_in[q], pack_host, VSMPC_PACK_DOUBLES, (size_t)h->B, row_stride, h->copy_stream));
    CK(cudaEventRecord(h->ev_h2d[q], h->copy_stream));
    int rc = run_linearise_overlapped(h, q, h->d_pack_in[q]);
    if (rc)
        return rc;
    h->has_state = true;
    return VSMPC_OK;
}

int vsmpc_set_state_device(vsmpc_handle* h, const double* pack_dev)
{
    if (!h || h->B <= 0 || !pack_dev)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_set_state_device: null argument");
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_set_state_device: configure first");
    CK(cudaSetDevice(h->device));
    double* saved = h->d_pack;
    h->d_pack = const_cast<double*>(pack_dev);
    int rc = run_linearise(h, 0);
    h->d_pack = saved;
    if (rc)
        return rc;
    h->has_state = true;
    return VSMPC_OK;
}

int vsmpc_solve_async(vsmpc_handle* h)
{
    if (!h || h->B <= 0)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_solve: bad handle");
    if (!h->has_state)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_solve: no state set since configure (call vsmpc_set_state)");
    CK(cudaSetDevice(h->device));
    {
        const int rc = solve_launch(h);
        if (rc)
            return rc;
    }
    h->has_state = false; // one solve per update, like the reference's tick
    return VSMPC_OK;
}

int vsmpc_wait(vsmpc_handle* h)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    {
        const int rc = join_linearise(h);      // a linearise kernel alone (vsmpc_linearise) runs on its own stream
        if (rc)
            return rc;
    }
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_solve(vsmpc_handle* h)
{
    int rc = vsmpc_solve_async(h);
    return rc ? rc : vsmpc_wait(h);
}

// SURVEY §8b names of the two inner seams (K1 alone, K2 alone); same host paths as the outer-surface entry points
int vsmpc_linearise(vsmpc_handle* h, const double* pack_host) { return vsmpc_set_state(h, pack_host); }
int vsmpc_solve_qp(vsmpc_handle* h) { return vsmpc_solve(h); }

int vsmpc_get_output(vsmpc_handle* h, double* out_rows_host, int* status_host)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (out_rows_host)
        CK(cudaMemcpyAsync(out_rows_host, h->d_out, (size_t)VSMPC_OUT_DOUBLES * h->B * 8, cudaMemcpyDeviceToHost, h->stream));
    if (status_host)
        CK(cudaMemcpyAsync(status_host, h->d_status, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_get_output_async(vsmpc_handle* h, double* out_rows_host, int* status_host, int* ticket)
{
    if (!h || h->B <= 0 || !ticket)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    // the rows are read back from a staging buffer on the read-back stream: neither the next linearise kernel nor the next QP
    // kernel (which rewrites d_out in place) waits for the PCIe copy.  The condensed kernels fill the staging buffer themselves
    // (solve_launch); after anything else (other solvers, rollouts, a second read-back of the same tick) a snapshot kernel does.
    const int q = h->out_idx ^= 1;
    const size_t nb_out = (size_t)VSMPC_OUT_DOUBLES * h->B * 8, nb_st = (size_t)h->B * 4;
    if (h->stage_written != q)
    {
        CK(cudaStreamWaitEvent(h->stream, h->ev_out[q], 0));   // the read-back that used this staging buffer two calls ago
        const int threads = 256;
        const int blocks = (int)std::min<size_t>(592, (nb_out / 16 + threads - 1) / threads);
        snapshot_outputs_kernel<<<blocks, threads, 0, h->stream>>>(nb_out / 8, h->B, h->d_out, h->d_status, h->d_out_stage[q],
                                                                   h->d_status_stage[q]);
        CK(cudaGetLastError());
    }
    h->stage_written = -1;
    CK(cudaEventRecord(h->ev_solved, h->stream));
    CK(cudaStreamWaitEvent(h->out_stream, h->ev_solved, 0));
    if (out_rows_host)
        CK(cudaMemcpyAsync(out_rows_host, h->d_out_stage[q], nb_out, cudaMemcpyDeviceToHost, h->out_stream));
    if (status_host)
        CK(cudaMemcpyAsync(status_host, h->d_status_stage[q], nb_st, cudaMemcpyDeviceToHost, h->out_stream));
    CK(cudaEventRecord(h->ev_out[q], h->out_stream));
    *ticket = q;
    return VSMPC_OK;
}

int vsmpc_wait_output(vsmpc_handle* h, int ticket)
{
    if (!h || h->B <= 0 || ticket < 0 || ticket > 1)
        return VSMPC_ERR_ARG;
    CK(cudaEventSynchronize(h->ev_out[ticket]));
    return VSMPC_OK;
}

int vsmpc_get_output_device(vsmpc_handle* h, double** out_rows_dev, int** status_dev)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    if (out_rows_dev)
        *out_rows_dev = h->d_out;
    if (status_dev)
        *status_dev = h->d_status;
    return VSMPC_OK;
}

int vsmpc_set_full_solution(vsmpc_handle* h, int enable)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    if (h->want_full != (enable != 0))
        drop_tick_graph(h);
    h->want_full = enable != 0;
    return VSMPC_OK;
}

int vsmpc_get_full_solution(vsmpc_handle* h, double* z_host)
{
    if (!h || h->B <= 0 || !z_host)
        return VSMPC_ERR_ARG;
    if (h->solver != 1 && !h->want_full)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_get_full_solution: enable it first with vsmpc_set_full_solution(h, 1)");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(z_host, h->d_z, (size_t)h->cfg.n_var * h->B * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_get_dynamics(vsmpc_handle* h, double* A, double* BJ, double* BT, double* c, double* dt)
{
    if (h && h->B > 0 && join_linearise(h) != VSMPC_OK)
        return VSMPC_ERR_CUDA;
    if (!h || h->B <= 0 || !A || !BJ || !BT || !c)
        return VSMPC_ERR_ARG;
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_get_dynamics: configure first");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B;
    double *dA = nullptr, *dBJ = nullptr, *dBT = nullptr, *dc = nullptr; // freed on every path below
    cudaError_t e = dalloc(&dA, B * NX * NX);
    if (e == cudaSuccess) e = dalloc(&dBJ, B * NX * NJ);
    if (e == cudaSuccess) e = dalloc(&dBT, B * NX * NT);
    if (e == cudaSuccess) e = dalloc(&dc, B * NX);
    if (e == cudaSuccess) e = launch_expand_dynamics(h->d_cfg, h->B, h->d_qd, dA, dBJ, dBT, dc, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(A, dA, B * NX * NX * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(BJ, dBJ, B * NX * NJ * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(BT, dBT, B * NX * NT * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(c, dc, B * NX * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dA); cudaFree(dBJ); cudaFree(dBT); cudaFree(dc);
    if (e != cudaSuccess)
        return cuda_fail(h, e, "vsmpc_get_dynamics");
    if (dt)
        std::memcpy(dt, h->cfg.dt, sizeof(double) * h->cfg.N);
    return VSMPC_OK;
}

int vsmpc_get_qp_vectors(vsmpc_handle* h, double* q, double* l, double* u)
{
    if (h && h->B > 0 && join_linearise(h) != VSMPC_OK)
        return VSMPC_ERR_CUDA;
    if (!h || h->B <= 0 || !q || !l || !u)
        return VSMPC_ERR_ARG;
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_get_qp_vectors: configure first");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B;
    double *dq = nullptr, *dl = nullptr, *du = nullptr; // freed on every path below
    cudaError_t e = dalloc(&dq, B * h->cfg.n_var);
    if (e == cudaSuccess) e = dalloc(&dl, B * h->cfg.n_con);
    if (e == cudaSuccess) e = dalloc(&du, B * h->cfg.n_con);
    if (e == cudaSuccess) e = launch_expand_qp_vectors(h->d_cfg, h->B, h->d_qd, dq, dl, du, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(q, dq, B * h->cfg.n_var * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(l, dl, B * h->cfg.n_con * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(u, du, B * h->cfg.n_con * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(dq); cudaFree(dl); cudaFree(du);
    if (e != cudaSuccess)
        return cuda_fail(h, e, "vsmpc_get_qp_vectors");
    return VSMPC_OK;
}

int vsmpc_get_counts(vsmpc_handle* h, int* n_factor, int* n_solve)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    if (n_factor)
        CK(cudaMemcpyAsync(n_factor, h->d_nf, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    if (n_solve)
        CK(cudaMemcpyAsync(n_solve, h->d_ns, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_get_pivot_counts(vsmpc_handle* h, int* n_pivot)
{
    if (!h || h->B <= 0 || !n_pivot)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(n_pivot, h->d_np, (size_t)h->B * 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_get_references(vsmpc_handle* h, double* refs_host)
{
    if (h && h->B > 0 && join_linearise(h) != VSMPC_OK)
        return VSMPC_ERR_CUDA;
    if (!h || h->B <= 0 || !refs_host)
        return VSMPC_ERR_ARG;
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_get_references: configure first");
    CK(cudaSetDevice(h->device));
    // the four published fields are 13 consecutive rows of the device-resident tick state (SoA): one copy, transposed here
    static_assert(ST_RPY_REF == ST_P_REF + 3 && ST_MOM_REF == ST_P_REF + 6 && ST_ALPHA == ST_P_REF + 12, "row order");
    const size_t B = h->B;
    std::vector<double> rows(13 * B);
    CK(cudaMemcpyAsync(rows.data(), h->d_st + (size_t)ST_P_REF * B, 13 * B * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (size_t i = 0; i < B; ++i)
    {
        double* o = refs_host + i * VSMPC_REF_DOUBLES;
        o[VSMPC_REF_ALPHA_GRAVITY] = rows[12 * B + i];
        for (int a = 0; a < 12; ++a)
            o[VSMPC_REF_POS_COM + a] = rows[(size_t)a * B + i];
    }
    return VSMPC_OK;
}

int vsmpc_get_hessian(vsmpc_handle* h, int instance, double* P_host)
{
    if (!h || h->B <= 0 || !P_host || instance < 0 || instance >= h->B)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_get_hessian: bad argument");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cfg.n_var * h->cfg.n_var;
    double* d = nullptr;
    cudaError_t e = dalloc(&d, n);
    if (e == cudaSuccess) e = launch_expand_hessian(h->d_cfg, h->cfg, d, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(P_host, d, n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess)
        return cuda_fail(h, e, "vsmpc_get_hessian");
    return VSMPC_OK;
}

int vsmpc_get_constraint_matrix(vsmpc_handle* h, int instance, double* A_host)
{
    if (h && h->B > 0 && join_linearise(h) != VSMPC_OK)
        return VSMPC_ERR_CUDA;
    if (!h || h->B <= 0 || !A_host || instance < 0 || instance >= h->B)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_get_constraint_matrix: bad argument");
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_get_constraint_matrix: configure first");
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)h->cfg.n_con * h->cfg.n_var, ns = NX * NX + NX * NJ + NX * NT + NX;
    double* d = nullptr;
    cudaError_t e = dalloc(&d, n + ns);
    if (e == cudaSuccess)
        e = launch_expand_constraint_matrix(h->d_cfg, h->cfg, h->d_qd + (size_t)instance * h->cfg.qd_stride, d + n, d, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(A_host, d, n * 8, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess)
        return cuda_fail(h, e, "vsmpc_get_constraint_matrix");
    return VSMPC_OK;
}

int vsmpc_set_fallback(vsmpc_handle* h, int mode)
{
    if (!h || h->B <= 0 || mode < 0 || mode > 2)
        return VSMPC_ERR_ARG;
    if (mode != 0 && !h->d_fb_list)
        return fail(h, VSMPC_ERR_UNSUPPORTED, "vsmpc_set_fallback: the fallback QP kernel belongs to the default solver");
    if (mode != h->fb_mode)
        drop_tick_graph(h);
    h->fb_mode = mode;
    return VSMPC_OK;
}

int vsmpc_debug_phase_clocks(long long* clocks_host, int n_instances)
{
    return g_last_qp_solver.load(std::memory_order_relaxed) == SOLVER_WIDE ? condensed_wide_phase_clocks(clocks_host, n_instances)
                                           : condensed_phase_clocks(clocks_host, n_instances);
}

int vsmpc_debug_set_counters(vsmpc_handle* h, int ref_counter, int throttle_counter)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    if (!h->configured)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_debug_set_counters: configure first");
    CK(cudaSetDevice(h->device));
    std::vector<int> v(h->B);
    if (ref_counter >= 0)
    {
        std::fill(v.begin(), v.end(), ref_counter);
        CK(cudaMemcpy(h->d_si + (size_t)SI_REF_COUNTER * h->B, v.data(), (size_t)h->B * 4, cudaMemcpyHostToDevice));
    }
    if (throttle_counter >= 0)
    {
        std::fill(v.begin(), v.end(), throttle_counter);
        CK(cudaMemcpy(h->d_si + (size_t)SI_THR_COUNTER * h->B, v.data(), (size_t)h->B * 4, cudaMemcpyHostToDevice));
    }
    return VSMPC_OK;
}

// ---- device-resident closed loop ----------------------------------------------------------------------------------
static_assert(sizeof(PlantModel) == sizeof(vsmpc_plant_model), "PlantModel must mirror vsmpc_plant_model");

static int solve_launch(vsmpc_handle* h)
{
    {
        const int rc = join_linearise(h);      // the QP kernel waits for the linearise kernel wherever it ran
        if (rc)
            return rc;
    }
    g_last_qp_solver.store(h->solver, std::memory_order_relaxed);
    // the condensed kernels (and the fallback kernel behind them) also write every instance's row + status into the staging
    // buffer the next vsmpc_get_output_async will read back: no snapshot kernel between this tick and the next.  The read-back
    // that used this staging buffer two calls ago must be over.
    double* out2 = nullptr;
    int* status2 = nullptr;
    h->stage_written = -1;
    if (!h->capturing && (h->solver == 0 || h->solver == SOLVER_WIDE))
    {
        const int q = h->out_idx ^ 1;
        CK(cudaStreamWaitEvent(h->stream, h->ev_out[q], 0));
        out2 = h->d_out_stage[q];
        status2 = h->d_status_stage[q];
        h->stage_written = q;
    }
    if (h->solver == 0)
        CK(launch_qp_condensed(h->d_cfg, h->cfg, h->B, h->d_qd, h->d_ws, h->d_z, h->d_st, h->d_out, h->d_status,
                               h->d_nf, h->d_ns, h->d_np, h->want_full ? 1 : 0, h->d_fb_list, h->d_fb_count, h->fb_mode,
                               out2, status2, h->warm ? h->d_jlset : nullptr, h->stream));
    else if (h->solver == SOLVER_WIDE)
        CK(launch_qp_condensed_wide(h->cfg, h->B, h->d_qd, h->d_ws, h->d_scratch, h->d_z, h->d_st, h->d_out, h->d_status,
                                    h->d_nf, h->d_ns, h->d_np, h->want_full ? 1 : 0, h->d_fb_list, h->d_fb_count,
                                    h->fb_mode, h->d_wset, h->warm ? 1 : 0, out2, status2, h->warm ? h->d_jlset : nullptr, h->stream));
    else if (h->solver == 1)
        CK(launch_qp_generic(h->d_cfg, h->cfg, h->B, h->d_qd, h->d_ws, h->d_scratch, h->d_z, h->d_st, h->d_out,
                             h->d_status, h->d_nf, h->d_ns, h->stream));
    else
        CK(launch_qp_structured(h->d_cfg, h->cfg, h->B, h->d_qd, h->d_ws, h->d_scratch, h->d_z, h->d_st, h->d_out,
                                h->d_status, h->d_nf, h->d_ns, h->want_full ? 1 : 0, h->stream));
    if (h->fb_mode != 0 && h->d_fb_list && (h->solver == 0 || h->solver == SOLVER_WIDE))
        CK(launch_qp_fallback(h->cfg, h->B, h->fb_slots, h->d_qd, h->d_fb_list, h->d_fb_count, h->d_fb_pos, h->d_fb_scratch,
                              h->d_z, h->d_st, h->d_out, h->d_status, h->d_nf, h->d_ns, h->d_np, h->want_full ? 1 : 0,
                              out2, status2, h->stream));
    if (!h->capturing)
        CK(cudaEventRecord(h->ev_qp[h->qd_idx], h->stream));     // QP buffer qd_idx may be refilled after this point
    return VSMPC_OK;
}

int vsmpc_rollout_init(vsmpc_handle* h, const vsmpc_plant_model* model, const double* plant_state_host,
                       const double* plant_param_host, const double* joint_pos_sel_host, const int* phase0_host)
{
    if (!h || h->B <= 0 || !model || !plant_state_host || !plant_param_host || !joint_pos_sel_host)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_init: null argument");
    if (!(model->dt_sim > 0) || model->n_sub < 1 || model->n_sub > MAX_SUB)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_init: dt_sim must be positive and 1 <= n_sub <= 16");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B;
    if (!h->d_pm)
    {
        CK(dalloc(&h->d_pm, 1));
        CK(dalloc(&h->d_ps, (size_t)PS_ROWS * B));
        CK(dalloc(&h->d_pp, (size_t)PP_ROWS * B));
    }
    drop_tick_graph(h);
    CK(cudaMemcpyAsync(h->d_pm, model, sizeof(PlantModel), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_ps, plant_state_host, PS_ROWS * B * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_pp, plant_param_host, PP_ROWS * B * 8, cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_jpos, joint_pos_sel_host, NJ * B * 8, cudaMemcpyHostToDevice, h->stream));
    if (phase0_host)
    {
        for (size_t i = 0; i < B; ++i)
            if (phase0_host[i] < 0 || phase0_host[i] >= h->cfg.ratio)
                return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_init: phase0 out of [0, ratio)");
        CK(cudaMemcpyAsync(h->d_phase, phase0_host, B * 4, cudaMemcpyHostToDevice, h->stream));
    }
    else
        CK(cudaMemsetAsync(h->d_phase, 0, B * 4, h->stream));
    // first pack from the plant state, then IMPCProblem::configure on it (tick 0 of every counter)
    CK(launch_plant(h->d_cfg, h->d_pm, h->B, 0, h->d_ps, h->d_pp, h->d_out, h->d_status, h->d_pack, nullptr,
                    h->use_ip ? h->d_ip : nullptr, nullptr, nullptr, h->stream));
    int rc = run_linearise(h, 1);
    if (rc)
        return rc;
    seed_outputs_kernel<<<(h->B + 127) / 128, 128, 0, h->stream>>>(h->B, h->d_jpos, h->d_ps, h->d_out, h->d_status);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(h->stream));
    h->configured = true;
    h->has_state = false;
    h->rollout_ready = true;
    return VSMPC_OK;
}

static int tick_launch(vsmpc_handle* h, double* rec)
{
    int rc = run_linearise(h, 0);
    if (rc)
        return rc;
    rc = solve_launch(h);
    if (rc)
        return rc;
    if (h->use_nn)
        CK(launch_jet_nn_ekf(h->d_cfg, h->d_pm, h->d_nn, h->B, h->d_ps, h->d_out, h->use_ip ? h->d_ip : nullptr, h->d_thr_sub,
                             h->stream));
    CK(launch_plant(h->d_cfg, h->d_pm, h->B, 1, h->d_ps, h->d_pp, h->d_out, h->d_status, h->d_pack, rec,
                    h->use_ip ? h->d_ip : nullptr, h->d_st, h->use_nn ? h->d_thr_sub : nullptr, h->stream));
    return VSMPC_OK;
}

int vsmpc_rollout_run(vsmpc_handle* h, int n_ticks, int record_every, double* rec_host, int use_graph)
{
    if (!h || h->B <= 0 || n_ticks < 0 || record_every < 0)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_run: bad argument");
    if (!h->rollout_ready)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_rollout_run: call vsmpc_rollout_init first");
    if (record_every > 0 && !rec_host)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_run: rec_host missing");
    CK(cudaSetDevice(h->device));
    const size_t B = h->B;
    const int n_rec = record_every > 0 ? n_ticks / record_every : 0;
    double* d_rec = nullptr;
    if (n_rec > 0)
        CK(dalloc(&d_rec, (size_t)n_rec * B * PLANT_REC));
    {
        const int rcj = join_linearise(h);      // outside any capture: a linearise kernel still running on k1_stream
        if (rcj)
            return rcj;
    }
    if (use_graph && !h->tick_graph)
    {
        cudaGraph_t g = nullptr;
        cudaError_t e = cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal);
        h->capturing = true;
        int rc = e == cudaSuccess ? tick_launch(h, nullptr) : VSMPC_ERR_CUDA;
        h->capturing = false;
        cudaError_t e2 = cudaStreamEndCapture(h->stream, &g);
        if (e != cudaSuccess || rc != VSMPC_OK || e2 != cudaSuccess)
        {
            if (g)
                cudaGraphDestroy(g);
            if (d_rec)
                cudaFree(d_rec);
            return fail(h, VSMPC_ERR_CUDA, "vsmpc_rollout_run: graph capture failed");
        }
        e = cudaGraphInstantiate(&h->tick_graph, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess)
        {
            if (d_rec)
                cudaFree(d_rec);
            return cuda_fail(h, e, "cudaGraphInstantiate");
        }
    }
    int rc = VSMPC_OK;
    int rec_i = 0;
    for (int t = 0; t < n_ticks && rc == VSMPC_OK; ++t)
    {
        const bool record = record_every > 0 && (t + 1) % record_every == 0 && rec_i < n_rec;
        if (use_graph && !record)
        {
            cudaError_t e = cudaGraphLaunch(h->tick_graph, h->stream);
            if (e != cudaSuccess)
                rc = cuda_fail(h, e, "cudaGraphLaunch");
        }
        else
        {
            rc = tick_launch(h, record ? d_rec + (size_t)rec_i * B * PLANT_REC : nullptr);
            if (record)
                ++rec_i;
        }
    }
    cudaError_t e = cudaSuccess;
    if (rc == VSMPC_OK && n_rec > 0)
        e = cudaMemcpyAsync(rec_host, d_rec, (size_t)n_rec * B * PLANT_REC * 8, cudaMemcpyDeviceToHost, h->stream);
    cudaError_t e3 = cudaStreamSynchronize(h->stream);
    if (d_rec)
        cudaFree(d_rec);
    if (rc != VSMPC_OK)
        return rc;
    if (e != cudaSuccess)
        return cuda_fail(h, e, "rollout record copy");
    if (e3 != cudaSuccess)
        return cuda_fail(h, e3, "vsmpc_rollout_run");
    h->has_state = false;
    return VSMPC_OK;
}

int vsmpc_rollout_set_jet_nn(vsmpc_handle* h, const float* w_ih, const float* b_ih, const float* b_hh, const float* fc_w,
                             const float* fc_b, const double* norm, const double* ekf_R, const double* ekf_Q)
{
    if (!h || h->B <= 0)
        return VSMPC_ERR_ARG;
    CK(cudaSetDevice(h->device));
    drop_tick_graph(h);
    if (!w_ih)
    {
        h->use_nn = false;
        return VSMPC_OK;
    }
    if (!b_ih || !b_hh || !fc_w || !fc_b || !norm || !ekf_R || !ekf_Q)
        return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_set_jet_nn: null argument");
    if (!(norm[1] > 0.0) || !(norm[3] > 0.0))
        return fail(h, VSMPC_ERR_ARG, "vsmpc_rollout_set_jet_nn: normalisation standard deviations must be positive");
    JetNN nn;
    std::memcpy(nn.w_ih, w_ih, sizeof(nn.w_ih));
    for (int e = 0; e < 4 * NN_HID; ++e)
        nn.b[e] = b_ih[e] + b_hh[e];
    std::memcpy(nn.fc_w, fc_w, sizeof(nn.fc_w));
    nn.fc_b = fc_b[0];
    nn.pad = 0.f;
    std::memcpy(nn.norm, norm, sizeof(nn.norm));
    std::memcpy(nn.R, ekf_R, sizeof(nn.R));
    std::memcpy(nn.Q, ekf_Q, sizeof(nn.Q));
    if (!h->d_nn)
    {
        CK(dalloc(&h->d_nn, 1));
        CK(dalloc(&h->d_thr_sub, (size_t)2 * MAX_SUB * NT * h->B));
    }
    CK(cudaMemcpy(h->d_nn, &nn, sizeof(nn), cudaMemcpyHostToDevice));
    h->use_nn = true;
    return VSMPC_OK;
}

int vsmpc_jet_nn_eval(vsmpc_handle* h, int n_groups, double dt, const float* T_host, const float* throttle_host, float* T_next_host,
                      float* T_dot_host)
{
    if (!h || h->B <= 0 || n_groups <= 0 || !T_host || !throttle_host || !T_next_host || !T_dot_host)
        return VSMPC_ERR_ARG;
    if (!h->use_nn)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_jet_nn_eval: call vsmpc_rollout_set_jet_nn first");
    CK(cudaSetDevice(h->device));
    float* d = nullptr;
    const size_t n = (size_t)n_groups * NT;
    CK(dalloc(&d, 4 * n));
    cudaError_t e = cudaMemcpyAsync(d, T_host, n * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d + n, throttle_host, n * 4, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = launch_jet_nn_eval(h->d_nn, n_groups, (float)dt, d, d + n, d + 2 * n, d + 3 * n, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(T_next_host, d + 2 * n, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(T_dot_host, d + 3 * n, n * 4, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d);
    if (e != cudaSuccess)
        return cuda_fail(h, e, "vsmpc_jet_nn_eval");
    return VSMPC_OK;
}

int vsmpc_rollout_get_state(vsmpc_handle* h, double* plant_state_host)
{
    if (!h || h->B <= 0 || !plant_state_host)
        return VSMPC_ERR_ARG;
    if (!h->rollout_ready)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_rollout_get_state: call vsmpc_rollout_init first");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(plant_state_host, h->d_ps, (size_t)PS_ROWS * h->B * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

int vsmpc_rollout_get_pack(vsmpc_handle* h, double* pack_host)
{
    if (!h || h->B <= 0 || !pack_host)
        return VSMPC_ERR_ARG;
    if (!h->rollout_ready)
        return fail(h, VSMPC_ERR_STATE, "vsmpc_rollout_get_pack: call vsmpc_rollout_init first");
    CK(cudaSetDevice(h->device));
    CK(cudaMemcpyAsync(pack_host, h->d_pack, (size_t)VSMPC_PACK_DOUBLES * h->B * 8, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return VSMPC_OK;
}

// ---- one process, several GPUs: contiguous instance ranges, one handle + stream per device ---------------
struct vsmpc_multi
{
    std::vector<vsmpc_handle*> h;
    std::vector<int> lo;       // first instance of every shard, and B at the end
    int B = 0;
    std::string err;
};

static int mfail(vsmpc_multi* m, int code, const std::string& msg)
{
    if (m)
        m->err = msg;
    return code;
}

int vsmpc_create_multi(const vsmpc_config* cfg, int n_instances, int n_gpus, const int* devices, vsmpc_multi** out)
{
    if (!cfg || !out || n_instances <= 0 || n_gpus <= 0)
        return VSMPC_ERR_ARG;
    *out = nullptr;
    vsmpc_multi* m = new (std::nothrow) vsmpc_multi();
    if (!m)
        return VSMPC_ERR_ARG;
    *out = m;      // kept alive on failure so that the message can be read
    m->B = n_instances;
    // shard g owns [g B / G, (g + 1) B / G); a device without instances (B < G) holds no handle
    for (int g = 0; g <= n_gpus; ++g)
        m->lo.push_back((int)(((long long)g * n_instances) / n_gpus));
    m->h.assign(n_gpus, nullptr);
    for (int g = 0; g < n_gpus; ++g)
    {
        const int n = m->lo[g + 1] - m->lo[g];
        if (n == 0)
            continue;
        const int rc = vsmpc_create(cfg, n, devices ? devices[g] : g, &m->h[g]);
        if (rc != VSMPC_OK)
        {
            m->err = std::string("shard ") + std::to_string(g) + ": " + (m->h[g] ? vsmpc_last_error(m->h[g]) : "vsmpc_create failed");
            for (vsmpc_handle*& q : m->h)
            {
                if (q)
                    vsmpc_destroy(q);
                q = nullptr;
            }
            m->B = 0;
            return rc;
        }
    }
    return VSMPC_OK;
}

int vsmpc_multi_destroy(vsmpc_multi* m)
{
    if (!m)
        return VSMPC_OK;
    for (vsmpc_handle* q : m->h)
        if (q)
            vsmpc_destroy(q);
    delete m;
    return VSMPC_OK;
}

const char* vsmpc_multi_last_error(const vsmpc_multi* m) { return m ? m->err.c_str() : "null handle"; }
int vsmpc_multi_n_shards(const vsmpc_multi* m) { return m ? (int)m->h.size() : -1; }
int vsmpc_multi_n_instances(const vsmpc_multi* m) { return m ? m->B : -1; }

int vsmpc_multi_shard(const vsmpc_multi* m, int shard, int* first, int* count, vsmpc_handle** handle)
{
    if (!m || shard < 0 || shard >= (int)m->h.size())
        return VSMPC_ERR_ARG;
    if (first)
        *first = m->lo[shard];
    if (count)
        *count = m->lo[shard + 1] - m->lo[shard];
    if (handle)
        *handle = m->h[shard];
    return VSMPC_OK;
}

#define MEACH(call, what)                                                                                     \
    for (size_t g = 0; g < m->h.size(); ++g)                                                                  \
        if (m->h[g])                                                                                          \
        {                                                                                                     \
            const int lo = m->lo[g];                                                                          \
            (void)lo;                                                                                         \
            const int rc__ = (call);                                                                          \
            if (rc__ != VSMPC_OK)                                                                             \
                return mfail(m, rc__, std::string(what) + ", shard " + std::to_string(g) + ": " + vsmpc_last_error(m->h[g])); \
        }

int vsmpc_multi_configure(vsmpc_multi* m, const double* pack_host, const double* joint_pos_sel_host, const int* phase0_host)
{
    if (!m || m->B <= 0 || !pack_host || !joint_pos_sel_host)
        return mfail(m, VSMPC_ERR_ARG, "vsmpc_multi_configure: null argument");
    MEACH(vsmpc_configure_strided(m->h[g], pack_host + lo, joint_pos_sel_host + lo, phase0_host ? phase0_host + lo : nullptr,
                                  (size_t)m->B), "vsmpc_multi_configure");
    return VSMPC_OK;
}

int vsmpc_multi_set_instance_params(vsmpc_multi* m, const double* instance_params_host)
{
    if (!m || m->B <= 0)
        return VSMPC_ERR_ARG;
    for (size_t g = 0; g < m->h.size(); ++g)
    {
        if (!m->h[g])
            continue;
        if (!instance_params_host)
        {
            vsmpc_set_instance_params(m->h[g], nullptr);
            continue;
        }
        // the per-device call takes a contiguous table: gather the shard's columns
        const int lo = m->lo[g], n = m->lo[g + 1] - lo;
        std::vector<double> t((size_t)IP_ROWS * n);
        for (int r = 0; r < IP_ROWS; ++r)
            std::memcpy(t.data() + (size_t)r * n, instance_params_host + (size_t)r * m->B + lo, sizeof(double) * n);
        const int rc = vsmpc_set_instance_params(m->h[g], t.data());
        if (rc != VSMPC_OK)
            return mfail(m, rc, std::string("vsmpc_multi_set_instance_params, shard ") + std::to_string(g) + ": " + vsmpc_last_error(m->h[g]));
    }
    return VSMPC_OK;
}

// update on every device: the copies and linearise kernels of all shards are in flight together
int vsmpc_multi_set_state(vsmpc_multi* m, const double* pack_host)
{
    if (!m || m->B <= 0 || !pack_host)
        return mfail(m, VSMPC_ERR_ARG, "vsmpc_multi_set_state: null argument");
    MEACH(vsmpc_set_state_strided(m->h[g], pack_host + lo, (size_t)m->B), "vsmpc_multi_set_state");
    return VSMPC_OK;
}

int vsmpc_multi_solve_async(vsmpc_multi* m)
{
    if (!m || m->B <= 0)
        return VSMPC_ERR_ARG;
    MEACH(vsmpc_solve_async(m->h[g]), "vsmpc_multi_solve");
    return VSMPC_OK;
}

int vsmpc_multi_wait(vsmpc_multi* m)
{
    if (!m || m->B <= 0)
        return VSMPC_ERR_ARG;
    MEACH(vsmpc_wait(m->h[g]), "vsmpc_multi_wait");
    return VSMPC_OK;
}

int vsmpc_multi_solve(vsmpc_multi* m)
{
    const int rc = vsmpc_multi_solve_async(m);
    return rc ? rc : vsmpc_multi_wait(m);
}

// the only gather of the path: every device copies its contiguous range of rows into the caller's arrays
int vsmpc_multi_get_output(vsmpc_multi* m, double* out_rows_host, int* status_host)
{
    if (!m || m->B <= 0)
        return VSMPC_ERR_ARG;
    int tickets[64];
    if (m->h.size() > 64)
        return mfail(m, VSMPC_ERR_ARG, "vsmpc_multi_get_output: more than 64 shards");
    MEACH(vsmpc_get_output_async(m->h[g], out_rows_host ? out_rows_host + (size_t)lo * VSMPC_OUT_DOUBLES : nullptr,
                                 status_host ? status_host + lo : nullptr, &tickets[g]), "vsmpc_multi_get_output");
    MEACH(vsmpc_wait_output(m->h[g], tickets[g]), "vsmpc_multi_get_output");
    return VSMPC_OK;
}

int vsmpc_multi_set_full_solution(vsmpc_multi* m, int enable)
{
    if (!m || m->B <= 0)
        return VSMPC_ERR_ARG;
    MEACH(vsmpc_set_full_solution(m->h[g], enable), "vsmpc_multi_set_full_solution");
    return VSMPC_OK;
}

int vsmpc_multi_get_full_solution(vsmpc_multi* m, double* z_host)
{
    if (!m || m->B <= 0 || !z_host)
        return VSMPC_ERR_ARG;
    MEACH(vsmpc_get_full_solution(m->h[g], z_host + (size_t)lo * vsmpc_n_var(m->h[g])), "vsmpc_multi_get_full_solution");
    return VSMPC_OK;
}
#undef MEACH

} // extern "C"
