// Shared layouts and constants of the vsmpc kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "vsmpc.h"

namespace vsmpc
{

constexpr int NX = VSMPC_NX; // 26
constexpr int NJ = VSMPC_NJ; // 8
constexpr int NT = VSMPC_NT; // 4
constexpr int NY = NX + NT;  // (x, v)      30
constexpr int NZ = NY + NJ;  // (x, v, dq)  38
constexpr int NU = NT + NJ;  // (v, dq)     12
constexpr int MAX_ITER = VSMPC_MAX_ITER;

// state layout, MPC/include/variableSamplingMPC/VSconstant.h:9-16
constexpr int IX_COM = 0, IX_LIN = 3, IX_RPY = 6, IX_ANG = 9, IX_T = 12, IX_TD = 16, IX_EP = 20, IX_ER = 23;

// ---- QP data block written by the linearise kernel, read by the QP kernel -----------------------
// per instance, contiguous (instance-major, stride = qd_stride doubles): the ~120 structural
// nonzeros of (A, B_J, B_T, c) + x0 + references + gradient pieces + bounds.
constexpr int QD_RM = 0;      // 9   A[com, lin]   = wRb / m
constexpr int QD_OMEGA = 9;   // 3   omega_B  (A[lin,lin] = A[ang,ang] = -S(omega_B))
constexpr int QD_ALIN = 12;   // 12  A[lin, T]     = A_momBody[0:3,:]   (3x4)
constexpr int QD_WI = 24;     // 9   A[rpy, ang]   = W^-1 I^-1
constexpr int QD_AANG = 33;   // 12  A[ang, T]     = A_momBody[3:6,:]
constexpr int QD_JA = 45;     // 4   A[Td_i, T_i]  = dh/dT
constexpr int QD_JB = 49;     // 4   A[Td_i, Td_i] = dh/dTdot
constexpr int QD_JG = 53;     // 4   B_T[Td_i, i]  = sigma_T g
constexpr int QD_LLIN = 57;   // 24  B_J[lin, :]   = Lambda_lin (3x8)
constexpr int QD_LANG = 81;   // 24  B_J[ang, :]   = Lambda_ang (3x8)
constexpr int QD_CL = 105;    // 3   c[lin]
constexpr int QD_CTD = 108;   // 4   c[Td]
constexpr int QD_CEP = 112;   // 3   c[posErr] = -p_ref
constexpr int QD_CER = 115;   // 3   c[rpyErr] = -rpy_init
constexpr int QD_X0 = 118;    // 26  measured initial state
constexpr int QD_GQ = 144;    // 8   gradient of every dq block
constexpr int QD_VBAR = 152;  // 4   v(std(u_prev))
constexpr int QD_VMIN = 156;  // 1
constexpr int QD_VMAX = 157;  // 1
constexpr int QD_PINNED = 158; // 1  1.0 when throttle block 0 is pinned to vbar this tick
constexpr int QD_JTT = 159;   // 1   A[T_i, Td_i] (1 with jet dynamics, 0 without)
constexpr int QD_JGT = 160;   // 1   B_T[T_i, i]  (0 with jet dynamics, 1 without)
constexpr int QD_JC12 = 161;  // 1   jet coefficient c12 of this instance   } throttle de-standardisation
constexpr int QD_UMEAN = 162; // 1   throttle mean                           } (JetModel.cpp:93-109) in the
constexpr int QD_USTD = 163;  // 1   throttle standard deviation             } epilogue of the QP kernels
constexpr int QD_JLIM = 164;  // 1   1.0 when the joint-limit rows are on (optional extension, constraintsVSMPC.cpp:388-468)
constexpr int QD_JLO = 165;   // 8   jointPos_min - q_cmd   (lower bound of every dq block, :450-451)
constexpr int QD_JHI = 173;   // 8   jointPos_max - q_cmd   (:452-453)
constexpr int QD_XREF = 184;  // 12*NC  rows: pos(3) linMom(3) rpy(3) angMom(3); NC reference columns each

// ---- per-instance persistent state (SoA, row = scalar, column = instance) ------------------------
constexpr int ST_P_INIT = 0;    // 3  ReferenceTrackingCost::m_initialCoMPos
constexpr int ST_RPY_INIT = 3;  // 3  m_initialRPY (cost) == AngularMomentumDynamicVS::m_rpyInit
constexpr int ST_RPY_OLD = 6;   // 3  ConstraintInitialState::m_rpyOld
constexpr int ST_NTURNS = 9;    // 3  m_nTurns
constexpr int ST_QREF0 = 12;    // 8  JointPositionRegularizationCost::m_jointPosReference
constexpr int ST_QACC = 20;     // 8  VariableSamplingMPC::m_jointsPositionReference[controlled]
constexpr int ST_P_REF = 28;    // 3  QPInput::m_posCoMReference
constexpr int ST_RPY_REF = 31;  // 3  QPInput::m_RPYReference
constexpr int ST_MOM_REF = 34;  // 6  QPInput::m_momentumReference
constexpr int ST_ALPHA = 40;    // 1  QPInput::m_alphaGravity
constexpr int ST_WIN = 41;      // 12*NC reference windows (same row order as QD_XREF)
// int state
constexpr int SI_REF_COUNTER = 0; // ReferenceTrackingCost::m_counter
constexpr int SI_THR_COUNTER = 1; // ThrottleConstraint::m_counter
constexpr int SI_ALPHA_IDX = 2;   // LinearMomentumDynamicVS::m_trajectoryManager cursor
constexpr int SI_REF_IDX = 3;     // ReferenceTrackingCost::m_trajManager cursor
constexpr int SI_COUNT = 4;

// ---- factorisation workspace per instance (QP kernel) --------------------------------------------
// per stage: K (NU x NY), Hinv (NU x NU), Ptt (NZ)
constexpr int WS_K = 0;
constexpr int WS_HINV = NU * NY;           // 360
constexpr int WS_PTT = WS_HINV + NU * NU;  // 504
constexpr int WS_STAGE = 544;              // padded (>= 504 + 38)

struct DeviceConfig
{
    int N, Ns, Nc, NC; // nIter, nIterSmall, controlHorizon, NC = N - Ns + 1 reference columns
    int nblk;          // throttle blocks Nc - Ns + 1
    int n_var, n_con;
    int ratio;         // round(periodLarge / periodSmall)
    int use_jet_dynamic, use_estimated_thrust;
    int use_jl;        // joint-limit rows on (JointPositionConstraint, optional)
    int qd_stride, st_rows;
    int alpha_len, traj_len;
    double Qd[NX];
    double Rqd[NJ];
    double w_reg_q;
    double w_t, w_i;
    double throttle_min, throttle_max;
    double jc[13];
    double jn[4];
    double jl_min[NJ], jl_max[NJ];   // handle-wide joint limits [rad]
    double dt[MAX_ITER];
};

// knot kinds of the backward recursion
enum : int { KIND_T = 0, KIND_H = 1, KIND_M = 2, KIND_0 = 3 };

__host__ __device__ inline int joint_block(int k, int Nc) { return k < Nc ? k : Nc - 1; }
__host__ __device__ inline int throttle_block(int k, int Ns, int Nc)
{
    return k < Ns ? 0 : (k < Nc ? k - (Ns - 1) : Nc - Ns);
}
__host__ __device__ inline int knot_kind(int k, int Ns, int Nc)
{
    if (k == 0)
        return KIND_0;
    const bool new_j = joint_block(k, Nc) != joint_block(k - 1, Nc);
    const bool new_t = throttle_block(k, Ns, Nc) != throttle_block(k - 1, Ns, Nc);
    if (new_j && new_t)
        return KIND_M;
    if (new_j)
        return KIND_H;
    return KIND_T;
}
// column of the reference window used by knot k (costsVSMPC.cpp:191-200)
__host__ __device__ inline int ref_col(int k, int Ns) { return k < Ns ? 0 : k - Ns; }

// per-instance parameter rows (optional SoA buffer [IP_ROWS][B]; absent: the handle's vsmpc_config values)
constexpr int IP_JC = VSMPC_IP_JET_COEFF, IP_JN = VSMPC_IP_JET_NORM, IP_TMIN = VSMPC_IP_THROTTLE_MIN,
              IP_TMAX = VSMPC_IP_THROTTLE_MAX, IP_ROWS = VSMPC_INSTANCE_PARAM_DOUBLES;

// jet model, UT/src/JetModel.cpp:29-79
struct Jet
{
    const double* c;
    const double* n;
    __device__ double f(double T, double Td) const
    {
        return c[0] + c[1] * T + c[2] * Td + c[3] * T * Td + c[4] * T * T + c[5] * Td * Td;
    }
    __device__ double g(double T, double Td) const
    {
        return c[6] + c[7] * T + c[8] * Td + c[9] * T * Td + c[10] * T * T + c[11] * Td * Td;
    }
    __device__ double df_dT(double T, double Td) const { return c[1] + c[3] * Td + 2 * c[4] * T; }
    __device__ double df_dTd(double T, double Td) const { return c[2] + c[3] * T + 2 * c[5] * Td; }
    __device__ double dg_dT(double T, double Td) const { return c[7] + c[9] * Td + 2 * c[10] * T; }
    __device__ double dg_dTd(double T, double Td) const { return c[8] + c[9] * T + 2 * c[11] * Td; }
    __device__ double v(double u) const { return u + c[12] * u * u; }
    __device__ double stdT(double T) const { return (T - n[0]) / n[1]; }
    __device__ double stdTd(double Td) const { return Td / n[1]; }
    __device__ double stdU(double u) const { return (u - n[2]) / n[3]; }
    // JetModel::destandardizeThrottle_u2T, JetModel.cpp:93-109
    __device__ double destdU(double vv) const
    {
        double u = (-1.0 + sqrt(1.0 + 4.0 * c[12] * vv)) / (2.0 * c[12]);
        u = u * n[3] + n[2];
        if (u < 0.0)
            u = 0.0;
        else if (u > 100.0)
            u = 100.0;
        return u;
    }
};

// dynamic shared memory opt-in is a per-device function attribute: set it once per device a kernel is launched on
template <typename K> inline cudaError_t ensure_dynamic_smem(K kernel, int bytes, bool (&done)[64])
{
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess)
        return e;
    if (dev >= 0 && dev < 64 && done[dev])
        return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e == cudaSuccess && dev >= 0 && dev < 64)
        done[dev] = true;
    return e;
}

// JetModel::destandardizeThrottle_u2T with the instance's constants from the QP data block
__device__ inline double destd_throttle_qd(const double* __restrict__ cf, double vv)
{
    const double c12 = cf[QD_JC12];
    double u = (-1.0 + sqrt(1.0 + 4.0 * c12 * vv)) / (2.0 * c12);
    u = u * cf[QD_USTD] + cf[QD_UMEAN];
    return u < 0.0 ? 0.0 : (u > 100.0 ? 100.0 : u);
}

__device__ __forceinline__ void skew3(const double* v, double* S)
{ // UT/src/FlightControlUtils.cpp:77-85
    S[0] = 0.0; S[1] = -v[2]; S[2] = v[1];
    S[3] = v[2]; S[4] = 0.0; S[5] = -v[0];
    S[6] = -v[1]; S[7] = v[0]; S[8] = 0.0;
}


// dense (A, B_J, B_T, c) of one instance from its QP data block (SystemDynamicVS::get{A,BJoints,BThrottle}Matrix / getCVector)
__device__ inline void expand_dense(const double* __restrict__ q, double* A, double* BJ, double* BT, double* c)
{
    for (int e = 0; e < NX * NX; ++e)
        A[e] = 0.0;
    for (int e = 0; e < NX * NJ; ++e)
        BJ[e] = 0.0;
    for (int e = 0; e < NX * NT; ++e)
        BT[e] = 0.0;
    for (int e = 0; e < NX; ++e)
        c[e] = 0.0;
    double S[9];
    skew3(q + QD_OMEGA, S);
    for (int a = 0; a < 3; ++a)
    {
        for (int b = 0; b < 3; ++b)
        {
            A[(IX_COM + a) * NX + IX_LIN + b] = q[QD_RM + a * 3 + b];
            A[(IX_LIN + a) * NX + IX_LIN + b] = (S[a * 3 + b] == 0.0) ? 0.0 : -S[a * 3 + b];
            A[(IX_RPY + a) * NX + IX_ANG + b] = q[QD_WI + a * 3 + b];
            A[(IX_ANG + a) * NX + IX_ANG + b] = (S[a * 3 + b] == 0.0) ? 0.0 : -S[a * 3 + b];
        }
        for (int j = 0; j < NT; ++j)
        {
            A[(IX_LIN + a) * NX + IX_T + j] = q[QD_ALIN + a * NT + j];
            A[(IX_ANG + a) * NX + IX_T + j] = q[QD_AANG + a * NT + j];
        }
        for (int b = 0; b < NJ; ++b)
        {
            BJ[(IX_LIN + a) * NJ + b] = q[QD_LLIN + a * NJ + b];
            BJ[(IX_ANG + a) * NJ + b] = q[QD_LANG + a * NJ + b];
        }
        A[(IX_EP + a) * NX + IX_COM + a] = 1.0;
        A[(IX_ER + a) * NX + IX_RPY + a] = 1.0;
        c[IX_LIN + a] = q[QD_CL + a];
        c[IX_EP + a] = q[QD_CEP + a];
        c[IX_ER + a] = q[QD_CER + a];
    }
    for (int j = 0; j < NT; ++j)
    {
        A[(IX_T + j) * NX + IX_TD + j] = q[QD_JTT];
        A[(IX_TD + j) * NX + IX_T + j] = q[QD_JA + j];
        A[(IX_TD + j) * NX + IX_TD + j] = q[QD_JB + j];
        BT[(IX_TD + j) * NT + j] = q[QD_JG + j];
        BT[(IX_T + j) * NT + j] = q[QD_JGT];
        c[IX_TD + j] = q[QD_CTD + j];
    }
}


} // namespace vsmpc
