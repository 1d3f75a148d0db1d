// K1 — linearise + discretise-prep kernel (sm_100a, FP64 CUDA cores).
//
// One warp per MPC instance.  Reads the instance's column of the structure-of-arrays pack into shared
// memory, splits the work across the lanes and replaces, for one controller tick, the host work of
//   IMPCProblem::update                       MPC/src/IMPCProblem/IMPCProblem.cpp:150-194
//     ReferenceTrackingCost::compute...       MPC/src/variableSamplingMPC/costsVSMPC.cpp:121-181
//     ThrottleInitialValueCost::compute...    costsVSMPC.cpp:468-487
//     JointPositionRegularizationCost::...    costsVSMPC.cpp:558-592
//     SystemDynamicVS::updateDynamicMatrices  systemDynamicsVSMPC.cpp:495-507 (Angular :72-226,
//                                             Linear :282-350, Jet :384-461)
//     ConstraintInitialState::updateInitial.. constraintsVSMPC.cpp:206-247
//     ThrottleConstraint::compute...          constraintsVSMPC.cpp:338-374
// It emits the ~120 structural nonzeros of (A, B_J, B_T, c) instead of the reference's dense
// 512x588 constraint matrix; the per-knot dt scaling (constraintsVSMPC.cpp:76-131) is applied by
// the QP kernel on the fly.  The instance-major QP data block is staged in shared memory and written
// with coalesced stores.
#include "vsmpc_common.cuh"

namespace vsmpc
{

__device__ __forceinline__ void mat3T_vec(const double* R, const double* v, double* o)
{ // o = R^T v, R row-major
    o[0] = R[0] * v[0] + R[3] * v[1] + R[6] * v[2];
    o[1] = R[1] * v[0] + R[4] * v[1] + R[7] * v[2];
    o[2] = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
}

__device__ __forceinline__ void mat3_mul(const double* A, const double* B, double* C)
{
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j)
            C[i * 3 + j] = A[i * 3 + 0] * B[0 + j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

__device__ __forceinline__ void mat3_inv(const double* A, double* I)
{
    const double c00 = A[4] * A[8] - A[5] * A[7];
    const double c01 = A[5] * A[6] - A[3] * A[8];
    const double c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    const double id = 1.0 / det;
    I[0] = c00 * id;
    I[1] = (A[2] * A[7] - A[1] * A[8]) * id;
    I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c01 * id;
    I[4] = (A[0] * A[8] - A[2] * A[6]) * id;
    I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c02 * id;
    I[7] = (A[1] * A[6] - A[0] * A[7]) * id;
    I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}

// (X^T M_b X).block(3,3,3,3) with X = Ad(G_H_B) = [[R, S(r)R],[0, R]]
// (systemDynamicsVSMPC.cpp:110-130, costsVSMPC.cpp:268-285)
__device__ void locked_inertia(const double* __restrict__ pack, int B, int i, const double* R, double* I3)
{
    double r[3], Sr[9], SR[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        r[a] = pack[(VSMPC_PK_P_COM + a) * (size_t)B + i] - pack[(VSMPC_PK_BASE_POS + a) * (size_t)B + i];
    skew3(r, Sr);
    mat3_mul(Sr, R, SR);
    // Xc = [SR; R] (6x3): I3 = Xc^T M_b Xc
    double MX[18]; // M_b * Xc  (6x3)
#pragma unroll
    for (int a = 0; a < 6; ++a)
    {
        double m[6];
#pragma unroll
        for (int b = 0; b < 6; ++b)
            m[b] = pack[(VSMPC_PK_MB + a * 6 + b) * (size_t)B + i];
#pragma unroll
        for (int c = 0; c < 3; ++c)
        {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += m[b] * SR[b * 3 + c];
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += m[3 + b] * R[b * 3 + c];
            MX[a * 3 + c] = s;
        }
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
        {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += SR[b * 3 + a] * MX[b * 3 + c];
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += R[b * 3 + a] * MX[(3 + b) * 3 + c];
            I3[a * 3 + c] = s;
        }
}

// W(rpy): costsVSMPC.cpp:276-282
__device__ __forceinline__ void W_of_rpy(double s0, double c0, double s1, double c1, double* W)
{
    W[0] = 1.0; W[1] = 0.0; W[2] = -s1;
    W[3] = 0.0; W[4] = c0; W[5] = c1 * s0;
    W[6] = 0.0; W[7] = -s0; W[8] = c0 * c1;
}

constexpr int K1_WARPS = 8; // instances per CTA (one warp each): 8 adjacent SoA columns = 64 contiguous bytes per pack row

// doubles of shared memory per instance: pk[360] | out[qd_stride] | col[12] | ipar[20] | stc[st_rows] | sic[4 ints], padded
// to 2 (mod 16) so that the tile-transposed staging stores (8 instances x 2 rows per half-warp) hit 16 distinct bank pairs
__host__ __device__ inline int k1_per_warp(const DeviceConfig& cfg)
{
    const int n = 360 + cfg.qd_stride + 12 + 20 + ((cfg.st_rows + 4 + 3) & ~3);
    return ((n + 13) & ~15) + 2;
}

// (X^T M_b X).block(3,3,3,3) from the pack row staged in shared memory
__device__ __forceinline__ void locked_inertia_sm(const double* __restrict__ pk, const double* R, double* I3)
{
    double r[3], Sr[9], SR[9];
#pragma unroll
    for (int a = 0; a < 3; ++a)
        r[a] = pk[VSMPC_PK_P_COM + a] - pk[VSMPC_PK_BASE_POS + a];
    skew3(r, Sr);
    mat3_mul(Sr, R, SR);
    double MX[18];
#pragma unroll
    for (int a = 0; a < 6; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
        {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += pk[VSMPC_PK_MB + a * 6 + b] * SR[b * 3 + c];
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += pk[VSMPC_PK_MB + a * 6 + 3 + b] * R[b * 3 + c];
            MX[a * 3 + c] = s;
        }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int c = 0; c < 3; ++c)
        {
            double s = 0.0;
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += SR[b * 3 + a] * MX[b * 3 + c];
#pragma unroll
            for (int b = 0; b < 3; ++b)
                s += R[b * 3 + a] * MX[(3 + b) * 3 + c];
            I3[a * 3 + c] = s;
        }
}

// One warp per MPC instance, eight instances per CTA.  The CTA stages the [359][8] tile of the SoA pack (and the
// [st_rows][8] tile of the persistent state) cooperatively: consecutive threads read the 8 adjacent columns of a row, i.e.
// 64 contiguous bytes = two fully used 32-byte sectors per row (4 doubles per sector; a warp-per-column read touched one
// sector per double), transposed into one shared-memory column per instance.  Then the lanes of a warp split the work
// (lane = 8*jet + joint for the Lambda matrices, one lane per jet / per state entry elsewhere), and the instance-major
// QP data block is written back with fully coalesced stores.
// mode 0: update tick.  mode 1: configure (initialise persistent state, then run tick 0).
__global__ void __launch_bounds__(32 * K1_WARPS)
linearise_kernel(const __grid_constant__ DeviceConfig cfgv, int B, int mode,
                 const double* __restrict__ pack, const double* __restrict__ joint_pos_sel,
                 const int* __restrict__ phase0, double* __restrict__ st, int* __restrict__ si,
                 const double* __restrict__ alpha_traj, const double* __restrict__ traj_pos,
                 const double* __restrict__ traj_vel, const double* __restrict__ traj_rpy,
                 const double* __restrict__ traj_rpyd, double* __restrict__ qd, const double* __restrict__ ip,
                 int* __restrict__ fb_count, const double* __restrict__ jl)
{
    extern __shared__ double k1_smem[]; // per warp: pk[360] | out[qd_stride] | col[12] | ipar[20] | stc[st_rows] | sic[4]
    const DeviceConfig& cfg = cfgv;   // kernel parameter space (constant bank): no global round trip for the configuration
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * K1_WARPS + warp;
    const int NC = cfg.NC;
    const int per_warp = k1_per_warp(cfg);
    const size_t Bs = (size_t)B;
    if (fb_count && blockIdx.x == 0 && threadIdx.x == 0)
        *fb_count = 0;   // the list of instances for the fallback QP kernel is refilled by this tick's QP kernel
    {
        // ---- cooperative tile staging: thread -> (row f, column c) with c fastest; every global load of the tick is
        // issued before the first one is consumed (one DRAM round trip)
        constexpr int TPB = 32 * K1_WARPS;
        constexpr int NPK = (VSMPC_PACK_DOUBLES * K1_WARPS + TPB - 1) / TPB;   // 12
        constexpr int NSV = 6;                                                   // state rows staged in registers: 6 * 32
        const int c = threadIdx.x & (K1_WARPS - 1), r0 = threadIdx.x / K1_WARPS; // r0 < 32
        const int ic = blockIdx.x * K1_WARPS + c;
        const bool on = ic < B;
        const int st_stage = 360 + cfg.qd_stride + 12 + 20;
        double pv[NPK], sv[NSV];
        int siv = 0;
#pragma unroll
        for (int t = 0; t < NPK; ++t)
        {
            const int f = r0 + 32 * t;
            pv[t] = (on && f < VSMPC_PACK_DOUBLES) ? pack[(size_t)f * Bs + ic] : 0.0;
        }
        const double ipv = (ip && on && r0 < IP_ROWS) ? ip[(size_t)r0 * Bs + ic] : 0.0;
        if (mode == 0)
        {
#pragma unroll
            for (int t = 0; t < NSV; ++t)
            {
                const int f = r0 + 32 * t;
                sv[t] = (on && f < cfg.st_rows) ? st[(size_t)f * Bs + ic] : 0.0;
            }
            if (on && r0 < SI_COUNT)
                siv = si[(size_t)r0 * Bs + ic];
        }
        double* pkc = k1_smem + (size_t)c * per_warp;
#pragma unroll
        for (int t = 0; t < NPK; ++t)
        {
            const int f = r0 + 32 * t;
            if (f < VSMPC_PACK_DOUBLES)
                pkc[f] = pv[t];
        }
        if (ip && r0 < IP_ROWS)
            pkc[st_stage - 20 + r0] = ipv;
        if (mode == 0)
        {
#pragma unroll
            for (int t = 0; t < NSV; ++t)
            {
                const int f = r0 + 32 * t;
                if (f < cfg.st_rows)
                    pkc[st_stage + f] = sv[t];
            }
            for (int f = r0 + 32 * NSV; f < cfg.st_rows; f += 32) // horizons with more than 13 reference columns
                pkc[st_stage + f] = on ? st[(size_t)f * Bs + ic] : 0.0;
            if (r0 < SI_COUNT)
                reinterpret_cast<int*>(pkc + st_stage + cfg.st_rows)[r0] = siv;
        }
    }
    __syncthreads();
    if (i >= B)
        return;
    double* pk = k1_smem + (size_t)warp * per_warp;
    double* out = pk + 360;
    double* colbuf = out + cfg.qd_stride;
    double* ipar = colbuf + 12;   // jet coefficients (13), normalisation (4), throttle min / max of this instance
    double* stc = ipar + 20;      // staged copy of the persistent state column: every global load of the tick is issued up
                                  // front; writes go to the copy and to global memory
    int* sic = reinterpret_cast<int*>(stc + cfg.st_rows);
#define STR(f) stc[(f)]
#define STW(f, v)                         \
    do                                    \
    {                                     \
        const double v__ = (v);           \
        stc[(f)] = v__;                   \
        st[(size_t)(f) * Bs + i] = v__;   \
    } while (0)
#define SI(f) si[(size_t)(f) * Bs + i]
    // joint limits of this instance (optional rows): lanes 0-7 lower, 8-15 upper; consumed at the end of the tick
    double jl_v = 0.0;
    if (cfg.use_jl && lane < 2 * NJ)
        jl_v = jl ? jl[(size_t)lane * Bs + i] : (lane < NJ ? cfg.jl_min[lane] : cfg.jl_max[lane - NJ]);
    if (!ip && lane < IP_ROWS)
        ipar[lane] = lane < IP_JN ? cfg.jc[lane] : (lane < IP_TMIN ? cfg.jn[lane - IP_JN] : (lane == IP_TMIN ? cfg.throttle_min : cfg.throttle_max));
    __syncwarp();
    const Jet jet{ipar + IP_JC, ipar + IP_JN};
    double R[9], rpy[3], pcom[3];
#pragma unroll
    for (int a = 0; a < 9; ++a)
        R[a] = pk[VSMPC_PK_WRB + a];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        rpy[a] = pk[VSMPC_PK_RPY + a];
        pcom[a] = pk[VSMPC_PK_P_COM + a];
    }
    const double mass = pk[VSMPC_PK_MASS];
    double I3[9], W[9];
    locked_inertia_sm(pk, R, I3);
    // one sincos for the whole warp (lanes 0,1: roll; lanes 2,3: pitch), shared by W(rpy) and W^-1(rpy) below
    double s0, c0, s1, c1;
    {
        double sv, cv;
        sincos((lane & 3) < 2 ? rpy[0] : rpy[1], &sv, &cv);
        s0 = __shfl_sync(0xffffffffu, sv, 0);
        c0 = __shfl_sync(0xffffffffu, cv, 0);
        s1 = __shfl_sync(0xffffffffu, sv, 2);
        c1 = __shfl_sync(0xffffffffu, cv, 2);
    }
    W_of_rpy(s0, c0, s1, c1, W);

    // new reference-window column at trajectory index idx (costsVSMPC.cpp:105-112,132-149) -> colbuf
    auto ref_column = [&](int idx, const double* pinit, const double* rinit) {
        double vel[3], rd[3], mv[3], Wr[3], col[12];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            col[a] = pinit[a] + traj_pos[3 * idx + a];
            vel[a] = traj_vel[3 * idx + a];
            col[6 + a] = rinit[a] + traj_rpy[3 * idx + a];
            rd[a] = traj_rpyd[3 * idx + a];
            mv[a] = mass * vel[a];
        }
        mat3T_vec(R, mv, col + 3);
#pragma unroll
        for (int a = 0; a < 3; ++a)
            Wr[a] = W[a * 3] * rd[0] + W[a * 3 + 1] * rd[1] + W[a * 3 + 2] * rd[2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            col[9 + a] = I3[a * 3] * Wr[0] + I3[a * 3 + 1] * Wr[1] + I3[a * 3 + 2] * Wr[2];
        if (lane == 0)
        {
#pragma unroll
            for (int a = 0; a < 12; ++a)
                colbuf[a] = col[a];
        }
        __syncwarp();
    };

    int rc, tc, aidx, ridx;
    if (mode == 1)
    {
        // configureDynVectorsSize of the costs/constraints (costsVSMPC.cpp:74-119,
        // constraintsVSMPC.cpp:184-204,326-336; systemDynamicsVSMPC.cpp:67; variableSamplingMPC.cpp:60)
        if (lane < 3)
        {
            STW(ST_P_INIT + lane, pcom[lane]);
            STW(ST_RPY_INIT + lane, rpy[lane]);
            STW(ST_RPY_OLD + lane, rpy[lane]);
            STW(ST_NTURNS + lane, 0.0);
            STW(ST_P_REF + lane, 0.0);
            STW(ST_RPY_REF + lane, 0.0);
        }
        if (lane < 6)
            STW(ST_MOM_REF + lane, 0.0);
        if (lane == 0)
            STW(ST_ALPHA, 0.0);
        if (lane < NJ)
        {
            const double q0 = joint_pos_sel[(size_t)lane * Bs + i];
            STW(ST_QREF0 + lane, q0);
            STW(ST_QACC + lane, q0);
        }
        ref_column(0, pcom, rpy);
        for (int e = lane; e < 12 * NC; e += 32)
            STW(ST_WIN + e, colbuf[e / NC]);
        const int ph = phase0 ? phase0[i] : 0;
        // both 20-tick counters start at ratio-1 (costsVSMPC.cpp:118, constraintsVSMPC.cpp:335)
        rc = tc = (cfg.ratio - 1 + ph) % cfg.ratio;
        aidx = ridx = 0;
        __syncwarp();
    }
    else
    {
        rc = sic[SI_REF_COUNTER];
        tc = sic[SI_THR_COUNTER];
        aidx = sic[SI_ALPHA_IDX];
        ridx = sic[SI_REF_IDX];
    }
    double pinit[3], rinit[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        pinit[a] = (mode == 1) ? pcom[a] : STR(ST_P_INIT + a);
        rinit[a] = (mode == 1) ? rpy[a] : STR(ST_RPY_INIT + a);
    }
    // ---------------- costs (evaluated before the constraints, IMPCProblem.cpp:157-192) -------------
    const bool shift = rc == cfg.ratio - 1; // ReferenceTrackingCost, costsVSMPC.cpp:124-165
    if (shift)
    {
        if (ridx < cfg.traj_len - 1) // TrajectoryManager::advanceTrajectory, TrajectoryManager.cpp:142-153
            ridx++;
        ref_column(ridx, pinit, rinit);
        rc = 0;
    }
    else
        rc++;
    {
        // window entries of this lane: read (shifted) sources first, then write
        double wv[5];
#pragma unroll
        for (int t = 0; t < 5; ++t)
        {
            const int e = lane + 32 * t;
            wv[t] = 0.0;
            if (e < 12 * NC)
            {
                const int r = e / NC, cidx = e - r * NC;
                if (shift)
                    wv[t] = (cidx + 1 < NC) ? STR(ST_WIN + e + 1) : colbuf[r];
                else
                    wv[t] = STR(ST_WIN + e);
            }
        }
        for (int e = lane + 160; e < 12 * NC; e += 32) // horizons with more than 13 reference columns
        {
            const int r = e / NC, cidx = e - r * NC;
            const double v = shift ? ((cidx + 1 < NC) ? STR(ST_WIN + e + 1) : colbuf[r]) : STR(ST_WIN + e);
            out[QD_XREF + e] = v;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 5; ++t)
        {
            const int e = lane + 32 * t;
            if (e < 12 * NC)
            {
                out[QD_XREF + e] = wv[t];
                if (shift)
                    STW(ST_WIN + e, wv[t]);
            }
        }
        if (shift)
            for (int e = lane + 160; e < 12 * NC; e += 32)
                STW(ST_WIN + e, out[QD_XREF + e]);
        __syncwarp();
    }
    if (shift && lane < 3)
    { // publish the references into "QPInput" (costsVSMPC.cpp:155-160)
        STW(ST_P_REF + lane, out[QD_XREF + (0 + lane) * NC]);
        STW(ST_RPY_REF + lane, out[QD_XREF + (6 + lane) * NC]);
        STW(ST_MOM_REF + lane, out[QD_XREF + (3 + lane) * NC]);
        STW(ST_MOM_REF + 3 + lane, out[QD_XREF + (9 + lane) * NC]);
    }
    __syncwarp();
    double pref[3], rref[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        pref[a] = STR(ST_P_REF + a);
        rref[a] = STR(ST_RPY_REF + a);
    }
    if (lane < NT) // ThrottleInitialValueCost gradient, costsVSMPC.cpp:479-485
        out[QD_VBAR + lane] = jet.v(jet.stdU(pk[VSMPC_PK_THROTTLE_PREV + lane]));
    if (lane < NJ) // JointPositionRegularizationCost gradient, costsVSMPC.cpp:574-590
        out[QD_GQ + lane] = cfg.w_reg_q * (pk[VSMPC_PK_Q_CMD + lane] - STR(ST_QREF0 + lane));
    // JointPositionConstraint bounds (optional rows, constraintsVSMPC.cpp:450-453): limits - q_cmd, the same for every block
    if (lane < 2 * NJ)
        out[QD_JLO + lane] = cfg.use_jl ? jl_v - pk[VSMPC_PK_Q_CMD + (lane & (NJ - 1))] : 0.0;
    if (lane == 2 * NJ)
        out[QD_JLIM] = cfg.use_jl ? 1.0 : 0.0;
    if (lane > 2 * NJ && lane < 2 * NJ + 4)
        out[QD_JLIM + 2 * NJ + (lane - 2 * NJ)] = 0.0;   // padding up to QD_XREF
    // ---------------- dynamics ---------------------------------------------------------------------------
    const double t1 = s1 / c1;
    if (lane == 0)
    {
        double omw[3], omB[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
            omw[a] = pk[VSMPC_PK_OMEGA_WORLD + a];
        mat3T_vec(R, omw, omB);
#pragma unroll
        for (int a = 0; a < 3; ++a)
            out[QD_OMEGA + a] = omB[a];
        // A[rpy, angMom] = W^-1 * I^-1     (systemDynamicsVSMPC.cpp:86-87,140-147)
        double Wi[9] = {1.0, s0 * t1, c0 * t1, 0.0, c0, -s0, 0.0, s0 / c1, c0 / c1};
        double Ii[9], WI[9];
        mat3_inv(I3, Ii);
        mat3_mul(Wi, Ii, WI);
#pragma unroll
        for (int a = 0; a < 9; ++a)
            out[QD_WI + a] = WI[a];
        // throttle box / pin (constraintsVSMPC.cpp:338-374)
        out[QD_PINNED] = (tc != cfg.ratio - 1) ? 1.0 : 0.0;
        out[QD_VMIN] = jet.v(jet.stdU(ipar[IP_TMIN]));
        out[QD_VMAX] = jet.v(jet.stdU(ipar[IP_TMAX]));
        out[QD_JTT] = cfg.use_jet_dynamic ? 1.0 : 0.0;
        out[QD_JGT] = cfg.use_jet_dynamic ? 0.0 : 1.0;
        out[QD_JC12] = ipar[IP_JC + 12];
        out[QD_UMEAN] = ipar[IP_JN + 2];
        out[QD_USTD] = ipar[IP_JN + 3];
    }
    if (lane < 9)
        out[QD_RM + lane] = (1.0 / mass) * pk[VSMPC_PK_WRB + lane]; // 1/m * wRb (:296-297)
    if (lane < 12)
    {
        out[QD_ALIN + lane] = pk[VSMPC_PK_AMOM_BODY + lane];
        out[QD_AANG + lane] = pk[VSMPC_PK_AMOM_BODY + 12 + lane];
    }
    { // Lambda_lin / Lambda_ang ("unfiltered", systemDynamicsVSMPC.cpp:166-186,338-346): lane = 8*jet + joint
        const int j = lane >> 3, b = lane & 7;
        double ax[3], ar[3], ab[3], rb[3], Sa[9], Srb[9], SrSa[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            ax[a] = pk[VSMPC_PK_JET_AXES + j * 3 + a];
            ar[a] = pk[VSMPC_PK_JET_ARMS + j * 3 + a];
        }
        mat3T_vec(R, ax, ab);
        mat3T_vec(R, ar, rb);
        skew3(ab, Sa);
        skew3(rb, Srb);
        mat3_mul(Srb, Sa, SrSa);
        const double T = pk[VSMPC_PK_THRUST + j];
        double Jw[3], dl[3], Jc[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            Jw[a] = pk[VSMPC_PK_J_REL_ANG + (j * 3 + a) * NJ + b];
            dl[a] = pk[VSMPC_PK_J_JET_LIN + (j * 3 + a) * NJ + b] - pk[VSMPC_PK_J_COM + a * NJ + b];
        }
        mat3T_vec(R, dl, Jc); // getRelativeJacobianCoM, :208-226
        double lin[3], ang[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            const double saJc = Sa[a * 3] * Jc[0] + Sa[a * 3 + 1] * Jc[1] + Sa[a * 3 + 2] * Jc[2];
            const double ssJw = SrSa[a * 3] * Jw[0] + SrSa[a * 3 + 1] * Jw[1] + SrSa[a * 3 + 2] * Jw[2];
            const double saJw = Sa[a * 3] * Jw[0] + Sa[a * 3 + 1] * Jw[1] + Sa[a * 3 + 2] * Jw[2];
            ang[a] = -(T * saJc) - T * ssJw;
            lin[a] = -(T * saJw);
        }
        // sum over the jets in the reference's order ((j0 + j1) + j2) + j3
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            double l1 = __shfl_sync(0xffffffffu, lin[a], b + 8), l2 = __shfl_sync(0xffffffffu, lin[a], b + 16),
                   l3 = __shfl_sync(0xffffffffu, lin[a], b + 24);
            double a1 = __shfl_sync(0xffffffffu, ang[a], b + 8), a2 = __shfl_sync(0xffffffffu, ang[a], b + 16),
                   a3 = __shfl_sync(0xffffffffu, ang[a], b + 24);
            if (lane < NJ)
            {
                out[QD_LLIN + a * NJ + b] = ((lin[a] + l1) + l2) + l3;
                out[QD_LANG + a * NJ + b] = ((ang[a] + a1) + a2) + a3;
            }
        }
    }
    if (lane < 3)
    { // c[linMom] = alpha_g * m * wRb^T g (:307-311) ; c[posErr], c[rpyErr]
        const double alpha = alpha_traj[aidx];
        const double s = alpha * mass;
        const double g0 = pk[VSMPC_PK_GRAVITY], g1 = pk[VSMPC_PK_GRAVITY + 1], g2 = pk[VSMPC_PK_GRAVITY + 2];
        out[QD_CL + lane] = (s * pk[VSMPC_PK_WRB + lane]) * g0 + (s * pk[VSMPC_PK_WRB + 3 + lane]) * g1
                            + (s * pk[VSMPC_PK_WRB + 6 + lane]) * g2;
        out[QD_CEP + lane] = -pref[lane];              // :316
        out[QD_CER + lane] = -rinit[lane];             // :100 (configure-time RPY)
        if (lane == 0)
            STW(ST_ALPHA, alpha);
    }
    if (aidx < cfg.alpha_len - 1)
        aidx++;
    if (lane < NT)
    { // jets (:384-429)
        const int j = lane;
        const double thrust = pk[VSMPC_PK_THRUST + j], Tdest = pk[VSMPC_PK_THRUST_DOT_EST + j];
        const double Tdes = pk[VSMPC_PK_THRUST_DES + j], Tddes = pk[VSMPC_PK_THRUST_DOT_DES + j];
        if (cfg.use_jet_dynamic)
        {
            const double T = cfg.use_estimated_thrust ? thrust : Tdes;
            const double Td = cfg.use_estimated_thrust ? Tdest : Tddes;
            const double Tb = jet.stdT(T), Tdb = jet.stdTd(Td);
            const double vu = jet.v(jet.stdU(pk[VSMPC_PK_THROTTLE_PREV + j]));
            const double dh_dT = jet.df_dT(Tb, Tdb) + jet.dg_dT(Tb, Tdb) * vu;
            const double dh_dTd = jet.df_dTd(Tb, Tdb) + jet.dg_dTd(Tb, Tdb) * vu;
            out[QD_JA + j] = dh_dT;
            out[QD_JB + j] = dh_dTd;
            out[QD_JG + j] = jet.g(jet.stdT(Tdes), jet.stdTd(Tddes)) * ipar[IP_JN + 1];
            out[QD_CTD + j] = jet.f(Tb, Tdb) * ipar[IP_JN + 1] - dh_dT * T - dh_dTd * Td;
        }
        else
            out[QD_JA + j] = out[QD_JB + j] = out[QD_JG + j] = out[QD_CTD + j] = 0.0;
        out[QD_X0 + IX_T + j] = cfg.use_estimated_thrust ? thrust : Tdes;
        out[QD_X0 + IX_TD + j] = cfg.use_estimated_thrust ? Tdest : Tddes;
    }
    if (lane < 3)
    { // initial state (constraintsVSMPC.cpp:206-247) incl. RPY unwrapping
        const double PI = 3.14159265358979323846;
        const int a = lane;
        double nt = (mode == 1) ? 0.0 : STR(ST_NTURNS + a);
        const double old = (mode == 1) ? rpy[a] : STR(ST_RPY_OLD + a);
        const double cur = pk[VSMPC_PK_RPY + a];
        if (cur - old > PI)
            nt -= 1.0;
        else if (cur - old < -PI)
            nt += 1.0;
        STW(ST_NTURNS + a, nt);
        STW(ST_RPY_OLD + a, cur);
        const double unw = cur + 2 * PI * nt;
        const double pc = pk[VSMPC_PK_P_COM + a];
        out[QD_X0 + IX_COM + a] = pc;
        out[QD_X0 + IX_LIN + a] = pk[VSMPC_PK_MOMENTUM_BODY + a];
        out[QD_X0 + IX_RPY + a] = unw;
        out[QD_X0 + IX_ANG + a] = pk[VSMPC_PK_MOMENTUM_BODY + 3 + a];
        out[QD_X0 + IX_EP + a] = pc - pref[a];
        out[QD_X0 + IX_ER + a] = unw - rref[a];
    }
    tc = (tc == cfg.ratio - 1) ? 0 : tc + 1;
    if (lane == 0)
    {
        SI(SI_REF_COUNTER) = rc;
        SI(SI_THR_COUNTER) = tc;
        SI(SI_ALPHA_IDX) = aidx;
        SI(SI_REF_IDX) = ridx;
    }
#undef STR
#undef STW
#undef SI
    for (int e = QD_XREF + 12 * NC + lane; e < cfg.qd_stride; e += 32)
        out[e] = 0.0; // padding
    __syncwarp();
    // coalesced write-out of the instance-major block
    double* dst = qd + (size_t)i * cfg.qd_stride;
    for (int e = lane; e < cfg.qd_stride; e += 32)
        dst[e] = out[e];
}


// ---- dense expansion for parity tests (vsmpc_get_dynamics / vsmpc_get_qp_vectors) ----------------
__global__ void expand_dynamics_kernel(const DeviceConfig* __restrict__ cfgp, int B, const double* __restrict__ qd,
                                       double* A, double* BJ, double* BT, double* c)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B)
        return;
    expand_dense(qd + (size_t)i * cfgp->qd_stride, A + (size_t)i * NX * NX, BJ + (size_t)i * NX * NJ,
                 BT + (size_t)i * NX * NT, c + (size_t)i * NX);
}

// gradient and bounds in the reference's dense ordering (IMPCProblem.cpp:156-192)
__global__ void expand_qp_vectors_kernel(const DeviceConfig* __restrict__ cfgp, int B, const double* __restrict__ qd,
                                         double* q, double* l, double* u)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B)
        return;
    const DeviceConfig& cfg = *cfgp;
    const double* d = qd + (size_t)i * cfg.qd_stride;
    double* qi = q + (size_t)i * cfg.n_var;
    double* li = l + (size_t)i * cfg.n_con;
    double* ui = u + (size_t)i * cfg.n_con;
    const int N = cfg.N, NC = cfg.NC;
    for (int e = 0; e < cfg.n_var; ++e)
        qi[e] = 0.0;
    for (int e = 0; e < cfg.n_con; ++e)
        li[e] = ui[e] = 0.0;
    // q[x_k] = -Q xref_{k-1}   (costsVSMPC.cpp:175-178)
    for (int k = 1; k <= N; ++k)
    {
        const int col = ref_col(k - 1, cfg.Ns);
        for (int r = 0; r < 12; ++r)
        {
            const double v = -cfg.Qd[r] * d[QD_XREF + r * NC + col];
            qi[k * NX + r] = (v == 0.0) ? 0.0 : v;
        }
    }
    const int base = NX * (N + 1);
    for (int j = 0; j < cfg.Nc; ++j)
        for (int a = 0; a < NJ; ++a)
            qi[base + j * NJ + a] = d[QD_GQ + a];
    const int tb = base + cfg.Nc * NJ;
    for (int a = 0; a < NT; ++a)
        qi[tb + a] = -cfg.w_i * d[QD_VBAR + a];
    // bounds: dynamics rows -dt_k c ; x0 rows ; throttle rows
    double c[NX];
    for (int e = 0; e < NX; ++e)
        c[e] = 0.0;
    for (int a = 0; a < 3; ++a)
    {
        c[IX_LIN + a] = d[QD_CL + a];
        c[IX_EP + a] = d[QD_CEP + a];
        c[IX_ER + a] = d[QD_CER + a];
    }
    for (int j = 0; j < NT; ++j)
        c[IX_TD + j] = d[QD_CTD + j];
    for (int k = 0; k < N; ++k)
        for (int r = 0; r < NX; ++r)
        {
            const double v = -cfg.dt[k] * c[r];
            li[k * NX + r] = ui[k * NX + r] = v;
        }
    for (int r = 0; r < NX; ++r)
        li[N * NX + r] = ui[N * NX + r] = d[QD_X0 + r];
    const int t0 = N * NX + NX;
    for (int b = 0; b < cfg.nblk; ++b)
        for (int a = 0; a < NT; ++a)
        {
            if (b == 0 && d[QD_PINNED] != 0.0)
                li[t0 + a] = ui[t0 + a] = d[QD_VBAR + a];
            else
            {
                li[t0 + b * NT + a] = d[QD_VMIN];
                ui[t0 + b * NT + a] = d[QD_VMAX];
            }
        }
    if (cfg.use_jl)
    { // JointPositionConstraint rows (optional), after the NT (N - Ns + 1) throttle rows: blocks 0 .. Nc-1, the rest zero
        const int j0 = t0 + NT * (N - cfg.Ns + 1);
        for (int b = 0; b < cfg.Nc; ++b)
            for (int a = 0; a < NJ; ++a)
            {
                li[j0 + b * NJ + a] = d[QD_JLO + a];
                ui[j0 + b * NJ + a] = d[QD_JHI + a];
            }
    }
}

// IMPCProblem::getHessian (IMPCProblem.h:88) of ONE instance, dense n_var x n_var row-major: the constant cost
// blocks of costsVSMPC.cpp:78-93,166-173 (tracking), :375-409 (joint increments, throttle Laplacian), :472-478
// (initial throttle), :564-571 (joint position regularisation).  One thread per row.
__global__ void expand_hessian_kernel(const DeviceConfig* __restrict__ cfgp, double* __restrict__ P)
{
    const DeviceConfig& cfg = *cfgp;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cfg.n_var)
        return;
    double* row = P + (size_t)r * cfg.n_var;
    for (int e = 0; e < cfg.n_var; ++e)
        row[e] = 0.0;
    const int nxs = NX * (cfg.N + 1), njs = NJ * cfg.Nc;
    if (r < nxs)
    {
        if (r >= NX)
            row[r] = cfg.Qd[r % NX];
    }
    else if (r < nxs + njs)
        row[r] = cfg.Rqd[(r - nxs) % NJ];
    else
    {
        const int e = r - nxs - njs, blk = e / NT;
        row[r] = cfg.w_t * ((blk > 0 ? 1.0 : 0.0) + (blk < cfg.nblk - 1 ? 1.0 : 0.0)) + (blk == 0 ? cfg.w_i : 0.0);
        if (blk > 0)
            row[r - NT] = -cfg.w_t;
        if (blk < cfg.nblk - 1)
            row[r + NT] = -cfg.w_t;
    }
}

// IMPCProblem::getLinearConstraintMatrix (IMPCProblem.h:100) of ONE instance, dense n_con x n_var row-major, in the
// reference's row order (variableSamplingMPC.cpp:77-84): dynamics rows (constraintsVSMPC.cpp:76-128), initial-state rows
// (IQPUtilsMPC.cpp:71-92), throttle rows (constraintsVSMPC.cpp:338-350; the rows beyond the nblk blocks stay zero).
// One thread per row; A, BJ, BT: the dense expansions of this instance (expand_dense).
__global__ void expand_constraint_matrix_kernel(const DeviceConfig* __restrict__ cfgp, const double* __restrict__ A,
                                                const double* __restrict__ BJ, const double* __restrict__ BT,
                                                double* __restrict__ M)
{
    const DeviceConfig& cfg = *cfgp;
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= cfg.n_con)
        return;
    double* row = M + (size_t)r * cfg.n_var;
    for (int e = 0; e < cfg.n_var; ++e)
        row[e] = 0.0;
    const int N = cfg.N, nxs = NX * (N + 1), njs = NJ * cfg.Nc;
    if (r < NX * N)
    {
        const int k = r / NX, i = r - k * NX;
        const double dt = cfg.dt[k];
        for (int j = 0; j < NX; ++j)
            row[k * NX + j] = (i == j ? 1.0 : 0.0) + dt * A[i * NX + j];
        row[(k + 1) * NX + i] = -1.0;
        const int jb = joint_block(k, cfg.Nc), tb = throttle_block(k, cfg.Ns, cfg.Nc);
        for (int a = 0; a < NJ; ++a)
            row[nxs + jb * NJ + a] = dt * BJ[i * NJ + a];
        for (int a = 0; a < NT; ++a)
            row[nxs + njs + tb * NT + a] = dt * BT[i * NT + a];
    }
    else if (r < NX * N + NX)
        row[r - NX * N] = 1.0;
    else
    {
        const int e = r - NX * N - NX;
        if (e < NT * cfg.nblk)
            row[nxs + njs + e] = 1.0;
        const int ej = e - NT * (N - cfg.Ns + 1);       // joint-limit rows (optional): identity on dq blocks 0 .. Nc-1
        if (cfg.use_jl && ej >= 0 && ej < NJ * cfg.Nc)
            row[nxs + ej] = 1.0;
    }
}

// ---- host-side launchers ---------------------------------------------------------------------------
cudaError_t launch_linearise(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, int mode,
                             const double* pack, const double* joint_pos_sel, const int* phase0, double* st,
                             int* si, const double* alpha_traj, const double* traj_pos, const double* traj_vel,
                             const double* traj_rpy, const double* traj_rpyd, double* qd, const double* ip,
                             int* fb_count, const double* jl, cudaStream_t s)
{
    const size_t smem = (size_t)K1_WARPS * k1_per_warp(h_cfg) * sizeof(double);
    static bool attr_set[64] = {};
    if (smem > 200 * 1024)
        return cudaErrorInvalidValue;
    {
        const cudaError_t e = ensure_dynamic_smem(linearise_kernel, 200 * 1024, attr_set);
        if (e != cudaSuccess)
            return e;
    }
    const int grid = (B + K1_WARPS - 1) / K1_WARPS;
    linearise_kernel<<<grid, 32 * K1_WARPS, smem, s>>>(h_cfg, B, mode, pack, joint_pos_sel, phase0, st, si,
                                                       alpha_traj, traj_pos, traj_vel, traj_rpy, traj_rpyd, qd, ip, fb_count, jl);
    return cudaGetLastError();
}

cudaError_t launch_expand_dynamics(const DeviceConfig* d_cfg, int B, const double* qd, double* A, double* BJ,
                                   double* BT, double* c, cudaStream_t s)
{
    expand_dynamics_kernel<<<(B + 63) / 64, 64, 0, s>>>(d_cfg, B, qd, A, BJ, BT, c);
    return cudaGetLastError();
}

cudaError_t launch_expand_qp_vectors(const DeviceConfig* d_cfg, int B, const double* qd, double* q, double* l,
                                     double* u, cudaStream_t s)
{
    expand_qp_vectors_kernel<<<(B + 63) / 64, 64, 0, s>>>(d_cfg, B, qd, q, l, u);
    return cudaGetLastError();
}

cudaError_t launch_expand_hessian(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, double* P, cudaStream_t s)
{
    expand_hessian_kernel<<<(h_cfg.n_var + 63) / 64, 64, 0, s>>>(d_cfg, P);
    return cudaGetLastError();
}

// scratch: NX*NX + NX*NJ + NX*NT + NX doubles
cudaError_t launch_expand_constraint_matrix(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, const double* qd_inst,
                                            double* scratch, double* M, cudaStream_t s)
{
    double *A = scratch, *BJ = A + NX * NX, *BT = BJ + NX * NJ, *c = BT + NX * NT;
    expand_dynamics_kernel<<<1, 64, 0, s>>>(d_cfg, 1, qd_inst, A, BJ, BT, c);
    expand_constraint_matrix_kernel<<<(h_cfg.n_con + 63) / 64, 64, 0, s>>>(d_cfg, A, BJ, BT, M);
    return cudaGetLastError();
}

} // namespace vsmpc
