// K-1 — batched reduced kinematics: the part of Robot::setState the MPC path consumes (UT/src/Robot.cpp:212-278,325-332;
// SURVEY §8 f-2), for a kinematic tree given as arrays (include/vsmpc.h: vsmpc_kin_model).  Replaces, per instance, the
// iDynTree calls setRobotState / getFreeFloatingMassMatrix / getCentroidalTotalMomentum / getCenterOfMassPosition /
// getCenterOfMassJacobian / getFrameFreeFloatingJacobian / getRelativeJacobian / getWorldTransform and the jet loop of
// Robot.cpp:236-278 by one kernel that writes the kinematic rows of the pack in place on the device.
//
// One warp per instance, lane = link (<= 32 links); a CTA of eight warps handles eight consecutive instances so that the
// SoA robot state is read and the SoA pack written as [row][8 instances] tiles (64-byte segments) through shared memory.
//   1. forward kinematics level by level of the tree (lanes of one depth compose their parent's pose from shared memory);
//   2. total mass, CoM, centroidal momentum and the base block of the mass matrix by warp reductions over the links;
//   3. subtree mass / first moment per controlled joint -> CoM Jacobian columns (a_k x (f_sub - m_sub p_k) / m);
//   4. lane = 8 jet + joint: free-floating linear and relative angular Jacobian columns of the jet frames (a joint moves
//      a jet iff its link is an ancestor of the jet's link: bit masks prepared on the host); lanes 0-3: axes, arms, A_mom.
// Conventions: iDynTree MIXED representation (oracle/kinematics_oracle.py).  The QPInput rows of the state are passed through.
#include <vector>

#include "vsmpc_common.cuh"

namespace vsmpc
{

constexpr int KIN_WARPS = 8;
constexpr int KIN_L = VSMPC_KIN_MAX_LINKS;
constexpr int KIN_PASS = 28;                 // QPInput rows passed through: VSMPC_PK_THRUST .. end of the pack
constexpr int KW_R = 0, KW_P = 9 * KIN_L, KW_C = 12 * KIN_L, KW_A = 15 * KIN_L, KW_DOUBLES = 18 * KIN_L;   // per-warp scratch

struct KinModelDev
{
    vsmpc_kin_model m;
    int depth[KIN_L];             // depth of link l in the tree (base 0)
    unsigned anc[KIN_L];          // bit k set: the joint of link k moves link l (k ancestor of l or l itself, k >= 1)
    int link_of_sel[NJ];          // link whose joint is controlled joint a
    int max_depth;
    int ks_rows;                  // rows of the robot-state SoA
};

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(32 * KIN_WARPS)
kinematics_kernel(const KinModelDev* __restrict__ mdl, int B, const double* __restrict__ ks, double* __restrict__ pack,
                  double* __restrict__ jpos)
{
    extern __shared__ double kin_smem[];   // in [ks_rows][8] | out [PACK][8] | per warp KW_DOUBLES
    const KinModelDev& K = *mdl;
    const vsmpc_kin_model& M = K.m;
    const int n = M.n_links, nd = M.n_dof, rows = K.ks_rows;
    double* tin = kin_smem;
    double* tout = tin + rows * KIN_WARPS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double* ws = tout + VSMPC_PACK_DOUBLES * KIN_WARPS + warp * KW_DOUBLES;
    const int i0 = blockIdx.x * KIN_WARPS;
    const int inst = i0 + warp;
    const bool live = inst < B;
    // ---- stage the robot states of the eight instances: rows of 8 consecutive doubles -----------------------------------
    for (int e = threadIdx.x; e < rows * KIN_WARPS; e += 32 * KIN_WARPS)
    {
        const int r = e >> 3, c = e & 7;
        tin[e] = (i0 + c < B) ? ks[(size_t)r * B + i0 + c] : 0.0;
    }
    __syncthreads();
#define IN(r) tin[(r) * KIN_WARPS + warp]
#define OUT(r) tout[(r) * KIN_WARPS + warp]
    const int l = lane;
    const bool isl = l < n;
    // ---- 1. forward kinematics ---------------------------------------------------------------------------------------------
    double R[9], p[3];
    if (l == 0)
    {
#pragma unroll
        for (int e = 0; e < 9; ++e)
            R[e] = IN(VSMPC_KS_WRB + e);
#pragma unroll
        for (int e = 0; e < 3; ++e)
            p[e] = IN(VSMPC_KS_BASE_POS + e);
#pragma unroll
        for (int e = 0; e < 9; ++e)
            ws[KW_R + e] = R[e];
#pragma unroll
        for (int e = 0; e < 3; ++e)
            ws[KW_P + e] = p[e];
    }
    __syncwarp();
    const int par = isl && l > 0 ? M.parent[l] : 0;
    const int dep = isl ? K.depth[l] : -1;
    double ax[3] = {0.0, 0.0, 0.0};
    if (isl && l > 0)
    {
#pragma unroll
        for (int e = 0; e < 3; ++e)
            ax[e] = M.axis[l][e];
    }
    for (int d = 1; d <= K.max_depth; ++d)
    {
        if (dep == d)
        {
            // joint frame at q = 0 in the parent, then the rotation by q about the joint axis (Rodrigues)
            const double q = IN(VSMPC_KS_Q + M.dof[l]);
            double s, c;
            sincos(q, &s, &c);
            const double v = 1.0 - c;
            double Rq[9];
            Rq[0] = c + v * ax[0] * ax[0];
            Rq[1] = v * ax[0] * ax[1] - s * ax[2];
            Rq[2] = v * ax[0] * ax[2] + s * ax[1];
            Rq[3] = v * ax[1] * ax[0] + s * ax[2];
            Rq[4] = c + v * ax[1] * ax[1];
            Rq[5] = v * ax[1] * ax[2] - s * ax[0];
            Rq[6] = v * ax[2] * ax[0] - s * ax[1];
            Rq[7] = v * ax[2] * ax[1] + s * ax[0];
            Rq[8] = c + v * ax[2] * ax[2];
            double Rj[9];     // R0 Rq
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    Rj[3 * a + b] = M.R0[l][3 * a] * Rq[b] + M.R0[l][3 * a + 1] * Rq[3 + b] + M.R0[l][3 * a + 2] * Rq[6 + b];
            const double* Rp = ws + KW_R + 9 * par;
            const double* pp = ws + KW_P + 3 * par;
#pragma unroll
            for (int a = 0; a < 3; ++a)
            {
#pragma unroll
                for (int b = 0; b < 3; ++b)
                    R[3 * a + b] = Rp[3 * a] * Rj[b] + Rp[3 * a + 1] * Rj[3 + b] + Rp[3 * a + 2] * Rj[6 + b];
                p[a] = pp[a] + Rp[3 * a] * M.p0[l][0] + Rp[3 * a + 1] * M.p0[l][1] + Rp[3 * a + 2] * M.p0[l][2];
            }
#pragma unroll
            for (int e = 0; e < 9; ++e)
                ws[KW_R + 9 * l + e] = R[e];
#pragma unroll
            for (int e = 0; e < 3; ++e)
                ws[KW_P + 3 * l + e] = p[e];
        }
        __syncwarp();
    }
    // ---- 2. link CoM, world axis, world inertia; totals --------------------------------------------------------------------
    double cw[3] = {0, 0, 0}, aw[3] = {0, 0, 0}, Iw[9], ml = 0.0;
#pragma unroll
    for (int e = 0; e < 9; ++e)
        Iw[e] = 0.0;
    if (isl)
    {
        ml = M.mass[l];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            cw[a] = p[a] + R[3 * a] * M.com[l][0] + R[3 * a + 1] * M.com[l][1] + R[3 * a + 2] * M.com[l][2];
            aw[a] = R[3 * a] * ax[0] + R[3 * a + 1] * ax[1] + R[3 * a + 2] * ax[2];
        }
        double RI[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                RI[3 * a + b] = R[3 * a] * M.inertia[l][b] + R[3 * a + 1] * M.inertia[l][3 + b] + R[3 * a + 2] * M.inertia[l][6 + b];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                Iw[3 * a + b] = RI[3 * a] * R[3 * b] + RI[3 * a + 1] * R[3 * b + 1] + RI[3 * a + 2] * R[3 * b + 2];
#pragma unroll
        for (int e = 0; e < 3; ++e)
        {
            ws[KW_C + 3 * l + e] = cw[e];
            ws[KW_A + 3 * l + e] = aw[e];
        }
    }
    __syncwarp();
    const double Mtot = warp_sum(ml);
    double pc[3], p0w[3], v0[3], w0[3];
#pragma unroll
    for (int e = 0; e < 3; ++e)
    {
        pc[e] = warp_sum(ml * cw[e]) / Mtot;
        p0w[e] = IN(VSMPC_KS_BASE_POS + e);
        v0[e] = IN(VSMPC_KS_BASE_LIN_VEL + e);
        w0[e] = IN(VSMPC_KS_OMEGA_WORLD + e);
    }
    // link velocity: base twist + the joints that move the link
    double wl[3] = {w0[0], w0[1], w0[2]}, vc[3];
    {
        double r0[3] = {cw[0] - p0w[0], cw[1] - p0w[1], cw[2] - p0w[2]}, t[3];
        cross3(w0, r0, t);
        vc[0] = v0[0] + t[0];
        vc[1] = v0[1] + t[1];
        vc[2] = v0[2] + t[2];
        const unsigned mask = isl ? K.anc[l] : 0u;
        for (int k = 1; k < n; ++k)
            if (mask >> k & 1u)
            {
                const double qd = IN(VSMPC_KS_Q + nd + M.dof[k]);
                const double* ak = ws + KW_A + 3 * k;
                const double* pk = ws + KW_P + 3 * k;
                const double rk[3] = {cw[0] - pk[0], cw[1] - pk[1], cw[2] - pk[2]};
                cross3(ak, rk, t);
#pragma unroll
                for (int e = 0; e < 3; ++e)
                {
                    vc[e] = fma(qd, t[e], vc[e]);
                    wl[e] = fma(qd, ak[e], wl[e]);
                }
            }
    }
    double hl[3], ha[3];
    {
        const double rc[3] = {cw[0] - pc[0], cw[1] - pc[1], cw[2] - pc[2]};
        double t[3];
        cross3(rc, vc, t);
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            hl[a] = warp_sum(ml * vc[a]);
            ha[a] = warp_sum(Iw[3 * a] * wl[0] + Iw[3 * a + 1] * wl[1] + Iw[3 * a + 2] * wl[2] + ml * t[a]);
        }
    }
    // total inertia about the base origin, world axes: sum I_w + m (|r|^2 1 - r r')
    double Io[9];
    {
        const double r[3] = {cw[0] - p0w[0], cw[1] - p0w[1], cw[2] - p0w[2]};
        const double rr = r[0] * r[0] + r[1] * r[1] + r[2] * r[2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = a; b < 3; ++b)
            {
                const double v = warp_sum(Iw[3 * a + b] + ml * ((a == b ? rr : 0.0) - r[a] * r[b]));
                Io[3 * a + b] = v;
                Io[3 * b + a] = v;
            }
    }
    const double* Rb = ws + KW_R;     // base rotation
    // ---- outputs that do not depend on the lane role -------------------------------------------------------------------------
    if (lane == 0)
    {
#pragma unroll
        for (int e = 0; e < 9; ++e)
            OUT(VSMPC_PK_WRB + e) = Rb[e];
#pragma unroll
        for (int e = 0; e < 3; ++e)
        {
            OUT(VSMPC_PK_OMEGA_WORLD + e) = w0[e];
            OUT(VSMPC_PK_GRAVITY + e) = M.gravity[e];
            OUT(VSMPC_PK_BASE_POS + e) = p0w[e];
            OUT(VSMPC_PK_P_COM + e) = pc[e];
        }
        // iDynTree::Rotation::asRPY
        double rpy[3];
        if (Rb[6] < 1.0)
        {
            if (Rb[6] > -1.0)
            {
                rpy[0] = atan2(Rb[7], Rb[8]);
                rpy[1] = asin(-Rb[6]);
                rpy[2] = atan2(Rb[3], Rb[0]);
            }
            else
            {
                rpy[0] = 0.0;
                rpy[1] = M_PI / 2.0;
                rpy[2] = -atan2(-Rb[5], Rb[4]);
            }
        }
        else
        {
            rpy[0] = 0.0;
            rpy[1] = -M_PI / 2.0;
            rpy[2] = atan2(-Rb[5], Rb[4]);
        }
#pragma unroll
        for (int e = 0; e < 3; ++e)
            OUT(VSMPC_PK_RPY + e) = rpy[e];
        OUT(VSMPC_PK_MASS) = (double)(float)Mtot;       // Robot::m_totalMass is a float (Robot.h:338)
        // base block of the mass matrix, mixed representation
        const double c[3] = {pc[0] - p0w[0], pc[1] - p0w[1], pc[2] - p0w[2]};
        const double S[9] = {0.0, -c[2], c[1], c[2], 0.0, -c[0], -c[1], c[0], 0.0};
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
            {
                OUT(VSMPC_PK_MB + 6 * a + b) = a == b ? Mtot : 0.0;
                OUT(VSMPC_PK_MB + 6 * a + 3 + b) = -Mtot * S[3 * a + b];
                OUT(VSMPC_PK_MB + 6 * (3 + a) + b) = Mtot * S[3 * a + b];
                OUT(VSMPC_PK_MB + 6 * (3 + a) + 3 + b) = Io[3 * a + b];
            }
        // Robot::getMomentum(true): both halves rotated into base axes (Robot.cpp:325-328)
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            OUT(VSMPC_PK_MOMENTUM_BODY + a) = Rb[a] * hl[0] + Rb[3 + a] * hl[1] + Rb[6 + a] * hl[2];
            OUT(VSMPC_PK_MOMENTUM_BODY + 3 + a) = Rb[a] * ha[0] + Rb[3 + a] * ha[1] + Rb[6 + a] * ha[2];
        }
    }
    if (lane < KIN_PASS)
        OUT(VSMPC_PK_THRUST + lane) = IN(VSMPC_KS_Q + 2 * nd + lane);
    // ---- 3. CoM Jacobian columns of the controlled joints: lanes 0-7 ------------------------------------------------------------
    if (lane < NJ)
    {
        const int k = K.link_of_sel[lane];
        double msub = 0.0, f[3] = {0, 0, 0};
        for (int j = 1; j < n; ++j)
            if (K.anc[j] >> k & 1u)
            {
                const double mj = M.mass[j];
                msub += mj;
#pragma unroll
                for (int e = 0; e < 3; ++e)
                    f[e] = fma(mj, ws[KW_C + 3 * j + e], f[e]);
            }
        const double* ak = ws + KW_A + 3 * k;
        const double* pk = ws + KW_P + 3 * k;
        const double r[3] = {(f[0] - msub * pk[0]) / Mtot, (f[1] - msub * pk[1]) / Mtot, (f[2] - msub * pk[2]) / Mtot};
        double t[3];
        cross3(ak, r, t);
#pragma unroll
        for (int e = 0; e < 3; ++e)
            OUT(VSMPC_PK_J_COM + e * NJ + lane) = t[e];
        if (jpos && live)
            jpos[(size_t)lane * B + inst] = IN(VSMPC_KS_Q + M.sel[lane]);     // joint_pos_sel of configure
    }
    // ---- 4. jets: lane = 8 jet + joint --------------------------------------------------------------------------------------
    {
        const int jet = lane >> 3, a = lane & 7;
        const int lj = M.jet_link[jet];
        const double* Rl = ws + KW_R + 9 * lj;
        const double* pl = ws + KW_P + 3 * lj;
        double pj[3], axw[3];
#pragma unroll
        for (int e = 0; e < 3; ++e)
        {
            pj[e] = pl[e] + Rl[3 * e] * M.jet_pos[jet][0] + Rl[3 * e + 1] * M.jet_pos[jet][1] + Rl[3 * e + 2] * M.jet_pos[jet][2];
            axw[e] = Rl[3 * e] * M.jet_axis[jet][0] + Rl[3 * e + 1] * M.jet_axis[jet][1] + Rl[3 * e + 2] * M.jet_axis[jet][2];
        }
        const int k = K.link_of_sel[a];
        const bool on = K.anc[lj] >> k & 1u;
        const double* ak = ws + KW_A + 3 * k;
        const double* pk = ws + KW_P + 3 * k;
        const double r[3] = {pj[0] - pk[0], pj[1] - pk[1], pj[2] - pk[2]};
        double t[3];
        cross3(ak, r, t);
#pragma unroll
        for (int e = 0; e < 3; ++e)
        {
            OUT(VSMPC_PK_J_JET_LIN + (jet * 3 + e) * NJ + a) = on ? t[e] : 0.0;
            // relative Jacobian, MIXED: angular rows in base axes
            OUT(VSMPC_PK_J_REL_ANG + (jet * 3 + e) * NJ + a) = on ? Rb[e] * ak[0] + Rb[3 + e] * ak[1] + Rb[6 + e] * ak[2] : 0.0;
        }
        if (a == 0)
        {
            // arms from the CoM with Robot::m_deltaCoM (Robot.cpp:253-259), A_mom and its base-axes form (:261-265, :329-330)
            double arm[3], am[3];
#pragma unroll
            for (int e = 0; e < 3; ++e)
                arm[e] = pj[e] - (pc[e] + Rb[3 * e] * M.delta_com[0] + Rb[3 * e + 1] * M.delta_com[1] + Rb[3 * e + 2] * M.delta_com[2]);
            cross3(arm, axw, am);
#pragma unroll
            for (int e = 0; e < 3; ++e)
            {
                OUT(VSMPC_PK_JET_AXES + jet * 3 + e) = axw[e];
                OUT(VSMPC_PK_JET_ARMS + jet * 3 + e) = arm[e];
                OUT(VSMPC_PK_AMOM_BODY + e * NT + jet) = Rb[e] * axw[0] + Rb[3 + e] * axw[1] + Rb[6 + e] * axw[2];
                OUT(VSMPC_PK_AMOM_BODY + (3 + e) * NT + jet) = Rb[e] * am[0] + Rb[3 + e] * am[1] + Rb[6 + e] * am[2];
            }
        }
    }
#undef IN
#undef OUT
    __syncthreads();
    // ---- the pack rows of the eight instances, 64-byte segments ---------------------------------------------------------------
    for (int e = threadIdx.x; e < VSMPC_PACK_DOUBLES * KIN_WARPS; e += 32 * KIN_WARPS)
    {
        const int r = e >> 3, c = e & 7;
        if (i0 + c < B)
            pack[(size_t)r * B + i0 + c] = tout[e];
    }
}

// host: validate the tree, derive depths / ancestor masks; returns an empty string or the reason it was rejected
const char* kin_prepare(const vsmpc_kin_model& m, KinModelDev& K)
{
    if (m.n_links < 1 || m.n_links > KIN_L || m.n_dof < 1 || m.n_dof > 64)
        return "n_links must be in [1, 32] and n_dof in [1, 64]";
    K.m = m;
    K.max_depth = 0;
    std::vector<int> link_of_dof(m.n_dof, -1);
    for (int l = 0; l < m.n_links; ++l)
    {
        if (l == 0)
        {
            K.depth[0] = 0;
            K.anc[0] = 0u;
            continue;
        }
        if (m.parent[l] < 0 || m.parent[l] >= l)
            return "parent[l] must be in [0, l): links in topological order";
        if (m.dof[l] < 0 || m.dof[l] >= m.n_dof || link_of_dof[m.dof[l]] >= 0)
            return "dof[l] must be a distinct position of the joint vector";
        link_of_dof[m.dof[l]] = l;
        K.depth[l] = K.depth[m.parent[l]] + 1;
        K.anc[l] = K.anc[m.parent[l]] | (1u << l);
        K.max_depth = K.depth[l] > K.max_depth ? K.depth[l] : K.max_depth;
        double a2 = 0.0;
        for (int e = 0; e < 3; ++e)
            a2 += m.axis[l][e] * m.axis[l][e];
        if (!(a2 > 0.999999 && a2 < 1.000001))
            return "joint axes must be unit vectors";
        if (!(m.mass[l] >= 0.0))
            return "link masses must be non-negative";
    }
    for (int l = m.n_links; l < KIN_L; ++l)
    {
        K.depth[l] = -1;
        K.anc[l] = 0u;
    }
    double mt = 0.0;
    for (int l = 0; l < m.n_links; ++l)
        mt += m.mass[l];
    if (!(mt > 0.0))
        return "total mass must be positive";
    for (int a = 0; a < NJ; ++a)
    {
        if (m.sel[a] < 0 || m.sel[a] >= m.n_dof || link_of_dof[m.sel[a]] < 0)
            return "sel[a] must name a joint of the tree";
        K.link_of_sel[a] = link_of_dof[m.sel[a]];
    }
    for (int i = 0; i < NT; ++i)
        if (m.jet_link[i] < 0 || m.jet_link[i] >= m.n_links)
            return "jet_link out of range";
    K.ks_rows = VSMPC_KS_Q + 2 * m.n_dof + KIN_PASS;
    return "";
}

size_t kin_model_bytes() { return sizeof(KinModelDev); }
int kin_state_rows(const KinModelDev& K) { return K.ks_rows; }

cudaError_t launch_kinematics(const KinModelDev* d_model, int ks_rows, int B, const double* ks, double* pack, double* jpos,
                              cudaStream_t s)
{
    const size_t smem = ((size_t)(ks_rows + VSMPC_PACK_DOUBLES) * KIN_WARPS + (size_t)KIN_WARPS * KW_DOUBLES) * sizeof(double);
    static bool attr_set[64] = {};
    cudaError_t e = ensure_dynamic_smem(kinematics_kernel, (int)(smem > 96 * 1024 ? smem : 96 * 1024), attr_set);
    if (e != cudaSuccess)
        return e;
    kinematics_kernel<<<(B + KIN_WARPS - 1) / KIN_WARPS, 32 * KIN_WARPS, smem, s>>>(d_model, B, ks, pack, jpos);
    return cudaGetLastError();
}

} // namespace vsmpc
