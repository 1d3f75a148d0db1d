// K2 (generic dense variant, config.solver = 1) — structured QP solve, one warp per MPC instance.
//
// Replaces IMPCProblem::solve -> OsqpEigen (MPC/src/IMPCProblem/IMPCProblem.cpp:196-298) and
// VariableSamplingMPC::solveMPC's output extraction (variableSamplingMPC.cpp:88-112).
//
// Algorithm (tools/riccati_model.py is the executable specification):
//   * backward Riccati recursion over the N knots on the augmented variable (x, v, dq) where v is
//     the throttle block in effect and dq the joint block in effect — move-blocked inputs are carried
//     as parameters of the value function and eliminated at the knot where they are introduced;
//     the throttle-rate Laplacian couples a new block to the previous one (kinds 0/H/M/T);
//   * Goldfarb–Idnani dual active set on the throttle boxes in the space of the throttle
//     variables only; every column of the reduced inverse Hessian is one homogeneous Riccati
//     back-solve, computed lazily; a last back-solve with the multipliers recovers the full primal.
// This variant treats the model matrices as dense (no use of the structural sparsity); it is the
// on-device cross-check of the structured kernel (vsmpc_qp_structured.cu) and handles any horizon.
#include "vsmpc_common.cuh"

namespace vsmpc
{

constexpr int LDP = NZ + 1;       // 39: odd leading dimension, conflict-free row/column access
constexpr int GEN_WARPS = 2;      // instances per CTA
constexpr int GEN_MAXV = 192;     // most throttle variables one instance may have (working set / column slots)

struct GenSmem
{
    double P[NZ * LDP];
    double W[NZ * LDP];
    double F[NX * LDP];
    double A[NX * NX];
    double BT[NX * NT];
    double BJ[NX * NJ];
    double c[NX];
    double H[NU * NU];      // H_uu / its inverse
    double Huy[NU * NY];
    double K[NU * NY];
    double p[NZ], s[NZ], phi[NZ], hu[NU], KtH[NY];
    double x[NX], xn[NX], y[NY], u[NU], vcur[NT];
    double P0vx[NT * NX];
    double M0inv[NT * NT];
    double p0v[NT];
    // active set
    double* GW;             // inverse of the signed working-set block of G, [nv][nv] (global scratch: grows with the horizon)
    double r[GEN_MAXV], lam[GEN_MAXV], sgn[GEN_MAXV];
    double gw[GEN_MAXV];    // S G[W, p] of the candidate / row copies during the down-date
    int W_idx[GEN_MAXV];
    int slotW[GEN_MAXV];    // column slot of every working-set member
    int slot_of[GEN_MAXV];  // variable -> column slot (-1: not computed yet)
    int wpos[GEN_MAXV];     // variable -> position in the working set (-1: inactive)
    int gi1;
    double gv1;
};

__device__ __forceinline__ double warp_max_arg(double v, int idx, int* arg)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        if (ov > v || (ov == v && oi < idx))
        {
            v = ov;
            idx = oi;
        }
    }
    *arg = idx;
    return v;
}

// in-place Gauss-Jordan inverse of an SPD n x n matrix (row-major, ld = n) by one warp
__device__ bool spd_inverse(double* H, int n, int lane)
{
    bool ok = true;
    for (int p = 0; p < n; ++p)
    {
        __syncwarp();
        const double d = H[p * n + p];
        if (!(d > 0.0) || !isfinite(d))
            ok = false;
        const double dinv = 1.0 / d;
        for (int e = lane; e < n * n; e += 32)
        {
            const int i = e / n, j = e - i * n;
            if (i != p && j != p)
                H[e] -= H[i * n + p] * H[p * n + j] * dinv;
        }
        __syncwarp();
        for (int e = lane; e < n; e += 32)
            if (e != p)
            {
                H[p * n + e] *= dinv;
                H[e * n + p] *= -dinv;
            }
        if (lane == 0)
            H[p * n + p] = dinv;
    }
    __syncwarp();
    return ok;
}

struct GenCtx
{
    const DeviceConfig& cfg;
    GenSmem& sm;
    const double* qd;
    double* ws;     // [N][WS_STAGE]
    double* gcols;  // [nv][nv]
    double* kff;    // [N][NU] feed-forward terms of the current solve (global scratch)
    double* zout;   // [n_var] full primal
    int lane;
};

__device__ void build_F(GenCtx& c, double dt)
{
    GenSmem& sm = c.sm;
    for (int e = c.lane; e < NX * NZ; e += 32)
    {
        const int r = e / NZ, j = e - r * NZ;
        double v;
        if (j < NX)
            v = (r == j ? 1.0 : 0.0) + dt * sm.A[r * NX + j];
        else if (j < NY)
            v = dt * sm.BT[r * NT + (j - NX)];
        else
            v = dt * sm.BJ[r * NJ + (j - NY)];
        sm.F[r * LDP + j] = v;
    }
    __syncwarp();
}

// matrix part of the backward recursion; returns false on a numerical failure
__device__ bool gen_factor(GenCtx& c)
{
    const DeviceConfig& cfg = c.cfg;
    GenSmem& sm = c.sm;
    const int lane = c.lane;
    const int N = cfg.N;
    bool ok = true;
    for (int e = lane; e < NZ * LDP; e += 32)
        sm.P[e] = 0.0;
    __syncwarp();
    for (int k = N - 1; k >= 0; --k)
    {
        const double dt = cfg.dt[k];
        double* wsk = c.ws + (size_t)k * WS_STAGE;
        build_F(c, dt);
        for (int e = lane; e < NX; e += 32)
            sm.P[e * LDP + e] += cfg.Qd[e];
        __syncwarp();
        // Ptt = Ptilde * t, t = dt*[c;0;0]
        for (int i = lane; i < NZ; i += 32)
        {
            double a = 0.0;
            for (int r = 0; r < NX; ++r)
                a += sm.P[i * LDP + r] * (dt * sm.c[r]);
            wsk[WS_PTT + i] = a;
        }
        // W = Ptilde * T
        for (int e = lane; e < NZ * NZ; e += 32)
        {
            const int i = e / NZ, j = e - i * NZ;
            double a = (j >= NX) ? sm.P[i * LDP + j] : 0.0;
            for (int r = 0; r < NX; ++r)
                a += sm.P[i * LDP + r] * sm.F[r * LDP + j];
            sm.W[i * LDP + j] = a;
        }
        __syncwarp();
        // Phi = T^T W  -> P
        for (int e = lane; e < NZ * NZ; e += 32)
        {
            const int i = e / NZ, j = e - i * NZ;
            double a = (i >= NX) ? sm.W[i * LDP + j] : 0.0;
            for (int r = 0; r < NX; ++r)
                a += sm.F[r * LDP + i] * sm.W[r * LDP + j];
            sm.P[i * LDP + j] = a;
        }
        __syncwarp();
        const int kind = knot_kind(k, cfg.Ns, cfg.Nc);
        if (kind == KIND_T)
            continue;
        const int nu = (kind == KIND_M) ? NU : NJ;
        const int u0 = NZ - nu; // first eliminated index (26 for M, 30 for H/0)
        // H_uu and H_uy
        for (int e = lane; e < nu * nu; e += 32)
        {
            const int a = e / nu, b = e - a * nu;
            double v = sm.P[(u0 + a) * LDP + u0 + b];
            if (a == b)
            {
                const int g = u0 + a; // global index in (x,v,dq)
                v += (g < NY) ? cfg.w_t : cfg.Rqd[g - NY];
            }
            sm.H[e] = v;
        }
        for (int e = lane; e < nu * NY; e += 32)
        {
            const int a = e / NY, j = e - a * NY;
            double v;
            if (kind == KIND_M)
                v = (j < NX) ? sm.P[(u0 + a) * LDP + j] : ((a == j - NX) ? -cfg.w_t : 0.0);
            else
                v = sm.P[(u0 + a) * LDP + j];
            sm.Huy[e] = v;
        }
        __syncwarp();
        ok = spd_inverse(sm.H, nu, lane) && ok;
        // K = Hinv * Huy
        for (int e = lane; e < nu * NY; e += 32)
        {
            const int a = e / NY, j = e - a * NY;
            double v = 0.0;
            for (int m = 0; m < nu; ++m)
                v += sm.H[a * nu + m] * sm.Huy[m * NY + j];
            sm.K[e] = v;
        }
        __syncwarp();
        // P_yy <- base - Huy^T K ; clear the eliminated rows/cols
        for (int e = lane; e < NZ * NZ; e += 32)
        {
            const int i = e / NZ, j = e - i * NZ;
            double v = 0.0;
            if (i < NY && j < NY)
            {
                if (kind == KIND_M)
                    v = (i < NX && j < NX) ? sm.P[i * LDP + j] : ((i >= NX && i == j) ? cfg.w_t : 0.0);
                else
                    v = sm.P[i * LDP + j];
                for (int m = 0; m < nu; ++m)
                    v -= sm.Huy[m * NY + i] * sm.K[m * NY + j];
            }
            sm.W[i * LDP + j] = v;
        }
        __syncwarp();
        for (int e = lane; e < NZ * NZ; e += 32)
        {
            const int i = e / NZ, j = e - i * NZ;
            sm.P[i * LDP + j] = sm.W[i * LDP + j];
        }
        for (int e = lane; e < nu * NY; e += 32)
            wsk[WS_K + e] = sm.K[e];
        for (int e = lane; e < nu * nu; e += 32)
            wsk[WS_HINV + e] = sm.H[e];
        __syncwarp();
    }
    // V_0(x0, v0): keep P0[v, x] and (P0[v,v] + w_i I)^-1
    for (int e = lane; e < NT * NX; e += 32)
    {
        const int a = e / NX, j = e - a * NX;
        sm.P0vx[e] = sm.P[(NX + a) * LDP + j];
    }
    for (int e = lane; e < NT * NT; e += 32)
    {
        const int a = e / NT, b = e - a * NT;
        sm.M0inv[e] = sm.P[(NX + a) * LDP + NX + b] + (a == b ? cfg.w_i : 0.0);
    }
    __syncwarp();
    ok = spd_inverse(sm.M0inv, NT, lane) && ok;
    return ok;
}

// vector pass + forward rollout.  gamma: extra linear cost on throttle variables, given as a list of
// (variable index in [0, 4*nblk), coefficient).  hom: linear response only.  If z != nullptr the full
// primal is written there; vout (4*nblk) always receives the throttle blocks.
__device__ void gen_solve(GenCtx& c, int n_gamma, const int* g_idx, const double* g_val, bool hom, double* vout,
                          double* z)
{
    const DeviceConfig& cfg = c.cfg;
    GenSmem& sm = c.sm;
    const int lane = c.lane;
    const int N = cfg.N, NC = cfg.NC;
    const double* qd = c.qd;
    const bool pinned = qd[QD_PINNED] != 0.0;
    for (int e = lane; e < NZ; e += 32)
        sm.p[e] = 0.0;
    __syncwarp();
    for (int k = N - 1; k >= 0; --k)
    {
        const double dt = cfg.dt[k];
        const double* wsk = c.ws + (size_t)k * WS_STAGE;
        const int kind = knot_kind(k, cfg.Ns, cfg.Nc);
        const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
        // s = p (+ stage gradient + Ptilde t)
        for (int i = lane; i < NZ; i += 32)
        {
            double v = sm.p[i];
            if (!hom)
            {
                if (i < 12)
                    v -= cfg.Qd[i] * qd[QD_XREF + i * NC + ref_col(k, cfg.Ns)];
                v += wsk[WS_PTT + i];
            }
            sm.s[i] = v;
        }
        __syncwarp();
        // phi = T^T s
        for (int j = lane; j < NZ; j += 32)
        {
            double a = sm.s[j];
            if (j < NX)
                for (int r = 0; r < NX; ++r)
                    a += dt * sm.A[r * NX + j] * sm.s[r];
            else if (j < NY)
                for (int r = 0; r < NX; ++r)
                    a += dt * sm.BT[r * NT + (j - NX)] * sm.s[r];
            else
                for (int r = 0; r < NX; ++r)
                    a += dt * sm.BJ[r * NJ + (j - NY)] * sm.s[r];
            sm.phi[j] = a;
        }
        __syncwarp();
        if (kind == KIND_T)
        {
            for (int e = lane; e < NZ; e += 32)
                sm.p[e] = sm.phi[e];
            __syncwarp();
            continue;
        }
        const int nu = (kind == KIND_M) ? NU : NJ;
        const int u0 = NZ - nu;
        for (int a = lane; a < nu; a += 32)
        {
            double v = sm.phi[u0 + a];
            const int g = u0 + a;
            if (g >= NY)
            {
                if (!hom)
                    v += qd[QD_GQ + g - NY];
            }
            else
            { // throttle component of an M knot: add gamma of this block
                for (int q = 0; q < n_gamma; ++q)
                    if (g_idx[q] == tb * NT + (g - NX))
                        v += g_val[q];
            }
            sm.hu[a] = v;
        }
        __syncwarp();
        // kff = Hinv hu ; KtH = K^T hu
        for (int a = lane; a < nu; a += 32)
        {
            double v = 0.0;
            for (int m = 0; m < nu; ++m)
                v += wsk[WS_HINV + a * nu + m] * sm.hu[m];
            c.kff[k * NU + a] = v;
        }
        for (int j = lane; j < NY; j += 32)
        {
            double v = 0.0;
            for (int m = 0; m < nu; ++m)
                v += wsk[WS_K + m * NY + j] * sm.hu[m];
            sm.KtH[j] = v;
        }
        __syncwarp();
        for (int j = lane; j < NZ; j += 32)
        {
            double v = 0.0;
            if (j < NY)
            {
                if (kind == KIND_M)
                    v = (j < NX ? sm.phi[j] : 0.0) - sm.KtH[j];
                else
                    v = sm.phi[j] - sm.KtH[j];
            }
            if (kind == KIND_0 && j >= NX && j < NY)
                for (int q = 0; q < n_gamma; ++q)
                    if (g_idx[q] == j - NX)
                        v += g_val[q];
            sm.p[j] = v;
        }
        __syncwarp();
    }
    // ---- forward ----
    for (int e = lane; e < NX; e += 32)
        sm.x[e] = hom ? 0.0 : qd[QD_X0 + e];
    __syncwarp();
    if (lane < NT)
    {
        double v0;
        const double vb = hom ? 0.0 : qd[QD_VBAR + lane];
        if (pinned)
            v0 = vb;
        else
        {
            v0 = 0.0;
            for (int b = 0; b < NT; ++b)
            {
                double rhs = sm.p[NX + b] - cfg.w_i * (hom ? 0.0 : qd[QD_VBAR + b]);
                for (int j = 0; j < NX; ++j)
                    rhs += sm.P0vx[b * NX + j] * sm.x[j];
                v0 -= sm.M0inv[lane * NT + b] * rhs;
            }
        }
        sm.vcur[lane] = v0;
        vout[lane] = v0;
    }
    __syncwarp();
    if (z)
        for (int e = lane; e < NX; e += 32)
            z[e] = sm.x[e];
    for (int k = 0; k < N; ++k)
    {
        const double dt = cfg.dt[k];
        const double* wsk = c.ws + (size_t)k * WS_STAGE;
        const int kind = knot_kind(k, cfg.Ns, cfg.Nc);
        const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
        const int jb = joint_block(k, cfg.Nc);
        if (kind != KIND_T)
        {
            const int nu = (kind == KIND_M) ? NU : NJ;
            for (int e = lane; e < NY; e += 32)
                sm.y[e] = e < NX ? sm.x[e] : sm.vcur[e - NX]; // vcur = previous block for M, v0 for H/0
            __syncwarp();
            for (int a = lane; a < nu; a += 32)
            {
                double v = -c.kff[k * NU + a];
                for (int j = 0; j < NY; ++j)
                    v -= wsk[WS_K + a * NY + j] * sm.y[j];
                sm.u[(NU - nu) + a] = v; // u = (v(4), dq(8)); H/0 fill only the dq part
            }
            __syncwarp();
            if (kind == KIND_M && lane < NT)
            {
                sm.vcur[lane] = sm.u[lane];
                vout[tb * NT + lane] = sm.u[lane];
            }
            if (z && lane < NJ)
                z[NX * (N + 1) + jb * NJ + lane] = sm.u[NT + lane];
            __syncwarp();
        }
        for (int r = lane; r < NX; r += 32)
        {
            double a = sm.x[r] + (hom ? 0.0 : dt * sm.c[r]);
            for (int j = 0; j < NX; ++j)
                a += dt * sm.A[r * NX + j] * sm.x[j];
            for (int j = 0; j < NT; ++j)
                a += dt * sm.BT[r * NT + j] * sm.vcur[j];
            for (int j = 0; j < NJ; ++j)
                a += dt * sm.BJ[r * NJ + j] * sm.u[NT + j];
            sm.xn[r] = a;
        }
        __syncwarp();
        for (int e = lane; e < NX; e += 32)
        {
            sm.x[e] = sm.xn[e];
            if (z)
                z[(k + 1) * NX + e] = sm.xn[e];
        }
        __syncwarp();
    }
    if (z)
    {
        const int base = NX * (N + 1) + cfg.Nc * NJ;
        __syncwarp();
        for (int e = lane; e < cfg.nblk * NT; e += 32)
            z[base + e] = vout[e];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * GEN_WARPS)
qp_generic_kernel(const DeviceConfig* __restrict__ cfgp, int B, const double* __restrict__ qd_all,
                  double* __restrict__ ws_all, double* __restrict__ scratch_all, double* __restrict__ z_all,
                  double* __restrict__ st, double* __restrict__ out_rows, int* __restrict__ status,
                  int* __restrict__ n_factor, int* __restrict__ n_solve, size_t scratch_stride)
{
    extern __shared__ unsigned char smem_raw[];
    const DeviceConfig& cfg = *cfgp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x * GEN_WARPS + warp;
    if (inst >= B)
        return;
    GenSmem& sm = *reinterpret_cast<GenSmem*>(smem_raw + (size_t)warp * sizeof(GenSmem));
    const double* qd = qd_all + (size_t)inst * cfg.qd_stride;
    const int N = cfg.N;
    const int nvtot = cfg.nblk * NT;
    double* scratch = scratch_all + (size_t)inst * scratch_stride;
    const int MAXACT = nvtot;                      // every throttle variable may be active / need its column
    double* gcols = scratch;                       // [nvtot][nvtot]
    double* kff = gcols + (size_t)MAXACT * nvtot;  // [N][NU]
    double* vv = kff + (size_t)N * NU;             // [nvtot] current throttle iterate
    double* vtmp = vv + nvtot;                     // [nvtot]
    if (lane == 0)
        sm.GW = vtmp + nvtot;                      // [nvtot][nvtot + 1]
    __syncwarp();
    double* z = z_all + (size_t)inst * cfg.n_var;
    GenCtx c{cfg, sm, qd, ws_all + (size_t)inst * N * WS_STAGE, gcols, kff, z, lane};

    // expand the structural nonzeros to dense A, B_T, B_J, c
    for (int e = lane; e < NX * NX; e += 32)
        sm.A[e] = 0.0;
    for (int e = lane; e < NX * NT; e += 32)
        sm.BT[e] = 0.0;
    for (int e = lane; e < NX * NJ; e += 32)
        sm.BJ[e] = 0.0;
    for (int e = lane; e < NX; e += 32)
        sm.c[e] = 0.0;
    __syncwarp();
    if (lane == 0)
    {
        const double w0 = qd[QD_OMEGA], w1 = qd[QD_OMEGA + 1], w2 = qd[QD_OMEGA + 2];
        const double mS[9] = {0.0, w2, -w1, -w2, 0.0, w0, w1, -w0, 0.0}; // -S(omega_B)
        for (int a = 0; a < 3; ++a)
        {
            for (int b = 0; b < 3; ++b)
            {
                sm.A[(IX_COM + a) * NX + IX_LIN + b] = qd[QD_RM + a * 3 + b];
                sm.A[(IX_LIN + a) * NX + IX_LIN + b] = mS[a * 3 + b];
                sm.A[(IX_RPY + a) * NX + IX_ANG + b] = qd[QD_WI + a * 3 + b];
                sm.A[(IX_ANG + a) * NX + IX_ANG + b] = mS[a * 3 + b];
            }
            for (int j = 0; j < NT; ++j)
            {
                sm.A[(IX_LIN + a) * NX + IX_T + j] = qd[QD_ALIN + a * NT + j];
                sm.A[(IX_ANG + a) * NX + IX_T + j] = qd[QD_AANG + a * NT + j];
            }
            for (int b = 0; b < NJ; ++b)
            {
                sm.BJ[(IX_LIN + a) * NJ + b] = qd[QD_LLIN + a * NJ + b];
                sm.BJ[(IX_ANG + a) * NJ + b] = qd[QD_LANG + a * NJ + b];
            }
            sm.A[(IX_EP + a) * NX + IX_COM + a] = 1.0;
            sm.A[(IX_ER + a) * NX + IX_RPY + a] = 1.0;
            sm.c[IX_LIN + a] = qd[QD_CL + a];
            sm.c[IX_EP + a] = qd[QD_CEP + a];
            sm.c[IX_ER + a] = qd[QD_CER + a];
        }
        for (int j = 0; j < NT; ++j)
        {
            sm.A[(IX_T + j) * NX + IX_TD + j] = qd[QD_JTT];
            sm.A[(IX_TD + j) * NX + IX_T + j] = qd[QD_JA + j];
            sm.A[(IX_TD + j) * NX + IX_TD + j] = qd[QD_JB + j];
            sm.BT[(IX_TD + j) * NT + j] = qd[QD_JG + j];
            sm.BT[(IX_T + j) * NT + j] = qd[QD_JGT];
            sm.c[IX_TD + j] = qd[QD_CTD + j];
        }
    }
    __syncwarp();

    int stat = VSMPC_STATUS_SOLVED;
    int nf = 1, ns = 0;
    // reject non-finite input data
    {
        bool fin = true;
        for (int e = lane; e < cfg.qd_stride; e += 32)
            fin = fin && isfinite(qd[e]);
        if (!__all_sync(0xffffffffu, fin))
            stat = VSMPC_STATUS_NUMERICAL;
    }
    if (stat == VSMPC_STATUS_SOLVED && !gen_factor(c))
        stat = VSMPC_STATUS_NUMERICAL;

    const bool pinned = qd[QD_PINNED] != 0.0;
    const int first = pinned ? NT : 0; // pinned block 0 is a parameter, not a variable
    const double lo = qd[QD_VMIN], up = qd[QD_VMAX];
    int nW = 0, ncols = 0;
    if (stat == VSMPC_STATUS_SOLVED)
    {
        gen_solve(c, 0, nullptr, nullptr, false, vv, z);
        ns++;
        // ---- Goldfarb-Idnani dual active set on the throttle boxes ----
        // working-set inverse Minv = (S G_WW S)^-1 kept in global scratch (ld = nvtot) and updated by bordering /
        // down-dating: O(nW^2) per iteration for any horizon; var -> slot / position maps avoid list searches
        const double tol = 1e-10;
        int iters = 0;
        int* g_idx = sm.W_idx; // reused as the gamma index list for the final solve
        double* Minv = sm.GW;
        for (int e = lane; e < nvtot; e += 32)
        {
            sm.slot_of[e] = -1;
            sm.wpos[e] = -1;
        }
        __syncwarp();
        while (true)
        {
            // most violated bound among variables not in W
            double best = -1.0;
            int barg = -1;
            for (int e = first + lane; e < nvtot; e += 32)
            {
                if (sm.wpos[e] < 0)
                {
                    const double v = fmax(vv[e] - up, lo - vv[e]);
                    if (v > best)
                    {
                        best = v;
                        barg = e;
                    }
                }
            }
            int p_idx;
            best = warp_max_arg(best, barg < 0 ? 0x7fffffff : barg, &p_idx);
            if (!(best > tol))
                break;
            const double s = (vv[p_idx] - up > lo - vv[p_idx]) ? 1.0 : -1.0;
            double lam_p = 0.0;
            bool fail = false;
            while (true)
            {
                if (++iters > 4 * MAXACT + 64)
                {
                    stat = VSMPC_STATUS_MAX_ITER;
                    fail = true;
                    break;
                }
                // column of G for p_idx (lazy: one homogeneous back-solve)
                int qp = sm.slot_of[p_idx];
                if (qp < 0)
                {
                    if (ncols >= MAXACT)
                    {
                        stat = VSMPC_STATUS_MAX_ITER;
                        fail = true;
                        break;
                    }
                    qp = ncols;
                    __syncwarp();
                    if (lane == 0)
                    {
                        sm.slot_of[p_idx] = qp;
                        sm.gi1 = p_idx;
                        sm.gv1 = 1.0;
                    }
                    __syncwarp();
                    gen_solve(c, 1, &sm.gi1, &sm.gv1, true, vtmp, nullptr);
                    for (int e = lane; e < nvtot; e += 32)
                        gcols[(size_t)qp * nvtot + e] = -vtmp[e];
                    __syncwarp();
                    ncols++;
                    ns++;
                }
                const double* gp = gcols + (size_t)qp * nvtot;
                // gw_a = sgn_a s G[W_a][p];  r = Minv gw  (Minv symmetric: column access is coalesced)
                for (int a2 = lane; a2 < nW; a2 += 32)
                    sm.gw[a2] = sm.sgn[a2] * s * gp[sm.W_idx[a2]];
                __syncwarp();
                double zpart = 0.0, t1 = INFINITY;
                int drop = -1;
                for (int a2 = lane; a2 < nW; a2 += 32)
                {
                    double acc = 0.0;
                    for (int b2 = 0; b2 < nW; ++b2)
                        acc = fma(Minv[(size_t)b2 * nvtot + a2], sm.gw[b2], acc);
                    sm.r[a2] = acc;
                    zpart = fma(acc, sm.gw[a2], zpart);
                    if (acc > 0.0 && sm.lam[a2] / acc < t1)
                    {
                        t1 = sm.lam[a2] / acc;
                        drop = a2;
                    }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    zpart += __shfl_xor_sync(0xffffffffu, zpart, o);
                    const double ot = __shfl_xor_sync(0xffffffffu, t1, o);
                    const int od = __shfl_xor_sync(0xffffffffu, drop, o);
                    if (ot < t1 || (ot == t1 && od >= 0 && (drop < 0 || od < drop)))
                    {
                        t1 = ot;
                        drop = od;
                    }
                }
                __syncwarp();
                const double zp = gp[p_idx] - zpart;
                const double t2 = (zp > 1e-300) ? (s * vv[p_idx] - s * (s > 0 ? up : lo)) / zp : INFINITY;
                const double t = fmin(t1, t2);
                if (!isfinite(t))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                // primal step along s G[:,p] - sum_a r_a sgn_a G[:,W_a]
                for (int e = lane; e < nvtot; e += 32)
                {
                    double zd = s * gp[e];
                    for (int a2 = 0; a2 < nW; ++a2)
                        zd = fma(-sm.r[a2] * sm.sgn[a2], gcols[(size_t)sm.slotW[a2] * nvtot + e], zd);
                    vv[e] -= t * zd;
                }
                for (int a2 = lane; a2 < nW; a2 += 32)
                    sm.lam[a2] -= t * sm.r[a2];
                lam_p += t;
                __syncwarp();
                if (t2 <= t1)
                { // full step: p becomes active; bordered update of Minv (Schur complement = zp)
                    if (nW >= MAXACT)
                    {
                        stat = VSMPC_STATUS_MAX_ITER;
                        fail = true;
                        break;
                    }
                    const double izp = 1.0 / zp;
                    for (int a2 = 0; a2 < nW; ++a2)
                    {
                        const double ra = sm.r[a2] * izp;
                        for (int b2 = lane; b2 < nW; b2 += 32)
                            Minv[(size_t)a2 * nvtot + b2] = fma(ra, sm.r[b2], Minv[(size_t)a2 * nvtot + b2]);
                    }
                    for (int a2 = lane; a2 < nW; a2 += 32)
                    {
                        Minv[(size_t)a2 * nvtot + nW] = -sm.r[a2] * izp;
                        Minv[(size_t)nW * nvtot + a2] = -sm.r[a2] * izp;
                    }
                    if (lane == 0)
                    {
                        Minv[(size_t)nW * nvtot + nW] = izp;
                        sm.W_idx[nW] = p_idx;
                        sm.sgn[nW] = s;
                        sm.lam[nW] = lam_p;
                        sm.slotW[nW] = qp;
                        sm.wpos[p_idx] = nW;
                    }
                    nW++;
                    __syncwarp();
                    break;
                }
                // partial step: position `drop` leaves the working set; down-date Minv, move the last position
                // into the hole
                {
                    const int last = nW - 1;
                    for (int b2 = lane; b2 < nW; b2 += 32)
                        sm.gw[b2] = Minv[(size_t)drop * nvtot + b2];      // row `drop`
                    __syncwarp();
                    const double imdd = 1.0 / sm.gw[drop];
                    for (int a2 = 0; a2 < nW; ++a2)
                    {
                        const double f = sm.gw[a2] * imdd;
                        for (int b2 = lane; b2 < nW; b2 += 32)
                            Minv[(size_t)a2 * nvtot + b2] = fma(-f, sm.gw[b2], Minv[(size_t)a2 * nvtot + b2]);
                    }
                    __syncwarp();
                    if (drop != last)
                    {
                        for (int b2 = lane; b2 < nW; b2 += 32)
                            sm.gw[b2] = Minv[(size_t)last * nvtot + b2];  // row `last`
                        __syncwarp();
                        for (int b2 = lane; b2 < nW; b2 += 32)
                        {
                            Minv[(size_t)drop * nvtot + b2] = sm.gw[b2];
                            Minv[(size_t)b2 * nvtot + drop] = sm.gw[b2];
                        }
                        __syncwarp();
                    }
                    if (lane == 0)
                    {
                        sm.wpos[sm.W_idx[drop]] = -1;
                        if (drop != last)
                        {
                            Minv[(size_t)drop * nvtot + drop] = sm.gw[last];
                            sm.W_idx[drop] = sm.W_idx[last];
                            sm.sgn[drop] = sm.sgn[last];
                            sm.lam[drop] = sm.lam[last];
                            sm.slotW[drop] = sm.slotW[last];
                            sm.wpos[sm.W_idx[drop]] = drop;
                        }
                    }
                    nW--;
                    __syncwarp();
                }
            }
            if (fail)
                break;
        }
        if (stat == VSMPC_STATUS_SOLVED && nW > 0)
        {
            // final primal recovery with the multipliers as linear cost
            if (lane == 0)
                for (int a = 0; a < nW; ++a)
                    sm.r[a] = sm.sgn[a] * sm.lam[a];
            __syncwarp();
            gen_solve(c, nW, g_idx, sm.r, false, vv, z);
            ns++;
            const int base = NX * (N + 1) + cfg.Nc * NJ;
            if (lane == 0)
                for (int a = 0; a < nW; ++a) // land exactly on the bound
                    z[base + sm.W_idx[a]] = sm.sgn[a] > 0 ? up : lo;
            __syncwarp();
        }
    }
    // ---- output extraction (variableSamplingMPC.cpp:88-112) ----
    if (lane == 0)
    {
        status[inst] = stat;
        n_factor[inst] = nf;
        n_solve[inst] = ns;
    }
    if (stat == VSMPC_STATUS_SOLVED)
    {
        double* o = out_rows + (size_t)inst * VSMPC_OUT_DOUBLES;
        const int ibase = NX * (N + 1);
        if (lane < NJ)
        {
            const double dq = z[ibase + lane];
            o[VSMPC_OUT_DELTA_Q + lane] = dq;
            const double acc = st[(size_t)(ST_QACC + lane) * B + inst] + dq;
            st[(size_t)(ST_QACC + lane) * B + inst] = acc;
            o[VSMPC_OUT_JOINTS_REF + lane] = acc;
        }
        if (lane < NT)
        {
            o[VSMPC_OUT_THROTTLE + lane] = destd_throttle_qd(qd, z[ibase + cfg.Nc * NJ + lane]);
            o[VSMPC_OUT_THRUST + lane] = z[NX + IX_T + lane];
            o[VSMPC_OUT_THRUST_DOT + lane] = z[NX + IX_TD + lane];
        }
        if (lane < NX)
            o[VSMPC_OUT_FINAL_STATE + lane] = z[N * NX + lane];
    }
}

bool generic_supported(const DeviceConfig& cfg)
{
    return NT * cfg.nblk <= GEN_MAXV;
}

size_t generic_scratch_doubles(const DeviceConfig& cfg)
{
    const int nvtot = cfg.nblk * NT;
    size_t n = (size_t)nvtot * nvtot + (size_t)cfg.N * NU + 2 * (size_t)nvtot + (size_t)nvtot * (nvtot + 1);
    return (n + 3) & ~(size_t)3;
}

cudaError_t launch_qp_generic(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                              double* ws, double* scratch, double* z, double* st, double* out_rows, int* status,
                              int* n_factor, int* n_solve, cudaStream_t s)
{
    const size_t smem = sizeof(GenSmem) * GEN_WARPS;
    static bool attr_set[64] = {};
    {
        const cudaError_t e = ensure_dynamic_smem(qp_generic_kernel, (int)smem, attr_set);
        if (e != cudaSuccess)
            return e;
    }
    const int grid = (B + GEN_WARPS - 1) / GEN_WARPS;
    qp_generic_kernel<<<grid, 32 * GEN_WARPS, smem, s>>>(d_cfg, B, qd, ws, scratch, z, st, out_rows, status, n_factor,
                                                         n_solve, generic_scratch_doubles(h_cfg));
    return cudaGetLastError();
}

} // namespace vsmpc
