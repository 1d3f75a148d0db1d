// K2 (structured, default) — Riccati / dual-active-set QP kernel specialised to the block structure
// of the linearised iRonCub model.  One warp per MPC instance, FP64 on the CUDA cores.
//
// Replaces IMPCProblem::solve -> OsqpEigen::Solver (MPC/src/IMPCProblem/IMPCProblem.cpp:196-298; OSQP
// 1.0.0 + QDLDL 0.1.8, not vendored) and the output extraction of VariableSamplingMPC::solveMPC
// (MPC/src/variableSamplingMPC/variableSamplingMPC.cpp:88-112) + JetModel::destandardizeThrottle_u2T
// (UT/src/JetModel.cpp:93-109).  Same algorithm as vsmpc_qp_generic.cu / tools/riccati_model.py.
//
// What is structured here:
//  * the knot map z+ = T z + t over z = (x, v, dq) is T = I + dt_k [A_c B_T B_J] where [A_c B_T B_J]
//    has ~120 structural nonzeros in fixed positions: T^T y is evaluated in
//    registers with compile-time indices (applyTt, 134 FMAs instead of 26*38 = 988);
//  * the value-function matrix (38 x 38, symmetric) lives in shared memory, one row/column per lane;
//    the congruence P <- T^T (P+Q) T is done in place as a row pass and a column pass;
//  * eliminated inputs (8 joint increments; 12 with a new throttle block) are removed with an explicit
//    12x12 SPD inverse and a rank-12 update, each lane holding its column of H_uy in registers;
//  * the active-set loop works on the <= 24 throttle variables only; the columns of the reduced inverse
//    Hessian are homogeneous back-solves, computed lazily, R right-hand sides at a time.
// Tensor cores are not used: the dense contractions that would map to DMMA disappear once the
// sparsity of T is used, and on B200 the measured DMMA peak (37 TF) equals the DFMA peak (34 TF).
#include "vsmpc_common.cuh"

namespace vsmpc
{

constexpr int SW = 4;            // instances (warps) per CTA
constexpr int LDW = NZ + 1;      // 39, odd: conflict-free row and column access
constexpr int MAXW = 24;         // cap on working-set size and on lazily computed columns
constexpr int RMAX = 4;          // right-hand sides per multi-RHS back-solve
constexpr int CF = 164;          // coefficient block copied from the QP data (QD_RM .. QD_JGT, padded)
constexpr int NVS = 24;          // throttle variables whose G columns fit in shared memory (reference horizon)
constexpr int NREC = 2 * NJ + NX; // per-solve recorded outputs besides v: dq_0 (8), x_1[T,Td] (8), x_N (26)

// solve-phase vectors, aliased onto StSmem::W (the matrix is dead once the factorisation is stored)
struct SolveVec
{
    double s[2][NZ][RMAX];   // backward: s = p + stage terms (ping-pong)
    double hu[NU][RMAX];
    double z[2][NY][RMAX];   // forward: (x, v in effect at the previous knot) (ping-pong)
    double ue[NU][RMAX];     // forward: (v, dq) in effect at the current knot
    double kff[1];           // [N][NU][RMAX] follows when it fits (reference horizon), else global
};

struct StSmem
{
    double W[NZ * LDW];          // value-function matrix (in place); later SolveVec / active-set matrix
    double Hs[NU * NU];          // inverse of H_uu (broadcast buffer)        } after the factorisation these
    double Ks[NU * 32];          // gain columns (broadcast buffer)            } 576 contiguous doubles hold the
    double hk_pad[MAXW * MAXW - NU * NU - NU * 32]; //                           } active-set inverse (MAXW x MAXW)
    double cf[CF];
    double P0vx[NT * NX];
    double M0inv[NT * NT];
    double gcols[MAXW * NVS];    // lazily computed columns of the reduced inverse Hessian
    double r[MAXW], lam[MAXW], sgn[MAXW], gval[MAXW];
    int W_idx[MAXW];
    int col_of[MAXW];
    int slot_of[MAXW];
    int gidx[RMAX];
    double vbuf[(1 + RMAX) * NVS];
};
constexpr int SV_FIXED = (2 * NZ + NU + 2 * NY + NU) * RMAX; // doubles before SolveVec::kff

// y <- T^T y,  T = I + dt [A_c B_T B_J]   (structure: SURVEY App. A-3)
__device__ __forceinline__ void applyTt(double (&y)[NZ], const double* __restrict__ cf, double dt)
{
    const double c0 = y[IX_COM], c1 = y[IX_COM + 1], c2 = y[IX_COM + 2];
    const double l0 = y[IX_LIN], l1 = y[IX_LIN + 1], l2 = y[IX_LIN + 2];
    const double r0 = y[IX_RPY], r1 = y[IX_RPY + 1], r2 = y[IX_RPY + 2];
    const double a0 = y[IX_ANG], a1 = y[IX_ANG + 1], a2 = y[IX_ANG + 2];
    const double w0 = cf[QD_OMEGA], w1 = cf[QD_OMEGA + 1], w2 = cf[QD_OMEGA + 2];
    const double jtt = cf[QD_JTT], jgt = cf[QD_JGT];
    double T[NT], Td[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        T[j] = y[IX_T + j];
        Td[j] = y[IX_TD + j];
    }
    // columns COM / RPY <- rows posErr / rpyErr (identity blocks)
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
        y[IX_COM + b] += dt * y[IX_EP + b];
        y[IX_RPY + b] += dt * y[IX_ER + b];
    }
    // columns LIN <- rows COM (wRb/m), LIN (-S(w)) ; columns ANG <- rows RPY (W^-1 I^-1), ANG (-S(w))
    {
        const double* Rm = cf + QD_RM;
        const double* WI = cf + QD_WI;
        y[IX_LIN + 0] = l0 + dt * (Rm[0] * c0 + Rm[3] * c1 + Rm[6] * c2 + (w1 * l2 - w2 * l1));
        y[IX_LIN + 1] = l1 + dt * (Rm[1] * c0 + Rm[4] * c1 + Rm[7] * c2 + (w2 * l0 - w0 * l2));
        y[IX_LIN + 2] = l2 + dt * (Rm[2] * c0 + Rm[5] * c1 + Rm[8] * c2 + (w0 * l1 - w1 * l0));
        y[IX_ANG + 0] = a0 + dt * (WI[0] * r0 + WI[3] * r1 + WI[6] * r2 + (w1 * a2 - w2 * a1));
        y[IX_ANG + 1] = a1 + dt * (WI[1] * r0 + WI[4] * r1 + WI[7] * r2 + (w2 * a0 - w0 * a2));
        y[IX_ANG + 2] = a2 + dt * (WI[2] * r0 + WI[5] * r1 + WI[8] * r2 + (w0 * a1 - w1 * a0));
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        const double ja = cf[QD_JA + j], jb = cf[QD_JB + j], jg = cf[QD_JG + j];
        const double al = cf[QD_ALIN + j] * l0 + cf[QD_ALIN + NT + j] * l1 + cf[QD_ALIN + 2 * NT + j] * l2;
        const double aa = cf[QD_AANG + j] * a0 + cf[QD_AANG + NT + j] * a1 + cf[QD_AANG + 2 * NT + j] * a2;
        y[IX_T + j] = T[j] + dt * (al + aa + ja * Td[j]);
        y[IX_TD + j] = Td[j] + dt * (jtt * T[j] + jb * Td[j]);
        y[NX + j] += dt * (jg * Td[j] + jgt * T[j]);
    }
#pragma unroll
    for (int b = 0; b < NJ; ++b)
    {
        const double ll = cf[QD_LLIN + b] * l0 + cf[QD_LLIN + NJ + b] * l1 + cf[QD_LLIN + 2 * NJ + b] * l2;
        const double la = cf[QD_LANG + b] * a0 + cf[QD_LANG + NJ + b] * a1 + cf[QD_LANG + 2 * NJ + b] * a2;
        y[NY + b] += dt * (ll + la);
    }
}

// in-register Gauss-Jordan inverse of an SPD n x n matrix: lane l < n holds row l; pivot rows travel
// by warp shuffle, every index is a compile-time constant
template <int n>
__device__ __forceinline__ bool spd_inverse_reg(double (&row)[n], int lane)
{
    bool ok = true;
#pragma unroll
    for (int p = 0; p < n; ++p)
    {
        double prow[n];
#pragma unroll
        for (int j = 0; j < n; ++j)
            prow[j] = __shfl_sync(0xffffffffu, row[j], p);
        const double d = prow[p];
        ok = ok && (d > 0.0) && isfinite(d);
        const double dinv = 1.0 / d;
        const double f = row[p] * dinv;
        const bool piv = lane == p;
#pragma unroll
        for (int j = 0; j < n; ++j)
        {
            if (j == p)
                row[j] = piv ? dinv : -f;
            else
                row[j] = piv ? prow[j] * dinv : fma(-f, prow[j], row[j]);
        }
    }
    return ok;
}

struct StCtx
{
    const DeviceConfig& cfg;
    StSmem& sm;
    const double* qd;
    double* ws;    // [N][WS_STAGE] factorisation
    double* kff;   // [N][NU][RMAX] feed-forward terms of the pass in flight (shared or global)
    double* rec;   // [1 + MAXW][NREC] recorded outputs of the base solve and of every column (global)
    int lane;
};

// ---- elimination of the inputs introduced at a knot (nu = 8: joint block; 12: throttle + joint block) ----
template <bool isM>
__device__ __forceinline__ bool st_eliminate(StCtx& c, double* wsk)
{
    constexpr int nu = isM ? NU : NJ;
    constexpr int u0 = NZ - nu;
    const DeviceConfig& cfg = c.cfg;
    StSmem& sm = c.sm;
    double* W = sm.W;
    const int lane = c.lane;
    // H_uu row of lanes < nu, inverted in registers
    double hrow[nu];
#pragma unroll
    for (int m = 0; m < nu; ++m)
        hrow[m] = (lane < nu) ? W[(u0 + lane) * LDW + u0 + m] : (m == lane % nu ? 1.0 : 0.0);
    if (lane < nu)
    {
        const int g = u0 + lane;
        const double dg = (g < NY) ? cfg.w_t : cfg.Rqd[(g - NY) & 7];
#pragma unroll
        for (int m = 0; m < nu; ++m)
            if (m == lane)
                hrow[m] += dg;
    }
    // H_uy column of this lane (registers)
    double hu[nu];
#pragma unroll
    for (int m = 0; m < nu; ++m)
        hu[m] = 0.0;
    if (lane < NY)
    {
        if (isM)
        {
            if (lane < NX)
            {
#pragma unroll
                for (int m = 0; m < nu; ++m)
                    hu[m] = W[(u0 + m) * LDW + lane];
            }
            else
            {
#pragma unroll
                for (int m = 0; m < nu; ++m)
                    if (m == lane - NX)
                        hu[m] = -cfg.w_t;
            }
        }
        else
        {
#pragma unroll
            for (int m = 0; m < nu; ++m)
                hu[m] = W[(u0 + m) * LDW + lane];
        }
    }
    const bool ok = spd_inverse_reg<nu>(hrow, lane);
    if (lane < nu)
    {
#pragma unroll
        for (int m = 0; m < nu; ++m)
        {
            sm.Hs[lane * nu + m] = hrow[m];
            wsk[WS_HINV + lane * nu + m] = hrow[m];
        }
    }
    __syncwarp();
    // K column = Hinv * hu
    if (lane < NY)
    {
#pragma unroll
        for (int a = 0; a < nu; ++a)
        {
            double v = 0.0;
#pragma unroll
            for (int m = 0; m < nu; ++m)
                v = fma(sm.Hs[a * nu + m], hu[m], v);
            sm.Ks[a * 32 + lane] = v;
            wsk[WS_K + a * NY + lane] = v;
        }
    }
    __syncwarp();
    // P_yy row of this lane <- base - hu^T K ; clear the eliminated rows / columns
    if (lane < NY)
    {
        double row[NY];
#pragma unroll
        for (int j = 0; j < NY; ++j)
        {
            if (isM)
                row[j] = (lane < NX && j < NX) ? W[lane * LDW + j] : ((lane >= NX && j == lane) ? cfg.w_t : 0.0);
            else
                row[j] = W[lane * LDW + j];
        }
#pragma unroll
        for (int m = 0; m < nu; ++m)
        {
            const double h = hu[m];
#pragma unroll
            for (int j = 0; j < NY; ++j)
                row[j] = fma(-h, sm.Ks[m * 32 + j], row[j]);
        }
#pragma unroll
        for (int j = 0; j < NY; ++j)
            W[lane * LDW + j] = row[j];
#pragma unroll
        for (int a = 0; a < NJ; ++a)
        {
            W[lane * LDW + NY + a] = 0.0;
            W[(NY + a) * LDW + lane] = 0.0;
        }
    }
#pragma unroll
    for (int t = 0; t < 2; ++t)
    {
        const int e = lane + 32 * t;
        W[(NY + (e >> 3)) * LDW + NY + (e & 7)] = 0.0;
    }
    __syncwarp();
    return ok;
}

// ---- factorisation: matrix part of the backward recursion --------------------------------------------
__device__ bool st_factor(StCtx& c)
{
    const DeviceConfig& cfg = c.cfg;
    StSmem& sm = c.sm;
    const int lane = c.lane;
    const int N = cfg.N;
    const double* cf = sm.cf;
    double* W = sm.W;
    bool ok = true;
    for (int e = lane; e < NZ * LDW; e += 32)
        W[e] = 0.0;
    __syncwarp();
    for (int k = N - 1; k >= 0; --k)
    {
        const double dt = cfg.dt[k];
        double* wsk = c.ws + (size_t)k * WS_STAGE;
        const int kind = knot_kind(k, cfg.Ns, cfg.Nc);
        const bool in_d = (k + 1 < N) && knot_kind(k + 1, cfg.Ns, cfg.Nc) == KIND_T;
        if (lane < NX)
            W[lane * LDW + lane] += cfg.Qd[lane];
        __syncwarp();
        // ---- row pass: W[i,:] <- W[i,:] T, and Ptt = Ptilde t (t = dt [c;0;0]) ----
#pragma unroll 1
        for (int rnd = 0; rnd < 2; ++rnd)
        {
            if (rnd == 1 && !in_d)
            {
                if (lane < NJ)
                    wsk[WS_PTT + NY + lane] = 0.0;
                break;
            }
            const int i = rnd == 0 ? lane : NY + lane;
            const bool act = rnd == 0 ? (lane < NY) : (lane < NJ);
            if (act)
            {
                double y[NZ];
#pragma unroll
                for (int j = 0; j < NZ; ++j)
                    y[j] = W[i * LDW + j];
                double acc = 0.0;
#pragma unroll
                for (int a = 0; a < 3; ++a)
                    acc += y[IX_LIN + a] * cf[QD_CL + a] + y[IX_EP + a] * cf[QD_CEP + a] + y[IX_ER + a] * cf[QD_CER + a];
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    acc += y[IX_TD + j] * cf[QD_CTD + j];
                wsk[WS_PTT + i] = dt * acc;
                applyTt(y, cf, dt);
#pragma unroll
                for (int j = 0; j < NZ; ++j)
                    W[i * LDW + j] = y[j];
            }
        }
        __syncwarp();
        // ---- column pass (in place, symmetric result): W[:,j] <- T^T W[:,j] for j < 30 ----
        if (lane < NY)
        {
            double y[NZ];
#pragma unroll
            for (int r = 0; r < NZ; ++r)
                y[r] = W[r * LDW + lane];
            applyTt(y, cf, dt);
#pragma unroll
            for (int r = 0; r < NZ; ++r)
                W[r * LDW + lane] = y[r];
        }
        // (dq,dq) block: the only entries of columns 30..37 not already given by symmetry
        double dd[2];
#pragma unroll
        for (int t = 0; t < 2; ++t)
        {
            const int e = lane + 32 * t; // 64 entries
            const int a = e >> 3, b = e & 7;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < 3; ++q)
                acc += cf[QD_LLIN + q * NJ + a] * W[(IX_LIN + q) * LDW + NY + b]
                       + cf[QD_LANG + q * NJ + a] * W[(IX_ANG + q) * LDW + NY + b];
            dd[t] = W[(NY + a) * LDW + NY + b] + dt * acc;
        }
        __syncwarp();
#pragma unroll
        for (int t = 0; t < 2; ++t)
        {
            const int e = lane + 32 * t;
            W[(NY + (e >> 3)) * LDW + NY + (e & 7)] = dd[t];
        }
        if (lane < NY)
        {
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                W[lane * LDW + NY + a] = W[(NY + a) * LDW + lane];
        }
        __syncwarp();
        if (kind == KIND_T)
            continue;
        if (kind == KIND_M)
            ok = st_eliminate<true>(c, wsk) && ok;
        else
            ok = st_eliminate<false>(c, wsk) && ok;
    }
    // V_0(x0, v0)
    for (int e = lane; e < NT * NX; e += 32)
    {
        const int a = e / NX, j = e - a * NX;
        sm.P0vx[e] = W[(NX + a) * LDW + j];
    }
    {
        double mrow[NT];
#pragma unroll
        for (int b = 0; b < NT; ++b)
            mrow[b] = (lane < NT) ? W[(NX + lane) * LDW + NX + b] + (b == lane ? cfg.w_i : 0.0) : (b == (lane & 3) ? 1.0 : 0.0);
        ok = spd_inverse_reg<NT>(mrow, lane) && ok;
        if (lane < NT)
        {
#pragma unroll
            for (int b = 0; b < NT; ++b)
                sm.M0inv[lane * NT + b] = mrow[b];
        }
    }
    __syncwarp();
    return ok;
}

// ---- per-lane sparse tables of T^T (backward) and T (forward), SURVEY App. A-3 --------------------------
constexpr int QB = 7, QD2 = 6, QF = 8;
struct LaneTab
{
    double cb[QB];   int ib[QB];    // lane j < 30: (T^T s)[j] = s[j] + dt sum_q cb[q] s[ib[q]]
    double cd[QD2];  int id[QD2];   // lane b < 8 : (T^T s)[30+b] = s[30+b] + dt sum_q cd[q] s[id[q]]
    double cw[QF];   int iw[QF];    // forward partial sums: lanes < 26 own row `lane`; lanes 26..31 carry the
                                    // joint-increment part of the linear/angular-momentum rows
    double cc;                      // affine term c_row
    int helper;                     // lane whose partial sum is added to this row (-1: none)
};

__device__ void build_tab(LaneTab& t, const double* __restrict__ cf, int lane)
{
#pragma unroll
    for (int q = 0; q < QB; ++q) { t.cb[q] = 0.0; t.ib[q] = 0; }
#pragma unroll
    for (int q = 0; q < QD2; ++q) { t.cd[q] = 0.0; t.id[q] = 0; }
#pragma unroll
    for (int q = 0; q < QF; ++q) { t.cw[q] = 0.0; t.iw[q] = 0; }
    t.cc = 0.0;
    t.helper = -1;
    const double w[3] = {cf[QD_OMEGA], cf[QD_OMEGA + 1], cf[QD_OMEGA + 2]};
    // mS = -S(w): mS[a][b]
    auto mS = [&](int a, int b) -> double {
        if (a == b) return 0.0;
        const int k = 3 - a - b;               // the remaining axis
        const double sgn = ((b - a + 3) % 3 == 1) ? 1.0 : -1.0; // mS[0][1]=w2, mS[1][2]=w0, mS[2][0]=w1
        return sgn * w[k];
    };
    auto setb = [&](int q, double cv, int iv) {
#pragma unroll
        for (int qq = 0; qq < QB; ++qq)
            if (qq == q) { t.cb[qq] = cv; t.ib[qq] = iv; }
    };
    auto setw = [&](int q, double cv, int iv) {
#pragma unroll
        for (int qq = 0; qq < QF; ++qq)
            if (qq == q) { t.cw[qq] = cv; t.iw[qq] = iv; }
    };
    // ---- backward: column `lane` of [A_c B_T] ----
    const int j = lane;
    if (j < IX_LIN) setb(0, 1.0, IX_EP + j);
    else if (j < IX_RPY)
    {
        const int b = j - IX_LIN;
        for (int a = 0; a < 3; ++a) { setb(a, cf[QD_RM + a * 3 + b], IX_COM + a); setb(3 + a, mS(a, b), IX_LIN + a); }
    }
    else if (j < IX_ANG) setb(0, 1.0, IX_ER + (j - IX_RPY));
    else if (j < IX_T)
    {
        const int b = j - IX_ANG;
        for (int a = 0; a < 3; ++a) { setb(a, cf[QD_WI + a * 3 + b], IX_RPY + a); setb(3 + a, mS(a, b), IX_ANG + a); }
    }
    else if (j < IX_TD)
    {
        const int q = j - IX_T;
        for (int a = 0; a < 3; ++a) { setb(a, cf[QD_ALIN + a * NT + q], IX_LIN + a); setb(3 + a, cf[QD_AANG + a * NT + q], IX_ANG + a); }
        setb(6, cf[QD_JA + q], IX_TD + q);
    }
    else if (j < IX_EP)
    {
        const int q = j - IX_TD;
        setb(0, cf[QD_JTT], IX_T + q);
        setb(1, cf[QD_JB + q], IX_TD + q);
    }
    else if (j >= NX && j < NY)
    {
        const int q = j - NX;
        setb(0, cf[QD_JG + q], IX_TD + q);
        setb(1, cf[QD_JGT], IX_T + q);
    }
    if (lane < NJ)
    {
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            t.cd[a] = cf[QD_LLIN + a * NJ + lane];     t.id[a] = IX_LIN + a;
            t.cd[3 + a] = cf[QD_LANG + a * NJ + lane]; t.id[3 + a] = IX_ANG + a;
        }
    }
    // ---- forward: row `lane` of [A_c B_T B_J c]; sources indexed in (x 0..25, v 26..29, dq 30..37) ----
    const int i = lane;
    if (i < IX_LIN)
        for (int b = 0; b < 3; ++b) setw(b, cf[QD_RM + i * 3 + b], IX_LIN + b);
    else if (i < IX_RPY || (i >= IX_ANG && i < IX_T))
    {
        const bool lin = i < IX_RPY;
        const int a = lin ? i - IX_LIN : i - IX_ANG;
        const int base = lin ? IX_LIN : IX_ANG;
        for (int b = 0; b < 3; ++b) setw(b, mS(a, b), base + b);
        for (int q = 0; q < NT; ++q) setw(3 + q, cf[(lin ? QD_ALIN : QD_AANG) + a * NT + q], IX_T + q);
        t.cc = lin ? cf[QD_CL + a] : 0.0;
        t.helper = (lin ? NX : NX + 3) + a;
    }
    else if (i < IX_ANG)
        for (int b = 0; b < 3; ++b) setw(b, cf[QD_WI + (i - IX_RPY) * 3 + b], IX_ANG + b);
    else if (i < IX_TD)
    {
        const int q = i - IX_T;
        setw(0, cf[QD_JTT], IX_TD + q);
        setw(1, cf[QD_JGT], NX + q);
    }
    else if (i < IX_EP)
    {
        const int q = i - IX_TD;
        setw(0, cf[QD_JA + q], IX_T + q);
        setw(1, cf[QD_JB + q], IX_TD + q);
        setw(2, cf[QD_JG + q], NX + q);
        t.cc = cf[QD_CTD + q];
    }
    else if (i < IX_ER)
    {
        setw(0, 1.0, IX_COM + (i - IX_EP));
        t.cc = cf[QD_CEP + (i - IX_EP)];
    }
    else if (i < NX)
    {
        setw(0, 1.0, IX_RPY + (i - IX_ER));
        t.cc = cf[QD_CER + (i - IX_ER)];
    }
    else
    { // helper lanes 26..28: Lambda_lin rows, 29..31: Lambda_ang rows
        const int a = (i - NX) % 3;
        const bool lin = i < NX + 3;
        for (int b = 0; b < NJ; ++b) setw(b, cf[(lin ? QD_LLIN : QD_LANG) + a * NJ + b], NY + b);
    }
}

// ---- vector pass + forward rollout for R right-hand sides ---------------------------------------------
// RHS r adds the linear cost gval[r] * v[gidx[r]] (gidx < 0: none); for !hom, RHS 0 additionally carries
// the multiplier list (n_list, lidx, lval).  Per RHS r the pass records v (4*nblk) into vout[r] and
// (dq_0, x_1[T,Td], x_N) into rec[r]; z (RHS 0, may be null) receives the full primal.
template <int R>
__device__ void st_solve(StCtx& c, const LaneTab& tab, bool hom, const int* gidx, const double* gval, int n_list,
                         const int* lidx, const double* lval, double* const* vout, double* const* rec, double* z)
{
    const DeviceConfig& cfg = c.cfg;
    StSmem& sm = c.sm;
    const int lane = c.lane;
    const int N = cfg.N, NC = cfg.NC;
    const double* qd = c.qd;
    const double* cf = sm.cf;
    SolveVec& sv = *reinterpret_cast<SolveVec*>(sm.W);
    const bool pinned = cf[QD_PINNED] != 0.0;

    auto gamma = [&](int var, int r) -> double {
        double g = 0.0;
        if (gidx && gidx[r] == var)
            g += gval[r];
        if (r == 0)
            for (int q = 0; q < n_list; ++q)
                if (lidx[q] == var)
                    g += lval[q];
        return g;
    };

    double p[R], pd[R]; // lane j < 30: p[j] ; lanes < 8: pd = p[30 + lane]
#pragma unroll
    for (int r = 0; r < R; ++r)
        p[r] = pd[r] = 0.0;
    int pp = 0;
    for (int k = N - 1; k >= 0; --k)
    {
        const double dt = cfg.dt[k];
        const double* __restrict__ wsk = c.ws + (size_t)k * WS_STAGE;
        const int kind = knot_kind(k, cfg.Ns, cfg.Nc);
        const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
        const bool isM = kind == KIND_M;
        const int nu = isM ? NU : NJ;
        // prefetch the gain column / Hinv row of this lane
        double kc[NU], hr[NU];
        if (kind != KIND_T)
        {
#pragma unroll
            for (int m = 0; m < NU; ++m)
            {
                kc[m] = (m < nu && lane < NY) ? wsk[WS_K + m * NY + lane] : 0.0;
                hr[m] = (m < nu && lane < nu) ? wsk[WS_HINV + lane * nu + m] : 0.0;
            }
        }
        double (*S)[RMAX] = sv.s[pp];
        pp ^= 1;
        double sr[R], sdr[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
            sr[r] = p[r];
            sdr[r] = pd[r];
        }
        if (!hom)
        {
            if (lane < 12)
                sr[0] -= cfg.Qd[lane] * qd[QD_XREF + lane * NC + ref_col(k, cfg.Ns)];
            if (lane < NY)
                sr[0] += wsk[WS_PTT + lane];
            if (lane < NJ)
                sdr[0] += wsk[WS_PTT + NY + lane];
        }
        if (lane < NY)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                S[lane][r] = sr[r];
        }
        if (lane < NJ)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                S[NY + lane][r] = sdr[r];
        }
        __syncwarp();
        // phi = T^T s through the per-lane sparse table
        double phi[R], phid[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
            double a = 0.0, ad = 0.0;
#pragma unroll
            for (int q = 0; q < QB; ++q)
                a = fma(tab.cb[q], S[tab.ib[q]][r], a);
#pragma unroll
            for (int q = 0; q < QD2; ++q)
                ad = fma(tab.cd[q], S[tab.id[q]][r], ad);
            phi[r] = fma(dt, a, sr[r]);
            phid[r] = fma(dt, ad, sdr[r]);
        }
        if (kind == KIND_T)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                p[r] = phi[r];
                pd[r] = phid[r];
            }
            continue;
        }
        if (isM && lane >= NX && lane < NY)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                sv.hu[lane - NX][r] = phi[r] + gamma(tb * NT + lane - NX, r);
        }
        if (lane < NJ)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                sv.hu[(isM ? NT : 0) + lane][r] = phid[r] + ((!hom && r == 0) ? cf[QD_GQ + lane] : 0.0);
        }
        __syncwarp();
        double kf[R], kth[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
            kf[r] = kth[r] = 0.0;
#pragma unroll
        for (int m = 0; m < NU; ++m)
        {
            if (m < nu)
            {
#pragma unroll
                for (int r = 0; r < R; ++r)
                {
                    const double h = sv.hu[m][r];
                    kf[r] = fma(hr[m], h, kf[r]);
                    kth[r] = fma(kc[m], h, kth[r]);
                }
            }
        }
        if (lane < nu)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                c.kff[((size_t)k * NU + lane) * RMAX + r] = kf[r];
        }
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
            if (isM)
                p[r] = (lane < NX ? phi[r] : 0.0) - kth[r];
            else
                p[r] = phi[r] - kth[r];
            if (kind == KIND_0 && lane >= NX && lane < NY)
                p[r] += gamma(lane - NX, r);
            pd[r] = 0.0;
        }
    }
    __syncwarp();
    // ---- forward ----
    int zp = 0;
    {
        double (*Z)[RMAX] = sv.z[0];
        if (lane < NX)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                Z[lane][r] = (hom || r > 0) ? 0.0 : cf[QD_X0 + lane];
        }
        if (lane >= NX && lane < NY)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                sv.hu[lane - NX][r] = p[r]; // p_v of V_0
        }
        if (lane < NU)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                sv.ue[lane][r] = 0.0;
        }
        __syncwarp();
        if (lane < NT)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
            {
                const bool inh = !hom && r == 0;
                double v0;
                if (pinned)
                    v0 = inh ? cf[QD_VBAR + lane] : 0.0;
                else
                {
                    v0 = 0.0;
#pragma unroll
                    for (int b = 0; b < NT; ++b)
                    {
                        double rhs = sv.hu[b][r] - (inh ? cfg.w_i * cf[QD_VBAR + b] : 0.0);
                        if (inh)
                            for (int j = 0; j < NX; ++j)
                                rhs += sm.P0vx[b * NX + j] * Z[j][r];
                        v0 -= sm.M0inv[lane * NT + b] * rhs;
                    }
                }
                Z[NX + lane][r] = v0;
                sv.ue[lane][r] = v0;
                vout[r][lane] = v0;
            }
        }
        if (z && lane < NX)
            z[lane] = (hom) ? 0.0 : cf[QD_X0 + lane];
        __syncwarp();
    }
    for (int k = 0; k < N; ++k)
    {
        const double dt = cfg.dt[k];
        const double* __restrict__ wsk = c.ws + (size_t)k * WS_STAGE;
        const int kind = knot_kind(k, cfg.Ns, cfg.Nc);
        double (*Z)[RMAX] = sv.z[zp];
        double (*Zn)[RMAX] = sv.z[zp ^ 1];
        zp ^= 1;
        if (kind != KIND_T)
        {
            const bool isM = kind == KIND_M;
            const int nu = isM ? NU : NJ;
            const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
            const int jb = joint_block(k, cfg.Nc);
            // u = -K y - kff ; lane = a + 12*half computes half of the dot product of row a
            const int a = lane % NU, half = lane / NU;
            double kv[NY / 2];
            if (half < 2 && a < nu)
            {
#pragma unroll
                for (int t = 0; t < NY / 2; ++t)
                    kv[t] = wsk[WS_K + a * NY + half * (NY / 2) + t];
            }
            else
            {
#pragma unroll
                for (int t = 0; t < NY / 2; ++t)
                    kv[t] = 0.0;
            }
            double part[R];
#pragma unroll
            for (int r = 0; r < R; ++r)
                part[r] = 0.0;
            const int j0 = (half < 2 ? half : 0) * (NY / 2);
#pragma unroll
            for (int t = 0; t < NY / 2; ++t)
            {
#pragma unroll
                for (int r = 0; r < R; ++r)
                    part[r] = fma(kv[t], Z[j0 + t][r], part[r]);
            }
#pragma unroll
            for (int r = 0; r < R; ++r)
                part[r] += __shfl_sync(0xffffffffu, part[r], (lane + NU) & 31);
            if (lane < nu)
            {
#pragma unroll
                for (int r = 0; r < R; ++r)
                {
                    const double uv = -part[r] - c.kff[((size_t)k * NU + lane) * RMAX + r];
                    sv.ue[(NU - nu) + lane][r] = uv;
                    if (isM && lane < NT)
                        vout[r][tb * NT + lane] = uv;
                    if (k == 0)
                        rec[r][lane] = uv; // dq_0 (kind 0 eliminates the joint block only)
                }
                if (z && lane >= nu - NJ)
                    z[NX * (N + 1) + jb * NJ + lane - (nu - NJ)] = -part[0] - c.kff[((size_t)k * NU + lane) * RMAX];
            }
            __syncwarp();
        }
        // x+ = x + dt (A_c x + B_T v + B_J dq + c) through the per-lane sparse table
        double xn[R];
#pragma unroll
        for (int r = 0; r < R; ++r)
        {
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < QF; ++q)
            {
                const int src = tab.iw[q];
                const double sval = (src < NX) ? Z[src][r] : sv.ue[src - NX][r];
                acc = fma(tab.cw[q], sval, acc);
            }
            const double other = __shfl_sync(0xffffffffu, acc, tab.helper < 0 ? lane : tab.helper);
            if (tab.helper >= 0)
                acc += other;
            if (!hom && r == 0)
                acc += tab.cc;
            xn[r] = (lane < NX) ? fma(dt, acc, Z[lane][r]) : ((lane < NY) ? sv.ue[lane - NX][r] : 0.0);
        }
        if (lane < NY)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                Zn[lane][r] = xn[r];
        }
        if (k == 0 && lane >= IX_T && lane < IX_EP)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                rec[r][NJ + lane - IX_T] = xn[r];
        }
        if (k == N - 1 && lane < NX)
        {
#pragma unroll
            for (int r = 0; r < R; ++r)
                rec[r][2 * NJ + lane] = xn[r];
        }
        if (z && lane < NX)
            z[(k + 1) * NX + lane] = xn[0];
        __syncwarp();
    }
    if (z)
    {
        const int base = NX * (N + 1) + cfg.Nc * NJ;
        for (int e = lane; e < cfg.nblk * NT; e += 32)
            z[base + e] = vout[0][e];
    }
    __syncwarp();
}

__global__ void __launch_bounds__(32 * SW)
qp_structured_kernel(const DeviceConfig* __restrict__ cfgp, int B, const double* __restrict__ qd_all,
                     double* __restrict__ ws_all, double* __restrict__ scratch_all, double* __restrict__ z_all,
                     double* __restrict__ st, double* __restrict__ out_rows, int* __restrict__ status,
                     int* __restrict__ n_factor, int* __restrict__ n_solve, size_t scratch_stride, int want_z)
{
    extern __shared__ unsigned char smem_raw[];
    const DeviceConfig& cfg = *cfgp;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x * SW + warp;
    if (inst >= B)
        return;
    StSmem& sm = *reinterpret_cast<StSmem*>(smem_raw + (size_t)warp * sizeof(StSmem));
    const double* qd = qd_all + (size_t)inst * cfg.qd_stride;
    const int N = cfg.N;
    const int nvtot = cfg.nblk * NT;
    // global scratch: [rec (1+MAXW) x NREC][kff N*NU*RMAX][gcols MAXW*nvtot][vv nvtot][vtmp RMAX*nvtot]
    double* scratch = scratch_all + (size_t)inst * scratch_stride;
    double* recb = scratch;
    double* kff_g = recb + (size_t)(1 + MAXW) * NREC;
    double* gcols_g = kff_g + (size_t)N * NU * RMAX;
    double* vbuf_g = gcols_g + (size_t)MAXW * nvtot;
    const bool small = nvtot <= NVS;
    const bool kff_sm = SV_FIXED + N * NU * RMAX <= NZ * LDW;
    double* gcols = small ? sm.gcols : gcols_g;
    double* vv = small ? sm.vbuf : vbuf_g;
    double* vtmp = vv + nvtot;
    double* z = want_z ? z_all + (size_t)inst * cfg.n_var : nullptr;
    SolveVec& svr = *reinterpret_cast<SolveVec*>(sm.W);
    StCtx c{cfg, sm, qd, ws_all + (size_t)inst * N * WS_STAGE, kff_sm ? svr.kff : kff_g, recb, lane};

    bool fin = true;
    for (int e = lane; e < CF; e += 32)
    {
        const double v = qd[e];
        sm.cf[e] = v;
        fin = fin && isfinite(v);
    }
    for (int e = CF + lane; e < cfg.qd_stride; e += 32)
        fin = fin && isfinite(qd[e]);
    __syncwarp();
    int stat = __all_sync(0xffffffffu, fin) ? VSMPC_STATUS_SOLVED : VSMPC_STATUS_NUMERICAL;
    int nf = 0, ns = 0;
    if (stat == VSMPC_STATUS_SOLVED)
    {
        nf = 1;
        if (!st_factor(c))
            stat = VSMPC_STATUS_NUMERICAL;
    }
    const bool pinned = sm.cf[QD_PINNED] != 0.0;
    const int first = pinned ? NT : 0;
    const double lo = sm.cf[QD_VMIN], up = sm.cf[QD_VMAX];
    int nW = 0, ncols = 0;
    if (stat == VSMPC_STATUS_SOLVED)
    {
        LaneTab tab;
        build_tab(tab, sm.cf, lane);
        {
            double* vo[1] = {vv};
            double* ro[1] = {recb};
            st_solve<1>(c, tab, false, nullptr, nullptr, 0, nullptr, nullptr, vo, ro, z);
            ns++;
        }
        // ---- Goldfarb-Idnani dual active set on the throttle boxes, one variable per lane ----
        // per-lane registers: value of variable `lane`, slot of its G column, position in the working set
        const double tol = 1e-10;
        const bool isvar = lane >= first && lane < nvtot;
        double v_e = lane < nvtot ? vv[lane] : 0.0;
        int slot_e = -1, wpos_e = -1;
        double lamW = 0.0;              // lane a < nW: multiplier of working-set position a
        double* Minv = sm.Hs;           // inverse of the signed G_WW (ld = MAXW); Hs/Ks are dead after the factorisation
        int iters = 0;
        bool fail = false;
        while (!fail)
        {
            double viol = (isvar && wpos_e < 0) ? fmax(v_e - up, lo - v_e) : -1.0;
            double best = viol;
            int p_idx = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const double ov = __shfl_xor_sync(0xffffffffu, best, o);
                const int oi = __shfl_xor_sync(0xffffffffu, p_idx, o);
                if (ov > best || (ov == best && oi < p_idx))
                {
                    best = ov;
                    p_idx = oi;
                }
            }
            if (!(best > tol))
                break;
            const double v_p0 = __shfl_sync(0xffffffffu, v_e, p_idx);
            const double s = (v_p0 - up > lo - v_p0) ? 1.0 : -1.0;
            const double bound = s > 0 ? up : lo;
            double lam_p = 0.0;
            while (true)
            {
                if (++iters > 6 * MAXW)
                {
                    stat = VSMPC_STATUS_MAX_ITER;
                    fail = true;
                    break;
                }
                int qp = __shfl_sync(0xffffffffu, slot_e, p_idx);
                if (qp < 0)
                {
                    // lazily compute up to RMAX columns in one multi-RHS homogeneous back-solve: the needed
                    // one plus the currently most violated variables that have no column yet
                    if (ncols >= MAXW)
                    {
                        stat = VSMPC_STATUS_MAX_ITER;
                        fail = true;
                        break;
                    }
                    int gi[RMAX];
                    gi[0] = p_idx;
                    int cnt = 1;
                    double cand = (isvar && slot_e < 0 && lane != p_idx && viol > tol) ? viol : -1.0;
#pragma unroll
                    for (int q = 1; q < RMAX; ++q)
                    {
                        double bv = cand;
                        int bi = lane;
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1)
                        {
                            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                            if (ov > bv || (ov == bv && oi < bi))
                            {
                                bv = ov;
                                bi = oi;
                            }
                        }
                        const bool take = bv > tol && ncols + cnt < MAXW;
                        gi[q] = take ? bi : -1;
                        cnt += take;
                        if (take && lane == bi)
                            cand = -1.0;
                    }
                    double* vo[RMAX];
                    double* ro[RMAX];
#pragma unroll
                    for (int q = 0; q < RMAX; ++q)
                    {
                        vo[q] = vtmp + (size_t)q * nvtot;
                        ro[q] = recb + (size_t)(1 + min(ncols + q, MAXW - 1)) * NREC;
                    }
                    if (lane < RMAX)
                        sm.gval[lane] = 1.0;
                    __syncwarp();
                    if (cnt == 1)
                        st_solve<1>(c, tab, true, gi, sm.gval, 0, nullptr, nullptr, vo, ro, nullptr);
                    else
                        st_solve<RMAX>(c, tab, true, gi, sm.gval, 0, nullptr, nullptr, vo, ro, nullptr);
                    ns += cnt;
#pragma unroll
                    for (int q = 0; q < RMAX; ++q)
                    {
                        if (q < cnt)
                        {
                            if (lane < nvtot)
                                gcols[(size_t)(ncols + q) * nvtot + lane] = -vo[q][lane];
                            if (lane == gi[q])
                                slot_e = ncols + q;
                        }
                    }
                    qp = ncols;
                    ncols += cnt;
                    __syncwarp();
                }
                const double gp_e = lane < nvtot ? gcols[(size_t)qp * nvtot + lane] : 0.0; // G[:, p]
                // r = Minv * gwp,  gwp_a = sgn_a s G[W_a][p]  (lane a < nW owns position a)
                const int widx_a = lane < nW ? sm.W_idx[lane] : 0;
                const double sgn_a = lane < nW ? sm.sgn[lane] : 0.0;
                const double gwp_a = sgn_a * s * __shfl_sync(0xffffffffu, gp_e, widx_a);
                double r_a = 0.0;
                for (int b = 0; b < nW; ++b)
                {
                    const double gb = __shfl_sync(0xffffffffu, gwp_a, b);
                    if (lane < nW)
                        r_a = fma(Minv[lane * MAXW + b], gb, r_a);
                }
                // zp = G_pp - sum_a r_a gwp_a ; t1 = min_{r_a > 0} lam_a / r_a
                double zsum = (lane < nW) ? r_a * gwp_a : 0.0;
                double t1 = (lane < nW && r_a > 0.0) ? lamW / r_a : INFINITY;
                int drop = lane;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
                    const double ot = __shfl_xor_sync(0xffffffffu, t1, o);
                    const int od = __shfl_xor_sync(0xffffffffu, drop, o);
                    if (ot < t1 || (ot == t1 && od < drop))
                    {
                        t1 = ot;
                        drop = od;
                    }
                }
                const double gpp = __shfl_sync(0xffffffffu, gp_e, p_idx);
                const double v_p = __shfl_sync(0xffffffffu, v_e, p_idx);
                const double zp = gpp - zsum;
                const double t2 = (zp > 1e-300) ? (s * v_p - s * bound) / zp : INFINITY;
                const double t = fmin(t1, t2);
                if (!isfinite(t))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                // primal step: v_e -= t * (s G[e][p] - sum_a r_a sgn_a G[e][W_a])
                if (lane < nW)
                    sm.r[lane] = r_a * sgn_a;
                __syncwarp();
                {
                    double zd = s * gp_e;
                    for (int a = 0; a < nW; ++a)
                        if (lane < nvtot)
                            zd = fma(-sm.r[a], gcols[(size_t)sm.slot_of[a] * nvtot + lane], zd);
                    v_e = fma(-t, zd, v_e);
                }
                if (lane < nW)
                    lamW -= t * r_a;
                lam_p += t;
                if (t2 <= t1)
                {
                    // full step: p joins the working set; bordered update of Minv (Schur complement = zp)
                    if (nW >= MAXW)
                    {
                        stat = VSMPC_STATUS_MAX_ITER;
                        fail = true;
                        break;
                    }
                    const double izp = 1.0 / zp;
                    if (lane < nW)
                        sm.lam[lane] = r_a; // plain r for the rank-1 term
                    __syncwarp();
                    if (lane < nW)
                    {
                        for (int b = 0; b < nW; ++b)
                            Minv[lane * MAXW + b] = fma(r_a * izp, sm.lam[b], Minv[lane * MAXW + b]);
                        Minv[lane * MAXW + nW] = -r_a * izp;
                        Minv[nW * MAXW + lane] = -r_a * izp;
                    }
                    if (lane == nW)
                    {
                        Minv[nW * MAXW + nW] = izp;
                        sm.W_idx[nW] = p_idx;
                        sm.slot_of[nW] = qp;
                        sm.sgn[nW] = s;
                        lamW = lam_p;
                    }
                    if (lane == p_idx)
                        wpos_e = nW;
                    nW++;
                    __syncwarp();
                    break;
                }
                // partial step: position `drop` leaves the working set; downdate Minv, move the last
                // position into the hole
                {
                    const int last = nW - 1;
                    const int var_d = sm.W_idx[drop], var_l = sm.W_idx[last];
                    const double mdd = Minv[drop * MAXW + drop];
                    __syncwarp();
                    const double f = lane < nW ? Minv[lane * MAXW + drop] / mdd : 0.0;
                    // row `drop` is needed by every lane: stage it
                    if (lane < nW)
                        sm.lam[lane] = Minv[drop * MAXW + lane];
                    __syncwarp();
                    if (lane < nW)
                        for (int b = 0; b < nW; ++b)
                            Minv[lane * MAXW + b] = fma(-f, sm.lam[b], Minv[lane * MAXW + b]);
                    __syncwarp();
                    // move position `last` into `drop`
                    if (drop != last)
                    {
                        if (lane < nW)
                            sm.lam[lane] = Minv[last * MAXW + lane]; // row last
                        __syncwarp();
                        if (lane < nW)
                        {
                            Minv[drop * MAXW + lane] = sm.lam[lane];
                            Minv[lane * MAXW + drop] = sm.lam[lane]; // symmetric
                        }
                        __syncwarp();
                        if (lane == 0)
                        {
                            Minv[drop * MAXW + drop] = sm.lam[last];
                            sm.W_idx[drop] = var_l;
                            sm.slot_of[drop] = sm.slot_of[last];
                            sm.sgn[drop] = sm.sgn[last];
                        }
                        const double lam_last = __shfl_sync(0xffffffffu, lamW, last);
                        if (lane == drop)
                            lamW = lam_last;
                        if (lane == var_l)
                            wpos_e = drop;
                    }
                    if (lane == var_d)
                        wpos_e = -1;
                    nW--;
                    __syncwarp();
                }
            }
        }
        // publish the final throttle iterate and the multipliers
        if (lane < nvtot)
            vv[lane] = v_e;
        if (lane < nW)
            sm.r[lane] = sm.sgn[lane] * lamW;
        __syncwarp();
        if (stat == VSMPC_STATUS_SOLVED && nW > 0)
        {
            if (want_z)
            {
                // full primal requested: one more inhomogeneous pass with the multipliers as linear cost
                double* vo[1] = {vv};
                double* ro[1] = {recb};
                st_solve<1>(c, tab, false, nullptr, nullptr, nW, sm.W_idx, sm.r, vo, ro, z);
                ns++;
                const int base = NX * (N + 1) + cfg.Nc * NJ;
                if (lane < nW)
                    z[base + sm.W_idx[lane]] = sm.sgn[lane] > 0 ? up : lo;
            }
            else
            {
                // outputs by superposition: base + sum_a (s_a lam_a) * response of column a
                for (int e = lane; e < NREC; e += 32)
                {
                    double v = recb[e];
                    for (int a = 0; a < nW; ++a)
                        v = fma(sm.r[a], recb[(size_t)(1 + sm.slot_of[a]) * NREC + e], v);
                    recb[e] = v;
                }
            }
            if (lane < nW) // land exactly on the bound
                vv[sm.W_idx[lane]] = sm.sgn[lane] > 0 ? up : lo;
            __syncwarp();
        }
    }
    if (lane == 0)
    {
        status[inst] = stat;
        n_factor[inst] = nf;
        n_solve[inst] = ns;
    }
    if (stat == VSMPC_STATUS_SOLVED)
    {
        // output extraction (variableSamplingMPC.cpp:88-112): recb = [dq_0, x_1[T,Td], x_N], vv = throttle blocks
        double* o = out_rows + (size_t)inst * VSMPC_OUT_DOUBLES;
        if (lane < NJ)
        {
            const double dq = recb[lane];
            o[VSMPC_OUT_DELTA_Q + lane] = dq;
            const double acc = st[(size_t)(ST_QACC + lane) * B + inst] + dq;
            st[(size_t)(ST_QACC + lane) * B + inst] = acc;
            o[VSMPC_OUT_JOINTS_REF + lane] = acc;
        }
        if (lane < NT)
        {
            o[VSMPC_OUT_THROTTLE + lane] = destd_throttle_qd(sm.cf, vv[lane]);
            o[VSMPC_OUT_THRUST + lane] = recb[NJ + lane];
            o[VSMPC_OUT_THRUST_DOT + lane] = recb[NJ + NT + lane];
        }
        if (lane < NX)
            o[VSMPC_OUT_FINAL_STATE + lane] = recb[2 * NJ + lane];
    }
}

size_t structured_scratch_doubles(const DeviceConfig& cfg)
{
    const int nvtot = cfg.nblk * NT;
    size_t n = (size_t)(1 + MAXW) * NREC + (size_t)cfg.N * NU * RMAX + (size_t)MAXW * nvtot + (size_t)nvtot * (1 + RMAX);
    return (n + 3) & ~(size_t)3;
}

cudaError_t launch_qp_structured(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                                 double* ws, double* scratch, double* z, double* st, double* out_rows, int* status,
                                 int* n_factor, int* n_solve, int want_z, cudaStream_t s)
{
    const size_t smem = sizeof(StSmem) * SW;
    static bool attr_set[64] = {};
    {
        const cudaError_t e = ensure_dynamic_smem(qp_structured_kernel, (int)smem, attr_set);
        if (e != cudaSuccess)
            return e;
    }
    const int grid = (B + SW - 1) / SW;
    qp_structured_kernel<<<grid, 32 * SW, smem, s>>>(d_cfg, B, qd, ws, scratch, z, st, out_rows, status, n_factor,
                                                     n_solve, structured_scratch_doubles(h_cfg), want_z);
    return cudaGetLastError();
}

} // namespace vsmpc
