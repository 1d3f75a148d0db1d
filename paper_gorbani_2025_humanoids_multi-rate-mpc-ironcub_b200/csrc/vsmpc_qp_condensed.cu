// K2 (default) — condensed-throttle Riccati QP kernel.  Two warps per MPC instance, FP64 on the CUDA cores.
//
// Replaces IMPCProblem::solve -> OsqpEigen::Solver (MPC/src/IMPCProblem/IMPCProblem.cpp:196-298; OSQP 1.0.0 +
// QDLDL 0.1.8, not vendored) and the output extraction of VariableSamplingMPC::solveMPC
// (MPC/src/variableSamplingMPC/variableSamplingMPC.cpp:88-112) + JetModel::destandardizeThrottle_u2T
// (UT/src/JetModel.cpp:93-109).  tools/condensed_model.py is the executable NumPy specification.
//
// Algorithm.  The value function at knot k is kept as
//       V_k(x; theta) = 1/2 x'P x + x'Psi theta + 1/2 theta'Om theta,   theta = (v_0..v_5, 1, held joint block)
// * warp A owns P (26 x 26, lane i = row i, in registers): the ordinary Riccati recursion of the
//   joint-increment LQR — congruence with T = I + dt A_c applied in registers (structure of SURVEY App. A-3),
//   transposition through shared memory, 8 x 8 elimination with an in-register SPD inverse;
// * warp B owns the parameter columns Psi (26 x 32, lane l = column l, in registers) and Om (25 x 25, shared
//   memory): linear propagation and rank-8 down-dates driven by what warp A publishes per knot
//   (P'D, H_ux, H_uu^-1) — software-pipelined one knot behind warp A, one __syncthreads per knot;
// * after knot 0 the reduced Hessian of the <= 24 throttle variables is explicit: it is inverted in registers
//   and the Goldfarb-Idnani dual active set runs on the boxes with all columns available (no back-solves);
// * one forward pass with the stored gains K_k (8 x 26) and F_k (8 x 32) produces the outputs.
// Compared with vsmpc_qp_structured.cu (one warp, 38-dim augmented state, one Riccati back-solve per active
// bound) the serial depth drops from (1 + n_s) x 34 knot steps to 17 + 17.
#include "vsmpc_common.cuh"

namespace vsmpc
{

constexpr int CD_THREADS = 64;
constexpr int NL = 32;        // parameter columns (lanes of warp B)
constexpr int NLO = 25;       // rows/columns of Om in use: 24 throttle variables + affine (held block aliases lanes)
constexpr int AFFL = 24;      // lane of the affine column
constexpr int LDM = NX + 1;   // 27, odd: conflict-free transposition
constexpr int NPD = 13;       // published P'D columns: throttle block (4), affine (1), joint block (8)
constexpr int LDH = 10;       // leading dimension of Hut rows (16-byte aligned)
constexpr int CCF = 164;      // coefficient block copied from the QP data (QD_RM .. QD_JGT, padded)
constexpr int CD_MAXNC = 16;  // reference columns kept in shared memory
constexpr int CD_MAXW = 24;
constexpr int WSC_K = 0;              // per elimination knot in the workspace: K [8][26]
constexpr int WSC_F = NJ * NX;        //                                     then F [8][32]
constexpr int WSC_H = NJ * NX + NJ * NL; //                                     then H_utheta [8][32]
constexpr int WSC_STAGE = NJ * NX + 2 * NJ * NL; // 720

struct alignas(16) CdSlot
{
    double Hux[NJ * NX];    // [m][j]
    double Hinv[NJ * NJ];   // [a][m]
    double PD[NX * NPD];    // [i][col]
};

constexpr int CD_MAXN = 32;   // knots kept in shared memory (dt grid)
constexpr int LDG = 26;       // leading dimension of the reduced Hessian / its inverse (16-byte aligned rows)

struct alignas(16) CdSmem
{
    double cf[CCF];
    alignas(16) double lam[6 * NJ];         // dt-free B_J rows: [q][a], q = 0..2 linear, 3..5 angular momentum
    double Qd[NX];
    double Rqd[NJ];
    double dtk[CD_MAXN];
    double xref[12 * CD_MAXNC];
    CdSlot slot[2];             // A -> B mailbox, slot = knot & 1; after the factorisation: G and the working-set inverse
    alignas(16) double Mt[NX * LDM];        // warp A: transposition buffer, then the gain rows K [8][26] of the knot in
                                // flight; after the factorisation: F theta, x, dq
    double Om[NLO * NLO];
    alignas(16) double Hut[4 * CD_MAXW];    // active-set vectors (r, lambda, sign, index) after the factorisation
    alignas(16) double theta[NL];
    int flags[4];
};

enum : int { TK_NONE = 0, TK_STAGE = 1, TK_PROP = 2, TK_SCHUR = 3 };

#ifdef VSMPC_PHASE_CLOCKS
__device__ long long g_phase_clk[4096][16];
#define SUBCLK(acc, t0) do { const long long t1__ = clock64(); acc += t1__ - t0; t0 = t1__; } while (0)
#define PHASE_CLK(slot) do { if (lane == 0 && warp == (slot >= 4 ? 1 : 0) && inst < 4096) g_phase_clk[inst][slot] = clock64(); } while (0)
#else
#define PHASE_CLK(slot) do { } while (0)
#define SUBCLK(acc, t0) do { } while (0)
#endif

// y <- T_x^T y,  T_x = I + dt A_c   (structure: SURVEY App. A-3)
__device__ __forceinline__ void applyTtx(double (&y)[NX], const double* __restrict__ cf, double dt)
{
    const double c0 = y[IX_COM], c1 = y[IX_COM + 1], c2 = y[IX_COM + 2];
    const double l0 = y[IX_LIN], l1 = y[IX_LIN + 1], l2 = y[IX_LIN + 2];
    const double r0 = y[IX_RPY], r1 = y[IX_RPY + 1], r2 = y[IX_RPY + 2];
    const double a0 = y[IX_ANG], a1 = y[IX_ANG + 1], a2 = y[IX_ANG + 2];
    const double w0 = cf[QD_OMEGA], w1 = cf[QD_OMEGA + 1], w2 = cf[QD_OMEGA + 2];
    const double jtt = cf[QD_JTT];
#pragma unroll
    for (int b = 0; b < 3; ++b)
    {
        y[IX_COM + b] += dt * y[IX_EP + b];
        y[IX_RPY + b] += dt * y[IX_ER + b];
    }
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
        const double T = y[IX_T + j], Td = y[IX_TD + j];
        const double al = cf[QD_ALIN + j] * l0 + cf[QD_ALIN + NT + j] * l1 + cf[QD_ALIN + 2 * NT + j] * l2;
        const double aa = cf[QD_AANG + j] * a0 + cf[QD_AANG + NT + j] * a1 + cf[QD_AANG + 2 * NT + j] * a2;
        y[IX_T + j] = T + dt * (al + aa + cf[QD_JA + j] * Td);
        y[IX_TD + j] = Td + dt * (jtt * T + cf[QD_JB + j] * Td);
    }
}

// out[a] = dt * B_J[:, a]' y  (B_J has the six momentum rows only; lam = [q][a] in shared memory, 16-byte aligned)
__device__ __forceinline__ void bjT_dot(const double (&y)[NX], const double* __restrict__ lam, double dt, double (&out)[NJ])
{
    const double2* l2 = reinterpret_cast<const double2*>(lam);
#pragma unroll
    for (int a = 0; a < NJ; ++a)
        out[a] = 0.0;
#pragma unroll
    for (int q = 0; q < 6; ++q)
    {
        const double yv = dt * y[(q < 3 ? IX_LIN : IX_ANG - 3) + q];
#pragma unroll
        for (int a2 = 0; a2 < NJ / 2; ++a2)
        {
            const double2 lv = l2[q * (NJ / 2) + a2];
            out[2 * a2] = fma(lv.x, yv, out[2 * a2]);
            out[2 * a2 + 1] = fma(lv.y, yv, out[2 * a2 + 1]);
        }
    }
}

// dt * c' y   (c: affine term of the dynamics; rows LIN, TD, EP, ER); three independent FMA chains
__device__ __forceinline__ double c_dot(const double (&y)[NX], const double* __restrict__ cf, double dt)
{
    double a0 = 0.0, a1 = 0.0, a2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        a0 = fma(y[IX_LIN + a], cf[QD_CL + a], a0);
        a1 = fma(y[IX_EP + a], cf[QD_CEP + a], a1);
        a2 = fma(y[IX_ER + a], cf[QD_CER + a], a2);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        if (j & 1)
            a1 = fma(y[IX_TD + j], cf[QD_CTD + j], a1);
        else
            a0 = fma(y[IX_TD + j], cf[QD_CTD + j], a0);
    }
    return dt * (a0 + a1 + a2);
}

// Gauss-Jordan inverse of an SPD 8 x 8 matrix through shared memory by one warp: lane (r = lane & 7, q = lane >> 3)
// owns element pair [r][2q..2q+1] in registers; every pivot step reads the pivot row / column from one buffer and
// writes the updated pairs to the other (ping-pong S <-> T: one __syncwarp per pivot, no divergent branches); eight
// pivots later the inverse is back in S.  S, T: row-major, ld 8, 16-byte aligned.
__device__ __forceinline__ bool gj8(double* __restrict__ S, double* __restrict__ T, double2 own, int lane)
{
    const int r = lane & 7, q = lane >> 3;
    reinterpret_cast<double2*>(S)[r * 4 + q] = own;
    __syncwarp();
    bool ok = true;
    double* src = S;
    double* dst = T;
#pragma unroll 2
    for (int p = 0; p < NJ; ++p)
    {
        const double d = src[p * NJ + p];
        const double f = src[r * NJ + p];
        const double2 pr = reinterpret_cast<const double2*>(src)[p * 4 + q];
        ok = ok && (d > 0.0) && (d < 1e300);
        const double dinv = __drcp_rn(d);
        const bool piv = r == p;
        const double coef = piv ? -dinv : f * dinv;      // pivot row: 0 - (-1/d) * row ; others: own - (f/d) * row
        const double bx = piv ? 0.0 : own.x, by = piv ? 0.0 : own.y;
        own.x = fma(-coef, pr.x, bx);
        own.y = fma(-coef, pr.y, by);
        const double val = piv ? dinv : -coef;           // column p of the inverse in progress
        const bool mine = q == (p >> 1);
        own.x = (mine && !(p & 1)) ? val : own.x;
        own.y = (mine && (p & 1)) ? val : own.y;
        reinterpret_cast<double2*>(dst)[r * 4 + q] = own;
        __syncwarp();
        double* t = src;
        src = dst;
        dst = t;
    }
    return ok;
}

// Gauss-Jordan inverse of an SPD matrix of order <= 24 in shared memory (leading dimension ld, rows 16-byte aligned),
// in place, pivots p0..p1-1 (the other rows/columns must be decoupled unit rows); lane l < 24 owns row l
__device__ __forceinline__ bool gj24(double* __restrict__ S, int ld, int lane, int p0, int p1)
{
    bool ok = true;
    const int l = lane < CD_MAXW ? lane : 0;
    const double2* rowl = reinterpret_cast<const double2*>(S + l * ld);
#pragma unroll 1
    for (int p = p0; p < p1; ++p)
    {
        const double d = S[p * ld + p];
        const double f = S[l * ld + p];
        const double2* rowp = reinterpret_cast<const double2*>(S + p * ld);
        double2 pr[CD_MAXW / 2], ow[CD_MAXW / 2];
#pragma unroll
        for (int j = 0; j < CD_MAXW / 2; ++j)
        {
            pr[j] = rowp[j];
            ow[j] = rowl[j];
        }
        ok = ok && (d > 0.0) && isfinite(d);
        const double dinv = 1.0 / d;
        __syncwarp();
        if (lane < CD_MAXW)
        {
            const bool piv = l == p;
            const double ff = piv ? -dinv : f * dinv;
            double2* wr = reinterpret_cast<double2*>(S + l * ld);
#pragma unroll
            for (int j = 0; j < CD_MAXW / 2; ++j)
            {
                double2 o = ow[j];
                if (piv)
                    o = make_double2(0.0, 0.0);
                o.x = fma(-ff, pr[j].x, o.x);
                o.y = fma(-ff, pr[j].y, o.y);
                wr[j] = o;
            }
            S[l * ld + p] = piv ? dinv : -ff;
        }
        __syncwarp();
    }
    return ok;
}

// exact max / min of NON-NEGATIVE doubles over the warp with two 32-bit REDUX each (non-negative doubles order like
// their bit patterns); arg = lowest lane attaining it
__device__ __forceinline__ double warp_max_nonneg(double v, int& arg)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
    arg = __ffs(__ballot_sync(0xffffffffu, hi == mhi && lo == mlo)) - 1;
    return __hiloint2double((int)mhi, (int)mlo);
}
__device__ __forceinline__ double warp_min_nonneg(double v, int& arg)
{
    const unsigned hi = (unsigned)__double2hiint(v), lo = (unsigned)__double2loint(v);
    const unsigned mhi = __reduce_min_sync(0xffffffffu, hi);
    const unsigned mlo = __reduce_min_sync(0xffffffffu, hi == mhi ? lo : 0xffffffffu);
    arg = __ffs(__ballot_sync(0xffffffffu, hi == mhi && lo == mlo)) - 1;
    return __hiloint2double((int)mhi, (int)mlo);
}

struct CdCtx
{
    const DeviceConfig& cfg;
    CdSmem& sm;
    double* ws;   // [Nc][WSC_STAGE]
    int lane;
    int D0;       // first lane of the held joint block during the tail
    int kS;       // knot where the held joint block is eliminated (-1: none)
};

// ---- warp A --------------------------------------------------------------------------------------------------
// steps i-k of a knot: invert H_uu (own = this lane's pair of it), K = H_uu^-1 H_ux, P <- P - H_ux' K
__device__ __forceinline__ bool a_eliminate(const CdCtx& c, CdSlot& sl, double (&p)[NX], const double (&hux)[NJ],
                                            double2 own, double* __restrict__ wsk)
{
    CdSmem& sm = c.sm;
    const int lane = c.lane;
    const bool ok = gj8(sl.Hinv, sm.Mt, own, lane);   // Mt: free between the transposition and the gain rows
    if (lane < NX)
    {
        const double2* hi = reinterpret_cast<const double2*>(sl.Hinv);
#pragma unroll 2
        for (int a = 0; a < NJ; ++a)
        {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int m = 0; m < NJ / 2; ++m)
            {
                const double2 hh = hi[a * (NJ / 2) + m];
                v0 = fma(hh.x, hux[2 * m], v0);
                v1 = fma(hh.y, hux[2 * m + 1], v1);
            }
            sm.Mt[a * NX + lane] = v0 + v1;
            wsk[WSC_K + a * NX + lane] = v0 + v1;
        }
    }
    __syncwarp();
    if (lane < NX)
    {
#pragma unroll 2
        for (int m = 0; m < NJ; ++m)
        {
            const double h = sl.Hux[m * NX + lane];
            const double2* kr = reinterpret_cast<const double2*>(sm.Mt + m * NX);
#pragma unroll
            for (int j = 0; j < NX / 2; ++j)
            {
                const double2 kk = kr[j];
                p[2 * j] = fma(-h, kk.x, p[2 * j]);
                p[2 * j + 1] = fma(-h, kk.y, p[2 * j + 1]);
            }
        }
    }
    __syncwarp();
    return ok;
}

// propagation of P through knot k: P' = P + Q, publish P'D, P <- T'P'T; with elim also H_ux (published) and this
// lane's pair of H_uu = R + B_u' P' B_u
__device__ __forceinline__ void a_prop(const CdCtx& c, int k, bool elim, double (&p)[NX], double qd_lane,
                                       double (&hux)[NJ], double2& own)
{
    CdSmem& sm = c.sm;
    const int lane = c.lane;
    const double* cf = sm.cf;
    const double dt = sm.dtk[k];
    CdSlot& sl = sm.slot[k & 1];
    // P' = P + Q on the diagonal element this lane owns (predicated add: keeps the compiler from turning the
    // 26-way ownership test into a divergent jump table)
#pragma unroll
    for (int j = 0; j < NX; ++j)
        asm("{ .reg .pred q; setp.eq.s32 q, %1, %2; @q add.f64 %0, %0, %3; }" : "+d"(p[j]) : "r"(lane), "r"(j), "d"(qd_lane));
    if (lane < NX)
    {
        // P'D for the throttle / affine columns of this knot
        const double jgt = cf[QD_JGT];
#pragma unroll
        for (int q = 0; q < NT; ++q)
            sl.PD[lane * NPD + q] = dt * (cf[QD_JG + q] * p[IX_TD + q] + jgt * p[IX_T + q]);
        sl.PD[lane * NPD + 4] = c_dot(p, cf, dt);
    }
    // pass 0: (P'D)_joint = dt P' B_J from row i of P', then row i of M = P'T, transposition;
    // pass 1: H_ux[:, i] = B_u' M[:, i] from column i of M, then column i of T'M = row i of T'P'T
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass)
    {
        if (lane < NX)
        {
            if (pass == 0 || elim)
            {
                bjT_dot(p, sm.lam, dt, hux);
                double* dst = pass == 0 ? sl.PD + lane * NPD + 5 : sl.Hux + lane;
                const int stride = pass == 0 ? 1 : NX;
#pragma unroll
                for (int a = 0; a < NJ; ++a)
                    dst[a * stride] = hux[a];
            }
            applyTtx(p, cf, dt);
            if (pass == 0)
            {
#pragma unroll
                for (int j = 0; j < NX; ++j)
                    sm.Mt[lane * LDM + j] = p[j];
            }
        }
        __syncwarp();
        if (pass == 0 && lane < NX)
        {
#pragma unroll
            for (int j = 0; j < NX; ++j)
                p[j] = sm.Mt[j * LDM + lane];
        }
    }
    if (elim)
    {
        // H_uu[r][2q..2q+1] = R + dt Lambda[:, r]' (P'D)_joint[momentum rows, 2q..2q+1]
        const int r = lane & 7, q = lane >> 3;
        double hx = (2 * q == r) ? sm.Rqd[r] : 0.0, hy = (2 * q + 1 == r) ? sm.Rqd[r] : 0.0;
#pragma unroll
        for (int m = 0; m < 6; ++m)
        {
            const double lv = dt * sm.lam[m * NJ + r];
            const double* pd = sl.PD + ((m < 3 ? IX_LIN : IX_ANG - 3) + m) * NPD + 5 + 2 * q;
            hx = fma(lv, pd[0], hx);
            hy = fma(lv, pd[1], hy);
        }
        own = make_double2(hx, hy);
    }
}

// ---- warp B --------------------------------------------------------------------------------------------------
// down-date of the parameter columns with the eliminated block: F = H_uu^-1 H_utheta, Psi -= H_ux' F.  The matching
// down-date of Om (Om -= H_utheta' F) is NOT done knot by knot: H_utheta and F of every elimination knot are stacked
// in the workspace and contracted once after the recursion on the FP64 tensor cores (cd_omega_downdate).
__device__ __forceinline__ void b_downdate(const CdCtx& c, CdSlot& sl, double (&s)[NX], const double (&hut)[NJ],
                                           double* __restrict__ wsk, bool clear_col)
{
    const int lane = c.lane;
    double* Fs = sl.PD;   // the P'D columns of this knot were consumed by b_prop: reuse as F [l][LDH]
#pragma unroll
    for (int m = 0; m < NJ; ++m)
        wsk[WSC_H + m * NL + lane] = hut[m];
    {
        const double2* hi = reinterpret_cast<const double2*>(sl.Hinv);
#pragma unroll 1
        for (int a = 0; a < NJ; ++a)
        {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int m = 0; m < NJ / 2; ++m)
            {
                const double2 hh = hi[a * (NJ / 2) + m];
                v0 = fma(hh.x, hut[2 * m], v0);
                v1 = fma(hh.y, hut[2 * m + 1], v1);
            }
            if (lane < NLO)
                Fs[lane * LDH + a] = v0 + v1;
            wsk[WSC_F + a * NL + lane] = v0 + v1;
        }
    }
    if (lane < NLO)
    {
#pragma unroll 2
        for (int m = 0; m < NJ; ++m)
        {
            const double f = Fs[lane * LDH + m];
            const double2* hr = reinterpret_cast<const double2*>(sl.Hux + m * NX);
#pragma unroll
            for (int j = 0; j < NX / 2; ++j)
            {
                const double2 hh = hr[j];
                s[2 * j] = fma(-f, hh.x, s[2 * j]);
                s[2 * j + 1] = fma(-f, hh.y, s[2 * j + 1]);
            }
        }
    }
    if (clear_col)
    {
#pragma unroll
        for (int j = 0; j < NX; ++j)
            s[j] = 0.0;
    }
    __syncwarp();
}

// Om -= sum_k H_utheta_k' F_k = H' F with H, F the (8 Nc) x 32 stacks of the workspace: a dense 25 x 25 x (8 Nc)
// contraction, run once on the FP64 tensor cores (mma.sync.m8n8k4.f64).  The symmetric result is computed on the ten
// upper 8 x 8 tiles: tile rows {0, 3} by warp 0, {1, 2} by warp 1 (five tiles each); A fragment = H' (lane l: row l >> 2 of the tile,
// stack row l & 3), B fragment = F (stack row l & 3, column l >> 2), C fragment: row l >> 2, columns 2 (l & 3) + {0, 1}.
template <int TA0, int TA1>
__device__ __forceinline__ void cd_omega_tile_rows(const double* __restrict__ ws, int n_rows, double* __restrict__ Om, int lane)
{
    // tile rows TA0 < TA1 of the upper triangle in one sweep over the stack, so that all their loads are in flight together
    constexpr int N0 = 4 - TA0, N1 = 4 - TA1;      // tiles (TA0, TA0..3) and (TA1, TA1..3)
    double c0[N0 + N1], c1[N0 + N1];
#pragma unroll
    for (int t = 0; t < N0 + N1; ++t)
        c0[t] = c1[t] = 0.0;
    const int lr = lane & 3, lc = lane >> 2;
#pragma unroll 8
    for (int r0 = 0; r0 < n_rows; r0 += 4)
    {
        const int r = r0 + lr;
        const double* __restrict__ row = ws + (size_t)(r >> 3) * WSC_STAGE + (r & 7) * NL;
        const double a0 = row[WSC_H + 8 * TA0 + lc];
        const double a1 = row[WSC_H + 8 * TA1 + lc];
        double b[N0];
#pragma unroll
        for (int t = 0; t < N0; ++t)
            b[t] = row[WSC_F + 8 * (TA0 + t) + lc];
#pragma unroll
        for (int t = 0; t < N0; ++t)
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0[t]), "+d"(c1[t])
                : "d"(a0), "d"(b[t]));
#pragma unroll
        for (int t = 0; t < N1; ++t)
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0[N0 + t]), "+d"(c1[N0 + t])
                : "d"(a1), "d"(b[TA1 - TA0 + t]));
    }
#pragma unroll
    for (int q = 0; q < N0 + N1; ++q)
    {
        const int ti = q < N0 ? TA0 : TA1;
        const int t = q < N0 ? q : q - N0;
        const int gi = 8 * ti + lc;
#pragma unroll
        for (int e = 0; e < 2; ++e)
        {
            const int gj = 8 * (ti + t) + 2 * lr + e;
            const double v = e == 0 ? c0[q] : c1[q];
            if (gi < NLO && gj < NLO)
            {
                Om[gi * NLO + gj] -= v;
                if (t > 0)
                    Om[gj * NLO + gi] -= v;
            }
        }
    }
}

__device__ __forceinline__ void cd_omega_downdate(const double* __restrict__ ws, int n_rows, double* __restrict__ Om,
                                                  int warp, int lane)
{
    if (warp == 0)
        cd_omega_tile_rows<0, 3>(ws, n_rows, Om, lane);
    else
        cd_omega_tile_rows<1, 2>(ws, n_rows, Om, lane);
}

// propagation of the parameter columns through knot k (Psi'' = Psi' + P'D, Om += D'Psi'' + Psi''D, Psi <- T'Psi'');
// returns dt B_J' Psi''[:, l] (= H_utheta column without the gradient term) in bj2
__device__ __forceinline__ void b_prop(const CdCtx& c, int k, bool tail, double (&s)[NX], double (&bj2)[NJ])
{
    const DeviceConfig& cfg = c.cfg;
    CdSmem& sm = c.sm;
    const int lane = c.lane;
    const double* cf = sm.cf;
    const double dt = sm.dtk[k];
    CdSlot& sl = sm.slot[k & 1];
    const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
    const bool isAff = lane == AFFL;
    const bool spV = lane < 4 * cfg.nblk && (lane >> 2) == tb;
    const bool isD = tail && lane >= c.D0 && lane < c.D0 + NJ;
    const double jgt = cf[QD_JGT];
    if (isAff)
    {
        const int rc = ref_col(k, cfg.Ns);
#pragma unroll
        for (int r = 0; r < 12; ++r)
            s[r] = fma(-sm.Qd[r], sm.xref[r * cfg.NC + rc], s[r]);   // tracking gradient of x_{k+1}
    }
    // b-terms: D_s' Psi'_l
    double bv[NT], bd[NJ];
#pragma unroll
    for (int q = 0; q < NT; ++q)
        bv[q] = dt * (cf[QD_JG + q] * s[IX_TD + q] + jgt * s[IX_T + q]);
    const double baff = c_dot(s, cf, dt);
    if (tail)
        bjT_dot(s, sm.lam, dt, bd);
    // Psi'' = Psi' + P'D
    const int col = spV ? (lane & 3) : (isAff ? 4 : (isD ? 5 + lane - c.D0 : -1));
    if (col >= 0)
    {
#pragma unroll
        for (int r = 0; r < NX; ++r)
            s[r] += sl.PD[r * NPD + col];
    }
    // a-terms: D_s' Psi''_l
    bjT_dot(s, sm.lam, dt, bj2);
    // Om += D'Psi'' (rows of the special columns) then Om += Psi''D (columns of the special columns): the loads of a
    // group before its stores (distinct entries), one warp barrier between the two phases
    if (lane < NLO)
    {
        {
            double ov[NT];
#pragma unroll
            for (int q = 0; q < NT; ++q)
                ov[q] = sm.Om[(4 * tb + q) * NLO + lane];
            const double oa = sm.Om[AFFL * NLO + lane];
#pragma unroll
            for (int q = 0; q < NT; ++q)
                sm.Om[(4 * tb + q) * NLO + lane] = ov[q] + dt * (cf[QD_JG + q] * s[IX_TD + q] + jgt * s[IX_T + q]);
            sm.Om[AFFL * NLO + lane] = oa + c_dot(s, cf, dt);
        }
        if (tail)
        {
            double od[NJ];
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                od[a] = sm.Om[(c.D0 + a) * NLO + lane];
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                sm.Om[(c.D0 + a) * NLO + lane] = od[a] + bj2[a];
        }
    }
    __syncwarp();
    if (lane < NLO)
    {
        {
            double ov[NT];
#pragma unroll
            for (int q = 0; q < NT; ++q)
                ov[q] = sm.Om[lane * NLO + 4 * tb + q];
            const double oa = sm.Om[lane * NLO + AFFL];
#pragma unroll
            for (int q = 0; q < NT; ++q)
                sm.Om[lane * NLO + 4 * tb + q] = ov[q] + bv[q];
            sm.Om[lane * NLO + AFFL] = oa + baff;
        }
        if (tail)
        {
            double od[NJ];
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                od[a] = sm.Om[lane * NLO + c.D0 + a];
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                sm.Om[lane * NLO + c.D0 + a] = od[a] + bd[a];
        }
    }
    __syncwarp();
    applyTtx(s, cf, dt);
}

// ---- per-lane sparse table of T_x (forward rollout), SURVEY App. A-3 ---------------------------------------------
constexpr int CQF = 8;
struct CdFwdTab
{
    double cw[CQF];
    int iw[CQF];   // source index: 0..25 state, 26..29 throttle in effect, 30..37 joint increment in effect
    double cc;
    int helper;
};

__device__ void cd_build_fwd(CdFwdTab& t, const double* __restrict__ cf, int lane)
{
#pragma unroll
    for (int q = 0; q < CQF; ++q) { t.cw[q] = 0.0; t.iw[q] = 0; }
    t.cc = 0.0;
    t.helper = -1;
    const double w[3] = {cf[QD_OMEGA], cf[QD_OMEGA + 1], cf[QD_OMEGA + 2]};
    auto mS = [&](int a, int b) -> double {   // -S(w)[a][b]
        if (a == b) return 0.0;
        const int k = 3 - a - b;
        const double sgn = ((b - a + 3) % 3 == 1) ? 1.0 : -1.0;
        return sgn * w[k];
    };
    auto setw = [&](int q, double cv, int iv) {
#pragma unroll
        for (int qq = 0; qq < CQF; ++qq)
            if (qq == q) { t.cw[qq] = cv; t.iw[qq] = iv; }
    };
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

// knot schedule of the two software-pipelined warps
__device__ __forceinline__ void cd_schedule(int t, int N, int kS, int& ta, int& ka, int& tb_, int& kb)
{
    ta = tb_ = TK_NONE;
    ka = kb = 0;
    if (kS < 0)
    {
        if (t < N) { ta = TK_STAGE; ka = N - 1 - t; }
        if (t >= 1 && t <= N) { tb_ = TK_STAGE; kb = N - t; }
        return;
    }
    const int nTail = N - 1 - kS;
    if (t < nTail) { ta = TK_STAGE; ka = N - 1 - t; }
    else if (t == nTail) { ta = TK_PROP; ka = kS; }
    else if (t == nTail + 2) { ta = TK_SCHUR; ka = kS; }
    else if (t >= nTail + 3 && t < nTail + 3 + kS) { ta = TK_STAGE; ka = kS - 1 - (t - nTail - 3); }
    if (t >= 1 && t <= nTail) { tb_ = TK_STAGE; kb = N - t; }
    else if (t == nTail + 1) { tb_ = TK_PROP; kb = kS; }
    else if (t == nTail + 3) { tb_ = TK_SCHUR; kb = kS; }
    else if (t > nTail + 3 && t <= nTail + 3 + kS) { tb_ = TK_STAGE; kb = kS - (t - nTail - 3); }
}

__global__ void __launch_bounds__(CD_THREADS, 8)
qp_condensed_kernel(const __grid_constant__ DeviceConfig cfgv, int B, const double* __restrict__ qd_all,
                    double* __restrict__ ws_all, double* __restrict__ z_all, double* __restrict__ st,
                    double* __restrict__ out_rows, int* __restrict__ status, int* __restrict__ n_factor,
                    int* __restrict__ n_solve, size_t ws_stride, int want_z)
{
    __shared__ CdSmem sm;
    const DeviceConfig& cfg = cfgv;   // kernel parameter space (constant bank)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x;
    const double* qd = qd_all + (size_t)inst * cfg.qd_stride;
    const int N = cfg.N, Nc = cfg.Nc, NC = cfg.NC;
    const int nv = 4 * cfg.nblk;
    const bool held = Nc - 1 < N - 1;
    CdCtx c{cfg, sm, ws_all + (size_t)inst * ws_stride, lane, cfg.nblk >= 3 ? 0 : 16, held ? Nc - 1 : -1};

    PHASE_CLK(0);
    // ---- stage the QP data: coefficients, reference window; finiteness gate --------------------------------------
    bool fin = true;
    for (int e = threadIdx.x; e < cfg.qd_stride; e += CD_THREADS)
    {
        const double v = qd[e];
        fin = fin && isfinite(v);
        if (e < CCF)
            sm.cf[e] = v;
        else if (e >= QD_XREF && e < QD_XREF + 12 * NC)
            sm.xref[e - QD_XREF] = v;
    }
    for (int e = threadIdx.x; e < NLO * NLO; e += CD_THREADS)
        sm.Om[e] = 0.0;
    if (threadIdx.x < NX)
        sm.Qd[threadIdx.x] = cfg.Qd[threadIdx.x];
    if (threadIdx.x < NJ)
        sm.Rqd[threadIdx.x] = cfg.Rqd[threadIdx.x];
    if (threadIdx.x < N)
        sm.dtk[threadIdx.x] = cfg.dt[threadIdx.x];
    for (int e = threadIdx.x; e < 6 * NJ; e += CD_THREADS)
        sm.lam[e] = qd[(e < 3 * NJ ? QD_LLIN : QD_LANG - 3 * NJ) + e];
    const bool all_fin = __syncthreads_and(fin);
    int stat = all_fin ? VSMPC_STATUS_SOLVED : VSMPC_STATUS_NUMERICAL;

    PHASE_CLK(1);
    // ---- factorisation: warp A = P recursion, warp B = parameter columns, one knot apart --------------------------
    double y[NX];   // warp A: row `lane` of P ; warp B: column `lane` of Psi
#pragma unroll
    for (int j = 0; j < NX; ++j)
        y[j] = 0.0;
    const double qd_lane = lane < NX ? sm.Qd[lane] : 0.0;
    bool ok = true;
    long long clkA0 = 0, clkA1 = 0, clkB0 = 0, clkB1 = 0;
    (void)clkA0; (void)clkA1; (void)clkB0; (void)clkB1;
    const int n_it = c.kS < 0 ? N + 1 : (N - 1 - c.kS) + 3 + c.kS + 1;
    for (int t = 0; t < n_it; ++t)
    {
        int ta, ka, tbk, kb;
        cd_schedule(t, N, c.kS, ta, ka, tbk, kb);
        if (warp == 0)
        {
            if (ta != TK_NONE)
            {
                CdSlot& sl = sm.slot[ka & 1];
                double hux[NJ];
                double2 own;
                const bool elim = ta == TK_STAGE && !(held && ka >= Nc - 1);
                long long tclk = clock64();
                (void)tclk;
                if (ta != TK_SCHUR)
                {
                    a_prop(c, ka, elim, y, qd_lane, hux, own);
                    SUBCLK(clkA0, tclk);
                }
                else
                {
                    // Schur step of the held joint block: H_ux = Psi_T[:, d]' (published by warp B),
                    // H_uu = Om_T[d, d] + R
                    const int r = lane & 7, q = lane >> 3;
#pragma unroll
                    for (int m = 0; m < NJ; ++m)
                        hux[m] = lane < NX ? sl.Hux[m * NX + lane] : 0.0;
                    own.x = sm.Om[(c.D0 + r) * NLO + c.D0 + 2 * q] + (2 * q == r ? sm.Rqd[r] : 0.0);
                    own.y = sm.Om[(c.D0 + r) * NLO + c.D0 + 2 * q + 1] + (2 * q + 1 == r ? sm.Rqd[r] : 0.0);
                }
                __syncwarp();
                if (elim || ta == TK_SCHUR)
                    ok = a_eliminate(c, sl, y, hux, own, c.ws + (size_t)ka * WSC_STAGE) && ok;
                SUBCLK(clkA1, tclk);
            }
        }
        else if (tbk != TK_NONE)
        {
            CdSlot& sl = sm.slot[kb & 1];
            const bool tail = held && kb >= Nc - 1;
            const bool isD = lane >= c.D0 && lane < c.D0 + NJ;
            double hut[NJ];
            bool down = false;
            long long tclk = clock64();
            (void)tclk;
            if (tbk != TK_SCHUR)
            {
                b_prop(c, kb, tail, y, hut);
                SUBCLK(clkB0, tclk);
                if (tbk == TK_PROP)
                {
                    // publish H_ux = Psi_T[:, d]' for warp A's Schur step
                    if (isD)
                    {
#pragma unroll
                        for (int j = 0; j < NX; ++j)
                            sl.Hux[(lane - c.D0) * NX + j] = y[j];
                    }
                }
                else
                    down = !tail;
            }
            else
            {
#pragma unroll
                for (int m = 0; m < NJ; ++m)
                    hut[m] = (lane < NLO && !isD) ? sm.Om[(c.D0 + m) * NLO + lane] : 0.0;
                __syncwarp();
                down = true;
            }
            if (down)
            {
                if (lane == AFFL)
                {
#pragma unroll
                    for (int m = 0; m < NJ; ++m)
                        hut[m] += sm.cf[QD_GQ + m];
                }
                const bool schur = tbk == TK_SCHUR;
                b_downdate(c, sl, y, hut, c.ws + (size_t)kb * WSC_STAGE, schur && isD);
                SUBCLK(clkB1, tclk);
                if (schur)
                {
                    if (lane < NLO)
                    {
#pragma unroll
                        for (int a = 0; a < NJ; ++a)
                        {
                            sm.Om[(c.D0 + a) * NLO + lane] = 0.0;
                            sm.Om[lane * NLO + c.D0 + a] = 0.0;
                        }
                    }
                    __syncwarp();
                }
            }
        }
        __syncthreads();
    }
    PHASE_CLK(2);
    PHASE_CLK(4);
#ifdef VSMPC_PHASE_CLOCKS
    if (lane == 0 && inst < 4096)
    {
        if (warp == 0)
        {
            unsigned smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            g_phase_clk[inst][8] = clkA0;
            g_phase_clk[inst][9] = clkA1;
            g_phase_clk[inst][12] = smid;
        }
        else { g_phase_clk[inst][10] = clkB0; g_phase_clk[inst][11] = clkB1; }
    }
#endif
    if (warp == 0 && lane == 0)
        sm.flags[0] = ok ? 0 : 1;
    // Psi_0' x0 while warp B still holds its column, then the deferred down-date of Om on the tensor cores (both
    // warps; the stores of H_utheta / F were made visible by the __syncthreads that closed the recursion)
    double g_psi = 0.0;
    if (warp == 1 && lane < nv)
    {
#pragma unroll
        for (int j = 0; j < NX; ++j)
            g_psi = fma(y[j], sm.cf[QD_X0 + j], g_psi);
    }
    cd_omega_downdate(c.ws, NJ * Nc, sm.Om, warp, lane);
    __syncthreads();

    // ---- warp B: reduced QP in the throttle variables + dual active set ---------------------------------------------
    // after the factorisation the mailbox slots are dead: G (24 x 25) and the working-set inverse live there
    double* G = reinterpret_cast<double*>(&sm.slot[0]);
    double* Minv = G + CD_MAXW * LDG;
    double* as_r = sm.Hut;
    double* as_lam = as_r + CD_MAXW;
    double* as_sgn = as_lam + CD_MAXW;
    int* as_widx = reinterpret_cast<int*>(as_sgn + CD_MAXW);
    const bool pinned = sm.cf[QD_PINNED] != 0.0;
    const int first = pinned ? NT : 0;
    const double lo = sm.cf[QD_VMIN], up = sm.cf[QD_VMAX];
    if (warp == 1)
    {
        // gradient and Hessian row of variable `lane`
        double g = g_psi;
        if (lane < nv)
            g += sm.Om[lane * NLO + AFFL];
        double h[CD_MAXW];
        const int blk = lane >> 2;
#pragma unroll
        for (int j = 0; j < CD_MAXW; ++j)
        {
            double v = (lane < nv && j < nv) ? sm.Om[lane * NLO + j] : ((j == lane % CD_MAXW && lane >= nv) ? 1.0 : 0.0);
            if (lane < nv)
            {
                if (j == lane)
                    v += cfg.w_t * ((blk > 0 ? 1.0 : 0.0) + (blk < cfg.nblk - 1 ? 1.0 : 0.0)) + (blk == 0 ? cfg.w_i : 0.0);
                if ((j == lane - NT && blk > 0) || (j == lane + NT && blk < cfg.nblk - 1))
                    v -= cfg.w_t;
            }
            h[j] = v;
        }
        if (lane < NT)
            g -= cfg.w_i * sm.cf[QD_VBAR + lane];
        if (pinned)
        {
            // block 0 is a parameter: fold it into the gradient and decouple it
            if (lane >= NT && lane < nv)
            {
#pragma unroll
                for (int j = 0; j < NT; ++j)
                    g = fma(h[j], sm.cf[QD_VBAR + j], g);
            }
#pragma unroll
            for (int j = 0; j < CD_MAXW; ++j)
            {
                if (lane < NT)
                    h[j] = (j == lane) ? 1.0 : 0.0;
                else if (j < NT)
                    h[j] = 0.0;
            }
            if (lane < NT)
                g = 0.0;
        }
        if (lane < CD_MAXW)
        {
            as_r[lane] = g;
            double2* gr = reinterpret_cast<double2*>(G + lane * LDG);
#pragma unroll
            for (int j = 0; j < CD_MAXW / 2; ++j)
                gr[j] = make_double2(h[2 * j], h[2 * j + 1]);
        }
        __syncwarp();
        const bool okG = gj24(G, LDG, lane, first, CD_MAXW);
        double v_e = 0.0;
        if (lane < CD_MAXW)
        {
            const double2* gr = reinterpret_cast<const double2*>(G + lane * LDG);
            const double2* g2 = reinterpret_cast<const double2*>(as_r);
#pragma unroll
            for (int j = 0; j < CD_MAXW / 2; ++j)
            {
                const double2 gg = gr[j], rr = g2[j];
                v_e = fma(-gg.x, rr.x, v_e);
                v_e = fma(-gg.y, rr.y, v_e);
            }
        }
        __syncwarp();
        if (!okG)
            stat = VSMPC_STATUS_NUMERICAL;
        // ---- Goldfarb-Idnani dual active set on the boxes, one variable per lane (G symmetric: column = row) ----
        const double tol = 1e-10;
        const bool isvar = lane >= first && lane < nv;
        int wpos_e = -1;
        double lamW = 0.0;
        int nW = 0, iters = 0;
        bool fail = stat != VSMPC_STATUS_SOLVED;
        while (!fail)
        {
            int p_idx;
            const double best = warp_max_nonneg((isvar && wpos_e < 0) ? fmax(fmax(v_e - up, lo - v_e), 0.0) : 0.0, p_idx);
            if (!(best > tol))
                break;
            const double v_p0 = __shfl_sync(0xffffffffu, v_e, p_idx);
            const double s = (v_p0 - up > lo - v_p0) ? 1.0 : -1.0;
            const double bound = s > 0 ? up : lo;
            double lam_p = 0.0;
            while (true)
            {
                if (++iters > 6 * CD_MAXW)
                {
                    stat = VSMPC_STATUS_MAX_ITER;
                    fail = true;
                    break;
                }
                const double gp_e = lane < CD_MAXW ? G[p_idx * LDG + lane] : 0.0; // G[:, p]
                const int widx_a = lane < nW ? as_widx[lane] : 0;
                const double sgn_a = lane < nW ? as_sgn[lane] : 0.0;
                const double gwp_a = sgn_a * s * __shfl_sync(0xffffffffu, gp_e, widx_a);
                // r = Minv gwp: gwp broadcast through shared memory, two FMA chains
                if (lane < nW)
                    as_lam[lane] = gwp_a;
                __syncwarp();
                double r_a = 0.0;
                if (lane < nW)
                {
                    const double* mrow = Minv + lane * CD_MAXW;
                    double r0 = 0.0, r1 = 0.0;
                    int b = 0;
#pragma unroll 2
                    for (; b + 1 < nW; b += 2)
                    {
                        r0 = fma(mrow[b], as_lam[b], r0);
                        r1 = fma(mrow[b + 1], as_lam[b + 1], r1);
                    }
                    if (b < nW)
                        r0 = fma(mrow[b], as_lam[b], r0);
                    r_a = r0 + r1;
                }
                double zsum = (lane < nW) ? r_a * gwp_a : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                    zsum += __shfl_xor_sync(0xffffffffu, zsum, o);
                int drop;
                const double t1 = warp_min_nonneg((lane < nW && r_a > 0.0) ? fmax(lamW, 0.0) / r_a : INFINITY, drop);
                const double gpp = __shfl_sync(0xffffffffu, gp_e, p_idx);
                const double v_p = __shfl_sync(0xffffffffu, v_e, p_idx);
                const double zp = gpp - zsum;
                const double izp = 1.0 / zp;
                const double t2 = (zp > 1e-300) ? (s * v_p - s * bound) * izp : INFINITY;
                const double tt = fmin(t1, t2);
                if (!isfinite(tt))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                if (lane < nW)
                    as_r[lane] = r_a * sgn_a;
                __syncwarp();
                {
                    double zd = s * gp_e, zd1 = 0.0;
                    if (lane < CD_MAXW)
                    {
                        int a = 0;
#pragma unroll 2
                        for (; a + 1 < nW; a += 2)
                        {
                            zd = fma(-as_r[a], G[as_widx[a] * LDG + lane], zd);
                            zd1 = fma(-as_r[a + 1], G[as_widx[a + 1] * LDG + lane], zd1);
                        }
                        if (a < nW)
                            zd = fma(-as_r[a], G[as_widx[a] * LDG + lane], zd);
                    }
                    v_e = fma(-tt, zd + zd1, v_e);
                }
                if (lane < nW)
                    lamW -= tt * r_a;
                lam_p += tt;
                if (t2 <= t1)
                {
                    if (nW >= CD_MAXW)
                    {
                        stat = VSMPC_STATUS_MAX_ITER;
                        fail = true;
                        break;
                    }
                    __syncwarp();
                    if (lane < nW)
                        as_lam[lane] = r_a;
                    __syncwarp();
                    if (lane < nW)
                    {
#pragma unroll 4
                        for (int b = 0; b < nW; ++b)
                            Minv[lane * CD_MAXW + b] = fma(r_a * izp, as_lam[b], Minv[lane * CD_MAXW + b]);
                        Minv[lane * CD_MAXW + nW] = -r_a * izp;
                        Minv[nW * CD_MAXW + lane] = -r_a * izp;
                    }
                    if (lane == nW)
                    {
                        Minv[nW * CD_MAXW + nW] = izp;
                        as_widx[nW] = p_idx;
                        as_sgn[nW] = s;
                        lamW = lam_p;
                    }
                    if (lane == p_idx)
                        wpos_e = nW;
                    nW++;
                    __syncwarp();
                    break;
                }
                {
                    const int last = nW - 1;
                    const int var_d = as_widx[drop], var_l = as_widx[last];
                    const double mdd = Minv[drop * CD_MAXW + drop];
                    __syncwarp();
                    const double f = lane < nW ? Minv[lane * CD_MAXW + drop] / mdd : 0.0;
                    if (lane < nW)
                        as_lam[lane] = Minv[drop * CD_MAXW + lane];
                    __syncwarp();
                    if (lane < nW)
#pragma unroll 4
                        for (int b = 0; b < nW; ++b)
                            Minv[lane * CD_MAXW + b] = fma(-f, as_lam[b], Minv[lane * CD_MAXW + b]);
                    __syncwarp();
                    if (drop != last)
                    {
                        if (lane < nW)
                            as_lam[lane] = Minv[last * CD_MAXW + lane];
                        __syncwarp();
                        if (lane < nW)
                        {
                            Minv[drop * CD_MAXW + lane] = as_lam[lane];
                            Minv[lane * CD_MAXW + drop] = as_lam[lane];
                        }
                        __syncwarp();
                        if (lane == 0)
                        {
                            Minv[drop * CD_MAXW + drop] = as_lam[last];
                            as_widx[drop] = var_l;
                            as_sgn[drop] = as_sgn[last];
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
        // theta*: throttle variables (active ones exactly on their bound), affine 1
        if (wpos_e >= 0)
            v_e = as_sgn[wpos_e] > 0 ? up : lo;
        double th = 0.0;
        if (lane < nv)
            th = (pinned && lane < NT) ? sm.cf[QD_VBAR + lane] : v_e;
        else if (lane == AFFL)
            th = 1.0;
        sm.theta[lane] = th;
        if (lane == 0)
            sm.flags[1] = stat;
        PHASE_CLK(5);
#ifdef VSMPC_PHASE_CLOCKS
        if (lane == 0 && inst < 4096)
            g_phase_clk[inst][7] = iters * 100 + nW;
#endif
    }
    __syncthreads();
    if (sm.flags[0] != 0)
        stat = VSMPC_STATUS_NUMERICAL;
    else if (stat == VSMPC_STATUS_SOLVED)
        stat = sm.flags[1];

    // ---- F_k theta* for every elimination knot (both warps) -----------------------------------------------------------
    double* fth = sm.Mt;              // [Nc][8]
    double* xs = sm.Mt + NJ * 32;     // x (26) then dq in effect (8); Nc <= 32 guaranteed by the launcher
    for (int e = threadIdx.x; e < Nc * NJ; e += CD_THREADS)
    {
        const int k = e >> 3, a = e & 7;
        const double2* fr = reinterpret_cast<const double2*>(c.ws + (size_t)k * WSC_STAGE + WSC_F + a * NL);
        const double2* thv = reinterpret_cast<const double2*>(sm.theta);
        double acc = 0.0;
#pragma unroll
        for (int l = 0; l < NL / 2; ++l)
        {
            const double2 f = fr[l];
            const double2 tv = thv[l];
            acc = fma(f.x, tv.x, acc);
            acc = fma(f.y, tv.y, acc);
        }
        fth[e] = acc;
    }
    __syncthreads();
    PHASE_CLK(6);
    if (warp != 0)
        return;

    // ---- warp A: forward rollout ---------------------------------------------------------------------------------------
    double* z = want_z ? z_all + (size_t)inst * cfg.n_var : nullptr;
    double* o = out_rows + (size_t)inst * VSMPC_OUT_DOUBLES;
    const bool solved = stat == VSMPC_STATUS_SOLVED;
    if (lane == 0)
    {
        status[inst] = stat;
        n_factor[inst] = 1;
        n_solve[inst] = 1;
    }
    if (!solved)
        return; // outputs and the joint accumulator are held (variableSamplingMPC.cpp:91)
    CdFwdTab tab;
    cd_build_fwd(tab, sm.cf, lane);
    double* dqs = xs + NY;
    double x = lane < NX ? sm.cf[QD_X0 + lane] : 0.0;
    if (lane < NX)
        xs[lane] = x;
    if (lane < NJ)
        dqs[lane] = 0.0;
    if (z && lane < NX)
        z[lane] = x;
    __syncwarp();
    const int ka = lane & 7, kq = lane >> 3;   // gain row / quarter of the state handled by this lane
    const int j0 = kq * 7, jn = kq == 3 ? 5 : 7;
    // gain rows are prefetched one knot ahead (their addresses do not depend on the state) into the register set the
    // other knot parity uses, so that no instruction of knot k waits for the loads of knot k+1
    double kA[7], kB[7];
#pragma unroll
    for (int t = 0; t < 7; ++t)
    {
        kA[t] = t < jn ? c.ws[WSC_K + ka * NX + j0 + t] : 0.0;
        kB[t] = 0.0;
    }
    auto knot = [&](int k, double (&kuse)[7], double (&kload)[7]) {
        const double dt = sm.dtk[k];
        const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
        if (k + 1 < Nc)
        {
            const double* __restrict__ Kn = c.ws + (size_t)(k + 1) * WSC_STAGE + WSC_K + ka * NX + j0;
#pragma unroll
            for (int t = 0; t < 7; ++t)
                kload[t] = t < jn ? Kn[t] : 0.0;
        }
        if (lane < NT)
            xs[NX + lane] = sm.theta[4 * tb + lane];
        if (k < Nc)
        {
            double part = 0.0, part2 = 0.0;
#pragma unroll
            for (int t = 0; t < 7; ++t)
            {
                if (t & 1)
                    part2 = fma(kuse[t], xs[j0 + t], part2);
                else
                    part = fma(kuse[t], xs[j0 + t], part);
            }
            part += part2;
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            part += __shfl_xor_sync(0xffffffffu, part, 16);
            const double u = -part - fth[k * NJ + ka];
            if (lane < NJ)
            {
                dqs[lane] = u;
                if (k == 0)
                    o[VSMPC_OUT_DELTA_Q + lane] = u;
                if (z)
                    z[NX * (N + 1) + k * NJ + lane] = u;
            }
        }
        __syncwarp();
        double acc = 0.0;
#pragma unroll
        for (int q = 0; q < CQF; ++q)
        {
            acc = fma(tab.cw[q], xs[tab.iw[q]], acc);   // xs = [x (26) | throttle block in effect (4) | dq in effect (8)]
        }
        const double other = __shfl_sync(0xffffffffu, acc, tab.helper < 0 ? lane : tab.helper);
        if (tab.helper >= 0)
            acc += other;
        acc += tab.cc;
        x = fma(dt, acc, x);
        __syncwarp();
        if (lane < NX)
            xs[lane] = x;
        if (k == 0 && lane >= IX_T && lane < IX_EP)
            o[(lane < IX_TD ? VSMPC_OUT_THRUST - IX_T : VSMPC_OUT_THRUST_DOT - IX_TD) + lane] = x;
        if (k == N - 1 && lane < NX)
            o[VSMPC_OUT_FINAL_STATE + lane] = x;
        if (z && lane < NX)
            z[(k + 1) * NX + lane] = x;
        __syncwarp();
    };
#pragma unroll 1
    for (int k = 0; k < N; k += 2)
    {
        knot(k, kA, kB);
        if (k + 1 < N)
            knot(k + 1, kB, kA);
    }
    // remaining outputs (variableSamplingMPC.cpp:96-108,138-151)
    if (lane < NT)
        o[VSMPC_OUT_THROTTLE + lane] = destd_throttle_qd(sm.cf, sm.theta[lane]);
    if (lane < NJ)
    {
        const double dq = o[VSMPC_OUT_DELTA_Q + lane];
        const double acc = st[(size_t)(ST_QACC + lane) * B + inst] + dq;
        st[(size_t)(ST_QACC + lane) * B + inst] = acc;
        o[VSMPC_OUT_JOINTS_REF + lane] = acc;
    }
    if (z)
    {
        const int base = NX * (N + 1) + Nc * NJ;
        if (lane < nv)
            z[base + lane] = sm.theta[lane];
    }
    PHASE_CLK(3);
}

int condensed_phase_clocks(long long* host, int n)
{
#ifdef VSMPC_PHASE_CLOCKS
    return cudaMemcpyFromSymbol(host, g_phase_clk, sizeof(long long) * 16 * (n < 4096 ? n : 4096)) == cudaSuccess ? 0 : 2;
#else
    (void)host; (void)n;
    return 3;
#endif
}

bool condensed_supported(const DeviceConfig& cfg)
{
    return cfg.nblk >= 1 && 4 * cfg.nblk <= CD_MAXW && cfg.NC <= CD_MAXNC && cfg.Nc <= 32 && cfg.N <= CD_MAXN;
}

size_t condensed_ws_doubles(const DeviceConfig& cfg)
{
    return (size_t)cfg.Nc * WSC_STAGE;
}

cudaError_t launch_qp_condensed(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                                double* ws, double* z, double* st, double* out_rows, int* status, int* n_factor,
                                int* n_solve, int want_z, cudaStream_t s)
{
    qp_condensed_kernel<<<B, CD_THREADS, 0, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status, n_factor, n_solve,
                                                 condensed_ws_doubles(h_cfg), want_z);
    return cudaGetLastError();
}

} // namespace vsmpc
