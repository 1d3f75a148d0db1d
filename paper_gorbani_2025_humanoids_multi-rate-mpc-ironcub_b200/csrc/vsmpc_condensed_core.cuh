// Shared device code of the condensed-throttle Riccati kernels (vsmpc_qp_condensed.cu: <= 6 throttle blocks, one column
// warp; vsmpc_qp_condensed_wide.cu: long horizons, several column warps): the structured application of
// T = I + dt A_c, warp A's P recursion (propagation, 8 x 8 elimination), the knot schedule of the software pipeline and
// the forward rollout.  The shared-memory struct is a template parameter (fields cf, lam, Rqd, dtk, slot, Mt).
#pragma once
#include "vsmpc_common.cuh"

namespace vsmpc
{

constexpr int LDM = NX + 1;   // 27, odd: conflict-free transposition
constexpr int NPD = 13;       // published P'D columns: throttle block (4), affine (1), joint block (8)
constexpr int CCF = 164;      // coefficient block copied from the QP data (QD_RM .. QD_JGT, padded)
constexpr int WSC_K = 0;      // per elimination knot in the workspace: K [8][26] first
constexpr int GJ_LD = 10;     // leading dimension of the 8 x 8 tiles of gj8 in doubles; GJ_LD2 = 5 in double2 units
constexpr int GJ_LD2 = GJ_LD / 2;

struct alignas(16) CdSlot
{
    double Hux[NJ * NX];    // [m][j]
    double Hinv[NJ * GJ_LD]; // [a][m], rows padded to GJ_LD doubles (conflict-free Gauss-Jordan, see gj8)
    double PD[NX * NPD];    // [i][col]
};

enum : int { TK_NONE = 0, TK_STAGE = 1, TK_PROP = 2, TK_SCHUR = 3 };

// mbarriers (shared memory) of the decoupled A -> B pipeline (small batches of the reference-horizon kernel; the long-horizon kernel, where FREE and B2A count every column warp): every lane of the posting
// warp arrives (count 32), the waiting warp polls the phase parity with mbarrier.try_wait (default .acquire.cta / .release.cta
// semantics order the published data).  FULL + s: slot s published (A posts, B waits); FREE + s: slot s consumed (B posts, A
// waits); B2A: warp B's propagation of the held-block knot done (A waits before the Schur step); SCHUR: Schur step published.
constexpr int CD_PIPE_SLOTS = 3;
constexpr int MB_FULL = 0, MB_FREE = CD_PIPE_SLOTS, MB_B2A = 2 * CD_PIPE_SLOTS, MB_SCHUR = 2 * CD_PIPE_SLOTS + 1,
              MB_COUNT = 2 * CD_PIPE_SLOTS + 2;
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_post(unsigned long long* b)
{
    asm volatile("{ .reg .b64 st; mbarrier.arrive.shared::cta.b64 st, [%0]; }" ::"r"((unsigned)__cvta_generic_to_shared(b)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, int parity)
{
    const unsigned a = (unsigned)__cvta_generic_to_shared(b);
    unsigned done;
    do
    {
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(a), "r"(parity)
                     : "memory");
    } while (!done);
}



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
// pivots later the inverse is back in S.  S, T: row-major, leading dimension GJ_LD = 10 doubles, 16-byte aligned: with
// rows 80 bytes apart the eight rows of a quarter-warp's STS.128 fall on eight distinct groups of four banks and the
// column reads S[r][p] on eight distinct bank pairs (ld 8 put rows 0, 2, 4, 6 on the same banks: the stores took 16
// wavefronts instead of 4 and the column reads 8 instead of 2 — 10 % of the kernel's shared-memory traffic, r01h capture).
__device__ __forceinline__ bool gj8(double* __restrict__ S, double* __restrict__ T, double2 own, int lane)
{
    const int r = lane & 7, q = lane >> 3;
    reinterpret_cast<double2*>(S)[r * GJ_LD2 + q] = own;
    __syncwarp();
    bool ok = true;
    double* src = S;
    double* dst = T;
#pragma unroll 2
    for (int p = 0; p < NJ; ++p)
    {
        const double d = src[p * GJ_LD + p];
        const double f = src[r * GJ_LD + p];
        const double2 pr = reinterpret_cast<const double2*>(src)[p * GJ_LD2 + q];
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
        reinterpret_cast<double2*>(dst)[r * GJ_LD2 + q] = own;
        __syncwarp();
        double* t = src;
        src = dst;
        dst = t;
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

template <class SM> struct CdCtxT
{
    const DeviceConfig& cfg;
    SM& sm;
    double* ws;   // per elimination knot: one stage (K first)
    int lane;
    int D0;       // first column of the held joint block during the tail
    int kS;       // knot where the held joint block is eliminated (-1: none)
};

// ---- joint increments held on a bound of their box (optional joint-limit rows; JL builds only) -----------------------
// Joint block k of an instance may have components CLAMPED: cm bit c = at the upper bound, bit 8 + c = at the lower one.
// A clamped component is a constant b_c in the elimination of its block (tools/condensed_model.py, ClampedCondensedQP):
//   * H_uu is inverted on the free components (clamped rows / columns replaced by the identity), so row c of
//     K = H_uu^-1 H_ux and of F = H_uu^-1 H_utheta come out as the RAW rows H_ux[c, :] and H_utheta[c, :] — exactly what the
//     multiplier of the bound needs in the forward pass — and P, Psi, Om skip them;
//   * the free rows see the constants through  H_utheta[m, affine] += sum_c H_uu[m][c] b_c  (hb, published by warp A);
//   * the constants enter the value function through  Psi[:, affine] += sum_c b_c H_ux[c, :]'  and
//     Om[:, affine] += b_c H_utheta[c, :]' (+ transpose).
struct CdClamp
{
    unsigned cm;              // 0: nothing clamped in this block (the only value outside the JL builds)
    const double* jb;         // lower bounds [8], upper bounds [8] of the joint increments (shared memory)
    double* hb;               // [8] out (warp A) / in (warp B): sum_c H_uu[m][c] b_c
    double* uraw;             // [8][8] out: H_uu (with R) of the block, for the multipliers (workspace)
};
constexpr int CD_JLSET_WORDS = 64;   // words per instance of the stored working set (one per joint block; both condensed kernels)
__device__ __forceinline__ unsigned cd_clamped(unsigned cm) { return (cm | (cm >> 8)) & 0xffu; }
__device__ __forceinline__ double cd_bval(const double* __restrict__ jb, unsigned cm, int c)
{
    return ((cm >> c) & 1u) ? jb[NJ + c] : (((cm >> (8 + c)) & 1u) ? jb[c] : 0.0);
}

// ---- warp A --------------------------------------------------------------------------------------------------
// steps i-k of a knot: invert H_uu (own = this lane's pair of it), K = H_uu^-1 H_ux, P <- P - H_ux' K
template <class SM, bool JL = false>
__device__ __forceinline__ bool a_eliminate(const CdCtxT<SM>& c, CdSlot& sl, double (&p)[NX], const double (&hux)[NJ],
                                            double2 own, double* __restrict__ wsk, CdClamp cl = CdClamp{0u, nullptr, nullptr, nullptr})
{
    SM& sm = c.sm;
    const int lane = c.lane;
    unsigned cany = 0u;
    if constexpr (JL)
    {
        if (cl.cm != 0u)
        {
            cany = cd_clamped(cl.cm);
            const int r = lane & 7, q = lane >> 3;
            cl.uraw[r * NJ + 2 * q] = own.x;
            cl.uraw[r * NJ + 2 * q + 1] = own.y;
            double part = own.x * cd_bval(cl.jb, cl.cm, 2 * q) + own.y * cd_bval(cl.jb, cl.cm, 2 * q + 1);
            part += __shfl_xor_sync(0xffffffffu, part, 8);
            part += __shfl_xor_sync(0xffffffffu, part, 16);
            if (lane < NJ)
                cl.hb[lane] = part;
            const bool rc = (cany >> r) & 1u;
            if (rc || ((cany >> (2 * q)) & 1u))
                own.x = (r == 2 * q) ? 1.0 : 0.0;
            if (rc || ((cany >> (2 * q + 1)) & 1u))
                own.y = (r == 2 * q + 1) ? 1.0 : 0.0;
        }
    }
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
                const double2 hh = hi[a * GJ_LD2 + m];
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
            if constexpr (JL)
            {
                if ((cany >> m) & 1u)
                    continue;        // clamped component: its K row holds the raw H_ux row, P does not see it
            }
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
template <class SM>
__device__ __forceinline__ void a_prop(const CdCtxT<SM>& c, int k, CdSlot& sl, bool elim, double (&p)[NX], double qd_lane,
                                       double (&hux)[NJ], double2& own)
{
    SM& sm = c.sm;
    const int lane = c.lane;
    const double* cf = sm.cf;
    const double dt = sm.dtk[k];
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

// ---- per-lane sparse table of T_x (forward rollout), SURVEY App. A-3 ---------------------------------------------
constexpr int CQF = 8;
struct CdFwdTab
{
    double cw[CQF];
    int iw[CQF];   // source index: 0..25 state, 26..29 throttle in effect, 30..37 joint increment in effect
    double cc;
    int helper;
};

static __device__ void cd_build_fwd(CdFwdTab& t, const double* __restrict__ cf, int lane)
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

// warp A: forward rollout with the stored gains (u_k = -K_k x - F_k theta*), outputs (variableSamplingMPC.cpp:88-112)
// and, with z != nullptr, the full primal.  xs: 40 doubles of shared memory (x, throttle block in effect, dq in effect);
// fth: F_k theta* [Nc][8]; theta: throttle variables (4 nblk); stage: doubles per elimination knot in ws (K first);
// ostage: 48 doubles of shared memory — the outputs are staged there and committed at the end, because with the optional
// joint-limit rows on (jl != nullptr: QD_JLO / QD_JHI of this instance) a joint increment outside its box means that the
// minimiser of the problem WITHOUT those rows is not the answer: nothing is committed, the function returns true and the
// caller hands the instance to the fallback kernel, which carries the joint boxes in its active set.
// JL builds (reference-horizon kernel with joint-limit rows): `clamp` holds the working set of the joint boxes per block (CdClamp);
// a clamped increment is set to its bound and its multiplier is read off the raw rows the elimination left in its place
// (H_ux[c, :] x in the gain row, H_utheta[c, :] theta in fth, H_uu[c, :] in uraw_all); the function writes the NEXT working set
// (violated free components join at the bound they crossed, clamped ones with the wrong multiplier sign leave — a primal-dual
// active-set step on all blocks at once) and returns 1 if it differs from the one this pass was factorised with, 0 at the
// fixed point (outputs committed), 2 on a non-finite increment.  Without JL: 1 = a joint box is violated (fallback kernel).
template <class SM, bool JL = false>
__device__ __forceinline__ int cd_forward(const DeviceConfig& cfg, SM& sm, const double* __restrict__ ws, int stage,
                                          const double* __restrict__ theta, const double* __restrict__ fth,
                                          double* __restrict__ xs, int lane, int B, int inst, double* __restrict__ z,
                                          double* __restrict__ o, double* __restrict__ st, double* __restrict__ ostage,
                                          const double* __restrict__ jl, unsigned* __restrict__ clamp = nullptr,
                                          const double* __restrict__ uraw_all = nullptr, unsigned* __restrict__ cand = nullptr,
                                          int mode = 0)
{
    const int N = cfg.N, Nc = cfg.Nc, nv = 4 * cfg.nblk;
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
    const double blo = jl ? jl[ka] : -INFINITY, bhi = jl ? jl[NJ + ka] : INFINITY;
    const double jlo = blo - 1e-9, jhi = bhi + 1e-9;
    bool viol = false, bad = false;
    // gain rows are prefetched one knot ahead (their addresses do not depend on the state) into the register set the
    // other knot parity uses, so that no instruction of knot k waits for the loads of knot k+1
    double kA[7], kB[7];
#pragma unroll
    for (int t = 0; t < 7; ++t)
    {
        kA[t] = t < jn ? ws[WSC_K + ka * NX + j0 + t] : 0.0;
        kB[t] = 0.0;
    }
    auto knot = [&](int k, double (&kuse)[7], double (&kload)[7]) {
        const double dt = sm.dtk[k];
        const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
        if (k + 1 < Nc)
        {
            const double* __restrict__ Kn = ws + (size_t)(k + 1) * stage + WSC_K + ka * NX + j0;
#pragma unroll
            for (int t = 0; t < 7; ++t)
                kload[t] = t < jn ? Kn[t] : 0.0;
        }
        if (lane < NT)
            xs[NX + lane] = theta[4 * tb + lane];
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
            double u = -part - fth[k * NJ + ka];
            unsigned cm = 0u;
            double g0 = 0.0;
            bool cup = false, clo = false;
            if constexpr (JL)
            {
                cm = clamp[k];
                cup = (cm >> ka) & 1u;
                clo = (cm >> (8 + ka)) & 1u;
                if (cup || clo)
                {
                    g0 = -u;                 // H_ux[c, :] x_k + H_utheta[c, :] theta (raw rows in place of K and F)
                    u = cup ? bhi : blo;
                }
                bad = bad || !isfinite(u);
            }
            else
                viol = viol || !(u >= jlo && u <= jhi);
            if (lane < NJ)
            {
                dqs[lane] = u;
                if (k == 0)
                    ostage[VSMPC_OUT_DELTA_Q + lane] = u;
                if (z)
                    z[NX * (N + 1) + k * NJ + lane] = u;
            }
            if constexpr (JL)
            {
                __syncwarp();
                double grad = g0, gmag = fabs(g0);
                if (cm != 0u && (cup || clo) && lane < NJ)
                {
                    const double* __restrict__ ur = uraw_all + k * (NJ * NJ) + ka * NJ;
#pragma unroll
                    for (int m = 0; m < NJ; ++m)
                    {
                        const double t = ur[m] * dqs[m];
                        grad += t;
                        gmag += fabs(t);
                    }
                }
                // candidates of this block: free increments outside their box (beyond 1e-9 rad) join at the bound they crossed;
                // clamped ones leave when their multiplier (upper bound: -grad, lower bound: +grad) is negative beyond the
                // rounding of its own terms — without that margin a degenerate increment (on its bound with a zero multiplier)
                // is released on noise, comes back 1e-9 outside, and the working set never settles
                const bool addu = !(cup || clo) && u > jhi, addl = !(cup || clo) && u < jlo;
                const bool rel = (cup && grad > 1e-9 * gmag) || (clo && grad < -1e-9 * gmag);
                const unsigned ad = (__ballot_sync(0xffffffffu, addu && lane < NJ) & 0xffu) |
                                    ((__ballot_sync(0xffffffffu, addl && lane < NJ) & 0xffu) << 8);
                const unsigned rl = __ballot_sync(0xffffffffu, rel && lane < NJ) & 0xffu;
                if (lane == 0)
                    cand[k] = ad | (rl << 16);
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
            ostage[(lane < IX_TD ? VSMPC_OUT_THRUST - IX_T : VSMPC_OUT_THRUST_DOT - IX_TD) + lane] = x;
        if (k == N - 1 && lane < NX)
            ostage[VSMPC_OUT_FINAL_STATE + lane] = x;
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
    if (__any_sync(0xffffffffu, bad))
        return 2;
    if constexpr (JL)
    {
        // next working set, all blocks at once (lane = block).  mode 0 (first passes): joins and leaves together (two to four
        // passes on the test workloads).  mode 2 (last pass of the first group): the same, but a set that still moves is
        // EMPTIED — a warm start from an unrelated working set can cycle where the cold start settles.  mode 1 (late passes):
        // leaves only once no free increment is outside its box — the joint add / release of the plain iteration can cycle
        // on general data
        const bool careful = mode == 1;
        __syncwarp();
        const int ncr = (Nc + 31) & ~31;
        bool any_add = false;
        for (int k0 = lane; k0 < ncr; k0 += 32)
            any_add = __any_sync(0xffffffffu, k0 < Nc && (cand[k0 < Nc ? k0 : 0] & 0xffffu) != 0u) || any_add;
        viol = false;
        for (int k0 = lane; k0 < Nc; k0 += 32)
        {
            const unsigned old = clamp[k0], cd = cand[k0];
            const unsigned ad = cd & 0xffffu, rl = (cd >> 16) & 0xffu;
            unsigned nm = old | ad;
            if (!(careful && any_add))
                nm &= ~(rl | (rl << 8));
            viol = viol || nm != old;
            clamp[k0] = nm;
        }
        if (mode == 2 && __any_sync(0xffffffffu, viol))
        {
            for (int k0 = lane; k0 < Nc; k0 += 32)
                clamp[k0] = 0u;
        }
        __syncwarp();
    }
    if (__any_sync(0xffffffffu, viol))
        return 1;     // a joint increment left its box / the working set moved: nothing committed
    // remaining outputs (variableSamplingMPC.cpp:96-108,138-151), then the commit of the staged ones
    __syncwarp();
    for (int e = lane; e < VSMPC_OUT_JOINTS_REF; e += 32)
        if (e < VSMPC_OUT_THROTTLE || e >= VSMPC_OUT_THRUST)
            o[e] = ostage[e];
    if (lane < NT)
        o[VSMPC_OUT_THROTTLE + lane] = destd_throttle_qd(sm.cf, theta[lane]);
    if (lane < NJ)
    {
        const double dq = ostage[VSMPC_OUT_DELTA_Q + lane];
        const double acc = st[(size_t)(ST_QACC + lane) * B + inst] + dq;
        st[(size_t)(ST_QACC + lane) * B + inst] = acc;
        o[VSMPC_OUT_JOINTS_REF + lane] = acc;
    }
    if (z)
    {
        const int base = NX * (N + 1) + Nc * NJ;
        for (int e = lane; e < nv; e += 32)
            z[base + e] = theta[e];
    }
    return 0;
}

// Copy of an instance's output row and status into the staging buffers of the asynchronous read-back (one warp; the row was
// written — or is being held — by this warp).  vsmpc_get_output_async then reads the staging buffer back without a snapshot
// kernel between the QP kernel and the next tick.
__device__ __forceinline__ void cd_stage_outputs(const double* o, const int* status, int inst, int lane,
                                                 double* __restrict__ out2, int* __restrict__ status2)
{
    if (!out2)
        return;
    __syncwarp();
    double* o2 = out2 + (size_t)inst * VSMPC_OUT_DOUBLES;
    for (int e = lane; e < VSMPC_OUT_DOUBLES; e += 32)
        o2[e] = o[e];
    if (lane == 0)
        status2[inst] = status[inst];
}

} // namespace vsmpc
