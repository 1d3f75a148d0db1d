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
// * optional joint-limit rows (JointPositionConstraint, constraintsVSMPC.cpp:388-468; JL build): a working set on the boxes of
//   the joint increments is carried through the eliminations (CdClamp in vsmpc_condensed_core.cuh) and a pass loop around
//   the solve re-factorises until the forward pass confirms it (ClampedCondensedQP in tools/condensed_model.py).
#include <cstdlib>

#include "vsmpc_condensed_core.cuh"

namespace vsmpc
{

constexpr int CD_THREADS = 64;
constexpr int NL = 32;        // parameter columns (lanes of warp B)
constexpr int NLO = 25;       // rows/columns of Om in use: 24 throttle variables + affine (held block aliases lanes)
constexpr int AFFL = 24;      // lane of the affine column
constexpr int LDH = 9;        // leading dimension of the F rows parked in shared memory (odd: conflict-free per lane)
constexpr int CD_MAXNC = 16;  // reference columns kept in shared memory
constexpr int CD_MAXW = 24;
constexpr int WSC_F = NJ * NX;        //                                     then F [8][32]
constexpr int WSC_H = NJ * NX + NJ * NL; //                                     then H_utheta [8][32]
constexpr int WSC_STAGE = NJ * NX + 2 * NJ * NL; // 720

constexpr int CD_MAXN = 32;   // knots kept in shared memory (dt grid)

// working set of the optional joint boxes (JL build only: the other builds keep their shared-memory footprint)
template <bool JL> struct CdJlSmem
{
    unsigned clamp[CD_MAXN];    // per joint block: bit c = increment c held at its upper bound, bit 8 + c at its lower bound
    unsigned cand[CD_MAXN];     // forward pass: increments that would join (bits 0-15, like clamp) / leave (bits 16-23)
    double jb[2 * NJ];          // bounds of the increments: lower [8], upper [8]  (QD_JLO / QD_JHI)
    double hb[2][NJ];           // warp A -> warp B, per mailbox slot: sum_c H_uu[m][c] b_c
};
template <> struct CdJlSmem<false>
{
};
constexpr int CD_JL_PASSES = 16;  // factorisations per solve before the instance is handed to the fallback kernel
constexpr int CD_JL_PLAIN = 5;    // passes 0..4 from the stored working set and 5..9 from the empty one apply joins and leaves together;
                                  // from pass 10 leaves wait for primal feasibility (cd_forward)
__device__ __forceinline__ int cd_jl_mode(int pass, int plain)
{
    return pass == plain - 1 ? 2 : (pass < 2 * plain ? 0 : 1);
}
constexpr int WSC_U = NJ * NJ;    // per joint block, behind the Nc stages of the workspace: raw H_uu [8][8]

template <int NSLOT, bool JL = false> struct alignas(16) CdSmemT : CdJlSmem<JL>
{
    double cf[CCF];
    alignas(16) double lam[6 * NJ];         // dt-free B_J rows: [q][a], q = 0..2 linear, 3..5 angular momentum
    double Qd[NX];
    double Rqd[NJ];
    double dtk[CD_MAXN];
    double xref[12 * CD_MAXNC];
    CdSlot slot[NSLOT];         // A -> B mailbox: slot = knot & 1 (lock step) or publication j % 3 (decoupled pipeline)
    alignas(16) double Mt[NX * LDM];        // warp A: transposition buffer, then the gain rows K [8][26] of the knot in
                                // flight; after the factorisation: F theta, x, dq
    double Om[NLO * NLO];
    alignas(16) double Hut[4 * CD_MAXW];    // active-set vectors (r, lambda, sign, index) after the factorisation
    alignas(16) double theta[NL];
    unsigned long long mbar[MB_COUNT];
    int flags[4];
};

#ifdef VSMPC_PHASE_CLOCKS
__device__ long long g_phase_clk[4096][16];
#define SUBCLK(acc, t0) do { const long long t1__ = clock64(); acc += t1__ - t0; t0 = t1__; } while (0)
#define PHASE_CLK(slot) do { if (lane == 0 && warp == (slot >= 4 ? 1 : 0) && inst < 4096) g_phase_clk[inst][slot] = clock64(); } while (0)
#else
#define PHASE_CLK(slot) do { } while (0)
#define SUBCLK(acc, t0) do { } while (0)
#endif

// lane `mine` publishes its row of T (registers) in shared memory for the warp
__device__ __forceinline__ void rp_publish(const double (&t)[CD_MAXW], double* __restrict__ rowq, bool mine)
{
    if (mine)
    {
        double2* w2 = reinterpret_cast<double2*>(rowq);
#pragma unroll
        for (int j = 0; j < CD_MAXW / 2; ++j)
            w2[j] = make_double2(t[2 * j], t[2 * j + 1]);
    }
    __syncwarp();
}

// one exchange pivot on index q (row q published in rowq): lane l < 24 owns row l of T in registers;
// sgn = +1 if lane l is in the same class as q (both exchanged or both not), -1 otherwise: T[l][q] = sgn T[q][l].
// The pivot is applied as the rank-1 change  T' = T - (T[:, q] + e_q)(T[q, :] - e_q)' / d,  d = T[q][q]  — it gives
// T'[q][q] = 1/d, T'[q][j] = -T[q][j]/d, T'[l][q] = T[l][q]/d and the Schur update of the rest in ONE uniform update of every
// row: lane q only lowers its published diagonal by one (shared memory, indexed by q at run time, which registers cannot
// be) and uses the multiplier (d + 1)/d.  The earlier form wrote row q / column q by case distinction, which cost 24
// predicated register writes per pivot — three instructions each, more than the 24 FMAs of the update itself.  The rank-1
// form is exact up to eps x max(d, 1/d): H_r is equilibrated to a unit diagonal first (the long-horizon kernel does the same).
// Tried and rejected (profiles/r02_k2_experiments.md): two indices per step (block pivot; same time — the phase is bound by
// instructions issued per index, not by round trips), row q by warp shuffles with a warp-uniform switch for the
// run-time register index (slower: the switches compile to predicated chains).
__device__ __forceinline__ bool rp_pivot(double (&t)[CD_MAXW], double* __restrict__ rowq, int q, int lane, double sgn)
{
    if (lane == q)
        rowq[q] -= 1.0;
    __syncwarp();
    const double dm1 = rowq[q];
    const double d = dm1 + 1.0;
    const double dinv = __drcp_rn(d);
    const double f = (lane == q) ? (d + 1.0) * dinv : sgn * rowq[lane < CD_MAXW ? lane : 0] * dinv;
    const double2* r2 = reinterpret_cast<const double2*>(rowq);
#pragma unroll
    for (int j = 0; j < CD_MAXW / 2; ++j)
    {
        const double2 rr = r2[j];
        t[2 * j] = fma(-f, rr.x, t[2 * j]);
        t[2 * j + 1] = fma(-f, rr.y, t[2 * j + 1]);
    }
    __syncwarp();
    return (d > 0.0) && (d < 1e300);
}

// ---- warp B --------------------------------------------------------------------------------------------------
// down-date of the parameter columns with the eliminated block: F = H_uu^-1 H_utheta, Psi -= H_ux' F.  The matching
// down-date of Om (Om -= H_utheta' F) is NOT done knot by knot: H_utheta and F of every elimination knot are stacked
// in the workspace and contracted once after the recursion on the FP64 tensor cores (cd_omega_downdate).
// cany (JL build): clamped components of the block — their F rows come out as the raw H_utheta rows (identity rows of the
// masked inverse; the forward pass reads the multiplier off them), their H_utheta rows are stacked as zeros (no share in the
// Om down-date) and Psi skips them.
template <bool JL = false, class CdCtx>
__device__ __forceinline__ void b_downdate(const CdCtx& c, CdSlot& sl, double (&s)[NX], const double (&hut)[NJ],
                                           double* __restrict__ wsk, bool clear_col, unsigned cany = 0u)
{
    const int lane = c.lane;
    double* Fs = sl.PD;   // the P'D columns of this knot were consumed by b_prop: reuse as F [l][LDH]
#pragma unroll
    for (int m = 0; m < NJ; ++m)
        wsk[WSC_H + m * NL + lane] = (JL && ((cany >> m) & 1u)) ? 0.0 : hut[m];
    {
        const double2* hi = reinterpret_cast<const double2*>(sl.Hinv);
#pragma unroll 1
        for (int a = 0; a < NJ; ++a)
        {
            double v0 = 0.0, v1 = 0.0;
#pragma unroll
            for (int m = 0; m < NJ / 2; ++m)
            {
                const double2 hh = hi[a * GJ_LD2 + m];
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
            if constexpr (JL)
            {
                if ((cany >> m) & 1u)
                    continue;
            }
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
template <class CdCtx>
__device__ __forceinline__ void b_prop(const CdCtx& c, int k, CdSlot& sl, bool tail, double (&s)[NX], double (&bj2)[NJ])
{
    const DeviceConfig& cfg = c.cfg;
    auto& sm = c.sm;
    const int lane = c.lane;
    const double* cf = sm.cf;
    const double dt = sm.dtk[k];
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

// PIPE = false: the two warps advance in lock step, one __syncthreads per knot (the throughput configuration: 25 KB of shared
// memory, eight CTAs per SM).  PIPE = true: decoupled pipeline for SMALL batches (at most four CTAs per SM, 232 registers: no spills) — warp A publishes
// knot N-1-j in slot j % 3 of a mailbox ring and may run two knots ahead, warp B consumes in order and hands the slot back
// (mbarriers); the one place where warp A needs warp B is the Schur step of the held joint block.  With one CTA per SM warp A
// spends a quarter of the recursion waiting at the common barrier (132 k cycles against 99 k busy); without contention that
// wait is pure latency of a single solve.  At B = 1024 (seven CTAs per SM) the same pipeline gained nothing and its 30 KB of
// shared memory cost the eighth CTA per SM (profiles/r02_k2_experiments.md), so the launcher uses it for small batches only.
template <bool PIPE, bool JL>
__device__ __forceinline__ void
qp_condensed_body(const DeviceConfig& cfgv, int B, const double* __restrict__ qd_all,
                  double* __restrict__ ws_all, double* __restrict__ z_all, double* __restrict__ st,
                  double* __restrict__ out_rows, int* __restrict__ status, int* __restrict__ n_factor,
                  int* __restrict__ n_solve, int* __restrict__ n_pivot, size_t ws_stride, int want_z,
                  int* __restrict__ fb_list, int* __restrict__ fb_count, int fb_mode, double* __restrict__ out2,
                  int* __restrict__ status2, unsigned* __restrict__ jlset)
{
    static_assert(!(PIPE && JL), "the joint-box working set is built into the lock-step loop only");
    using CdSmem = CdSmemT<PIPE ? CD_PIPE_SLOTS : 2, JL>;
    using CdCtx = CdCtxT<CdSmem>;
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
    // (three loads in flight per thread: the block is read once, a load per trip would cost an L2 round trip each)
    for (int e0 = threadIdx.x; e0 < cfg.qd_stride; e0 += 3 * CD_THREADS)
    {
        double v[3];
#pragma unroll
        for (int u = 0; u < 3; ++u)
            v[u] = e0 + u * CD_THREADS < cfg.qd_stride ? qd[e0 + u * CD_THREADS] : 0.0;
#pragma unroll
        for (int u = 0; u < 3; ++u)
        {
            const int e = e0 + u * CD_THREADS;
            fin = fin && isfinite(v[u]);
            if (e < CCF)
                sm.cf[e] = v[u];
            else if (e >= QD_XREF && e < QD_XREF + 12 * NC)
                sm.xref[e - QD_XREF] = v[u];
        }
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
    if (PIPE && threadIdx.x < MB_COUNT)
        mbar_init(&sm.mbar[threadIdx.x], 32);
    if constexpr (JL)
    {
        if (threadIdx.x < 2 * NJ)
            sm.jb[threadIdx.x] = qd[QD_JLO + threadIdx.x];
        // warm start: the working set the last solve of this instance ended with (zeros after configure)
        if (threadIdx.x < CD_MAXN)
            sm.clamp[threadIdx.x] = (jlset && threadIdx.x < Nc) ? jlset[(size_t)inst * CD_JLSET_WORDS + threadIdx.x] : 0u;
    }
    const bool all_fin = __syncthreads_and(fin);
    // JL build: primal-dual active set on the joint boxes around the whole solve — pass p factorises with the working set pass
    // p - 1 left in sm.clamp (empty at pass 0), the forward pass writes the next one; a solve without an active joint bound is
    // one pass, like the other builds
    for (int pass = 0;; ++pass)
    {
    int stat = all_fin ? VSMPC_STATUS_SOLVED : VSMPC_STATUS_NUMERICAL;
    if constexpr (JL)
    {
        if (pass > 0)
        {
            for (int e = threadIdx.x; e < NLO * NLO; e += CD_THREADS)
                sm.Om[e] = 0.0;
            __syncthreads();
        }
    }

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
    if constexpr (PIPE)
    {
        if (warp == 0)
        {
#pragma unroll 1
            for (int j = 0; j < N; ++j)
            {
                const int ka = N - 1 - j, s = j % CD_PIPE_SLOTS;
                CdSlot& sl = sm.slot[s];
                double hux[NJ];
                double2 own;
                const bool elim = !(held && ka >= Nc - 1);
                if (j >= CD_PIPE_SLOTS)
                    mbar_wait(&sm.mbar[MB_FREE + s], (j / CD_PIPE_SLOTS - 1) & 1);
                long long tclk = clock64();
                (void)tclk;
                a_prop(c, ka, sl, elim, y, qd_lane, hux, own);
                SUBCLK(clkA0, tclk);
                __syncwarp();
                if (elim)
                    ok = a_eliminate(c, sl, y, hux, own, c.ws + (size_t)ka * WSC_STAGE) && ok;
                SUBCLK(clkA1, tclk);
                mbar_post(&sm.mbar[MB_FULL + s]);
                if (ka == c.kS)
                {
                    // Schur step of the held joint block: H_ux = Psi_T[:, d]' (published by warp B in this slot),
                    // H_uu = Om_T[d, d] + R
                    mbar_wait(&sm.mbar[MB_B2A], 0);
                    tclk = clock64();
                    const int r = lane & 7, q = lane >> 3;
#pragma unroll
                    for (int m = 0; m < NJ; ++m)
                        hux[m] = lane < NX ? sl.Hux[m * NX + lane] : 0.0;
                    own.x = sm.Om[(c.D0 + r) * NLO + c.D0 + 2 * q] + (2 * q == r ? sm.Rqd[r] : 0.0);
                    own.y = sm.Om[(c.D0 + r) * NLO + c.D0 + 2 * q + 1] + (2 * q + 1 == r ? sm.Rqd[r] : 0.0);
                    __syncwarp();
                    ok = a_eliminate(c, sl, y, hux, own, c.ws + (size_t)ka * WSC_STAGE) && ok;
                    SUBCLK(clkA1, tclk);
                    mbar_post(&sm.mbar[MB_SCHUR]);
                }
            }
        }
        else
        {
#pragma unroll 1
            for (int j = 0; j < N; ++j)
            {
                const int kb = N - 1 - j, s = j % CD_PIPE_SLOTS;
                CdSlot& sl = sm.slot[s];
                const bool tail = held && kb >= Nc - 1;
                const bool schur = kb == c.kS;
                const bool isD = lane >= c.D0 && lane < c.D0 + NJ;
                double hut[NJ];
                mbar_wait(&sm.mbar[MB_FULL + s], (j / CD_PIPE_SLOTS) & 1);
                long long tclk = clock64();
                (void)tclk;
                b_prop(c, kb, sl, tail, y, hut);
                SUBCLK(clkB0, tclk);
                if (schur)
                {
                    // publish H_ux = Psi_T[:, d]' for warp A's Schur step, wait for its H_uu^-1
                    if (isD)
                    {
#pragma unroll
                        for (int jj = 0; jj < NX; ++jj)
                            sl.Hux[(lane - c.D0) * NX + jj] = y[jj];
                    }
                    mbar_post(&sm.mbar[MB_B2A]);
                    mbar_wait(&sm.mbar[MB_SCHUR], 0);
                    tclk = clock64();
#pragma unroll
                    for (int m = 0; m < NJ; ++m)
                        hut[m] = (lane < NLO && !isD) ? sm.Om[(c.D0 + m) * NLO + lane] : 0.0;
                    __syncwarp();
                }
                if (schur || !tail)
                {
                    if (lane == AFFL)
                    {
#pragma unroll
                        for (int m = 0; m < NJ; ++m)
                            hut[m] += sm.cf[QD_GQ + m];
                    }
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
                mbar_post(&sm.mbar[MB_FREE + s]);
            }
        }
        __syncthreads();
    }
    else
    {
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
                        a_prop(c, ka, sl, elim, y, qd_lane, hux, own);
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
                    {
                        CdClamp cl{0u, nullptr, nullptr, nullptr};
                        if constexpr (JL)
                            cl = CdClamp{sm.clamp[ka], sm.jb, sm.hb[ka & 1], c.ws + (size_t)Nc * WSC_STAGE + ka * WSC_U};
                        ok = a_eliminate<CdSmem, JL>(c, sl, y, hux, own, c.ws + (size_t)ka * WSC_STAGE, cl) && ok;
                    }
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
                    b_prop(c, kb, sl, tail, y, hut);
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
                    unsigned cany = 0u;
                    if constexpr (JL)
                    {
                        const unsigned cm = sm.clamp[kb];
                        if (cm != 0u)
                        {
                            cany = cd_clamped(cm);
                            // the constants of the clamped components in the value function and in the free rows (CdClamp)
#pragma unroll
                            for (int cc = 0; cc < NJ; ++cc)
                            {
                                if (!((cany >> cc) & 1u))
                                    continue;
                                const double bc = cd_bval(sm.jb, cm, cc);
                                if (lane < NLO)
                                {
                                    const double v = bc * hut[cc];
                                    sm.Om[lane * NLO + AFFL] += v;
                                    sm.Om[AFFL * NLO + lane] += v;
                                }
                                if (lane == AFFL)
                                {
#pragma unroll
                                    for (int j = 0; j < NX; ++j)
                                        y[j] = fma(bc, sl.Hux[cc * NX + j], y[j]);
                                }
                            }
                            if (lane == AFFL)
                            {
#pragma unroll
                                for (int m = 0; m < NJ; ++m)
                                    if (!((cany >> m) & 1u))
                                        hut[m] += sm.hb[kb & 1][m];
                            }
                            __syncwarp();
                        }
                    }
                    b_downdate<JL>(c, sl, y, hut, c.ws + (size_t)kb * WSC_STAGE, schur && isD, cany);
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
    PHASE_CLK(13);

    // ---- warp B: reduced QP in the throttle variables + dual active set ---------------------------------------------
    double* gvec = sm.Hut;                 // gradient of the reduced QP
    double* rowq = sm.Hut + CD_MAXW;       // the published pivot row (16-byte aligned)
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
        double hd = 1.0;       // diagonal entry of this lane's row
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
                if (j == lane)
                    hd = v;
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
        // equilibration to a unit diagonal, v = S v~ with S = diag(H_r)^-1/2 (pinned block and unused lanes: 1): the pivots
        // stay O(1), which the rank-1 pivot form needs; bounds and gradient of the scaled variable per lane
        const bool isvar = lane >= first && lane < nv;
        const double S_e = (isvar && hd > 0.0 && hd < 1e300) ? rsqrt(hd) : 1.0;
        const double iS_e = 1.0 / S_e;
        double* ssm = sm.Hut + 2 * CD_MAXW;
        if (lane < CD_MAXW)
            ssm[lane] = S_e;
        __syncwarp();
        {
            const double2* s2 = reinterpret_cast<const double2*>(ssm);
#pragma unroll
            for (int j = 0; j < CD_MAXW / 2; ++j)
            {
                const double2 sj = s2[j];
                h[2 * j] *= S_e * sj.x;
                h[2 * j + 1] *= S_e * sj.y;
            }
        }
        g *= S_e;
        const double lo_e = lo * iS_e, up_e = up * iS_e;
        if (lane < CD_MAXW)
            gvec[lane] = g;
        __syncwarp();
        // ---- inverse of H_r by exchange pivots: lane l keeps row l of T in registers (h[]) -------------------------
        // In (outputs) = T (inputs) a pivot on q swaps input q and output q of y = H_r v; after all of them T = H_r^-1.
        // T is symmetric within the exchanged set and within the rest, antisymmetric across: T[l][q] = +- T[q][l], so the
        // published row q gives every lane its entry of column q without indexing its registers by q.
        bool okG = __all_sync(0xffffffffu, isvar ? (hd > 0.0 && hd < 1e300) : true);
#pragma unroll 1
        for (int p = first; p < nv; ++p)
        {
            rp_publish(h, rowq, lane == p);
            // exchanged already: first .. p - 1
            okG = rp_pivot(h, rowq, p, lane, (lane >= first && lane < p) ? -1.0 : 1.0) && okG;
        }
        PHASE_CLK(14);
        double v_e = 0.0;
        if (lane < CD_MAXW)
        {
            const double2* g2 = reinterpret_cast<const double2*>(gvec);
            double v1 = 0.0;
#pragma unroll
            for (int j = 0; j < CD_MAXW / 2; ++j)
            {
                const double2 rr = g2[j];
                v_e = fma(-h[2 * j], rr.x, v_e);
                v1 = fma(-h[2 * j + 1], rr.y, v1);
            }
            v_e += v1;
        }
        if (!okG)
            stat = VSMPC_STATUS_NUMERICAL;
        // ---- Goldfarb-Idnani dual active set on the boxes, one variable per lane ---------------------------------------
        // T stays the principal pivot transform of H_r over the free set: activating a bound / dropping it is one more
        // pivot on that index.  For a violated free p with sign s, raising its multiplier by t moves v_F by -t s T[F, p]
        // and lambda_a by -t r_a, r_a = -s_a s T[a, p]; T[p][p] is the step denominator.  No working-set inverse.
        const double tol = 1e-10;
        int act = 0;           // 0 free, +1 / -1 active at the upper / lower bound
        double lam_e = 0.0;
        int iters = 0;
        bool fail = stat != VSMPC_STATUS_SOLVED;
        while (!fail)
        {
            int p_idx;
            // violation measured in the unscaled variable; the iteration itself runs on the scaled box QP (bounds per lane)
            const double best = warp_max_nonneg((isvar && act == 0) ? S_e * fmax(fmax(v_e - up_e, lo_e - v_e), 0.0) : 0.0, p_idx);
            if (!(best > tol))
                break;
            const double s = __shfl_sync(0xffffffffu, (v_e - up_e > lo_e - v_e) ? 1.0 : -1.0, p_idx);
            const double bound = __shfl_sync(0xffffffffu, (v_e - up_e > lo_e - v_e) ? up_e : lo_e, p_idx);
            double lam_p = 0.0;
            while (true)
            {
                if (++iters > 6 * CD_MAXW)
                {
                    stat = VSMPC_STATUS_MAX_ITER;
                    fail = true;
                    break;
                }
                rp_publish(h, rowq, lane == p_idx);
                const double rp = rowq[lane < CD_MAXW ? lane : 0];
                const double zp = rowq[p_idx];
                const double c_e = isvar ? (act == 0 ? rp : -rp) : 0.0;      // T[e][p], p free
                const double r_e = act != 0 ? -(double)act * s * c_e : 0.0;
                int drop;
                const double t1 = warp_min_nonneg((act != 0 && r_e > 0.0) ? fmax(lam_e, 0.0) / r_e : INFINITY, drop);
                const double v_p = __shfl_sync(0xffffffffu, v_e, p_idx);
                const double t2 = (zp > 1e-300) ? (s * v_p - s * bound) / zp : INFINITY;
                const double tt = fmin(t1, t2);
                if (!isfinite(tt) || !isfinite(zp))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                if (isvar)
                {
                    if (act == 0)
                        v_e = fma(-tt * s, c_e, v_e);
                    else
                        lam_e -= tt * r_e;
                }
                lam_p += tt;
                if (t2 <= t1)
                {
                    // full step: p becomes active (row p is published already)
                    rp_pivot(h, rowq, p_idx, lane, act == 0 ? 1.0 : -1.0);
                    if (lane == p_idx)
                    {
                        act = s > 0 ? 1 : -1;
                        lam_e = lam_p;
                        v_e = bound;
                    }
                    __syncwarp();
                    break;
                }
                // blocked step: the blocking bound leaves the working set
                __syncwarp();
                rp_publish(h, rowq, lane == drop);
                if (!rp_pivot(h, rowq, drop, lane, act != 0 ? 1.0 : -1.0))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                if (lane == drop)
                {
                    act = 0;
                    lam_e = 0.0;
                }
                __syncwarp();
            }
        }
        // theta*: throttle variables back in their own scale (active ones exactly on their bound), affine 1
        v_e = act != 0 ? (act > 0 ? up : lo) : S_e * v_e;
        // a NaN iterate never shows up as a violated bound: gate it here (the status holds the outputs)
        if (stat == VSMPC_STATUS_SOLVED && __any_sync(0xffffffffu, lane < nv && !isfinite(v_e)))
            stat = VSMPC_STATUS_NUMERICAL;
        double th = 0.0;
        if (lane < nv)
            th = (pinned && lane < NT) ? sm.cf[QD_VBAR + lane] : v_e;
        else if (lane == AFFL)
            th = 1.0;
        sm.theta[lane] = th;
        if (lane == 0)
        {
            sm.flags[1] = stat;
            sm.flags[2] = (nv - first) + iters;   // exchange pivots executed: inverse + one per active-set iteration
        }
        PHASE_CLK(5);
#ifdef VSMPC_PHASE_CLOCKS
        {
            const int n_act = __popc(__ballot_sync(0xffffffffu, act != 0));     // all lanes: not under the lane-0 test
            if (lane == 0 && inst < 4096)
                g_phase_clk[inst][7] = iters * 100 + n_act;
        }
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
    if constexpr (!JL)
    {
        if (warp != 0)
            return;
    }

    // ---- warp A: forward rollout ---------------------------------------------------------------------------------------
    int again = 0;
    if (warp == 0)
    {
        double* z = want_z ? z_all + (size_t)inst * cfg.n_var : nullptr;
        double* o = out_rows + (size_t)inst * VSMPC_OUT_DOUBLES;
        if (fb_mode == 2 && all_fin)
            stat = VSMPC_STATUS_NUMERICAL;   // test hook: every instance goes through the fallback kernel
        const bool solved = stat == VSMPC_STATUS_SOLVED;
        int fwd = 0;
        if (solved)
        {
            const double* jl = qd[QD_JLIM] != 0.0 ? qd + QD_JLO : nullptr;     // optional joint-limit rows
            if constexpr (JL)
            {
                fwd = cd_forward<CdSmem, true>(cfg, sm, c.ws, WSC_STAGE, sm.theta, fth, xs, lane, B, inst, z, o, st, sm.Mt + 304, jl,
                                               sm.clamp, c.ws + (size_t)Nc * WSC_STAGE, sm.cand, cd_jl_mode(pass, CD_JL_PLAIN));
                // the working set of the joint boxes moved: factorise again with it
                again = (fwd == 1 && pass + 1 < CD_JL_PASSES) ? 1 : 0;
            }
            else
                fwd = cd_forward(cfg, sm, c.ws, WSC_STAGE, sm.theta, fth, xs, lane, B, inst, z, o, st, sm.Mt + 304, jl);
        }
        if (!again)
        {
            if (lane == 0)
            {
                // handed to the pivoted-LU kernel that runs behind this one (vsmpc_qp_fallback.cu): the recursion broke down on
                // finite data (expanding open-loop dynamics), or a joint box is active at the minimiser and this build does not
                // carry the joint boxes (or its working set did not settle); until the fallback succeeds the outputs and the
                // joint accumulator are held (variableSamplingMPC.cpp:91)
                if (((!solved && all_fin) || fwd != 0) && fb_mode != 0)
                    fb_list[atomicAdd(fb_count, 1)] = inst;
                status[inst] = (solved && fwd != 0) ? VSMPC_STATUS_NUMERICAL : stat;
                n_factor[inst] = pass + 1;                       // backward recursions (one unless joint boxes became active)
                n_solve[inst] = (solved && fwd == 0) ? 1 : 0;    // committed forward pass
                n_pivot[inst] = sm.flags[2];
            }
            cd_stage_outputs(o, status, inst, lane, out2, status2);
            if constexpr (JL)
            {
                // next tick's guess: the working set of a committed solve, nothing otherwise
                if (jlset)
                    jlset[(size_t)inst * CD_JLSET_WORDS + lane] = (solved && fwd == 0) ? sm.clamp[lane] : 0u;
            }
            PHASE_CLK(3);
        }
    }
    if constexpr (JL)
    {
        if (warp == 0 && lane == 0)
            sm.flags[3] = again;
        __syncthreads();
        if (sm.flags[3] == 0)
            return;
    }
    else
        return;
    }   // pass
}

// The builds of the kernel (one body): register budget and pipeline per batch size
#define CD_KERNEL_ARGS                                                                                                     \
    const __grid_constant__ DeviceConfig cfgv, int B, const double* __restrict__ qd_all, double* __restrict__ ws_all,      \
        double* __restrict__ z_all, double* __restrict__ st, double* __restrict__ out_rows, int* __restrict__ status,       \
        int* __restrict__ n_factor, int* __restrict__ n_solve, int* __restrict__ n_pivot, size_t ws_stride, int want_z,     \
        int* __restrict__ fb_list, int* __restrict__ fb_count, int fb_mode, double* __restrict__ out2, int* __restrict__ status2,    \
        unsigned* __restrict__ jlset
#define CD_KERNEL_PASS cfgv, B, qd_all, ws_all, z_all, st, out_rows, status, n_factor, n_solve, n_pivot, ws_stride, want_z, fb_list, fb_count, fb_mode, out2, status2, jlset
// large batches: eight CTAs per SM, 128 registers
__global__ void __launch_bounds__(CD_THREADS, 8) qp_condensed_kernel(CD_KERNEL_ARGS) { qp_condensed_body<false, false>(CD_KERNEL_PASS); }
// (a 168-register build with six CTAs per SM for many-wave batches was tried for this body as well: 2.42 ms against 2.39 ms at
// B = 16 384, ahead only where the wave count favours it — profiles/r02_k2_experiments.md; the JL build below, whose 128-register
// form spills more, does gain from it)
// handles with joint-limit rows (vsmpc_config.use_joint_limits): the same lock-step body with the working set of the joint boxes
// carried through the elimination (CdClamp) and a pass loop around the solve; any batch size.  168 registers, six CTAs per SM:
// 3.57 M against 3.21 M closed-loop solves/s in the parameter sweep at 2048 instances, 4.50 M against 3.75 M at 16 384, compared
// with the 128-register build (profiles/r02bo_time_jl.txt)
__global__ void __launch_bounds__(CD_THREADS, 6) qp_condensed_kernel_jl(CD_KERNEL_ARGS) { qp_condensed_body<false, true>(CD_KERNEL_PASS); }
// (a 144-register build for the one wave of seven CTAs per SM at B = 1024 was tried: 207 us instead of 165 — the register file
// is split over the four sub-partitions, 14 warps put four on two of them and four warps of 144 registers do not fit 16 384,
// so the SM holds six CTAs and the launch takes two waves; 128 registers is the cap for anything above twelve warps per SM)
// small batches, at most four CTAs per SM: decoupled pipeline, no register cap that matters (232 registers, no spills)
__global__ void __launch_bounds__(CD_THREADS, 4) qp_condensed_kernel_pipe(CD_KERNEL_ARGS) { qp_condensed_body<true, false>(CD_KERNEL_PASS); }

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

size_t condensed_jlset_words()
{
    return CD_JLSET_WORDS;   // one word per joint block
}

size_t condensed_ws_doubles(const DeviceConfig& cfg)
{
    return (size_t)cfg.Nc * (WSC_STAGE + WSC_U);   // the stages, then raw H_uu per joint block (written by the JL build only)
}

cudaError_t launch_qp_condensed(const DeviceConfig* d_cfg, const DeviceConfig& h_cfg, int B, const double* qd,
                                double* ws, double* z, double* st, double* out_rows, int* status, int* n_factor,
                                int* n_solve, int* n_pivot, int want_z, int* fb_list, int* fb_count, int fb_mode,
                                double* out2, int* status2, unsigned* jlset, cudaStream_t s)
{
    // small batches (at most four CTAs per SM): the decoupled pipeline, which shortens a single solve; otherwise lock step
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static int pipe_ctas = -1;      // development: VSMPC_K2_PIPE_CTAS overrides the number of CTAs per SM up to which PIPE is used
    if (pipe_ctas < 0)
    {
        const char* e = getenv("VSMPC_K2_PIPE_CTAS");
        pipe_ctas = e ? atoi(e) : 4;
    }
    const size_t wsd = condensed_ws_doubles(h_cfg);
    const int fbm = fb_list && fb_count ? fb_mode : 0;
    if (h_cfg.use_jl)
        qp_condensed_kernel_jl<<<B, CD_THREADS, 0, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status, n_factor, n_solve, n_pivot, wsd,
                                                        want_z, fb_list, fb_count, fbm, out2, status2, jlset);
    else if (B <= pipe_ctas * sms)
        qp_condensed_kernel_pipe<<<B, CD_THREADS, 0, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status, n_factor, n_solve, n_pivot, wsd,
                                                          want_z, fb_list, fb_count, fbm, out2, status2, jlset);
    else
        qp_condensed_kernel<<<B, CD_THREADS, 0, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status, n_factor, n_solve, n_pivot, wsd,
                                                     want_z, fb_list, fb_count, fbm, out2, status2, jlset);
    return cudaGetLastError();
}

} // namespace vsmpc
