// K2 (long horizons) — condensed-throttle Riccati QP kernel with several column warps.  FP64 on the CUDA cores, the
// deferred down-date of Om on the FP64 tensor cores.
//
// Same algorithm and the same replaced reference code as vsmpc_qp_condensed.cu (IMPCProblem::solve -> OsqpEigen,
// MPC/src/IMPCProblem/IMPCProblem.cpp:196-298; VariableSamplingMPC::solveMPC output extraction,
// variableSamplingMPC.cpp:88-112), for horizons beyond its 6 throttle blocks / 32 knots (BASELINE configs[3]: 2-4x
// the reference knot count).  The value function V_k(x; theta) = 1/2 x'P x + x'Psi theta + 1/2 theta'Om theta now has
// theta = (v_0 .. v_{nblk-1}, 1, held joint block) with up to 192 columns:
// * warp 0 owns P and runs exactly the recursion of the narrow kernel (shared code: vsmpc_condensed_core.cuh);
// * warps 1..G own 32 parameter columns each (lane = column, Psi column in registers).  The columns are independent
//   given what warp 0 publishes per knot, so all column warps read one mailbox, one knot behind warp 0; only the
//   updates of Om (dynamic shared memory, odd leading dimension) need a named barrier among them;
// * the rank-8 down-dates of Om are stacked in the workspace and contracted once by all warps with
//   mma.sync.m8n8k4.f64 over the upper 8 x 8 tiles;
// * the reduced Hessian H_r (in place in Om) is inverted by exchange pivots (x_q <-> y_q in y = H_r x); the
//   Goldfarb-Idnani dual active set then keeps the SAME matrix as the principal pivot transform of H_r over the free
//   variables: adding a bound to / dropping it from the working set is one more pivot on that index, and column p of
//   the matrix holds both the primal direction (free rows) and the multiplier direction (active rows) — no working-set
//   inverse, no back-solves; one variable per thread.  A pivot is a rank-1 change T' = T - (T[:,q] + e_q)(T[q,:] - e_q)'/d:
//   pivots are DEFERRED as rows of two 8 x nv stacks (entries of T are read through the stacks) and applied eight at a
//   time as one rank-8 update of the whole matrix on the FP64 tensor cores, which cuts the shared-memory traffic per
//   pivot by eight;
// * F_k theta* by all warps, forward rollout by warp 0.
#include "vsmpc_condensed_core.cuh"

namespace vsmpc
{

constexpr int CW_MAXG = 6;                          // column warps
constexpr int CW_MAXTHREADS = 256;                  // 1 + CW_MAXG warps, rounded up to two warps per SM sub-partition
constexpr int CW_SMEM_LIMIT = 227 * 1024;           // opt-in shared memory per CTA on sm_100
constexpr int CW_MD = 8;                            // pivots deferred before one rank-8 update of the matrix

// working set of the optional joint boxes (JL builds; see CdClamp in vsmpc_condensed_core.cuh and the reference-horizon kernel)
template <bool JL> struct CwJlSmem
{
    unsigned clamp[MAX_ITER];   // per joint block: bit c = increment c held at its upper bound, bit 8 + c at its lower bound
    unsigned cand[MAX_ITER];    // forward pass: increments that would join / leave
    double jb[2 * NJ];          // bounds of the increments: lower [8], upper [8]
    double hb[2][NJ];           // warp 0 -> column warps, per mailbox slot: sum_c H_uu[m][c] b_c
};
template <> struct CwJlSmem<false>
{
};
constexpr int CW_JL_PASSES = 24, CW_JL_PLAIN = 6;   // more boxes than at the reference horizon: more passes before the hand-over
constexpr int CW_WSU = NJ * NJ;                       // per joint block, behind the Nc stages of the workspace: raw H_uu [8][8]

template <bool JL> struct alignas(16) CwSmemT : CwJlSmem<JL>
{
    double cf[CCF];
    alignas(16) double lam[6 * NJ];   // dt-free B_J rows: [q][a]
    double Qd[NX];
    double Rqd[NJ];
    double dtk[MAX_ITER];
    CdSlot slot[2];                   // warp 0 -> column warps mailbox, slot = knot & 1
    alignas(16) double Mt[NX * LDM];  // warp 0: transposition buffer, gain rows of the knot in flight
    double xs[40];                    // forward rollout: x (26), throttle block in effect (4), dq in effect (8)
    double ostage[48];                // outputs staged until the forward rollout has checked the optional joint boxes
    int flags[4];
};

constexpr int CW_ALIAS_DOUBLES = (2 * (int)sizeof(CdSlot) + NX * LDM * (int)sizeof(double)) / (int)sizeof(double);   // slot[2] + Mt

struct CwLayout
{
    int G, ldc;                 // column warps, columns carried (32 G)
    int nv, aff, D0, nlo, ldo;  // throttle variables, affine column, first held-joint column, order of Om, its ld (odd)
    int nt;                     // 8 x 8 tiles per side of Om
    int wsF, wsH, stage;        // workspace per elimination knot: K [8][26] | F [8][ldc] | H_utheta [8][ldc]
    int nvp, nvs;               // nv rounded up to 8; stride of the deferred-pivot stacks
    int om, fth, theta, vv, grad, stk, xref, rb, actg, total;   // offsets (doubles) in dynamic shared memory
};

__host__ __device__ inline CwLayout cw_layout(const DeviceConfig& cfg)
{
    CwLayout L;
    const bool held = cfg.Nc - 1 < cfg.N - 1;
    L.nv = NT * cfg.nblk;
    L.aff = L.nv;
    L.D0 = L.nv + 1;
    L.nlo = L.nv + 1 + (held ? NJ : 0);
    L.G = (L.nlo + 31) / 32;
    L.ldc = 32 * L.G;
    L.ldo = L.nlo | 1;
    L.nt = (L.nlo + 7) / 8;
    L.nvp = (L.nv + 7) & ~7;
    L.nvs = L.nvp + 4;          // + 4: the tensor-core fragments of the stacks are read without bank conflicts
    L.wsF = NJ * NX;
    L.wsH = L.wsF + NJ * L.ldc;
    L.stage = L.wsH + NJ * L.ldc;
    int o = 0;
    L.om = o;     o += (L.nlo * L.ldo + 1) & ~1;
    L.fth = o;    o += cfg.Nc * NJ;
    L.theta = o;  o += L.ldc;
    L.vv = o;     o += L.nv;
    L.grad = o;   o += L.nv;
    // deferred pivots: A [8][nvs] then B [8][nvs].  They are used only after the recursion, when the mailbox slots and the
    // transposition buffer of the static part are free: up to 29 throttle blocks the stacks live THERE (stk = -1) — at 3x the
    // reference knot count that takes the CTA from 118 KB to 105 KB, i.e. from one CTA per SM to two
    if (2 * CW_MD * L.nvs <= CW_ALIAS_DOUBLES)
        L.stk = -1;
    else
    {
        L.stk = o;
        o += 2 * CW_MD * L.nvs;
    }
    L.xref = o;   o += 12 * cfg.NC;
    L.rb = o;     o += 4 * 8;      // two block-reduction buffers: value [8], index [8] each
    L.actg = o;   o += (L.nvp + 1) / 2 + 2;   // working-set guess, int [nvp]; then two doubles broadcast by the dropping thread
    L.total = o;
    return L;
}

using CwSmem = CwSmemT<false>;

#ifdef VSMPC_PHASE_CLOCKS
__device__ long long g_wide_clk[4096][16];
#define WCLK(slot) do { if (threadIdx.x == 0 && inst < 4096) g_wide_clk[inst][slot] = clock64(); } while (0)
#define WSUB(acc, t0) do { const long long t1__ = clock64(); acc += t1__ - t0; t0 = t1__; } while (0)
#else
#define WCLK(slot) do { } while (0)
#define WSUB(acc, t0) do { } while (0)
#endif

__device__ __forceinline__ void cw_bar_columns(int n_threads)
{
    asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// propagation of the parameter columns through knot k (Psi'' = Psi' + P'D, Om += D'Psi'' + Psi''D, Psi <- T'Psi'');
// returns dt B_J' Psi''[:, gc] in bj2.  gc: column of this lane
template <class CwCtx>
__device__ __forceinline__ void w_prop(const CwCtx& c, const CwLayout& L, double* __restrict__ Om,
                                       const double* __restrict__ xref, int k, bool tail, int gc, double (&s)[NX],
                                       double (&bj2)[NJ])
{
    const DeviceConfig& cfg = c.cfg;
    auto& sm = c.sm;
    const double* cf = sm.cf;
    const double dt = sm.dtk[k];
    CdSlot& sl = sm.slot[k & 1];
    const int tb = throttle_block(k, cfg.Ns, cfg.Nc);
    const int ldo = L.ldo;
    const bool isAff = gc == L.aff;
    const bool spV = gc < L.nv && (gc >> 2) == tb;
    const bool isD = tail && gc >= L.D0 && gc < L.D0 + NJ;
    const double jgt = cf[QD_JGT];
    if (isAff)
    {
        const int rc = ref_col(k, cfg.Ns);
#pragma unroll
        for (int r = 0; r < 12; ++r)
            s[r] = fma(-sm.Qd[r], xref[r * cfg.NC + rc], s[r]);   // tracking gradient of x_{k+1}
    }
    double bv[NT], bd[NJ];
#pragma unroll
    for (int q = 0; q < NT; ++q)
        bv[q] = dt * (cf[QD_JG + q] * s[IX_TD + q] + jgt * s[IX_T + q]);
    const double baff = c_dot(s, cf, dt);
    if (tail)
        bjT_dot(s, sm.lam, dt, bd);
    const int col = spV ? (gc & 3) : (isAff ? 4 : (isD ? 5 + gc - L.D0 : -1));
    if (col >= 0)
    {
#pragma unroll
        for (int r = 0; r < NX; ++r)
            s[r] += sl.PD[r * NPD + col];
    }
    bjT_dot(s, sm.lam, dt, bj2);
    // Om += D'Psi'' (rows of the special columns: every lane its own column), barrier among the column warps, then
    // Om += Psi'D (columns of the special columns: every lane its own row)
    if (gc < L.nlo)
    {
        double ov[NT];
#pragma unroll
        for (int q = 0; q < NT; ++q)
            ov[q] = Om[(4 * tb + q) * ldo + gc];
        const double oa = Om[L.aff * ldo + gc];
#pragma unroll
        for (int q = 0; q < NT; ++q)
            Om[(4 * tb + q) * ldo + gc] = ov[q] + dt * (cf[QD_JG + q] * s[IX_TD + q] + jgt * s[IX_T + q]);
        Om[L.aff * ldo + gc] = oa + c_dot(s, cf, dt);
        if (tail)
        {
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                Om[(L.D0 + a) * ldo + gc] += bj2[a];
        }
    }
    cw_bar_columns(32 * L.G);
    if (gc < L.nlo)
    {
        double ov[NT];
#pragma unroll
        for (int q = 0; q < NT; ++q)
            ov[q] = Om[gc * ldo + 4 * tb + q];
        const double oa = Om[gc * ldo + L.aff];
#pragma unroll
        for (int q = 0; q < NT; ++q)
            Om[gc * ldo + 4 * tb + q] = ov[q] + bv[q];
        Om[gc * ldo + L.aff] = oa + baff;
        if (tail)
        {
#pragma unroll
            for (int a = 0; a < NJ; ++a)
                Om[gc * ldo + L.D0 + a] += bd[a];
        }
    }
    applyTtx(s, cf, dt);
}

// down-date of the parameter columns with the eliminated block: F = H_uu^-1 H_utheta, Psi -= H_ux' F; H_utheta and F go
// to the workspace stacks for the deferred down-date of Om
// cany (JL builds): clamped components of the block — raw H_utheta rows come out in their F rows, zero rows go to the H stack,
// Psi skips them (see b_downdate of the reference-horizon kernel)
template <bool JL = false>
__device__ __forceinline__ void w_downdate(const CdSlot& sl, const CwLayout& L, int gc, double (&s)[NX],
                                           const double (&hut)[NJ], double* __restrict__ wsk, bool clear_col, unsigned cany = 0u)
{
    double f[NJ];
#pragma unroll
    for (int m = 0; m < NJ; ++m)
        wsk[L.wsH + m * L.ldc + gc] = (JL && ((cany >> m) & 1u)) ? 0.0 : hut[m];
    const double2* hi = reinterpret_cast<const double2*>(sl.Hinv);
#pragma unroll
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
        f[a] = v0 + v1;
        wsk[L.wsF + a * L.ldc + gc] = f[a];
    }
    if (gc < L.nlo)
    {
#pragma unroll
        for (int m = 0; m < NJ; ++m)
        {
            if constexpr (JL)
            {
                if ((cany >> m) & 1u)
                    continue;
            }
            const double2* hr = reinterpret_cast<const double2*>(sl.Hux + m * NX);
#pragma unroll
            for (int j = 0; j < NX / 2; ++j)
            {
                const double2 hh = hr[j];
                s[2 * j] = fma(-f[m], hh.x, s[2 * j]);
                s[2 * j + 1] = fma(-f[m], hh.y, s[2 * j + 1]);
            }
        }
    }
    if (clear_col)
    {
#pragma unroll
        for (int j = 0; j < NX; ++j)
            s[j] = 0.0;
    }
}

// Om -= H'F with H, F the (8 Nc) x ldc stacks of the workspace, on the upper 8 x 8 tiles; one work item = tile row ti,
// four tile columns from tj0; A fragment = H' (lane l: row l >> 2 of the tile, stack row l & 3), B fragment = F (stack
// row l & 3, column l >> 2), C fragment: row l >> 2, columns 2 (l & 3) + {0, 1}
__device__ __forceinline__ void cw_omega_item(const double* __restrict__ ws, const CwLayout& L, int n_rows,
                                              double* __restrict__ Om, int ti, int tj0, int lane)
{
    double c0[4], c1[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
        c0[t] = c1[t] = 0.0;
    const int lr = lane & 3, lc = lane >> 2;
    int colb[4];
#pragma unroll
    for (int t = 0; t < 4; ++t)
        colb[t] = 8 * min(tj0 + t, L.nt - 1) + lc;
#pragma unroll 8
    for (int r0 = 0; r0 < n_rows; r0 += 4)
    {
        const int r = r0 + lr;
        const double* __restrict__ row = ws + (size_t)(r >> 3) * L.stage + (r & 7) * L.ldc;
        const double a = row[L.wsH + 8 * ti + lc];
        double b[4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
            b[t] = row[L.wsF + colb[t]];
#pragma unroll
        for (int t = 0; t < 4; ++t)
            asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                : "+d"(c0[t]), "+d"(c1[t])
                : "d"(a), "d"(b[t]));
    }
#pragma unroll
    for (int t = 0; t < 4; ++t)
    {
        const int tj = tj0 + t;
        const int gi = 8 * ti + lc;
#pragma unroll
        for (int e = 0; e < 2; ++e)
        {
            const int gj = 8 * tj + 2 * lr + e;
            const double v = e == 0 ? c0[t] : c1[t];
            if (tj < L.nt && gi < L.nlo && gj < L.nlo)
            {
                Om[gi * L.ldo + gj] -= v;
                if (tj > ti)
                    Om[gj * L.ldo + gi] -= v;
            }
        }
    }
}

// ---- exchange pivots on T = Om[first:nv, first:nv], deferred ------------------------------------------------------
// In (outputs) = T (inputs) a pivot on q swaps the roles of input q and output q: pivoting every index turns H into
// H^-1, pivoting q again undoes it.  T_eff = T0 - sum_{m < M} A[m][:]' B[m][:] with T0 in Om and the M <= 8 pivots
// since the last flush in the stacks A, B (rows >= M are zero).
struct CwPiv
{
    double* Om;
    double* As;
    double* Bs;
    int ldo, nvs, first, nv, nvp;
};

__device__ __forceinline__ double cw_teff(const CwPiv& P, int M, int i, int j)
{
    double t = P.Om[i * P.ldo + j];
#pragma unroll
    for (int m = 0; m < CW_MD; ++m)
        if (m < M)
            t = fma(-P.As[m * P.nvs + i], P.Bs[m * P.nvs + j], t);
    return t;
}

// T0 -= A'B over all 8 x 8 tiles (mma.sync.m8n8k4.f64: A fragment = -A' (lane l: row l >> 2 of the tile, pivot l & 3),
// B fragment = B (pivot l & 3, column l >> 2), C fragment = T0 tile (row l >> 2, columns 2 (l & 3) + {0, 1})), then the
// stacks are cleared
__device__ __forceinline__ void cw_flush(const CwPiv& P, int& M, int warp, int lane, int nwarps, int nthr)
{
    const int nt8 = P.nvp >> 3, ngrp = (nt8 + 3) >> 2;
    const int lr = lane & 3, lc = lane >> 2;
    // one work item = tile row ti, four tile columns: their loads and tensor-core instructions overlap
    for (int item = warp; item < nt8 * ngrp; item += nwarps)
    {
        const int ti = item / ngrp, tj0 = 4 * (item - ti * ngrp);
        const int gi = 8 * ti + lc;
        double a[CW_MD / 4];
#pragma unroll
        for (int h = 0; h < CW_MD / 4; ++h)
            a[h] = -P.As[(4 * h + lr) * P.nvs + gi];
        double c0[4], c1[4], b[4][CW_MD / 4];
#pragma unroll
        for (int t = 0; t < 4; ++t)
        {
            const int tj = min(tj0 + t, nt8 - 1);
            const int gj = 8 * tj + 2 * lr;
            const bool on = tj0 + t < nt8 && gi < P.nv;
            c0[t] = (on && gj < P.nv) ? P.Om[gi * P.ldo + gj] : 0.0;
            c1[t] = (on && gj + 1 < P.nv) ? P.Om[gi * P.ldo + gj + 1] : 0.0;
#pragma unroll
            for (int h = 0; h < CW_MD / 4; ++h)
                b[t][h] = P.Bs[(4 * h + lr) * P.nvs + 8 * tj + lc];
        }
#pragma unroll
        for (int h = 0; h < CW_MD / 4; ++h)
#pragma unroll
            for (int t = 0; t < 4; ++t)
                asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                    : "+d"(c0[t]), "+d"(c1[t])
                    : "d"(a[h]), "d"(b[t][h]));
#pragma unroll
        for (int t = 0; t < 4; ++t)
        {
            const int gj = 8 * (tj0 + t) + 2 * lr;
            const bool on = tj0 + t < nt8 && gi < P.nv;
            if (on && gj < P.nv)
                P.Om[gi * P.ldo + gj] = c0[t];
            if (on && gj + 1 < P.nv)
                P.Om[gi * P.ldo + gj + 1] = c1[t];
        }
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 2 * CW_MD * P.nvs; e += nthr)
        P.As[e] = 0.0;      // A and B are contiguous
    __syncthreads();
    M = 0;
}

// pivot on q as the rank-1 change T' = T - (T[:, q] + e_q)(T[q, :] - e_q)' / T[q][q]: thread e appends its entries of the
// two vectors to the stacks; one barrier, plus the flush every CW_MD pivots
__device__ __forceinline__ bool cw_pivot(const CwPiv& P, int& M, int q, int warp, int lane, int nwarps, int nthr)
{
    const int e = threadIdx.x;
    const double d = cw_teff(P, M, q, q);
    const double dinv = 1.0 / d;
    if (e >= P.first && e < P.nv)
    {
        const double u = cw_teff(P, M, e, q), v = cw_teff(P, M, q, e);
        const double one = e == q ? 1.0 : 0.0;
        P.As[M * P.nvs + e] = (u + one) * dinv;
        P.Bs[M * P.nvs + e] = v - one;
    }
    __syncthreads();
    ++M;
    if (M == CW_MD)
        cw_flush(P, M, warp, lane, nwarps, nthr);
    return (d > 0.0) && isfinite(d);
}

// block-wide max / min of non-negative doubles with the lowest thread attaining it (one barrier; rb: value [8] and
// index [8], not reused before the next barrier)
template <bool MAX>
__device__ __forceinline__ double cw_block_best(double v, double* __restrict__ rb, int warp, int lane, int nwarps, int& arg)
{
    int a;
    const double w = MAX ? warp_max_nonneg(v, a) : warp_min_nonneg(v, a);
    int* rbi = reinterpret_cast<int*>(rb + 8);
    if (lane == 0)
    {
        rb[warp] = w;
        rbi[warp] = 32 * warp + a;
    }
    __syncthreads();
    double best = rb[0];
    arg = rbi[0];
    for (int q = 1; q < nwarps; ++q)
    {
        const double o = rb[q];
        if (MAX ? (o > best) : (o < best))
        {
            best = o;
            arg = rbi[q];
        }
    }
    return best;
}

// THREADS / MINB: launch bounds per number of column warps, so that the register cap follows what shared memory allows
// anyway (<96, 4>: two column warps, 168 registers; <128, 2> and <224, 1>: 255 registers) — the column warps keep a Psi
// column plus three 8-vectors in registers, and spills there sit on the critical path of every knot
template <int THREADS, int MINB, bool JL>
__global__ void __launch_bounds__(THREADS, MINB)
qp_condensed_wide_kernel(const __grid_constant__ DeviceConfig cfgv, int B, const double* __restrict__ qd_all,
                         double* __restrict__ ws_all, double* __restrict__ z_all,
                         double* __restrict__ st, double* __restrict__ out_rows, int* __restrict__ status,
                         int* __restrict__ n_factor, int* __restrict__ n_solve, int* __restrict__ n_pivot, size_t ws_stride,
                         int want_z, int* __restrict__ fb_list, int* __restrict__ fb_count, int fb_mode,
                         signed char* __restrict__ wset_all, int warm, double* __restrict__ out2, int* __restrict__ status2,
                         unsigned* __restrict__ jlset)
{
    using CwSm = CwSmemT<JL>;
    using CwCtx = CdCtxT<CwSm>;
    extern __shared__ __align__(16) unsigned char cw_raw[];
    CwSm& sm = *reinterpret_cast<CwSm*>(cw_raw);
    double* dyn = reinterpret_cast<double*>(cw_raw + sizeof(CwSm));
    const DeviceConfig& cfg = cfgv;
    const CwLayout L = cw_layout(cfg);
    double* Om = dyn + L.om;
    double* fth = dyn + L.fth;
    double* theta = dyn + L.theta;
    double* vv = dyn + L.vv;
    double* grad = dyn + L.grad;
    double* stkp = L.stk < 0 ? reinterpret_cast<double*>(&sm.slot[0]) : dyn + L.stk;
    const CwPiv P{Om, stkp, stkp + CW_MD * L.nvs, L.ldo, L.nvs, 0, L.nv, L.nvp};
    double* xref = dyn + L.xref;
    double* rbA = dyn + L.rb;
    double* rbB = dyn + L.rb + 16;

    const int nthr = blockDim.x, nwarps = nthr >> 5;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int inst = blockIdx.x;
    const double* qd = qd_all + (size_t)inst * cfg.qd_stride;
    const int N = cfg.N, Nc = cfg.Nc;
    const int nv = L.nv, ldo = L.ldo;
    const bool held = Nc - 1 < N - 1;
    const int gc = (warp - 1) * 32 + lane;     // column of this lane (warps >= 1)
    CwCtx c{cfg, sm, ws_all + (size_t)inst * ws_stride, lane, L.D0, held ? Nc - 1 : -1};

    WCLK(0);
    // ---- stage the QP data; finiteness gate ----------------------------------------------------------------------
    bool fin = true;
    for (int e = threadIdx.x; e < cfg.qd_stride; e += nthr)
    {
        const double v = qd[e];
        fin = fin && isfinite(v);
        if (e < CCF)
            sm.cf[e] = v;
        else if (e >= QD_XREF && e < QD_XREF + 12 * cfg.NC)
            xref[e - QD_XREF] = v;
    }
    for (int e = threadIdx.x; e < L.nlo * ldo; e += nthr)
        Om[e] = 0.0;
    for (int e = threadIdx.x; e < 2 * CW_MD * L.nvs; e += nthr)
        P.As[e] = 0.0;
    if (threadIdx.x < NX)
        sm.Qd[threadIdx.x] = cfg.Qd[threadIdx.x];
    if (threadIdx.x < NJ)
        sm.Rqd[threadIdx.x] = cfg.Rqd[threadIdx.x];
    for (int e = threadIdx.x; e < N; e += nthr)
        sm.dtk[e] = cfg.dt[e];
    for (int e = threadIdx.x; e < 6 * NJ; e += nthr)
        sm.lam[e] = qd[(e < 3 * NJ ? QD_LLIN : QD_LANG - 3 * NJ) + e];
    if constexpr (JL)
    {
        if (threadIdx.x < 2 * NJ)
            sm.jb[threadIdx.x] = qd[QD_JLO + threadIdx.x];
        // warm start: the working set of the joint boxes the last solve of this instance ended with (zeros after configure)
        for (int e = threadIdx.x; e < MAX_ITER; e += nthr)
            sm.clamp[e] = (jlset && e < Nc && e < CD_JLSET_WORDS) ? jlset[(size_t)inst * CD_JLSET_WORDS + e] : 0u;
    }
    const bool all_fin = __syncthreads_and(fin);
    // JL builds: primal-dual active set on the joint boxes around the whole solve, as in the reference-horizon kernel
    for (int pass = 0;; ++pass)
    {
    int stat = all_fin ? VSMPC_STATUS_SOLVED : VSMPC_STATUS_NUMERICAL;
    if constexpr (JL)
    {
        if (pass > 0)
        {
            for (int e = threadIdx.x; e < L.nlo * ldo; e += nthr)
                Om[e] = 0.0;
            for (int e = threadIdx.x; e < 2 * CW_MD * L.nvs; e += nthr)
                P.As[e] = 0.0;
            __syncthreads();
        }
    }

    WCLK(1);
    // ---- factorisation: warp 0 = P recursion, warps 1..G = parameter columns, one knot apart ----------------------
    double y[NX];   // warp 0: row `lane` of P ; column warps: column gc of Psi
#pragma unroll
    for (int j = 0; j < NX; ++j)
        y[j] = 0.0;
    const double qd_lane = lane < NX ? sm.Qd[lane] : 0.0;
    bool ok = true;
    const int n_it = c.kS < 0 ? N + 1 : (N - 1 - c.kS) + 3 + c.kS + 1;
    long long wk0 = 0, wk1 = 0;
    (void)wk0; (void)wk1;
    for (int t = 0; t < n_it; ++t)
    {
        int ta, ka, tbk, kb;
        cd_schedule(t, N, c.kS, ta, ka, tbk, kb);
        long long tclk = clock64();
        (void)tclk;
        if (warp == 0)
        {
            if (ta != TK_NONE)
            {
                CdSlot& sl = sm.slot[ka & 1];
                double hux[NJ];
                double2 own;
                const bool elim = ta == TK_STAGE && !(held && ka >= Nc - 1);
                if (ta != TK_SCHUR)
                {
                    a_prop(c, ka, sl, elim, y, qd_lane, hux, own);
                    WSUB(wk0, tclk);
                }
                else
                {
                    // Schur step of the held joint block: H_ux = Psi_T[:, d]' (published by the column warps),
                    // H_uu = Om_T[d, d] + R
                    const int r = lane & 7, q = lane >> 3;
#pragma unroll
                    for (int m = 0; m < NJ; ++m)
                        hux[m] = lane < NX ? sl.Hux[m * NX + lane] : 0.0;
                    own.x = Om[(L.D0 + r) * ldo + L.D0 + 2 * q] + (2 * q == r ? sm.Rqd[r] : 0.0);
                    own.y = Om[(L.D0 + r) * ldo + L.D0 + 2 * q + 1] + (2 * q + 1 == r ? sm.Rqd[r] : 0.0);
                }
                __syncwarp();
                if (elim || ta == TK_SCHUR)
                {
                    CdClamp cl{0u, nullptr, nullptr, nullptr};
                    if constexpr (JL)
                        cl = CdClamp{sm.clamp[ka], sm.jb, sm.hb[ka & 1], c.ws + (size_t)Nc * L.stage + ka * CW_WSU};
                    ok = a_eliminate<CwSm, JL>(c, sl, y, hux, own, c.ws + (size_t)ka * L.stage, cl) && ok;
                }
                WSUB(wk1, tclk);
            }
        }
        else if (tbk != TK_NONE && warp <= L.G)   // (a CTA may carry idle warps that only join the block-wide phases)
        {
            CdSlot& sl = sm.slot[kb & 1];
            const bool tail = held && kb >= Nc - 1;
            const bool isD = held && gc >= L.D0 && gc < L.D0 + NJ;
            double hut[NJ];
            bool down = false;
            if (tbk != TK_SCHUR)
            {
                w_prop(c, L, Om, xref, kb, tail, gc, y, hut);
                WSUB(wk0, tclk);
                if (tbk == TK_PROP)
                {
                    if (isD)
                    {
#pragma unroll
                        for (int j = 0; j < NX; ++j)
                            sl.Hux[(gc - L.D0) * NX + j] = y[j];
                    }
                }
                else
                    down = !tail;
            }
            else
            {
#pragma unroll
                for (int m = 0; m < NJ; ++m)
                    hut[m] = (gc < L.nlo && !isD) ? Om[(L.D0 + m) * ldo + gc] : 0.0;
                down = true;
            }
            if (down)
            {
                if (gc == L.aff)
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
                        // the constants of the clamped components in the value function and in the free rows (CdClamp).  Only
                        // column `aff` of Om is ever read (gradient of the reduced QP, H_utheta of the Schur step): its entry
                        // in row gc belongs to this lane after the column barrier of w_prop
#pragma unroll
                        for (int cc = 0; cc < NJ; ++cc)
                        {
                            if (!((cany >> cc) & 1u))
                                continue;
                            const double bc = cd_bval(sm.jb, cm, cc);
                            if (gc < L.nlo && !isD)
                                Om[gc * ldo + L.aff] += bc * hut[cc];
                            if (gc == L.aff)
                            {
#pragma unroll
                                for (int j = 0; j < NX; ++j)
                                    y[j] = fma(bc, sl.Hux[cc * NX + j], y[j]);
                            }
                        }
                        if (gc == L.aff)
                        {
#pragma unroll
                            for (int m = 0; m < NJ; ++m)
                                if (!((cany >> m) & 1u))
                                    hut[m] += sm.hb[kb & 1][m];
                        }
                    }
                }
                w_downdate<JL>(sl, L, gc, y, hut, c.ws + (size_t)kb * L.stage, schur && isD, cany);
                WSUB(wk1, tclk);
                if (schur && gc < L.nlo)
                {
#pragma unroll
                    for (int a = 0; a < NJ; ++a)
                    {
                        Om[(L.D0 + a) * ldo + gc] = 0.0;
                        Om[gc * ldo + L.D0 + a] = 0.0;
                    }
                }
            }
        }
        __syncthreads();
    }
    WCLK(2);
#ifdef VSMPC_PHASE_CLOCKS
    if (lane == 0 && inst < 4096 && (warp == 0 || warp == 1 || warp == nwarps - 1))
    {
        const int b = warp == 0 ? 9 : (warp == 1 ? 13 : 10);
        g_wide_clk[inst][b] = wk0;
        g_wide_clk[inst][b == 13 ? 14 : b + 2] = wk1;     // 9/11: warp 0; 10/12..: see the tool
    }
#endif
    if (warp == 0 && lane == 0)
        sm.flags[0] = ok ? 0 : 1;
    if (L.stk < 0)
    {
        // the deferred-pivot stacks share the mailbox slots / transposition buffer (cw_layout): the recursion is over (the
        // __syncthreads that closed its last knot), they start empty; the barrier after the Om down-date publishes the zeros
        for (int e = threadIdx.x; e < 2 * CW_MD * L.nvs; e += nthr)
            P.As[e] = 0.0;
    }
    // Psi_0' x0 while the column warps still hold their columns
    double g_psi = 0.0;
    if (warp >= 1 && gc < nv)
    {
#pragma unroll
        for (int j = 0; j < NX; ++j)
            g_psi = fma(y[j], sm.cf[QD_X0 + j], g_psi);
    }
    // deferred down-date of Om on the tensor cores, work items dealt round-robin to the warps
    {
        int item = 0;
        for (int ti = 0; ti < L.nt; ++ti)
            for (int tj0 = ti; tj0 < L.nt; tj0 += 4, ++item)
                if (item % nwarps == warp)
                {
                    // columns of tile row ti are zero in the stacks of the knots whose throttle block comes before
                    // them (block b enters the value function at the last knot it acts on): staircase contraction
                    const int k_end = (8 * ti + 7 >= nv) ? Nc : min(Nc, 2 * ti + cfg.Ns + 1);
                    cw_omega_item(c.ws, L, NJ * k_end, Om, ti, tj0, lane);
                }
    }
    __syncthreads();

    WCLK(3);
    // ---- reduced QP in the throttle variables: gradient, Hessian (in place in Om), inverse ---------------------------
    const bool pinned = sm.cf[QD_PINNED] != 0.0;
    const int first = pinned ? NT : 0;
    const double lo = sm.cf[QD_VMIN], up = sm.cf[QD_VMAX];
    if (warp >= 1 && gc < nv)
        grad[gc] = g_psi + Om[gc * ldo + L.aff] - (gc < NT ? cfg.w_i * sm.cf[QD_VBAR + gc] : 0.0);
    for (int e = threadIdx.x; e < nv; e += nthr)
    {
        const int blk = e >> 2;
        Om[e * ldo + e] += cfg.w_t * ((blk > 0 ? 1.0 : 0.0) + (blk < cfg.nblk - 1 ? 1.0 : 0.0)) + (blk == 0 ? cfg.w_i : 0.0);
        if (blk > 0)
            Om[e * ldo + e - NT] -= cfg.w_t;
        if (blk < cfg.nblk - 1)
            Om[e * ldo + e + NT] -= cfg.w_t;
    }
    __syncthreads();
    if (pinned)
    {
        // block 0 is a parameter: fold it into the gradient
        for (int e = NT + threadIdx.x; e < nv; e += nthr)
        {
            double g = grad[e];
#pragma unroll
            for (int j = 0; j < NT; ++j)
                g = fma(Om[e * ldo + j], sm.cf[QD_VBAR + j], g);
            grad[e] = g;
        }
        __syncthreads();
    }
    // equilibration to a unit diagonal (v = S v~, S = diag(H_r)^-1/2): the pivots stay O(1), which the deferred rank-1
    // form needs (it rebuilds the pivot row / column by subtraction, exact only up to eps x max(d, 1 / d))
    double S_e = 1.0;
    {
        const int e = threadIdx.x;
        if (e >= first && e < nv)
        {
            const double hd = Om[e * ldo + e];
            S_e = (hd > 0.0 && isfinite(hd)) ? rsqrt(hd) : 1.0;
            vv[e] = S_e;
            grad[e] *= S_e;
        }
        __syncthreads();
        for (int i = first + warp; i < nv; i += nwarps)
        {
            const double si = vv[i];
            for (int j = first + lane; j < nv; j += 32)
                Om[i * ldo + j] *= si * vv[j];
        }
        __syncthreads();
    }
    // Working-set guess: the working set this instance ended its previous solve with (device-resident state, all-lower
    // after configure; `warm` off: empty, which is the cold method — invert H_r, then add bounds one at a time).
    // tools/condensed_model.box_qp_pivot_warm is the specification: T starts as H_r itself (every index in the working
    // set) and only the guessed-free indices are pivoted — |F0| pivots instead of n + |W|; a drop phase removes the
    // wrong-signed multipliers of the guess (it ends on an S-pair whatever the guess was), then the dual iterations add
    // the bounds still violated.  Any guess gives the same minimiser.
    int* actg = reinterpret_cast<int*>(dyn + L.actg);
    double* bc = dyn + L.actg + (L.nvp + 1) / 2;      // two doubles broadcast by the dropping thread
    signed char* wset = wset_all + (size_t)inst * L.nvp;
    int act = 0;            // 0 free, +1 / -1 in the working set at the upper / lower bound
    if (threadIdx.x < nv)
    {
        const int a = (warm && threadIdx.x >= first) ? (int)wset[threadIdx.x] : 0;
        actg[threadIdx.x] = a;
        act = a;
    }
    __syncthreads();
    CwPiv Pv = P;
    Pv.first = first;
    int M = 0;
    int n_piv0 = 0;
    bool okG = true;
    for (int p = first; p < nv; ++p)
        if (actg[p] == 0)
        {
            okG = cw_pivot(Pv, M, p, warp, lane, nwarps, nthr) && okG;
            ++n_piv0;
        }
    if (M > 0)
        cw_flush(Pv, M, warp, lane, nwarps, nthr);
    if (sm.flags[0] != 0 || !okG)
        stat = stat == VSMPC_STATUS_SOLVED ? VSMPC_STATUS_NUMERICAL : stat;
    // sub-problem on the guessed set, one matrix-vector product (scaled variables): out = T (-g_F, b_W) = v on the free
    // set, y = H_r v on the working set; afterwards grad holds the inverse scale factors for all threads
    const double iS_e = 1.0 / S_e;
    const double lo_e = lo * iS_e, up_e = up * iS_e;    // bounds of the scaled variable
    const bool isvar = threadIdx.x >= first && threadIdx.x < nv;
    const double g_e = isvar ? grad[threadIdx.x] : 0.0;
    double b_e = act > 0 ? up_e : lo_e;
    if (threadIdx.x < nv)
        vv[threadIdx.x] = isvar ? (act == 0 ? -g_e : b_e) : 0.0;
    __syncthreads();
    double out_e = 0.0;
    if (isvar)
    {
        const double* row = Om + threadIdx.x * ldo;
        for (int j = first; j < nv; ++j)
            out_e = fma(row[j], vv[j], out_e);
    }
    __syncthreads();
    double* iscl = grad;          // 1 / S for all threads
    if (threadIdx.x < nv)
        iscl[threadIdx.x] = iS_e;
    // drop phase: while a multiplier of the guessed working set has the wrong sign, pivot the worst one out.  The pivot only
    // relabels input a (was v_a = b_a) and output a (was y_a); then the new input y_a moves to its free-variable value -g_a
    // and every output follows column a of the new transform: no second matrix-vector product.
    double lam_e = 0.0;
    int iters = 0;
    bool fail = stat != VSMPC_STATUS_SOLVED;
    while (!fail)
    {
        lam_e = (isvar && act != 0) ? -(double)act * (out_e + g_e) : 0.0;      // multiplier of the scaled variable
        int a;
        const double worst = cw_block_best<true>(fmax(-lam_e * S_e, 0.0), rbA, warp, lane, nwarps, a);
        if (!(worst > 1e-10))
            break;
        if (++iters > 4 * nv + 64)
        {
            stat = VSMPC_STATUS_MAX_ITER;
            fail = true;
            break;
        }
        if (threadIdx.x == a)
        {
            bc[0] = -g_e - out_e;
            bc[1] = b_e;
        }
        if (!cw_pivot(Pv, M, a, warp, lane, nwarps, nthr))
        {
            stat = VSMPC_STATUS_NUMERICAL;
            fail = true;
            break;
        }
        const double delta = bc[0];
        const double col_e = isvar ? cw_teff(Pv, M, threadIdx.x, a) : 0.0;
        if (threadIdx.x == a)
        {
            out_e = fma(col_e, delta, bc[1]);
            act = 0;
        }
        else
            out_e = fma(col_e, delta, out_e);
    }
    lam_e = fmax(lam_e, 0.0);
    if (threadIdx.x < nv)
        vv[threadIdx.x] = isvar ? (act == 0 ? out_e : b_e) : 0.0;
    __syncthreads();

    WCLK(4);
    // ---- Goldfarb-Idnani dual active set on the boxes, one variable per thread -----------------------------------------
    // T = Om[first:nv, first:nv] is kept as the principal pivot transform of H_r over the free set F (W = working set):
    // (v_F, y_W) = T (y_F, v_W) with y = H_r v = -g - sum_a s_a lambda_a e_a.  For a violated free p with sign s, raising
    // its multiplier by t moves v_F by -t s T[F, p] and lambda_a by -t r_a, r_a = -s_a s T[a, p]; T[p, p] is the step
    // denominator.  A full step pivots p into W, a blocked step pivots the blocking index back into F.
    {
        const int e = threadIdx.x;
        const double tol = 1e-10;
        while (!fail)
        {
            double viol = 0.0;
            if (isvar && act == 0)
            {
                const double v = vv[e];
                viol = S_e * fmax(fmax(v - up_e, lo_e - v), 0.0);     // measured in the unscaled variable
            }
            int p;
            const double best = cw_block_best<true>(viol, rbA, warp, lane, nwarps, p);
            if (!(best > tol))
                break;
            const double vp0 = vv[p], isp = iscl[p];
            const double s = (vp0 - up * isp > lo * isp - vp0) ? 1.0 : -1.0;
            const double bound = (s > 0 ? up : lo) * isp;
            double lam_p = 0.0;
            while (true)
            {
                if (++iters > 4 * nv + 64)
                {
                    stat = VSMPC_STATUS_MAX_ITER;
                    fail = true;
                    break;
                }
                const double c_e = isvar ? cw_teff(Pv, M, e, p) : 0.0;
                const double zp = cw_teff(Pv, M, p, p);
                const double v_p = vv[p];
                const double r_e = act != 0 ? -(double)act * s * c_e : 0.0;
                int drop;
                const double t1 = cw_block_best<false>((act != 0 && r_e > 0.0) ? fmax(lam_e, 0.0) / r_e : INFINITY, rbB,
                                                       warp, lane, nwarps, drop);
                const double t2 = (zp > 1e-300) ? (s * v_p - s * bound) / zp : INFINITY;
                const double tt = fmin(t1, t2);
                if (!isfinite(tt))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                if (isvar)
                {
                    if (act == 0)
                        vv[e] = fma(-tt * s, c_e, vv[e]);
                    else
                        lam_e -= tt * r_e;
                }
                lam_p += tt;
                const bool full = t2 <= t1;
                const int q = full ? p : drop;
                if (e == q)
                {
                    if (full)
                    {
                        act = s > 0 ? 1 : -1;
                        lam_e = lam_p;
                        vv[e] = bound;      // exactly on the bound
                    }
                    else
                    {
                        act = 0;
                        lam_e = 0.0;
                    }
                }
                if (!cw_pivot(Pv, M, q, warp, lane, nwarps, nthr))
                {
                    stat = VSMPC_STATUS_NUMERICAL;     // non-positive pivot: same value in every thread
                    fail = true;
                    break;
                }
                if (full)
                    break;
            }
        }
        WCLK(5);
#ifdef VSMPC_PHASE_CLOCKS
        {
            const int n_act = __syncthreads_count(act != 0);
            if (threadIdx.x == 0 && inst < 4096)
                g_wide_clk[inst][14] = n_act;
        }
        if (threadIdx.x == 0 && inst < 4096)
        {
            unsigned smid;
            asm("mov.u32 %0, %%smid;" : "=r"(smid));
            g_wide_clk[inst][8] = iters;
            g_wide_clk[inst][15] = smid;
        }
#endif
        // a NaN iterate never shows up as a violated bound: gate it here (the status holds the outputs)
        if (__syncthreads_or(isvar && !isfinite(vv[e])) && stat == VSMPC_STATUS_SOLVED)
            stat = VSMPC_STATUS_NUMERICAL;
        if (threadIdx.x == 0)
            sm.flags[2] = n_piv0 + iters;         // exchange pivots executed: guessed-free set + one per active-set iteration
        // the working set this solve ended with is the guess of the next one
        if (stat == VSMPC_STATUS_SOLVED && e < nv)
            wset[e] = (signed char)(e >= first ? act : 0);
        // theta*: throttle variables, affine 1, held block 0
        if (e < L.ldc)
        {
            double th = 0.0;
            if (e < nv)
                th = (pinned && e < NT) ? sm.cf[QD_VBAR + e] : (act != 0 ? (act > 0 ? up : lo) : S_e * vv[e]);
            else if (e == L.aff)
                th = 1.0;
            theta[e] = th;
        }
    }
    __syncthreads();

    // ---- F_k theta* for every elimination knot: one warp per (knot, row), lanes over the columns ---------------------
    constexpr int FB = 8;     // (knot, row) items per step: their loads are in flight together
    for (int e0 = FB * warp; e0 < Nc * NJ; e0 += FB * nwarps)
    {
        double acc[FB];
        const double* fr[FB];
#pragma unroll
        for (int u = 0; u < FB; ++u)
        {
            const int e = min(e0 + u, Nc * NJ - 1);
            fr[u] = c.ws + (size_t)(e >> 3) * L.stage + L.wsF + (e & 7) * L.ldc + lane;
            acc[u] = 0.0;
        }
        for (int l = 0; l < L.ldc; l += 32)
        {
            const double th = theta[l + lane];
            double v[FB];
#pragma unroll
            for (int u = 0; u < FB; ++u)
                v[u] = fr[u][l];
#pragma unroll
            for (int u = 0; u < FB; ++u)
                acc[u] = fma(v[u], th, acc[u]);
        }
#pragma unroll
        for (int u = 0; u < FB; ++u)
        {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
                acc[u] += __shfl_xor_sync(0xffffffffu, acc[u], o);
            if (lane == 0 && e0 + u < Nc * NJ)
                fth[e0 + u] = acc[u];
        }
    }
    __syncthreads();
    WCLK(6);
    if constexpr (!JL)
    {
        if (warp != 0)
            return;
    }

    // ---- warp 0: forward rollout and outputs ----------------------------------------------------------------------
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
                fwd = cd_forward<CwSm, true>(cfg, sm, c.ws, L.stage, theta, fth, sm.xs, lane, B, inst, z, o, st, sm.ostage, jl,
                                             sm.clamp, c.ws + (size_t)Nc * L.stage, sm.cand,
                                             pass == CW_JL_PLAIN - 1 ? 2 : (pass < 2 * CW_JL_PLAIN + 2 ? 0 : 1));
                again = (fwd == 1 && pass + 1 < CW_JL_PASSES) ? 1 : 0;      // the working set of the joint boxes moved
            }
            else
                fwd = cd_forward(cfg, sm, c.ws, L.stage, theta, fth, sm.xs, lane, B, inst, z, o, st, sm.ostage, jl);
        }
        if (!again)
        {
            if (lane == 0)
            {
                // see vsmpc_qp_condensed.cu: recursion broken down on finite data, or a joint box active that this build does not
                // carry (or whose working set did not settle) -> fallback kernel; outputs and the joint accumulator are held
                // (variableSamplingMPC.cpp:91)
                if (((!solved && all_fin) || fwd != 0) && fb_mode != 0)
                    fb_list[atomicAdd(fb_count, 1)] = inst;
                status[inst] = (solved && fwd != 0) ? VSMPC_STATUS_NUMERICAL : stat;
                n_factor[inst] = pass + 1;
                n_solve[inst] = (solved && fwd == 0) ? 1 : 0;
                n_pivot[inst] = sm.flags[2];
            }
            cd_stage_outputs(o, status, inst, lane, out2, status2);
            if constexpr (JL)
            {
                if (jlset)
                    for (int e = lane; e < CD_JLSET_WORDS; e += 32)
                        jlset[(size_t)inst * CD_JLSET_WORDS + e] = (solved && fwd == 0 && e < Nc) ? sm.clamp[e] : 0u;
            }
            WCLK(7);
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

int condensed_wide_phase_clocks(long long* host, int n)
{
#ifdef VSMPC_PHASE_CLOCKS
    return cudaMemcpyFromSymbol(host, g_wide_clk, sizeof(long long) * 16 * (n < 4096 ? n : 4096)) == cudaSuccess ? 0 : 2;
#else
    (void)host; (void)n;
    return 3;
#endif
}

static size_t cw_smem_bytes(const DeviceConfig& cfg, bool jl = false)
{
    return (jl ? sizeof(CwSmemT<true>) : sizeof(CwSmem)) + (size_t)cw_layout(cfg).total * sizeof(double);
}

// the build that carries the joint boxes itself: up to two column warps (2x the reference knot count), which is as far as the
// fallback kernel — the net behind a working set that does not settle — covers joint-limit rows (vsmpc_create refuses them
// beyond).  Tried at 3x (four column warps, 288 boxes, no net): on the tight-box test workload a tenth of the instances did not
// settle within 24 passes, so the rows stay refused there.
static bool cw_jl_build(const DeviceConfig& cfg)
{
    const CwLayout L = cw_layout(cfg);
    return cfg.use_jl && L.G <= 2 && cfg.Nc <= CD_JLSET_WORDS && cw_smem_bytes(cfg, true) <= (size_t)CW_SMEM_LIMIT;
}

bool condensed_wide_supported(const DeviceConfig& cfg)
{
    if (cfg.nblk < 1 || cfg.N > MAX_ITER || cfg.Nc < 1)
        return false;
    const CwLayout L = cw_layout(cfg);
    return L.G <= CW_MAXG && cw_smem_bytes(cfg) <= (size_t)CW_SMEM_LIMIT;
}

size_t condensed_wide_ws_doubles(const DeviceConfig& cfg)
{
    return (size_t)cfg.Nc * (cw_layout(cfg).stage + CW_WSU);   // the stages, then raw H_uu per joint block (JL builds)
}

size_t condensed_wide_wset_bytes(const DeviceConfig& cfg)
{
    return (size_t)cw_layout(cfg).nvp;     // working set of the last solve per instance: one signed char per throttle variable
}

size_t condensed_wide_scratch_doubles(const DeviceConfig& cfg)
{
    (void)cfg;
    return 4;   // no global scratch beyond the workspace stacks
}

cudaError_t launch_qp_condensed_wide(const DeviceConfig& h_cfg, int B, const double* qd, double* ws, double* scratch,
                                     double* z, double* st, double* out_rows, int* status, int* n_factor, int* n_solve,
                                     int* n_pivot, int want_z, int* fb_list, int* fb_count, int fb_mode, signed char* wset,
                                     int warm, double* out2, int* status2, unsigned* jlset, cudaStream_t s)
{
    static bool attr_a[64] = {}, attr_b[64] = {}, attr_c[64] = {}, attr_d[64] = {}, attr_g[64] = {};
    const CwLayout L = cw_layout(h_cfg);
    const bool jlb = cw_jl_build(h_cfg);
    const size_t smem = cw_smem_bytes(h_cfg, jlb);
    const size_t wsd = condensed_wide_ws_doubles(h_cfg);
    (void)scratch;     // no global scratch beyond the workspace stacks
    const int fbm = fb_list && fb_count ? fb_mode : 0;
    cudaError_t e;
    if (jlb && L.G <= 2)
    {
        if ((e = ensure_dynamic_smem(qp_condensed_wide_kernel<96, 4, true>, CW_SMEM_LIMIT, attr_d)) != cudaSuccess)
            return e;
        qp_condensed_wide_kernel<96, 4, true><<<B, 32 * (1 + L.G), smem, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status,
                                                                             n_factor, n_solve, n_pivot, wsd, want_z, fb_list, fb_count, fbm, wset, warm, out2, status2, jlset);
    }
    else if (L.G <= 2)
    {
        if ((e = ensure_dynamic_smem(qp_condensed_wide_kernel<96, 4, false>, CW_SMEM_LIMIT, attr_a)) != cudaSuccess)
            return e;
        qp_condensed_wide_kernel<96, 4, false><<<B, 32 * (1 + L.G), smem, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status,
                                                                              n_factor, n_solve, n_pivot, wsd, want_z, fb_list, fb_count, fbm, wset, warm, out2, status2, nullptr);
    }
    else if (L.G == 3)
    {
        if ((e = ensure_dynamic_smem(qp_condensed_wide_kernel<128, 2, false>, CW_SMEM_LIMIT, attr_b)) != cudaSuccess)
            return e;
        qp_condensed_wide_kernel<128, 2, false><<<B, 32 * (1 + L.G), smem, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status,
                                                                               n_factor, n_solve, n_pivot, wsd, want_z, fb_list, fb_count, fbm, wset, warm, out2, status2, nullptr);
    }
    else if (L.G == 4 && 2 * (smem + 1024) <= (size_t)CW_SMEM_LIMIT + 1024)
    {
        // four column warps and a CTA small enough for two per SM (3x the reference knot count, with the deferred-pivot stacks
        // in the static part): five warps per CTA, 168 registers (launch bounds of 192 threads: 65 536 / 384), no idle warps
        if ((e = ensure_dynamic_smem(qp_condensed_wide_kernel<192, 2, false>, CW_SMEM_LIMIT, attr_g)) != cudaSuccess)
            return e;
        qp_condensed_wide_kernel<192, 2, false><<<B, 32 * (1 + L.G), smem, s>>>(h_cfg, B, qd, ws, z, st, out_rows, status,
                                                                               n_factor, n_solve, n_pivot, wsd, want_z, fb_list, fb_count, fbm, wset, warm, out2, status2, nullptr);
    }
    else
    {
        if ((e = ensure_dynamic_smem(qp_condensed_wide_kernel<CW_MAXTHREADS, 1, false>, CW_SMEM_LIMIT, attr_c)) != cudaSuccess)
            return e;
        // eight warps whatever G: the block-wide phases (tensor-core contractions, pivots, active set) are bound by the
        // per-sub-partition FP64 / shared-memory throughput, which 5-7 warps load unevenly
        qp_condensed_wide_kernel<CW_MAXTHREADS, 1, false><<<B, CW_MAXTHREADS, smem, s>>>(
            h_cfg, B, qd, ws, z, st, out_rows, status, n_factor, n_solve, n_pivot, wsd, want_z, fb_list, fb_count, fbm, wset, warm, out2, status2, nullptr);
    }
    return cudaGetLastError();
}

} // namespace vsmpc
