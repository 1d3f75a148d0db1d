// K2f — fallback QP kernel: pivoted LU of the equality-constrained KKT system in a bordered-band ordering.
//
// Replaces, for the instances the Riccati kernels give up on, the same reference code as they do (IMPCProblem::solve ->
// OsqpEigen::Solver, MPC/src/IMPCProblem/IMPCProblem.cpp:196-298; output extraction of VariableSamplingMPC::solveMPC,
// variableSamplingMPC.cpp:88-112).  tools/kkt_lu_model.py is the executable NumPy specification.
//
// Why.  The Riccati recursion is a block elimination without pivoting in backward time order.  When the open-loop
// transition T_k = I + dt_k A_c expands — a vehicle spinning at |omega_B| >~ 30 rad/s makes the explicit-Euler momentum
// block I - dt S(omega_B) grow by |1 + i dt omega| ~ 10 per coarse knot — the cost-to-go grows by that factor squared per
// knot and the rank-8 down-dates cancel catastrophically: non-positive pivots, status 2 (0.2 % of the Monte Carlo loops of
// BASELINE configs[2]; profiles/r02_nonsolved_adjudication.md).  The QP stays well posed — the oracle's pivoted sparse-KKT
// solve and the reference's OSQP both return its minimiser — so the condensed kernels append such an instance to a device
// list and this kernel solves it with ROW pivoting, which decides per unknown whether a dynamics row is resolved forward
// (for x_{k+1}) or backward (for x_k).
//
// Ordering  [ mu | stage 0 | ... | stage N | border ],  stage k = [ x_k | dq_k | v_b(k) | nu_k ]:  half bandwidth 64;
// the border (<= 20 unknowns) holds the input blocks acting over several knots and the pin rows of throttle block 0.
// One CTA per listed instance: the matrix [K | rhs, e_v...] lives in global scratch (dense storage, only the band window and
// the border are touched), elimination column by column with the pivot searched inside the band window, back-substitution
// with one warp per right-hand side, then the SAME Goldfarb-Idnani dual active set as the condensed kernels on
// T = (K^-1)_vv (= the fully exchanged principal pivot transform of the reduced throttle Hessian) and
// z = z_unc - sum_a s_a lam_a K^-1 e_a.
// With the optional joint-limit rows the boxes of the joint increments join the throttle boxes in that active set: this kernel is
// the net behind the condensed kernels' own working set on the joint boxes when it does not settle (vsmpc_qp_condensed.cu, JL
// builds), and the only carrier of the rows where the long-horizon kernel has no JL build.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>

#include "vsmpc_common.cuh"

namespace vsmpc
{

constexpr int FB_THREADS = 512;   // 16 warps: the rank-1 updates of a column are spread over the rows of the band window
constexpr int FB_MAXBOX = 256;    // boxed variables (throttle + joint increments with the joint-limit rows)
constexpr size_t FB_WINDOW_LIMIT = 200 * 1024;   // dynamic shared memory for the elimination window (227 KB per CTA on sm_100, ~15 KB static)

struct FbLayout
{
    int n, nb, bw, nv, ld;        // unknowns, band part, half bandwidth, throttle variables, leading dimension of [K | R]
    int nq, nbox;                 // joint-increment variables; boxed variables at most (nv, + nq with the joint-limit rows)
    int o_x, o_dq, o_v, o_nu, o_mu, o_pin, n_pos;   // offsets into the position table
};

static void fb_host_layout(const DeviceConfig& g, FbLayout& L, std::vector<int>& pos)
{
    const int N = g.N, Nc = g.Nc, Ns = g.Ns, nblk = g.nblk;
    std::vector<int> span_j(Nc, 0), span_t(nblk, 0);
    for (int k = 0; k < N; ++k)
    {
        span_j[joint_block(k, Nc)]++;
        span_t[throttle_block(k, Ns, Nc)]++;
    }
    L.o_x = 0;
    L.o_dq = L.o_x + (N + 1) * NX;
    L.o_v = L.o_dq + Nc * NJ;
    L.o_nu = L.o_v + nblk * NT;
    L.o_mu = L.o_nu + N * NX;
    L.o_pin = L.o_mu + NX;
    L.n_pos = L.o_pin + NT;
    pos.assign(L.n_pos, -1);
    int o = 0;
    for (int i = 0; i < NX; ++i)
        pos[L.o_mu + i] = o++;
    for (int k = 0; k <= N; ++k)
    {
        for (int i = 0; i < NX; ++i)
            pos[L.o_x + k * NX + i] = o++;
        if (k < N)
        {
            const int j = joint_block(k, Nc), b = throttle_block(k, Ns, Nc);
            if (span_j[j] == 1)
                for (int a = 0; a < NJ; ++a)
                    pos[L.o_dq + j * NJ + a] = o++;
            if (span_t[b] == 1 && b != 0)
                for (int a = 0; a < NT; ++a)
                    pos[L.o_v + b * NT + a] = o++;
            for (int i = 0; i < NX; ++i)
                pos[L.o_nu + k * NX + i] = o++;
        }
    }
    L.nb = o;
    for (int e = 0; e < Nc * NJ; ++e)
        if (pos[L.o_dq + e] < 0)
            pos[L.o_dq + e] = o++;
    for (int e = 0; e < nblk * NT; ++e)
        if (pos[L.o_v + e] < 0)
            pos[L.o_v + e] = o++;
    for (int a = 0; a < NT; ++a)
        pos[L.o_pin + a] = o++;
    L.n = o;
    L.nv = NT * nblk;
    L.nq = NJ * Nc;
    L.nbox = L.nv + (g.use_jl ? L.nq : 0);
    L.ld = (L.n + 1 + L.nbox + 3) & ~3;
    // half bandwidth of the band part: a dynamics row of knot k reaches from x_k to x_{k+1} (and the in-stage inputs)
    int bw = 0;
    for (int k = 0; k < N; ++k)
        for (int i = 0; i < NX; ++i)
        {
            const int row = pos[L.o_nu + k * NX + i];
            auto reach = [&](int p) { if (p < L.nb) bw = std::max(bw, std::abs(row - p)); };
            for (int j = 0; j < NX; ++j)
                reach(pos[L.o_x + k * NX + j]);
            reach(pos[L.o_x + (k + 1) * NX + i]);
            for (int a = 0; a < NJ; ++a)
                reach(pos[L.o_dq + joint_block(k, Nc) * NJ + a]);
            for (int a = 0; a < NT; ++a)
                reach(pos[L.o_v + throttle_block(k, Ns, Nc) * NT + a]);
        }
    for (int i = 0; i < NX; ++i)
        bw = std::max(bw, std::abs(pos[L.o_mu + i] - pos[L.o_x + i]));
    for (int b = 0; b + 1 < nblk; ++b)      // throttle Laplacian between in-band blocks
        for (int a = 0; a < NT; ++a)
        {
            const int p = pos[L.o_v + b * NT + a], q = pos[L.o_v + (b + 1) * NT + a];
            if (p < L.nb && q < L.nb)
                bw = std::max(bw, std::abs(p - q));
        }
    L.bw = bw;
}

__device__ __forceinline__ void fb_set(double* __restrict__ M, int ld, int r, int c, double v)
{
    M[(size_t)r * ld + c] = v;
    M[(size_t)c * ld + r] = v;
}

// per-slot scratch: M [n][ld] | T [nbox][nbox] | zs [n]
__global__ void __launch_bounds__(FB_THREADS)
qp_fallback_kernel(const __grid_constant__ DeviceConfig cfgv, const FbLayout L, int B, const double* __restrict__ qd_all,
                   const int* __restrict__ fb_list, const int* __restrict__ fb_count, const int* __restrict__ pos,
                   double* __restrict__ scratch, size_t slot_doubles, double* __restrict__ z_all, double* __restrict__ st,
                   double* __restrict__ out_rows, int* __restrict__ status, int* __restrict__ n_factor,
                   int* __restrict__ n_solve, int* __restrict__ n_pivot, int want_z, int use_window, double* __restrict__ out2,
                   int* __restrict__ status2)
{
    const DeviceConfig& cfg = cfgv;
    __shared__ double A[NX * NX], BJ[NX * NJ], BT[NX * NT], cv[NX];
    extern __shared__ double fb_dyn[];               // the sliding elimination window, (bw + 1 + border) x (2 bw + 1 + border + rhs)
    __shared__ double vbuf[FB_MAXBOX], cbuf[FB_MAXBOX], rbuf[FB_MAXBOX];   // active set: one boxed variable per thread
    __shared__ double red_v[FB_THREADS / 32];
    __shared__ int red_i[FB_THREADS / 32];
    __shared__ double s_val[4];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int count = min(*fb_count, B);
    const int n = L.n, nb = L.nb, bw = L.bw, ld = L.ld, N = cfg.N, Nc = cfg.Nc, nblk = cfg.nblk;
    double* M = scratch + (size_t)blockIdx.x * slot_doubles;
    double* T = M + (size_t)n * ld;
    double* zs = T + (size_t)L.nbox * L.nbox;
    const int* px = pos + L.o_x;
    const int* pq = pos + L.o_dq;
    const int* pv = pos + L.o_v;
    const int* pnu = pos + L.o_nu;
    const int* pmu = pos + L.o_mu;
    const int* ppin = pos + L.o_pin;

#ifdef VSMPC_FB_CLOCKS
#define FBCLK(name) do { if (tid == 0 && blockIdx.x == 0) { const long long t__ = clock64(); printf("fallback %-28s %10lld cycles\n", name, t__ - fbt0); fbt0 = t__; } } while (0)
    long long fbt0 = clock64();
#else
#define FBCLK(name) do { } while (0)
#endif
    for (int li = blockIdx.x; li < count; li += gridDim.x)
    {
        const int inst = fb_list[li];
        const double* qd = qd_all + (size_t)inst * cfg.qd_stride;
        const bool pinned = qd[QD_PINNED] != 0.0;
        const int first = pinned ? NT : 0;
        const int nvf = L.nv - first;
        // boxed variables: the free throttle variables and, with the optional joint-limit rows on, every joint increment
        // (box e < nqb: joint variable e, bounds QD_JLO / QD_JHI; else throttle variable first + e - nqb)
        const int nqb = (cfg.use_jl && qd[QD_JLIM] != 0.0) ? L.nq : 0;
        const int nbx = nqb + nvf, nrhs = 1 + nbx;
        const int ncol = n + nrhs;             // columns in use (<= ld)
        if (tid == 0)
            expand_dense(qd, A, BJ, BT, cv);
        // ---- zero the band window, the border and the right-hand sides --------------------------------------------------
        for (int r = warp; r < n; r += FB_THREADS / 32)
        {
            double* row = M + (size_t)r * ld;
            if (r < nb)
            {
                const int c0 = max(0, r - bw), c1 = min(nb - 1, r + 2 * bw);
                for (int c = c0 + lane; c <= c1; c += 32)
                    row[c] = 0.0;
                for (int c = nb + lane; c < ncol; c += 32)
                    row[c] = 0.0;
            }
            else
                for (int c = lane; c < ncol; c += 32)
                    row[c] = 0.0;
        }
        __syncthreads();
        // ---- assemble K and the right-hand sides (tools/kkt_lu_model.assemble_kkt) ---------------------------------------
        for (int e = tid; e < N * NX; e += FB_THREADS)
        { // tracking cost on x_1 .. x_N: K = Q, rhs = Q xref
            const int k = 1 + e / NX, i = e - (k - 1) * NX;
            const int p = px[k * NX + i];
            M[(size_t)p * ld + p] = cfg.Qd[i];
            M[(size_t)p * ld + n] = i < 12 ? cfg.Qd[i] * qd[QD_XREF + i * cfg.NC + ref_col(k - 1, cfg.Ns)] : 0.0;
        }
        for (int e = tid; e < Nc * NJ; e += FB_THREADS)
        {
            const int p = pq[e];
            M[(size_t)p * ld + p] = cfg.Rqd[e % NJ];
            M[(size_t)p * ld + n] = -qd[QD_GQ + e % NJ];
            if (nqb)
                M[(size_t)p * ld + n + 1 + e] = 1.0;                 // unit right-hand side of joint variable e
        }
        for (int e = tid; e < L.nv; e += FB_THREADS)
        {
            const int b = e / NT, p = pv[e];
            M[(size_t)p * ld + p] = cfg.w_t * ((b > 0 ? 1.0 : 0.0) + (b < nblk - 1 ? 1.0 : 0.0)) + (b == 0 ? cfg.w_i : 0.0);
            if (b > 0)
                M[(size_t)p * ld + pv[e - NT]] = -cfg.w_t;
            if (b < nblk - 1)
                M[(size_t)p * ld + pv[e + NT]] = -cfg.w_t;
            if (b == 0)
                M[(size_t)p * ld + n] = cfg.w_i * qd[QD_VBAR + e];
            if (e >= first)
                M[(size_t)p * ld + n + 1 + nqb + (e - first)] = 1.0;   // unit right-hand side of throttle variable e
        }
        for (int e = tid; e < N * NX; e += FB_THREADS)
        { // dynamics rows: T x_k - x_{k+1} + dt B_J dq + dt B_T v = -dt c
            const int k = e / NX, i = e - k * NX;
            const double dt = cfg.dt[k];
            const int row = pnu[e];
            for (int j = 0; j < NX; ++j)
            {
                const double t = (i == j ? 1.0 : 0.0) + dt * A[i * NX + j];
                if (t != 0.0)
                    fb_set(M, ld, row, px[k * NX + j], t);
            }
            fb_set(M, ld, row, px[(k + 1) * NX + i], -1.0);
            const int jb = joint_block(k, Nc), tb = throttle_block(k, cfg.Ns, Nc);
            for (int a = 0; a < NJ; ++a)
                if (BJ[i * NJ + a] != 0.0)
                    fb_set(M, ld, row, pq[jb * NJ + a], dt * BJ[i * NJ + a]);
            for (int a = 0; a < NT; ++a)
                if (BT[i * NT + a] != 0.0)
                    fb_set(M, ld, row, pv[tb * NT + a], dt * BT[i * NT + a]);
            M[(size_t)row * ld + n] = -dt * cv[i];
        }
        if (tid < NX)
        {
            fb_set(M, ld, pmu[tid], px[tid], 1.0);
            M[(size_t)pmu[tid] * ld + n] = qd[QD_X0 + tid];
        }
        if (tid < NT)
        {
            if (pinned)
            {
                fb_set(M, ld, ppin[tid], pv[tid], 1.0);
                M[(size_t)ppin[tid] * ld + n] = qd[QD_VBAR + tid];
            }
            else
                M[(size_t)ppin[tid] * ld + ppin[tid]] = 1.0;
        }
        __syncthreads();
        FBCLK("zero + assemble");
        // ---- elimination with row pivoting inside the band window ----------------------------------------------------------
        // use_window: the window lives in SHARED memory — rows k .. k + bw of the band part and the border rows, columns
        // k .. k + 2 bw of the band part (circular slots c mod (2 bw + 1)), the border columns and the right-hand sides: 85 x 175
        // doubles at the reference horizon.  Per column: every warp finds the pivot by itself (no block reduction), updates
        // its rows against the pivot row where it lies (row interchanges are entries of a slot map, not copies), one barrier;
        // then the finished row goes to global memory for the back-substitution and the next band row / column (original
        // data, not touched so far, prefetched at the start of the step) slides into the freed slot, second barrier.
        // Otherwise (window larger than the shared memory: joint-limit rows on a long horizon) the matrix is updated in global
        // memory.  profiles/r02_fallback.md has the times of the versions.
        bool ok = true;
        if (use_window)
        {
            const int WB = 2 * bw + 1, RB = bw + 1, nbord = n - nb;
            const int wcols = WB + nbord + nrhs;            // column slots in use
            const int WCp = (WB + nbord + 1 + L.nbox) | 1;  // leading dimension (odd), sized for the largest right-hand side
            double* W = fb_dyn;
            int* smap = reinterpret_cast<int*>(fb_dyn + (size_t)(RB + nbord) * WCp);   // physical slot of band row r: smap[r % RB]
            // initial window: band rows 0 .. bw, border rows; band columns 0 .. 2 bw
            for (int e = tid; e < (RB + nbord) * wcols; e += FB_THREADS)
            {
                const int rs = e / wcols, cs = e - rs * wcols;
                const int r = rs < RB ? rs : nb + (rs - RB);
                const int c = cs < WB ? cs : nb + (cs - WB);
                W[rs * WCp + cs] = ((cs >= WB || c < nb) && (rs >= RB || r < nb)) ? M[(size_t)r * ld + c] : 0.0;
            }
            if (tid < RB)
                smap[tid] = tid;
            __syncthreads();
#ifdef VSMPC_FB_CLOCKS
            long long cA = 0, cB = 0, cC = 0, cD = 0, ct = clock64();
#define FBSUB(acc) do { const long long t__ = clock64(); acc += t__ - ct; ct = t__; } while (0)
#else
#define FBSUB(acc) do { } while (0)
#endif
            // residues kept incrementally: an integer division per index costs more than the arithmetic it serves here
            int kWB = 0, kRB = 0;     // k mod WB, k mod RB
            for (int k = 0; k < n && ok; ++k, kWB = kWB + 1 == WB ? 0 : kWB + 1, kRB = kRB + 1 == RB ? 0 : kRB + 1)
            {
                const bool band = k < nb;
                const int hi = band ? min(k + bw, nb - 1) : n - 1;
                const int csk = band ? kWB : WB + (k - nb);              // column slot of column k
                // prefetch what slides in at the end of this step (original entries, not touched by the elimination so far)
                const int rn = k + bw + 1, cn = k + 2 * bw + 1;
                double pf_row = 0.0, pf_col = 0.0;
                if (band && rn < nb && tid < wcols)
                {
                    int c;      // slot tid of the new row: the band column of [k + 1, k + 2 bw + 1] with c mod WB = tid, or border / rhs
                    if (tid < WB)
                    {
                        const int k1 = kWB + 1 == WB ? 0 : kWB + 1;          // (k + 1) mod WB
                        const int off = tid - k1;
                        c = k + 1 + (off < 0 ? off + WB : off);
                        c = c < nb ? c : -1;
                    }
                    else
                        c = nb + (tid - WB);
                    pf_row = c >= 0 ? M[(size_t)rn * ld + c] : 0.0;
                }
                if (band && cn < nb && tid >= FB_THREADS - nbord)
                    pf_col = M[(size_t)(nb + (tid - (FB_THREADS - nbord))) * ld + cn];
                // pivot: largest |W[r][k]|, r in [k, hi]; every warp by itself
                double best = -1.0;
                int arg = k;
                for (int r = k + lane; r <= hi; r += 32)
                {
                    int rr = kRB + (r - k);                                  // r mod RB: r - k <= bw < RB
                    rr = rr >= RB ? rr - RB : rr;
                    const int rs = band ? smap[rr] : RB + (r - nb);
                    const double a = fabs(W[rs * WCp + csk]);
                    if (a > best) { best = a; arg = r; }
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                }
                if (!(best > 0.0) || !isfinite(best))
                {
                    ok = false;
                    break;
                }
                FBSUB(cA);
                // physical slots of the pivot row and of the row that sat at position k; in the border phase rows are
                // interchanged by copying (20 rows at most)
                int ps, ks;
                int aRB = kRB + (arg - k);
                aRB = aRB >= RB ? aRB - RB : aRB;                            // arg mod RB
                if (band)
                {
                    ps = smap[aRB];
                    ks = smap[kRB];
                }
                else
                {
                    ps = RB + (k - nb);
                    ks = ps;
                    const int as = RB + (arg - nb);
                    if (arg != k)
                    {
                        __syncthreads();
                        if (tid < wcols)
                        {
                            const double a = W[as * WCp + tid];
                            W[as * WCp + tid] = W[ps * WCp + tid];
                            W[ps * WCp + tid] = a;
                        }
                        __syncthreads();
                    }
                }
                const double* __restrict__ prw = W + ps * WCp;
                const double ipiv = 1.0 / prw[csk];
                // rank-1 update of the other rows of the window and of the border rows: one warp per row, the pivot row's
                // entries of this lane's columns in registers (the window fills in within a few steps: nearly every row has a
                // non-zero in column k, the update is bound by the shared-memory bytes of the rows)
                constexpr int PC = 16;      // column slots per lane: wcols <= 512
                double pr[PC];
#pragma unroll
                for (int t = 0; t < PC; ++t)
                {
                    const int cs = lane + 32 * t;
                    pr[t] = (cs < wcols && cs != csk) ? prw[cs] : 0.0;
                }
                const int nr1 = band ? RB : hi - k, nr2 = band ? nbord : 0;
                for (int e = warp; e < nr1 + nr2; e += FB_THREADS / 32)
                {
                    const int rs = band ? (e < nr1 ? e : RB + (e - nr1)) : RB + (k + 1 + e - nb);
                    if (rs == ps)
                        continue;
                    double* __restrict__ row = W + rs * WCp;
                    const double l = row[csk] * ipiv;
                    if (l != 0.0)
                    {
#pragma unroll
                        for (int t = 0; t < PC; ++t)
                        {
                            const int cs = lane + 32 * t;
                            if (cs < wcols)
                                row[cs] = fma(-l, pr[t], row[cs]);
                        }
                    }
                }
                FBSUB(cB);
                __syncthreads();
                FBSUB(cC);
                // the finished row k: to global memory (row k of U and its right-hand sides) for the back-substitution
                if (tid < wcols)
                {
                    int c;
                    if (tid < WB)
                    {
                        const int off = tid - csk;
                        c = band ? k + (off < 0 ? off + WB : off) : -1;
                        c = c < nb ? c : -1;
                    }
                    else
                        c = nb + (tid - WB);
                    if (c >= k)
                        M[(size_t)k * ld + c] = prw[tid];
                }
                if (band)
                {
                    // slide: column k leaves (its slot takes column k + 2 bw + 1: zero in the band rows of the window, original
                    // entries in the border rows), row k leaves (the pivot row's slot takes row k + bw + 1)
                    if (tid < RB && tid != ps)
                        W[tid * WCp + csk] = 0.0;
                    if (tid >= FB_THREADS - nbord)
                        W[(RB + tid - (FB_THREADS - nbord)) * WCp + csk] = cn < nb ? pf_col : 0.0;
                    if (tid < wcols)
                        W[ps * WCp + tid] = rn < nb ? pf_row : 0.0;
                    if (tid == 0)
                    {
                        smap[aRB] = ks;           // the row that sat at position k now sits at position arg
                        smap[kRB] = ps;           // position k + bw + 1 (same residue) gets the freed slot
                    }
                    __syncthreads();
                }
                FBSUB(cD);
            }
#ifdef VSMPC_FB_CLOCKS
            if (tid == 0 && blockIdx.x == 0)
                printf("fallback elimination: prefetch + pivot search %lld, update %lld, barrier %lld, write + slide %lld cycles\n", cA, cB, cC, cD);
#endif
        }
        else
        {
            double* prow = fb_dyn;                                        // pivot row cache
            double* lmul = prow + 2 * bw + 1 + (n - nb) + 1 + L.nbox + 8;
            int* lrow = reinterpret_cast<int*>(lmul + bw + (n - nb) + 8);
                for (int k = 0; k < n && ok; ++k)
            {
                const int hi = k < nb ? min(k + bw, nb - 1) : n - 1;
                // pivot: largest |M[r][k]|, r in [k, hi]
                double best = -1.0;
                int arg = k;
                for (int r = k + tid; r <= hi; r += FB_THREADS)
                {
                    const double a = fabs(M[(size_t)r * ld + k]);
                    if (a > best) { best = a; arg = r; }
                }
    #pragma unroll
                for (int o = 16; o > 0; o >>= 1)
                {
                    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
                    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
                    if (ob > best || (ob == best && oa < arg)) { best = ob; arg = oa; }
                }
                if (lane == 0) { red_v[warp] = best; red_i[warp] = arg; }
                __syncthreads();
                best = red_v[0]; arg = red_i[0];
                for (int w = 1; w < FB_THREADS / 32; ++w)
                    if (red_v[w] > best || (red_v[w] == best && red_i[w] < arg)) { best = red_v[w]; arg = red_i[w]; }
                if (!(best > 0.0) || !isfinite(best))
                {
                    ok = false;
                    break;
                }
                // columns touched by this step: (k, cmax] in the band part, then the border and the right-hand sides
                const int cmax = k < nb ? min(k + 2 * bw, nb - 1) : n - 1;
                const int nc1 = cmax - k;                       // band columns k+1 .. cmax
                const int c2 = k < nb ? nb : n;                 // first column of the second segment
                const int c2s = max(c2, k + 1);
                const int nc2 = ncol - c2s;
                const int nct = nc1 + nc2;
                double* rk = M + (size_t)k * ld;
                double* rp = M + (size_t)arg * ld;
                const double piv = rp[k];
                // swap rows k <-> arg over the touched columns, cache the pivot row
                for (int e = tid; e < nct + 1; e += FB_THREADS)
                {
                    const int c = e == nct ? k : (e < nc1 ? k + 1 + e : c2s + (e - nc1));
                    const double a = rp[c];
                    if (arg != k)
                    {
                        rp[c] = rk[c];
                        rk[c] = a;
                    }
                    if (e < nct)
                        prow[e] = a;
                }
                // rows to eliminate: band window below k, then the border rows
                const int nr1 = hi - k;
                const int r2s = k < nb ? nb : n;
                const int nr2 = k < nb ? n - nb : 0;
                __syncthreads();
                const double ipiv = 1.0 / piv;
                for (int e = tid; e < nr1 + nr2; e += FB_THREADS)
                {
                    const int r = e < nr1 ? k + 1 + e : r2s + (e - nr1);
                    const double a = M[(size_t)r * ld + k];
                    lmul[e] = a * ipiv;
                    lrow[e] = a != 0.0 ? r : -1;
                }
                __syncthreads();
                // rank-1 update: one warp per row, lanes over the columns
                for (int e = warp; e < nr1 + nr2; e += FB_THREADS / 32)
                {
                    const int r = lrow[e];
                    if (r < 0)
                        continue;
                    const double l = lmul[e];
                    double* __restrict__ row = M + (size_t)r * ld;
                    // four columns per lane in flight: the loads of a row are independent, the matrix lives in L2 / HBM
                    for (int q0 = lane; q0 < nct; q0 += 128)
                    {
                        double v[4];
                        int cc[4];
    #pragma unroll
                        for (int u = 0; u < 4; ++u)
                        {
                            const int q = q0 + 32 * u;
                            cc[u] = q < nct ? (q < nc1 ? k + 1 + q : c2s + (q - nc1)) : -1;
                            v[u] = cc[u] >= 0 ? row[cc[u]] : 0.0;
                        }
    #pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (cc[u] >= 0)
                                row[cc[u]] = fma(-l, prow[q0 + 32 * u], v[u]);
                    }
                }
                __syncthreads();
            }
        }
        if (!ok)
        { // structurally or numerically singular: the instance keeps status 2 and its held outputs
            __syncthreads();
            continue;
        }
        FBCLK("elimination");
        // ---- back-substitution: one warp per right-hand side, X overwrites the right-hand-side columns --------------------
        if (use_window)
        {
            // the part of x a row needs (2 bw band entries behind it, the border) sits in shared memory per warp, circular like
            // the elimination window; row k - 1 of U is loaded while row k is reduced.  (Reading x back through global memory
            // cost a dependent L2 round trip per row: 3.0 of the 7.7 ms of a fallback solve.)
            const int WB = 2 * bw + 1, nbord = n - nb;
            const int XL = WB + nbord;
            constexpr int UE = 8;      // entries of a U row per lane: 2 bw + border <= 32 UE
            __syncthreads();
            for (int j = warp; j < nrhs; j += 2 * (FB_THREADS / 32))
            {
                // two right-hand sides per warp and pass (they share the loads of the rows of U)
                const int j2 = j + FB_THREADS / 32;
                const bool two = j2 < nrhs;
                const int cj = n + j, cj2 = n + (two ? j2 : j);
                double* xw = fb_dyn + (size_t)(2 * warp) * XL;
                double* xw2 = xw + XL;
                auto load_row = [&](int k, double (&u)[UE], double& dk, double& bk, double& bk2) {
                    const double* rk = M + (size_t)k * ld;
                    const int nbnd = k < nb ? min(2 * bw, nb - 1 - k) : 0;          // band entries behind the diagonal
                    const int ntot = k < nb ? nbnd + nbord : n - 1 - k;
#pragma unroll
                    for (int t = 0; t < UE; ++t)
                    {
                        const int e = lane + 32 * t;
                        const int c = k < nb ? (e < nbnd ? k + 1 + e : nb + (e - nbnd)) : k + 1 + e;
                        u[t] = e < ntot ? rk[c] : 0.0;
                    }
                    dk = rk[k];
                    bk = rk[cj];
                    bk2 = rk[cj2];
                };
                double u[UE], un[UE], dk, bk, bk2, dn = 1.0, bn = 0.0, bn2 = 0.0;
                load_row(n - 1, u, dk, bk, bk2);
                int k1WB = nb % WB;        // (k + 1) mod WB for the band rows, kept incrementally from k = nb - 1 down
                for (int k = n - 1; k >= 0; --k)
                {
                    if (k > 0)
                        load_row(k - 1, un, dn, bn, bn2);
                    const int nbnd = k < nb ? min(2 * bw, nb - 1 - k) : 0;
                    const int ntot = k < nb ? nbnd + nbord : n - 1 - k;
                    double acc = 0.0, acc2 = 0.0;
#pragma unroll
                    for (int t = 0; t < UE; ++t)
                    {
                        const int e = lane + 32 * t;
                        if (e < ntot)
                        {
                            int xs;      // slot of x[c] in this warp's windows
                            if (k < nb)
                            {
                                const int b = k1WB + e;                     // (k + 1 + e) mod WB, e < 2 bw < WB
                                xs = e < nbnd ? (b >= WB ? b - WB : b) : WB + (e - nbnd);
                            }
                            else
                                xs = WB + (k + 1 + e - nb);
                            acc = fma(u[t], xw[xs], acc);
                            acc2 = fma(u[t], xw2[xs], acc2);
                        }
                    }
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                    {
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                        acc2 += __shfl_xor_sync(0xffffffffu, acc2, o);
                    }
                    const double idk = 1.0 / dk;
                    const double xk = (bk - acc) * idk, xk2 = (bk2 - acc2) * idk;
                    if (k < nb)
                        k1WB = k1WB == 0 ? WB - 1 : k1WB - 1;           // now k mod WB = the (k + 1) mod WB of the next row
                    if (lane == 0)
                    {
                        const int xs = k < nb ? k1WB : WB + (k - nb);
                        xw[xs] = xk;
                        xw2[xs] = xk2;
                        M[(size_t)k * ld + cj] = xk;
                        if (two)
                            M[(size_t)k * ld + cj2] = xk2;
                    }
                    __syncwarp();
#pragma unroll
                    for (int t = 0; t < UE; ++t)
                        u[t] = un[t];
                    dk = dn;
                    bk = bn;
                    bk2 = bn2;
                }
            }
        }
        else
            for (int j = warp; j < nrhs; j += FB_THREADS / 32)
            {
                const int cj = n + j;
                for (int k = n - 1; k >= 0; --k)
                {
                    const double* rk = M + (size_t)k * ld;
                    const int cmax = k < nb ? min(k + 2 * bw, nb - 1) : n - 1;
                    double acc = 0.0;
                    for (int c = k + 1 + lane; c <= cmax; c += 32)
                        acc = fma(rk[c], M[(size_t)c * ld + cj], acc);
                    if (k < nb)
                        for (int c = nb + lane; c < n; c += 32)
                            acc = fma(rk[c], M[(size_t)c * ld + cj], acc);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        acc += __shfl_xor_sync(0xffffffffu, acc, o);
                    if (lane == 0)
                        M[(size_t)k * ld + cj] = (rk[cj] - acc) / rk[k];
                    __syncwarp();
                }
            }
        __syncthreads();
        FBCLK("back-substitution");
        // ---- dual active set on T = (K^-1)_vv (tools/condensed_model._dual_pivot_loop), one variable per thread ----------
        const double tol = 1e-10;
        const bool isvar = tid < nbx;
        const int pos_e = isvar ? (tid < nqb ? pq[tid] : pv[first + tid - nqb]) : 0;     // KKT position of this thread's variable
        const double lo = !isvar ? 0.0 : (tid < nqb ? qd[QD_JLO + tid % NJ] : qd[QD_VMIN]);
        const double up = !isvar ? 0.0 : (tid < nqb ? qd[QD_JHI + tid % NJ] : qd[QD_VMAX]);
        if (isvar)
            vbuf[tid] = (double)pos_e;
        __syncthreads();
        for (int e = tid; e < nbx * nbx; e += FB_THREADS)
        {
            const int i = e / nbx, j = e - i * nbx;
            T[e] = 0.5 * (M[(size_t)(int)vbuf[i] * ld + n + 1 + j] + M[(size_t)(int)vbuf[j] * ld + n + 1 + i]);
        }
        double v_e = isvar ? M[(size_t)pos_e * ld + n] : 0.0;
        double lam_e = 0.0;
        int act = 0, iters = 0, stat = VSMPC_STATUS_SOLVED;
        __syncthreads();
        auto block_best = [&](double v, bool want_max, int& arg_out) -> double {
            // exact block-wide max / min with the lowest thread attaining it
            double bv = v;
            int ba = tid;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
            {
                const double ob = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oa = __shfl_xor_sync(0xffffffffu, ba, o);
                if (want_max ? (ob > bv || (ob == bv && oa < ba)) : (ob < bv || (ob == bv && oa < ba))) { bv = ob; ba = oa; }
            }
            __syncthreads();
            if (lane == 0) { red_v[warp] = bv; red_i[warp] = ba; }
            __syncthreads();
            bv = red_v[0]; ba = red_i[0];
            for (int w = 1; w < FB_THREADS / 32; ++w)
                if (want_max ? (red_v[w] > bv || (red_v[w] == bv && red_i[w] < ba)) : (red_v[w] < bv || (red_v[w] == bv && red_i[w] < ba)))
                { bv = red_v[w]; ba = red_i[w]; }
            arg_out = ba;
            return bv;
        };
        auto pivot = [&](int q) -> bool {
            // exchange pivot on q: T' = T - u v'/d off row / column q, T'[q,:] = -v/d, T'[:,q] = u/d, T'[q,q] = 1/d
            __syncthreads();
            if (tid < nbx)
            {
                cbuf[tid] = T[tid * nbx + q];
                rbuf[tid] = T[q * nbx + tid];
            }
            __syncthreads();
            const double d = rbuf[q];
            if (!(d > 0.0) || !isfinite(d))
                return false;
            const double id = 1.0 / d;
            for (int e = tid; e < nbx * nbx; e += FB_THREADS)
            {
                const int i = e / nbx, j = e - i * nbx;
                double t;
                if (i == q)
                    t = j == q ? id : -rbuf[j] * id;
                else if (j == q)
                    t = cbuf[i] * id;
                else
                    t = fma(-cbuf[i] * id, rbuf[j], T[e]);
                T[e] = t;
            }
            __syncthreads();
            return true;
        };
        bool fail = false;
        while (!fail && nbx > 0)
        {
            int p;
            const double best = block_best((isvar && act == 0) ? fmax(fmax(v_e - up, lo - v_e), 0.0) : 0.0, true, p);
            if (!(best > tol))
                break;
            if (tid == p)
            { // sign and bound of the violated variable (its own box)
                s_val[0] = (v_e - up > lo - v_e) ? 1.0 : -1.0;
                s_val[2] = (v_e - up > lo - v_e) ? up : lo;
            }
            __syncthreads();
            const double s = s_val[0];
            const double bound = s_val[2];
            double lam_p = 0.0;
            while (true)
            {
                if (++iters > 6 * L.nbox + 64)
                {
                    stat = VSMPC_STATUS_MAX_ITER;
                    fail = true;
                    break;
                }
                const double c_e = isvar ? T[tid * nbx + p] : 0.0;
                const double zp = T[p * nbx + p];
                const double r_e = act != 0 ? -(double)act * s * c_e : 0.0;
                int drop;
                const double t1 = block_best((act != 0 && r_e > 0.0) ? fmax(lam_e, 0.0) / r_e : INFINITY, false, drop);
                if (tid == p)
                    s_val[1] = v_e;
                __syncthreads();
                const double v_p = s_val[1];
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
                        v_e = fma(-tt * s, c_e, v_e);
                    else
                        lam_e -= tt * r_e;
                }
                lam_p += tt;
                const bool full = t2 <= t1;
                const int q = full ? p : drop;
                if (tid == q)
                {
                    if (full)
                    {
                        act = s > 0 ? 1 : -1;
                        lam_e = lam_p;
                        v_e = bound;
                    }
                    else
                    {
                        act = 0;
                        lam_e = 0.0;
                    }
                }
                if (!pivot(q))
                {
                    stat = VSMPC_STATUS_NUMERICAL;
                    fail = true;
                    break;
                }
                if (full)
                    break;
            }
        }
        FBCLK("active set");
        // ---- z = z_unc - sum_a s_a lam_a K^-1 e_a; active variables exactly on their bound ----------------------------------
        __syncthreads();
        __syncthreads();
        if (isvar)
            vbuf[tid] = act != 0 ? (double)act * lam_e : 0.0;
        __syncthreads();
        bool fin = true;
        for (int r = tid; r < n; r += FB_THREADS)
        {
            const double* row = M + (size_t)r * ld + n;
            double z = row[0];
            for (int a = 0; a < nbx; ++a)
            {
                const double w = vbuf[a];
                if (w != 0.0)
                    z = fma(-w, row[1 + a], z);
            }
            zs[r] = z;
            fin = fin && isfinite(z);
        }
        __syncthreads();
        if (isvar && act != 0)
            zs[pos_e] = act > 0 ? up : lo;
        const bool all_fin = __syncthreads_and(fin);
        if (!all_fin && stat == VSMPC_STATUS_SOLVED)
            stat = VSMPC_STATUS_NUMERICAL;
        FBCLK("z");
        // ---- outputs (variableSamplingMPC.cpp:88-112,138-151) ------------------------------------------------------------
        if (tid == 0)
        {
            status[inst] = stat;
            n_factor[inst] = 2;       // the Riccati attempt + this LU factorisation
            n_solve[inst] = nrhs;     // right-hand sides back-substituted
            n_pivot[inst] = iters;
        }
        if (stat == VSMPC_STATUS_SOLVED)
        {
            double* o = out_rows + (size_t)inst * VSMPC_OUT_DOUBLES;
            if (tid < NJ)
            {
                const double dq = zs[pq[tid]];
                o[VSMPC_OUT_DELTA_Q + tid] = dq;
                const double acc = st[(size_t)(ST_QACC + tid) * B + inst] + dq;
                st[(size_t)(ST_QACC + tid) * B + inst] = acc;
                o[VSMPC_OUT_JOINTS_REF + tid] = acc;
            }
            if (tid < NT)
            {
                o[VSMPC_OUT_THROTTLE + tid] = destd_throttle_qd(qd, zs[pv[tid]]);
                o[VSMPC_OUT_THRUST + tid] = zs[px[NX + IX_T + tid]];
                o[VSMPC_OUT_THRUST_DOT + tid] = zs[px[NX + IX_TD + tid]];
            }
            if (tid < NX)
                o[VSMPC_OUT_FINAL_STATE + tid] = zs[px[N * NX + tid]];
            if (want_z)
            {
                double* z = z_all + (size_t)inst * cfg.n_var;
                for (int e = tid; e < cfg.n_var; e += FB_THREADS)
                    z[e] = zs[pos[L.o_x + e]];     // x | dq | v are consecutive in the position table, in the reference's order
            }
        }
        __syncthreads();
        if (out2)
        { // staging buffers of the asynchronous read-back (cd_stage_outputs): the row as it stands now
            if (tid < VSMPC_OUT_DOUBLES)
                out2[(size_t)inst * VSMPC_OUT_DOUBLES + tid] = out_rows[(size_t)inst * VSMPC_OUT_DOUBLES + tid];
            if (tid == 0)
                status2[inst] = status[inst];
        }
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
struct FallbackPlan
{
    FbLayout L;
    std::vector<int> pos;
    size_t slot_doubles;
};

static FallbackPlan fb_plan(const DeviceConfig& cfg)
{
    FallbackPlan P;
    fb_host_layout(cfg, P.L, P.pos);
    P.slot_doubles = (size_t)P.L.n * P.L.ld + (size_t)P.L.nbox * P.L.nbox + P.L.n + 8;
    return P;
}

static size_t fb_window_bytes(const FbLayout& L)
{ // (bw + 1 + border rows) x (2 bw + 1 + border + 1 + nbox | 1) doubles + the slot map, as the kernel lays it out
    const size_t rows = (size_t)L.bw + 1 + (L.n - L.nb);
    const size_t cols = (size_t)((2 * L.bw + 1 + (L.n - L.nb) + 1 + L.nbox) | 1);
    return rows * cols * sizeof(double) + (((size_t)L.bw + 1) * sizeof(int) + 15) / 16 * 16;
}

// the elimination window fits the shared memory and the per-thread loops of the window path cover it
static bool fb_use_window(const FbLayout& L)
{
    const int nbord = L.n - L.nb;
    return fb_window_bytes(L) <= FB_WINDOW_LIMIT && 2 * L.bw + 1 + nbord + 1 + L.nbox <= FB_THREADS && L.bw + 1 <= FB_THREADS - nbord
           && 2 * L.bw + nbord <= 256 && L.bw + 1 + nbord <= 128;
}

// dynamic shared memory of the launch: the window, or the pivot-row cache / multipliers / row list of the global-memory path
static size_t fb_dyn_bytes(const FbLayout& L)
{
    if (fb_use_window(L))
        return fb_window_bytes(L);
    const size_t nbord = L.n - L.nb;
    return (2 * L.bw + 1 + nbord + 1 + L.nbox + 8 + 2 * (L.bw + nbord + 8)) * sizeof(double);
}

bool fallback_supported(const DeviceConfig& cfg)
{
    const FallbackPlan P = fb_plan(cfg);
    // the pivot-row cache and the per-thread active set bound the sizes (nv <= 160: up to 40 throttle blocks)
    // the active set has one boxed variable per thread; the elimination runs in a shared-memory window when it fits
    return P.L.nbox <= FB_MAXBOX && P.L.nbox <= FB_THREADS && fb_dyn_bytes(P.L) <= FB_WINDOW_LIMIT;
}

size_t fallback_slot_doubles(const DeviceConfig& cfg) { return fb_plan(cfg).slot_doubles; }
size_t fallback_pos_ints(const DeviceConfig& cfg) { return fb_plan(cfg).pos.size(); }
void fallback_positions(const DeviceConfig& cfg, int* out)
{
    const FallbackPlan P = fb_plan(cfg);
    for (size_t i = 0; i < P.pos.size(); ++i)
        out[i] = P.pos[i];
}

cudaError_t launch_qp_fallback(const DeviceConfig& h_cfg, int B, int n_slots, const double* qd, const int* fb_list,
                               const int* fb_count, const int* pos, double* scratch, double* z, double* st, double* out_rows,
                               int* status, int* n_factor, int* n_solve, int* n_pivot, int want_z, double* out2, int* status2,
                               cudaStream_t s)
{
    const FallbackPlan P = fb_plan(h_cfg);
    static bool attr_set[64] = {};
    cudaError_t e = ensure_dynamic_smem(qp_fallback_kernel, (int)FB_WINDOW_LIMIT, attr_set);
    if (e != cudaSuccess)
        return e;
    qp_fallback_kernel<<<n_slots, FB_THREADS, fb_dyn_bytes(P.L), s>>>(h_cfg, P.L, B, qd, fb_list, fb_count, pos, scratch, P.slot_doubles,
                                                                     z, st, out_rows, status, n_factor, n_solve, n_pivot, want_z,
                                                                     fb_use_window(P.L) ? 1 : 0, out2, status2);
    return cudaGetLastError();
}

} // namespace vsmpc
