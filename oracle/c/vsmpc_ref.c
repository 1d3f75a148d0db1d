/*
 * vsmpc_ref.c — CPU restatement (plain C, FP64) of the reference's MPC tick:
 *   IMPCProblem::update (dense assembly)  MPC/src/IMPCProblem/IMPCProblem.cpp:150-194
 *   IMPCProblem::solve  (dense -> sparse scan, OSQP data update, solve, polish)  :196-298
 *   VariableSamplingMPC::solveMPC (output extraction)  MPC/src/variableSamplingMPC/variableSamplingMPC.cpp:88-112
 *
 * TEST INFRASTRUCTURE / CPU BASELINE ONLY.  Only tests/, __graft_entry__.smoke() and the cpu_baseline /
 * `--impl reference` legs of bench.py may load this library; the product never does.
 *
 * PARITY UNPINNED: the reference has no tests or golden vectors for this path and its solver stack is
 * not installable here.  The QP solve lives in an un-vendored third-party dependency:
 *   OSQP 1.0.0 + QDLDL 0.1.8 through osqp-eigen 0.11.0 (pixi.lock:232,237,317), called at
 *   IMPCProblem.cpp:140-145 (settings: warm start, polish; everything else OSQP defaults),
 *   :225-255 (setup), :263-277 (update q/A/l/u), :279 (solve), :296 (solution).
 * What is restated here is OSQP's *published* algorithm (Stellato et al., "OSQP: an operator splitting
 * solver for quadratic programs", 2020) with its documented default settings — ADMM with per-constraint
 * rho, sigma regularisation, alpha relaxation, Ruiz equilibration, adaptive rho, residual-based
 * termination every 25 iterations, warm start, polishing with iterative refinement — on a sparse
 * quasi-definite KKT system factorised by an up-looking LDL^T (Davis, "Algorithm 849: a concise sparse
 * Cholesky factorization package", the algorithm QDLDL implements) under a minimum-degree ordering.
 * It is pinned against the exact solver of oracle/vsmpc_oracle.py in tests/test_oracle_c.py.
 *
 * The assembly half follows the reference's own sources line by line (dense blocks and all), see the
 * citations at each function; the rigid-body quantities enter as data (the pack of include/vsmpc.h).
 */
#include <malloc.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NX 26
#define NJ 8
#define NT 4
#define PACKN 359

/* pack offsets = include/vsmpc.h VSMPC_PK_* (kept in sync by tests/test_oracle_c.py) */
enum {
    PK_WRB = 0, PK_OMEGA = 9, PK_RPY = 12, PK_MASS = 15, PK_GRAV = 16, PK_MB = 19, PK_BASE = 55, PK_PCOM = 58,
    PK_MOM = 61, PK_AMOM = 67, PK_AXES = 91, PK_ARMS = 103, PK_JW = 115, PK_JL = 211, PK_JC = 307, PK_T = 331,
    PK_TDE = 335, PK_TDES = 339, PK_TDDES = 343, PK_UPREV = 347, PK_QCMD = 351
};

typedef struct
{
    int n_iter, n_small, n_ctrl;
    double period_mpc, period_large, period_small;
    int use_jet_dynamic, use_estimated_thrust;
    double w_com[3], w_com_err[3], w_lin[3], w_rpy[3], w_rpy_err[3], w_ang[3];
    double w_dq[NJ], w_throttle, w_init_throttle, w_reg_q;
    double throttle_min, throttle_max;
    double jc[13], jn[4];
    const double* alpha; int alpha_len;              /* already resampled to 1/periodMPC           */
    const double *tpos, *tvel, *trpy, *trpyd; int traj_len; /* sample-major [3*s+a], at 1/periodLarge */
    /* OSQP settings (defaults of OSQP 1.0.0 unless the reference sets them) */
    double rho, sigma, alpha_relax, eps_abs, eps_rel, delta;
    int max_iter, check_termination, scaling_iters, adaptive_rho, adaptive_rho_interval, polish, polish_refine_iter;
    double adaptive_rho_tolerance;
} ref_config;

/* ---------------------------------------------------------------------------------------------------- */
/* sparse helpers (CSC)                                                                                   */
typedef struct { int m, n, nnz; int* p; int* i; double* x; } csc;

static csc csc_alloc(int m, int n, int nnz)
{
    csc A; A.m = m; A.n = n; A.nnz = nnz;
    A.p = (int*)calloc(n + 1, sizeof(int)); A.i = (int*)malloc(sizeof(int) * (nnz > 0 ? nnz : 1));
    A.x = (double*)malloc(sizeof(double) * (nnz > 0 ? nnz : 1));
    return A;
}
static void csc_free(csc* A) { free(A->p); free(A->i); free(A->x); A->p = 0; A->i = 0; A->x = 0; }

/* dense (row-major m x n) -> CSC dropping exact zeros: Eigen's sparseView() (IMPCProblem.cpp:211) */
static csc dense_to_csc(const double* M, int m, int n)
{
    int nnz = 0;
    for (int k = 0; k < m * n; ++k) nnz += (M[k] != 0.0);
    csc A = csc_alloc(m, n, nnz);
    int c = 0;
    for (int j = 0; j < n; ++j)
    {
        A.p[j] = c;
        for (int i = 0; i < m; ++i)
            if (M[(size_t)i * n + j] != 0.0) { A.i[c] = i; A.x[c] = M[(size_t)i * n + j]; ++c; }
    }
    A.p[n] = c;
    return A;
}
static int csc_same_pattern(const csc* A, const csc* B)
{
    if (A->m != B->m || A->n != B->n || A->nnz != B->nnz) return 0;
    return memcmp(A->p, B->p, sizeof(int) * (A->n + 1)) == 0 && memcmp(A->i, B->i, sizeof(int) * A->nnz) == 0;
}
static void csc_mv(const csc* A, const double* x, double* y) /* y = A x */
{
    memset(y, 0, sizeof(double) * A->m);
    for (int j = 0; j < A->n; ++j) { const double xj = x[j]; for (int p = A->p[j]; p < A->p[j + 1]; ++p) y[A->i[p]] += A->x[p] * xj; }
}
static void csc_mtv(const csc* A, const double* x, double* y) /* y = A^T x */
{
    for (int j = 0; j < A->n; ++j) { double s = 0; for (int p = A->p[j]; p < A->p[j + 1]; ++p) s += A->x[p] * x[A->i[p]]; y[j] = s; }
}
static double vnorm_inf(const double* v, int n) { double m = 0; for (int k = 0; k < n; ++k) { double a = fabs(v[k]); if (a > m) m = a; } return m; }

/* ---------------------------------------------------------------------------------------------------- */
/* LDL^T of a symmetric quasi-definite matrix given by its upper triangle in CSC (Davis' up-looking LDL) */
typedef struct
{
    int n; int* Lp; int* Li; double* Lx; double* D; double* Dinv; int* Parent; int* Lnz; int* Flag; int* Pattern; double* Y;
    int* perm; int* iperm; /* fill-reducing ordering */
    /* permuted upper-triangular K */
    int* Kp; int* Ki; double* Kx; int* Kmap; /* Kmap: entry of the un-permuted KKT -> position in Kx */
    int knnz; double* work;
} ldl;

static void ldl_symbolic(ldl* F)
{
    const int n = F->n;
    for (int k = 0; k < n; ++k)
    {
        F->Parent[k] = -1; F->Flag[k] = k; F->Lnz[k] = 0;
        for (int p = F->Kp[k]; p < F->Kp[k + 1]; ++p)
        {
            int i = F->Ki[p];
            if (i < k)
                for (; F->Flag[i] != k; i = F->Parent[i])
                {
                    if (F->Parent[i] == -1) F->Parent[i] = k;
                    F->Lnz[i]++; F->Flag[i] = k;
                }
        }
    }
    F->Lp[0] = 0;
    for (int k = 0; k < n; ++k) F->Lp[k + 1] = F->Lp[k] + F->Lnz[k];
}
static int ldl_numeric(ldl* F)
{
    const int n = F->n;
    for (int k = 0; k < n; ++k)
    {
        F->Y[k] = 0.0; int top = n; F->Flag[k] = k; F->Lnz[k] = 0;
        for (int p = F->Kp[k]; p < F->Kp[k + 1]; ++p)
        {
            int i = F->Ki[p];
            if (i <= k)
            {
                F->Y[i] += F->Kx[p];
                int len = 0;
                for (; F->Flag[i] != k; i = F->Parent[i]) { F->Pattern[len++] = i; F->Flag[i] = k; }
                while (len > 0) F->Pattern[--top] = F->Pattern[--len];
            }
        }
        F->D[k] = F->Y[k]; F->Y[k] = 0.0;
        for (; top < n; ++top)
        {
            const int i = F->Pattern[top];
            const double yi = F->Y[i]; F->Y[i] = 0.0;
            const int p2 = F->Lp[i] + F->Lnz[i];
            for (int p = F->Lp[i]; p < p2; ++p) F->Y[F->Li[p]] -= F->Lx[p] * yi;
            const double lki = yi / F->D[i];
            F->D[k] -= lki * yi;
            F->Li[p2] = k; F->Lx[p2] = lki; F->Lnz[i]++;
        }
        if (F->D[k] == 0.0) return k;
        F->Dinv[k] = 1.0 / F->D[k];
    }
    return n;
}
static void ldl_solve(const ldl* F, double* b) /* in place, b in original ordering */
{
    const int n = F->n; double* x = F->work;
    for (int k = 0; k < n; ++k) x[k] = b[F->perm[k]];
    for (int j = 0; j < n; ++j) { const double xj = x[j]; for (int p = F->Lp[j]; p < F->Lp[j + 1]; ++p) x[F->Li[p]] -= F->Lx[p] * xj; }
    for (int j = 0; j < n; ++j) x[j] *= F->Dinv[j];
    for (int j = n - 1; j >= 0; --j) { double s = x[j]; for (int p = F->Lp[j]; p < F->Lp[j + 1]; ++p) s -= F->Lx[p] * x[F->Li[p]]; x[j] = s; }
    for (int k = 0; k < n; ++k) b[F->perm[k]] = x[k];
}
static void ldl_free(ldl* F)
{
    free(F->Lp); free(F->Li); free(F->Lx); free(F->D); free(F->Dinv); free(F->Parent); free(F->Lnz); free(F->Flag);
    free(F->Pattern); free(F->Y); free(F->perm); free(F->iperm); free(F->Kp); free(F->Ki); free(F->Kx); free(F->Kmap); free(F->work);
    memset(F, 0, sizeof(*F));
}

/* greedy minimum-degree ordering on the pattern of a symmetric matrix given as upper CSC (bitset graph) */
static void min_degree(int n, const int* Up, const int* Ui, int* perm)
{
    const int W = (n + 63) / 64;
    uint64_t* adj = (uint64_t*)calloc((size_t)n * W, sizeof(uint64_t));
    char* done = (char*)calloc(n, 1);
    int* deg = (int*)calloc(n, sizeof(int));
    for (int j = 0; j < n; ++j)
        for (int p = Up[j]; p < Up[j + 1]; ++p)
        {
            const int i = Ui[p];
            if (i != j) { adj[(size_t)i * W + j / 64] |= 1ull << (j % 64); adj[(size_t)j * W + i / 64] |= 1ull << (i % 64); }
        }
    for (int v = 0; v < n; ++v) { int d = 0; for (int w = 0; w < W; ++w) d += __builtin_popcountll(adj[(size_t)v * W + w]); deg[v] = d; }
    for (int k = 0; k < n; ++k)
    {
        int best = -1;
        for (int v = 0; v < n; ++v) if (!done[v] && (best < 0 || deg[v] < deg[best])) best = v;
        perm[k] = best; done[best] = 1;
        uint64_t* av = adj + (size_t)best * W;
        for (int w = 0; w < W; ++w)
        {
            uint64_t bits = av[w];
            while (bits)
            {
                const int u = w * 64 + __builtin_ctzll(bits); bits &= bits - 1;
                uint64_t* au = adj + (size_t)u * W;
                for (int t = 0; t < W; ++t) au[t] |= av[t];
                au[best / 64] &= ~(1ull << (best % 64));
                au[u / 64] &= ~(1ull << (u % 64));
                int d = 0; for (int t = 0; t < W; ++t) d += __builtin_popcountll(au[t]); deg[u] = d;
            }
        }
        for (int v = 0; v < n; ++v) if (!done[v]) adj[(size_t)v * W + best / 64] &= ~(1ull << (best % 64));
    }
    free(adj); free(done); free(deg);
}

/* ---------------------------------------------------------------------------------------------------- */
typedef struct
{
    ref_config cfg;
    int N, Ns, Nc, NC, nblk, nvar, ncon, ratio;
    double dt[256];
    double vmin, vmax;
    /* persistent state of the costs / constraints */
    int ref_counter, thr_counter, alpha_idx, ref_idx;
    double p_init[3], rpy_init[3], rpy_old[3], nturns[3], qref0[NJ], qacc[NJ];
    double p_ref[3], rpy_ref[3];
    double* win; /* 12 x NC reference windows */
    /* dense problem data, as the reference holds it */
    double *P, *q, *A, *l, *u;        /* P nvar x nvar, A ncon x nvar (row-major) */
    double Ad[NX * NX], BJd[NX * NJ], BTd[NX * NT], cd[NX];
    int first_update;
    /* OSQP-like workspace */
    int initialised;
    csc Ps, As;                        /* scaled problem data (P upper triangle) */
    double *D, *E, *Dinv, *Einv, cscale;
    double *qs, *ls, *us;              /* scaled vectors */
    double *rho_vec; int* ctype; double rho;
    ldl F;
    double *x, *z, *y, *xt, *zt, *xprev, *zprev, *rhs, *tmpn, *tmpm, *tmpm2;
    double* sol; double* ysol;
    int status, iters, polished, n_refactor;
    /* outputs */
    double out[54];
} ref_inst;

/* ---- jet model, UT/src/JetModel.cpp:29-109 ----------------------------------------------------------- */
static double jf(const double* c, double T, double Td) { return c[0] + c[1] * T + c[2] * Td + c[3] * T * Td + c[4] * pow(T, 2.0) + c[5] * pow(Td, 2.0); }
static double jg(const double* c, double T, double Td) { return c[6] + c[7] * T + c[8] * Td + c[9] * T * Td + c[10] * pow(T, 2.0) + c[11] * pow(Td, 2.0); }
static double jv(const double* c, double u) { return u + c[12] * pow(u, 2.0); }
static double destd_u(const double* c, const double* n, double v)
{
    double u = (-1 + sqrt(1 + 4 * c[12] * v)) / (2 * c[12]);
    u = u * n[3] + n[2];
    if (u < 0) u = 0; else if (u > 100) u = 100;
    return u;
}

static void skew(const double* v, double* S) { S[0] = 0; S[1] = -v[2]; S[2] = v[1]; S[3] = v[2]; S[4] = 0; S[5] = -v[0]; S[6] = -v[1]; S[7] = v[0]; S[8] = 0; }
static void m3mul(const double* A, const double* B, double* C) { for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j]; }
static void m3Tv(const double* R, const double* v, double* o) { for (int i = 0; i < 3; ++i) o[i] = R[i] * v[0] + R[3 + i] * v[1] + R[6 + i] * v[2]; }
static void m3inv(const double* A, double* I)
{
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
    I[0] = c00 * id; I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[2] = (A[1] * A[5] - A[2] * A[4]) * id;
    I[3] = c01 * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id;
    I[6] = c02 * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}
/* (X^T M_b X).block(3,3,3,3), X = Ad(G_H_B)  (systemDynamicsVSMPC.cpp:110-130, costsVSMPC.cpp:268-285) */
static void locked_inertia(const double* pk, double* I3)
{
    const double* R = pk + PK_WRB; double r[3], Sr[9], SR[9], X[36], MX[36];
    for (int a = 0; a < 3; ++a) r[a] = pk[PK_PCOM + a] - pk[PK_BASE + a];
    skew(r, Sr); m3mul(Sr, R, SR);
    memset(X, 0, sizeof(X));
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { X[a * 6 + b] = R[a * 3 + b]; X[a * 6 + 3 + b] = SR[a * 3 + b]; X[(3 + a) * 6 + 3 + b] = R[a * 3 + b]; }
    for (int a = 0; a < 6; ++a) for (int b = 0; b < 6; ++b) { double s = 0; for (int k = 0; k < 6; ++k) s += pk[PK_MB + a * 6 + k] * X[k * 6 + b]; MX[a * 6 + b] = s; }
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) { double s = 0; for (int k = 0; k < 6; ++k) s += X[k * 6 + 3 + a] * MX[k * 6 + 3 + b]; I3[a * 3 + b] = s; }
}
static void W_of_rpy(const double* rpy, double* W)
{
    memset(W, 0, 9 * sizeof(double));
    W[0] = 1.0; W[4] = cos(rpy[0]); W[7] = -sin(rpy[0]); W[2] = -sin(rpy[1]); W[5] = cos(rpy[1]) * sin(rpy[0]); W[8] = cos(rpy[0]) * cos(rpy[1]);
}
static void ref_column(const ref_inst* I, const double* pk, int idx, double* col)
{ /* costsVSMPC.cpp:105-112,132-149 */
    const ref_config* c = &I->cfg; const double* R = pk + PK_WRB; const double mass = pk[PK_MASS];
    double I3[9], W[9], mv[3], Wr[3];
    locked_inertia(pk, I3); W_of_rpy(pk + PK_RPY, W);
    for (int a = 0; a < 3; ++a) { col[a] = I->p_init[a] + c->tpos[3 * idx + a]; col[6 + a] = I->rpy_init[a] + c->trpy[3 * idx + a]; mv[a] = mass * c->tvel[3 * idx + a]; }
    m3Tv(R, mv, col + 3);
    for (int a = 0; a < 3; ++a) Wr[a] = W[a * 3] * c->trpyd[3 * idx] + W[a * 3 + 1] * c->trpyd[3 * idx + 1] + W[a * 3 + 2] * c->trpyd[3 * idx + 2];
    for (int a = 0; a < 3; ++a) col[9 + a] = I3[a * 3] * Wr[0] + I3[a * 3 + 1] * Wr[1] + I3[a * 3 + 2] * Wr[2];
}
static int ref_colidx(int k, int Ns) { return k < Ns ? 0 : k - Ns; } /* costsVSMPC.cpp:191-200 */

/* one evaluation of all costs and constraints = IMPCProblem::update (IMPCProblem.cpp:150-194) */
static void assemble(ref_inst* I, const double* pk)
{
    const ref_config* c = &I->cfg;
    const int N = I->N, Ns = I->Ns, Nc = I->Nc, NC = I->NC, nvar = I->nvar, ncon = I->ncon;
    const int base = NX * (N + 1), tbase = base + Nc * NJ;
    const double* R = pk + PK_WRB;
    /* ---- costs ---- */
    memset(I->q, 0, sizeof(double) * nvar);                      /* m_gradient.setZero(), :156 */
    if (I->ref_counter == I->ratio - 1)
    { /* ReferenceTrackingCost, costsVSMPC.cpp:124-161 */
        if (I->ref_idx < c->traj_len - 1) I->ref_idx++;
        double col[12]; ref_column(I, pk, I->ref_idx, col);
        for (int r = 0; r < 12; ++r) { for (int k = 0; k + 1 < NC; ++k) I->win[r * NC + k] = I->win[r * NC + k + 1]; I->win[r * NC + NC - 1] = col[r]; }
        for (int a = 0; a < 3; ++a) { I->p_ref[a] = I->win[a * NC]; I->rpy_ref[a] = I->win[(6 + a) * NC]; }
        I->ref_counter = 0;
    }
    else I->ref_counter++;
    double Qd[NX]; memset(Qd, 0, sizeof(Qd));
    for (int a = 0; a < 3; ++a) { Qd[a] = c->w_com[a]; Qd[3 + a] = c->w_lin[a]; Qd[6 + a] = c->w_rpy[a]; Qd[9 + a] = c->w_ang[a]; Qd[20 + a] = c->w_com_err[a]; Qd[23 + a] = c->w_rpy_err[a]; }
    if (I->first_update)
    { /* Hessian: summed once, then frozen (IMPCProblem.cpp:152-175; costsVSMPC.cpp:166-174,371-411,470-478,560-573) */
        memset(I->P, 0, sizeof(double) * (size_t)nvar * nvar);
        for (int k = 1; k <= N; ++k) for (int r = 0; r < NX; ++r) I->P[(size_t)(k * NX + r) * nvar + k * NX + r] += Qd[r];
        for (int j = 0; j < Nc; ++j) for (int a = 0; a < NJ; ++a) I->P[(size_t)(base + j * NJ + a) * nvar + base + j * NJ + a] += c->w_dq[a] + c->w_reg_q;
        for (int b = 0; b + 1 < I->nblk; ++b)
            for (int a = 0; a < NT; ++a)
            {
                const int i0 = tbase + b * NT + a, i1 = tbase + (b + 1) * NT + a;
                I->P[(size_t)i0 * nvar + i0] += c->w_throttle; I->P[(size_t)i1 * nvar + i1] += c->w_throttle;
                I->P[(size_t)i0 * nvar + i1] -= c->w_throttle; I->P[(size_t)i1 * nvar + i0] -= c->w_throttle;
            }
        for (int a = 0; a < NT; ++a) I->P[(size_t)(tbase + a) * nvar + tbase + a] += c->w_init_throttle;
        I->first_update = 0;
    }
    for (int k = 1; k <= N; ++k) /* q[x_k] = -Q xref_{k-1}, costsVSMPC.cpp:175-178 */
        for (int r = 0; r < 12; ++r) I->q[k * NX + r] = -Qd[r] * I->win[r * NC + ref_colidx(k - 1, Ns)];
    double vbar[NT];
    for (int j = 0; j < NT; ++j)
    { /* ThrottleInitialValueCost, costsVSMPC.cpp:479-485 */
        vbar[j] = jv(c->jc, (pk[PK_UPREV + j] - c->jn[2]) / c->jn[3]);
        I->q[tbase + j] += -c->w_init_throttle * vbar[j];
    }
    for (int j = 0; j < Nc; ++j) /* JointPositionRegularizationCost, costsVSMPC.cpp:574-590 */
        for (int a = 0; a < NJ; ++a) I->q[base + j * NJ + a] += c->w_reg_q * (pk[PK_QCMD + a] - I->qref0[a]);
    /* ---- dynamics: Angular, Linear, Jet (systemDynamicsVSMPC.cpp:72-461), summed (:509-585) ---- */
    double* A = I->Ad; double* BJ = I->BJd; double* BT = I->BTd; double* cv = I->cd;
    memset(A, 0, sizeof(I->Ad)); memset(BJ, 0, sizeof(I->BJd)); memset(BT, 0, sizeof(I->BTd)); memset(cv, 0, sizeof(I->cd));
    const double mass = pk[PK_MASS];
    double omB[3]; m3Tv(R, pk + PK_OMEGA, omB);
    double So[9]; skew(omB, So);
    {
        double I3[9], Ii[9], WI[9]; locked_inertia(pk, I3); m3inv(I3, Ii);
        const double* rpy = pk + PK_RPY;
        const double s0 = sin(rpy[0]), c0 = cos(rpy[0]), t1 = tan(rpy[1]), c1 = cos(rpy[1]);
        const double Wi[9] = {1.0, s0 * t1, c0 * t1, 0.0, c0, -s0, 0.0, s0 / c1, c0 / c1};
        m3mul(Wi, Ii, WI);
        for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b)
        {
            A[(6 + a) * NX + 9 + b] = WI[a * 3 + b];
            A[(9 + a) * NX + 9 + b] -= So[a * 3 + b];
            A[(3 + a) * NX + 3 + b] -= So[a * 3 + b];
            A[a * NX + 3 + b] = (1 / mass) * R[a * 3 + b];
        }
    }
    for (int a = 0; a < 3; ++a)
    {
        for (int j = 0; j < NT; ++j) { A[(3 + a) * NX + 12 + j] = pk[PK_AMOM + a * NT + j]; A[(9 + a) * NX + 12 + j] = pk[PK_AMOM + (3 + a) * NT + j]; }
        A[(20 + a) * NX + a] = 1.0; A[(23 + a) * NX + 6 + a] = 1.0;
        cv[20 + a] = -I->p_ref[a]; cv[23 + a] = -I->rpy_init[a];
    }
    for (int j = 0; j < NT; ++j)
    { /* Lambda matrices, "unfiltered" (:166-186,338-346) */
        double ab[3], rb[3], Sa[9], Sr[9], SrSa[9];
        m3Tv(R, pk + PK_AXES + 3 * j, ab); m3Tv(R, pk + PK_ARMS + 3 * j, rb);
        skew(ab, Sa); skew(rb, Sr); m3mul(Sr, Sa, SrSa);
        const double T = pk[PK_T + j];
        for (int b = 0; b < NJ; ++b)
        {
            double Jw[3], dl[3], Jc[3];
            for (int a = 0; a < 3; ++a) { Jw[a] = pk[PK_JW + (j * 3 + a) * NJ + b]; dl[a] = pk[PK_JL + (j * 3 + a) * NJ + b] - pk[PK_JC + a * NJ + b]; }
            m3Tv(R, dl, Jc);
            for (int a = 0; a < 3; ++a)
            {
                BJ[(9 + a) * NJ + b] -= T * (Sa[a * 3] * Jc[0] + Sa[a * 3 + 1] * Jc[1] + Sa[a * 3 + 2] * Jc[2]);
                BJ[(9 + a) * NJ + b] -= T * (SrSa[a * 3] * Jw[0] + SrSa[a * 3 + 1] * Jw[1] + SrSa[a * 3 + 2] * Jw[2]);
                BJ[(3 + a) * NJ + b] -= T * (Sa[a * 3] * Jw[0] + Sa[a * 3 + 1] * Jw[1] + Sa[a * 3 + 2] * Jw[2]);
            }
        }
    }
    {
        const double alpha = c->alpha[I->alpha_idx], s = alpha * mass;
        for (int a = 0; a < 3; ++a) cv[3 + a] = (s * R[a]) * pk[PK_GRAV] + (s * R[3 + a]) * pk[PK_GRAV + 1] + (s * R[6 + a]) * pk[PK_GRAV + 2];
        if (I->alpha_idx < c->alpha_len - 1) I->alpha_idx++;
    }
    if (c->use_jet_dynamic)
        for (int j = 0; j < NT; ++j)
        {
            const double T = c->use_estimated_thrust ? pk[PK_T + j] : pk[PK_TDES + j];
            const double Td = c->use_estimated_thrust ? pk[PK_TDE + j] : pk[PK_TDDES + j];
            const double Tb = (T - c->jn[0]) / c->jn[1], Tdb = Td / c->jn[1];
            const double vu = jv(c->jc, (pk[PK_UPREV + j] - c->jn[2]) / c->jn[3]);
            const double dT = (c->jc[1] + c->jc[3] * Tdb + 2 * c->jc[4] * Tb) + (c->jc[7] + c->jc[9] * Tdb + 2 * c->jc[10] * Tb) * vu;
            const double dTd = (c->jc[2] + c->jc[3] * Tb + 2 * c->jc[5] * Tdb) + (c->jc[8] + c->jc[9] * Tb + 2 * c->jc[11] * Tdb) * vu;
            A[(12 + j) * NX + 16 + j] = 1.0;
            A[(16 + j) * NX + 12 + j] = dT; A[(16 + j) * NX + 16 + j] += dTd;
            BT[(16 + j) * NT + j] = jg(c->jc, (pk[PK_TDES + j] - c->jn[0]) / c->jn[1], pk[PK_TDDES + j] / c->jn[1]) * c->jn[1];
            cv[16 + j] = jf(c->jc, Tb, Tdb) * c->jn[1] - dT * T - dTd * Td;
        }
    else
        for (int j = 0; j < NT; ++j) BT[(12 + j) * NT + j] = 1.0;
    /* ---- dense constraint matrix, as the reference fills and copies it (constraintsVSMPC.cpp:73-131;
     *      IMPCProblem.cpp:177-192).  m_linearMatrix.setZero() + block writes every tick. ---- */
    memset(I->A, 0, sizeof(double) * (size_t)ncon * nvar);
    memset(I->l, 0, sizeof(double) * ncon); memset(I->u, 0, sizeof(double) * ncon);
    for (int k = 0; k < N; ++k)
    {
        const double dT = I->dt[k];
        const int jb = k < Nc ? k : Nc - 1;
        const int tb = k < Ns ? 0 : (k < Nc ? k - (Ns - 1) : Nc - Ns);
        for (int r = 0; r < NX; ++r)
        {
            double* row = I->A + (size_t)(k * NX + r) * nvar;
            for (int j = 0; j < NX; ++j) row[k * NX + j] = (r == j ? 1.0 : 0.0) + dT * A[r * NX + j];
            row[(k + 1) * NX + r] = -1.0;
            for (int j = 0; j < NJ; ++j) row[base + jb * NJ + j] = dT * BJ[r * NJ + j];
            for (int j = 0; j < NT; ++j) row[tbase + tb * NT + j] = dT * BT[r * NT + j];
            I->l[k * NX + r] = I->u[k * NX + r] = -dT * cv[r];
        }
    }
    { /* ConstraintInitialState (constraintsVSMPC.cpp:206-247; IQPUtilsMPC.cpp:71-92) */
        const double PI = 3.14159265358979323846; double x0[NX];
        for (int a = 0; a < 3; ++a)
        {
            const double cur = pk[PK_RPY + a];
            if (cur - I->rpy_old[a] > PI) I->nturns[a] -= 1; else if (cur - I->rpy_old[a] < -PI) I->nturns[a] += 1;
            I->rpy_old[a] = cur;
            const double unw = cur + 2 * PI * I->nturns[a];
            x0[a] = pk[PK_PCOM + a]; x0[3 + a] = pk[PK_MOM + a]; x0[6 + a] = unw; x0[9 + a] = pk[PK_MOM + 3 + a];
            x0[20 + a] = pk[PK_PCOM + a] - I->p_ref[a]; x0[23 + a] = unw - I->rpy_ref[a];
        }
        for (int j = 0; j < NT; ++j)
        {
            x0[12 + j] = c->use_estimated_thrust ? pk[PK_T + j] : pk[PK_TDES + j];
            x0[16 + j] = c->use_estimated_thrust ? pk[PK_TDE + j] : pk[PK_TDDES + j];
        }
        for (int r = 0; r < NX; ++r) { I->A[(size_t)(N * NX + r) * nvar + r] = 1.0; I->l[N * NX + r] = I->u[N * NX + r] = x0[r]; }
    }
    { /* ThrottleConstraint (constraintsVSMPC.cpp:338-374) */
        const int r0 = N * NX + NX;
        for (int b = 0; b < I->nblk; ++b)
            for (int a = 0; a < NT; ++a)
            {
                I->A[(size_t)(r0 + b * NT + a) * nvar + tbase + b * NT + a] = 1.0;
                if (I->thr_counter != I->ratio - 1 && b == 0) I->l[r0 + a] = I->u[r0 + a] = vbar[a];
                else { I->l[r0 + b * NT + a] = I->vmin; I->u[r0 + b * NT + a] = I->vmax; }
            }
        I->thr_counter = (I->thr_counter == I->ratio - 1) ? 0 : I->thr_counter + 1;
    }
}

/* ---------------------------------------------------------------------------------------------------- */
/* OSQP-like solver                                                                                      */
static void osqp_free_ws(ref_inst* I)
{
    if (!I->initialised) return;
    csc_free(&I->Ps); csc_free(&I->As); ldl_free(&I->F);
    free(I->D); free(I->E); free(I->Dinv); free(I->Einv); free(I->qs); free(I->ls); free(I->us); free(I->rho_vec); free(I->ctype);
    free(I->xt); free(I->zt); free(I->xprev); free(I->zprev); free(I->rhs); free(I->tmpn); free(I->tmpm); free(I->tmpm2);
    I->initialised = 0;
}

/* KKT = [[P + sigma I, A^T],[A, -diag(1/rho)]], upper triangle, built once (pattern) then refreshed */
static void kkt_build(ref_inst* I)
{
    const int n = I->nvar, m = I->ncon, nk = n + m;
    ldl* F = &I->F;
    /* count: P upper incl. full diagonal, A^T block (columns n..n+m-1 hold column i = row i of A), -1/rho diagonal */
    int* cnt = (int*)calloc(nk + 1, sizeof(int));
    /* P (upper) per column, force diagonal */
    for (int j = 0; j < n; ++j) { int hasdiag = 0; for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p) { cnt[j + 1]++; if (I->Ps.i[p] == j) hasdiag = 1; } if (!hasdiag) cnt[j + 1]++; }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p) cnt[n + I->As.i[p] + 1]++;
    for (int i = 0; i < m; ++i) cnt[n + i + 1]++;
    int* Up = (int*)calloc(nk + 1, sizeof(int));
    for (int j = 0; j < nk; ++j) Up[j + 1] = Up[j] + cnt[j + 1];
    const int nnz = Up[nk];
    int* Ui = (int*)malloc(sizeof(int) * nnz); double* Ux = (double*)calloc(nnz, sizeof(double));
    int* fill = (int*)calloc(nk, sizeof(int));
    for (int j = 0; j < n; ++j)
    {
        int hasdiag = 0;
        for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p) { Ui[Up[j] + fill[j]] = I->Ps.i[p]; fill[j]++; if (I->Ps.i[p] == j) hasdiag = 1; }
        if (!hasdiag) { Ui[Up[j] + fill[j]] = j; fill[j]++; }
    }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p) { const int col = n + I->As.i[p]; Ui[Up[col] + fill[col]] = j; fill[col]++; }
    for (int i = 0; i < m; ++i) { Ui[Up[n + i] + fill[n + i]] = n + i; fill[n + i]++; }
    /* ordering + permuted upper triangle */
    F->n = nk;
    F->perm = (int*)malloc(sizeof(int) * nk); F->iperm = (int*)malloc(sizeof(int) * nk);
    min_degree(nk, Up, Ui, F->perm);
    for (int k = 0; k < nk; ++k) F->iperm[F->perm[k]] = k;
    F->Kp = (int*)calloc(nk + 1, sizeof(int)); F->Ki = (int*)malloc(sizeof(int) * nnz); F->Kx = (double*)calloc(nnz, sizeof(double));
    F->Kmap = (int*)malloc(sizeof(int) * nnz); F->knnz = nnz;
    int* c2 = (int*)calloc(nk + 1, sizeof(int));
    for (int j = 0; j < nk; ++j) for (int p = Up[j]; p < Up[j + 1]; ++p) { const int a = F->iperm[Ui[p]], b = F->iperm[j]; c2[(a > b ? a : b) + 1]++; }
    for (int j = 0; j < nk; ++j) F->Kp[j + 1] = F->Kp[j] + c2[j + 1];
    memset(fill, 0, sizeof(int) * nk);
    for (int j = 0; j < nk; ++j) for (int p = Up[j]; p < Up[j + 1]; ++p)
    {
        const int a = F->iperm[Ui[p]], b = F->iperm[j]; const int col = a > b ? a : b, row = a > b ? b : a;
        const int pos = F->Kp[col] + fill[col]++; F->Ki[pos] = row; F->Kmap[p] = pos;
    }
    F->Lp = (int*)calloc(nk + 1, sizeof(int)); F->Parent = (int*)malloc(sizeof(int) * nk); F->Lnz = (int*)malloc(sizeof(int) * nk);
    F->Flag = (int*)malloc(sizeof(int) * nk); F->Pattern = (int*)malloc(sizeof(int) * nk); F->Y = (double*)calloc(nk, sizeof(double));
    F->D = (double*)malloc(sizeof(double) * nk); F->Dinv = (double*)malloc(sizeof(double) * nk); F->work = (double*)malloc(sizeof(double) * nk);
    ldl_symbolic(F);
    F->Li = (int*)malloc(sizeof(int) * (F->Lp[nk] > 0 ? F->Lp[nk] : 1)); F->Lx = (double*)malloc(sizeof(double) * (F->Lp[nk] > 0 ? F->Lp[nk] : 1));
    free(cnt); free(c2); free(fill); free(Ux); free(Up); free(Ui);
}

/* refresh KKT values following exactly the traversal order of kkt_build */
static void kkt_refresh(ref_inst* I)
{
    const int n = I->nvar, m = I->ncon, nk = n + m; ldl* F = &I->F;
    memset(F->Kx, 0, sizeof(double) * F->knnz);
    /* rebuild the column pointers of the un-permuted upper KKT (cheap) */
    int* Up = (int*)calloc(nk + 1, sizeof(int));
    for (int j = 0; j < n; ++j) { int hasdiag = 0; for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p) { Up[j + 1]++; if (I->Ps.i[p] == j) hasdiag = 1; } if (!hasdiag) Up[j + 1]++; }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p) Up[n + I->As.i[p] + 1]++;
    for (int i = 0; i < m; ++i) Up[n + i + 1]++;
    for (int j = 0; j < nk; ++j) Up[j + 1] += Up[j];
    int* fill = (int*)calloc(nk, sizeof(int));
    for (int j = 0; j < n; ++j)
    {
        int hasdiag = 0;
        for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p)
        {
            double v = I->Ps.x[p]; if (I->Ps.i[p] == j) { v += I->cfg.sigma; hasdiag = 1; }
            F->Kx[F->Kmap[Up[j] + fill[j]]] = v; fill[j]++;
        }
        if (!hasdiag) { F->Kx[F->Kmap[Up[j] + fill[j]]] = I->cfg.sigma; fill[j]++; }
    }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p) { const int col = n + I->As.i[p]; F->Kx[F->Kmap[Up[col] + fill[col]]] = I->As.x[p]; fill[col]++; }
    for (int i = 0; i < m; ++i) { F->Kx[F->Kmap[Up[n + i] + fill[n + i]]] = -1.0 / I->rho_vec[i]; fill[n + i]++; }
    free(Up); free(fill);
    ldl_numeric(F);
    I->n_refactor++;
}

static void set_rho_vec(ref_inst* I)
{
    for (int i = 0; i < I->ncon; ++i)
    {
        const double lo = I->ls[i], up = I->us[i];
        if (lo < -1e19 && up > 1e19) { I->ctype[i] = -1; I->rho_vec[i] = 1e-6; }
        else if (up - lo < 1e-4) { I->ctype[i] = 1; I->rho_vec[i] = 1e3 * I->rho; }
        else { I->ctype[i] = 0; I->rho_vec[i] = I->rho; }
    }
}

/* Ruiz equilibration of (P, q, A) as in OSQP's scale_data (10 iterations), cost scaling included */
static void scale_problem(ref_inst* I, csc* P, csc* A, double* q)
{
    const int n = I->nvar, m = I->ncon;
    for (int j = 0; j < n; ++j) I->D[j] = 1.0;
    for (int i = 0; i < m; ++i) I->E[i] = 1.0;
    I->cscale = 1.0;
    double* dn = (double*)malloc(sizeof(double) * n); double* en = (double*)malloc(sizeof(double) * m);
    for (int it = 0; it < I->cfg.scaling_iters; ++it)
    {
        for (int j = 0; j < n; ++j) dn[j] = 0; for (int i = 0; i < m; ++i) en[i] = 0;
        /* column inf-norms of [P A^T; A 0] (P stored upper: symmetric contributions) */
        for (int j = 0; j < n; ++j) for (int p = P->p[j]; p < P->p[j + 1]; ++p) { const double a = fabs(P->x[p]); const int i = P->i[p]; if (a > dn[j]) dn[j] = a; if (a > dn[i]) dn[i] = a; }
        for (int j = 0; j < n; ++j) for (int p = A->p[j]; p < A->p[j + 1]; ++p) { const double a = fabs(A->x[p]); if (a > dn[j]) dn[j] = a; if (a > en[A->i[p]]) en[A->i[p]] = a; }
        for (int j = 0; j < n; ++j) { double v = dn[j]; if (v < 1e-4) v = 1.0; else if (v > 1e4) v = 1e4; dn[j] = 1.0 / sqrt(v); }
        for (int i = 0; i < m; ++i) { double v = en[i]; if (v < 1e-4) v = 1.0; else if (v > 1e4) v = 1e4; en[i] = 1.0 / sqrt(v); }
        for (int j = 0; j < n; ++j) for (int p = P->p[j]; p < P->p[j + 1]; ++p) P->x[p] *= dn[j] * dn[P->i[p]];
        for (int j = 0; j < n; ++j) for (int p = A->p[j]; p < A->p[j + 1]; ++p) A->x[p] *= dn[j] * en[A->i[p]];
        for (int j = 0; j < n; ++j) { q[j] *= dn[j]; I->D[j] *= dn[j]; }
        for (int i = 0; i < m; ++i) I->E[i] *= en[i];
        /* cost scaling */
        double colmean = 0.0;
        for (int j = 0; j < n; ++j) dn[j] = 0;
        for (int j = 0; j < n; ++j) for (int p = P->p[j]; p < P->p[j + 1]; ++p) { const double a = fabs(P->x[p]); const int i = P->i[p]; if (a > dn[j]) dn[j] = a; if (a > dn[i]) dn[i] = a; }
        for (int j = 0; j < n; ++j) colmean += dn[j]; colmean /= n;
        double qn = vnorm_inf(q, n);
        double ct = colmean > qn ? colmean : qn;
        if (ct < 1e-4) ct = 1.0; else if (ct > 1e4) ct = 1e4;
        ct = 1.0 / ct;
        for (int p = 0; p < P->nnz; ++p) P->x[p] *= ct;
        for (int j = 0; j < n; ++j) q[j] *= ct;
        I->cscale *= ct;
    }
    for (int j = 0; j < n; ++j) I->Dinv[j] = 1.0 / I->D[j];
    for (int i = 0; i < m; ++i) I->Einv[i] = 1.0 / I->E[i];
    free(dn); free(en);
}

static csc upper_from_dense(const double* P, int n)
{
    int nnz = 0;
    for (int j = 0; j < n; ++j) for (int i = 0; i <= j; ++i) nnz += (P[(size_t)i * n + j] != 0.0);
    csc U = csc_alloc(n, n, nnz); int c = 0;
    for (int j = 0; j < n; ++j) { U.p[j] = c; for (int i = 0; i <= j; ++i) if (P[(size_t)i * n + j] != 0.0) { U.i[c] = i; U.x[c] = P[(size_t)i * n + j]; ++c; } }
    U.p[n] = c; return U;
}
static void sym_upper_mv(const csc* U, const double* x, double* y)
{
    memset(y, 0, sizeof(double) * U->n);
    for (int j = 0; j < U->n; ++j) for (int p = U->p[j]; p < U->p[j + 1]; ++p) { const int i = U->i[p]; y[i] += U->x[p] * x[j]; if (i != j) y[j] += U->x[p] * x[i]; }
}

/* setup = OsqpEigen initSolver (IMPCProblem.cpp:221-255): scaling, rho vector, KKT ordering + factorisation */
static void osqp_setup(ref_inst* I, csc Anew)
{
    const int n = I->nvar, m = I->ncon;
    osqp_free_ws(I);
    I->Ps = upper_from_dense(I->P, n);
    I->As = Anew;
    I->D = (double*)malloc(sizeof(double) * n); I->Dinv = (double*)malloc(sizeof(double) * n);
    I->E = (double*)malloc(sizeof(double) * m); I->Einv = (double*)malloc(sizeof(double) * m);
    I->qs = (double*)malloc(sizeof(double) * n); I->ls = (double*)malloc(sizeof(double) * m); I->us = (double*)malloc(sizeof(double) * m);
    I->rho_vec = (double*)malloc(sizeof(double) * m); I->ctype = (int*)malloc(sizeof(int) * m);
    I->xt = (double*)calloc(n, sizeof(double)); I->zt = (double*)calloc(m, sizeof(double));
    I->xprev = (double*)calloc(n, sizeof(double)); I->zprev = (double*)calloc(m, sizeof(double));
    I->rhs = (double*)calloc(n + m, sizeof(double)); I->tmpn = (double*)calloc(n + m, sizeof(double));
    I->tmpm = (double*)calloc(m, sizeof(double)); I->tmpm2 = (double*)calloc(m, sizeof(double));
    memcpy(I->qs, I->q, sizeof(double) * n);
    scale_problem(I, &I->Ps, &I->As, I->qs);
    for (int i = 0; i < m; ++i) { I->ls[i] = I->E[i] * I->l[i]; I->us[i] = I->E[i] * I->u[i]; }
    I->rho = I->cfg.rho;
    set_rho_vec(I);
    I->initialised = 1;
    kkt_build(I);
    kkt_refresh(I);
}

/* data update = updateGradient / updateLinearConstraintsMatrix / updateBounds (IMPCProblem.cpp:263-277) */
static void osqp_update(ref_inst* I, csc Anew)
{
    const int n = I->nvar, m = I->ncon;
    if (!csc_same_pattern(&I->As, &Anew))
    { /* OsqpEigen re-initialises the solver when the sparsity pattern of A changes; the warm start is kept */
        osqp_setup(I, Anew);
        return;
    }
    for (int j = 0; j < n; ++j) for (int p = Anew.p[j]; p < Anew.p[j + 1]; ++p) I->As.x[p] = Anew.x[p] * I->D[j] * I->E[Anew.i[p]];
    csc_free(&Anew);
    for (int j = 0; j < n; ++j) I->qs[j] = I->cscale * I->D[j] * I->q[j];
    for (int i = 0; i < m; ++i) { I->ls[i] = I->E[i] * I->l[i]; I->us[i] = I->E[i] * I->u[i]; }
    set_rho_vec(I);
    kkt_refresh(I);
}

static void compute_residuals(ref_inst* I, double* prim, double* dual, double* eps_p, double* eps_d, double* nprim, double* ndual)
{
    const int n = I->nvar, m = I->ncon;
    /* primal: || E^-1 (A x - z) ||_inf */
    csc_mv(&I->As, I->x, I->tmpm);
    double r = 0, nax = 0, nz = 0;
    for (int i = 0; i < m; ++i) { const double a = fabs(I->Einv[i] * (I->tmpm[i] - I->z[i])); if (a > r) r = a; const double b = fabs(I->Einv[i] * I->tmpm[i]); if (b > nax) nax = b; const double c = fabs(I->Einv[i] * I->z[i]); if (c > nz) nz = c; }
    *prim = r; *nprim = nax > nz ? nax : nz;
    *eps_p = I->cfg.eps_abs + I->cfg.eps_rel * (*nprim);
    /* dual: || D^-1 (P x + q + A^T y) ||_inf / c */
    double* Px = I->tmpn; double* Aty = I->tmpn + n;
    sym_upper_mv(&I->Ps, I->x, Px);
    csc_mtv(&I->As, I->y, I->rhs); /* rhs as scratch (n entries) */
    double d = 0, npx = 0, naty = 0, nq = 0;
    for (int j = 0; j < n; ++j)
    {
        const double a = fabs(I->Dinv[j] * (Px[j] + I->qs[j] + I->rhs[j])); if (a > d) d = a;
        const double b = fabs(I->Dinv[j] * Px[j]); if (b > npx) npx = b;
        const double c2 = fabs(I->Dinv[j] * I->rhs[j]); if (c2 > naty) naty = c2;
        const double e = fabs(I->Dinv[j] * I->qs[j]); if (e > nq) nq = e;
    }
    (void)Aty;
    double nd = npx > naty ? npx : naty; if (nq > nd) nd = nq;
    *dual = d / I->cscale; *ndual = nd / I->cscale;
    *eps_d = I->cfg.eps_abs + I->cfg.eps_rel * (*ndual);
}

static void polish(ref_inst* I);

/* solveProblem (IMPCProblem.cpp:279): warm-started ADMM + polish.  Returns status (1 solved, 2 inaccurate/max iter) */
static int osqp_solve(ref_inst* I)
{
    const int n = I->nvar, m = I->ncon; const ref_config* c = &I->cfg;
    const double alpha = c->alpha_relax, sigma = c->sigma;
    /* warm start: x, y from the previous solution (scaled); z = A x */
    for (int j = 0; j < n; ++j) I->x[j] = I->Dinv[j] * I->sol[j];
    for (int i = 0; i < m; ++i) I->y[i] = I->cscale * I->Einv[i] * I->ysol[i];
    csc_mv(&I->As, I->x, I->z);
    int it, status = 2;
    for (it = 1; it <= c->max_iter; ++it)
    {
        memcpy(I->xprev, I->x, sizeof(double) * n); memcpy(I->zprev, I->z, sizeof(double) * m);
        for (int j = 0; j < n; ++j) I->rhs[j] = sigma * I->xprev[j] - I->qs[j];
        for (int i = 0; i < m; ++i) I->rhs[n + i] = I->zprev[i] - I->y[i] / I->rho_vec[i];
        ldl_solve(&I->F, I->rhs);
        for (int j = 0; j < n; ++j) I->xt[j] = I->rhs[j];
        for (int i = 0; i < m; ++i) I->zt[i] = I->zprev[i] + (I->rhs[n + i] - I->y[i]) / I->rho_vec[i];
        for (int j = 0; j < n; ++j) I->x[j] = alpha * I->xt[j] + (1 - alpha) * I->xprev[j];
        for (int i = 0; i < m; ++i)
        {
            const double zr = alpha * I->zt[i] + (1 - alpha) * I->zprev[i];
            double zn = zr + I->y[i] / I->rho_vec[i];
            if (zn < I->ls[i]) zn = I->ls[i]; else if (zn > I->us[i]) zn = I->us[i];
            I->z[i] = zn;
            I->y[i] += I->rho_vec[i] * (zr - zn);
        }
        const int check = (c->check_termination > 0 && it % c->check_termination == 0);
        const int adapt = (c->adaptive_rho && c->adaptive_rho_interval > 0 && it % c->adaptive_rho_interval == 0);
        if (check || adapt)
        {
            double pr, du, ep, ed, npn, ndn;
            compute_residuals(I, &pr, &du, &ep, &ed, &npn, &ndn);
            if (check && pr <= ep && du <= ed) { status = 1; break; }
            if (adapt)
            {
                const double pn = pr / (npn + 1e-10), dn = du / (ndn + 1e-10);
                double rn = I->rho * sqrt(pn / (dn + 1e-10));
                if (rn < 1e-6) rn = 1e-6; if (rn > 1e6) rn = 1e6;
                if (rn > I->rho * c->adaptive_rho_tolerance || rn < I->rho / c->adaptive_rho_tolerance)
                {
                    I->rho = rn; set_rho_vec(I); kkt_refresh(I);
                }
            }
        }
    }
    if (it > c->max_iter)
    { /* last check */
        double pr, du, ep, ed, npn, ndn; compute_residuals(I, &pr, &du, &ep, &ed, &npn, &ndn);
        status = (pr <= ep && du <= ed) ? 1 : 2; it = c->max_iter;
    }
    I->iters = it;
    /* unscale */
    for (int j = 0; j < n; ++j) I->sol[j] = I->D[j] * I->x[j];
    for (int i = 0; i < m; ++i) I->ysol[i] = I->E[i] * I->y[i] / I->cscale;
    I->polished = 0;
    if (status == 1 && c->polish) polish(I);
    I->status = status;
    return status;
}

/* OSQP polishing: guess the active set from (z, y), solve the reduced KKT system regularised by delta with
 * iterative refinement, accept if the residuals do not get worse */
static void polish(ref_inst* I)
{
    const int n = I->nvar, m = I->ncon; const double delta = I->cfg.delta;
    int* act = (int*)malloc(sizeof(int) * m); int na = 0; double* bact = (double*)malloc(sizeof(double) * m);
    for (int i = 0; i < m; ++i)
    {
        if (I->z[i] - I->ls[i] < -I->y[i]) { act[na] = i; bact[na] = I->ls[i]; na++; }
        else if (I->us[i] - I->z[i] < I->y[i]) { act[na] = i; bact[na] = I->us[i]; na++; }
    }
    /* reduced KKT (scaled data): [[P + delta I, Ared^T],[Ared, -delta I]] */
    const int nk = n + na;
    int* rowmap = (int*)malloc(sizeof(int) * m); for (int i = 0; i < m; ++i) rowmap[i] = -1;
    for (int a = 0; a < na; ++a) rowmap[act[a]] = a;
    int nnz = 0;
    for (int j = 0; j < n; ++j) { int hd = 0; for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p) { nnz++; if (I->Ps.i[p] == j) hd = 1; } if (!hd) nnz++; }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p) if (rowmap[I->As.i[p]] >= 0) nnz++;
    nnz += na;
    /* build upper CSC (unpermuted) */
    int* Up = (int*)calloc(nk + 1, sizeof(int)); int* Ui = (int*)malloc(sizeof(int) * nnz); double* Ux = (double*)malloc(sizeof(double) * nnz);
    double* Ux0 = (double*)malloc(sizeof(double) * nnz); /* unregularised values for refinement */
    for (int j = 0; j < n; ++j) { int hd = 0; for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p) { Up[j + 1]++; if (I->Ps.i[p] == j) hd = 1; } if (!hd) Up[j + 1]++; }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p) if (rowmap[I->As.i[p]] >= 0) Up[n + rowmap[I->As.i[p]] + 1]++;
    for (int a = 0; a < na; ++a) Up[n + a + 1]++;
    for (int j = 0; j < nk; ++j) Up[j + 1] += Up[j];
    int* fill = (int*)calloc(nk, sizeof(int));
    for (int j = 0; j < n; ++j)
    {
        int hd = 0;
        for (int p = I->Ps.p[j]; p < I->Ps.p[j + 1]; ++p)
        {
            const int pos = Up[j] + fill[j]++; Ui[pos] = I->Ps.i[p]; Ux0[pos] = I->Ps.x[p]; Ux[pos] = I->Ps.x[p] + (I->Ps.i[p] == j ? delta : 0.0);
            if (I->Ps.i[p] == j) hd = 1;
        }
        if (!hd) { const int pos = Up[j] + fill[j]++; Ui[pos] = j; Ux0[pos] = 0.0; Ux[pos] = delta; }
    }
    for (int j = 0; j < n; ++j) for (int p = I->As.p[j]; p < I->As.p[j + 1]; ++p)
    {
        const int a = rowmap[I->As.i[p]]; if (a < 0) continue;
        const int pos = Up[n + a] + fill[n + a]++; Ui[pos] = j; Ux[pos] = Ux0[pos] = I->As.x[p];
    }
    for (int a = 0; a < na; ++a) { const int pos = Up[n + a] + fill[n + a]++; Ui[pos] = n + a; Ux0[pos] = 0.0; Ux[pos] = -delta; }
    ldl F; memset(&F, 0, sizeof(F)); F.n = nk;
    F.perm = (int*)malloc(sizeof(int) * nk); F.iperm = (int*)malloc(sizeof(int) * nk);
    { /* the reduced KKT is a principal submatrix of the full one: restrict the full ordering to it */
        int k2 = 0;
        for (int k = 0; k < I->F.n; ++k)
        {
            const int v = I->F.perm[k];
            if (v < n) F.perm[k2++] = v;
            else if (rowmap[v - n] >= 0) F.perm[k2++] = n + rowmap[v - n];
        }
    }
    for (int k = 0; k < nk; ++k) F.iperm[F.perm[k]] = k;
    F.Kp = (int*)calloc(nk + 1, sizeof(int)); F.Ki = (int*)malloc(sizeof(int) * nnz); F.Kx = (double*)calloc(nnz, sizeof(double)); F.Kmap = (int*)malloc(sizeof(int) * nnz);
    int* c2 = (int*)calloc(nk + 1, sizeof(int));
    for (int j = 0; j < nk; ++j) for (int p = Up[j]; p < Up[j + 1]; ++p) { const int a = F.iperm[Ui[p]], b = F.iperm[j]; c2[(a > b ? a : b) + 1]++; }
    for (int j = 0; j < nk; ++j) F.Kp[j + 1] = F.Kp[j] + c2[j + 1];
    memset(fill, 0, sizeof(int) * nk);
    for (int j = 0; j < nk; ++j) for (int p = Up[j]; p < Up[j + 1]; ++p)
    {
        const int a = F.iperm[Ui[p]], b = F.iperm[j]; const int col = a > b ? a : b, row = a > b ? b : a;
        const int pos = F.Kp[col] + fill[col]++; F.Ki[pos] = row; F.Kx[pos] = Ux[p];
    }
    F.knnz = nnz;
    F.Lp = (int*)calloc(nk + 1, sizeof(int)); F.Parent = (int*)malloc(sizeof(int) * nk); F.Lnz = (int*)malloc(sizeof(int) * nk);
    F.Flag = (int*)malloc(sizeof(int) * nk); F.Pattern = (int*)malloc(sizeof(int) * nk); F.Y = (double*)calloc(nk, sizeof(double));
    F.D = (double*)malloc(sizeof(double) * nk); F.Dinv = (double*)malloc(sizeof(double) * nk); F.work = (double*)malloc(sizeof(double) * nk);
    ldl_symbolic(&F);
    F.Li = (int*)malloc(sizeof(int) * (F.Lp[nk] > 0 ? F.Lp[nk] : 1)); F.Lx = (double*)malloc(sizeof(double) * (F.Lp[nk] > 0 ? F.Lp[nk] : 1));
    ldl_numeric(&F);
    /* rhs = [-q; b_act]; iterative refinement on the unregularised system */
    double* rhs = (double*)malloc(sizeof(double) * nk); double* sol = (double*)calloc(nk, sizeof(double)); double* res = (double*)malloc(sizeof(double) * nk);
    for (int j = 0; j < n; ++j) rhs[j] = -I->qs[j];
    for (int a = 0; a < na; ++a) rhs[n + a] = bact[a];
    memcpy(sol, rhs, sizeof(double) * nk); ldl_solve(&F, sol);
    for (int itr = 0; itr < I->cfg.polish_refine_iter; ++itr)
    {
        /* res = rhs - K0 sol (K0 symmetric from upper Ux0) */
        memcpy(res, rhs, sizeof(double) * nk);
        for (int j = 0; j < nk; ++j) for (int p = Up[j]; p < Up[j + 1]; ++p) { const int i = Ui[p]; res[i] -= Ux0[p] * sol[j]; if (i != j) res[j] -= Ux0[p] * sol[i]; }
        ldl_solve(&F, res);
        for (int k = 0; k < nk; ++k) sol[k] += res[k];
    }
    /* candidate (x_pol, z_pol = A x_pol, y_pol) ; accept if residuals improve */
    double* xs = (double*)malloc(sizeof(double) * n); double* ys = (double*)calloc(m, sizeof(double)); double* zs = (double*)malloc(sizeof(double) * m);
    memcpy(xs, sol, sizeof(double) * n);
    for (int a = 0; a < na; ++a) ys[act[a]] = sol[n + a];
    csc_mv(&I->As, xs, zs);
    for (int i = 0; i < m; ++i) { if (zs[i] < I->ls[i]) zs[i] = I->ls[i]; else if (zs[i] > I->us[i]) zs[i] = I->us[i]; }
    double pr0, du0, ep, ed, a1, a2, pr1, du1;
    compute_residuals(I, &pr0, &du0, &ep, &ed, &a1, &a2);
    double* xk = I->x; double* yk = I->y; double* zk = I->z;
    I->x = xs; I->y = ys; I->z = zs;
    compute_residuals(I, &pr1, &du1, &ep, &ed, &a1, &a2);
    const int ok = (pr1 < pr0 && du1 < du0) || (pr1 < pr0 && du0 < 1e-10) || (du1 < du0 && pr0 < 1e-10);
    if (ok)
    {
        memcpy(xk, xs, sizeof(double) * n); memcpy(yk, ys, sizeof(double) * m); memcpy(zk, zs, sizeof(double) * m);
        I->polished = 1;
    }
    I->x = xk; I->y = yk; I->z = zk;
    if (ok)
    {
        for (int j = 0; j < n; ++j) I->sol[j] = I->D[j] * I->x[j];
        for (int i = 0; i < m; ++i) I->ysol[i] = I->E[i] * I->y[i] / I->cscale;
    }
    free(xs); free(ys); free(zs); free(rhs); free(sol); free(res); free(Up); free(Ui); free(Ux); free(Ux0); free(fill); free(c2);
    free(act); free(bact); free(rowmap); ldl_free(&F);
}

/* ---------------------------------------------------------------------------------------------------- */
/* public C interface (ctypes)                                                                           */
void ref_default_settings(ref_config* c)
{
    c->rho = 0.1; c->sigma = 1e-6; c->alpha_relax = 1.6; c->eps_abs = 1e-3; c->eps_rel = 1e-3; c->delta = 1e-6;
    c->max_iter = 4000; c->check_termination = 25; c->scaling_iters = 10; c->adaptive_rho = 1; c->adaptive_rho_interval = 50;
    c->adaptive_rho_tolerance = 5.0; c->polish = 1; c->polish_refine_iter = 3;
}

void* ref_create(const ref_config* cfg)
{
    /* keep the per-tick scratch allocations inside the malloc arenas: mmap/munmap of >128 KiB blocks on every
     * tick serialises the OpenMP threads on the process address-space lock */
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);
    ref_inst* I = (ref_inst*)calloc(1, sizeof(ref_inst));
    I->cfg = *cfg;
    I->N = cfg->n_iter; I->Ns = cfg->n_small; I->Nc = cfg->n_ctrl; I->NC = I->N - I->Ns + 1; I->nblk = I->Nc - I->Ns + 1;
    I->nvar = NX * (I->N + 1) + NJ * I->Nc + NT * I->nblk;
    I->ncon = NX * I->N + NX + NT * (I->N - I->Ns + 1);
    I->ratio = (int)lround(cfg->period_large / cfg->period_small);
    const double beta2 = (cfg->period_large - I->Ns * cfg->period_small) / (I->Ns * (I->Ns - 1)), beta1 = cfg->period_small - beta2;
    for (int k = 0; k < I->N; ++k)
        I->dt[k] = k < I->Ns ? (beta1 * (k + 1) + beta2 * (k + 1) * (k + 1)) - (beta1 * k + beta2 * k * k) : cfg->period_large;
    I->vmax = jv(cfg->jc, (cfg->throttle_max - cfg->jn[2]) / cfg->jn[3]);
    I->vmin = jv(cfg->jc, (cfg->throttle_min - cfg->jn[2]) / cfg->jn[3]);
    I->win = (double*)calloc(12 * I->NC, sizeof(double));
    I->P = (double*)calloc((size_t)I->nvar * I->nvar, sizeof(double)); I->q = (double*)calloc(I->nvar, sizeof(double));
    I->A = (double*)calloc((size_t)I->ncon * I->nvar, sizeof(double)); I->l = (double*)calloc(I->ncon, sizeof(double)); I->u = (double*)calloc(I->ncon, sizeof(double));
    I->x = (double*)calloc(I->nvar, sizeof(double)); I->z = (double*)calloc(I->ncon, sizeof(double)); I->y = (double*)calloc(I->ncon, sizeof(double));
    I->sol = (double*)calloc(I->nvar, sizeof(double)); I->ysol = (double*)calloc(I->ncon, sizeof(double));
    return I;
}

void ref_destroy(void* h)
{
    ref_inst* I = (ref_inst*)h; if (!I) return;
    osqp_free_ws(I);
    free(I->win); free(I->P); free(I->q); free(I->A); free(I->l); free(I->u); free(I->x); free(I->z); free(I->y); free(I->sol); free(I->ysol);
    free(I);
}

/* IMPCProblem::configure: initialise the persistent state and run tick 0 of every counter */
void ref_configure(void* h, const double* pk, const double* joint_pos_sel)
{
    ref_inst* I = (ref_inst*)h;
    for (int a = 0; a < 3; ++a) { I->p_init[a] = pk[PK_PCOM + a]; I->rpy_init[a] = I->rpy_old[a] = pk[PK_RPY + a]; I->nturns[a] = 0; I->p_ref[a] = I->rpy_ref[a] = 0; }
    for (int a = 0; a < NJ; ++a) I->qref0[a] = I->qacc[a] = joint_pos_sel[a];
    I->ref_idx = 0; I->alpha_idx = 0;
    double col[12]; ref_column(I, pk, 0, col);
    for (int r = 0; r < 12; ++r) for (int k = 0; k < I->NC; ++k) I->win[r * I->NC + k] = col[r];
    I->ref_counter = I->thr_counter = I->ratio - 1;
    I->first_update = 1;
    assemble(I, pk);
    I->first_update = 1; /* IMPCProblem::update re-sums the Hessian on its first call (:152-175) */
    memset(I->out, 0, sizeof(I->out));
    memset(I->sol, 0, sizeof(double) * I->nvar); memset(I->ysol, 0, sizeof(double) * I->ncon);
}

void ref_update(void* h, const double* pk) { assemble((ref_inst*)h, pk); }

/* IMPCProblem::solve + VariableSamplingMPC::solveMPC */
int ref_solve(void* h)
{
    ref_inst* I = (ref_inst*)h;
    csc An = dense_to_csc(I->A, I->ncon, I->nvar);   /* m_linearMatrix.sparseView(), :211 */
    if (!I->initialised) osqp_setup(I, An); else osqp_update(I, An);
    const int status = osqp_solve(I);
    if (status == 1)
    { /* outputs only when Solved (variableSamplingMPC.cpp:91) */
        const int N = I->N, base = NX * (N + 1), tbase = base + I->Nc * NJ;
        for (int a = 0; a < NJ; ++a) { I->out[a] = I->sol[base + a]; I->qacc[a] += I->sol[base + a]; I->out[46 + a] = I->qacc[a]; }
        for (int j = 0; j < NT; ++j)
        {
            I->out[8 + j] = destd_u(I->cfg.jc, I->cfg.jn, I->sol[tbase + j]);
            I->out[12 + j] = I->sol[NX + 12 + j]; I->out[16 + j] = I->sol[NX + 16 + j];
        }
        for (int r = 0; r < NX; ++r) I->out[20 + r] = I->sol[N * NX + r];
    }
    return status;
}

void ref_get_output(void* h, double* out54) { memcpy(out54, ((ref_inst*)h)->out, sizeof(double) * 54); }
void ref_get_solution(void* h, double* z) { ref_inst* I = (ref_inst*)h; memcpy(z, I->sol, sizeof(double) * I->nvar); }
int ref_nvar(void* h) { return ((ref_inst*)h)->nvar; }
int ref_ncon(void* h) { return ((ref_inst*)h)->ncon; }
int ref_iters(void* h) { return ((ref_inst*)h)->iters; }
int ref_polished(void* h) { return ((ref_inst*)h)->polished; }
int ref_refactors(void* h) { return ((ref_inst*)h)->n_refactor; }
void ref_get_qp(void* h, double* P, double* q, double* A, double* l, double* u)
{
    ref_inst* I = (ref_inst*)h;
    if (P) memcpy(P, I->P, sizeof(double) * (size_t)I->nvar * I->nvar);
    if (q) memcpy(q, I->q, sizeof(double) * I->nvar);
    if (A) memcpy(A, I->A, sizeof(double) * (size_t)I->ncon * I->nvar);
    if (l) memcpy(l, I->l, sizeof(double) * I->ncon);
    if (u) memcpy(u, I->u, sizeof(double) * I->ncon);
}

/* batch driver for the CPU baseline: n independent instances (packs are AoS rows of 359 doubles), `threads`
 * OpenMP threads over instances; each does configure(nominal) once outside the timed part (done by the
 * caller through ref_configure) and then update + solve per tick. */
void ref_tick_batch(void** handles, int n, const double* packs, int threads, int* status_out)
{
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
#endif
    for (int i = 0; i < n; ++i)
    {
        ref_update(handles[i], packs + (size_t)i * PACKN);
        const int s = ref_solve(handles[i]);
        if (status_out) status_out[i] = s;
    }
}
