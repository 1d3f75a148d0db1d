// Device-resident closed loop: surrogate plant + pack builder, one thread per instance.
//
// Replaces, for batched rollouts, the loop body of the reference driver around the MPC tick
// (src/variable_sampling_mpc.py:106-161): sim.update_robot_state() -> [update + solveMPC] -> feedback of
// throttle / desired thrust / desired thrust rate / joint references into QPInput (:124-131) -> sim.step(n).
// MuJoCo, the URDF and iDynTree are outside the hot path and not available: the plant is a
// SURROGATE that integrates the MPC's own nonlinear model with frozen body-frame kinematics:
//   jets      : Tdd = sigma_T (f(T,Td) + g(T,Td) v(u)), semi-implicit Euler exactly as
//               src/mujoco_lib/jet_kalman_filter.py:30-45 (Td += Tdd dt; T += Td dt)
//   momentum  : h_lin^w' = alpha_g m g + sum_i (T_i + dT_i) R a_i ;  h_ang^B' = -w_B x h_ang^B + sum_i (T_i + dT_i) r_i x a_i
//   pose      : p' = h_lin^w / m ;  rpy' = W^-1(rpy) w_B,  w_B = I_B^-1 h_ang^B
//   joints    : position-controlled, q = q_cmd (the accumulated MPC joint reference); the jet frames follow them through
//               first-order kinematics about q0 with frozen relative Jacobians (the sensitivities behind the MPC's
//               Lambda matrices, systemDynamicsVSMPC.cpp:159-226,321-350):
//               a_i(q) = a_i0 + (J^w_rel,i dq) x a_i0 ,  r_i(q) = r_i0 + (J^lin_i - J^CoM) dq ,  dq = q - q0
// and rebuilds the getter-level pack (include/vsmpc.h VSMPC_PK_*) from the plant state every tick.
#include "vsmpc_common.cuh"
#include "vsmpc_plant.cuh"

namespace vsmpc
{

__device__ __forceinline__ void rpy_to_R(const double* rpy, double* R)
{
    const double cr = cos(rpy[0]), sr = sin(rpy[0]), cp = cos(rpy[1]), sp = sin(rpy[1]), cy = cos(rpy[2]), sy = sin(rpy[2]);
    R[0] = cy * cp; R[1] = cy * sp * sr - sy * cr; R[2] = cy * sp * cr + sy * sr;
    R[3] = sy * cp; R[4] = sy * sp * sr + cy * cr; R[5] = sy * sp * cr - cy * sr;
    R[6] = -sp;     R[7] = cp * sr;                R[8] = cp * cr;
}

__device__ __forceinline__ void mat3_vec(const double* M, const double* v, double* o)
{
    o[0] = M[0] * v[0] + M[1] * v[1] + M[2] * v[2];
    o[1] = M[3] * v[0] + M[4] * v[1] + M[5] * v[2];
    o[2] = M[6] * v[0] + M[7] * v[1] + M[8] * v[2];
}
__device__ __forceinline__ void mat3T_vec(const double* M, const double* v, double* o)
{
    o[0] = M[0] * v[0] + M[3] * v[1] + M[6] * v[2];
    o[1] = M[1] * v[0] + M[4] * v[1] + M[7] * v[2];
    o[2] = M[2] * v[0] + M[5] * v[1] + M[8] * v[2];
}
__device__ __forceinline__ void cross3(const double* a, const double* b, double* o)
{
    o[0] = a[1] * b[2] - a[2] * b[1];
    o[1] = a[2] * b[0] - a[0] * b[2];
    o[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ bool inv3(const double* M, double* o)
{
    const double c0 = M[4] * M[8] - M[5] * M[7], c1 = M[5] * M[6] - M[3] * M[8], c2 = M[3] * M[7] - M[4] * M[6];
    const double det = M[0] * c0 + M[1] * c1 + M[2] * c2;
    const double id = 1.0 / det;
    o[0] = c0 * id; o[1] = (M[2] * M[7] - M[1] * M[8]) * id; o[2] = (M[1] * M[5] - M[2] * M[4]) * id;
    o[3] = c1 * id; o[4] = (M[0] * M[8] - M[2] * M[6]) * id; o[5] = (M[2] * M[3] - M[0] * M[5]) * id;
    o[6] = c2 * id; o[7] = (M[1] * M[6] - M[0] * M[7]) * id; o[8] = (M[0] * M[4] - M[1] * M[3]) * id;
    return det != 0.0 && isfinite(id);
}

// thrust axes / arms in the body frame at joint command q (first-order kinematics about pm.q0)
__device__ __forceinline__ void jet_frames_body(const PlantModel& pm, const double* q, double (&aB)[NT][3], double (&rB)[NT][3])
{
    double dq[NJ];
#pragma unroll
    for (int b = 0; b < NJ; ++b)
        dq[b] = q[b] - pm.q0[b];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        double w[3], d[3];
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            double sw = 0.0, sd = 0.0;
#pragma unroll
            for (int b = 0; b < NJ; ++b)
            {
                sw = fma(pm.J_rel_ang_body[j * 24 + a * NJ + b], dq[b], sw);
                sd = fma(pm.J_jet_lin_body[j * 24 + a * NJ + b] - pm.J_com_body[a * NJ + b], dq[b], sd);
            }
            w[a] = sw;
            d[a] = sd;
        }
        double wxa[3];
        cross3(w, pm.jet_axes_body + 3 * j, wxa);
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            aB[j][a] = pm.jet_axes_body[3 * j + a] + wxa[a];
            rB[j][a] = pm.jet_pos_body[3 * j + a] + d[a];
        }
    }
}

// mode 0: build the pack from the plant state only (tick 0 / configure)
// mode 1: apply the MPC outputs of the tick just solved (feedback, src/variable_sampling_mpc.py:124-131),
//         integrate n_sub plant steps, record, build the next pack
__global__ void __launch_bounds__(128)
plant_kernel(const DeviceConfig* __restrict__ cfgp, const PlantModel* __restrict__ pmp, int B, int mode,
             double* __restrict__ ps, const double* __restrict__ pp, const double* __restrict__ out_rows,
             const int* __restrict__ status, double* __restrict__ pack, double* __restrict__ rec,
             const double* __restrict__ ip, const double* __restrict__ st, const double* __restrict__ thr_sub)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B)
        return;
    const DeviceConfig& cfg = *cfgp;
    const PlantModel& pm = *pmp;
    double jpar[IP_TMIN];   // jet coefficients (13) + normalisation (4) of this instance
#pragma unroll
    for (int a = 0; a < IP_TMIN; ++a)
        jpar[a] = ip ? ip[(size_t)a * B + i] : (a < IP_JN ? cfg.jc[a] : cfg.jn[a - IP_JN]);
    const Jet jet{jpar + IP_JC, jpar + IP_JN};
#define PS(r) ps[(size_t)(r) * B + i]
#define PP(r) pp[(size_t)(r) * B + i]
#define PK(r) pack[(size_t)(r) * B + i]
    double p[3], hl[3], rpy[3], ha[3], T[NT], Td[NT], u[NT], Tdes[NT], Tddes[NT], q[NJ];
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        p[a] = PS(PS_PCOM + a);
        hl[a] = PS(PS_HLIN_W + a);
        rpy[a] = PS(PS_RPY + a);
        ha[a] = PS(PS_HANG_B + a);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        T[j] = PS(PS_T + j);
        Td[j] = PS(PS_TD + j);
        u[j] = PS(PS_THROTTLE + j);
        Tdes[j] = PS(PS_TDES + j);
        Tddes[j] = PS(PS_TDDES + j);
    }
#pragma unroll
    for (int a = 0; a < NJ; ++a)
        q[a] = PS(PS_QCMD + a);
    const double mass = PP(PP_MASS);
    double Ib[9], Ibinv[9], dT[NT];
#pragma unroll
    for (int a = 0; a < 9; ++a)
        Ib[a] = PP(PP_INERTIA + a);
#pragma unroll
    for (int j = 0; j < NT; ++j)
        dT[j] = PP(PP_DTHRUST + j);
    inv3(Ib, Ibinv);

    if (mode == 1)
    {
        // feedback of the tick just solved; a non-solved instance holds its previous outputs, which is what
        // out_rows still contains (variableSamplingMPC.cpp:91)
        const double* o = out_rows + (size_t)i * VSMPC_OUT_DOUBLES;
#pragma unroll
        for (int j = 0; j < NT; ++j)
        {
            u[j] = o[VSMPC_OUT_THROTTLE + j];
            Tdes[j] = o[VSMPC_OUT_THRUST + j];
            Tddes[j] = o[VSMPC_OUT_THRUST_DOT + j];
        }
#pragma unroll
        for (int a = 0; a < NJ; ++a)
            q[a] = o[VSMPC_OUT_JOINTS_REF + a];
        // jet frames at the new joint command, torque arms r_i x a_i (body)
        double aBm[NT][3], rBm[NT][3], rxa[NT][3];
        jet_frames_body(pm, q, aBm, rBm);
#pragma unroll
        for (int j = 0; j < NT; ++j)
            cross3(rBm[j], aBm[j], rxa[j]);
        const double dt = pm.dt_sim;
        // the ground carries the share (1 - alpha_g) of the weight during take-off: the plant uses the gravity
        // compensation factor the MPC published for this tick (QPInput::setAlphaGravity, systemDynamicsVSMPC.cpp:307-311)
        const double alpha_g = st ? st[(size_t)ST_ALPHA * B + i] : 1.0;
        for (int sstep = 0; sstep < pm.n_sub; ++sstep)
        {
            double R[9];
            rpy_to_R(rpy, R);
            if (thr_sub)
            { // jet-NN mode: the EKF estimate of this plant step (jet_nn_ekf_kernel) is the applied thrust
#pragma unroll
                for (int j = 0; j < NT; ++j)
                {
                    T[j] = thr_sub[((size_t)(2 * sstep) * NT + j) * B + i];
                    Td[j] = thr_sub[((size_t)(2 * sstep + 1) * NT + j) * B + i];
                }
            }
            else
            { // jets (jet_kalman_filter.py:30-45)
#pragma unroll
                for (int j = 0; j < NT; ++j)
                {
                    const double Ts = jet.stdT(T[j]), Tds = jet.stdTd(Td[j]);
                    const double tdd = jet.f(Ts, Tds) + jet.g(Ts, Tds) * jet.v(jet.stdU(u[j]));
                    Td[j] += tdd * jpar[IP_JN + 1] * dt;
                    T[j] += Td[j] * dt;
                }
            }
            // momentum
            double fB[3] = {0, 0, 0}, tauB[3] = {0, 0, 0};
#pragma unroll
            for (int j = 0; j < NT; ++j)
            {
                const double Tj = T[j] + dT[j];
#pragma unroll
                for (int a = 0; a < 3; ++a)
                {
                    fB[a] += Tj * aBm[j][a];
                    tauB[a] += Tj * rxa[j][a];
                }
            }
            double fW[3], wB[3], wxh[3];
            mat3_vec(R, fB, fW);
            mat3_vec(Ibinv, ha, wB);
            cross3(wB, ha, wxh);
#pragma unroll
            for (int a = 0; a < 3; ++a)
            {
                hl[a] += dt * (alpha_g * mass * pm.gravity[a] + fW[a]);
                ha[a] += dt * (tauB[a] - wxh[a]);
            }
            // pose
            mat3_vec(Ibinv, ha, wB);
            const double s0 = sin(rpy[0]), c0 = cos(rpy[0]), t1 = tan(rpy[1]), c1 = cos(rpy[1]);
            const double rd0 = wB[0] + s0 * t1 * wB[1] + c0 * t1 * wB[2];
            const double rd1 = c0 * wB[1] - s0 * wB[2];
            const double rd2 = (s0 * wB[1] + c0 * wB[2]) / c1;
#pragma unroll
            for (int a = 0; a < 3; ++a)
                p[a] += dt * hl[a] / mass;
            rpy[0] += dt * rd0;
            rpy[1] += dt * rd1;
            rpy[2] += dt * rd2;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            PS(PS_PCOM + a) = p[a];
            PS(PS_HLIN_W + a) = hl[a];
            PS(PS_RPY + a) = rpy[a];
            PS(PS_HANG_B + a) = ha[a];
        }
#pragma unroll
        for (int j = 0; j < NT; ++j)
        {
            PS(PS_T + j) = T[j];
            PS(PS_TD + j) = Td[j];
            PS(PS_THROTTLE + j) = u[j];
            PS(PS_TDES + j) = Tdes[j];
            PS(PS_TDDES + j) = Tddes[j];
        }
#pragma unroll
        for (int a = 0; a < NJ; ++a)
            PS(PS_QCMD + a) = q[a];
        if (rec)
        {
            double* r = rec + (size_t)i * PLANT_REC;
#pragma unroll
            for (int a = 0; a < 3; ++a)
            {
                r[a] = p[a];
                r[3 + a] = rpy[a];
            }
#pragma unroll
            for (int j = 0; j < NT; ++j)
            {
                r[6 + j] = T[j];
                r[10 + j] = u[j];
            }
            r[14] = (double)status[i];
            r[15] = 0.0;
        }
    }

    // ---- pack of the current plant state (the formulas of the synthetic robot, synthetic.py::make_states) ----
    double R[9], wB[3], v3[3], c[3];
    double aBp[NT][3], rBp[NT][3];
    jet_frames_body(pm, q, aBp, rBp);
    rpy_to_R(rpy, R);
    mat3_vec(Ibinv, ha, wB);
#pragma unroll
    for (int a = 0; a < 9; ++a)
        PK(VSMPC_PK_WRB + a) = R[a];
    mat3_vec(R, wB, v3);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        PK(VSMPC_PK_OMEGA_WORLD + a) = v3[a];
        PK(VSMPC_PK_RPY + a) = rpy[a];
        PK(VSMPC_PK_GRAVITY + a) = pm.gravity[a];
        PK(VSMPC_PK_P_COM + a) = p[a];
    }
    PK(VSMPC_PK_MASS) = mass;
    mat3_vec(R, pm.com_from_base_body, c);     // p_com - p_base, world
#pragma unroll
    for (int a = 0; a < 3; ++a)
        PK(VSMPC_PK_BASE_POS + a) = p[a] - c[a];
    {
        // M_b = [m I, -m S(c); m S(c), R I_B R' + m S(c)'S(c)]
        const double S[9] = {0, -c[2], c[1], c[2], 0, -c[0], -c[1], c[0], 0};
        double RI[9], Iw[9];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                RI[a * 3 + b] = R[a * 3] * Ib[b] + R[a * 3 + 1] * Ib[3 + b] + R[a * 3 + 2] * Ib[6 + b];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
                Iw[a * 3 + b] = RI[a * 3] * R[b * 3] + RI[a * 3 + 1] * R[b * 3 + 1] + RI[a * 3 + 2] * R[b * 3 + 2];
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < 3; ++b)
            {
                const double StS = S[0 * 3 + a] * S[0 * 3 + b] + S[1 * 3 + a] * S[1 * 3 + b] + S[2 * 3 + a] * S[2 * 3 + b];
                PK(VSMPC_PK_MB + a * 6 + b) = a == b ? mass : 0.0;
                PK(VSMPC_PK_MB + a * 6 + 3 + b) = -mass * S[a * 3 + b];
                PK(VSMPC_PK_MB + (3 + a) * 6 + b) = mass * S[a * 3 + b];
                PK(VSMPC_PK_MB + (3 + a) * 6 + 3 + b) = Iw[a * 3 + b] + mass * StS;
            }
    }
    mat3T_vec(R, hl, v3);
#pragma unroll
    for (int a = 0; a < 3; ++a)
    {
        PK(VSMPC_PK_MOMENTUM_BODY + a) = v3[a];
        PK(VSMPC_PK_MOMENTUM_BODY + 3 + a) = ha[a];
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        double aw[3], rw[3], rxaw[3], t3[3];
        mat3_vec(R, aBp[j], aw);
        mat3_vec(R, rBp[j], rw);
        cross3(rw, aw, rxaw);
        // A_mom_body = [R'a_w ; R'(r_w x a_w)]  (Robot::getMatrixAmomJets(true), Robot.cpp:325-329)
        mat3T_vec(R, aw, t3);
        double t4[3];
        mat3T_vec(R, rxaw, t4);
#pragma unroll
        for (int a = 0; a < 3; ++a)
        {
            PK(VSMPC_PK_JET_AXES + 3 * j + a) = aw[a];
            PK(VSMPC_PK_JET_ARMS + 3 * j + a) = rw[a];
            PK(VSMPC_PK_AMOM_BODY + a * NT + j) = t3[a];
            PK(VSMPC_PK_AMOM_BODY + (3 + a) * NT + j) = t4[a];
        }
        // relative Jacobians: angular part is body-frame data; linear Jacobians are world-frame (R J_body)
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int b = 0; b < NJ; ++b)
            {
                PK(VSMPC_PK_J_REL_ANG + j * 24 + a * NJ + b) = pm.J_rel_ang_body[j * 24 + a * NJ + b];
                PK(VSMPC_PK_J_JET_LIN + j * 24 + a * NJ + b) = R[a * 3] * pm.J_jet_lin_body[j * 24 + b]
                                                              + R[a * 3 + 1] * pm.J_jet_lin_body[j * 24 + NJ + b]
                                                              + R[a * 3 + 2] * pm.J_jet_lin_body[j * 24 + 2 * NJ + b];
            }
        PK(VSMPC_PK_THRUST + j) = T[j];
        PK(VSMPC_PK_THRUST_DOT_EST + j) = Td[j];
        PK(VSMPC_PK_THRUST_DES + j) = Tdes[j];
        PK(VSMPC_PK_THRUST_DOT_DES + j) = Tddes[j];
        PK(VSMPC_PK_THROTTLE_PREV + j) = u[j];
    }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
        for (int b = 0; b < NJ; ++b)
            PK(VSMPC_PK_J_COM + a * NJ + b) = R[a * 3] * pm.J_com_body[b] + R[a * 3 + 1] * pm.J_com_body[NJ + b]
                                              + R[a * 3 + 2] * pm.J_com_body[2 * NJ + b];
#pragma unroll
    for (int a = 0; a < NJ; ++a)
        PK(VSMPC_PK_Q_CMD + a) = q[a];
#undef PS
#undef PP
#undef PK
}

// NeuralJetModel.get_state for four (thrust, normalised throttle) pairs by one warp (nn_jet_model.py:21-30,86-109): the
// LSTM cell from a zero state in float32 — hidden units over the lanes, the fc dot product by warp reduction.
// T is advanced in place; Td receives the thrust rate.
__device__ __forceinline__ void nn_jet_step4(const JetNN& nn, float (&T)[NT], const float (&un)[NT], float dtf, int lane,
                                             float (&Td)[NT])
{
    const float stdT = (float)nn.norm[1], meanT = (float)nn.norm[0];
    float part[NT] = {0.f, 0.f, 0.f, 0.f};
    float Tn[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j)
        Tn[j] = (float)(((double)T[j] - nn.norm[0]) / nn.norm[1]);
    for (int hh = lane; hh < NN_HID; hh += 32)
    {
        const float wi0 = nn.w_ih[hh * 2], wi1 = nn.w_ih[hh * 2 + 1], bi = nn.b[hh];
        const float wg0 = nn.w_ih[(2 * NN_HID + hh) * 2], wg1 = nn.w_ih[(2 * NN_HID + hh) * 2 + 1], bg = nn.b[2 * NN_HID + hh];
        const float wo0 = nn.w_ih[(3 * NN_HID + hh) * 2], wo1 = nn.w_ih[(3 * NN_HID + hh) * 2 + 1], bo = nn.b[3 * NN_HID + hh];
        const float fw = nn.fc_w[hh];
#pragma unroll
        for (int j = 0; j < NT; ++j)
        {
            const float gi = 1.f / (1.f + expf(-(wi0 * Tn[j] + wi1 * un[j] + bi)));
            const float gg = tanhf(wg0 * Tn[j] + wg1 * un[j] + bg);
            const float go = 1.f / (1.f + expf(-(wo0 * Tn[j] + wo1 * un[j] + bo)));
            part[j] = fmaf(fw, go * tanhf(gi * gg), part[j]);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
#pragma unroll
        for (int j = 0; j < NT; ++j)
            part[j] += __shfl_xor_sync(0xffffffffu, part[j], o);
    }
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        const float td = part[j] + nn.fc_b;
        T[j] = (Tn[j] + td * dtf) * stdT + meanT;
        Td[j] = td * stdT;
    }
}

// parity seam: one network step for n groups of four (thrust, throttle) pairs, one warp per group
__global__ void __launch_bounds__(128)
jet_nn_eval_kernel(const JetNN* __restrict__ nnp, int n_groups, float dt, const float* __restrict__ T_in,
                   const float* __restrict__ u_in, float* __restrict__ T_out, float* __restrict__ Td_out)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int gidx = blockIdx.x * 4 + warp;
    if (gidx >= n_groups)
        return;
    const JetNN& nn = *nnp;
    float T[NT], un[NT], Td[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        T[j] = T_in[gidx * NT + j];
        un[j] = (float)(((double)u_in[gidx * NT + j] - nn.norm[2]) / nn.norm[3]);
    }
    nn_jet_step4(nn, T, un, dt, lane, Td);
    if (lane < NT)
    {
        T_out[gidx * NT + lane] = lane == 0 ? T[0] : lane == 1 ? T[1] : lane == 2 ? T[2] : T[3];
        Td_out[gidx * NT + lane] = lane == 0 ? Td[0] : lane == 1 ? Td[1] : lane == 2 ? Td[2] : Td[3];
    }
}

cudaError_t launch_jet_nn_eval(const JetNN* d_nn, int n_groups, float dt, const float* T_in, const float* u_in, float* T_out,
                               float* Td_out, cudaStream_t s)
{
    jet_nn_eval_kernel<<<(n_groups + 3) / 4, 128, 0, s>>>(d_nn, n_groups, dt, T_in, u_in, T_out, Td_out);
    return cudaGetLastError();
}

// ---- neural jet plant + per-jet EKF, one warp per instance (jet-NN mode of the rollout) ----------------------------------
// Per plant step: (1) NeuralJetModel.get_state for the four jets (nn_jet_model.py:21-30,86-109): the LSTM cell from a zero
// state, float32 — hidden units over the lanes, the fc dot product by warp reduction; (2) SecondOrderJetModel.update per jet
// (jet_kalman_filter.py:56-65) on lanes 0..3, float64; the estimate of every plant step goes to thr_sub for plant_kernel.
__global__ void __launch_bounds__(128)
jet_nn_ekf_kernel(const DeviceConfig* __restrict__ cfgp, const PlantModel* __restrict__ pmp, const JetNN* __restrict__ nnp,
                  int B, double* __restrict__ ps, const double* __restrict__ out_rows, const double* __restrict__ ip,
                  double* __restrict__ thr_sub)
{
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int i = blockIdx.x * 4 + warp;
    if (i >= B)
        return;
    const DeviceConfig& cfg = *cfgp;
    const PlantModel& pm = *pmp;
    const JetNN& nn = *nnp;
#define PS(r) ps[(size_t)(r) * B + i]
    double jpar[IP_TMIN];
#pragma unroll
    for (int a = 0; a < IP_TMIN; ++a)
        jpar[a] = ip ? ip[(size_t)a * B + i] : (a < IP_JN ? cfg.jc[a] : cfg.jn[a - IP_JN]);
    const Jet jet{jpar + IP_JC, jpar + IP_JN};
    // this tick's throttle command (the feedback plant_kernel applies right after this kernel)
    float un[NT], Tnn[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j)
    {
        // the simulator keeps the throttle as a float32 tensor (ironcub_mujoco_simulator.py:190)
        const double uj = (double)(float)out_rows[(size_t)i * VSMPC_OUT_DOUBLES + VSMPC_OUT_THROTTLE + j];
        un[j] = (float)((uj - nn.norm[2]) / nn.norm[3]);
        Tnn[j] = (float)PS(PS_TNN + j);
    }
    // EKF state of jet `lane` (lanes 0..3)
    const int jl = lane & 3;
    double xe0 = PS(PS_T + jl), xe1 = PS(PS_TD + jl);
    double P00 = PS(PS_EKFP + 4 * jl), P01 = PS(PS_EKFP + 4 * jl + 1), P10 = PS(PS_EKFP + 4 * jl + 2), P11 = PS(PS_EKFP + 4 * jl + 3);
    const double ue = (double)(float)out_rows[(size_t)i * VSMPC_OUT_DOUBLES + VSMPC_OUT_THROTTLE + jl];
    const double dt = pm.dt_sim;
    const float dtf = (float)dt;
    for (int sstep = 0; sstep < pm.n_sub; ++sstep)
    {
        // ---- neural jet plant ----
        float Tdn[NT];
        nn_jet_step4(nn, Tnn, un, dtf, lane, Tdn);
        // ---- EKF of jet `lane` (lanes >= 4 compute a copy that is never stored) ----
        const double z0 = (double)(jl == 0 ? Tnn[0] : jl == 1 ? Tnn[1] : jl == 2 ? Tnn[2] : Tnn[3]);
        const double z1 = (double)(jl == 0 ? Tdn[0] : jl == 1 ? Tdn[1] : jl == 2 ? Tdn[2] : Tdn[3]);
        {
            const double v = jet.v(jet.stdU(ue));
            double Ts = jet.stdT(xe0), Tds = jet.stdTd(xe1);
            const double tdd = jet.f(Ts, Tds) + jet.g(Ts, Tds) * v;
            xe1 += tdd * jpar[IP_JN + 1] * dt;          // predict (jet_kalman_filter.py:30-45)
            xe0 += xe1 * dt;
            Ts = jet.stdT(xe0);                          // Jacobian at the predicted state (:58)
            Tds = jet.stdTd(xe1);
            const double hT = jet.df_dT(Ts, Tds) + jet.dg_dT(Ts, Tds) * v, hTd = jet.df_dTd(Ts, Tds) + jet.dg_dTd(Ts, Tds) * v;
            const double a10 = dt * hT, a11 = 1.0 + dt * hTd, a00 = 1.0 + dt * a10, a01 = dt * a11;
            // P <- A P A' + Q
            const double t00 = a00 * P00 + a01 * P10, t01 = a00 * P01 + a01 * P11, t10 = a10 * P00 + a11 * P10, t11 = a10 * P01 + a11 * P11;
            P00 = t00 * a00 + t01 * a01 + nn.Q[0];
            P01 = t00 * a10 + t01 * a11 + nn.Q[1];
            P10 = t10 * a00 + t11 * a01 + nn.Q[2];
            P11 = t10 * a10 + t11 * a11 + nn.Q[3];
            // K = P (P + R)^-1 ; x += K (z - x) ; P <- (I - K) P
            const double s00 = P00 + nn.R[0], s01 = P01 + nn.R[1], s10 = P10 + nn.R[2], s11 = P11 + nn.R[3];
            const double idet = 1.0 / (s00 * s11 - s01 * s10);
            const double i00 = s11 * idet, i01 = -s01 * idet, i10 = -s10 * idet, i11 = s00 * idet;
            const double k00 = P00 * i00 + P01 * i10, k01 = P00 * i01 + P01 * i11, k10 = P10 * i00 + P11 * i10, k11 = P10 * i01 + P11 * i11;
            const double e0 = z0 - xe0, e1 = z1 - xe1;
            xe0 += k00 * e0 + k01 * e1;
            xe1 += k10 * e0 + k11 * e1;
            const double n00 = (1.0 - k00) * P00 - k01 * P10, n01 = (1.0 - k00) * P01 - k01 * P11;
            const double n10 = -k10 * P00 + (1.0 - k11) * P10, n11 = -k10 * P01 + (1.0 - k11) * P11;
            P00 = n00; P01 = n01; P10 = n10; P11 = n11;
        }
        if (lane < NT)
        {
            thr_sub[((size_t)(2 * sstep) * NT + lane) * B + i] = xe0;
            thr_sub[((size_t)(2 * sstep + 1) * NT + lane) * B + i] = xe1;
        }
    }
    if (lane < NT)
    {
        PS(PS_TNN + lane) = (double)(lane == 0 ? Tnn[0] : lane == 1 ? Tnn[1] : lane == 2 ? Tnn[2] : Tnn[3]);
        PS(PS_EKFP + 4 * lane) = P00;
        PS(PS_EKFP + 4 * lane + 1) = P01;
        PS(PS_EKFP + 4 * lane + 2) = P10;
        PS(PS_EKFP + 4 * lane + 3) = P11;
    }
#undef PS
}

cudaError_t launch_jet_nn_ekf(const DeviceConfig* d_cfg, const PlantModel* d_pm, const JetNN* d_nn, int B, double* ps,
                              const double* out_rows, const double* ip, double* thr_sub, cudaStream_t s)
{
    jet_nn_ekf_kernel<<<(B + 3) / 4, 128, 0, s>>>(d_cfg, d_pm, d_nn, B, ps, out_rows, ip, thr_sub);
    return cudaGetLastError();
}

cudaError_t launch_plant(const DeviceConfig* d_cfg, const PlantModel* d_pm, int B, int mode, double* ps,
                         const double* pp, const double* out_rows, const int* status, double* pack, double* rec,
                         const double* ip, const double* st, const double* thr_sub, cudaStream_t s)
{
    const int threads = 128;
    plant_kernel<<<(B + threads - 1) / threads, threads, 0, s>>>(d_cfg, d_pm, B, mode, ps, pp, out_rows, status, pack, rec, ip, st, thr_sub);
    return cudaGetLastError();
}

} // namespace vsmpc
