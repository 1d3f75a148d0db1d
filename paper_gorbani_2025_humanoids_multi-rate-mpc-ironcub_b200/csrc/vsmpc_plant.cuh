// Surrogate plant of the device-resident closed loop: layouts shared by vsmpc_plant.cu and vsmpc_api.cu.
#pragma once
#include "vsmpc.h"

namespace vsmpc
{
// plant state rows (SoA, row = scalar, column = instance) — VSMPC_PS_* of include/vsmpc.h
constexpr int PS_PCOM = VSMPC_PS_P_COM, PS_HLIN_W = VSMPC_PS_LIN_MOM_WORLD, PS_RPY = VSMPC_PS_RPY,
              PS_HANG_B = VSMPC_PS_ANG_MOM_BODY, PS_T = VSMPC_PS_THRUST, PS_TD = VSMPC_PS_THRUST_DOT,
              PS_THROTTLE = VSMPC_PS_THROTTLE, PS_TDES = VSMPC_PS_THRUST_DES, PS_TDDES = VSMPC_PS_THRUST_DOT_DES,
              PS_QCMD = VSMPC_PS_Q_CMD, PS_TNN = VSMPC_PS_THRUST_NN, PS_EKFP = VSMPC_PS_EKF_P,
              PS_ROWS = VSMPC_PLANT_STATE_DOUBLES;
constexpr int NN_HID = 80;     // hidden units of the jet network (nn_jet_model.py:46)
constexpr int MAX_SUB = 16;    // plant steps per tick the jet-NN path records
// per-instance plant parameters
constexpr int PP_MASS = VSMPC_PP_MASS, PP_INERTIA = VSMPC_PP_INERTIA_BODY, PP_DTHRUST = VSMPC_PP_THRUST_DISTURBANCE,
              PP_ROWS = VSMPC_PLANT_PARAM_DOUBLES;
constexpr int PLANT_REC = VSMPC_ROLLOUT_REC_DOUBLES;

// neural jet plant + EKF constants (device copy)
struct JetNN
{
    float w_ih[4 * NN_HID * 2];
    float b[4 * NN_HID];      // b_ih + b_hh
    float fc_w[NN_HID];
    float fc_b;
    float pad;
    double norm[4];
    double R[4], Q[4];
};

struct PlantModel
{
    double com_from_base_body[3];
    double jet_pos_body[12];
    double jet_axes_body[12];
    double J_rel_ang_body[96];
    double J_jet_lin_body[96];
    double J_com_body[24];
    double gravity[3];
    double q0[8];
    double dt_sim;
    int n_sub;
};
} // namespace vsmpc
