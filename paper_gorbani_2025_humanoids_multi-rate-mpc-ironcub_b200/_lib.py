"""ctypes binding of libvsmpc.so (the C-ABI declared in include/vsmpc.h).

There is NO fallback: if the CUDA library is missing or cannot be loaded the import of any compute
entry point raises — the product path never routes through the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libvsmpc.so")

NX, NJ, NT = 26, 8, 4
PACK_DOUBLES = 359
OUT_DOUBLES = 54
OUT_DELTA_Q, OUT_THROTTLE, OUT_THRUST, OUT_THRUST_DOT, OUT_FINAL_STATE, OUT_JOINTS_REF = 0, 8, 12, 16, 20, 46
STATUS_SOLVED, STATUS_MAX_ITER, STATUS_NUMERICAL = 0, 1, 2
OK, ERR_ARG, ERR_CUDA, ERR_UNSUPPORTED, ERR_STATE = 0, 1, 2, 3, 4
REF_DOUBLES = 13
REF_ALPHA_GRAVITY, REF_POS_COM, REF_RPY, REF_MOMENTUM = 0, 1, 4, 7

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class VsmpcConfig(C.Structure):
    _fields_ = [
        ("n_iter", C.c_int), ("n_iter_small", C.c_int), ("control_horizon", C.c_int),
        ("period_mpc", C.c_double), ("period_large", C.c_double), ("period_small", C.c_double),
        ("use_jet_dynamic", C.c_int), ("use_estimated_thrust", C.c_int), ("joints_lambda_option", C.c_int),
        ("weight_com_pos", C.c_double * 3), ("weight_com_pos_error", C.c_double * 3),
        ("weight_lin_mom", C.c_double * 3), ("weight_rpy", C.c_double * 3),
        ("weight_rpy_error", C.c_double * 3), ("weight_ang_mom", C.c_double * 3),
        ("weight_delta_joint", C.c_double * 8),
        ("weight_throttle", C.c_double), ("weight_initial_throttle", C.c_double),
        ("weight_regularization_joint_pos", C.c_double),
        ("throttle_min", C.c_double), ("throttle_max", C.c_double),
        ("jet_coeff", C.c_double * 13), ("jet_norm", C.c_double * 4),
        ("alpha_gravity", c_double_p), ("alpha_len", C.c_int), ("alpha_fps", C.c_int),
        ("position_com", c_double_p), ("velocity_com", c_double_p), ("rpy", c_double_p),
        ("rpy_dot", c_double_p), ("traj_len", C.c_int), ("traj_fps", C.c_int),
        ("solver", C.c_int),
        ("use_joint_limits", C.c_int), ("joint_pos_min_deg", C.c_double * 8), ("joint_pos_max_deg", C.c_double * 8),
    ]


INSTANCE_PARAM_DOUBLES = 19
IP_JET_COEFF, IP_JET_NORM, IP_THROTTLE_MIN, IP_THROTTLE_MAX = 0, 13, 17, 18
PLANT_STATE_DOUBLES, PLANT_PARAM_DOUBLES, ROLLOUT_REC_DOUBLES = 60, 14, 16
PS_THRUST_NN, PS_EKF_P = 40, 44
PS_P_COM, PS_LIN_MOM_WORLD, PS_RPY, PS_ANG_MOM_BODY, PS_THRUST, PS_THRUST_DOT = 0, 3, 6, 9, 12, 16
PS_THROTTLE, PS_THRUST_DES, PS_THRUST_DOT_DES, PS_Q_CMD = 20, 24, 28, 32
PP_MASS, PP_INERTIA_BODY, PP_THRUST_DISTURBANCE = 0, 1, 10


class VsmpcPlantModel(C.Structure):
    _fields_ = [
        ("com_from_base_body", C.c_double * 3), ("jet_pos_body", C.c_double * 12),
        ("jet_axes_body", C.c_double * 12), ("J_rel_ang_body", C.c_double * 96),
        ("J_jet_lin_body", C.c_double * 96), ("J_com_body", C.c_double * 24), ("gravity", C.c_double * 3), ("q0", C.c_double * 8),
        ("dt_sim", C.c_double), ("n_sub", C.c_int),
    ]


EXPORTS = [
    "vsmpc_host_alloc", "vsmpc_host_free",
    "vsmpc_create", "vsmpc_destroy", "vsmpc_last_error", "vsmpc_set_stream", "vsmpc_n_var",
    "vsmpc_n_constraints", "vsmpc_n_instances", "vsmpc_configure", "vsmpc_set_state",
    "vsmpc_set_state_device", "vsmpc_solve", "vsmpc_solve_async", "vsmpc_wait", "vsmpc_get_output",
    "vsmpc_get_output_device", "vsmpc_get_output_async", "vsmpc_wait_output", "vsmpc_set_full_solution", "vsmpc_get_full_solution", "vsmpc_get_dynamics", "vsmpc_get_qp_vectors",
    "vsmpc_linearise", "vsmpc_solve_qp",
    "vsmpc_get_counts", "vsmpc_get_pivot_counts", "vsmpc_get_references", "vsmpc_get_hessian", "vsmpc_get_constraint_matrix",
    "vsmpc_debug_set_counters", "vsmpc_debug_phase_clocks", "vsmpc_microbench_fp64", "vsmpc_set_fallback",
    "vsmpc_configure_strided", "vsmpc_set_state_strided",
    "vsmpc_create_multi", "vsmpc_multi_destroy", "vsmpc_multi_last_error", "vsmpc_multi_n_shards", "vsmpc_multi_n_instances",
    "vsmpc_multi_shard", "vsmpc_multi_configure", "vsmpc_multi_set_instance_params", "vsmpc_multi_set_state", "vsmpc_multi_solve",
    "vsmpc_multi_solve_async", "vsmpc_multi_wait", "vsmpc_multi_get_output", "vsmpc_multi_set_full_solution",
    "vsmpc_multi_get_full_solution",
    "vsmpc_set_instance_params", "vsmpc_set_joint_limits", "vsmpc_set_warm_start", "vsmpc_set_kin_model", "vsmpc_kin_state_doubles", "vsmpc_configure_kinematics", "vsmpc_set_state_kinematics", "vsmpc_get_kinematics_pack", "vsmpc_debug_set_working_set", "vsmpc_rollout_init", "vsmpc_rollout_run", "vsmpc_rollout_get_state", "vsmpc_rollout_set_jet_nn", "vsmpc_jet_nn_eval", "vsmpc_rollout_get_pack",
]

_lib = None


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    H = C.c_void_p
    lib.vsmpc_create.argtypes = [C.POINTER(VsmpcConfig), C.c_int, C.c_int, C.POINTER(H)]
    lib.vsmpc_destroy.argtypes = [H]
    lib.vsmpc_last_error.argtypes = [H]
    lib.vsmpc_last_error.restype = C.c_char_p
    lib.vsmpc_set_stream.argtypes = [H, C.c_void_p]
    for f in ("vsmpc_n_var", "vsmpc_n_constraints", "vsmpc_n_instances"):
        getattr(lib, f).argtypes = [H]
    lib.vsmpc_configure.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vsmpc_set_state.argtypes = [H, C.c_void_p]
    lib.vsmpc_linearise.argtypes = [H, C.c_void_p]
    lib.vsmpc_solve_qp.argtypes = [H]
    lib.vsmpc_set_state_device.argtypes = [H, C.c_void_p]
    for f in ("vsmpc_solve", "vsmpc_solve_async", "vsmpc_wait"):
        getattr(lib, f).argtypes = [H]
    lib.vsmpc_get_output.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.vsmpc_get_output_async.argtypes = [H, C.c_void_p, C.c_void_p, c_int_p]
    lib.vsmpc_wait_output.argtypes = [H, C.c_int]
    lib.vsmpc_get_output_device.argtypes = [H, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p)]
    lib.vsmpc_set_full_solution.argtypes = [H, C.c_int]
    lib.vsmpc_get_full_solution.argtypes = [H, C.c_void_p]
    lib.vsmpc_get_dynamics.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vsmpc_get_qp_vectors.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vsmpc_get_counts.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.vsmpc_get_pivot_counts.argtypes = [H, C.c_void_p]
    lib.vsmpc_get_references.argtypes = [H, C.c_void_p]
    lib.vsmpc_get_hessian.argtypes = [H, C.c_int, C.c_void_p]
    lib.vsmpc_get_constraint_matrix.argtypes = [H, C.c_int, C.c_void_p]
    lib.vsmpc_debug_set_counters.argtypes = [H, C.c_int, C.c_int]
    lib.vsmpc_set_fallback.argtypes = [H, C.c_int]
    lib.vsmpc_set_warm_start.argtypes = [H, C.c_int]
    lib.vsmpc_set_kin_model.argtypes = [H, C.c_void_p]
    lib.vsmpc_kin_state_doubles.argtypes = [H]
    lib.vsmpc_configure_kinematics.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.vsmpc_set_state_kinematics.argtypes = [H, C.c_void_p]
    lib.vsmpc_get_kinematics_pack.argtypes = [H, C.c_void_p]
    lib.vsmpc_debug_set_working_set.argtypes = [H, C.c_void_p]
    lib.vsmpc_debug_phase_clocks.argtypes = [C.c_void_p, C.c_int]
    lib.vsmpc_microbench_fp64.argtypes = [C.c_int, C.c_int, c_double_p]
    lib.vsmpc_set_instance_params.argtypes = [H, C.c_void_p]
    lib.vsmpc_set_joint_limits.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.vsmpc_rollout_init.argtypes = [H, C.POINTER(VsmpcPlantModel), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vsmpc_rollout_run.argtypes = [H, C.c_int, C.c_int, C.c_void_p, C.c_int]
    lib.vsmpc_rollout_get_state.argtypes = [H, C.c_void_p]
    lib.vsmpc_rollout_set_jet_nn.argtypes = [H] + [C.c_void_p] * 8
    lib.vsmpc_jet_nn_eval.argtypes = [H, C.c_int, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vsmpc_rollout_get_pack.argtypes = [H, C.c_void_p]
    lib.vsmpc_configure_strided.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
    lib.vsmpc_set_state_strided.argtypes = [H, C.c_void_p, C.c_size_t]
    lib.vsmpc_create_multi.argtypes = [C.POINTER(VsmpcConfig), C.c_int, C.c_int, C.c_void_p, C.POINTER(H)]
    lib.vsmpc_multi_last_error.argtypes = [H]
    lib.vsmpc_multi_last_error.restype = C.c_char_p
    for f in ("vsmpc_multi_destroy", "vsmpc_multi_n_shards", "vsmpc_multi_n_instances", "vsmpc_multi_solve",
              "vsmpc_multi_solve_async", "vsmpc_multi_wait"):
        getattr(lib, f).argtypes = [H]
    lib.vsmpc_multi_shard.argtypes = [H, C.c_int, c_int_p, c_int_p, C.POINTER(H)]
    lib.vsmpc_multi_configure.argtypes = [H, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.vsmpc_multi_set_instance_params.argtypes = [H, C.c_void_p]
    lib.vsmpc_multi_set_state.argtypes = [H, C.c_void_p]
    lib.vsmpc_multi_get_output.argtypes = [H, C.c_void_p, C.c_void_p]
    lib.vsmpc_multi_set_full_solution.argtypes = [H, C.c_int]
    lib.vsmpc_multi_get_full_solution.argtypes = [H, C.c_void_p]
    for f in EXPORTS:
        if f not in ("vsmpc_last_error", "vsmpc_multi_last_error"):
            getattr(lib, f).restype = C.c_int
    _lib = lib
    return lib
