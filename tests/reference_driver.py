"""Drive the reference's OWN VariableSamplingMPC (oracle/_ref/libvsmpc_reference.so, compiled from /root/reference by
oracle/build_ref.py against stand-in headers) on the same getter-level states the oracle takes.  Test infrastructure."""
import ctypes

import numpy as np

from helpers import load_trajectories
from oracle import build_ref
from oracle import vsmpc_oracle as O

_QP_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double),
                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double))
_DP = ctypes.POINTER(ctypes.c_double)
_lib = None
_glue = False
_keep = []


def _p(a):
    return a.ctypes.data_as(_DP)


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def lib(glue=False):
    """glue=True: oracle/_ref/libvsmpc_reference_glue.so — the same library compiled together with the product's reference-side
    binding (include/vsmpc_reference_glue.hpp); it links libvsmpc.so.  One library per process (both define the same symbols)."""
    global _lib, _glue
    if _lib is not None and glue and not _glue:
        raise RuntimeError("reference_driver: the plain reference library is already loaded in this process")
    if _lib is None:
        path = build_ref.build_glue() if glue else build_ref.build_mpc()
        if not path:
            return None
        _glue = bool(glue)
        L = ctypes.CDLL(path)
        if glue:
            for f in ("ref_glue_configure", "ref_glue_update", "ref_glue_solve", "ref_glue_status", "ref_glue_nvar", "ref_glue_ncon"):
                getattr(L, f).argtypes = [ctypes.c_void_p]
            L.ref_glue_get_output.argtypes = [ctypes.c_void_p] + [_DP] * 7
            L.ref_glue_clear_published.argtypes = [ctypes.c_void_p, ctypes.c_double]
            L.ref_glue_fill_pack.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int), _DP]
        L.ref_mpc_create.restype = ctypes.c_void_p
        L.ref_mpc_create.argtypes = [ctypes.c_char_p, ctypes.c_char_p]
        for f in ("ref_mpc_destroy", "ref_mpc_configure", "ref_mpc_update", "ref_mpc_solve", "ref_mpc_nvar", "ref_mpc_ncon",
                  "ref_mpc_status"):
            getattr(L, f).argtypes = [ctypes.c_void_p]
        L.ref_mpc_param_numbers.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, _DP, ctypes.c_int]
        L.ref_mpc_param_strings.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p]
        L.ref_mpc_mat_variable.argtypes = [ctypes.c_char_p, ctypes.c_char_p, _DP, ctypes.c_int, ctypes.c_int]
        L.ref_mpc_set_robot.argtypes = [ctypes.c_void_p] + [_DP] * 15
        L.ref_mpc_set_input.argtypes = [ctypes.c_void_p, ctypes.c_char_p, _DP, ctypes.c_int]
        L.ref_mpc_set_qp_solver.argtypes = [_QP_FN]
        L.ref_mpc_get_qp.argtypes = [ctypes.c_void_p] + [_DP] * 5
        L.ref_mpc_get_output.argtypes = [ctypes.c_void_p] + [_DP] * 7

        def solve(n, m, P, q, A, l, u, z):
            # the data arrive column-major, exactly as IMPCProblem::solve handed them to OsqpEigen
            Pm = np.ctypeslib.as_array(P, (n * n,)).reshape(n, n).T
            Am = np.ctypeslib.as_array(A, (n * m,)).reshape(n, m).T
            try:
                zz, _, _ = O.solve_qp_exact(Pm, np.ctypeslib.as_array(q, (n,)), Am, np.ctypeslib.as_array(l, (m,)),
                                            np.ctypeslib.as_array(u, (m,)))
            except Exception:
                return -3
            np.ctypeslib.as_array(z, (n,))[:] = zz
            return 1                                              # OsqpEigen::Status::Solved

        cb = _QP_FN(solve)
        _keep.append(cb)
        L.ref_mpc_set_qp_solver(cb)
        _lib = L
    return _lib


def register_trajectories(L, traj):
    """The two .mat files of the reference as the matio stand-in's tables (variables are (dim, samples), column-major)."""
    for group, fname in (("TRAJECTORY_MANAGER", "alphaGravity.mat"), ("POSITION_TRAJECTORY", "minimumJerkTrajectory.mat")):
        t = traj[group]
        fps = _c([float(t["fps"])])
        L.ref_mpc_mat_variable(fname.encode(), b"fps", _p(fps), 1, 1)
        for k, a in t["arrays"].items():
            a = _c(np.asarray(a, float).T)                       # C-order of the transpose = column-major of (dim, N)
            L.ref_mpc_mat_variable(fname.encode(), k.encode(), _p(a), a.shape[1], a.shape[0])


class ReferenceInstance:
    """Same call sequence as tests/oracle_driver.OracleInstance, on the compiled reference."""

    def __init__(self, nominal_state, i, params=None, trajectories=None, glue=False):
        self.L = lib(glue)
        self.i = i
        p = dict(O.default_params())
        p.update(params or {})
        register_trajectories(self.L, trajectories or load_trajectories())
        self.nJ = nominal_state["joint_pos"].shape[1]
        self.h = self.L.ref_mpc_create(",".join(O.AXES_LIST[:self.nJ]).encode(), ",".join(O.JETS_LIST).encode())
        for k, v in p.items():
            if isinstance(v, str):
                self.L.ref_mpc_param_strings(self.h, b"", k.encode(), v.encode())
            elif isinstance(v, (list, tuple)) and v and isinstance(v[0], str):
                self.L.ref_mpc_param_strings(self.h, b"", k.encode(), ",".join(v).encode())
            else:
                a = _c(np.atleast_1d(np.asarray(v, float)))
                self.L.ref_mpc_param_numbers(self.h, b"", k.encode(), _p(a), a.size)
        self.L.ref_mpc_param_strings(self.h, b"TRAJECTORY_MANAGER", b"trajectoryFile", b"alphaGravity.mat")
        self.L.ref_mpc_param_strings(self.h, b"POSITION_TRAJECTORY", b"trajectoryFile", b"minimumJerkTrajectory.mat")
        self._set(nominal_state)
        assert self.L.ref_mpc_configure(self.h) == 0, "reference configure() returned false"
        self.n_var, self.n_con = self.L.ref_mpc_nvar(self.h), self.L.ref_mpc_ncon(self.h)

    def _set(self, s):
        i = self.i
        a = [_c(s[k][i]) for k in ("wRb", "base_pos", "omega_world", "M_b", "p_com", "momentum_body", "A_mom_body", "jet_axes",
                                   "jet_arms", "J_rel_body", "J_jet_lin", "J_com", "thrust", "joint_pos", "gravity")]
        self.L.ref_mpc_set_robot(self.h, *[_p(x) for x in a])
        for name, key in (("ThrottleMPC", "throttle_prev"), ("ThrustDesMPC", "thrust_des"), ("ThrustDotDesMPC", "thrust_dot_des"),
                          ("EstimatedThrustDot", "thrust_dot_est"), ("OutputQPJointsPosition", "q_cmd")):
            v = _c(s[key][i])
            assert self.L.ref_mpc_set_input(self.h, name.encode(), _p(v), v.size) == 0

    def update(self, state):
        self._set(state)
        assert self.L.ref_mpc_update(self.h) == 0

    def qp(self):
        n, m = self.n_var, self.n_con
        P, q, A, l, u = np.zeros((n, n)), np.zeros(n), np.zeros((m, n)), np.zeros(m), np.zeros(m)
        self.L.ref_mpc_get_qp(self.h, _p(P), _p(q), _p(A), _p(l), _p(u))
        return P, q, A, l, u

    def solve(self):
        assert self.L.ref_mpc_solve(self.h) == 0
        return self.output()["solution"]

    def output(self):
        o = dict(joints=np.zeros(self.nJ), throttle=np.zeros(4), thrust=np.zeros(4), thrust_dot=np.zeros(4),
                 final=np.zeros(12), solution=np.zeros(self.n_var), qp_input=np.zeros(13))
        self.L.ref_mpc_get_output(self.h, *[_p(o[k]) for k in ("joints", "throttle", "thrust", "thrust_dot", "final", "solution",
                                                               "qp_input")])
        o["status"] = self.L.ref_mpc_status(self.h)
        return o

    # ---- the product's reference-side binding on the same QPInput object (glue library only) ------------------------------
    def glue_configure(self):
        return self.L.ref_glue_configure(self.h) == 0

    def glue_update(self, state=None):
        if state is not None:
            self._set(state)
        return self.L.ref_glue_update(self.h) == 0

    def glue_solve(self):
        return self.L.ref_glue_solve(self.h) == 0

    def glue_output(self):
        n = self.L.ref_glue_nvar(self.h)
        o = dict(joints=np.zeros(self.nJ), throttle=np.zeros(4), thrust=np.zeros(4), thrust_dot=np.zeros(4),
                 final=np.zeros(12), solution=np.zeros(n), qp_input=np.zeros(13))
        rc = self.L.ref_glue_get_output(self.h, *[_p(o[k]) for k in ("joints", "throttle", "thrust", "thrust_dot", "final",
                                                                     "solution", "qp_input")])
        o["ok"] = rc == 0
        o["status"] = self.L.ref_glue_status(self.h)
        return o

    def glue_fill_pack(self, sel):
        pk = np.zeros(359)
        s = (ctypes.c_int * 8)(*[int(j) for j in sel])
        assert self.L.ref_glue_fill_pack(self.h, s, _p(pk)) == 0
        return pk

    def clear_published(self, v=-7.0):
        self.L.ref_glue_clear_published(self.h, float(v))

    def set_state(self, state):
        self._set(state)

    def close(self):
        if self.h:
            self.L.ref_mpc_destroy(self.h)
            self.h = None
