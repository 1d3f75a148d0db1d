"""B200-native batched variable-sampling ("multi-rate") MPC for iRonCub — host-side package.

Public surface: BatchedVSMPC (B instances on one GPU), VariableSamplingMPC / QPInput (single-instance
mirror of the reference's Python bindings), synthetic state generator, config helpers.
"""
from .config import default_params, read_xml_config, load_trajectories_npz, hover_trajectories  # noqa: F401
from .pack import PACK_FIELDS, PACK_OFFSETS, PACK_DOUBLES, build_pack  # noqa: F401


def __getattr__(name):
    # compute classes are imported lazily so that pack/config/synthetic stay importable on machines
    # where the CUDA library has not been built yet (they raise loudly when actually used)
    if name in ("BatchedVSMPC", "VsmpcError"):
        from . import batched
        return getattr(batched, name)
    if name in ("VariableSamplingMPC", "QPInput", "RobotState"):
        from . import mpc
        return getattr(mpc, name)
    raise AttributeError(name)
