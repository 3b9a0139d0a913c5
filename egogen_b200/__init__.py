"""egogen_b200 - B200-native implementation of EgoGen's crowd_ppo hot path.

Host side mirrors the reference's Python operator surface (SMPLXParser, calc_sdf, CrowdEnv,
GAMMAPPOPolicy, main_ppo); the arithmetic runs in hand-written sm_100a CUDA kernels behind the
C ABI declared in include/egogen_b200.h (egogen_b200/libegogen_b200.so, built by egogen_b200.build).
"""
from .sdf import calc_sdf, penetration_count, ego_depth  # noqa: F401
from .smplx_parser import SMPLXParser, LbsModel, get_lbs_model  # noqa: F401

__version__ = "0.1.0"
