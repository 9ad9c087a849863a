"""collision-detect-gjk-epa_b200: B200-native (sm_100a) batched GJK + EPA narrow phase.

Drop-in for the ``GJKEPA`` subroutine of xiejihong0306/collision-detect-GJK-EPA
(src/GCLIB_GJKEPA.f90:39-52) behind a C ABI (include/gjkepa.h).  See DESIGN.md.
"""
from .api import GJKEPA, Engine, guard_check, GjkEpaError, EXIT_NAMES, MODE_GJK, MODE_EPA, MODE_FULL  # noqa: F401
from . import workloads  # noqa: F401

__all__ = ["GJKEPA", "Engine", "guard_check", "GjkEpaError", "EXIT_NAMES", "MODE_GJK", "MODE_EPA", "MODE_FULL", "workloads"]
