"""phylocsf_b200 — B200-native (sm_100a) implementation of PhyloCSF's scoring hot path:
P(t) construction and Felsenstein pruning under the 64-state ECMs, behind a C ABI
(include/phylocsf_b200.h). See DESIGN.md."""
from ._native import LIB_PATH, NativeLibraryMissing, SYMBOLS  # noqa: F401
from .api import Context, PcsfError  # noqa: F401

__version__ = "0.1"
