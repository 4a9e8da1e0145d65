"""cha1_mcmc_b200 -- B200-native LTE likelihood engine for the emcee walker log-probability of
KahaanGandhi/Cha1-MCMC's inference.py.  Host API mirrors the reference (config dict, MolCat,
SpectralFitMCMC); the arithmetic runs in hand-written sm_100a CUDA behind a C-ABI (include/chalte.h)."""
from .catalog import MolCat, resolve_q_mode, find_catalog          # noqa: F401
from .engine import LTEEngine, ModelSpec, EngineError, load_library  # noqa: F401

__all__ = ["MolCat", "LTEEngine", "ModelSpec", "EngineError", "load_library", "resolve_q_mode", "find_catalog"]
__version__ = "0.1.0"
