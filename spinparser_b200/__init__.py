"""spinparser_b200: B200-native pf-FRG flow-equation core behind SpinParser's FrgCore plugin surface.

The package is a thin host mirror of the reference interface (``FrgCoreFactory.newFrgCore`` -> ``FrgCore`` with
``computeStep`` / ``finalizeStep`` / ``flowingFunctional`` / ``flow``) over the C ABI of ``include/pffrg.h``.
All compute happens in hand-written CUDA kernels for sm_100a inside ``libpffrg.so``; there is no CPU fallback.
"""
from .pfd import read_pfd, write_pfd  # noqa: F401

__all__ = ["read_pfd", "write_pfd", "FrgCore", "FrgCoreFactory", "EffectiveAction", "ProblemTables", "PffrgError"]


def __getattr__(name):
    # the CUDA library is loaded on first use of the core classes so that pure-host helpers (pfd) stay importable
    if name in ("FrgCore", "FrgCoreFactory", "EffectiveAction", "ProblemTables", "PffrgError"):
        from . import frgcore
        return getattr(frgcore, name)
    raise AttributeError(name)
