"""harmonypy_b200 -- Blackwell-native engine for the Harmony inner loop.

Public surface mirrors slowkow/harmonypy (harmonypy/__init__.py:1-4).
"""
__version__ = "0.1.0"


def __getattr__(name):
    # Lazy so that ``harmonypy_b200.synthetic`` can be used without touching CUDA.
    if name in ("run_harmony", "Harmony"):
        from . import harmony
        return getattr(harmony, name)
    if name == "pinned_empty":            # page-locked host arrays: inputs / results that move with one DMA
        from ._cabi import pinned_empty
        return pinned_empty
    if name == "compute_lisi":
        from .lisi import compute_lisi
        return compute_lisi
    raise AttributeError(name)
