"""Loads the UNMODIFIED reference (build container only; /root/reference does not exist on
the GPU box).  Test infrastructure: used by oracle/make_golden.py and by the CPU-side
"oracle vs live reference" tests, which skip when the reference is absent.

The reference does not import as-is on this image: utils.py:2 imports an
unused name from tkinter (absent) and bbme.py:143... use np.infty (removed in NumPy 2).
Both are shimmed here without touching the reference tree.
"""
import importlib
import os
import sys
import types

import numpy as np

REFERENCE_DIR = "/root/reference/global_motion_estimation"


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "bbme.py"))


def load():
    """Returns (utils, bbme, motion) of the reference, imported under private names so they
    never shadow (or get shadowed by) the drop-in modules of the same bare names."""
    if not available():
        raise RuntimeError("reference not present at " + REFERENCE_DIR)
    t = types.ModuleType("tkinter")
    t.image_names = lambda *a, **k: ()
    sys.modules.setdefault("tkinter", t)                 # utils.py:2
    if not hasattr(np, "infty"):
        np.infty = np.inf                                # bbme.py:143, 226, 264, 304, 383, 495, 516
    saved = {k: sys.modules.pop(k, None) for k in ("utils", "bbme", "motion")}
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", SyntaxWarning)   # '\g' in a docstring, bbme.py:654
            mods = tuple(importlib.import_module(n) for n in ("utils", "bbme", "motion"))
    finally:
        sys.path.remove(REFERENCE_DIR)
        for k in ("utils", "bbme", "motion"):
            sys.modules.pop(k, None)
            if saved[k] is not None:
                sys.modules[k] = saved[k]
    return mods


def with_search_override(motion_mod, bbme_mod, procedure, window):
    """configs 4/5 override the search on the bs=16 levels only; the dense bs=2
    first estimate stays diamond and the cost stays MSE.  Returns an undo callable."""
    original = motion_mod.get_motion_field

    def patched(previous, current, block_size, searching_procedure, **kw):
        sp = procedure if block_size == motion_mod.BBME_BLOCK_SIZE else searching_procedure
        return bbme_mod.get_motion_field(previous, current, block_size=block_size,
                                         searching_procedure=sp, search_window=window)

    motion_mod.get_motion_field = patched
    return lambda: setattr(motion_mod, "get_motion_field", original)
