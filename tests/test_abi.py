"""CPU: the C-ABI library loads and exports every symbol include/gme_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import gme_native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "gme_b200.h")).read()
    return set(re.findall(r"GME_API[^;(]*?\b(gme_\w+)\s*\(", text))


def test_header_symbols_exported_and_bound():
    names = declared_symbols()
    assert len(names) >= 14
    lib = ctypes.CDLL(N.LIB_PATH)
    for name in names:
        assert hasattr(lib, name), f"{name} declared in gme_b200.h but not exported"
        assert name in N.SYMBOLS, f"{name} has no ctypes prototype in gme_native.SYMBOLS"
    assert set(N.SYMBOLS) == names


def test_version_and_error_strings():
    assert N.ABI_VERSION == 2
    assert N.lib.gme_error_string(0) == b"ok"
    for code in (-1, -2, -3, -4, -5):
        assert N.lib.gme_error_string(code) not in (b"ok", b"unknown error")


def test_argument_validation_needs_no_gpu():
    # null pointers / bad enums are rejected before any CUDA call
    assert N.lib.gme_bbme_motion_field(None, 0, None, 0, 1, 16, 16, 16, 4, 2, 0, 0, None, None) == N.GME_ERR_INVALID_ARGUMENT
    assert N.lib.gme_pyr_down(None, 0, 0, None, 0, 0, 1, 8, 8, None) == N.GME_ERR_INVALID_ARGUMENT
    assert N.lib.gme_pipeline_workspace_bytes(4, 1080, 1920) > 4 * (540 * 960 + 270 * 480) * 2
    assert N.lib.gme_pipeline_workspace_bytes(0, 10, 10) == 0


def test_product_path_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under the package may import or load it."""
    pkg = os.path.join(ROOT, "global-motion-estimation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                assert "libgme_oracle" not in text and "oracle/_" not in text, f
                if f.endswith(".py"):
                    assert "gme_oracle" not in text and "ref_shim" not in text, f
