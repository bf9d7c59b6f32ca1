"""CPU: the C-ABI library loads and exports every symbol include/bout_b200.h declares; without a CUDA
device the compute entry points fail loudly (no CPU fallback exists)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "bout_b200.h")
LIB = os.path.join(ROOT, "bout-dev_b200", "csrc", "libbout_b200.so")


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    return sorted(set(re.findall(r"\b(bout_b200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_hot_path():
    syms = declared_symbols()
    for need in ("bout_b200_stencil", "bout_b200_bracket_arakawa", "bout_b200_delp2_fft",
                 "bout_b200_laplace_cyclic_solve", "bout_b200_rk_stage", "bout_b200_halo_self"):
        assert need in syms


def test_library_exports_every_declared_symbol():
    if not os.path.exists(LIB):
        pytest.fail(f"{LIB} missing: run python -c 'import __graft_entry__ as g; g.build()'")
    lib = C.CDLL(LIB)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.bout_b200_version() >= 100


def test_product_never_touches_the_oracle():
    """The product tree must not import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "bout-dev_b200")
    offenders = []
    for dirpath, _, files in os.walk(pkg):
        if os.sep + "build" in dirpath:
            continue
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".cxx", ".hxx", ".h", ".cpp", "Makefile")):
                text = open(os.path.join(dirpath, fn), errors="replace").read()
                # nor the host-memory test double of tests/fake_device (CPU tests of the host logic only)
                if re.search(r"oracle[/.]|bout_oracle|libboutref|liboracle_port|fake_device|libfake_b200|LD_PRELOAD", text):
                    offenders.append(os.path.join(dirpath, fn))
    assert not offenders, offenders


def test_compute_call_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    lib = C.CDLL(LIB)
    lib.bout_b200_last_error.restype = C.c_char_p

    class G(C.Structure):
        _fields_ = [(n, C.c_int) for n in ("nx", "ny", "nz", "xstart", "xend", "ystart", "yend")]

    g = G(8, 4, 8, 1, 6, 1, 2)
    buf = (C.c_double * (8 * 4 * 8))()
    rc = lib.bout_b200_unary(C.byref(g), 10, buf, C.c_double(0.0), buf, None)
    assert rc == -2, rc   # BOUT_B200_ERR_CUDA
    assert b"" != lib.bout_b200_last_error()


def test_header_is_plain_c():
    """The drop-in boundary is a C ABI: include/bout_b200.h must compile as C99 (no C++/torch types)."""
    import shutil
    import subprocess
    import tempfile
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    with tempfile.NamedTemporaryFile("w", suffix=".c", delete=False) as f:
        f.write('#include "bout_b200.h"\nint main(void) { bout_b200_elmpb_rhs_args a; bout_b200_laplace l; (void)a; (void)l; return 0; }\n')
        path = f.name
    try:
        p = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                            "-fsyntax-only", path], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr
    finally:
        os.unlink(path)
