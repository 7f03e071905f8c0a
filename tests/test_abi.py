"""CPU-side checks of the drop-in boundary: the shared library loads, exports
every symbol include/s2mv.h declares plus the reference's own C++ symbols
(SURVEY §8b), and refuses to compute without a GPU (no fallback path)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "s2mv.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s2mv_[a-z0-9_]+)\s*\(", text)))


def test_header_and_python_symbol_lists_agree(s2mv):
    assert header_symbols() == sorted(s2mv.EXPORTED_SYMBOLS)


def test_library_exports_every_declared_symbol(s2mv):
    L = ctypes.CDLL(s2mv.LIB_PATH)
    for name in header_symbols():
        assert hasattr(L, name), name


def test_library_exports_reference_cxx_symbols(s2mv):
    # exact Itanium-mangled names of the reference's entry points (d_io.h:32-40 and the stage headers)
    L = ctypes.CDLL(s2mv.LIB_PATH)
    for name in s2mv.COMPAT_SYMBOLS:
        assert hasattr(L, name), name


def test_compat_header_lists_same_functions(s2mv):
    text = open(os.path.join(ROOT, "include", "s2mv_compat.h")).read()
    for mangled in s2mv.COMPAT_SYMBOLS:
        assert mangled in text, mangled


def test_params_struct_matches_header(s2mv):
    text = open(os.path.join(ROOT, "include", "s2mv.h")).read()
    body = re.search(r"typedef struct \{(.*?)\} s2mv_params;", text, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            names += [n.strip() for n in decl.split(None, 1)[1].split(",")]
    assert names == [f[0] for f in s2mv.Params._fields_]


def test_no_cpu_fallback(s2mv):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present: the no-device path cannot be exercised")
    with pytest.raises(s2mv.S2mvError, match="no usable CUDA device"):
        s2mv.Pipeline(0)


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "stereo-to-multiview-cuda_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".inl", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.lower(), os.path.join(dirpath, f)


def test_headers_compile_and_struct_sizes_match_ctypes(s2mv, tmp_path):
    """include/s2mv.h is plain C (any FFI can bind it), include/s2mv_compat.h is the reference's C++ surface; the
    ctypes mirrors of the two structs have the C compiler's size and field offsets."""
    import subprocess
    from s2mv_b200_pkg import rowband
    inc = os.path.join(ROOT, "include")
    src = tmp_path / "t.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "s2mv.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(s2mv_params), offsetof(s2mv_params, thresh_h), '
                   'offsetof(s2mv_params, mask_blur_sigma), sizeof(s2mv_band_ipc), offsetof(s2mv_band_ipc, frame_y0), '
                   'offsetof(s2mv_band_ipc, device)); return 0; }\n')
    exe = tmp_path / "t"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", inc, str(src), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    P, B = s2mv.Params, rowband.BandIpc
    assert got == [ctypes.sizeof(P), P.thresh_h.offset, P.mask_blur_sigma.offset, ctypes.sizeof(B), B.frame_y0.offset,
                   B.device.offset]
    cxx = tmp_path / "t.cpp"
    cxx.write_text('#include "s2mv_compat.h"\n#include "s2mv.h"\nint main() { return 0; }\n')
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Werror", "-fsyntax-only", "-I", inc, str(cxx)])
