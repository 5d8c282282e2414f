"""The C-ABI library loads and exports every symbol include/mrcal_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_loads_and_exports_header_symbols():
    from mrcal_b200 import _capi
    header = open(os.path.join(ROOT, "include", "mrcal_b200.h")).read()
    # function declarations: a name followed by '(' at the start of a declaration line
    declared = set(re.findall(r"^\s*(?:[A-Za-z_][\w\s\*]*?[\s\*])?(_?mrcal_\w+)\s*\(", header, re.M))
    declared = {d for d in declared if not d.endswith("_t")}
    assert len(declared) > 50
    missing = [s for s in declared if not hasattr(_capi.lib, s)]
    assert not missing, f"declared in the header but not exported: {missing}"
    # and the binding's own list agrees with the header
    assert set(_capi.EXPORTED_SYMBOLS) == declared, set(_capi.EXPORTED_SYMBOLS) ^ declared


def test_header_is_plain_c(tmp_path):
    """include/mrcal_b200.h is the drop-in boundary: it has to compile as C (gnu11, the reference's dialect: its types use
    anonymous structs, basic-geometry.h:19-60) and as C++"""
    import shutil
    import subprocess
    src = tmp_path / "h.c"
    src.write_text('#include "mrcal_b200.h"\nint main(void) { return (int)sizeof(mrcal_lensmodel_t) - (int)sizeof(mrcal_lensmodel_t); }\n')
    inc = os.path.join(ROOT, "include")
    for cc, std in (("gcc", "-std=gnu11"), ("g++", "-std=c++17")):
        if shutil.which(cc) is None:
            pytest.skip(cc + " not found")
        r = subprocess.run([cc, std, "-Wall", "-Wextra", "-Werror", "-fsyntax-only", "-x", "c" if cc == "gcc" else "c++", "-I", inc, str(src)],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr


def test_struct_sizes_match_reference_abi():
    from mrcal_b200 import _capi
    assert ctypes.sizeof(_capi.Lensmodel) == 16      # types.h:122-136
    assert ctypes.sizeof(_capi.Selections) == 1      # types.h:283-307
    assert ctypes.sizeof(_capi.Stats) == 16          # types.h:320-343
    assert ctypes.sizeof(_capi.Sparse) == 88         # public cholmod_sparse layout


def test_no_gpu_fails_loudly():
    """Without a CUDA device the hot path must raise, never fall back."""
    import mrcal_b200
    from mrcal_b200 import synthetic
    if mrcal_b200.device_count() > 0:
        pytest.skip("a GPU is present")
    inp, _ = synthetic.make_problem(Ncameras=1, Nframes=3, W=4, H=4)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        mrcal_b200.optimizer_callback(**inp)
    with pytest.raises(RuntimeError, match="no usable CUDA device"):
        mrcal_b200.optimize(**inp)


def test_product_does_not_import_oracle():
    """The product must never route through the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "mrcal_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h", ".cuh")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "libmrcal_ref" not in txt, f
