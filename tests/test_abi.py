"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every symbol that
include/cvgs_b200.h declares, the ctypes mirror has the C layout, and argument validation that happens
before any CUDA call behaves as documented."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

from cvgpuspeedup_b200 import _abi
from tests import util

HEADER = os.path.join(util.ROOT, "include", "cvgs_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cvgs_b200_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = _abi.load()
    declared = _declared_functions()
    assert len(declared) >= 10
    assert sorted(_abi.SYMBOLS) == declared, "ctypes table and header disagree"
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.cvgs_b200_version() == 105


def test_struct_layouts_match_c(tmp_path):
    prog = tmp_path / "layout.c"
    prog.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "cvgs_b200.h"\n'
                    'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(cvgs_crop_t), sizeof(cvgs_op_t),'
                    'sizeof(cvgs_pipeline_t), offsetof(cvgs_pipeline_t, ops), offsetof(cvgs_pipeline_t, out),'
                    'offsetof(cvgs_pipeline_t, out_plane_stride), sizeof(cvgs_rect_t));return 0;}\n')
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.dirname(HEADER), str(prog), "-o", str(exe)])
    got = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    want = [C.sizeof(_abi.Crop), C.sizeof(_abi.Op), C.sizeof(_abi.Pipeline), _abi.Pipeline.ops.offset,
            _abi.Pipeline.out.offset, _abi.Pipeline.out_plane_stride.offset, C.sizeof(_abi.Rect)]
    assert got == want


def test_header_is_plain_c():
    """The boundary must be consumable from C (cgo / JNI / ctypes style bindings)."""
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-fsyntax-only", "-x", "c", HEADER])


def test_validation_without_gpu():
    lib = _abi.load()
    assert lib.cvgs_b200_preproc_launch(None, 1, 1, None, None) == 1
    assert b"pipeline" in lib.cvgs_b200_last_error()
    p = util.make_pipeline((0, 10), [])
    assert lib.cvgs_b200_preproc_launch(None, 1, 1, C.byref(p), None) == 1
    p = util.make_pipeline((64, 128), [("mul", (1, 1, 1))])
    p.ops[0].kind = 99
    p.out = 16
    crops = (_abi.Crop * 1)()
    assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) == 1
    assert b"op kind" in lib.cvgs_b200_last_error()
    p = util.make_pipeline((64, 128), [("reorder", (0, 0, 1))], out_ptr=16)
    assert lib.cvgs_b200_preproc_launch(crops, 1, 1, C.byref(p), None) == 1
    assert b"permutation" in lib.cvgs_b200_last_error()
    assert lib.cvgs_b200_ct_update(None, None, None, None) == 1
    assert lib.cvgs_b200_ct_data(None) is None
    assert lib.cvgs_b200_ct_destroy(None) == 0


def test_python_mirror_builds_reference_style_chains():
    import cvgpuspeedup_b200 as cvgs
    with pytest.raises(cvgs.CvgsError):
        cvgs.executeOperations(None, cvgs.multiply(1.0))
    p = cvgs.build_pipeline((64, 128), [cvgs.cvtColor(), cvgs.convertTo(0.5, 1.0), cvgs.subtract((1, 2, 3))])
    assert p.n_ops == 4 and [p.ops[i].kind for i in range(4)] == [5, 1, 4, 2]
    assert list(p.ops[0].perm)[:3] == [2, 1, 0]
    g = cvgs.GpuMat(4096, 100, 50, 512)
    r = g.roi(10, 5, 20, 30)
    assert (r.data, r.cols, r.rows, r.step) == (4096 + 5 * 512 + 30, 20, 30, 512)
    with pytest.raises(ValueError):
        g.roi(90, 0, 20, 10)


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    code = ("import cvgpuspeedup_b200._abi as a; a.LIB_PATH='/nonexistent/lib.so'; a._lib=None\n"
            "try:\n a.load()\nexcept RuntimeError as e:\n print('LOUD', 'no CPU fallback' in str(e))\n")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=util.ROOT).decode()
    assert "LOUD True" in out
