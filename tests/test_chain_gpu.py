"""The fused kernel against a multi-kernel chain, the way the reference tests itself (fused call vs the sequence of
cv::cuda calls on the same inputs, reference tests/resize/test_resize_x_split.cu:72-84, README.md:91-97).

OpenCV-CUDA is not installable here, so the chain is a restatement (oracle/chain_kernels.cu, test infrastructure):
resize on 8-bit -> convertTo(alpha) -> [cvtColor] -> subtract -> divide -> split, 5-6 launches per crop.  The
product reproduces it bit for bit in (fp_contract = SEPARATE, interp_mode = ROUND_U8) and the oracle agrees."""
import numpy as np
import pytest

from cvgpuspeedup_b200 import _abi
from tests import gpu_util, util

pytestmark = pytest.mark.gpu

MUL, SUB, DIV = (0.3, 0.3, 0.3), (1.0, 4.0, 3.2), (3.2, 0.6, 11.8)


@pytest.mark.parametrize("swap", [False, True])
@pytest.mark.parametrize("variant", [1, 2])
def test_fused_equals_multi_kernel_chain(swap, variant):
    if gpu_util.chain_lib() is None:
        pytest.skip("oracle/libchain.so not built")
    w = util.workload_c2(seed=21, n=50, pitch=6144)
    ops = ([("reorder", (2, 1, 0))] if swap else []) + [("mul", MUL), ("sub", SUB), ("div", DIV)]
    chain, launches = gpu_util.run_chain(w.image, w.rects, w.dsize, swap, MUL, SUB, DIV)
    assert launches == len(w.rects) * (6 if swap else 5)
    kw = dict(fp_contract=_abi.FP_SEPARATE, interp_mode=_abi.INTERP_ROUND_U8)
    fused = gpu_util.run_cvgs(w.image, w.rects, w.dsize, ops, variant=variant, **kw)
    util.assert_bit_equal(fused, chain, "fused kernel vs multi-kernel chain")
    util.assert_bit_equal(chain, util.run_oracle(w.image, w.rects, w.dsize, ops, **kw), "chain vs oracle")


def test_default_mode_differs_from_the_8bit_chain_as_documented():
    """the reference's fused kernel keeps the interpolated value in float and contracts mul+sub, so on
    a real image it is NOT the 8-bit chain; the difference is bounded by half a grey level times alpha/div."""
    if gpu_util.chain_lib() is None:
        pytest.skip("oracle/libchain.so not built")
    w = util.workload_c2(seed=22, n=20, pitch=6144)
    ops = [("mul", MUL), ("sub", SUB), ("div", DIV)]
    chain, _ = gpu_util.run_chain(w.image, w.rects, w.dsize, False, MUL, SUB, DIV)
    fused = gpu_util.run_cvgs(w.image, w.rects, w.dsize, ops, variant=2)
    diff = np.abs(fused - chain)
    assert diff.max() > 0
    bound = 0.5 * np.array(MUL) / np.array(DIV) * 1.0001 + 1e-6
    for c in range(3):
        assert diff[:, c].max() <= bound[c], (c, diff[:, c].max(), bound[c])
