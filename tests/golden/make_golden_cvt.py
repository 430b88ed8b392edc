"""Golden vectors for the channel-count changing colour conversions, produced by the REFERENCE's own kernel
(oracle/_ref/libfkref_16.so: Resize + ColorConversion<code> + Mul + Sub + write) on a GPU box:
    python tests/golden/make_golden_cvt.py gpurun_out/      then copy cvt_code*.npz into tests/golden/.
A 96x64 random source; destinations 96x64 (scale 1: exact pixel values reach the conversion, which is what separates
the two FMUL orders of the gray codes) and 40x24."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from tests import gpu_util  # noqa: E402

MUL, SUB = (0.3, 1.7, 0.05, 0.25), (1.0, -4.0, 3.2, 0.5)


def main(out_dir):
    for code, (nc, _) in sorted(gpu_util.CVT_CODES.items()):
        rng = np.random.default_rng(4000 + code)
        w, h = 96, 64
        img = rng.integers(0, 256, size=(h, w * nc + 32), dtype=np.uint8)
        outs = {f"out_{dw}x{dh}": gpu_util.run_fkref_cvt(code, img, w, h, (dw, dh), MUL, SUB) for dw, dh in [(96, 64), (40, 24)]}
        np.savez_compressed(os.path.join(out_dir, f"cvt_code{code}.npz"), image=img, width=w, height=h, code=code, channels=nc,
                            mul=np.array(MUL, dtype=np.float32), sub=np.array(SUB, dtype=np.float32), **outs)
        print(code, {k: v.shape for k, v in outs.items()})


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
