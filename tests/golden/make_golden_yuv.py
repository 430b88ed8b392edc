"""Golden vectors for the other YUV readers (NV21, P010, P210, Y210), produced by the REFERENCE's own kernel
(oracle/_ref/libfkref_16.so, -DFKREF_YUV) on a GPU box:
    python tests/golden/make_golden_yuv.py gpurun_out/      then copy yuv_*.npz into tests/golden/.
Same keys as nv12_std*.npz plus `src_type`; read by tests/test_oracle_golden.py."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from cvgpuspeedup_b200 import _abi  # noqa: E402
from tests import gpu_util  # noqa: E402

MUL, SUB, DIV = (1 / 255.0,) * 3, (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


def main(out_dir):
    rng = np.random.default_rng(2026)
    names = {_abi.CVGS_NV21: "nv21", _abi.CVGS_P010: "p010", _abi.CVGS_P210: "p210", _abi.CVGS_Y210: "y210"}
    for fmt, name in names.items():
        for standard in (0, 3):
            w, h, pitch = 98, 66, 512
            rows = {_abi.CVGS_NV21: h + (h + 1) // 2, _abi.CVGS_P010: h + (h + 1) // 2, _abi.CVGS_P210: 2 * h, _abi.CVGS_Y210: h}[fmt]
            frame = rng.integers(0, 256, size=(rows, pitch), dtype=np.uint8)
            out = gpu_util.run_fkref_yuv(fmt, frame, w, h, (40, 56), standard, MUL, SUB, DIV)
            np.savez_compressed(os.path.join(out_dir, f"yuv_{name}_std{standard}.npz"), image=frame, width=w, height=h,
                                dsize=np.array((40, 56), dtype=np.int32), standard=standard, src_type=fmt,
                                mul=np.array(MUL, dtype=np.float32), sub=np.array(SUB, dtype=np.float32),
                                div=np.array(DIV, dtype=np.float32), out=out)
            print(name, standard, out.shape, float(np.nanmean(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out")
