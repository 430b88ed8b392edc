"""Generate tests/golden/*.npz by running the REFERENCE's own fused kernel on a GPU.

Run on a B200 box (no /root/reference needed there: oracle/_ref/libfkref_16.so travels with the repo):
    gpurun -- 'python tests/golden/make_golden.py gpurun_out/golden && ls gpurun_out/golden'
then copy gpurun_out/golden/*.npz to tests/golden/.  The CPU test
tests/test_oracle_golden.py::test_golden_vectors_from_reference_kernel checks oracle/oracle.c against them.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import gpu_util, util  # noqa: E402

MUL, SUB, DIV = (0.3, 0.3, 0.3), (1.0, 4.0, 3.2), (3.2, 0.6, 11.8)


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    rng = np.random.default_rng(2024)
    cases = []
    img = util.make_image(rng, 160, 120, pitch=512)
    cases.append(("random_mixed", img, [(0, 0, 160, 120), (3, 5, 24, 48), (50, 2, 101, 33), (17, 90, 7, 5),
                                        (140, 0, 20, 120), (1, 1, 64, 118), (80, 60, 80, 60)], (64, 128), 1, 1, (0, 0, 0), 7))
    img2 = util.make_image(rng, 96, 200, smooth=True)
    cases.append(("smooth_ar", img2, [(0, 0, 30, 120), (10, 20, 60, 30), (0, 0, 96, 200)], (64, 128), 0, 0,
                  (128.0, 7.5, 250.0), 3))
    cases.append(("smooth_ar_even_left", img2, [(1, 1, 31, 121), (10, 20, 61, 29)], (64, 128), 2, 1, (1.0, 2.0, 3.0), 2))
    cases.append(("smooth_ar_left", img2, [(1, 1, 31, 121), (10, 20, 61, 29)], (64, 128), 3, 0, (1.0, 2.0, 3.0), 2))
    cases.append(("partial_batch", img, [(0, 0, 60, 120), (1, 1, 60, 120), (2, 2, 60, 118)], (32, 40), 1, 1,
                  (9.0, 8.0, 7.0), 2))
    for name, image, rects, dsize, aspect, swap, bg, used in cases:
        out = gpu_util.run_fkref(image, rects, dsize, swap, MUL, SUB, DIV, aspect=aspect, bg=bg, batch=16, used=used)
        np.savez_compressed(os.path.join(out_dir, f"{name}.npz"), image=image, rects=np.array(rects, dtype=np.int32),
                            dsize=np.array(dsize, dtype=np.int32), aspect=aspect, swap=swap,
                            bg=np.array(bg, dtype=np.float32), used=used, n_planes=16,
                            mul=np.array(MUL, dtype=np.float32), sub=np.array(SUB, dtype=np.float32),
                            div=np.array(DIV, dtype=np.float32), out=out)
        print(name, out.shape, float(np.nanmean(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__)))
