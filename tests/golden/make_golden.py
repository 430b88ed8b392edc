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

    # other source types of the same chain: 16-bit and 4-channel pixels (the reference's ushort3 / short3 / uchar4 /
    # ushort4 / short4 instantiations) and NV12 frames (ReadYUV + ConvertYUVToRGB in front of the resize)
    from cvgpuspeedup_b200 import _abi
    mul4, sub4, div4, bg4 = (0.3, 0.3, 0.3, 0.25), (1.0, 4.0, 3.2, 0.5), (3.2, 0.6, 11.8, 2.0), (128.0, 3.5, 250.0, 7.0)
    rects = [(0, 0, 96, 72), (5, 7, 24, 48), (40, 3, 50, 33), (17, 50, 7, 5), (95, 71, 1, 1)]
    for tname, st in [("16uc3", _abi.CVGS_16UC3), ("16sc3", _abi.CVGS_16SC3), ("8uc4", _abi.CVGS_8UC4),
                      ("16uc4", _abi.CVGS_16UC4), ("16sc4", _abi.CVGS_16SC4)]:
        nc = util.channels_of(st)
        image = rng.integers(0, 256, size=(72, 96 * util.px_bytes_of(st) + 32), dtype=np.uint8)
        out = gpu_util.run_fkref(image, rects, (40, 56), 1, mul4[:nc], sub4[:nc], div4[:nc], aspect=0, bg=bg4[:nc], batch=16,
                                 used=5, src_type=st)
        np.savez_compressed(os.path.join(out_dir, f"type_{tname}.npz"), image=image, rects=np.array(rects, dtype=np.int32),
                            dsize=np.array((40, 56), dtype=np.int32), aspect=0, swap=1, bg=np.array(bg4[:nc], dtype=np.float32),
                            used=5, n_planes=16, mul=np.array(mul4[:nc], dtype=np.float32),
                            sub=np.array(sub4[:nc], dtype=np.float32), div=np.array(div4[:nc], dtype=np.float32),
                            src_type=st, out=out)
        print("type", tname, out.shape, float(np.nanmean(out)))
    for standard in range(4):
        w, h, pitch = 98, 66, 128
        frame = rng.integers(0, 256, size=(h + (h + 1) // 2, pitch), dtype=np.uint8)
        out = gpu_util.run_fkref_nv12(frame, w, h, (40, 56), standard, MUL, SUB, DIV)
        np.savez_compressed(os.path.join(out_dir, f"nv12_std{standard}.npz"), image=frame, width=w, height=h,
                            dsize=np.array((40, 56), dtype=np.int32), standard=standard, mul=np.array(MUL, dtype=np.float32),
                            sub=np.array(SUB, dtype=np.float32), div=np.array(DIV, dtype=np.float32), out=out)
        print("nv12", standard, out.shape, float(np.nanmean(out)))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__)))
