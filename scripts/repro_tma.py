import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util, gpu_util
sel = [int(a) for a in sys.argv[1].split(",")]
w = util.workload_c2(n=50, pitch=6144)
rects = [w.rects[i] for i in sel]
try:
    got = gpu_util.run_cvgs(w.image, rects, w.dsize, w.ops, variant=2)
except Exception as e:
    print("crops", sel, rects, "FAILED", str(e).split("\n")[0]); sys.exit(0)
want = util.run_oracle(w.image, rects, w.dsize, w.ops)
bad = (got.view(np.uint32) != want.view(np.uint32))
print("crops", sel, "mismatches", int(bad.sum()), "of", bad.size)
