"""BASELINE config 5: n crops (default 8192) of one 4K frame sharded over the ranks, per-GPU fused kernel writing its
slab in place + ONE in-place NCCL all-gather into a contiguous NCHW tensor on every rank.
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P scripts/run_c5.py [n] [reps]
Prints kernel-only and kernel+gather throughput (device time, max over ranks) and checks planes against the oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import cvgpuspeedup_b200 as cvGS
from cvgpuspeedup_b200 import sharding
from tests import util

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cvGS._abi.load().cvgs_b200_set_overlap(1)
w = util.workload_c3(seed=5, n=n)                     # same frame and rect list on every rank
d_img = torch.from_numpy(w.image).cuda()
frame = cvGS.GpuMat(d_img.data_ptr(), w.width, w.height, w.pitch, owner=d_img)
crops = [frame.roi(*r) for r in w.rects]
out = torch.full((n, 3, w.dsize[1], w.dsize[0]), float("nan"), dtype=torch.float32, device="cuda")
ops = [cvGS.cvtColor(cvGS.COLOR_BGR2RGB), cvGS.multiply((1 / 255.0,) * 3), cvGS.subtract((0.485, 0.456, 0.406)),
       cvGS.divide((0.229, 0.224, 0.225))]
lo, hi = sharding.shard_range(n, rank, world)
# one cvGS::executeOperations call per `chunk` crops (argv[3], default: the whole shard in one launch)
chunk = int(sys.argv[3]) if len(sys.argv) > 3 else max(1, hi - lo)
prebuilt = [(cvGS.resize(crops[a:min(hi, a + chunk)], w.dsize, min(hi, a + chunk) - a), cvGS.split(out[a:min(hi, a + chunk)], w.dsize))
            for a in range(lo, hi, chunk)]
stream = torch.cuda.current_stream()

def kernels():
    for rd, wr in prebuilt:
        cvGS.executeOperations(stream, rd, *ops, wr)

def timed(fn):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

kernels(); sharding.gather_slabs(out, n); torch.cuda.synchronize()
ms_k = timed(kernels)
ms_kg = timed(lambda: (kernels(), sharding.gather_slabs(out, n)))
idx = sorted(set([0, n // 3, n // 2, n - 1, lo, max(lo, hi - 1)]))
want = util.run_oracle(w.image, [w.rects[i] for i in idx], w.dsize, w.ops)
util.assert_bit_equal(out[idx].cpu().numpy(), want, f"rank {rank}: gathered planes vs oracle")
if rank == 0:
    gb = n * 3 * w.dsize[0] * w.dsize[1] * 4 / 1e9
    print(f"c5: {n} crops -> {w.dsize}, {world} GPU(s): kernels {ms_k:.3f} ms ({n / ms_k / 1e3:.2f} Mcrops/s), "
          f"kernels + all-gather of {gb:.2f} GB {ms_kg:.3f} ms ({n / ms_kg / 1e3:.2f} Mcrops/s); parity ok on planes {idx}")
if world > 1:
    dist.destroy_process_group()
