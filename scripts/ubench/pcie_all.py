"""Aggregate PCIe ceiling of the box: the same H2D / D2H copies as pcie.py on every GPU at once (one process per GPU
under torchrun, barrier-aligned), so that the end-to-end scaling of bench.py at N = 8 can be read against what the
host-memory / PCIe path of the VM delivers in total.
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29530 scripts/ubench/pcie_all.py"""
import os, time, torch, torch.distributed as dist
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
MB_UP, MB_DOWN = 6.0, 4.9  # per frame of the bench's e2e arm
h = torch.empty(int(MB_UP * 1e6), dtype=torch.uint8).pin_memory()
d = torch.empty_like(h, device="cuda")
ho = torch.empty(int(MB_DOWN * 1e6), dtype=torch.uint8).pin_memory()
do = torch.empty_like(ho, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def run(up, down, n=300):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        if up:
            with torch.cuda.stream(s1):
                d.copy_(h, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                ho.copy_(do, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / n


for name, up, down in (("H2D only", True, False), ("D2H only", False, True), ("H2D + D2H", True, True)):
    run(up, down, 20)
    s = run(up, down)
    if rank == 0:
        gu = world * MB_UP * 1e6 / s / 1e9 if up else 0.0
        gd = world * MB_DOWN * 1e6 / s / 1e9 if down else 0.0
        print(f"{world} GPU(s), {name}: {s * 1e6:.1f} us per round; aggregate H2D {gu:.1f} GB/s, D2H {gd:.1f} GB/s "
              f"({gu / world:.1f} / {gd / world:.1f} per GPU)" + (f"; = {world / s / 1e3 * 50:.0f} K crops/s if this were the e2e arm" if up and down else ""), flush=True)
if world > 1:
    dist.destroy_process_group()
