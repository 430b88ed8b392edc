import torch, time
h = torch.empty((1080, 6144), dtype=torch.uint8).pin_memory()
d = torch.empty((1080, 6144), dtype=torch.uint8, device="cuda")
ho = torch.empty((50*3*128*64,), dtype=torch.float32).pin_memory()
do = torch.empty_like(ho, device="cuda")
def t(fn, n=200):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
us = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1D {h.numel()/1e6:.2f} MB: {us:.1f} us = {h.numel()/us/1e3:.1f} GB/s")
us = t(lambda: d[:, :5760].copy_(h[:, :5760], non_blocking=True)); print(f"H2D 2D 5760 of 6144: {us:.1f} us = {1080*5760/us/1e3:.1f} GB/s")
us = t(lambda: ho.copy_(do, non_blocking=True)); print(f"D2H 1D {ho.numel()*4/1e6:.2f} MB: {us:.1f} us = {ho.numel()*4/us/1e3:.1f} GB/s")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): ho.copy_(do, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200): both()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 200 * 1e6
print(f"H2D 6.6MB + D2H 4.9MB concurrently: {dt:.1f} us per pair -> H2D {h.numel()/dt/1e3:.1f} GB/s, D2H {ho.numel()*4/dt/1e3:.1f} GB/s")
