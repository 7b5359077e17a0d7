import torch, time
dev = torch.device("cuda", 0)
n = 115_000_000 // 4
src = torch.empty(n, dtype=torch.float32).pin_memory()
dst = torch.empty(n, dtype=torch.float32, device=dev)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
def one(): dst.copy_(src, non_blocking=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
h = n // 2
def two():
    with torch.cuda.stream(s1): dst[:h].copy_(src[:h], non_blocking=True)
    with torch.cuda.stream(s2): dst[h:].copy_(src[h:], non_blocking=True)
ss = [torch.cuda.Stream() for _ in range(4)]
q = n // 4
def four():
    for k, st in enumerate(ss):
        with torch.cuda.stream(st): dst[k*q:(k+1)*q].copy_(src[k*q:(k+1)*q], non_blocking=True)
for name, fn in (("one", one), ("two", two), ("four", four)):
    t = timeit(fn)
    print(name, f"{t*1e3:.3f} ms  {n*4/t/1e9:.1f} GB/s")
