"""Where does the host pipeline lose against the copy-only ceiling? e2e with fewer fits (cheaper
kernel), and the copy pattern of the pipeline alone (same chunking, no kernel)."""
import sys, time
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
import bench
lib = SakuraLib(); R, n, K = 262144, 4096, 6
dev = torch.device("cuda", 0)
data, mask = bench.generate_on_device(torch, R, n, seed=1, device=dev)
h = {k: torch.empty(R, n, dtype=t, pin_memory=True) for k, t in (("data", torch.float32), ("mask", torch.bool), ("resid", torch.float32), ("fmask", torch.bool))}
h["data"].copy_(data); h["mask"].copy_(mask)
co = torch.empty(R, K, dtype=torch.float64, pin_memory=True); rms = torch.empty(R, dtype=torch.float32, pin_memory=True)
st = torch.empty(R, dtype=torch.int32, pin_memory=True); ls = torch.empty(R, dtype=torch.int32, pin_memory=True)
s, ctx = lib.create_polynomial_context(5, n)
def call(nfit, resid=True):
    rc = lib.lib.sakura_LSQFitPolynomialFloatBatch(ctx, 5, R, n, h["data"].data_ptr(), n, h["mask"].data_ptr(), n, 3.0, nfit, K, co.data_ptr(), None,
                                                   h["resid"].data_ptr() if resid else None, h["fmask"].data_ptr(), rms.data_ptr(), st.data_ptr(), ls.data_ptr())
    assert rc == 0, lib.last_error()
def timeit(fn, reps=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
for nfit in (5, 1):
    dt = timeit(lambda: call(nfit)); print(f"e2e nfit {nfit}: {dt*1e3:.1f} ms, {R*n*5/dt/1e9:.1f} GB/s per direction")
dt = timeit(lambda: call(5, False)); print(f"e2e nfit 5, no residual: {dt*1e3:.1f} ms, up {R*n*5/dt/1e9:.1f} GB/s down {R*n/dt/1e9:.1f} GB/s")
# the pipeline's copy pattern alone: chunks of 24576 rows, data+mask up, resid+fmask down, 4 slots
rows = 24576
d = [dict(data=torch.empty(rows, n, dtype=torch.float32, device=dev), mask=torch.empty(rows, n, dtype=torch.bool, device=dev),
          resid=torch.empty(rows, n, dtype=torch.float32, device=dev), fmask=torch.empty(rows, n, dtype=torch.bool, device=dev),
          up=torch.cuda.Stream(), down=torch.cuda.Stream()) for _ in range(4)]
def pattern(one_copy=False):
    for ci, r0 in enumerate(range(0, R, rows)):
        sl = d[ci % 4]; r1 = min(R, r0 + rows)
        with torch.cuda.stream(sl["up"]):
            sl["data"][:r1 - r0].copy_(h["data"][r0:r1], non_blocking=True); sl["mask"][:r1 - r0].copy_(h["mask"][r0:r1], non_blocking=True)
        with torch.cuda.stream(sl["down"]):
            h["resid"][r0:r1].copy_(sl["resid"][:r1 - r0], non_blocking=True); h["fmask"][r0:r1].copy_(sl["fmask"][:r1 - r0], non_blocking=True)
dt = timeit(pattern); print(f"copy pattern only (4 slots x 2 streams, chunked): {dt*1e3:.1f} ms, {R*n*5/dt/1e9:.1f} GB/s per direction")
big = torch.empty(R * n * 5, dtype=torch.uint8, device=dev); big2 = torch.empty(R * n * 5, dtype=torch.uint8, device=dev)
hu = torch.empty(R * n * 5, dtype=torch.uint8, pin_memory=True); hd = torch.empty(R * n * 5, dtype=torch.uint8, pin_memory=True)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def two():
    with torch.cuda.stream(s1): big.copy_(hu, non_blocking=True)
    with torch.cuda.stream(s2): hd.copy_(big2, non_blocking=True)
dt = timeit(two); print(f"two single copies: {dt*1e3:.1f} ms, {R*n*5/dt/1e9:.1f} GB/s per direction")
# does a running fit kernel slow the copy engines down? the same copy pattern with the fit kernel looping on
# device-resident rows in another stream
import threading
rows2 = 24576
d2, m2 = data[:rows2].contiguous(), mask[:rows2].contiguous()
from sakura_b200.device_api import DeviceFits
fits = DeviceFits(lib)
ks = torch.cuda.Stream()
out2 = fits.polynomial(ctx, 5, d2, m2, 3.0, 5, stream=ks)
stop = False
def spin():
    while not stop:
        for _ in range(8):
            fits.polynomial(ctx, 5, d2, m2, 3.0, 5, out=out2, stream=ks)
        ks.synchronize()
t = threading.Thread(target=spin); t.start()
dt = timeit(pattern); print(f"copy pattern with the fit kernel running all the time: {dt*1e3:.1f} ms, {R*n*5/dt/1e9:.1f} GB/s per direction")
stop = True; t.join()
