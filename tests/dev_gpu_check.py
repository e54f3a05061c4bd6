"""Ad-hoc development check (not collected by pytest): GPU library vs oracle on small inputs."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from sakura_b200.capi import SakuraLib
from sakura_b200 import synth
from oracle.refbind import Oracle

lib = SakuraLib()
ora = Oracle()
print("oracle kind:", ora.kind)

def cmp(name, a, b):
    ok = True
    for k in ("status", "lsqfit_status"):
        if not np.array_equal(a[k], b[k]):
            print(name, k, "MISMATCH", np.nonzero(a[k] != b[k])[0][:10]); ok = False
    good = (b["status"] == 0)
    fm = (a["final_mask"][good] != b["final_mask"][good]).sum()
    print(f"{name}: rows={len(good)} ok_rows={good.sum()} final_mask mismatches={fm}")
    for k in ("coeff", "residual", "best_fit", "rms"):
        if a.get(k) is None or b.get(k) is None: continue
        x = a[k][good].astype(np.float64); y = b[k][good].astype(np.float64)
        d = np.abs(x - y); rel = d / np.maximum(np.abs(y), 1e-300)
        print(f"   {k}: max abs {np.nanmax(d):.3e} max rel {np.nanmax(np.where(d>1e-6, rel, 0)):.3e} nan_mismatch={(np.isnan(x)!=np.isnan(y)).sum()}")
    if "boundary" in a:
        print("   boundary equal:", np.array_equal(a["boundary"][good], b["boundary"][good]))
    return ok

R, n = 64, 4096
data, mask = synth.make_spectra(R, n)
names = synth.add_edge_case_rows(data, mask, 6)
st, ctx = lib.create_polynomial_context(5, n)
assert st == 0
t=time.time(); g = lib.lsqfit_polynomial_batch(ctx, 5, data, mask, 3.0, 3, want_best_fit=True); print("gpu call", time.time()-t, "status", g["call_status"], lib.last_kernel_name(), lib.last_error())
o = ora.lsqfit_polynomial_batch(5, data, mask, 3.0, 3, want_best_fit=True)
print("statuses gpu", g["status"][:8], g["lsqfit_status"][:8]); print("statuses ref", o["status"][:8], o["lsqfit_status"][:8])
cmp("poly5", g, o)
lib.destroy_context(ctx)

n = 1024
data, mask = synth.make_spectra(32, n, ripple=True)
st, ctx = lib.create_cubicspline_context(8, n); assert st == 0
g = lib.lsqfit_cubicspline_batch(ctx, 8, data, mask, 3.0, 3, want_best_fit=True); print("spline call", g["call_status"], lib.last_kernel_name())
o = ora.lsqfit_cubicspline_batch(8, data, mask, 3.0, 3, want_best_fit=True)
cmp("spline8", g, o)
lib.destroy_context(ctx)

st, ctx = lib.create_sinusoid_context(10, n); assert st == 0
nw = list(range(0, 11))
g = lib.lsqfit_sinusoid_batch(ctx, nw, data, mask, 3.0, 3, want_best_fit=True); print("sinusoid call", g["call_status"], lib.last_kernel_name())
o = ora.lsqfit_sinusoid_batch(nw, data, mask, 3.0, 3, want_best_fit=True)
cmp("sinusoid10", g, o)
lib.destroy_context(ctx)

d = np.arange(256, dtype=np.float32); m = np.ones(256, dtype=bool)
print("stats gpu", lib.compute_statistics(d, m)); print("stats ref", ora.statistics_batch(d, m)[0])
print("launches", lib.kernel_launch_count())

# ---- quick device-side timing of the polynomial path
import torch, ctypes as C
R, n = 65536, 4096
g = torch.Generator(device="cuda"); g.manual_seed(1)
x = torch.linspace(0, 1, n, device="cuda", dtype=torch.float64)
co = (torch.rand(R, 6, device="cuda", dtype=torch.float64, generator=g) - 0.5) * 10
y = torch.zeros(R, n, device="cuda", dtype=torch.float64)
for k in range(5, -1, -1):
    y = y * x + co[:, k:k+1]
y += torch.randn(R, n, device="cuda", dtype=torch.float64, generator=g)
spk = torch.rand(R, n, device="cuda", generator=g) < 0.001
y += spk * 30.0
d_data = y.float().contiguous(); del y
d_mask = (torch.rand(R, n, device="cuda", generator=g) >= 0.02)
d_mask[:, :32] = False; d_mask[:, -32:] = False
d_coeff = torch.empty(R, 6, device="cuda", dtype=torch.float64)
d_res = torch.empty(R, n, device="cuda", dtype=torch.float32)
d_fm = torch.empty(R, n, device="cuda", dtype=torch.bool)
d_rms = torch.empty(R, device="cuda", dtype=torch.float32)
d_st = torch.empty(R, device="cuda", dtype=torch.int32)
d_ls = torch.empty(R, device="cuda", dtype=torch.int32)
st, ctx = lib.create_polynomial_context(5, n)
def run(nfit, res=True):
    return lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(ctx, 5, R, n, d_data.data_ptr(), n, d_mask.data_ptr(), n,
        3.0, nfit, 6, d_coeff.data_ptr(), None, d_res.data_ptr() if res else None, d_fm.data_ptr(), d_rms.data_ptr(),
        d_st.data_ptr(), d_ls.data_ptr(), None)
for nfit in (1, 3, 5):
    for res in (True, False):
        assert run(nfit, res) == 0, lib.last_error()
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): run(nfit, res)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        byts = R * (n * (10 if res else 6) + 56)
        print(f"nfit={nfit} residual={res}: {ms:.3f} ms  {R/ms*1e3:.3e} spectra/s  {byts/ms/1e6:.0f} GB/s algorithmic  kernel={lib.last_kernel_name()}")
print("status ok rows:", int((d_st == 0).sum()), "mean clipped:", float((d_mask & ~d_fm).sum(dim=1).float().mean()))
# parity of the device path on a slice vs oracle
sl = slice(0, 512)
o = ora.lsqfit_polynomial_batch(5, d_data[sl].cpu().numpy(), d_mask[sl].cpu().numpy(), 3.0, 5)
run(5, True); torch.cuda.synchronize()
fm = d_fm[sl].cpu().numpy(); print("device-path final_mask mismatches vs oracle:", (fm != o["final_mask"]).sum(), "rms max diff", np.abs(d_rms[sl].cpu().numpy() - o["rms"]).max(), "resid max diff", np.abs(d_res[sl].cpu().numpy() - o["residual"]).max(), "coeff max rel", np.max(np.abs(d_coeff[sl].cpu().numpy() - o["coeff"]) / np.abs(o["coeff"])))
