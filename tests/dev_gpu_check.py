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
