"""Development check of the warp-per-spectrum polynomial kernel (csrc/lsqfit_poly_warp.cu): against the
per-channel kernel on a large device batch (final_mask / status bit-exact, rms, residual, coefficients), against
the oracle on a small batch with the fixed edge-case rows, and a timing of the C2 shape.
    python tools/dev_warp_check.py [rows] [nfit]"""
import os, sys
sys.path.insert(0, ".")
import numpy as np, torch
from sakura_b200.capi import SakuraLib
from sakura_b200 import synth
import bench

lib = SakuraLib()
dev = torch.device("cuda", 0)
R = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nfit = int(sys.argv[2]) if len(sys.argv) > 2 else 5


def run(kernel, data, mask, order, nfit, clip=3.0, time_it=False):
    R, n = data.shape
    K = order + 1
    os.environ["SAKURA_B200_POLY_KERNEL"] = kernel
    out = dict(res=torch.full((R, n), float("nan"), dtype=torch.float32, device=dev),
               fm=torch.zeros(R, n, dtype=torch.bool, device=dev),
               rms=torch.full((R,), float("nan"), dtype=torch.float32, device=dev),
               st=torch.full((R,), -1, dtype=torch.int32, device=dev), ls=torch.full((R,), -1, dtype=torch.int32, device=dev),
               co=torch.full((R, K), float("nan"), dtype=torch.float64, device=dev))
    s, ctx = lib.create_polynomial_context(order, n)
    assert s == 0
    ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
    reps = 3 if time_it else 1
    for it in range(reps):
        ev0.record()
        rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(ctx, order, R, n, data.data_ptr(), n, mask.data_ptr(), n, clip, nfit, K,
                                                            out["co"].data_ptr(), None, out["res"].data_ptr(), out["fm"].data_ptr(),
                                                            out["rms"].data_ptr(), out["st"].data_ptr(), out["ls"].data_ptr(), None)
        ev1.record()
        assert rc == 0, (rc, lib.last_error())
    torch.cuda.synchronize()
    out["ms"] = ev0.elapsed_time(ev1)
    out["kernel"] = lib.last_kernel_name()
    out["handed"] = lib.last_handover_count(ctx)
    lib.destroy_context(ctx)
    del os.environ["SAKURA_B200_POLY_KERNEL"]
    return out


def compare(a, b, tag):
    ok = a["st"] == 0
    fm = int((a["fm"] != b["fm"]).sum())
    stq = int((a["st"] != b["st"]).sum()) + int((a["ls"] != b["ls"]).sum())
    rms = float(((a["rms"][ok] - b["rms"][ok]).abs() / b["rms"][ok].abs().clamp_min(1e-30)).max()) if ok.any() else 0.0
    rd = (a["res"][ok] - b["res"][ok]).abs()
    rd = float(rd[~torch.isnan(rd)].max()) if ok.any() else 0.0
    nanmis = int((torch.isnan(a["res"]) != torch.isnan(b["res"])).sum())
    cd = float(((a["co"][ok] - b["co"][ok]).abs() / b["co"][ok].abs().clamp_min(1e-12)).max()) if ok.any() else 0.0
    print(f"{tag}: final_mask mismatches {fm}, status mismatches {stq}, rms max rel {rms:.2e}, residual max abs {rd:.2e} (nan mismatch {nanmis}), coeff max rel {cd:.2e}")
    return fm == 0 and stq == 0 and nanmis == 0


allok = True
for n, order, nf in ((4096, 5, nfit), (2048, 5, nfit), (4096, 3, 4), (4096, 7, 6), (2048, 0, 3), (4096, 1, 10)):
    data, mask = bench.generate_on_device(torch, R if n == 4096 else R // 2, n, seed=20240611 + n + order, device=dev)
    w = run("warp", data, mask, order, nf, time_it=True)
    c = run("channel", data, mask, order, nf, time_it=True)
    print(f"n={n} order={order} nfit={nf}: warp {w['kernel']} {w['ms']:.3f} ms ({data.shape[0] / w['ms'] * 1e3:.3e}/s, handed {w['handed']}), "
          f"channel {c['kernel']} {c['ms']:.3f} ms ({data.shape[0] / c['ms'] * 1e3:.3e}/s)")
    allok &= compare(w, c, "  warp vs channel")
    del data, mask

# oracle on a small batch with the edge-case rows
from oracle.refbind import Oracle
ora = Oracle()
for n in (4096, 2048):
    data, mask = synth.make_spectra(512, n)
    synth.add_edge_case_rows(data, mask, 6)
    o = ora.lsqfit_polynomial_batch(5, data, mask, 3.0, 5)
    w = run("warp", torch.from_numpy(data).to(dev), torch.from_numpy(mask).to(dev), 5, 5)
    ok = o["status"] == 0
    fm = int((w["fm"].cpu().numpy()[ok] != o["final_mask"][ok]).sum())
    st = int((w["st"].cpu().numpy() != o["status"]).sum()) + int((w["ls"].cpu().numpy() != o["lsqfit_status"]).sum())
    rms = np.abs(w["rms"].cpu().numpy()[ok] - o["rms"][ok]).max()
    rd = np.nanmax(np.abs(w["res"].cpu().numpy()[ok] - o["residual"][ok]))
    print(f"oracle n={n}: kernel {w['kernel']} handed {w['handed']} final_mask mismatches {fm} status mismatches {st} rms max abs {rms:.2e} residual max abs {rd:.2e}")
    allok &= fm == 0 and st == 0
print("ALL OK" if allok else "FAILURES")
