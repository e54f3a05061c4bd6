"""One-off stress: screening vs per-channel kernel on large device batches, several shapes."""
import os, sys
sys.path.insert(0, ".")
import torch
from sakura_b200.capi import SakuraLib
from sakura_b200.device_api import DeviceFits
import bench
lib = SakuraLib(); dev = DeviceFits(lib)
device = torch.device("cuda", 0)
for rows, n, order, nfit, clip, seed in ((1 << 19, 4096, 3, 10, 2.5, 1), (1 << 19, 4096, 5, 7, 3.5, 2), (1 << 18, 8192, 5, 5, 3.0, 3),
                                         (1 << 19, 4096, 1, 6, 2.0, 4), (1 << 19, 4096, 6, 5, 3.0, 5)):
    data, mask = bench.generate_on_device(torch, rows, n, seed=seed, device=device)
    st, ctx = lib.create_polynomial_context(order, n)
    os.environ["SAKURA_B200_POLY_SCREEN"] = "1"
    a = dev.polynomial(ctx, order, data, mask, clip, nfit)
    handed = lib.last_handover_count(ctx)
    os.environ["SAKURA_B200_POLY_SCREEN"] = "0"
    b = dev.polynomial(ctx, order, data, mask, clip, nfit)
    mism = int((a["final_mask"] != b["final_mask"]).sum())
    stm = int((a["status"] != b["status"]).sum())
    rel = float(((a["rms"] - b["rms"]).abs() / b["rms"].abs()).max())
    print(f"rows {rows} n {n} order {order} nfit {nfit} clip {clip}: mask mismatches {mism}, status mismatches {stm}, "
          f"rms max rel diff {rel:.2e}, handed over {handed}, clipped/row {float((mask & ~a['final_mask']).sum()) / rows:.1f}")
    lib.destroy_context(ctx)
    del data, mask, a, b
