"""Small fits through every fit kernel and host path, for compute-sanitizer (memcheck / racecheck /
synccheck): polynomial (per-channel, screening and warp kernels, forced in turn), cubic spline,
sinusoid (tensor-core kernel), the general kernel, and the three host paths (single arena, serial,
pipeline)."""
import os
import sys
sys.path.insert(0, ".")
import numpy as np
from sakura_b200 import synth
from sakura_b200.capi import SakuraLib
lib = SakuraLib()
for n, order, nfit, kernel in ((4096, 5, 5, "warp"), (4096, 5, 5, "screen"), (4096, 5, 5, "channel"), (8192, 3, 4, None),
                               (2048, 5, 5, "warp"), (4096, 7, 6, None), (1024, 9, 3, None)):
    data, mask = synth.make_spectra(48, n, seed=n + order)
    synth.add_edge_case_rows(data, mask, order + 1)
    st, ctx = lib.create_polynomial_context(order, n)
    if kernel is not None:
        os.environ["SAKURA_B200_POLY_KERNEL"] = kernel
    for chunk, packed in ((4096, 0), (1, 0), (4096, 1 << 20)):
        lib.set_host_pipeline_tuning(chunk, packed)
        g = lib.lsqfit_polynomial_batch(ctx, order, data, mask, 3.0, nfit, want_best_fit=True)
        print("poly", n, order, nfit, lib.last_kernel_name(), g["call_status"], int((g["status"] == 0).sum()), lib.last_handover_count(ctx))
    os.environ.pop("SAKURA_B200_POLY_KERNEL", None)
    lib.destroy_context(ctx)
lib.set_host_pipeline_tuning(96, 2048)
data, mask = synth.make_spectra(70, 2048, seed=5, ripple=True)
synth.add_edge_case_rows(data, mask, 9)
st, ctx = lib.create_cubicspline_context(6, 2048)
g = lib.lsqfit_cubicspline_batch(ctx, 6, data, mask, 3.0, 3, want_best_fit=True)
print("spline", lib.last_kernel_name(), g["call_status"], int((g["status"] == 0).sum()))
lib.destroy_context(ctx)
st, ctx = lib.create_sinusoid_context(6, 2048)
g = lib.lsqfit_sinusoid_batch(ctx, [0, 1, 2, 3, 5, 6], data, mask, 3.0, 3, want_best_fit=True)
print("sinusoid", lib.last_kernel_name(), g["call_status"], int((g["status"] == 0).sum()))
lib.destroy_context(ctx)
