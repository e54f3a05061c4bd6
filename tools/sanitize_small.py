"""Small polynomial fits through both kernels, for compute-sanitizer (memcheck / racecheck)."""
import sys
sys.path.insert(0, ".")
import numpy as np
from sakura_b200 import synth
from sakura_b200.capi import SakuraLib
lib = SakuraLib()
for n, order, nfit in ((4096, 5, 5), (8192, 3, 4), (2048, 5, 5), (4096, 7, 6)):
    data, mask = synth.make_spectra(48, n, seed=n + order)
    synth.add_edge_case_rows(data, mask, order + 1)
    st, ctx = lib.create_polynomial_context(order, n)
    g = lib.lsqfit_polynomial_batch(ctx, order, data, mask, 3.0, nfit, want_best_fit=True)
    print(n, order, nfit, lib.last_kernel_name(), g["call_status"], int((g["status"] == 0).sum()), lib.last_handover_count(ctx))
    lib.destroy_context(ctx)
