"""Small reproducer: warp kernel, clip 0.5, order 4, 12 fits (for compute-sanitizer)."""
import os, sys
sys.path.insert(0, ".")
import numpy as np
from sakura_b200.capi import SakuraLib
from sakura_b200 import synth
lib = SakuraLib()
clip = float(sys.argv[1]) if len(sys.argv) > 1 else 0.5
data, mask = synth.make_spectra(128, 4096, seed=int(clip * 10))
os.environ["SAKURA_B200_POLY_KERNEL"] = "warp"
st, ctx = lib.create_polynomial_context(4, 4096)
g = lib.lsqfit_polynomial_batch(ctx, 4, data, mask, clip, 12)
print("status", g["call_status"], lib.last_kernel_name(), "handed", lib.last_handover_count(ctx), lib.last_error())
print("rows ok", int((g["status"] == 0).sum()))
