#!/usr/bin/env python
"""Headline benchmark: masked LSQ baseline fit with sigma-clipping (BASELINE.json north_star).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is ONE pass of the hot path over the whole synthetic batch resident in HBM:
one call of sakura_LSQFitPolynomialFloatBatchDevice = one launch of the screening fit kernel plus
one launch of the per-channel kernel over the (few) rows the former hands over.
Workload (BASELINE.json configs[1]): 1,048,576 spectra x 4,096 channels float32 + bool mask,
polynomial order 5, num_fitting_max = 5 clip iterations at 3 sigma; outputs coeff, residual,
final_mask, rms, status.  The 3-iteration figure named in the metric string is reported
next to it under "also".  Rows are independent: with N ranks every rank fits its own
1M-row shard (weak scaling), no data-path collective; timing is the max over ranks.

`--impl reference` times the reference's own CPU code (oracle/_ref, built from
/root/reference by oracle/Makefile; falls back to the C restatement) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

METRIC = "baseline-fit spectra/s (4096ch, poly-5, 3 clip iters); % of HBM roofline"
UNIT = "spectra/s"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1 << 20, help="spectra per GPU")
    ap.add_argument("--nchan", type=int, default=4096)
    ap.add_argument("--order", type=int, default=5)
    ap.add_argument("--nfit", type=int, default=5, help="num_fitting_max of the headline value")
    ap.add_argument("--clip", type=float, default=3.0)
    ap.add_argument("--e2e-rows", type=int, default=1 << 18, help="spectra per e2e step (host buffers)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-rows", type=int, default=1 << 16, help="spectra in the CPU sample")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-also", action="store_true")
    return ap.parse_args()


def bytes_per_spectrum(n: int, ncoeff: int, residual: bool = True) -> int:
    """Algorithmic HBM bytes of one whole fit: data + mask in, final_mask
    (+ residual) + coeff + rms + status out; independent of the iteration count."""
    return n * 4 + n + n + (n * 4 if residual else 0) + 8 * ncoeff + 4 + 4


def streaming_equivalent(n: int, nfit: int, rows: int, ms: float) -> dict:
    """SURVEY 8(d) secondary figure (north_star: "each clip iteration streams the data once"): the
    HBM bandwidth a kernel that re-streamed the row every clip iteration would need for the same
    spectra/s -- 6n bytes per iteration (data + mask in, mask out) + 4n once for the residual. The
    row-resident kernel moves fewer bytes than this; it is NOT the graded roofline fraction.
    Iterations executed are taken as num_fitting_max: with 4096 Gaussian channels clipped at 3 sigma
    an iteration that clips nothing (the only early exit) practically never happens."""
    from_peak = measured_peak_gbs()[0]
    b = nfit * 6 * n + 4 * n
    gbs = rows * b / (ms * 1e-3) / 1e9
    return {"bytes_per_spectrum": b, "iterations_assumed": nfit, "equivalent_gbs": gbs,
            "frac_of_hbm_peak": gbs / from_peak}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def measured_peaks() -> dict:
    try:
        with open(os.path.join(HERE, "MEASURED_PEAKS.json")) as fh:
            return json.load(fh)
    except Exception:
        return {}


def measured_peak_gbs():
    path = os.path.join(HERE, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)),
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- host placement
def gpu_numa_cpus(torch, index: int):
    """(NUMA node, CPUs of that node this process may run on) of CUDA device `index`, from sysfs; (None, None)
    when the box does not say (single node, container without sysfs)."""
    try:
        p = torch.cuda.get_device_properties(index)
        addr = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{addr}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None, None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            spec = fh.read().strip()
        cpus = set()
        for part in spec.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        return node, (cpus or None)
    except Exception:
        return None, None


class NearGpu:
    """Runs the enclosed allocations on the CPUs of the GPU's NUMA node, so that page-locked staging
    buffers (placed where the allocating thread runs) sit on the memory the GPU's PCIe root port is
    attached to; the previous affinity is restored on exit (the CPU baseline uses every core)."""

    def __init__(self, torch, index: int):
        self.node, self.cpus = gpu_numa_cpus(torch, index)
        self.saved = None

    def __enter__(self):
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except OSError:
                self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            os.sched_setaffinity(0, self.saved)


# ----------------------------------------------------------------------------- synthetic data
def generate_on_device(torch, rows: int, n: int, seed: int, device, ripple: bool = False):
    """Synthetic spectra of SURVEY 8(d) generated on the GPU (quintic continuum, Gaussian
    emission lines, single-channel spikes, N(0,1) noise; edges + 2 % random flags + one flagged
    block).  Returns (data float32 [rows, n], mask bool [rows, n])."""
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    data = torch.empty(rows, n, dtype=torch.float32, device=device)
    mask = torch.empty(rows, n, dtype=torch.bool, device=device)
    x = torch.linspace(0.0, 1.0, n, dtype=torch.float64, device=device)
    chan = torch.arange(n, dtype=torch.float32, device=device)[None, :]
    step = 16384
    for r0 in range(0, rows, step):
        r = min(step, rows - r0)
        co = (torch.rand(r, 6, dtype=torch.float64, device=device, generator=g) - 0.5) * 10.0
        y = torch.zeros(r, n, dtype=torch.float64, device=device)
        for k in range(5, -1, -1):
            y = y * x + co[:, k:k + 1]
        if ripple:  # standing wave of configs C3/C4 (as sakura_b200/synth.py)
            amp = 0.5 + 1.5 * torch.rand(r, 1, dtype=torch.float64, device=device, generator=g)
            kk = torch.randint(1, 21, (r, 1), device=device, generator=g).double()
            ph = 2.0 * np.pi * torch.rand(r, 1, dtype=torch.float64, device=device, generator=g)
            y = y + amp * torch.sin(2.0 * np.pi * kk * x[None, :] + ph)
        y = y.float()
        y += torch.randn(r, n, dtype=torch.float32, device=device, generator=g)
        nl = torch.randint(0, 4, (r, 1), device=device, generator=g)
        for j in range(3):
            peak = 5.0 + 45.0 * torch.rand(r, 1, device=device, generator=g)
            fwhm = 4.0 + 60.0 * torch.rand(r, 1, device=device, generator=g)
            cen = (n - 1.0) * torch.rand(r, 1, device=device, generator=g)
            sig = fwhm / 2.354820045
            y += torch.where(nl > j, peak * torch.exp(-0.5 * ((chan - cen) / sig) ** 2),
                             torch.zeros((), device=device))
        ns = torch.randint(0, 9, (r,), device=device, generator=g)
        rows_idx = torch.arange(r, device=device)
        for j in range(8):
            pos = torch.randint(0, n, (r,), device=device, generator=g)
            amp = (10.0 + 90.0 * torch.rand(r, device=device, generator=g)) * \
                (torch.randint(0, 2, (r,), device=device, generator=g).float() * 2.0 - 1.0)
            y[rows_idx, pos] += torch.where(ns > j, amp, torch.zeros((), device=device))
        data[r0:r0 + r] = y
        m = torch.rand(r, n, device=device, generator=g) >= 0.02
        m[:, :32] = False
        m[:, n - 32:] = False
        blk = torch.randint(0, 257, (r, 1), device=device, generator=g)
        start = torch.randint(0, n, (r, 1), device=device, generator=g)
        ci = torch.arange(n, device=device)[None, :]
        m &= ~((ci >= start) & (ci < start + blk))
        mask[r0:r0 + r] = m
        del y, m
    return data, mask


def generate_on_host(rows: int, n: int, seed: int):
    from sakura_b200 import synth
    return synth.make_spectra(rows, n, seed=seed)


# ----------------------------------------------------------------------------- reference arm
def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle.refbind import Oracle
    try:
        ora = Oracle()
    except OSError as exc:
        print(json.dumps({"impl": "reference", "unavailable": str(exc).splitlines()[0]}))
        return
    n, K = args.nchan, args.order + 1
    rows = min(args.cpu_rows, args.rows)
    data, mask = generate_on_host(rows, n, seed=20240611)
    # torchrun exports OMP_NUM_THREADS=1; the CPU arm must use every host core it can get
    threads = host_cores()
    for _ in range(max(args.warmup, 1)):
        ora.lsqfit_polynomial_batch(args.order, data[:4096], mask[:4096], args.clip, args.nfit, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ora.lsqfit_polynomial_batch(args.order, data, mask, args.clip, args.nfit, nthreads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    value = rows / dt
    sample = (f"{rows} spectra x {n} ch of the same synthetic workload per step, "
              f"OpenMP over rows, one context per thread")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "impl": "reference",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": workload_config(args),
        "sample_spectra_per_step": rows,
        "values": {"num_fitting_max_%d (configs[1], headline)" % args.nfit: value},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": ora.kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args) -> dict:
    """The same dictionary in both arms (the driver compares them): the workload is BASELINE configs[1];
    how many of its spectra a step of the CPU arm samples is reported outside `config`."""
    return {
        "workload": ("BASELINE configs[1]: ALMA-TP-like spectra x 4096 ch float32 + bool mask, "
                     "sakura_LSQFitPolynomialFloat semantics, polynomial order 5, 5 clip iterations "
                     "(num_fitting_max=5) at 3 sigma; outputs coeff+residual+final_mask+rms+status. The metric "
                     "string is BASELINE.json's (it says 3 clip iterations); `value` is this 5-iteration "
                     "config, the 3-iteration figure is `values.num_fitting_max_3`"),
        "spectra_per_gpu": args.rows,
        "num_data": args.nchan, "order": args.order, "num_fitting_max": args.nfit,
        "clip_threshold_sigma": args.clip,
        "l2": "inputs (20 GiB per step) far exceed the 126 MB L2; no explicit flush needed",
        "parallelism": "rows sharded per GPU, no collective",
    }


def side_configs(lib, torch, device, host=None, rows: int = 1 << 18, n: int = 8192, reps: int = 2) -> dict:
    """BASELINE configs[2] and [3] at their full size (262 144 x 8192 ch, 20 pieces / wave numbers 0..20,
    3 fits): device-resident spectra/s of the cubic-spline and sinusoid fits, and -- through the
    host-pointer C ABI with pinned host buffers, copies inside the timed region -- their e2e figure on
    as many rows as the pinned buffers of the headline e2e hold (`host`: dict of pinned tensors).
    Not bench lines of their own; reported under "also"."""
    out = {}
    data, mask = generate_on_device(torch, rows, n, seed=20240611 + 7, device=device, ripple=True)
    res = torch.empty(rows, n, dtype=torch.float32, device=device)
    fm = torch.empty(rows, n, dtype=torch.bool, device=device)
    rms = torch.empty(rows, dtype=torch.float32, device=device)
    st = torch.empty(rows, dtype=torch.int32, device=device)
    ls = torch.empty(rows, dtype=torch.int32, device=device)
    stream = torch.cuda.current_stream(device)

    def timeit(fn):
        assert fn() == 0, lib.last_error()
        torch.cuda.synchronize(device)
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize(device)
        return e0.elapsed_time(e1) / reps

    def host_views(kind, ncoeff_doubles):
        """Views of the pinned e2e buffers as rows of n channels, filled from the device workload."""
        if host is None:
            return None
        er = min(host["data"].numel() // n, rows)
        hv = {"rows": er,
              "data": host["data"].view(-1)[:er * n].view(er, n), "mask": host["mask"].view(-1)[:er * n].view(er, n),
              "resid": host["resid"].view(-1)[:er * n].view(er, n), "fmask": host["fmask"].view(-1)[:er * n].view(er, n),
              "rms": torch.empty(er, dtype=torch.float32, pin_memory=True),
              "status": torch.empty(er, dtype=torch.int32, pin_memory=True),
              "lsq": torch.empty(er, dtype=torch.int32, pin_memory=True),
              "coeff": torch.empty(er, ncoeff_doubles, dtype=torch.float64, pin_memory=True)}
        hv["data"].copy_(data[:er])
        hv["mask"].copy_(mask[:er])
        torch.cuda.synchronize(device)
        return hv

    def time_host(fn, hv, d2h_extra):
        assert fn() == 0, lib.last_error()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        dt = (time.perf_counter() - t0) / reps
        er = hv["rows"]
        same = bool(torch.equal(hv["fmask"], fm[:er].cpu())) and bool(torch.equal(hv["status"], st[:er].cpu()))
        packed = n % 32 == 0 and os.environ.get("SAKURA_B200_PACK_MASKS", "1")[:1] != "0"
        mb = n // 8 if packed else n       # masks over PCIe as bits (sakura_b200/csrc/host_bits.cc)
        return {"value": er / dt, "unit": UNIT, "spectra_per_step": er, "ms_per_step": dt * 1e3,
                "h2d_bytes_per_step": er * (n * 4 + mb), "d2h_bytes_per_step": er * (n * 4 + mb + d2h_extra),
                "final_mask_and_status_match_device_path": same}

    peak_hbm = measured_peak_gbs()[0]
    # C3: cubic spline, 20 pieces, 3 fits
    npiece = 20
    s, ctx = lib.create_cubicspline_context(npiece, n)
    co = torch.empty(rows, npiece, 4, dtype=torch.float64, device=device)
    bd = torch.empty(rows, npiece + 1, dtype=torch.int64, device=device)
    ms = timeit(lambda: lib.lib.sakura_LSQFitCubicSplineFloatBatchDevice(
        ctx, npiece, rows, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, 3, co.data_ptr(), None,
        res.data_ptr(), fm.data_ptr(), rms.data_ptr(), bd.data_ptr(), st.data_ptr(), ls.data_ptr(),
        stream.cuda_stream))
    bps = n * 10 + 8 * 4 * npiece + 8 * (npiece + 1) + 8
    out["config_c3_cubic_spline"] = {
        "workload": f"BASELINE configs[2]: {rows} x {n} ch, 20 pieces, 3 fits, residual out", "value": rows / (ms * 1e-3),
        "unit": UNIT, "ms": ms, "kernel": lib.last_kernel_name(), "rows_ok": int((st == 0).sum().item()),
        "roofline": {"bound": "hbm", "bytes_per_spectrum": bps,
                     "achieved": rows * bps / (ms * 1e-3) / 1e9, "peak": peak_hbm, "unit": "GB/s",
                     "frac": rows * bps / (ms * 1e-3) / 1e9 / peak_hbm}}
    hv = host_views("spline", npiece * 4)
    if hv is not None:
        h_bd = torch.empty(hv["rows"], npiece + 1, dtype=torch.int64, pin_memory=True)
        out["config_c3_cubic_spline"]["e2e"] = time_host(lambda: lib.lib.sakura_LSQFitCubicSplineFloatBatch(
            ctx, npiece, hv["rows"], n, hv["data"].data_ptr(), n, hv["mask"].data_ptr(), n, 3.0, 3,
            hv["coeff"].data_ptr(), None, hv["resid"].data_ptr(), hv["fmask"].data_ptr(), hv["rms"].data_ptr(),
            h_bd.data_ptr(), hv["status"].data_ptr(), hv["lsq"].data_ptr()), hv, 8 * 4 * npiece + 8 * (npiece + 1) + 12)
    lib.destroy_context(ctx)
    # C4: sinusoid, wave numbers 0..20 (41 bases), 3 fits
    s, ctx = lib.create_sinusoid_context(20, n)
    nw = np.arange(21, dtype=np.uint64)
    co = torch.empty(rows, 41, dtype=torch.float64, device=device)
    ms = timeit(lambda: lib.lib.sakura_LSQFitSinusoidFloatBatchDevice(
        ctx, 21, nw.ctypes.data, rows, n, data.data_ptr(), n, mask.data_ptr(), n, 3.0, 3, 41,
        co.data_ptr(), None, res.data_ptr(), fm.data_ptr(), rms.data_ptr(), st.data_ptr(),
        ls.data_ptr(), stream.cuda_stream))
    # FP64 tensor-core work of the kernel: data moments n x 48, model n x 48 per fit (the masked trig sums
    # of the Gram matrix no longer are a contraction: "all channels minus flagged", lsqfit_sinusoid.cu)
    flop = 2.0 * n * (48 + 3 * 48)
    sm_mhz = measured_peaks().get("sm_max_mhz", 1965.0)
    peak_tf = 148 * 64 * 2 * sm_mhz * 1e6 / 1e12
    bps4 = n * 10 + 8 * 41 + 8
    out["config_c4_sinusoid"] = {
        "workload": f"BASELINE configs[3]: {rows} x {n} ch, nwave 0..20 (41 bases), 3 fits, residual out",
        "value": rows / (ms * 1e-3), "unit": UNIT, "ms": ms, "kernel": lib.last_kernel_name(),
        "rows_ok": int((st == 0).sum().item()),
        "roofline": {"bound": "tensor", "flop_per_spectrum": flop,
                     "achieved": rows * flop / (ms * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": rows * flop / (ms * 1e-3) / 1e12 / peak_tf,
                     "peak_source": "FP64 DMMA/DFMA issue rate measured on this GPU (64 FMA/clk/SM, "
                                    "tools/microbench.cu; ncu dmma sub-pipe) x 148 SMs x max SM clock",
                     "hbm_bytes_per_spectrum": bps4,
                     "hbm_frac": rows * bps4 / (ms * 1e-3) / 1e9 / peak_hbm}}
    hv = host_views("sinusoid", 41)
    if hv is not None:
        out["config_c4_sinusoid"]["e2e"] = time_host(lambda: lib.lib.sakura_LSQFitSinusoidFloatBatch(
            ctx, 21, nw.ctypes.data, hv["rows"], n, hv["data"].data_ptr(), n, hv["mask"].data_ptr(), n, 3.0, 3, 41,
            hv["coeff"].data_ptr(), None, hv["resid"].data_ptr(), hv["fmask"].data_ptr(), hv["rms"].data_ptr(),
            hv["status"].data_ptr(), hv["lsq"].data_ptr()), hv, 8 * 41 + 12)
    lib.destroy_context(ctx)
    return out


def chain_config(lib, torch, device, rows: int = 1 << 19, n: int = 2048, reps: int = 3) -> dict:
    """BASELINE configs[4] (the e2eReduce chain, bench/e2eReduce/src/main.cc:80-240) on device arrays
    at its per-GPU size (4 M spectra over 8 GPUs = 524 288 per rank): linear Tsys interpolation to one row per spectrum, uint8 flags -> mask with
    edge flags, calibration fused into the poly-5 fit (NaN/Inf flags merged, 5 fits), clipping of the
    residual merged into the mask, statistics. One timed pass = all five steps back to back."""
    from sakura_b200.device_api import DeviceFits
    dev = DeviceFits(lib)
    sky, mask0 = generate_on_device(torch, rows, n, seed=20240611 + 11, device=device)
    off = 300.0 + 20.0 * torch.randn(rows, n, device=device)
    on = off * (1.0 + sky / 250.0)
    flags = (~mask0).to(torch.uint8) * 4
    del sky, mask0
    t_base = torch.linspace(0.0, 100.0, 6, dtype=torch.float64, device=device)
    tsys_base = (150.0 + 10.0 * torch.randn(6, n, device=device)).float()
    tsys_valid = torch.rand(6, n, device=device) < 0.98
    t_rows = torch.sort(torch.rand(rows, dtype=torch.float64, device=device) * 110.0 - 5.0).values
    chan = torch.arange(n, dtype=torch.int32, device=device).repeat(rows, 1)
    st, ctx = lib.create_polynomial_context(5, n)
    tsys = torch.empty(rows, n, dtype=torch.float32, device=device)
    tsys_m = torch.empty(rows, n, dtype=torch.bool, device=device)
    mask = torch.empty(rows, n, dtype=torch.bool, device=device)
    clipm = torch.empty(rows, n, dtype=torch.bool, device=device)
    res = dev.polynomial(ctx, 5, on, flags == 0, 3.0, 5)
    res.pop("call_status")
    kernels = []

    def chain(record=False):
        dev.interpolate_y(1, t_base, tsys_base, tsys_valid, t_rows, out=tsys, out_mask=tsys_m)
        record and kernels.append(lib.last_kernel_name())
        dev.uint8_to_bool(flags, invert=True, out=mask)
        dev.in_ranges(chan, [31], [n - 32], inclusive=False, and_with=mask, out=mask)
        record and kernels.append(lib.last_kernel_name())
        dev.calibrate_and_polynomial(ctx, 5, tsys, on, off, mask, 3.0, 5, flag_nan_or_inf=True, out=res)
        record and kernels.append(lib.last_kernel_name())
        dev.in_ranges(res["residual"], [-5.0], [5.0], inclusive=False, and_with=res["final_mask"], out=clipm)
        stats = dev.statistics(res["residual"], clipm)
        record and kernels.append(lib.last_kernel_name())
        return stats

    chain(record=True)
    chain()
    torch.cuda.synchronize(device)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        chain()
    e1.record()
    torch.cuda.synchronize(device)
    ms = e0.elapsed_time(e1) / reps
    rows_ok = int((res["status"] == 0).sum().item())
    lib.destroy_context(ctx)
    bps = n * (4 + 1 + 1 + 4) + 48 + 8 + 48  # SURVEY 8(d), C5 fused figure
    return {"workload": f"{rows} x {n} ch: Tsys interpolation + mask builders + calibrate&fit (poly-5, 5 fits) "
                        "+ residual clipping + statistics, device arrays",
            "value": rows / (ms * 1e-3), "unit": UNIT, "ms": ms, "rows": rows, "kernels": kernels, "rows_ok": rows_ok,
            "roofline": {"bound": "hbm", "bytes_per_spectrum": bps,
                         "achieved": rows * bps / (ms * 1e-3) / 1e9, "peak": measured_peak_gbs()[0],
                         "unit": "GB/s", "frac": rows * bps / (ms * 1e-3) / 1e9 / measured_peak_gbs()[0]}}


# ----------------------------------------------------------------------------- our arm
def run_ours(args) -> None:
    import torch
    from sakura_b200.capi import SakuraLib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout when the first communicator is created: send
        # file descriptor 1 to stderr meanwhile, so that stdout carries the one JSON line only
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist_mod.init_process_group("nccl", device_id=device)
            dist_mod.barrier()
            torch.cuda.synchronize(device)
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
        dist = dist_mod

    lib = SakuraLib()
    n, order, K = args.nchan, args.order, args.order + 1
    rows = args.rows
    data, mask = generate_on_device(torch, rows, n, seed=20240611 + rank, device=device)
    coeff = torch.empty(rows, K, dtype=torch.float64, device=device)
    resid = torch.empty(rows, n, dtype=torch.float32, device=device)
    fmask = torch.empty(rows, n, dtype=torch.bool, device=device)
    rms = torch.empty(rows, dtype=torch.float32, device=device)
    status = torch.empty(rows, dtype=torch.int32, device=device)
    lsq = torch.empty(rows, dtype=torch.int32, device=device)
    st, ctx = lib.create_polynomial_context(order, n)
    assert st == 0
    stream = torch.cuda.current_stream(device)

    def step(nfit):
        rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(
            ctx, order, rows, n, data.data_ptr(), n, mask.data_ptr(), n, args.clip, nfit, K,
            coeff.data_ptr(), None, resid.data_ptr(), fmask.data_ptr(), rms.data_ptr(),
            status.data_ptr(), lsq.data_ptr(), stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(f"fit call failed: {rc} {lib.last_error()}")

    def step_no_residual(nfit):
        rc = lib.lib.sakura_LSQFitPolynomialFloatBatchDevice(
            ctx, order, rows, n, data.data_ptr(), n, mask.data_ptr(), n, args.clip, nfit, K,
            coeff.data_ptr(), None, None, fmask.data_ptr(), rms.data_ptr(),
            status.data_ptr(), lsq.data_ptr(), stream.cuda_stream)
        if rc != 0:
            raise RuntimeError(f"fit call failed: {rc} {lib.last_error()}")

    def barrier():
        torch.cuda.synchronize(device)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)

    def timed(nfit, steps, warmup, step=step):
        for _ in range(warmup):
            step(nfit)
        barrier()
        launches0 = lib.kernel_launch_count()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            step(nfit)
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        launches = lib.kernel_launch_count() - launches0
        if dist is not None:
            t = torch.tensor([ms], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_step, launches = timed(args.nfit, args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if rank == 0 else None
    kernel_name = lib.last_kernel_name()
    # rows the screening kernel did not decide itself and handed to the per-channel kernel (second
    # launch of the same step, inside the timed region)
    handed_over = lib.last_handover_count(ctx)
    ok_rows = int((status == 0).sum().item())
    mean_clipped = float((mask & ~fmask).sum(dim=1).float().mean().item())
    # snapshot of the headline run's results for the parity check against the CPU sample
    snap_rows = min(args.cpu_rows, rows)
    snap_fmask = fmask[:snap_rows].cpu().numpy().copy()
    snap_status = status[:snap_rows].cpu().numpy().copy()
    snap_rms = rms[:snap_rows].cpu().numpy().copy()

    also = {}
    if not args.no_also and args.nfit != 3:
        ms3, _ = timed(3, max(args.steps // 2, 2), 3)
        also["num_fitting_max_3"] = {
            "value": world * rows / (ms3 * 1e-3), "unit": UNIT, "ms_per_step": ms3,
            "hbm_frac": rows * bytes_per_spectrum(n, K) / (ms3 * 1e-3) / 1e9 / measured_peak_gbs()[0],
            "per_iteration_streaming": streaming_equivalent(n, 3, rows, ms3)}
    if not args.no_also:
        # the "fit + mask" mode, no residual written (24 632 algorithmic B/spectrum)
        ms_nr, _ = timed(args.nfit, max(args.steps // 3, 2), 1, step=step_no_residual)
        also["fit_and_mask_only"] = {
            "value": world * rows / (ms_nr * 1e-3), "unit": UNIT, "ms_per_step": ms_nr,
            "bytes_per_spectrum": bytes_per_spectrum(n, K, residual=False),
            "hbm_frac": rows * bytes_per_spectrum(n, K, residual=False) / (ms_nr * 1e-3) / 1e9
            / measured_peak_gbs()[0]}
    if not args.no_also and world == 1:
        # the same workload with every row through the per-channel kernel (screening off), and a
        # bit-for-bit comparison of its clip decisions with the headline run's
        step(args.nfit)
        torch.cuda.synchronize(device)
        fm_screen = fmask.clone()
        os.environ["SAKURA_B200_POLY_SCREEN"] = "0"
        try:
            ms_pc, _ = timed(args.nfit, max(args.steps // 3, 2), 1)
            also["per_channel_kernel_only"] = {
                "value": rows / (ms_pc * 1e-3), "unit": UNIT, "ms_per_step": ms_pc,
                "kernel": lib.last_kernel_name(),
                "final_mask_mismatches_vs_headline": int((fm_screen != fmask).sum().item())}
        finally:
            del os.environ["SAKURA_B200_POLY_SCREEN"]
            del fm_screen

    # ---- e2e: host buffers through the C ABI, copies inside the timed region --------------
    e2e = None
    host = None
    if not args.no_e2e:
        er = min(args.e2e_rows, rows)
        # page-locked staging on the NUMA node of this rank's GPU (see NearGpu)
        with NearGpu(torch, local_rank) as near:
            h_data = torch.empty(er, n, dtype=torch.float32, pin_memory=True)
            h_mask = torch.empty(er, n, dtype=torch.bool, pin_memory=True)
            h_coeff = torch.empty(er, K, dtype=torch.float64, pin_memory=True)
            h_resid = torch.empty(er, n, dtype=torch.float32, pin_memory=True)
            h_fmask = torch.empty(er, n, dtype=torch.bool, pin_memory=True)
            h_rms = torch.empty(er, dtype=torch.float32, pin_memory=True)
            h_status = torch.empty(er, dtype=torch.int32, pin_memory=True)
            h_lsq = torch.empty(er, dtype=torch.int32, pin_memory=True)
        h_data.copy_(data[:er])
        h_mask.copy_(mask[:er])
        host = {"data": h_data, "mask": h_mask, "resid": h_resid, "fmask": h_fmask}
        st2, ctx2 = lib.create_polynomial_context(order, n)
        assert st2 == 0
        import ctypes
        dev_ids = (ctypes.c_int * 1)(local_rank)

        def e2e_step():
            # the library's own sharding entry point, here with the one GPU this rank owns (a
            # single-process caller passes every GPU of the box: tools/bench_multidevice.py)
            rc = lib.lib.sakura_LSQFitPolynomialFloatBatchMultiDevice(
                ctx2, order, er, n, h_data.data_ptr(), n, h_mask.data_ptr(), n, args.clip, args.nfit,
                K, h_coeff.data_ptr(), None, h_resid.data_ptr(), h_fmask.data_ptr(), h_rms.data_ptr(),
                h_status.data_ptr(), h_lsq.data_ptr(), 1, ctypes.cast(dev_ids, ctypes.c_void_p))
            if rc != 0:
                raise RuntimeError(f"e2e fit call failed: {rc} {lib.last_error()}")

        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        torch.cuda.synchronize(device)
        dt = (time.perf_counter() - t0) / args.e2e_steps
        if dist is not None:
            t = torch.tensor([dt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # copy-only ceiling of the box for the same step: the same bytes up and down on the same pinned
        # buffers with no kernel, every rank at once (what the host feed allows at this GPU count)
        # (the pipeline's own pattern: 96 MiB chunks through four slots, one stream per direction and slot)
        crow = 24576
        slots = [dict(data=torch.empty(crow, n, dtype=torch.float32, device=device),
                      mask=torch.empty(crow, n, dtype=torch.bool, device=device),
                      resid=torch.empty(crow, n, dtype=torch.float32, device=device),
                      fmask=torch.empty(crow, n, dtype=torch.bool, device=device),
                      up=torch.cuda.Stream(device), down=torch.cuda.Stream(device)) for _ in range(4)]

        def copy_only():
            for ci, r0 in enumerate(range(0, er, crow)):
                sl = slots[ci % 4]
                r1 = min(er, r0 + crow)
                with torch.cuda.stream(sl["up"]):
                    sl["data"][:r1 - r0].copy_(h_data[r0:r1], non_blocking=True)
                    sl["mask"][:r1 - r0].copy_(h_mask[r0:r1], non_blocking=True)
                with torch.cuda.stream(sl["down"]):
                    h_resid[r0:r1].copy_(sl["resid"][:r1 - r0], non_blocking=True)
                    h_fmask[r0:r1].copy_(sl["fmask"][:r1 - r0], non_blocking=True)
            torch.cuda.synchronize(device)

        copy_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(2):
            copy_only()
        cdt = (time.perf_counter() - t0) / 2
        if dist is not None:
            t = torch.tensor([cdt], dtype=torch.float64, device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            cdt = float(t.item())
        del slots
        e2e_step()  # the probe overwrote the result buffers: put the fit's results back
        torch.cuda.synchronize(device)
        # the library moves the boolean masks over PCIe as one bit per channel (host codec on worker threads,
        # sakura_b200/csrc/host_bits.cc) unless SAKURA_B200_PACK_MASKS=0: count what is actually copied
        packed = n % 32 == 0 and os.environ.get("SAKURA_B200_PACK_MASKS", "1")[:1] != "0"
        mask_bytes = n // 8 if packed else n
        e2e = {"value": world * er / dt, "unit": UNIT,
               "h2d_bytes_per_step": er * (n * 4 + mask_bytes),
               "d2h_bytes_per_step": er * (n * 4 + mask_bytes + 8 * K + 4 + 4 + 4 + 4),
               "spectra_per_step": er, "ms_per_step": dt * 1e3,
               "api": "sakura_LSQFitPolynomialFloatBatchMultiDevice (pinned host buffers, per GPU a 4-slot "
                      "upload | kernel | download pipeline" + (", masks bit-packed by host worker threads)"
                                                                if packed else ")"),
               "masks_bit_packed": packed,
               "pcie_gbs_per_direction_per_gpu": er * (n * 4 + mask_bytes) / dt / 1e9,
               "copy_only_ceiling": {"spectra_per_s": world * er / cdt, "ms_per_step": cdt * 1e3,
                                     "gbs_per_direction_per_gpu": er * n * 5 / cdt / 1e9,
                                     "what": "the caller's arrays as they are (5 bytes per channel each way) "
                                             "copied up and down concurrently on every rank, no kernel"},
               "frac_of_copy_only_ceiling": cdt / dt,
               "pinned_buffers_numa_node": near.node,
               "status_match_device_path": bool(np.array_equal(h_status.numpy()[:snap_rows],
                                                                 snap_status[:er])),
               "final_mask_match_device_path": bool(np.array_equal(h_fmask.numpy()[:snap_rows],
                                                                     snap_fmask[:er]))}
        lib.destroy_context(ctx2)

    # ---- the other BASELINE configs (side figures) ------------------------------------------
    if not args.no_also:
        if world == 1:
            try:
                also.update(side_configs(lib, torch, device, host=host))
            except Exception as exc:  # the headline line must not depend on the side configs
                also["side_configs_error"] = repr(exc)
        try:
            # configs[4] is defined over 8 GPUs: every rank runs its 524 288-row share
            c5 = chain_config(lib, torch, device)
            if dist is not None:
                t = torch.tensor([c5["ms"]], dtype=torch.float64, device=device)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                c5["ms"] = float(t.item())
                c5["value"] = world * c5["rows"] / (c5["ms"] * 1e-3)
                c5["workload"] += f" (x {world} ranks, max over ranks)"
            also["config_c5_chain"] = c5
        except Exception as exc:
            also["config_c5_error"] = repr(exc)
    host = None
    if not args.no_e2e:
        del h_data, h_mask, h_resid, h_fmask

    # ---- CPU baseline: the reference's own code on this box's host cores ---------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            from oracle.refbind import Oracle
            ora = Oracle()
            cr = min(args.cpu_rows, rows)
            cd = data[:cr].cpu().numpy()
            cm = mask[:cr].cpu().numpy()
            ncores = host_cores()
            ora.lsqfit_polynomial_batch(order, cd[:2048], cm[:2048], args.clip, args.nfit, nthreads=ncores)
            best = None
            for _ in range(2):
                t0 = time.perf_counter()
                o = ora.lsqfit_polynomial_batch(order, cd, cm, args.clip, args.nfit, nthreads=ncores)
                dt = time.perf_counter() - t0
                best = dt if best is None else min(best, dt)
            good = o["status"] == 0
            cpu = {"value": cr / best, "unit": UNIT, "cores": ncores, "kind": ora.kind,
                   "sample": f"first {cr} spectra of the same workload, OpenMP over rows, best of 2",
                   "final_mask_mismatches_vs_gpu": int((snap_fmask[good] != o["final_mask"][good]).sum()),
                   "channels_compared": int(good.sum()) * n,
                   "rms_max_rel_diff_vs_gpu": float(np.max(np.abs(snap_rms[good] - o["rms"][good])
                                                           / np.abs(o["rms"][good]))),
                   "status_equal": bool(np.array_equal(o["status"], snap_status))}
        except OSError as exc:
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": str(exc)}

    if rank == 0:
        peak, peak_src = measured_peak_gbs()
        bps = bytes_per_spectrum(n, K)
        achieved = rows * bps / (ms_step * 1e-3) / 1e9
        # DRAM bytes of one launch: the ncu capture (profiles/) measured bytes per spectrum on a
        # 131072-row launch of the same kernel and shape; scaled to this launch's row count
        traffic = None
        traffic_src = None
        tpath = os.path.join(HERE, "profiles", "traffic_per_launch.json")
        if os.path.exists(tpath) and n == 4096 and K == 6:
            try:
                with open(tpath) as fh:
                    t = json.load(fh)
                traffic = float(t["dram_bytes_per_spectrum"]) * rows
                traffic_src = t.get("source")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": world * rows / (ms_step * 1e-3), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args),
            "values": {"num_fitting_max_%d (configs[1], headline)" % args.nfit: world * rows / (ms_step * 1e-3),
                       "num_fitting_max_3 (the metric string's iteration count)":
                           also.get("num_fitting_max_3", {}).get("value")},
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "kernel": kernel_name, "bytes_per_spectrum": bps,
                         "spectra_per_launch": rows,
                         "per_iteration_streaming": streaming_equivalent(n, args.nfit, rows, ms_step)},
            "cpu_baseline": cpu,
            "also": also,
            "check": {"rows_ok": ok_rows, "rows": rows, "mean_clipped_channels": mean_clipped,
                      "rows_handed_to_per_channel_kernel": handed_over},
        }
        print(json.dumps(line))
    lib.destroy_context(ctx)
    if dist is not None:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
