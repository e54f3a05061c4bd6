"""One process, one C-ABI call, G GPUs: sakura_LSQFitPolynomialFloatBatchMultiDevice on a batch in
pinned host memory, next to the copy-only ceiling of the same box (the same bytes moved up and down
concurrently on the same G GPUs with no kernel at all) -- VERDICT r01 item 2.

    python tools/bench_multidevice.py [--rows-per-gpu 65536] [--nchan 4096] [--numa {none,shard}]

`--numa shard` places the rows of each device's block on the NUMA node that device hangs off (mbind
on an anonymous mapping, then cudaHostRegister); `none` lets one thread allocate everything
(cudaHostAlloc through torch: pages land where that thread runs)."""
import argparse
import ctypes as C
import json
import mmap
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, HERE)
import bench  # noqa: E402
from sakura_b200.capi import SakuraLib  # noqa: E402

libc = C.CDLL(None, use_errno=True)
SYS_MBIND = 237  # x86-64
MPOL_BIND = 2


def mbind(addr: int, length: int, node: int) -> bool:
    mask = C.c_ulong(1 << node)
    rc = libc.syscall(SYS_MBIND, C.c_void_p(addr), C.c_ulong(length), C.c_int(MPOL_BIND), C.byref(mask),
                      C.c_ulong(64), C.c_uint(0))
    return rc == 0


class HostArena:
    """Page-aligned anonymous memory, optionally bound per row block to NUMA nodes, page-locked for CUDA."""

    def __init__(self, nbytes: int):
        self.nbytes = (nbytes + 4095) // 4096 * 4096
        self.map = mmap.mmap(-1, self.nbytes, flags=mmap.MAP_PRIVATE | mmap.MAP_ANONYMOUS)
        self.arr = np.frombuffer(self.map, dtype=np.uint8)
        self.addr = self.arr.ctypes.data
        self.registered = False

    def bind(self, offset: int, length: int, node: int) -> bool:
        lo = offset // 4096 * 4096
        hi = (offset + length + 4095) // 4096 * 4096
        return mbind(self.addr + lo, min(hi, self.nbytes) - lo, node)

    def pin(self):
        self.arr[::4096] = 0  # first touch under the policy set by bind()
        rt = torch.cuda.cudart()
        rc = rt.cudaHostRegister(self.addr, self.nbytes, 1)  # cudaHostRegisterPortable
        assert int(rc) == 0, rc
        self.registered = True

    def close(self):
        if self.registered:
            torch.cuda.cudart().cudaHostUnregister(self.addr)
        del self.arr
        try:
            self.map.close()
        except BufferError:  # a view is still alive somewhere: the mapping goes with the process
            pass


def copy_ceiling(G, views, reps=3):
    """Duplex copy of every device's block, all devices at once, nothing else: GB/s per direction (sum)."""
    devs = []
    for g in range(G):
        with torch.cuda.device(g):
            up, down = torch.cuda.Stream(g), torch.cuda.Stream(g)
            d_in = torch.empty(views[g]["in"].numel(), dtype=torch.uint8, device=f"cuda:{g}")
            d_out = torch.empty(views[g]["out"].numel(), dtype=torch.uint8, device=f"cuda:{g}")
            devs.append((up, down, d_in, d_out))

    def once():
        for g in range(G):
            up, down, d_in, d_out = devs[g]
            with torch.cuda.stream(up):
                d_in.copy_(views[g]["in"], non_blocking=True)
            with torch.cuda.stream(down):
                views[g]["out"].copy_(d_out, non_blocking=True)
        for g in range(G):
            torch.cuda.synchronize(g)

    once()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    dt = (time.perf_counter() - t0) / reps
    nbytes = sum(v["in"].numel() for v in views)
    return nbytes / dt / 1e9, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-per-gpu", type=int, default=65536)
    ap.add_argument("--nchan", type=int, default=4096)
    ap.add_argument("--nfit", type=int, default=5)
    ap.add_argument("--numa", default="both", choices=["none", "shard", "both"])
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    n, K = args.nchan, 6
    visible = torch.cuda.device_count()
    lib = SakuraLib()
    report = {"visible_gpus": visible, "host_cores": len(os.sched_getaffinity(0)),
              "gpu_numa_nodes": [bench.gpu_numa_cpus(torch, g)[0] for g in range(visible)], "runs": []}
    try:
        report["topology"] = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True).stdout
    except Exception:
        pass
    modes = ["none", "shard"] if args.numa == "both" else [args.numa]
    if all(node is None for node in report["gpu_numa_nodes"]) and "none" in modes:
        modes = ["none"]  # the box does not expose NUMA nodes: nothing to place
    for numa in modes:
        for G in [g for g in (1, 2, 4, 8) if g <= visible]:
            R = args.rows_per_gpu * G
            sizes = {"data": R * n * 4, "mask": R * n, "resid": R * n * 4, "fmask": R * n, "coeff": R * K * 8,
                     "rms": R * 4, "status": R * 4, "lsq": R * 4}
            arenas = {k: HostArena(v) for k, v in sizes.items()}
            per_row = {"data": n * 4, "mask": n, "resid": n * 4, "fmask": n, "coeff": K * 8, "rms": 4, "status": 4, "lsq": 4}
            bound = True
            if numa == "shard":
                for g in range(G):
                    node = report["gpu_numa_nodes"][g]
                    if node is None:
                        bound = False
                        continue
                    r0, r1 = R * g // G, R * (g + 1) // G
                    for k, a in arenas.items():
                        bound = a.bind(r0 * per_row[k], (r1 - r0) * per_row[k], node) and bound
            for a in arenas.values():
                a.pin()
            # synthetic workload generated on GPU 0 block by block
            h_data = torch.from_numpy(arenas["data"].arr[:sizes["data"]].view(np.float32).reshape(R, n))
            h_mask = torch.from_numpy(arenas["mask"].arr[:sizes["mask"]].view(np.bool_).reshape(R, n))
            for r0 in range(0, R, 32768):
                d, m = bench.generate_on_device(torch, min(32768, R - r0), n, seed=5 + r0, device=torch.device("cuda", 0))
                h_data[r0:r0 + d.shape[0]].copy_(d)
                h_mask[r0:r0 + d.shape[0]].copy_(m)
            torch.cuda.synchronize(0)
            st, ctx = lib.create_polynomial_context(5, n)
            assert st == 0
            ids = (C.c_int * G)(*range(G))

            def call():
                rc = lib.lib.sakura_LSQFitPolynomialFloatBatchMultiDevice(
                    ctx, 5, R, n, arenas["data"].addr, n, arenas["mask"].addr, n, 3.0, args.nfit, K,
                    arenas["coeff"].addr, None, arenas["resid"].addr, arenas["fmask"].addr, arenas["rms"].addr,
                    arenas["status"].addr, arenas["lsq"].addr, G, C.cast(ids, C.c_void_p))
                assert rc == 0, lib.last_error()

            call()
            t0 = time.perf_counter()
            for _ in range(args.reps):
                call()
            dt = (time.perf_counter() - t0) / args.reps
            status = arenas["status"].arr[:sizes["status"]].view(np.int32)
            # copy-only ceiling on the same buffers, same split
            views = []
            for g in range(G):
                r0, r1 = R * g // G, R * (g + 1) // G
                vin = torch.from_numpy(arenas["data"].arr[r0 * n * 4:r1 * n * 4])
                vout = torch.from_numpy(arenas["resid"].arr[r0 * n * 4:r1 * n * 4])
                views.append({"in": vin, "out": vout})
            gbs, cdt = copy_ceiling(G, views, args.reps)
            # the fit moves 5 bytes per channel each way where the probe moved 4
            ceiling = gbs * 1e9 / (n * 5)
            report["runs"].append({
                "numa": numa, "numa_bound": bound if numa == "shard" else None, "gpus": G, "spectra": R,
                "call_ms": dt * 1e3, "spectra_per_s": R / dt, "rows_ok": int((status == 0).sum()),
                "gbs_per_direction_in_call": R * n * 5 / dt / 1e9,
                "copy_only_gbs_per_direction": gbs, "copy_only_ceiling_spectra_per_s": ceiling,
                "fraction_of_copy_ceiling": (R / dt) / ceiling})
            print(json.dumps(report["runs"][-1]), file=sys.stderr, flush=True)
            lib.destroy_context(ctx)
            del h_data, h_mask, views, status
            for a in arenas.values():
                a.close()
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
