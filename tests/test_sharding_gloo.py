"""N>1 host logic on CPU: two gloo ranks shard the rows, fit their blocks (the oracle stands in
for the device kernel here -- tests may use it), gather the per-row outputs on rank 0, which
must equal the single-process result; the max-over-ranks timing reduction is exercised too."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, tmpdir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle.refbind import Oracle
    from sakura_b200 import sharding, synth
    R, n, order = 37, 256, 3  # odd on purpose: uneven shards
    data, mask = synth.make_spectra(R, n, seed=99)
    s, e = sharding.shard_rows(R, world, rank)
    ora = Oracle("port")
    r = ora.lsqfit_polynomial_batch(order, data[s:e], mask[s:e], 3.0, 3)
    outs = {"coeff": torch.from_numpy(r["coeff"].copy()), "rms": torch.from_numpy(r["rms"].copy()),
            "status": torch.from_numpy(r["status"].copy())}
    g = sharding.gather_row_outputs(outs, R, dst=0)
    slow = sharding.max_over_ranks(1.0 + rank)
    assert slow == float(world)
    if rank == 0:
        full = ora.lsqfit_polynomial_batch(order, data, mask, 3.0, 3)
        assert np.array_equal(g["coeff"].numpy(), full["coeff"])
        assert np.array_equal(g["rms"].numpy(), full["rms"])
        assert np.array_equal(g["status"].numpy(), full["status"])
        open(os.path.join(tmpdir, "ok"), "w").write("ok")
    dist.barrier()
    dist.destroy_process_group()


def test_shard_rows_cover_exactly():
    from sakura_b200.sharding import shard_rows
    for R in (0, 1, 7, 8, 1000, 1 << 20):
        for G in (1, 2, 3, 4, 8):
            blocks = [shard_rows(R, G, g) for g in range(G)]
            assert blocks[0][0] == 0 and blocks[-1][1] == R
            for (s0, e0), (s1, e1) in zip(blocks, blocks[1:]):
                assert e0 == s1 and s0 <= e0
    with pytest.raises(ValueError):
        shard_rows(10, 2, 2)


def test_two_rank_gloo(tmp_path):
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok").exists()
