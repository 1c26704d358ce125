"""CPU suite for the N > 1 layout: world_size-2/3 gloo processes shard a batch by records, compute their
range (the oracle stands in for the GPU here — this tests the host logic only) and rank 0 reassembles the
table in input order."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

import hx_synth
from hyperex_b200 import shard
from oracle import oracle as O


def test_shard_ranges_cover_in_order():
    rng = np.random.default_rng(0)
    for _ in range(50):
        n = int(rng.integers(0, 40))
        lens = rng.integers(0, 2000, n)
        off = np.zeros(n + 1, np.uint64)
        off[1:] = np.cumsum(lens)
        off += np.uint64(rng.integers(0, 100))
        for world in (1, 2, 3, 8):
            rs = shard.shard_ranges(off, world)
            assert len(rs) == world and rs[0][0] == 0 and rs[-1][1] == n
            assert all(rs[i][1] == rs[i + 1][0] for i in range(world - 1))
            assert all(a <= b for a, b in rs)
            if n >= 8 * world and lens.sum() > 0:
                per = [int(off[b] - off[a]) for a, b in rs]
                assert max(per) - min(per) <= 2 * lens.max() + 1


def _worker(rank, world, port, out_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    a, o = hx_synth.gen_reads_mutated(61, seed=7, max_edits=2, lmin=1, lmax=900)
    pairs = O.default_pairs()
    res = shard.run_sharded(lambda aa, oo: O.run_batch(aa, oo, pairs, 2), a, o, rank, world, len(pairs),
                            O.RESULT_DTYPE)
    if rank == 0:
        np.save(out_path, res)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_run_equals_single_process(tmp_path, world):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "res.npy")
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    a, o = hx_synth.gen_reads_mutated(61, seed=7, max_edits=2, lmin=1, lmax=900)
    ref = O.run_batch(a, o, O.default_pairs(), 2)
    got = np.load(out)
    assert got.tobytes() == ref.tobytes()
