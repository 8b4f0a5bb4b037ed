"""The N>1 host path on CPU: world_size-2 (and 3) gloo process groups run the package's
sharding logic — contiguous unit ranges per rank, one all-gather of byte/record counts,
pwrite at the exchanged offsets — and the single output file must be byte-identical to the
one-rank output.  No GPU here, so each rank's shard bytes come from the CPU oracle (test
infrastructure); on the GPU box the same logic runs over Engine.simulate (test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
import simmr_b200 as S
from oracle import pyoracle as po
from simmr_b200 import sharding

READS = [2001, 0, 640, 3000, 7]


def _inputs():
    gs = H.synth_genomes(5, (1, 3), (1500, 6000), seed=3)
    prof = S.MinimalShortErrorProfile(150, 150, 30)
    return gs, prof


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _rank_main(rank, world, port, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        gs, prof = _inputs()
        oprof = H.oracle_profile(prof)
        total = po.total_units(oprof, READS)
        ub, ue = sharding.shard_range(total, rank, world)
        data, st = (b"", po.OrcStats()) if ub == ue else po.simulate(po.Genomes.from_set(gs), oprof, READS, 42, unit_begin=ub,
                                                                      unit_end=ue, want_stats=True)
        off, first_rec, all_bytes, all_recs = sharding.exchange_counts(len(data), st.records, dist)
        w = sharding.ShardWriter(out_path, off)
        half = len(data) // 2
        w(np.frombuffer(data[:half], np.uint8))
        w(np.frombuffer(data[half:], np.uint8))
        w.close()
        dist.barrier()
        if rank == 0:
            assert os.path.getsize(out_path) == all_bytes
            assert all_recs == 2 * total
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_output_is_byte_identical(tmp_path, world):
    gs, prof = _inputs()
    want = po.simulate(po.Genomes.from_set(gs), H.oracle_profile(prof), READS, 42)
    out = str(tmp_path / f"reads_w{world}.fq")
    mp.spawn(_rank_main, args=(world, _free_port(), out), nprocs=world, join=True)
    assert open(out, "rb").read() == want


def test_shard_ranges_tile_the_unit_space():
    for total in (0, 1, 7, 1000, 10 ** 9 + 7):
        for world in (1, 2, 3, 8):
            r = [sharding.shard_range(total, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            assert max(e - b for b, e in r) - min(e - b for b, e in r) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)
    assert sharding.exchange_counts(123, 4) == (0, 0, 123, 4)
