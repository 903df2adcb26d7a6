"""Multi-rank host logic on CPU (gloo, world_size 2 and 3): every rank derives its own contiguous range of
the pair triangle from the sequence lengths alone (pa_partition_by_length, no device needed), fills the
records of its range, and the ranges tile the triangle exactly.  The compute inside each rank is the CUDA
module on the GPU box; here the oracle stands in for it so that the sharding, ordering and gathering
logic is what is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from phylommand_b200 import synth


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_seq, seed, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from phylommand_b200 import capi
        from tests import oracle_lib
        oracle = oracle_lib.load()
        _, seqs, _ = synth.make_its_like(n_seq, seed)
        enc = [synth.to_masks(s[:60 + (k * 7) % 90]) for k, s in enumerate(seqs)]      # short, very ragged
        lens = np.array([len(e) for e in enc], dtype=np.uint32)
        total = n_seq * (n_seq - 1) // 2
        bounds, cells = capi.partition_by_length(lens, 0, total, world)
        first, last = int(bounds[rank]), int(bounds[rank + 1])
        masks, offsets = capi.pack(enc)
        mine = oracle.all_pairs(masks, offsets, first=first, last=last, threads=2)
        # exchange: sizes, then padded records (no collective is needed in the product; this is the test's gather)
        sizes = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([last - first], dtype=torch.int64))
        cap = int(max(int(s) for s in sizes))
        buf = torch.zeros(cap * 20, dtype=torch.uint8)
        buf[: (last - first) * 20] = torch.from_numpy(np.frombuffer(mine.tobytes(), dtype=np.uint8).copy())
        gathered = [torch.zeros(cap * 20, dtype=torch.uint8) for _ in range(world)]
        dist.all_gather(gathered, buf)
        cell_t = torch.tensor([int(cells[rank])], dtype=torch.int64)
        dist.all_reduce(cell_t)
        if rank == 0:
            whole = b"".join(gathered[r][: int(sizes[r]) * 20].numpy().tobytes() for r in range(world))
            want = oracle.all_pairs(masks, offsets, threads=2)
            ok = whole == want.tobytes()
            lens64 = lens.astype(np.int64)
            all_cells = int((lens64.sum() ** 2 - (lens64 ** 2).sum()) // 2)
            balanced = int(cells.max() - cells.min()) <= 2 * int(lens64.max()) ** 2
            with open(os.path.join(out_dir, "result.txt"), "w") as fh:
                fh.write(f"{int(ok)} {int(int(cell_t) == all_cells)} {int(balanced)} {int(sum(int(s) for s in sizes) == total)}")
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n_seq", [(2, 40), (3, 31)])
def test_triangle_sharding_over_ranks(tmp_path, capi, world, n_seq):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_seq, 1004, str(tmp_path)), nprocs=world, join=True)
    flags = (tmp_path / "result.txt").read_text().split()
    assert flags == ["1", "1", "1", "1"], f"records equal / cells add up / balanced / ranges tile: {flags}"


def test_partition_by_length_edge_cases(capi):
    lens = np.array([10, 20, 30, 40, 50], dtype=np.uint32)
    total = 10
    b, c = capi.partition_by_length(lens, 0, total, 1)
    assert b.tolist() == [0, 10] and int(c[0]) == 10 * 140 + 20 * 120 + 30 * 90 + 40 * 50
    b, c = capi.partition_by_length(lens, 0, total, 4)
    assert b[0] == 0 and b[-1] == total and np.all(np.diff(b.astype(np.int64)) >= 0)
    assert int(c.sum()) == 10 * 140 + 20 * 120 + 30 * 90 + 40 * 50
    b, c = capi.partition_by_length(lens, 3, 4, 2)            # a sub-range
    assert b[0] == 3 and b[-1] == 7
    b, c = capi.partition_by_length(lens, 0, total, 16)       # more parts than pairs: empty parts are fine
    assert b[-1] == total and int(c.sum()) > 0
    with pytest.raises(capi.PairalignError):
        capi.partition_by_length(lens, 8, 5, 2)
    b, c = capi.partition_by_length(np.array([5], dtype=np.uint32), 0, 0, 2)   # a single sequence has no pairs
    assert b.tolist() == [0, 0, 0]


def test_partition_by_length_properties(capi):
    """Random length sets, sub-ranges and part counts: the parts tile the range in order, their cells add up to the
    cells of the range (recounted pair by pair), and no part exceeds the ideal share by more than one pair's cells."""
    rng = np.random.default_rng(12)
    for trial in range(200):
        n = int(rng.integers(2, 60))
        lens = rng.integers(0 if trial % 5 == 0 else 1, 3000, size=n).astype(np.uint32)
        total = n * (n - 1) // 2
        first = int(rng.integers(0, total + 1))
        count = int(rng.integers(0, total - first + 1))
        parts = int(rng.integers(1, 12))
        b, c = capi.partition_by_length(lens, first, count, parts)
        b = b.astype(np.int64)
        assert b[0] == first and b[-1] == first + count and np.all(np.diff(b) >= 0)
        pair_cells = np.array([int(lens[x]) * int(lens[y]) for x in range(n) for y in range(x + 1, n)], dtype=np.int64)
        want = [int(pair_cells[b[p]:b[p + 1]].sum()) for p in range(parts)]
        assert [int(v) for v in c] == want
        whole = int(pair_cells[first:first + count].sum())
        biggest = int(pair_cells[first:first + count].max()) if count else 0
        assert max(want) <= whole / parts + biggest + 1
