"""CPU, world_size 2 over gloo: the N > 1 encode path: every rank takes a contiguous, byte-balanced
range of chunks (mbpe_plan_shards), encodes it with no communication, and the ranks' id streams concatenated in
rank order are the id stream of the whole input. On the CPU box the per-rank compute is the oracle (the checker);
on the GPU box bench.py runs the same plan with the CUDA encoder."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import GOLDEN, ROOT, golden_data, load_package


def _worker(rank, world, port, fname, model, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    pkg = load_package()
    text = golden_data(fname)
    _, _, merges = O.read_model(os.path.join(GOLDEN, "models", model + ".model"))
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    off = np.concatenate([s, e[-1:]])
    part, poff, c0 = pkg.shard_for_rank(text, off, rank, world)
    ids, _ = O.encode_chunks(merges, part, poff[:-1], poff[1:])
    # the only collective of the path: sizes, so rank 0 can lay the parts out (the id streams themselves stay put)
    counts = [None] * world
    dist.all_gather_object(counts, (c0, len(poff) - 1, len(part), len(ids)))
    parts = [None] * world
    dist.gather_object(ids, parts if rank == 0 else None, dst=0)
    if rank == 0:
        whole, _ = O.encode_chunks(merges, text, s, e)
        cat = np.concatenate(parts)
        ret["equal"] = bool(np.array_equal(cat, whole))
        ret["counts"] = counts
        ret["n_chunks"] = len(s)
        ret["n_bytes"] = len(text)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("fname,model", [("taylorswift.txt", "ts512_gpt4_first"),
                                         ("shakespeare.txt", "shk4096_gpt4_lexical_special")])
def test_sharded_encode_two_ranks_gloo(fname, model):
    world, port = 2, 29500 + os.getpid() % 1000
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, fname, model, ret), nprocs=world, join=True)
        assert ret["equal"]
        counts = ret["counts"]
        assert counts[0][0] == 0 and counts[1][0] == counts[0][1]            # contiguous chunk ranges
        assert counts[0][1] + counts[1][1] == ret["n_chunks"]
        assert counts[0][2] + counts[1][2] == ret["n_bytes"]
        assert abs(counts[0][2] - counts[1][2]) < 64                         # balanced by bytes (within one chunk)


def test_plan_shards_edge_cases(pkg):
    off = np.asarray([0, 3, 3, 10, 11, 20], np.uint64)
    assert pkg.plan_shards(off, 1).tolist() == [0, 5]
    p = pkg.plan_shards(off, 3).tolist()
    assert p[0] == 0 and p[-1] == 5 and p == sorted(p)
    assert pkg.plan_shards(np.asarray([0], np.uint64), 4).tolist() == [0, 0, 0, 0, 0]   # nothing to encode
    assert pkg.plan_shards(off, 8).tolist()[-1] == 5                                     # more ranks than chunks
