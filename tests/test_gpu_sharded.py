"""GPU, 2 ranks over NCCL: sharded training (mbpe_train_sharded) gives every rank the reference's merge list.
Needs two GPUs; the single-GPU test box skips it (tools/sharded_check.py runs the same thing under torchrun)."""
import os
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_data, load_package

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, ret):
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    sys.path.insert(0, ROOT)
    from oracle import oracle as O
    pkg = load_package()

    def bcast(ident):
        t = torch.from_numpy(ident.copy()).cuda()
        dist.broadcast(t, 0)
        return t.cpu().numpy()

    comm = pkg.Comm(rank, world, rank, bcast)
    ok = True
    for name in ("ts512_gpt4_first", "ts512_gpt4_lexical", "ts400_basic_first"):
        import json
        e = json.load(open(os.path.join(GOLDEN, "manifest.json")))["train"][name]
        _, _, gm = O.read_model(os.path.join(GOLDEN, "models", name + ".model"))
        text = golden_data(e["input"])
        tok, off, w, _ = pkg.split_dedup(pkg.patterns()[e["encoder"]], text)
        m, c, st = comm.train(tok, off, w, e["vocab_size"], e["mode"])
        ok = ok and m.shape == gm.shape and bool((m == gm).all())
    ret[rank] = ok
    comm.close()
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_training_two_gpus(pkg):
    if pkg.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    world, port = 2, 29600 + os.getpid() % 1000
    with mp.Manager() as mgr:
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert ret[0] and ret[1]
