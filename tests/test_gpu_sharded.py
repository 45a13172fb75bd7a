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
    import json
    for name in ("ts512_gpt4_first", "ts512_gpt4_lexical", "ts400_basic_first", "str_runs_c_gpt4_first",
                 "str_exhaust_basic_lexical", "sample512_gpt4_first"):
        e = json.load(open(os.path.join(GOLDEN, "manifest.json")))["train"][name]
        _, _, gm = O.read_model(os.path.join(GOLDEN, "models", name + ".model"))
        text = golden_data(e["input"])
        tok, off, w, _ = pkg.split_dedup(pkg.patterns()[e["encoder"]], text)
        _, oc = O.train(tok, off, w, e["vocab_size"], e["mode"])
        tr = pkg.ShardedTrainer(comm, tok, off, w)
        # resident engine (one CTA per rank, exchange through the peers' mapped memory) and host-driven engine
        for engine in ("persistent", "stepwise", "persistent"):
            m, c, st = tr.run(e["vocab_size"], e["mode"], engine=engine)
            ok = ok and m.shape == gm.shape and bool((m == gm).all()) and bool((c == oc).all())
        tr.close()
    # a corpus large enough for early merges above the resident limit (host-driven steps in between resident ones)
    text = pkg.synth_corpus(0x5EED0021, 24 << 20).tobytes()
    tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text)
    for mode in ("lexical", "first"):
        om, oc = O.train(tok, off, w, 256 + 1500, mode)
        m, c, st = comm.train(tok, off, w, 256 + 1500, mode)
        ok = ok and m.shape == om.shape and bool((m == om).all()) and bool((c == oc).all())
        # the text itself sharded (parts cut at 1 MiB block ends = chunk boundaries): split + dedup per rank, one all-gather
        # of the unique chunks, merge loop on the combined corpus
        cut = [0, 11 << 20, len(text)] if world == 2 else [len(text) * r // world // (1 << 20) * (1 << 20) for r in range(world)] + [len(text)]
        m, c, st = comm.train_text(text[cut[rank]:cut[rank + 1]], 256 + 1500, mode)
        ok = ok and m.shape == om.shape and bool((m == om).all()) and bool((c == oc).all())
    ret[rank] = ok
    ret["resident"] = comm.resident()
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
        print("resident exchange available:", ret["resident"])
