#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/sharded_check.py [MiB] [vocab]: sharded training on N GPUs vs the oracle and vs
one GPU, with timings (evidence for DESIGN.md section 5)."""
import importlib.util, json, os, sys, time
import numpy as np
import torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
from oracle import oracle as O
rank, lr, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 64
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
def bcast(ident):
    t = torch.from_numpy(ident.copy()).cuda(); dist.broadcast(t, 0); return t.cpu().numpy()
comm = pkg.Comm(rank, world, lr, bcast)
text = pkg.synth_corpus(0x5EED0001, mib << 20).tobytes()
tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text)
res = {"world": world, "corpus_mib": mib, "vocab": vocab, "n_positions": int(len(tok)), "resident_exchange": comm.resident()}
for mode in ("lexical", "first"):
    om, oc = O.train(tok, off, w, vocab, mode)
    dist.barrier(); t0 = time.time()
    m, c, st = comm.train(tok, off, w, vocab, mode)
    torch.cuda.synchronize(); dist.barrier(); t_sh = time.time() - t0
    ok = m.shape == om.shape and bool((m == om).all()) and bool((c == oc).all())
    flag = torch.tensor([int(ok)], device="cuda"); dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    m1, c1, st1 = pkg.train(tok, off, w, vocab, mode, device=lr)
    t_sw, ok2 = None, None
    if not os.environ.get("CHECK_SKIP_STEPWISE"):
        os.environ["MBPE_SHARDED_STEPWISE"] = "1"  # the round-1 path: grid kernels + one NCCL all-gather per merge
        dist.barrier(); t0 = time.time()
        m2, c2, st2 = comm.train(tok, off, w, vocab, mode)
        torch.cuda.synchronize(); dist.barrier(); t_sw = time.time() - t0
        del os.environ["MBPE_SHARDED_STEPWISE"]
        ok2 = m2.shape == om.shape and bool((m2 == om).all()) and bool((c2 == oc).all())
    res[mode] = {"all_ranks_equal_oracle": bool(flag.item()), "sharded_wall_s": round(t_sh, 3), "sharded_gpu_ms": round(st["gpu_ms"], 1),
                 "host_driven_merges": st["n_big_merges"], "launches": st["n_launches"], "single_gpu_ms": round(st1["gpu_ms"], 1),
                 "resident_steps": st["resident_cycles"]["steps"], "resident_cycles_per_step": round(st["resident_cycles"]["total"] / max(st["resident_cycles"]["steps"], 1)),
                 "stepwise_nccl_wall_s": None if t_sw is None else round(t_sw, 3), "stepwise_equal_oracle": ok2,
                 "us_per_merge_sharded": round(1e3 * st["gpu_ms"] / max(len(m), 1), 1),
                 "us_per_merge_single": round(1e3 * st1["gpu_ms"] / max(len(m1), 1), 1)}
if rank == 0:
    print(json.dumps(res))
comm.close()
dist.destroy_process_group()
