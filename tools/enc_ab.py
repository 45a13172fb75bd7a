#!/usr/bin/env python3
"""A/B the encode kernel configurations on the GPU box: python tools/enc_ab.py [MiB] [cfg ...]"""
import importlib.util, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [int(x) for x in sys.argv[2:]] or list(range(8))
train_text = pkg.synth_corpus(0x5EED0001, 256 << 20).tobytes()
tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], train_text)
merges, _, st = pkg.train(tok, off, w, 32768, "lexical")
print("trained", len(merges), "merges in", round(st["gpu_ms"], 1), "ms", flush=True)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
batches = []
for seed in (0x5EED0002, 0x5EED0003):  # two DISTINCT texts: B is new text for caches that have only seen A
    text = pkg.synth_corpus(seed, mib << 20)
    s, e = pkg.split(pkg.patterns()["gpt4"], text.tobytes())
    off32 = np.concatenate([s, e[-1:]]).astype(np.uint32)
    batches.append((len(text), len(s), torch.from_numpy(text).to(dev), torch.from_numpy(off32.view(np.int32)).to(dev)))
d_out = torch.empty(mib << 20, dtype=torch.int32, device=dev)
d_n = torch.zeros(1, dtype=torch.int64, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def run(enc, b):
    nb, nc, d_bytes, d_off = batches[b]
    flush.fill_(1)  # evict L2 between runs
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    enc.encode_device(d_bytes.data_ptr(), nb, d_off.data_ptr(), nc, d_out.data_ptr(), mib << 20, d_n.data_ptr(), stream)
    t1.record(); torch.cuda.synchronize()
    return t0.elapsed_time(t1)


ref = None
for cfg in cfgs:
    for env in (os.environ.get("AB_ENV", "").split(";") if os.environ.get("AB_ENV") else [""]):
        for kv in filter(None, env.split(",")):
            k, v = kv.split("=")
            if v == "-":
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
        os.environ["MBPE_ENC_CFG"] = str(cfg)
        enc = pkg.Encoder(merges)
        enc.reserve(mib << 20, max(b[1] for b in batches))
        cold_a = run(enc, 0)
        new_b = run(enc, 1)
        n_b = int(d_n.item())
        warm = [run(enc, 0) for _ in range(4)]
        n = int(d_n.item())
        ids = d_out[:n].cpu().numpy()
        if ref is None:
            ref = ids.copy()
        ok = np.array_equal(ids, ref)
        best = min(warm)
        nb, nc = batches[0][0], batches[0][1]
        b_enc = nb + 4 * nc + 4 * n
        print(f"cfg {cfg} {env:34s} cold A {cold_a:7.2f}  new text B {new_b:7.2f}  warm A best {best:6.2f} ms = {nb/1e6/best:7.1f} GB/s text, "
              f"B_enc {b_enc/1e6/best:7.1f} GB/s  same_ids={ok}", flush=True)
        enc.close()
