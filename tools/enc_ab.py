#!/usr/bin/env python3
"""A/B the encode kernel configurations on the GPU box: python tools/enc_ab.py [MiB] [cfg ...]"""
import importlib.util, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
cfgs = [int(x) for x in sys.argv[2:]] or list(range(7))
train_text = pkg.synth_corpus(0x5EED0001, 256 << 20).tobytes()
tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], train_text)
merges, _, st = pkg.train(tok, off, w, 32768, "lexical")
print("trained", len(merges), "merges in", round(st["gpu_ms"], 1), "ms", flush=True)
text = pkg.synth_corpus(0x5EED0002, mib << 20)
s, e = pkg.split(pkg.patterns()["gpt4"], text.tobytes())
off32 = np.concatenate([s, e[-1:]]).astype(np.uint32)
dev = torch.device("cuda", 0)
d_bytes = torch.from_numpy(text).to(dev)
d_off = torch.from_numpy(off32.view(np.int32)).to(dev)
d_out = torch.empty(len(text), dtype=torch.int32, device=dev)
d_n = torch.zeros(1, dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
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
        enc.reserve(len(text), len(s))
        times = []
        for it in range(5):
            t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
            t0.record()
            enc.encode_device(d_bytes.data_ptr(), len(text), d_off.data_ptr(), len(s), d_out.data_ptr(), len(text), d_n.data_ptr(), stream)
            t1.record(); torch.cuda.synchronize()
            times.append(t0.elapsed_time(t1))
        n = int(d_n.item())
        ids = d_out[:n].cpu().numpy()
        if ref is None:
            ref = ids.copy()
        ok = np.array_equal(ids, ref)
        best = min(times[2:])
        print(f"cfg {cfg} {env:30s} cold {times[0]:7.2f} ms  warm best {best:6.2f} ms  {len(text)/1e6/best:7.1f} GB/s  same_ids={ok}", flush=True)
        enc.close()
