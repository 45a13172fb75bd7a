#!/usr/bin/env python3
"""A/B the decode kernel on the GPU box: python tools/dec_ab.py [MiB]  (AB_ENV="K=V;K=V" variants)"""
import importlib.util, os, sys
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], pkg.synth_corpus(0x5EED0001, 256 << 20).tobytes())
merges, _, _ = pkg.train(tok, off, w, 32768, "lexical")
text = pkg.synth_corpus(0x5EED0002, mib << 20)
dev = torch.device("cuda", 0)
d_text = torch.from_numpy(text).to(dev)
pt = pkg.Pretok(); enc = pkg.Encoder(merges)
d_off = torch.empty(len(text) + 2, dtype=torch.int32, device=dev)
nck = pt.split_device(d_text.data_ptr(), len(text), d_off.data_ptr(), len(text) + 2)
d_ids = torch.empty(len(text), dtype=torch.int32, device=dev); d_n = torch.zeros(1, dtype=torch.int64, device=dev)
enc.encode_device(d_text.data_ptr(), len(text), d_off.data_ptr(), nck, d_ids.data_ptr(), len(text), d_n.data_ptr())
n_ids = int(d_n.item())
d_out = torch.empty(len(text) + 64, dtype=torch.uint8, device=dev)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for env in (os.environ.get("AB_ENV", "").split(";") if os.environ.get("AB_ENV") else [""]) * 2:
    for kv in filter(None, env.split(",")):
        k, v = kv.split("=")
        os.environ.pop(k, None) if v == "-" else os.environ.__setitem__(k, v)
    ts = []
    enc.close(); enc = pkg.Encoder(merges)  # (the kernel shape is read when the encoder is created)
    for i in range(6):
        flush.fill_(i)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); enc.decode_device(d_ids.data_ptr(), n_ids, d_out.data_ptr(), len(text) + 64, d_n.data_ptr()); b.record()
        torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ok = int(d_n.item()) == len(text) and bool(torch.equal(d_out[:len(text)], d_text))
    print(f"{env:24s} best {min(ts[2:]):6.2f} ms  {len(text)/1e6/min(ts[2:]):7.1f} GB/s of text  ok={ok}", flush=True)
