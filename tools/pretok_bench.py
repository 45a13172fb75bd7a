#!/usr/bin/env python3
"""Time the GPU pre-tokeniser and dedup on the GPU box: python tools/pretok_bench.py [MiB]"""
import importlib.util, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
text = pkg.synth_corpus(0x5EED0001, mib << 20)
dev = torch.device("cuda", 0)
d_text = torch.from_numpy(text).to(dev)
d_off = torch.empty(len(text) + 2, dtype=torch.int32, device=dev)
pt = pkg.Pretok()
def timed(f, n=4):
    ts = []
    for _ in range(n):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = f(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return r, min(ts[1:])
n_chunks, t_split = timed(lambda: pt.split_device(d_text.data_ptr(), len(text), d_off.data_ptr(), len(text) + 2))
print(f"split_device: {n_chunks} chunks in {t_split*1e3:.2f} ms = {len(text)/t_split/1e9:.1f} GB/s", flush=True)
def dd():
    c = pt.dedup_device(d_text.data_ptr(), len(text), d_off.data_ptr(), n_chunks)
    r = (c.n_unique, c.n_tokens); c.free(); return r
(nu, nt), t_dd = timed(dd)
print(f"dedup_device: {nu} unique, {nt} tokens in {t_dd*1e3:.2f} ms", flush=True)
tb = text.tobytes()
def corpus():
    c = pt.corpus(tb); r = c.n_unique; c.free(); return r
_, t_c = timed(corpus, 3)
print(f"corpus (pageable host text -> device corpus): {t_c*1e3:.1f} ms", flush=True)
t0 = time.perf_counter(); tok, off, w, nc = pkg.split_dedup(pkg.patterns()["gpt4"], tb); t_host = time.perf_counter() - t0
print(f"host split_dedup: {t_host*1e3:.0f} ms ({nc} chunks, {len(w)} unique)", flush=True)
assert nc == n_chunks and len(w) == nu and len(tok) == nt
