#!/usr/bin/env python3
"""Tokenizer::encode of a text with a special token every few KiB: device front end vs host front end."""
import importlib.util, os, sys, tempfile, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 512
base = pkg.synth_corpus(0x5EED0002, mib << 20)
tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], pkg.synth_corpus(0x5EED0001, 64 << 20).tobytes())
merges, _, _ = pkg.train(tok, off, w, 32768, "lexical")
# documents of ~4 KiB separated by <|endoftext|>, cut after newlines
b = base.tobytes(); docs = []; pos = 0
while pos < len(b):
    nxt = b.find(b"\n", min(pos + 4096, len(b)))
    nxt = len(b) if nxt < 0 else nxt + 1
    docs.append(b[pos:nxt]); pos = nxt
text = b"<|endoftext|>".join(docs)
print(f"{len(text)/2**20:.0f} MiB, {len(docs)} documents", flush=True)
pin = torch.empty(len(text), dtype=torch.uint8, pin_memory=True); pin.numpy()[:] = np.frombuffer(text, np.uint8)
out = torch.empty(len(text), dtype=torch.int32, pin_memory=True)
with tempfile.TemporaryDirectory() as td:
    mp = os.path.join(td, "m.model")
    pkg.write_model(mp, pkg.patterns()["gpt4"], b"<|endoftext|> 100257\n", merges)
    res = {}
    for mode in ("1", "0"):
        os.environ["MBPE_GPU_SPLIT"] = mode
        tk = pkg.Tokenizer(pkg.patterns()["gpt4"]); tk.load(mp)
        for i in range(3 if mode == "1" else 1):
            t0 = time.time(); ids = tk.encode(pin.numpy(), out=out.numpy().view(np.uint32)); dt = time.time() - t0
            print(f"MBPE_GPU_SPLIT={mode} #{i}: {dt:.3f} s = {len(text)/dt/1e6:.0f} MB/s, {len(ids)} ids", flush=True)
        res[mode] = ids.copy()
print("same ids:", np.array_equal(res["0"], res["1"]))
