#!/usr/bin/env python3
"""Where the end-to-end time goes (run with MBPE_DEBUG=1): python tools/e2e_probe.py [MiB]"""
import importlib.util, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
text = pkg.synth_corpus(0x5EED0001, mib << 20)
pinned = torch.empty(len(text), dtype=torch.uint8, pin_memory=True); pinned.numpy()[:] = text
tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
only_encode = len(sys.argv) > 2 and sys.argv[2] == "encode"
if only_encode:
    tk.train(text[:64 << 20], 32768, "lexical")
for name, buf in [] if only_encode else (("pageable", text), ("pinned", pinned.numpy())):
    for i in range(3):
        t0 = time.time(); tk.train(buf, 32768, "lexical"); torch.cuda.synchronize()
        print(f"train {name} #{i}: {time.time()-t0:.3f} s", flush=True)
etext = pkg.synth_corpus(0x5EED0002, mib << 20)
epin = torch.empty(len(etext), dtype=torch.uint8, pin_memory=True); epin.numpy()[:] = etext
opin = torch.empty(len(etext), dtype=torch.int32, pin_memory=True)
opage = np.empty(len(etext), np.uint32)
for name, buf, out in (("pageable", etext, opage), ("pinned", epin.numpy(), opin.numpy().view(np.uint32))):
    for i in range(3):
        t0 = time.time(); ids = tk.encode(buf, out=out); dt = time.time() - t0
        print(f"encode {name} #{i}: {dt:.3f} s = {len(etext)/dt/1e6:.0f} MB/s ({len(ids)} ids)", flush=True)
