#!/usr/bin/env python3
"""debug: one kernel shape, cold + warm, first mismatch against the default shape.  python tools/enc_dbg.py MiB CFG [reps]"""
import importlib.util, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
mib, cfg = int(sys.argv[1]), sys.argv[2]
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
text = pkg.synth_corpus(0x5EED0013, mib << 20).tobytes()
tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text[: 4 << 20])
merges, _, _ = pkg.train(tok, off, w, 256 + 3000, "lexical")
s, e = pkg.split(pkg.patterns()["gpt4"], text)
chunk_off = np.concatenate([s, e[-1:]]).astype(np.uint64)
os.environ["MBPE_ENC_CFG"] = "0"
enc = pkg.Encoder(merges)
ref, ref_off = enc.encode(text, chunk_off, want_off=True)
enc.close()
os.environ["MBPE_ENC_CFG"] = cfg
for r in range(reps):
    enc = pkg.Encoder(merges)
    for rep in range(2):
        ids, out_off = enc.encode(text, chunk_off, want_off=True)
        same = len(ids) == len(ref) and np.array_equal(ids, ref)
        msg = f"cfg {cfg} run {r} rep {rep}: n={len(ids)} ref={len(ref)} same={same}"
        if not same:
            n = min(len(ids), len(ref))
            bad = np.nonzero(ids[:n] != ref[:n])[0]
            msg += f" mismatching ids: {len(bad)}"
            shown = set()
            for i in bad[:2000]:
                c = int(np.searchsorted(ref_off, i, side="right")) - 1
                if c in shown:
                    continue
                shown.add(c)
                if len(shown) <= 6:
                    a0, a1 = int(ref_off[c]), int(ref_off[c + 1])
                    msg += (f"\n   chunk {c} (tile {c // 1024}/{c // 512}, in tile {c % 1024}) bytes={text[int(chunk_off[c]):int(chunk_off[c+1])]!r} "
                            f"got={ids[a0:a1].tolist()} want={ref[a0:a1].tolist()} counts_equal={bool((out_off[c:c+2] == ref_off[c:c+2]).all())}")
            msg += f"\n   distinct chunks among the first 2000 bad ids: {len(shown)}; tiles: {sorted(set(c // 1024 for c in shown))[:12]}"
        print(msg, flush=True)
    enc.close()
