#!/usr/bin/env python3
"""GPU side of profiles/reference_full_16MiB.json: Tokenizer::train on the 16 MiB slice the reference trained to the full
32768 vocabulary (2671 s on one CPU core); prints wall time, device time and whether the .model equals the golden."""
import hashlib, importlib.util, json, os, re, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
text = pkg.synth_corpus(0x5EED0001, 16 << 20).tobytes()
lo = len(text) - 65536
last = None
for m in re.finditer(rb"\n[\x21-\x7e]", text[lo:]):
    last = m
text = text[:lo + last.start() + 1]
golden = open(os.path.join(ROOT, "tests", "golden", "models", "synth16m_gpt4_lexical_32768.model"), "rb").read()
tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
res = {"input_sha256": hashlib.sha256(text).hexdigest(), "runs": []}
for i in range(3):
    t0 = time.time()
    tk.train(text, 32768, "lexical")
    dt = time.time() - t0
    st = tk.last_train_stats()
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "m.model")
        tk.save(p)
        same = open(p, "rb").read() == golden
    res["runs"].append({"wall_s_text_to_merges": round(dt, 4), "gpu_ms_merge_loop": round(st["gpu_ms"], 2), "model_equals_reference": same})
ref = json.load(open(os.path.join(ROOT, "profiles", "reference_full_16MiB.json")))
best = min(r["wall_s_text_to_merges"] for r in res["runs"][1:])
res["reference_seconds"] = ref["seconds"]
res["speedup_wall"] = round(ref["seconds"] / best, 1)
print(json.dumps(res))
