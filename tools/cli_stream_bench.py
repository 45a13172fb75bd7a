#!/usr/bin/env python3
"""disk -> disk through the CLI on the GPU box: streaming --encode vs the whole-file path with the host front end."""
import importlib.util, os, resource, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
pkg = importlib.util.module_from_spec(spec); spec.loader.exec_module(pkg)
D = "/tmp/mbpe_cli"; os.makedirs(D, exist_ok=True)
mib = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
pkg.synth_corpus(0x5EED0002, mib << 20).tofile(f"{D}/big.txt")
pkg.synth_corpus(0x5EED0001, 64 << 20).tofile(f"{D}/train.txt")
CLI = os.path.join(ROOT, "minbpe-cc_b200", "bin", "minbpe-cc")
def run(label, args, env=None):
    r0 = resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss
    t0 = time.time()
    r = subprocess.run([CLI] + args, capture_output=True, text=True, env={**os.environ, **(env or {})})
    dt = time.time() - t0
    assert r.returncode == 0, (r.stdout[-500:], r.stderr[-500:])
    for line in r.stderr.splitlines():
        if "encode_file" in line or "cli:" in line:
            print("   ", line)
    print(f"{label}: {dt:.2f} s wall, max child rss so far {max(r0, resource.getrusage(resource.RUSAGE_CHILDREN).ru_maxrss) >> 10} MiB", flush=True)
    return dt
run("train 64 MiB, vocab 32768 (text file -> .model)", ["--train", "-i", f"{D}/train.txt", "-m", f"{D}/m.model", "--vocab-size", "32768", "--encoder", "gpt4", "-c", "lexical"])
for i in range(2):
    dt = run(f"encode {mib} MiB, streaming (block-wise disk -> GPU -> disk)", ["--encode", "-i", f"{D}/big.txt", "-m", f"{D}/m.model", "-o", f"{D}/s.enc"])
    print(f"   = {mib * 1.048576 / dt:.0f} MB/s of text incl. process start, model load, CUDA init", flush=True)
run(f"encode {mib} MiB, whole file in memory, device front end", ["--encode", "-i", f"{D}/big.txt", "-m", f"{D}/m.model", "-o", f"{D}/d.enc"], {"MBPE_ENCODE_SEG_BYTES": str(64 << 20), "MBPE_CLI_NO_STREAM": "1"})
run(f"encode {mib} MiB, whole file in memory, host front end (PCRE2)", ["--encode", "-i", f"{D}/big.txt", "-m", f"{D}/m.model", "-o", f"{D}/w.enc"], {"MBPE_GPU_SPLIT": "0"})
a, b = open(f"{D}/s.enc", "rb").read(), open(f"{D}/w.enc", "rb").read()
print("outputs identical:", a == b, len(a), "bytes", flush=True)
run("decode", ["--decode", "-i", f"{D}/s.enc", "-m", f"{D}/m.model", "-o", f"{D}/back.txt"])
print("round trip exact:", open(f"{D}/back.txt", "rb").read() == open(f"{D}/big.txt", "rb").read())
