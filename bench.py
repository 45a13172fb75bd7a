#!/usr/bin/env python3
"""bench.py -- the reference's headline metric on B200: BPE train merges/s (+ wall-s) on a 1 GiB synthetic
Zipfian UTF-8 corpus at vocab 32768 (BASELINE.json configs[2]), and encode MB/s with that model (configs[3]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (GPU)
    python bench.py --impl reference [...]                         the reference's own CPU code, bounded sample

One JSON line on stdout (rank 0). A "step" of the train workload is one complete merge loop (vocab-256 merges)
over the resident deduplicated corpus; a step of the encode workload is one pass over the resident text batch.
PyTorch is used for plumbing only: torch.distributed (NCCL) barrier/max across ranks, device buffers for the
encode batch, streams and events. All compute goes through the C ABI of libminbpe_b200.so.

Multi-GPU (N > 1, launched by torchrun): encode shards chunks across ranks with no communication (each rank
encodes its own batch: weak scaling). Train: `value` counts N independent replicas of the merge loop (weak scaling);
the sharded trainer (chunks sharded, per-merge NCCL exchange of pair-count deltas) is run once and reported
under train.sharded -- it is exact but one collective per merge makes it slower than a single GPU.
"""
import argparse
import hashlib
import importlib.util
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_TRAIN, SEED_ENCODE = 0x5EED0001, 0x5EED0002


def load_pkg():
    spec = importlib.util.spec_from_file_location("minbpe_cc_b200", os.path.join(ROOT, "minbpe-cc_b200", "__init__.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["minbpe_cc_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None
        self.marks = []

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self, t0, t1):
        self.marks.append((t0, t1))

    def summary(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, line in self.rows:
            if self.marks and not any(a - 0.2 <= t <= b + 0.2 for a, b in self.marks):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def safe_prefix(text: bytes, n: int) -> bytes:
    """largest prefix <= n bytes ending at a regex-safe cut: after '\\n', before a printable ASCII byte."""
    n = min(n, len(text))
    lo = max(0, n - 65536)
    last = None
    for m in re.finditer(rb"\n[\x21-\x7e]", text[lo:n]):
        last = m
    return text[:lo + last.start() + 1] if last else text[:n]


# --------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation (oracle/_ref/ref_driver = reference headers compiled
# verbatim), on a bounded sample of the same workload.
# --------------------------------------------------------------------------------------------------------
def ref_train_sample(pkg, corpus_mib, sample_mib, n_merges, mode):
    from oracle import oracle as O
    text = pkg.synth_corpus(SEED_TRAIN, sample_mib << 20).tobytes()  # the generator is blockwise: same first MiBs
    text = safe_prefix(text, len(text))
    with tempfile.TemporaryDirectory() as td:
        inp, model = os.path.join(td, "in.txt"), os.path.join(td, "m.model")
        open(inp, "wb").write(text)
        if os.path.exists(O.REF_DRIVER):
            kind = "reference"
            p = subprocess.run([O.REF_DRIVER, "train", inp, model, str(256 + n_merges), "gpt4", mode],
                               capture_output=True, text=True)
            secs = float(re.search(r"REF_TIME_S ([0-9.eE+-]+)", p.stderr).group(1))
            merges = O.read_model(model)[2]
        else:  # the compiled reference did not travel: time the C restatement's literal walk instead
            kind = "port"
            t0 = time.time()
            t, o, w = O.flatten(O.chunks_of(text, "gpt4"), dedup=False)
            merges, _ = O.train(t, o, w, 256 + n_merges, mode, impl="rescan")
            secs = time.time() - t0
    mps = n_merges / secs
    return {"value": mps, "unit": "merges/s", "cores": 1, "kind": kind,
            "sample": f"Tokenizer::train (regex split + lists + merge loop) on the first {len(text)} bytes "
                      f"({sample_mib} MiB) of the same synthetic corpus, {n_merges} merges, {mode}, 1 thread "
                      f"(the reference is single-threaded); took {secs:.2f} s",
            "seconds": secs, "sample_bytes": len(text),
            "value_extrapolated_to_workload": mps * len(text) / float(corpus_mib << 20),
            "extrapolation": "per-merge cost of the reference is linear in corpus bytes (it walks every chunk every "
                             "merge, Tokenizer.h:309-320): merges/s x sample_bytes / workload_bytes",
            }, merges


def ref_encode_sample(pkg, merges_path, sample_mib):
    from oracle import oracle as O
    text = pkg.synth_corpus(SEED_ENCODE, sample_mib << 20).tobytes()
    with tempfile.TemporaryDirectory() as td:
        inp, out = os.path.join(td, "in.txt"), os.path.join(td, "o.enc")
        open(inp, "wb").write(text)
        if os.path.exists(O.REF_DRIVER):
            kind = "reference"
            p = subprocess.run([O.REF_DRIVER, "encode", inp, merges_path, out], capture_output=True, text=True)
            secs = float(re.search(r"REF_TIME_S ([0-9.eE+-]+)", p.stderr).group(1))
            digest = hashlib.sha256(open(out, "rb").read()).hexdigest()
        else:
            kind = "port"
            pat, sp, m = O.read_model(merges_path)
            t0 = time.time()
            ids = O.encode_text(text, pat, sp, m)
            secs = time.time() - t0
            digest = hashlib.sha256(ids.tobytes()).hexdigest()
    return {"value": len(text) / 1e6 / secs, "unit": "MB/s", "cores": 1, "kind": kind,
            "sample": f"Tokenizer::encode (regex split + merge scan) on {sample_mib} MiB of the synthetic encode corpus, "
                      f"1 thread; took {secs:.2f} s", "seconds": secs, "sha256": digest}, text


def run_reference_arm(a, rank, world):
    if rank != 0:
        return
    pkg = load_pkg()
    vals = []
    for _ in range(a.warmup + a.steps):
        cb, _ = ref_train_sample(pkg, a.corpus_mib, a.ref_sample_mib, a.ref_merges, a.mode)
        vals.append(cb)
    vals = vals[a.warmup:] or vals
    v = sum(x["value"] for x in vals) / len(vals)
    cb = dict(vals[-1], value=v)
    line = {"impl": "reference", "metric": "bpe_train_merges_per_sec", "value": v, "unit": "merges/s",
            "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * cb["seconds"],
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
            "config": workload_config(a), "cpu_baseline": cb,
            "e2e": {"value": v, "unit": "merges/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(a):
    return {"workload": f"train {a.corpus_mib} MiB synthetic Zipfian UTF-8 corpus (seed 0x{SEED_TRAIN:X}), "
                        f"--vocab-size {a.vocab} --encoder gpt4 -c {a.mode}",
            "vocab_size": a.vocab, "corpus_bytes": a.corpus_mib << 20, "mode": a.mode, "encoder": "gpt4",
            "engine": a.engine, "l2": "each step re-streams the whole working set (> L2) and a 256 MiB buffer is "
                                      "written between timed steps"}


# --------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--corpus-mib", type=int, default=1024)
    ap.add_argument("--vocab", type=int, default=32768)
    ap.add_argument("--mode", default="lexical", choices=["first", "lexical"])
    ap.add_argument("--engine", default="persistent", choices=["persistent", "stepwise"])
    ap.add_argument("--encode-mib", type=int, default=1024)
    ap.add_argument("--encode-batches", type=int, default=1, help="resident batches of --encode-mib per step")
    ap.add_argument("--ref-sample-mib", type=int, default=4)
    ap.add_argument("--ref-merges", type=int, default=128)
    ap.add_argument("--ref-encode-mib", type=int, default=64)
    ap.add_argument("--e2e-budget-s", type=float, default=90.0)
    ap.add_argument("--skip-encode", action="store_true")
    ap.add_argument("--skip-cpu-baseline", action="store_true")
    ap.add_argument("--no-check", action="store_true", help="skip the full-size diff of the merge list against the CPU oracle")
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if a.impl == "reference":
        run_reference_arm(a, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    pkg = load_pkg()
    if pkg.device_count() < 1 or not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libminbpe_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    clocks = ClockSampler(local_rank) if rank == 0 else None
    peak, peak_src = measured_peaks()
    n_merges_target = a.vocab - 256

    # ---------------- host prep (outside every timed region): corpus, regex split, dedup -----------------
    t0 = time.time()
    text = pkg.synth_corpus(SEED_TRAIN, a.corpus_mib << 20)  # N > 1: every rank holds the same corpus
    t_gen = time.time() - t0
    tb = text.tobytes()
    t0 = time.time()
    tok, off, w, n_chunks = pkg.split_dedup(pkg.patterns()["gpt4"], tb)
    t_split = time.time() - t0
    t_dedup = 0.0  # fused into the split pass

    # ---------------- train: device-resident steps -------------------------------------------------------
    trainer = pkg.Trainer(tok, off, w, device=local_rank)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * (a.warmup + a.steps))]
    times, stats, merges = [], None, None
    for i in range(a.warmup + a.steps):
        flush_buf.fill_(i & 0xFF)  # evict L2 between steps
        if i == a.warmup:
            barrier()
            wall0 = time.time()
        ev[2 * i].record()
        merges, counts, stats = trainer.run(a.vocab, a.mode, a.engine, stream)
        ev[2 * i + 1].record()
    barrier()
    wall1 = time.time()
    if clocks:
        clocks.mark(wall0, wall1)
    for i in range(a.warmup, a.warmup + a.steps):
        times.append(ev[2 * i].elapsed_time(ev[2 * i + 1]))
    ms_per_step = max_over_ranks(sum(times) / len(times))
    n_done = len(merges)
    value = world * n_done / (ms_per_step / 1e3)
    launches_train = stats["n_launches"] * a.steps
    model_sha = hashlib.sha256(merges.tobytes()).hexdigest()

    # ---------------- e2e: reference-facing call, HOST text in -> merges out -----------------------------
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"], device=local_rank)
    tk.set_engine(a.engine)
    text_pinned = torch.empty(len(text), dtype=torch.uint8, pin_memory=True)  # the step's input, in pinned host memory
    text_pinned.numpy()[:] = text
    e2e_times = []
    t_budget = time.time()
    for i in range(1 + max(1, a.steps)):  # the first call is a warm-up: it sizes the tokenizer's resident device buffers
        barrier()
        t0 = time.time()
        tk.train(text_pinned.numpy(), a.vocab, a.mode)
        torch.cuda.synchronize()
        if i:
            e2e_times.append(time.time() - t0)
        if time.time() - t_budget > a.e2e_budget_s and e2e_times:
            break
    e2e_s = max_over_ranks(sum(e2e_times) / len(e2e_times))
    st2 = tk.last_train_stats()
    assert hashlib.sha256(tk.merges().tobytes()).hexdigest() == model_sha, "tokenizer path and trainer path disagree"
    on_gpu = bool(st2.get("split_on_gpu"))
    h2d = int(len(tb)) if on_gpu else int(st2["n_positions"] * 4 + (st2["n_unique"] + 1) * 8 + st2["n_unique"] * 4)
    d2h = int(n_done * 12)
    # the same call with pre-tokenisation kept on the host (PCRE2 on all cores), once, for the record
    e2e_host_split_s = None
    if on_gpu and rank == 0:
        os.environ["MBPE_GPU_SPLIT"] = "0"
        try:
            tkh = pkg.Tokenizer(pkg.patterns()["gpt4"], device=local_rank)
            tkh.set_engine(a.engine)
            t0 = time.time()
            tkh.train(tb, a.vocab, a.mode)
            e2e_host_split_s = time.time() - t0
            assert hashlib.sha256(tkh.merges().tobytes()).hexdigest() == model_sha, "host-split and device-split models differ"
            del tkh
        finally:
            del os.environ["MBPE_GPU_SPLIT"]
    # the C-ABI hot-path boundary with deduplicated HOST buffers (H2D + merge loop + D2H per step)
    abi_times = []
    for i in range(a.steps):
        barrier()
        t0 = time.time()
        m2, _, _ = pkg.train(tok, off, w, a.vocab, a.mode, a.engine, device=local_rank)
        abi_times.append(time.time() - t0)
    abi_s = max_over_ranks(sum(abi_times) / len(abi_times))

    line = {
        "metric": "bpe_train_merges_per_sec", "value": value, "unit": "merges/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": workload_config(a),
        "e2e": {"value": world * n_done / e2e_s, "unit": "merges/s", "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": len(e2e_times), "wall_s_per_step": e2e_s,
                "wall_s_steps": [round(t, 4) for t in e2e_times],
                "what": "mbpe_tokenizer_train(text) = Tokenizer::train, text in pinned host memory: " +
                        ("H2D text + device split (GPT-4 matcher) + device dedup + merge loop + D2H merges" if on_gpu else
                         "host regex split + dedup + H2D + merge loop + D2H merges"),
                "split": "device" if on_gpu else "host", "split_dedup_s": st2["split_s"], "gpu_ms": st2["gpu_ms"],
                "wall_s_with_host_split": e2e_host_split_s},
        "e2e_abi": {"value": world * n_done / abi_s, "unit": "merges/s", "wall_s_per_step": abi_s,
                    "what": "mbpe_train(deduplicated host buffers): H2D + merge loop + D2H"},
        "gpu_launches": int(launches_train),
        "train": {"merges": int(n_done), "merges_sha256": model_sha, "gpu_s_per_run": ms_per_step / 1e3,
                  "wall_s_text_to_model": e2e_s, "host_prep_s": {"generate": t_gen, "split": t_split, "dedup": t_dedup},
                  "n_chunks": int(n_chunks), "n_unique_chunks": int(len(w)), "stats": stats,
                  "multi_gpu": None if world == 1 else "value = N independent replicas of the merge loop (weak scaling); the "
                                                       "sharded trainer is reported under train.sharded"},
    }
    # roofline of the merge loop: SURVEY 8(d) full-rescan algorithmic volume / device time
    ach = stats["rescan_bytes"] / 1e9 / (ms_per_step / 1e3)
    line["roofline"] = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                        "traffic": None, "peak_source": peak_src, "kernel": "k_persistent (+ per-phase grid kernels)",
                        "algorithmic_bytes_per_step": int(stats["rescan_bytes"]),
                        "note": "algorithmic bytes = sum over merges of 12*T_m + 16*P_m, what a full rescan per merge "
                                "would move (SURVEY 8(d)); the incremental kernels move far fewer real bytes, so the "
                                "fraction can exceed 1 and the loop is latency-bound, not HBM-bound"}

    if world > 1:
        # the north star's multi-GPU train: unique chunks sharded over the ranks, per-merge exchange of pair-count
        # deltas over NCCL/NVLink. It is exact but latency-bound (one collective per merge), so it is reported beside
        # the replica number, not instead of it.
        def bcast(ident):
            t = torch.from_numpy(ident.copy()).to(dev)
            dist.broadcast(t, 0)
            return t.cpu().numpy()
        comm = pkg.Comm(rank, world, local_rank, bcast)
        barrier()
        t0 = time.time()
        sm, sc, sst = comm.train(tok, off, w, a.vocab, a.mode, stream)
        torch.cuda.synchronize()
        barrier()
        sh_s = max_over_ranks(time.time() - t0)
        same = torch.tensor([int(sm.shape == merges.shape and bool((sm == merges).all()))], device=dev)
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
        line["train"]["sharded"] = {"merges_per_sec": len(sm) / sh_s, "wall_s": sh_s, "gpu_ms": sst["gpu_ms"],
                                    "exchanges": sst["n_big_merges"], "launches": sst["n_launches"],
                                    "equals_single_gpu_merges_on_every_rank": bool(same.item()),
                                    "what": "mbpe_train_sharded: replicated pair table, chunks sharded over ranks, "
                                            "one NCCL all-gather of count deltas per merge"}
        comm.close()

    if not a.no_check and rank == 0:
        # parity at the FULL workload size: the CPU oracle (indexed restatement, validated against the compiled
        # reference on the small configs) trains the same deduplicated corpus; merge lists and counts must be equal
        from oracle import oracle as O
        t0 = time.time()
        om, oc = O.train(tok, off, w, a.vocab, a.mode)
        line["train"]["oracle_check"] = {"merges_equal": bool(om.shape == merges.shape and (om == merges).all()),
                                         "counts_equal": bool(oc.shape == counts.shape and (oc == counts).all()),
                                         "oracle_s": time.time() - t0, "oracle": "oracle_train_indexed, 1 thread"}

    # ---------------- encode ----------------------------------------------------------------------------
    if not a.skip_encode:
        # B distinct resident batches of encode_mib each (a device batch is < 4 GiB: u32 boundaries); one timed step =
        # one pass over all of them. --encode-batches 10 --encode-mib 1024 is BASELINE config 4 (10 GiB).
        enc = pkg.Encoder(merges, device=local_rank)
        pt = pkg.Pretok(device=local_rank)
        batches, t_esplit, n_echunks_total, n_bytes_total = [], 0.0, 0, 0
        for b in range(a.encode_batches):
            etext = pkg.synth_corpus(SEED_ENCODE + 1000 * rank + b, a.encode_mib << 20)
            d_bytes = torch.from_numpy(etext).to(dev)
            d_off_full = torch.empty(len(etext) + 2, dtype=torch.int32, device=dev)
            torch.cuda.synchronize()
            t0 = time.time()
            nck = pt.split_device(d_bytes.data_ptr(), len(etext), d_off_full.data_ptr(), len(etext) + 2)  # device matcher
            t_esplit += time.time() - t0
            d_off = d_off_full[:nck + 1].clone()
            del d_off_full
            batches.append({"n_chunks": nck, "n_bytes": len(etext), "d_bytes": d_bytes, "d_off": d_off})
            n_echunks_total += nck
            n_bytes_total += len(etext)
        etb = etext.tobytes()  # the last batch also goes through the host paths below
        t0 = time.time()
        es, ee = pkg.split(pkg.patterns()["gpt4"], etb)  # PCRE2 on the host: must give the device matcher's offsets
        t_host_split = time.time() - t0
        eoff64 = np.concatenate([es, ee[-1:]]).astype(np.uint64)
        split_equal = bool(np.array_equal(batches[-1]["d_off"].cpu().numpy().view(np.uint32).astype(np.uint64), eoff64))
        n_echunks = batches[-1]["n_chunks"]
        d_out = torch.empty(a.encode_mib << 20, dtype=torch.int32, device=dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)
        enc.reserve(a.encode_mib << 20, max(bt["n_chunks"] for bt in batches))
        eev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * (a.warmup + a.steps))]
        n_ids_total = 0
        for i in range(a.warmup + a.steps):
            if i == a.warmup:
                barrier()
                ew0 = time.time()
            eev[2 * i].record()
            for bt in batches:
                enc.encode_device(bt["d_bytes"].data_ptr(), bt["n_bytes"], bt["d_off"].data_ptr(), bt["n_chunks"],
                                  d_out.data_ptr(), a.encode_mib << 20, d_n.data_ptr(), stream)
            eev[2 * i + 1].record()
        barrier()
        if clocks:
            clocks.mark(ew0, time.time())
        etimes = [eev[2 * i].elapsed_time(eev[2 * i + 1]) for i in range(a.warmup, a.warmup + a.steps)]
        ems = max_over_ranks(sum(etimes) / len(etimes))
        n_ids = int(d_n.item())  # ids of the last batch (the one still in d_out)
        n_ids_total = n_ids
        for bt in batches[:-1]:  # untimed: id counts of the other batches for the algorithmic-byte figure
            enc.encode_device(bt["d_bytes"].data_ptr(), bt["n_bytes"], bt["d_off"].data_ptr(), bt["n_chunks"],
                              d_out.data_ptr(), a.encode_mib << 20, d_n.data_ptr(), stream)
            n_ids_total += int(d_n.item())
        if len(batches) > 1:  # leave the last batch's ids in d_out for the checks below
            bt = batches[-1]
            enc.encode_device(bt["d_bytes"].data_ptr(), bt["n_bytes"], bt["d_off"].data_ptr(), bt["n_chunks"],
                              d_out.data_ptr(), a.encode_mib << 20, d_n.data_ptr(), stream)
            torch.cuda.synchronize()
        b_enc = n_bytes_total + 4 * n_echunks_total + 4 * n_ids_total
        etext_len = n_bytes_total
        # e2e: host text in, host ids out through the Tokenizer mirror (H2D + device split + merge scan + D2H)
        etext_pinned = torch.empty(len(etext), dtype=torch.uint8, pin_memory=True)
        etext_pinned.numpy()[:] = etext
        ids_pinned = torch.empty(len(etext), dtype=torch.int32, pin_memory=True)
        ids_out = ids_pinned.numpy().view(np.uint32)
        tk.encode(etext_pinned.numpy()[:1 << 20], out=ids_out)  # builds the tokenizer's device tables outside the timed calls
        tk.encode(etext_pinned.numpy(), out=ids_out)             # and sizes its resident buffers
        e2e_enc_times = []
        for i in range(max(1, a.steps)):
            barrier()
            t0 = time.time()
            ids = tk.encode(etext_pinned.numpy(), out=ids_out)  # the last batch: pinned text in, pinned ids out
            e2e_enc_times.append(time.time() - t0)
        e2e_enc_s = max_over_ranks(sum(e2e_enc_times) / len(e2e_enc_times))
        ids_dev = d_out[:n_ids].cpu().numpy().view(np.uint32)
        assert np.array_equal(ids, ids_dev)
        # and the hot-path boundary alone: mbpe_encode(host bytes + host chunk offsets)
        t0 = time.time()
        ids_abi = enc.encode(etb, eoff64)
        e2e_abi_s = max_over_ranks(time.time() - t0)
        assert np.array_equal(ids_abi, ids_dev)
        # size-independent property at full size: decode(encode(x)) == x
        roundtrip = enc.decode(ids) == etb
        # decode gather (Tokenizer.h:725-751), ids and bytes resident: the ids of the last batch are still in d_out
        d_txt = torch.empty(len(etext) + 64, dtype=torch.uint8, device=dev)
        d_ntxt = torch.zeros(1, dtype=torch.int64, device=dev)
        dev_ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * (a.warmup + a.steps))]
        for i in range(a.warmup + a.steps):
            flush_buf.fill_(i & 0xFF)
            dev_ev[2 * i].record()
            enc.decode_device(d_out.data_ptr(), n_ids, d_txt.data_ptr(), len(etext) + 64, d_ntxt.data_ptr(), stream)
            dev_ev[2 * i + 1].record()
        barrier()
        dms = max_over_ranks(sum(dev_ev[2 * i].elapsed_time(dev_ev[2 * i + 1]) for i in range(a.warmup, a.warmup + a.steps)) / a.steps)
        n_txt = int(d_ntxt.item())
        decode_ok = n_txt == len(etext) and bool(torch.equal(d_txt[:n_txt], batches[-1]["d_bytes"]))
        b_dec = 4 * n_ids + n_txt  # read every id once, write every byte once (the vocabulary tables stay in cache)
        line["decode"] = {
            "metric": "bpe_decode_mb_per_sec", "value": world * n_txt / 1e6 / (dms / 1e3), "unit": "MB/s", "ms_per_step": dms,
            "n_ids": n_ids, "bytes_out": n_txt, "equals_input_text": decode_ok,
            "roofline": {"bound": "hbm", "achieved": b_dec / 1e9 / (dms / 1e3), "peak": peak, "unit": "GB/s",
                         "frac": b_dec / 1e9 / (dms / 1e3) / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "k_decode_tiles", "algorithmic_bytes_per_step": int(b_dec)},
        }
        line["gpu_launches"] += a.steps
        line["encode"] = {
            "metric": "bpe_encode_mb_per_sec", "value": world * etext_len / 1e6 / (ems / 1e3), "unit": "MB/s",
            "ms_per_step": ems, "bytes_per_step": etext_len, "n_chunks": int(n_echunks_total), "n_tokens": int(n_ids_total),
            "workload": f"encode {a.encode_batches} x {a.encode_mib} MiB synthetic text per rank (seeds 0x{SEED_ENCODE:X}+1000*rank+b) "
                        f"with the {a.vocab}-vocab model just trained; inputs and output resident in HBM, every batch >> L2",
            "roofline": {"bound": "hbm", "achieved": b_enc / 1e9 / (ems / 1e3), "peak": peak, "unit": "GB/s",
                         "frac": b_enc / 1e9 / (ems / 1e3) / peak, "traffic": None, "peak_source": peak_src,
                         "kernel": "k_encode_tiles", "algorithmic_bytes_per_step": int(b_enc)},
            "e2e": {"value": world * len(etext) / 1e6 / e2e_enc_s, "unit": "MB/s", "h2d_bytes_per_step": int(len(etext)),
                    "bytes": len(etext), "d2h_bytes_per_step": int(4 * n_ids),
                    "steps": len(e2e_enc_times),
                    "what": "mbpe_tokenizer_encode(text) = Tokenizer::encode, text and ids in pinned host memory: H2D text + "
                            "device split + merge scan + D2H ids"},
            "e2e_abi": {"value": world * len(etext) / 1e6 / e2e_abi_s, "unit": "MB/s",
                        "what": "mbpe_encode(host bytes + host chunk offsets): H2D + merge scan + D2H"},
            "device_split_s": t_esplit, "host_split_s_last_batch": t_host_split, "device_split_equals_pcre2": split_equal,
            "roundtrip_ok": bool(roundtrip), "ids_sha256": hashlib.sha256(ids.tobytes()).hexdigest(),
            "gpu_launches": 3 * a.steps * a.encode_batches * (1 + (n_echunks >> 22)),
        }
        line["gpu_launches"] += line["encode"]["gpu_launches"]

    # ---------------- CPU baseline (rank 0, N == 1 only) -------------------------------------------------
    if rank == 0 and world == 1 and not a.skip_cpu_baseline:
        cb, ref_merges = ref_train_sample(pkg, a.corpus_mib, a.ref_sample_mib, a.ref_merges, a.mode)
        # parity on the same sample: our merge list for that slice must equal the reference's
        sample = safe_prefix(pkg.synth_corpus(SEED_TRAIN, a.ref_sample_mib << 20).tobytes(), a.ref_sample_mib << 20)
        ss, se = pkg.split(pkg.patterns()["gpt4"], sample)
        st, so, sw = pkg.dedup(sample, ss, se)
        sm, _, _ = pkg.train(st, so, sw, 256 + a.ref_merges, a.mode, a.engine)
        cb["gpu_equals_reference_on_sample"] = bool(sm.shape == ref_merges.shape and (sm == ref_merges).all())
        line["cpu_baseline"] = cb
        if not a.skip_encode:
            with tempfile.TemporaryDirectory() as td:
                mp = os.path.join(td, "bench.model")
                pkg.write_model(mp, pkg.patterns()["gpt4"], None, merges)
                ecb, sample_text = ref_encode_sample(pkg, mp, a.ref_encode_mib)
            tk2 = pkg.Tokenizer(pkg.patterns()["gpt4"])
            # same model, same sample through our Tokenizer::encode: the id stream must be identical
            import numpy as _np
            enc_ids = pkg.Encoder(merges).encode(sample_text, _np.concatenate(
                [(lambda z: z[0])(pkg.split(pkg.patterns()["gpt4"], sample_text)),
                 _np.asarray([len(sample_text)], _np.uint64)]))
            ecb["gpu_equals_reference_on_sample"] = hashlib.sha256(enc_ids.tobytes()).hexdigest() == ecb["sha256"]
            line["encode"]["cpu_baseline"] = ecb
            tk2.close()

    if clocks:
        line["clocks"] = clocks.summary()
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
