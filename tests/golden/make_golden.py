#!/usr/bin/env python3
"""Generate the golden vectors under tests/golden/ by RUNNING THE REFERENCE.

The reference (justinhj/minbpe-cc, headers compiled verbatim by oracle/Makefile into
oracle/_ref/ref_driver) cannot travel to the GPU box, so its outputs are committed here:
  tests/golden/data/      inputs (copied reference data files + strings written for this repo)
  tests/golden/models/    .model / .model.vocab written by Tokenizer::save   (Tokenizer.h:875-926)
  tests/golden/enc/       raw little-endian u32 streams from Tokenizer::encode (Tokenizer.h:653-722)
  tests/golden/manifest.json   case table + SHA-256 of every artefact + reference CPU seconds

Usage:  make -C oracle ref && python tests/golden/make_golden.py [--only NAME_SUBSTR] [--skip-slow]
"""
import argparse
import hashlib
import json
import os
import re
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "ref_driver")
DATA = os.path.join(HERE, "data")

KEEP_ENC_MAX = 512 * 1024

# Small adversarial strings written for this repo (the reference's own list lives in
# code/examples/train.cpp:17-30; these exercise the same things: a==b runs, exhaustion,
# multi-byte UTF-8, emoji + ZWNJ, digits, contractions, CR/LF and trailing whitespace).
STRINGS = {
    "exhaust": "abcdebce",
    "runs_a": "aaaa",
    "runs_b": "aaabdaaabac",
    "runs_c": "aaaaaaa aaaaaa aaaaa aaaa aaa aa a abababab ababa aabbaabb",
    "kat": "abcbcde",
    "unicode": "Ｆｕｌｌｗｉｄｔｈ ｔｅｘｔ! 🅣🅔🅢🅣 ‽ 🇩‌🇪‌🇫 naïve café 😄 “quotes” — dashes… 안녕하세요 世界 مرحبا "
    "Привет мир! 30 years, 1234567 items, 3.14159; we'll they're I'M DON'T x'S",
    "hello": "hello world!!!? (안녕하세요!) lol123 😉",
    "ws": "line one\r\nline two\n\n\n   indented\t\ttabs  \n trailing   \n nbsp ls　ideographic  end   ",
    "prose": "But tokenizers can be abstruse plus we know we are still finding the whole thing mysterious, "
    "the theory and the thing and then the other thing; there there, their theme then.",
}

# (name, input file, vocab, encoder, mode, special file or None, write_vocab, slow)
TRAIN = [
    ("ts512_gpt4_first", "taylorswift.txt", 512, "gpt4", "first", None, True, False),
    ("ts512_gpt4_lexical", "taylorswift.txt", 512, "gpt4", "lexical", None, True, False),
    ("ts512_gpt4_first_special", "taylorswift.txt", 512, "gpt4", "first", "special1.txt", False, False),
    ("ts512_gpt2_lexical", "taylorswift.txt", 512, "gpt2", "lexical", None, False, False),
    ("ts400_basic_first", "taylorswift.txt", 400, "basic", "first", None, False, False),
    ("sample512_gpt4_first", "sample.txt", 512, "gpt4", "first", None, True, False),
    ("sample512_gpt4_lexical", "sample.txt", 512, "gpt4", "lexical", None, True, False),
    ("sample700_gpt2_first", "sample.txt", 700, "gpt2", "first", None, False, False),
    ("sample400_basic_lexical", "sample.txt", 400, "basic", "lexical", None, False, False),
    ("shk512_basic_lexical", "shakespeare.txt", 512, "basic", "lexical", None, False, False),
    ("shk4096_gpt4_lexical_special", "shakespeare.txt", 4096, "gpt4", "lexical", "special1.txt", False, False),
    ("shk4096_gpt4_first_special", "shakespeare.txt", 4096, "gpt4", "first", "special1.txt", False, True),
]
for _k in STRINGS:
    for _enc in ("basic", "gpt4"):
        for _mode in ("first", "lexical"):
            TRAIN.append((f"str_{_k}_{_enc}_{_mode}", f"str_{_k}.txt", 300, _enc, _mode, None, False, False))

# (name, input file, model case name)
ENCODE = [
    ("sample__shk512_basic_lexical", "sample.txt", "shk512_basic_lexical"),
    ("specialtokensample__ts512_gpt4_first_special", "specialtokensample.txt", "ts512_gpt4_first_special"),
    ("taylorswift__ts512_gpt4_first", "taylorswift.txt", "ts512_gpt4_first"),
    ("shakespeare__ts512_gpt4_lexical", "shakespeare.txt", "ts512_gpt4_lexical"),
    ("shakespeare__shk4096_gpt4_lexical_special", "shakespeare.txt", "shk4096_gpt4_lexical_special"),
    ("sample__sample512_gpt4_lexical", "sample.txt", "sample512_gpt4_lexical"),
    ("sample__ts512_gpt2_lexical", "sample.txt", "ts512_gpt2_lexical"),
    ("taylorswift__ts400_basic_first", "taylorswift.txt", "ts400_basic_first"),
    ("str_unicode__sample512_gpt4_first", "str_unicode.txt", "sample512_gpt4_first"),
    ("str_ws__ts512_gpt4_lexical", "str_ws.txt", "ts512_gpt4_lexical"),
    ("str_runs_c__str_runs_c_basic_first", "str_runs_c.txt", "str_runs_c_basic_first"),
    ("str_runs_c__str_runs_c_gpt4_lexical", "str_runs_c.txt", "str_runs_c_gpt4_lexical"),
    ("str_exhaust__str_exhaust_basic_lexical", "str_exhaust.txt", "str_exhaust_basic_lexical"),
]


def sha(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


def run(args):
    t0 = time.time()
    p = subprocess.run([REF] + args, capture_output=True, text=True)
    wall = time.time() - t0
    m = re.search(r"REF_TIME_S ([0-9.eE+-]+)", p.stderr)
    return p.returncode, (float(m.group(1)) if m else None), wall, p


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-slow", action="store_true")
    a = ap.parse_args()
    if not os.path.exists(REF):
        sys.exit("build the reference first: make -C oracle ref")
    for d in ("models", "enc"):
        os.makedirs(os.path.join(HERE, d), exist_ok=True)
    for k, s in STRINGS.items():
        with open(os.path.join(DATA, f"str_{k}.txt"), "wb") as f:
            f.write(s.encode("utf-8"))

    mpath = os.path.join(HERE, "manifest.json")
    manifest = json.load(open(mpath)) if os.path.exists(mpath) else {"train": {}, "encode": {}}

    for name, inp, vocab, enc, mode, special, wv, slow in TRAIN:
        if a.only and a.only not in name:
            continue
        if slow and a.skip_slow:
            continue
        model = os.path.join(HERE, "models", name + ".model")
        args = ["train", os.path.join(DATA, inp), model, str(vocab), enc, mode]
        if special:
            args += ["-s", os.path.join(DATA, special)]
        if wv:
            args += ["-w"]
        rc, secs, wall, p = run(args)
        ent = {"input": inp, "vocab_size": vocab, "encoder": enc, "mode": mode, "special": special,
               "write_vocab": wv, "rc": rc, "ref_train_s": secs}
        if rc == 0:
            ent["model_sha256"] = sha(model)
            with open(model) as f:
                ent["n_merge_lines"] = len(f.read().split("\n")) - 4 - (5 if special else 0)
            if wv:
                ent["vocab_sha256"] = sha(model + ".vocab")
        else:  # aborts/asserts in the reference are recorded, not golden
            ent["stderr_tail"] = p.stderr[-200:]
            if os.path.exists(model):
                os.remove(model)
        manifest["train"][name] = ent
        print(name, rc, secs, flush=True)

    for name, inp, mcase in ENCODE:
        if a.only and a.only not in name:
            continue
        model = os.path.join(HERE, "models", mcase + ".model")
        if not os.path.exists(model):
            continue
        out = os.path.join(HERE, "enc", name + ".enc")
        rc, secs, wall, p = run(["encode", os.path.join(DATA, inp), model, out])
        ent = {"input": inp, "model": mcase, "rc": rc, "ref_encode_s": secs}
        if rc == 0:
            ent["enc_sha256"] = sha(out)
            ent["n_tokens"] = os.path.getsize(out) // 4
            dec = out + ".dec"
            rc2, dsecs, _, _ = run(["decode", out, model, dec])
            ent["decode_roundtrip"] = (rc2 == 0 and open(dec, "rb").read() == open(os.path.join(DATA, inp), "rb").read())
            os.remove(dec)
            if os.path.getsize(out) > KEEP_ENC_MAX:  # big streams are pinned by SHA-256 only
                os.remove(out)
                ent["enc_file"] = None
            else:
                ent["enc_file"] = "enc/" + name + ".enc"
        manifest["encode"][name] = ent
        print(name, rc, secs, flush=True)

    with open(mpath, "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
