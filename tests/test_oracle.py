"""CPU: pins the oracle (oracle/bpe_oracle.c) to the reference.

 * known-answer tests restated from the reference's own suite, code/test/test.cpp:15-106 (PairCount tie-breaks)
   and :136-186 (3-step merge trace on "abcbcde");
 * every golden .model / .enc under tests/golden/, which were produced by the reference itself compiled here
   (oracle/_ref/ref_driver, tests/golden/make_golden.py).
"""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_data


# ---- code/test/test.cpp:15-48 ---------------------------------------------------------------------------
def test_paircount_insert_order_same_rank(oracle):
    top, n = oracle.paircount_top([(10, 20, 1), (30, 40, 1)], "first")
    assert n == 2 and top[:2] == (10, 20)


def test_paircount_add_and_count(oracle):
    assert oracle.paircount_top([(1, 2, 1)], "first")[1] == 1
    assert oracle.paircount_top([(1, 2, 1), (1, 2, 1)], "first")[1] == 1
    assert oracle.paircount_top([(1, 2, 1), (1, 2, 1), (2, 3, 1)], "first")[1] == 2


# ---- code/test/test.cpp:50-80 ---------------------------------------------------------------------------
def test_paircount_insert_order_most_frequent(oracle):
    ops = []
    assert oracle.paircount_top(ops, "first")[0] is None
    ops += [(1, 2, 1)]
    assert oracle.paircount_top(ops, "first")[0][:2] == (1, 2)
    ops += [(1, 2, 1), (2, 3, 1)]
    assert oracle.paircount_top(ops, "first")[0][:2] == (1, 2)
    ops += [(2, 3, 1), (2, 3, 1)]
    assert oracle.paircount_top(ops, "first")[0][:2] == (2, 3)
    ops += [(1, 2, 1)]  # 3 vs 3: the one inserted first wins
    assert oracle.paircount_top(ops, "first")[0][:2] == (1, 2)
    ops += [(1, 2, 1)]
    assert oracle.paircount_top(ops, "first")[0][:2] == (1, 2)


# ---- code/test/test.cpp:82-106 --------------------------------------------------------------------------
def test_paircount_lexical_most_frequent(oracle):
    ops = []
    assert oracle.paircount_top(ops, "lexical")[0] is None
    ops += [(1, 2, 1)]
    assert oracle.paircount_top(ops, "lexical")[0][:2] == (1, 2)
    ops += [(1, 2, 1), (2, 3, 1)]
    assert oracle.paircount_top(ops, "lexical")[0][:2] == (1, 2)
    ops += [(2, 3, 1)]  # 2 vs 2: smaller pair wins
    assert oracle.paircount_top(ops, "lexical")[0][:2] == (1, 2)
    ops += [(0, 1, 3)]
    assert oracle.paircount_top(ops, "lexical")[0][:2] == (0, 1)


# ---- code/test/test.cpp:136-186 -------------------------------------------------------------------------
@pytest.mark.parametrize("impl", ["rescan", "indexed"])
def test_training_trace_abcbcde(oracle, impl):
    t, o, w = oracle.flatten([b"abcbcde"], dedup=False)
    m, c = oracle.train(t, o, w, 259, "first", impl=impl)
    assert m.tolist() == [[98, 99], [97, 256], [257, 256]]
    assert c.tolist() == [2, 1, 1]


# ---- goldens from the reference itself ------------------------------------------------------------------
def _train_params(manifest_path=os.path.join(GOLDEN, "manifest.json")):
    man = json.load(open(manifest_path))
    return sorted(k for k, e in man["train"].items() if e["rc"] == 0)


@pytest.mark.parametrize("name", _train_params())
def test_train_matches_reference_model(oracle, manifest, name):
    e = manifest["train"][name]
    pattern, specials, gm = oracle.read_model(os.path.join(GOLDEN, "models", name + ".model"))
    assert pattern == oracle.PATTERNS[e["encoder"]]
    text = golden_data(e["input"])
    chunks = oracle.chunks_of(text, e["encoder"])
    big = len(text) > 500_000
    for dedup in (True, False):
        t, o, w = oracle.flatten(chunks, dedup)
        impls = ["indexed"]
        # the literal per-merge walk is quadratic: keep it to the sizes it finishes in seconds
        if not big or (dedup and e["mode"] == "lexical"):
            impls.append("rescan")
        if big and not dedup:
            continue
        for impl in impls:
            m, c = oracle.train(t, o, w, e["vocab_size"], e["mode"], impl=impl)
            assert m.shape == gm.shape and (m == gm).all(), (name, dedup, impl)
    # the writer restatement reproduces the golden bytes
    assert oracle.model_bytes(pattern, specials, gm) == open(os.path.join(GOLDEN, "models", name + ".model"), "rb").read()
    if e.get("write_vocab"):
        assert oracle.vocab_bytes(gm) == open(os.path.join(GOLDEN, "models", name + ".model.vocab"), "rb").read()
    assert hashlib.sha256(oracle.model_bytes(pattern, specials, gm)).hexdigest() == e["model_sha256"]


def _encode_params():
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    return sorted(k for k, e in man["encode"].items() if e["rc"] == 0)


@pytest.mark.parametrize("name", _encode_params())
def test_encode_matches_reference_stream(oracle, manifest, name):
    e = manifest["encode"][name]
    pattern, specials, merges = oracle.read_model(os.path.join(GOLDEN, "models", e["model"] + ".model"))
    text = golden_data(e["input"])
    ids = oracle.encode_text(text, pattern, specials, merges)
    assert len(ids) == e["n_tokens"]
    assert hashlib.sha256(ids.tobytes()).hexdigest() == e["enc_sha256"]
    if e.get("enc_file"):
        assert ids.tobytes() == open(os.path.join(GOLDEN, e["enc_file"]), "rb").read()
    assert oracle.decode(merges, ids, {i: t for t, i in specials}) == text  # endtoend-test.sh round trip


def test_decode_skips_invalid_and_prefers_specials(oracle):
    merges = np.asarray([[97, 98]], np.uint32)
    assert oracle.decode(merges, [97, 256, 999999, 98], {}) == b"aabb"
    assert oracle.decode(merges, [97, 256, 65], {256: b"<S>", 65: b"<A>"}) == b"a<S><A>"


def test_exhaustion_modes_differ(oracle, manifest):
    # first stops early, lexical repeats the smallest zero-count pair
    assert manifest["train"]["str_exhaust_basic_first"]["n_merge_lines"] == 6
    assert manifest["train"]["str_exhaust_basic_lexical"]["n_merge_lines"] == 44
