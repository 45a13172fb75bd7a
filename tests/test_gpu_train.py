"""GPU: parity of the train merge loop (through the C ABI) with the reference goldens and the oracle."""
import hashlib
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden_data

pytestmark = pytest.mark.gpu


def _cases(pred=lambda k, e: True):
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    return sorted(k for k, e in man["train"].items() if e["rc"] == 0 and pred(k, e))


# ---- PairCount seam: the reference's own known-answer tests (code/test/test.cpp:15-106) on the device table ----
def test_paircount_insert_order_same_rank(pkg):
    pc = pkg.PairCount("first")
    pc.create_or_modify_pair(10, 20, 1)
    assert pc.get_count() == 1
    pc.create_or_modify_pair(30, 40, 1)
    assert pc.get_count() == 2
    assert pc.get_pair((10, 20)) == 1 and pc.get_pair((30, 40)) == 1
    assert pc.get_pair((7, 7)) is None


def test_paircount_add_and_count(pkg):
    pc = pkg.PairCount("first")
    assert pc.get_count() == 0
    pc.create_or_modify_pair(1, 2, 1)
    assert pc.get_count() == 1
    pc.create_or_modify_pair(1, 2, 1)
    assert pc.get_count() == 1
    pc.create_or_modify_pair(2, 3, 1)
    assert pc.get_count() == 2


def test_paircount_insert_order_most_frequent(pkg):
    pc = pkg.PairCount("first")
    assert pc.get_top_pair_count() is None
    pc.create_or_modify_pair(1, 2, 1)
    assert pc.get_top_pair_count() == (1, 2)
    pc.create_or_modify_pair(1, 2, 1)
    pc.create_or_modify_pair(2, 3, 1)
    assert pc.get_top_pair_count() == (1, 2)
    pc.create_or_modify_pair(2, 3, 1)
    pc.create_or_modify_pair(2, 3, 1)
    assert pc.get_top_pair_count() == (2, 3)
    pc.create_or_modify_pair(1, 2, 1)  # 3 vs 3: inserted first wins
    assert pc.get_top_pair_count() == (1, 2)
    pc.create_or_modify_pair(1, 2, 1)
    assert pc.get_top_pair_count() == (1, 2)


def test_paircount_lexical_most_frequent(pkg):
    pc = pkg.PairCount("lexical")
    assert pc.get_top_pair_count() is None
    pc.create_or_modify_pair(1, 2, 1)
    assert pc.get_top_pair_count() == (1, 2)
    pc.create_or_modify_pair(1, 2, 1)
    pc.create_or_modify_pair(2, 3, 1)
    assert pc.get_top_pair_count() == (1, 2)
    pc.create_or_modify_pair(2, 3, 1)  # 2 vs 2: smaller pair wins
    assert pc.get_top_pair_count() == (1, 2)
    pc.create_or_modify_pair(0, 1, 3)
    assert pc.get_top_pair_count() == (0, 1)


@pytest.mark.parametrize("mode", ["first", "lexical"])
def test_paircount_batched_random_vs_oracle(pkg, oracle, mode):
    rng = np.random.default_rng(3)
    ops = [(int(a), int(b), int(d)) for a, b, d in
           zip(rng.integers(0, 40, 5000), rng.integers(0, 40, 5000), rng.integers(1, 4, 5000))]
    pc = pkg.PairCount(mode)
    for i in range(0, len(ops), 700):  # several batches: table growth + insertion order across calls
        pc.add(ops[i:i + 700])
    top, n = oracle.paircount_top(ops, mode)
    assert pc.get_count() == n and pc.get_top_pair_count() == top[:2] and pc.get_pair(top[:2]) == top[2]


# ---- merge loop: code/test/test.cpp:136-186 -----------------------------------------------------------------
@pytest.mark.parametrize("engine", ["stepwise", "persistent"])
def test_training_trace_abcbcde(pkg, engine):
    t = np.frombuffer(b"abcbcde", np.uint8).astype(np.uint32)
    m, c, _ = pkg.train(t, np.asarray([0, 7], np.uint64), None, 259, "first", engine)
    assert m.tolist() == [[98, 99], [97, 256], [257, 256]] and c.tolist() == [2, 1, 1]


# ---- merge loop vs. every golden model, both engines -------------------------------------------------------
@pytest.mark.parametrize("engine", ["persistent", "stepwise"])
@pytest.mark.parametrize("name", _cases())
def test_merges_match_reference(pkg, oracle, manifest, name, engine):
    e = manifest["train"][name]
    if engine == "stepwise" and e["vocab_size"] > 1024:
        pytest.skip("stepwise engine is the slow debug path; covered on the smaller configs")
    _, _, gm = oracle.read_model(os.path.join(GOLDEN, "models", name + ".model"))
    text = golden_data(e["input"])
    s, en = pkg.split(pkg.patterns()[e["encoder"]], text)
    tok, off, w = pkg.dedup(text, s, en)
    m, c, st = pkg.train(tok, off, w, e["vocab_size"], e["mode"], engine)
    assert m.shape == gm.shape and (m == gm).all(), (name, engine, st)
    _, oc = oracle.train(tok, off, w, e["vocab_size"], e["mode"])
    assert (c == oc).all()


def test_without_dedup_is_the_reference_layout(pkg, oracle, manifest):
    """Every regex match its own chunk, weight 1: the reference's own data layout."""
    for name in ("ts512_gpt4_first", "ts512_gpt4_lexical"):
        e = manifest["train"][name]
        _, _, gm = oracle.read_model(os.path.join(GOLDEN, "models", name + ".model"))
        text = golden_data(e["input"])
        t, o, w = oracle.flatten(oracle.chunks_of(text, e["encoder"]), dedup=False)
        m, _, _ = pkg.train(t, o, None, e["vocab_size"], e["mode"])
        assert (m == gm).all()


def test_small_tables_force_growth_and_rebuilds(pkg, oracle, monkeypatch):
    text = golden_data("taylorswift.txt")
    t, o, w = oracle.flatten(oracle.chunks_of(text, "gpt4"), True)
    om, oc = oracle.train(t, o, w, 1500, "lexical")
    monkeypatch.setenv("MBPE_INIT_SLOTS", "8192")
    monkeypatch.setenv("MBPE_CAND_WANT", "8")
    monkeypatch.setenv("MBPE_BIG_LIMIT", "64")
    m, c, st = pkg.train(t, o, w, 1500, "lexical", "persistent")
    assert (m == om).all() and (c == oc).all()
    assert st["n_grows"] >= 1 and st["n_rebuilds"] >= 2 and st["n_big_merges"] >= 10


@pytest.mark.parametrize("mode", ["first", "lexical"])
def test_random_small_alphabets(pkg, oracle, mode):
    rng = np.random.default_rng(5)
    for trial in range(25):
        k = int(rng.integers(1, 4))
        chunks = [bytes(rng.integers(97, 97 + k, int(rng.integers(1, 40))).astype(np.uint8))
                  for _ in range(int(rng.integers(1, 60)))]
        t, o, w = oracle.flatten(chunks, True)
        vocab = 256 + int(rng.integers(1, 60))
        om, oc = oracle.train(t, o, w, vocab, mode, impl="rescan")
        m, c, _ = pkg.train(t, o, w, vocab, mode)
        assert m.shape == om.shape and (m == om).all() and (c == oc).all(), (trial, chunks)


def test_empty_pairless_and_invalid_inputs(pkg):
    m, c, _ = pkg.train(np.zeros(0, np.uint32), np.asarray([0], np.uint64), None, 300, "lexical")
    assert len(m) == 0
    m, c, _ = pkg.train(np.asarray([97, 98, 99], np.uint32), np.asarray([0, 1, 2, 3], np.uint64), None, 300, "first")
    assert len(m) == 0
    with pytest.raises(pkg.MbpeError):
        pkg.train(np.asarray([97, 300], np.uint32), np.asarray([0, 2], np.uint64), None, 300, "first")
    with pytest.raises(pkg.MbpeError):
        pkg.train(np.asarray([97, 98], np.uint32), np.asarray([0, 2], np.uint64), None, 255, "first")


def test_one_long_chunk_basic_encoder(pkg, oracle):
    """encoder 'basic': the whole text is one chunk (Tokenizer.h:541-544)."""
    text = golden_data("taylorswift.txt")
    t = np.frombuffer(text, np.uint8).astype(np.uint32)
    o = np.asarray([0, len(t)], np.uint64)
    w = np.ones(1, np.uint32)
    for mode in ("first", "lexical"):
        om, oc = oracle.train(t, o, w, 400, mode)
        m, c, _ = pkg.train(t, o, w, 400, mode)
        assert (m == om).all() and (c == oc).all()


# ---- synthetic Zipfian corpus, sizes the oracle finishes in seconds ---------------------------------------
@pytest.mark.parametrize("mode", ["lexical", "first"])
def test_synthetic_corpus_vs_oracle(pkg, oracle, mode):
    text = pkg.synth_corpus(0x5EED0001, 24 << 20).tobytes()
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    tok, off, w = pkg.dedup(text, s, e)
    om, oc = oracle.train(tok, off, w, 256 + 3000, mode)
    m, c, st = pkg.train(tok, off, w, 256 + 3000, mode)
    assert m.shape == om.shape and (m == om).all() and (c == oc).all(), st
    # property that holds at any size: counts never increase along the merge list
    assert (np.diff(c.astype(np.int64)) <= 0).all()


# ---- Tokenizer mirror: train + save == golden .model / .vocab bytes ---------------------------------------
@pytest.mark.parametrize("name", _cases(lambda k, e: e["vocab_size"] <= 4096))
def test_tokenizer_train_save_bytes(pkg, manifest, name, tmp_path):
    e = manifest["train"][name]
    tk = pkg.Tokenizer(pkg.patterns()[e["encoder"]])
    if e["special"]:
        tk.set_special_tokens_from_file(golden_data(e["special"]))
    tk.train(golden_data(e["input"]), e["vocab_size"], e["mode"])
    out = tmp_path / "out.model"
    tk.save(out, write_vocab=bool(e["write_vocab"]))
    assert hashlib.sha256(out.read_bytes()).hexdigest() == e["model_sha256"]
    if e["write_vocab"]:
        assert hashlib.sha256((tmp_path / "out.model.vocab").read_bytes()).hexdigest() == e["vocab_sha256"]


@pytest.mark.parametrize("mode", ["lexical", "first"])
def test_config3_shape_256mib_32k_vocab_vs_oracle(pkg, oracle, mode):
    """BASELINE config 3 at a quarter of its corpus (256 MiB, vocab 32768): merge list and counts equal the oracle's.
    bench.py repeats this diff at the full 1 GiB on every run."""
    text = pkg.synth_corpus(0x5EED0001, 256 << 20).tobytes()
    tok, off, w, n_chunks = pkg.split_dedup(pkg.patterns()["gpt4"], text)
    assert int(w.sum()) == n_chunks
    om, oc = oracle.train(tok, off, w, 32768, mode)
    m, c, st = pkg.train(tok, off, w, 32768, mode)
    assert m.shape == om.shape and (m == om).all() and (c == oc).all(), st


def test_full_32k_vocab_model_equals_the_reference_on_a_16mib_slice(pkg, tmp_path):
    """the un-extrapolated comparison. The compiled reference trained the first 16 MiB (cut at a regex-safe
    point) of the synthetic corpus to the FULL 32768 vocabulary once (44.5 min of one CPU core, profiles/reference_full_16MiB.json);
    its .model is a golden. Tokenizer::train on the same slice must write the same file, byte for byte."""
    import re
    text = pkg.synth_corpus(0x5EED0001, 16 << 20).tobytes()
    lo = len(text) - 65536
    last = None
    for m in re.finditer(rb"\n[\x21-\x7e]", text[lo:]):
        last = m
    text = text[:lo + last.start() + 1]
    assert len(text) == 16777011 and hashlib.sha256(text).hexdigest() == "abf7b4e4880ca7f70bac62d1b0755a313fd0da49adc4a01323cc2c05c223889f"
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
    tk.train(text, 32768, "lexical")
    out = tmp_path / "m.model"
    tk.save(out)
    golden = open(os.path.join(GOLDEN, "models", "synth16m_gpt4_lexical_32768.model"), "rb").read()
    assert hashlib.sha256(golden).hexdigest() == "38103cba13ea420b7077d1baa80125551fa7cbb03d5d5566c5ddf39087a64119"
    assert out.read_bytes() == golden
