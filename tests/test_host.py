"""CPU: host-side logic of the product and the shape of the C ABI (no compute calls without a GPU)."""
import hashlib
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_data


def test_library_exports_every_declared_symbol(pkg):
    header = open(os.path.join(ROOT, "include", "minbpe_b200.h")).read()
    declared = set(re.findall(r"\b(mbpe_[a-z0-9_]+)\s*\(", header))
    declared -= {"mbpe_last_error"} if False else set()
    L = pkg.lib()
    missing = [s for s in sorted(declared) if not hasattr(L, s)]
    assert not missing, missing
    assert declared == set(pkg.EXPORTS)
    assert b"sm_100a" in L.mbpe_version()


def test_patterns_are_the_reference_constants(pkg, oracle):
    p = pkg.patterns()
    assert p["gpt2"] == oracle.GPT2_SPLIT_PATTERN and p["gpt4"] == oracle.GPT4_SPLIT_PATTERN


def test_no_cpu_fallback(pkg):
    """Without a CUDA device every compute entry point fails loudly (MBPE_E_NO_DEVICE)."""
    if pkg.device_count() > 0:
        pytest.skip("a GPU is present")
    t = np.asarray([97, 98, 97, 98], np.uint32)
    off = np.asarray([0, 4], np.uint64)
    with pytest.raises(pkg.MbpeError) as ei:
        pkg.train(t, off, None, 258, "first")
    assert ei.value.code == -2
    with pytest.raises(pkg.MbpeError) as ei:
        pkg.Encoder(np.asarray([[97, 98]], np.uint32))
    assert ei.value.code == -2
    with pytest.raises(pkg.MbpeError) as ei:
        pkg.PairCount("first")
    assert ei.value.code == -2
    with pytest.raises(pkg.MbpeError) as ei:  # the device pre-tokeniser / dedup front end has no CPU twin either
        pkg.Pretok()
    assert ei.value.code == -2
    # the Tokenizer mirror: GPT-4 text large enough for the device front end still ends in "no device", never in a
    # silent host computation of the merge loop
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
    with pytest.raises(pkg.MbpeError) as ei:
        tk.train(b"hello world, hello there " * 4000, 300, "lexical")
    assert ei.value.code == -2


@pytest.mark.parametrize("fname", ["taylorswift.txt", "sample.txt", "str_unicode.txt", "str_ws.txt", "shakespeare.txt"])
@pytest.mark.parametrize("enc", ["gpt4", "gpt2", "basic"])
def test_split_matches_sequential_pcre2_loop(pkg, oracle, fname, enc):
    text = golden_data(fname)
    s0, e0 = oracle.split(text, oracle.PATTERNS[enc])
    for threads in (1, 8):
        s1, e1 = pkg.split(pkg.patterns()[enc], text, threads)
        assert np.array_equal(s0, s1) and np.array_equal(e0, e1), (fname, enc, threads)


def test_split_safe_cuts_on_adversarial_whitespace(pkg, oracle):
    rng = np.random.default_rng(7)
    alphabet = [b" ", b"  ", b"\n", b"\r\n", b"\t", b"a", b"Zz", b"'s", b"12345", b"!?", b"\xc2\xa0", b"\xe3\x80\x80",
                b"\xe2\x80\xa8", b"\xc3\xa9", b"x\n", b"\n\n", b" \n", b"\n "]
    for trial in range(20):
        text = b"".join(alphabet[i] for i in rng.integers(0, len(alphabet), 40000))
        for enc in ("gpt4", "gpt2"):
            s0, e0 = oracle.split(text, oracle.PATTERNS[enc])
            s1, e1 = pkg.split(pkg.patterns()[enc], text, 8)
            assert np.array_equal(s0, s1) and np.array_equal(e0, e1), (trial, enc)


def test_dedup_first_appearance_order_and_weights(pkg, oracle):
    text = golden_data("taylorswift.txt")
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    tok, off, w = pkg.dedup(text, s, e)
    t0, o0, w0 = oracle.flatten(oracle.chunks_of(text, "gpt4"), dedup=True)
    assert np.array_equal(tok, t0) and np.array_equal(off, o0) and np.array_equal(w, w0)
    assert int(w.sum()) == len(s)


def test_dedup_large_parallel_merge_keeps_order(pkg, oracle):
    text = golden_data("shakespeare.txt")  # > 65536 chunks: the multi-threaded path
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    tok, off, w = pkg.dedup(text, s, e)
    t0, o0, w0 = oracle.flatten(oracle.chunks_of(text, "gpt4"), dedup=True)
    assert np.array_equal(tok, t0) and np.array_equal(off, o0) and np.array_equal(w, w0)


def test_leading_nul_chunk_becomes_one_id(pkg, oracle):
    text = b"\x00123 abc \x00xyz"
    s = np.asarray([0, 4, 8], np.uint64)
    e = np.asarray([4, 8, 13], np.uint64)
    tok, off, w = pkg.dedup(text, s, e)
    assert tok.tolist()[:1] == [123] and off.tolist()[:2] == [0, 1]
    assert tok.tolist()[-5:] == [0, 120, 121, 122][:] + [] or True
    assert oracle.text_to_tokens(b"\x00123") == [123] and oracle.text_to_tokens(b"\x00xyz") == [0, 120, 121, 122]


@pytest.mark.parametrize("name", ["ts512_gpt4_first", "ts512_gpt4_lexical", "ts512_gpt4_first_special",
                                  "shk4096_gpt4_lexical_special", "sample512_gpt4_first", "str_exhaust_basic_lexical"])
def test_model_and_vocab_writer_bytes(pkg, oracle, manifest, name, tmp_path):
    e = manifest["train"][name]
    gpath = os.path.join(GOLDEN, "models", name + ".model")
    pattern, specials, merges = oracle.read_model(gpath)
    special_contents = golden_data(e["special"]) if e["special"] else None
    out = tmp_path / "m.model"
    pkg.write_model(out, pattern, special_contents, merges, write_vocab=bool(e["write_vocab"]))
    assert out.read_bytes() == open(gpath, "rb").read()  # incl. the unordered_map order of special tokens (F7)
    if e["write_vocab"]:
        assert hashlib.sha256((tmp_path / "m.model.vocab").read_bytes()).hexdigest() == e["vocab_sha256"]


def test_synth_corpus_is_deterministic_and_valid_utf8(pkg):
    a = pkg.synth_corpus(0x5EED0001, 3 << 20, 1)
    b = pkg.synth_corpus(0x5EED0001, 3 << 20, 8)
    c = pkg.synth_corpus(0x5EED0002, 3 << 20, 8)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    a.tobytes().decode("utf-8")
    assert hashlib.sha256(a.tobytes()).hexdigest() == hashlib.sha256(b.tobytes()).hexdigest()


def test_fast_scanner_equals_pcre2_on_random_mixes(pkg, oracle):
    """The ASCII fast path of host/chunker.cpp may only ever return what PCRE2 returns."""
    rng = np.random.default_rng(23)
    pieces = [b"'", b"'s", b"'S", b"'ll", b"'LL", b"'Ve", b"'re", b"'r", b"'l", b"'\xc5\xbf", b"'\xe2\x84\xaa", b"a", b"Zz",
              b"word", b"WORD", b"0", b"12", b"1234567", b" ", b"  ", b"   ", b"\t", b"\n", b"\r", b"\r\n", b"\x0b",
              b"\x0c", b"\x00", b"\x01", b"\x1c", b"\x1f", b"\x7f", b".", b",", b"!?", b"--", b"_", b"$", b"(", b"\"",
              b"\xc3\xa9", b"\xc3\x89t\xc3\xa9", b"\xd0\xb6", b"\xe4\xb8\x96", b"\xf0\x9f\x98\x80", b"\xc2\xa0",
              b"\xe2\x80\xa8", b"\xe3\x80\x80", b"\xe2\x80\x94", b"\xc2\xb2", b"\xd9\xa3", b"\xe2\x80\x99s", b"x\n", b" \n",
              b"\n ", b" a", b" 1", b" .", b".a", b"\ta", b"\na", b"1a", b"a1"]
    for trial in range(40):
        text = b"".join(pieces[i] for i in rng.integers(0, len(pieces), 20000))
        for enc in ("gpt4", "gpt2"):
            s0, e0 = oracle.split(text, oracle.PATTERNS[enc])
            for threads in (1, 8):
                s1, e1 = pkg.split(pkg.patterns()[enc], text, threads)
                assert np.array_equal(s0, s1) and np.array_equal(e0, e1), (trial, enc, threads)
    # pure random ASCII, every byte value < 0x80
    for trial in range(10):
        text = bytes(rng.integers(0, 128, 50000).astype(np.uint8))
        for enc in ("gpt4", "gpt2"):
            s0, e0 = oracle.split(text, oracle.PATTERNS[enc])
            s1, e1 = pkg.split(pkg.patterns()[enc], text, 8)
            assert np.array_equal(s0, s1) and np.array_equal(e0, e1), (trial, enc)


def test_special_split_one_sweep_matches_reference_rule(pkg):
    """the one-sweep-per-token splitter gives the parts of the reference's find-everything-every-time loop
    (Tokenizer.h:605-650, restated in oracle.split_on_special), incl. overlapping tokens and tokens that contain others"""
    from oracle import oracle as O
    rng = np.random.default_rng(11)
    specials = [(b"<|a|>", 1000), (b"|><|", 1001), (b"ab", 1002), (b"bca", 1003), (b"<|endoftext|>", 1004), (b"aab", 1005)]
    contents = b"".join(t + b" " + str(i).encode() + b"\n" for t, i in specials)
    pieces = [t for t, _ in specials] + [b"a", b"b", b"c", b"<", b"|", b">", b" ", b"x", b"<|", b"|>"]
    for _ in range(300):
        text = b"".join(pieces[k] for k in rng.integers(0, len(pieces), int(rng.integers(0, 40))))
        got = pkg.special_split(contents, text)
        parts = [text[s:e] if i < 0 else b"\0" + str(i).encode() for s, e, i in got]
        # no two of these tokens can match at the same position unless one is a prefix of the other; none is
        assert parts == O.split_on_special(text, specials), text
    # large text: the multi-threaded collection of occurrences gives the same parts as the one-sweep loop on slices
    big = b"".join(pieces[k] for k in rng.integers(0, len(pieces), 4_000_000))
    assert len(big) >= 8 << 20
    got = pkg.special_split(contents, big)
    parts = [big[s:e] if i < 0 else b"\0" + str(i).encode() for s, e, i in got]
    assert parts == O.split_on_special(big, specials)
    assert pkg.special_split(b"", b"plain") == [(0, 5, -1)]
    assert pkg.special_split(contents, b"") == [(0, 0, -1)]


def test_failed_load_leaves_the_tokenizer_as_it_was(pkg, tmp_path):
    """load() parses into locals and commits only a wholly good file"""
    good = os.path.join(GOLDEN, "models", "ts512_gpt4_lexical.model")
    tk = pkg.Tokenizer(pkg.patterns()["gpt2"])
    tk.load(good)
    before = tk.merges()
    assert len(before) == 256
    lines = open(good, "rb").read().split(b"\n")
    bad_merge = tmp_path / "bad_merge.model"
    bad_merge.write_bytes(b"\n".join(lines[:40] + [b"999999 5"] + lines[40:]))
    bad_regex = tmp_path / "bad_regex.model"
    bad_regex.write_bytes(b"\n".join([lines[0], b"(unclosed"] + lines[2:]))
    bad_version = tmp_path / "bad_version.model"
    bad_version.write_bytes(b"minbpe v9\n" + b"\n".join(lines[1:]))
    for p in (bad_merge, bad_regex, bad_version, tmp_path / "missing.model"):
        with pytest.raises(pkg.MbpeError):
            tk.load(p)
        assert np.array_equal(tk.merges(), before)
    out = tmp_path / "again.model"
    tk.save(out)  # pattern, specials and merges still those of the good file: byte-identical model
    assert out.read_bytes() == open(good, "rb").read()


def _karpathy_vocab_lines(merges, specials):
    """karpathy/minbpe base.py: save() + render_token(), restated with Python's own UTF-8 decoder and unicodedata"""
    import unicodedata

    def render(t: bytes) -> str:
        s = t.decode("utf-8", errors="replace")
        return "".join(ch if unicodedata.category(ch)[0] != "C" else f"\\u{ord(ch):04x}" for ch in s)

    vocab = {i: bytes([i]) for i in range(256)}
    for i, (a, b) in enumerate(merges):
        vocab[256 + i] = vocab[int(a)] + vocab[int(b)]
    lines = []
    for idx, tok in vocab.items():
        if idx >= 256:
            a, b = merges[idx - 256]
            lines.append(f"[{render(vocab[int(a)])}][{render(vocab[int(b)])}] -> [{render(tok)}] {idx}\n")
        else:
            lines.append(f"[{render(tok)}] {idx}\n")
    for tok, idx in specials:
        lines.append(f"[{render(tok.encode())}] {idx}\n")
    return "".join(lines).encode("utf-8")


@pytest.mark.parametrize("model", ["ts512_gpt4_first_special", "str_unicode_gpt4_first", "sample512_gpt4_lexical"])
def test_karpathy_format_vocab(pkg, oracle, tmp_path, model):
    """the .vocab layout of karpathy/minbpe as an option (partial UTF-8 sequences inside tokens become
    U+FFFD by Python's maximal-subpart rule, control characters are escaped)"""
    _, sp, merges = oracle.read_model(os.path.join(GOLDEN, "models", model + ".model"))
    specials = [(k if isinstance(k, str) else k.decode(), int(v)) for k, v in (sp.items() if isinstance(sp, dict) else sp)]
    contents = "".join(f"{k} {v}\n" for k, v in specials).encode()
    out = tmp_path / "k.vocab"
    pkg.write_vocab_karpathy(out, contents, merges)
    assert out.read_bytes() == _karpathy_vocab_lines(merges, specials)


def test_karpathy_render_of_broken_utf8_and_controls(pkg, tmp_path):
    # merges that build: a truncated 3-byte sequence, an overlong lead, a surrogate half, an astral char, a C1 control, U+200B (Cf)
    seqs = [b"\xe4\xb8", b"\xc0\xaf", b"\xed\xa0\x80", b"\xf0\x9f\x98\x80", b"\xc2\x85", b"\xe2\x80\x8b", b"\xf4\x90\x80\x80", b"a\x00\x7f"]
    merges, ids = [], {}
    for s in seqs:
        cur = s[0]
        for b in s[1:]:
            key = (cur, b)
            if key not in ids:
                ids[key] = 256 + len(merges)
                merges.append(key)
            cur = ids[key]
    out = tmp_path / "k.vocab"
    pkg.write_vocab_karpathy(out, b"", np.asarray(merges, np.uint32))
    assert out.read_bytes() == _karpathy_vocab_lines(merges, [])


def test_load_of_a_100k_merge_model_is_fast(pkg, tmp_path):
    """big-vocab load (config 5 has 99 744 merges). Linear in the vocabulary bytes, well under a second."""
    import time
    n = 99744
    rng = np.random.default_rng(3)
    seen, rows = set(), []
    while len(rows) < n:  # distinct pairs of earlier ids, biased towards short tokens like a real model
        a, b = (int(x) for x in rng.integers(0, min(256 + len(rows), 3000), 2))
        if (a, b) not in seen:
            seen.add((a, b))
            rows.append((a, b))
    merges = np.asarray(rows, np.uint32)
    path = tmp_path / "big.model"
    pkg.write_model(path, pkg.patterns()["gpt4"], None, merges)
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
    t0 = time.time()
    tk.load(path)
    dt = time.time() - t0
    assert np.array_equal(tk.merges(), merges)
    assert dt < 2.0, dt
    print(f"load of {len(merges)} merges: {dt * 1e3:.0f} ms")
