"""The hand-written matcher of the GPT-4 split pattern (csrc/pretok_core.cuh) against PCRE2.

CPU tier: the same pretok_window() the CUDA kernel runs, executed window by window on the host
(tests/emu/pretok_emu.cpp), must mark exactly the chunk starts that the reference's pcre2_match loop produces
(mbpe_split, Tokenizer.h:506-540) -- for every window size, on the fixtures and on fuzzed Unicode."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

EMU_SRC = os.path.join(ROOT, "tests", "emu", "pretok_emu.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libpretok_emu.so")


@pytest.fixture(scope="module")
def emu(pkg):
    deps = [EMU_SRC, os.path.join(ROOT, "minbpe-cc_b200", "csrc", "pretok_core.cuh")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-x", "c++", "-o", EMU_LIB, EMU_SRC])
    L = C.CDLL(EMU_LIB)
    L.emu_pretok.restype = C.c_uint32
    table = pkg.pretok_class_table()

    def run(text: bytes, window=32, max_crawl=1 << 40, order=0, kind="gpt4"):
        L.emu_pretok_kind({"gpt4": 0, "gpt2": 1}[kind])
        buf = np.frombuffer(text, np.uint8) if len(text) else np.zeros(1, np.uint8)
        marks = np.zeros(max(len(text), 1), np.uint8)
        err = L.emu_pretok(buf.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(len(text)),
                           table.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_uint64(window), C.c_uint64(max_crawl), order,
                           marks.ctypes.data_as(C.POINTER(C.c_uint8)))
        return np.flatnonzero(marks[:len(text)]).astype(np.uint64), err

    return run


@pytest.fixture(scope="module")
def pcre2_starts(pkg):
    return lambda text, kind="gpt4": _pcre2_starts(pkg, text, kind)


def _pcre2_starts(pkg, text: bytes, kind="gpt4"):
    s, e = pkg.split(pkg.patterns()[kind], text, 1)
    assert len(s) == 0 or (s[0] == 0 and e[-1] == len(text) and np.array_equal(s[1:], e[:-1])), "chunks tile the text"
    return s


def test_class_table_matches_known_points(pkg):
    t = pkg.pretok_class_table()
    cls = lambda cp: (int(t[cp >> 2]) >> ((cp & 3) * 2)) & 3
    assert [cls(ord(c)) for c in "aZ09 \n\t!_'"] == [1, 1, 2, 2, 3, 3, 3, 0, 0, 0]
    assert cls(0xE9) == 1 and cls(0x4E2D) == 1 and cls(0x0416) == 1      # e-acute, CJK, Cyrillic: letters
    assert cls(0x0663) == 2 and cls(0xBD) == 2 and cls(0x2167) == 2      # Arabic-Indic digit, 1/2, roman numeral: numbers
    assert cls(0xA0) == 3 and cls(0x3000) == 3 and cls(0x2028) == 3 and cls(0x85) == 3  # Unicode spaces
    assert cls(0x1F600) == 0 and cls(0x3002) == 0 and cls(0x200B) == 0   # emoji, CJK full stop, zero-width space
    assert cls(0x17F) == 1


@pytest.mark.parametrize("name", ["taylorswift.txt", "shakespeare.txt"])
@pytest.mark.parametrize("window", [32, 64, 7])
def test_fixtures(emu, name, window, pcre2_starts):
    path = os.path.join(ROOT, "tests", "golden", "data", name)
    text = open(path, "rb").read()
    if name == "shakespeare.txt" and window == 7:
        text = text[:200000]
    got, err = emu(text, window)
    assert err == 0
    assert np.array_equal(got, pcre2_starts(text))


# every class, the caseless partners of alternative 1 (long s U+017F, Kelvin U+212A), Unicode blanks, combining
# marks, astral code points, control characters -- and the multi-character shapes the alternatives care about
ALPHABET = (list(" \n\r\t'.,!?-_0123456789") * 3 + list("abcdefghijklmnopqrstuvwxyzSDMTLVRE") * 2 +
            ["é", "ß", "ſ", "K", "Ж", "中", "。", " ", "　", " ", " ",
             "", "٣", "½", "Ⅷ", "\U0001f600", "\U00020000", "​", "́", "\v", "\f", "\x00",
             "\x1c", "'s", "'S", "'ſ", "'ll", "'LL", "'Ve", "'re", "'t", " '", "  ", "\n\n", "\r\n", " \n ", "123456",
             "　　"])


@pytest.mark.parametrize("seed", range(6))
def test_fuzz_unicode(emu, seed, pcre2_starts):
    rng = np.random.default_rng(1000 + seed)
    for _ in range(150):
        n = int(rng.integers(1, 200))
        text = "".join(ALPHABET[i] for i in rng.integers(0, len(ALPHABET), n)).encode("utf-8")
        want = pcre2_starts(text)
        for window in (int(rng.integers(1, 12)), 32):
            for order in (0, 1):
                got, err = emu(text, window, order=order)
                assert err == 0, text
                assert np.array_equal(got, want), (text, window, got.tolist(), want.tolist())


@pytest.mark.parametrize("seed", range(3))
def test_gpt2_pattern_fuzz_and_fixture(emu, seed, pcre2_starts):
    rng = np.random.default_rng(2000 + seed)
    for _ in range(150):
        text = "".join(ALPHABET[i] for i in rng.integers(0, len(ALPHABET), int(rng.integers(1, 200)))).encode("utf-8")
        want = pcre2_starts(text, "gpt2")
        for window in (int(rng.integers(1, 12)), 32):
            got, err = emu(text, window, order=seed & 1, kind="gpt2")
            assert err == 0 and np.array_equal(got, want), (text, window, got.tolist(), want.tolist())
    if seed == 0:
        text = open(os.path.join(ROOT, "tests", "golden", "data", "taylorswift.txt"), "rb").read()
        got, err = emu(text, 32, kind="gpt2")
        assert err == 0 and np.array_equal(got, pcre2_starts(text, "gpt2"))


def _expected_starts_with_specials(pkg, contents: bytes, text: bytes, kind="gpt4"):
    """chunk starts the reference produces: special tokens cut out (Tokenizer.h:605-650), every ordinary part split on
    its own (:664-704), each special token one chunk"""
    starts, sp = [], []
    for s0, e0, sid in pkg.special_split(contents, text):
        if sid >= 0:
            starts.append(s0)
            sp.append((s0, e0))
        elif e0 > s0:
            ps, _ = pkg.split(pkg.patterns()[kind], text[s0:e0], 1)
            starts.extend(int(x) + s0 for x in ps)
    return np.asarray(sorted(starts), np.uint64), sp


@pytest.mark.parametrize("kind", ["gpt4", "gpt2"])
def test_subjects_cut_by_special_tokens(pkg, emu, kind):
    L = C.CDLL(EMU_LIB)
    L.emu_pretok_parts.restype = C.c_uint32
    L.emu_pretok_kind({"gpt4": 0, "gpt2": 1}[kind])
    table = pkg.pretok_class_table()
    specials = [(b"<|endoftext|>", 100257), (b"<|x|>", 100258), (b"ab", 7)]
    contents = b"".join(t + b" " + str(i).encode() + b"\n" for t, i in specials)
    pieces = [t for t, _ in specials] * 2 + [b"hello", b" world", b"  ", b"\n", b" \n ", b"a", b"b", b"'s", b"123", b"!!", b"<|",
                                             b"|>", " é".encode(), "中文".encode(), b" "]
    rng = np.random.default_rng(31)
    P = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    for it in range(400):
        text = b"".join(pieces[k] for k in rng.integers(0, len(pieces), int(rng.integers(0, 40))))
        if not text:
            continue
        want, sp = _expected_starts_with_specials(pkg, contents, text, kind)
        sb = np.asarray([a for a, _ in sp] or [0], np.uint32)
        se = np.asarray([b for _, b in sp] or [0], np.uint32)
        buf = np.frombuffer(text, np.uint8)
        for window in (int(rng.integers(1, 20)), 32):
            marks = np.zeros(len(text), np.uint8)
            err = L.emu_pretok_parts(P(buf, C.c_uint8), C.c_uint64(len(text)), P(table, C.c_uint8), P(sb, C.c_uint32),
                                     P(se, C.c_uint32), C.c_uint32(len(sp)), C.c_uint64(window), it & 1, P(marks, C.c_uint8))
            got = np.flatnonzero(marks).astype(np.uint64)
            assert err == 0 and np.array_equal(got, want), (text, window, got.tolist(), want.tolist(), sp)


def test_fuzz_class_runs(emu, pcre2_starts):
    """long runs of one class (digits, symbols, blanks, newlines) across many windows"""
    rng = np.random.default_rng(7)
    pieces = ["7" * 50, "-" * 70, " " * 40, "\n" * 33, "x" * 90, "中" * 30, " \n" * 20, " " * 9, "a1" * 20,
              "!a" * 20, "'s" * 10, " !" * 10, "\r\n\r\n", "\t\t\t", "1 2  3   4", "　\n　"]
    for _ in range(200):
        text = "".join(pieces[i] for i in rng.integers(0, len(pieces), int(rng.integers(1, 8)))).encode()
        want = pcre2_starts(text)
        for window in (5, 32, 64):
            got, err = emu(text, window)
            assert err == 0 and np.array_equal(got, want), (text, window)


def test_edges(emu, pcre2_starts):
    for text in [b"", b"a", b" ", b"\n", b"'", b"'s", b" a", b"  ", b"  a", b"a  ", b"\n ", b" \n", b"1234", b"a'", b"''s",
                 " ".encode(), " a".encode(), "   a".encode(), "a　　b".encode()]:
        got, err = emu(text, 4)
        assert err == 0 and np.array_equal(got, pcre2_starts(text)), text


def test_invalid_utf8_and_pathological_runs_are_reported(emu):
    for bad in [b"abc \xff def", b"abc \xc3", b"x \xe4\xb8 y", b"\xc0\x80 overlong", b"\xed\xa0\x80 surrogate", b"a \x80 b"]:
        _, err = emu(bad * 3, 8)
        assert err & 1, bad
    _, err = emu(b"a " + b"7" * 5000 + b" b", 32, max_crawl=1024)
    assert err & 2
    _, err = emu(b"a " + b"7" * 500 + b" b", 32, max_crawl=1024)
    assert err == 0


# ---------------------------------------------------------------------------------------------------------
# GPU tier: the same matcher as a kernel, through the C ABI, against PCRE2 (mbpe_split) and the host dedup
# ---------------------------------------------------------------------------------------------------------
def _texts(pkg):
    ts = [open(os.path.join(ROOT, "tests", "golden", "data", n), "rb").read() for n in ("taylorswift.txt", "shakespeare.txt")]
    ts.append(pkg.synth_corpus(0x5EED0001, 8 << 20).tobytes())
    rng = np.random.default_rng(5)
    ts.append("".join(ALPHABET[i] for i in rng.integers(0, len(ALPHABET), 200000)).encode())
    return ts


@pytest.mark.gpu
def test_gpu_split_matches_pcre2(pkg):
    pt = pkg.Pretok()
    for text in _texts(pkg) + [b"", b"a", b" ", b"'s", b"\n\n", " a".encode()]:
        s, e = pkg.split(pkg.patterns()["gpt4"], text, 0)
        want = np.concatenate([s, [len(text)]]).astype(np.uint64) if len(s) else np.asarray([len(text)], np.uint64)
        got = pt.split(text)
        assert np.array_equal(got, want), (len(text), len(got), len(want))
    pt.close()


@pytest.mark.gpu
def test_gpu_split_gpt2_pattern(pkg):
    pt = pkg.Pretok(pattern=pkg.patterns()["gpt2"])
    for text in _texts(pkg)[:2] + _texts(pkg)[3:] + [b"", b"a 'll b", b"  x"]:
        s, e = pkg.split(pkg.patterns()["gpt2"], text, 0)
        want = np.concatenate([s, [len(text)]]).astype(np.uint64) if len(s) else np.asarray([len(text)], np.uint64)
        assert np.array_equal(pt.split(text), want)
    pt.close()
    with pytest.raises(pkg.MbpeError) as ei:
        pkg.Pretok(pattern="\\w+")
    assert ei.value.code == -8


@pytest.mark.gpu
def test_gpu_split_refuses_what_it_cannot_take(pkg):
    pt = pkg.Pretok()
    for bad in [b"abc \xff def" * 100, b"x" * 70 + b"\xe4\xb8"]:
        with pytest.raises(pkg.MbpeError) as ei:
            pt.split(bad)
        assert ei.value.code == -8
    os.environ["MBPE_PRETOK_MAX_CRAWL"] = "1024"
    try:
        pt2 = pkg.Pretok()
        with pytest.raises(pkg.MbpeError) as ei:
            pt2.split(b"a " + b"7" * 100000 + b" b")
        assert ei.value.code == -8
        pt2.close()
    finally:
        del os.environ["MBPE_PRETOK_MAX_CRAWL"]
    assert len(pt.split(b"a " + b"7" * 3000 + b" b")) == 1 + 1000 + 2 + 1  # default limit takes it: a, 1000 x 777, ' b'... + end
    pt.close()


@pytest.mark.gpu
def test_gpu_dedup_matches_host(pkg):
    pt = pkg.Pretok()
    for text in _texts(pkg) + [b"", b"aaaa", b"a a a a b b a"]:
        tok, off, w, n_chunks = pkg.split_dedup(pkg.patterns()["gpt4"], text)
        c = pt.corpus(text)
        assert (c.n_chunks, c.n_unique, c.n_tokens) == (n_chunks, len(w), len(tok))
        gt, go, gw = c.download()
        assert np.array_equal(go, off) and np.array_equal(gw, w) and np.array_equal(gt, tok)
        c.free()
    pt.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["lexical", "first"])
def test_gpu_corpus_trains_to_the_same_merges(pkg, mode):
    text = pkg.synth_corpus(0x5EED0001, 4 << 20).tobytes()
    tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text)
    want, wc, _ = pkg.train(tok, off, w, 1024, mode)
    pt = pkg.Pretok()
    c = pt.corpus(text)
    t = c.trainer()
    c.free()
    got, gc, _ = t.run(1024, mode)
    t.close()
    pt.close()
    assert np.array_equal(got, want) and np.array_equal(gc, wc)


def _train_cases(manifest):
    return sorted(k for k, e in manifest["train"].items() if e["rc"] == 0 and e["vocab_size"] <= 4096)


@pytest.mark.gpu
@pytest.mark.parametrize("gpu_split", ["0", "1"])
def test_tokenizer_goldens_with_host_and_device_split(pkg, manifest, tmp_path, monkeypatch, gpu_split):
    """Tokenizer::train + save and Tokenizer::encode give the reference's bytes whichever side pre-tokenises"""
    import hashlib
    from conftest import GOLDEN, golden_data
    monkeypatch.setenv("MBPE_GPU_SPLIT", gpu_split)
    for name in _train_cases(manifest):
        e = manifest["train"][name]
        tk = pkg.Tokenizer(pkg.patterns()[e["encoder"]])
        if e["special"]:
            tk.set_special_tokens_from_file(golden_data(e["special"]))
        tk.train(golden_data(e["input"]), e["vocab_size"], e["mode"])
        assert tk.last_train_stats()["split_on_gpu"] == (gpu_split == "1" and e["encoder"] in ("gpt4", "gpt2")), name
        out = tmp_path / "out.model"
        tk.save(out, write_vocab=False)
        assert hashlib.sha256(out.read_bytes()).hexdigest() == e["model_sha256"], name
    for name, e in sorted(manifest["encode"].items()):
        if e["rc"] != 0:
            continue
        tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
        tk.load(os.path.join(GOLDEN, "models", e["model"] + ".model"))
        ids = tk.encode(golden_data(e["input"]))
        assert len(ids) == e["n_tokens"] and hashlib.sha256(ids.tobytes()).hexdigest() == e["enc_sha256"], name


@pytest.mark.gpu
def test_segmented_corpus_and_encode_text(pkg, monkeypatch):
    """host text brought over in many small segments (cut at matcher cuts) == one piece == host path"""
    text = open(os.path.join(ROOT, "tests", "golden", "data", "shakespeare.txt"), "rb").read()
    tok, off, w, n_chunks = pkg.split_dedup(pkg.patterns()["gpt4"], text)
    merges, _, _ = pkg.train(tok, off, w, 600, "lexical")
    enc = pkg.Encoder(merges)
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    want_ids = enc.encode(text, np.concatenate([s, e[-1:]]).astype(np.uint64))
    for seg in ("100000", "4096", str(1 << 30)):
        monkeypatch.setenv("MBPE_PRETOK_SEG_BYTES", seg)
        monkeypatch.setenv("MBPE_ENCODE_SEG_BYTES", seg)
        pt = pkg.Pretok()
        if seg == "4096":  # more than DD_MAX_SEGS segments: the corpus path declines, the encode path streams them
            with pytest.raises(pkg.MbpeError) as ei:
                pt.corpus(text)
            assert ei.value.code == -8
        else:
            c = pt.corpus(text)
            gt, go, gw = c.download()
            assert c.n_chunks == n_chunks and np.array_equal(go, off) and np.array_equal(gw, w) and np.array_equal(gt, tok)
            c.free()
        assert np.array_equal(pt.encode_text(enc, text), want_ids), seg
        pt.close()
    enc.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seg", ["65536", "1048576", str(64 << 20)])
def test_streaming_file_encode_equals_in_memory(pkg, tmp_path, monkeypatch, seg):
    """mbpe_encode_file (reader thread -> device pipeline -> writer thread, blocks cut at matcher cuts with carried
    tails) writes the same .enc bytes as encoding the whole text in memory"""
    monkeypatch.setenv("MBPE_ENCODE_SEG_BYTES", seg)
    text = pkg.synth_corpus(0x5EED0009, 12 << 20).tobytes()
    tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text[:4 << 20])
    merges, _, _ = pkg.train(tok, off, w, 1500, "lexical")
    model = tmp_path / "m.model"
    pkg.write_model(str(model), pkg.patterns()["gpt4"], None, merges)
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
    tk.load(model)
    want = tk.encode(text)
    src, dst = tmp_path / "in.txt", tmp_path / "out.enc"
    src.write_bytes(text)
    n = tk.encode_file(src, dst)
    got = np.fromfile(dst, np.uint32)
    assert n == len(want) == len(got) and np.array_equal(got, want)
    if seg == "65536":  # and back: .enc file -> text file in blocks
        back = tmp_path / "back.txt"
        assert tk.decode_file(dst, back) == (len(want), len(text)) and back.read_bytes() == text
        (tmp_path / "odd.enc").write_bytes(want[:5].tobytes() + b"\x01\x02")  # trailing partial word is dropped
        assert tk.decode_file(tmp_path / "odd.enc", back)[0] == 5 and back.read_bytes() == tk.decode(want[:5])
    # empty and tiny files
    (tmp_path / "e.txt").write_bytes(b"")
    assert tk.encode_file(tmp_path / "e.txt", tmp_path / "e.enc") == 0 and (tmp_path / "e.enc").stat().st_size == 0
    (tmp_path / "t.txt").write_bytes(b"hello world")
    assert tk.encode_file(tmp_path / "t.txt", tmp_path / "t.enc") == len(tk.encode(b"hello world"))


@pytest.mark.gpu
def test_cli_streams_big_inputs_and_matches_whole_file_path(pkg, tmp_path):
    cli = os.path.join(ROOT, "minbpe-cc_b200", "bin", "minbpe-cc")
    text = pkg.synth_corpus(0x5EED000A, 9 << 20).tobytes()  # >= 8 MiB: the CLI takes the streaming path
    src = tmp_path / "in.txt"
    src.write_bytes(text)
    model = tmp_path / "m.model"
    small = tmp_path / "small.txt"
    small.write_bytes(text[:1 << 20])
    r = subprocess.run([cli, "--train", "-i", str(small), "-m", str(model), "--vocab-size", "700", "--encoder", "gpt4",
                        "-c", "lexical"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    outs = {}
    for name, env in (("stream", {}), ("whole", {"MBPE_GPU_SPLIT": "0"})):
        out = tmp_path / (name + ".enc")
        r = subprocess.run([cli, "--encode", "-i", str(src), "-m", str(model), "-o", str(out)], capture_output=True,
                           text=True, env={**os.environ, **env})
        assert r.returncode == 0 and "Success" in r.stdout, (r.stdout, r.stderr)
        outs[name] = out.read_bytes()
    assert outs["stream"] == outs["whole"] and len(outs["stream"]) > 0
    back = tmp_path / "back.txt"
    r = subprocess.run([cli, "--decode", "-i", str(tmp_path / "stream.enc"), "-m", str(model), "-o", str(back)],
                       capture_output=True, text=True)
    assert r.returncode == 0 and back.read_bytes() == text


@pytest.mark.gpu
def test_encode_with_special_tokens_on_the_device(pkg, monkeypatch):
    """Tokenizer::encode of a text full of special tokens: device front end (occurrences found on the host, ordinary parts
    split as independent subjects by the kernel, special chunks resolved through the seeded chunk cache) == host path"""
    base = pkg.synth_corpus(0x5EED000B, 3 << 20).tobytes()
    tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], base[:1 << 20])
    merges, _, _ = pkg.train(tok, off, w, 900, "lexical")
    specials = b"<|endoftext|> 100257\n<|fim|> 100258\nab 100259\n"
    rng = np.random.default_rng(4)
    cuts = np.sort(rng.integers(0, len(base), 4000))
    arr = np.frombuffer(base, np.uint8)
    cuts = np.asarray([c for c in cuts if (arr[c] & 0xC0) != 0x80])  # never inside a UTF-8 sequence: the text stays well formed
    toks = [b"<|endoftext|>", b"<|fim|>", b"ab", b"<|endoftext|><|fim|>", b""]
    text = b"".join(base[a:b] + toks[int(k)] for a, b, k in zip(np.r_[0, cuts], np.r_[cuts, len(base)], rng.integers(0, len(toks), len(cuts) + 1)))
    text = b"<|fim|>" + text + b"<|endoftext|>"
    results = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("MBPE_GPU_SPLIT", mode)
        monkeypatch.setenv("MBPE_ENCODE_SEG_BYTES", str(1 << 20))  # several segments: boundaries at occurrences and at cuts
        tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            mp = os.path.join(td, "m.model")
            pkg.write_model(mp, pkg.patterns()["gpt4"], specials, merges)
            tk.load(mp)
        results[mode] = tk.encode(text)
        assert tk.decode(results[mode]) == text
        if mode == "1":  # the device front end really ran: straight through the C ABI as well
            parts = pkg.special_split(specials, text)
            sb = np.asarray([a for a, _, i in parts if i >= 0], np.uint64)
            se = np.asarray([b for _, b, i in parts if i >= 0], np.uint64)
            enc = pkg.Encoder(merges)
            toks_ = [(b"<|endoftext|>", 100257), (b"<|fim|>", 100258), (b"ab", 100259)]
            enc.seed_special_chunks(toks_)
            pt = pkg.Pretok()
            direct = pt.encode_text_special(enc, text, sb, se)
            assert np.array_equal(direct, results[mode])
            pt.close()
            enc.close()
    assert np.array_equal(results["0"], results["1"])
    assert (results["1"] == 100257).sum() > 1000 and (results["1"] == 100259).sum() > 1000
