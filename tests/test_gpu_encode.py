"""GPU: parity of the encode merge scan and the decode gather (through the C ABI) with the reference."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, golden_data

pytestmark = pytest.mark.gpu


def _cases():
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    return sorted(k for k, e in man["encode"].items() if e["rc"] == 0)


@pytest.mark.parametrize("name", _cases())
def test_tokenizer_encode_matches_reference_stream(pkg, manifest, name):
    e = manifest["encode"][name]
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])  # replaced by the pattern stored in the model
    tk.load(os.path.join(GOLDEN, "models", e["model"] + ".model"))
    text = golden_data(e["input"])
    ids = tk.encode(text)
    assert len(ids) == e["n_tokens"]
    assert hashlib.sha256(ids.tobytes()).hexdigest() == e["enc_sha256"]
    assert tk.decode(ids) == text  # endtoend-test.sh round trip


def test_expected_ids_of_special_token_sample(pkg):
    tk = pkg.Tokenizer("")
    tk.load(os.path.join(GOLDEN, "models", "ts512_gpt4_first_special.model"))
    ids = tk.encode(golden_data("specialtokensample.txt"))
    assert ids.tolist() == [84, 104, 355, 32, 355, 306, 288, 101, 261, 101, 120, 116, 434, 287, 349, 262, 116, 97, 259,
                            115, 32, 100258, 261, 119, 111, 306, 112, 310, 478, 108, 32, 100257, 348, 107, 290, 115, 46]


def test_chunk_level_encode_vs_oracle_incl_offsets(pkg, oracle):
    _, _, merges = oracle.read_model(os.path.join(GOLDEN, "models", "shk4096_gpt4_lexical_special.model"))
    text = golden_data("shakespeare.txt")
    s, e = oracle.split(text, oracle.GPT4_SPLIT_PATTERN)
    off = np.concatenate([s, e[-1:]])
    enc = pkg.Encoder(merges)
    ids, out_off = enc.encode(text, off, want_off=True)
    oids, ooff = oracle.encode_chunks(merges, text, s, e)
    assert np.array_equal(ids, oids) and np.array_equal(out_off, ooff)


def test_ragged_empty_and_long_chunks(pkg, oracle):
    _, _, merges = oracle.read_model(os.path.join(GOLDEN, "models", "ts512_gpt4_lexical.model"))
    enc = pkg.Encoder(merges)
    assert len(enc.encode(b"", [0])) == 0
    text = golden_data("taylorswift.txt")[:50000]
    rng = np.random.default_rng(9)
    cuts = np.unique(np.concatenate([[0, len(text)], rng.integers(0, len(text), 3000),
                                     np.arange(1000, 1200)]))  # 1-byte chunks, mixed with chunks > 64 bytes
    cuts = np.sort(np.concatenate([cuts, cuts[5:15]]))            # and a few empty chunks
    ids, out_off = enc.encode(text, cuts, want_off=True)
    oids, ooff = oracle.encode_chunks(merges, text, cuts[:-1], cuts[1:])
    assert np.array_equal(ids, oids) and np.array_equal(out_off, ooff)
    # one chunk = whole text (encoder "basic")
    ids = enc.encode(text, [0, len(text)])
    oids, _ = oracle.encode_chunks(merges, text, [0], [len(text)])
    assert np.array_equal(ids, oids)


def test_duplicate_merge_lines_overwrite(pkg, oracle):
    """SURVEY F4/Tokenizer.h:835: a pair listed twice maps to the LATER id."""
    _, _, merges = oracle.read_model(os.path.join(GOLDEN, "models", "str_exhaust_basic_lexical.model"))
    assert len(merges) == 44
    enc = pkg.Encoder(merges)
    text = b"abcdebce abab"
    ids = enc.encode(text, [0, len(text)])
    oids, _ = oracle.encode_chunks(merges, text, [0], [len(text)])
    assert np.array_equal(ids, oids) and 299 in ids.tolist()
    assert enc.decode(ids) == text


def test_decode_invalid_ids_and_special_override(pkg, oracle):
    merges = np.asarray([[97, 98], [256, 99]], np.uint32)
    enc = pkg.Encoder(merges)
    ids = [97, 256, 999999, 257, 98, 258]
    assert enc.decode(ids) == oracle.decode(merges, ids, {}) == b"aababcb"
    sp = {256: b"<S>", 100257: b"<|endoftext|>", 65: b"<A>"}
    enc.set_specials(sp)
    ids = [65, 256, 100257, 257, 5000]
    assert enc.decode(ids) == oracle.decode(merges, ids, sp) == b"<A><S><|endoftext|>abc"
    assert enc.decode([]) == b""


def test_decode_device_matches_host_api_and_reports_overflow(pkg):
    """resident ids -> resident bytes (one pass, staged coalesced stores) == mbpe_decode; undersized buffers are
    reported through the size, bytes past the capacity are not written"""
    import torch
    text = pkg.synth_corpus(0x5EED0007, 6 << 20).tobytes()
    tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text)
    merges, _, _ = pkg.train(tok, off, w, 2000, "lexical")
    enc = pkg.Encoder(merges)
    enc.set_specials({100257: b"<|endoftext|>", 7: b"SEVEN"})
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    ids = enc.encode(text, np.concatenate([s, e[-1:]]).astype(np.uint64))
    rng = np.random.default_rng(1)
    ids = ids.copy()
    ids[rng.integers(0, len(ids), 2000)] = 100257      # special ids override
    ids[rng.integers(0, len(ids), 2000)] = 7            # ... also ids inside the vocabulary
    ids[rng.integers(0, len(ids), 2000)] = 4000000      # unknown ids contribute nothing
    want = enc.decode(ids)
    dev = torch.device("cuda", 0)
    for n in (len(ids), 2048, 2049, 5, 1, 0):
        d_ids = torch.from_numpy(ids[:n].view(np.int32).copy()).to(dev) if n else torch.zeros(4, dtype=torch.int32, device=dev)
        w_n = enc.decode(ids[:n])
        d_out = torch.full((len(w_n) + 64,), 0xAA, dtype=torch.uint8, device=dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)
        enc.decode_device(d_ids.data_ptr(), n, d_out.data_ptr(), len(w_n), d_n.data_ptr())
        torch.cuda.synchronize()
        assert int(d_n.item()) == len(w_n)
        assert d_out[:len(w_n)].cpu().numpy().tobytes() == w_n and bool((d_out[len(w_n):] == 0xAA).all())
    assert want == enc.decode(ids)
    # size only, and a buffer that is too small
    d_ids = torch.from_numpy(ids.view(np.int32).copy()).to(dev)
    d_n = torch.zeros(1, dtype=torch.int64, device=dev)
    enc.decode_device(d_ids.data_ptr(), len(ids), 0, 0, d_n.data_ptr())
    assert int(d_n.item()) == len(want)
    small = torch.full((1000 + 64,), 0xAA, dtype=torch.uint8, device=dev)
    enc.decode_device(d_ids.data_ptr(), len(ids), small.data_ptr(), 1000, d_n.data_ptr())
    torch.cuda.synchronize()
    assert int(d_n.item()) == len(want) and bool((small[1000:] == 0xAA).all())
    enc.close()


def test_large_synthetic_roundtrip_and_oracle_slice(pkg, oracle):
    """Size-independent property at a bench-like size: decode(encode(x)) == x; plus an oracle diff on a slice."""
    text = pkg.synth_corpus(0x5EED0002, 64 << 20).tobytes()
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
    tk.train(text[: 8 << 20], 256 + 2000, "lexical")
    ids = tk.encode(text)
    assert tk.decode(ids) == text
    merges = tk.merges()
    cut = text.rfind(b"\nq", 0, 4 << 20) + 1 or (4 << 20)
    s, e = oracle.split(text[:cut], oracle.GPT4_SPLIT_PATTERN)
    oids, _ = oracle.encode_chunks(merges, text[:cut], s, e)
    assert np.array_equal(ids[:len(oids)], oids)


def test_cli_end_to_end(pkg, manifest, tmp_path):
    """endtoend-test.sh, both scenarios, through the flag-compatible CLI; models must equal the goldens."""
    cli, data = pkg.CLI_PATH, os.path.join(GOLDEN, "data")
    run = lambda *a: subprocess.run([cli, *a], check=True, capture_output=True)
    m1 = str(tmp_path / "basic-model")
    run("--train", "--input", f"{data}/shakespeare.txt", "--model-path", m1, "--vocab-size", "512", "--encoder", "basic",
        "--conflict-resolution", "lexical")
    assert hashlib.sha256(open(m1, "rb").read()).hexdigest() == manifest["train"]["shk512_basic_lexical"]["model_sha256"]
    run("--encode", "--input", f"{data}/sample.txt", "--model-path", m1, "--output", str(tmp_path / "s.enc"))
    assert (hashlib.sha256((tmp_path / "s.enc").read_bytes()).hexdigest()
            == manifest["encode"]["sample__shk512_basic_lexical"]["enc_sha256"])
    run("--decode", "--input", str(tmp_path / "s.enc"), "--model-path", m1, "--output", str(tmp_path / "s.txt"))
    assert (tmp_path / "s.txt").read_bytes() == golden_data("sample.txt")
    m2 = str(tmp_path / "gpt4-model")
    run("-t", "-i", f"{data}/taylorswift.txt", "-s", f"{data}/special1.txt", "-m", m2, "--vocab-size=512",
        "--encoder", "gpt4", "-c", "first")
    assert hashlib.sha256(open(m2, "rb").read()).hexdigest() == manifest["train"]["ts512_gpt4_first_special"]["model_sha256"]
    run("-e", "-i", f"{data}/specialtokensample.txt", "-m", m2, "-o", str(tmp_path / "t.enc"))
    run("-d", "-i", str(tmp_path / "t.enc"), "-m", m2, "-o", str(tmp_path / "t.txt"))
    assert (tmp_path / "t.txt").read_bytes() == golden_data("specialtokensample.txt")


def _long_chunk_text(n_bytes, specials, seed=5):
    """text whose chunks average well over 10 bytes (CJK runs, long identifiers, URLs) with special tokens sprinkled in:
    the tile kernel cannot stage such tiles in shared memory and takes its slow path for every chunk"""
    rng = np.random.default_rng(seed)
    cjk = "".join(chr(int(c)) for c in rng.integers(0x4E00, 0x9FA5, 4000))
    parts, size = [], 0
    while size < n_bytes:
        k = int(rng.integers(0, 5))
        if k == 0:
            a = int(rng.integers(0, 3900))
            p = cjk[a:a + int(rng.integers(8, 30))]
        elif k == 1:
            p = "https://example.org/" + "".join(rng.choice(list("abcdefghij_"), int(rng.integers(20, 60))))
        elif k == 2:
            p = "".join(rng.choice(list("ABCDEFabcdef"), int(rng.integers(16, 40))))
        elif k == 3:
            p = specials[int(rng.integers(0, len(specials)))]
        else:
            p = " " if rng.random() < 0.7 else "\n"
        parts.append(p)
        size += len(p.encode())
    return "".join(parts).encode()


def test_special_tokens_in_long_chunk_text_and_of_any_length(pkg, monkeypatch):
    """specials must come out as their ids wherever they sit -- in tiles too wide for shared memory,
    with the caches off, and when the token itself is longer than a cache key (31 bytes) or than ENC_SHORT_MAX (64)."""
    specials = ["<|endoftext|>", "<|fim_prefix|>", "<|a_special_token_that_is_longer_than_31_bytes|>",
                "<|" + "x" * 80 + "|>"]
    contents = "".join(f"{s} {100257 + i}\n" for i, s in enumerate(specials))
    text = _long_chunk_text(3 << 20, specials)
    train_text = pkg.synth_corpus(0x5EED0011, 4 << 20).tobytes() + text[: 1 << 20]
    results = {}
    for name, env in (("device", {}), ("device_nocache", {"MBPE_ENCODE_CACHE": "0"}), ("host_split", {"MBPE_GPU_SPLIT": "0"})):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
        tk.train(train_text, 256 + 600, "lexical")
        tk.set_special_tokens_from_file(contents)
        ids = tk.encode(text)
        assert tk.decode(ids) == text
        results[name] = ids
        for k in env:
            monkeypatch.delenv(k)
        tk.close()
    assert np.array_equal(results["device"], results["host_split"])
    assert np.array_equal(results["device_nocache"], results["host_split"])
    for i, s in enumerate(specials):  # every occurrence became its one id (the host path resolves them on the host)
        assert int((results["device"] == 100257 + i).sum()) == text.count(s.encode()) > 0


def test_load_reselects_the_device_matcher(pkg, tmp_path, monkeypatch):
    """load() replaces the pattern; large texts must then split under the NEW pattern on
    the device too. GPT-2 and GPT-4 split "Hello's 12345   world" differently, so a stale matcher shows in the ids."""
    text = pkg.synth_corpus(0x5EED0012, 2 << 20).tobytes() + b" it's 1234567 HELLO'S   x\n" * 4000
    t2 = pkg.Tokenizer(pkg.patterns()["gpt2"])
    t2.train(text, 256 + 300, "lexical")
    t2.save(tmp_path / "gpt2.model")
    tk = pkg.Tokenizer(pkg.patterns()["gpt4"])
    tk.train(text, 256 + 50, "lexical")  # the device matcher now matches the GPT-4 pattern
    tk.load(tmp_path / "gpt2.model")
    ids = tk.encode(text)
    assert np.array_equal(ids, t2.encode(text))
    monkeypatch.setenv("MBPE_GPU_SPLIT", "0")
    th = pkg.Tokenizer(pkg.patterns()["gpt4"])
    th.load(tmp_path / "gpt2.model")
    assert np.array_equal(ids, th.encode(text))
    monkeypatch.delenv("MBPE_GPU_SPLIT")
    # and training after the load follows the loaded pattern as well
    tk.train(text, 256 + 300, "lexical")
    assert np.array_equal(tk.merges(), t2.merges())


def test_unaligned_device_buffers_take_the_cooperative_path(pkg, oracle):
    """mbpe_encode_device with text / boundary pointers that are not 16-byte aligned (no bulk copies possible)"""
    import torch
    _, _, merges = oracle.read_model(os.path.join(GOLDEN, "models", "shk4096_gpt4_lexical_special.model"))
    text = golden_data("shakespeare.txt")
    s, e = oracle.split(text, oracle.GPT4_SPLIT_PATTERN)
    off32 = np.concatenate([s, e[-1:]]).astype(np.uint32)
    oids, _ = oracle.encode_chunks(merges, text, s, e)
    dev = torch.device("cuda", 0)
    enc = pkg.Encoder(merges)
    for shift_b, shift_o in ((0, 0), (1, 0), (0, 1), (3, 3)):
        d_bytes = torch.zeros(len(text) + 64, dtype=torch.uint8, device=dev)
        d_bytes[shift_b:shift_b + len(text)] = torch.from_numpy(np.frombuffer(text, np.uint8).copy()).to(dev)
        d_off = torch.zeros(len(off32) + 8, dtype=torch.int32, device=dev)
        d_off[shift_o:shift_o + len(off32)] = torch.from_numpy(off32.view(np.int32).copy()).to(dev)
        d_out = torch.zeros(len(text), dtype=torch.int32, device=dev)
        d_n = torch.zeros(1, dtype=torch.int64, device=dev)
        enc.encode_device(d_bytes.data_ptr() + shift_b, len(text), d_off.data_ptr() + 4 * shift_o, len(s), d_out.data_ptr(),
                          len(text), d_n.data_ptr())
        torch.cuda.synchronize()
        n = int(d_n.item())
        assert n == len(oids) and np.array_equal(d_out[:n].cpu().numpy().view(np.uint32), oids), (shift_b, shift_o)
    enc.close()


def test_encode_is_the_same_for_every_kernel_shape_and_without_caches(pkg, oracle, monkeypatch):
    """every k_encode_tiles configuration and both k_encode_hot shapes (8, 9), bulk and cooperative staging, caches on / off / tiny: identical ids, cold and warm"""
    text = pkg.synth_corpus(0x5EED0013, 6 << 20).tobytes() + golden_data("sample.txt") * 20
    tok, off, w, _ = pkg.split_dedup(pkg.patterns()["gpt4"], text[: 4 << 20])
    merges, _, _ = pkg.train(tok, off, w, 256 + 3000, "lexical")
    s, e = pkg.split(pkg.patterns()["gpt4"], text)
    chunk_off = np.concatenate([s, e[-1:]]).astype(np.uint64)
    cut = text.rfind(b"\nq", 0, 1 << 20) + 1 or (1 << 20)
    so, eo = oracle.split(text[:cut], oracle.GPT4_SPLIT_PATTERN)
    oids, _ = oracle.encode_chunks(merges, text[:cut], so, eo)
    ref = None
    envs = [{"MBPE_ENC_CFG": str(c)} for c in range(10)] + [{"MBPE_ENC_NO_BULK": "1"}, {"MBPE_ENCODE_CACHE": "0"},
                                                            {"MBPE_ENCODE_CACHE": "12"}, {"MBPE_ENCODE_SUBBATCH": "8192"}]
    for env in envs:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        enc = pkg.Encoder(merges)
        for rep in range(2):  # cold caches, then warm
            ids = enc.encode(text, chunk_off)
            if ref is None:
                ref = ids
                assert np.array_equal(ids[:len(oids)], oids)
            assert np.array_equal(ids, ref), (env, rep)
        enc.close()
        for k in env:
            monkeypatch.delenv(k)


def test_decode_is_the_same_for_every_kernel_shape(pkg, oracle, monkeypatch):
    """every decode configuration (thread-per-8-ids, lane-per-token, k_decode_lean, no packed table): identical bytes for a
    stream with tokens of 1..40 bytes, special tokens inside and outside the vocabulary (one of them 3000 bytes long, one
    with the id 0xFFFFFFFF), unknown ids, and a last tile that is not full"""
    # a vocabulary with long tokens: chain merges build "abcdefgh..." of up to 40 bytes
    merges = [[97 + i, 98 + i] for i in range(0, 20, 2)]          # 10 two-byte tokens 256..265
    merges += [[256 + i, 257 + i] for i in range(0, 10, 2)]      # 5 four-byte tokens 266..270
    merges += [[266, 267], [268, 269], [271, 272], [273, 270], [274, 274]]  # 8, 8, 16, 20, 40 bytes: 271..275
    merges = np.asarray(merges, np.uint32)
    sp = {7: b"SEVEN", 260: b"<two-sixty>", 100257: b"<|endoftext|>", 0xFFFFFFFF: b"<max>", 5000: b"L" * 3000}
    rng = np.random.default_rng(5)
    ids = rng.integers(0, 256 + len(merges), 300001).astype(np.uint32)
    ids[rng.integers(0, len(ids), 3000)] = 100257
    ids[rng.integers(0, len(ids), 3000)] = 7
    ids[rng.integers(0, len(ids), 3000)] = 260
    ids[rng.integers(0, len(ids), 300)] = 0xFFFFFFFF
    ids[rng.integers(0, len(ids), 30)] = 5000
    ids[rng.integers(0, len(ids), 3000)] = 4000000   # unknown
    ids[rng.integers(0, len(ids), 3000)] = 300       # unknown, just past the vocabulary
    want = oracle.decode(merges, ids[:5000].tolist(), sp)
    ref = None
    envs = [{"MBPE_DEC_CFG": str(c)} for c in range(16)] + [{"MBPE_DEC_NOPACK": "1"}]
    for env in envs:
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        enc = pkg.Encoder(merges)
        assert enc.decode(ids[:5000]) == oracle.decode(merges, ids[:5000].tolist(), {}), env  # before any special token is set
        enc.set_specials(sp)
        assert enc.decode(ids[:5000]) == want, env
        got = enc.decode(ids)
        if ref is None:
            ref = got
        assert got == ref, env
        enc.set_specials({})  # and the table forgets them again
        assert enc.decode(ids[:5000]) == oracle.decode(merges, ids[:5000].tolist(), {}), env
        enc.close()
        for k in env:
            monkeypatch.delenv(k)
