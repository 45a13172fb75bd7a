"""ctypes front end of the CPU oracle (oracle/bpe_oracle.c). TEST INFRASTRUCTURE ONLY.

Imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "_ref", "libbpe_oracle.so")
REF_DRIVER = os.path.join(HERE, "_ref", "ref_driver")

# Tokenizer.h:59-60 (public constants of the reference API; the product mirrors them in C++)
GPT2_SPLIT_PATTERN = r"'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"
GPT4_SPLIT_PATTERN = (r"'(?i:[sdmt]|ll|ve|re)|[^\r\n\p{L}\p{N}]?+\p{L}+|\p{N}{1,3}| ?[^\s\p{L}\p{N}]++[\r\n]*"
                      r"|\s*[\r\n]|\s+(?!\S)|\s+")
PATTERNS = {"basic": "", "gpt2": GPT2_SPLIT_PATTERN, "gpt4": GPT4_SPLIT_PATTERN}
MODES = {"first": 0, "lexical": 1}

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(HERE, "bpe_oracle.c")):
        subprocess.check_call(["make", "-C", HERE, "port"], stdout=subprocess.DEVNULL)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        # bpe_oracle.c calls PCRE2; the image has the runtime .so only
        C.CDLL("libpcre2-8.so.0", mode=C.RTLD_GLOBAL)
        _lib = C.CDLL(LIB)
        _lib.oracle_split.restype = C.c_int64
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def split(text: bytes, pattern: str, jit=True):
    """Tokenizer.h:500-544. Returns (starts, ends) uint64 arrays."""
    L = lib()
    buf = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(0, np.uint8)
    n = L.oracle_split(_p(buf, C.c_uint8), C.c_uint64(len(text)), pattern.encode(), int(jit), None, None, C.c_uint64(0))
    if n < 0:
        raise RuntimeError(f"oracle_split failed: {n}")
    starts = np.zeros(max(n, 1), np.uint64)
    ends = np.zeros(max(n, 1), np.uint64)
    n2 = L.oracle_split(_p(buf, C.c_uint8), C.c_uint64(len(text)), pattern.encode(), int(jit),
                        _p(starts, C.c_uint64), _p(ends, C.c_uint64), C.c_uint64(n))
    assert n2 == n
    return starts[:n], ends[:n]


def chunks_of(text: bytes, encoder: str):
    s, e = split(text, PATTERNS[encoder])
    return [text[int(a):int(b)] for a, b in zip(s, e)]


def text_to_tokens(chunk: bytes):
    """Tokenizer.h:85-100 incl. the leading-NUL quirk."""
    if chunk[:1] == b"\0":
        try:
            # std::stoi: optional whitespace, sign, then digits; stops at the first non-digit
            import re
            m = re.match(rb"[ \t\n\v\f\r]*([+-]?[0-9]+)", chunk[1:])
            if m:
                v = int(m.group(1))
                if -2**31 <= v < 2**31:
                    return [v & 0xFFFFFFFF]
        except ValueError:
            pass
    return list(chunk)


def flatten(chunk_list, dedup):
    """chunk list -> (tokens u32, off u64, weight u32). dedup keeps first-appearance order."""
    if dedup:
        seen = {}
        for c in chunk_list:
            seen[c] = seen.get(c, 0) + 1
        uniq, weight = list(seen.keys()), list(seen.values())
    else:
        uniq, weight = chunk_list, [1] * len(chunk_list)
    toks, off = [], [0]
    for c in uniq:
        toks.extend(text_to_tokens(c))
        off.append(len(toks))
    return (np.asarray(toks, np.uint32), np.asarray(off, np.uint64), np.asarray(weight, np.uint32))


def train(tokens, off, weight, vocab_size, mode, impl="indexed"):
    """Returns (merges [n,2] u32, counts [n] i32)."""
    L = lib()
    n_max = max(vocab_size - 256, 0)
    merges = np.zeros((max(n_max, 1), 2), np.uint32)
    counts = np.zeros(max(n_max, 1), np.int32)
    nm = C.c_uint32(0)
    fn = L.oracle_train_indexed if impl == "indexed" else L.oracle_train_rescan
    tokens = np.ascontiguousarray(tokens, np.uint32)
    off = np.ascontiguousarray(off, np.uint64)
    weight = np.ascontiguousarray(weight, np.uint32)
    rc = fn(_p(tokens, C.c_uint32), C.c_uint64(len(tokens)), _p(off, C.c_uint64), C.c_uint64(len(off) - 1),
            _p(weight, C.c_uint32), C.c_uint32(vocab_size), MODES[mode] if isinstance(mode, str) else mode,
            _p(merges, C.c_uint32), _p(counts, C.c_int32), C.byref(nm))
    assert rc == 0
    return merges[:nm.value].copy(), counts[:nm.value].copy()


def train_text(text: bytes, vocab_size, encoder, mode, dedup=True, impl="indexed"):
    t, o, w = flatten(chunks_of(text, encoder), dedup)
    return train(t, o, w, vocab_size, mode, impl)


def encode_chunks(merges, text: bytes, starts, ends):
    L = lib()
    merges = np.ascontiguousarray(merges, np.uint32).reshape(-1, 2)
    buf = np.frombuffer(text, dtype=np.uint8) if len(text) else np.zeros(1, np.uint8)
    starts = np.ascontiguousarray(starts, np.uint64)
    ends = np.ascontiguousarray(ends, np.uint64)
    total = int((ends - starts).sum())
    out = np.zeros(max(total, 1), np.uint32)
    out_off = np.zeros(len(starts) + 1, np.uint64)
    n_out = C.c_uint64(0)
    rc = L.oracle_encode(_p(merges, C.c_uint32), C.c_uint32(len(merges)), _p(buf, C.c_uint8), _p(starts, C.c_uint64),
                         _p(ends, C.c_uint64), C.c_uint64(len(starts)), _p(out, C.c_uint32), _p(out_off, C.c_uint64),
                         C.byref(n_out))
    assert rc == 0
    return out[:n_out.value].copy(), out_off


def decode(merges, ids, specials=None):
    """specials: dict id -> bytes."""
    L = lib()
    merges = np.ascontiguousarray(merges, np.uint32).reshape(-1, 2)
    ids = np.ascontiguousarray(ids, np.uint32)
    specials = specials or {}
    sid = np.asarray(list(specials.keys()) or [0], np.uint32)
    sb = b"".join(specials.values())
    soff = np.zeros(len(specials) + 1, np.uint64)
    np.cumsum([len(v) for v in specials.values()], out=soff[1:]) if specials else None
    sbuf = np.frombuffer(sb, dtype=np.uint8) if sb else np.zeros(1, np.uint8)
    n_out = C.c_uint64(0)
    args = (_p(merges, C.c_uint32), C.c_uint32(len(merges)), _p(sid, C.c_uint32), _p(sbuf, C.c_uint8),
            _p(soff, C.c_uint64), C.c_uint32(len(specials)), _p(ids, C.c_uint32), C.c_uint64(len(ids)))
    rc = L.oracle_decode(*args, None, C.byref(n_out))
    if rc != 0:
        raise RuntimeError("oracle_decode: malformed merges")
    out = np.zeros(max(n_out.value, 1), np.uint8)
    L.oracle_decode(*args, _p(out, C.c_uint8), C.byref(n_out))
    return out[:n_out.value].tobytes()


def paircount_top(ops, mode):
    """ops: list of (a, b, delta). Mirrors PairCount::create_or_modify_pair + get_top_pair_count."""
    L = lib()
    arr = np.asarray(ops, np.int32).reshape(-1, 3)
    a, b, c, n = C.c_uint32(), C.c_uint32(), C.c_int32(), C.c_uint64()
    rc = L.oracle_paircount_top(_p(arr, C.c_int32), C.c_uint64(len(arr)), MODES[mode], C.byref(a), C.byref(b),
                                C.byref(c), C.byref(n))
    return (None if rc else (a.value, b.value, c.value)), n.value


# ---- file formats, restated for the checker (Tokenizer.h:754-872 load, :875-926 save) -------------------------

def read_model(path):
    """Returns (pattern, specials [(token, id)] in file order, merges [n,2])."""
    with open(path, "rb") as f:
        lines = f.read().split(b"\n")
    assert lines[0] == b"minbpe v1"
    pattern = lines[1].decode()
    ns = int(lines[2])
    specials = []
    for i in range(ns):
        tok, sid = lines[3 + i].rsplit(b" ", 1)
        specials.append((tok, int(sid)))
    merges = [tuple(int(x) for x in ln.split()) for ln in lines[3 + ns:] if ln.strip()]
    return pattern, specials, np.asarray(merges, np.uint32).reshape(-1, 2)


def model_bytes(pattern: str, specials_in_order, merges):
    """.model writer (Tokenizer.h:881-891). specials_in_order must already be in the reference's
    unordered_map iteration order; the checker takes it from the golden file."""
    out = [b"minbpe v1", pattern.encode(), str(len(specials_in_order)).encode()]
    out += [t + b" " + str(i).encode() for t, i in specials_in_order]
    out += [f"{a} {b}".encode() for a, b in np.asarray(merges).reshape(-1, 2)]
    return b"\n".join(out) + b"\n"


def vocab_bytes(merges):
    """.vocab writer (Tokenizer.h:905-917)."""
    vocab = [bytes([i]) for i in range(256)]
    for a, b in np.asarray(merges).reshape(-1, 2):
        vocab.append(vocab[a] + vocab[b])
    out = []
    for i, v in enumerate(vocab):
        s = b"".join(bytes([c]) if 32 <= c <= 126 else b"\xef\xbf\xbd" for c in v)
        out.append(str(i).ljust(6).encode() + b': "' + s + b'"\n')
    return b"".join(out)


def split_on_special(text: bytes, specials_in_map_order):
    """Tokenizer.h:605-650. specials_in_map_order: [(token bytes, id)] in unordered_map iteration order
    (only matters when two specials match at the same position). Returns parts; special parts are
    b'\\0' + decimal id."""
    if not specials_in_map_order:
        return [text]
    out, pos, last = [], 0, 0
    while pos < len(text):
        found, ftok, fid = -1, None, 0
        for tok, sid in specials_in_map_order:
            p = text.find(tok, pos)
            if p != -1 and (found == -1 or p < found):
                found, ftok, fid = p, tok, sid
        if found == -1:
            break
        if found > last:
            out.append(text[last:found])
        out.append(b"\0" + str(fid).encode())
        pos = found + len(ftok)
        last = pos
    if last < len(text):
        out.append(text[last:])
    if not out:
        out.append(text)
    return out


def encode_text(text: bytes, pattern: str, specials_in_map_order, merges):
    """Tokenizer::encode (Tokenizer.h:653-722) on top of oracle_split / oracle_encode."""
    ids = []
    for part in split_on_special(text, specials_in_map_order):
        if part[:1] == b"\0":
            ids.append(np.asarray(text_to_tokens(part), np.uint32))
            continue
        if pattern:
            s, e = split(part, pattern, jit=False)
        else:
            s, e = np.asarray([0], np.uint64), np.asarray([len(part)], np.uint64)
        # a regex chunk that starts with NUL + digits becomes one ready-made id
        quirk = [i for i in range(len(s)) if part[int(s[i]):int(s[i]) + 1] == b"\0"
                 and len(text_to_tokens(part[int(s[i]):int(e[i])])) == 1 and int(e[i]) - int(s[i]) > 1]
        if quirk:
            for i in range(len(s)):
                ch = part[int(s[i]):int(e[i])]
                if i in quirk:
                    ids.append(np.asarray(text_to_tokens(ch), np.uint32))
                else:
                    ids.append(encode_chunks(merges, ch, [0], [len(ch)])[0])
        else:
            ids.append(encode_chunks(merges, part, s, e)[0])
    return np.concatenate(ids) if ids else np.zeros(0, np.uint32)
