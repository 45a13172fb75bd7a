"""ctypes binding of libminbpe_b200.so (include/minbpe_b200.h) for the tests and bench.py.

The directory name has a hyphen (it mirrors the reference's name), so import it with
`importlib` -- see `load_package()` in tests/conftest.py / __graft_entry__.py -- as module `minbpe_cc_b200`.

This is plumbing only: every compute entry point is a C-ABI call into hand-written sm_100a kernels. There is no
Python or CPU implementation behind it; a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libminbpe_b200.so")
CLI_PATH = os.path.join(HERE, "bin", "minbpe-cc")

MODE = {"first": 0, "lexical": 1}
ENGINE = {"stepwise": 0, "persistent": 1}

EXPORTS = [
    "mbpe_version", "mbpe_last_error", "mbpe_device_count",
    "mbpe_paircount_create", "mbpe_paircount_destroy", "mbpe_paircount_add", "mbpe_paircount_top",
    "mbpe_paircount_get", "mbpe_paircount_size",
    "mbpe_trainer_create", "mbpe_trainer_run", "mbpe_trainer_destroy", "mbpe_train",
    "mbpe_comm_unique_id", "mbpe_comm_create", "mbpe_comm_destroy", "mbpe_comm_resident", "mbpe_train_sharded",
    "mbpe_write_vocab_karpathy", "mbpe_tokenizer_save_vocab_karpathy", "mbpe_train_text_sharded", "mbpe_pretok_merge_corpora", "mbpe_sharded_trainer_create", "mbpe_sharded_trainer_run", "mbpe_sharded_trainer_destroy", "mbpe_encoder_launches",
    "mbpe_encoder_create", "mbpe_encoder_destroy", "mbpe_encoder_set_specials", "mbpe_encode", "mbpe_encode_device",
    "mbpe_encode_reserve", "mbpe_decode", "mbpe_decode_device", "mbpe_encoder_seed_special_chunks",
    "mbpe_gpt2_split_pattern", "mbpe_gpt4_split_pattern", "mbpe_tokenizer_create", "mbpe_tokenizer_destroy",
    "mbpe_tokenizer_set_special_tokens", "mbpe_tokenizer_train", "mbpe_tokenizer_save", "mbpe_tokenizer_load",
    "mbpe_tokenizer_encode", "mbpe_tokenizer_decode", "mbpe_tokenizer_get_merges",
    "mbpe_tokenizer_last_train_stats", "mbpe_tokenizer_last_split_on_gpu", "mbpe_tokenizer_encode_file", "mbpe_encode_file", "mbpe_tokenizer_decode_file", "mbpe_decode_file", "mbpe_tokenizer_set_engine", "mbpe_tokenizer_set_threads",
    "mbpe_split", "mbpe_special_split", "mbpe_pretok_class_table", "mbpe_dedup",
    "mbpe_pretok_create", "mbpe_pretok_select", "mbpe_pretok_destroy", "mbpe_pretok_split_device", "mbpe_pretok_split",
    "mbpe_pretok_dedup_device", "mbpe_pretok_dedup_segments", "mbpe_pretok_corpus", "mbpe_encode_text",
    "mbpe_encode_text_special", "mbpe_pretok_split_device_parts", "mbpe_device_corpus_download", "mbpe_device_corpus_free",
    "mbpe_trainer_create_device", "mbpe_split_dedup", "mbpe_plan_shards", "mbpe_write_model", "mbpe_synth_corpus", "mbpe_synth_corpus_at",
]


class MbpeError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"mbpe error {code}: {msg}")
        self.code = code


class TrainStats(C.Structure):
    _fields_ = [("gpu_ms", C.c_double), ("build_ms", C.c_double), ("n_positions", C.c_uint64),
                ("n_pairs", C.c_uint64), ("table_slots", C.c_uint64), ("n_launches", C.c_uint64),
                ("n_big_merges", C.c_uint64), ("n_rebuilds", C.c_uint64), ("n_grows", C.c_uint64),
                ("rescan_bytes", C.c_uint64), ("resident_cycles", C.c_uint64 * 8)]

    def as_dict(self):
        d = {k: getattr(self, k) for k, _ in self._fields_ if k != "resident_cycles"}
        d["resident_cycles"] = dict(zip(["select", "hits", "mutate_alloc", "seg_fill", "fin", "steps", "total", "_"],
                                        list(self.resident_cycles)))
        return d


_lib = None


def lib():
    """Load the library or raise: there is no fallback implementation."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python minbpe-cc_b200/build.py` (nvcc, sm_100a)")
        L = C.CDLL(LIB_PATH)
        L.mbpe_version.restype = C.c_char_p
        L.mbpe_last_error.restype = C.c_char_p
        L.mbpe_gpt2_split_pattern.restype = C.c_char_p
        L.mbpe_gpt4_split_pattern.restype = C.c_char_p
        L.mbpe_tokenizer_destroy.restype = None
        L.mbpe_trainer_destroy.restype = None
        L.mbpe_encoder_destroy.restype = None
        L.mbpe_paircount_destroy.restype = None
        L.mbpe_comm_destroy.restype = None
        L.mbpe_tokenizer_set_engine.restype = None
        L.mbpe_tokenizer_set_threads.restype = None
        _lib = L
    return _lib


def _ck(rc):
    if rc != 0:
        raise MbpeError(rc, lib().mbpe_last_error().decode(errors="replace"))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _u8(b):
    return np.frombuffer(b, dtype=np.uint8) if len(b) else np.zeros(1, np.uint8)


def device_count():
    return lib().mbpe_device_count()


def patterns():
    L = lib()
    return {"basic": "", "gpt2": L.mbpe_gpt2_split_pattern().decode(), "gpt4": L.mbpe_gpt4_split_pattern().decode()}


# ---- 1. PairCount seam ----------------------------------------------------------------------------------
class PairCount:
    """PairCount<T> (PairCount.h:27-47) on the device."""

    def __init__(self, mode, device=0):
        self.h = C.c_void_p()
        _ck(lib().mbpe_paircount_create(MODE[mode], device, C.byref(self.h)))

    def create_or_modify_pair(self, a, b, freq):
        self.add([(a, b, freq)])

    def add(self, ops):
        arr = np.asarray(ops, np.int64).reshape(-1, 3)
        a = np.ascontiguousarray(arr[:, 0], np.uint32)
        b = np.ascontiguousarray(arr[:, 1], np.uint32)
        d = np.ascontiguousarray(arr[:, 2], np.int32)
        _ck(lib().mbpe_paircount_add(self.h, _p(a, C.c_uint32), _p(b, C.c_uint32), _p(d, C.c_int32), C.c_uint64(len(a))))

    def get_top_pair_count(self):
        a, b, c, f = C.c_uint32(), C.c_uint32(), C.c_int32(), C.c_int()
        _ck(lib().mbpe_paircount_top(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(f)))
        return (a.value, b.value) if f.value else None

    def get_pair(self, pair):
        c, f = C.c_int32(), C.c_int()
        _ck(lib().mbpe_paircount_get(self.h, C.c_uint32(pair[0]), C.c_uint32(pair[1]), C.byref(c), C.byref(f)))
        return c.value if f.value else None

    def get_count(self):
        n = C.c_uint64()
        _ck(lib().mbpe_paircount_size(self.h, C.byref(n)))
        return n.value

    def close(self):
        if self.h:
            lib().mbpe_paircount_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- 2. train merge loop --------------------------------------------------------------------------------
class Trainer:
    """Resident deduplicated corpus + repeatable merge loop (Tokenizer.h:551-589)."""

    def __init__(self, tokens, off, weight=None, device=0):
        self.tokens = np.ascontiguousarray(tokens, np.uint32)
        self.off = np.ascontiguousarray(off, np.uint64)
        self.weight = None if weight is None else np.ascontiguousarray(weight, np.uint32)
        self.h = C.c_void_p()
        tok = self.tokens if len(self.tokens) else np.zeros(1, np.uint32)
        _ck(lib().mbpe_trainer_create(_p(tok, C.c_uint32), C.c_uint64(len(self.tokens)), _p(self.off, C.c_uint64),
                                      C.c_uint64(len(self.off) - 1),
                                      None if self.weight is None else _p(self.weight, C.c_uint32), device,
                                      C.byref(self.h)))

    def run(self, vocab_size, mode, engine="persistent", stream=None):
        n = max(vocab_size - 256, 1)
        merges = np.zeros((n, 2), np.uint32)
        counts = np.zeros(n, np.int32)
        nm = C.c_uint32()
        st = TrainStats()
        _ck(lib().mbpe_trainer_run(self.h, C.c_uint32(vocab_size), MODE[mode] if isinstance(mode, str) else mode,
                                   ENGINE[engine] if isinstance(engine, str) else engine,
                                   C.c_void_p(stream or 0), _p(merges, C.c_uint32), _p(counts, C.c_int32),
                                   C.byref(nm), C.byref(st)))
        return merges[:nm.value].copy(), counts[:nm.value].copy(), st.as_dict()

    def close(self):
        if self.h:
            lib().mbpe_trainer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceCorpus(C.Structure):
    """mbpe_device_corpus: unique chunks resident on the GPU in the trainer's input layout."""
    _fields_ = [("d_tokens", C.c_void_p), ("d_off", C.c_void_p), ("d_weight", C.c_void_p), ("n_tokens", C.c_uint64),
                ("n_unique", C.c_uint64), ("n_chunks", C.c_uint64), ("device", C.c_int)]

    def download(self):
        tokens = np.zeros(max(self.n_tokens, 1), np.uint32)
        off = np.zeros(self.n_unique + 1, np.uint64)
        w = np.zeros(max(self.n_unique, 1), np.uint32)
        _ck(lib().mbpe_device_corpus_download(C.byref(self), _p(tokens, C.c_uint32), _p(off, C.c_uint64), _p(w, C.c_uint32)))
        return tokens[:self.n_tokens], off, w[:self.n_unique]

    def trainer(self):
        t = Trainer.__new__(Trainer)
        t.h = C.c_void_p()
        _ck(lib().mbpe_trainer_create_device(C.byref(self), C.byref(t.h)))
        return t

    def free(self):
        lib().mbpe_device_corpus_free(C.byref(self))


class Pretok:
    """GPU pre-tokeniser + chunk dedup for the GPT-4 split pattern (Tokenizer.h:500-544)."""

    def __init__(self, device=0, pattern=None):
        self.h = C.c_void_p()
        _ck(lib().mbpe_pretok_create(device, C.byref(self.h)))
        if pattern is not None:
            _ck(lib().mbpe_pretok_select(self.h, pattern.encode()))

    def split(self, text: bytes):
        """chunk offsets (n_chunks + 1, u64): chunk c = text[off[c]:off[c+1]]"""
        buf = _u8(text)
        off = np.zeros(len(text) + 2, np.uint64)
        n = C.c_uint64()
        _ck(lib().mbpe_pretok_split(self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)), _p(off, C.c_uint64),
                                    C.c_uint64(len(off)), C.byref(n)))
        return off[:n.value + 1].copy()

    def split_device(self, d_text, n_bytes, d_off, off_cap, stream=None):
        n = C.c_uint64()
        _ck(lib().mbpe_pretok_split_device(self.h, C.c_void_p(d_text), C.c_uint64(n_bytes), C.c_void_p(d_off),
                                           C.c_uint64(off_cap), C.byref(n), C.c_void_p(stream or 0)))
        return n.value

    def dedup_device(self, d_text, n_bytes, d_off, n_chunks, stream=None):
        c = DeviceCorpus()
        _ck(lib().mbpe_pretok_dedup_device(self.h, C.c_void_p(d_text), C.c_uint64(n_bytes), C.c_void_p(d_off),
                                           C.c_uint64(n_chunks), C.byref(c), C.c_void_p(stream or 0)))
        return c

    def corpus(self, text: bytes):
        buf = _u8(text)
        c = DeviceCorpus()
        _ck(lib().mbpe_pretok_corpus(self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)), C.byref(c)))
        return c

    def encode_text(self, encoder, text: bytes):
        """GPT-4 split + merge scan of host text, ids back in host memory"""
        buf = _u8(text)
        out = np.zeros(max(len(text), 1), np.uint32)
        n = C.c_uint64()
        _ck(lib().mbpe_encode_text(encoder.h, self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)), _p(out, C.c_uint32),
                                   C.c_uint64(len(out)), C.byref(n)))
        return out[:n.value].copy()

    def encode_text_special(self, encoder, text: bytes, sp_begin, sp_end):
        """... with special-token occurrences [sp_begin[i], sp_end[i]) in text order (encoder.seed_special_chunks first)"""
        buf = _u8(text)
        sb, se = np.ascontiguousarray(sp_begin, np.uint64), np.ascontiguousarray(sp_end, np.uint64)
        out = np.zeros(max(len(text), 1), np.uint32)
        n = C.c_uint64()
        _ck(lib().mbpe_encode_text_special(encoder.h, self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)),
                                           _p(sb if len(sb) else np.zeros(1, np.uint64), C.c_uint64),
                                           _p(se if len(se) else np.zeros(1, np.uint64), C.c_uint64), C.c_uint64(len(sb)),
                                           _p(out, C.c_uint32), C.c_uint64(len(out)), C.byref(n)))
        return out[:n.value].copy()

    def close(self):
        if self.h:
            lib().mbpe_pretok_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def train(tokens, off, weight, vocab_size, mode, engine="persistent", device=0):
    t = Trainer(tokens, off, weight, device)
    try:
        return t.run(vocab_size, mode, engine)
    finally:
        t.close()


class Comm:
    """NCCL communicator for sharded training (one process per GPU). `bcast` moves rank 0's 128-byte id to the others."""

    def __init__(self, rank, world, device, bcast):
        ident = np.zeros(128, np.uint8)
        if rank == 0:
            _ck(lib().mbpe_comm_unique_id(_p(ident, C.c_uint8)))
        ident = np.ascontiguousarray(bcast(ident), np.uint8)
        self.h = C.c_void_p()
        self.rank, self.world, self.device = rank, world, device
        _ck(lib().mbpe_comm_create(_p(ident, C.c_uint8), rank, world, device, C.byref(self.h)))

    def train(self, tokens, off, weight, vocab_size, mode, stream=None):
        tokens = np.ascontiguousarray(tokens, np.uint32)
        off = np.ascontiguousarray(off, np.uint64)
        weight = None if weight is None else np.ascontiguousarray(weight, np.uint32)
        n = max(vocab_size - 256, 1)
        merges = np.zeros((n, 2), np.uint32)
        counts = np.zeros(n, np.int32)
        nm = C.c_uint32()
        st = TrainStats()
        tok = tokens if len(tokens) else np.zeros(1, np.uint32)
        _ck(lib().mbpe_train_sharded(self.h, _p(tok, C.c_uint32), C.c_uint64(len(tokens)), _p(off, C.c_uint64),
                                     C.c_uint64(len(off) - 1), None if weight is None else _p(weight, C.c_uint32),
                                     C.c_uint32(vocab_size), MODE[mode] if isinstance(mode, str) else mode,
                                     C.c_void_p(stream or 0), _p(merges, C.c_uint32), _p(counts, C.c_int32), C.byref(nm),
                                     C.byref(st)))
        return merges[:nm.value].copy(), counts[:nm.value].copy(), st.as_dict()

    def train_text(self, text_part, vocab_size, mode, pretok=None):
        """Tokenizer::train over a text sharded over the ranks: text_part = this rank's contiguous part (bytes or uint8 array)"""
        buf = text_part if isinstance(text_part, np.ndarray) else _u8(text_part)
        own = pretok is None
        pt = Pretok(device=self.device) if own else pretok
        n = max(vocab_size - 256, 1)
        merges = np.zeros((n, 2), np.uint32)
        counts = np.zeros(n, np.int32)
        nm = C.c_uint32()
        st = TrainStats()
        fe = C.c_double()
        try:
            _ck(lib().mbpe_train_text_sharded(self.h, pt.h, _p(buf, C.c_uint8), C.c_uint64(len(text_part)), C.c_uint32(vocab_size),
                                              MODE[mode] if isinstance(mode, str) else mode, _p(merges, C.c_uint32),
                                              _p(counts, C.c_int32), C.byref(nm), C.byref(st), C.byref(fe)))
        finally:
            if own:
                pt.close()
        d = st.as_dict()
        d["front_end_s"] = fe.value
        return merges[:nm.value].copy(), counts[:nm.value].copy(), d

    def resident(self):
        """True: the ranks mapped each other's inboxes, merges run in one resident CTA per rank that exchanges by itself"""
        return bool(lib().mbpe_comm_resident(self.h))

    def close(self):
        if self.h:
            lib().mbpe_comm_destroy(self.h)
            self.h = C.c_void_p()


class ShardedTrainer:
    """This rank's share of the deduplicated corpus resident on its GPU (every rank passes the same corpus); run() can be repeated."""

    def __init__(self, comm, tokens, off, weight):
        tokens = np.ascontiguousarray(tokens, np.uint32)
        off = np.ascontiguousarray(off, np.uint64)
        weight = None if weight is None else np.ascontiguousarray(weight, np.uint32)
        self.h = C.c_void_p()
        tok = tokens if len(tokens) else np.zeros(1, np.uint32)
        _ck(lib().mbpe_sharded_trainer_create(comm.h, _p(tok, C.c_uint32), C.c_uint64(len(tokens)), _p(off, C.c_uint64),
                                              C.c_uint64(len(off) - 1), None if weight is None else _p(weight, C.c_uint32),
                                              C.byref(self.h)))

    def run(self, vocab_size, mode, stream=None, engine="persistent"):
        n = max(vocab_size - 256, 1)
        merges = np.zeros((n, 2), np.uint32)
        counts = np.zeros(n, np.int32)
        nm = C.c_uint32()
        st = TrainStats()
        _ck(lib().mbpe_sharded_trainer_run(self.h, C.c_uint32(vocab_size), MODE[mode] if isinstance(mode, str) else mode,
                                           ENGINE[engine] if isinstance(engine, str) else engine, C.c_void_p(stream or 0),
                                           _p(merges, C.c_uint32), _p(counts, C.c_int32), C.byref(nm), C.byref(st)))
        return merges[:nm.value].copy(), counts[:nm.value].copy(), st.as_dict()

    def close(self):
        if self.h:
            lib().mbpe_sharded_trainer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- 3. encode / decode ---------------------------------------------------------------------------------
class Encoder:
    def __init__(self, merges, device=0):
        m = np.ascontiguousarray(merges, np.uint32).reshape(-1, 2)
        self.n_merges = len(m)
        self.h = C.c_void_p()
        mm = m if len(m) else np.zeros((1, 2), np.uint32)
        _ck(lib().mbpe_encoder_create(_p(mm, C.c_uint32), C.c_uint32(len(m)), device, C.byref(self.h)))

    def set_specials(self, specials):
        """specials: dict id -> bytes"""
        ids = np.asarray(list(specials.keys()) or [0], np.uint32)
        blob = b"".join(specials.values())
        off = np.zeros(len(specials) + 1, np.uint64)
        if specials:
            off[1:] = np.cumsum([len(v) for v in specials.values()])
        _ck(lib().mbpe_encoder_set_specials(self.h, _p(ids, C.c_uint32), _p(_u8(blob), C.c_uint8), _p(off, C.c_uint64),
                                            C.c_uint32(len(specials))))

    def encode(self, data: bytes, chunk_off, want_off=False, out=None):
        """out: optional caller-owned uint32 buffer (>= len(data) words are always enough); the ids are then returned as a
        view of it -- no allocation, no first-touch page faults and no copy inside the call"""
        off = np.ascontiguousarray(chunk_off, np.uint64)
        if out is not None:
            if want_off:
                raise ValueError("out= and want_off=True are mutually exclusive")
            if out.dtype != np.uint32 or not out.flags["C_CONTIGUOUS"]:
                raise ValueError("out must be a C-contiguous uint32 array")
            n = C.c_uint64()
            _ck(lib().mbpe_encode(self.h, _p(_u8(data), C.c_uint8), C.c_uint64(len(data)), _p(off, C.c_uint64),
                                  C.c_uint64(len(off) - 1), _p(out, C.c_uint32), C.c_uint64(len(out)), C.byref(n), None))
            return out[:n.value]
        out = np.zeros(max(len(data), 1), np.uint32)
        out_off = np.zeros(len(off), np.uint64) if want_off else None
        n = C.c_uint64()
        _ck(lib().mbpe_encode(self.h, _p(_u8(data), C.c_uint8), C.c_uint64(len(data)), _p(off, C.c_uint64),
                              C.c_uint64(len(off) - 1), _p(out, C.c_uint32), C.c_uint64(len(out)), C.byref(n),
                              None if out_off is None else _p(out_off, C.c_uint64)))
        return (out[:n.value].copy(), out_off) if want_off else out[:n.value].copy()

    def seed_special_chunks(self, tokens):
        """tokens: [(bytes, id)]: a chunk with exactly these bytes encodes to this one id (special tokens)"""
        ids = np.asarray([i for _, i in tokens] or [0], np.uint32)
        blob = b"".join(t for t, _ in tokens)
        off = np.zeros(len(tokens) + 1, np.uint64)
        off[1:] = np.cumsum([len(t) for t, _ in tokens])
        _ck(lib().mbpe_encoder_seed_special_chunks(self.h, _p(ids, C.c_uint32), _p(_u8(blob), C.c_uint8), _p(off, C.c_uint64),
                                                   C.c_uint32(len(tokens))))

    def decode_device(self, d_ids, n_ids, d_out, out_cap, d_n_out, stream=None):
        """All pointers are device addresses (ints); d_out may be 0 to size only."""
        _ck(lib().mbpe_decode_device(self.h, C.c_void_p(d_ids), C.c_uint64(n_ids), C.c_void_p(d_out), C.c_uint64(out_cap),
                                     C.c_void_p(d_n_out), C.c_void_p(stream or 0)))

    def encode_device(self, d_bytes, n_bytes, d_off32, n_chunks, d_out, out_cap, d_n_out, stream=None):
        """All pointers are device addresses (ints)."""
        _ck(lib().mbpe_encode_device(self.h, C.c_void_p(d_bytes), C.c_uint64(n_bytes), C.c_void_p(d_off32),
                                     C.c_uint64(n_chunks), C.c_void_p(d_out), C.c_uint64(out_cap), C.c_void_p(d_n_out),
                                     C.c_void_p(stream or 0)))

    def reserve(self, n_bytes, n_chunks):
        _ck(lib().mbpe_encode_reserve(self.h, C.c_uint64(n_bytes), C.c_uint64(n_chunks)))

    def launches(self):
        """kernels launched through this encoder so far"""
        lib().mbpe_encoder_launches.restype = C.c_uint64
        return int(lib().mbpe_encoder_launches(self.h))

    def decode(self, ids):
        ids = np.ascontiguousarray(ids, np.uint32)
        src = ids if len(ids) else np.zeros(1, np.uint32)
        n = C.c_uint64()
        _ck(lib().mbpe_decode(self.h, _p(src, C.c_uint32), C.c_uint64(len(ids)), None, C.c_uint64(0), C.byref(n)))
        out = np.zeros(max(n.value, 1), np.uint8)
        if n.value:
            _ck(lib().mbpe_decode(self.h, _p(src, C.c_uint32), C.c_uint64(len(ids)), _p(out, C.c_uint8),
                                  C.c_uint64(n.value), C.byref(n)))
        return out[:n.value].tobytes()

    def close(self):
        if self.h:
            lib().mbpe_encoder_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- 4. Tokenizer mirror --------------------------------------------------------------------------------
class Tokenizer:
    """MinBpeCC::Tokenizer::Tokenizer (Tokenizer.h:379-927), same method names and argument meaning."""

    FIRST, LEXICAL = 0, 1

    def __init__(self, pattern="", device=0):
        self.h = C.c_void_p()
        _ck(lib().mbpe_tokenizer_create(pattern.encode(), device, C.byref(self.h)))

    def set_special_tokens_from_file(self, contents):
        b = contents if isinstance(contents, bytes) else contents.encode()
        _ck(lib().mbpe_tokenizer_set_special_tokens(self.h, b, C.c_uint64(len(b))))

    def train(self, text, vocab_size, conflict_resolution, verbose=False):
        """text: bytes or a uint8 array (e.g. pinned memory)"""
        mode = MODE[conflict_resolution] if isinstance(conflict_resolution, str) else conflict_resolution
        buf = text if isinstance(text, np.ndarray) else _u8(text)
        _ck(lib().mbpe_tokenizer_train(self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)), int(vocab_size), mode,
                                       int(verbose)))

    def save(self, path, write_vocab=False):
        _ck(lib().mbpe_tokenizer_save(self.h, str(path).encode(), int(write_vocab)))

    def load(self, path, verbose=False):
        _ck(lib().mbpe_tokenizer_load(self.h, str(path).encode(), int(verbose)))

    def encode(self, text, verbose=False, out=None):
        """text: bytes or a uint8 array (e.g. pinned memory); out: optional uint32 array to receive the ids (a view of
        it is returned), otherwise a fresh array."""
        n = C.c_uint64()
        buf = text if isinstance(text, np.ndarray) else _u8(text)
        if out is None:
            tmp = np.empty(max(len(buf), 1), np.uint32)
            _ck(lib().mbpe_tokenizer_encode(self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)), _p(tmp, C.c_uint32),
                                            C.c_uint64(len(tmp)), C.byref(n)))
            return tmp[:n.value].copy()
        _ck(lib().mbpe_tokenizer_encode(self.h, _p(buf, C.c_uint8), C.c_uint64(len(text)), _p(out, C.c_uint32),
                                        C.c_uint64(len(out)), C.byref(n)))
        return out[:n.value]

    def encode_file(self, in_path, out_path):
        """streaming file -> .enc file; returns the number of ids written"""
        n = C.c_uint64()
        _ck(lib().mbpe_tokenizer_encode_file(self.h, str(in_path).encode(), str(out_path).encode(), C.byref(n)))
        return n.value

    def decode_file(self, in_path, out_path):
        """streaming .enc file -> text file; returns (ids read, bytes written)"""
        ni, nb = C.c_uint64(), C.c_uint64()
        _ck(lib().mbpe_tokenizer_decode_file(self.h, str(in_path).encode(), str(out_path).encode(), C.byref(ni), C.byref(nb)))
        return ni.value, nb.value

    def decode(self, ids, verbose=False):
        ids = np.ascontiguousarray(ids, np.uint32)
        src = ids if len(ids) else np.zeros(1, np.uint32)
        n = C.c_uint64()
        _ck(lib().mbpe_tokenizer_decode(self.h, _p(src, C.c_uint32), C.c_uint64(len(ids)), None, C.c_uint64(0),
                                        C.byref(n)))
        out = np.zeros(max(n.value, 1), np.uint8)
        _ck(lib().mbpe_tokenizer_decode(self.h, _p(src, C.c_uint32), C.c_uint64(len(ids)), _p(out, C.c_uint8),
                                        C.c_uint64(len(out)), C.byref(n)))
        return out[:n.value].tobytes()

    def merges(self):
        n = C.c_uint32()
        _ck(lib().mbpe_tokenizer_get_merges(self.h, None, 0, C.byref(n)))
        m = np.zeros((max(n.value, 1), 2), np.uint32)
        _ck(lib().mbpe_tokenizer_get_merges(self.h, _p(m, C.c_uint32), C.c_uint32(len(m)), C.byref(n)))
        return m[:n.value].copy()

    def last_train_stats(self):
        st = TrainStats()
        split, dedup, nc, nu = C.c_double(), C.c_double(), C.c_uint64(), C.c_uint64()
        _ck(lib().mbpe_tokenizer_last_train_stats(self.h, C.byref(st), C.byref(split), C.byref(dedup), C.byref(nc),
                                                  C.byref(nu)))
        d = st.as_dict()
        d.update(split_s=split.value, dedup_s=dedup.value, n_chunks=nc.value, n_unique=nu.value,
                 split_on_gpu=bool(lib().mbpe_tokenizer_last_split_on_gpu(self.h)))
        return d

    def set_engine(self, engine):
        lib().mbpe_tokenizer_set_engine(self.h, ENGINE[engine] if isinstance(engine, str) else engine)

    def set_threads(self, n):
        lib().mbpe_tokenizer_set_threads(self.h, int(n))

    def close(self):
        if self.h:
            lib().mbpe_tokenizer_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- 5. host-only helpers -------------------------------------------------------------------------------
def split(pattern: str, text: bytes, n_threads=0):
    n = C.c_uint64()
    buf = _u8(text)
    cap = len(text) + 1  # a chunk is at least one byte; untouched pages of the zeroed arrays cost nothing
    s = np.zeros(cap, np.uint64)
    e = np.zeros(cap, np.uint64)
    _ck(lib().mbpe_split(pattern.encode(), _p(buf, C.c_uint8), C.c_uint64(len(text)), n_threads, _p(s, C.c_uint64),
                         _p(e, C.c_uint64), C.c_uint64(cap), C.byref(n)))
    return s[:n.value].copy(), e[:n.value].copy()


def special_split(special_contents: bytes, text: bytes):
    """[(start, end, id)]; id < 0 = ordinary text (Tokenizer.h:605-650)"""
    n = C.c_uint64()
    buf = _u8(text)
    args = (special_contents, C.c_uint64(len(special_contents)), _p(buf, C.c_uint8), C.c_uint64(len(text)))
    _ck(lib().mbpe_special_split(*args, None, None, None, C.c_uint64(0), C.byref(n)))
    s, e, i = np.zeros(max(n.value, 1), np.uint64), np.zeros(max(n.value, 1), np.uint64), np.zeros(max(n.value, 1), np.int64)
    _ck(lib().mbpe_special_split(*args, _p(s, C.c_uint64), _p(e, C.c_uint64), _p(i, C.c_int64), C.c_uint64(len(s)), C.byref(n)))
    return [(int(s[k]), int(e[k]), int(i[k])) for k in range(n.value)]


def pretok_class_table():
    """2-bit class per code point as the linked PCRE2 sees it (0 other, 1 letter, 2 number, 3 white space)."""
    t = np.zeros(0x110000 // 4, np.uint8)
    _ck(lib().mbpe_pretok_class_table(_p(t, C.c_uint8)))
    return t


def dedup(text: bytes, starts, ends):
    starts = np.ascontiguousarray(starts, np.uint64)
    ends = np.ascontiguousarray(ends, np.uint64)
    total = int((ends - starts).sum())
    tokens = np.zeros(max(total, 1), np.uint32)
    off = np.zeros(len(starts) + 1, np.uint64)
    w = np.zeros(max(len(starts), 1), np.uint32)
    nt, nu = C.c_uint64(), C.c_uint64()
    _ck(lib().mbpe_dedup(_p(_u8(text), C.c_uint8), _p(starts, C.c_uint64), _p(ends, C.c_uint64),
                         C.c_uint64(len(starts)), _p(tokens, C.c_uint32), C.byref(nt), _p(off, C.c_uint64),
                         _p(w, C.c_uint32), C.byref(nu)))
    return tokens[:nt.value].copy(), off[:nu.value + 1].copy(), w[:nu.value].copy()


def split_dedup(pattern: str, text: bytes, n_threads=0):
    """regex split + dedup fused (the train front end). Returns (tokens, off, weight, n_chunks)."""
    buf = _u8(text)
    nt, nu, nc = C.c_uint64(), C.c_uint64(), C.c_uint64()
    args = (pattern.encode(), _p(buf, C.c_uint8), C.c_uint64(len(text)), n_threads)
    _ck(lib().mbpe_split_dedup(*args, None, C.c_uint64(0), C.byref(nt), None, None, C.c_uint64(0), C.byref(nu),
                               C.byref(nc)))
    tokens = np.zeros(max(nt.value, 1), np.uint32)
    off = np.zeros(nu.value + 1, np.uint64)
    w = np.zeros(max(nu.value, 1), np.uint32)
    _ck(lib().mbpe_split_dedup(*args, _p(tokens, C.c_uint32), C.c_uint64(len(tokens)), C.byref(nt), _p(off, C.c_uint64),
                               _p(w, C.c_uint32), C.c_uint64(len(w)), C.byref(nu), C.byref(nc)))
    return tokens[:nt.value], off, w[:nu.value], nc.value


def plan_shards(chunk_off, n_parts):
    """contiguous, byte-balanced chunk ranges for n_parts ranks; returns n_parts+1 chunk indices"""
    off = np.ascontiguousarray(chunk_off, np.uint64)
    out = np.zeros(n_parts + 1, np.uint64)
    _ck(lib().mbpe_plan_shards(_p(off, C.c_uint64), C.c_uint64(len(off) - 1), C.c_uint32(n_parts), _p(out, C.c_uint64)))
    return out


def shard_for_rank(text: bytes, chunk_off, rank, world):
    """this rank's slice of an encode job: (bytes, rebased chunk offsets, first chunk index)"""
    plan = plan_shards(chunk_off, world)
    c0, c1 = int(plan[rank]), int(plan[rank + 1])
    off = np.ascontiguousarray(chunk_off, np.uint64)[c0:c1 + 1]
    b0, b1 = int(off[0]), int(off[-1])
    return text[b0:b1], off - np.uint64(b0), c0


def write_model(path, pattern, special_contents, merges, write_vocab=False):
    m = np.ascontiguousarray(merges, np.uint32).reshape(-1, 2)
    sp = special_contents or b""
    _ck(lib().mbpe_write_model(str(path).encode(), pattern.encode(), sp, C.c_uint64(len(sp)), _p(m, C.c_uint32),
                               C.c_uint32(len(m)), int(write_vocab)))


def write_vocab_karpathy(path, special_contents, merges):
    """.vocab in karpathy/minbpe's layout"""
    m = np.ascontiguousarray(merges, np.uint32).reshape(-1, 2)
    sp = special_contents or b""
    _ck(lib().mbpe_write_vocab_karpathy(str(path).encode(), sp, C.c_uint64(len(sp)), _p(m, C.c_uint32), C.c_uint32(len(m))))


def synth_corpus(seed, n_bytes, n_threads=0, first_block=0, out=None):
    """bytes [first_block MiB, first_block MiB + n_bytes) of the synthetic corpus `seed` (into `out` if given)"""
    if out is None:
        out = np.empty(n_bytes, np.uint8)
    _ck(lib().mbpe_synth_corpus_at(C.c_uint64(seed), C.c_uint64(first_block), _p(out, C.c_uint8), C.c_uint64(n_bytes), n_threads))
    return out[:n_bytes]
