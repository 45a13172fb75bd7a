"""CPU: the train phases (minbpe-cc_b200/csrc/train_phases.cuh) and their orchestration (train_driver.hpp),
compiled for the host with a sequential backend (tests/emu/train_emu.cpp), against the reference goldens and
the oracle. This checks the algorithm the kernels run -- delta rules, a==b runs, both tie-breaks, candidate
list / rebuild / table growth, big-merge hand-off -- under ascending, descending and shuffled thread orders.
It is a development check, not a product path."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_data

EMU_SRC = os.path.join(ROOT, "tests", "emu", "train_emu.cpp")
EMU_LIB = os.path.join(ROOT, "tests", "emu", "libtrain_emu.so")


@pytest.fixture(scope="module")
def emu():
    deps = [EMU_SRC] + [os.path.join(ROOT, "minbpe-cc_b200", "csrc", f) for f in ("train_phases.cuh", "train_driver.hpp")]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-pthread", "-x", "c++", "-o", EMU_LIB, EMU_SRC])
    L = C.CDLL(EMU_LIB)

    def run(t, o, w, vocab, mode, engine=0, nth=64, order=0, big=1 << 30, want=1024, init=0):
        n = max(vocab - 256, 1)
        m = np.zeros((n, 2), np.uint32)
        c = np.zeros(n, np.int32)
        nm = C.c_uint32()
        st = np.zeros(8, np.uint64)
        P = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
        tt = t if len(t) else np.zeros(1, np.uint32)
        rc = L.emu_train(P(tt, C.c_uint32), C.c_uint64(len(t)), P(o, C.c_uint64), C.c_uint64(len(o) - 1),
                         P(w, C.c_uint32), C.c_uint32(vocab), {"first": 0, "lexical": 1}[mode], engine, nth, order,
                         C.c_uint32(big), C.c_uint32(want), C.c_uint32(init), P(m, C.c_uint32), P(c, C.c_int32),
                         C.byref(nm), P(st, C.c_uint64))
        assert rc == 0
        return m[:nm.value], c[:nm.value], dict(zip(["n_pairs", "slots", "n_big", "n_rebuilds", "n_grows",
                                                       "rescan_bytes", "launches", "device_merges"], st.tolist()))
    return run


CONFIGS = [  # engine, threads, order, big_limit, cand_want, init_slots
    (0, 64, 0, 1 << 30, 1024, 0),
    (1, 32, 2, 16, 8, 8192),
    (1, 7, 1, 1 << 30, 3, 16384),
    (0, 1, 0, 1 << 30, 1, 0),
]


def _cases():
    import json
    man = json.load(open(os.path.join(GOLDEN, "manifest.json")))
    return sorted(k for k, e in man["train"].items() if e["rc"] == 0)


@pytest.mark.parametrize("name", _cases())
def test_phases_match_reference(emu, oracle, manifest, name):
    e = manifest["train"][name]
    _, _, gm = oracle.read_model(os.path.join(GOLDEN, "models", name + ".model"))
    text = golden_data(e["input"])
    t, o, w = oracle.flatten(oracle.chunks_of(text, e["encoder"]), True)
    _, oc = oracle.train(t, o, w, e["vocab_size"], e["mode"])
    for cfg in CONFIGS:
        m, c, st = emu(t, o, w, e["vocab_size"], e["mode"], *cfg)
        assert m.shape == gm.shape and (m == gm).all(), (name, cfg)
        assert (c == oc).all(), (name, cfg)


def test_growth_rebuild_and_big_path_are_exercised(emu, oracle):
    text = golden_data("taylorswift.txt")
    t, o, w = oracle.flatten(oracle.chunks_of(text, "gpt4"), True)
    m, c, st = emu(t, o, w, 1500, "lexical", 1, 32, 2, 16, 8, 8192)
    assert st["n_grows"] >= 1 and st["n_rebuilds"] >= 2 and st["n_big"] >= 10
    om, oc = oracle.train(t, o, w, 1500, "lexical")
    assert (m == om).all() and (c == oc).all()


@pytest.mark.parametrize("mode", ["first", "lexical"])
def test_random_small_alphabets(emu, oracle, mode):
    """Tiny alphabets force a==b runs, overlapping occurrences ("abab") and heavy count ties."""
    rng = np.random.default_rng(11)
    for trial in range(60):
        k = int(rng.integers(1, 4))
        n_chunks = int(rng.integers(1, 30))
        chunks = [bytes(rng.integers(97, 97 + k, int(rng.integers(1, 25))).astype(np.uint8)) for _ in range(n_chunks)]
        t, o, w = oracle.flatten(chunks, True)
        vocab = 256 + int(rng.integers(1, 40))
        om, oc = oracle.train(t, o, w, vocab, mode, impl="rescan")
        for cfg in CONFIGS[:3]:
            m, c, _ = emu(t, o, w, vocab, mode, *cfg)
            assert m.shape == om.shape and (m == om).all() and (c == oc).all(), (trial, cfg, chunks)


def test_empty_and_pairless_inputs(emu):
    z = np.zeros(0, np.uint32)
    m, c, _ = emu(z, np.asarray([0], np.uint64), np.zeros(1, np.uint32), 300, "lexical")
    assert len(m) == 0
    t = np.asarray([97, 98, 99], np.uint32)  # three single-token chunks: no pairs at all
    m, c, _ = emu(t, np.asarray([0, 1, 2, 3], np.uint64), np.ones(3, np.uint32), 300, "first")
    assert len(m) == 0


# ---- sharded training: `world` emulated ranks, one host thread each -------------------------
def _emu_sharded(t, o, w, vocab, mode, world, nth=16, order=2, want=8, resident_limit=0):
    L = C.CDLL(EMU_LIB)
    n = max(vocab - 256, 1)
    m = np.zeros((n, 2), np.uint32)
    c = np.zeros(n, np.int32)
    nm = C.c_uint32()
    P = lambda a, ty: a.ctypes.data_as(C.POINTER(ty))
    rc = L.emu_train_sharded(P(t, C.c_uint32), C.c_uint64(len(t)), P(o, C.c_uint64), C.c_uint64(len(o) - 1),
                             P(w, C.c_uint32), C.c_uint32(vocab), {"first": 0, "lexical": 1}[mode], world, nth, order,
                             C.c_uint32(want), C.c_uint32(resident_limit), P(m, C.c_uint32), P(c, C.c_int32), C.byref(nm))
    return rc, m[:nm.value], c[:nm.value]


@pytest.mark.parametrize("name", ["ts512_gpt4_first", "ts512_gpt4_lexical", "sample512_gpt4_first",
                                  "sample512_gpt4_lexical", "str_runs_c_gpt4_first", "str_runs_c_basic_lexical",
                                  "str_exhaust_basic_first", "str_exhaust_basic_lexical", "str_unicode_gpt4_first",
                                  "ts400_basic_first"])
def test_sharded_ranks_agree_and_match_reference(emu, oracle, manifest, name):
    """Every rank keeps a replica of the pair table with global counts and its share of the chunks; after the
    per-step exchange of count deltas all ranks must pick the same merges -- and they must be the reference's.
    rc 7 = ranks disagreed. encoder 'basic' = one chunk: one rank owns everything, the others own nothing."""
    e = manifest["train"][name]
    _, _, gm = oracle.read_model(os.path.join(GOLDEN, "models", name + ".model"))
    text = golden_data(e["input"])
    t, o, w = oracle.flatten(oracle.chunks_of(text, e["encoder"]), True)
    _, oc = oracle.train(t, o, w, e["vocab_size"], e["mode"])
    # resident_limit 0: every step driven from the host; 40: merges with a count <= 40 run in the resident program (the
    # exchange happens inside it, FIRST mode's extra exchange too) and the others fall back; 1 << 30: all resident
    for world, limit in ((2, 0), (3, 0), (2, 40), (3, 1 << 30), (2, 1 << 30)):
        rc, m, c = _emu_sharded(t, o, w, e["vocab_size"], e["mode"], world, resident_limit=limit)
        assert rc == 0, (name, world, limit, rc)
        assert m.shape == gm.shape and (m == gm).all() and (c == oc).all(), (name, world, limit)
