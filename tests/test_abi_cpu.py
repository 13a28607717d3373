"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/b200zk.h
declares, and refuses to compute without a GPU (no CPU fallback)."""
import os, re, ctypes
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import _lib
    return _lib.lib()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "b200zk.h")).read()
    names = set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 20
    for n in sorted(names):
        assert hasattr(lib, n), "symbol %s declared in include/b200zk.h is not exported" % n
    from eigen_zkvm_b200 import _lib
    assert set(_lib.EXPORTS) == names


def test_version_and_n_nodes(lib):
    assert b"sm_100a" in lib.b200_version()
    # merklehash.rs:47-61
    for h, exp in [(1, 1), (2, 3), (33, 34 + 18 + 10 + 6 + 4 + 2 + 1), (256, 511)]:
        n = h; acc = 0
        nn = (n - 1) // 2 + 1; acc = nn * 2
        while n > 1:
            n = nn; nn = (n - 1) // 2 + 1
            acc += nn * 2 if n > 1 else 1
        assert lib.b200_gl_merkle_n_nodes(h) == acc
    assert lib.b200_gl_merkle_n_nodes(256) == 511


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    a = np.arange(8, dtype=np.uint64); out = np.zeros(8, dtype=np.uint64)
    rc = lib.b200_gl_ntt(a.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), 1, 3)
    assert rc == -2 and b"no CPU fallback" in lib.b200_last_error()
    h = ctypes.c_void_p()
    rc = lib.b200_setup_new(b"{}", a.ctypes.data_as(ctypes.c_void_p), 8, 1, ctypes.byref(h))
    assert rc != 0


def test_transcript_poseidon_host_code_matches_the_reference_kats(lib):
    """csrc/poseidon_host.cpp (the transcript's permutation, host code inside the product library) against the KATs of
    starky/src/poseidon_opt.rs:219-262 and against the oracle on random states."""
    from oracle import gl
    P = gl.P
    def run(st):
        a = np.array(st, dtype=np.uint64); o = np.zeros(12, dtype=np.uint64)
        assert lib.b200_debug_transcript_poseidon(a.ctypes.data_as(ctypes.c_void_p), o.ctypes.data_as(ctypes.c_void_p)) == 0
        return [int(x) for x in o]
    out = run([0] * 12)
    assert out[:4] == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca]      # poseidon_opt.rs:225-230
    out = run(list(range(8)) + [0] * 4)
    assert out[:4] == [int(x) for x in gl.poseidon(list(range(8)), [0, 0, 0, 0])[:4]]
    rng = np.random.default_rng(7)
    for _ in range(10):
        st = [int(x) % P for x in rng.integers(0, 2**63, size=12, dtype=np.uint64) * 2 + 1]
        assert run(st) == [int(x) for x in gl.poseidon(st[:8], st[8:])]


def test_msm_window_choice_host_logic(lib):
    """The window rule of the multiexp (msm.cu msm_pick_c) is host logic: windows cover the scalar plus the signed-digit carry, the table mode
    never takes a width whose TOP window is narrow (all n digits of that window would share a handful of buckets: serialised atomics and
    serial bucket runs -- measured 43 ms instead of 12 for 2^22 points at c = 19), and the sizes of the 1 / 2 / 4 / 8-GPU shares of the
    2^22-point benchmark get the measured choices."""
    import ctypes
    bits = {0: 254, 1: 254, 2: 255, 3: 255}
    c = ctypes.c_uint(); w = ctypes.c_uint()
    for curve, sb in bits.items():
        for lg in range(0, 27):
            for table in (0, 1):
                assert lib.b200_debug_msm_window(curve, 1 << lg, table, ctypes.byref(c), ctypes.byref(w)) == 0
                assert 6 <= c.value <= 22 and w.value * c.value >= sb + 1 and (w.value - 1) * c.value < sb + 1
                if table and c.value > 12:
                    assert sb - (w.value - 1) * c.value >= 12, (curve, lg, c.value, w.value)
                if lg >= 1:          # wider windows for larger problems (monotone)
                    prev = ctypes.c_uint(); pw = ctypes.c_uint()
                    lib.b200_debug_msm_window(curve, 1 << (lg - 1), table, ctypes.byref(prev), ctypes.byref(pw))
                    assert prev.value <= c.value
    got = {}
    for lg in (19, 20, 21, 22):
        lib.b200_debug_msm_window(0, 1 << lg, 1, ctypes.byref(c), ctypes.byref(w)); got[lg] = (c.value, w.value)
    assert got == {19: (16, 16), 20: (17, 15), 21: (17, 15), 22: (17, 15)}
    assert lib.b200_debug_msm_window(7, 16, 1, ctypes.byref(c), ctypes.byref(w)) != 0          # unknown curve id
