"""Host-only checks of the step-program JIT: the generated CUDA for every step program of the fixtures compiles with NVRTC
to an sm_100a cubin on the CPU box (no GPU needed); the GPU parity tests then run the proofs through those kernels."""
import ctypes, json, os
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import _lib
    return _lib.lib()


def _source(L, pilf, ssf, which):
    from eigen_zkvm_b200 import _lib, starkinfo as si
    pil = si.load_pil(os.path.join(G, pilf)); ss = json.load(open(os.path.join(G, ssf)))
    info, prog = si.new_starkinfo(pil, ss)
    out = ctypes.c_void_p(); ln = ctypes.c_size_t()
    _lib.check(L.b200_debug_step_program_source(si.setup_json(info, prog, ss).encode(), which.encode(), ctypes.byref(out), ctypes.byref(ln)))
    return _lib.take_string(out, ln), prog


@pytest.mark.parametrize("pilf,ssf", [("fib.pil.json.gl", "starkStruct.json.gl"), ("plookup.pil.json.gl", "starkStruct.json.gl"),
                                      ("connection.pil.json", "starkStruct.json.gl"), ("fib.pil.json", "starkStruct.json")])
def test_generated_sources_compile(L, pilf, ssf):
    try:
        ctypes.CDLL("libnvrtc.so.12")
    except OSError:
        pytest.skip("libnvrtc not installed")
    for which in ("step2prev", "step3prev", "step3", "step42ns", "step52ns"):
        src, prog = _source(L, pilf, ssf, which)
        n_ops = len(prog[which]["first"])
        assert src.count("\n    { ") == n_ops                      # one statement block per op of the step program
        if n_ops == 0:
            continue
        n = ctypes.c_size_t()
        rc = L.b200_debug_jit_compile(src.encode(), ctypes.byref(n))
        assert rc == 0 and n.value > 1000, L.b200_last_error()
    with pytest.raises(Exception):
        from eigen_zkvm_b200 import _lib
        _lib.check(L.b200_debug_jit_compile(b"this is not cuda", ctypes.byref(ctypes.c_size_t())))


def test_columns_a_program_writes_are_never_read_through_the_read_only_path(L):
    """plookup step3prev / step3 store to tmpExp / cm columns and read them back in the same program; a non-coherent
    load (__ldg, ld.global.nc) of data written by the same kernel is undefined.  The generator must use ordinary loads for every
    (section, column) that is a destination somewhere in the program, and may keep __ldg for the rest."""
    import re
    for which in ("step3prev", "step3"):
        src, _ = _source(L, "plookup.pil.json.gl", "starkStruct.json.gl", which)
        stores = set(re.findall(r"\*\((secs\.s\[\d+\]\.base \+ \(size_t\)\d+ \* secs\.s\[\d+\]\.rows) \+ ip?\) = ", src))
        ldg = set(re.findall(r"__ldg\((secs\.s\[\d+\]\.base \+ \(size_t\)\d+ \* secs\.s\[\d+\]\.rows) \+ ip?\)", src))
        plain = set(re.findall(r"\(\*\(const volatile u64\*\)\((secs\.s\[\d+\]\.base \+ \(size_t\)\d+ \* secs\.s\[\d+\]\.rows) \+ ip?\)\)", src))
        assert stores and not (stores & ldg), "a written column is loaded with __ldg"
        assert plain and plain <= stores | plain
        assert stores & plain, "the program reads back what it wrote (that is the case this test is about)"
    src, _ = _source(L, "fib.pil.json.gl", "starkStruct.json.gl", "step42ns")
    assert "volatile" not in src and "__ldg(secs" in src          # nothing read is written: all loads stay on the read-only path
