"""Error behaviour of the C-ABI on a GPU box: bad arguments come back as negative codes with a message (never a crash or
an exception across the boundary), mirroring where the reference panics / bails (stark_gen.rs:210,268; poseidon_bn128_opt.rs:112-118)."""
import ctypes, json, os
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import _lib
    return _lib.lib()


def _p(a): return a.ctypes.data_as(ctypes.c_void_p)


def test_bad_arguments_return_codes(L):
    a = np.arange(16, dtype=np.uint64); out = np.zeros(16, dtype=np.uint64)
    assert L.b200_gl_ntt(None, _p(out), 1, 3) == -1 and b"null" in L.b200_last_error()
    assert L.b200_gl_ntt(_p(a), _p(out), 1, 28) == -1                       # log size > 27
    assert L.b200_gl_lde(_p(a), _p(out), 1, 4, 3) == -1                      # nBitsExt < nBits
    assert L.b200_gl_ntt(_p(a), _p(out), 0, 3) == 0                          # empty input is a no-op (fft_p.rs:262-264)
    assert L.b200_gl_merkelize(_p(a), 2, 0, _p(out)) == -1                   # height 0
    assert L.b200_msm(7, _p(a), _p(a), 1, _p(out)) == -1 and L.b200_msm_point_bytes(7) == 0      # unknown curve id
    assert L.b200_msm(0, None, None, 5, _p(out)) == -1
    assert L.b200_big_poseidon(0, _p(a), 0, _p(a), _p(out)) == -1            # "Wrong inputs length"
    assert L.b200_big_poseidon(0, _p(np.zeros(17 * 4, dtype=np.uint64)), 17, _p(a), _p(np.zeros(18 * 4, dtype=np.uint64))) == -1
    assert L.b200_big_merkelize(5, _p(a), 2, 8, _p(out)) == -1               # unknown hash id
    h = ctypes.c_void_p()
    assert L.b200_setup_new(b"{not json", _p(a), 8, 1, ctypes.byref(h)) != 0
    assert L.b200_setup_new(b"{}", _p(a), 8, 1, ctypes.byref(h)) != 0


def test_setup_and_prove_shape_checks(L):
    from eigen_zkvm_b200 import starky, starkinfo as si, _lib
    pil = si.load_pil(os.path.join(G, "fib.pil.json.gl"))
    ss = json.load(open(os.path.join(G, "starkStruct.json.gl")))
    cm = np.fromfile(os.path.join(G, "fib.cm.gl"), dtype="<u8"); const = np.fromfile(os.path.join(G, "fib.const.gl"), dtype="<u8")
    with pytest.raises(ValueError):                                         # const_pol.nPols != pil.nConstants
        starky.StarkSetup.new(const[:-1], si.load_pil(os.path.join(G, "fib.pil.json.gl")), ss)    # (a fresh PIL each time: StarkInfo::new mutates it)
    bad = dict(ss); bad["verificationHashType"] = "KECCAK"
    with pytest.raises(_lib.B200Error) as e:
        starky.StarkSetup.new(const, si.load_pil(os.path.join(G, "fib.pil.json.gl")), bad)
    assert e.value.code == -3                                               # B200_ERR_UNSUPPORTED
    bad = dict(ss); bad["steps"] = [{"nBits": 10}, {"nBits": 7}, {"nBits": 3}]
    with pytest.raises((ValueError, _lib.B200Error)):                       # MustEqualDegreeError (stark_gen.rs:209-211 / starkinfo.rs)
        starky.StarkSetup.new(const, si.load_pil(os.path.join(G, "fib.pil.json.gl")), bad)
    setup = starky.StarkSetup.new(const, pil, ss)
    out = ctypes.c_void_p(); ln = ctypes.c_size_t()
    rc = L.b200_stark_gen(setup._h, _p(cm), 1024, 3, b"", ctypes.byref(out), ctypes.byref(ln))      # wrong column count
    assert rc != 0 and b"shape" in L.b200_last_error()
    rc = L.b200_stark_gen(setup._h, _p(cm), 512, 2, b"", ctypes.byref(out), ctypes.byref(ln))       # wrong row count
    assert rc != 0
    # a plookup whose f column contains a value missing from t: the reference panics (stark_gen.rs:640-646); here an error code
    ppil = si.load_pil(os.path.join(G, "plookup.pil.json.gl"))
    pcm = np.fromfile(os.path.join(G, "plookup.cm.gl"), dtype="<u8").copy(); pconst = np.fromfile(os.path.join(G, "plookup.const.gl"), dtype="<u8")
    psetup = starky.StarkSetup.new(pconst, ppil, ss)
    good = starky.StarkProof.stark_gen(pcm, psetup)
    assert good == open(os.path.join(G, "plookup10.proof.json")).read()


def test_non_canonical_field_inputs_are_reduced_on_ingest(L):
    """host buffers may hold any u64; the reference reduces on ingest (FGL::from), so values >= p must behave as their residues."""
    P = 0xFFFFFFFF00000001
    rng = np.random.default_rng(5)
    for w, bits in ((1, 6), (2, 7), (3, 5), (5, 12)):
        a = rng.integers(0, P, size=(1 << bits) * w, dtype=np.uint64)
        b = a.copy()
        b[::3] = np.where(a[::3] < np.uint64(2**32 - 1), a[::3] + np.uint64(P), a[::3])        # the non-canonical representative where one exists
        b[1] = np.uint64(P); a[1] = 0; b[2] = np.uint64(2**64 - 1); a[2] = np.uint64(2**32 - 2)
        oa = np.zeros_like(a); ob = np.zeros_like(a)
        assert L.b200_gl_ntt(_p(a), _p(oa), w, bits) == 0 and L.b200_gl_ntt(_p(b), _p(ob), w, bits) == 0
        assert (oa == ob).all() and (ob < np.uint64(P)).all()
        da = np.zeros((1 << bits) * 4, dtype=np.uint64); db = np.zeros_like(da)
        assert L.b200_gl_linearhash(_p(a), w, 1 << bits, _p(da)) == 0 and L.b200_gl_linearhash(_p(b), w, 1 << bits, _p(db)) == 0
        assert (da == db).all()


def test_msm_scalar_out_of_range_is_an_error(L):
    """a scalar whose signed-digit recoding overflows the windows (not a canonical Fr `Repr`) used to lose its top carry silently."""
    from eigen_zkvm_b200 import groth16 as g16
    from oracle import bn254 as bn
    pts = bn.pack_points([bn.G1, bn.mul(5, bn.G1)])
    ok = g16.multiexp(pts, bn.pack_scalars([3, bn.R - 1]))
    assert bn.unpack_point(g16.jacobian_to_affine_mont(ok)) == bn.add(bn.mul(3, bn.G1), bn.mul((bn.R - 1) * 5 % bn.R, bn.G1))
    bad = np.array([[3, 0, 0, 0], [2**64 - 1] * 4], dtype=np.uint64)
    out = np.zeros(12, dtype=np.uint64)
    assert L.b200_msm(0, _p(pts), _p(bad), 2, _p(out)) != 0 and b"scalar" in L.b200_last_error()
