"""GPU parity tests proper: every call goes through the C-ABI (ctypes) and is compared bit-exactly with the
CPU oracle on the same seeded inputs; golden vectors of the reference are re-asserted on the device path."""
import json, os, random
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def sk():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import starky
    return starky


def _rand(shape, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2**63, size=shape, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=shape, dtype=np.uint64)
    return a % np.uint64(P)


def test_poseidon_kats(sk):
    # starky/src/poseidon_opt.rs:219-262
    h = sk.Poseidon()
    assert h.hash([0] * 8, [0] * 4) == [0x3c18a9786cb0b359, 0xc4055e3364a246c3, 0x7953db0ab48808f4, 0xc71603f33a1144ca]
    assert h.hash(list(range(8)), list(range(8, 12))) == [0xd64e1e3efc5b8e9e, 0x53666633020aaa47, 0xd40285597c6a8825, 0x613a4f81e81231d2]
    assert h.hash([P - 1] * 8, [P - 1] * 4) == [0xbe0085cfc57a8357, 0xd95af71847d05c09, 0xcf55a13d33c1c953, 0x95803a74f4530e82]
    from oracle import gl
    rnd = random.Random(11)
    for _ in range(20):
        i8 = [rnd.randrange(P) for _ in range(8)]; c4 = [rnd.randrange(P) for _ in range(4)]
        assert h.hash(i8, c4, 12) == gl.poseidon(i8, c4)
    with pytest.raises(ValueError):
        h.hash([0] * 7, [0] * 4)


def test_linearhash_kats_and_widths(sk):
    from oracle import gl
    lh = sk.LinearHash()
    # starky/src/linearhash.rs:311-362
    assert lh.hash(list(range(1, 28))) == [17618903473682537397, 11844743283521766961, 185773432536380223, 6083210164459944430]
    assert lh.hash([1, 2, 3]) == [1, 2, 3, 0]
    for w in [1, 2, 4, 5, 8, 9, 12, 16, 17, 31, 32, 33, 36, 40, 48, 50, 64, 65, 100]:
        rows = _rand((37, w), 100 + w)
        d = lh.hash_rows(rows, w)
        for r in (0, 17, 36):
            assert [int(x) for x in d[r]] == gl.linearhash(rows[r]), "width %d row %d" % (w, r)


def _cols(n, n_pols):
    return np.array([[i + j * 1000 for j in range(n_pols)] for i in range(n)], dtype=np.uint64)


def test_merkle_kats(sk):
    from oracle import gl
    # starky/src/merklehash.rs:469-497
    t = sk.MerkleTreeGL(); t.merkelize(_cols(256, 9), 9, 256)
    assert t.root() == [11508832812350783315, 5044133147279090978, 6335412741057168694, 12530816673814004438]
    v, mp = t.get_group_proof(3)
    assert gl.merkle_root_from_proof(v, np.array(mp, dtype=np.uint64), 3) == t.root()
    # starky/src/merklehash.rs:519-545 (non power of two)
    t = sk.MerkleTreeGL(); t.merkelize(_cols(33, 6), 6, 33)
    assert t.root() == [10952823080416094333, 14127307315435918656, 18155557507084305090, 4650815682547343351]
    assert (t.nodes == gl.merkelize(_cols(33, 6), 6, 33)).all()
    # starky/src/merklehash.rs:548-564 (2^16 x 50): full node array equals the oracle's
    a = _cols(1 << 16, 50)
    t = sk.MerkleTreeGL(); t.merkelize(a, 50, 1 << 16)
    assert (t.nodes == gl.merkelize(a, 50, 1 << 16)).all()
    with pytest.raises(IndexError):
        t.get_group_proof(1 << 16)
    # width 0 (empty section): tree of zero digests, merklehash.rs:311-343
    t = sk.MerkleTreeGL(); t.merkelize(np.zeros(0, dtype=np.uint64), 0, 64)
    assert (t.nodes == gl.merkelize(np.zeros(0, dtype=np.uint64), 0, 64)).all()
    # height 1: get_n_nodes(1) = 2 and root() = nodes[len-1] is the (zero) pad slot -- reference quirk, kept
    t = sk.MerkleTreeGL(); t.merkelize(np.array([5, 6, 7], dtype=np.uint64), 3, 1)
    assert (t.nodes == gl.merkelize(np.array([5, 6, 7], dtype=np.uint64), 3, 1)).all()
    assert [int(x) for x in t.nodes[0]] == [5, 6, 7, 0]


@pytest.mark.parametrize("bits,w", [(1, 1), (3, 2), (5, 3), (9, 2), (10, 1), (11, 5), (12, 2), (13, 3), (14, 1), (15, 4), (16, 2), (17, 1), (18, 5), (19, 2), (20, 3), (21, 2), (22, 1)])
def test_ntt_intt_match_oracle(sk, bits, w):
    # starky/src/fft_p.rs:372-477 sizes (2^5, 2^18 x 5, ...) against the simple single-vector FFT
    from oracle import gl
    a = _rand((1 << bits, w), bits * 31 + w)
    f = sk.fft(a, w, bits)
    assert (f == gl.ntt(a, w, bits).reshape(-1)).all()
    assert (sk.ifft(f, w, bits) == a.reshape(-1)).all()
    assert (sk.ifft(a, w, bits) == gl.intt(a, w, bits).reshape(-1)).all()


@pytest.mark.parametrize("bits,ext,w", [(2, 3, 1), (5, 6, 3), (8, 11, 2), (10, 11, 1), (12, 13, 2), (12, 15, 2), (13, 14, 3), (15, 19, 1), (16, 17, 3), (18, 19, 5), (17, 20, 2), (19, 21, 1), (20, 23, 2)])
def test_interpolate_matches_oracle(sk, bits, ext, w):
    from oracle import gl
    a = _rand((1 << bits, w), bits * 17 + ext)
    assert (sk.interpolate(a, w, bits, ext) == gl.lde(a, w, bits, ext)).all()
    assert sk.interpolate(np.zeros(0, dtype=np.uint64), 0, bits, ext).size == 0


def test_const_tree_root_kat(sk, golden_dir):
    # starky/src/stark_setup.rs:100-116
    const = np.fromfile(os.path.join(golden_dir, "fib.const.gl"), dtype="<u8")
    ext = sk.interpolate(const, 1, 10, 11)
    t = sk.MerkleTreeGL(); t.merkelize(ext, 1, 2048)
    assert t.root() == [15302509084042343527, 985081440042889555, 14692153289195851822, 1611894784155222896]


def _gl_add_np(x, y):
    t = x + y
    t = np.where(t < x, t + np.uint64(0xFFFFFFFF), t)
    return np.where(t >= np.uint64(P), t - np.uint64(P), t)


def test_ntt_large_roundtrip_and_linearity(sk):
    # size-independent properties at BASELINE's full height (2^24 x 2), where the oracle would take minutes:
    # iNTT(NTT(a)) = a and NTT(a + b) = NTT(a) + NTT(b)
    bits, w = 24, 2
    a = _rand((1 << bits, w), 7).reshape(-1); b = _rand((1 << bits, w), 8).reshape(-1)
    fa = sk.fft(a, w, bits)
    assert (sk.ifft(fa, w, bits) == a).all()
    fb = sk.fft(b, w, bits)
    assert (sk.fft(_gl_add_np(a, b), w, bits) == _gl_add_np(fa, fb)).all()


def test_sharded_lde_merkle_world1_matches_oracle(sk):
    # the multi-GPU entry point with one rank (GpuBackend over the C-ABI device calls) against the oracle
    import torch
    from eigen_zkvm_b200 import sharded
    from oracle import gl
    W, nbits, nbits_ext = 12, 8, 11
    full = _rand((W, 1 << nbits), 77)                      # column-major
    cols = torch.from_numpy(full.reshape(-1).view(np.int64)).cuda()
    root, nodes, shard = sharded.lde_merkle_sharded(cols, W, nbits, nbits_ext, sharded.GpuBackend())
    ext = gl.lde(np.ascontiguousarray(full.T), W, nbits, nbits_ext)
    assert root == [int(x) for x in gl.merkelize(ext, W, 1 << nbits_ext)[-1]]
    assert (shard.cpu().numpy().view(np.uint64).reshape(W, -1) == ext.reshape(-1, W).T).all()
