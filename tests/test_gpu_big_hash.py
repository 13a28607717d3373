"""GPU parity for the BN128 / BLS12-381 Poseidon / LinearHash / 16-ary Merkle back-ends through the C-ABI: the
reference's KATs directly, and the oracle (oracle/poseidon_big.py, itself pinned to those KATs) on seeded inputs."""
import random
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P_GL = 0xFFFFFFFF00000001


@pytest.fixture(scope="module")
def mb():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import merklehash_big
    return merklehash_big


def test_reference_kats_on_device(mb):
    bn, bls = mb.Poseidon("BN128"), mb.Poseidon("BLS12381")
    # poseidon_bn128_opt.rs:232-300
    assert bn.hash([1]) == 0x29176100eaa962bdc1fe6c654d6a3c130e96a4d1168b33848b897dc502820133
    assert bn.hash([1, 2]) == 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a
    assert bn.hash([1, 2, 3, 4, 5, 6]) == 0x2d1a03850084442813c8ebf094dea47538490a68b05f2239134a4cca2f6302e1
    assert bn.hash(list(range(16))) == 0x1b733f2ff41971b23819a16bc8c16bbe13d98173358429fcc12f6f0826407a56
    # poseidon_bls12381_opt.rs:237-310
    assert bls.hash([1]) == 0x164efff6c8a32ef98836c868f8c8dedcbe3068d16ba6098f282a6d185edb551f
    assert bls.hash([1, 2, 3, 4]) == 0x6f5f297b0ab0d1e7400501b9bdd4c3be2fe676b6a05deb845143b87355167a8d
    assert bls.hash(list(range(16))) == 0x12d374bbdb8d3c1c0230b20b8fe1572f1e652a616d16e834718a982574106405
    with pytest.raises(ValueError):
        bn.hash([])
    with pytest.raises(ValueError):
        bn.hash(list(range(17)))
    # merklehash_bn128.rs:270-292 and merklehash_bls12381.rs:274-293
    t = mb.MerkleTree("BN128")
    t.merkelize(np.array([[i + j * 1000 for j in range(9)] for i in range(256)], dtype=np.uint64), 9, 256)
    assert t.root() == 2052732265221205192391066587135329070685482706470940527184785165917406935559
    t = mb.MerkleTree("BLS12381")
    t.merkelize(np.array([[i + j * 10 + 1 for j in range(3)] for i in range(4)], dtype=np.uint64), 3, 4)
    assert t.root() == 32227206116237215740162377531481191838063909532381497804787245624658969614932


@pytest.mark.parametrize("field", ["bn128", "bls12381"])
def test_poseidon_all_widths_against_oracle(mb, field):
    from oracle import poseidon_big as pb
    rnd = random.Random(7)
    h = mb.Poseidon(field)
    for n in range(1, 17):
        inp = [rnd.randrange(pb.MOD[field]) for _ in range(n)]
        init = rnd.randrange(pb.MOD[field]) if n % 2 else 0
        assert h.hash_ex(inp, init, n + 1) == pb.permute(field, inp, init)
        assert h.hash(inp, init) == pb.hash(field, inp, init)
    edge = [0, 1, pb.MOD[field] - 1]
    assert h.hash_ex(edge, pb.MOD[field] - 1, 4) == pb.permute(field, edge, pb.MOD[field] - 1)


@pytest.mark.parametrize("field", ["bn128", "bls12381"])
def test_linearhash_widths_against_oracle(mb, field):
    from oracle import poseidon_big as pb
    rng = np.random.default_rng(3)
    lh = mb.LinearHash(field)
    for width in [1, 2, 3, 4, 5, 6, 9, 12, 47, 48, 49, 50, 96, 100]:
        n = 5
        rows = rng.integers(0, P_GL, size=(n, width), dtype=np.uint64)
        rows[0, :] = P_GL - 1          # maximum limbs: width <= 4 exercises the raw-integer reduction (to_bn128_mont corner case)
        got = lh.hash_element_array(rows, width)
        assert got == [pb.hash_element_array(field, [int(v) for v in r]) for r in rows], width
    # the reference's corner-case KAT (linearhash_bn128.rs:155-175): Montgomery limbs of the digest
    d = lh.hash_element_array(np.array([[6188675464075253840, 2608530331018891925]], dtype=np.uint64), 2)[0]
    exp = {"bn128": [15714769047018385385, 14080511166848616671, 11411897157942048316, 1802287360671936077],
           "bls12381": [664572115127318441, 16413352647427919515, 17253685441004911215, 6212100569330953807]}[field]
    assert pb.to_ref_limbs(field, d) == exp


@pytest.mark.parametrize("field", ["bn128", "bls12381"])
@pytest.mark.parametrize("height,width", [(1, 5), (2, 3), (16, 6), (17, 6), (33, 6), (256, 9), (300, 50), (1000, 2)])
def test_merkle_nodes_against_oracle(mb, field, height, width):
    from oracle import poseidon_big as pb
    rng = np.random.default_rng(height * 131 + width)
    rows = rng.integers(0, P_GL, size=(height, width), dtype=np.uint64)
    t = mb.MerkleTree(field)
    t.merkelize(rows, width, height)
    exp = pb.merkelize(field, [[int(v) for v in r] for r in rows])
    assert len(t.nodes) == len(exp) == pb.get_n_nodes(height)
    assert t.nodes == exp
    idx = height - 1
    v, mp = t.get_group_proof(idx)
    assert v == [int(x) for x in rows[idx]]
    # recompute the root from the opening (merklehash_bn128.rs:108-129)
    cur = pb.hash_element_array(field, v); i = idx
    for sibs in mp:
        assert sibs[i & 15] == cur
        cur = pb.hash(field, sibs, 0); i >>= 4
    assert cur == t.root() or height == 1


@pytest.mark.parametrize("field", ["bn128", "bls12381"])
def test_large_tree_properties(mb, field):
    """2^17 x 12 (the final stark's shape: 2^16 rows, blowup 2, compressor12 width): device-resident call equals the host
    call, a sampled opening recomputes to the root, and changing one leaf element changes the root."""
    import ctypes, torch
    from oracle import poseidon_big as pb
    from eigen_zkvm_b200 import _lib
    h, w = 1 << 17, 12
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    cols = torch.randint(0, 2**62, (w, h), dtype=torch.int64, device="cuda", generator=g)       # column-major
    nn = _lib.lib().b200_big_merkle_n_nodes(h)
    d_nodes = torch.empty(nn * 4, dtype=torch.int64, device="cuda")
    fid = mb.FIELD_IDS[field]
    _lib.check(_lib.lib().b200_big_merkelize_dev(fid, ctypes.c_void_p(cols.data_ptr()), w, h, ctypes.c_void_p(d_nodes.data_ptr())))
    rows = cols.t().contiguous().cpu().numpy().view(np.uint64)
    t = mb.MerkleTree(field); t.merkelize(rows, w, h)
    dev_nodes = d_nodes.cpu().numpy().view(np.uint64).reshape(-1, 4)
    assert [sum(int(r[i]) << (64 * i) for i in range(4)) for r in dev_nodes[-3:]] == t.nodes[-3:]
    idx = 54321
    v, mp = t.get_group_proof(idx)
    cur = pb.hash_element_array(field, v); i = idx
    for sibs in mp:
        assert sibs[i & 15] == cur
        cur = pb.hash(field, sibs, 0); i >>= 4
    assert cur == t.root()
    rows2 = rows.copy(); rows2[idx, 3] ^= np.uint64(1)
    t2 = mb.MerkleTree(field); t2.merkelize(rows2, w, h)
    assert t2.root() != t.root()
