"""Pins oracle/poseidon_big.py (BN128 / BLS12-381 Poseidon, LinearHash, 16-ary Merkle) to the reference's own KATs."""
import pytest
from oracle import poseidon_big as pb


def test_poseidon_bn128_kats():
    # starky/src/poseidon_bn128_opt.rs:232-300
    h = lambda v: pb.hash("bn128", v, 0)
    assert h([1]) == 0x29176100eaa962bdc1fe6c654d6a3c130e96a4d1168b33848b897dc502820133
    assert h([1, 2]) == 0x115cc0f5e7d690413df64c6b9662e9cf2a3617f2743245519e19607a4417189a
    assert h([1, 2, 0, 0, 0]) == 0x024058dd1e168f34bac462b6fffe58fd69982807e9884c1c6148182319cee427
    assert h([1, 2, 0, 0, 0, 0]) == 0x21e82f465e00a15965e97a44fe3c30f3bf5279d8bf37d4e65765b6c2550f42a1
    assert h([3, 4, 0, 0, 0]) == 0x0cd93f1bab9e8c9166ef00f2a1b0e1d66d6a4145e596abe0526247747cc71214
    assert h([3, 4, 0, 0, 0, 0]) == 0x1b1caddfc5ea47e09bb445a7447eb9694b8d1b75a97fff58e884398c6b22825a
    assert h([1, 2, 3, 4, 5, 6]) == 0x2d1a03850084442813c8ebf094dea47538490a68b05f2239134a4cca2f6302e1
    assert h(list(range(16))) == 0x1b733f2ff41971b23819a16bc8c16bbe13d98173358429fcc12f6f0826407a56
    with pytest.raises(ValueError):
        h([])
    with pytest.raises(ValueError):
        h(list(range(17)))


def test_poseidon_bls12381_kats():
    # starky/src/poseidon_bls12381_opt.rs:237-310 (output lane 1, "Neptune convention", :95-103)
    h = lambda v: pb.hash("bls12381", v, 0)
    assert h([1]) == 0x164efff6c8a32ef98836c868f8c8dedcbe3068d16ba6098f282a6d185edb551f
    assert h([1, 0]) == 0x59220c0fc5748e83c141c7bb8dae0a2bd5bbb227c778ede87296ba07960ec3d8
    assert h([1, 0, 0]) == 0x73584296b068384db6028b55d995108518d4483ab177197274effe979b91526e
    assert h([1, 2, 0, 0, 0]) == 0x385acd94e53a8c6f981809c2201582beceaec12250200f1e75ba93e6cf5ec736
    assert h([1, 2, 0, 0, 0, 0]) == 0x023dd8aecc0967c0588754eebd39af39bdae2bbf4195fee1208613c909aaa29b
    assert h([3, 4, 0, 0, 0]) == 0x19c96d726da9e3df4e5d0da19f324f7bf376dc7bf97efbf37082473f7fa24af8
    assert h([3, 4, 0, 0, 0, 0]) == 0x0cb7b1761b9abe661847a10701c6eae7c631ff580c5b7f3ac2f8be1088d22bba
    assert h([1, 2, 3, 4]) == 0x6f5f297b0ab0d1e7400501b9bdd4c3be2fe676b6a05deb845143b87355167a8d
    assert h(list(range(16))) == 0x12d374bbdb8d3c1c0230b20b8fe1572f1e652a616d16e834718a982574106405


def test_linearhash_kats():
    # linearhash_bn128.rs:140-153, linearhash_bls12381.rs:139-168 (hash_element_matrix)
    m100 = [[e, e * 1000, e * 1000000] for e in range(100)]
    assert pb.hash_element_matrix("bn128", m100) == 0x29c2ac38b7b8d18b9c1b575369cb4ab930ef71ebd5e4631b3916360233a29cae
    assert pb.hash_element_matrix("bls12381", m100) == 0x1aea10165e8c452045633835341291832bf7d46ace4bd6e8b1a2ddb9f257c2be
    assert pb.hash_element_matrix("bls12381", [[e, e, e] for e in range(9)]) == 0x683f0b0c6f1a15d7715cbac061ca80f1f30a28920d32993c2f9cd307aee7bcbb
    # corner case, width <= 4: the reference asserts the 4 Montgomery limbs of the digest (linearhash_bn128.rs:155-175,
    # linearhash_bls12381.rs:170-192)
    for field, rows in (("bn128", [([6188675464075253840, 2608530331018891925], [15714769047018385385, 14080511166848616671, 11411897157942048316, 1802287360671936077]),
                                   ([18440682777423237490, 1156220815552880681], [12850950522295690944, 15045028186447136619, 11701297961637547631, 875058675367281598])]),
                        ("bls12381", [([6188675464075253840, 2608530331018891925], [664572115127318441, 16413352647427919515, 17253685441004911215, 6212100569330953807]),
                                      ([18440682777423237490, 1156220815552880681], [13796980492452026086, 13318555390970742201, 9516443056151387241, 7411250834153264701])])):
        for vals, limbs in rows:
            assert pb.to_ref_limbs(field, pb.hash_element_array(field, vals)) == limbs
    # the "R2" constants of to_bn128_mont are R^2 mod r (linearhash_bn128.rs:80-85, linearhash_bls12381.rs:79-84)
    lim = lambda v: [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    assert lim(pow(2, 512, pb.MOD["bn128"])) == [1997599621687373223, 6052339484930628067, 10108755138030829701, 150537098327114917]
    assert lim(pow(2, 512, pb.MOD["bls12381"])) == [14526898881837571181, 3129137299524312099, 419701826671360399, 524908885293268753]


def test_merkle_kats():
    # merklehash_bn128.rs:270-292: 256 x 9, value i + 1000 j
    rows = [[i + j * 1000 for j in range(9)] for i in range(256)]
    nodes = pb.merkelize("bn128", rows)
    assert len(nodes) == pb.get_n_nodes(256) == 256 + 16 + 1
    assert nodes[-1] == 2052732265221205192391066587135329070685482706470940527184785165917406935559
    # merklehash_bls12381.rs:274-293: 4 x 3, value i + 10 j + 1
    rows = [[i + j * 10 + 1 for j in range(3)] for i in range(4)]
    nodes = pb.merkelize("bls12381", rows)
    assert len(nodes) == 16 + 1
    assert nodes[-1] == 32227206116237215740162377531481191838063909532381497804787245624658969614932
    # non power of 16 heights pad levels with zero digests (merklehash_bn128.rs:26-40)
    assert pb.get_n_nodes(33) == 48 + 16 + 1 and pb.get_n_nodes(1) == 16 and pb.get_n_nodes(17) == 32 + 16 + 1
