"""`Groth16::prove` in one call on the GPU (b200_groth16_pk_read + b200_groth16_prove) against oracle/groth16_oracle.py.
Parity is UNPINNED by the reference (bellman_ce is un-vendored; its tests only check prove -> verify, groth16/src/groth16.rs:134-262).
Pinned here: the setup is generated with a known trapdoor, so (A, B, C) have closed forms in the exponent and the GPU's points must
equal them EXACTLY; the pairing equation is then checked in the exponent.  Circuits: the reference's own test/multiplier.r1cs
(BN128) and groth16/test-vectors/mycircuit_bls12381.r1cs (BLS12-381), and a synthetic 2^10-constraint product chain."""
import os, random
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "groth16")
TRAPDOOR = (0x1234567890ABCDEF1234567890ABCDEF, 0x2222222222222222333333333333, 0x9999999999AAAAAAAAAAAAAABBBB, 0x1111111100000000FFFFFFFF, 0x7777777755555555333333331111)


@pytest.fixture(scope="module")
def g16():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import groth16
    return groth16


def _check(g16, r1cs, curve, witness, seed):
    from oracle import groth16_oracle as G, fr_domain as D
    cv = G.CURVE[curve]; p = D.MOD[cv["field"]]
    cid = g16.BN128 if curve == "BN128" else g16.BLS12381
    S, P = G.setup(r1cs, curve, TRAPDOOR)
    params = g16.Parameters(G.write_parameters(P, curve), cid)
    assert (params.n_h, params.n_l, params.n_ic) == (S["m"] - 1, r1cs["num_aux"], r1cs["num_inputs"])
    syn = G.synthesize(r1cs, witness, p)
    rnd = random.Random(seed)
    for r, s in ((rnd.randrange(p), rnd.randrange(p)), (0, 0), (1, p - 1)):
        A, B, Cc = g16.prove(params, syn["a"], syn["b"], syn["c"], syn["inputs"], syn["aux"], syn["a_aux_density"], syn["b_input_density"], syn["b_aux_density"], r, s)
        A_s, B_s, C_s, ic_s = G.prove_in_the_exponent(r1cs, curve, S, witness, r, s)
        g1, g2 = cv["g1"], cv["g2"]
        assert g1.affine_from_words(A) == g1.mul(A_s, g1.gen)
        assert g2.affine_from_words(B) == g2.mul(B_s, g2.gen)
        assert g1.affine_from_words(Cc) == g1.mul(C_s, g1.gen)
        assert G.pairing_equation_holds(S, p, A_s, B_s, C_s, ic_s)          # the verifier's equation, in the exponent
    params.free()


def test_reference_multiplier_circuit_bn128(g16):
    from oracle import groth16_oracle as G, fr_domain as D
    r1cs = G.read_r1cs(open(os.path.join(HERE, "multiplier.r1cs"), "rb").read())
    p = D.MOD["bn254"]
    w = g16.read_wtns(G.write_wtns([1, 33, 3, 11], p), g16.BN128)          # test/multiplier.input.json: a = 3, b = 11
    _check(g16, r1cs, "BN128", w, 1)


def test_reference_mycircuit_bls12381(g16):
    from oracle import groth16_oracle as G, fr_domain as D
    r1cs = G.read_r1cs(open(os.path.join(HERE, "mycircuit_bls12381.r1cs"), "rb").read())
    p = D.MOD["bls12381"]
    _check(g16, r1cs, "BLS12381", [1, 7 * 9, 7, 9], 2)


@pytest.mark.parametrize("curve", ["BN128", "BLS12381"])
def test_product_chain_150_constraints(g16, curve):
    """x_{k+1} = x_k * x_k + x_0 (one constraint each): 150 constraints (+ 2 input rows), 1 public output, domain 2^8: the table MSMs and the device H
    composed (the scalar-field transforms themselves are checked up to 2^12 / 2^20 in tests/test_gpu_fr_domain.py); the python
    big-int setup of a longer chain would cost minutes of GPU-box time."""
    from oracle import groth16_oracle as G, fr_domain as D
    p = D.MOD[G.CURVE[curve]["field"]]
    n = 150
    # wires: 0 = ONE, 1 = out (public), 2 = x_0 (private input), 3.. = x_1 .. x_n ; out = x_n
    cons = []
    x = [5]
    for k in range(n):
        nxt = (x[-1] * x[-1] + x[0]) % p
        # x_k * x_k = x_{k+1} - x_0
        cons.append(([(2 + k, 1)], [(2 + k, 1)], [(3 + k, 1), (2, p - 1)]))
        x.append(nxt)
    cons.append(([(2 + n, 1)], [(0, 1)], [(1, 1)]))              # out = x_n
    r1cs = dict(prime=p, n_wires=3 + n, n_pub_out=1, n_pub_in=0, n_prv_in=1, num_inputs=2, num_aux=1 + n, constraints=cons)
    w = [1, x[-1]] + x
    _check(g16, r1cs, curve, w, 3)


def test_errors(g16):
    from eigen_zkvm_b200 import _lib
    from oracle import groth16_oracle as G, fr_domain as D
    r1cs = G.read_r1cs(open(os.path.join(HERE, "multiplier.r1cs"), "rb").read())
    p = D.MOD["bn254"]
    S, P = G.setup(r1cs, "BN128", TRAPDOOR)
    blob = G.write_parameters(P, "BN128")
    # truncated, trailing byte, a coordinate that is not below the field modulus, a compressed-point flag
    for bad in (blob[:-1], blob + b"\0", bytes([0x3F]) + b"\xff" * 31 + blob[32:], bytes([blob[0] | 0x80]) + blob[1:]):
        with pytest.raises(_lib.B200Error):
            g16.Parameters(bad, g16.BN128)
    params = g16.Parameters(blob, g16.BN128)
    syn = G.synthesize(r1cs, [1, 33, 3, 11], p)
    with pytest.raises(_lib.B200Error):       # densities that do not select as many scalars as the key has points
        g16.prove(params, syn["a"], syn["b"], syn["c"], syn["inputs"], syn["aux"], [True, True], syn["b_input_density"], syn["b_aux_density"], 1, 2)
    with pytest.raises(_lib.B200Error):       # a domain the key was not made for
        g16.prove(params, syn["a"] * 3, syn["b"] * 3, syn["c"] * 3, syn["inputs"], syn["aux"], syn["a_aux_density"], syn["b_input_density"], syn["b_aux_density"], 1, 2)
