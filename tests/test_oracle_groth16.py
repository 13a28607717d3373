"""The groth16 prove() oracle (oracle/groth16_oracle.py) pinned on the CPU: the reference's own .r1cs fixtures parse; the trapdoor
setup is consistent (every parameter point is the claimed multiple, H's coefficients satisfy A B - C = H Z, the closed-form proof
satisfies the pairing equation in the exponent and a wrong witness does not); the `.wtns` reader of the product library
(b200_wtns_read, host code) reads what algebraic/src/reader.rs:87-138 reads and rejects what it rejects."""
import ctypes, os, random, struct
import numpy as np
import pytest
from oracle import groth16_oracle as G, fr_domain as D, curves as C

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "groth16")
TRAPDOOR = (0x1234567890ABCDEF1234567890ABCDEF, 0x2222222222222222333333333333, 0x9999999999AAAAAAAAAAAAAABBBB, 0x1111111100000000FFFFFFFF, 0x7777777755555555333333331111)


def multiplier_witness(p, a=3, b=11):
    # test/multiplier.input.json {"a": 3, "b": 11}; wires: ONE, c (public output), a, b
    return [1, a * b % p, a, b]


@pytest.mark.parametrize("fn,curve", [("multiplier.r1cs", "BN128"), ("mycircuit_bls12381.r1cs", "BLS12381")])
def test_setup_and_closed_form_proof(fn, curve):
    r1cs = G.read_r1cs(open(os.path.join(HERE, fn), "rb").read())
    cv = G.CURVE[curve]; p = D.MOD[cv["field"]]
    assert r1cs["prime"] == p and r1cs["num_inputs"] == 2 and r1cs["num_aux"] == 2 and len(r1cs["constraints"]) == 1
    w = multiplier_witness(p)
    syn = G.synthesize(r1cs, w, p)
    assert len(syn["a"]) == 3 and syn["a"][0] * syn["b"][0] % p == syn["c"][0]          # the constraint holds on this witness
    assert syn["a_aux_density"] == [True, False] and syn["b_aux_density"] == [False, True] and syn["b_input_density"] == [False, False]
    S, P = G.setup(r1cs, curve, TRAPDOOR)
    assert S["m"] == 4 and len(P["h"]) == 3 and len(P["l"]) == 2 and len(P["ic"]) == 2
    # bellman keeps the non-zero a / b points: inputs (consistency rows) + dense aux
    assert len(P["a"]) == 2 + sum(syn["a_aux_density"]) and len(P["b_g1"]) == sum(syn["b_input_density"]) + sum(syn["b_aux_density"]) == len(P["b_g2"])
    for q in [P["alpha_g1"], P["beta_g1"], P["delta_g1"]] + P["ic"] + P["h"] + P["l"] + P["a"] + P["b_g1"]:
        assert cv["g1"].is_on_curve(q)
    for q in [P["beta_g2"], P["gamma_g2"], P["delta_g2"]] + P["b_g2"]:
        assert cv["g2"].is_on_curve(q)
    # H by the FFT route equals the definition at tau: h(tau) Z(tau) = a(tau) b(tau) - c(tau)
    h, m = G.h_coefficients(r1cs, curve, w)
    assert m == 4 and len(h) == 3
    tau = TRAPDOOR[0] % p
    a_t = sum(w[i] * S["At"][i] for i in range(4)) % p; b_t = sum(w[i] * S["Bt"][i] for i in range(4)) % p; c_t = sum(w[i] * S["Ct"][i] for i in range(4)) % p
    assert sum(h[i] * pow(tau, i, p) for i in range(3)) % p * S["z"] % p == (a_t * b_t - c_t) % p
    rnd = random.Random(4)
    r, s = rnd.randrange(p), rnd.randrange(p)
    A_s, B_s, C_s, ic_s = G.prove_in_the_exponent(r1cs, curve, S, w, r, s)
    assert G.pairing_equation_holds(S, p, A_s, B_s, C_s, ic_s)
    # composing the parameter POINTS the way bellman's create_proof does gives the same elements
    g1 = cv["g1"]
    msm = lambda pts, sc: g1.msm_naive(pts, sc)
    a_sc = syn["inputs"] + [x for x, d in zip(syn["aux"], syn["a_aux_density"]) if d]
    g_a = g1.add(g1.add(g1.mul(r, P["delta_g1"]), P["alpha_g1"]), msm(P["a"], a_sc))
    assert g_a == g1.mul(A_s, g1.gen)
    # a witness that violates the constraint cannot satisfy the equation (h(tau) is then not a polynomial value)
    bad = list(w); bad[1] = (bad[1] + 1) % p
    A2, B2, C2, ic2 = G.prove_in_the_exponent(r1cs, curve, S, bad, r, s)
    hb, _ = G.h_coefficients(r1cs, curve, bad)
    a_t = sum(bad[i] * S["At"][i] for i in range(4)) % p; b_t = sum(bad[i] * S["Bt"][i] for i in range(4)) % p; c_t = sum(bad[i] * S["Ct"][i] for i in range(4)) % p
    assert sum(hb[i] * pow(tau, i, p) for i in range(3)) % p * S["z"] % p != (a_t * b_t - c_t) % p
    # Parameters::write layout: the vk prefix has the size of the reference's verification_key fixtures (708 / 1060 B with 2 IC)
    blob = G.write_parameters(P, curve)
    n = cv["nbytes"]
    vk_len = 3 * 2 * n + 3 * 4 * n + 4 + 2 * 2 * n
    assert vk_len == {"BN128": 708, "BLS12381": 1060}[curve]
    assert struct.unpack(">I", blob[vk_len:vk_len + 4])[0] == 3


def test_wtns_reader_of_the_library():
    import __graft_entry__ as g
    g.build()
    from eigen_zkvm_b200 import groth16 as g16, _lib
    for curve, cid in (("BN128", g16.BN128), ("BLS12381", g16.BLS12381)):
        p = D.MOD[G.CURVE[curve]["field"]]
        w = [1, 33, 3, 11, p - 1, 0]
        data = G.write_wtns(w, p)
        assert G.read_wtns(data) == (p, w)
        assert g16.read_wtns(data, cid) == w
        for bad in (b"wtnx" + data[4:], data[:4] + struct.pack("<I", 3) + data[8:], data[:-1], data[:28] + bytes(32) + data[60:]):
            with pytest.raises(_lib.B200Error):
                g16.read_wtns(bad, cid)
    with pytest.raises(_lib.B200Error):
        g16.read_wtns(G.write_wtns([1], D.MOD["bn254"]), g16.BLS12381)            # prime of the other curve
