"""The library's host-side verifier (csrc/verify.cpp: stark_verify.rs:21-121 + fri.rs:187-297) on the committed Goldilocks proofs.
Runs WITHOUT a GPU (the Goldilocks transcript and Merkle hashing of a verifier are host code in the reference too); no oracle code
is involved: the acceptance of proofs the oracle generated and the oracle's verifier accepted is an independent cross-check of both."""
import json, os
import pytest
from eigen_zkvm_b200 import starkinfo as si, starky

FIXTURES = [("fib", "fib.pil.json.gl", "fib10.proof.json"), ("plookup", "plookup.pil.json.gl", "plookup10.proof.json"),
            ("pe", "pe.pil.json", "pe10.proof.json"), ("connection", "connection.pil.json", "connection10.proof.json")]


def _case(golden_dir, pil_name, proof_name):
    pil = si.load_pil(os.path.join(golden_dir, pil_name))
    ss = json.load(open(os.path.join(golden_dir, "starkStruct.json.gl")))
    info, prog = si.new_starkinfo(pil, ss)
    proof = open(os.path.join(golden_dir, proof_name)).read()
    return info, prog, ss, proof


@pytest.mark.parametrize("name,pil_name,proof_name", FIXTURES)
def test_library_verifier_accepts_the_golden_proofs(golden_dir, name, pil_name, proof_name):
    info, prog, ss, proof = _case(golden_dir, pil_name, proof_name)
    root = [int(x) for x in json.loads(proof)["rootC"]]
    why = []
    assert starky.stark_verify(proof, root, info, ss, prog, why), why
    # a different constant root: the constant-tree openings no longer match
    bad_root = list(root); bad_root[0] = (bad_root[0] + 1) % (2**64 - 2**32 + 1)
    why = []
    assert not starky.stark_verify(proof, bad_root, info, ss, prog, why)
    assert "tree C" in why[0]


def _bump(s):
    return str((int(s) + 1) % (2**64 - 2**32 + 1))


@pytest.mark.parametrize("name,pil_name,proof_name", FIXTURES)
def test_library_verifier_rejects_tampered_proofs(golden_dir, name, pil_name, proof_name):
    info, prog, ss, proof = _case(golden_dir, pil_name, proof_name)
    root = [int(x) for x in json.loads(proof)["rootC"]]
    nq = ss["nQueries"]

    def rejected(mut, expect=None):
        p = json.loads(proof); mut(p)
        why = []
        ok = starky.stark_verify(json.dumps(p), root, info, ss, prog, why)
        assert not ok
        if expect: assert expect in why[0], why
    def m_eval(p): p["evals"][0][0] = _bump(p["evals"][0][0])
    def m_root1(p): p["root1"][0] = _bump(p["root1"][0])
    def m_val(p): p["s0_vals1"][nq - 1][0] = _bump(p["s0_vals1"][nq - 1][0])
    def m_sib(p): p["s0_siblings4"][0][2][1] = _bump(p["s0_siblings4"][0][2][1])
    def m_const(p): p["s0_valsC"][1][0] = _bump(p["s0_valsC"][1][0])
    def m_fri_val(p): p["s1_vals"][0][3] = _bump(p["s1_vals"][0][3])
    def m_fri_sib(p): p["s1_siblings"][2][0][0] = _bump(p["s1_siblings"][2][0][0])
    def m_final(p): p["finalPol"][0][0] = _bump(p["finalPol"][0][0])
    def m_public(p):
        if p["publics"]: p["publics"][0] = _bump(p["publics"][0])
        else: p["evals"][1][2] = _bump(p["evals"][1][2])
    def m_drop_query(p): p["s0_vals1"].pop()
    def m_range(p): p["evals"][0][0] = str(2**64 - 2**32 + 1)
    def m_short(p): p["finalPol"].pop()
    for mut in (m_eval, m_root1, m_val, m_sib, m_const, m_fri_val, m_fri_sib, m_final, m_public, m_drop_query, m_range, m_short):
        rejected(mut)
    rejected(m_sib, "tree 4")
    # the folding check proper: the first FRI layer's opening is changed TOGETHER with nothing else, so its Merkle proof fails first;
    # swapping two whole queries keeps every Merkle proof valid for its own index but not for the index the transcript derives
    def m_swap(p):
        for k in list(p):
            if k.startswith("s0_") or (k.startswith("s") and k.endswith(("_vals", "_siblings"))):
                p[k][0], p[k][1] = p[k][1], p[k][0]
    rejected(m_swap)


def test_library_verifier_rejects_garbage_without_failing(golden_dir):
    info, prog, ss, proof = _case(golden_dir, *FIXTURES[0][1:])
    root = [int(x) for x in json.loads(proof)["rootC"]]
    for junk in ("", "{", "[]", "{}", proof[: len(proof) // 2], proof.replace('"evals"', '"evalz"')):
        why = []
        assert not starky.stark_verify(junk, root, info, ss, prog, why)
        assert why and why[0]


def test_library_verifier_agrees_with_the_oracle_verifier_on_a_fresh_oracle_proof(golden_dir):
    """a proof produced NOW by the CPU oracle for another size (2^12 rows, config 0) is accepted by the library's verifier"""
    from oracle import stark_oracle as so
    ss = {"nBits": 12, "nBitsExt": 13, "nQueries": 8, "verificationHashType": "GL", "steps": [{"nBits": 13}, {"nBits": 9}, {"nBits": 5}]}
    pil = si.load_pil(os.path.join(golden_dir, "fib.pil.json.gl"))
    for ref in pil["references"].values():
        ref["polDeg"] = 1 << 12
    pil["publics"][0]["idx"] = (1 << 12) - 1
    cm, const = so.fibonacci_inputs(12)
    setup = so.stark_setup(const, pil, ss)
    proof = so.stark_gen(cm, const, setup, ss)
    js = so.proof_to_json(proof)
    assert so.stark_verify(proof, setup["const_root"], setup["starkinfo"], ss, setup["program"])
    why = []
    assert starky.stark_verify(js, setup["const_root"], setup["starkinfo"], ss, setup["program"], why), why
    bad = json.loads(js); bad["s2_vals"][3][1] = _bump(bad["s2_vals"][3][1])
    assert not so.stark_verify(so.proof_from_json(json.dumps(bad)), setup["const_root"], setup["starkinfo"], ss, setup["program"])
    assert not starky.stark_verify(json.dumps(bad), setup["const_root"], setup["starkinfo"], ss, setup["program"])
