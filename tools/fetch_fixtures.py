#!/usr/bin/env python3
"""Copy the reference's GL *data* fixtures (no source code) into tests/golden/ so that parity tests
can run on the GPU box, where /root/reference does not exist.

Source: /root/reference/starky/data/{fib,plookup}.{pil.json,const,cm}.gl + starkStruct.json.gl --
(and the BN128-hash variants fib/plookup.{pil.json,cm,const} + starkStruct.json{,.bls12381} of stark_gen.rs:981-1022,1093-1148
and the const-root KAT of stark_setup.rs:83-98) -- the inputs of the reference's own end-to-end tests (starky/src/stark_gen.rs:1149-1195; pe.* / connection.* are the
permutation / connection fixtures of stark_gen.rs:1023-1148, there run with the BN128 hash, here with GL;
starky/src/stark_setup.rs:100-116).  .cm/.const are row-major little-endian canonical u64
(starky/src/polsarray.rs:137-217).
"""
import shutil, pathlib
src = pathlib.Path("/root/reference/starky/data")
dst = pathlib.Path(__file__).resolve().parent.parent / "tests" / "golden"
dst.mkdir(parents=True, exist_ok=True)
for n in ["fib.pil.json", "fib.cm", "fib.const", "plookup.pil.json", "plookup.cm", "plookup.const", "starkStruct.json", "starkStruct.json.bls12381", "pe.pil.json", "pe.const", "pe.cm", "connection.pil.json", "connection.const", "connection.cm", "fib.pil.json.gl", "fib.const.gl", "fib.cm.gl", "plookup.pil.json.gl", "plookup.const.gl", "plookup.cm.gl", "starkStruct.json.gl"]:
    shutil.copyfile(src / n, dst / n)
    (dst / n).chmod(0o644)
print("copied to", dst)

# groth16 boundary fixtures (groth16/src/json_utils.rs:350-429 round-trips them)
g = pathlib.Path("/root/reference/groth16/test-vectors")
for n in ["proof.bin", "proof.json", "verification_key.bin", "verification_key.json", "verification_key_bls12381.bin", "verification_key_bls12381.json"]:
    shutil.copyfile(g / n, dst / ("groth16_" + n)); (dst / ("groth16_" + n)).chmod(0o644)
print("copied groth16 test vectors")
