"""B200-native prover core for eigen-zkvm's two data-parallel hot paths: the STARK pipeline of `starky::stark_gen`
(Goldilocks NTT / LDE, step programs, Poseidon Merkle trees over GL / BN128 / BLS12-381, FRI) and the groth16 final layer
(G1 / G2 multi-scalar multiplication and the scalar-field domain on BN254 / BLS12-381).

Layout: csrc/ holds the CUDA kernels and the C-ABI (libb200zk.so, include/b200zk.h); the modules here are the thin
host-side mirrors of the reference's interfaces over that ABI:
  starky          StarkSetup / StarkProof.stark_gen / MerkleTreeGL / LinearHash / Poseidon / fft, ifft, interpolate
  starkinfo       port of the reference's PIL codegen (StarkInfo::new), host side of the stark_gen boundary
  merklehash_big  Poseidon / LinearHash / MerkleTree for the BN128 and BLS12381 back-ends
  groth16         multiexp (4 groups), fr_fft, groth16_h
  groth16_formats bellman's VerifyingKey / Proof binary layouts, JSON twins, base packing
  sharded         multi-GPU sharding over torch.distributed (NCCL)
There is no CPU fallback: without the built library or without a CUDA device the compute calls fail loudly.
"""
