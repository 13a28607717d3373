"""B200-native prover core for eigen-zkvm's Goldilocks STARK hot path (and the groth16 MSM).

Layout: csrc/ holds the CUDA kernels and the C-ABI (libb200zk.so); this package is the thin host-side
mirror of the reference's starky interface (StarkSetup / stark_gen / MerkleTree / fft) over that ABI.
"""
