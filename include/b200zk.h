/*
 * b200zk.h -- C-ABI of the B200-native prover core (libb200zk.so).
 *
 * Drop-in boundary for eigen-zkvm's Goldilocks STARK hot path.  Every entry point names the reference
 * interface (file:line under 0xEigenLabs/eigen-zkvm) it replaces; INTEGRATION.md shows the Rust
 * `extern "C"` binding a maintainer would add.  Conventions:
 *   - plain pointers and sizes only; all field elements are CANONICAL u64 in [0, p), p = 2^64 - 2^32 + 1
 *     (the reference's Montgomery form is internal: convert with `as_int()` / `Fr::from`,
 *     fields/src/field_gl.rs:496-507,542-544);
 *   - matrices are row-major [row][col], the layout of `PolsArray::write_buff`
 *     (starky/src/polsarray.rs:219-227) and of the `.cm/.const` files (:137-217);
 *   - return 0 on success, negative on error; the message is in b200_last_error() (thread local);
 *     no exception or panic crosses the boundary;
 *   - host-pointer entry points copy in/out themselves; `_dev` variants take device pointers
 *     (inputs already resident in HBM) and run on the library stream;
 *   - there is no CPU fallback: without a CUDA device every compute entry point returns B200_ERR_CUDA.
 */
#ifndef B200ZK_H
#define B200ZK_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_ARG (-1)
#define B200_ERR_CUDA (-2)
#define B200_ERR_UNSUPPORTED (-3)
#define B200_ERR_INTERNAL (-4)

/* ---- library ---------------------------------------------------------------------------------------- */
const char* b200_last_error(void);
const char* b200_version(void);
void b200_free(void* p);                       /* frees buffers returned through char** / void** outputs */
int b200_device_count(void);
int b200_set_device(int device);               /* one process per GPU: call once before anything else */
int b200_set_stream(void* cuda_stream);        /* cudaStream_t; default: the legacy default stream */
/* per-kernel device timings (CUDA events on the launch stream), JSON: [{"name","launches","ms","bytes"}] */
int b200_timing_enable(int on);
int b200_timing_report(char** json_out, size_t* len_out);
uint64_t b200_kernel_launches(void);           /* number of kernels this library launched so far */

/* ---- Goldilocks NTT family: starky/src/fft_p.rs:242-261 (`fft`, `ifft`, `interpolate`);
 *      call sites stark_gen.rs:377,395,724 and stark_setup.rs:43.  n_cols interleaved columns. ----------- */
int b200_gl_ntt(const uint64_t* in, uint64_t* out, size_t n_cols, unsigned log_n);
int b200_gl_intt(const uint64_t* in, uint64_t* out, size_t n_cols, unsigned log_n);
int b200_gl_lde(const uint64_t* in, uint64_t* out, size_t n_cols, unsigned log_n, unsigned log_n_ext);
/* device-resident, COLUMN-MAJOR ([col][row]) variants used for kernel benchmarking */
int b200_gl_ntt_dev(const uint64_t* d_in, uint64_t* d_out, size_t n_cols, unsigned log_n, int inverse);
int b200_gl_lde_dev(const uint64_t* d_in, uint64_t* d_out, size_t n_cols, unsigned log_n, unsigned log_n_ext);

/* ---- Poseidon / LinearHash / MerkleTreeGL: starky/src/poseidon_opt.rs:76-78 (`Poseidon::hash`),
 *      linearhash.rs:79-110 (`LinearHash::hash`), traits.rs:24-55 + merklehash.rs:293-346,430-458 --------- */
int b200_gl_poseidon(const uint64_t in8[8], const uint64_t cap4[4], uint64_t out12[12]);
int b200_gl_linearhash(const uint64_t* rows, size_t width, size_t n_rows, uint64_t* digests_out /* n_rows x 4 */);
size_t b200_gl_merkle_n_nodes(size_t height);                                  /* merklehash.rs:47-61 */
int b200_gl_merkelize(const uint64_t* leaves, size_t width, size_t height, uint64_t* nodes_out /* n_nodes x 4 */);
int b200_gl_merkelize_dev(const uint64_t* d_leaves_colmajor, size_t width, size_t height, uint64_t* d_nodes_out);

/* ---- BN128 / BLS12-381 Poseidon commitment back-ends (`verificationHashType` "BN128" / "BLS12381", the last stark
 *      before the snark; starky/src/prove.rs:52-89): `Poseidon::hash_ex` / `hash` (poseidon_bn128_opt.rs:94-118,
 *      poseidon_bls12381_opt.rs:95-113), `LinearHash*::hash_element_array` (linearhash_bn128.rs:105-131),
 *      `MerkleTree*::merkelize` (merklehash_bn128.rs:26-40,176-224; BLS12-381 twins alike).  A field element /
 *      digest is 4 x u64 little-endian CANONICAL (the reference's `ElementDigest<4, Fr>` holds the Montgomery limbs
 *      of the same scalar: convert with `into_repr()` / `from_repr`, digest.rs:45-65).  `nodes` keeps the reference
 *      layout: levels concatenated, each padded with zero digests to a multiple of 16, root last. --------------- */
#define B200_HASH_BN128 0
#define B200_HASH_BLS12381 1
/* full permutation state for inputs (1..16 elements) and init_state: state_out = (n_inputs + 1) x 4 (`hash_ex(.., t)`) */
int b200_big_poseidon(int field, const uint64_t* inputs, size_t n_inputs, const uint64_t init4[4], uint64_t* state_out);
/* `Poseidon::hash`: lane 0 of the state for BN128, lane 1 for BLS12-381 */
int b200_big_hash(int field, const uint64_t* inputs, size_t n_inputs, const uint64_t init4[4], uint64_t out4[4]);
int b200_big_linearhash(int field, const uint64_t* rows, size_t width, size_t n_rows, uint64_t* digests_out /* n_rows x 4 */);
size_t b200_big_merkle_n_nodes(size_t height);
int b200_big_merkelize(int field, const uint64_t* leaves, size_t width, size_t height, uint64_t* nodes_out /* n_nodes x 4 */);
int b200_big_merkelize_dev(int field, const uint64_t* d_leaves_colmajor, size_t width, size_t height, uint64_t* d_nodes_out);

/* ---- STARK: starky/src/stark_setup.rs:27-66 (`StarkSetup::new`) and stark_gen.rs:193-202
 *      (`StarkProof::<MerkleTreeGL>::stark_gen::<TranscriptGL>`); proof = serde_json of StarkProof
 *      (serializer.rs:137-270), byte-identical to `serde_json::to_string(&starkproof)` (prove.rs:153). ---- */
typedef struct b200_setup b200_setup_t;
/* setup_json = {"starkinfo": <serde StarkInfo>, "program": <serde Program>, "stark_struct": <serde StarkStruct>} */
int b200_setup_new(const char* setup_json, const uint64_t* const_rowmajor, size_t n_rows, size_t n_consts, b200_setup_t** out);
int b200_setup_const_root(const b200_setup_t* s, uint64_t root_out[4]);        /* StarkSetup.const_root */
/* `StarkSetup` is serde-serializable in the reference (stark_setup.rs:13-19: const_tree, const_root, starkinfo, program), so the
 * constant LDE and tree are built once per CIRCUIT.  Export writes one file (setup JSON + constant polynomials, their extension and
 * the tree nodes as they sit in device memory); import uploads it without recomputing anything.  A setup serves one proof at a time
 * (its workspace is shared; calls on the same setup serialise on an internal lock), different setups may be used from different
 * host threads concurrently. */
/* nBits, nBitsExt, number of stage-1 committed columns (the width of the trace b200_stark_gen expects), number of constant columns */
int b200_setup_shape(const b200_setup_t* s, size_t shape_out[4]);
int b200_setup_export(const b200_setup_t* s, const char* path);
int b200_setup_import(const char* path, b200_setup_t** out);
void b200_setup_free(b200_setup_t* s);
/* The reference's `stark_prove` asserts `stark_verify` on every proof it produces (starky/src/prove.rs:124-132).  With the flag on,
 * b200_stark_gen does the same before it returns (B200_ERR_INTERNAL and a reason in b200_last_error() when the verifier rejects).
 * Off by default: the check is host work on the proof, about 20 ms for a 2^24-row Goldilocks proof. */
int b200_setup_set_self_verify(b200_setup_t* s, int on);
/* `stark_verify` (starky/src/stark_verify.rs:21-121) with `FRI::verify` (fri.rs:187-297) and the Merkle `verify_group_proof` of the
 * setup's hash back-end.  setup_json as for b200_setup_new; const_root = StarkSetup.const_root (4 x u64; BN128 / BLS12-381: the
 * canonical scalar, little-endian).  *accepted_out = 1 / 0; *reason_out (optional, b200_free) names the first failed check; a proof
 * that does not parse is rejected, not an error.  Host code for "GL" (runs without a GPU); the 254 / 255-bit Poseidon of "BN128" /
 * "BLS12381" runs on the device. */
int b200_stark_verify(const char* setup_json, const uint64_t const_root[4], const char* proof_json, int* accepted_out, char** reason_out);
/* Step programs (`calculate_exps*`, stark_gen.rs:752-963) run as kernels specialised per program: the library generates
 * straight-line CUDA for each one and compiles it at first use with NVRTC (B200_JIT=0 or a missing libnvrtc selects the
 * generic interpreter kernel instead; both run on the GPU and give identical results).  Host-only hooks for inspection: */
int b200_debug_step_program_source(const char* setup_json, const char* which /* "step2prev" .. "step52ns" */, char** source_out, size_t* len_out);
int b200_debug_jit_compile(const char* source, size_t* cubin_bytes_out);
/* the window width / window count the multiexp would pick for n points (host logic, no GPU): table_mode = b200_msm_table_*, else b200_msm_* */
int b200_debug_msm_window(int curve, size_t n, int table_mode, unsigned* window_bits_out, unsigned* windows_out);
/* The Fiat-Shamir transcript (`TranscriptGL`, starky/src/transcript.rs:45-75) hashes a few 32-byte roots and evaluations per proof:
 * those single permutations run on the HOST inside the library (csrc/poseidon_host.cpp), like in the reference; everything that
 * hashes data runs on the device.  Host-only hook so that the CPU test-suite can check that code against the reference KATs. */
int b200_debug_transcript_poseidon(const uint64_t in12[12], uint64_t out12[12]);
int b200_stark_gen(b200_setup_t* s, const uint64_t* cm_rowmajor, size_t n_rows, size_t n_cols, const char* prover_addr,
                   char** proof_json_out, size_t* len_out);
int b200_stark_gen_dev(b200_setup_t* s, const uint64_t* d_cm_rowmajor, size_t n_rows, size_t n_cols, const char* prover_addr,
                       char** proof_json_out, size_t* len_out);

/* ---- Multi-scalar multiplication on G1 / G2 of BN254 and BLS12-381: the multiexp calls of `create_random_proof`
 *      behind `Groth16::prove` (groth16/src/groth16.rs:88-96 -> bellman_ce, BN254; groth16.rs:45-57 ->
 *      bellperson + blstrs, BLS12-381; CLI groth16/src/api.rs:144-177,247-271).  In-memory forms of those libraries:
 *      bases = affine (x, y), little-endian MONTGOMERY limbs (R = 2^256 for BN254, 2^384 for BLS12-381; a G2
 *      coordinate is c0 || c1), the all-zero pair = point at infinity; scalars = canonical 4 x u64 `Repr`;
 *      result = Jacobian (X, Y, Z) in Montgomery form, written to HOST memory (normalised: Z = R, or (0, R, 0)
 *      for the point at infinity).  Sizes: b200_msm_point_bytes(curve) per base (64 / 128 / 96 / 192), 1.5x that
 *      for a result. ---------------------------------------------------------------------------------------- */
#define B200_CURVE_BN254_G1 0
#define B200_CURVE_BN254_G2 1
#define B200_CURVE_BLS12381_G1 2
#define B200_CURVE_BLS12381_G2 3
size_t b200_msm_point_bytes(int curve);        /* 0 for an unknown curve id */
int b200_msm(int curve, const void* bases_affine, const void* scalars, size_t n, void* out_jacobian);
int b200_msm_dev(int curve, const void* d_bases_affine, const void* d_scalars, size_t n, void* out_jacobian);
int b200_msm_bn254_g1(const void* bases_affine, const void* scalars, size_t n, void* out_jacobian96);
int b200_msm_bn254_g1_dev(const void* d_bases_affine, const void* d_scalars, size_t n, void* out_jacobian96);
int b200_msm_bn254_g2(const void* bases_affine, const void* scalars, size_t n, void* out_jacobian192);
int b200_msm_bn254_g2_dev(const void* d_bases_affine, const void* d_scalars, size_t n, void* out_jacobian192);
int b200_msm_bls12381_g1(const void* bases_affine, const void* scalars, size_t n, void* out_jacobian144);
int b200_msm_bls12381_g1_dev(const void* d_bases_affine, const void* d_scalars, size_t n, void* out_jacobian144);
int b200_msm_bls12381_g2(const void* bases_affine, const void* scalars, size_t n, void* out_jacobian288);
int b200_msm_bls12381_g2_dev(const void* d_bases_affine, const void* d_scalars, size_t n, void* out_jacobian288);
/* Per-circuit tables.  The bases of every groth16 multiexp are the proving key (`Parameters::{a, b_g1, b_g2, h, l}`,
 * groth16/src/api.rs:161,545-550): fixed per circuit, uploaded once.  b200_msm_table_new keeps them on the device together with
 * their shifted copies 2^(c w) P_i (w < windows), computed once; every window digit of every scalar is then a small scalar for one
 * table entry and all of them share ONE bucket set: no per-window bucket reduction, no doublings, wider windows.  Memory:
 * windows x n x b200_msm_point_bytes(curve).  `on_device` != 0: the pointer is device memory.  b200_msm_table_run moves only the
 * scalars (32 B each) and returns the same normalised Jacobian triple as b200_msm. */
typedef struct b200_msm_table b200_msm_table_t;
int b200_msm_table_new(int curve, const void* bases_affine, size_t n, int on_device, b200_msm_table_t** out);
int b200_msm_table_info(const b200_msm_table_t* t, unsigned* window_bits_out, unsigned* windows_out, size_t* n_out);
int b200_msm_table_run(const b200_msm_table_t* t, const void* scalars, int on_device, void* out_jacobian);
/* Partial sums (one chunk of a multiexp per GPU, combined by b200_points_sum_dev after the all-gather): with the flag on the result is an
 * UN-normalised Jacobian triple of the same point (Z != 1) -- the normalisation is a field inversion, 0.17 ms of single-thread latency
 * per call, which only the combined sum needs. */
int b200_msm_table_set_partial_output(b200_msm_table_t* t, int on);
void b200_msm_table_free(b200_msm_table_t* t);
/* sum of `count` (X, Y, Z) triples held in DEVICE memory (the all-gathered per-GPU partial sums), normalised, to host memory */
int b200_points_sum_dev(int curve, const void* d_points_jacobian, size_t count, void* out_jacobian);
/* out = a + b on (X, Y, Z) triples (host memory): combines per-GPU partial sums after the all-gather */
int b200_point_add(int curve, const void* a, const void* b, void* out);
int b200_bn254_g1_add(const void* a96, const void* b96, void* out96);
/* bench/test utility: n deterministic pseudo-random points.  G1: x from SplitMix64(seed, i), y = sqrt(x^3 + b);
 * G2: [k_i] G2_generator with k_i = SplitMix64(seed, i) | 1 */
int b200_random_points_dev(int curve, void* d_bases_affine, size_t n, uint64_t seed);
int b200_bn254_g1_random_points_dev(void* d_bases_affine, size_t n, uint64_t seed);

/* ---- groth16 scalar-field domain: bellman_ce `domain::EvaluationDomain::{fft, ifft, coset_fft,
 *      icoset_fft}` and the quotient computation of `create_random_proof` (a, b, c evaluations -> the coefficients fed to
 *      the `h` multiexp), behind `Groth16::prove` (groth16/src/groth16.rs:88-96; bellperson twin :45-57).  Elements are the
 *      libraries' in-memory `Fr`: 4 x u64 little-endian MONTGOMERY limbs (R = 2^256), natural order in and out;
 *      omega_m = (7^t)^(2^(S - log2 m)), r - 1 = 2^S t, coset generator 7 (un-vendored upstream constants, see
 *      oracle/fr_domain.py).  `b200_groth16_h` returns m - 1 CANONICAL `Repr`s, the form `b200_msm*` takes as scalars. ---- */
#define B200_FR_BN254 0
#define B200_FR_BLS12381 1
#define B200_FFT 0
#define B200_IFFT 1
#define B200_COSET_FFT 2
#define B200_ICOSET_FFT 3
int b200_fr_fft(int field, void* data /* 2^log_n x 32 B, in place */, unsigned log_n, int mode);
int b200_fr_fft_dev(int field, void* d_data, unsigned log_n, int mode);
int b200_groth16_h(int field, const void* a, const void* b, const void* c, unsigned log_m, void* h_out /* (2^log_m - 1) x 32 B */);
int b200_groth16_h_dev(int field, void* d_a /* overwritten */, void* d_b /* overwritten */, void* d_c /* overwritten */, unsigned log_m, void* d_h_out);

/* ---- `Groth16::prove` in one call (groth16/src/groth16.rs:88-96; CLI groth16/src/api.rs:144-203).
 *      b200_groth16_pk_read = `read_pk_from_file` -> `Parameters::read(reader, false)` (api.rs:161,545-550) on the bytes of the
 *      file: bellman's `Parameters::write` layout, vk (alpha_g1, beta_g1, beta_g2, gamma_g2, delta_g1, delta_g2, u32 BE n, ic[n])
 *      then h, l, a, b_g1, b_g2, each a u32 BE count + uncompressed big-endian points (G2: c1 before c0; byte 0 bit 6 = infinity).
 *      The five vectors are uploaded ONCE and kept as resident MSM tables.
 *      b200_groth16_prove = the body of bellman's `create_proof` after `circuit.synthesize(&mut prover)`: the quotient H
 *      (3 ifft + 3 coset_fft + pointwise + icoset_fft) chained on the device into the h multiexp, the l / a / b_g1 / b_g2
 *      multiexps over the resident tables, and the final linear combinations with r, s (the values `create_random_proof` draws).
 *        a, b, c          : prover.a / b / c, n_constraints in-memory `Fr`s each (4 x u64 MONTGOMERY), incl. the input-consistency rows
 *        inputs, aux      : input_assignment / aux_assignment as canonical `Repr`s (4 x u64), what bellman passes to multiexp
 *        *_density        : one byte per variable (DensityTracker bits): a_aux_density, b_input_density, b_aux_density
 *        proof_out        : A (G1 affine) || B (G2 affine) || C (G1 affine), in-memory Montgomery words (2+4+2 coordinates)
 *      b200_wtns_read     = `load_witness_from_bin_reader` (algebraic/src/reader.rs:87-138): n_out values, 4 x u64 canonical each;
 *      call with out = NULL to size the buffer. ---------------------------------------------------------------------------------- */
#define B200_G16_BN128 0
#define B200_G16_BLS12381 1
typedef struct b200_groth16_pk b200_groth16_pk_t;
int b200_groth16_pk_read(int curve, const void* parameters_bytes, size_t len, b200_groth16_pk_t** out);
int b200_groth16_pk_info(const b200_groth16_pk_t* pk, size_t counts_out[6] /* h, l, a, b_g1, b_g2, ic */);
void b200_groth16_pk_free(b200_groth16_pk_t* pk);
int b200_groth16_prove(const b200_groth16_pk_t* pk, const void* a, const void* b, const void* c, size_t n_constraints,
                       const uint64_t* inputs, size_t n_inputs, const uint64_t* aux, size_t n_aux,
                       const unsigned char* a_aux_density, const unsigned char* b_input_density, const unsigned char* b_aux_density,
                       const uint64_t r[4], const uint64_t s[4], void* proof_out);
int b200_wtns_read(const void* bytes, size_t len, int curve, uint64_t* out, size_t out_capacity, size_t* n_out);

/* ---- compressor12 exec phase without the file round trip (recursion/src/compressor12/compressor12_exec.rs:19-108): extends the
 *      circom witness with the PlonkAdd rows of the `.exec` vector (its JSON array of u64: adds_len, map_rows, adds[4 adds_len],
 *      s_map[12 map_rows]; coefficients are raw Montgomery limbs) and fills the 12 committed columns Compressor.a[0..12] through the
 *      signal map into a row-major n_rows x 12 matrix -- the `.cm` contents `exec` would save and `stark_prove` would load again.
 *      The `_dev` variant leaves the trace in device memory, ready for b200_stark_gen_dev.  witness: canonical-or-not u64 (reduced
 *      like `FGL::from`).
 *      b200_pols_load_dev streams an existing `.cm` / `.const` file (`PolsArray::load`, starky/src/polsarray.rs:137-217: row-major
 *      little-endian u64) into device memory through pinned staging buffers. ------------------------------------------------------ */
int b200_c12_exec(const uint64_t* exec_vec, size_t exec_len, const uint64_t* witness, size_t n_witness, size_t n_rows, uint64_t* cm_rowmajor_out);
int b200_c12_exec_dev(const uint64_t* exec_vec, size_t exec_len, const uint64_t* witness, size_t n_witness, size_t n_rows, uint64_t* d_cm_rowmajor_out);
int b200_pols_load_dev(const char* path, size_t n_rows, size_t n_cols, uint64_t* d_rowmajor_out);

/* ---- bench/test utility: the Fibonacci trace behind starky/data/fib.cm.gl (row i = (F_i, F_{i+1}), F_0=1, F_1=2),
 *      written row-major (2^log_n x 2) into device memory. -------------------------------------------------- */
int b200_fib_trace_dev(uint64_t* d_cm_rowmajor, unsigned log_n);

#ifdef __cplusplus
}
#endif
#endif
