// Internal (C++) interfaces between the translation units of libb200zk.so.  Not part of the C-ABI.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string>
#include <vector>
#include <array>
#include <stdexcept>
#include <cuda_runtime.h>

typedef uint64_t u64;
typedef uint32_t u32;

#define B200_CUDA_CHECK(x)                                                                             \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) throw std::runtime_error(std::string("CUDA error: ") + cudaGetErrorString(e_) + " at " + __FILE__ + ":" + std::to_string(__LINE__)); \
    } while (0)

namespace b200 {

// ------------------------------------------------------------------------------------------------ layout
// Device matrices are COLUMN-MAJOR: column c of an (rows x w) matrix is the contiguous range [c*rows, (c+1)*rows).
// (The reference's host buffers are row-major [row][col], starky/src/polsarray.rs:219-227; the C-ABI transposes.)
// A "column view" lets the hashing kernels read logical column c of a virtual matrix at
//     base + (c / a) * s1 + (c % a) * s2           (elements)
// plain column-major: a = 1, s1 = rows.  FRI layer leaves (fri.rs:299-317): a = 3, s1 = n_groups, s2 = n.
struct ColView { const u64* base; u32 a; u64 s1; u64 s2; };
static inline ColView colview_plain(const u64* base, size_t rows) { ColView v{base, 1u, (u64)rows, 0}; return v; }

// ------------------------------------------------------------------------------------------------ devices / streams / timing
// Threading model: every compute entry point of the C-ABI holds a per-DEVICE lock for its whole duration (capi.cpp `guard`), so
// calls on one device serialise (they share that device's grow-only workspaces and its launch stream) while host threads that
// drive different devices run concurrently.  Per-device state is indexed by current_device() (bounds-checked, B200_MAX_DEVICES);
// the caches shared between devices (twiddle / power tables, JIT kernels) carry their own mutexes.
#define B200_MAX_DEVICES 16
int current_device();                       // cudaGetDevice, throws when the index is outside [0, B200_MAX_DEVICES)
cudaStream_t stream();                      // the stream every kernel of the library is launched on (per device)
void set_stream(cudaStream_t s);            // for the current device
struct KernelTimer;                         // see timing.cpp
void timing_enable(bool on);
void timing_reset();
// name -> (launches, total ms, algorithmic bytes); valid after timing_collect()
struct TimingRow { std::string name; int launches; double ms; double bytes; };
std::vector<TimingRow> timing_collect();
void timing_begin(const char* name, double algo_bytes);
void timing_end();
struct ScopedTimer { ScopedTimer(const char* n, double bytes = 0) { timing_begin(n, bytes); } ~ScopedTimer() { timing_end(); } };
u64 launch_count();
void launch_count_add(u64 n);

// ------------------------------------------------------------------------------------------------ ntt.cu
void transpose_rm_to_cm(const u64* d_in, u64* d_out, size_t rows, size_t w);   // [row][col] -> [col][row]
void transpose_cm_to_rm(const u64* d_in, u64* d_out, size_t rows, size_t w);
// In-place-capable batched transforms over column-major data (w columns of 2^log_n each).
void ntt_cols(const u64* d_in, u64* d_out, size_t w, unsigned log_n, bool inverse);
// forward transform of size 2^log_n whose input columns hold only `in_rows` (<= 2^log_n) rows, the rest being zero
void ntt_cols_padded(const u64* d_in, size_t in_rows, u64* d_out, size_t w, unsigned log_n);
// Coset LDE (starky/src/fft_p.rs:255-355): out[c][j] = P_c(49 * w_ext^j), P_c interpolating in[c][.] on w_n^i.
void lde_cols(const u64* d_in, u64* d_out, size_t w, unsigned log_n, unsigned log_n_ext);
struct DevPowTab { const u64* lo; const u64* hi; };
DevPowTab powtab(u64 base, unsigned log_range);          // cached on device: base^e, e < 2^log_range
u64 h_root(unsigned k); u64 h_root_inv(unsigned k);      // constant.rs:54-68
u64 h_mul(u64 a, u64 b); u64 h_pow(u64 a, u64 e); u64 h_inv(u64 a); u64 h_add(u64 a, u64 b); u64 h_sub(u64 a, u64 b);

// ------------------------------------------------------------------------------------------------ merkle.cu
size_t merkle_n_nodes(size_t height);                   // merklehash.rs:47-61
void poseidon12_host(const u64 in12[12], u64 out12[12]);          // poseidon_host.cpp: one permutation on the host (transcript)
void poseidon_perm_host(const u64 in12[12], u64 out12[12]);       // = poseidon12_host
void poseidon_perm_device(const u64 in12[12], u64 out12[12]);     // one permutation by the device kernel, result to host
void linearhash_rows(ColView cols, size_t width, size_t height, u64* d_digests /* height x 4 */);
// nodes[0..height) = leaf digests already in place; leaf_width 1..3 = they are zero-padded rows of that width (0: unknown)
void merkle_levels(u64* d_nodes, size_t height, size_t leaf_width = 0);
struct DevTree {
    int hash = 0;                       // 0 = GL (binary, merkle.cu), 1 = BN128, 2 = BLS12-381 (16-ary, merkle_big.cu); set before merkelize()
    ColView cols{nullptr, 1, 0, 0};     // leaves (device), logical width x height
    size_t width = 0, height = 0;
    u64* nodes = nullptr;               // merkle_n_nodes(height) x 4 (device); nullptr when degenerate
    bool degenerate = false;            // width == 0: every level is one repeated digest (merklehash.rs:311-343 on empty input)
    std::vector<std::array<u64, 4>> level_digest;   // degenerate trees: digest of level l (l = 0: zero leaf)
    u64 root[4] = {0, 0, 0, 0};
};
void merkelize(DevTree& t, ColView cols, size_t width, size_t height, u64* d_nodes);
// openings for n_idx leaves: vals (n_idx x width) and siblings (n_idx x depth x 4; 16-ary trees: n_idx x depth x 16 x 4) on the host
void merkle_open(const DevTree& t, const std::vector<u64>& idx, std::vector<u64>& vals, std::vector<u64>& sibs, size_t& depth);
void big_merkelize_tree(DevTree& t, ColView cols, size_t width, size_t height, u64* d_nodes);      // merkle_big.cu (t.hash = 1, 2)
void big_merkle_open(const DevTree& t, const std::vector<u64>& idx, std::vector<u64>& vals, std::vector<u64>& sibs, size_t& depth);

// ------------------------------------------------------------------------------------------------ merkle_big.cu
// BN128 / BLS12-381 Poseidon back-ends (field ids: 0 = BN128, 1 = BLS12-381); digests are canonical 4 x u64
size_t big_merkle_n_nodes(size_t height);                          // merklehash_bn128.rs:26-40
int big_out_lane(int field);
void big_poseidon_host(int field, const u64* h_state_in /* t x 4: init, inputs */, int t, u64* h_state_out /* t x 4 */);
void big_leaves(int field, const u64* d_cols /* column-major GL */, size_t width, size_t height, u64* d_digests);
void big_leaves_view(int field, ColView d_cols, size_t width, size_t height, u64* d_digests);
void big_merkle_levels(int field, u64* d_nodes, size_t height);    // nodes[0..height) = leaf digests already in place

// ------------------------------------------------------------------------------------------------ evaluator.cu
struct EvOperand { u32 kind; u32 a; u32 b; u32 prime; u32 dim; };   // see evaluator.cu
struct EvOp { u32 opc; EvOperand d, s0, s1; };
struct EvSection { u64* base; u64 rows; };
struct EvProgram {
    std::vector<EvOp> ops;
    std::vector<u64> consts;        // dim-1 constants (numbers, publics)
    u32 n_slots = 0;                // tmp slots (u64 each) after liveness allocation
    EvOp* d_ops = nullptr; u64* d_consts = nullptr;
};
std::string eval_jit_source(const EvProgram& p);        // CUDA source of the specialised kernel for one step program (evaluator.cu)
std::string jit_compile_cubin(const std::string& src, std::string& err);     // NVRTC, host only (jit.cpp)
void eval_program(const EvProgram& p, const EvSection* secs, int n_secs, const u64* h_f3consts, int n_f3,
                  DevPowTab x_tab, u64 x_start, const u64* d_zi, u32 zi_mask, size_t n, size_t next, double algo_bytes);
void zh_inv_table(u64* d_out, unsigned nbits, unsigned ext_bits);                       // stark_gen.rs:575-592
void quotient_split(const u64* d_qq1, u64* d_qq2, size_t n, size_t n_ext, size_t q_dim, size_t q_deg, unsigned nbits);
void f3_powers(const u64 base3[3], u64* d_out /* 3 x n col-major */, size_t n);          // LEv (stark_gen.rs:416-427)
void eval_dot(const u64* d_col0, size_t col_stride, int dim, unsigned ext_bits, const u64* d_L /* 3 x n */, size_t n, u64 out3[3]);
void eval_dot_multi(const u64* const* d_col0, const size_t* col_stride, const int* dim, int m /* <= 8 */, unsigned ext_bits, const u64* d_L, size_t n, u64* out /* m x 3 */);
void xdivxsub(DevPowTab x_tab, u64 x_start, size_t n_ext, const u64 pt3[3], u64* d_out /* 3 x n_ext */, const u64* scale3 = nullptr, const char* timer_name = "xdivxsub");
void lagrange_row(DevPowTab x_n_tab, unsigned nbits, const u64 y3[3], u64* d_out /* 3 x 2^nbits */);   // = iNTT of the powers of y (LEv / LpEv, stark_gen.rs:416-436)
void fib_trace(u64* d_out_rowmajor, size_t n);   // bench/test utility: the generator behind starky/data/fib.cm.gl
void fri_fold(const u64* d_pol /* 3 x n */, u64* d_out /* 3 x n>>red */, unsigned pol_bits, unsigned red_bits, u64 sinv0, const u64 sx3[3]);

// ------------------------------------------------------------------------------------------------ lookup.cu
// stark_gen.rs:625-666: column-major polynomials with `dim` consecutive columns of n rows
void calculate_H1H2(const u64* f, const u64* t, u32 dim, size_t n, u64* h1, u64* h2);
void calculate_Z(const u64* num, u32 num_dim, const u64* den, u32 den_dim, u64* z, u32 z_dim, size_t n, u64* d_tmp);
size_t calculate_Z_tmp_u64(size_t n);

// ------------------------------------------------------------------------------------------------ msm.cu
// curve ids: 0 = BN254 G1, 1 = BN254 G2, 2 = BLS12-381 G1, 3 = BLS12-381 G2 (B200_CURVE_* in include/b200zk.h).
// bases: n affine points (x, y) in Montgomery limbs, all-zero = infinity; scalars: n x 32 B canonical;
// out: Jacobian (X, Y, Z) on the host, 3 coordinates of msm_point_bytes(curve) / 2 bytes each.
size_t msm_point_bytes(int curve);
void msm_dev(int curve, const void* d_bases, const void* d_scalars, size_t n, void* h_out);
void msm_host_buffers(int curve, const void* bases, const void* scalars, size_t n, void* h_out);
void msm_point_add(int curve, const void* a, const void* b, void* out);
void msm_random_points_dev(int curve, void* d_bases, size_t n, u64 seed);
// per-circuit table of shifted bases 2^(c w) P_i (device resident): every window of every scalar then shares one bucket set
struct MsmTable;
MsmTable* msm_table_new(int curve, const void* d_bases, size_t n);
void msm_table_free(MsmTable* t);
void msm_table_info(const MsmTable* t, u32* c, u32* nwin, size_t* n);
void msm_window_choice(int curve, size_t n, bool table_mode, u32* c_out, u32* nwin_out);
void msm_table_set_partial_output(MsmTable* t, bool on);   // results stay un-normalised Jacobian triples (no inversion): for partial sums
void msm_table_run(const MsmTable* t, const void* d_scalars, void* h_out);
void msm_table_run_host(const MsmTable* t, const void* scalars, void* h_out);      // scalars in host memory (copied in: 32 B per scalar)
void msm_points_sum_dev(int curve, const void* d_points /* count Jacobian triples on the device */, size_t count, void* h_out);

// ------------------------------------------------------------------------------------------------ fr_ntt.cu
// scalar-field domain (field ids: 0 = BN254 Fr, 1 = BLS12-381 Fr); elements are 32-byte Montgomery `Fr`s on the device
void fr_fft_dev(int field, void* d_data, unsigned log_n, int mode /* 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft */);
void groth16_h_dev(int field, void* d_a, void* d_b, void* d_c, unsigned log_m, void* d_h_out /* (2^log_m - 1) canonical */);

// ------------------------------------------------------------------------------------------------ groth16.cu
struct G16Pk;
G16Pk* groth16_pk_read(int curve /* 0 BN128, 1 BLS12381 */, const void* data, size_t len);       // bellman `Parameters::read`
void groth16_pk_free(G16Pk* pk);
void groth16_pk_info(const G16Pk* pk, size_t out[6]);                                            // n_h, n_l, n_a, n_b_g1, n_b_g2, n_ic
void groth16_prove(const G16Pk* pk, const void* a_evals, const void* b_evals, const void* c_evals, size_t n_constraints,
                   const u64* input_assignment, size_t n_inputs, const u64* aux_assignment, size_t n_aux,
                   const unsigned char* a_aux_density, const unsigned char* b_input_density, const unsigned char* b_aux_density,
                   const u64 r[4], const u64 s[4], void* proof_out);
size_t wtns_read(const void* data, size_t len, const unsigned char prime_le32[32], u64* out, size_t out_cap);

// ------------------------------------------------------------------------------------------------ c12_exec.cu
void c12_extend_witness(const u64* adds, size_t adds_len, std::vector<u64>& w);           // host: the PlonkAdd chain of compressor12_exec.rs:58-64
void c12_exec_dev(const u64* exec, size_t exec_len, const u64* witness, size_t n_witness, size_t n_rows, u64* d_cm_rowmajor /* n_rows x 12 */);
void pols_load_dev(const char* path, size_t n_u64, u64* d_out);

// ------------------------------------------------------------------------------------------------ arena
struct Arena {
    char* base = nullptr; size_t cap = 0, off = 0, high = 0;
    void reserve(size_t bytes);
    void reset() { off = 0; }
    u64* alloc_u64(size_t n);        // 256-byte aligned; throws when exhausted
    void release();
};

}  // namespace b200
