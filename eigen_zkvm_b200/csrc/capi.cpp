// extern "C" surface of libb200zk.so (declared in include/b200zk.h).  Translates exceptions to error codes,
// moves host buffers (row-major, reference layout) to the device and back.
#include "../../include/b200zk.h"
#include "b200_internal.h"
#include "stark.h"
#include <cstring>
#include <cstdlib>
#include <sstream>
#include <mutex>

static thread_local std::string g_err;
struct b200_setup { b200::Setup* s; };

namespace {
// one lock per device: a call holds it from entry to return (see b200_internal.h, "Threading model")
static std::recursive_mutex g_dev_mu[B200_MAX_DEVICES + 1];
static std::recursive_mutex& device_mutex() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= B200_MAX_DEVICES) { (void)cudaGetLastError(); return g_dev_mu[B200_MAX_DEVICES]; }     // no GPU: host-only hooks
    return g_dev_mu[dev];
}
template <class F> int guard(F&& f) {
    try { std::lock_guard<std::recursive_mutex> lk(device_mutex()); f(); return B200_OK; }
    catch (const std::invalid_argument& e) { g_err = e.what(); return B200_ERR_ARG; }
    catch (const std::exception& e) {
        g_err = e.what();
        if (g_err.find("CUDA error") != std::string::npos) return B200_ERR_CUDA;
        if (g_err.find("not implemented") != std::string::npos || g_err.find("not supported") != std::string::npos) return B200_ERR_UNSUPPORTED;
        return B200_ERR_INTERNAL;
    } catch (...) { g_err = "unknown error"; return B200_ERR_INTERNAL; }
}
void need_device() {
    int n = 0; cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) throw std::runtime_error(std::string("CUDA error: no CUDA device available (") + cudaGetErrorString(e) + "); libb200zk has no CPU fallback");
}
struct DevBuf {
    u64* p = nullptr;
    explicit DevBuf(size_t n) { B200_CUDA_CHECK(cudaMalloc(&p, (n ? n : 1) * 8)); }
    ~DevBuf() { if (p) cudaFree(p); }
};
char* dup_out(const std::string& s, size_t* len) { char* o = (char*)malloc(s.size() + 1); memcpy(o, s.data(), s.size()); o[s.size()] = 0; if (len) *len = s.size(); return o; }
}  // namespace

extern "C" {

const char* b200_last_error(void) { return g_err.c_str(); }
const char* b200_version(void) { return "b200zk 0.1 (sm_100a)"; }
void b200_free(void* p) { free(p); }
int b200_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
int b200_set_device(int device) { return guard([&] { B200_CUDA_CHECK(cudaSetDevice(device)); }); }
int b200_set_stream(void* s) { return guard([&] { need_device(); b200::set_stream((cudaStream_t)s); }); }
int b200_timing_enable(int on) { b200::timing_reset(); b200::timing_enable(on != 0); return B200_OK; }
int b200_timing_report(char** json_out, size_t* len_out) {
    return guard([&] {
        auto rows = b200::timing_collect();
        std::ostringstream o; o << '[';
        for (size_t i = 0; i < rows.size(); i++) { if (i) o << ','; o << "{\"name\":\"" << rows[i].name << "\",\"launches\":" << rows[i].launches << ",\"ms\":" << rows[i].ms << ",\"bytes\":" << (double)rows[i].bytes << '}'; }
        o << ']';
        b200::timing_reset();
        *json_out = dup_out(o.str(), len_out);
    });
}
uint64_t b200_kernel_launches(void) { return b200::launch_count(); }

// ---------------------------------------------------------------------------------------------- NTT family
static void ntt_host(const uint64_t* in, uint64_t* out, size_t w, unsigned log_n, unsigned log_out, int mode) {
    need_device();
    if (!in || !out) throw std::invalid_argument("null buffer");
    if (log_out > 27 || log_n > log_out) throw std::invalid_argument("log size out of range (<= 27)");
    if (w == 0) return;       // fft_p.rs:262-264: empty input is a no-op
    size_t n = (size_t)1 << log_n, no = (size_t)1 << log_out;
    DevBuf rm(std::max(n, no) * w), cm_in(n * w), cm_out(no * w);
    B200_CUDA_CHECK(cudaMemcpyAsync(rm.p, in, n * w * 8, cudaMemcpyHostToDevice, b200::stream()));
    b200::transpose_rm_to_cm(rm.p, cm_in.p, n, w);
    if (mode == 0) b200::ntt_cols(cm_in.p, cm_out.p, w, log_n, false);
    else if (mode == 1) b200::ntt_cols(cm_in.p, cm_out.p, w, log_n, true);
    else b200::lde_cols(cm_in.p, cm_out.p, w, log_n, log_out);
    b200::transpose_cm_to_rm(cm_out.p, rm.p, no, w);
    B200_CUDA_CHECK(cudaMemcpyAsync(out, rm.p, no * w * 8, cudaMemcpyDeviceToHost, b200::stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
}
int b200_gl_ntt(const uint64_t* in, uint64_t* out, size_t n_cols, unsigned log_n) { return guard([&] { ntt_host(in, out, n_cols, log_n, log_n, 0); }); }
int b200_gl_intt(const uint64_t* in, uint64_t* out, size_t n_cols, unsigned log_n) { return guard([&] { ntt_host(in, out, n_cols, log_n, log_n, 1); }); }
int b200_gl_lde(const uint64_t* in, uint64_t* out, size_t n_cols, unsigned log_n, unsigned log_n_ext) { return guard([&] { ntt_host(in, out, n_cols, log_n, log_n_ext, 2); }); }
int b200_gl_ntt_dev(const uint64_t* d_in, uint64_t* d_out, size_t n_cols, unsigned log_n, int inverse) {
    return guard([&] { need_device(); if (log_n > 27) throw std::invalid_argument("log size out of range (<= 27)"); b200::ntt_cols(d_in, d_out, n_cols, log_n, inverse != 0); B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream())); });
}
int b200_gl_lde_dev(const uint64_t* d_in, uint64_t* d_out, size_t n_cols, unsigned log_n, unsigned log_n_ext) {
    return guard([&] { need_device(); if (log_n_ext > 27 || log_n > log_n_ext) throw std::invalid_argument("log size out of range"); b200::lde_cols(d_in, d_out, n_cols, log_n, log_n_ext); B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream())); });
}

// ---------------------------------------------------------------------------------------------- hashing
int b200_gl_poseidon(const uint64_t in8[8], const uint64_t cap4[4], uint64_t out12[12]) {
    return guard([&] {
        need_device();
        u64 in[12]; for (int i = 0; i < 8; i++) in[i] = in8[i] % GL_P_HOST; for (int i = 0; i < 4; i++) in[8 + i] = cap4[i] % GL_P_HOST;
        b200::poseidon_perm_device(in, out12);
    });
}
int b200_gl_linearhash(const uint64_t* rows, size_t width, size_t n_rows, uint64_t* digests_out) {
    return guard([&] {
        need_device();
        if (n_rows == 0) return;
        DevBuf rm(width * n_rows), cm(width * n_rows), dg(n_rows * 4);
        if (width) { B200_CUDA_CHECK(cudaMemcpyAsync(rm.p, rows, width * n_rows * 8, cudaMemcpyHostToDevice, b200::stream())); b200::transpose_rm_to_cm(rm.p, cm.p, n_rows, width); }
        b200::linearhash_rows(b200::colview_plain(cm.p, n_rows), width, n_rows, dg.p);
        B200_CUDA_CHECK(cudaMemcpyAsync(digests_out, dg.p, n_rows * 32, cudaMemcpyDeviceToHost, b200::stream()));
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}
size_t b200_gl_merkle_n_nodes(size_t height) { return height ? b200::merkle_n_nodes(height) : 0; }
int b200_gl_merkelize(const uint64_t* leaves, size_t width, size_t height, uint64_t* nodes_out) {
    return guard([&] {
        need_device();
        if (height == 0) throw std::invalid_argument("height must be > 0");
        size_t nn = b200::merkle_n_nodes(height);
        DevBuf rm(width * height), cm(width * height), nodes(nn * 4);
        B200_CUDA_CHECK(cudaMemsetAsync(nodes.p, 0, nn * 32, b200::stream()));
        if (width) { B200_CUDA_CHECK(cudaMemcpyAsync(rm.p, leaves, width * height * 8, cudaMemcpyHostToDevice, b200::stream())); b200::transpose_rm_to_cm(rm.p, cm.p, height, width);
            b200::linearhash_rows(b200::colview_plain(cm.p, height), width, height, nodes.p); }
        b200::merkle_levels(nodes.p, height);      // width 0: levels over zero digests, like merklehash.rs:311-343
        B200_CUDA_CHECK(cudaMemcpyAsync(nodes_out, nodes.p, nn * 32, cudaMemcpyDeviceToHost, b200::stream()));
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}
int b200_gl_merkelize_dev(const uint64_t* d_leaves_colmajor, size_t width, size_t height, uint64_t* d_nodes_out) {
    return guard([&] {
        need_device();
        if (height == 0 || width == 0) throw std::invalid_argument("width and height must be > 0");
        b200::linearhash_rows(b200::colview_plain(d_leaves_colmajor, height), width, height, d_nodes_out);
        b200::merkle_levels(d_nodes_out, height);
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}

// ---------------------------------------------------------------------------------------------- BN128 / BLS12-381 hashing
int b200_big_poseidon(int field, const uint64_t* inputs, size_t n_inputs, const uint64_t init4[4], uint64_t* state_out) {
    return guard([&] {
        need_device();
        if (!inputs || !init4 || !state_out) throw std::invalid_argument("null buffer");
        if (n_inputs == 0 || n_inputs > 16) throw std::invalid_argument("Wrong inputs length");    // poseidon_bn128_opt.rs:112-118
        std::vector<u64> st(4 * (n_inputs + 1));
        memcpy(st.data(), init4, 32); memcpy(st.data() + 4, inputs, 32 * n_inputs);
        b200::big_poseidon_host(field, st.data(), (int)n_inputs + 1, state_out);
    });
}
int b200_big_hash(int field, const uint64_t* inputs, size_t n_inputs, const uint64_t init4[4], uint64_t out4[4]) {
    uint64_t st[17 * 4];
    int rc = b200_big_poseidon(field, inputs, n_inputs, init4, st);
    if (rc) return rc;
    memcpy(out4, st + 4 * b200::big_out_lane(field), 32);
    return B200_OK;
}
int b200_big_linearhash(int field, const uint64_t* rows, size_t width, size_t n_rows, uint64_t* digests_out) {
    return guard([&] {
        need_device();
        b200::big_out_lane(field);
        if (n_rows == 0) return;
        if (!digests_out || (!rows && width)) throw std::invalid_argument("null buffer");
        DevBuf rm(std::max<size_t>(width * n_rows, 1)), cm(std::max<size_t>(width * n_rows, 1)), dg(n_rows * 4);
        if (width) { B200_CUDA_CHECK(cudaMemcpyAsync(rm.p, rows, width * n_rows * 8, cudaMemcpyHostToDevice, b200::stream())); b200::transpose_rm_to_cm(rm.p, cm.p, n_rows, width); }
        b200::big_leaves(field, cm.p, width, n_rows, dg.p);
        B200_CUDA_CHECK(cudaMemcpyAsync(digests_out, dg.p, n_rows * 32, cudaMemcpyDeviceToHost, b200::stream()));
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}
size_t b200_big_merkle_n_nodes(size_t height) { return height ? b200::big_merkle_n_nodes(height) : 0; }
int b200_big_merkelize(int field, const uint64_t* leaves, size_t width, size_t height, uint64_t* nodes_out) {
    return guard([&] {
        need_device();
        b200::big_out_lane(field);
        if (height == 0) throw std::invalid_argument("height must be > 0");
        if (!nodes_out || (!leaves && width)) throw std::invalid_argument("null buffer");
        size_t nn = b200::big_merkle_n_nodes(height);
        DevBuf rm(std::max<size_t>(width * height, 1)), cm(std::max<size_t>(width * height, 1)), nodes(nn * 4);
        B200_CUDA_CHECK(cudaMemsetAsync(nodes.p, 0, nn * 32, b200::stream()));
        if (width) { B200_CUDA_CHECK(cudaMemcpyAsync(rm.p, leaves, width * height * 8, cudaMemcpyHostToDevice, b200::stream())); b200::transpose_rm_to_cm(rm.p, cm.p, height, width);
            b200::big_leaves(field, cm.p, width, height, nodes.p); }       // empty buffer: leaves stay zero digests (merklehash_bn128.rs:191-203)
        b200::big_merkle_levels(field, nodes.p, height);
        B200_CUDA_CHECK(cudaMemcpyAsync(nodes_out, nodes.p, nn * 32, cudaMemcpyDeviceToHost, b200::stream()));
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}
int b200_big_merkelize_dev(int field, const uint64_t* d_leaves_colmajor, size_t width, size_t height, uint64_t* d_nodes_out) {
    return guard([&] {
        need_device();
        if (height == 0 || width == 0) throw std::invalid_argument("width and height must be > 0");
        B200_CUDA_CHECK(cudaMemsetAsync(d_nodes_out, 0, b200::big_merkle_n_nodes(height) * 32, b200::stream()));
        b200::big_leaves(field, d_leaves_colmajor, width, height, d_nodes_out);
        b200::big_merkle_levels(field, d_nodes_out, height);
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}

// ---------------------------------------------------------------------------------------------- STARK
int b200_setup_new(const char* setup_json, const uint64_t* const_rowmajor, size_t n_rows, size_t n_consts, b200_setup_t** out) {
    return guard([&] {
        need_device();
        if (!setup_json || !out) throw std::invalid_argument("null argument");
        b200::Setup* s = b200::setup_new(setup_json, const_rowmajor, false, n_rows, n_consts);
        *out = new b200_setup{s};
    });
}
int b200_setup_const_root(const b200_setup_t* s, uint64_t root_out[4]) { return guard([&] { if (!s) throw std::invalid_argument("null setup"); b200::setup_const_root(s->s, root_out); }); }
void b200_setup_free(b200_setup_t* s) { if (s) { b200::setup_free(s->s); delete s; } }
int b200_setup_shape(const b200_setup_t* s, size_t shape_out[4]) { return guard([&] { if (!s || !shape_out) throw std::invalid_argument("null argument"); b200::setup_shape(s->s, shape_out); }); }
int b200_setup_export(const b200_setup_t* s, const char* path) {
    return guard([&] { need_device(); if (!s || !path) throw std::invalid_argument("null argument"); b200::setup_export(s->s, path); });
}
int b200_setup_import(const char* path, b200_setup_t** out) {
    return guard([&] { need_device(); if (!path || !out) throw std::invalid_argument("null argument"); *out = new b200_setup{b200::setup_import(path)}; });
}
int b200_setup_set_self_verify(b200_setup_t* s, int on) { return guard([&] { if (!s) throw std::invalid_argument("null setup"); b200::setup_set_self_verify(s->s, on != 0); }); }
// host code for Goldilocks proofs (no GPU needed); the BN128 / BLS12-381 back-ends hash on the device
int b200_stark_verify(const char* setup_json, const uint64_t const_root[4], const char* proof_json, int* accepted_out, char** reason_out) {
    return guard([&] {
        if (!setup_json || !const_root || !proof_json || !accepted_out) throw std::invalid_argument("null argument");
        std::string why;
        bool ok;
        try { ok = b200::stark_verify(setup_json, const_root, proof_json, why); }
        catch (const std::invalid_argument&) { throw; }
        catch (const std::exception& e) {          // a proof that does not even parse is a rejected proof, not a library failure -- unless it is the device that failed
            if (std::string(e.what()).find("CUDA error") != std::string::npos) throw;
            ok = false; why = e.what();
        }
        *accepted_out = ok ? 1 : 0;
        if (reason_out) *reason_out = dup_out(why, nullptr);
    });
}
static int gen(b200_setup_t* s, const uint64_t* cm, bool dev, size_t n_rows, size_t n_cols, const char* prover_addr, char** proof_json_out, size_t* len_out) {
    return guard([&] {
        need_device();
        if (!s || !cm || !proof_json_out) throw std::invalid_argument("null argument");
        std::string js = b200::stark_gen(s->s, cm, dev, n_rows, n_cols, prover_addr);     // prover_addr: serialized for BN128/BLS12381 only (serializer.rs:262-267)
        *proof_json_out = dup_out(js, len_out);
    });
}
// host-only hooks of the step-program JIT (no GPU needed): generated source, and its NVRTC compilation to an sm_100a cubin
int b200_debug_step_program_source(const char* setup_json, const char* which, char** source_out, size_t* len_out) {
    return guard([&] { if (!setup_json || !which || !source_out) throw std::invalid_argument("null argument");
        *source_out = dup_out(b200::step_program_source(setup_json, which), len_out); });
}
int b200_debug_transcript_poseidon(const uint64_t in12[12], uint64_t out12[12]) {
    return guard([&] { if (!in12 || !out12) throw std::invalid_argument("null argument"); b200::poseidon12_host(in12, out12); });
}
int b200_debug_jit_compile(const char* source, size_t* cubin_bytes_out) {
    return guard([&] { if (!source || !cubin_bytes_out) throw std::invalid_argument("null argument");
        std::string err; std::string cubin = b200::jit_compile_cubin(source, err);
        if (cubin.empty()) throw std::runtime_error("JIT compilation failed: " + err);
        *cubin_bytes_out = cubin.size(); });
}
int b200_stark_gen(b200_setup_t* s, const uint64_t* cm_rowmajor, size_t n_rows, size_t n_cols, const char* prover_addr, char** proof_json_out, size_t* len_out) {
    return gen(s, cm_rowmajor, false, n_rows, n_cols, prover_addr, proof_json_out, len_out);
}
int b200_stark_gen_dev(b200_setup_t* s, const uint64_t* d_cm_rowmajor, size_t n_rows, size_t n_cols, const char* prover_addr, char** proof_json_out, size_t* len_out) {
    return gen(s, d_cm_rowmajor, true, n_rows, n_cols, prover_addr, proof_json_out, len_out);
}

// ---------------------------------------------------------------------------------------------- MSM
int b200_c12_exec_dev(const uint64_t* exec_vec, size_t exec_len, const uint64_t* witness, size_t n_witness, size_t n_rows, uint64_t* d_out) {
    return guard([&] { need_device(); if (!exec_vec || !witness || !d_out) throw std::invalid_argument("null argument"); b200::c12_exec_dev(exec_vec, exec_len, witness, n_witness, n_rows, d_out); });
}
int b200_c12_exec(const uint64_t* exec_vec, size_t exec_len, const uint64_t* witness, size_t n_witness, size_t n_rows, uint64_t* out) {
    return guard([&] {
        need_device();
        if (!exec_vec || !witness || !out) throw std::invalid_argument("null argument");
        DevBuf d(n_rows * 12);
        b200::c12_exec_dev(exec_vec, exec_len, witness, n_witness, n_rows, d.p);
        B200_CUDA_CHECK(cudaMemcpy(out, d.p, n_rows * 96, cudaMemcpyDeviceToHost));
    });
}
int b200_pols_load_dev(const char* path, size_t n_rows, size_t n_cols, uint64_t* d_out) {
    return guard([&] { need_device(); if (!path || !d_out) throw std::invalid_argument("null argument"); b200::pols_load_dev(path, n_rows * n_cols, d_out); });
}
struct b200_groth16_pk { b200::G16Pk* pk; };
int b200_groth16_pk_read(int curve, const void* bytes, size_t len, b200_groth16_pk_t** out) {
    return guard([&] { need_device(); if (!bytes || !out) throw std::invalid_argument("null argument"); *out = new b200_groth16_pk{b200::groth16_pk_read(curve, bytes, len)}; });
}
int b200_groth16_pk_info(const b200_groth16_pk_t* pk, size_t counts_out[6]) {
    return guard([&] { if (!pk || !counts_out) throw std::invalid_argument("null argument"); b200::groth16_pk_info(pk->pk, counts_out); });
}
void b200_groth16_pk_free(b200_groth16_pk_t* pk) { if (pk) { b200::groth16_pk_free(pk->pk); delete pk; } }
int b200_groth16_prove(const b200_groth16_pk_t* pk, const void* a, const void* b, const void* c, size_t n_constraints,
                       const uint64_t* inputs, size_t n_inputs, const uint64_t* aux, size_t n_aux,
                       const unsigned char* a_aux_density, const unsigned char* b_input_density, const unsigned char* b_aux_density,
                       const uint64_t r[4], const uint64_t s[4], void* proof_out) {
    return guard([&] {
        need_device();
        if (!pk || !a || !b || !c || !r || !s || !proof_out || (n_inputs && !inputs) || (n_aux && !aux)) throw std::invalid_argument("null argument");
        b200::groth16_prove(pk->pk, a, b, c, n_constraints, inputs, n_inputs, aux, n_aux, a_aux_density, b_input_density, b_aux_density, r, s, proof_out);
    });
}
int b200_wtns_read(const void* bytes, size_t len, int curve, uint64_t* out, size_t out_capacity, size_t* n_out) {
    return guard([&] {
        if (!bytes || !n_out) throw std::invalid_argument("null argument");
        // scalar-field moduli, little-endian (reader.rs:117 checks the BN254 one)
        static const unsigned char bn[32] = {0x01,0x00,0x00,0xf0,0x93,0xf5,0xe1,0x43,0x91,0x70,0xb9,0x79,0x48,0xe8,0x33,0x28,0x5d,0x58,0x81,0x81,0xb6,0x45,0x50,0xb8,0x29,0xa0,0x31,0xe1,0x72,0x4e,0x64,0x30};
        static const unsigned char bls[32] = {0x01,0x00,0x00,0x00,0xff,0xff,0xff,0xff,0xfe,0x5b,0xfe,0xff,0x02,0xa4,0xbd,0x53,0x05,0xd8,0xa1,0x09,0x08,0xd8,0x39,0x33,0x48,0x7d,0x9d,0x29,0x53,0xa7,0xed,0x73};
        if (curve != 0 && curve != 1) throw std::invalid_argument("unknown curve (0 = BN128, 1 = BLS12381)");
        *n_out = b200::wtns_read(bytes, len, curve == 0 ? bn : bls, out, out_capacity);
    });
}
struct b200_msm_table { b200::MsmTable* t; };
int b200_msm_table_new(int curve, const void* bases_affine, size_t n, int on_device, b200_msm_table_t** out) {
    return guard([&] {
        need_device();
        if (!out || !bases_affine || n == 0) throw std::invalid_argument("null or empty argument");
        const size_t pb = b200::msm_point_bytes(curve);
        const void* d = bases_affine; void* tmp = nullptr;
        if (!on_device) { B200_CUDA_CHECK(cudaMalloc(&tmp, n * pb)); B200_CUDA_CHECK(cudaMemcpy(tmp, bases_affine, n * pb, cudaMemcpyHostToDevice)); d = tmp; }
        b200::MsmTable* t = nullptr;
        try { t = b200::msm_table_new(curve, d, n); } catch (...) { if (tmp) cudaFree(tmp); throw; }
        if (tmp) cudaFree(tmp);
        *out = new b200_msm_table{t};
    });
}
int b200_msm_table_info(const b200_msm_table_t* t, unsigned* c, unsigned* w, size_t* n) {
    return guard([&] { if (!t || !c || !w || !n) throw std::invalid_argument("null argument"); u32 cc, ww; b200::msm_table_info(t->t, &cc, &ww, n); *c = cc; *w = ww; });
}
int b200_debug_msm_window(int curve, size_t n, int table_mode, unsigned* window_bits_out, unsigned* windows_out) {
    return guard([&] { if (!window_bits_out || !windows_out || n == 0) throw std::invalid_argument("null or empty argument"); u32 c, w; b200::msm_window_choice(curve, n, table_mode != 0, &c, &w); *window_bits_out = c; *windows_out = w; });
}
int b200_msm_table_set_partial_output(b200_msm_table_t* t, int on) { return guard([&] { if (!t) throw std::invalid_argument("null argument"); b200::msm_table_set_partial_output(t->t, on != 0); }); }
int b200_msm_table_run(const b200_msm_table_t* t, const void* scalars, int on_device, void* out_jacobian) {
    return guard([&] {
        need_device();
        if (!t || !scalars || !out_jacobian) throw std::invalid_argument("null argument");
        if (on_device) b200::msm_table_run(t->t, scalars, out_jacobian); else b200::msm_table_run_host(t->t, scalars, out_jacobian);
    });
}
void b200_msm_table_free(b200_msm_table_t* t) { if (t) { b200::msm_table_free(t->t); delete t; } }
int b200_points_sum_dev(int curve, const void* d_points, size_t count, void* out_jacobian) {
    return guard([&] { need_device(); if (!d_points || !out_jacobian || !count) throw std::invalid_argument("null or empty argument"); b200::msm_points_sum_dev(curve, d_points, count, out_jacobian); });
}
size_t b200_msm_point_bytes(int curve) { try { return b200::msm_point_bytes(curve); } catch (...) { return 0; } }
int b200_msm(int curve, const void* bases_affine, const void* scalars, size_t n, void* out_jacobian) {
    return guard([&] { need_device(); if ((!bases_affine || !scalars) && n) throw std::invalid_argument("null buffer"); if (!out_jacobian) throw std::invalid_argument("null output");
        b200::msm_host_buffers(curve, bases_affine, scalars, n, out_jacobian); });
}
int b200_msm_dev(int curve, const void* d_bases_affine, const void* d_scalars, size_t n, void* out_jacobian) {
    return guard([&] { need_device(); if ((!d_bases_affine || !d_scalars) && n) throw std::invalid_argument("null buffer"); if (!out_jacobian) throw std::invalid_argument("null output");
        b200::msm_dev(curve, d_bases_affine, d_scalars, n, out_jacobian); });
}
int b200_point_add(int curve, const void* a, const void* b, void* out) {
    return guard([&] { need_device(); if (!a || !b || !out) throw std::invalid_argument("null argument"); b200::msm_point_add(curve, a, b, out); });
}
int b200_random_points_dev(int curve, void* d_bases_affine, size_t n, uint64_t seed) {
    return guard([&] { need_device(); if (!d_bases_affine && n) throw std::invalid_argument("null buffer"); b200::msm_random_points_dev(curve, d_bases_affine, n, seed); });
}
#define B200_MSM_NAMED(NAME, ID)                                                                                                              \
    int b200_msm_##NAME(const void* bases, const void* scalars, size_t n, void* out) { return b200_msm(ID, bases, scalars, n, out); }          \
    int b200_msm_##NAME##_dev(const void* d_bases, const void* d_scalars, size_t n, void* out) { return b200_msm_dev(ID, d_bases, d_scalars, n, out); }
B200_MSM_NAMED(bn254_g1, B200_CURVE_BN254_G1)
B200_MSM_NAMED(bn254_g2, B200_CURVE_BN254_G2)
B200_MSM_NAMED(bls12381_g1, B200_CURVE_BLS12381_G1)
B200_MSM_NAMED(bls12381_g2, B200_CURVE_BLS12381_G2)
int b200_bn254_g1_add(const void* a96, const void* b96, void* out96) { return b200_point_add(B200_CURVE_BN254_G1, a96, b96, out96); }
int b200_bn254_g1_random_points_dev(void* d_bases_affine, size_t n, uint64_t seed) { return b200_random_points_dev(B200_CURVE_BN254_G1, d_bases_affine, n, seed); }

// ---------------------------------------------------------------------------------------------- groth16 scalar-field domain
int b200_fr_fft_dev(int field, void* d_data, unsigned log_n, int mode) {
    return guard([&] { need_device(); if (!d_data) throw std::invalid_argument("null buffer"); b200::fr_fft_dev(field, d_data, log_n, mode); B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream())); });
}
int b200_fr_fft(int field, void* data, unsigned log_n, int mode) {
    return guard([&] {
        need_device();
        if (!data) throw std::invalid_argument("null buffer");
        if (log_n > 27) throw std::invalid_argument("log size out of range (<= 27)");
        const size_t n = (size_t)1 << log_n;
        DevBuf d(n * 4);
        B200_CUDA_CHECK(cudaMemcpyAsync(d.p, data, n * 32, cudaMemcpyHostToDevice, b200::stream()));
        b200::fr_fft_dev(field, d.p, log_n, mode);
        B200_CUDA_CHECK(cudaMemcpyAsync(data, d.p, n * 32, cudaMemcpyDeviceToHost, b200::stream()));
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}
int b200_groth16_h_dev(int field, void* d_a, void* d_b, void* d_c, unsigned log_m, void* d_h_out) {
    return guard([&] { need_device(); if (!d_a || !d_b || !d_c || !d_h_out) throw std::invalid_argument("null buffer");
        b200::groth16_h_dev(field, d_a, d_b, d_c, log_m, d_h_out); B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream())); });
}
int b200_groth16_h(int field, const void* a, const void* b, const void* c, unsigned log_m, void* h_out) {
    return guard([&] {
        need_device();
        if (!a || !b || !c || !h_out) throw std::invalid_argument("null buffer");
        if (log_m > 27) throw std::invalid_argument("log size out of range (<= 27)");
        const size_t m = (size_t)1 << log_m;
        DevBuf d(4 * m * 4);
        u64 *da = d.p, *db = d.p + 4 * m, *dc = d.p + 8 * m, *dh = d.p + 12 * m;
        B200_CUDA_CHECK(cudaMemcpyAsync(da, a, m * 32, cudaMemcpyHostToDevice, b200::stream()));
        B200_CUDA_CHECK(cudaMemcpyAsync(db, b, m * 32, cudaMemcpyHostToDevice, b200::stream()));
        B200_CUDA_CHECK(cudaMemcpyAsync(dc, c, m * 32, cudaMemcpyHostToDevice, b200::stream()));
        b200::groth16_h_dev(field, da, db, dc, log_m, dh);
        if (m > 1) B200_CUDA_CHECK(cudaMemcpyAsync(h_out, dh, (m - 1) * 32, cudaMemcpyDeviceToHost, b200::stream()));
        B200_CUDA_CHECK(cudaStreamSynchronize(b200::stream()));
    });
}

int b200_fib_trace_dev(uint64_t* d_cm_rowmajor, unsigned log_n) {
    return guard([&] { need_device(); if (log_n > 30) throw std::invalid_argument("log_n too large"); b200::fib_trace(d_cm_rowmajor, (size_t)1 << log_n); });
}

}  // extern "C"
