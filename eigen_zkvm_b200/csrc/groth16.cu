// One call for `Groth16::prove` (groth16/src/groth16.rs:88-96, CLI groth16/src/api.rs:144-203): everything bellman_ce's
// `create_random_proof` does AFTER circuit synthesis -- the quotient H and the five multiexps -- chained on the device over a
// proving key that was uploaded once.
//
//   b200_groth16_pk_read   = `Parameters::read(reader, checked = false)` (api.rs:161,545-550): bellman's `Parameters::write`
//                            layout (vk || h || l || a || b_g1 || b_g2; u32 BE counts; uncompressed big-endian points, G2 as
//                            c1 || c0; byte 0 bit 6 = infinity).  Every point vector becomes a resident MSM table (msm.cu).
//   b200_groth16_prove     = the body of `create_proof` after `circuit.synthesize(&mut prover)`: takes prover.a / b / c (the
//                            per-constraint evaluations, in-memory Montgomery `Fr`), the input / aux assignments (canonical
//                            `Repr`s, what bellman hands to multiexp) and the three density bitmaps its DensityTracker keeps;
//                            r, s are the caller's randomness (`create_random_proof` draws them from its RNG).
//   b200_wtns_read         = `load_witness_from_bin_reader` (algebraic/src/reader.rs:87-138).
// The coefficients of H never leave the device: groth16_h_dev writes them where the `h` multiexp reads them.
// bellman_ce is un-vendored: the algebra is the published Groth16 prover; parity is pinned by oracle/groth16_oracle.py
// (trapdoor-exact proof elements), see tests/test_gpu_groth16.py.
#include "b200_internal.h"
#include "curve.cuh"
#include <cstring>
#include <memory>

namespace b200 {

struct G16Pk {
    int curve = 0;                       // 0 = BN254 (G1 id 0, G2 id 1, Fr id 0), 1 = BLS12-381 (G1 id 2, G2 id 3, Fr id 1)
    size_t fq_bytes = 32;                // bytes of a base-field element on disk
    std::vector<u32> alpha_g1, beta_g1, delta_g1, beta_g2, gamma_g2, delta_g2;     // affine Montgomery words
    std::vector<std::vector<u32>> ic;
    size_t n_h = 0, n_l = 0, n_a = 0, n_b1 = 0, n_b2 = 0;
    MsmTable *t_h = nullptr, *t_l = nullptr, *t_a = nullptr, *t_b1 = nullptr, *t_b2 = nullptr;
    ~G16Pk() { for (MsmTable* t : {t_h, t_l, t_a, t_b1, t_b2}) msm_table_free(t); }
};
static int g1_id(int curve) { return curve == 0 ? 0 : 2; }
static int g2_id(int curve) { return curve == 0 ? 1 : 3; }

// big-endian base-field element -> Montgomery limbs (host; mont.cuh emulates the carry chains on the CPU)
template <class P> static void fq_from_be(const unsigned char* be, size_t nbytes, u32* out) {
    Fp<P> x = Fp<P>::zero();
    for (size_t i = 0; i < nbytes; i++) { size_t bit = 8 * (nbytes - 1 - i); if (bit / 32 < (size_t)P::N) x.l[bit / 32] |= (u32)be[i] << (bit % 32); }
    // must be canonical
    bool lt = false;
    for (int i = P::N - 1; i >= 0; i--) { if (x.l[i] < P::mod(i)) { lt = true; break; } if (x.l[i] > P::mod(i)) break; }
    if (!lt) throw std::invalid_argument("proving key: coordinate is not below the field modulus");
    Fp<P> m = x.to_mont();
    for (int i = 0; i < P::N; i++) out[i] = m.l[i];
}
// one uncompressed point -> affine Montgomery words (all-zero = infinity).  g2: x.c1 || x.c0 || y.c1 || y.c0 on disk, c0 || c1 in memory
template <class P> static void point_from_be(const unsigned char* p, size_t fq_bytes, bool g2, u32* out) {
    const size_t coords = g2 ? 4 : 2, words = P::N;
    if (p[0] & 0x80) throw std::invalid_argument("proving key: compressed point where an uncompressed one is expected");
    if (p[0] & 0x40) { memset(out, 0, coords * words * 4); return; }
    if (!g2) { fq_from_be<P>(p, fq_bytes, out); fq_from_be<P>(p + fq_bytes, fq_bytes, out + words); }
    else {
        fq_from_be<P>(p, fq_bytes, out + words);                  // x.c1
        fq_from_be<P>(p + fq_bytes, fq_bytes, out);               // x.c0
        fq_from_be<P>(p + 2 * fq_bytes, fq_bytes, out + 3 * words);
        fq_from_be<P>(p + 3 * fq_bytes, fq_bytes, out + 2 * words);
    }
}
struct Cursor {
    const unsigned char* d; size_t len, off = 0;
    const unsigned char* take(size_t n) { if (off + n > len) throw std::invalid_argument("proving key: unexpected end of data"); const unsigned char* p = d + off; off += n; return p; }
    u32 be32() { const unsigned char* p = take(4); return ((u32)p[0] << 24) | ((u32)p[1] << 16) | ((u32)p[2] << 8) | p[3]; }
};
template <class P> static std::vector<u32> read_point(Cursor& c, size_t fq_bytes, bool g2) {
    std::vector<u32> w((g2 ? 4 : 2) * P::N);
    point_from_be<P>(c.take((g2 ? 4 : 2) * fq_bytes), fq_bytes, g2, w.data());
    return w;
}
// a length-prefixed vector of points -> resident MSM table (nullptr when empty)
template <class P> static MsmTable* read_vector(Cursor& c, size_t fq_bytes, bool g2, int curve_id, size_t& n_out) {
    const size_t n = c.be32(); n_out = n;
    if (n == 0) return nullptr;
    const size_t pw = (g2 ? 4 : 2) * P::N;
    std::vector<u32> host(n * pw);
    const unsigned char* src = c.take(n * (g2 ? 4 : 2) * fq_bytes);
    for (size_t i = 0; i < n; i++) point_from_be<P>(src + i * (g2 ? 4 : 2) * fq_bytes, fq_bytes, g2, host.data() + i * pw);
    void* d = nullptr;
    B200_CUDA_CHECK(cudaMalloc(&d, host.size() * 4));
    MsmTable* t = nullptr;
    try {
        B200_CUDA_CHECK(cudaMemcpy(d, host.data(), host.size() * 4, cudaMemcpyHostToDevice));
        t = msm_table_new(curve_id, d, n);
    } catch (...) { cudaFree(d); throw; }
    cudaFree(d);
    return t;
}
template <class P> static G16Pk* pk_read_t(int curve, const unsigned char* data, size_t len) {
    std::unique_ptr<G16Pk> pk(new G16Pk());
    pk->curve = curve; pk->fq_bytes = P::N * 4;
    Cursor c{data, len};
    const size_t fb = pk->fq_bytes;
    pk->alpha_g1 = read_point<P>(c, fb, false); pk->beta_g1 = read_point<P>(c, fb, false); pk->beta_g2 = read_point<P>(c, fb, true);
    pk->gamma_g2 = read_point<P>(c, fb, true); pk->delta_g1 = read_point<P>(c, fb, false); pk->delta_g2 = read_point<P>(c, fb, true);
    const size_t n_ic = c.be32();
    for (size_t i = 0; i < n_ic; i++) pk->ic.push_back(read_point<P>(c, fb, false));
    pk->t_h = read_vector<P>(c, fb, false, g1_id(curve), pk->n_h);
    pk->t_l = read_vector<P>(c, fb, false, g1_id(curve), pk->n_l);
    pk->t_a = read_vector<P>(c, fb, false, g1_id(curve), pk->n_a);
    pk->t_b1 = read_vector<P>(c, fb, false, g1_id(curve), pk->n_b1);
    pk->t_b2 = read_vector<P>(c, fb, true, g2_id(curve), pk->n_b2);
    if (c.off != len) throw std::invalid_argument("proving key: trailing bytes");
    if (pk->n_b1 != pk->n_b2) throw std::invalid_argument("proving key: b_g1 and b_g2 differ in length");
    return pk.release();
}
G16Pk* groth16_pk_read(int curve, const void* data, size_t len) {
    if (curve == 0) return pk_read_t<Bn254Fq>(0, (const unsigned char*)data, len);
    if (curve == 1) return pk_read_t<Bls381Fq>(1, (const unsigned char*)data, len);
    throw std::invalid_argument("unknown curve (0 = BN128, 1 = BLS12381)");
}
void groth16_pk_free(G16Pk* pk) { delete pk; }
void groth16_pk_info(const G16Pk* pk, size_t out[6]) { out[0] = pk->n_h; out[1] = pk->n_l; out[2] = pk->n_a; out[3] = pk->n_b1; out[4] = pk->n_b2; out[5] = pk->ic.size(); }

// scalars selected by a density bitmap (one byte per variable), appended to `dst`
static void append_dense(std::vector<u64>& dst, const u64* src, size_t n, const unsigned char* density) {
    for (size_t i = 0; i < n; i++) if (!density || density[i]) dst.insert(dst.end(), src + 4 * i, src + 4 * i + 4);
}
static std::vector<u32> jac_affine(const std::vector<u32>& jac, size_t coord_words) {      // normalised triple -> affine words (all-zero = infinity)
    std::vector<u32> a(2 * coord_words, 0);
    bool inf = true; for (size_t i = 0; i < coord_words; i++) inf &= jac[2 * coord_words + i] == 0;
    if (!inf) memcpy(a.data(), jac.data(), 2 * coord_words * 4);
    return a;
}
// sum_k scalars[k] * points[k] for a handful of host points (the linear combinations around the big multiexps)
static std::vector<u32> small_msm(int curve_id, const std::vector<const std::vector<u32>*>& pts, const std::vector<const u64*>& sc) {
    const size_t pw = msm_point_bytes(curve_id) / 4;
    std::vector<u32> bases(pts.size() * pw); std::vector<u64> scal(pts.size() * 4);
    for (size_t k = 0; k < pts.size(); k++) { memcpy(bases.data() + k * pw, pts[k]->data(), pw * 4); memcpy(scal.data() + 4 * k, sc[k], 32); }
    std::vector<u32> out(pw / 2 * 3);
    msm_host_buffers(curve_id, bases.data(), scal.data(), pts.size(), out.data());
    return out;
}
// r * s mod the scalar field, canonical 4 x u64 (host, through the Montgomery code)
template <class P> static void fr_mul_canon(const u64 a[4], const u64 b[4], u64 out[4]) {
    Fp<P> x, y;
    for (int i = 0; i < 4; i++) { x.l[2 * i] = (u32)a[i]; x.l[2 * i + 1] = (u32)(a[i] >> 32); y.l[2 * i] = (u32)b[i]; y.l[2 * i + 1] = (u32)(b[i] >> 32); }
    Fp<P> z = (x.to_mont() * y.to_mont()).from_mont();
    for (int i = 0; i < 4; i++) out[i] = (u64)z.l[2 * i] | ((u64)z.l[2 * i + 1] << 32);
}

// proof_out: A (G1 affine) || B (G2 affine) || C (G1 affine), Montgomery words
void groth16_prove(const G16Pk* pk, const void* a_evals, const void* b_evals, const void* c_evals, size_t n_constraints,
                   const u64* input_assignment, size_t n_inputs, const u64* aux_assignment, size_t n_aux,
                   const unsigned char* a_aux_density, const unsigned char* b_input_density, const unsigned char* b_aux_density,
                   const u64 r[4], const u64 s[4], void* proof_out) {
    const int G1 = g1_id(pk->curve), G2 = g2_id(pk->curve);
    const size_t w1 = msm_point_bytes(G1) / 8, w2 = msm_point_bytes(G2) / 8;              // coordinate words (u32) of G1 / G2 = pw / 2
    if (n_constraints == 0) throw std::invalid_argument("no constraints");
    unsigned log_m = 0; while (((size_t)1 << log_m) < n_constraints) log_m++;
    const size_t m = (size_t)1 << log_m;
    if (m - 1 != pk->n_h) throw std::invalid_argument("proving key: h has " + std::to_string(pk->n_h) + " points, the domain needs " + std::to_string(m - 1));
    if (n_aux != pk->n_l) throw std::invalid_argument("proving key: l does not match the number of aux variables");
    if (n_inputs != pk->ic.size()) throw std::invalid_argument("proving key: ic does not match the number of inputs");
    cudaStream_t st = stream();
    // ---- H on the device: a, b, c padded to the domain, quotient, and straight into the h multiexp
    static char* g_buf[16] = {nullptr}; static size_t g_cap[16] = {0};
    int dev = current_device();
    const size_t need = 4 * m * 32;
    if (g_cap[dev] < need) { if (g_buf[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(st)); B200_CUDA_CHECK(cudaFree(g_buf[dev])); } B200_CUDA_CHECK(cudaMalloc(&g_buf[dev], need)); g_cap[dev] = need; }
    char* d_a = g_buf[dev]; char* d_b = d_a + m * 32; char* d_c = d_b + m * 32; char* d_h = d_c + m * 32;
    B200_CUDA_CHECK(cudaMemsetAsync(d_a, 0, 3 * m * 32, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(d_a, a_evals, n_constraints * 32, cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(d_b, b_evals, n_constraints * 32, cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(d_c, c_evals, n_constraints * 32, cudaMemcpyHostToDevice, st));
    groth16_h_dev(pk->curve, d_a, d_b, d_c, log_m, d_h);
    std::vector<u32> h_pt(3 * w1, 0), l_pt(3 * w1, 0), a_pt(3 * w1, 0), b1_pt(3 * w1, 0), b2_pt(3 * w2, 0);
    auto set_inf = [&](std::vector<u32>& p, size_t cw, int cid) { (void)cid; std::fill(p.begin(), p.end(), 0u); std::vector<u32> one(3 * cw); msm_dev(cid, nullptr, nullptr, 0, one.data()); p = one; };
    if (pk->t_h) msm_table_run(pk->t_h, d_h, h_pt.data()); else set_inf(h_pt, w1, G1);
    // ---- the four assignment multiexps
    if (pk->t_l) msm_table_run_host(pk->t_l, aux_assignment, l_pt.data()); else set_inf(l_pt, w1, G1);
    std::vector<u64> sc;
    sc.reserve(4 * (n_inputs + n_aux));
    append_dense(sc, input_assignment, n_inputs, nullptr); append_dense(sc, aux_assignment, n_aux, a_aux_density);
    if (sc.size() / 4 != pk->n_a) throw std::invalid_argument("proving key: a has " + std::to_string(pk->n_a) + " points, the densities select " + std::to_string(sc.size() / 4));
    if (pk->t_a) msm_table_run_host(pk->t_a, sc.data(), a_pt.data()); else set_inf(a_pt, w1, G1);
    sc.clear();
    append_dense(sc, input_assignment, n_inputs, b_input_density); append_dense(sc, aux_assignment, n_aux, b_aux_density);
    if (sc.size() / 4 != pk->n_b1) throw std::invalid_argument("proving key: b has " + std::to_string(pk->n_b1) + " points, the densities select " + std::to_string(sc.size() / 4));
    if (pk->t_b1) { msm_table_run_host(pk->t_b1, sc.data(), b1_pt.data()); msm_table_run_host(pk->t_b2, sc.data(), b2_pt.data()); }
    else { set_inf(b1_pt, w1, G1); set_inf(b2_pt, w2, G2); }
    // ---- g_a = r delta + alpha + A;  g_b = s delta_2 + beta_2 + B_2;  g_c = rs delta + s alpha + r beta + s A + r B_1 + H + L
    u64 rs[4];
    if (pk->curve == 0) fr_mul_canon<Bn254Fr>(r, s, rs); else fr_mul_canon<Bls381Fr>(r, s, rs);
    const u64 one[4] = {1, 0, 0, 0};
    std::vector<u32> A = jac_affine(a_pt, w1), B1 = jac_affine(b1_pt, w1), B2 = jac_affine(b2_pt, w2), H = jac_affine(h_pt, w1), Lp = jac_affine(l_pt, w1);
    std::vector<u32> ga = small_msm(G1, {&pk->delta_g1, &pk->alpha_g1, &A}, {r, one, one});
    std::vector<u32> gb = small_msm(G2, {&pk->delta_g2, &pk->beta_g2, &B2}, {s, one, one});
    std::vector<u32> gc = small_msm(G1, {&pk->delta_g1, &pk->alpha_g1, &pk->beta_g1, &A, &B1, &H, &Lp}, {rs, s, r, s, r, one, one});
    std::vector<u32> oa = jac_affine(ga, w1), ob = jac_affine(gb, w2), oc = jac_affine(gc, w1);
    char* o = (char*)proof_out;
    memcpy(o, oa.data(), oa.size() * 4); memcpy(o + oa.size() * 4, ob.data(), ob.size() * 4); memcpy(o + oa.size() * 4 + ob.size() * 4, oc.data(), oc.size() * 4);
}

// `.wtns` (iden3, version <= 2): returns the number of witness values; out (if not null) receives n x 4 canonical u64
size_t wtns_read(const void* data, size_t len, const unsigned char prime_le32[32], u64* out, size_t out_cap) {
    const unsigned char* p = (const unsigned char*)data;
    auto rd32 = [&](size_t o) { if (o + 4 > len) throw std::invalid_argument("wtns: truncated"); return (u32)p[o] | ((u32)p[o + 1] << 8) | ((u32)p[o + 2] << 16) | ((u32)p[o + 3] << 24); };
    auto rd64 = [&](size_t o) { return (u64)rd32(o) | ((u64)rd32(o + 4) << 32); };
    if (len < 76 || memcmp(p, "wtns", 4)) throw std::invalid_argument("wtns: invalid file header");
    if (rd32(4) > 2) throw std::invalid_argument("wtns: unsupported file version");
    if (rd32(8) != 2) throw std::invalid_argument("wtns: invalid num sections");
    if (rd32(12) != 1) throw std::invalid_argument("wtns: invalid section type");
    if (rd64(16) != 4 + 32 + 4) throw std::invalid_argument("wtns: invalid section len");
    if (rd32(24) != 32) throw std::invalid_argument("wtns: invalid field byte size");
    if (prime_le32 && memcmp(p + 28, prime_le32, 32)) throw std::invalid_argument("wtns: invalid curve prime");
    const size_t n = rd32(60);
    if (rd32(64) != 2) throw std::invalid_argument("wtns: invalid section type");
    if (rd64(68) != (u64)n * 32 || 76 + n * 32 > len) throw std::invalid_argument("wtns: invalid witness section size");
    if (out) {
        if (out_cap < n) throw std::invalid_argument("wtns: output buffer too small");
        memcpy(out, p + 76, n * 32);
    }
    return n;
}

}  // namespace b200
