// Scalar-field (BN254 Fr / BLS12-381 Fr) evaluation domain on the device and the groth16 quotient `H`.
//
// Boundary: bellman_ce's `domain::EvaluationDomain::{fft, ifft, coset_fft, icoset_fft, mul_assign, sub_assign,
// divide_by_z_on_coset}` as used by `groth16::create_random_proof`, reached from `Groth16::prove`
// (groth16/src/groth16.rs:88-96; bellperson twin for BLS12-381 :45-57) -- SURVEY.md 8f rank 2.  These crates are not
// vendored by the reference; the semantics restated in oracle/fr_domain.py are: omega_m = (7^t)^(2^(S - log m)) with
// r - 1 = 2^S t, cosets g <omega> with g = 7, natural order in and out.
// Values are the in-memory `Fr` of those libraries: 4 x u64 little-endian MONTGOMERY limbs (R = 2^256); the `h`
// coefficients come back as canonical `Repr`s, the form the following multiexp consumes.
//
// Kernels: bit-reversal swap, 2^10-point shared-memory blocks for the first 10 stages, then three stages per pass in registers
// (radix 8, twiddles from a table of n/2 powers of omega), fused `x_i *= c g^i` for the coset shifts and 1/m.
// Roofline class: INT (one 256-bit Montgomery product per butterfly, 180 instructions) -- about 2 ms per 2^22-point
// transform against 0.3 ms of HBM time; the seven transforms of a proof are small next to its five multiexps.
#include "b200_internal.h"
#include "curve_params.h"
#include <map>
#include <tuple>
#include <mutex>

namespace b200 {

template <class P> __device__ __forceinline__ Fp<P> fr_load(const Fp<P>* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = q[0], b = q[1];
    Fp<P> r; r.l[0] = a.x; r.l[1] = a.y; r.l[2] = a.z; r.l[3] = a.w; r.l[4] = b.x; r.l[5] = b.y; r.l[6] = b.z; r.l[7] = b.w;
    return r;
}
template <class P> __device__ __forceinline__ void fr_store(Fp<P>* p, const Fp<P>& v) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]); q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}
template <class P> __device__ __forceinline__ Fp<P> fr_pow_u32(Fp<P> base, u32 e) {
    Fp<P> acc = Fp<P>::one();
    while (e) { if (e & 1) acc = acc * base; base = base.sqr(); e >>= 1; }
    return acc;
}

// out[i] = base^i
template <class P> __global__ void k_fr_powers(Fp<P>* __restrict__ out, Fp<P> base, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store<P>(out + i, fr_pow_u32<P>(base, i));
}
// in-place bit reversal
template <class P> __global__ void k_fr_bitrev(Fp<P>* __restrict__ a, u32 log_n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1u << log_n)) return;
    u32 r = __brev(i) >> (32 - log_n);
    if (log_n == 0 || i >= r) return;
    Fp<P> x = fr_load<P>(a + i), y = fr_load<P>(a + r);
    fr_store<P>(a + i, y); fr_store<P>(a + r, x);
}
// stages 0 .. ls-1 (butterfly spans 1 .. 2^(ls-1)) on blocks of 2^ls consecutive elements held in shared memory
#define FR_LOCAL_BITS 10
template <class P> __global__ void __launch_bounds__(256) k_fr_local(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ tw, u32 log_n, u32 ls) {
    __shared__ Fp<P> sh[1 << FR_LOCAL_BITS];
    const u32 bn = 1u << ls, base = blockIdx.x * bn;
    for (u32 t = threadIdx.x; t < bn; t += blockDim.x) sh[t] = fr_load<P>(a + base + t);
    __syncthreads();
    for (u32 s = 0; s < ls; s++) {
        const u32 m = 1u << s;
        for (u32 t = threadIdx.x; t < bn / 2; t += blockDim.x) {
            const u32 j = t & (m - 1), k = ((t >> s) << (s + 1)) + j;
            Fp<P> u = sh[k], v = sh[k + m];
            if (j) v = v * fr_load<P>(tw + ((size_t)j << (log_n - 1 - s)));
            sh[k] = u + v; sh[k + m] = u - v;
        }
        __syncthreads();
    }
    for (u32 t = threadIdx.x; t < bn; t += blockDim.x) fr_store<P>(a + base + t, sh[t]);
}
// K consecutive stages s .. s+K-1 in one pass over the array (radix 2^K, decimation in time): a thread owns the 2^K elements
// k + i * 2^s, runs the K butterfly layers in registers and writes them back -- 2^K - 1 twiddle loads and K * 2^(K-1) products
// per 2^K elements, one read and one write of the array per K stages (the first version made one pass per stage: thirteen passes
// over 128 MB for 2^22 points, 3.5 % of the HBM roofline).  Consecutive threads own consecutive k: every access is coalesced.
template <class P, int K> __global__ void __launch_bounds__(256) k_fr_stages(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ tw, u32 log_n, u32 s) {
    const size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ((size_t)1 << (log_n - K))) return;
    const size_t m = (size_t)1 << s, j = t & (m - 1), k = ((t >> s) << (s + K)) + j;
    Fp<P> x[1 << K];
#pragma unroll
    for (int i = 0; i < (1 << K); i++) x[i] = fr_load<P>(a + k + (size_t)i * m);
#pragma unroll
    for (int q = 0; q < K; q++) {
        const int h = 1 << q;
#pragma unroll
        for (int i = 0; i < (1 << K); i++) {
            if (i & h) continue;
            const size_t e = (j + (size_t)(i & (h - 1)) * m) << (log_n - 1 - s - q);
            Fp<P> v = x[i + h];
            if (e) v = v * fr_load<P>(tw + e);
            const Fp<P> u = x[i];
            x[i] = u + v; x[i + h] = u - v;
        }
    }
#pragma unroll
    for (int i = 0; i < (1 << K); i++) fr_store<P>(a + k + (size_t)i * m, x[i]);
}
// a[i] *= c * g^i   (distribute_powers and the 1/m of the inverse transform, fused).  g^i = lo[i & 2047] * hi[i >> 11] from
// two small tables built per call (c is folded into `hi`): 2 loads + 2 products per element instead of a 30-product pow.
#define FR_POW_LO_BITS 11
template <class P> __global__ void k_fr_pow_tables(Fp<P>* __restrict__ lo, Fp<P>* __restrict__ hi, Fp<P> g, Fp<P> c, u32 n_lo, u32 n_hi) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lo) fr_store<P>(lo + i, fr_pow_u32<P>(g, i));
    else if (i < n_lo + n_hi) fr_store<P>(hi + (i - n_lo), c * fr_pow_u32<P>(g, (i - n_lo) << FR_POW_LO_BITS));
}
template <class P> __global__ void __launch_bounds__(256) k_fr_scale_powers(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ lo, const Fp<P>* __restrict__ hi, Fp<P> c, u32 n, int use_g) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Fp<P> f = use_g ? fr_load<P>(lo + (i & ((1u << FR_POW_LO_BITS) - 1))) * fr_load<P>(hi + (i >> FR_POW_LO_BITS)) : c;
    fr_store<P>(a + i, fr_load<P>(a + i) * f);
}
// a = (a * b - c) * zinv     (mul_assign, sub_assign, divide_by_z_on_coset)
template <class P> __global__ void __launch_bounds__(256) k_fr_quotient(Fp<P>* __restrict__ a, const Fp<P>* __restrict__ b, const Fp<P>* __restrict__ c, Fp<P> zinv, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fr_store<P>(a + i, (fr_load<P>(a + i) * fr_load<P>(b + i) - fr_load<P>(c + i)) * zinv);
}
template <class P> __global__ void __launch_bounds__(256) k_fr_from_mont(const Fp<P>* __restrict__ a, Fp<P>* __restrict__ out, u32 n) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) fr_store<P>(out + i, fr_load<P>(a + i).from_mont());
}

// ------------------------------------------------------------------------------------------------ host
// small host-side Montgomery helpers (the same mont.cuh code, compiled for the host)
template <class P> static Fp<P> h_from_u64(u64 v) { Fp<P> x = Fp<P>::zero(); x.l[0] = (u32)v; x.l[1] = (u32)(v >> 32); return x.to_mont(); }
template <class P> static Fp<P> h_pow(Fp<P> b, const u32* e, int n_limbs) { return b.pow(e, n_limbs); }
template <class P> struct FrConsts { Fp<P> g, ginv, root; unsigned s; };
template <class P> static FrConsts<P> fr_consts() {
    FrConsts<P> c;
    c.g = h_from_u64<P>(7); c.ginv = c.g.inv();
    // r - 1 = 2^S t
    u32 e[P::N]; for (int i = 0; i < P::N; i++) e[i] = P::mod(i);
    e[0] -= 1;                                   // r is odd
    unsigned s = 0;
    while (!(e[0] & 1)) { for (int i = 0; i < P::N; i++) e[i] = (e[i] >> 1) | (i + 1 < P::N ? e[i + 1] << 31 : 0); s++; }
    c.s = s; c.root = c.g.pow(e, P::N);          // ROOT_OF_UNITY = g^t
    return c;
}
template <class P> static Fp<P> fr_omega(unsigned log_m, bool inverse) {
    static FrConsts<P> C = fr_consts<P>();
    if (log_m > C.s) throw std::invalid_argument("PolynomialDegreeTooLarge");
    Fp<P> w = C.root;
    for (unsigned i = log_m; i < C.s; i++) w = w.sqr();
    return inverse ? w.inv() : w;
}
template <class P> static const Fp<P>* fr_twiddles(unsigned log_n, bool inverse) {
    static std::map<std::tuple<int, unsigned, bool>, const Fp<P>*> cache; static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = current_device();
    auto key = std::make_tuple(dev, log_n, inverse);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    u32 n = log_n ? 1u << (log_n - 1) : 1;
    Fp<P>* p; B200_CUDA_CHECK(cudaMalloc(&p, (size_t)n * sizeof(Fp<P>)));
    k_fr_powers<P><<<(n + 255) / 256, 256, 0, stream()>>>(p, fr_omega<P>(log_n, inverse), n); launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
    cache[key] = p;
    return p;
}
// a[i] *= c * g^i on the device (use_g = false: plain scaling by c)
template <class P> static void fr_scale(Fp<P>* d, u32 n, const Fp<P>& g, const Fp<P>& c, bool use_g) {
    static Fp<P>* tab[16] = {nullptr};           // per device: 2^11 + 2^16 entries, rebuilt per call (a 70 k-thread kernel)
    int dev = current_device();
    const u32 n_lo = 1u << FR_POW_LO_BITS, n_hi_max = 1u << 16;
    if (!tab[dev]) B200_CUDA_CHECK(cudaMalloc(&tab[dev], (size_t)(n_lo + n_hi_max) * sizeof(Fp<P>)));
    Fp<P>* lo = tab[dev]; Fp<P>* hi = lo + n_lo;
    if (use_g) {
        const u32 n_hi = (n + n_lo - 1) >> FR_POW_LO_BITS;
        if (n_hi > n_hi_max) throw std::invalid_argument("log size out of range (<= 27)");
        k_fr_pow_tables<P><<<(n_lo + n_hi + 255) / 256, 256, 0, stream()>>>(lo, hi, g, c, n_lo, n_hi); launch_count_add(1);
    }
    k_fr_scale_powers<P><<<(n + 255) / 256, 256, 0, stream()>>>(d, lo, hi, c, n, use_g ? 1 : 0); launch_count_add(1);
}
// mode: 0 fft, 1 ifft, 2 coset_fft, 3 icoset_fft; in place on n = 2^log_n Montgomery elements
template <class P> static void fr_fft_dev_t(Fp<P>* d, unsigned log_n, int mode) {
    if (log_n > 27) throw std::invalid_argument("log size out of range (<= 27)");
    static FrConsts<P> C = fr_consts<P>();
    const u32 n = 1u << log_n;
    const bool inverse = (mode == 1 || mode == 3);
    cudaStream_t st = stream();
    ScopedTimer tm("fr_ntt", 64.0 * n);
    if (mode == 2) fr_scale<P>(d, n, C.g, Fp<P>::one(), true);
    if (log_n > 0) {
        const Fp<P>* tw = fr_twiddles<P>(log_n, inverse);
        k_fr_bitrev<P><<<(n + 255) / 256, 256, 0, st>>>(d, log_n);
        const unsigned ls = log_n < FR_LOCAL_BITS ? log_n : FR_LOCAL_BITS;
        k_fr_local<P><<<n >> ls, 256, 0, st>>>(d, tw, log_n, ls);
        unsigned s = ls, launches = 2;
        while (s < log_n) {
            const unsigned rem = log_n - s;
            if (rem >= 3) { k_fr_stages<P, 3><<<((n >> 3) + 255) / 256, 256, 0, st>>>(d, tw, log_n, s); s += 3; }
            else if (rem == 2) { k_fr_stages<P, 2><<<((n >> 2) + 255) / 256, 256, 0, st>>>(d, tw, log_n, s); s += 2; }
            else { k_fr_stages<P, 1><<<((n >> 1) + 255) / 256, 256, 0, st>>>(d, tw, log_n, s); s += 1; }
            launches++;
        }
        launch_count_add(launches);
    }
    if (inverse) {
        Fp<P> minv = h_from_u64<P>(n).inv();
        fr_scale<P>(d, n, C.ginv, minv, mode == 3);
    }
    B200_CUDA_CHECK(cudaGetLastError());
}
// a, b, c: m Montgomery elements each on the device (a is overwritten); h_out: m - 1 canonical elements (device)
template <class P> static void groth16_h_dev_t(Fp<P>* a, Fp<P>* b, Fp<P>* c, unsigned log_m, Fp<P>* h_out) {
    static FrConsts<P> C = fr_consts<P>();
    const u32 m = 1u << log_m;
    for (Fp<P>* x : {a, b, c}) { fr_fft_dev_t<P>(x, log_m, 1); fr_fft_dev_t<P>(x, log_m, 2); }
    // Z(g w^i) = g^m - 1
    Fp<P> gm = C.g; for (unsigned i = 0; i < log_m; i++) gm = gm.sqr();
    Fp<P> zinv = (gm - Fp<P>::one()).inv();
    k_fr_quotient<P><<<(m + 255) / 256, 256, 0, stream()>>>(a, b, c, zinv, m); launch_count_add(1);
    fr_fft_dev_t<P>(a, log_m, 3);
    if (m > 1) { k_fr_from_mont<P><<<(m - 1 + 255) / 256, 256, 0, stream()>>>(a, h_out, m - 1); launch_count_add(1); }
    B200_CUDA_CHECK(cudaGetLastError());
}

// field ids of the C-ABI: 0 = BN254 Fr, 1 = BLS12-381 Fr
void fr_fft_dev(int field, void* d_data, unsigned log_n, int mode) {
    if (mode < 0 || mode > 3) throw std::invalid_argument("unknown transform mode");
    if (field == 0) fr_fft_dev_t<Bn254Fr>((Fp<Bn254Fr>*)d_data, log_n, mode);
    else if (field == 1) fr_fft_dev_t<Bls381Fr>((Fp<Bls381Fr>*)d_data, log_n, mode);
    else throw std::invalid_argument("unknown scalar field id");
}
void groth16_h_dev(int field, void* d_a, void* d_b, void* d_c, unsigned log_m, void* d_h_out) {
    if (log_m > 27) throw std::invalid_argument("log size out of range (<= 27)");
    if (field == 0) groth16_h_dev_t<Bn254Fr>((Fp<Bn254Fr>*)d_a, (Fp<Bn254Fr>*)d_b, (Fp<Bn254Fr>*)d_c, log_m, (Fp<Bn254Fr>*)d_h_out);
    else if (field == 1) groth16_h_dev_t<Bls381Fr>((Fp<Bls381Fr>*)d_a, (Fp<Bls381Fr>*)d_b, (Fp<Bls381Fr>*)d_c, log_m, (Fp<Bls381Fr>*)d_h_out);
    else throw std::invalid_argument("unknown scalar field id");
}

}  // namespace b200
