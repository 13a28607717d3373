// BN128 / BLS12-381 Poseidon commitment back-ends: the Merkle stack of the LAST stark before the snark
// (`verificationHashType` "BN128" / "BLS12381", starky/src/prove.rs:52-89).
// Reference: Poseidon (variable width t = inputs + 1 <= 17, x^5, optimised rounds) starky/src/poseidon_bn128_opt.rs:94-225,
// poseidon_bls12381_opt.rs:95-231; leaves starky/src/linearhash_bn128.rs:105-131 (3 GL elements per scalar, 16 scalars
// per absorption, last block unpadded); 16-ary tree starky/src/merklehash_bn128.rs:26-87,176-224 (BLS twins alike).
//
// Layout: digests are field elements, 4 x u64 little-endian CANONICAL at rest (the reference's Montgomery limbs,
// digest.rs:45-65, are its internal form); arithmetic is mont.cuh's 8 x 32-bit Montgomery.  One permutation per
// thread with the (up to 17-lane) state in thread-local memory; round constants live in global memory and are read
// at warp-uniform addresses (one broadcast transaction per constant).
// Roofline class: INT (IMAD.WIDE issue): a t = 17 permutation is about 5.6k 256-bit products against 544 B of traffic.
#include "b200_internal.h"
#include "curve_params.h"
#include <cstring>
#include <cstdio>
#include <dlfcn.h>
#include <mutex>

namespace b200 {

template <class P> struct PosTab {      // device pointers, Montgomery form
    const Fp<P>* C[18]; const Fp<P>* S[18]; const Fp<P>* M[18]; const Fp<P>* Pm[18];
    u32 rp[18];
};

// x^5
template <class P> __device__ __forceinline__ Fp<P> pow5(const Fp<P>& x) { Fp<P> x2 = x.sqr(); Fp<P> x4 = x2.sqr(); return x4 * x; }

// Dot products of up to 17 terms are accumulated UNREDUCED (FpWide, mont.cuh) and reduced once: a term costs the 64 limb
// products of a * b instead of 128 (product + Montgomery reduction) -- the linear layers are 85 % of a t = 17 permutation's products.
// LOGK: the reduced value is < p (terms p / R + 2) with p / R < (top limb + 1) / 2^32  (3 for BN254's r, 4 for BLS12-381's)
template <class P> __host__ __device__ constexpr int dot_logk(int terms) {
    double b = terms * ((double)P::mod(P::N - 1) + 1.0) / 4294967296.0 + 2.0;
    int k = 1; while ((double)(1 << k) < b) k++;
    return k;
}
// st' = Mat^T st : st'[i] = sum_j Mat[j][i] st[j]
template <class P> __device__ __noinline__ void pos_mix(Fp<P>* st, Fp<P>* tmp, const Fp<P>* __restrict__ mat, int t) {
    for (int i = 0; i < t; i++) {
        FpWide<P> acc; acc.clear();
#pragma unroll 1
        for (int j = 0; j < t; j++) acc.mad(mat[j * t + i], st[j]);
        tmp[i] = acc.template reduce<dot_logk<P>(17)>();
    }
    for (int i = 0; i < t; i++) st[i] = tmp[i];
}
// sum_{j < t} row[j] * st[j]
template <class P> __device__ __noinline__ Fp<P> pos_dot(const Fp<P>* __restrict__ row, const Fp<P>* st, int t) {
    FpWide<P> acc; acc.clear();
#pragma unroll 1
    for (int j = 0; j < t; j++) acc.mad(row[j], st[j]);
    return acc.template reduce<dot_logk<P>(17)>();
}
// poseidon_bn128_opt.rs:111-225 (hash_inner): st = [init, inputs...] in Montgomery form, permuted in place
template <class P> __device__ __noinline__ void poseidon_big(Fp<P>* st, Fp<P>* tmp, int t, const PosTab<P>& T) {
    const Fp<P>* C = T.C[t]; const Fp<P>* S = T.S[t]; const Fp<P>* M = T.M[t]; const Fp<P>* Pm = T.Pm[t];
    const int rp = (int)T.rp[t];
    for (int i = 0; i < t; i++) st[i] = st[i] + C[i];
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < t; i++) st[i] = pow5(st[i]) + C[(r + 1) * t + i];
        pos_mix<P>(st, tmp, M, t);
    }
    for (int i = 0; i < t; i++) st[i] = pow5(st[i]) + C[4 * t + i];
    pos_mix<P>(st, tmp, Pm, t);
    for (int r = 0; r < rp; r++) {
        const Fp<P>* Sr = S + (size_t)(2 * t - 1) * r;
        Fp<P> x0 = pow5(st[0]) + C[5 * t + r];
        st[0] = x0;
        Fp<P> s0 = pos_dot<P>(Sr, st, t);
        for (int k = 1; k < t; k++) st[k] = st[k] + Sr[t + k - 1] * x0;
        st[0] = s0;
    }
    for (int r = 0; r < 3; r++) {
        for (int i = 0; i < t; i++) st[i] = pow5(st[i]) + C[5 * t + rp + r * t + i];
        pos_mix<P>(st, tmp, M, t);
    }
    for (int i = 0; i < t; i++) st[i] = pow5(st[i]);
    pos_mix<P>(st, tmp, M, t);
}

// ---- warp-resident permutation: lane i holds state element i (t <= 17 lanes active) ------------------------------------------
// The thread-per-permutation form above is the throughput form (32 permutations per warp), but ONE t = 17 permutation is about
// 10^6 dependent instructions on one thread -- 2 ms -- and the top levels of every 16-ary tree, the FRI layer trees and the
// transcript are chains of a few such permutations.  Here the state lives across the lanes of a warp: S-boxes of a full round run
// in parallel, a mix is t shuffled elements per lane (each lane owns one output row), a partial round is one S-box on lane 0, one
// product per lane and a 5-step shuffle reduction.  About 520 dependent products per permutation instead of 5457: 25 x lower
// latency, at 3.6 x the instruction cost -- used below a few thousand permutations per launch.
template <class P> __device__ __forceinline__ Fp<P> shfl_fp(const Fp<P>& x, int src) {
    Fp<P> r;
#pragma unroll
    for (int k = 0; k < P::N; k++) r.l[k] = __shfl_sync(0xffffffffu, x.l[k], src);
    return r;
}
template <class P> __device__ __forceinline__ Fp<P> shfl_down_fp(const Fp<P>& x, int off) {
    Fp<P> r;
#pragma unroll
    for (int k = 0; k < P::N; k++) r.l[k] = __shfl_down_sync(0xffffffffu, x.l[k], off);
    return r;
}
template <class P> __device__ __noinline__ void warp_mix(Fp<P>& s, const Fp<P>* __restrict__ mat, int t, int lane) {
    FpWide<P> acc; acc.clear();
    const int col = lane < t ? lane : 0;             // idle lanes compute a copy of column 0 (no divergence inside the product)
#pragma unroll 1
    for (int j = 0; j < t; j++) {
        Fp<P> v = shfl_fp<P>(s, j);
        acc.mad(mat[j * t + col], v);
    }
    s = acc.template reduce<dot_logk<P>(17)>();
}
// all 32 lanes must call; s = this lane's state element (Montgomery; lanes >= t: ignored), permuted in place
template <class P> __device__ __noinline__ void poseidon_big_warp(Fp<P>& s, int t, const PosTab<P>& T) {
    const int lane = threadIdx.x & 31;
    const bool act = lane < t;
    const Fp<P>* C = T.C[t]; const Fp<P>* S = T.S[t]; const Fp<P>* M = T.M[t]; const Fp<P>* Pm = T.Pm[t];
    const int rp = (int)T.rp[t];
    if (!act) s = Fp<P>::zero();
    if (act) s = s + C[lane];
    for (int r = 0; r < 3; r++) {
        if (act) s = pow5(s) + C[(r + 1) * t + lane];
        warp_mix<P>(s, M, t, lane);
    }
    if (act) s = pow5(s) + C[4 * t + lane];
    warp_mix<P>(s, Pm, t, lane);
    for (int r = 0; r < rp; r++) {
        const Fp<P>* Sr = S + (size_t)(2 * t - 1) * r;
        if (lane == 0) s = pow5(s) + C[5 * t + r];
        Fp<P> x0 = shfl_fp<P>(s, 0);
        Fp<P> term = act ? Sr[lane] * s : Fp<P>::zero();
        for (int off = 16; off > 0; off >>= 1) { Fp<P> o = shfl_down_fp<P>(term, off); term = term + o; }
        if (act && lane >= 1) s = s + Sr[t + lane - 1] * x0;
        if (lane == 0) s = term;
    }
    for (int r = 0; r < 3; r++) {
        if (act) s = pow5(s) + C[5 * t + rp + r * t + lane];
        warp_mix<P>(s, M, t, lane);
    }
    if (act) s = pow5(s);
    warp_mix<P>(s, M, t, lane);
}

template <class P> __device__ __forceinline__ Fp<P> load_canon(const u64* p4) {      // canonical 4 x u64 -> Montgomery
    Fp<P> x;
#pragma unroll
    for (int i = 0; i < 4; i++) { u64 v = p4[i]; x.l[2 * i] = (u32)v; x.l[2 * i + 1] = (u32)(v >> 32); }
    return x.to_mont();
}
template <class P> __device__ __forceinline__ void store_canon(u64* p4, const Fp<P>& m) {
    Fp<P> x = m.from_mont();
#pragma unroll
    for (int i = 0; i < 4; i++) p4[i] = (u64)x.l[2 * i] | ((u64)x.l[2 * i + 1] << 32);
}
// any 256-bit integer -> canonical residue (at most 5 subtractions: 2^256 < 6 r for both fields)
template <class P> __device__ __forceinline__ Fp<P> reduce_raw(Fp<P> x) {
    for (int k = 0; k < 6; k++) {
        Fp<P> t; t.l[0] = mp_sub_cc(x.l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < 8; i++) t.l[i] = mp_subc_cc(x.l[i], P::mod(i));
        u32 bw = mp_subc(0, 0);
        if (bw) break;
        x = t;
    }
    return x;
}

// constants conversion at load time
template <class P> __global__ void k_big_to_mont(const u64* __restrict__ in, Fp<P>* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = load_canon<P>(in + 4 * i);
}
// one permutation on one warp: in = init, inputs (canonical); out = full state (canonical)
template <class P> __global__ void __launch_bounds__(32) k_big_poseidon_single(const u64* __restrict__ in, u64* __restrict__ out, int t, PosTab<P> T) {
    const int lane = threadIdx.x;
    Fp<P> s = lane < t ? load_canon<P>(in + 4 * lane) : Fp<P>::zero();
    poseidon_big_warp<P>(s, t, T);
    if (lane < t) store_canon<P>(out + 4 * lane, s);
}
// one level, one WARP per parent (small levels): parent[i] = Poseidon_17(children[16 i .. 16 i + 16), init 0)
template <class P, int LANE> __global__ void __launch_bounds__(128) k_big_level_warp(const u64* __restrict__ in, u64* __restrict__ out, size_t n_ops, PosTab<P> T) {
    const size_t i = (size_t)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n_ops) return;                      // warp-uniform
    Fp<P> s = (lane >= 1 && lane <= 16) ? load_canon<P>(in + 4 * (16 * i + lane - 1)) : Fp<P>::zero();
    poseidon_big_warp<P>(s, 17, T);
    if (lane == LANE) store_canon<P>(out + 4 * i, s);
}
// leaf digests: linearhash_bn128.rs:105-131 (`hash_element_array`) on column-major GL data
__device__ __forceinline__ u64 big_col_load(const ColView& v, u32 c, size_t row) {       // same view as merkle.cu's col_load
    u64 off = (u64)(c / v.a) * v.s1 + (u64)(c % v.a) * v.s2;
    return __ldg(v.base + off + row);
}
template <class P, int LANE> __global__ void __launch_bounds__(128) k_big_leaves(ColView cols, u32 width, size_t height, u64* __restrict__ digests, PosTab<P> T) {
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= height) return;
    if (width <= 4) {
        Fp<P> x = Fp<P>::zero();
        for (u32 c = 0; c < width; c++) { u64 v = big_col_load(cols, c, row); x.l[2 * c] = (u32)v; x.l[2 * c + 1] = (u32)(v >> 32); }
        x = reduce_raw<P>(x);
#pragma unroll
        for (int i = 0; i < 4; i++) digests[4 * row + i] = (u64)x.l[2 * i] | ((u64)x.l[2 * i + 1] << 32);
        return;
    }
    Fp<P> st[17], tmp[17];
    Fp<P> d = Fp<P>::zero();
    const u32 n3 = (width + 2) / 3;
    for (u32 i = 0; i < n3; i += 16) {
        const u32 cnt = n3 - i < 16 ? n3 - i : 16;
        st[0] = d;
        for (u32 k = 0; k < cnt; k++) {
            Fp<P> x = Fp<P>::zero();
            for (u32 e = 0; e < 3; e++) { u32 c = 3 * (i + k) + e; if (c < width) { u64 v = big_col_load(cols, c, row); x.l[2 * e] = (u32)v; x.l[2 * e + 1] = (u32)(v >> 32); } }
            st[1 + k] = x.to_mont();
        }
        poseidon_big<P>(st, tmp, (int)cnt + 1, T);
        d = st[LANE];
    }
    store_canon<P>(digests + 4 * row, d);
}
// one level: parent[i] = Poseidon_17(children[16 i .. 16 i + 16), init 0)   (merklehash_bn128.rs:68-87)
template <class P, int LANE> __global__ void __launch_bounds__(128) k_big_level(const u64* __restrict__ in, u64* __restrict__ out, size_t n_ops, PosTab<P> T) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_ops) return;
    Fp<P> st[17], tmp[17];
    st[0] = Fp<P>::zero();
    for (int k = 0; k < 16; k++) st[1 + k] = load_canon<P>(in + 4 * (16 * i + k));
    poseidon_big<P>(st, tmp, 17, T);
    store_canon<P>(out + 4 * i, st[LANE]);
}

// ------------------------------------------------------------------------------------------------ host
static std::string data_dir() {
    if (const char* e = getenv("B200ZK_DATA")) return e;
    Dl_info info;
    if (dladdr((void*)&data_dir, &info) && info.dli_fname) {
        std::string p = info.dli_fname;
        size_t k = p.find_last_of('/');
        return (k == std::string::npos ? std::string(".") : p.substr(0, k)) + "/data";
    }
    return "data";
}
template <class P> struct TabHolder { bool ready[16] = {false}; PosTab<P> tab[16]; };
template <class P> static const PosTab<P>& tables(const char* name) {
    static TabHolder<P> H; static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = current_device();
    if (H.ready[dev]) return H.tab[dev];
    std::string path = data_dir() + "/poseidon_" + name + ".bin";
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("cannot open " + path + " (set B200ZK_DATA)");
    std::vector<unsigned char> buf;
    { fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET); buf.resize((size_t)sz); if (fread(buf.data(), 1, buf.size(), f) != buf.size()) { fclose(f); throw std::runtime_error("short read: " + path); } fclose(f); }
    if (buf.size() < 12 || memcmp(buf.data(), "PSDB", 4)) throw std::runtime_error("bad constants file " + path);
    u32 nt; memcpy(&nt, buf.data() + 8, 4);
    // one device allocation: raw canonical copy + Montgomery table
    const size_t n_elems = (buf.size() - 12 - 16 * (size_t)nt) / 32;
    u64* d_raw; Fp<P>* d_m;
    B200_CUDA_CHECK(cudaMalloc(&d_raw, n_elems * 32)); B200_CUDA_CHECK(cudaMalloc(&d_m, n_elems * sizeof(Fp<P>)));
    std::vector<unsigned char> packed(n_elems * 32);
    PosTab<P> T{}; size_t off = 12, e = 0;
    for (u32 k = 0; k < nt; k++) {
        u32 h[4]; memcpy(h, buf.data() + off, 16); off += 16;
        const u32 t = h[0], nc = h[2], ns = h[3];
        if (t < 2 || t > 17) throw std::runtime_error("bad width in " + path);
        const size_t cnt = (size_t)nc + ns + 2 * (size_t)t * t;
        if (off + cnt * 32 > buf.size() || e + cnt > n_elems) throw std::runtime_error("truncated " + path);
        memcpy(packed.data() + e * 32, buf.data() + off, cnt * 32); off += cnt * 32;
        T.rp[t] = h[1]; T.C[t] = d_m + e; T.S[t] = d_m + e + nc; T.M[t] = d_m + e + nc + ns; T.Pm[t] = d_m + e + nc + ns + (size_t)t * t;
        e += cnt;
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(d_raw, packed.data(), e * 32, cudaMemcpyHostToDevice, stream()));
    k_big_to_mont<P><<<(unsigned)((e + 255) / 256), 256, 0, stream()>>>(d_raw, d_m, e);
    B200_CUDA_CHECK(cudaGetLastError()); B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    cudaFree(d_raw);
    H.tab[dev] = T; H.ready[dev] = true;
    return H.tab[dev];
}

size_t big_merkle_n_nodes(size_t n_) {       // merklehash_bn128.rs:26-40
    size_t n = n_, next_n = (n - 1) / 16 + 1, acc = next_n * 16;
    while (n > 1) { n = next_n; next_n = (n - 1) / 16 + 1; if (n > 1) acc += next_n * 16; else acc += 1; }
    return acc;
}

#ifndef BIG_WARP_MAX
#define BIG_WARP_MAX 8192
#endif
template <class P, int LANE> static void big_poseidon_t(const char* name, const u64* h_in, int t, u64* h_out) {
    const PosTab<P>& T = tables<P>(name);
    static u64* g_buf[16] = {nullptr}; static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = current_device();
    if (!g_buf[dev]) B200_CUDA_CHECK(cudaMalloc(&g_buf[dev], 2 * 17 * 32));
    u64* d = g_buf[dev];
    B200_CUDA_CHECK(cudaMemcpyAsync(d, h_in, (size_t)t * 32, cudaMemcpyHostToDevice, stream()));
    k_big_poseidon_single<P><<<1, 32, 0, stream()>>>(d, d + 17 * 4, t, T); launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(h_out, d + 17 * 4, (size_t)t * 32, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}
template <class P, int LANE> static void big_leaves_t(const char* name, ColView d_cols, size_t width, size_t height, u64* d_digests) {
    const PosTab<P>& T = tables<P>(name);
    if (height == 0) return;
    const size_t n3 = (width + 2) / 3;
    ScopedTimer tm("big_leaves", (8.0 * width + 32.0) * height);
    (void)n3;
    k_big_leaves<P, LANE><<<(unsigned)((height + 127) / 128), 128, 0, stream()>>>(d_cols, (u32)width, height, d_digests, T); launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}
template <class P, int LANE> static void big_levels_t(const char* name, u64* d_nodes, size_t height) {
    const PosTab<P>& T = tables<P>(name);
    size_t n = height, next_n = (n - 1) / 16 + 1, p_in = 0, p_out = next_n * 16;
    while (n > 1) {
        ScopedTimer tm("big_merkle_level", 544.0 * next_n);
        // below BIG_WARP_MAX parents the launch cannot fill the machine with one permutation per thread: one warp per parent
        if (next_n <= BIG_WARP_MAX) k_big_level_warp<P, LANE><<<(unsigned)((next_n + 3) / 4), 128, 0, stream()>>>(d_nodes + 4 * p_in, d_nodes + 4 * p_out, next_n, T);
        else k_big_level<P, LANE><<<(unsigned)((next_n + 127) / 128), 128, 0, stream()>>>(d_nodes + 4 * p_in, d_nodes + 4 * p_out, next_n, T);
        launch_count_add(1);
        B200_CUDA_CHECK(cudaGetLastError());
        n = next_n; next_n = (n - 1) / 16 + 1; p_in = p_out; p_out = p_in + next_n * 16;
    }
}

// field ids of the C-ABI: 0 = BN128 (output lane 0, poseidon_bn128_opt.rs:94-97), 1 = BLS12-381 (lane 1, poseidon_bls12381_opt.rs:95-103)
void big_poseidon_host(int field, const u64* h_state_in, int t, u64* h_state_out) {
    if (t < 2 || t > 17) throw std::invalid_argument("Wrong inputs length");      // the reference bails the same way
    if (field == 0) big_poseidon_t<Bn254Fr, 0>("bn128", h_state_in, t, h_state_out);
    else if (field == 1) big_poseidon_t<Bls381Fr, 1>("bls12381", h_state_in, t, h_state_out);
    else throw std::invalid_argument("unknown hash field id");
}
int big_out_lane(int field) { if (field == 0) return 0; if (field == 1) return 1; throw std::invalid_argument("unknown hash field id"); }
void big_leaves(int field, const u64* d_cols, size_t width, size_t height, u64* d_digests) { big_leaves_view(field, colview_plain(d_cols, height), width, height, d_digests); }
void big_leaves_view(int field, ColView d_cols, size_t width, size_t height, u64* d_digests) {
    if (field == 0) big_leaves_t<Bn254Fr, 0>("bn128", d_cols, width, height, d_digests);
    else if (field == 1) big_leaves_t<Bls381Fr, 1>("bls12381", d_cols, width, height, d_digests);
    else throw std::invalid_argument("unknown hash field id");
}
void big_merkle_levels(int field, u64* d_nodes, size_t height) {
    if (field == 0) big_levels_t<Bn254Fr, 0>("bn128", d_nodes, height);
    else if (field == 1) big_levels_t<Bls381Fr, 1>("bls12381", d_nodes, height);
    else throw std::invalid_argument("unknown hash field id");
}


// ------------------------------------------------------------------------------------------------ DevTree interface (stark.cpp)
// t.hash = 1 (BN128) / 2 (BLS12-381).  Same role as merkle.cu's merkelize / merkle_open for the GL tree.
void big_merkelize_tree(DevTree& t, ColView cols, size_t width, size_t height, u64* d_nodes) {
    const int field = t.hash - 1;
    t.cols = cols; t.width = width; t.height = height; t.nodes = d_nodes; t.degenerate = false; t.level_digest.clear();
    if (width == 0) {
        // empty section: zero leaf digests, every level still hashed (merklehash_bn128.rs:191-224 with an empty buffer).
        // All nodes of a level are equal while the level is a multiple of 16 wide; the last level (n < 16) mixes n copies
        // with 16 - n zero pads.  log16(height) permutations reproduce the reference's nodes.
        t.degenerate = true; t.nodes = nullptr;
        std::array<u64, 4> cur = {0, 0, 0, 0};
        t.level_digest.push_back(cur);
        size_t n = height;
        while (n > 1) {
            if (n >= 16 && (n % 16)) throw std::runtime_error("degenerate 16-ary tree needs a power-of-two height");
            const size_t real = n >= 16 ? 16 : n;
            u64 st[17 * 4], out[17 * 4];
            memset(st, 0, sizeof st);
            for (size_t k = 0; k < real; k++) memcpy(st + 4 * (1 + k), cur.data(), 32);
            big_poseidon_host(field, st, 17, out);
            memcpy(cur.data(), out + 4 * big_out_lane(field), 32);
            t.level_digest.push_back(cur);
            n = (n - 1) / 16 + 1;
        }
        memcpy(t.root, cur.data(), 32);
        return;
    }
    const size_t nn = big_merkle_n_nodes(height);
    B200_CUDA_CHECK(cudaMemsetAsync(d_nodes, 0, nn * 32, stream()));          // level padding = zero digests
    big_leaves_view(field, cols, width, height, d_nodes);
    big_merkle_levels(field, d_nodes, height);
    B200_CUDA_CHECK(cudaMemcpyAsync(t.root, d_nodes + 4 * (nn - 1), 32, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}

// merklehash_bn128.rs:89-106,226-243: leaf row + the 16 nodes of the leaf's group on every level, bottom-up
__global__ void k_big_open(ColView v, u32 width, size_t height, const u64* __restrict__ nodes, const u64* __restrict__ idxs,
                           u64* __restrict__ vals, u64* __restrict__ sibs, u32 depth) {
    size_t q = blockIdx.x;
    size_t idx = idxs[q];
    for (u32 c = threadIdx.x; c < width; c += blockDim.x) vals[q * width + c] = big_col_load(v, c, idx);
    if (threadIdx.x < 64) {
        size_t n = height, off = 0, id = idx; u32 d = 0;
        while (n > 1) {
            size_t si = id & ~(size_t)15;
            sibs[(q * depth + d) * 64 + threadIdx.x] = nodes[4 * (off + si) + threadIdx.x];
            size_t next_n = (n - 1) / 16 + 1;
            off += next_n * 16; id >>= 4; n = next_n; d++;
        }
    }
}
void big_merkle_open(const DevTree& t, const std::vector<u64>& idx, std::vector<u64>& vals, std::vector<u64>& sibs, size_t& depth) {
    const size_t nq = idx.size();
    depth = 0; { size_t n = t.height; while (n > 1) { n = (n - 1) / 16 + 1; depth++; } }
    vals.assign(nq * t.width, 0); sibs.assign(nq * depth * 64, 0);
    if (nq == 0) return;
    if (t.degenerate) {
        for (size_t q = 0; q < nq; q++) {
            size_t n = t.height;
            for (size_t d = 0; d < depth; d++) {
                const size_t real = n >= 16 ? 16 : n;
                for (size_t k = 0; k < real; k++) memcpy(&sibs[(q * depth + d) * 64 + 4 * k], t.level_digest[d].data(), 32);
                n = (n - 1) / 16 + 1;
            }
        }
        return;
    }
    const size_t nv = nq * t.width, ns = nq * depth * 64;
    u64* buf; B200_CUDA_CHECK(cudaMalloc(&buf, (nq + nv + ns + 1) * 8));
    u64 *d_idx = buf, *d_vals = buf + nq, *d_sibs = d_vals + nv;
    B200_CUDA_CHECK(cudaMemcpyAsync(d_idx, idx.data(), nq * 8, cudaMemcpyHostToDevice, stream()));
    k_big_open<<<(unsigned)nq, 128, 0, stream()>>>(t.cols, (u32)t.width, t.height, t.nodes, d_idx, d_vals, d_sibs, (u32)depth);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
    if (nv) B200_CUDA_CHECK(cudaMemcpyAsync(vals.data(), d_vals, nv * 8, cudaMemcpyDeviceToHost, stream()));
    if (ns) B200_CUDA_CHECK(cudaMemcpyAsync(sibs.data(), d_sibs, ns * 8, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    cudaFree(buf);
}

}  // namespace b200
