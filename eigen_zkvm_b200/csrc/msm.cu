// Multi-scalar multiplication (Pippenger) on BN254 and BLS12-381, G1 and G2, for sm_100a.
//
// Boundary: the multiexp calls inside `create_random_proof`, reached from `Groth16::prove`
// (groth16/src/groth16.rs:88-96 -> bellman_ce, BN254; groth16.rs:45-57 -> bellperson + blstrs, BLS12-381) /
// `groth16_prove` (groth16/src/api.rs:144-177).  The libraries' in-memory forms are kept: bases = affine (x, y) as
// little-endian MONTGOMERY limbs (R = 2^256 / 2^384; G2 coordinates are c0 || c1), the all-zero pair = infinity;
// scalars = canonical 4 x u64 `Repr`; the result is a Jacobian triple (X, Y, Z) in Montgomery form.
//
// Pipeline (all on the device; kernels are templates over the curve, see curve.cuh / mont.cuh):
//   1. signed c-bit window digits per scalar (buckets 1..2^(c-1), sign folded into the point index)   k_msm_digits
//   2. counting sort of point indices by bucket, per window (histogram -> scan -> scatter)              k_msm_scan / k_msm_scatter
//   3. one thread per (window, bucket, chunk of <= ch points): XYZZ accumulator += +-P over its contiguous
//      index run; chunks of a bucket are summed by a second small kernel (load balance for skewed scalars)  k_msm_accumulate / k_msm_bucket_finish
//   4. per window: sum_b b * B_b by two levels of segmented running sums + shared-memory tree           k_msm_reduce1/2
//   5. Horner over windows (c doublings each), normalisation to affine                                  k_msm_final
// Roofline class: INT (IMAD.WIDE issue): about 10 field products per mixed addition; HBM traffic is one affine
// point + 32 B per (point, scalar) plus 8 B per (point, window) of index traffic.
#include "b200_internal.h"
#include <cstring>
#include <algorithm>
#include "curve.cuh"

namespace b200 {

#define MSM_D __device__ __forceinline__

// ------------------------------------------------------------------------------------------------ kernels
// digits: dig[w * n + i] = bucket | sign << 31 (bucket 0 = nothing to add)
// merged = 1 (table mode): all windows share ONE bucket set (the table holds 2^(c w) P, so a digit of any window is just a small
// scalar for the table entry w * n + i); counts then has a single row of nb entries.
__global__ void k_msm_digits(const u32* __restrict__ scalars, const u32* __restrict__ bases, u32 base_words, size_t n, u32 c, u32 nw, u32 nb,
                             u32* __restrict__ dig, u32* __restrict__ counts, u32 merged, u32 scalar_bits, u32* __restrict__ range_err) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s[9];
    for (int k = 0; k < 8; k++) s[k] = scalars[8 * i + k];
    s[8] = 0;
    u32 any = 0;
    for (u32 k = 0; k < base_words; k++) any |= bases[(size_t)base_words * i + k];
    u32 carry = 0;
    const u32 half = 1u << (c - 1);
    for (u32 w = 0; w < nw; w++) {
        u32 off = w * c, limb = off >> 5, sh = off & 31;
        u32 v = 0;
        if (limb < 8) { u64 two = (u64)s[limb] | ((u64)s[limb + 1] << 32); v = (u32)(two >> sh) & ((1u << c) - 1); }
        v += carry;
        u32 sign = 0; carry = 0;
        if (v > half) { v = (1u << c) - v; sign = 1; carry = 1; }
        if (!any) v = 0;
        dig[(size_t)w * n + i] = v | (sign << 31);
        if (v) atomicAdd(&counts[merged ? (size_t)v : (size_t)w * nb + v], 1u);
    }
    // the windows cover SCALAR_BITS + 1 bits (the scalar and the signed-digit carry): a scalar with a bit at or above SCALAR_BITS is not a
    // canonical `Repr` of the curve's Fr and could lose its top carry -- flag it, the host turns the flag into an error
    u32 hi_bits = carry;
    { const u32 limb = scalar_bits >> 5; hi_bits |= s[limb] >> (scalar_bits & 31); for (u32 k = limb + 1; k < 8; k++) hi_bits |= s[k]; }
    if (hi_bits && any) atomicOr(range_err, 1u);
}
// exclusive scan of counts per window -> offsets (and a copy used as scatter cursors)
__global__ void k_msm_scan(const u32* __restrict__ counts, u32* __restrict__ offsets, u32* __restrict__ cursors, u32 nb) {
    __shared__ u32 part[1024];
    u32 w = blockIdx.x, t = threadIdx.x;
    u32 per = (nb + 1023) / 1024, lo = t * per, hi = lo + per < nb ? lo + per : nb;
    u32 s = 0;
    for (u32 b = lo; b < hi; b++) s += counts[(size_t)w * nb + b];
    part[t] = s;
    __syncthreads();
    for (u32 d = 1; d < 1024; d <<= 1) { u32 v = t >= d ? part[t - d] : 0; __syncthreads(); part[t] += v; __syncthreads(); }
    u32 run = t ? part[t - 1] : 0;
    for (u32 b = lo; b < hi; b++) { offsets[(size_t)w * nb + b] = run; cursors[(size_t)w * nb + b] = run; run += counts[(size_t)w * nb + b]; }
}
__global__ void k_msm_scatter(const u32* __restrict__ dig, u32* __restrict__ cursors, u32* __restrict__ sorted, size_t n, u32 nb) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 w = blockIdx.y;
    if (i >= n) return;
    u32 d = dig[(size_t)w * n + i];
    u32 b = d & 0x7fffffffu;
    if (!b) return;
    u32 pos = atomicAdd(&cursors[(size_t)w * nb + b], 1u);
    sorted[(size_t)w * n + pos] = (u32)i | (d & 0x80000000u);
}
// Load balancing.  A bucket's index run is cut into chunks of at most `ch` points; every (bucket, chunk) pair is one work
// item, so a witness with a skewed digit distribution (many scalars equal to 0 / 1 / a few small values -- the normal case
// for a groth16 witness) costs the same as a uniform one instead of serialising a million additions in one thread.
// nch[w][b] = ceil(count / ch); coff = its exclusive scan (k_msm_scan); item t of window w belongs to the last bucket b with
// coff[b] <= t.  For uniform scalars every bucket has one chunk and the second kernel is a copy.
__global__ void k_msm_chunks(const u32* __restrict__ counts, u32* __restrict__ nch, u32 nb, u32 ch, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    nch[i] = (i % nb) ? (counts[i] + ch - 1) / ch : 0;
}
template <class C> __global__ void __launch_bounds__(128) k_msm_accumulate(const Affine<typename C::F>* __restrict__ bases, const u32* __restrict__ sorted,
        const u32* __restrict__ offsets, const u32* __restrict__ counts, const u32* __restrict__ coff, const u32* __restrict__ nch,
        Xyzz<typename C::F>* __restrict__ partial, size_t n, u32 nb, u32 ch, u32 max_items) {
    typedef typename C::F F;
    const u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    const u32 w = blockIdx.y;
    const u32* co = coff + (size_t)w * nb;
    const u32 total = co[nb - 1] + nch[(size_t)w * nb + nb - 1];
    if (t >= total) return;
    u32 lo = 0, hi = nb;                    // upper_bound(co, t) - 1
    while (hi - lo > 1) { u32 mid = (lo + hi) >> 1; if (co[mid] <= t) lo = mid; else hi = mid; }
    const u32 b = lo, j = t - co[b];
    const u32 start = offsets[(size_t)w * nb + b] + j * ch, left = counts[(size_t)w * nb + b] - j * ch, cnt = left < ch ? left : ch;
    const u32* idx = sorted + (size_t)w * n + start;
    Xyzz<F> acc = Xyzz<F>::inf();
    for (u32 k = 0; k < cnt; k++) {
        u32 e = idx[k];
        Affine<F> p = bases[e & 0x7fffffffu];
        if (p.x.is_zero() && p.y.is_zero()) continue;      // table entries can be the point at infinity (2^k P = O); plain bases are screened by k_msm_digits
        if (e >> 31) p.y = p.y.neg();
        acc = acc.add_affine(p.x, p.y);
    }
    partial[(size_t)w * max_items + t] = acc;
}
template <class F> __global__ void __launch_bounds__(128) k_msm_bucket_finish(const Xyzz<F>* __restrict__ partial, const u32* __restrict__ coff, const u32* __restrict__ nch,
        Xyzz<F>* __restrict__ buckets, u32 nb, u32 max_items) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
    if (b >= nb) return;
    const u32 k = nch[(size_t)w * nb + b];
    const Xyzz<F>* p = partial + (size_t)w * max_items + coff[(size_t)w * nb + b];
    Xyzz<F> acc = k ? p[0] : Xyzz<F>::inf();
    for (u32 i = 1; i < k; i++) acc = acc.add(p[i]);
    buckets[(size_t)w * nb + b] = acc;
}
// window sum W = sum_{b=1}^{nb-1} b * B_b, two stages.  Buckets are cut into segments of RED_L; with b = s*RED_L + i
// (i in [1, RED_L]):  W = sum_s acc_s + RED_L * sum_s s * run_s,  run_s = sum_i B,  acc_s = sum_i i * B  (running sums).
#define RED_L 16
#define RED_T 128
// red_l (a power of two <= RED_L) = segment length: short segments for the small sets of the row / column split, where the
// 2 x red_l sequential additions of a thread are pure latency
template <class F> __global__ void __launch_bounds__(128) k_msm_reduce1(const Xyzz<F>* __restrict__ buckets, Xyzz<F>* __restrict__ seg_run, Xyzz<F>* __restrict__ seg_acc, u32 nb, u32 nseg, u32 red_l) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
    if (s >= nseg) return;
    u32 lo = 1 + s * red_l, hi = lo + red_l < nb ? lo + red_l : nb;
    Xyzz<F> run = Xyzz<F>::inf(), acc = Xyzz<F>::inf();
    // missing top buckets of a short last segment count as empty: start the running sum at the segment's nominal top
    for (u32 b = lo + red_l; b-- > lo;) { if (b < hi) run = run.add(buckets[(size_t)w * nb + b]); acc = acc.add(run); }
    seg_run[(size_t)w * nseg + s] = run; seg_acc[(size_t)w * nseg + s] = acc;
}
template <class F> __global__ void __launch_bounds__(RED_T) k_msm_reduce2(const Xyzz<F>* __restrict__ seg_run, const Xyzz<F>* __restrict__ seg_acc, Xyzz<F>* __restrict__ wsum, u32 nseg, u32 red_l) {
    extern __shared__ __align__(16) unsigned char sh_raw[];
    Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(sh_raw);
    u32 w = blockIdx.x, t = threadIdx.x;
    u32 G = (nseg + RED_T - 1) / RED_T, lo = t * G, hi = lo + G < nseg ? lo + G : nseg;
    Xyzz<F> A = Xyzz<F>::inf(), r = Xyzz<F>::inf(), a = Xyzz<F>::inf();
    for (u32 s = hi; s-- > lo;) { A = A.add(seg_acc[(size_t)w * nseg + s]); r = r.add(seg_run[(size_t)w * nseg + s]); a = a.add(r); }
    // sum_s s * run_s over this thread's range = (a - r) + lo * r
    Xyzz<F> Cs = a.add(r.neg());
    if (lo && lo < hi) Cs = Cs.add(r.mul_small(lo));
    for (u32 k = 1; k < red_l; k <<= 1) Cs = Cs.dbl();      // * red_l
    sh[t] = A.add(Cs);
    __syncthreads();
    for (u32 st = RED_T / 2; st > 0; st >>= 1) { if (t < st) sh[t] = sh[t].add(sh[t + st]); __syncthreads(); }
    if (t == 0) wsum[w] = sh[0];
}
// Large bucket sets (c > 13): sum_b b B_b with b = b1 L + b0 is  sum_b0 b0 S0[b0] + L sum_b1 b1 S1[b1]  with the column sums
// S0[b0] = sum_b1 B and the row sums S1[b1] = sum_b0 B: two fully parallel passes over the buckets (one CTA per output point,
// shared-memory tree), then the running-sum kernels above on L and nb / L entries.  grid: (rows or cols, window).
// out layout: [window][2][P] (0 = column sums, 1 = row sums), entries past `count` are the point at infinity, so that ONE launch of each
// running-sum kernel handles both weighted sums of all windows.  grid: (P, window, 2).
template <class F> __global__ void __launch_bounds__(128) k_msm_rowcol(const Xyzz<F>* __restrict__ buckets, Xyzz<F>* __restrict__ out_all, u32 nb, u32 L, u32 rows, u32 P) {
    extern __shared__ __align__(16) unsigned char sh_raw[];
    Xyzz<F>* sh = reinterpret_cast<Xyzz<F>*>(sh_raw);
    const u32 o = blockIdx.x, w = blockIdx.y, t = threadIdx.x;
    const int cols = blockIdx.z == 0;
    Xyzz<F>* out = out_all + ((size_t)w * 2 + blockIdx.z) * P;
    if (o >= (cols ? L : rows)) { if (t == 0) out[o] = Xyzz<F>::inf(); return; }
    const Xyzz<F>* B = buckets + (size_t)w * nb;
    Xyzz<F> acc = Xyzz<F>::inf();
    if (cols) { for (u32 b = o + t * L; b < nb; b += 128 * L) if (b) acc = acc.add(B[b]); }          // column o: b = b1 L + o
    else { for (u32 b0 = t; b0 < L; b0 += 128) { u32 b = o * L + b0; if (b && b < nb) acc = acc.add(B[b]); } }   // row o
    sh[t] = acc;
    __syncthreads();
    for (u32 st = 64; st > 0; st >>= 1) { if (t < st) sh[t] = sh[t].add(sh[t + st]); __syncthreads(); }
    if (t == 0) out[o] = sh[0];
}
// table[w * n + i] = 2^(c w) * P_i as affine points (all-zero = infinity): the per-circuit precomputation that lets every window of
// every scalar share one bucket set.  One thread per point, c doublings and one inversion per window.
template <class C> __global__ void __launch_bounds__(128) k_msm_table(const Affine<typename C::F>* __restrict__ bases, Affine<typename C::F>* __restrict__ tab, size_t n, u32 c, u32 nwin) {
    typedef typename C::F F;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = bases[i];
    tab[i] = p;
    bool inf = p.x.is_zero() && p.y.is_zero();
    for (u32 w = 1; w < nwin; w++) {
        if (!inf) {
            Xyzz<F> q = Xyzz<F>::dbl_affine(p.x, p.y);
            for (u32 k = 1; k < c; k++) q = q.dbl();
            if (q.is_inf()) inf = true;
            else { Jacobian<F> j = q.to_jacobian(); p.x = j.x; p.y = j.y; }
        }
        if (inf) { p.x = F::zero(); p.y = F::zero(); }
        tab[(size_t)w * n + i] = p;
    }
}
// sum of `count` Jacobian triples (the per-GPU partial results after the all-gather), normalised
template <class F> __global__ void k_points_sum(const Jacobian<F>* __restrict__ pts, u32 count, Jacobian<F>* __restrict__ out3) {
    if (threadIdx.x || blockIdx.x) return;
    Xyzz<F> tot = Xyzz<F>::inf();
    for (u32 k = 0; k < count; k++) tot = tot.add(Xyzz<F>::from_jacobian(pts[k]));
    *out3 = tot.to_jacobian();
}
// Horner over windows + affine normalisation; out = (X, Y, Z) Montgomery, Z = R (finite) or (0, R, 0)

// paired: the window sum is wsum[2 w] + 2^l0 * wsum[2 w + 1] (row / column split), else wsum[w]
template <class F> __global__ void k_msm_final(const Xyzz<F>* __restrict__ wsum, u32 paired, u32 l0, u32 nw, u32 c, Jacobian<F>* __restrict__ out3, u32 normalise) {
    if (threadIdx.x || blockIdx.x) return;
    Xyzz<F> tot = Xyzz<F>::inf();
    for (int w = (int)nw - 1; w >= 0; w--) {
        if (!tot.is_inf()) for (u32 k = 0; k < c; k++) tot = tot.dbl();
        if (paired) { Xyzz<F> h = wsum[2 * w + 1]; for (u32 k = 0; k < l0; k++) h = h.dbl(); tot = tot.add(h); tot = tot.add(wsum[2 * w]); }
        else tot = tot.add(wsum[w]);
    }
    *out3 = normalise ? tot.to_jacobian() : tot.to_jacobian_raw();
}

// out = a + b for two (X, Y, Z) Jacobian triples
template <class F> __global__ void k_point_add(const Jacobian<F>* __restrict__ a3, const Jacobian<F>* __restrict__ b3, Jacobian<F>* __restrict__ out3) {
    if (threadIdx.x || blockIdx.x) return;
    *out3 = Xyzz<F>::from_jacobian(*a3).add(Xyzz<F>::from_jacobian(*b3)).to_jacobian();
}

// deterministic pseudo-random curve points (bench / tests).
//   G1: x from SplitMix64(seed, i, attempt) taken as a Montgomery representative, y = sqrt(x^3 + b) (p = 3 mod 4)
//   G2: [k_i] G for a 64-bit k_i = SplitMix64(seed, i) | 1 (a square root in Fp2 is not worth the code here)
MSM_D u64 splitmix(u64 x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }
template <class C> __global__ void __launch_bounds__(128) k_random_points(Affine<typename C::F>* __restrict__ out, size_t n, u64 seed) {
    typedef typename C::F F;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if constexpr (!C::IS_G2) {
        typedef typename C::Base P;
        F b; for (int k = 0; k < P::N; k++) b.l[k] = C::coeff_b(k);
        u32 e[P::N]; for (int k = 0; k < P::N; k++) e[k] = P::sqrt_e(k);
        for (u32 attempt = 0;; attempt++) {
            F x;
            for (int k = 0; k < P::N / 2; k++) { u64 v = splitmix(seed ^ splitmix((u64)i * 8 + k + ((u64)attempt << 48))); x.l[2 * k] = (u32)v; x.l[2 * k + 1] = (u32)(v >> 32); }
            x.l[P::N - 1] &= (1u << ((P::BITS - 1) & 31)) - 1;        // < 2^(BITS-1) < p: a valid Montgomery representative
            F rhs = x.sqr() * x + b;
            F y = rhs.pow(e, P::N);
            if (y.sqr() == rhs) { out[i].x = x; out[i].y = y; return; }
        }
    } else {
        F gx, gy; u32* wx = reinterpret_cast<u32*>(&gx); u32* wy = reinterpret_cast<u32*>(&gy);
        for (int k = 0; k < F::N; k++) { wx[k] = C::gen_x(k); wy[k] = C::gen_y(k); }
        u64 k = splitmix(seed ^ splitmix((u64)i)) | 1;
        Xyzz<F> acc = Xyzz<F>::inf();
        for (int bit = 63; bit >= 0; bit--) { acc = acc.dbl(); if ((k >> bit) & 1) acc = acc.add_affine(gx, gy); }
        Jacobian<F> j = acc.to_jacobian();
        out[i].x = j.x; out[i].y = j.y;
    }
}

// ------------------------------------------------------------------------------------------------ host
static char* g_msm_ws[16] = {nullptr}; static size_t g_msm_ws_cap[16] = {0};
static char* msm_workspace(size_t bytes) {
    int dev = current_device();
    if (g_msm_ws_cap[dev] < bytes) {
        if (g_msm_ws[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_msm_ws[dev])); g_msm_ws[dev] = nullptr; g_msm_ws_cap[dev] = 0; }
        B200_CUDA_CHECK(cudaMalloc(&g_msm_ws[dev], bytes)); g_msm_ws_cap[dev] = bytes;
    }
    return g_msm_ws[dev];
}
// Window width: minimise (mixed additions) + (cost of a bucket in additions) x (buckets); table mode has ONE bucket set, so it affords wider
// windows.  The per-bucket cost of the table mode is FITTED (B200, 2^19 .. 2^22 points): accumulate = 0.18 ms per 10^6 additions + 3.2 ms per
// 10^6 buckets (a work item per bucket: the search for its bucket, its offsets, a 128-byte XYZZ store), reduction 0.9 ms per 10^6 buckets --
// about 20 additions per bucket, not the 3 of the textbook count; with 3 a 2^20-point share of a 4-GPU run took the 2^22-point window
// (2^19 buckets of 26 entries) and 6.1 ms.
static u32 msm_pick_c(size_t n, u32 scalar_bits, bool merged) {
    u32 best = 8; double best_cost = 1e300;
    for (u32 c = 6; c <= 22; c++) {
        const double nwin = (double)((scalar_bits + 1 + c - 1) / c), nbk = (double)(1u << (c - 1));
        // a narrow TOP window (254 - 13 * 19 = 7 bits at c = 19) puts all n of its digits into a few buckets: the histogram and scatter atomics
        // of those buckets serialise (digits 0.35 -> 0.99 ms, sort 1.43 -> 1.75 ms at 2^22) -- such widths are not candidates in table mode
        const int top_bits = (int)scalar_bits - (int)(nwin - 1) * (int)c;
        if (merged && top_bits < 12 && c > 12) continue;
        const double cost = (double)n * nwin + (merged ? 20.0 * nbk : 3.0 * nbk * nwin);
        if (cost < best_cost) { best_cost = cost; best = c; }
    }
    return best;
}
struct MsmTable { int curve; size_t n; u32 c, nwin; void* d_tab; int device; bool partial_out = false; };

// tab == nullptr: plain Pippenger over `nwin` windows with their own bucket sets.
// tab != nullptr: the table holds 2^(c w) P_i, every (scalar, window) digit is a small scalar for table entry w n + i and all of them
// share one bucket set: n * nwin "points", 1 "window", no doublings at the end.
template <class C> static void msm_run(const void* d_bases, const void* d_scalars, size_t n, void* h_out, const MsmTable* tab = nullptr) {
    typedef typename C::F F;
    typedef Xyzz<F> XY;
    const size_t out_bytes = sizeof(Jacobian<F>);
    if (n == 0) {       // (0, R, 0)
        std::vector<u32> r(out_bytes / 4, 0);
        for (int k = 0; k < C::Base::N; k++) r[sizeof(F) / 4 + k] = C::Base::one(k);
        memcpy(h_out, r.data(), out_bytes); return;
    }
    if (n >= (1ull << 31)) throw std::runtime_error("msm: n too large");
    const bool merged = tab != nullptr;
    const u32 normalise = (tab && tab->partial_out) ? 0u : 1u;
    const u32 c = merged ? tab->c : msm_pick_c(n, C::SCALAR_BITS, false);
    const u32 nwin_s = (C::SCALAR_BITS + 1 + c - 1) / c;              // digits per scalar: scalar bits + the signed-digit carry
    if (merged && (tab->n != n || tab->nwin != nwin_s)) throw std::invalid_argument("msm: table does not match the call");
    const size_t n_eff = merged ? n * nwin_s : n;                     // points per bucket set
    const u32 nwin = merged ? 1 : nwin_s;                             // bucket sets
    if (n_eff >= (1ull << 31)) throw std::runtime_error("msm: n * windows too large for the table mode");
    const u32 nb = (1u << (c - 1)) + 1;
    cudaStream_t st = stream();
    const u32 nseg = (nb - 1 + RED_L - 1) / RED_L;
    // row / column split of the bucket reduction for large bucket sets
    const bool split = c > 13;
    const u32 l0 = split ? (c - 1) / 2 : 0, L = 1u << l0, rows = split ? (nb - 1) / L + 1 : 0, P = split ? std::max(L, rows) : 0;
    const u32 red_l = split ? (P <= 2048 ? 4u : 16u) : (u32)RED_L;
    const u32 nseg_rc = split ? (P + red_l - 1) / red_l : 0;
    // one grow-only workspace per device (cudaMalloc/cudaFree per call cost far more than the kernels on multi-GPU hosts)
    // chunk length of a bucket's index run.  Two jobs: (i) a giant bucket (skewed witness scalars) becomes many work items instead of one
    // serial thread; (ii) small problems (a rank's 2^19-point share at 8 GPUs: 65 k buckets of ~120 entries) still get a few waves of work
    // items, so the launch ends at the MEAN bucket length rather than at the longest one (one thread per bucket was a single 0.86-full
    // wave there: accumulate 2.1 ms against 1.66 ms of additions).  About 300 k items over all windows (4 waves of 128-thread CTAs at 4 per SM) at most.
    // With enough buckets to fill the machine only the outliers are split: twice the mean run length.  Outliers are the rule, not the
    // exception: the TOP window of a 254-bit scalar is narrow (7 bits at c = 19: 128 buckets share all n entries -- with the former
    // "n / 8192" chunks those were 7168 serial additions per thread and a 2^22-point MSM took 43 ms instead of 12).
    const size_t mean_run = n_eff / nb + 1;
    const u32 ch = (size_t)nwin * nb >= 250000 ? (u32)std::min<size_t>(1024, std::max<size_t>(64, 2 * mean_run))
                                               : (u32)std::max<size_t>(32, (size_t)nwin * n_eff / 300000);
    const u32 max_items = nb + (u32)(n_eff / ch) + 1;                 // sum_b ceil(count_b / ch) <= nb + n / ch
    const size_t b_idx = (size_t)nwin * n_eff * 4, b_cnt = (size_t)nwin * nb * 4 * 5,
                 b_pts = ((size_t)nwin * nb + 2 * nwin + 2 * (size_t)nwin * (nseg + 2 * nseg_rc) + (size_t)nwin * max_items + 2 * (size_t)nwin * P) * sizeof(XY) + out_bytes + 256;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char* ws = msm_workspace(al(b_idx) * 2 + al(b_cnt) + al(b_pts));
    u32* dig = reinterpret_cast<u32*>(ws); u32* sorted = reinterpret_cast<u32*>(ws + al(b_idx)); u32* counts = reinterpret_cast<u32*>(ws + 2 * al(b_idx));
    u32* offsets = counts + (size_t)nwin * nb; u32* cursors = offsets + (size_t)nwin * nb; u32* nch = cursors + (size_t)nwin * nb; u32* coff = nch + (size_t)nwin * nb;
    XY* buckets = reinterpret_cast<XY*>(ws + 2 * al(b_idx) + al(b_cnt));
    XY* wsum = buckets + (size_t)nwin * nb;                       // 2 * nwin entries (pairs when split)
    XY* seg_run = wsum + 2 * nwin; XY* seg_acc = seg_run + (size_t)nwin * (nseg + 2 * nseg_rc);
    XY* partial = seg_acc + (size_t)nwin * (nseg + 2 * nseg_rc);
    XY* rc = partial + (size_t)nwin * max_items;                  // [nwin][2][P]
    Jacobian<F>* d_out = reinterpret_cast<Jacobian<F>*>(rc + 2 * (size_t)nwin * P);
    u32* d_range_err = reinterpret_cast<u32*>(reinterpret_cast<char*>(d_out) + out_bytes);      // inside the 256 spare bytes of b_pts
    B200_CUDA_CHECK(cudaMemsetAsync(d_range_err, 0, 4, st));
    B200_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)nwin * nb * 4, st));
    const double pair_bytes = (double)sizeof(Affine<F>) + 32.0;
    const Affine<F>* pts = merged ? (const Affine<F>*)tab->d_tab : (const Affine<F>*)d_bases;
    {
        ScopedTimer t("msm_digits", 32.0 * n);
        k_msm_digits<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const u32*)d_scalars, (const u32*)pts, (u32)(sizeof(Affine<F>) / 4), n, c, nwin_s, nb, dig, counts, merged ? 1u : 0u, (u32)C::SCALAR_BITS, d_range_err);
    }
    { ScopedTimer t("msm_sort", 8.0 * n * nwin_s); k_msm_scan<<<nwin, 1024, 0, st>>>(counts, offsets, cursors, nb);
      k_msm_scatter<<<dim3((unsigned)((n_eff + 255) / 256), nwin), 256, 0, st>>>(dig, cursors, sorted, n_eff, nb); }
    { ScopedTimer t("msm_accumulate", pair_bytes * n);
      k_msm_chunks<<<(unsigned)(((size_t)nwin * nb + 255) / 256), 256, 0, st>>>(counts, nch, nb, ch, (size_t)nwin * nb);
      k_msm_scan<<<nwin, 1024, 0, st>>>(nch, coff, cursors, nb);            // cursors: scratch output (the scatter is done)
      k_msm_accumulate<C><<<dim3((max_items + 127) / 128, nwin), 128, 0, st>>>(pts, sorted, offsets, counts, coff, nch, partial, n_eff, nb, ch, max_items);
      k_msm_bucket_finish<F><<<dim3((nb + 127) / 128, nwin), 128, 0, st>>>(partial, coff, nch, buckets, nb, max_items); }
    { ScopedTimer t("msm_reduce", (double)sizeof(XY) * nb * nwin);
      static bool attr_done[16] = {false};
      int dev = current_device();
      B200_CUDA_CHECK(cudaFuncSetAttribute(k_msm_reduce2<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(RED_T * sizeof(XY))));
      B200_CUDA_CHECK(cudaFuncSetAttribute(k_msm_rowcol<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(128 * sizeof(XY))));
      (void)attr_done;
      if (!split) {
          k_msm_reduce1<F><<<dim3((nseg + 127) / 128, nwin), 128, 0, st>>>(buckets, seg_run, seg_acc, nb, nseg, red_l);
          k_msm_reduce2<F><<<nwin, RED_T, RED_T * sizeof(XY), st>>>(seg_run, seg_acc, wsum, nseg, red_l);
          k_msm_final<F><<<1, 32, 0, st>>>(wsum, 0, 0, nwin, c, d_out, normalise);
          launch_count_add(3);
      } else {
          // column and row sums of every window in one launch, then both weighted sums of every window in one launch of each
          // running-sum kernel (2 * nwin "windows" of P entries; entry 0 has weight 0)
          k_msm_rowcol<F><<<dim3(P, nwin, 2), 128, 128 * sizeof(XY), st>>>(buckets, rc, nb, L, rows, P);
          k_msm_reduce1<F><<<dim3((nseg_rc + 127) / 128, 2 * nwin), 128, 0, st>>>(rc, seg_run, seg_acc, P, nseg_rc, red_l);
          k_msm_reduce2<F><<<2 * nwin, RED_T, RED_T * sizeof(XY), st>>>(seg_run, seg_acc, wsum, nseg_rc, red_l);
          k_msm_final<F><<<1, 32, 0, st>>>(wsum, 1, l0, nwin, c, d_out, normalise);
          launch_count_add(4);
      }
    }
    launch_count_add(7);
    B200_CUDA_CHECK(cudaGetLastError());
    std::vector<unsigned char> res(out_bytes + 4);
    B200_CUDA_CHECK(cudaMemcpyAsync(res.data(), d_out, out_bytes + 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    u32 range_err; memcpy(&range_err, res.data() + out_bytes, 4);
    if (range_err) throw std::invalid_argument("msm: a scalar has a bit at or above the curve's scalar size (scalars must be canonical `Repr`s, < r)");
    memcpy(h_out, res.data(), out_bytes);
}
// ---- per-circuit table of shifted bases (groth16: the bases are the proving key, fixed per circuit)
template <class C> static MsmTable* msm_table_build(int curve, const void* d_bases, size_t n) {
    typedef typename C::F F;
    if (n == 0 || n >= (1ull << 31)) throw std::invalid_argument("msm table: bad n");
    MsmTable* t = new MsmTable();
    t->curve = curve; t->n = n; t->c = msm_pick_c(n, C::SCALAR_BITS, true); t->nwin = (C::SCALAR_BITS + 1 + t->c - 1) / t->c; t->d_tab = nullptr;
    if ((unsigned long long)n * t->nwin >= (1ull << 31)) { t->c = 16; t->nwin = (C::SCALAR_BITS + 1 + 15) / 16; }
    B200_CUDA_CHECK(cudaGetDevice(&t->device));
    cudaError_t e = cudaMalloc(&t->d_tab, (size_t)t->nwin * n * sizeof(Affine<F>));
    if (e != cudaSuccess) { delete t; throw std::runtime_error(std::string("msm table: cudaMalloc failed: ") + cudaGetErrorString(e)); }
    k_msm_table<C><<<(unsigned)((n + 127) / 128), 128, 0, stream()>>>((const Affine<F>*)d_bases, (Affine<F>*)t->d_tab, n, t->c, t->nwin);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    return t;
}
template <class C> static void points_sum_run(const void* d_points, size_t count, void* h_out) {
    typedef Jacobian<typename C::F> J;
    J* d_out = reinterpret_cast<J*>(msm_workspace(4096));
    k_points_sum<typename C::F><<<1, 32, 0, stream()>>>((const J*)d_points, (u32)count, d_out); launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(h_out, d_out, sizeof(J), cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}
template <class C> static void msm_host(const void* bases, const void* scalars, size_t n, void* h_out) {
    // staging buffers are grow-only and per device, like the workspace
    static char* g_in[16] = {nullptr}; static size_t g_in_cap[16] = {0};
    const size_t pb = sizeof(Affine<typename C::F>);
    int dev = current_device();
    size_t need = (n ? n : 1) * (pb + 32);
    if (g_in_cap[dev] < need) { if (g_in[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_in[dev])); } B200_CUDA_CHECK(cudaMalloc(&g_in[dev], need)); g_in_cap[dev] = need; }
    char* db = g_in[dev]; char* ds = db + (n ? n : 1) * pb;
    B200_CUDA_CHECK(cudaMemcpyAsync(db, bases, n * pb, cudaMemcpyHostToDevice, stream()));
    B200_CUDA_CHECK(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, stream()));
    msm_run<C>(db, ds, n, h_out);
}
template <class C> static void point_add_host(const void* a, const void* b, void* out) {
    typedef Jacobian<typename C::F> J;
    J* d; B200_CUDA_CHECK(cudaMalloc(&d, 3 * sizeof(J)));
    B200_CUDA_CHECK(cudaMemcpyAsync(d, a, sizeof(J), cudaMemcpyHostToDevice, stream())); B200_CUDA_CHECK(cudaMemcpyAsync(d + 1, b, sizeof(J), cudaMemcpyHostToDevice, stream()));
    k_point_add<typename C::F><<<1, 32, 0, stream()>>>(d, d + 1, d + 2); launch_count_add(1);
    B200_CUDA_CHECK(cudaMemcpyAsync(out, d + 2, sizeof(J), cudaMemcpyDeviceToHost, stream())); B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    cudaFree(d);
}
template <class C> static void random_points(void* d_bases, size_t n, u64 seed) {
    k_random_points<C><<<(unsigned)((n + 127) / 128), 128, 0, stream()>>>((Affine<typename C::F>*)d_bases, n, seed); launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError()); B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}

// curve ids of the C-ABI (include/b200zk.h: B200_CURVE_*)
size_t msm_point_bytes(int curve) {
    switch (curve) { case 0: return sizeof(Affine<Bn254G1::F>); case 1: return sizeof(Affine<Bn254G2::F>); case 2: return sizeof(Affine<Bls381G1::F>); case 3: return sizeof(Affine<Bls381G2::F>); }
    throw std::invalid_argument("unknown curve id");
}
#define MSM_DISPATCH(curve, CALL)                                                    \
    switch (curve) {                                                                 \
    case 0: { typedef Bn254G1 C; CALL; break; }                                      \
    case 1: { typedef Bn254G2 C; CALL; break; }                                      \
    case 2: { typedef Bls381G1 C; CALL; break; }                                     \
    case 3: { typedef Bls381G2 C; CALL; break; }                                     \
    default: throw std::invalid_argument("unknown curve id");                        \
    }
void msm_dev(int curve, const void* d_bases, const void* d_scalars, size_t n, void* h_out) { MSM_DISPATCH(curve, msm_run<C>(d_bases, d_scalars, n, h_out)); }
// the window width and the number of windows msm_run / msm_table_build would use (host logic only; tests/test_abi_cpu.py)
void msm_window_choice(int curve, size_t n, bool table_mode, u32* c_out, u32* nwin_out) {
    MSM_DISPATCH(curve, { const u32 c = msm_pick_c(n, C::SCALAR_BITS, table_mode); *c_out = c; *nwin_out = (C::SCALAR_BITS + 1 + c - 1) / c; });
}
void msm_host_buffers(int curve, const void* bases, const void* scalars, size_t n, void* h_out) { MSM_DISPATCH(curve, msm_host<C>(bases, scalars, n, h_out)); }
void msm_point_add(int curve, const void* a, const void* b, void* out) { MSM_DISPATCH(curve, point_add_host<C>(a, b, out)); }
void msm_random_points_dev(int curve, void* d_bases, size_t n, u64 seed) { MSM_DISPATCH(curve, random_points<C>(d_bases, n, seed)); }
MsmTable* msm_table_new(int curve, const void* d_bases, size_t n) { MsmTable* t = nullptr; MSM_DISPATCH(curve, t = msm_table_build<C>(curve, d_bases, n)); return t; }
void msm_table_free(MsmTable* t) { if (!t) return; if (t->d_tab) cudaFree(t->d_tab); delete t; }
void msm_table_set_partial_output(MsmTable* t, bool on) { t->partial_out = on; }
void msm_table_info(const MsmTable* t, u32* c, u32* nwin, size_t* n) { *c = t->c; *nwin = t->nwin; *n = t->n; }
void msm_table_run(const MsmTable* t, const void* d_scalars, void* h_out) { MSM_DISPATCH(t->curve, msm_run<C>(nullptr, d_scalars, t->n, h_out, t)); }
void msm_table_run_host(const MsmTable* t, const void* scalars, void* h_out) {
    static char* g_sc[16] = {nullptr}; static size_t g_sc_cap[16] = {0};
    int dev = current_device();
    size_t need = t->n * 32;
    if (g_sc_cap[dev] < need) { if (g_sc[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_sc[dev])); } B200_CUDA_CHECK(cudaMalloc(&g_sc[dev], need)); g_sc_cap[dev] = need; }
    B200_CUDA_CHECK(cudaMemcpyAsync(g_sc[dev], scalars, need, cudaMemcpyHostToDevice, stream()));
    msm_table_run(t, g_sc[dev], h_out);
}
void msm_points_sum_dev(int curve, const void* d_points, size_t count, void* h_out) { MSM_DISPATCH(curve, points_sum_run<C>(d_points, count, h_out)); }

}  // namespace b200
