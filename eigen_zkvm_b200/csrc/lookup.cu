// Device versions of the two serial helpers of stark_gen's rounds 2-3:
//   calculate_H1H2 (starky/src/stark_gen.rs:625-651): plookup sorted vectors h1, h2 from f and t
//   calculate_Z    (starky/src/stark_gen.rs:653-666) + batch_inverse (polutils.rs:35-53): grand-product column
// Both are exact field/ordering computations, so parallel reformulations give identical results:
//   Z:    z[i] = prod_{j<i} num[j]/den[j]  -> per-element inverse (Montgomery batches in registers), then a three
//         phase exclusive scan under GF(p^3) multiplication; the wrap-around check z[N-1]*num/den == 1 is kept.
//   H1H2: the reference builds idx_t[value] = last index of value in t (HashMap), then stably sorts
//         [(t_i, i)] ++ [(f_i, idx_t[f_i])] by index.  Here: radix-sort t by value (3 stable 64-bit passes), binary
//         search every f_i for the last equal element, histogram + exclusive scan of the hit counts, one more stable
//         radix sort of the f entries by hit index; then entry positions are closed-form:
//           pos(t_j) = j + #{f : idx < j},   pos(f with sorted rank q, idx j) = q + j + 1.
// Polynomials are column-major device sections; a dim-1 polynomial is one column, dim-3 three consecutive columns.
#include "b200_internal.h"
#include "field.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

namespace b200 {

// ------------------------------------------------------------------------------------------------ helpers
struct PolRef { const u64* base; u64 rows; u32 dim; };      // lanes at base + l*rows
struct PolOut { u64* base; u64 rows; u32 dim; };
GL_D f3 pol_get(const PolRef& p, size_t i) { return p.dim == 1 ? f3_make(p.base[i], 0, 0) : f3_make(p.base[i], p.base[p.rows + i], p.base[2 * p.rows + i]); }
GL_D void pol_set(const PolOut& p, size_t i, f3 v) { p.base[i] = v.c[0]; if (p.dim == 3) { p.base[p.rows + i] = v.c[1]; p.base[2 * p.rows + i] = v.c[2]; } }

// ------------------------------------------------------------------------------------------------ calculate_Z
#define ZB 8
// r[i] = num[i] / den[i], ZB elements per thread (strided), one inversion per thread
__global__ void __launch_bounds__(128) k_ratio(PolRef num, PolRef den, u64* __restrict__ r /* 3 x n */, size_t n) {
    size_t stride = (size_t)gridDim.x * blockDim.x, t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    f3 pre[ZB], d[ZB]; f3 acc = f3_make(1, 0, 0); int cnt = 0;
#pragma unroll
    for (int j = 0; j < ZB; j++) { size_t k = t + (size_t)j * stride; if (k < n) { d[j] = pol_get(den, k); pre[j] = acc; acc = f3_mul(acc, d[j]); cnt = j + 1; } }
    if (!cnt) return;
    f3 inv = f3_inv(acc);
#pragma unroll
    for (int j = ZB - 1; j >= 0; j--) if (j < cnt) {
        size_t k = t + (size_t)j * stride;
        f3 di = f3_mul(inv, pre[j]); inv = f3_mul(inv, d[j]);
        f3 q = f3_mul(pol_get(num, k), di);
        r[k] = q.c[0]; r[n + k] = q.c[1]; r[2 * n + k] = q.c[2];
    }
}
#define SC_T 256
#define SC_E 8
// phase A: product of each block's SC_T*SC_E chunk
__global__ void __launch_bounds__(SC_T) k_scan_blockprod(const u64* __restrict__ r, size_t n, u64* __restrict__ bprod) {
    __shared__ u64 sh[3][SC_T];
    size_t base = ((size_t)blockIdx.x * SC_T + threadIdx.x) * SC_E;
    f3 acc = f3_make(1, 0, 0);
    for (int e = 0; e < SC_E; e++) { size_t k = base + e; if (k < n) acc = f3_mul(acc, f3_make(r[k], r[n + k], r[2 * n + k])); }
    for (int l = 0; l < 3; l++) sh[l][threadIdx.x] = acc.c[l];
    __syncthreads();
    for (int s = SC_T / 2; s > 0; s >>= 1) {
        if ((int)threadIdx.x < s) { f3 a = f3_make(sh[0][threadIdx.x], sh[1][threadIdx.x], sh[2][threadIdx.x]), b = f3_make(sh[0][threadIdx.x + s], sh[1][threadIdx.x + s], sh[2][threadIdx.x + s]);
            f3 c = f3_mul(a, b); for (int l = 0; l < 3; l++) sh[l][threadIdx.x] = c.c[l]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) for (int l = 0; l < 3; l++) bprod[3 * blockIdx.x + l] = sh[l][0];
}
// phase B: exclusive scan of the block products (single block, serial chunks + Hillis-Steele); total -> bprod[3*nb..]
__global__ void __launch_bounds__(SC_T) k_scan_blocks(u64* __restrict__ bprod, size_t nb) {
    __shared__ u64 sh[3][SC_T];
    size_t per = (nb + SC_T - 1) / SC_T, lo = threadIdx.x * per, hi = lo + per < nb ? lo + per : nb;
    f3 acc = f3_make(1, 0, 0);
    for (size_t b = lo; b < hi; b++) acc = f3_mul(acc, f3_make(bprod[3 * b], bprod[3 * b + 1], bprod[3 * b + 2]));
    for (int l = 0; l < 3; l++) sh[l][threadIdx.x] = acc.c[l];
    __syncthreads();
    for (int d = 1; d < SC_T; d <<= 1) {
        f3 v = f3_make(1, 0, 0);
        if ((int)threadIdx.x >= d) v = f3_make(sh[0][threadIdx.x - d], sh[1][threadIdx.x - d], sh[2][threadIdx.x - d]);
        __syncthreads();
        if ((int)threadIdx.x >= d) { f3 c = f3_mul(v, f3_make(sh[0][threadIdx.x], sh[1][threadIdx.x], sh[2][threadIdx.x])); for (int l = 0; l < 3; l++) sh[l][threadIdx.x] = c.c[l]; }
        __syncthreads();
    }
    f3 run = threadIdx.x ? f3_make(sh[0][threadIdx.x - 1], sh[1][threadIdx.x - 1], sh[2][threadIdx.x - 1]) : f3_make(1, 0, 0);
    if (threadIdx.x == SC_T - 1) for (int l = 0; l < 3; l++) bprod[3 * nb + l] = sh[l][SC_T - 1];     // grand total
    for (size_t b = lo; b < hi; b++) { f3 cur = f3_make(bprod[3 * b], bprod[3 * b + 1], bprod[3 * b + 2]); for (int l = 0; l < 3; l++) bprod[3 * b + l] = run.c[l]; run = f3_mul(run, cur); }
}
// phase C: z[i] = prefix(block) * prod of the chunk's earlier elements
__global__ void __launch_bounds__(SC_T) k_scan_apply(const u64* __restrict__ r, size_t n, const u64* __restrict__ bprod, PolOut z) {
    __shared__ u64 sh[3][SC_T];
    size_t base = ((size_t)blockIdx.x * SC_T + threadIdx.x) * SC_E;
    f3 loc[SC_E]; f3 acc = f3_make(1, 0, 0);
    for (int e = 0; e < SC_E; e++) { size_t k = base + e; loc[e] = acc; if (k < n) acc = f3_mul(acc, f3_make(r[k], r[n + k], r[2 * n + k])); }
    for (int l = 0; l < 3; l++) sh[l][threadIdx.x] = acc.c[l];
    __syncthreads();
    for (int d = 1; d < SC_T; d <<= 1) {
        f3 v = f3_make(1, 0, 0);
        if ((int)threadIdx.x >= d) v = f3_make(sh[0][threadIdx.x - d], sh[1][threadIdx.x - d], sh[2][threadIdx.x - d]);
        __syncthreads();
        if ((int)threadIdx.x >= d) { f3 c = f3_mul(v, f3_make(sh[0][threadIdx.x], sh[1][threadIdx.x], sh[2][threadIdx.x])); for (int l = 0; l < 3; l++) sh[l][threadIdx.x] = c.c[l]; }
        __syncthreads();
    }
    f3 pre = f3_make(bprod[3 * blockIdx.x], bprod[3 * blockIdx.x + 1], bprod[3 * blockIdx.x + 2]);
    if (threadIdx.x) pre = f3_mul(pre, f3_make(sh[0][threadIdx.x - 1], sh[1][threadIdx.x - 1], sh[2][threadIdx.x - 1]));
    for (int e = 0; e < SC_E; e++) { size_t k = base + e; if (k < n) pol_set(z, k, f3_mul(pre, loc[e])); }
}

void calculate_Z(const u64* num, u32 num_dim, const u64* den, u32 den_dim, u64* z, u32 z_dim, size_t n, u64* d_tmp /* 3n + 3*(nblocks+1) u64 */) {
    if (z_dim != 3) throw std::runtime_error("calculate_Z: Z polynomial must have dim 3");
    PolRef pn{num, n, num_dim}, pd{den, n, den_dim}; PolOut pz{z, n, z_dim};
    u64* r = d_tmp; size_t nblk = (n + (size_t)SC_T * SC_E - 1) / ((size_t)SC_T * SC_E); u64* bprod = d_tmp + 3 * n;
    ScopedTimer t("calculate_Z", 24.0 * 3 * (double)n);
    size_t nth = (n + ZB - 1) / ZB;
    k_ratio<<<(unsigned)((nth + 127) / 128), 128, 0, stream()>>>(pn, pd, r, n);
    k_scan_blockprod<<<(unsigned)nblk, SC_T, 0, stream()>>>(r, n, bprod);
    k_scan_blocks<<<1, SC_T, 0, stream()>>>(bprod, nblk);
    k_scan_apply<<<(unsigned)nblk, SC_T, 0, stream()>>>(r, n, bprod, pz);
    launch_count_add(4);
    B200_CUDA_CHECK(cudaGetLastError());
    u64 tot[3];
    B200_CUDA_CHECK(cudaMemcpyAsync(tot, bprod + 3 * nblk, 24, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    if (!(tot[0] == 1 && tot[1] == 0 && tot[2] == 0)) throw std::runtime_error("calculate_Z: grand product does not wrap to 1 (stark_gen.rs:663-664 assertion)");
}
size_t calculate_Z_tmp_u64(size_t n) { return 3 * n + 3 * ((n + (size_t)SC_T * SC_E - 1) / ((size_t)SC_T * SC_E) + 2); }

// ------------------------------------------------------------------------------------------------ calculate_H1H2
__global__ void k_iota(u32* a, size_t n) { size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = (u32)i; }
__global__ void k_gather_lane(const u64* __restrict__ lane, const u32* __restrict__ idx, u64* __restrict__ out, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = lane ? lane[idx[i]] : 0;
}
GL_D int key_cmp(u64 a0, u64 a1, u64 a2, u64 b0, u64 b1, u64 b2) {    // lexicographic, lane 0 most significant
    if (a0 != b0) return a0 < b0 ? -1 : 1;
    if (a1 != b1) return a1 < b1 ? -1 : 1;
    if (a2 != b2) return a2 < b2 ? -1 : 1;
    return 0;
}
// idx_f[i] = original index of the LAST element of t equal to f[i]; counts[idx]++ ; flag on miss
__global__ void k_lookup(PolRef f, PolRef t, const u32* __restrict__ tperm /* t sorted by value, stable */, size_t n, u32* __restrict__ idx_f, u32* __restrict__ counts, u32* __restrict__ miss) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    f3 v = pol_get(f, i);
    size_t lo = 0, hi = n;                       // upper bound: first position with key > v
    while (lo < hi) { size_t mid = (lo + hi) >> 1; f3 w = pol_get(t, tperm[mid]); if (key_cmp(w.c[0], w.c[1], w.c[2], v.c[0], v.c[1], v.c[2]) <= 0) lo = mid + 1; else hi = mid; }
    if (lo == 0) { atomicExch(miss, 1u); idx_f[i] = 0; return; }
    u32 j = tperm[lo - 1]; f3 w = pol_get(t, j);
    if (key_cmp(w.c[0], w.c[1], w.c[2], v.c[0], v.c[1], v.c[2]) != 0) { atomicExch(miss, 1u); idx_f[i] = 0; return; }
    idx_f[i] = j; atomicAdd(&counts[j], 1u);
}
// scatter s = sorted-by-index list into h1 (even positions) / h2 (odd positions)
__global__ void k_place_t(PolRef t, const u32* __restrict__ offs, size_t n, PolOut h1, PolOut h2) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (j >= n) return;
    size_t pos = j + offs[j]; f3 v = pol_get(t, j);
    if (pos & 1) pol_set(h2, pos >> 1, v); else pol_set(h1, pos >> 1, v);
}
__global__ void k_place_f(PolRef f, const u32* __restrict__ idx_sorted, const u32* __restrict__ fi_sorted, size_t n, PolOut h1, PolOut h2) {
    size_t q = (size_t)blockIdx.x * blockDim.x + threadIdx.x; if (q >= n) return;
    size_t pos = q + (size_t)idx_sorted[q] + 1; f3 v = pol_get(f, fi_sorted[q]);
    if (pos & 1) pol_set(h2, pos >> 1, v); else pol_set(h1, pos >> 1, v);
}

void calculate_H1H2(const u64* f, const u64* t, u32 dim, size_t n, u64* h1, u64* h2) {
    if (n >= (1ull << 32)) throw std::runtime_error("calculate_H1H2: n too large");
    PolRef pf{f, n, dim}, pt{t, n, dim}; PolOut o1{h1, n, dim}, o2{h2, n, dim};
    cudaStream_t st = stream();
    ScopedTimer tm("calculate_H1H2", 8.0 * dim * 4 * (double)n);
    unsigned g = (unsigned)((n + 255) / 256);
    u32 *perm_a, *perm_b, *idx_f, *counts, *offs, *miss, *fi, *idx_s, *fi_s; u64 *key_a, *key_b;
    size_t words32 = 8 * n + 16;
    B200_CUDA_CHECK(cudaMalloc(&perm_a, words32 * 4)); perm_b = perm_a + n; idx_f = perm_b + n; counts = idx_f + n; offs = counts + n; fi = offs + n; idx_s = fi + n; fi_s = idx_s + n; miss = fi_s + n;
    B200_CUDA_CHECK(cudaMalloc(&key_a, 2 * n * 8)); key_b = key_a + n;
    size_t tmp_bytes = 0, tb2 = 0, tb3 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, key_a, key_b, perm_a, perm_b, (int)n, 0, 64, st);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, idx_f, idx_s, fi, fi_s, (int)n, 0, 32, st);
    cub::DeviceScan::ExclusiveSum(nullptr, tb3, counts, offs, (int)n, st);
    if (tb2 > tmp_bytes) tmp_bytes = tb2; if (tb3 > tmp_bytes) tmp_bytes = tb3;
    void* d_tmp; B200_CUDA_CHECK(cudaMalloc(&d_tmp, tmp_bytes));
    try {
        // t sorted by value: LSD over lanes 2, 1, 0 with stable 64-bit radix passes
        k_iota<<<g, 256, 0, st>>>(perm_a, n);
        for (int lane = (int)dim - 1; lane >= 0; lane--) {
            k_gather_lane<<<g, 256, 0, st>>>(t + (size_t)lane * n, perm_a, key_a, n);
            cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, key_a, key_b, perm_a, perm_b, (int)n, 0, 64, st);
            std::swap(perm_a, perm_b);
        }
        B200_CUDA_CHECK(cudaMemsetAsync(counts, 0, n * 4, st)); B200_CUDA_CHECK(cudaMemsetAsync(miss, 0, 4, st));
        k_lookup<<<g, 256, 0, st>>>(pf, pt, perm_a, n, idx_f, counts, miss);
        u32 h_miss = 0; B200_CUDA_CHECK(cudaMemcpyAsync(&h_miss, miss, 4, cudaMemcpyDeviceToHost, st)); B200_CUDA_CHECK(cudaStreamSynchronize(st));
        if (h_miss) throw std::runtime_error("calculate_H1H2: Number not included (a looked-up value is missing from the table, stark_gen.rs:637-639)");
        cub::DeviceScan::ExclusiveSum(d_tmp, tmp_bytes, counts, offs, (int)n, st);
        k_iota<<<g, 256, 0, st>>>(fi, n);
        cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, idx_f, idx_s, fi, fi_s, (int)n, 0, 32, st);
        k_place_t<<<g, 256, 0, st>>>(pt, offs, n, o1, o2);
        k_place_f<<<g, 256, 0, st>>>(pf, idx_s, fi_s, n, o1, o2);
        launch_count_add(8 + 2 * dim);
        B200_CUDA_CHECK(cudaGetLastError());
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    } catch (...) { cudaFree(d_tmp); cudaFree(key_a); cudaFree(perm_a < perm_b ? perm_a : perm_b); throw; }
    cudaFree(d_tmp); cudaFree(key_a); cudaFree(perm_a < perm_b ? perm_a : perm_b);
}

}  // namespace b200
