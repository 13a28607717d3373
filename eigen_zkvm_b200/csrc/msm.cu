// BN254 G1 multi-scalar multiplication (Pippenger) for sm_100a.
//
// Boundary: the multiexp calls inside `bellman_ce::groth16::create_random_proof`, reached from
// `Groth16::prove` (groth16/src/groth16.rs:88-96) / `groth16_prove` (groth16/src/api.rs:144-177).  bellman's
// in-memory forms are kept: bases = affine (x, y) as 4 x u64 little-endian MONTGOMERY limbs (R = 2^256),
// scalars = canonical 4 x u64 `Repr`; the result is a Jacobian triple (X, Y, Z) in Montgomery form.
//
// Pipeline (all on the device):
//   1. signed c-bit window digits per scalar (buckets 1..2^(c-1), sign folded into the point index)   k_msm_digits
//   2. counting sort of point indices by bucket, per window (histogram -> scan -> scatter)              k_msm_scan / k_msm_scatter
//   3. one thread per (window, bucket): XYZZ accumulator += +-P over its contiguous index run           k_msm_accumulate
//   4. per window: sum_b b * B_b by two levels of segmented running sums + shared-memory tree           k_msm_reduce1/2
//   5. Horner over windows (c doublings each), normalisation to affine                                  k_msm_final
// Roofline class: INT-ALU (about 10 Fq products per mixed addition, 8x32-bit limb CIOS Montgomery);
// HBM traffic is 96 B per (point, scalar) plus 8 B per (point, window) of index traffic.
#include "b200_internal.h"
#include <cstring>

namespace b200 {

struct fq { u32 l[8]; };
// q = 21888242871839275222246405745257275088696311157297823662689037894645226208583 (groth16/src/api.rs:636)
__device__ __constant__ u32 FQ_Q[8] = {0xd87cfd47, 0x3c208c16, 0x6871ca8d, 0x97816a91, 0x8181585d, 0xb85045b6, 0xe131a029, 0x30644e72};
#define FQ_QINV 0xe4866389u   /* -q^-1 mod 2^32 */
// R = 2^256 mod q (Montgomery one)
__device__ __constant__ u32 FQ_R[8] = {0xc58f0d9d, 0xd35d438d, 0xf5c70b3d, 0x0a78eb28, 0x7879462c, 0x666ea36f, 0x9a07df2f, 0x0e0a77c1};
// 3 * R mod q (curve constant b = 3)
__device__ __constant__ u32 FQ_B3[8] = {0x50ad28d7, 0x7a17caa9, 0xe15521b9, 0x1f6ac17a, 0x696bd284, 0x334bea4e, 0xce179d8e, 0x2a1f6744};
// (q + 1) / 4, exponent of the square root (q = 3 mod 4)
__device__ __constant__ u32 FQ_SQRT_E[8] = {0xb61f3f52, 0x4f082305, 0x5a1c72a3, 0x65e05aa4, 0xa0605617, 0x6e14116d, 0xb84c680a, 0x0c19139c};

#define MSM_D __device__ __forceinline__

MSM_D bool fq_is_zero(const fq& a) { u32 o = 0; for (int i = 0; i < 8; i++) o |= a.l[i]; return o == 0; }
MSM_D bool fq_eq(const fq& a, const fq& b) { u32 o = 0; for (int i = 0; i < 8; i++) o |= a.l[i] ^ b.l[i]; return o == 0; }
MSM_D fq fq_zero() { fq r; for (int i = 0; i < 8; i++) r.l[i] = 0; return r; }
MSM_D fq fq_one() { fq r; for (int i = 0; i < 8; i++) r.l[i] = FQ_R[i]; return r; }
// r = a - q if a >= q
MSM_D fq fq_cond_sub(const fq& a) {
    fq t; u32 bw;
    asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\tsubc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\tsubc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\tsubc.u32 %8, 0, 0;"
        : "=r"(t.l[0]), "=r"(t.l[1]), "=r"(t.l[2]), "=r"(t.l[3]), "=r"(t.l[4]), "=r"(t.l[5]), "=r"(t.l[6]), "=r"(t.l[7]), "=r"(bw)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(FQ_Q[0]), "r"(FQ_Q[1]), "r"(FQ_Q[2]), "r"(FQ_Q[3]), "r"(FQ_Q[4]), "r"(FQ_Q[5]), "r"(FQ_Q[6]), "r"(FQ_Q[7]));
    fq r;
    for (int i = 0; i < 8; i++) r.l[i] = bw ? a.l[i] : t.l[i];
    return r;
}
MSM_D fq fq_add(const fq& a, const fq& b) {     // a, b < q < 2^254: no carry out of 256 bits
    fq s;
    asm("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, %23;"
        : "=r"(s.l[0]), "=r"(s.l[1]), "=r"(s.l[2]), "=r"(s.l[3]), "=r"(s.l[4]), "=r"(s.l[5]), "=r"(s.l[6]), "=r"(s.l[7])
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    return fq_cond_sub(s);
}
MSM_D fq fq_sub(const fq& a, const fq& b) {
    fq d; u32 bw;
    asm("sub.cc.u32 %0, %9, %17;\n\tsubc.cc.u32 %1, %10, %18;\n\tsubc.cc.u32 %2, %11, %19;\n\tsubc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\tsubc.cc.u32 %5, %14, %22;\n\tsubc.cc.u32 %6, %15, %23;\n\tsubc.cc.u32 %7, %16, %24;\n\tsubc.u32 %8, 0, 0;"
        : "=r"(d.l[0]), "=r"(d.l[1]), "=r"(d.l[2]), "=r"(d.l[3]), "=r"(d.l[4]), "=r"(d.l[5]), "=r"(d.l[6]), "=r"(d.l[7]), "=r"(bw)
        : "r"(a.l[0]), "r"(a.l[1]), "r"(a.l[2]), "r"(a.l[3]), "r"(a.l[4]), "r"(a.l[5]), "r"(a.l[6]), "r"(a.l[7]),
          "r"(b.l[0]), "r"(b.l[1]), "r"(b.l[2]), "r"(b.l[3]), "r"(b.l[4]), "r"(b.l[5]), "r"(b.l[6]), "r"(b.l[7]));
    // borrow: add q back (bw is all-ones then)
    fq r;
    asm("add.cc.u32 %0, %8, %16;\n\taddc.cc.u32 %1, %9, %17;\n\taddc.cc.u32 %2, %10, %18;\n\taddc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\taddc.cc.u32 %5, %13, %21;\n\taddc.cc.u32 %6, %14, %22;\n\taddc.u32 %7, %15, %23;"
        : "=r"(r.l[0]), "=r"(r.l[1]), "=r"(r.l[2]), "=r"(r.l[3]), "=r"(r.l[4]), "=r"(r.l[5]), "=r"(r.l[6]), "=r"(r.l[7])
        : "r"(d.l[0]), "r"(d.l[1]), "r"(d.l[2]), "r"(d.l[3]), "r"(d.l[4]), "r"(d.l[5]), "r"(d.l[6]), "r"(d.l[7]),
          "r"(FQ_Q[0] & bw), "r"(FQ_Q[1] & bw), "r"(FQ_Q[2] & bw), "r"(FQ_Q[3] & bw), "r"(FQ_Q[4] & bw), "r"(FQ_Q[5] & bw), "r"(FQ_Q[6] & bw), "r"(FQ_Q[7] & bw));
    return r;
}
MSM_D fq fq_neg(const fq& a) { return fq_is_zero(a) ? a : fq_sub(fq_zero(), a); }
MSM_D fq fq_dbl(const fq& a) { return fq_add(a, a); }
// CIOS Montgomery product, 8 x 32-bit limbs, 64-bit multiply-accumulate (IMAD.WIDE.U32)
MSM_D fq fq_mul(const fq& a, const fq& b) {
    u32 t[10];
#pragma unroll
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        u64 c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) { c = (u64)a.l[j] * b.l[i] + t[j] + c; t[j] = (u32)c; c >>= 32; }
        c += t[8]; t[8] = (u32)c; t[9] = (u32)(c >> 32);
        u32 m = t[0] * FQ_QINV;
        c = ((u64)m * FQ_Q[0] + t[0]) >> 32;
#pragma unroll
        for (int j = 1; j < 8; j++) { c = (u64)m * FQ_Q[j] + t[j] + c; t[j - 1] = (u32)c; c >>= 32; }
        c += t[8]; t[7] = (u32)c; t[8] = t[9] + (u32)(c >> 32);
    }
    fq r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = t[i];
    return fq_cond_sub(r);     // result < 2q < 2^255, so t[8] == 0
}
MSM_D fq fq_sqr(const fq& a) { return fq_mul(a, a); }
MSM_D fq fq_pow(const fq& a, const u32* e) {
    fq acc = fq_one();
    for (int i = 255; i >= 0; i--) { acc = fq_sqr(acc); if ((e[i >> 5] >> (i & 31)) & 1) acc = fq_mul(acc, a); }
    return acc;
}
MSM_D fq fq_inv(const fq& a) {     // a^(q-2)
    u32 e[8];
    for (int i = 0; i < 8; i++) e[i] = FQ_Q[i];
    e[0] -= 2;                      // q's low limb is > 2
    return fq_pow(a, e);
}

// ------------------------------------------------------------------------------------------------ curve: y^2 = x^3 + 3
struct affine { fq x, y; };                 // (0, 0) = infinity
struct xyzz { fq x, y, zz, zzz; };          // x = X/ZZ, y = Y/ZZZ; ZZ = 0 <=> infinity
MSM_D xyzz xyzz_inf() { xyzz p; p.x = fq_zero(); p.y = fq_zero(); p.zz = fq_zero(); p.zzz = fq_zero(); return p; }
MSM_D bool xyzz_is_inf(const xyzz& p) { return fq_is_zero(p.zz); }
MSM_D xyzz xyzz_dbl_affine(const fq& x, const fq& y) {      // mdbl-2008-s-1 (a = 0)
    fq u = fq_dbl(y), v = fq_sqr(u), w = fq_mul(u, v), s = fq_mul(x, v);
    fq x2 = fq_sqr(x), m = fq_add(fq_dbl(x2), x2);
    xyzz r;
    r.x = fq_sub(fq_sqr(m), fq_dbl(s));
    r.y = fq_sub(fq_mul(m, fq_sub(s, r.x)), fq_mul(w, y));
    r.zz = v; r.zzz = w;
    return r;
}
MSM_D xyzz xyzz_dbl(const xyzz& p) {                          // dbl-2008-s-1
    if (xyzz_is_inf(p)) return p;
    fq u = fq_dbl(p.y), v = fq_sqr(u), w = fq_mul(u, v), s = fq_mul(p.x, v);
    fq x2 = fq_sqr(p.x), m = fq_add(fq_dbl(x2), x2);
    xyzz r;
    r.x = fq_sub(fq_sqr(m), fq_dbl(s));
    r.y = fq_sub(fq_mul(m, fq_sub(s, r.x)), fq_mul(w, p.y));
    r.zz = fq_mul(v, p.zz); r.zzz = fq_mul(w, p.zzz);
    return r;
}
MSM_D xyzz xyzz_add_affine(const xyzz& p, const fq& x2, const fq& y2) {   // madd-2008-s; (x2, y2) finite
    if (xyzz_is_inf(p)) { xyzz r; r.x = x2; r.y = y2; r.zz = fq_one(); r.zzz = fq_one(); return r; }
    fq u2 = fq_mul(x2, p.zz), s2 = fq_mul(y2, p.zzz);
    fq pp_ = fq_sub(u2, p.x), rr = fq_sub(s2, p.y);
    if (fq_is_zero(pp_)) { if (fq_is_zero(rr)) return xyzz_dbl_affine(x2, y2); return xyzz_inf(); }
    fq pp = fq_sqr(pp_), ppp = fq_mul(pp_, pp), q = fq_mul(p.x, pp);
    xyzz r;
    r.x = fq_sub(fq_sub(fq_sqr(rr), ppp), fq_dbl(q));
    r.y = fq_sub(fq_mul(rr, fq_sub(q, r.x)), fq_mul(p.y, ppp));
    r.zz = fq_mul(p.zz, pp); r.zzz = fq_mul(p.zzz, ppp);
    return r;
}
MSM_D xyzz xyzz_add(const xyzz& p, const xyzz& q_) {                      // add-2008-s
    if (xyzz_is_inf(p)) return q_;
    if (xyzz_is_inf(q_)) return p;
    fq u1 = fq_mul(p.x, q_.zz), u2 = fq_mul(q_.x, p.zz), s1 = fq_mul(p.y, q_.zzz), s2 = fq_mul(q_.y, p.zzz);
    fq pp_ = fq_sub(u2, u1), rr = fq_sub(s2, s1);
    if (fq_is_zero(pp_)) { if (fq_is_zero(rr)) return xyzz_dbl(p); return xyzz_inf(); }
    fq pp = fq_sqr(pp_), ppp = fq_mul(pp_, pp), q = fq_mul(u1, pp);
    xyzz r;
    r.x = fq_sub(fq_sub(fq_sqr(rr), ppp), fq_dbl(q));
    r.y = fq_sub(fq_mul(rr, fq_sub(q, r.x)), fq_mul(s1, ppp));
    r.zz = fq_mul(fq_mul(p.zz, q_.zz), pp); r.zzz = fq_mul(fq_mul(p.zzz, q_.zzz), ppp);
    return r;
}
MSM_D xyzz xyzz_mul_small(xyzz p, u32 k) {     // k * p, double-and-add
    xyzz acc = xyzz_inf();
    while (k) { if (k & 1) acc = xyzz_add(acc, p); p = xyzz_dbl(p); k >>= 1; }
    return acc;
}

// ------------------------------------------------------------------------------------------------ kernels
// digits: dig[w * n + i] = bucket | sign << 31 (bucket 0 = nothing to add)
__global__ void k_msm_digits(const u32* __restrict__ scalars, const u32* __restrict__ bases, size_t n, u32 c, u32 nw, u32 nb,
                             u32* __restrict__ dig, u32* __restrict__ counts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    u32 s[9];
    for (int k = 0; k < 8; k++) s[k] = scalars[8 * i + k];
    s[8] = 0;
    u32 any = 0;
    for (int k = 0; k < 16; k++) any |= bases[16 * i + k];
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
        if (v) atomicAdd(&counts[(size_t)w * nb + v], 1u);
    }
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
__global__ void __launch_bounds__(128) k_msm_accumulate(const affine* __restrict__ bases, const u32* __restrict__ sorted, const u32* __restrict__ offsets,
                                                         const u32* __restrict__ counts, xyzz* __restrict__ buckets, size_t n, u32 nb) {
    u32 b = blockIdx.x * blockDim.x + threadIdx.x;
    u32 w = blockIdx.y;
    if (b >= nb) return;
    xyzz acc = xyzz_inf();
    if (b) {
        u32 start = offsets[(size_t)w * nb + b], cnt = counts[(size_t)w * nb + b];
        const u32* idx = sorted + (size_t)w * n + start;
        for (u32 k = 0; k < cnt; k++) {
            u32 e = idx[k];
            affine p = bases[e & 0x7fffffffu];
            if (e >> 31) p.y = fq_neg(p.y);
            acc = xyzz_add_affine(acc, p.x, p.y);
        }
    }
    buckets[(size_t)w * nb + b] = acc;
}
// window sum W = sum_{b=1}^{nb-1} b * B_b, two stages.  Buckets are cut into segments of RED_L; with b = s*RED_L + i
// (i in [1, RED_L]):  W = sum_s acc_s + RED_L * sum_s s * run_s,  run_s = sum_i B,  acc_s = sum_i i * B  (running sums).
#define RED_L 16
#define RED_T 128
MSM_D xyzz xyzz_neg(xyzz p) { p.y = fq_neg(p.y); return p; }
__global__ void __launch_bounds__(128) k_msm_reduce1(const xyzz* __restrict__ buckets, xyzz* __restrict__ seg_run, xyzz* __restrict__ seg_acc, u32 nb, u32 nseg) {
    u32 s = blockIdx.x * blockDim.x + threadIdx.x, w = blockIdx.y;
    if (s >= nseg) return;
    u32 lo = 1 + s * RED_L, hi = lo + RED_L < nb ? lo + RED_L : nb;
    xyzz run = xyzz_inf(), acc = xyzz_inf();
    // missing top buckets of a short last segment count as empty: start the running sum at the segment's nominal top
    for (u32 b = lo + RED_L; b-- > lo;) { if (b < hi) run = xyzz_add(run, buckets[(size_t)w * nb + b]); acc = xyzz_add(acc, run); }
    seg_run[(size_t)w * nseg + s] = run; seg_acc[(size_t)w * nseg + s] = acc;
}
__global__ void __launch_bounds__(RED_T) k_msm_reduce2(const xyzz* __restrict__ seg_run, const xyzz* __restrict__ seg_acc, xyzz* __restrict__ wsum, u32 nseg) {
    __shared__ xyzz sh[RED_T];
    u32 w = blockIdx.x, t = threadIdx.x;
    u32 G = (nseg + RED_T - 1) / RED_T, lo = t * G, hi = lo + G < nseg ? lo + G : nseg;
    xyzz A = xyzz_inf(), r = xyzz_inf(), a = xyzz_inf();
    for (u32 s = hi; s-- > lo;) { A = xyzz_add(A, seg_acc[(size_t)w * nseg + s]); r = xyzz_add(r, seg_run[(size_t)w * nseg + s]); a = xyzz_add(a, r); }
    // sum_s s * run_s over this thread's range = (a - r) + lo * r
    xyzz C = xyzz_add(a, xyzz_neg(r));
    if (lo && lo < hi) C = xyzz_add(C, xyzz_mul_small(r, lo));
    for (u32 k = 1; k < RED_L; k <<= 1) C = xyzz_dbl(C);      // * RED_L
    sh[t] = xyzz_add(A, C);
    __syncthreads();
    for (u32 st = RED_T / 2; st > 0; st >>= 1) { if (t < st) sh[t] = xyzz_add(sh[t], sh[t + st]); __syncthreads(); }
    if (t == 0) wsum[w] = sh[0];
}
// Horner over windows + affine normalisation; out = (X, Y, Z) Montgomery, Z = R (finite) or (0, R, 0)
__global__ void k_msm_final(const xyzz* __restrict__ wsum, u32 nw, u32 c, fq* __restrict__ out3) {
    if (threadIdx.x || blockIdx.x) return;
    xyzz tot = xyzz_inf();
    for (int w = (int)nw - 1; w >= 0; w--) { for (u32 k = 0; k < c; k++) tot = xyzz_dbl(tot); tot = xyzz_add(tot, wsum[w]); }
    if (xyzz_is_inf(tot)) { out3[0] = fq_zero(); out3[1] = fq_one(); out3[2] = fq_zero(); return; }
    fq zi = fq_inv(tot.zzz);                  // 1/z^3
    fq zinv = fq_mul(zi, tot.zz);             // z^2 / z^3 = 1/z
    fq zinv2 = fq_sqr(zinv);
    out3[0] = fq_mul(tot.x, zinv2);           // X / zz
    out3[1] = fq_mul(tot.y, zi);              // Y / zzz
    out3[2] = fq_one();
}

// out = a + b for two (X, Y, Z) Jacobian triples with Z in {0, R} or general: normalise through XYZZ
__global__ void k_g1_add(const fq* __restrict__ a3, const fq* __restrict__ b3, fq* __restrict__ out3) {
    if (threadIdx.x || blockIdx.x) return;
    xyzz p[2];
    const fq* in[2] = {a3, b3};
    for (int k = 0; k < 2; k++) {
        fq z = in[k][2];
        if (fq_is_zero(z)) { p[k] = xyzz_inf(); continue; }
        p[k].x = in[k][0]; p[k].y = in[k][1]; p[k].zz = fq_sqr(z); p[k].zzz = fq_mul(p[k].zz, z);
    }
    xyzz tot = xyzz_add(p[0], p[1]);
    if (xyzz_is_inf(tot)) { out3[0] = fq_zero(); out3[1] = fq_one(); out3[2] = fq_zero(); return; }
    fq zi = fq_inv(tot.zzz), zinv = fq_mul(zi, tot.zz);
    out3[0] = fq_mul(tot.x, fq_sqr(zinv)); out3[1] = fq_mul(tot.y, zi); out3[2] = fq_one();
}

// deterministic pseudo-random curve points (bench / tests): x from SplitMix64(seed, i, attempt), y = sqrt(x^3 + 3)
MSM_D u64 splitmix(u64 x) { x += 0x9E3779B97F4A7C15ULL; x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL; x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL; return x ^ (x >> 31); }
__global__ void __launch_bounds__(128) k_random_points(affine* __restrict__ out, size_t n, u64 seed) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (u32 attempt = 0;; attempt++) {
        fq x;
        for (int k = 0; k < 4; k++) { u64 v = splitmix(seed ^ splitmix((u64)i * 8 + k + ((u64)attempt << 48))); x.l[2 * k] = (u32)v; x.l[2 * k + 1] = (u32)(v >> 32); }
        x.l[7] &= 0x1fffffffu;               // < 2^253 < q: a valid Montgomery representative
        fq rhs = fq_add(fq_mul(fq_sqr(x), x), *reinterpret_cast<const fq*>(FQ_B3));
        fq y = fq_pow(rhs, FQ_SQRT_E);
        if (fq_eq(fq_sqr(y), rhs)) { out[i].x = x; out[i].y = y; return; }
    }
}

// ------------------------------------------------------------------------------------------------ host
static char* g_msm_ws[16] = {nullptr}; static size_t g_msm_ws_cap[16] = {0};
static char* msm_workspace(size_t bytes) {
    int dev = 0; B200_CUDA_CHECK(cudaGetDevice(&dev));
    if (g_msm_ws_cap[dev] < bytes) {
        if (g_msm_ws[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_msm_ws[dev])); g_msm_ws[dev] = nullptr; g_msm_ws_cap[dev] = 0; }
        B200_CUDA_CHECK(cudaMalloc(&g_msm_ws[dev], bytes)); g_msm_ws_cap[dev] = bytes;
    }
    return g_msm_ws[dev];
}
static void msm_run(const void* d_bases, const void* d_scalars, size_t n, void* h_out96) {
    if (n == 0) { u32 r[24]; memset(r, 0, sizeof r); B200_CUDA_CHECK(cudaMemcpyFromSymbol(r + 8, FQ_R, 32)); memcpy(h_out96, r, 96); return; }
    if (n >= (1ull << 31)) throw std::runtime_error("msm: n too large");
    const u32 c = n >= (1u << 18) ? 16 : (n >= (1u << 12) ? 12 : 8);
    const u32 nwin = (254 + 1 + c - 1) / c;                          // ceil(255 / c): 254 scalar bits + the signed-digit carry
    const u32 nb = (1u << (c - 1)) + 1;
    cudaStream_t st = stream();
    u32 *dig, *sorted, *counts, *offsets, *cursors; xyzz *buckets, *wsum; fq* d_out;
    const u32 nseg = (nb - 1 + RED_L - 1) / RED_L;
    xyzz *seg_run, *seg_acc;
    // one grow-only workspace per device (cudaMalloc/cudaFree per call cost far more than the kernels on multi-GPU hosts)
    const size_t b_idx = (size_t)nwin * n * 4, b_cnt = (size_t)nwin * nb * 4 * 3, b_pts = ((size_t)nwin * nb + nwin + 2 * (size_t)nwin * nseg) * sizeof(xyzz) + 256;
    auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
    char* ws = msm_workspace(al(b_idx) * 2 + al(b_cnt) + al(b_pts));
    dig = reinterpret_cast<u32*>(ws); sorted = reinterpret_cast<u32*>(ws + al(b_idx)); counts = reinterpret_cast<u32*>(ws + 2 * al(b_idx));
    offsets = counts + (size_t)nwin * nb; cursors = offsets + (size_t)nwin * nb;
    buckets = reinterpret_cast<xyzz*>(ws + 2 * al(b_idx) + al(b_cnt));
    wsum = buckets + (size_t)nwin * nb; seg_run = wsum + nwin; seg_acc = seg_run + (size_t)nwin * nseg; d_out = reinterpret_cast<fq*>(seg_acc + (size_t)nwin * nseg);
    B200_CUDA_CHECK(cudaMemsetAsync(counts, 0, (size_t)nwin * nb * 4, st));
    {
        ScopedTimer t("msm_digits", 32.0 * n);
        k_msm_digits<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((const u32*)d_scalars, (const u32*)d_bases, n, c, nwin, nb, dig, counts);
    }
    { ScopedTimer t("msm_sort", 8.0 * n * nwin); k_msm_scan<<<nwin, 1024, 0, st>>>(counts, offsets, cursors, nb);
      k_msm_scatter<<<dim3((unsigned)((n + 255) / 256), nwin), 256, 0, st>>>(dig, cursors, sorted, n, nb); }
    { ScopedTimer t("msm_accumulate", 96.0 * n);
      k_msm_accumulate<<<dim3((nb + 127) / 128, nwin), 128, 0, st>>>((const affine*)d_bases, sorted, offsets, counts, buckets, n, nb); }
    { ScopedTimer t("msm_reduce", 128.0 * nb * nwin); k_msm_reduce1<<<dim3((nseg + 127) / 128, nwin), 128, 0, st>>>(buckets, seg_run, seg_acc, nb, nseg);
      k_msm_reduce2<<<nwin, RED_T, 0, st>>>(seg_run, seg_acc, wsum, nseg); k_msm_final<<<1, 32, 0, st>>>(wsum, nwin, c, d_out); }
    launch_count_add(7);
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(h_out96, d_out, 96, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
}
void msm_bn254_g1_dev(const void* d_bases, const void* d_scalars, size_t n, void* h_out96) { msm_run(d_bases, d_scalars, n, h_out96); }
void msm_bn254_g1_host(const void* bases, const void* scalars, size_t n, void* h_out96) {
    // staging buffers are grow-only and per device, like the workspace
    static char* g_in[16] = {nullptr}; static size_t g_in_cap[16] = {0};
    int dev = 0; B200_CUDA_CHECK(cudaGetDevice(&dev));
    size_t need = (n ? n : 1) * 96;
    if (g_in_cap[dev] < need) { if (g_in[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_in[dev])); } B200_CUDA_CHECK(cudaMalloc(&g_in[dev], need)); g_in_cap[dev] = need; }
    char* db = g_in[dev]; char* ds = db + (n ? n : 1) * 64;
    B200_CUDA_CHECK(cudaMemcpyAsync(db, bases, n * 64, cudaMemcpyHostToDevice, stream()));
    B200_CUDA_CHECK(cudaMemcpyAsync(ds, scalars, n * 32, cudaMemcpyHostToDevice, stream()));
    msm_run(db, ds, n, h_out96);
}
void bn254_g1_add_host(const void* a96, const void* b96, void* out96) {
    fq* d; B200_CUDA_CHECK(cudaMalloc(&d, 9 * sizeof(fq)));
    B200_CUDA_CHECK(cudaMemcpyAsync(d, a96, 96, cudaMemcpyHostToDevice, stream())); B200_CUDA_CHECK(cudaMemcpyAsync(d + 3, b96, 96, cudaMemcpyHostToDevice, stream()));
    k_g1_add<<<1, 32, 0, stream()>>>(d, d + 3, d + 6); launch_count_add(1);
    B200_CUDA_CHECK(cudaMemcpyAsync(out96, d + 6, 96, cudaMemcpyDeviceToHost, stream())); B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    cudaFree(d);
}
void bn254_g1_random_points_dev(void* d_bases, size_t n, u64 seed) {
    k_random_points<<<(unsigned)((n + 127) / 128), 128, 0, stream()>>>((affine*)d_bases, n, seed); launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError()); B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}

}  // namespace b200
