// Batched Goldilocks NTT / iNTT / coset LDE over column-major device matrices.
// Reference semantics: starky/src/fft_p.rs:174-355 (`fft`, `ifft`, `interpolate`), roots from
// starky/src/constant.rs:52-68 (w_{2^k} = 7^((p-1)/2^k), coset shift 49).  Natural order in, natural order out.
//
// Design (B200): a size-2^k transform is 1..3 HBM passes (k <= 9 / 18 / 27).  Each pass is a mixed-radix Cooley-Tukey step over one
// index digit of up to 9 bits, done as one or two register-resident shift-only radix-8/16/32 rounds with one shared-memory exchange
// between them, followed by the inter-pass twiddle.  The last pass tiles over the *first* output digit instead, which folds the
// digit-reversal permutation into its (still coalesced) store.  No bit-reversal pass, no transposes.
//   k >= 12: k_ntt3 -- tiles staged by TMA (cp.async.bulk.tensor + mbarrier), twiddles streamed as a tile of a table laid out like the
//            output, TMA store; the LDE's zero padding is the tensor map's out-of-bounds fill.
//   k <  12 (and unaligned buffers, or B200_NTT_TMA=0): k_ntt2 -- register-staged loads and stores, twiddles gathered from a power table.
// Fusions: 1/N and the coset factor 49^i ride on the iNTT's last store.
// Roofline class: HBM by traffic (16 B per element per pass), ALU-pipe bound in practice; see DESIGN.md 3.1.
#include "b200_internal.h"
#include "field.cuh"
#include <cuda.h>      // CUtensorMap types only; the encode entry point is looked up through the runtime (no libcuda link)
#include <cstdlib>
#include <map>
#include <tuple>
#include <mutex>

namespace b200 {

// ------------------------------------------------------------------------------------------------ host field helpers
static inline u64 hred(unsigned __int128 x) { return (u64)(x % GL_P); }
u64 h_mul(u64 a, u64 b) { return hred((unsigned __int128)a * b); }
u64 h_add(u64 a, u64 b) { return hred((unsigned __int128)a + b); }
u64 h_sub(u64 a, u64 b) { return hred((unsigned __int128)a + GL_P - b); }
u64 h_pow(u64 a, u64 e) { u64 r = 1; while (e) { if (e & 1) r = h_mul(r, a); a = h_mul(a, a); e >>= 1; } return r; }
u64 h_inv(u64 a) { return h_pow(a, GL_P - 2); }
static u64 g_w[33], g_wi[33]; static std::once_flag g_roots_once;
static void roots_init() {
    std::call_once(g_roots_once, [] {
        g_w[32] = h_pow(7, 0xFFFFFFFFULL); g_wi[32] = h_inv(g_w[32]);
        for (int n = 31; n >= 0; n--) { g_w[n] = h_mul(g_w[n + 1], g_w[n + 1]); g_wi[n] = h_mul(g_wi[n + 1], g_wi[n + 1]); }
    });
}
u64 h_root(unsigned k) { roots_init(); return g_w[k]; }
u64 h_root_inv(unsigned k) { roots_init(); return g_wi[k]; }

// ------------------------------------------------------------------------------------------------ stream
int current_device() {
    int dev = 0; B200_CUDA_CHECK(cudaGetDevice(&dev));
    if (dev < 0 || dev >= B200_MAX_DEVICES) throw std::runtime_error("device index " + std::to_string(dev) + " is outside the library's per-device tables (B200_MAX_DEVICES)");
    return dev;
}
static cudaStream_t g_stream[B200_MAX_DEVICES] = {0};
cudaStream_t stream() { int dev = 0; if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= B200_MAX_DEVICES) return 0; return g_stream[dev]; }
void set_stream(cudaStream_t s) { g_stream[current_device()] = s; }
static std::mutex g_tab_mu;                 // guards the table caches below (shared between devices, keyed by device)

// ------------------------------------------------------------------------------------------------ power tables
__global__ void k_powtab(u64* lo, u64* hi, u64 base, u64 scale, u32 n_lo, u32 n_hi) {
    u32 i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_lo) lo[i] = gl_mul(scale, gl_pow(base, i));
    else if (i < n_lo + n_hi) hi[i - n_lo] = gl_pow(base, (u64)(i - n_lo) << POW_LO_BITS);
}
static std::map<std::tuple<int, u64, unsigned, u64>, DevPowTab> g_powtabs;
static DevPowTab powtab_scaled(u64 base, unsigned log_range, u64 scale) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    int dev = current_device();
    auto key = std::make_tuple(dev, base, log_range, scale);
    auto it = g_powtabs.find(key);
    if (it != g_powtabs.end()) return it->second;
    u32 n_lo = 1u << POW_LO_BITS, n_hi = log_range > POW_LO_BITS ? 1u << (log_range - POW_LO_BITS) : 1u;
    u64* p; B200_CUDA_CHECK(cudaMalloc(&p, (size_t)(n_lo + n_hi) * 8));
    k_powtab<<<(n_lo + n_hi + 255) / 256, 256, 0, stream()>>>(p, p + n_lo, base, scale, n_lo, n_hi);
    B200_CUDA_CHECK(cudaGetLastError());
    DevPowTab t{p, p + n_lo};
    g_powtabs[key] = t;
    return t;
}
DevPowTab powtab(u64 base, unsigned log_range) { return powtab_scaled(base, log_range, 1); }

// ------------------------------------------------------------------------------------------------ transposes
// 32x32 tiles through shared memory; both sides coalesced.  in: rows_in x cols_in row-major -> out: cols_in x rows_in.
// The longer dimension rides on gridDim.x (2^31 limit), the shorter on gridDim.y (65535 limit).
__global__ void k_transpose(const u64* __restrict__ in, u64* __restrict__ out, size_t n_rows_in, size_t n_cols_in, int rows_on_x) {
    __shared__ u64 tile[32][33];
    size_t r0 = (size_t)(rows_on_x ? blockIdx.x : blockIdx.y) * 32, c0 = (size_t)(rows_on_x ? blockIdx.y : blockIdx.x) * 32;
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        size_t r = r0 + j, c = c0 + threadIdx.x;
        if (r < n_rows_in && c < n_cols_in) tile[j][threadIdx.x] = gl_canon(in[r * n_cols_in + c]);     // ingest reduces like FGL::from (field_gl.rs), free on an HBM-bound kernel
    }
    __syncthreads();
    for (int j = threadIdx.y; j < 32; j += blockDim.y) {
        size_t c = c0 + j, r = r0 + threadIdx.x;
        if (r < n_rows_in && c < n_cols_in) out[c * n_rows_in + r] = tile[threadIdx.x][j];
    }
}
// Narrow matrices (the Fibonacci trace is N x 2): one thread per row, the row read as 16-byte vectors, every column written
// coalesced -- no shared-memory tile.  W = row length, in: rows x W row-major -> out: W x rows.
template <int W> __global__ void __launch_bounds__(256) k_transpose_narrow(const u64* __restrict__ in, u64* __restrict__ out, size_t rows) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    u64 v[W];
    if (W % 2 == 0) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + r * W);
#pragma unroll
        for (int c = 0; c < W / 2; c++) { ulonglong2 q = p[c]; v[2 * c] = gl_canon(q.x); v[2 * c + 1] = gl_canon(q.y); }
    } else {
#pragma unroll
        for (int c = 0; c < W; c++) v[c] = gl_canon(in[r * W + c]);
    }
#pragma unroll
    for (int c = 0; c < W; c++) out[(size_t)c * rows + r] = v[c];
}
// the inverse: in: W x rows -> out: rows x W row-major
template <int W> __global__ void __launch_bounds__(256) k_untranspose_narrow(const u64* __restrict__ in, u64* __restrict__ out, size_t rows) {
    size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    u64 v[W];
#pragma unroll
    for (int c = 0; c < W; c++) v[c] = in[(size_t)c * rows + r];
    if (W % 2 == 0) {
        ulonglong2* p = reinterpret_cast<ulonglong2*>(out + r * W);
#pragma unroll
        for (int c = 0; c < W / 2; c++) p[c] = make_ulonglong2(v[2 * c], v[2 * c + 1]);
    } else {
#pragma unroll
        for (int c = 0; c < W; c++) out[r * W + c] = v[c];
    }
}
template <int W> static void transpose_narrow(const u64* in, u64* out, size_t rows, bool to_colmajor) {
    unsigned blocks = (unsigned)((rows + 255) / 256);
    if (to_colmajor) k_transpose_narrow<W><<<blocks, 256, 0, stream()>>>(in, out, rows);
    else k_untranspose_narrow<W><<<blocks, 256, 0, stream()>>>(in, out, rows);
}
static void transpose_any(const u64* in, u64* out, size_t rows_in, size_t cols_in) {
    if (rows_in == 0 || cols_in == 0) return;
    ScopedTimer t("transpose", 16.0 * (double)rows_in * (double)cols_in);
    // rows x W with W <= 4 (either orientation): the narrow kernels
    const bool tall = cols_in <= 4 && rows_in > 4, wide = rows_in <= 4 && cols_in > 4;
    if (tall || wide) {
        size_t W = tall ? cols_in : rows_in, rows = tall ? rows_in : cols_in;
        switch (W) {
        case 1: transpose_narrow<1>(in, out, rows, tall); break;
        case 2: transpose_narrow<2>(in, out, rows, tall); break;
        case 3: transpose_narrow<3>(in, out, rows, tall); break;
        default: transpose_narrow<4>(in, out, rows, tall); break;
        }
        launch_count_add(1);
        B200_CUDA_CHECK(cudaGetLastError());
        return;
    }
    dim3 block(32, 8);
    size_t gr = (rows_in + 31) / 32, gc = (cols_in + 31) / 32;
    int rows_on_x = gr >= gc;
    size_t gx = rows_on_x ? gr : gc, gy = rows_on_x ? gc : gr;
    if (gy > 65535) throw std::runtime_error("transpose: matrix too large in both dimensions");
    k_transpose<<<dim3((unsigned)gx, (unsigned)gy), block, 0, stream()>>>(in, out, rows_in, cols_in, rows_on_x);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}
void transpose_rm_to_cm(const u64* d_in, u64* d_out, size_t rows, size_t w) { transpose_any(d_in, d_out, rows, w); }
void transpose_cm_to_rm(const u64* d_in, u64* d_out, size_t rows, size_t w) { transpose_any(d_in, d_out, w, rows); }

// ------------------------------------------------------------------------------------------------ NTT pass
// One pass = one index digit of r = RA + RB bits, done as one or two register-resident radix-2^RA / 2^RB rounds.
//
// Inside a round every twiddle is a power of two: 2 has multiplicative order 192 in GL (2^96 = -1), so
// w'_n = 2^(192/n) is a primitive n-th root for n | 64 and a size-n DFT with root w'_n needs only shifts.  The
// reference's root w_n = 7^((p-1)/n) (constant.rs:54-68) is an odd power (w'_n)^k, hence
//     X_ref[j] = X'[k*j mod n]:
// the shift-DFT computes the same numbers, and the register that holds X'[brev(p)] simply gets the reference
// frequency label perm[p] = k^-1 * brev(p) mod n (table built on the host, used for addresses and twiddles only).
// Between the two rounds one full product with w_R^(k1*d0) (shared-memory table) and one shared-memory exchange;
// after the pass the inter-pass twiddle w_Nj^(kd*low) from the two-level power table.  Global loads and stores go
// straight from/to registers in runs of T consecutive elements.
// All within-column offsets are < 2^27 (split_digits rejects larger transforms), so the index arithmetic is 32-bit: the
// 64-bit multiplies and compares it replaces were ~10 % of the instructions of a pass (profiles/ncu_r1c.md).
struct Pass2 {
    u32 T;            // tile width (power of two)
    u32 n;            // transform size
    u32 S, Nj;        // non-last: blockIdx.x = hi*(S/T) + lowtile ; addr = hi*Nj + d*S + lowtile*T + t
    u32 R1, M;        // last: blockIdx.x = mid*(R1/T) + atile ; read = (atile*T+t)*(n/R1) + mid*R + d ; write = (atile*T+t) + R1*mid + R1*M*kd
    u32 n_in;         // valid input length (elements >= n_in read as 0)
    const u64* twR;   // w_R^e, e < R (direction aware)
    PowTab tw;        // inter-pass twiddle base w_Nj (non-last), two-level table
    const u64* tw1;   // the same as ONE table of Nj entries when Nj <= 2^20 (L2 resident): one load instead of two loads + a product
    PowTab post;      // last pass: X[k] *= post^k (scale folded in) when has_post
    u32 has_post;
    u64 post_scale;   // last pass: constant factor (1 = none) when !has_post
    unsigned char permA[32], permB[32];
};

// x * 2^e mod p for a compile-time e in [0, 96); x canonical
template <int E> GL_D u64 gl_mul_pow2(u64 x) {
    static_assert(E >= 0 && E < 96, "shift out of range");
    if constexpr (E == 0) return x;
    else if constexpr (E < 32) return gl_red96(x << E, (u32)(x >> (64 - E)));
    else if constexpr (E == 32) return gl_red96(x << 32, (u32)(x >> 32));
    else if constexpr (E < 64) return gl_red128(x << E, x >> (64 - E));
    else if constexpr (E == 64) return gl_red128(0, x);
    else {
        // 64 < E < 96: x*2^(E-64) = yhi*2^64 + ylo ;  times 2^64:  ylo*2^64 + yhi*2^128,  2^128 = -2^32 (mod p)
        u64 ylo = x << (E - 64); u32 yhi = (u32)(x >> (128 - E));
        return gl_sub(gl_red128(0, ylo), (u64)yhi << 32);
    }
}
// stage with half = 2^SH of a size-2^A DIF: v' = (u - v) * 2^(96*j/half)
template <int A, int SH, int PI> struct ShiftStage {
    GL_D static void run(u64* x) {
        constexpr int half = 1 << SH, j = PI & (half - 1), p = ((PI >> SH) << (SH + 1)) | j;
        u64 u = x[p], v = x[p + half];
        x[p] = gl_add(u, v);
        x[p + half] = gl_mul_pow2<(96 * j) / half>(gl_sub(u, v));
        ShiftStage<A, SH, PI + 1>::run(x);
    }
};
template <int A, int SH> struct ShiftStage<A, SH, (1 << A) / 2> { GL_D static void run(u64*) {} };
template <int A, int SH> struct ShiftDft { GL_D static void run(u64* x) { ShiftStage<A, SH, 0>::run(x); ShiftDft<A, SH - 1>::run(x); } };
template <int A> struct ShiftDft<A, -1> { GL_D static void run(u64*) {} };
// in: natural order; out: register p holds X'[brev_A(p)] for the root 2^(192/2^A)
template <int A> GL_D void shift_dft(u64* x) { ShiftDft<A, A - 1>::run(x); }

template <int RA, int RB, bool LAST>
__global__ void __launch_bounds__(256) k_ntt2(const u64* __restrict__ in, u64* __restrict__ out, u64 col_stride_in, u64 col_stride_out, Pass2 pp) {
    constexpr int NA = 1 << RA, NB = 1 << RB, R = NA * NB;
    extern __shared__ u64 sm[];
    const u32 T = pp.T, TP = T + 1;
    u64* twR = sm + (size_t)R * TP;
    const u64* src = in + (u64)blockIdx.y * col_stride_in;
    u64* dst = out + (u64)blockIdx.y * col_stride_out;
    const u32 tid = threadIdx.x;
    if (RB > 0) for (u32 e = tid; e < (u32)R; e += blockDim.x) twR[e] = pp.twR[e];

    u32 hi = 0, low0 = 0, mid = 0, a0 = 0;
    if (!LAST) { u32 tiles_per_hi = pp.S / T; hi = blockIdx.x / tiles_per_hi; low0 = (blockIdx.x % tiles_per_hi) * T; }
    else { u32 tiles = pp.R1 / T; mid = blockIdx.x / tiles; a0 = (blockIdx.x % tiles) * T; }
    const u32 rowlen = pp.n / pp.R1;

    u64 x[NA];
    u32 t = 0, d0 = 0;
    const bool actA = tid < (u32)NB * T;
    if (actA) {
        if (!LAST) { d0 = tid / T; t = tid % T; } else { t = tid / NB; d0 = tid % NB; }
#pragma unroll
        for (int m = 0; m < NA; m++) {
            u32 d = m * NB + d0;
            u32 a = LAST ? (a0 + t) * rowlen + mid * R + d : hi * pp.Nj + d * pp.S + low0 + t;
            x[m] = a < pp.n_in ? __ldg(src + a) : 0;
        }
        shift_dft<RA>(x);
    }
    if (RB > 0) {
        __syncthreads();      // twR (filled above by the first R threads) must be visible before any warp multiplies by it:
                              // without this barrier a fast warp could read a stale entry (seen once in ~10 full test runs)
        if (actA) {
#pragma unroll
            for (int m = 0; m < NA; m++) {
                u32 k1 = pp.permA[m];
                u32 e = (k1 * d0) & (R - 1);
                u64 v = x[m];
                v = gl_mul(v, twR[e]);
                sm[(size_t)(k1 * NB + d0) * TP + t] = v;
            }
        }
        __syncthreads();
        if (tid >= (u32)NA * T) return;
        const u32 k1 = tid / T; t = tid % T;
        u64 y[NB > 1 ? NB : 1];
#pragma unroll
        for (int m = 0; m < NB; m++) y[m] = sm[(size_t)(k1 * NB + m) * TP + t];
        shift_dft<RB>(y);
#pragma unroll
        for (int m = 0; m < NB; m++) {
            u32 kd = k1 + NA * (u32)pp.permB[m];
            u64 v = y[m];
            if (!LAST) {
                u32 low = low0 + t;
                v = gl_mul(v, pp.tw1 ? __ldg(pp.tw1 + kd * low) : powtab_getw(pp.tw, kd * low));
                dst[hi * pp.Nj + kd * pp.S + low] = v;
            } else {
                u32 k = (a0 + t) + pp.R1 * mid + pp.R1 * pp.M * kd;
                if (pp.has_post) v = gl_mul(v, powtab_getw(pp.post, k));
                else if (pp.post_scale != 1) v = gl_mul(v, pp.post_scale);
                dst[k] = v;
            }
        }
    } else {
        if (!actA) return;
#pragma unroll
        for (int m = 0; m < NA; m++) {
            u32 kd = pp.permA[m];
            u64 v = x[m];
            if (!LAST) {
                u32 low = low0 + t;
                v = gl_mul(v, pp.tw1 ? __ldg(pp.tw1 + kd * low) : powtab_getw(pp.tw, kd * low));
                dst[hi * pp.Nj + kd * pp.S + low] = v;
            } else {
                u32 k = (a0 + t) + pp.R1 * mid + pp.R1 * pp.M * kd;
                if (pp.has_post) v = gl_mul(v, powtab_getw(pp.post, k));
                else if (pp.post_scale != 1) v = gl_mul(v, pp.post_scale);
                dst[k] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------ NTT pass, TMA-staged (r3)
// Same mathematics as k_ntt2 (one index digit of RA + RB bits per pass, shift-only radix-2^RA / 2^RB rounds), different data
// movement.  A CTA owns one tile of R rows x T consecutive positions (at most 2048 elements, 16 KB).  One thread issues
// `cp.async.bulk.tensor` loads (TMA) for the data tile and for the tile of inter-pass twiddles -- the twiddles w_Nj^(kd*low) are kept as
// a table with the SAME [kd][low] layout as the output, so their tile is a box at the output's coordinates, streamed once per
// CTA instead of gathered 8 bytes at a time from 32-byte sectors.  The zero padding of the LDE is the tensor map's out-of-bounds
// fill.  Results go to shared memory and leave with one TMA store.  With the shift-DFT output permutations as compile-time
// constants every shared-memory address is `thread base + immediate`: no per-element address arithmetic is left (k_ntt2 spent
// ~80 of its 218 instructions per element-pass on it, profiles/ncu_r2a.md).  Buffer A holds the data tile, then (in place) the
// exchange between the two rounds; buffer B holds the twiddle tile and is overwritten element by element with the output.
// Tile = R rows x T consecutive positions, T = min(8, NTT3_TILE / R): 64-byte rows for R <= 256 (tiles of 512 .. 2048 elements), 32-byte
// rows for R = 512.  Measured on B200 (tools/gpu_runs/_run25.sh, ms per pass of 2^24 x 2 / 2^25 x 2 / 2^21 x 64): 4096-element tiles 0.196 / 0.439 /
// 0.765, 2048-element tiles 0.190 / 0.429 / 0.708, T = 8 everywhere 0.190 / 0.433 / 0.680, T = 4 everywhere 0.193 / 0.433 / 1.02: the same
// number of resident threads in finer grains overlaps the load, compute and store phases of different CTAs better, as long as a row stays
// a whole 32-byte sector (and 64 bytes where the tile allows).  NTT3_T > 0 forces T.
#ifndef NTT3_TILE
#define NTT3_TILE 2048
#endif
#ifndef NTT3_T
#define NTT3_T 0
#endif
#define NTT3_TSEL(R) (NTT3_T > 0 ? NTT3_T : (NTT3_TILE / (R) < 8 ? NTT3_TILE / (R) : 8))
GL_HD constexpr u32 brev_c(u32 p, int a) { u32 b = 0; for (int i = 0; i < a; i++) if (p & (1u << i)) b |= 1u << (a - 1 - i); return b; }
// perm[p] = kinv * brev(p) mod 2^a with root_ref(2^a) = (2^(192/2^a))^k, kinv = k^-1 (checked against shift_perm() at first use)
GL_HD constexpr u32 shift_kinv(int a, bool inv) {
    return a <= 1 ? 1u : (!inv ? (a == 2 ? 1u : 5u) : (a == 2 ? 3u : (a == 3 ? 3u : (a == 4 ? 11u : 27u))));
}
template <int A, bool INV> GL_HD constexpr u32 shift_perm_c(u32 p) { return A == 0 ? 0u : (shift_kinv(A, INV) * brev_c(p, A)) & ((1u << A) - 1); }

GL_D u32 smem_addr(const void* p) { return (u32)__cvta_generic_to_shared(p); }
GL_D void mbar_init(u64* bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory"); }
GL_D void mbar_expect_tx(u64* bar, u32 bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory"); }
GL_D void mbar_wait(u64* bar, u32 phase) {
    asm volatile("{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}" ::"r"(smem_addr(bar)), "r"(phase) : "memory");
}
GL_D void tma_load3(void* dst, const void* desc, u64* bar, u32 c0, u32 c1, u32 c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_addr(dst)), "l"((u64)desc), "r"(smem_addr(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
GL_D void tma_store3(const void* desc, const void* src, u32 c0, u32 c1, u32 c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((u64)desc), "r"(smem_addr(src)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

struct Pass3 {
    u32 log_tiles;    // non-last: log2(S / T) (blockIdx.y = hi * (S/T) + lowtile); last: log2(R1 / T) (blockIdx.y = mid * (R1/T) + atile); blockIdx.x = column
    u32 Rprod;        // non-last: number of `hi` values per column (tensor coordinate 2 = column * Rprod + hi)
    u32 R1;           // last: first digit's radix
    u64 scale;        // last pass without a post table: constant factor (1 = none)
    const u64* twr;   // [k1][d0] -> w_R^(k1 d0)
};

template <int RA, int RB> struct Ntt3Cfg {
    static constexpr int NA = 1 << RA, NB = 1 << RB, R = NA * NB, T = NTT3_TSEL(R), TILE = T * R, RW = R > 256 ? 256 : R, NBOX = R / RW;
    static constexpr int NT = TILE / (NA < NB ? NA : NB);
    __host__ __device__ static constexpr size_t a_words(bool last) { return last ? (size_t)R * (T + 1) + 15 & ~(size_t)15 : (size_t)TILE; }
    __host__ __device__ static constexpr size_t smem(bool last) { return (a_words(last) + TILE + 2) * 8; }
};

// TW: a table tile multiplies the output (inter-pass twiddles, or the last pass's post factors scale * g^k)
template <int RA, int RB, bool LAST, bool INV, bool TW>
__global__ void __launch_bounds__(Ntt3Cfg<RA, RB>::NT) k_ntt3(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_tw,
                                                              const __grid_constant__ CUtensorMap tm_out, Pass3 pp) {
    typedef Ntt3Cfg<RA, RB> C;
    constexpr int NA = C::NA, NB = C::NB, R = C::R, T = C::T, RW = C::RW, NBOX = C::NBOX, NT = C::NT;
    constexpr int TP = LAST ? T + 1 : T;
    extern __shared__ __align__(1024) unsigned char sm_raw[];
    u64* A = reinterpret_cast<u64*>(sm_raw);
    u64* B = A + C::a_words(LAST);
    u64* bars = B + C::TILE;
    const u32 tid = threadIdx.x;

    u32 cin0, cin1, cin2, cout0, cout2;
    if (!LAST) {
        const u32 hi = blockIdx.y >> pp.log_tiles, low0 = (blockIdx.y & ((1u << pp.log_tiles) - 1)) * T;
        cin0 = low0; cin1 = 0; cin2 = blockIdx.x * pp.Rprod + hi; cout0 = low0; cout2 = cin2;
    } else {
        const u32 mid = blockIdx.y >> pp.log_tiles, a0 = (blockIdx.y & ((1u << pp.log_tiles) - 1)) * T;
        cin0 = mid * R; cin1 = a0; cin2 = blockIdx.x; cout0 = a0 + pp.R1 * mid; cout2 = blockIdx.x;
    }
    if (tid == 0) {
        mbar_init(bars, 1); mbar_init(bars + 1, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(bars, C::TILE * 8);
#pragma unroll
        for (int b = 0; b < NBOX; b++) {
            if (!LAST) tma_load3(A + b * RW * T, &tm_in, bars, cin0, cin1 + b * RW, cin2);
            else tma_load3(A + b * RW * T, &tm_in, bars, cin0 + b * RW, cin1, cin2);
        }
        if (TW) {
            mbar_expect_tx(bars + 1, C::TILE * 8);
#pragma unroll
            for (int b = 0; b < NBOX; b++) tma_load3(B + b * RW * T, &tm_tw, bars + 1, cout0, b * RW, 0);
        }
    }
    mbar_wait(bars, 0);

    // ---- round A: thread (d0, t) holds x[m] = tile[d = m * NB + d0][t]
    constexpr bool ALL_A = (NB * T == NT);
    const bool actA = ALL_A || tid < (u32)(NB * T);
    u64 x[NA];
    u32 d0 = 0, tA = 0;
    if (!LAST) { d0 = tid / T; tA = tid % T; } else { tA = tid / NB; d0 = tid % NB; }
    if (actA) {
#pragma unroll
        for (int m = 0; m < NA; m++) {
            if (!LAST) x[m] = A[m * NB * T + tid];
            else x[m] = A[((m * NB) / RW) * T * RW + ((m * NB) % RW) + tA * RW + d0];
        }
    }
    __syncthreads();                      // every thread has its inputs: the tile buffer becomes the exchange buffer
    if (actA) {
        shift_dft<RA>(x);
        const u64* twr = pp.twr + d0;
#pragma unroll
        for (int m = 0; m < NA; m++) {
            const u32 k1 = shift_perm_c<RA, INV>(m);
            u64 v = x[m];
            if (k1 != 0) v = gl_mul(v, __ldg(twr + k1 * NB));
            if (!LAST) A[k1 * NB * T + tid] = v;
            else A[(k1 * NB + d0) * TP + tA] = v;
        }
    }
    __syncthreads();
    // ---- round B: thread (k1, t) holds y[m] = exchange[k1 * NB + m][t]; output row kd = k1 + NA * perm(m)
    constexpr bool ALL_B = (NA * T == NT);
    if (ALL_B || tid < (u32)(NA * T)) {
        const u32 k1 = tid / T, t = tid % T;
        u64 y[NB];
#pragma unroll
        for (int m = 0; m < NB; m++) y[m] = A[(k1 * NB + m) * TP + t];
        shift_dft<RB>(y);
        if (TW) mbar_wait(bars + 1, 0);
#pragma unroll
        for (int m = 0; m < NB; m++) {
            const u32 o = tid + NA * T * shift_perm_c<RB, INV>(m);
            u64 v = y[m];
            if (TW) v = gl_mul(v, B[o]);
            else if (LAST && pp.scale != 1) v = gl_mul(v, pp.scale);
            B[o] = v;
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0) {
#pragma unroll
        for (int b = 0; b < NBOX; b++) tma_store3(&tm_out, B + b * RW * T, cout0, b * RW, cout2);
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}

// w_R^e tables (e < R), cached per (r, inverse, device)
__global__ void k_root_tab(u64* out, u64 w, u32 n) { u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) out[i] = gl_pow(w, i); }
static std::map<std::tuple<int, unsigned, bool>, const u64*> g_root_tab;
static const u64* root_tab(unsigned r, bool inverse) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    int dev = current_device();
    auto key = std::make_tuple(dev, r, inverse);
    auto it = g_root_tab.find(key);
    if (it != g_root_tab.end()) return it->second;
    u32 n = 1u << r;
    u64* p; B200_CUDA_CHECK(cudaMalloc(&p, (size_t)n * 8));
    k_root_tab<<<(n + 255) / 256, 256, 0, stream()>>>(p, inverse ? h_root_inv(r) : h_root(r), n);
    B200_CUDA_CHECK(cudaGetLastError());
    g_root_tab[key] = p;
    return p;
}
// full tables w^e, e < 2^log_range, for log_range <= FULLTAB_MAX_BITS; cached per (root, device)
#ifndef FULLTAB_MAX_BITS
#define FULLTAB_MAX_BITS 20
#endif
static std::map<std::tuple<int, u64, unsigned>, const u64*> g_full_tab;
static const u64* full_tab(u64 w, unsigned log_range) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    int dev = current_device();
    auto key = std::make_tuple(dev, w, log_range);
    auto it = g_full_tab.find(key);
    if (it != g_full_tab.end()) return it->second;
    u32 n = 1u << log_range;
    u64* p; B200_CUDA_CHECK(cudaMalloc(&p, (size_t)n * 8));
    k_root_tab<<<(n + 255) / 256, 256, 0, stream()>>>(p, w, n);
    B200_CUDA_CHECK(cudaGetLastError());
    g_full_tab[key] = p;
    return p;
}
// perm[p] = k^-1 * brev_a(p) mod 2^a where root_ref(a) = (2^(192/2^a))^k
static void shift_perm(unsigned a, bool inverse, unsigned char* perm) {
    if (a == 0) { perm[0] = 0; return; }
    const u32 n = 1u << a;
    u64 base = h_pow(2, 192 / n), target = inverse ? h_root_inv(a) : h_root(a);
    u32 k = 0; u64 acc = 1;
    for (k = 0; k < n; k++) { if (acc == target) break; acc = h_mul(acc, base); }
    if (k == n) throw std::runtime_error("ntt: reference root is not a power of two root");
    u32 kinv = 1; for (u32 c = 1; c < n; c += 2) if ((c * k) % n == 1) { kinv = c; break; }
    for (u32 p = 0; p < n; p++) { u32 b = 0; for (u32 i = 0; i < a; i++) if (p & (1u << i)) b |= 1u << (a - 1 - i); perm[p] = (unsigned char)((kinv * b) % n); }
}

// scratch (grow-only, per device)
static u64* g_scratch[16] = {nullptr}; static size_t g_scratch_cap[16] = {0};
static u64* scratch(size_t n_u64) {
    int dev = current_device();
    if (g_scratch_cap[dev] < n_u64) {
        if (g_scratch[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_scratch[dev])); }
        B200_CUDA_CHECK(cudaMalloc(&g_scratch[dev], n_u64 * 8));
        g_scratch_cap[dev] = n_u64;
    }
    return g_scratch[dev];
}

static void split_digits(unsigned k, unsigned* r, int& m) {
    if (k > 27) throw std::runtime_error("ntt: log size > 27 not supported");
    m = k <= 9 ? 1 : (k <= 18 ? 2 : 3);
    unsigned base = k / m, rem = k % m;
    for (int j = 0; j < m; j++) r[j] = base + (j < (int)rem ? 1 : 0);
}

typedef void (*ntt2_fn)(const u64*, u64*, u64, u64, Pass2);
template <int RA, int RB> static ntt2_fn pick(bool last) { return last ? k_ntt2<RA, RB, true> : k_ntt2<RA, RB, false>; }
static ntt2_fn kernel_for(unsigned r, bool last, unsigned& ra, unsigned& rb) {
    switch (r) {
    case 0: ra = 0; rb = 0; return pick<0, 0>(last);
    case 1: ra = 1; rb = 0; return pick<1, 0>(last);
    case 2: ra = 2; rb = 0; return pick<2, 0>(last);
    case 3: ra = 3; rb = 0; return pick<3, 0>(last);
    case 4: ra = 4; rb = 0; return pick<4, 0>(last);
    case 5: ra = 5; rb = 0; return pick<5, 0>(last);
    case 6: ra = 3; rb = 3; return pick<3, 3>(last);
    case 7: ra = 4; rb = 3; return pick<4, 3>(last);
    case 8: ra = 4; rb = 4; return pick<4, 4>(last);
    case 9: ra = 5; rb = 4; return pick<5, 4>(last);
    }
    throw std::runtime_error("ntt: bad digit width");
}

// ---- host side of the TMA path
typedef CUresult (*tm_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                 const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static tm_encode_fn tm_encode() {      // the library links libcudart only: the driver entry point comes through the runtime
    static tm_encode_fn f = [] {
        void* p = nullptr; cudaDriverEntryPointQueryResult q = cudaDriverEntryPointSymbolNotFound;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return (tm_encode_fn)p;
    }();
    return f;
}
static bool ntt3_enabled() {
    static int on = [] { const char* e = getenv("B200_NTT_TMA"); return (e && e[0] == '0') ? 0 : 1; }();
    return on && tm_encode() != nullptr;
}
// rank-3 u64 tensor: dims d0 (contiguous), d1 (stride s1 elements), d2 (stride s2 elements); box b0 x b1 x 1; out-of-bounds reads give 0
static CUtensorMap make_map3(const void* base, u64 d0, u64 d1, u64 d2, u64 s1, u64 s2, u32 b0, u32 b1) {
    alignas(64) CUtensorMap tm;
    cuuint64_t dims[3] = {d0, d1, d2}, strides[2] = {s1 * 8, s2 * 8};
    cuuint32_t box[3] = {b0, b1, 1}, es[3] = {1, 1, 1};
    CUresult rc = tm_encode()(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, const_cast<void*>(base), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) throw std::runtime_error("ntt: cuTensorMapEncodeTiled failed (" + std::to_string((int)rc) + ")");
    return tm;
}
typedef void (*ntt3_fn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, Pass3);
template <int RA, int RB> static ntt3_fn pick3(bool last, bool inv, bool tw, size_t& smem, unsigned& nt) {
    smem = Ntt3Cfg<RA, RB>::smem(last); nt = Ntt3Cfg<RA, RB>::NT;
    if (!last) return inv ? k_ntt3<RA, RB, false, true, true> : k_ntt3<RA, RB, false, false, true>;
    if (tw) return inv ? k_ntt3<RA, RB, true, true, true> : k_ntt3<RA, RB, true, false, true>;
    return inv ? k_ntt3<RA, RB, true, true, false> : k_ntt3<RA, RB, true, false, false>;
}
static ntt3_fn kernel3_for(unsigned r, bool last, bool inv, bool tw, unsigned& ra, unsigned& rb, size_t& smem, unsigned& nt) {
    switch (r) {
    case 6: ra = 3; rb = 3; return pick3<3, 3>(last, inv, tw, smem, nt);
    case 7: ra = 4; rb = 3; return pick3<4, 3>(last, inv, tw, smem, nt);
    case 8: ra = 4; rb = 4; return pick3<4, 4>(last, inv, tw, smem, nt);
    case 9: ra = 5; rb = 4; return pick3<5, 4>(last, inv, tw, smem, nt);
    }
    throw std::runtime_error("ntt: bad digit width for the TMA path");
}
// [k1][d0] -> w_R^(k1 d0), cached per (r, ra, inverse, device)
__global__ void k_twr3(u64* out, u64 w, u32 NB, u32 R) { u32 i = blockIdx.x * blockDim.x + threadIdx.x; if (i < R) out[i] = gl_pow(w, ((i / NB) * (i % NB)) & (R - 1)); }
static std::map<std::tuple<int, unsigned, unsigned, bool>, const u64*> g_twr3;
static const u64* twr3_tab(unsigned r, unsigned ra, bool inverse) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_tuple(current_device(), r, ra, inverse);
    auto it = g_twr3.find(key);
    if (it != g_twr3.end()) return it->second;
    u32 R = 1u << r;
    u64* p; B200_CUDA_CHECK(cudaMalloc(&p, (size_t)R * 8));
    k_twr3<<<(R + 255) / 256, 256, 0, stream()>>>(p, inverse ? h_root_inv(r) : h_root(r), R >> ra, R);
    B200_CUDA_CHECK(cudaGetLastError());
    g_twr3[key] = p;
    return p;
}
// streamed tables.  mode 0: out[kd * S + low] = g^(kd * low) (inter-pass twiddles, same layout as the pass's output);
// mode 1: out[k] = t(k) for the (scaled) power table t (post factors of the LDE's inverse transform)
__global__ void k_stream_tab(u64* out, PowTab t, u32 log_s, u64 count, int mode) {
    u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = mode == 0 ? powtab_get(t, (i >> log_s) * (i & ((1ull << log_s) - 1))) : powtab_get(t, i);
}
static std::map<std::tuple<int, const u64*, unsigned, unsigned, int>, const u64*> g_stream_tab;
static const u64* stream_tab(DevPowTab t, unsigned log_count, unsigned log_s, int mode) {
    std::lock_guard<std::mutex> lk(g_tab_mu);
    auto key = std::make_tuple(current_device(), t.lo, log_count, log_s, mode);
    auto it = g_stream_tab.find(key);
    if (it != g_stream_tab.end()) return it->second;
    const u64 count = 1ull << log_count;
    u64* p; B200_CUDA_CHECK(cudaMalloc(&p, count * 8));
    PowTab pt; pt.lo = t.lo; pt.hi = t.hi;
    k_stream_tab<<<(unsigned)((count + 255) / 256), 256, 0, stream()>>>(p, pt, log_s, count, mode);
    B200_CUDA_CHECK(cudaGetLastError());
    g_stream_tab[key] = p;
    return p;
}
static void ntt3_check_perms() {
    static std::once_flag once;
    std::call_once(once, [] {
        for (unsigned a = 1; a <= 5; a++) for (int inv = 0; inv < 2; inv++) {
            unsigned char perm[32]; shift_perm(a, inv != 0, perm);
            for (u32 p = 0; p < (1u << a); p++) {
                u32 c = (shift_kinv((int)a, inv != 0) * brev_c(p, (int)a)) & ((1u << a) - 1);
                if (c != perm[p]) throw std::runtime_error("ntt: compile-time shift permutation disagrees with the computed one");
            }
        }
    });
}
static unsigned ilog2(u64 v) { unsigned l = 0; while ((1ull << l) < v) l++; return l; }

// One transform through the TMA kernels; returns false (nothing launched) when the shape or the alignment does not qualify.
static bool ntt_run_tma(const u64* in, u64 si, u64 n_in, u64* out, u64 so, u64* tmp, size_t w, unsigned k, bool inverse, bool has_post, DevPowTab post, u64 post_scale, const char* name) {
    if (k < 12 || k > 27 || !ntt3_enabled()) return false;
    const u64 n = 1ull << k;
    unsigned r[3]; int m; split_digits(k, r, m);
    const u64 S0 = n >> r[0];
    if ((((uintptr_t)in | (uintptr_t)out | (uintptr_t)tmp) & 15) || (si & 1) || (so & 1) || n_in == 0 || n_in > n || n_in % S0 != 0) return false;
    if ((u64)w << (k - 6) > 0xffffffffull) return false;          // tensor dimension 2 (columns x hi) must stay below 2^32
    ntt3_check_perms();
    u64 Rprod = 1;
    for (int j = 0; j < m; j++) {
        const bool last = (j == m - 1);
        const u32 R = 1u << r[j], T = (u32)NTT3_TSEL(R), RW = R > 256 ? 256 : R;
        const bool tw = !last || has_post;
        unsigned ra, rb, nt; size_t smem;
        ntt3_fn fn = kernel3_for(r[j], last, inverse, tw, ra, rb, smem, nt);
        Pass3 pp{}; pp.twr = twr3_tab(r[j], ra, inverse); pp.scale = 1; pp.Rprod = (u32)Rprod; pp.R1 = 1;
        const u64* src = (j == 0) ? in : tmp; const u64 ssrc = (j == 0) ? si : n;
        u64* dst = last ? out : tmp; const u64 sdst = last ? so : n;
        CUtensorMap tm_in, tm_tw, tm_out; u64 n_tiles;
        if (!last) {
            const u64 Nj = n / Rprod, S = Nj / R;
            const unsigned lognj = ilog2(Nj);
            const u64 wj = inverse ? h_root_inv(lognj) : h_root(lognj);
            const u64* tab = stream_tab(powtab(wj, lognj), lognj, ilog2(S), 0);
            if (j == 0) tm_in = make_map3(src, S, n_in / S, w, S, ssrc, T, RW);
            else tm_in = make_map3(src, S, R, w * Rprod, S, Nj, T, RW);
            tm_out = make_map3(dst, S, R, w * Rprod, S, Nj, T, RW);
            tm_tw = make_map3(tab, S, R, 1, S, Nj, T, RW);
            pp.log_tiles = ilog2(S / T); n_tiles = Rprod * (S / T);
        } else {
            const u64 R1 = 1ull << r[0], M = (m == 3) ? (1ull << r[1]) : 1, rowlen = n / R1;
            tm_in = make_map3(src, rowlen, R1, w, rowlen, ssrc, RW, T);
            tm_out = make_map3(dst, R1 * M, R, w, R1 * M, sdst, T, RW);
            if (has_post) tm_tw = make_map3(stream_tab(post, k, 0, 1), R1 * M, R, 1, R1 * M, n, T, RW);
            else { tm_tw = tm_out; pp.scale = post_scale; }
            pp.R1 = (u32)R1; pp.log_tiles = ilog2(R1 / T); n_tiles = M * (R1 / T);
        }
        {
            static std::mutex mu; static std::map<std::pair<int, const void*>, bool> done;
            std::lock_guard<std::mutex> lk(mu);
            auto key = std::make_pair(current_device(), (const void*)fn);
            if (!done[key]) { B200_CUDA_CHECK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); done[key] = true; }
        }
        ScopedTimer tmr(name, 16.0 * (double)n * (double)w);
        fn<<<dim3((unsigned)w, (unsigned)n_tiles), nt, smem, stream()>>>(tm_in, tm_tw, tm_out, pp);     // columns on x: the CTAs that share a twiddle tile run together (L2)
        launch_count_add(1);
        Rprod *= R;
    }
    B200_CUDA_CHECK(cudaGetLastError());
    return true;
}

// One transform of `w` columns: in (col stride si, valid rows n_in) -> out (col stride so), size 2^k.
// tmp: scratch of w * 2^k u64 (needed when m >= 2).  post: optional power table applied as X[i] *= post^i.
static void ntt_run(const u64* in, u64 si, u64 n_in, u64* out, u64 so, u64* tmp, size_t w, unsigned k, bool inverse, bool has_post, DevPowTab post, u64 post_scale, const char* name) {
    if (w == 0) return;
    if (ntt_run_tma(in, si, n_in, out, so, tmp, w, k, inverse, has_post, post, post_scale, name)) return;
    const u64 n = 1ull << k;
    unsigned r[3]; int m; split_digits(k, r, m);
    u64 Rprod = 1;
    for (int j = 0; j < m; j++) {
        Pass2 pp{};
        const bool last = (j == m - 1);
        unsigned ra, rb;
        ntt2_fn fn = kernel_for(r[j], last, ra, rb);
        const u32 R = 1u << r[j], NA = 1u << ra;
        pp.n = n; pp.twR = rb ? root_tab(r[j], inverse) : nullptr;
        shift_perm(ra, inverse, pp.permA); shift_perm(rb, inverse, pp.permB);
        pp.n_in = (j == 0) ? n_in : n;
        pp.has_post = 0; pp.post_scale = 1; pp.R1 = 1; pp.M = 1; pp.S = 1; pp.Nj = n;
        u32 T = 256 / NA; if (T > 32) T = 32;
        u64 n_tiles;
        if (!last) {
            pp.Nj = n / Rprod; pp.S = pp.Nj / R;
            if ((u64)T > pp.S) T = (u32)pp.S;
            unsigned lognj = 0; while ((1ull << lognj) < pp.Nj) lognj++;
            const u64 wj = inverse ? h_root_inv(lognj) : h_root(lognj);
            DevPowTab t = powtab(wj, lognj);
            pp.tw.lo = t.lo; pp.tw.hi = t.hi;
            pp.tw1 = lognj <= FULLTAB_MAX_BITS ? full_tab(wj, lognj) : nullptr;
            n_tiles = Rprod * (pp.S / T);
        } else {
            pp.R1 = (m == 1) ? 1 : (1ull << r[0]);
            pp.M = (m == 3) ? (1ull << r[1]) : 1;
            if ((u64)T > pp.R1) T = (u32)pp.R1;
            pp.has_post = has_post ? 1 : 0; pp.post.lo = post.lo; pp.post.hi = post.hi; pp.post_scale = post_scale;
            n_tiles = pp.M * (pp.R1 / T);
        }
        pp.T = T;
        const u64* src; u64* dst; u64 ssrc, sdst;
        if (m == 1) { src = in; ssrc = si; dst = out; sdst = so; }
        else if (j == 0) { src = in; ssrc = si; dst = tmp; sdst = n; }
        else if (!last) { src = tmp; ssrc = n; dst = tmp; sdst = n; }
        else { src = tmp; ssrc = n; dst = out; sdst = so; }
        size_t smem = rb ? ((size_t)R * (T + 1) + R) * 8 : 8;
        unsigned nt = NA * T; if (nt < 32) nt = 32;
        ScopedTimer tmr(name, 16.0 * (double)n * (double)w);
        if (w > 65535) throw std::runtime_error("ntt: too many columns in one call");
        fn<<<dim3((unsigned)n_tiles, (unsigned)w), nt, smem, stream()>>>(src, dst, ssrc, sdst, pp);
        launch_count_add(1);
        Rprod *= R;
    }
    B200_CUDA_CHECK(cudaGetLastError());
}

void ntt_cols(const u64* d_in, u64* d_out, size_t w, unsigned log_n, bool inverse) {
    const u64 n = 1ull << log_n;
    u64* tmp = log_n > 9 ? scratch(w * n) : nullptr;
    u64 scale = inverse ? h_inv(n % GL_P) : 1;
    ntt_run(d_in, n, n, d_out, n, tmp, w, log_n, inverse, false, DevPowTab{nullptr, nullptr}, scale, inverse ? "intt_pass" : "ntt_pass");
}

void ntt_cols_padded(const u64* d_in, size_t in_rows, u64* d_out, size_t w, unsigned log_n) {
    const u64 n = 1ull << log_n;
    u64* tmp = log_n > 9 ? scratch(w * n) : nullptr;
    ntt_run(d_in, in_rows, in_rows, d_out, n, tmp, w, log_n, false, false, DevPowTab{nullptr, nullptr}, 1, "ntt_pass");
}

void lde_cols(const u64* d_in, u64* d_out, size_t w, unsigned log_n, unsigned log_n_ext) {
    if (w == 0) return;
    const u64 n = 1ull << log_n, ne = 1ull << log_n_ext;
    // scratch: coefficient buffer (w*n) + iNTT pass scratch (w*n) + forward scratch (w*ne)
    u64* sc = scratch(w * (2 * n + ne));
    u64* coef = sc; u64* tmp_i = sc + w * n; u64* tmp_f = sc + 2 * w * n;
    DevPowTab shift_tab = powtab_scaled(49, log_n, h_inv(n % GL_P));   // 49^i / N  (fft_p.rs:144-172)
    ntt_run(d_in, n, n, coef, n, tmp_i, w, log_n, true, true, shift_tab, 1, "lde_intt_pass");
    ntt_run(coef, n, n, d_out, ne, tmp_f, w, log_n_ext, false, false, DevPowTab{nullptr, nullptr}, 1, "lde_ntt_pass");
}

}  // namespace b200
