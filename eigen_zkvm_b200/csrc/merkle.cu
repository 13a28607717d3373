// Poseidon-GL commitment kernels: LinearHash leaves, binary Merkle levels, openings.
// Reference: starky/src/linearhash.rs:79-145, merklehash.rs:47-134,293-346, poseidon_opt.rs:80-200.
//
// Roofline class: INT-ALU bound (about 1.1k 64x64 modular products + 2k small MACs per permutation against
// 96 B of node traffic), so the kernels are laid out for issue rate: one permutation per thread, state in
// registers, constants through the constant bank, 32-byte digests read/written as 2 x 16-byte vectors.
#include "b200_internal.h"
#include "poseidon.cuh"
#include "poseidon_gl_params.h"
#include <cstring>
#include <vector>
#include <map>
#include <mutex>

namespace b200 {

static bool g_pos_ready[B200_MAX_DEVICES] = {false};
static void pos_init() {
    int dev = current_device();          // under the device lock of the calling entry point
    if (g_pos_ready[dev]) return;
    u64 c[118], p[144], s[506]; u32 m[144];
    for (int i = 0; i < 118; i++) c[i] = POS_C[i] % GL_P;
    for (int i = 0; i < 144; i++) { p[i] = POS_P[i] % GL_P; m[i] = (u32)POS_M[i]; if (POS_M[i] >> 8) throw std::runtime_error("MDS entry not small"); }
    for (int i = 0; i < 506; i++) s[i] = POS_S[i] % GL_P;
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_C, c, sizeof c));
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_P, p, sizeof p));
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_S, s, sizeof s));
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_M, m, sizeof m));
    u64 rc[96];
    for (int r = 0; r < 8; r++) for (int i = 0; i < 12; i++) rc[r * 12 + i] = r < 4 ? c[12 * (r + 1) + i] : (r < 7 ? c[82 + 12 * (r - 4) + i] : 0);
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_RC, rc, sizeof rc));
    u64 k0[12]; u32 mt[144];
    for (int i = 0; i < 12; i++) { u64 x = c[i], x2 = gl_mul(x, x), x3 = gl_mul(x2, x), x6 = gl_mul(x3, x3); k0[i] = gl_add(gl_mul(x6, x), rc[i]); }
    for (int i = 0; i < 12; i++) for (int j = 0; j < 12; j++) {
        mt[i * 12 + j] = m[j * 12 + i];
        // pos_mds3 relies on: entries < 64 (three 32-bit limb sums cannot overflow) and the circulant shape
        int k = (i - j + 12) % 12;
        u32 expect = (i == 0 && j == 0) ? m[0] : (k == 0 ? m[13] : m[k]);      // M[j][i] = c[(i - j) mod 12], c[k] = M[0][k], c[0] = M[1][1]
        if (m[j * 12 + i] >= 64 || m[j * 12 + i] != expect) throw std::runtime_error("MDS matrix is not the expected small circulant");
    }
    // blocked partial rounds: see pos_partial_block (poseidon.cuh)
    {
        std::vector<u64> blk;
        const int sizes[6] = {4, 4, 4, 4, 4, 2};
        int r0 = 0;
        for (int b = 0; b < 6; b++) {
            const int B = sizes[b];
            for (int i = 0; i < B; i++) {
                const u64* Sr = s + 23 * (r0 + i);
                for (int j = 0; j < 12; j++) blk.push_back(Sr[j]);
                for (int m2 = 0; m2 < i; m2++) {
                    const u64* Sm = s + 23 * (r0 + m2);
                    u64 w = 0;
                    for (int j = 1; j < 12; j++) w = gl_add(w, gl_mul(Sr[j], Sm[11 + j]));
                    blk.push_back(w);
                }
            }
            for (int j = 1; j < 12; j++) for (int m2 = 0; m2 < B; m2++) blk.push_back(s[23 * (r0 + m2) + 11 + j]);
            r0 += B;
        }
        if (blk.size() != sizeof(cPOS_BLK) / 8) throw std::runtime_error("internal: blocked Poseidon table size");
        B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_BLK, blk.data(), blk.size() * 8));
    }
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_K0, k0, sizeof k0));
    B200_CUDA_CHECK(cudaMemcpyToSymbol(cPOS_MT, mt, sizeof mt));
    {   // per-lane constants of poseidon12_warp in global memory (see poseidon.cuh)
        std::vector<u64> wt(POS_WT_WORDS, 0);
        for (int i = 0; i < 12; i++) wt[POS_WT_C + i] = c[i];
        for (int i = 0; i < 96; i++) wt[POS_WT_RC + i] = rc[i];
        for (int i = 0; i < 144; i++) wt[POS_WT_P + i] = p[i];
        for (int i = 0; i < 506; i++) wt[POS_WT_S + i] = s[i];
        memcpy(&wt[POS_WT_MT], mt, sizeof mt);
        B200_CUDA_CHECK(cudaMemcpyToSymbol(gPOS_WT, wt.data(), wt.size() * 8));
    }
    g_pos_ready[dev] = true;
}

size_t merkle_n_nodes(size_t n_) {
    size_t n = n_, next_n = (n - 1) / 2 + 1, acc = next_n * 2;
    while (n > 1) { n = next_n; next_n = (n - 1) / 2 + 1; if (n > 1) acc += next_n * 2; else acc += 1; }
    return acc;
}

// ------------------------------------------------------------------------------------------------ single permutation
__global__ void k_poseidon_single(const u64* __restrict__ in, u64* __restrict__ out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    u64 st[12];
#pragma unroll
    for (int i = 0; i < 12; i++) st[i] = in[i];
    poseidon12(st);
#pragma unroll
    for (int i = 0; i < 12; i++) out[i] = gl_canon(st[i]);
}
// one permutation of the DEVICE kernels with the result on the host: b200_gl_poseidon (the finer seam of `Poseidon::hash`) and the
// parity tests go through this; the transcript uses poseidon12_host (poseidon_host.cpp)
void poseidon_perm_device(const u64 in12[12], u64 out12[12]) {
    pos_init();
    static u64* g_perm_buf[16] = {nullptr}; static std::mutex mu;
    std::lock_guard<std::mutex> lk(mu);
    int dev = current_device();
    if (!g_perm_buf[dev]) B200_CUDA_CHECK(cudaMalloc(&g_perm_buf[dev], 24 * sizeof(u64)));
    u64* b = g_perm_buf[dev];
    B200_CUDA_CHECK(cudaMemcpyAsync(b, in12, 96, cudaMemcpyHostToDevice, stream()));
    k_poseidon_single<<<1, 32, 0, stream()>>>(b, b + 12);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaMemcpyAsync(out12, b + 12, 96, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}
void poseidon_perm_host(const u64 in12[12], u64 out12[12]) { poseidon12_host(in12, out12); }

// ------------------------------------------------------------------------------------------------ leaves
GL_D u64 col_load(const ColView& v, u32 c, size_t row) {
    u64 off = (u64)(c / v.a) * v.s1 + (u64)(c % v.a) * v.s2;
    return __ldg(v.base + off + row);
}
// sponge over `len` logical columns starting at c0 (linearhash.rs:112-145, `_hash`)
GL_D void lh_sponge_cols(const ColView& v, u32 c0, u32 len, size_t row, u64* out4) {
    if (len <= 4) {
#pragma unroll
        for (int k = 0; k < 4; k++) out4[k] = (u32)k < len ? col_load(v, c0 + k, row) : 0;
        return;
    }
    u64 st[12];
    u64 cap0 = 0, cap1 = 0, cap2 = 0, cap3 = 0;
    for (u32 i = 0; i < len; i += 8) {
#pragma unroll
        for (int k = 0; k < 8; k++) st[k] = (i + k < len) ? col_load(v, c0 + i + k, row) : 0;
        st[8] = cap0; st[9] = cap1; st[10] = cap2; st[11] = cap3;
        poseidon12<true, 4>(st, i == 0);
        cap0 = st[0]; cap1 = st[1]; cap2 = st[2]; cap3 = st[3];
    }
    out4[0] = gl_canon(cap0); out4[1] = gl_canon(cap1); out4[2] = gl_canon(cap2); out4[3] = gl_canon(cap3);
}
#ifndef POS_LH_MIN_BLOCKS
#define POS_LH_MIN_BLOCKS 8
#endif
__global__ void __launch_bounds__(128, POS_LH_MIN_BLOCKS) k_linearhash(ColView v, u32 width, size_t height, u64* __restrict__ digests) {
    size_t row = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= height) return;
    u64 out[4];
    if (width <= 4) {
        lh_sponge_cols(v, 0, width, row, out);
    } else {
        u32 bs = (width + 3) / 4; if (bs < 8) bs = 8;           // linearhash.rs:80-83
        u32 hsz = (width + bs - 1) / bs;
        u64 h[16];
#pragma unroll 1
        for (u32 c = 0; c < hsz; c++) {
            u32 len = width - c * bs < bs ? width - c * bs : bs;
            u64 o[4];
            lh_sponge_cols(v, c * bs, len, row, o);
            h[4 * c] = o[0]; h[4 * c + 1] = o[1]; h[4 * c + 2] = o[2]; h[4 * c + 3] = o[3];
        }
        if (hsz == 1) { out[0] = h[0]; out[1] = h[1]; out[2] = h[2]; out[3] = h[3]; }
        else {
            // second sponge over the 4*hsz (8, 12 or 16) chunk digests
            u64 st[12];
            u64 cap[4] = {0, 0, 0, 0};
            u32 n = 4 * hsz;
#pragma unroll 1
            for (u32 i = 0; i < n; i += 8) {
#pragma unroll
                for (int k = 0; k < 8; k++) st[k] = (i + k < n) ? h[i + k] : 0;
                st[8] = cap[0]; st[9] = cap[1]; st[10] = cap[2]; st[11] = cap[3];
                poseidon12<true, 4>(st, i == 0);
                cap[0] = st[0]; cap[1] = st[1]; cap[2] = st[2]; cap[3] = st[3];
            }
            out[0] = gl_canon(cap[0]); out[1] = gl_canon(cap[1]); out[2] = gl_canon(cap[2]); out[3] = gl_canon(cap[3]);
        }
    }
    ulonglong2* o = reinterpret_cast<ulonglong2*>(digests + 4 * row);
    o[0] = make_ulonglong2(out[0], out[1]);
    o[1] = make_ulonglong2(out[2], out[3]);
}
void linearhash_rows(ColView cols, size_t width, size_t height, u64* d_digests) {
    pos_init();
    if (height == 0) return;
    size_t perms = 0;
    if (width > 4) { size_t bs = (width + 3) / 4; if (bs < 8) bs = 8; size_t hsz = (width + bs - 1) / bs;
        for (size_t c = 0; c < hsz; c++) { size_t len = width - c * bs < bs ? width - c * bs : bs; if (len > 4) perms += (len + 7) / 8; }
        if (hsz > 1) perms += (4 * hsz + 7) / 8; }
    (void)perms;
    ScopedTimer t("linearhash_leaves", (double)height * (8.0 * width + 32.0));
    unsigned blocks = (unsigned)((height + 127) / 128);
    k_linearhash<<<blocks, 128, 0, stream()>>>(cols, (u32)width, height, d_digests);
    launch_count_add(1);
    B200_CUDA_CHECK(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------------ levels
// out[i] = Poseidon(in[2i] || in[2i+1], cap = 0)[0..4]   (merklehash.rs:110-134)
#ifndef POS_LEVEL_MIN_BLOCKS
#define POS_LEVEL_MIN_BLOCKS 8
#endif
// ZMASK: input lanes known to be zero -- always the capacity (0xF00); for the first level over leaves of width 1..3
// (digest = the zero-padded row, linearhash.rs:86-95) also lanes w..3 and 4+w..7
template <u32 ZMASK>
__global__ void __launch_bounds__(128, POS_LEVEL_MIN_BLOCKS) k_merkle_level(const u64* __restrict__ in, u64* __restrict__ out, size_t n_out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_out) return;
    const ulonglong2* p = reinterpret_cast<const ulonglong2*>(in + 8 * i);
    ulonglong2 a = p[0], b = p[1], c = p[2], d = p[3];
    u64 st[12] = {a.x, a.y, b.x, b.y, c.x, c.y, d.x, d.y, 0, 0, 0, 0};
    poseidon12<false, 4, ZMASK, true>(st);       // inlined S-boxes, unrolled small-MDS layers: best for this kernel (profiles/README.md)
    ulonglong2* o = reinterpret_cast<ulonglong2*>(out + 4 * i);
    o[0] = make_ulonglong2(gl_canon(st[0]), gl_canon(st[1]));
    o[1] = make_ulonglong2(gl_canon(st[2]), gl_canon(st[3]));
}
// All remaining levels once a level has <= MERKLE_TOP nodes: one CTA, a barrier between levels (the last ten levels of every tree
// were ten launches of a fraction of a wave each).
#define MERKLE_TOP 32          /* levels with at most this many nodes are fused into one launch */
#define MERKLE_WARP_MAX 8192   /* levels with at most this many nodes run one WARP per node (latency form), larger ones one thread per node */
// one level, one warp per node
__global__ void __launch_bounds__(256) k_merkle_level_warp(const u64* __restrict__ in, u64* __restrict__ out, size_t n_out) {
    const size_t i = (size_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= n_out) return;                                         // warp-uniform
    u64 s = lane < 8 ? __ldg(in + 8 * i + lane) : 0;
    s = poseidon12_warp(s);
    if (lane < 4) out[4 * i + lane] = gl_canon(s);
}
// all remaining levels in one launch, one WARP per node (poseidon12_warp): lane j < 8 loads child word j, lanes 8..11 are the zero capacity
__global__ void __launch_bounds__(1024) k_merkle_top(u64* nodes, size_t n64, size_t p_in) {
    size_t next = (n64 - 1) / 2 + 1, p_out = p_in + next * 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, n_warps = blockDim.x >> 5;
    while (n64 > 1) {
        if ((n64 & 1) && threadIdx.x < 4) nodes[4 * (p_in + n64) + threadIdx.x] = 0;
        __syncthreads();
        for (size_t i = warp; i < next; i += n_warps) {            // warp-uniform trip count
            u64 s = lane < 8 ? nodes[4 * (p_in + 2 * i) + lane] : 0;
            s = poseidon12_warp(s);
            if (lane < 4) nodes[4 * (p_out + i) + lane] = gl_canon(s);
        }
        __syncthreads();
        n64 = next; next = (n64 - 1) / 2 + 1; p_in = p_out; p_out = p_in + next * 2;
    }
}
void merkle_levels(u64* d_nodes, size_t height, size_t leaf_width) {
    pos_init();
    size_t n64 = height, next = (n64 - 1) / 2 + 1, p_in = 0, p_out = next * 2;
    bool first = true;
    while (n64 > 1) {
        if (next <= MERKLE_WARP_MAX && next > MERKLE_TOP && !first) {
            if (n64 & 1) B200_CUDA_CHECK(cudaMemsetAsync(d_nodes + 4 * (p_in + n64), 0, 32, stream()));
            ScopedTimer t("merkle_level_warp", 96.0 * (double)next);
            k_merkle_level_warp<<<(unsigned)((next + 7) / 8), 256, 0, stream()>>>(d_nodes + 4 * p_in, d_nodes + 4 * p_out, next);
            launch_count_add(1);
            n64 = next; next = (n64 - 1) / 2 + 1; p_in = p_out; p_out = p_in + next * 2;
            continue;
        }
        if (next <= MERKLE_TOP && !first) {
            ScopedTimer t("merkle_top", 96.0 * (double)(n64 - 1));
            k_merkle_top<<<1, 1024, 0, stream()>>>(d_nodes, n64, p_in);
            launch_count_add(1);
            break;
        }
        if (n64 & 1) B200_CUDA_CHECK(cudaMemsetAsync(d_nodes + 4 * (p_in + n64), 0, 32, stream()));   // zero pad digest
        {
            ScopedTimer t("merkle_level", 96.0 * (double)next);
            unsigned blocks = (unsigned)((next + 127) / 128);
            const u64* in = d_nodes + 4 * p_in; u64* out = d_nodes + 4 * p_out;
            size_t lw = first ? leaf_width : 0;
            if (lw == 1) k_merkle_level<0xFEEu><<<blocks, 128, 0, stream()>>>(in, out, next);
            else if (lw == 2) k_merkle_level<0xFCCu><<<blocks, 128, 0, stream()>>>(in, out, next);
            else if (lw == 3) k_merkle_level<0xF88u><<<blocks, 128, 0, stream()>>>(in, out, next);
            else k_merkle_level<0xF00u><<<blocks, 128, 0, stream()>>>(in, out, next);
            launch_count_add(1);
            first = false;
        }
        n64 = next; next = (n64 - 1) / 2 + 1; p_in = p_out; p_out = p_in + next * 2;
    }
    B200_CUDA_CHECK(cudaGetLastError());
}

void merkelize(DevTree& t, ColView cols, size_t width, size_t height, u64* d_nodes) {
    if (t.hash != 0) { big_merkelize_tree(t, cols, width, height, d_nodes); return; }
    pos_init();
    t.cols = cols; t.width = width; t.height = height; t.nodes = d_nodes; t.degenerate = false; t.level_digest.clear();
    if (width == 0) {
        // Every leaf digest is 0^4 and every level repeats one value (the reference hashes all of them:
        // merklehash.rs:311-343 with an empty buffer); log2(height) permutations give the same nodes.
        t.degenerate = true; t.nodes = nullptr;
        // the digests depend on the height alone: computed once per process (saves ~2 log2(N) device round trips per proof)
        static std::map<size_t, std::vector<std::array<u64, 4>>> cache; static std::mutex mu;
        {
            std::lock_guard<std::mutex> lk(mu);
            auto it = cache.find(height);
            if (it != cache.end()) { t.level_digest = it->second; memcpy(t.root, t.level_digest.back().data(), 32); return; }
        }
        std::array<u64, 4> cur = {0, 0, 0, 0};
        t.level_digest.push_back(cur);
        size_t n = height;
        while (n > 1) {
            if (n & 1) throw std::runtime_error("degenerate tree needs a power-of-two height");
            u64 in[12] = {cur[0], cur[1], cur[2], cur[3], cur[0], cur[1], cur[2], cur[3], 0, 0, 0, 0}, out[12];
            poseidon_perm_host(in, out);
            cur = {out[0], out[1], out[2], out[3]};
            t.level_digest.push_back(cur);
            n >>= 1;
        }
        memcpy(t.root, cur.data(), 32);
        { std::lock_guard<std::mutex> lk(mu); cache[height] = t.level_digest; }
        return;
    }
    linearhash_rows(cols, width, height, d_nodes);
    merkle_levels(d_nodes, height, width <= 3 ? width : 0);
    size_t nn = merkle_n_nodes(height);
    B200_CUDA_CHECK(cudaMemcpyAsync(t.root, d_nodes + 4 * (nn - 1), 32, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}

// ------------------------------------------------------------------------------------------------ openings
// merklehash.rs:64-77 + get_group_proof (:430-438): leaf row + sibling digests bottom-up
__global__ void k_merkle_open(ColView v, u32 width, size_t height, const u64* __restrict__ nodes, const u64* __restrict__ idxs,
                              u64* __restrict__ vals, u64* __restrict__ sibs, u32 depth) {
    size_t q = blockIdx.x;
    size_t idx = idxs[q];
    for (u32 c = threadIdx.x; c < width; c += blockDim.x) vals[q * width + c] = col_load(v, c, idx);
    if (threadIdx.x == 0) {
        size_t n = height, off = 0, id = idx; u32 d = 0;
        while (n > 1) {
            size_t si = id ^ 1;
            for (int k = 0; k < 4; k++) sibs[(q * depth + d) * 4 + k] = nodes[4 * (off + si) + k];
            size_t next_n = (n - 1) / 2 + 1;
            off += next_n * 2; id >>= 1; n = next_n; d++;
        }
    }
}
void merkle_open(const DevTree& t, const std::vector<u64>& idx, std::vector<u64>& vals, std::vector<u64>& sibs, size_t& depth) {
    if (t.hash != 0) { big_merkle_open(t, idx, vals, sibs, depth); return; }
    size_t nq = idx.size();
    depth = 0; { size_t n = t.height; while (n > 1) { n = (n - 1) / 2 + 1; depth++; } }
    vals.assign(nq * t.width, 0); sibs.assign(nq * depth * 4, 0);
    if (nq == 0) return;
    if (t.degenerate) {
        for (size_t q = 0; q < nq; q++) for (size_t d = 0; d < depth; d++) memcpy(&sibs[(q * depth + d) * 4], t.level_digest[d].data(), 32);
        return;
    }
    u64 *d_idx, *d_vals, *d_sibs;
    size_t nv = nq * t.width, ns = nq * depth * 4;
    static u64* g_buf[16] = {nullptr}; static size_t g_cap[16] = {0};
    int dev0 = current_device();
    size_t need = nq + nv + ns + 1;
    if (g_cap[dev0] < need) { if (g_buf[dev0]) B200_CUDA_CHECK(cudaFree(g_buf[dev0])); size_t cap = need < (1u << 16) ? (1u << 16) : need; B200_CUDA_CHECK(cudaMalloc(&g_buf[dev0], cap * 8)); g_cap[dev0] = cap; }
    d_idx = g_buf[dev0];
    d_vals = d_idx + nq; d_sibs = d_vals + nv;
    B200_CUDA_CHECK(cudaMemcpyAsync(d_idx, idx.data(), nq * 8, cudaMemcpyHostToDevice, stream()));
    k_merkle_open<<<(unsigned)nq, 128, 0, stream()>>>(t.cols, (u32)t.width, t.height, t.nodes, d_idx, d_vals, d_sibs, (u32)depth);
    launch_count_add(1);
    if (nv) B200_CUDA_CHECK(cudaMemcpyAsync(vals.data(), d_vals, nv * 8, cudaMemcpyDeviceToHost, stream()));
    if (ns) B200_CUDA_CHECK(cudaMemcpyAsync(sibs.data(), d_sibs, ns * 8, cudaMemcpyDeviceToHost, stream()));
    B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
}

}  // namespace b200
