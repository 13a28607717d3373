// compressor12 exec phase on the device and trace-file streaming.
//
// Reference: recursion/src/compressor12/compressor12_exec.rs:19-108 -- `exec` extends the circom witness with the PlonkAdd
// rows of the `.exec` file (w.push(w[a] * ka + w[b] * kb), :58-64; the coefficients are stored as the RAW Montgomery limb of
// the field element, compressor12_setup.rs:68-73 / field_gl.rs:337-356,503-507) and scatters it into the 12 committed columns
// `Compressor.a[0..12]` through the signal map (cm[i][c] = w[s_map[12 i + c]], 0 where the map is 0 and on the padding rows
// :71-92), then SAVES the polynomials to a `.cm` file that `stark_prove` loads again (starky/src/polsarray.rs:137-217).
// Here the fill writes straight into a device buffer in the prover's input layout (row-major N x 12), so the trace of a
// recursion stage never touches the disk; `pols_load_dev` streams an existing `.cm/.const` file through two pinned staging
// buffers for the stages that still come from files.
#include "b200_internal.h"
#include "field.cuh"
#include <cstdio>
#include <cstring>

namespace b200 {

// the PlonkAdd chain is sequential by construction (a row may use the result of any earlier row): host code, exact GL arithmetic
void c12_extend_witness(const u64* adds, size_t adds_len, std::vector<u64>& w) {
    const u64 r_inv = h_inv(GL_EPS);                 // R = 2^64 = 2^32 - 1 (mod p): raw Montgomery limb x  ->  x / R
    w.reserve(w.size() + adds_len);
    for (size_t i = 0; i < adds_len; i++) {
        const u64 ia = adds[4 * i], ib = adds[4 * i + 1], ra = adds[4 * i + 2], rb = adds[4 * i + 3];
        if (ia >= w.size() || ib >= w.size()) throw std::invalid_argument("exec file: PlonkAdd refers to a signal that does not exist yet");
        if (ra >= GL_P || rb >= GL_P) throw std::invalid_argument("exec file: coefficient is not a valid field representation");   // from_raw_repr fails (is_valid)
        const u64 ka = h_mul(ra, r_inv), kb = h_mul(rb, r_inv);
        w.push_back(h_add(h_mul(w[ia], ka), h_mul(w[ib], kb)));
    }
}

// out[i][c] = w[s_map[12 i + c]] (0 where the map entry is 0), rows >= map_rows are zero; out: row-major n_rows x 12
__global__ void __launch_bounds__(256) k_c12_fill(const u64* __restrict__ w, size_t n_w, const u64* __restrict__ s_map, size_t map_rows, u64* __restrict__ out, size_t n_rows, int* __restrict__ bad) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_rows * 12) return;
    u64 v = 0;
    if (t < map_rows * 12) {
        const u64 s = s_map[t];
        if (s >= n_w) { *bad = 1; }
        else if (s) v = __ldg(w + s);
    }
    out[t] = v;
}

void c12_exec_dev(const u64* exec, size_t exec_len, const u64* witness, size_t n_witness, size_t n_rows, u64* d_cm_rowmajor) {
    if (exec_len < 2) throw std::invalid_argument("exec file: too short");
    const size_t adds_len = exec[0], map_rows = exec[1];
    if (exec_len != 2 + adds_len * 4 + map_rows * 12) throw std::invalid_argument("exec file: length does not match its header");      // assert_eq!(new_buff.len(), size)
    if (map_rows > n_rows) throw std::invalid_argument("exec file: more mapped rows than the polynomial degree");
    std::vector<u64> w(n_witness);
    for (size_t i = 0; i < n_witness; i++) w[i] = witness[i] % GL_P;            // FGL::from(u64)
    c12_extend_witness(exec + 2, adds_len, w);
    const u64* s_map = exec + 2 + adds_len * 4;
    static char* g_buf[16] = {nullptr}; static size_t g_cap[16] = {0};
    int dev = current_device();
    const size_t need = (w.size() + map_rows * 12 + 2) * 8;
    if (g_cap[dev] < need) { if (g_buf[dev]) { B200_CUDA_CHECK(cudaStreamSynchronize(stream())); B200_CUDA_CHECK(cudaFree(g_buf[dev])); } B200_CUDA_CHECK(cudaMalloc(&g_buf[dev], need)); g_cap[dev] = need; }
    u64* d_w = reinterpret_cast<u64*>(g_buf[dev]); u64* d_map = d_w + w.size(); int* d_bad = reinterpret_cast<int*>(d_map + map_rows * 12);
    cudaStream_t st = stream();
    B200_CUDA_CHECK(cudaMemcpyAsync(d_w, w.data(), w.size() * 8, cudaMemcpyHostToDevice, st));
    if (map_rows) B200_CUDA_CHECK(cudaMemcpyAsync(d_map, s_map, map_rows * 96, cudaMemcpyHostToDevice, st));
    B200_CUDA_CHECK(cudaMemsetAsync(d_bad, 0, 4, st));
    {
        ScopedTimer t("c12_fill", 8.0 * 12 * (double)(n_rows + map_rows));
        k_c12_fill<<<(unsigned)((n_rows * 12 + 255) / 256), 256, 0, st>>>(d_w, w.size(), d_map, map_rows, d_cm_rowmajor, n_rows, d_bad);
        launch_count_add(1);
    }
    B200_CUDA_CHECK(cudaGetLastError());
    int bad = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    if (bad) throw std::invalid_argument("exec file: the signal map refers to a signal that does not exist");
}

// `.cm/.const` file (row-major little-endian u64, polsarray.rs:137-217) -> device buffer, through two pinned staging buffers so
// that the file read of chunk k+1 overlaps the H2D copy of chunk k.  Values are checked to be canonical (< p).
void pols_load_dev(const char* path, size_t n_u64, u64* d_out) {
    FILE* f = fopen(path, "rb");
    if (!f) throw std::invalid_argument(std::string("cannot open ") + path);
    const size_t CH = (size_t)8 << 20;                  // 8 Mi u64 = 64 MiB per staging buffer
    u64* pin[2] = {nullptr, nullptr}; cudaEvent_t ev[2];
    try {
        for (int k = 0; k < 2; k++) { B200_CUDA_CHECK(cudaMallocHost(&pin[k], CH * 8)); B200_CUDA_CHECK(cudaEventCreate(&ev[k])); }
        size_t done = 0; int k = 0;
        while (done < n_u64) {
            const size_t n = n_u64 - done < CH ? n_u64 - done : CH;
            B200_CUDA_CHECK(cudaEventSynchronize(ev[k]));                 // the copy that last used this buffer has finished
            if (fread(pin[k], 8, n, f) != n) throw std::invalid_argument(std::string("short read: ") + path);
            for (size_t i = 0; i < n; i++) if (pin[k][i] >= GL_P) throw std::invalid_argument(std::string("non-canonical field element in ") + path);
            B200_CUDA_CHECK(cudaMemcpyAsync(d_out + done, pin[k], n * 8, cudaMemcpyHostToDevice, stream()));
            B200_CUDA_CHECK(cudaEventRecord(ev[k], stream()));
            done += n; k ^= 1;
        }
        char extra;
        if (fread(&extra, 1, 1, f) == 1) throw std::invalid_argument(std::string("file is longer than n_rows x n_cols: ") + path);
        B200_CUDA_CHECK(cudaStreamSynchronize(stream()));
    } catch (...) {
        fclose(f); for (int k = 0; k < 2; k++) if (pin[k]) { cudaFreeHost(pin[k]); cudaEventDestroy(ev[k]); }
        throw;
    }
    fclose(f);
    for (int k = 0; k < 2; k++) { cudaFreeHost(pin[k]); cudaEventDestroy(ev[k]); }
}

}  // namespace b200
