// Run-time compilation of step programs: NVRTC (dlopen'ed, so the library still loads where it is absent) turns the CUDA
// source that evaluator.cu generates for one step program into an sm_100a cubin; the driver API (dlopen'ed libcuda) loads
// and launches it.  Nothing here is a fallback to the CPU: when NVRTC or the driver entry points are missing, or a
// compilation fails, eval_program() keeps using the interpreter kernel k_eval on the GPU.
#include "b200_internal.h"
#include <dlfcn.h>
#include <cstring>
#include <cstdlib>
#include <mutex>

namespace b200 {

namespace {
typedef int nvrtcResult;
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
    void* h = nullptr;
    nvrtcResult (*create)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
    nvrtcResult (*compile)(nvrtcProgram, int, const char* const*) = nullptr;
    nvrtcResult (*cubin_size)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*cubin)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*log_size)(nvrtcProgram, size_t*) = nullptr;
    nvrtcResult (*log)(nvrtcProgram, char*) = nullptr;
    nvrtcResult (*destroy)(nvrtcProgram*) = nullptr;
    bool ok = false;
};
typedef int CUresult;
typedef struct CUmod_st* CUmodule;
typedef struct CUfunc_st* CUfunction;
struct Driver {
    void* h = nullptr;
    CUresult (*module_load)(CUmodule*, const void*) = nullptr;
    CUresult (*get_function)(CUfunction*, CUmodule, const char*) = nullptr;
    CUresult (*launch)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, void*, void**, void**) = nullptr;
    bool ok = false;
};
Nvrtc& nvrtc() {
    static Nvrtc N; static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"}) { N.h = dlopen(name, RTLD_NOW | RTLD_LOCAL); if (N.h) break; }
        if (!N.h) return;
        *(void**)&N.create = dlsym(N.h, "nvrtcCreateProgram"); *(void**)&N.compile = dlsym(N.h, "nvrtcCompileProgram");
        *(void**)&N.cubin_size = dlsym(N.h, "nvrtcGetCUBINSize"); *(void**)&N.cubin = dlsym(N.h, "nvrtcGetCUBIN");
        *(void**)&N.log_size = dlsym(N.h, "nvrtcGetProgramLogSize"); *(void**)&N.log = dlsym(N.h, "nvrtcGetProgramLog");
        *(void**)&N.destroy = dlsym(N.h, "nvrtcDestroyProgram");
        N.ok = N.create && N.compile && N.cubin_size && N.cubin && N.log_size && N.log && N.destroy;
    });
    return N;
}
Driver& driver() {
    static Driver D; static std::once_flag once;
    std::call_once(once, [] {
        D.h = dlopen("libcuda.so.1", RTLD_NOW | RTLD_LOCAL);
        if (!D.h) return;
        *(void**)&D.module_load = dlsym(D.h, "cuModuleLoadData"); *(void**)&D.get_function = dlsym(D.h, "cuModuleGetFunction");
        *(void**)&D.launch = dlsym(D.h, "cuLaunchKernel");
        D.ok = D.module_load && D.get_function && D.launch;
    });
    return D;
}
std::string lib_dir() {
    Dl_info info;
    if (dladdr((void*)&lib_dir, &info) && info.dli_fname) { std::string p = info.dli_fname; size_t k = p.find_last_of('/'); return k == std::string::npos ? std::string(".") : p.substr(0, k); }
    return ".";
}
}  // namespace

struct JitKernel { CUmodule mod = nullptr; CUfunction fn = nullptr; };

bool jit_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("B200_JIT"); v = (e && *e == '0') ? 0 : 1; }
    return v == 1;
}

// source -> cubin (host only: works without a GPU, used by the CPU-side tests); empty string + `err` on failure
std::string jit_compile_cubin(const std::string& src, std::string& err) {
    Nvrtc& N = nvrtc();
    if (!N.ok) { err = "libnvrtc not available"; return std::string(); }
    nvrtcProgram prog = nullptr;
    if (N.create(&prog, src.c_str(), "step_program.cu", 0, nullptr, nullptr) != 0) { err = "nvrtcCreateProgram failed"; return std::string(); }
    const char* e = getenv("B200ZK_CSRC");
    std::string inc = std::string("-I") + (e ? std::string(e) : lib_dir() + "/csrc");
    const char* opts[] = {"--gpu-architecture=sm_100a", "-std=c++17", inc.c_str(), "-lineinfo"};
    nvrtcResult rc = N.compile(prog, 4, opts);
    if (rc != 0) {
        size_t ls = 0; N.log_size(prog, &ls); std::string log(ls, 0); if (ls) N.log(prog, &log[0]);
        err = "nvrtcCompileProgram failed: " + log.substr(0, 1500);
        N.destroy(&prog); return std::string();
    }
    size_t sz = 0; N.cubin_size(prog, &sz);
    std::string out(sz, 0);
    if (sz) N.cubin(prog, &out[0]);
    N.destroy(&prog);
    if (!sz) err = "empty cubin";
    return out;
}

JitKernel* jit_load(const std::string& cubin, const char* name, std::string& err) {
    Driver& D = driver();
    if (!D.ok) { err = "libcuda entry points not available"; return nullptr; }
    cudaFree(0);        // make sure the runtime's primary context exists and is current on this thread
    JitKernel* k = new JitKernel();
    CUresult rc = D.module_load(&k->mod, cubin.data());
    if (rc != 0) { err = "cuModuleLoadData failed: " + std::to_string(rc); delete k; return nullptr; }
    rc = D.get_function(&k->fn, k->mod, name);
    if (rc != 0) { err = "cuModuleGetFunction failed: " + std::to_string(rc); delete k; return nullptr; }
    return k;
}

void jit_launch(JitKernel* k, unsigned grid, unsigned block, cudaStream_t st, void** args) {
    CUresult rc = driver().launch(k->fn, grid, 1, 1, block, 1, 1, 0, (void*)st, args, nullptr);
    if (rc != 0) throw std::runtime_error("CUDA error: cuLaunchKernel failed for a JIT-compiled step program: " + std::to_string(rc));
}

}  // namespace b200
