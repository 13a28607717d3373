// Integer pipe throughput micro-benchmark, second version (sm_100a).  Every measured instruction depends on its own previous
// result, so nothing is loop invariant (the first version's IMAD.WIDE line was hoisted by the compiler and measured adds).
// 8 independent chains per thread, 64 warps per SM.  Output: warp-instructions per clock per SM for each SASS form.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes2 int_pipes2.cu ; check the SASS with cuobjdump before trusting a line.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define ITERS 4096
#define CH 8
template <int KIND> __global__ void __launch_bounds__(256) k(u32* out, u32 seed) {
    u32 a[CH], b[CH], c[CH], d[CH]; u64 w[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 5 + 1 + threadIdx.x * 11; c[i] = (seed ^ (i * 0x9e3779b9u)) + threadIdx.x; d[i] = seed + i + threadIdx.x * 3; w[i] = ((u64)a[i] << 32) | b[i]; }
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CH; i++) {
            if (KIND == 0) asm volatile("{ .reg .u32 x, y; mov.b64 {x, y}, %0; mul.wide.u32 %0, x, %1; }" : "+l"(w[i]) : "r"(b[i]));                       // IMAD.WIDE.U32 R, R, R, RZ
            if (KIND == 1) asm volatile("{ .reg .u32 x, y; mov.b64 {x, y}, %0; mad.wide.u32 %0, x, %1, %0; }" : "+l"(w[i]) : "r"(c[i]));   // IMAD.WIDE.U32 R, R, R, R (64-bit addend)
            if (KIND == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));                 // IMAD.HI.U32
            if (KIND == 3) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));                 // IMAD
            if (KIND == 4) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                                      // IADD3
            if (KIND == 5) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]), "r"(d[i]));   // IADD3 + IADD3.X
            if (KIND == 6) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));              // LOP3
            if (KIND == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i]));              // SHF
            if (KIND == 8) asm volatile("mad.lo.cc.u32 %0, %0, %2, %0; madc.hi.u32 %1, %0, %2, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(c[i]));   // carry pair
            if (KIND == 9) { asm volatile("{ .reg .u32 x, y; mov.b64 {x, y}, %0; mul.wide.u32 %0, x, %1; }" : "+l"(w[i]) : "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(d[i])); }   // WIDE : IADD3 = 1 : 1
            if (KIND == 10) { asm volatile("{ .reg .u32 x, y; mov.b64 {x, y}, %0; mul.wide.u32 %0, x, %1; }" : "+l"(w[i]) : "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(d[i])); asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(d[i]) : "r"(c[i])); }   // 1 : 2
            if (KIND == 11) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(d[i])); }   // IMAD : IADD3 = 1 : 1
            if (KIND == 12) { asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(c[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(c[i]) : "r"(d[i])); asm volatile("lop3.b32 %0, %0, %1, %1, 0x96;" : "+r"(d[i]) : "r"(c[i])); }   // 1 : 2
            if (KIND == 13) { asm volatile("{ .reg .u32 x, y; mov.b64 {x, y}, %0; mul.wide.u32 %0, x, %1; }" : "+l"(w[i]) : "r"(b[i])); asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(c[i]) : "r"(d[i])); }   // WIDE : IMAD = 1 : 1 (same pipe?)
            if (KIND == 14) asm volatile("mad.hi.cc.u32 %0, %0, %2, %1; addc.u32 %1, %1, 0;" : "+r"(a[i]), "+r"(c[i]) : "r"(b[i]));   // IMAD.HI with carry out + IADD3.X
        }
    }
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) acc ^= a[i] ^ b[i] ^ c[i] ^ d[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int KIND> void run(const char* name, int instr_per_step, u32* d) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k<KIND><<<blocks, 256>>>(d, 1); cudaDeviceSynchronize();
    float best = 1e9;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0); k<KIND><<<blocks, 256>>>(d, 2 + rep); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    double warp_instr = (double)blocks * 8 * ITERS * CH * instr_per_step;
    double cycles = best * 1e-3 * clk * 1e3;
    // SMSP-cycles one warp's loop iteration (CH chains, see the SASS loop body for the exact instruction list) occupies: 16 warps share an SMSP
    printf("%-44s %8.3f ms  %6.2f nominal warp-instr/clk/SM  %7.2f SMSP-clk per loop iteration\n", name, best, warp_instr / cycles / sms, cycles / ITERS / 16.0);
}
int main() {
    u32* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock rate attribute %d kHz\n", clk);
    run<0>("IMAD.WIDE.U32 (RZ addend)", 1, d); run<1>("IMAD.WIDE.U32 (64-bit addend)", 1, d); run<2>("IMAD.HI.U32", 1, d); run<3>("IMAD (32-bit, addend)", 1, d);
    run<4>("IADD3", 1, d); run<5>("IADD3 + IADD3.X", 2, d); run<6>("LOP3", 1, d); run<7>("SHF", 1, d); run<8>("mad.lo.cc + madc.hi", 2, d);
    run<9>("WIDE : IADD3 = 1:1", 2, d); run<10>("WIDE : ALU = 1:2", 3, d); run<11>("IMAD : IADD3 = 1:1", 2, d); run<12>("IMAD : ALU = 1:2", 3, d); run<13>("WIDE : IMAD = 1:1", 2, d);
    run<14>("IMAD.HI.cc + addc", 2, d);
    printf("done\n");
    return 0;
}
