// Integer pipe throughput micro-benchmark for sm_100a: warp-instructions per cycle per SM for the instruction
// kinds the Goldilocks / Montgomery kernels are made of.  Build: nvcc -arch=sm_100a -O3 -o int_pipes int_pipes.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
typedef uint32_t u32; typedef uint64_t u64;
#define ITERS 2048
#define CHAINS 8
template <int KIND> __global__ void __launch_bounds__(256) k(u32* out, u32 seed) {
    u32 a[CHAINS], b[CHAINS]; u64 w[CHAINS];
#pragma unroll
    for (int i = 0; i < CHAINS; i++) { a[i] = seed + threadIdx.x * 7 + i; b[i] = seed * 3 + i * 5 + 1; w[i] = a[i]; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < CHAINS; i++) {
            if (KIND == 0) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i]));                 // IMAD.WIDE.U32
            if (KIND == 1) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));                  // IMAD
            if (KIND == 2) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));                  // IMAD.HI
            if (KIND == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));                                       // IADD3
            if (KIND == 4) asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed)); // IADD3 + IADD3.X
            if (KIND == 5) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(seed)); }  // 1 IMAD.WIDE : 1 IADD3
            if (KIND == 6) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(seed)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed)); }  // 1 : 2
            if (KIND == 7) asm volatile("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.u32 %1, %2, %3, %1;" : "+r"(a[i]), "+r"(b[i]) : "r"(seed), "r"(seed + i));   // fused IMAD.WIDE with carry
            if (KIND == 8) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));               // SHF
            if (KIND == 9) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(b[i]), "r"(seed));               // LOP3
            if (KIND == 10) { asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(a[i]), "r"(b[i])); asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(seed)); asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(seed)); asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i])); }  // 1 : 3
        }
    }
    u32 acc = 0;
#pragma unroll
    for (int i = 0; i < CHAINS; i++) acc ^= a[i] ^ b[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int KIND> void run(const char* name, int instr_per_step, u32* d) {
    int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = sms * 8;
    k<KIND><<<blocks, 256>>>(d, 1); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<KIND><<<blocks, 256>>>(d, 2); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * 8 * ITERS * CHAINS * instr_per_step;
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-34s %8.3f ms  %6.2f warp-instr/clk/SM  (%5.1f lanes/clk/SM)\n", name, ms, warp_instr / cycles / sms, 32 * warp_instr / cycles / sms);
}
int main() {
    u32* d; cudaMalloc(&d, 148 * 8 * 256 * 4 * 4);
    run<0>("IMAD.WIDE.U32", 1, d); run<1>("IMAD", 1, d); run<2>("IMAD.HI", 1, d); run<3>("IADD3", 1, d); run<4>("IADD3 + IADD3.X", 2, d);
    run<7>("mad.lo.cc+madc.hi (fused)", 1, d); run<8>("SHF", 1, d); run<9>("LOP3", 1, d);
    run<5>("IMAD.WIDE : IADD3 = 1:1", 2, d); run<6>("IMAD.WIDE : ALU = 1:2", 3, d); run<10>("IMAD.WIDE : ALU = 1:3", 4, d);
    printf("done\n");
    return 0;
}
