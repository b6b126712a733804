// ubench.cu -- instruction-rate probes used to choose the accumulate strategy.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ubench tools/ubench.cu && /tmp/ubench
// Prints warp-instructions per clock per SM for a few integer ops and for mma.sync IMMA.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096, kUnroll = 16;

template <int OP>
__global__ void probe(uint32_t* out, uint32_t seed) {
    uint32_t a[kUnroll];
    for (int i = 0; i < kUnroll; ++i) a[i] = seed + threadIdx.x * 7 + i;
    uint32_t b = seed ^ 0x00FF00FF;
    int c[4][4] = {};
    for (int it = 0; it < kIters; ++it) {
#pragma unroll
        for (int i = 0; i < kUnroll; ++i) {
            if (OP == 0) a[i] = __vadd2(a[i], b);                       // VIADD.16x2
            if (OP == 1) a[i] = a[i] + b + (uint32_t)it;               // IADD3
            if (OP == 2) a[i] = __byte_perm(a[i], b, 0x4341);           // PRMT
            if (OP == 3) a[i] = a[i] * 3u + b;                          // IMAD
            if (OP == 4) a[i] = __vmaxs2(a[i], b);                      // VIMNMX.S16x2
            if (OP == 5) a[i] = (a[i] & 0x00FF00FFu) ^ b;               // LOP3
            if (OP == 6) {                                              // IMMA m16n8k32 u8 x s8
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+r"(c[i & 3][0]), "+r"(c[i & 3][1]), "+r"(c[i & 3][2]), "+r"(c[i & 3][3])
                             : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b), "r"(seed));
            }
        }
    }
    uint32_t r = 0;
    for (int i = 0; i < kUnroll; ++i) r ^= a[i];
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) r ^= c[i][j];
    if (r == 0x12345678) out[0] = r;
}

template <int OP>
void run(const char* name, int sms, int mhz) {
    uint32_t* out;
    cudaMalloc(&out, 4);
    const int blocks = sms * 2, threads = 512;  // 32 warps / SM
    probe<OP><<<blocks, threads>>>(out, 1);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, 2);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = double(blocks) * (threads / 32) * kIters * kUnroll;
    const double clocks = ms * 1e-3 * mhz * 1e6;
    printf("%-14s %8.3f ms  %6.3f warp-instr/clk/SM\n", name, ms, warp_instr / clocks / sms);
    cudaFree(out);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int mhz = p.clockRate / 1000;
    printf("%s, %d SMs, %d MHz nominal (rates assume the nominal clock)\n", p.name, p.multiProcessorCount, mhz);
    run<0>("VIADD.16x2", p.multiProcessorCount, mhz);
    run<1>("IADD3", p.multiProcessorCount, mhz);
    run<2>("PRMT", p.multiProcessorCount, mhz);
    run<3>("IMAD", p.multiProcessorCount, mhz);
    run<4>("VIMNMX.S16x2", p.multiProcessorCount, mhz);
    run<5>("LOP3", p.multiProcessorCount, mhz);
    run<6>("IMMA.16832", p.multiProcessorCount, mhz);
    return 0;
}
