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

// ---- on-chip bandwidth probes: the denominators of roofline.onchip in bench.py (profiles/onchip_peaks.json)
// mode 0: L2 -> SM, ld.global.cg.v4 (no L1 allocation) over a 64 MiB buffer that stays L2-resident
// mode 1: L1 hits, ld.global.ca.v4 over a 64 KiB window per CTA
// mode 2: shared memory, ld.shared.v4
// mode 3: L2 -> shared memory, cp.async.cg 16 B per lane (LDGSTS.BYPASS), the path ft_group_kernel stages rows with
template <int MODE>
__global__ void __launch_bounds__(512) bw_probe(const uint4* __restrict__ buf, size_t n_vec, int passes, uint32_t* out) {
    __shared__ __align__(16) uint4 tile[2048]; /* 32 KiB */
    uint32_t acc = 0;
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    if (MODE == 2)
        for (int i = threadIdx.x; i < 2048; i += blockDim.x) tile[i] = make_uint4(i, 1, 2, 3);
    __syncthreads();
    for (int pass = 0; pass < passes; ++pass) {
        if (MODE == 0) {
            for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i + 3 * stride < n_vec; i += 4 * stride) {
                uint4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(buf + i + u * stride));
#pragma unroll
                for (int u = 0; u < 4; ++u) acc += v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            }
        } else if (MODE == 1) {
            const uint4* win = buf + static_cast<size_t>(blockIdx.x % 64) * 4096; /* 64 KiB window */
            for (int rep = 0; rep < 64; ++rep)
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    uint4 v;
                    asm volatile("ld.global.ca.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "l"(win + ((u * 512 + threadIdx.x + rep * 32) & 4095)) : "memory");
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                }
        } else if (MODE == 2) {
            for (int rep = 0; rep < 256; ++rep)
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    uint4 v;
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                                 : "r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tile[(u * 512 + threadIdx.x + rep) & 2047]))));
                    acc += v.x ^ v.y ^ v.z ^ v.w;
                }
        } else {
            for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i + 3 * stride < n_vec; i += 4 * stride) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(&tile[u * 512 + threadIdx.x]))), "l"(buf + i + u * stride) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            }
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            acc += tile[threadIdx.x].x;
        }
    }
    if (acc == 0x12345678) out[0] = acc;
}

template <int MODE>
double run_bw(const char* name, int sms, double* clk_bytes, int mhz) {
    const size_t bytes = 64ull << 20, n_vec = bytes / 16;
    uint4* buf;
    uint32_t* out;
    cudaMalloc(&buf, bytes), cudaMalloc(&out, 4);
    cudaMemset(buf, 1, bytes);
    const int blocks = sms * (MODE == 3 ? 4 : 2), threads = 512, passes = MODE == 0 || MODE == 3 ? 20 : 200;
    bw_probe<MODE><<<blocks, threads>>>(buf, n_vec, 2, out); /* warms L2 */
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bw_probe<MODE><<<blocks, threads>>>(buf, n_vec, passes, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double moved;
    if (MODE == 0 || MODE == 3) moved = double(passes) * (n_vec / (4ull * blocks * threads)) * (4ull * blocks * threads) * 16;
    else if (MODE == 1) moved = double(passes) * 64 * 8 * 16.0 * blocks * threads;
    else moved = double(passes) * 256 * 4 * 16.0 * blocks * threads;
    const double gbs = moved / (ms * 1e-3) / 1e9;
    *clk_bytes = moved / (ms * 1e-3 * mhz * 1e6) / sms;
    printf("%-28s %8.3f ms  %9.1f GB/s  %6.1f B/clk/SM (at the nominal clock)\n", name, ms, gbs, *clk_bytes);
    cudaFree(buf), cudaFree(out);
    return gbs;
}

int main(int argc, char** argv) {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int mhz = p.clockRate / 1000;
    printf("%s, %d SMs, %d MHz nominal (rates assume the nominal clock)\n", p.name, p.multiProcessorCount, mhz);
    {
        double c[4];
        const double l2 = run_bw<0>("L2 -> SM (ld.global.cg.v4)", p.multiProcessorCount, &c[0], mhz);
        const double l1 = run_bw<1>("L1 hit (ld.global.ca.v4)", p.multiProcessorCount, &c[1], mhz);
        const double sm = run_bw<2>("shared memory (ld.shared.v4)", p.multiProcessorCount, &c[2], mhz);
        const double cp = run_bw<3>("L2 -> smem (cp.async.cg 16 B)", p.multiProcessorCount, &c[3], mhz);
        if (argc > 1) { /* argv[1] = path of the JSON record bench.py reads */
            if (FILE* f = fopen(argv[1], "w")) {
                fprintf(f, "{\"gpu\": \"%s\", \"sms\": %d, \"nominal_mhz\": %d, \"l2_read_gbs\": %.1f, \"l1_read_gbs\": %.1f, \"smem_read_gbs\": %.1f, "
                           "\"l2_to_smem_cp_async_gbs\": %.1f, \"how\": \"tools/ubench.cu: 64 MiB L2-resident buffer (ld.global.cg.v4 / cp.async.cg), 64 KiB window per CTA "
                           "(ld.global.ca.v4), 32 KiB tile (ld.shared.v4); CUDA events, 2 to 4 CTAs of 512 threads per SM\"}\n",
                        p.name, p.multiProcessorCount, mhz, l2, l1, sm, cp);
                fclose(f);
            }
        }
    }
    run<0>("VIADD.16x2", p.multiProcessorCount, mhz);
    run<1>("IADD3", p.multiProcessorCount, mhz);
    run<2>("PRMT", p.multiProcessorCount, mhz);
    run<3>("IMAD", p.multiProcessorCount, mhz);
    run<4>("VIMNMX.S16x2", p.multiProcessorCount, mhz);
    run<5>("LOP3", p.multiProcessorCount, mhz);
    run<6>("IMMA.16832", p.multiProcessorCount, mhz);
    return 0;
}
