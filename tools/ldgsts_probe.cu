// ldgsts_probe.cu -- how long does ONE warp take to ISSUE n back-to-back cp.async (16 B per lane) + commit, and how long until
// they have landed?  Pattern 0: 512 contiguous bytes per instruction; pattern 1: four 128-byte segments of four different
// rows (the tile-owner copy of ft_group_kernel).  `warps` warps per CTA do the same concurrently (different rows).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/ldgsts_probe tools/ldgsts_probe.cu && /tmp/ldgsts_probe
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__global__ void __launch_bounds__(512) probe(const uint8_t* __restrict__ table, int n, int pattern, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint8_t* ring = smem + warp * 8192;
    const int rr = lane >> 3, c8 = lane & 7;
    __syncthreads();
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 4; ++rep) { /* last repetition is reported (L2 warm) */
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < n; ++i) {
            const uint32_t row = (blockIdx.x * 16 + warp) * 64 + rep * 0 + i * 4 + (pattern ? rr : 0);
            const uint8_t* src = pattern ? table + static_cast<size_t>(row) * 1024 + warp * 64 + c8 * 16 : table + static_cast<size_t>(row) * 1024 + lane * 16;
            const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(ring)) + (pattern ? rr * 1024 + (i & 7) * 128 + ((c8 ^ (i & 7)) << 4) : (i & 15) * 512 + lane * 16);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        t1 = clock64();
        asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        t2 = clock64();
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) out[0] = t1 - t0, out[1] = t2 - t0;
}

int main() {
    uint8_t* table;
    long long* d_out;
    cudaMalloc(&table, 64ull << 20);
    cudaMemset(table, 1, 64ull << 20);
    cudaMalloc(&d_out, 16);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192);
    for (int pattern = 0; pattern < 2; ++pattern)
        for (int warps : {1, 8, 16})
            for (int n : {1, 2, 4, 8, 16}) {
                probe<<<148, warps * 32, 16 * 8192>>>(table, n, pattern, d_out);
                long long h[2];
                cudaDeviceSynchronize();
                cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
                printf("pattern %d  warps/CTA %2d  n %2d : issue %5lld clk  landed %5lld clk\n", pattern, warps, n, h[0], h[1]);
            }
    return 0;
}
