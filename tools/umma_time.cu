// umma_time.cu -- how long does tcgen05.mma.kind::i8 (M = 128, K = 32, A MN-major SWIZZLE_128B from shared memory) take as a
// function of N, and how does it compare with a K-major A?  One CTA, one issuing thread, `reps` MMAs back to back on the
// same operands, clock64 around issue ... tcgen05.commit -> mbarrier wait.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_time tools/umma_time.cu && /tmp/umma_time
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    return static_cast<uint64_t>((addr >> 4) & 0x3FFF) | static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16 | static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32
         | uint64_t{1} << 46 | static_cast<uint64_t>(layout) << 61;
}

// mode 0: A MN-major SW128, B MN-major NONE ([k][N] bytes, N <= 16 only -> for larger N use mode 1 layouts)
// mode 1: A K-major SW128 (128 rows x 32 B... one swizzle row of 128 B holds 4 k-steps), B K-major SW128
__global__ void __launch_bounds__(128) timing(int n_dim, int mode, int reps, int tiles, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_base;
    uint8_t* base = smem + ((1024u - (smem_u32(smem) & 1023u)) & 1023u);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(base)[i] = 0x01010101u * (i & 3);
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(512));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | ((mode == 0 ? 1u : 0u) << 15) | ((mode == 0 ? 1u : 0u) << 16) | ((n_dim >> 3) << 17) | ((128u >> 4) << 24);
        uint8_t* a = base;               /* `tiles` A tiles of 4 KB */
        uint8_t* b = base + 64 * 1024;   /* B: up to 256 x 32 bytes */
        uint32_t phase = 0;
        for (int round = 0; round < 3; ++round) {
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const int tile = r % tiles;
                const uint64_t da = mode == 0 ? make_desc(smem_u32(a + tile * 4096), 1024, 1024, 2) : make_desc(smem_u32(a + tile * 16384), 16, 1024, 2);
                const uint64_t db = mode == 0 ? make_desc(smem_u32(b), 128, 128, 0) : make_desc(smem_u32(b), 16, 1024, 2);
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}" ::"r"(tmem + (tile % 8) * 32),
                    "l"(da), "l"(db), "r"(idesc), "r"(1), "r"(0), "r"(0), "r"(0), "r"(0)
                    : "memory");
            }
            const long long t1 = clock64();
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
            uint32_t ok = 0;
            while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done)), "r"(phase) : "memory");
            phase ^= 1;
            const long long t2 = clock64();
            out[round * 2] = t1 - t0;
            out[round * 2 + 1] = t2 - t0;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512));
}

int main() {
    long long* d_out;
    cudaMalloc(&d_out, 64);
    cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 256;
    printf("tcgen05.mma kind::i8 M=128 K=32, %d back-to-back MMAs from one thread: clocks per MMA (issue only | until commit completes)\n", reps);
    for (int mode = 0; mode < 2; ++mode)
        for (int n : {16, 32, 64, 128, 256}) {
            if (mode == 0 && n > 16) continue; /* the plain [k][16] B layout is one 16-byte unit wide */
            for (int tiles : {1, 8}) {
                timing<<<1, 128, 100 * 1024>>>(n, mode, reps, tiles, d_out);
                long long h[6];
                if (cudaDeviceSynchronize() != cudaSuccess) { printf("mode %d N %d: CUDA error %s\n", mode, n, cudaGetErrorString(cudaGetLastError())); return 1; }
                cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
                printf("%s  N=%3d  A tiles=%d : issue %.1f  complete %.1f clk/MMA\n", mode == 0 ? "A MN-major SW128, B MN-major" : "A K-major  SW128, B K-major ", n, tiles,
                       double(h[4]) / reps, double(h[5]) / reps);
            }
        }
    return 0;
}
