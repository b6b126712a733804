// umma_probe.cu -- checks, on a B200, the tcgen05 operand layouts that ft_group_kernel relies on
// (DESIGN.md section 4): D[128 x 16] (s32, TMEM) += A[128 x 32] (u8, MN-major in shared memory: weight
// rows copied 16 bytes at a time by cp.async) * B[32 x 16] (u8, MN-major, no swizzle: a plain [k][16] array).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_probe tools/umma_probe.cu
//   /tmp/umma_probe <variant>      0: A SWIZZLE_128B   1: A SWIZZLE_NONE (interleaved core matrices)
// Prints the number of mismatching outputs against a CPU contraction (0 = the layout is understood).
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int kM = 128, kN = 16, kRows = 96; /* three k-steps of 32 rows */

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46; /* descriptor version: Blackwell */
    d |= static_cast<uint64_t>(layout) << 61;
    return d;
}

// rows: [kRows][128] u8 in global memory (one M tile of a weight row each), sel: [kRows][16] u8
__global__ void __launch_bounds__(128) probe(const uint8_t* __restrict__ rows, const uint8_t* __restrict__ sel, int32_t* __restrict__ out, int variant) {
    __shared__ __align__(1024) uint8_t a_smem[kRows * kM];   /* per k-step: 4 k-groups x 1 KB */
    __shared__ __align__(128) uint8_t b_smem[kRows * kN];
    __shared__ __align__(8) uint64_t done;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base)), "r"(32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // A: every row is 128 B = eight 16-byte pieces; thread t copies piece (t & 7) of rows t >> 3, + 16, ...
    for (int r = tid >> 3; r < kRows; r += 16) {
        const int c = tid & 7, kg = r >> 3, k8 = r & 7;
        uint32_t off;
        if (variant == 0) off = kg * 1024 + k8 * 128 + ((c ^ k8) * 16);     /* SW128: 8 k x 128 B atom, 16-byte pieces XORed with k % 8 */
        else off = (r >> 5) * 4096 + c * 128 + k8 * 16 + ((kg & 3) * 1024); /* NONE: core matrix = 8 k x 16 B; m-groups 128 B apart, k-groups 1 KB */
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(a_smem + off)), "l"(rows + r * kM + c * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < kRows * kN / 16; i += 128) reinterpret_cast<uint4*>(b_smem)[i] = reinterpret_cast<const uint4*>(sel)[i];
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); /* generic-proxy writes -> visible to the tensor core's async proxy */
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        // instruction descriptor: D = s32, A = u8, B = u8, both MN-major, N = 16, M = 128
        const uint32_t idesc = (2u << 4) | (0u << 7) | (0u << 10) | (1u << 15) | (1u << 16) | ((kN >> 3) << 17) | ((kM >> 4) << 24);
        for (int ks = 0; ks < kRows / 32; ++ks) {
            uint64_t da, db;
            if (variant == 0) da = make_desc(smem_u32(a_smem + ks * 4096), 1024, 1024, 2);
            else da = make_desc(smem_u32(a_smem + ks * 4096), 1024, 128, 0); /* NONE, MN-major: SBO = m-group stride, LBO = k-group stride */
            db = make_desc(smem_u32(b_smem + ks * 32 * kN), 128, 128, 0);
            const uint32_t accumulate = ks > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    // everybody waits for the MMAs
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    uint32_t v[16];
    const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                   "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int n = 0; n < kN; ++n) out[tid * kN + n] = ok ? static_cast<int32_t>(v[n]) : -1;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main(int argc, char** argv) {
    const int variant = argc > 1 ? atoi(argv[1]) : 0;
    std::vector<uint8_t> rows(kRows * kM), sel(kRows * kN);
    uint32_t s = 12345;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto& x : rows) x = static_cast<uint8_t>(rnd());
    for (auto& x : sel) x = static_cast<uint8_t>(rnd() % 3);
    std::vector<int32_t> want(kM * kN, 0), got(kM * kN, 0);
    for (int k = 0; k < kRows; ++k)
        for (int m = 0; m < kM; ++m)
            for (int n = 0; n < kN; ++n) want[m * kN + n] += int(rows[k * kM + m]) * int(sel[k * kN + n]);
    uint8_t *d_rows, *d_sel;
    int32_t* d_out;
    cudaMalloc(&d_rows, rows.size()), cudaMalloc(&d_sel, sel.size()), cudaMalloc(&d_out, got.size() * 4);
    cudaMemcpy(d_rows, rows.data(), rows.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_sel, sel.data(), sel.size(), cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(d_rows, d_sel, d_out, variant);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("variant %d: CUDA error: %s\n", variant, cudaGetErrorString(e));
        return 2;
    }
    cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t i = 0; i < got.size(); ++i) bad += got[i] != want[i];
    printf("variant %d: %d of %zu outputs differ (got[0..3] = %d %d %d %d, want %d %d %d %d)\n", variant, bad, got.size(), got[0], got[1], got[2],
           got[3], want[0], want[1], want[2], want[3]);
    return bad ? 1 : 0;
}
