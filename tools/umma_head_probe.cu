// umma_head_probe.cu -- checks, on a B200, what head_umma_kernel relies on (dense head):
//   D[128 positions x 32 outputs] (s32, TMEM) += A[128 x K] (u8, K-major, SWIZZLE_128B: 8-row x 128-byte atoms written 16 bytes at a
//   time by cp.async) * B[32 x K] (s8, K-major, SWIZZLE_128B), K = 128 as four tcgen05.mma.kind::i8 of K = 32 whose descriptors
//   advance 32 bytes inside the atom; and that tcgen05.ld.16x256b.x4 hands a warp the mma.sync C-fragment layout
//   (lane (g, t): registers 4 nt + 2 h + c = row g + 8 h, column 8 nt + 2 t + c of its 16 lanes).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/umma_head_probe tools/umma_head_probe.cu && /tmp/umma_head_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

constexpr int kM = 128, kN = 32, kK = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((addr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>((lbo >> 4) & 0x3FFF) << 16;
    d |= static_cast<uint64_t>((sbo >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(layout) << 61;
    return d;
}

__global__ void __launch_bounds__(128) probe(const uint8_t* __restrict__ a, const int8_t* __restrict__ b, int32_t* __restrict__ out) {
    __shared__ __align__(1024) uint8_t a_smem[kM * kK];
    __shared__ __align__(1024) uint8_t b_smem[kN * kK];
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
    // row r, 16-byte chunk c -> atom r / 8 (1 KB), line r % 8 (128 B), chunk c ^ (r % 8)
    for (int i = tid; i < kM * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(a_smem + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4))), "l"(a + r * kK + c * 16) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    for (int i = tid; i < kN * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(b_smem + (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(b + r * kK + c * 16);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;
    if (tid == 0) {
        // D = s32, A = u8 (format 0), B = s8 (format 1), both K-major, N = 32, M = 128
        const uint32_t idesc = (2u << 4) | (0u << 7) | (1u << 10) | (0u << 15) | (0u << 16) | ((kN >> 3) << 17) | ((kM >> 4) << 24);
        for (int kk = 0; kk < kK / 32; ++kk) {
            const uint64_t da = make_desc(smem_u32(a_smem) + 32 * kk, 16, 1024, 2);
            const uint64_t db = make_desc(smem_u32(b_smem) + 32 * kk, 16, 1024, 2);
            const uint32_t accumulate = kk > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
                ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0)
                : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done)) : "memory");
    }
    uint32_t ok = 0;
    for (int spin = 0; spin < (1 << 22) && !ok; ++spin)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(&done)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const int g = lane >> 2, t = lane & 3;
    for (int mt = 0; mt < 2; ++mt) {
        uint32_t v[16];
        const uint32_t taddr = tmem + (static_cast<uint32_t>(warp * 32 + 16 * mt) << 16);
        asm volatile("tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                       "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                     : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int nt = 0; nt < 4; ++nt)
            for (int h = 0; h < 2; ++h)
                for (int c = 0; c < 2; ++c) {
                    const int row = warp * 32 + 16 * mt + g + 8 * h, col = 8 * nt + 2 * t + c;
                    out[row * kN + col] = ok ? static_cast<int32_t>(v[4 * nt + 2 * h + c]) : -1;
                }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(32));
}

int main() {
    std::vector<uint8_t> a(kM * kK);
    std::vector<int8_t> b(kN * kK);
    uint32_t s = 777;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto& x : a) x = static_cast<uint8_t>(rnd());
    for (auto& x : b) x = static_cast<int8_t>(rnd());
    std::vector<int32_t> want(kM * kN, 0), got(kM * kN, 0);
    for (int m = 0; m < kM; ++m)
        for (int n = 0; n < kN; ++n)
            for (int k = 0; k < kK; ++k) want[m * kN + n] += int(a[m * kK + k]) * int(b[n * kK + k]);
    uint8_t* d_a;
    int8_t* d_b;
    int32_t* d_out;
    cudaMalloc(&d_a, a.size()), cudaMalloc(&d_b, b.size()), cudaMalloc(&d_out, got.size() * 4);
    cudaMemcpy(d_a, a.data(), a.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(d_b, b.data(), b.size(), cudaMemcpyHostToDevice);
    probe<<<1, 128>>>(d_a, d_b, d_out);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        printf("CUDA error: %s\n", cudaGetErrorString(e));
        return 2;
    }
    cudaMemcpy(got.data(), d_out, got.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t i = 0; i < got.size(); ++i) bad += got[i] != want[i];
    printf("K-major SW128 u8 x s8, 16x256b.x4 readback: %d of %zu outputs differ (got[0..3] = %d %d %d %d, want %d %d %d %d)\n", bad, got.size(), got[0],
           got[1], got[2], got[3], want[0], want[1], want[2], want[3]);
    return bad ? 1 : 0;
}
