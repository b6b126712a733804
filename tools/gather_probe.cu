// gather_probe.cu -- what can a B200 sustain when every SM gathers scattered 1 KB rows from HBM with cp.async (16 B per lane,
// LDGSTS.BYPASS) into shared memory, as head_umma_kernel's gather warps do?  No consumer: the copies land in a per-warp ring and are
// only waited for.  Parameters: bytes of a row requested per copy instruction sequence ("piece": the row is fetched in 1024 / piece
// passes, each over all 128 rows of the tile; 128, 256 or 512), warps per SM, groups in flight per warp.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/gather_probe tools/gather_probe.cu && /tmp/gather_probe
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <random>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// One "tile" = 128 rows.  A warp takes tiles warp_global, + n_warps, ...; per tile it makes 1024 / PIECE passes over the 128 rows and in
// each pass copies PIECE bytes of every row (PIECE / 16 lanes per row, 512 / PIECE rows per instruction).
template <int PIECE, int INFLIGHT>
__global__ void __launch_bounds__(512) gather(const uint8_t* __restrict__ act, const uint32_t* __restrict__ order, uint32_t n_tiles, int warps_per_cta) {
    extern __shared__ __align__(16) uint8_t ring[]; // per warp: INFLIGHT + 1 slots of 2 KB (one instruction group = 4 instructions = 2 KB)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (warp >= warps_per_cta) return;
    constexpr int kLanesPerRow = PIECE / 16, kRowsPerCopy = 32 / kLanesPerRow;
    const int rr = lane / kLanesPerRow, piece = lane % kLanesPerRow;
    uint8_t* mine = ring + warp * (INFLIGHT + 1) * 2048;
    const uint32_t n_warps = gridDim.x * warps_per_cta, me = blockIdx.x * warps_per_cta + warp;
    uint32_t slot = 0;
    for (uint32_t tile = me; tile < n_tiles; tile += n_warps) {
        const uint32_t* ord = order + static_cast<size_t>(tile) * 128;
        for (int pass = 0; pass < 1024 / PIECE; ++pass) {
            for (int r0 = 0; r0 < 128; r0 += 4 * kRowsPerCopy) { // one group = 4 copy instructions
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t row = __ldg(ord + r0 + i * kRowsPerCopy + rr);
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(mine + slot * 2048 + i * 512 + lane * 16)),
                                 "l"(act + static_cast<size_t>(row) * 1024 + pass * PIECE + piece * 16)
                                 : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group %0;" ::"n"(INFLIGHT) : "memory");
                slot = (slot + 1) % (INFLIGHT + 1);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
}

template <int PIECE, int INFLIGHT>
void run(const uint8_t* act, const uint32_t* order, uint32_t n_rows, int sms, int warps_per_cta, int ctas_per_sm) {
    const size_t smem = static_cast<size_t>(warps_per_cta) * (INFLIGHT + 1) * 2048;
    if (smem * ctas_per_sm > 200 * 1024) {
        printf("piece %d, %d warps x %d, %d in flight: does not fit shared memory\n", PIECE, warps_per_cta, ctas_per_sm, INFLIGHT);
        return;
    }
    cudaFuncSetAttribute(gather<PIECE, INFLIGHT>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    float best = 1e9f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(e0);
        gather<PIECE, INFLIGHT><<<sms * ctas_per_sm, 512, smem>>>(act, order, n_rows / 128, warps_per_cta);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        best = std::min(best, ms);
    }
    const cudaError_t e = cudaGetLastError();
    printf("piece %4d B  %2d warps x %d CTA/SM  %2d groups (2 KB) in flight per warp = %5.0f KB per SM : %7.1f us  %6.0f GB/s%s\n", PIECE, warps_per_cta,
           ctas_per_sm, INFLIGHT, warps_per_cta * ctas_per_sm * INFLIGHT * 2.0, best * 1e3, n_rows * 1024.0 / (best * 1e-3) / 1e9,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
    fflush(stdout);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const uint32_t n_rows = 1u << 20;
    uint8_t* act;
    uint32_t* order;
    cudaMalloc(&act, static_cast<size_t>(n_rows) * 1024), cudaMalloc(&order, n_rows * 4);
    cudaMemset(act, 1, static_cast<size_t>(n_rows) * 1024);
    std::vector<uint32_t> h(n_rows);
    for (uint32_t i = 0; i < n_rows; ++i) h[i] = i;
    std::mt19937 rng(5);
    std::shuffle(h.begin(), h.end(), rng);
    cudaMemcpy(order, h.data(), n_rows * 4, cudaMemcpyHostToDevice);
    printf("%s: gather of %u scattered 1 KB rows (1 GiB, larger than L2) by cp.async 16 B per lane\n", p.name, n_rows);
    const int sms = p.multiProcessorCount;
    fflush(stdout);
    run<128, 6>(act, order, n_rows, sms, 3, 1);
    run<256, 6>(act, order, n_rows, sms, 3, 1);
    run<512, 6>(act, order, n_rows, sms, 3, 1);
    run<128, 3>(act, order, n_rows, sms, 3, 1);
    run<128, 6>(act, order, n_rows, sms, 2, 1);
    run<128, 6>(act, order, n_rows, sms, 4, 1);
    run<128, 6>(act, order, n_rows, sms, 8, 1);
    run<128, 4>(act, order, n_rows, sms, 16, 1);
    run<128, 2>(act, order, n_rows, sms, 16, 2);
    run<256, 6>(act, order, n_rows, sms, 8, 1);
    run<512, 6>(act, order, n_rows, sms, 8, 1);
    run<512, 4>(act, order, n_rows, sms, 16, 1);
    return 0;
}
