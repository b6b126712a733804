// tma_gather_probe.cu -- cp.async.bulk.tensor.2d ... tile::gather4 on a B200: semantics and cost.
// A 2-D u8 tensor [rows][1024] (the threat weight table's shape); one instruction fetches the 128-byte column segment `tile` of
// FOUR arbitrary rows into four consecutive 128-byte lines of shared memory, with the tensor map's SWIZZLE_128B applied by
// the TMA engine -- exactly the MN-major atom tcgen05.mma reads (tools/umma_probe.cu).  Checks the landed bytes against the
// expected swizzled layout and times batches of 8 instructions (one K = 32 batch of one column tile) from 1 / 8 / 16 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_gather_probe tools/tma_gather_probe.cu && /tmp/tma_gather_probe [box_rows]
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void gather4(void* dst, const CUtensorMap* map, int col, int r0, int r1, int r2, int r3, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(col), "r"(r0), "r"(r1), "r"(r2), "r"(r3), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}

// each warp: `batches` batches of 32 rows (8 gather4) of column tile `warp % 8` into its own 2 x 4 KB ring
__global__ void __launch_bounds__(512) probe(const __grid_constant__ CUtensorMap map, const uint32_t* __restrict__ rows, int batches, uint8_t* out, long long* clocks) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ __align__(8) uint64_t bars[16][2];
    __shared__ uint32_t keys[16][16 * 32]; /* row indices staged in shared memory, as ft_group_kernel has them */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = lane; i < batches * 32; i += 32) keys[warp][i] = rows[(blockIdx.x * 16 + warp) * batches * 32 + i];
    uint8_t* ring = smem + warp * 8192;
    if (lane == 0) {
        for (int s = 0; s < 2; ++s) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[warp][s])), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    long long t0 = clock64(), t_issue = 0;
    if (lane == 0) {
        const uint32_t* my = keys[warp];
        for (int b = 0; b < batches; ++b) {
            const int st = b & 1;
            if (b >= 2) wait(&bars[warp][st], ((b - 2) >> 1) & 1);
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[warp][st])), "r"(4096) : "memory");
            const long long ti = clock64();
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint4 k = *reinterpret_cast<const uint4*>(my + b * 32 + q * 4);
                gather4(ring + st * 4096 + q * 512, &map, (warp & 7) * 128, k.x, k.y, k.z, k.w, &bars[warp][st]);
            }
            t_issue += clock64() - ti;
        }
        for (int b = batches - 2 < 0 ? 0 : batches - 2; b < batches; ++b) wait(&bars[warp][b & 1], (b >> 1) & 1);
    }
    __syncwarp();
    const long long t1 = clock64();
    if (blockIdx.x == 0 && threadIdx.x == 0) clocks[0] = t1 - t0, clocks[1] = t_issue;
    // dump warp 0's last batch (stage (batches - 1) & 1) of block 0
    if (blockIdx.x == 0 && warp == 0)
        for (int i = lane; i < 4096; i += 32) out[i] = ring[((batches - 1) & 1) * 4096 + i];
}

int main(int argc, char** argv) {
    const int box_rows = argc > 1 ? atoi(argv[1]) : 1; /* 1 is what tile::gather4 wants (4 = illegal instruction at run time) */
    const int n_rows = 65536, batches = 16;
    std::vector<uint8_t> table(size_t(n_rows) * 1024);
    uint32_t s = 7;
    auto rnd = [&]() { s = s * 1664525u + 1013904223u; return s >> 8; };
    for (auto& x : table) x = uint8_t(rnd());
    uint8_t* d_table;
    cudaMalloc(&d_table, table.size());
    cudaMemcpy(d_table, table.data(), table.size(), cudaMemcpyHostToDevice);
    // tensor map through the driver entry point (no link-time dependency on libcuda)
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                 CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    CUtensorMap map;
    const cuuint64_t dims[2] = {1024, cuuint64_t(n_rows)}, strides[1] = {1024};
    const cuuint32_t box[2] = {128, cuuint32_t(box_rows)}, estr[2] = {1, 1};
    const CUresult rc = reinterpret_cast<EncodeFn>(fn)(&map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_table, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                       CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (rc != CUDA_SUCCESS) { printf("cuTensorMapEncodeTiled failed: %d (box rows %d)\n", int(rc), box_rows); return 2; }
    std::vector<uint32_t> rows(size_t(148) * 16 * batches * 32);
    for (auto& r : rows) r = rnd() % n_rows;
    uint32_t* d_rows;
    uint8_t* d_out;
    long long* d_clk;
    cudaMalloc(&d_rows, rows.size() * 4), cudaMalloc(&d_out, 4096), cudaMalloc(&d_clk, 16);
    cudaMemcpy(d_rows, rows.data(), rows.size() * 4, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 8192 + 1024);
    for (int warps : {1, 8, 16}) {
        probe<<<148, warps * 32, 16 * 8192 + 1024>>>(map, d_rows, batches, d_out, d_clk);
        const cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("warps %d: CUDA error %s\n", warps, cudaGetErrorString(e)); return 1; }
        long long clk[2];
        std::vector<uint8_t> got(4096);
        cudaMemcpy(clk, d_clk, sizeof(clk), cudaMemcpyDeviceToHost);
        cudaMemcpy(got.data(), d_out, 4096, cudaMemcpyDeviceToHost);
        // expected: line L (0..31) = row my[(batches-1)*32 + L] of block 0 / warp 0, segment tile 0; 16-byte piece c at piece (c ^ (L & 7))
        int bad = 0;
        const uint32_t* my = rows.data() + size_t(batches - 1) * 32;
        for (int L = 0; L < 32; ++L)
            for (int c = 0; c < 8; ++c)
                for (int b = 0; b < 16; ++b) bad += got[L * 128 + ((c ^ (L & 7)) << 4) + b] != table[size_t(my[L]) * 1024 + c * 16 + b];
        printf("box rows %d, %2d warps/CTA: %d of 4096 bytes differ from the swizzled layout; %lld clk per batch of 32 rows x 128 B (issue of 8 gather4: %lld clk)\n", box_rows, warps,
               bad, clk[0] / batches, clk[1] / batches);
    }
    return 0;
}
