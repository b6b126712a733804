/*
 * kernels.cu -- sm_100a kernels of the batched NNUE evaluator.
 *
 * Work decomposition:
 *   ft_*   one WARP owns one position.  Lane l decodes piece l of the packed board, the warp
 *          builds both perspectives' feature-index lists in shared memory, then sums the
 *          feature-transformer rows with 128-bit coalesced loads: a warp-wide load covers 512
 *          contiguous bytes of one row, lane l accumulates logical elements [16l, 16l+16) and
 *          [512+16l, 512+16l+16) -- both members of the 16 activation pairs it multiplies.
 *          Replaces resetPsqAccumulator / addThreatFeatures / applyThreatRows / activateFt
 *          (src/eval/nnue_state.cpp:89-145,309-354,440-456; src/eval/nnue/arch/multilayer.h:92-152).
 *   head   one WARP owns 16 positions.  L1 is an int8 tensor-core contraction
 *          (mma.sync m16n8k32 u8 x s8 -> s32, exactly vpdpbusd's arithmetic), L2/L3 are int32
 *          CUDA-core loops.  Replaces propagateL1/L2/L3 (multilayer.h:154-490).
 *
 * Integer semantics: all accumulator arithmetic is modulo 2^16, all dense-layer arithmetic
 * modulo 2^32 (unsigned types are used wherever a sum may wrap), shifts of signed values are
 * arithmetic, the final division truncates toward zero.
 */
#include "kernels.cuh"

#include "sp_delta.h"

namespace sp::gpu {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kPsqGroup = 4;      /* PSQ rows fetched per batch of loads (16 x LDG.128 in flight per lane) */
constexpr int kThrGroup = 8;      /* threat rows per batch (16 x LDG.128 in flight per lane) */
constexpr int kPsqListCap = 40;   /* 32 pieces + bias row, padded to a multiple of kPsqGroup */
constexpr int kThrListCap = SP_MAX_THREAT_INDICES;
constexpr unsigned kFull = 0xFFFFFFFFu;

struct WarpScratch {
    uint16_t thr[2][kThrListCap];
    uint16_t psq[2][kPsqListCap];
    uint8_t mailbox[64];
    int n_thr[2];
    int n_psq[2];
};

/* Board as the shared feature code (sp_features.h, sp_delta.h) wants to see it. */
struct BoardView {
    const uint8_t* mailbox;
    uint64_t occ;
    uint64_t pawns[2];
    int king[2];
    int stm;
};

/* What lane `lane` knows after decoding: warp-uniform board facts + its own piece. */
struct Decoded {
    BoardView view;
    int n_pieces;
    int piece; /* kNoPiece for lanes >= n_pieces */
    int sq;
    bool ok;
};

__device__ __forceinline__ int nth_set_bit(uint64_t m, int n) {
    const uint32_t lo = static_cast<uint32_t>(m), hi = static_cast<uint32_t>(m >> 32);
    const int c = __popc(lo);
    return n < c ? static_cast<int>(__fns(lo, 0, n + 1)) : 32 + static_cast<int>(__fns(hi, 0, n - c + 1));
}

__device__ __forceinline__ uint64_t warp_or64(uint64_t v) {
    const uint32_t lo = __reduce_or_sync(kFull, static_cast<uint32_t>(v));
    const uint32_t hi = __reduce_or_sync(kFull, static_cast<uint32_t>(v >> 32));
    return static_cast<uint64_t>(hi) << 32 | lo;
}

/* Decode one marlinformat record (src/datagen/marlinformat.h:43-77) cooperatively.
 * The mailbox is written to shared memory; everything else stays in registers. */
__device__ __forceinline__ Decoded decode_board(const SpPackedBoard* board, int lane, uint8_t* mailbox) {
    const uint4* p = reinterpret_cast<const uint4*>(board);
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    Decoded d;
    d.view.mailbox = mailbox;
    d.view.occ = static_cast<uint64_t>(lo.y) << 32 | lo.x;
    d.view.stm = (hi.z & 0x80) ? kBlack : kWhite;
    d.n_pieces = __popcll(d.view.occ);
    d.ok = d.n_pieces <= 32 && d.n_pieces >= 2;
    d.piece = kNoPiece;
    d.sq = 0;
    reinterpret_cast<uint16_t*>(mailbox)[lane] = static_cast<uint16_t>(kNoPiece | kNoPiece << 8);
    __syncwarp();
    if (lane < d.n_pieces && d.ok) {
        d.sq = nth_set_bit(d.view.occ, lane);
        const uint32_t word = lane < 8 ? lo.z : (lane < 16 ? lo.w : (lane < 24 ? hi.x : hi.y));
        const uint32_t nib = (word >> ((lane & 7) * 4)) & 0xF;
        uint32_t type = nib & 7;
        if (type == 6) type = kRook;
        if (type > kKing) {
            d.ok = false;
        } else {
            d.piece = static_cast<int>(type << 1 | ((nib & 8) ? kBlack : kWhite));
            mailbox[d.sq] = static_cast<uint8_t>(d.piece);
        }
    }
    d.ok = __all_sync(kFull, d.ok);
    const unsigned bk = __ballot_sync(kFull, d.piece == (kKing << 1 | kBlack));
    const unsigned wk = __ballot_sync(kFull, d.piece == (kKing << 1 | kWhite));
    if (__popc(bk) != 1 || __popc(wk) != 1) d.ok = false;
    d.view.king[kBlack] = __shfl_sync(kFull, d.sq, bk ? __ffs(bk) - 1 : 0);
    d.view.king[kWhite] = __shfl_sync(kFull, d.sq, wk ? __ffs(wk) - 1 : 0);
    d.view.pawns[kBlack] = warp_or64(d.piece == (kPawn << 1 | kBlack) ? bit(d.sq) : 0);
    d.view.pawns[kWhite] = warp_or64(d.piece == (kPawn << 1 | kWhite) ? bit(d.sq) : 0);
    __syncwarp();
    return d;
}

/* Build both perspectives' full feature lists (nnue_state.cpp:309-354, 440-449).  Returns false
 * if a threat list would exceed the reference's own bound of 256 entries. */
__device__ __forceinline__ bool build_full_lists(const FeatureTables& t, const Decoded& d, int lane, WarpScratch& ws) {
    if (lane < 2) ws.n_thr[lane] = 0;
    __syncwarp();
    if (d.piece != kNoPiece) {
        ws.psq[kBlack][lane] = static_cast<uint16_t>(psq_index(t, kBlack, d.piece, d.sq, d.view.king[kBlack]));
        ws.psq[kWhite][lane] = static_cast<uint16_t>(psq_index(t, kWhite, d.piece, d.sq, d.view.king[kWhite]));
        square_threat_features(t, d.view, d.sq, [&](int c, uint32_t idx) {
            const int at = atomicAdd(&ws.n_thr[c], 1);
            if (at < kThrListCap) ws.thr[c][at] = static_cast<uint16_t>(idx);
        });
    }
    __syncwarp();
    const int n0 = ws.n_thr[0], n1 = ws.n_thr[1];
    if (n0 > kThrListCap || n1 > kThrListCap) return false;
    /* bias row + zero-row padding so the row loops run in whole batches */
    const int np = d.n_pieces;
    const int np_pad = (np + 1 + kPsqGroup - 1) / kPsqGroup * kPsqGroup;
    if (lane < 2) {
        ws.psq[lane][np] = kPsqBiasRow;
        for (int i = np + 1; i < np_pad; ++i) ws.psq[lane][i] = kPsqZeroRow;
        ws.n_psq[lane] = np_pad;
        const int n = lane == 0 ? n0 : n1;
        const int n_pad = (n + kThrGroup - 1) / kThrGroup * kThrGroup;
        for (int i = n; i < n_pad; ++i) ws.thr[lane][i] = static_cast<uint16_t>(kThrZeroRow);
        ws.n_thr[lane] = n_pad;
    }
    __syncwarp();
    return true;
}

/* ------------------------------------------------------------------ accumulators in registers */

/* One perspective, one lane: 32 logical elements.
 *   plo/phi  int16 PSQ sums, one 32-bit register per element; only the low 16 bits are meaningful
 *            (plo += word leaves garbage above bit 15, phi += word >> 16 likewise)
 *   te/to    sums of BIASED (+128) threat bytes, two 16-bit fields per register: even bytes of a
 *            word in te, odd bytes in to.  <= 256 rows x 255 < 2^16, so fields never carry. */
struct LaneAcc {
    uint32_t plo[16], phi[16];
    uint32_t te[8], to[8];
    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < 16; ++i) plo[i] = phi[i] = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) te[i] = to[i] = 0;
    }
};

template <int kSign>
__device__ __forceinline__ void add_psq_chunks(LaneAcc& a, const uint4 (&c)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint32_t w[4] = {c[k].x, c[k].y, c[k].z, c[k].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (kSign > 0) {
                a.plo[k * 4 + t] += w[t];
                a.phi[k * 4 + t] += w[t] >> 16;
            } else {
                a.plo[k * 4 + t] -= w[t];
                a.phi[k * 4 + t] -= w[t] >> 16;
            }
        }
    }
}

__device__ __forceinline__ void add_thr_chunks(uint32_t (&te)[8], uint32_t (&to)[8], const uint4 (&c)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const uint32_t w[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            te[u * 4 + t] += w[t] & 0x00FF00FFu;
            to[u * 4 + t] += __byte_perm(w[t], 0, 0x4341); /* bytes 1 and 3 into the two 16-bit fields */
        }
    }
}

__device__ __forceinline__ void load_psq_row(const DeviceNet& net, uint32_t row, int lane, uint4 (&c)[4]) {
    const uint4* r = net.psq + static_cast<size_t>(row) * 128 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = __ldg(r + 32 * k);
}

__device__ __forceinline__ void load_thr_row(const DeviceNet& net, uint32_t row, int lane, uint4 (&c)[2]) {
    const uint4* r = net.thr + static_cast<size_t>(row) * 64 + lane;
    c[0] = __ldg(r);
    c[1] = __ldg(r + 32);
}

/* Sum `n_psq` PSQ rows and `n_thr` threat rows (both multiples of their batch size). */
__device__ __forceinline__ void accumulate_lists(
    const DeviceNet& net, const uint16_t* psq_list, int n_psq, const uint16_t* thr_list, int n_thr, int lane, LaneAcc& a) {
    for (int i = 0; i < n_psq; i += kPsqGroup) {
        uint4 c[kPsqGroup][4];
#pragma unroll
        for (int j = 0; j < kPsqGroup; ++j) load_psq_row(net, psq_list[i + j], lane, c[j]);
#pragma unroll
        for (int j = 0; j < kPsqGroup; ++j) add_psq_chunks<1>(a, c[j]);
    }
    for (int i = 0; i < n_thr; i += kThrGroup) {
        uint4 c[kThrGroup][2];
#pragma unroll
        for (int j = 0; j < kThrGroup; ++j) load_thr_row(net, thr_list[i + j], lane, c[j]);
#pragma unroll
        for (int j = 0; j < kThrGroup; ++j) add_thr_chunks(a.te, a.to, c[j]);
    }
}

/* Collapse a LaneAcc into 32 wrapped int16 values: v[0..15] = first-half elements 16l + j,
 * v[16..31] = second-half elements 512 + 16l + j.  `n_thr` rows contributed a +128 bias each. */
__device__ __forceinline__ void finalize(const LaneAcc& a, int n_thr, int (&v)[32]) {
    const uint32_t corr = static_cast<uint32_t>(n_thr) * 128u;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const int pi = (h * 2 + (j >> 3)) * 4 + ((j & 7) >> 1);
            const uint32_t psq = (j & 1) ? a.phi[pi] : a.plo[pi];
            const uint32_t field = (j & 1) ? a.to[h * 4 + (j >> 2)] : a.te[h * 4 + (j >> 2)];
            const uint32_t thr = field >> (((j & 3) >> 1) * 16);
            v[h * 16 + j] = static_cast<int16_t>(static_cast<uint16_t>(psq + thr - corr));
        }
    }
}

/* activateFt, multilayer.h:92-152: 16 outputs of this lane for one perspective, packed little-endian. */
__device__ __forceinline__ uint4 activate(const int (&v)[32]) {
    uint32_t out[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 16; ++j) {
        int a = v[j], d = v[16 + j];
        a = min(max(a, 0), 255);
        d = min(d, 255);
        int p = ((a << 7) * d) >> 16; /* signed mulhi: arithmetic shift floors */
        p = min(max(p, 0), 255);      /* packus */
        out[j >> 2] |= static_cast<uint32_t>(p) << ((j & 3) * 8);
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ void pack_acc(const int (&v)[32], uint4 (&q)[4]) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i)
        w[i] = (static_cast<uint32_t>(v[2 * i]) & 0xFFFFu) | (static_cast<uint32_t>(v[2 * i + 1]) << 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) q[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
}

__device__ __forceinline__ void unpack_acc(const uint4 (&q)[4], int (&v)[32]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t w[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            v[(i * 4 + t) * 2] = static_cast<int16_t>(w[t] & 0xFFFFu);
            v[(i * 4 + t) * 2 + 1] = static_cast<int16_t>(w[t] >> 16);
        }
    }
}

__device__ __forceinline__ void flag_error(DeviceStatus* status, int bits) { atomicOr(&status->error, bits); }

/* ------------------------------------------------------------------ full refresh: boards -> activations */

__global__ void __launch_bounds__(kThreads)
ft_full_kernel(DeviceNet net, const SpPackedBoard* __restrict__ boards, size_t n, uint8_t* __restrict__ act,
               uint8_t* __restrict__ bucket, DeviceStatus* status) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const FeatureTables& t = *net.tables;
    const size_t stride = static_cast<size_t>(gridDim.x) * kWarpsPerCta;
    for (size_t pos = static_cast<size_t>(blockIdx.x) * kWarpsPerCta + warp; pos < n; pos += stride) {
        const Decoded d = decode_board(boards + pos, lane, ws.mailbox);
        bool ok = d.ok;
        if (!ok) {
            if (lane == 0) flag_error(status, kErrBadBoard);
        } else if (!build_full_lists(t, d, lane, ws)) {
            ok = false;
            if (lane == 0) flag_error(status, kErrCapacity);
        }
        if (!ok) {
            if (lane == 0) bucket[pos] = 0xFF;
            __syncwarp();
            continue;
        }
        uint4* row = reinterpret_cast<uint4*>(act + pos * SP_L1_SIZE);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            LaneAcc a;
            a.clear();
            accumulate_lists(net, ws.psq[c], ws.n_psq[c], ws.thr[c], ws.n_thr[c], lane, a);
            int v[32];
            finalize(a, ws.n_thr[c], v);
            const int half = c == d.view.stm ? 0 : 1; /* side to move first, nnue_state.cpp:405-419 */
            row[half * 32 + lane] = activate(v);
        }
        if (lane == 0) bucket[pos] = static_cast<uint8_t>(output_bucket(d.view.occ));
        __syncwarp();
    }
}

/* ------------------------------------------------------------------ dense head */

__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kHeadWarps = 4;
constexpr int kSkipStride = 65; /* int32 words per row, padded against bank conflicts */

/*
 * The contraction index k may be visited in any order as long as A and B agree.  Per 64-wide
 * k-step, lane (g = lane / 4, t = lane % 4) loads 16 contiguous activation bytes of rows g and
 * g + 8 (k = 64 s + 16 t ...) and, for each of the 4 k-quads inside them, 16 contiguous weight
 * bytes = outputs 4 g .. 4 g + 3 of that quad ([k/4][o][k%4] layout, multilayer.h:180-196).  MMA
 * column g of n-tile nt is therefore output o = 4 g + nt, and the C fragment of lane (g, t)
 * holds outputs 8 t + nt and 8 t + 4 + nt of rows g and g + 8.
 */
__global__ void __launch_bounds__(kHeadWarps * 32)
head_kernel(DeviceNet net, const uint8_t* __restrict__ act, const uint8_t* __restrict__ bucket, size_t n,
            int32_t* __restrict__ out) {
    __shared__ int skip[kHeadWarps][16][kSkipStride];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = lane >> 2, t = lane & 3;
    const size_t tile = static_cast<size_t>(blockIdx.x) * kHeadWarps + warp;
    const size_t base = tile * 16;
    if (base >= n) return;
    const size_t last = n - 1;
    const size_t r0 = min(base + g, last), r1 = min(base + g + 8, last);
    /* buckets of the 16 rows: lane i < 16 holds row i's */
    int my_bucket = 0xFF;
    if (lane < 16) my_bucket = bucket[min(base + lane, last)];
    const int b0 = __shfl_sync(kFull, my_bucket, g), b1 = __shfl_sync(kFull, my_bucket, g + 8);
    unsigned present = __reduce_or_sync(kFull, my_bucket < SP_OUTPUT_BUCKETS ? 1u << my_bucket : 0u);

    const uint4* a_row0 = reinterpret_cast<const uint4*>(act + r0 * SP_L1_SIZE) + t;
    const uint4* a_row1 = reinterpret_cast<const uint4*>(act + r1 * SP_L1_SIZE) + t;

    while (present) {
        const int b = __ffs(present) - 1;
        present &= present - 1;
        int c[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) c[i][j] = 0;
        const uint4* w = reinterpret_cast<const uint4*>(net.l1_w + static_cast<size_t>(b) * SP_L1_SIZE * SP_L2_SIZE) + g;
#pragma unroll 4
        for (int s = 0; s < SP_L1_SIZE / 64; ++s) {
            const uint4 alo = __ldg(a_row0 + s * 4), ahi = __ldg(a_row1 + s * 4);
            uint4 q[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) q[j] = __ldg(w + static_cast<size_t>(s * 16 + t * 4 + j) * 8);
            const uint32_t al[4] = {alo.x, alo.y, alo.z, alo.w}, ah[4] = {ahi.x, ahi.y, ahi.z, ahi.w};
#pragma unroll
            for (int m = 0; m < 2; ++m) {
                const uint32_t q0[4] = {q[2 * m].x, q[2 * m].y, q[2 * m].z, q[2 * m].w};
                const uint32_t q1[4] = {q[2 * m + 1].x, q[2 * m + 1].y, q[2 * m + 1].z, q[2 * m + 1].w};
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) mma_u8s8(c[nt], al[2 * m], ah[2 * m], al[2 * m + 1], ah[2 * m + 1], q0[nt], q1[nt]);
            }
        }
        /* L1 epilogue + dual activation, multilayer.h:219-256 (kShift = -2) */
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            if ((hrow ? b1 : b0) != b) continue;
            const int r = g + 8 * hrow;
#pragma unroll
            for (int nt = 0; nt < 4; ++nt)
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    const int o = 8 * t + 4 * cc + nt;
                    const int x = static_cast<int>(static_cast<uint32_t>(c[nt][hrow * 2 + cc] >> 2)
                                                   + static_cast<uint32_t>(__ldg(net.l1_b + b * SP_L2_SIZE + o)));
                    const int cr = min(max(x, 0), 4096);
                    int sq = static_cast<int>(static_cast<uint32_t>(x) * static_cast<uint32_t>(x)); /* wraps BEFORE the min */
                    sq = min(sq, 16777216);
                    skip[warp][r][o] = cr << 6;
                    skip[warp][r][SP_L2_SIZE + o] = sq >> 6;
                }
        }
    }
    __syncwarp();

    /* L2 + L3: lanes 2r, 2r+1 share row r; each owns 32 of the 64 L2 outputs. multilayer.h:261-447 */
    const int r = lane >> 1, half = lane & 1;
    const size_t pos = base + r;
    const int rb = __shfl_sync(kFull, my_bucket, r);
    const bool valid = pos < n && rb < SP_OUTPUT_BUCKETS;
    const int wb = valid ? rb : 0;
    uint32_t acc[32];
    {
        const int4* bias = reinterpret_cast<const int4*>(net.l2_b + wb * SP_L3_SIZE + half * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int4 v = __ldg(bias + i);
            acc[4 * i] = v.x, acc[4 * i + 1] = v.y, acc[4 * i + 2] = v.z, acc[4 * i + 3] = v.w;
        }
    }
    const int* srow = skip[warp][r];
    if (valid) {
        const int4* w2 = reinterpret_cast<const int4*>(net.l2_w + static_cast<size_t>(wb) * 2 * SP_L2_SIZE * SP_L3_SIZE + half * 32);
#pragma unroll 2
        for (int i = 0; i < 2 * SP_L2_SIZE; ++i) {
            const uint32_t in = static_cast<uint32_t>(srow[i] >> 6);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int4 v = __ldg(w2 + i * (SP_L3_SIZE / 4) + q);
                acc[4 * q] += in * static_cast<uint32_t>(v.x);
                acc[4 * q + 1] += in * static_cast<uint32_t>(v.y);
                acc[4 * q + 2] += in * static_cast<uint32_t>(v.z);
                acc[4 * q + 3] += in * static_cast<uint32_t>(v.w);
            }
        }
    }
    uint32_t l3 = 0;
    if (valid) {
        const int* w3 = net.l3_w + wb * SP_L3_SIZE + half * 32;
#pragma unroll
        for (int p = 0; p < 32; ++p) {
            const int cl = min(max(static_cast<int>(acc[p]), 0), 262144);
            l3 += (static_cast<uint32_t>(cl) + static_cast<uint32_t>(srow[half * 32 + p])) * static_cast<uint32_t>(__ldg(w3 + p));
        }
    }
    l3 += __shfl_xor_sync(kFull, l3, 1);
    if (half == 0 && pos < n) {
        int32_t result = INT32_MIN;
        if (valid) {
            const int32_t sum = static_cast<int32_t>(l3 + static_cast<uint32_t>(__ldg(net.l3_b + wb)));
            result = static_cast<int32_t>(static_cast<int64_t>(sum) * 400 / 16777216); /* truncates toward zero */
        }
        out[pos] = result;
    }
}

int grid_for(size_t n_warp_items, int warps_per_cta, int sm_count, int ctas_per_sm) {
    const size_t want = (n_warp_items + warps_per_cta - 1) / warps_per_cta;
    const size_t cap = static_cast<size_t>(sm_count) * ctas_per_sm;
    return static_cast<int>(want < cap ? (want ? want : 1) : cap);
}

} // namespace

void launch_ft_full(
    const DeviceNet& net, const SpPackedBoard* boards, size_t n, uint8_t* act, uint8_t* bucket, DeviceStatus* status,
    int sm_count, cudaStream_t stream) {
    if (!n) return;
    ft_full_kernel<<<grid_for(n, kWarpsPerCta, sm_count, 2), kThreads, 0, stream>>>(net, boards, n, act, bucket, status);
}

void launch_head(
    const DeviceNet& net, const uint8_t* act, const uint8_t* bucket, size_t n, int32_t* out, DeviceStatus*, int,
    cudaStream_t stream) {
    if (!n) return;
    const size_t tiles = (n + 15) / 16;
    const unsigned grid = static_cast<unsigned>((tiles + kHeadWarps - 1) / kHeadWarps);
    head_kernel<<<grid, kHeadWarps * 32, 0, stream>>>(net, act, bucket, n, out);
}

} // namespace sp::gpu
