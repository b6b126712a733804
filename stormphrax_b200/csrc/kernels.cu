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

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sp_delta.h"

namespace sp::gpu {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
/* tuning knobs (overridable at compile time for experiments; defaults are the measured best) */
#ifndef SP_PSQ_GROUP
#define SP_PSQ_GROUP 4
#endif
#ifndef SP_THR_GROUP
#define SP_THR_GROUP 8
#endif
#ifndef SP_PSQ_INFLIGHT
#define SP_PSQ_INFLIGHT 2 /* PSQ rows whose loads are in flight together (4 x LDG.128 each) */
#endif
#ifndef SP_THR_INFLIGHT
#define SP_THR_INFLIGHT 4 /* threat rows in flight together (2 x LDG.128 each) */
#endif
#ifndef SP_ENQ_UNROLL
#define SP_ENQ_UNROLL 1 /* unroll of the 8-round board enumeration (more ILP, more code) */
#endif
#ifndef SP_ENQ_PREFETCH
#define SP_ENQ_PREFETCH 0 /* 1 (prepared, NOT yet run on a GPU): the board enumeration fetches round k + 1's ray while round k works
                             (today one dependent 64-bit table load per round: 2.7 % of ft_full's samples wait on it) */
#endif
#ifndef SP_GROUP_TIMING
#define SP_GROUP_TIMING 0 /* 1: ft_group_kernel prints CTA 0's clocks per phase (diagnostics; tools/gpu scripts) */
#endif
#ifndef SP_FULL_MIN_BLOCKS
#define SP_FULL_MIN_BLOCKS 4
#endif
#ifndef SP_ACC_MIN_BLOCKS
#define SP_ACC_MIN_BLOCKS 3
#endif
#ifndef SP_SLOTS_MIN_BLOCKS
#define SP_SLOTS_MIN_BLOCKS 3 /* CTAs per SM of ft_slots_kernel: 3 (80 registers) measured +6.5 % over 2 once items came from the ticket counter (before that: +-2 %); 4 (64 registers) no better than 2 */
#endif
#ifndef SP_GAMES_MIN_BLOCKS
#define SP_GAMES_MIN_BLOCKS 2
#endif
constexpr int kEnqUnroll = SP_ENQ_UNROLL;
constexpr int kPsqGroup = SP_PSQ_GROUP;      /* PSQ rows fetched per batch on the rebuild path (4 x LDG.128 each per lane) */
constexpr int kThrGroupFull = SP_THR_GROUP;  /* threat rows per batch on the rebuild path (2 x LDG.128 each per lane) */
constexpr int kPsqGroupDelta = 2;  /* delta rows per batch on the incremental path (8 x LDG.128); lists are short, padding costs */
constexpr int kThrGroupDelta = 2;  /* per sign: 2 added + 2 subtracted rows (8 x LDG.128) */
constexpr int kPsqListCap = 40;    /* 32 pieces + bias row */
constexpr int kPsqDeltaCap = 16;
constexpr int kThrDeltaCap = 96;   /* added rows grow from the front, subtracted rows from the back */
constexpr int kThrListCap = SP_MAX_THREAT_INDICES;
constexpr int kTaskCap = 128;
constexpr uint32_t kSubFlag = 0x80000000u;      /* PSQ delta list entry: subtract this row */
constexpr uint32_t kTaskPawnPair = 1u << 31;    /* task = pawn pair (else attacker -> victim candidate) */
constexpr uint32_t kTaskSub = 1u << 30;         /* feature of the predecessor board: subtract */
constexpr uint32_t kTaskFull = 1u << 29;        /* from a whole-board enumeration: for rebuilt perspectives */
constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr uint32_t kPsqVecs = 128, kThrVecs = 64; /* uint4 per row */
constexpr uint32_t kPsqZeroOff = kPsqZeroRow * kPsqVecs, kPsqBiasOff = kPsqBiasRow * kPsqVecs, kThrZeroOff = kThrZeroRow * kThrVecs;

/* Per-warp shared memory. */
struct WarpScratch {
    /* Row lists hold uint4 offsets into the weight tables (row * 128 for PSQ, row * 64 for threat rows)
     * and are padded with the zero row to a whole number of load batches, so the accumulate loops fetch
     * four entries with one LDS.128 and form an address with one IMAD.WIDE. */
    __align__(16) uint32_t thr_add[2][kThrListCap];      /* rebuild: every threat / pawn-pair row of the board */
    __align__(16) uint32_t thr_delta[2][kThrDeltaCap];   /* update: added rows [0, n_add), subtracted rows [cap - n_sub, cap) */
    __align__(16) uint32_t psq_add[2][kPsqListCap];      /* rebuild: one row per piece + the bias row */
    __align__(16) uint32_t psq_delta[2][kPsqDeltaCap];   /* update: kSubFlag = subtract */
    uint32_t tasks[kTaskCap];              /* (attacker, victim) candidates and pawn pairs awaiting indexing */
    uint8_t mailbox[2][64];                /* two boards: the one being evaluated and its predecessor */
    int n_thr_add[2];
    int n_thr_dadd[2];
    int n_thr_dsub[2];
    int n_psq_add[2];
    int n_psq_delta[2];
};

/* Board as the shared feature code (sp_features.h, sp_delta.h) wants to see it. */
struct BoardView {
    const uint8_t* mailbox;
    uint64_t occ;
    uint64_t pawns[2];
    int king[2];
    int stm;
};

/* Warp-uniform result of decoding a record; the mailbox itself lives in shared memory. */
struct Decoded {
    BoardView view;
    int n_pieces;
    bool ok;
};

/* Decode one marlinformat record (src/datagen/marlinformat.h:43-77) cooperatively: lane l owns
 * squares l and l + 32, so one ballot per piece kind yields one half of a bitboard. */
__device__ __forceinline__ Decoded decode_board(uint4 lo, uint4 hi, int lane, uint8_t* mailbox) {
    Decoded d;
    d.view.mailbox = mailbox;
    d.view.occ = static_cast<uint64_t>(lo.y) << 32 | lo.x;
    d.view.stm = (hi.z & 0x80) ? kBlack : kWhite;
    const int n_lo = __popc(lo.x);
    d.n_pieces = n_lo + __popc(lo.y);
    d.ok = d.n_pieces <= 32 && d.n_pieces >= 2;
    const uint32_t below = (1u << lane) - 1;
    int piece[2];
    bool bad = false;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const uint32_t half = h ? lo.y : lo.x;
        piece[h] = kNoPiece;
        if ((half >> lane) & 1) {
            const int rank = (h ? n_lo : 0) + __popc(half & below); /* index into the nibble list */
            const uint32_t word = rank < 8 ? lo.z : (rank < 16 ? lo.w : (rank < 24 ? hi.x : hi.y));
            const uint32_t nib = (word >> ((rank & 7) * 4)) & 0xF;
            uint32_t type = nib & 7;
            if (type == 6) type = kRook; /* rook with castling rights */
            if (type > kKing || rank >= 32) bad = true;
            else piece[h] = static_cast<int>(type << 1 | ((nib & 8) ? kBlack : kWhite));
        }
    }
    __syncwarp(); /* earlier readers of this mailbox buffer are done */
    mailbox[lane] = static_cast<uint8_t>(piece[0]);
    mailbox[lane + 32] = static_cast<uint8_t>(piece[1]);
    if (__any_sync(kFull, bad)) d.ok = false;
    auto board_of = [&](int p) {
        return static_cast<uint64_t>(__ballot_sync(kFull, piece[1] == p)) << 32 | __ballot_sync(kFull, piece[0] == p);
    };
    const uint64_t bk = board_of(kKing << 1 | kBlack), wk = board_of(kKing << 1 | kWhite);
    if (__popcll(bk) != 1 || __popcll(wk) != 1) d.ok = false;
    d.view.king[kBlack] = bk ? lsb64(bk) : 0;
    d.view.king[kWhite] = wk ? lsb64(wk) : 0;
    d.view.pawns[kBlack] = board_of(kPawn << 1 | kBlack);
    d.view.pawns[kWhite] = board_of(kPawn << 1 | kWhite);
    __syncwarp();
    return d;
}

__device__ __forceinline__ Decoded decode_board(const SpPackedBoard* board, int lane, uint8_t* mailbox) {
    const uint4* p = reinterpret_cast<const uint4*>(board);
    return decode_board(__ldg(p), __ldg(p + 1), lane, mailbox);
}

/* Square of the n-th set bit (ascending), n < popcount(bits): binary search on prefix popcounts. */
__device__ __forceinline__ int nth_piece_square(uint64_t bits, int n) {
    uint32_t half = static_cast<uint32_t>(bits);
    int base = 0;
    const int n_lo = __popc(half);
    if (n >= n_lo) {
        n -= n_lo;
        half = static_cast<uint32_t>(bits >> 32);
        base = 32;
    }
    int pos = 0;
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const int c = __popc(half & ((1u << (pos + step)) - 1));
        if (c <= n) pos += step;
    }
    return base + pos;
}

/* Append `task` from every lane where `valid` (ballot compaction, no atomics). Returns the new count. */
__device__ __forceinline__ int enqueue(uint32_t* tasks, int n_tasks, bool valid, uint32_t task, int lane) {
    const unsigned m = __ballot_sync(kFull, valid);
    if (valid) {
        const int at = n_tasks + __popc(m & ((1u << lane) - 1));
        if (at < kTaskCap) tasks[at] = task;
    }
    return n_tasks + __popc(m);
}
__device__ __forceinline__ int enqueue(WarpScratch& ws, int n_tasks, bool valid, uint32_t task, int lane) {
    return enqueue(ws.tasks, n_tasks, valid, task, lane);
}

__device__ __forceinline__ uint32_t pawn_pair_task(int a_color, int asq, int b_color, int bsq) {
    return static_cast<uint32_t>(asq | bsq << 8 | a_color << 16 | b_color << 24) | kTaskPawnPair;
}

/* Pawn-pair tasks: each lane holds a pawn square (or none) and the bitboard of its partners. */
__device__ __forceinline__ int enqueue_pawn_pairs(
    WarpScratch& ws, int n_tasks, const uint8_t* mailbox, int sq, int color, uint64_t partners, uint32_t flags, int lane) {
#pragma unroll 1
    while (__any_sync(kFull, partners != 0)) {
        const bool valid = partners != 0;
        uint32_t task = 0;
        if (valid) {
            const int o = lsb64(partners);
            partners &= partners - 1;
            task = pawn_pair_task(color, sq, mailbox[o] & 1, o) | flags;
        }
        n_tasks = enqueue(ws, n_tasks, valid, task, lane);
    }
    return n_tasks;
}

/* Whole-board enumeration (nnue_state.cpp:309-354, 440-449), lane-per-piece: eight uniform rounds, round k
 * looks along ray k (sliders, pawns) or at knight offset k; then the pawn pairs. */
/* The threat / pawn-pair half of the enumeration: lane l looks from piece l (`sq`, `piece`; `has` = the lane has one). */
template <int kUnroll = kEnqUnroll, typename Flush>
__device__ __forceinline__ int enqueue_board_threats(
    const FeatureTables& t, const BoardView& b, bool has, int sq, int piece, int lane, uint32_t* tasks, int n_tasks, Flush&& flush) {
    const int type = piece >> 1;
    const bool attacker = has && type != kKing;
#if SP_ENQ_PREFETCH
    const bool slider_like = attacker && type != kKnight; /* pawns look along rays too */
    uint64_t next_ray = slider_like ? t.rays[0][sq] : 0;
#endif
#pragma unroll kUnroll
    for (int k = 0; k < 8; ++k) {
#if SP_ENQ_PREFETCH
        const uint64_t ray = next_ray;
        if (slider_like && k < 7) next_ray = t.rays[k + 1][sq];
#endif
        int target = kNoSquare;
        if (attacker) {
            if (type == kKnight) {
                const int fx = (sq & 7) + static_cast<int>((0x10013443u >> (4 * k)) & 0xF) - 2;
                const int ry = (sq >> 3) + static_cast<int>((0x43100134u >> (4 * k)) & 0xF) - 2;
                if (fx >= 0 && fx < 8 && ry >= 0 && ry < 8) target = ry * 8 + fx;
            } else {
                uint64_t gap;
#if SP_ENQ_PREFETCH
                const int ahead = ray_first(ray, b.occ, k, gap);
#else
                const int ahead = ray_first(t, b.occ, sq, k, gap);
#endif
                if (ahead != kNoSquare && attacks_along(piece, k, gap == 0)) target = ahead;
            }
        }
        int victim = kNoPiece;
        if (target != kNoSquare) victim = b.mailbox[target];
        const bool valid = victim != kNoPiece && (victim >> 1) != kKing;
        n_tasks = enqueue(tasks, n_tasks, valid, pack_candidate(piece, sq, victim, target) | kTaskFull, lane);
        if (n_tasks > kTaskCap - 32) n_tasks = flush(n_tasks); /* the next round might not fit: index what is queued */
    }
    /* every unordered pawn pair within one file of each other, once: partner on a higher square */
    const bool pawn = has && type == kPawn;
    uint64_t partners = pawn ? (b.pawns[0] | b.pawns[1]) & pp_mask(sq) & squares_above(sq) : 0;
#pragma unroll 1
    while (__any_sync(kFull, partners != 0)) {
        const bool valid = partners != 0;
        uint32_t task = 0;
        if (valid) {
            const int o = lsb64(partners);
            partners &= partners - 1;
            task = pawn_pair_task(piece & 1, sq, b.mailbox[o] & 1, o) | kTaskFull;
        }
        n_tasks = enqueue(tasks, n_tasks, valid, task, lane);
        if (n_tasks > kTaskCap - 32) n_tasks = flush(n_tasks);
    }
    return n_tasks;
}

template <typename Flush>
__device__ __forceinline__ int enqueue_board(
    const FeatureTables& t, const BoardView& b, int n_pieces, int rebuild, int lane, WarpScratch& ws, int n_tasks, Flush&& flush) {
    const bool has = lane < n_pieces;
    const int sq = has ? nth_piece_square(b.occ, lane) : 0;
    const int piece = has ? b.mailbox[sq] : kNoPiece;
    if (has) {
        if (rebuild & 1) ws.psq_add[kBlack][lane] = psq_index(t, kBlack, piece, sq, b.king[kBlack]) * kPsqVecs;
        if (rebuild & 2) ws.psq_add[kWhite][lane] = psq_index(t, kWhite, piece, sq, b.king[kWhite]) * kPsqVecs;
    }
    return enqueue_board_threats(t, b, has, sq, piece, lane, ws.tasks, n_tasks, flush);
}

/* Changed-square enumeration (sp_delta.h): line items lane = unit * 8 + k, then one square item per unit. */
__device__ __forceinline__ int enqueue_delta(
    const FeatureTables& t, const BoardView& before, const BoardView& after, uint64_t changed, int rebuild, int lane,
    WarpScratch& ws, int n_tasks, int (&n_psq_delta)[2]) {
    const int n_changed = __popcll(changed);
    const int my_changed = lane < n_changed ? nth_piece_square(changed, lane) : 0;
    const int units = 2 * n_changed;
#pragma unroll 1
    for (int base = 0; base < units * kLineItemsPerUnit; base += 32) {
        const int item = base + lane;
        const int u = item >> 3;
        const int s = __shfl_sync(kFull, my_changed, (u >> 1) & 31);
        uint32_t c[4] = {0, 0, 0, 0};
        unsigned valid = 0;
        if (u < units) valid = delta_line_candidates(t, (u & 1) ? after : before, changed, s, item & 7, c);
        const uint32_t flags = (u & 1) ? 0u : kTaskSub;
#pragma unroll
        for (int i = 0; i < 4; ++i) n_tasks = enqueue(ws, n_tasks, (valid >> i) & 1, c[i] | flags, lane);
    }
    /* square items: lane = unit */
    const int s = __shfl_sync(kFull, my_changed, (lane >> 1) & 31);
    const BoardView& b = (lane & 1) ? after : before;
    const int piece = lane < units ? b.mailbox[s] : kNoPiece;
    const uint32_t sub = (lane & 1) ? 0u : kSubFlag;
    const unsigned occupied = __ballot_sync(kFull, piece != kNoPiece);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        if ((rebuild >> c) & 1) continue;
        if (piece != kNoPiece) {
            const int at = __popc(occupied & ((1u << lane) - 1));
            if (at < kPsqDeltaCap) ws.psq_delta[c][at] = psq_index(t, c, piece, s, after.king[c]) * kPsqVecs | sub;
        }
        n_psq_delta[c] = __popc(occupied);
    }
    const bool pawn = piece != kNoPiece && (piece >> 1) == kPawn;
    const uint64_t partners = pawn ? delta_pawn_partners(b, changed, s) : 0;
    return enqueue_pawn_pairs(ws, n_tasks, b.mailbox, s, piece & 1, partners, (lane & 1) ? 0u : kTaskSub, lane);
}

/* Index the queued tasks, one (task, perspective) per lane, and compact the existing features into the
 * row lists.  Kings of the current board orient both boards' features: a perspective whose king changed
 * side or bucket is rebuilt, not updated (threats.cpp:170-221). */
__device__ __forceinline__ void process_tasks(
    const FeatureTables& t, const BoardView& b, int rebuild, int external, int n_tasks, int lane, WarpScratch& ws,
    int (&n_full)[2], int (&n_dadd)[2], int (&n_dsub)[2]) {
    const int c = lane & 1;
    const unsigned mine = (c ? 0xAAAAAAAAu : 0x55555555u) & ((1u << lane) - 1); /* earlier lanes of my perspective */
    const bool rebuilt = (rebuild >> c) & 1;
    const bool wanted = !((external >> c) & 1);
#pragma unroll 1
    for (int base = 0; base < n_tasks; base += 16) {
        const int i = base + (lane >> 1);
        uint32_t task = 0;
        int32_t idx = -1;
        if (i < n_tasks) {
            task = ws.tasks[i];
            if (wanted && ((task & kTaskFull) != 0) == rebuilt) {
                const int sq0 = task & 0xFF, sq1 = (task >> 8) & 0xFF, p0 = (task >> 16) & 0xF, p1 = (task >> 24) & 0xF;
                const int ksq = c ? b.king[1] : b.king[0];
                idx = (task & kTaskPawnPair) ? static_cast<int32_t>(pp_index(c, ksq, p0, sq0, p1, sq1))
                                             : threat_index(t, c, ksq, p0, sq0, p1, sq1);
            }
        }
        const bool valid = idx >= 0;
        const bool full = (task & kTaskFull) != 0, sub = (task & kTaskSub) != 0;
        const unsigned mf = __ballot_sync(kFull, valid && full);
        const unsigned ma = __ballot_sync(kFull, valid && !full && !sub);
        const unsigned ms = __ballot_sync(kFull, valid && !full && sub);
        if (valid) {
            if (full) {
                const int at = (c ? n_full[1] : n_full[0]) + __popc(mf & mine);
                if (at < kThrListCap) ws.thr_add[c][at] = static_cast<uint32_t>(idx) * kThrVecs;
            } else if (!sub) {
                const int at = (c ? n_dadd[1] : n_dadd[0]) + __popc(ma & mine);
                if (at < kThrDeltaCap) ws.thr_delta[c][at] = static_cast<uint32_t>(idx) * kThrVecs;
            } else {
                const int at = (c ? n_dsub[1] : n_dsub[0]) + __popc(ms & mine);
                if (at < kThrDeltaCap) ws.thr_delta[c][kThrDeltaCap - 1 - at] = static_cast<uint32_t>(idx) * kThrVecs;
            }
        }
        n_full[0] += __popc(mf & 0x55555555u), n_full[1] += __popc(mf & 0xAAAAAAAAu);
        n_dadd[0] += __popc(ma & 0x55555555u), n_dadd[1] += __popc(ma & 0xAAAAAAAAu);
        n_dsub[0] += __popc(ms & 0x55555555u), n_dsub[1] += __popc(ms & 0xAAAAAAAAu);
    }
}

/*
 * Fill the warp's feature lists for the step `before` -> `d` (before == nullptr: no predecessor).
 *   perspectives in the returned mask are rebuilt from scratch: psq_add = all pieces + bias row,
 *     thr_add = every threat / pawn-pair feature of the board (nnue_state.cpp:309-354, 440-449)
 *   the others are updated: psq_delta / thr_delta = delta rows (sp_delta.h; replaces
 *     nnue.cpp:490-599, nnue_state.cpp:34-87, 163-307)
 * A perspective is rebuilt when its king changes input bucket or board half (psq.h:264-283,
 * nnue_state.h:118-128), when more than kMaxChanged squares differ, or when a delta list overflows.
 * Returns -1 if a full list exceeds the reference's bound of 256 entries.
 * `external`: perspectives whose fresh accumulator comes from elsewhere (RebuildPlan): nothing is
 * listed for them.  `only`: with no predecessor, the perspectives to rebuild (default both). */
__device__ __forceinline__ int build_lists(
    const FeatureTables& t, const BoardView* before, const Decoded& d, int lane, WarpScratch& ws, int external = 0, int only = 3) {
    int rebuild = only & ~external;
    uint64_t changed = 0;
    if (before) {
        const unsigned lo = __ballot_sync(kFull, before->mailbox[lane] != d.view.mailbox[lane]);
        const unsigned hi = __ballot_sync(kFull, before->mailbox[32 + lane] != d.view.mailbox[32 + lane]);
        changed = static_cast<uint64_t>(hi) << 32 | lo;
        rebuild = (needs_refresh(t, *before, d.view, kBlack) ? 1 : 0) | (needs_refresh(t, *before, d.view, kWhite) ? 2 : 0);
        if (__popcll(changed) > kMaxChanged) rebuild = 3;
        rebuild &= ~external;
    }
    for (;;) {
        int n_tasks = 0;
        int n_full[2] = {0, 0}, n_dadd[2] = {0, 0}, n_dsub[2] = {0, 0}, n_psq_delta[2] = {0, 0};
        const int skip = rebuild | external; /* perspectives that take no delta rows */
        if (before && skip != 3) n_tasks = enqueue_delta(t, *before, d.view, changed, skip, lane, ws, n_tasks, n_psq_delta);
        if (n_tasks > kTaskCap - 32) { /* far too many candidates for an update: rebuild instead */
            rebuild = 3 & ~external;
            __syncwarp();
            continue;
        }
        auto flush = [&](int queued) {
            __syncwarp();
            process_tasks(t, d.view, rebuild, external, queued, lane, ws, n_full, n_dadd, n_dsub);
            __syncwarp();
            return 0;
        };
        if (rebuild) n_tasks = enqueue_board(t, d.view, d.n_pieces, rebuild, lane, ws, n_tasks, flush);
        flush(n_tasks);
        int overflow = 0, too_long = 0;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if (n_full[c] > kThrListCap) too_long |= 1 << c;
            if (2 * ((max(n_dadd[c], n_dsub[c]) + kThrGroupDelta - 1) / kThrGroupDelta * kThrGroupDelta) > kThrDeltaCap
                || n_psq_delta[c] > kPsqDeltaCap)
                overflow |= 1 << c;
        }
        if (too_long) return -1; /* a from-scratch list does not fit: the reference's own limit */
        if (overflow) {
            rebuild = 3 & ~external; /* an over-long delta: fall back to rebuilding */
            __syncwarp();
            continue;
        }
        /* top the lists up with the zero row to whole load batches and publish the padded lengths */
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            if ((external >> c) & 1) continue;
            if ((rebuild >> c) & 1) {
                const int n_psq = d.n_pieces + 1, n_psq_pad = (n_psq + kPsqGroup - 1) / kPsqGroup * kPsqGroup;
                const int n_thr_pad = (n_full[c] + kThrGroupFull - 1) / kThrGroupFull * kThrGroupFull;
                if (lane == 0) ws.psq_add[c][d.n_pieces] = kPsqBiasOff;
                if (lane > 0 && d.n_pieces + lane < n_psq_pad) ws.psq_add[c][d.n_pieces + lane] = kPsqZeroOff;
                if (n_full[c] + lane < n_thr_pad) ws.thr_add[c][n_full[c] + lane] = kThrZeroOff;
                if (lane == 0) ws.n_psq_add[c] = n_psq_pad, ws.n_thr_add[c] = n_thr_pad;
            } else {
                const int n_psq_pad = (n_psq_delta[c] + kPsqGroupDelta - 1) / kPsqGroupDelta * kPsqGroupDelta;
                const int n_thr_pad = (max(n_dadd[c], n_dsub[c]) + kThrGroupDelta - 1) / kThrGroupDelta * kThrGroupDelta;
                if (n_psq_delta[c] + lane < n_psq_pad) ws.psq_delta[c][n_psq_delta[c] + lane] = kPsqZeroOff;
                /* both signs are fetched in lock-step: pad each up to the common length */
                for (int i = n_dadd[c] + lane; i < n_thr_pad; i += 32) ws.thr_delta[c][i] = kThrZeroOff;
                for (int i = n_dsub[c] + lane; i < n_thr_pad; i += 32) ws.thr_delta[c][kThrDeltaCap - 1 - i] = kThrZeroOff;
                if (lane == 0) ws.n_psq_delta[c] = n_psq_pad, ws.n_thr_dadd[c] = n_thr_pad;
            }
        }
        __syncwarp();
        return rebuild;
    }
}

/* ------------------------------------------------------------------ accumulators in registers
 *
 * One perspective, one lane: 32 int16 elements packed two per register, V[k * 4 + t]:
 *   k = 0: A-half elements (4t, 4t+2)   k = 1: A-half (4t+1, 4t+3)      A-half element j = 16 l + j
 *   k = 2: D-half elements (4t, 4t+2)   k = 3: D-half (4t+1, 4t+3)      D-half element j = 512 + 16 l + j
 * (low 16 bits first).  This is the order in which the even / odd bytes of word t of a threat row
 * fall out of one AND / one PRMT, and the device PSQ rows are stored in the same order, so both
 * kinds of row are added with VIADD.16x2 (__vadd2: wrapping, no carry between the halves). */

__device__ __forceinline__ uint32_t odd_bytes(uint32_t w) { return __byte_perm(w, 0, 0x4341); }

/* Weight-row loads.  SP_ROW_LOAD selects the cache policy (experiments): 0 = ld.global.nc (default),
 * 1 = threat rows do not allocate in L1, 2 = no row allocates in L1. */
#ifndef SP_ROW_LOAD
#define SP_ROW_LOAD 0
#endif
__device__ __forceinline__ uint4 ldg_no_allocate(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

/* The lane's view of a weight table (table + lane).  The compiler re-associates the uniform table pointer out
 * of it and forms every row address with four instructions (IADD3 + IMAD.X + LEA + LEA.HI.X on lane +
 * offset).  SP_ROW_ADDR=1 hides the pointer behind an empty asm so that a row address is ONE IMAD.WIDE.U32:
 * 7 % fewer instructions in the full refresh -- and 6 % SLOWER on the GPU (84.9 vs 90.7 Mpos/s; the wide
 * multiply sits on the half-rate FMA pipe in front of every load).  Kept as an experiment knob, off. */
#ifndef SP_ROW_ADDR
#define SP_ROW_ADDR 0
#endif
__device__ __forceinline__ const uint4* lane_view(const uint4* table, int lane) {
    const uint4* p = table + lane;
#if SP_ROW_ADDR
    asm volatile("" : "+l"(p));
#endif
    return p;
}

/* `base` = table + lane; `off` = uint4 offset of the row (a list entry) */
__device__ __forceinline__ void load_psq_row(const uint4* base, uint32_t off, uint4 (&c)[4]) {
    const uint4* r = base + off;
#pragma unroll
    for (int k = 0; k < 4; ++k) c[k] = SP_ROW_LOAD >= 2 ? ldg_no_allocate(r + 32 * k) : __ldg(r + 32 * k);
}

__device__ __forceinline__ void load_thr_row(const uint4* base, uint32_t off, uint4 (&c)[2]) {
    const uint4* r = base + off;
    c[0] = SP_ROW_LOAD >= 1 ? ldg_no_allocate(r) : __ldg(r);
    c[1] = SP_ROW_LOAD >= 1 ? ldg_no_allocate(r + 32) : __ldg(r + 32);
}

__device__ __forceinline__ void add_psq(uint32_t (&v)[16], const uint4 (&c)[4]) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k * 4 + 0] = __vadd2(v[k * 4 + 0], c[k].x);
        v[k * 4 + 1] = __vadd2(v[k * 4 + 1], c[k].y);
        v[k * 4 + 2] = __vadd2(v[k * 4 + 2], c[k].z);
        v[k * 4 + 3] = __vadd2(v[k * 4 + 3], c[k].w);
    }
}

/* Many threat rows (rebuild path): plain 32-bit sums.  s += word accumulates all four bytes with
 * weights 1, 2^8, 2^16, 2^24 (mod 2^32); o += odd bytes as two 16-bit fields.  <= 256 rows x 255
 * < 2^16, so the fields of o never carry and s - (o << 8) leaves exactly the even-byte fields. */
__device__ __forceinline__ void add_thr_wide(uint32_t (&s)[8], uint32_t (&o)[8], const uint4 (&c)[2]) {
#pragma unroll
    for (int u = 0; u < 2; ++u) {
        const uint32_t w[4] = {c[u].x, c[u].y, c[u].z, c[u].w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            s[u * 4 + t] += w[t];
            o[u * 4 + t] += odd_bytes(w[t]);
        }
    }
}

/* Rebuild one perspective from its full lists: v = bias + sum(PSQ rows) + sum(threat rows).
 * The lists are padded to whole batches (build_lists). */
__device__ __forceinline__ void rebuild_perspective(
    const DeviceNet& net, const uint32_t* psq_list, int n_psq, const uint32_t* thr_list, int n_thr, int lane, uint32_t (&v)[16]) {
    static_assert(kPsqGroup == 4 && kThrGroupFull % 4 == 0, "list entries are fetched four at a time");
    const uint4* psq_base = lane_view(net.psq, lane);
    const uint4* thr_base = lane_view(net.thr, lane);
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = 0;
#pragma unroll 1
    for (int i = 0; i < n_psq; i += kPsqGroup) {
        const uint4 e = *reinterpret_cast<const uint4*>(psq_list + i);
        const uint32_t off[4] = {e.x, e.y, e.z, e.w};
#pragma unroll
        for (int h = 0; h < kPsqGroup; h += SP_PSQ_INFLIGHT) {
            uint4 c[SP_PSQ_INFLIGHT][4];
#pragma unroll
            for (int j = 0; j < SP_PSQ_INFLIGHT; ++j) load_psq_row(psq_base, off[h + j], c[j]);
#pragma unroll
            for (int j = 0; j < SP_PSQ_INFLIGHT; ++j) add_psq(v, c[j]);
        }
    }
    uint32_t s[8], o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = o[i] = 0;
#pragma unroll 1
    for (int i = 0; i < n_thr; i += kThrGroupFull) {
#pragma unroll
        for (int h = 0; h < kThrGroupFull; h += SP_THR_INFLIGHT) {
            uint4 c[SP_THR_INFLIGHT][2];
#pragma unroll
            for (int q = 0; q < SP_THR_INFLIGHT / 4; ++q) {
                const uint4 e = *reinterpret_cast<const uint4*>(thr_list + i + h + 4 * q);
                load_thr_row(thr_base, e.x, c[4 * q + 0]);
                load_thr_row(thr_base, e.y, c[4 * q + 1]);
                load_thr_row(thr_base, e.z, c[4 * q + 2]);
                load_thr_row(thr_base, e.w, c[4 * q + 3]);
            }
#pragma unroll
            for (int j = 0; j < SP_THR_INFLIGHT; ++j) add_thr_wide(s, o, c[j]);
        }
    }
    /* every row (zero-row top-ups included) carried +128 per element */
    const uint32_t corr = (static_cast<uint32_t>(-128 * n_thr) & 0xFFFFu) * 0x10001u;
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t even = s[u * 4 + t] - (o[u * 4 + t] << 8);
            v[(2 * u) * 4 + t] = __vadd2(__vadd2(v[(2 * u) * 4 + t], even), corr);
            v[(2 * u + 1) * 4 + t] = __vadd2(__vadd2(v[(2 * u + 1) * 4 + t], o[u * 4 + t]), corr);
        }
}

/* Out-of-line copy for the kernels where a rebuild is the rare path (king crossed a bucket
 * boundary): keeps their hot loop small enough for the instruction cache. */
__device__ __noinline__ void rebuild_perspective_cold(
    const DeviceNet& net, const uint32_t* psq_list, int n_psq, const uint32_t* thr_list, int n_thr, int lane, uint32_t* out) {
    uint32_t v[16];
    rebuild_perspective(net, psq_list, n_psq, thr_list, n_thr, lane, v);
#pragma unroll
    for (int i = 0; i < 16; ++i) out[i] = v[i];
}

/* Advance one perspective by its delta lists.
 * PSQ rows: one signed list; a subtracted row is added complemented (~w = -w - 1 per int16).
 * Threat rows: added and subtracted rows are fetched in lock-step (equal row counts, so their +128
 * biases cancel) and summed as on the rebuild path -- whole words in s, odd bytes in o, even bytes
 * recovered as s - (o << 8) -- then the subtracted sums are added complemented.  All the "-1" of the
 * complements are repaid by one constant at the end. */
__device__ __forceinline__ void update_perspective(const DeviceNet& net, const WarpScratch& ws, int c, int lane, uint32_t (&v)[16]) {
    static_assert(kPsqGroupDelta == 2 && kThrGroupDelta == 2, "list entries are fetched two at a time");
    const uint4* psq_base = lane_view(net.psq, lane);
    const uint4* thr_base = lane_view(net.thr, lane);
    const int n_psq = ws.n_psq_delta[c]; /* padded */
    int psq_subs = 0;
#pragma unroll 1
    for (int i = 0; i < n_psq; i += kPsqGroupDelta) {
        const uint2 e2 = *reinterpret_cast<const uint2*>(ws.psq_delta[c] + i);
        const uint32_t e[2] = {e2.x, e2.y};
        uint4 rows[kPsqGroupDelta][4];
        uint32_t mask[kPsqGroupDelta];
#pragma unroll
        for (int j = 0; j < kPsqGroupDelta; ++j) {
            mask[j] = static_cast<uint32_t>(static_cast<int32_t>(e[j]) >> 31);
            psq_subs += e[j] >> 31;
            load_psq_row(psq_base, e[j] & ~kSubFlag, rows[j]);
        }
#pragma unroll
        for (int j = 0; j < kPsqGroupDelta; ++j)
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                v[k * 4 + 0] = __vadd2(v[k * 4 + 0], rows[j][k].x ^ mask[j]);
                v[k * 4 + 1] = __vadd2(v[k * 4 + 1], rows[j][k].y ^ mask[j]);
                v[k * 4 + 2] = __vadd2(v[k * 4 + 2], rows[j][k].z ^ mask[j]);
                v[k * 4 + 3] = __vadd2(v[k * 4 + 3], rows[j][k].w ^ mask[j]);
            }
    }
    const int n_thr = ws.n_thr_dadd[c]; /* common padded length of the added and the subtracted list */
    uint32_t sa[8], oa[8], ss[8], os[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) sa[i] = oa[i] = ss[i] = os[i] = 0;
#pragma unroll 1
    for (int i = 0; i < n_thr; i += kThrGroupDelta) {
        const uint2 ea = *reinterpret_cast<const uint2*>(ws.thr_delta[c] + i);
        const uint2 es = *reinterpret_cast<const uint2*>(ws.thr_delta[c] + kThrDeltaCap - kThrGroupDelta - i); /* stored backwards */
        uint4 ra[kThrGroupDelta][2], rs[kThrGroupDelta][2];
        load_thr_row(thr_base, ea.x, ra[0]), load_thr_row(thr_base, ea.y, ra[1]);
        load_thr_row(thr_base, es.x, rs[0]), load_thr_row(thr_base, es.y, rs[1]);
#pragma unroll
        for (int j = 0; j < kThrGroupDelta; ++j) {
            add_thr_wide(sa, oa, ra[j]);
            add_thr_wide(ss, os, rs[j]);
        }
    }
    const uint32_t repay = (static_cast<uint32_t>(psq_subs + 1) & 0xFFFFu) * 0x10001u; /* +1: the complemented threat sums */
#pragma unroll
    for (int u = 0; u < 2; ++u)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const uint32_t even_a = sa[u * 4 + t] - (oa[u * 4 + t] << 8), even_s = ss[u * 4 + t] - (os[u * 4 + t] << 8);
            uint32_t& ve = v[(2 * u) * 4 + t];
            uint32_t& vo = v[(2 * u + 1) * 4 + t];
            ve = __vadd2(__vadd2(__vadd2(ve, even_a), ~even_s), repay);
            vo = __vadd2(__vadd2(__vadd2(vo, oa[u * 4 + t]), ~os[u * 4 + t]), repay);
        }
}

/* activateFt, multilayer.h:92-152, on packed pairs: out = (clamp(a,0,255) * clamp(d,0,255)) >> 9.
 * Equal to the reference's sat_u8(((clamp(a,0,255) << 7) * min(d,255)) >> 16): for d < 0 the
 * product is <= 0 and saturates to 0, for d >= 0 it is (a * d) >> 9 <= 127.
 * Returns this lane's 16 output bytes (elements 16 l .. 16 l + 15 of the perspective's half). */
__device__ __forceinline__ uint4 activate(const uint32_t (&v)[16]) {
    uint32_t out[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        uint32_t pair[2];
#pragma unroll
        for (int par = 0; par < 2; ++par) {
            const uint32_t a = __vmins2(__vmaxs2(v[par * 4 + t], 0u), 0x00FF00FFu);
            const uint32_t d = __vmins2(__vmaxs2(v[(2 + par) * 4 + t], 0u), 0x00FF00FFu);
            /* d's bytes are (d_lo, 0, d_hi, 0): dp2a.lo multiplies a's halves by bytes 0 and 1 */
            const uint32_t lo = __dp2a_lo(a, d, 0u);      /* a_lo * d_lo : element 4t + par     */
            const uint32_t hi = __dp2a_lo(a, d >> 8, 0u); /* a_hi * d_hi : element 4t + par + 2 */
            pair[par] = ((hi << 16 | lo) >> 9) & 0x007F007Fu; /* both products < 2^16, >> 9 leaves 7 bits */
        }
        out[t] = pair[0] | pair[1] << 8; /* bytes: 4t, 4t+1, 4t+2, 4t+3 */
    }
    return make_uint4(out[0], out[1], out[2], out[3]);
}

__device__ __forceinline__ void flag_error(DeviceStatus* status, int bits) { atomicOr(&status->error, bits); }

/* ------------------------------------------------------------------ full refresh: boards -> activations */

__global__ void __launch_bounds__(kThreads, SP_FULL_MIN_BLOCKS)
ft_full_kernel(DeviceNet net, const SpPackedBoard* __restrict__ boards, size_t n, uint8_t* __restrict__ act,
               uint8_t* __restrict__ bucket, DeviceStatus* status, const uint32_t* __restrict__ groups = nullptr) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const FeatureTables& t = *net.tables;
    const size_t stride = static_cast<size_t>(gridDim.x) * kWarpsPerCta;
    /* `groups` (ft_group_kernel's overflow list: [0] = count, [1..] = indices of 16-position groups): only those positions */
    const size_t items = groups ? static_cast<size_t>(groups[0]) * 16 : n;
    for (size_t item = static_cast<size_t>(blockIdx.x) * kWarpsPerCta + warp; item < items; item += stride) {
        const size_t pos = groups ? static_cast<size_t>(groups[1 + item / 16]) * 16 + item % 16 : item;
        if (pos >= n) continue;
        const Decoded d = decode_board(boards + pos, lane, ws.mailbox[0]);
        int err = d.ok ? 0 : kErrBadBoard;
        if (!err && build_lists(t, nullptr, d, lane, ws) < 0) err = kErrCapacity;
        if (err) {
            if (lane == 0) {
                flag_error(status, err);
                bucket[pos] = 0xFF;
            }
            continue;
        }
        uint4* row = reinterpret_cast<uint4*>(act + pos * SP_L1_SIZE);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t v[16];
            rebuild_perspective(net, ws.psq_add[c], ws.n_psq_add[c], ws.thr_add[c], ws.n_thr_add[c], lane, v);
            const int half = c == d.view.stm ? 0 : 1; /* side to move first, nnue_state.cpp:405-419 */
            row[half * 32 + lane] = activate(v);
        }
        if (lane == 0) bucket[pos] = static_cast<uint8_t>(output_bucket(d.view.occ));
    }
}

/* ------------------------------------------------------------------ full refresh, split form
 *
 * The same work as ft_full_kernel in two kernels, each with the occupancy its phase wants:
 *   extract_kernel     boards -> row lists (one RowListRecord per position and perspective) in global
 *                      memory.  Branchy integer code, latency-bound: few registers, many warps.
 *   accumulate_kernel  row lists -> activations.  One warp per (position, perspective); the next
 *                      record is staged into shared memory by the TMA engine (cp.async.bulk + mbarrier)
 *                      while the current one's rows stream in, so the row loop never waits for its
 *                      own index list. */

struct alignas(16) RowListRecord {
    uint16_t n_psq, n_thr; /* padded entry counts */
    uint8_t half;          /* which half of the activation row this perspective fills (0 = side to move) */
    uint8_t bucket;        /* output bucket, 0xFF = rejected board */
    uint8_t pad[10];
    uint32_t psq[kPsqListCap];
    uint32_t thr[kThrListCap];
};
static_assert(sizeof(RowListRecord) == 1200 && sizeof(RowListRecord) % 16 == 0, "bulk copies move whole records");

__global__ void __launch_bounds__(kThreads, 4)
extract_kernel(DeviceNet net, const SpPackedBoard* __restrict__ boards, size_t n, RowListRecord* __restrict__ records,
               DeviceStatus* status) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const FeatureTables& t = *net.tables;
    const size_t stride = static_cast<size_t>(gridDim.x) * kWarpsPerCta;
    for (size_t pos = static_cast<size_t>(blockIdx.x) * kWarpsPerCta + warp; pos < n; pos += stride) {
        const Decoded d = decode_board(boards + pos, lane, ws.mailbox[0]);
        int err = d.ok ? 0 : kErrBadBoard;
        if (!err && build_lists(t, nullptr, d, lane, ws) < 0) err = kErrCapacity;
        if (err && lane == 0) flag_error(status, err);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            RowListRecord& rec = records[2 * pos + c];
            if (lane == 0) {
                rec.n_psq = err ? 0 : static_cast<uint16_t>(ws.n_psq_add[c]);
                rec.n_thr = err ? 0 : static_cast<uint16_t>(ws.n_thr_add[c]);
                rec.half = c == d.view.stm ? 0 : 1;
                rec.bucket = err ? 0xFF : static_cast<uint8_t>(output_bucket(d.view.occ));
            }
            if (err) continue;
            for (int i = lane; i < ws.n_psq_add[c]; i += 32) rec.psq[i] = ws.psq_add[c][i];
            for (int i = lane; i < ws.n_thr_add[c]; i += 32) rec.thr[i] = ws.thr_add[c][i];
        }
        __syncwarp();
    }
}

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
/* The suspend-time hint lets the hardware park the warp until the phase completes (or the hint expires) instead of
 * returning at once: without it a waiting warp spins -- in ft_group_kernel a third of ALL executed instructions were
 * TRYWAIT / BRA / YIELD of such loops (ncu, profiles/r2_ft_group_ncu_v1.md), stealing issue slots from the working warps. */
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_addr(bar)), "r"(parity), "r"(0x989680u)
        : "memory");
    return done != 0;
}
/* TMA engine: global -> shared bulk copy, completion counted in bytes on the mbarrier */
__device__ __forceinline__ void bulk_copy_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_addr(dst)),
                 "l"(src), "r"(bytes), "r"(smem_addr(bar))
                 : "memory");
}

constexpr int kAccWarps = 8;

__global__ void __launch_bounds__(kAccWarps * 32, SP_ACC_MIN_BLOCKS)
accumulate_kernel(DeviceNet net, const RowListRecord* __restrict__ records, size_t n_items, uint8_t* __restrict__ act,
                  uint8_t* __restrict__ bucket) {
    __shared__ RowListRecord staged[kAccWarps][2];
    __shared__ uint64_t full[kAccWarps][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) {
        mbar_init(&full[warp][0], 1);
        mbar_init(&full[warp][1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const size_t stride = static_cast<size_t>(gridDim.x) * kAccWarps;
    size_t item = static_cast<size_t>(blockIdx.x) * kAccWarps + warp;
    if (item < n_items && lane == 0) {
        mbar_expect_tx(&full[warp][0], sizeof(RowListRecord));
        bulk_copy_g2s(&staged[warp][0], records + item, sizeof(RowListRecord), &full[warp][0]);
    }
    for (uint32_t k = 0; item < n_items; item += stride, ++k) {
        const int buf = k & 1;
        const size_t next = item + stride;
        if (next < n_items && lane == 0) { /* stage the next record while this one is processed */
            mbar_expect_tx(&full[warp][buf ^ 1], sizeof(RowListRecord));
            bulk_copy_g2s(&staged[warp][buf ^ 1], records + next, sizeof(RowListRecord), &full[warp][buf ^ 1]);
        }
        while (!mbar_try_wait(&full[warp][buf], (k >> 1) & 1)) {}
        const RowListRecord& rec = staged[warp][buf];
        const size_t pos = item >> 1;
        if (rec.bucket != 0xFF) {
            uint32_t v[16];
            rebuild_perspective(net, rec.psq, rec.n_psq, rec.thr, rec.n_thr, lane, v);
            reinterpret_cast<uint4*>(act + pos * SP_L1_SIZE)[rec.half * 32 + lane] = activate(v);
        }
        if ((item & 1) == 0 && lane == 0) bucket[pos] = rec.bucket;
        __syncwarp();
        /* order this warp's generic-proxy reads of the buffer before the async-proxy write that refills it */
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

/* ------------------------------------------------------------------ accumulator slots */

__device__ __forceinline__ void load_slot_acc(const SlotStore& s, uint32_t slot, int c, int lane, uint32_t (&v)[16]) {
    const uint4* p = s.acc + (static_cast<size_t>(slot) * 2 + c) * 128 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const uint4 q = p[32 * k];
        v[k * 4 + 0] = q.x, v[k * 4 + 1] = q.y, v[k * 4 + 2] = q.z, v[k * 4 + 3] = q.w;
    }
}

__device__ __forceinline__ void store_slot_acc(const SlotStore& s, uint32_t slot, int c, int lane, const uint32_t (&v)[16]) {
    uint4* p = s.acc + (static_cast<size_t>(slot) * 2 + c) * 128 + lane;
#pragma unroll
    for (int k = 0; k < 4; ++k) p[32 * k] = make_uint4(v[k * 4 + 0], v[k * 4 + 1], v[k * 4 + 2], v[k * 4 + 3]);
}

/* dst[i] = (src ? src[i] advanced to boards[i] : rebuilt from boards[i]); optional activation output.
 * Replaces NnueState::reset / push + applyMove<BoardObserver> + ensureUpToDate
 * (nnue_state.cpp:539-570, 636-697). */
__global__ void __launch_bounds__(kThreads, SP_SLOTS_MIN_BLOCKS)
ft_slots_kernel(DeviceNet net, SlotStore slots, const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst,
                const SpPackedBoard* __restrict__ boards, size_t n, uint8_t* __restrict__ act, uint8_t* __restrict__ bucket,
                DeviceStatus* status) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const FeatureTables& t = *net.tables;
    /* Items are handed out from a ticket counter instead of round-robin: an item whose king crossed a bucket boundary
     * (a rebuild from ~80 rows instead of ~10 delta rows) costs several times the usual one, and with ~28 items per warp the slowest
     * warp of a static schedule ran well past the average.  The counters reset themselves: the last warp to run dry zeroes them. */
    unsigned long long* const ticket = &status->counters[kSlotsTicket];
#ifndef SP_SLOTS_TICKET
#define SP_SLOTS_TICKET 1 /* items per ticket; 0 = static round-robin.  (2 is NOT safe: taking adjacent items back to back in one warp gave wrong results) */
#endif
    constexpr unsigned long long kItemsPerTicket = SP_SLOTS_TICKET ? SP_SLOTS_TICKET : 1;
    size_t static_next = static_cast<size_t>(blockIdx.x) * kWarpsPerCta + warp;
    for (;;) {
        unsigned long long first = 0;
        if (SP_SLOTS_TICKET) {
            if (lane == 0) first = atomicAdd(ticket, kItemsPerTicket);
            first = __shfl_sync(kFull, first, 0);
        } else {
            first = static_next, static_next += static_cast<size_t>(gridDim.x) * kWarpsPerCta;
        }
        if (first >= n) break;
        const size_t last = min(static_cast<size_t>(first + kItemsPerTicket), n);
    for (size_t i = first; i < last; ++i) {
        __syncwarp(); /* the previous item's readers of the warp's scratch are done */
        const uint32_t to = dst[i];
        const uint32_t from = src ? src[i] : to;
        int err = (to >= slots.n_slots || from >= slots.n_slots) ? kErrBadSlot : 0;
        Decoded d{};
        Decoded prev{};
        int rebuild = 3;
        if (!err) {
            d = decode_board(boards + i, lane, ws.mailbox[0]);
            if (!d.ok) err = kErrBadBoard;
        }
        if (!err && src) {
            const uint4* rec = reinterpret_cast<const uint4*>(slots.boards + from); /* written by earlier launches: plain loads */
            prev = decode_board(rec[0], rec[1], lane, ws.mailbox[1]);
            if (!prev.ok) err = kErrBadBoard; /* the source slot was never filled */
        }
        if (!err) {
            rebuild = build_lists(t, src ? &prev.view : nullptr, d, lane, ws);
            if (rebuild < 0) err = kErrCapacity;
        }
        if (err) {
            if (lane == 0) {
                flag_error(status, err);
                if (bucket) bucket[i] = 0xFF;
            }
            continue;
        }
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t v[16];
            if ((rebuild >> c) & 1) {
                rebuild_perspective_cold(net, ws.psq_add[c], ws.n_psq_add[c], ws.thr_add[c], ws.n_thr_add[c], lane, v);
            } else {
                load_slot_acc(slots, from, c, lane, v);
                update_perspective(net, ws, c, lane, v);
            }
            store_slot_acc(slots, to, c, lane, v);
            if (act) {
                const int half = c == d.view.stm ? 0 : 1;
                reinterpret_cast<uint4*>(act + i * SP_L1_SIZE)[half * 32 + lane] = activate(v);
            }
        }
        if (lane < 2) reinterpret_cast<uint4*>(slots.boards + to)[lane] = __ldg(reinterpret_cast<const uint4*>(boards + i) + lane);
        if (act && lane == 0) bucket[i] = static_cast<uint8_t>(output_bucket(d.view.occ));
    }
    }
    if (SP_SLOTS_TICKET && lane == 0) {
        const unsigned long long finished = atomicAdd(&status->counters[kSlotsTicket + 1], 1ull) + 1;
        if (finished == static_cast<unsigned long long>(gridDim.x) * kWarpsPerCta) { /* every warp has drawn its last (empty) ticket */
            status->counters[kSlotsTicket] = 0, status->counters[kSlotsTicket + 1] = 0;
            __threadfence();
        }
    }
}

/* slots -> activations for an explicit side to move (NnueState::evaluate, nnue_state.cpp:598-610) */
__global__ void __launch_bounds__(kThreads)
slot_activate_kernel(SlotStore slots, const uint32_t* __restrict__ ids, const uint8_t* __restrict__ stm, size_t n,
                     uint8_t* __restrict__ act, uint8_t* __restrict__ bucket, DeviceStatus* status) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t stride = static_cast<size_t>(gridDim.x) * kWarpsPerCta;
    for (size_t i = static_cast<size_t>(blockIdx.x) * kWarpsPerCta + warp; i < n; i += stride) {
        const uint32_t slot = ids[i];
        if (slot >= slots.n_slots) {
            if (lane == 0) {
                flag_error(status, kErrBadSlot);
                bucket[i] = 0xFF;
            }
            continue;
        }
        const uint4* rec = reinterpret_cast<const uint4*>(slots.boards + slot);
        const uint4 lo = rec[0], hi = rec[1];
        const uint64_t occ = static_cast<uint64_t>(lo.y) << 32 | lo.x;
        const int n_pieces = __popcll(occ);
        if (n_pieces < 2 || n_pieces > 32) {
            if (lane == 0) {
                flag_error(status, kErrBadBoard);
                bucket[i] = 0xFF;
            }
            continue;
        }
        const int side = stm ? (stm[i] & 1) : ((hi.z & 0x80) ? kBlack : kWhite);
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
            uint32_t v[16];
            load_slot_acc(slots, slot, c, lane, v);
            const int half = c == side ? 0 : 1;
            reinterpret_cast<uint4*>(act + i * SP_L1_SIZE)[half * 32 + lane] = activate(v);
        }
        if (lane == 0) bucket[i] = static_cast<uint8_t>(output_bucket(occ));
    }
}

/* ------------------------------------------------------------------ playout walker */

/* King squares of a packed record without decoding the board: nibble i belongs to the i-th occupied
 * square (marlinformat.h:43-68).  Returns false for a malformed record. */
__device__ __forceinline__ bool find_kings(const SpPackedBoard* board, int (&king)[2]) {
    const uint4* p = reinterpret_cast<const uint4*>(board);
    const uint4 lo = __ldg(p), hi = __ldg(p + 1);
    const uint64_t occ = static_cast<uint64_t>(lo.y) << 32 | lo.x;
    const int n = __popcll(occ);
    if (n < 2 || n > 32) return false;
    const uint32_t words[4] = {lo.z, lo.w, hi.x, hi.y};
    int idx[2] = {-1, -1}, found = 0;
#pragma unroll
    for (int w = 0; w < 4; ++w) {
        /* a nibble equals 5 (white king) or 13 (black king) iff its low three bits are 101 */
        uint32_t m = words[w] ^ 0x22222222u; /* low three bits 101 -> 111 */
        m = m & (m >> 1) & (m >> 2) & 0x11111111u;
        while (m) {
            const int nib = (__ffs(m) - 1) >> 2;
            m &= m - 1;
            const int i = w * 8 + nib;
            if (i < n) {
                idx[(words[w] >> (nib * 4 + 3)) & 1 ? kBlack : kWhite] = i;
                ++found;
            }
        }
    }
    if (found != 2 || idx[0] < 0 || idx[1] < 0) return false;
    king[kBlack] = nth_piece_square(occ, idx[kBlack]);
    king[kWhite] = nth_piece_square(occ, idx[kWhite]);
    return true;
}

struct KingPair {
    int king[2];
};

/* One warp per game, one lane per ply: which (board, perspective) pairs need a fresh accumulator. */
__global__ void __launch_bounds__(256)
plan_rebuilds_kernel(DeviceNet net, RebuildPlan plan, const SpPackedBoard* __restrict__ boards,
                     const uint32_t* __restrict__ game_start, uint32_t n_games, size_t n_boards) {
    const FeatureTables& t = *net.tables;
    const int lane = threadIdx.x & 31;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    for (uint32_t g = warp; g < n_games; g += n_warps) {
        const size_t first = min(static_cast<size_t>(game_start[g]), n_boards), last = min(static_cast<size_t>(game_start[g + 1]), n_boards);
        KingPair carry{};      /* kings of the board before this round of 32 plies */
        bool carry_ok = false; /* false at the start of the game: the first board is always rebuilt */
        for (size_t base = first; base < last; base += 32) {
            const size_t pos = base + lane;
            KingPair cur{};
            const bool ok = pos < last && find_kings(boards + pos, cur.king);
            KingPair prev;
            prev.king[0] = __shfl_up_sync(kFull, cur.king[0], 1);
            prev.king[1] = __shfl_up_sync(kFull, cur.king[1], 1);
            bool have_prev = __shfl_up_sync(kFull, ok ? 1 : 0, 1) != 0;
            if (lane == 0) prev = carry, have_prev = carry_ok;
            /* one reservation per round: a game's items sit next to each other in `items`, so the warps
             * that rebuild them run side by side and share the rows the positions have in common */
            const bool want0 = ok && (!have_prev || needs_refresh(t, prev, cur, kBlack));
            const bool want1 = ok && (!have_prev || needs_refresh(t, prev, cur, kWhite));
            const unsigned m0 = __ballot_sync(kFull, want0), m1 = __ballot_sync(kFull, want1);
            uint32_t first_slot = 0;
            if (lane == 0 && (m0 | m1)) first_slot = atomicAdd(&plan.counters[0], static_cast<uint32_t>(__popc(m0) + __popc(m1)));
            first_slot = __shfl_sync(kFull, first_slot, 0);
            if (pos < last) {
                const unsigned lt = (1u << lane) - 1;
                uint32_t at = first_slot + __popc(m0 & lt) + __popc(m1 & lt);
                uint2 slots = make_uint2(kNoRebuildSlot, kNoRebuildSlot);
                if (want0) {
                    if (at < plan.capacity) plan.items[at] = static_cast<uint32_t>(pos) * 2, slots.x = at;
                    ++at;
                }
                if (want1 && at < plan.capacity) plan.items[at] = static_cast<uint32_t>(pos) * 2 + 1, slots.y = at;
                *reinterpret_cast<uint2*>(plan.slot + 2 * pos) = slots;
            }
            carry.king[0] = __shfl_sync(kFull, cur.king[0], 31);
            carry.king[1] = __shfl_sync(kFull, cur.king[1], 31);
            carry_ok = __shfl_sync(kFull, ok ? 1 : 0, 31) != 0; /* after a rejected record the walker restarts the chain by itself */
        }
    }
}

/* One warp per planned item: rebuild that perspective exactly as the full refresh would. */
__global__ void __launch_bounds__(kThreads, 4)
run_rebuilds_kernel(DeviceNet net, RebuildPlan plan, const SpPackedBoard* __restrict__ boards, DeviceStatus* status) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const FeatureTables& t = *net.tables;
    const uint32_t first = plan.counters[1], last = min(plan.counters[0], plan.capacity);
    const uint32_t stride = gridDim.x * kWarpsPerCta;
    for (uint32_t i = first + blockIdx.x * kWarpsPerCta + warp; i < last; i += stride) {
        const uint32_t item = plan.items[i];
        const int c = item & 1;
        const Decoded d = decode_board(boards + (item >> 1), lane, ws.mailbox[0]);
        uint32_t v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) v[k] = 0;
        if (d.ok && build_lists(t, nullptr, d, lane, ws, 0, 1 << c) >= 0)
            rebuild_perspective(net, ws.psq_add[c], ws.n_psq_add[c], ws.thr_add[c], ws.n_thr_add[c], lane, v);
        /* a bad record is reported by the walker when it gets there */
        uint4* out = plan.acc + static_cast<size_t>(i) * 128 + lane;
#pragma unroll
        for (int k = 0; k < 4; ++k) out[32 * k] = make_uint4(v[k * 4 + 0], v[k * 4 + 1], v[k * 4 + 2], v[k * 4 + 3]);
    }
}

__global__ void rebuilds_done_kernel(RebuildPlan plan) { plan.counters[1] = min(plan.counters[0], plan.capacity); }

/* One warp plays through one game: both accumulators stay in registers / shared memory from ply to
 * ply (datagen form, src/datagen/datagen.cpp:257-262: applyMove + applyImmediately + evaluate). */
__global__ void __launch_bounds__(kThreads, SP_GAMES_MIN_BLOCKS)
ft_games_kernel(DeviceNet net, const SpPackedBoard* __restrict__ boards, const uint32_t* __restrict__ game_start,
                uint32_t n_games, size_t n_boards, uint8_t* __restrict__ act, uint8_t* __restrict__ bucket, RebuildPlan plan, DeviceStatus* status) {
    __shared__ WarpScratch scratch[kWarpsPerCta];
    __shared__ uint32_t parked[kWarpsPerCta][16][32]; /* one perspective's registers, parked between passes */
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    WarpScratch& ws = scratch[warp];
    const FeatureTables& t = *net.tables;
    const uint32_t stride = gridDim.x * kWarpsPerCta;
    for (uint32_t g = blockIdx.x * kWarpsPerCta + warp; g < n_games; g += stride) {
        /* the scratch rows are sized from the caller's n_boards: offsets past it are reported, never followed */
        if (game_start[g + 1] > n_boards && lane == 0) flag_error(status, kErrBadSlot);
        const size_t first = min(static_cast<size_t>(game_start[g]), n_boards), last = min(static_cast<size_t>(game_start[g + 1]), n_boards);
        uint32_t v[16];
        int in_regs = 0; /* which perspective `v` currently holds */
        BoardView prev{};
        bool have_prev = false;
        /* the next record (and its rebuild slots) are fetched one ply ahead so their latency hides behind
         * this ply's work */
        uint4 lo = make_uint4(0, 0, 0, 0), hi = lo;
        uint2 slots = make_uint2(kNoRebuildSlot, kNoRebuildSlot);
        if (first < last) {
            const uint4* p = reinterpret_cast<const uint4*>(boards + first);
            lo = __ldg(p), hi = __ldg(p + 1);
            if (plan.slot) slots = *reinterpret_cast<const uint2*>(plan.slot + 2 * first);
        }
#pragma unroll 1
        for (size_t pos = first; pos < last; ++pos) {
            const int buf = static_cast<int>(pos - first) & 1;
            const Decoded d = decode_board(lo, hi, lane, ws.mailbox[buf]);
            const uint2 my_slots = slots;
            if (pos + 1 < last) {
                const uint4* p = reinterpret_cast<const uint4*>(boards + pos + 1);
                lo = __ldg(p), hi = __ldg(p + 1);
                if (plan.slot) slots = *reinterpret_cast<const uint2*>(plan.slot + 2 * (pos + 1));
            }
            const int external = (my_slots.x != kNoRebuildSlot ? 1 : 0) | (my_slots.y != kNoRebuildSlot ? 2 : 0);
            int err = d.ok ? 0 : kErrBadBoard;
            int rebuild = 3;
            if (!err) {
                rebuild = build_lists(t, have_prev ? &prev : nullptr, d, lane, ws, external);
                if (rebuild < 0) err = kErrCapacity;
            }
            if (err) {
                if (lane == 0) {
                    flag_error(status, err);
                    bucket[pos] = 0xFF;
                }
                have_prev = false; /* the next good board restarts the chain */
                continue;
            }
            /* One copy of the loop body.  `v` holds the perspective being advanced, the other one is
             * parked in shared memory; they swap once per ply and the order alternates, so that each
             * ply starts with the perspective that is already in registers. */
#pragma unroll 1
            for (int pass = 0; pass < 2; ++pass) {
                const int c = in_regs;
                if ((external >> c) & 1) {
                    const uint4* fresh = plan.acc + static_cast<size_t>(c ? my_slots.y : my_slots.x) * 128 + lane;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const uint4 q = fresh[32 * k];
                        v[k * 4 + 0] = q.x, v[k * 4 + 1] = q.y, v[k * 4 + 2] = q.z, v[k * 4 + 3] = q.w;
                    }
                } else if ((rebuild >> c) & 1) {
                    uint32_t fresh[16];
                    rebuild_perspective_cold(net, ws.psq_add[c], ws.n_psq_add[c], ws.thr_add[c], ws.n_thr_add[c], lane, fresh);
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fresh[i];
                } else {
                    update_perspective(net, ws, c, lane, v);
                }
                const int half = c == d.view.stm ? 0 : 1;
                reinterpret_cast<uint4*>(act + pos * SP_L1_SIZE)[half * 32 + lane] = activate(v);
                if (pass == 0) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t other = parked[warp][i][lane];
                        parked[warp][i][lane] = v[i];
                        v[i] = other;
                    }
                    in_regs ^= 1;
                }
            }
            if (lane == 0) bucket[pos] = static_cast<uint8_t>(output_bucket(d.view.occ));
            prev = d.view;
            have_prev = true;
        }
    }
}

/* ------------------------------------------------------------------ dense head */

__device__ __forceinline__ void mma_u8s8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void mma_u8u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

constexpr int kHeadWarps = 8;
constexpr int kHeadRows = 256;  /* positions per CTA */
constexpr int kW1Bytes = SP_L1_SIZE * SP_L2_SIZE;                  /* one bucket's L1 weights: 32 KB */
constexpr int kW2Words = 2 * SP_L2_SIZE * SP_L3_SIZE;              /* one bucket's L2 weights: 4096 int32 */
constexpr int kLimbStride = 80; /* bytes per row of a limb plane: 64 inputs + padding so A-fragment reads hit 32 banks */

struct HeadShared {
    __align__(16) int8_t w1[kW1Bytes];                      /* current bucket, reference layout [k/4][o][k%4] */
    __align__(16) uint32_t w2[kW2Words];                    /* current bucket: byte limbs as B fragments (l2_limb_index) */
    __align__(16) uint8_t l2in[kHeadWarps][4][16][kLimbStride]; /* L2 inputs (skip >> 6) as four byte limbs: [limb][row][input] */
    uint32_t rows[kHeadRows];                               /* this CTA's slice of the bucket-grouped order */
    uint8_t tile_bucket[kHeadRows / 16];                    /* bucket of each 16-row tile (0xFF = empty tile) */
};

/* ---- counting sort of a launch's positions by output bucket (three tiny kernels) */
__device__ __forceinline__ void head_span(const uint32_t* range, uint32_t range_len, size_t& first, size_t& n) {
    first = 0;
    if (range) { /* positions [range[0], range[range_len]), never past the n boards the caller says exist */
        const size_t bound = n, last = range[range_len] < bound ? range[range_len] : bound;
        first = range[0] < last ? range[0] : last;
        n = last - first;
    }
}

/* Sixteen consecutive bucket bytes of this lane (rows i .. i + 15 of the span; 0xFE past the end), as per-bucket counts packed in 16-bit
 * fields: lo = buckets 0..3, hi = buckets 4..7.  One 128-bit load when the span allows it. */
struct Buckets16 {
    uint8_t b[16];
};
/* `p` is 16-byte aligned; the span's rows are bytes [lead, n) of it (lead < 16: a span may start anywhere, e.g. at a game boundary of
 * a playout stream, and is then read from the aligned address below it) */
__device__ __forceinline__ Buckets16 load_buckets16(const uint8_t* __restrict__ p, size_t i, size_t lead, size_t n) {
    Buckets16 r;
    if (i >= lead && i + 16 <= n) {
        const uint4 v = __ldg(reinterpret_cast<const uint4*>(p + i));
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 16; ++k) r.b[k] = static_cast<uint8_t>(w[k >> 2] >> (8 * (k & 3)));
    } else {
#pragma unroll
        for (int k = 0; k < 16; ++k) r.b[k] = (i + k >= lead && i + k < n) ? __ldg(p + i + k) : 0xFE;
    }
    return r;
}
__device__ __forceinline__ void count_buckets16(const Buckets16& r, uint64_t& lo, uint64_t& hi) {
    lo = hi = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const uint32_t b = r.b[k];
        if (b < 4) lo += uint64_t{1} << (16 * b);
        else if (b < SP_OUTPUT_BUCKETS) hi += uint64_t{1} << (16 * (b - 4));
    }
}
constexpr int kSortRowsPerWarp = 512; /* 16 rows per lane and iteration */

/* Counting sort, pass 1: rows per bucket.  A warp counts 512 rows per iteration in registers (no atomics), a block adds its eight
 * totals to the global counters once. */
__global__ void head_hist_kernel(const uint8_t* __restrict__ bucket, size_t n, const uint32_t* __restrict__ range, uint32_t range_len,
                                 HeadSort sort) {
    __shared__ unsigned hist[SP_OUTPUT_BUCKETS];
    size_t first;
    head_span(range, range_len, first, n);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < SP_OUTPUT_BUCKETS) hist[threadIdx.x] = 0;
    __syncthreads();
    uint32_t mine[SP_OUTPUT_BUCKETS] = {0, 0, 0, 0, 0, 0, 0, 0};
    const size_t warp_global = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, n_warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t lead = reinterpret_cast<uintptr_t>(bucket + first) & 15, end = lead + n; /* byte range of the span below its aligned base */
    const uint8_t* const aligned = bucket + first - lead;
    for (size_t base = warp_global * kSortRowsPerWarp; base < end; base += n_warps * kSortRowsPerWarp) {
        uint64_t lo, hi;
        count_buckets16(load_buckets16(aligned, base + 16 * lane, lead, end), lo, hi);
#pragma unroll
        for (int b = 0; b < 4; ++b) mine[b] += static_cast<uint32_t>(lo >> (16 * b)) & 0xFFFFu, mine[4 + b] += static_cast<uint32_t>(hi >> (16 * b)) & 0xFFFFu;
    }
#pragma unroll
    for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b) {
        uint32_t v = mine[b];
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) v += __shfl_xor_sync(kFull, v, d);
        if (lane == 0 && v) atomicAdd(&hist[b], v);
    }
    __syncthreads();
    if (threadIdx.x < SP_OUTPUT_BUCKETS && hist[threadIdx.x]) atomicAdd(&sort.counters[threadIdx.x], hist[threadIdx.x]);
}

/* Pass 2.  Group starts (each group padded to whole tiles) from the histogram; every block derives them for itself, block 0
 * publishes them for the head kernel and fills the padding slots.  A warp then places 512 rows per iteration: per-lane counts ->
 * exclusive prefix over the lanes (packed 16-bit fields, five shuffle steps) -> ONE reservation per bucket for the whole warp (the
 * first version reserved per warp and 32 rows: 262,144 atomics on eight addresses at M = 2^20, 26 us) -> each lane writes its 16
 * rows. */
__global__ void head_scatter_kernel(const uint8_t* __restrict__ bucket, size_t n, const uint32_t* __restrict__ range, uint32_t range_len,
                                    HeadSort sort, int32_t* __restrict__ out) {
    __shared__ uint32_t start[SP_OUTPUT_BUCKETS + 1];
    if (threadIdx.x == 0) {
        uint32_t at = 0;
        for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b) {
            start[b] = at;
            at += (sort.counters[b] + kHeadGroupPad - 1) & ~(kHeadGroupPad - 1);
        }
        start[SP_OUTPUT_BUCKETS] = at;
    }
    __syncthreads();
    if (blockIdx.x == 0) {
        if (threadIdx.x <= SP_OUTPUT_BUCKETS) sort.counters[16 + threadIdx.x] = start[threadIdx.x];
        for (uint32_t i = threadIdx.x; i < SP_OUTPUT_BUCKETS * kHeadGroupPad; i += blockDim.x) {
            const uint32_t b = i / kHeadGroupPad, at = start[b] + sort.counters[b] + i % kHeadGroupPad;
            if (at < start[b + 1]) sort.order[at] = kHeadNoRow;
        }
    }
    size_t first;
    head_span(range, range_len, first, n);
    const int lane = threadIdx.x & 31;
    const size_t warp_global = (static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5, n_warps = (static_cast<size_t>(gridDim.x) * blockDim.x) >> 5;
    const size_t lead = reinterpret_cast<uintptr_t>(bucket + first) & 15, end = lead + n;
    const uint8_t* const aligned = bucket + first - lead;
    for (size_t base = warp_global * kSortRowsPerWarp; base < end; base += n_warps * kSortRowsPerWarp) {
        const size_t i0 = base + 16 * lane;
        const Buckets16 rows = load_buckets16(aligned, i0, lead, end);
        uint64_t lo, hi;
        count_buckets16(rows, lo, hi);
        /* inclusive prefix over the lanes; a field holds at most 512 */
        uint64_t plo = lo, phi = hi;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t qlo = __shfl_up_sync(kFull, plo, d), qhi = __shfl_up_sync(kFull, phi, d);
            if (lane >= d) plo += qlo, phi += qhi;
        }
        const uint64_t tlo = __shfl_sync(kFull, plo, 31), thi = __shfl_sync(kFull, phi, 31); /* the warp's totals */
        /* lane b reserves the warp's rows of bucket b */
        uint32_t reserved = 0;
        if (lane < SP_OUTPUT_BUCKETS) {
            const uint32_t total = static_cast<uint32_t>((lane < 4 ? tlo : thi) >> (16 * (lane & 3))) & 0xFFFFu;
            if (total) reserved = atomicAdd(&sort.counters[8 + lane], total);
        }
        uint32_t at[SP_OUTPUT_BUCKETS]; /* where this lane's next row of bucket b goes */
#pragma unroll
        for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b) {
            const uint32_t before = (static_cast<uint32_t>(((b < 4 ? plo : phi) - (b < 4 ? lo : hi)) >> (16 * (b & 3)))) & 0xFFFFu;
            at[b] = start[b] + __shfl_sync(kFull, reserved, b) + before;
        }
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t b = rows.b[k];
            if (b < SP_OUTPUT_BUCKETS) {
                uint32_t slot = 0;
#pragma unroll
                for (int j = 0; j < SP_OUTPUT_BUCKETS; ++j)
                    if (b == static_cast<uint32_t>(j)) slot = at[j]++;
                sort.order[slot] = static_cast<uint32_t>(first + i0 + k - lead);
            } else if (i0 + k >= lead && i0 + k < end) {
                out[first + i0 + k - lead] = INT32_MIN; /* rejected board */
            }
        }
    }
}
/* The same sort for small launches (latency matters there: search-style batches of a few thousand
 * positions): one block does histogram, group starts, padding and scatter, and needs no zeroed counters. */
constexpr size_t kHeadSortSmall = 8192;
__global__ void __launch_bounds__(1024)
head_sort_small_kernel(const uint8_t* __restrict__ bucket, size_t n, const uint32_t* __restrict__ range, uint32_t range_len, HeadSort sort,
                       int32_t* __restrict__ out) {
    __shared__ uint32_t hist[SP_OUTPUT_BUCKETS], cursor[SP_OUTPUT_BUCKETS], start[SP_OUTPUT_BUCKETS + 1];
    size_t first;
    head_span(range, range_len, first, n);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < SP_OUTPUT_BUCKETS) hist[threadIdx.x] = cursor[threadIdx.x] = 0;
    __syncthreads();
    for (size_t i = threadIdx.x; i < n; i += blockDim.x) {
        const int b = bucket[first + i];
        if (b < SP_OUTPUT_BUCKETS) atomicAdd(&hist[b], 1u);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t at = 0;
        for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b) {
            start[b] = at;
            at += (hist[b] + kHeadGroupPad - 1) & ~(kHeadGroupPad - 1);
        }
        start[SP_OUTPUT_BUCKETS] = at;
    }
    __syncthreads();
    if (threadIdx.x <= SP_OUTPUT_BUCKETS) sort.counters[16 + threadIdx.x] = start[threadIdx.x];
    for (uint32_t i = threadIdx.x; i < SP_OUTPUT_BUCKETS * kHeadGroupPad; i += blockDim.x) {
        const uint32_t b = i / kHeadGroupPad, at = start[b] + hist[b] + i % kHeadGroupPad;
        if (at < start[b + 1]) sort.order[at] = kHeadNoRow;
    }
    const size_t n_round = (n + 31) & ~size_t{31};
    for (size_t i = threadIdx.x; i < n_round; i += blockDim.x) {
        const int b = i < n ? bucket[first + i] : 0xFE;
        const unsigned peers = __match_any_sync(kFull, b);
        if (b < SP_OUTPUT_BUCKETS) {
            const int leader = __ffs(peers) - 1;
            uint32_t slot = 0;
            if (lane == leader) slot = atomicAdd(&cursor[b], static_cast<uint32_t>(__popc(peers)));
            slot = __shfl_sync(peers, slot, leader) + __popc(peers & ((1u << lane) - 1));
            sort.order[start[b] + slot] = static_cast<uint32_t>(first + i);
        } else if (i < n) {
            out[first + i] = INT32_MIN; /* rejected board */
        }
    }
}
constexpr uint32_t kNoRow = kHeadNoRow;

/*
 * One CTA = 256 rows of the bucket-grouped order: sixteen 16-row tiles, two per warp, that (except
 * where two groups meet) all belong to ONE bucket, whose L1 and L2 weights (48 KB) are staged in shared
 * memory once.
 *
 * L1: the contraction index k may be visited in any order as long as A and B agree.  Per 64-wide
 * k-step, lane (g = lane / 4, t = lane % 4) loads 16 contiguous activation bytes of rows g and
 * g + 8 (k = 64 s + 16 t ...) and, for each of the 4 k-quads inside them, 16 contiguous weight
 * bytes = outputs 4 g .. 4 g + 3 of that quad ([k/4][o][k%4] layout, multilayer.h:180-196).  MMA
 * column g of n-tile nt is therefore output o = 4 g + nt, and the C fragment of lane (g, t)
 * holds outputs 8 t + nt and 8 t + 4 + nt of rows g and g + 8.
 *
 * L2 (int32 weights, multilayer.h:261-343) also runs on the tensor cores, as ten byte-limb contractions
 * that are recombined modulo 2^32; L3 is a 64-term dot product closed with two shuffles.
 */
static_assert(sizeof(HeadShared) * 2 <= 227 * 1024, "two CTAs per SM");
__global__ void __launch_bounds__(kHeadWarps * 32, 2)
head_kernel(DeviceNet net, const uint8_t* __restrict__ act, const uint8_t* __restrict__ bucket, int32_t* __restrict__ out, HeadSort sort) {
    extern __shared__ __align__(16) unsigned char head_smem[];
    HeadShared& sh = *reinterpret_cast<HeadShared*>(head_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const uint32_t total = sort.counters[16 + SP_OUTPUT_BUCKETS]; /* padded: a multiple of 16 */
    const uint32_t slice = blockIdx.x * kHeadRows;
    if (slice >= total) return;
    constexpr size_t base = 0; /* `rows` holds absolute position indices */

    sh.rows[tid] = slice + tid < total ? sort.order[slice + tid] : kNoRow;
    __syncthreads();
    if (tid < kHeadRows / 16) {
        /* a group is padded at its end only, so a tile's first row tells its bucket */
        const uint32_t first = sh.rows[tid * 16];
        sh.tile_bucket[tid] = first == kNoRow ? 0xFF : bucket[first];
    }
    __syncthreads();
    unsigned present = 0;
#pragma unroll
    for (int i = 0; i < kHeadRows / 16; ++i)
        if (sh.tile_bucket[i] < SP_OUTPUT_BUCKETS) present |= 1u << sh.tile_bucket[i];

    for (unsigned todo = present; todo; todo &= todo - 1) {
        const int b = __ffs(todo) - 1;
        /* ---- stage this bucket's weights */
        {
            const uint4* src1 = reinterpret_cast<const uint4*>(net.l1_w + static_cast<size_t>(b) * kW1Bytes);
            uint4* dst1 = reinterpret_cast<uint4*>(sh.w1);
            for (int i = tid; i < kW1Bytes / 16; i += kHeadWarps * 32) dst1[i] = __ldg(src1 + i);
            const uint4* src2 = reinterpret_cast<const uint4*>(net.l2_limbs + static_cast<size_t>(b) * kW2Words);
            uint4* dst2 = reinterpret_cast<uint4*>(sh.w2);
            for (int i = tid; i < kW2Words / 4; i += kHeadWarps * 32) dst2[i] = __ldg(src2 + i);
        }
        __syncthreads();
        for (int tile = warp; tile < kHeadRows / 16; tile += kHeadWarps) {
            if (sh.tile_bucket[tile] != b) continue;
            const uint32_t* ord = sh.rows + tile * 16;
            const uint32_t row0 = ord[g], row1 = ord[g + 8];
            /* padding slots read the tile's first row; their results are never written */
            const uint4* a_row0 = reinterpret_cast<const uint4*>(act + (base + (row0 == kNoRow ? ord[0] : row0)) * SP_L1_SIZE) + t;
            const uint4* a_row1 = reinterpret_cast<const uint4*>(act + (base + (row1 == kNoRow ? ord[0] : row1)) * SP_L1_SIZE) + t;
            int c[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) c[i][j] = 0;
            const uint4* w = reinterpret_cast<const uint4*>(sh.w1) + g;
            /* Activation rows come from global memory (L2 or HBM): the loads of the next four k-steps are
             * issued before the MMAs of the current four, so only the first group's latency is exposed. */
            constexpr int kGroup = 4, kGroups = SP_L1_SIZE / 64 / kGroup;
            uint4 a_next[kGroup][2];
#pragma unroll
            for (int i = 0; i < kGroup; ++i) a_next[i][0] = __ldg(a_row0 + i * 4), a_next[i][1] = __ldg(a_row1 + i * 4);
#pragma unroll
            for (int grp = 0; grp < kGroups; ++grp) {
                uint4 a_cur[kGroup][2];
#pragma unroll
                for (int i = 0; i < kGroup; ++i) a_cur[i][0] = a_next[i][0], a_cur[i][1] = a_next[i][1];
                if (grp + 1 < kGroups) {
#pragma unroll
                    for (int i = 0; i < kGroup; ++i) {
                        a_next[i][0] = __ldg(a_row0 + ((grp + 1) * kGroup + i) * 4);
                        a_next[i][1] = __ldg(a_row1 + ((grp + 1) * kGroup + i) * 4);
                    }
                }
#pragma unroll
                for (int i = 0; i < kGroup; ++i) {
                    const int s = grp * kGroup + i;
                    uint4 q[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) q[j] = w[(s * 16 + t * 4 + j) * 8];
                    const uint32_t al[4] = {a_cur[i][0].x, a_cur[i][0].y, a_cur[i][0].z, a_cur[i][0].w};
                    const uint32_t ah[4] = {a_cur[i][1].x, a_cur[i][1].y, a_cur[i][1].z, a_cur[i][1].w};
#pragma unroll
                    for (int m = 0; m < 2; ++m) {
                        const uint32_t q0[4] = {q[2 * m].x, q[2 * m].y, q[2 * m].z, q[2 * m].w};
                        const uint32_t q1[4] = {q[2 * m + 1].x, q[2 * m + 1].y, q[2 * m + 1].z, q[2 * m + 1].w};
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) mma_u8s8(c[nt], al[2 * m], ah[2 * m], al[2 * m + 1], ah[2 * m + 1], q0[nt], q1[nt]);
                    }
                }
            }
            /* L1 epilogue + dual activation, multilayer.h:219-256 (kShift = -2).
             * The L1 outputs also feed L3 directly (skip connection, multilayer.h:353-446).  All sums are
             * modulo 2^32, so that term is summed here, where the outputs are in registers.
             * The L2 inputs (skip >> 6, range [-2^19, 4096]: the wrapped square may be negative) leave as
             * four unsigned byte limbs of their 32-bit two's complement, ready to be IMMA A fragments. */
            uint32_t skip_dot[2];
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int r = g + 8 * hrow;
                uint32_t dot = 0;
#pragma unroll
                for (int cc = 0; cc < 2; ++cc) {
                    uint32_t cr_limbs[4] = {0, 0, 0, 0}, sq_limbs[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        const int o = 8 * t + 4 * cc + nt;
                        const int x = static_cast<int>(static_cast<uint32_t>(c[nt][hrow * 2 + cc] >> 2)
                                                       + static_cast<uint32_t>(__ldg(net.l1_b + b * SP_L2_SIZE + o)));
                        const int cr = min(max(x, 0), 4096);
                        int sq = static_cast<int>(static_cast<uint32_t>(x) * static_cast<uint32_t>(x)); /* wraps BEFORE the min */
                        sq = min(sq, 16777216);
                        dot += static_cast<uint32_t>(cr << 6) * static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + o))
                             + static_cast<uint32_t>(sq >> 6) * static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + SP_L2_SIZE + o));
                        const uint32_t in_cr = static_cast<uint32_t>(cr);       /* (cr << 6) >> 6 */
                        const uint32_t in_sq = static_cast<uint32_t>(sq >> 12); /* (sq >> 6) >> 6, arithmetic */
#pragma unroll
                        for (int limb = 0; limb < 4; ++limb) {
                            cr_limbs[limb] |= ((in_cr >> (8 * limb)) & 0xFFu) << (8 * nt);
                            sq_limbs[limb] |= ((in_sq >> (8 * limb)) & 0xFFu) << (8 * nt);
                        }
                    }
                    /* inputs 8t + 4cc .. + 3 (CReLU half) and 32 + 8t + 4cc .. + 3 (squared half) of row r */
#pragma unroll
                    for (int limb = 0; limb < 4; ++limb) {
                        *reinterpret_cast<uint32_t*>(&sh.l2in[warp][limb][r][8 * t + 4 * cc]) = cr_limbs[limb];
                        *reinterpret_cast<uint32_t*>(&sh.l2in[warp][limb][r][SP_L2_SIZE + 8 * t + 4 * cc]) = sq_limbs[limb];
                    }
                }
                dot += __shfl_xor_sync(kFull, dot, 1);
                dot += __shfl_xor_sync(kFull, dot, 2);
                skip_dot[hrow] = dot;
            }
            __syncwarp();

            /* L2 on the tensor cores, multilayer.h:261-343: out = bias + sum_k in[k] * W2[k][o] (mod 2^32).
             * With in = sum_i a_i 2^(8i) and W2 = sum_j w_j 2^(8j) (unsigned byte limbs), the product modulo
             * 2^32 is sum over i + j <= 3 of (a_i w_j) << 8 (i + j): ten u8 x u8 -> s32 contractions of
             * length 64, each exact (64 * 255 * 255 < 2^31).  Lane (g, t) ends up with outputs
             * o = 8 nt + 2 t, + 1 of rows g and g + 8. */
            uint32_t l2[8][4];
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const uint32_t b0v = static_cast<uint32_t>(__ldg(net.l2_b + b * SP_L3_SIZE + nt * 8 + 2 * t));
                const uint32_t b1v = static_cast<uint32_t>(__ldg(net.l2_b + b * SP_L3_SIZE + nt * 8 + 2 * t + 1));
                l2[nt][0] = b0v, l2[nt][1] = b1v, l2[nt][2] = b0v, l2[nt][3] = b1v;
            }
#pragma unroll
            for (int shift = 0; shift < 4; ++shift) {
                int part[8][4];
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int k = 0; k < 4; ++k) part[nt][k] = 0;
#pragma unroll
                for (int i = 0; i <= shift; ++i) { /* input limb i with weight limb shift - i */
                    const int j = shift - i;
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint32_t a0 = *reinterpret_cast<const uint32_t*>(&sh.l2in[warp][i][g][ks * 32 + 4 * t]);
                        const uint32_t a1 = *reinterpret_cast<const uint32_t*>(&sh.l2in[warp][i][g + 8][ks * 32 + 4 * t]);
                        const uint32_t a2 = *reinterpret_cast<const uint32_t*>(&sh.l2in[warp][i][g][ks * 32 + 16 + 4 * t]);
                        const uint32_t a3 = *reinterpret_cast<const uint32_t*>(&sh.l2in[warp][i][g + 8][ks * 32 + 16 + 4 * t]);
#pragma unroll
                        for (int nt = 0; nt < 8; ++nt) {
                            const uint32_t* wl = sh.w2 + ((j * 8 + nt) * 16 + ks * 8 + t) * 8 + g; /* l2_limb_index */
                            mma_u8u8(part[nt], a0, a1, a2, a3, wl[0], wl[4 * 8]);
                        }
                    }
                }
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int k = 0; k < 4; ++k) l2[nt][k] += static_cast<uint32_t>(part[nt][k]) << (8 * shift);
            }

            /* L3 + scale, multilayer.h:345-447, 484-489 */
            uint32_t dot_lo = 0, dot_hi = 0; /* rows g and g + 8 */
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const uint32_t w3a = static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + nt * 8 + 2 * t));
                const uint32_t w3b = static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + nt * 8 + 2 * t + 1));
                dot_lo += static_cast<uint32_t>(min(max(static_cast<int>(l2[nt][0]), 0), 262144)) * w3a
                        + static_cast<uint32_t>(min(max(static_cast<int>(l2[nt][1]), 0), 262144)) * w3b;
                dot_hi += static_cast<uint32_t>(min(max(static_cast<int>(l2[nt][2]), 0), 262144)) * w3a
                        + static_cast<uint32_t>(min(max(static_cast<int>(l2[nt][3]), 0), 262144)) * w3b;
            }
            dot_lo += __shfl_xor_sync(kFull, dot_lo, 1), dot_hi += __shfl_xor_sync(kFull, dot_hi, 1);
            dot_lo += __shfl_xor_sync(kFull, dot_lo, 2), dot_hi += __shfl_xor_sync(kFull, dot_hi, 2);
            if (t == 0) {
                const uint32_t bias3 = static_cast<uint32_t>(__ldg(net.l3_b + b));
                const int32_t l3_lo = static_cast<int32_t>(dot_lo + skip_dot[0] + bias3), l3_hi = static_cast<int32_t>(dot_hi + skip_dot[1] + bias3);
                /* the final division truncates toward zero */
                if (row0 != kNoRow) out[base + row0] = static_cast<int32_t>(static_cast<int64_t>(l3_lo) * 400 / 16777216);
                if (row1 != kNoRow) out[base + row1] = static_cast<int32_t>(static_cast<int64_t>(l3_hi) * 400 / 16777216);
            }
            __syncwarp();
        }
        __syncthreads(); /* the next bucket overwrites the staged weights */
    }
}

/* ------------------------------------------------------------------ dense head, streaming form
 *
 * The same arithmetic as head_kernel, organised around the HBM stream (the head is the one kernel of
 * this library that is bound by it: 1 KB of activations per position, read once).
 *
 *   - one persistent CTA per SM works on a contiguous run of 32-row tiles of the bucket-grouped order;
 *   - kHeadStages tiles are kept in flight in a ring of stages: one bulk copy per activation row (TMA
 *     engine, cp.async.bulk global -> shared, 1 KB), completion counted in bytes on the stage's `full`
 *     mbarrier; rows are padded to 1040 B so that the A-fragment reads (ldmatrix: 8 rows x 16 B per matrix)
 *     touch every bank group once;
 *   - every warp is a consumer; warp w takes tiles w, w + kHeadWarpsStream, ...  A warp owns its tile's
 *     32 rows (two m16 tiles), so every weight fragment it fetches from shared memory feeds two IMMAs.
 *     As soon as L1 is done it releases the stage and itself requests the tile that follows on it -- the
 *     epilogue, L2 and L3 then run from registers while the TMA engine refills the stage.  (A dedicated
 *     producer warp was the first design: ncu showed that single issuing warp to be the bottleneck,
 *     profiles/r1_head_stream_ncu_v7.md.)
 *   - L1 -> L2 without a shared-memory round trip: the contraction index of an IMMA may be visited in
 *     any order as long as A and B agree, so L2's k-slots are DEFINED as the order in which L1's C
 *     fragment leaves the outputs in a lane (lane (g, t) holds outputs 8t .. 8t+7 of rows g and g+8 =
 *     k-slots 4t..4t+3 and 16+4t..16+4t+3), and the L2 weight limbs are laid out to match at upload
 *     (l2_fragment_index);
 *   - L2 by Horner's rule over the limb weight: acc = (acc << 8) + sum_{i+j = s} a_i w_j for s = 3..0,
 *     the IMMA accumulating straight into acc (s32 accumulation wraps, as everything here must);
 *     zero limbs are skipped (see the three forms at the L2 step below).
 */
#ifndef SP_HEAD_CONSUMERS
#define SP_HEAD_CONSUMERS 8 /* warps per CTA: two per scheduler at up to 255 registers (measured best; 12 = three at 168) */
#endif
#ifndef SP_HEAD_STAGES
#define SP_HEAD_STAGES 5
#endif
constexpr int kHeadConsumers = SP_HEAD_CONSUMERS;
constexpr int kStreamThreads = kHeadConsumers * 32;
constexpr int kTileRows = 32;
/* SP_HEAD_LDMATRIX (default): the A fragments of L1 come from ldmatrix.x4, which delivers the four fragment
 * registers of a lane directly, and the B fragment pairs are stored adjacent, so the L1 loop is 2 LDSM + 2
 * LDS.128 + 8 IMMA per 32-wide k-step.  The first version (SP_HEAD_LDMATRIX=0) packed A quads out of two
 * LDS.128 of different rows: that packing was 30 % of the kernel's executed instructions (IMAD.MOV,
 * profiles/r1_head_stream_ncu_v9.md); 307 -> 272 us at M = 2^20.  ldmatrix reads 8 rows x 16 B per matrix, so
 * the row padding is 16 B (8 consecutive rows on 8 bank groups) and the contraction index is visited in
 * natural order. */
#ifndef SP_HEAD_LDMATRIX
#define SP_HEAD_LDMATRIX 1
#endif
constexpr int kTileRowStride = SP_L1_SIZE + (SP_HEAD_LDMATRIX ? 16 : 64);
constexpr int kHeadStages = SP_HEAD_STAGES;

struct HeadStreamShared {
    __align__(128) uint8_t a[kHeadStages][kTileRows * kTileRowStride];
    __align__(16) uint4 w1[kW1Bytes / 16]; /* chunk (k-quad q, output quad g) at q * 8 + (g ^ ((q >> 2 & 3) << 1)) */
    __align__(16) uint2 w2[kW2Words / 2];  /* l2_fragment_index */
    __align__(8) uint64_t full[kHeadStages];
    uint32_t gen[kHeadStages]; /* tiles consumed so far on each stage */
    uint32_t rows[kHeadStages][kTileRows];
};
static_assert(sizeof(HeadStreamShared) <= 227 * 1024, "one CTA per SM");
static_assert(kHeadConsumers >= kHeadStages, "the first kHeadStages warps request the first tiles");

__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
__device__ __forceinline__ void consumers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kHeadConsumers * 32) : "memory"); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.b32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

/* Requests tile n of the CTA (lane l holds the position index of its row l, or kHeadNoRow) into stage
 * n % kHeadStages: one 1 KB bulk copy per row.  UBLKCP takes its operands from uniform registers; issued
 * per lane the compiler serialises the warp in a loop of ~64 clk per copy (measured: the issuing warp
 * was the kernel's bottleneck).  Broadcasting each row index with a shuffle and letting one elected lane
 * issue all the copies gives straight-line code whose R2UR latencies overlap. */
__device__ __noinline__ void head_issue_fill(HeadStreamShared& sh, const uint8_t* __restrict__ act, uint32_t n, uint32_t row, int lane) {
    const int stage = n % kHeadStages;
    sh.rows[stage][lane] = row;
    const uint32_t bytes = __popc(__ballot_sync(kFull, row != kHeadNoRow)) * SP_L1_SIZE;
    __syncwarp();
    const bool leader = elect_one();
    if (leader) mbar_expect_tx(&sh.full[stage], bytes);
    uint8_t* dst = &sh.a[stage][0];
#pragma unroll
    for (int i = 0; i < kTileRows; ++i) {
        const uint32_t r = __shfl_sync(kFull, row, i);
        if (r != kHeadNoRow && leader) bulk_copy_g2s(dst + i * kTileRowStride, act + static_cast<size_t>(r) * SP_L1_SIZE, SP_L1_SIZE, &sh.full[stage]);
    }
}

/* Everything after L1 for a warp's 32 rows (two m16 tiles), shared by head_stream_kernel (L1 by mma.sync) and head_umma_kernel (L1 by
 * tcgen05, read back with tcgen05.ld.16x256b -- the same C-fragment layout): c[mt][nt][2 hrow + cc] = L1 sum of row 16 mt + g + 8 hrow,
 * output 8 t + 4 cc + nt.  w2_frags = the bucket's L2 weights in shared memory (l2_fragment_index). */
__device__ __forceinline__ void head_tail(
    const DeviceNet& net, int b, bool narrow_w2, const int (&c)[2][4][4], const uint32_t (&row_id)[2][2], const uint2* w2_frags,
    int32_t* __restrict__ out, int lane) {
    const int t = lane & 3;
    /* ---- L1 epilogue (multilayer.h:219-256), skip term of L3, L2 inputs as byte limbs in A-fragment order:
     * register index hrow + 2 cc = a0..a3 of an IMMA (row g | g+8, k-slots 4t.. | 16+4t..) */
    uint32_t cr_l[2][2][4], sq_l[2][4][4], skip_dot[2][2];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            uint32_t dot = 0;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                uint32_t in_cr[4], in_sq[4];
#pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    const int o = 8 * t + 4 * cc + nt;
                    const int x = static_cast<int>(static_cast<uint32_t>(c[mt][nt][hrow * 2 + cc] >> 2)
                                                   + static_cast<uint32_t>(__ldg(net.l1_b + b * SP_L2_SIZE + o)));
                    const int cr = min(max(x, 0), 4096);
                    int sq = static_cast<int>(static_cast<uint32_t>(x) * static_cast<uint32_t>(x)); /* wraps BEFORE the min */
                    sq = min(sq, 16777216);
                    dot += static_cast<uint32_t>(cr << 6) * static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + o))
                         + static_cast<uint32_t>(sq >> 6) * static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + SP_L2_SIZE + o));
                    in_cr[nt] = static_cast<uint32_t>(cr);       /* (cr << 6) >> 6: at most 0x1000 */
                    in_sq[nt] = static_cast<uint32_t>(sq >> 12); /* (sq >> 6) >> 6, arithmetic */
                }
                /* 4 x 4 byte transposes: word `limb` = byte `limb` of the four values */
                {
                    const uint32_t lo01 = __byte_perm(in_cr[0], in_cr[1], 0x5140), lo23 = __byte_perm(in_cr[2], in_cr[3], 0x5140);
                    cr_l[mt][0][hrow + 2 * cc] = __byte_perm(lo01, lo23, 0x5410);
                    cr_l[mt][1][hrow + 2 * cc] = __byte_perm(lo01, lo23, 0x7632);
                }
                {
                    const uint32_t lo01 = __byte_perm(in_sq[0], in_sq[1], 0x5140), lo23 = __byte_perm(in_sq[2], in_sq[3], 0x5140);
                    const uint32_t hi01 = __byte_perm(in_sq[0], in_sq[1], 0x7362), hi23 = __byte_perm(in_sq[2], in_sq[3], 0x7362);
                    sq_l[mt][0][hrow + 2 * cc] = __byte_perm(lo01, lo23, 0x5410);
                    sq_l[mt][1][hrow + 2 * cc] = __byte_perm(lo01, lo23, 0x7632);
                    sq_l[mt][2][hrow + 2 * cc] = __byte_perm(hi01, hi23, 0x5410);
                    sq_l[mt][3][hrow + 2 * cc] = __byte_perm(hi01, hi23, 0x7632);
                }
            }
            dot += __shfl_xor_sync(kFull, dot, 1);
            dot += __shfl_xor_sync(kFull, dot, 2);
            skip_dot[mt][hrow] = dot;
        }

    /* ---- L2 (multilayer.h:261-343) by Horner over the limb weight; lane (g, t) ends up with outputs
     * 8 nt + 2t, + 1 of rows g and g + 8 */
    int acc[2][8][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 8; ++nt)
#pragma unroll
            for (int k = 0; k < 4; ++k) acc[mt][nt][k] = 0;
    const uint2* w2 = w2_frags + lane;
    auto shl8 = [&]() {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                for (int k = 0; k < 4; ++k) acc[mt][nt][k] = static_cast<int>(static_cast<uint32_t>(acc[mt][nt][k]) << 8);
    };
    /* The squared half of the inputs is negative only after a wrapped square (|x| > 46340): without
     * one in this tile its limbs 2 and 3 are zero like those of the CReLU half. */
    uint32_t high = 0;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int r = 0; r < 4; ++r) high |= sq_l[mt][2][r] | sq_l[mt][3][r];
    const bool wide_inputs = __any_sync(kFull, high != 0);
    if (narrow_w2 && !wide_inputs) {
        /* Weights that fit int16 (true of the whole bucket, found at upload) are lo + 256 hi with lo = byte 0
         * unsigned and hi = byte 1 SIGNED -- the same stored bytes, read by a u8 x s8 IMMA -- and inputs below
         * 2^16 are a0 + 256 a1: in * w = a0 lo + 2^8 (a0 hi + a1 lo) + 2^16 a1 hi, four contractions per
         * k-half instead of ten / seven. */
#pragma unroll
        for (int level = 2; level >= 0; --level) {
            if (level < 2) shl8();
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int j = level - i; /* input limb i with weight limb j */
                if (j < 0 || j > 1) continue;
#pragma unroll
                for (int nt = 0; nt < 8; ++nt)
#pragma unroll
                    for (int ks = 0; ks < 2; ++ks) {
                        const uint2 wf = w2[((j * 8 + nt) * 2 + ks) * 32];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt) {
                            const uint32_t* a = ks == 0 ? cr_l[mt][i] : sq_l[mt][i];
                            if (j == 1) mma_u8s8(acc[mt][nt], a[0], a[1], a[2], a[3], wf.x, wf.y);
                            else mma_u8u8(acc[mt][nt], a[0], a[1], a[2], a[3], wf.x, wf.y);
                        }
                    }
            }
        }
    } else {
#pragma unroll
        for (int shift = 3; shift >= 0; --shift) {
            if (shift < 3) shl8();
#pragma unroll
            for (int i = 0; i <= shift; ++i) { /* input limb i with weight limb shift - i */
                const int j = shift - i;
                if (i >= 2 && !wide_inputs) continue; /* warp-uniform */
#pragma unroll
                for (int nt = 0; nt < 8; ++nt) {
                    if (i < 2) {
                        const uint2 wf = w2[((j * 8 + nt) * 2 + 0) * 32];
#pragma unroll
                        for (int mt = 0; mt < 2; ++mt)
                            mma_u8u8(acc[mt][nt], cr_l[mt][i][0], cr_l[mt][i][1], cr_l[mt][i][2], cr_l[mt][i][3], wf.x, wf.y);
                    }
                    const uint2 wf = w2[((j * 8 + nt) * 2 + 1) * 32];
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt)
                        mma_u8u8(acc[mt][nt], sq_l[mt][i][0], sq_l[mt][i][1], sq_l[mt][i][2], sq_l[mt][i][3], wf.x, wf.y);
                }
            }
        }
    }

    /* ---- L3 + scale, multilayer.h:345-447, 484-489 */
    uint32_t dot[2][2] = {{0, 0}, {0, 0}};
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
        const uint32_t b2a = static_cast<uint32_t>(__ldg(net.l2_b + b * SP_L3_SIZE + nt * 8 + 2 * t));
        const uint32_t b2b = static_cast<uint32_t>(__ldg(net.l2_b + b * SP_L3_SIZE + nt * 8 + 2 * t + 1));
        const uint32_t w3a = static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + nt * 8 + 2 * t));
        const uint32_t w3b = static_cast<uint32_t>(__ldg(net.l3_w + b * SP_L3_SIZE + nt * 8 + 2 * t + 1));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
            for (int hrow = 0; hrow < 2; ++hrow) {
                const int va = static_cast<int>(static_cast<uint32_t>(acc[mt][nt][2 * hrow]) + b2a);
                const int vb = static_cast<int>(static_cast<uint32_t>(acc[mt][nt][2 * hrow + 1]) + b2b);
                dot[mt][hrow] += static_cast<uint32_t>(min(max(va, 0), 262144)) * w3a + static_cast<uint32_t>(min(max(vb, 0), 262144)) * w3b;
            }
    }
    const uint32_t bias3 = static_cast<uint32_t>(__ldg(net.l3_b + b));
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int hrow = 0; hrow < 2; ++hrow) {
            uint32_t d = dot[mt][hrow];
            d += __shfl_xor_sync(kFull, d, 1);
            d += __shfl_xor_sync(kFull, d, 2);
            const int32_t l3 = static_cast<int32_t>(d + skip_dot[mt][hrow] + bias3);
            /* the final division truncates toward zero */
            if (t == 0 && row_id[mt][hrow] != kHeadNoRow) out[row_id[mt][hrow]] = static_cast<int32_t>(static_cast<int64_t>(l3) * 400 / 16777216);
        }
}

__global__ void __launch_bounds__(kStreamThreads, 1)
head_stream_kernel(DeviceNet net, const uint8_t* __restrict__ act, int32_t* __restrict__ out, HeadSort sort) {
    extern __shared__ __align__(16) unsigned char head_smem[];
    HeadStreamShared& sh = *reinterpret_cast<HeadStreamShared*>(head_smem);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    /* this CTA's run of tiles */
    const uint32_t n_tiles = sort.counters[16 + SP_OUTPUT_BUCKETS] / kTileRows; /* groups are padded to whole tiles */
    const uint32_t t0 = static_cast<uint32_t>(static_cast<uint64_t>(n_tiles) * blockIdx.x / gridDim.x);
    const uint32_t t1 = static_cast<uint32_t>(static_cast<uint64_t>(n_tiles) * (blockIdx.x + 1) / gridDim.x);
    if (t0 >= t1) return;

    if (tid == 0) {
        for (int i = 0; i < kHeadStages; ++i) mbar_init(&sh.full[i], 1), sh.gen[i] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    auto issue_fill = [&](uint32_t n, uint32_t row) { head_issue_fill(sh, act, n, row, lane); };
    const uint32_t my_tiles = t1 - t0;
    if (warp < kHeadStages && static_cast<uint32_t>(warp) < my_tiles) {
        /* the first kHeadStages tiles are requested here, every later tile by the warp that releases its stage */
        issue_fill(warp, sort.order[static_cast<size_t>(t0 + warp) * kTileRows + lane]);
    }

    /* ---- consumers */
    uint32_t cur = t0;
    while (cur < t1) {
        /* the bucket of tile `cur` and the end of its group, from the group starts */
        int b = 0;
#pragma unroll
        for (int i = 1; i < SP_OUTPUT_BUCKETS; ++i)
            if (sort.counters[16 + i] <= cur * kTileRows) b = i;
        const uint32_t seg_end = min(t1, sort.counters[16 + b + 1] / kTileRows);
        const bool narrow_w2 = (net.l2_narrow >> b) & 1;

        consumers_sync(); /* nobody still reads the previous bucket's weights */
        {
#if SP_HEAD_LDMATRIX
            /* B fragments as the IMMAs want them: for k-step S (32 wide), lane (g, t) and n-tile nt the pair
             * (b0, b1) = k-quads 8S + t and 8S + 4 + t of output 4g + nt sit in adjacent words, eight words per
             * lane and k-step = two LDS.128; 16-byte chunks XOR-swizzled by t so that the eight lanes of a
             * quarter warp (two g, four t) hit eight bank groups. */
            const uint32_t* src1 = reinterpret_cast<const uint32_t*>(net.l1_w + static_cast<size_t>(b) * kW1Bytes);
            uint32_t* dst1 = reinterpret_cast<uint32_t*>(sh.w1);
            for (int i = tid; i < kW1Bytes / 4; i += kHeadConsumers * 32) {
                const int q = i >> 5, o = i & 31; /* source word: k-quad q, output o */
                const int S = q >> 3, half = (q >> 2) & 1, tt = q & 3, gg = o >> 2, nt = o & 3;
                const int chunk = (((S * 4 + tt) * 8 + gg) * 2 + (nt >> 1)) ^ ((tt & 1) | ((tt >> 1) << 2));
                dst1[chunk * 4 + (nt & 1) * 2 + half] = __ldg(src1 + i);
            }
#else
            const uint4* src1 = reinterpret_cast<const uint4*>(net.l1_w + static_cast<size_t>(b) * kW1Bytes);
            for (int i = tid; i < kW1Bytes / 16; i += kHeadConsumers * 32) {
                const int q = i >> 3, og = i & 7;
                sh.w1[q * 8 + (og ^ (((q >> 2) & 3) << 1))] = __ldg(src1 + i);
            }
#endif
            const uint4* src2 = reinterpret_cast<const uint4*>(net.l2_frags + static_cast<size_t>(b) * kW2Words);
            uint4* dst2 = reinterpret_cast<uint4*>(sh.w2);
            for (int i = tid; i < kW2Words / 4; i += kHeadConsumers * 32) dst2[i] = __ldg(src2 + i);
        }
        consumers_sync();

        for (uint32_t tile = cur + warp; tile < seg_end; tile += kHeadConsumers) {
            const uint32_t n = tile - t0;
            const int stage = n % kHeadStages;
            /* A phase parity tells two consecutive uses of a stage apart, no more, and there are more
             * consumer warps than stages: before waiting for ITS fill a warp makes sure every earlier tile
             * of this stage has been consumed (then the barrier can only be in this tile's phase). */
            while (*reinterpret_cast<volatile uint32_t*>(&sh.gen[stage]) != n / kHeadStages) __nanosleep(64);
            /* rows of the tile that will follow this one on the stage (requested once L1 is done) */
            const uint32_t refill = n + kHeadStages;
            uint32_t refill_row = kHeadNoRow;
            if (refill < my_tiles) refill_row = sort.order[static_cast<size_t>(t0 + refill) * kTileRows + lane];
            mbar_wait(&sh.full[stage], (n / kHeadStages) & 1);

            /* ---- L1: 32 rows x 32 outputs, k = 1024 */
            int c[2][4][4];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt)
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[mt][i][j] = 0;
#if SP_HEAD_LDMATRIX
            /* lane l addresses row (l & 15) of an m-tile at byte 16 (l >> 4) of the 32-byte k-step: matrices 0..3 =
             * (rows 0-7 | 8-15) x (bytes 0-15 | 16-31) = a0..a3 of mma.m16n8k32 */
            const uint32_t a_lane = smem_addr(&sh.a[stage][(lane & 15) * kTileRowStride + 16 * (lane >> 4)]);
            /* chunk of (k-step ks, lane) = ((ks * 4 + t) * 8 + g) * 2 + pair, swizzled in bits 0 and 2: 64 ks apart */
            const int w_swz = (t & 1) | ((t >> 1) << 2);
            const uint4* w01 = sh.w1 + (((t * 8 + g) * 2) ^ w_swz);
            const uint4* w23 = sh.w1 + (((t * 8 + g) * 2 + 1) ^ w_swz);
#pragma unroll 4
            for (int ks = 0; ks < SP_L1_SIZE / 32; ++ks) {
                uint32_t a[2][4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                                 : "=r"(a[mt][0]), "=r"(a[mt][1]), "=r"(a[mt][2]), "=r"(a[mt][3])
                                 : "r"(a_lane + (16 * mt) * kTileRowStride + 32 * ks));
                const uint4 b01 = w01[64 * ks], b23 = w23[64 * ks]; /* (b0, b1) of n-tiles 0, 1 | 2, 3 */
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    mma_u8s8(c[mt][0], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b01.x, b01.y);
                    mma_u8s8(c[mt][1], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b01.z, b01.w);
                    mma_u8s8(c[mt][2], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b23.x, b23.y);
                    mma_u8s8(c[mt][3], a[mt][0], a[mt][1], a[mt][2], a[mt][3], b23.z, b23.w);
                }
            }
#else
            const uint8_t* a_base = &sh.a[stage][g * kTileRowStride + 16 * t];
            const uint4* w_base = sh.w1 + (g ^ (t << 1));
#pragma unroll 4
            for (int s = 0; s < SP_L1_SIZE / 64; ++s) {
                uint4 al[2], ah[2], q[4];
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) {
                    al[mt] = *reinterpret_cast<const uint4*>(a_base + (16 * mt) * kTileRowStride + 64 * s);
                    ah[mt] = *reinterpret_cast<const uint4*>(a_base + (16 * mt + 8) * kTileRowStride + 64 * s);
                }
#pragma unroll
                for (int j = 0; j < 4; ++j) q[j] = w_base[(s * 16 + t * 4 + j) * 8];
#pragma unroll
                for (int m = 0; m < 2; ++m) {
                    const uint32_t q0[4] = {q[2 * m].x, q[2 * m].y, q[2 * m].z, q[2 * m].w};
                    const uint32_t q1[4] = {q[2 * m + 1].x, q[2 * m + 1].y, q[2 * m + 1].z, q[2 * m + 1].w};
#pragma unroll
                    for (int mt = 0; mt < 2; ++mt) {
                        const uint32_t lo[4] = {al[mt].x, al[mt].y, al[mt].z, al[mt].w};
                        const uint32_t hi[4] = {ah[mt].x, ah[mt].y, ah[mt].z, ah[mt].w};
#pragma unroll
                        for (int nt = 0; nt < 4; ++nt) mma_u8s8(c[mt][nt], lo[2 * m], hi[2 * m], lo[2 * m + 1], hi[2 * m + 1], q0[nt], q1[nt]);
                    }
                }
            }
#endif
            uint32_t row_id[2][2];
#pragma unroll
            for (int mt = 0; mt < 2; ++mt) row_id[mt][0] = sh.rows[stage][16 * mt + g], row_id[mt][1] = sh.rows[stage][16 * mt + g + 8];
            __syncwarp();
            /* order this warp's generic-proxy reads of the stage before the async-proxy writes that refill it */
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (lane == 0) *reinterpret_cast<volatile uint32_t*>(&sh.gen[stage]) = n / kHeadStages + 1;
            if (refill < my_tiles) issue_fill(refill, refill_row);

            head_tail(net, b, narrow_w2, c, row_id, sh.w2, out, lane);
        }
        cur = seg_end;
    }
}

/* ------------------------------------------------------------------ eval post-processing */

/* adjustStatic + adjustEval, src/eval/eval.cpp:25-67.  All arithmetic is the reference's: int32, C++
 * division (truncating), std::clamp to +-(kScoreWin - 1). */
__global__ void adjust_kernel(const SpPackedBoard* __restrict__ boards, const int32_t* __restrict__ raw,
                              const int32_t* __restrict__ correction, size_t n, SpAdjustParams p, int32_t* __restrict__ out) {
    constexpr int kLimit = 25000 - 1; /* kScoreWin - 1, src/core.h:708 */
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4* rec = reinterpret_cast<const uint4*>(boards + i);
        const uint4 lo = __ldg(rec), hi = __ldg(rec + 1);
        const int n_pieces = __popcll(static_cast<uint64_t>(lo.y) << 32 | lo.x);
        const int stm = (hi.z & 0x80) ? kBlack : kWhite;
        const int halfmove = (hi.z >> 8) & 0xFF;
        const uint32_t words[4] = {lo.z, lo.w, hi.x, hi.y};
        int material = 0;
        for (int k = 0; k < n_pieces && k < 32; ++k) {
            uint32_t type = (words[k >> 3] >> ((k & 7) * 4)) & 7;
            if (type == 6) type = kRook; /* rook with castling rights */
            if (type < kKing) material += p.scaling_value[type];
        }
        int eval = raw[i];
        if (eval == INT32_MIN) { /* rejected record: keep the marker */
            out[i] = eval;
            continue;
        }
        eval = min(max(eval + p.contempt[stm], -kLimit), kLimit);                                 /* adjustStatic */
        eval = (eval * (p.material_scaling_base + material)
                + p.optimism[stm] * (p.optimism_base + material * p.optimism_material_scale / 1024))
             / 32768;
        eval = eval * (200 - halfmove) / 200;
        if (correction) eval += correction[i] / 2048;
        out[i] = min(max(eval, -kLimit), kLimit);
    }
}

/* wdl::normalizeScore<false> and wdl::wdlModel for a batch, src/wdl.cpp:28-80.  The material is
 * Position::classicalMaterial (position.h:515-523) counted from the record's nibbles.  Every double operation is
 * an explicitly rounded one (no contraction into FMAs), so the normalised score is the reference's bit for bit;
 * the win / loss per-mille figures go through exp(), whose last bit may differ between libm and the device
 * (tests allow one per mille). */
__global__ void wdl_kernel(const SpPackedBoard* __restrict__ boards, const int32_t* __restrict__ scores, size_t n,
                           int32_t* __restrict__ normalized, int32_t* __restrict__ win, int32_t* __restrict__ loss) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const uint4* rec = reinterpret_cast<const uint4*>(boards + i);
        const uint4 lo = __ldg(rec), hi = __ldg(rec + 1);
        const int n_pieces = __popcll(static_cast<uint64_t>(lo.y) << 32 | lo.x);
        const uint32_t words[4] = {lo.z, lo.w, hi.x, hi.y};
        int material = 0;
        for (int k = 0; k < n_pieces && k < 32; ++k) {
            const uint32_t type = (words[k >> 3] >> ((k & 7) * 4)) & 7;
            material += (0x5095331 >> (4 * type)) & 0xF; /* P 1, N 3, B 3, R 5, Q 9, K 0, castling rook (6) 5 */
        }
        const int32_t score = scores[i];
        const double m = static_cast<double>(min(max(material, 17), 78)) / 58.0;
        auto cubic = [m](double c0, double c1, double c2, double c3) {
            double v = __dadd_rn(__dmul_rn(c0, m), c1);
            v = __dadd_rn(__dmul_rn(v, m), c2);
            return __dadd_rn(__dmul_rn(v, m), c3);
        };
        const double a = cubic(-244.97139595, 687.39969858, -654.38002091, 608.47087786);
        if (normalized) {
            const bool keep = score == 0 || abs(score) > 25000; /* zero or decisive: core.h:722-724 */
            normalized[i] = keep ? score : static_cast<int32_t>(round(__dmul_rn(100.0, __ddiv_rn(static_cast<double>(score), a))));
        }
        if (win && loss) {
            const double b = cubic(68.24072080, -111.17718819, 74.50316570, 71.16566713);
            const double x = static_cast<double>(score);
            win[i] = static_cast<int32_t>(round(__ddiv_rn(1000.0, __dadd_rn(1.0, exp(__ddiv_rn(__dsub_rn(a, x), b))))));
            loss[i] = static_cast<int32_t>(round(__ddiv_rn(1000.0, __dadd_rn(1.0, exp(__ddiv_rn(__dadd_rn(a, x), b))))));
        }
    }
}

#include "ft_group.inc"
#include "small_batch.inc"
#include "head_umma.inc"

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
    ft_full_kernel<<<grid_for(n, kWarpsPerCta, sm_count, SP_FULL_MIN_BLOCKS), kThreads, 0, stream>>>(net, boards, n, act, bucket, status);
}

void launch_small_batch(const DeviceNet& net, SlotStore slots, const SmallBatchArgs& a, int sm_count, cudaStream_t stream) {
    const uint32_t total = a.n_refresh + a.n_update + a.n_eval;
    if (!total) return;
    SmallBatch sb{a.boards, a.dst, a.src, a.stm, a.out, a.error, a.item_error, a.n_refresh, a.n_update, a.n_eval, a.want};
    small_batch_kernel<<<grid_for(total, kWarpsPerCta, sm_count, 2), kThreads, 0, stream>>>(net, slots, sb);
}

size_t ft_group_scratch_words(size_t n_positions) { return 1 + (n_positions + kGroupN - 1) / kGroupN; }

cudaError_t launch_ft_group(
    const DeviceNet& net, const SpPackedBoard* boards, size_t n, uint8_t* act, uint8_t* bucket, DeviceStatus* status, uint32_t* overflow,
    int sm_count, cudaStream_t stream) {
    if (!n) return cudaSuccess;
    constexpr int kSmem = static_cast<int>(sizeof(GroupShared)) + 1024;
    static std::atomic<uint64_t> configured{0};
    int device = 0;
    cudaGetDevice(&device);
    if (!((configured.load(std::memory_order_relaxed) >> (device & 63)) & 1)) {
        const cudaError_t e = cudaFuncSetAttribute(ft_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
        if (e != cudaSuccess) return e;
        configured.fetch_or(uint64_t{1} << (device & 63), std::memory_order_relaxed);
    }
    cudaMemsetAsync(overflow, 0, sizeof(uint32_t), stream);
    const size_t n_groups = (n + kGroupN - 1) / kGroupN;
    static const int ctas_per_sm = [] { /* experiments: SP_NNUE_GROUP_CTAS=1 leaves every CTA an SM of its own */
        const char* v = std::getenv("SP_NNUE_GROUP_CTAS");
        return v && std::atoi(v) == 1 ? 1 : 2;
    }();
    const unsigned grid = static_cast<unsigned>(std::min<size_t>(n_groups, static_cast<size_t>(sm_count) * ctas_per_sm));
    /* SP_NNUE_GROUP_LIMIT=<threat rows>: a smaller union bound, so that tests can drive groups down the overflow path */
    static const uint32_t thr_limit = [] {
        const char* v = std::getenv("SP_NNUE_GROUP_LIMIT");
        return v ? static_cast<uint32_t>(std::min<long>(std::max<long>(std::atol(v), 1), kThrSetLimit)) : static_cast<uint32_t>(kThrSetLimit);
    }();
    ft_group_kernel<<<grid, kGroupThreads, kSmem, stream>>>(net, boards, n, act, bucket, status, overflow, kPsqSetLimit, thr_limit);
    /* groups whose row union did not fit the shared-memory sets: the per-position kernel (normally none: it returns at once) */
    ft_full_kernel<<<sm_count * SP_FULL_MIN_BLOCKS, kThreads, 0, stream>>>(net, boards, n, act, bucket, status, overflow);
    return cudaPeekAtLastError();
}

size_t row_list_bytes(size_t n_positions) { return n_positions * 2 * sizeof(RowListRecord); }

void launch_ft_full_split(
    const DeviceNet& net, const SpPackedBoard* boards, size_t n, void* row_lists, uint8_t* act, uint8_t* bucket,
    DeviceStatus* status, int sm_count, cudaStream_t stream) {
    if (!n) return;
    RowListRecord* records = static_cast<RowListRecord*>(row_lists);
    extract_kernel<<<grid_for(n, kWarpsPerCta, sm_count, 4), kThreads, 0, stream>>>(net, boards, n, records, status);
    accumulate_kernel<<<grid_for(2 * n, kAccWarps, sm_count, SP_ACC_MIN_BLOCKS), kAccWarps * 32, 0, stream>>>(net, records, 2 * n, act, bucket);
}

void launch_extract(const DeviceNet& net, const SpPackedBoard* boards, size_t n, void* row_lists, DeviceStatus* status, int sm_count, cudaStream_t stream) {
    if (!n) return;
    extract_kernel<<<grid_for(n, kWarpsPerCta, sm_count, 4), kThreads, 0, stream>>>(net, boards, n, static_cast<RowListRecord*>(row_lists), status);
}

void launch_accumulate(const DeviceNet& net, const void* row_lists, size_t n, uint8_t* act, uint8_t* bucket, int sm_count, cudaStream_t stream) {
    if (!n) return;
    accumulate_kernel<<<grid_for(2 * n, kAccWarps, sm_count, SP_ACC_MIN_BLOCKS), kAccWarps * 32, 0, stream>>>(
        net, static_cast<const RowListRecord*>(row_lists), 2 * n, act, bucket);
}

void launch_ft_slots(
    const DeviceNet& net, SlotStore slots, const uint32_t* src, const uint32_t* dst, const SpPackedBoard* boards,
    size_t n, uint8_t* act, uint8_t* bucket, DeviceStatus* status, int sm_count, cudaStream_t stream) {
    if (!n) return;
    ft_slots_kernel<<<grid_for(n, kWarpsPerCta, sm_count, SP_SLOTS_MIN_BLOCKS), kThreads, 0, stream>>>(net, slots, src, dst, boards, n, act, bucket, status);
}

void launch_plan_rebuilds(
    const DeviceNet& net, RebuildPlan plan, const SpPackedBoard* boards, const uint32_t* game_start, uint32_t n_games, size_t n_boards,
    int sm_count, cudaStream_t stream) {
    if (!n_games) return;
    const unsigned grid = static_cast<unsigned>(std::min<size_t>((n_games + 7) / 8, static_cast<size_t>(sm_count) * 8));
    plan_rebuilds_kernel<<<grid, 256, 0, stream>>>(net, plan, boards, game_start, n_games, n_boards);
}

void launch_run_rebuilds(
    const DeviceNet& net, RebuildPlan plan, const SpPackedBoard* boards, DeviceStatus* status, int sm_count, cudaStream_t stream) {
    run_rebuilds_kernel<<<sm_count * 4, kThreads, 0, stream>>>(net, plan, boards, status);
    rebuilds_done_kernel<<<1, 1, 0, stream>>>(plan);
}

void launch_ft_games(
    const DeviceNet& net, const SpPackedBoard* boards, const uint32_t* game_start, uint32_t n_games, size_t n_boards, uint8_t* act,
    uint8_t* bucket, RebuildPlan plan, DeviceStatus* status, int sm_count, cudaStream_t stream) {
    if (!n_games) return;
    ft_games_kernel<<<grid_for(n_games, kWarpsPerCta, sm_count, SP_GAMES_MIN_BLOCKS), kThreads, 0, stream>>>(
        net, boards, game_start, n_games, n_boards, act, bucket, plan, status);
}

void launch_slot_activate(
    SlotStore slots, const uint32_t* slot_ids, const uint8_t* stm, size_t n, uint8_t* act, uint8_t* bucket,
    DeviceStatus* status, int sm_count, cudaStream_t stream) {
    if (!n) return;
    slot_activate_kernel<<<grid_for(n, kWarpsPerCta, sm_count, 4), kThreads, 0, stream>>>(slots, slot_ids, stm, n, act, bucket, status);
}

/* SP_NNUE_HEAD: umma (default) = head_umma_kernel (L1 on tcgen05), stream = head_stream_kernel (L1 by mma.sync from bulk-copied
 * rows), tiles = the first head_kernel; the older two are kept for A/B measurements and as cross-checks in the tests. */
int head_variant_from_env() {
    const char* v = std::getenv("SP_NNUE_HEAD");
    if (v && std::strcmp(v, "tiles") == 0) return 2;
    if (v && std::strcmp(v, "stream") == 0) return 1;
    return 0;
}
/* SP_NNUE_HEAD_DIRECT: up to this many positions take the one-launch warp-per-position kernel (0 = never) */
uint32_t head_direct_max_from_env() {
    const char* v = std::getenv("SP_NNUE_HEAD_DIRECT");
    return v ? static_cast<uint32_t>(std::strtoul(v, nullptr, 10)) : 2048u;
}

int head_kernel_launches(size_t n, const HeadSort& sort) { return n <= sort.direct_max ? 1 : (n <= kHeadSortSmall ? 2 : 3); }

cudaError_t launch_head(
    const DeviceNet& net, const uint8_t* act, const uint8_t* bucket, size_t n, int32_t* out, const uint32_t* range, HeadSort sort,
    int sm_count, cudaStream_t stream, uint32_t range_len) {
    if (!n) return cudaSuccess;
    if (n <= sort.direct_max) {
        if (sort.ev_main_begin) cudaEventRecord(sort.ev_main_begin, stream);
        head_direct_kernel<<<grid_for(n, kWarpsPerCta, sm_count, 2), kThreads, 0, stream>>>(net, act, bucket, n, range, range_len, out);
        if (sort.ev_main_end) cudaEventRecord(sort.ev_main_end, stream);
        return cudaPeekAtLastError();
    }
    const size_t slots = std::min(sort.capacity, n + kHeadGroupPad * SP_OUTPUT_BUCKETS); /* n bounds the rows of this launch */
    if (n <= kHeadSortSmall) {
        head_sort_small_kernel<<<1, 1024, 0, stream>>>(bucket, n, range, range_len, sort, out);
    } else {
        cudaMemsetAsync(sort.counters, 0, kHeadSortCounters * sizeof(uint32_t), stream);
        const unsigned sort_grid = static_cast<unsigned>(std::min<size_t>((n + 4095) / 4096, static_cast<size_t>(sm_count) * 4)); /* 8 warps x 512 rows */
        head_hist_kernel<<<sort_grid, 256, 0, stream>>>(bucket, n, range, range_len, sort);
        head_scatter_kernel<<<sort_grid, 256, 0, stream>>>(bucket, n, range, range_len, sort, out);
    }
    /* opt in to > 48 KB of dynamic shared memory: a per-device attribute, set once per device */
    static std::atomic<uint64_t> configured{0};
    int device = 0;
    cudaGetDevice(&device);
    if (!((configured.load(std::memory_order_relaxed) >> (device & 63)) & 1)) {
        /* fails with "no kernel image" on a device this sm_100a-only library cannot run on: report it here, not at the launch */
        cudaError_t e = cudaFuncSetAttribute(head_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(HeadShared)));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(head_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(HeadStreamShared)));
        if (e == cudaSuccess) e = cudaFuncSetAttribute(head_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(HeadUmmaShared) + 1024));
        if (e != cudaSuccess) return e;
        configured.fetch_or(uint64_t{1} << (device & 63), std::memory_order_relaxed);
    }
    if (sort.ev_main_begin) cudaEventRecord(sort.ev_main_begin, stream);
    if (sort.variant == 2) {
        const unsigned grid = static_cast<unsigned>((slots + kHeadRows - 1) / kHeadRows);
        head_kernel<<<grid, kHeadWarps * 32, sizeof(HeadShared), stream>>>(net, act, bucket, out, sort);
    } else if (sort.variant == 0) {
        const unsigned grid = static_cast<unsigned>(std::min<size_t>((slots + kUmmaRows - 1) / kUmmaRows, static_cast<size_t>(sm_count)));
        head_umma_kernel<<<grid, kUmmaThreads, sizeof(HeadUmmaShared) + 1024, stream>>>(net, act, out, sort);
    } else {
        const unsigned grid = static_cast<unsigned>(std::min<size_t>((slots + kTileRows - 1) / kTileRows, static_cast<size_t>(sm_count)));
        head_stream_kernel<<<grid, kStreamThreads, sizeof(HeadStreamShared), stream>>>(net, act, out, sort);
    }
    if (sort.ev_main_end) cudaEventRecord(sort.ev_main_end, stream);
    return cudaPeekAtLastError();
}

void launch_adjust(
    const SpPackedBoard* boards, const int32_t* raw, const int32_t* correction, size_t n, const SpAdjustParams& params,
    int32_t* out, int sm_count, cudaStream_t stream) {
    if (!n) return;
    const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(sm_count) * 16));
    adjust_kernel<<<grid, 256, 0, stream>>>(boards, raw, correction, n, params, out);
}

void launch_wdl(
    const SpPackedBoard* boards, const int32_t* scores, size_t n, int32_t* normalized, int32_t* win, int32_t* loss, int sm_count,
    cudaStream_t stream) {
    if (!n) return;
    const unsigned grid = static_cast<unsigned>(std::min<size_t>((n + 255) / 256, static_cast<size_t>(sm_count) * 16));
    wdl_kernel<<<grid, 256, 0, stream>>>(boards, scores, n, normalized, win, loss);
}

} // namespace sp::gpu
