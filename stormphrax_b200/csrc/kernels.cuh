/*
 * kernels.cuh -- launch interface between the C-ABI layer (capi.cu) and the sm_100a kernels
 * (kernels.cu).  Device memory layouts are documented in DESIGN.md section 3.
 */
#ifndef SP_KERNELS_CUH
#define SP_KERNELS_CUH

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sp_nnue.h"
#include "sp_features.h"

namespace sp::gpu {

/* Row counts of the device feature-transformer tables (extra rows are synthetic). */
constexpr int kPsqBiasRow = SP_PSQ_FEATURES;        /* FT bias stored as one more PSQ row */
constexpr int kPsqZeroRow = SP_PSQ_FEATURES + 1;    /* all-zero row: pads index lists to x4 */
constexpr int kPsqRows = SP_PSQ_FEATURES + 2;
constexpr int kThrZeroRow = SP_THREAT_FEATURES;     /* "zero" row (0x80 bytes in the biased table) */
constexpr int kThrRows = SP_THREAT_FEATURES + 1;

constexpr int kSlotBytes = 2 * SP_L1_SIZE * 2;      /* two perspectives of int16[1024] */
constexpr int kSlotVec = kSlotBytes / 16;           /* uint4 per slot */

struct DeviceNet {
    const uint4* psq;   /* [kPsqRows][128]: int16 rows in lane order (see lane_order_element) */
    const uint4* thr;   /* [kThrRows][64] : int8 rows + 128 (stored unsigned), natural order */
    const uint8_t* psq_planes; /* [kPsqRows][2][1024]: the same int16 rows as two byte planes (low bytes, high bytes), natural
                                  column order: tensor-core operands of ft_group_kernel (u8 x u8 -> s32, recombined mod 2^16) */
    const int8_t* l1_w; /* [8][256][32][4] reference order (multilayer.h:180-196) */
    const int32_t* l1_b;
    const int32_t* l2_w; /* [8][64][64] */
    const uint32_t* l2_limbs; /* [8][4 limbs][8 n-tiles][16 k-quads][8]: byte limbs of l2_w as IMMA B fragments (l2_limb_index) */
    const uint32_t* l2_frags; /* [8][4 limbs][8 n-tiles][2 k-halves][32 lanes][2]: the same limbs for head_stream_kernel (l2_fragment_index) */
    const int32_t* l2_b;
    const int32_t* l3_w; /* [8][64] */
    const int32_t* l3_b;
    const FeatureTables* tables;
    uint32_t l2_narrow; /* bit b: every L2 weight of bucket b fits int16 (head_stream_kernel then needs 2 weight limbs, not 4) */
};

struct SlotStore {
    uint4* acc;            /* [n_slots][2][4][32] uint4 : lane-order accumulators (psq + threat, wrapped) */
    SpPackedBoard* boards; /* [n_slots] */
    uint32_t n_slots;
};

enum : int {
    kErrBadBoard = 1,
    kErrCapacity = 2,
    kErrBadSlot = 4,
};

/* Device-resident status block. */
struct DeviceStatus {
    int error; /* OR of kErr* bits */
    int pad;
    unsigned long long counters[8]; /* zero between launches: the work queues of ft_slots_kernel ([kSlotsTicket], + 1) and ft_group_kernel ([kGroupTicket], + 1) */
};
constexpr int kSlotsTicket = 2, kGroupTicket = 4; /* each + 1: warps / CTAs that have run dry */

/* Logical element index held at (chunk k, lane l, int16 slot e) of a device PSQ row / stored
 * accumulator.  Lane l owns logical elements [16 l, 16 l + 16) ("A half") and
 * [512 + 16 l, 512 + 16 l + 16) ("D half"): both members of the 16 activation pairs (i, i + 512) it
 * multiplies (multilayer.h:118-135).  Inside a lane the int16 are paired the way the even / odd
 * bytes of a 4-byte threat-row word unpack: chunk k = 0: A even, 1: A odd, 2: D even, 3: D odd;
 * word t = e / 2 of the chunk holds elements 4t + (k & 1) (low half) and 4t + (k & 1) + 2 (high). */
inline int lane_order_element(int k, int lane, int e) {
    return (k >= 2 ? 512 : 0) + lane * 16 + 4 * (e >> 1) + (k & 1) + 2 * (e & 1);
}

/* Word index, inside one bucket's 4096-word block, of the B-fragment word that holds byte limb `limb` of
 * the L2 weights W2[k][o] for k = 4 kq .. 4 kq + 3 (one byte each, k ascending).  Within an n-tile the 32
 * lanes (k-quad t, column g) of one IMMA read 32 consecutive words: no bank conflicts. */
inline int l2_limb_index(int limb, int kq, int o) { return ((limb * 8 + (o >> 3)) * 16 + kq) * 8 + (o & 7); }

/* The same limbs in the order head_stream_kernel wants them.  Its L2 contraction visits the inputs in the
 * order in which the L1 IMMAs leave their outputs in a lane: lane (g, t) of the warp holds outputs
 * 8t .. 8t + 7 of its rows, so k-slot 4t + m of a 32-wide k-half is input 8t + m and k-slot 16 + 4t + m is
 * input 8t + 4 + m (m = 0..3).  Word index, inside one bucket's 4096-word block, of the B-fragment register
 * `reg` (0: k-slots 4t.., 1: k-slots 16 + 4t..) of lane `lane` for (limb, n-tile nt, k-half ks); byte m of
 * that word is limb `limb` of W2[32 ks + 8t + 4 reg + m][8 nt + g]. */
inline int l2_fragment_index(int limb, int nt, int ks, int lane, int reg) { return ((((limb * 8 + nt) * 2 + ks) * 32) + lane) * 2 + reg; }

/* boards[i] -> act[i][1024], bucket[i]; every position rebuilt from scratch */
void launch_ft_full(
    const DeviceNet& net, const SpPackedBoard* boards, size_t n, uint8_t* act, uint8_t* bucket, DeviceStatus* status,
    int sm_count, cudaStream_t stream);

/* The same on the tensor cores (ft_group.inc): groups of 16 positions, every weight row of a group's union fetched once,
 * summed by tcgen05.mma.kind::i8.  `overflow` = ft_group_scratch_words(n) uint32 of scratch (groups handed to ft_full_kernel). */
size_t ft_group_scratch_words(size_t n_positions);
cudaError_t launch_ft_group(
    const DeviceNet& net, const SpPackedBoard* boards, size_t n, uint8_t* act, uint8_t* bucket, DeviceStatus* status, uint32_t* overflow,
    int sm_count, cudaStream_t stream);

/* The same in two kernels: boards -> row lists (row_list_bytes(n) of scratch) -> activations. */
size_t row_list_bytes(size_t n_positions);
void launch_extract(
    const DeviceNet& net, const SpPackedBoard* boards, size_t n, void* row_lists, DeviceStatus* status, int sm_count,
    cudaStream_t stream);
void launch_accumulate(
    const DeviceNet& net, const void* row_lists, size_t n, uint8_t* act, uint8_t* bucket, int sm_count, cudaStream_t stream);

/* slot refresh / update. src == nullptr: every dst slot is rebuilt from boards[i].
 * act may be nullptr (no evaluation wanted). */
void launch_ft_slots(
    const DeviceNet& net, SlotStore slots, const uint32_t* src, const uint32_t* dst, const SpPackedBoard* boards,
    size_t n, uint8_t* act, uint8_t* bucket, DeviceStatus* status, int sm_count, cudaStream_t stream);

/* Rebuilds taken out of the playout walker's loop.  A perspective must be rebuilt from scratch at the
 * first board of a game and whenever its king changes input bucket or board half (about 6 % of
 * perspective-plies in random playouts); done inside the walker those rare, long steps cost a quarter
 * of its time.  Instead: launch_plan_rebuilds finds them (king squares only), launch_run_rebuilds
 * computes their accumulators at full-refresh efficiency into `acc`, and the walker just loads them.
 * When `acc` is full the remaining ones simply stay with the walker (slot = kNoRebuildSlot). */
constexpr uint32_t kNoRebuildSlot = 0xFFFFFFFFu;
struct RebuildPlan {
    uint32_t* slot;     /* [2 * n_boards]: index into acc for (board, perspective), or kNoRebuildSlot */
    uint32_t* items;    /* [capacity]: board * 2 + perspective, in reservation order */
    uint4* acc;         /* [capacity][4][32]: rebuilt accumulators, lane order */
    uint32_t* counters; /* [0] = items reserved so far (may run past capacity), [1] = items already computed */
    uint32_t capacity;
};
void launch_plan_rebuilds(
    const DeviceNet& net, RebuildPlan plan, const SpPackedBoard* boards, const uint32_t* game_start, uint32_t n_games,
    size_t n_boards, int sm_count, cudaStream_t stream);
void launch_run_rebuilds(
    const DeviceNet& net, RebuildPlan plan, const SpPackedBoard* boards, DeviceStatus* status, int sm_count, cudaStream_t stream);

/* one warp walks one game; act/bucket rows are indexed like boards.  plan.slot == nullptr: every
 * rebuild happens inside the walker. */
void launch_ft_games(
    const DeviceNet& net, const SpPackedBoard* boards, const uint32_t* game_start, uint32_t n_games, size_t n_boards,
    uint8_t* act, uint8_t* bucket, RebuildPlan plan, DeviceStatus* status, int sm_count, cudaStream_t stream);

/* A search-sized round in ONE launch (small_batch.inc): refresh items [0, n_refresh), update items, evaluate-only items; a warp
 * takes an item from the board record to the evaluation.  All pointers are device pointers (one staged block). */
struct SmallBatchArgs {
    const SpPackedBoard* boards;
    const uint32_t* dst;
    const uint32_t* src;
    const uint8_t* stm;
    int32_t* out;
    int* error;
    uint8_t* item_error;
    uint32_t n_refresh, n_update, n_eval;
    uint32_t want; /* bit 0: evaluate refresh items, bit 1: evaluate update items */
};
void launch_small_batch(const DeviceNet& net, SlotStore slots, const SmallBatchArgs& args, int sm_count, cudaStream_t stream);

/* slots[i] -> act[i], bucket[i]; stm may be nullptr (use the stored board's side to move) */
void launch_slot_activate(
    SlotStore slots, const uint32_t* slot_ids, const uint8_t* stm, size_t n, uint8_t* act, uint8_t* bucket,
    DeviceStatus* status, int sm_count, cudaStream_t stream);

/* Scratch of the dense head: the launch's positions grouped by output bucket (counting sort), so that
 * every CTA works on rows of one bucket with that bucket's weights staged in shared memory once. */
struct HeadSort {
    uint32_t* order;    /* [capacity]: position indices grouped by bucket, each group padded to kHeadGroupPad with kHeadNoRow */
    uint32_t* counters; /* [0..7] rows per bucket, [8..15] scatter cursors, [16..24] group starts (24 = padded total) */
    size_t capacity;    /* entries in `order`: at least n + kHeadGroupPad * 8 */
    int variant = 0;    /* SP_NNUE_HEAD at sp_nnue_create: 0 umma (default), 1 stream, 2 tiles */
    uint32_t direct_max = 2048; /* SP_NNUE_HEAD_DIRECT: launches of up to this many positions skip the sort (head_direct_kernel) */
    cudaEvent_t ev_main_begin = nullptr, ev_main_end = nullptr; /* when set: recorded around the head kernel proper (profiling) */
};
/* kernels launch_head starts for n positions: 1 (head_direct_kernel), 2 (single-block sort + head) or 3 (histogram, scatter, head) */
int head_kernel_launches(size_t n, const HeadSort& sort);
int head_variant_from_env();
uint32_t head_direct_max_from_env();
constexpr uint32_t kHeadNoRow = 0xFFFFFFFFu;
constexpr uint32_t kHeadGroupPad = 128; /* rows of a head tile (head_umma_kernel: M of the MMA): groups are padded to whole tiles */
constexpr int kHeadSortCounters = 32;

/* act[i], bucket[i] -> out[i] : L1 (int8 IMMA) + L2 (byte-limb IMMA) + L3 + scale.  bucket[i] > 7 marks
 * a position whose board was rejected: out[i] = INT32_MIN.
 * range == nullptr: positions [0, n).  Otherwise positions [range[0], range[range_len]) read on the
 * device (a span of a game_start array) and n is only an upper bound of their count for the grids. */
cudaError_t launch_head(
    const DeviceNet& net, const uint8_t* act, const uint8_t* bucket, size_t n, int32_t* out, const uint32_t* range,
    HeadSort sort, int sm_count, cudaStream_t stream, uint32_t range_len = 0);

/* raw network outputs -> adjusted static evals (eval.cpp:25-67); one thread per position.
 * `correction` may be null. */
void launch_adjust(
    const SpPackedBoard* boards, const int32_t* raw, const int32_t* correction, size_t n, const SpAdjustParams& params,
    int32_t* out, int sm_count, cudaStream_t stream);

/* scores -> wdl::normalizeScore<false> and / or wdl::wdlModel per mille (src/wdl.cpp:28-80); one thread per position.
 * `normalized` may be null; `win` and `loss` are both given or both null. */
void launch_wdl(
    const SpPackedBoard* boards, const int32_t* scores, size_t n, int32_t* normalized, int32_t* win, int32_t* loss, int sm_count,
    cudaStream_t stream);

} // namespace sp::gpu

#endif
