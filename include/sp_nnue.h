/*
 * sp_nnue.h -- C-ABI of the B200-native batched NNUE evaluator (libsp_nnue.so).
 *
 * The reference engine (Stormphrax 8.0.2) has no plugin/FFI layer: its evaluation is the C++
 * API of src/eval, one position per call, statically linked.  This header
 * is the boundary a replacement plugs in at.  Every entry point names the reference interface it
 * replaces (file:line, relative to the reference tree).  The C++ mirror of that API
 * (stormphrax_b200/csrc/host/nnue_state.h: eval::init, NnueState::push/pop/evaluate, ...) is a
 * thin host-side layer over these calls; INTEGRATION.md shows how an engine build binds to it.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++ or torch types cross this boundary
 *   - every call returns an SpStatus (0 = ok); nothing throws; the message of the last failure
 *     on a context is available from sp_nnue_last_error
 *   - "host" pointers are ordinary process memory (pageable is fine); "device" pointers are CUDA
 *     device memory on the context's GPU; `stream` is a cudaStream_t passed as void* (NULL = the
 *     context's own stream)
 *   - a context is thread-compatible, not thread-safe: like a reference NnueState
 *     (src/eval/nnue_state.h:85-116) it must be driven by one host thread at a time
 *   - evaluations are the reference's raw network output: int32, side-to-move relative, before
 *     contempt / clamping / material scaling (src/eval/nnue_state.cpp:598-610)
 *   - results are bit-exact with the reference CPU path by contract
 */
#ifndef SP_NNUE_H
#define SP_NNUE_H

#include "sp_types.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SpNnue SpNnue; /* opaque; one per GPU; owns the device copy of the network */

typedef enum SpStatus {
    SP_OK = 0,
    SP_ERR_INVALID = 1,     /* null pointer, bad size, slot out of range ... */
    SP_ERR_BAD_NETWORK = 2, /* header validation failed (src/eval/nnue.cpp:85-185) */
    SP_ERR_CUDA = 3,        /* a CUDA runtime call failed; see sp_nnue_last_error */
    SP_ERR_NO_DEVICE = 4,   /* no usable CUDA device: there is NO CPU fallback on this path */
    SP_ERR_BAD_BOARD = 5,   /* a position record is malformed (king count, piece codes, > 32 pieces) */
    SP_ERR_CAPACITY = 6     /* a per-position feature list exceeded the reference's own bound (256) */
} SpStatus;

/* ---------------------------------------------------------------- lifecycle
 * Replaces eval::init / eval::shutdown / eval::getNetwork / eval::isNetworkLoaded
 * (src/eval/nnue.h:38-45, src/eval/nnue.cpp:200-321).  `net_image` is a LOGICAL (un-permuted)
 * network file: 64-byte CBNF header (src/eval/header.h:38-52) + raw arrays
 * (src/eval/nnue/input.h:359-361, src/eval/nnue/arch/multilayer.h:492-496), or, when the header's
 * kZstdCompressed flag is set, one zstd frame of those arrays (nnue.cpp:226-247; decoded with the system's
 * libzstd.so.1).  The reference replicates the network per NUMA node (nnue.cpp:266-284); here it is one
 * copy per GPU. */
int sp_nnue_create(const void* net_image, size_t len, int device, SpNnue** out);
void sp_nnue_destroy(SpNnue* ctx);
const char* sp_nnue_last_error(const SpNnue* ctx); /* ctx may be NULL: error of the last failed create */
int sp_nnue_device(const SpNnue* ctx);
/* Wait for `stream` (NULL = the context's own) and report what the kernels flagged since the last
 * check: SP_ERR_BAD_BOARD, SP_ERR_CAPACITY, SP_ERR_INVALID (slot out of range) or SP_OK.  The
 * host-pointer entry points do this themselves; the *_device ones are asynchronous and do not. */
int sp_nnue_sync(SpNnue* ctx, void* stream);
/* Run the host-pointer entry points on the caller's stream (a cudaStream_t as void*) instead of
 * the context's private one, so that a caller can bracket them with its own events.
 * NULL restores a private stream. */
int sp_nnue_set_stream(SpNnue* ctx, void* stream);

/* ---------------------------------------------------------------- full refresh
 * Replaces NnueState::evaluateOnce / eval::staticEvalOnce for a batch
 * (src/eval/nnue_state.cpp:612-634, src/eval/eval.cpp:109-112): both accumulators are rebuilt
 * from scratch for every position.  Side to move and output bucket are taken from the record.
 * Host variant: copies boards H2D, runs, copies evals D2H, returns when `out` is filled. */
int sp_nnue_eval_full(SpNnue* ctx, const SpPackedBoard* boards, size_t n, int32_t* out);
int sp_nnue_eval_full_device(SpNnue* ctx, const SpPackedBoard* d_boards, size_t n, int32_t* d_out, void* stream);

/* ---------------------------------------------------------------- accumulator slots
 * Replaces the per-thread accumulator stack of NnueState (src/eval/nnue_state.h:68-116): a slot
 * holds the two perspective accumulators (int16[2][1024], PSQ and threat parts summed -- only
 * their wrapped sum reaches the network, src/eval/nnue/arch/multilayer.h:118-124) plus the board
 * they describe.  Slots live in device memory and are addressed by index.
 *
 *   sp_nnue_refresh      NnueState::reset (nnue_state.cpp:539-560): rebuild slots from boards
 *   sp_nnue_update       push + Position::applyMove<BoardObserver> + ensureUpToDate
 *                        (nnue_state.cpp:562-570, 636-697; delta generation nnue.cpp:490-599,
 *                        nnue_state.cpp:34-87, 163-307, 356-394): dst = src advanced to `after`.
 *                        The library derives add/sub feature lists on the GPU from the board stored
 *                        in src and the `after` record; when a king changes input bucket or board
 *                        half (psq.h:264-283, nnue_state.h:118-128) that perspective is rebuilt.
 *                        src == dst (in-place, the datagen applyImmediately form,
 *                        nnue_state.cpp:572-591) is allowed.
 *   sp_nnue_eval_slots   NnueState::evaluate (nnue_state.cpp:598-610).  stm[i] = 0 black, 1 white;
 *                        NULL = side to move of the stored board (explicit stm exists because the
 *                        engine evaluates null-move children on the parent's accumulators).
 *   sp_nnue_update_eval  update followed by eval of the dst slots, one H2D and one D2H.
 * All array arguments of the host variants are host pointers. */
int sp_nnue_slots_reserve(SpNnue* ctx, size_t n_slots);
int sp_nnue_refresh(SpNnue* ctx, const uint32_t* slots, const SpPackedBoard* boards, size_t n);
int sp_nnue_update(
    SpNnue* ctx, const uint32_t* src_slots, const uint32_t* dst_slots, const SpPackedBoard* after, size_t n);
int sp_nnue_eval_slots(SpNnue* ctx, const uint32_t* slots, const uint8_t* stm, size_t n, int32_t* out);
/* One round of a batched driver (many NnueStates, one evaluation each) in ONE submission and one wait:
 * a refresh group, an update group -- each also evaluated, with the side to move of its boards, when its
 * output array is given -- and an evaluate-only group (stm as in sp_nnue_eval_slots).  No item may read or
 * write a slot that another item of the same call writes (different states; an in-place update src == dst is
 * one item).  Any group may be empty.  Up to 1,024 items run as ONE kernel launch (small_batch_kernel: slot
 * update, activation and the dense head per warp) between one host-to-device and one device-to-host copy --
 * the size of a round of a few hundred searches; SP_NNUE_SMALL=0 sends them through the general kernels. */
int sp_nnue_batch(
    SpNnue* ctx, const uint32_t* refresh_slots, const SpPackedBoard* refresh_boards, size_t n_refresh, int32_t* refresh_out,
    const uint32_t* src_slots, const uint32_t* dst_slots, const SpPackedBoard* after, size_t n_update, int32_t* update_out,
    const uint32_t* eval_slots, const uint8_t* stm, size_t n_eval, int32_t* eval_out);
int sp_nnue_update_eval(
    SpNnue* ctx,
    const uint32_t* src_slots,
    const uint32_t* dst_slots,
    const SpPackedBoard* after,
    size_t n,
    int32_t* out);
/* device-pointer variants (all arrays in device memory, asynchronous on `stream`) */
int sp_nnue_refresh_device(SpNnue* ctx, const uint32_t* d_slots, const SpPackedBoard* d_boards, size_t n, void* stream);
int sp_nnue_update_eval_device(
    SpNnue* ctx,
    const uint32_t* d_src_slots,
    const uint32_t* d_dst_slots,
    const SpPackedBoard* d_after,
    size_t n,
    int32_t* d_out, /* NULL = update only */
    void* stream);

/* ---------------------------------------------------------------- playout streams
 * The incremental workload of BASELINE.json config 3 and of src/datagen/datagen.cpp:206-262: game g
 * starts from starts[g] and plays moves[game_start[g] - g ... ) (one move fewer than positions);
 * out[game_start[g] + i] receives the evaluation after i moves (i = 0 is the start position), for
 * the side to move there.  Accumulators never leave the SM between plies of one game. */
int sp_nnue_eval_playouts(
    SpNnue* ctx,
    const SpPackedBoard* boards, /* host; every position of every game, as sp_host_playouts writes them */
    const uint32_t* game_start,  /* host; n_games + 1 offsets into boards/out */
    uint32_t n_games,
    int32_t* out);
int sp_nnue_eval_playouts_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, const uint32_t* d_game_start, uint32_t n_games, size_t n_boards,
    int32_t* d_out, void* stream);

/* ---------------------------------------------------------------- dense head in isolation
 * BASELINE.json config 4: L1 (int8 IMMA) -> L2 -> L3 from already-activated FT outputs.
 * Replaces PairwiseMultilayerCReLUSCReLUCReLU::propagateL1/L2/L3
 * (src/eval/nnue/arch/multilayer.h:154-490).  d_act is uint8[n][1024] (stm half first),
 * d_bucket uint8[n]. */
int sp_nnue_forward_device(
    SpNnue* ctx, const uint8_t* d_act, const uint8_t* d_bucket, size_t n, int32_t* d_out, void* stream);
/* FT only: boards -> activations + buckets (for tests and for feeding sp_nnue_forward_device) */
int sp_nnue_activations_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, size_t n, uint8_t* d_act, uint8_t* d_bucket, void* stream);

/* ---------------------------------------------------------------- eval post-processing
 * Replaces eval::staticEval's adjustStatic (src/eval/eval.cpp:25-28: contempt, clamp to +-24999) followed
 * by eval::adjustEval (src/eval/eval.cpp:31-67: material scaling, optimism, 50-move damping, optional
 * correction, clamp) for a batch.  `raw` are the network outputs the eval entry points return.
 * The reference reads its correction from per-thread history tables (src/correction.h); that state stays
 * with the caller, which passes the looked-up correction per position (NULL = adjustEval<false>). */
typedef struct SpAdjustParams {
    int32_t scaling_value[5];        /* pawn, knight, bishop, rook, queen: src/tunable.h:161-165 */
    int32_t material_scaling_base;   /* src/tunable.h:167 */
    int32_t optimism_base;           /* src/tunable.h:168 */
    int32_t optimism_material_scale; /* src/tunable.h:169 */
    int32_t contempt[2];             /* eval::Contempt, [black, white] (src/eval/eval.h:31) */
    int32_t optimism[2];             /* eval::Optimism, [black, white] (src/eval/eval.h:32) */
} SpAdjustParams;
void sp_nnue_adjust_defaults(SpAdjustParams* params); /* the reference's default tunables, zero contempt / optimism */
int sp_nnue_adjust(
    SpNnue* ctx, const SpPackedBoard* boards, const int32_t* raw, const int32_t* correction, size_t n,
    const SpAdjustParams* params, int32_t* out);
int sp_nnue_adjust_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, const int32_t* d_raw, const int32_t* d_correction, size_t n,
    const SpAdjustParams* params, int32_t* d_out, void* stream);

/* Score normalisation and the win / draw / loss model for a batch, on the device.
 * Replaces wdl::normalizeScore<false>(score, pos.classicalMaterial()) and wdl::wdlModel(povScore, material)
 * (src/wdl.cpp:28-80, src/position.h:515-523): what datagen's adjudication (src/datagen/datagen.cpp:224-253) and the
 * UCI score output apply to a search score.  `scores` are white-relative or side-to-move-relative as the caller's
 * use demands (the functions do not care).  normalized[i] is bit-exact with the reference; win[i] / loss[i] are per
 * mille and pass through exp(): the last bit of a device exp may differ from libm's (at most one per mille).
 * `normalized` may be NULL; `win` and `loss` are both given or both NULL. */
int sp_nnue_wdl(
    SpNnue* ctx, const SpPackedBoard* boards, const int32_t* scores, size_t n, int32_t* normalized, int32_t* win, int32_t* loss);
int sp_nnue_wdl_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, const int32_t* d_scores, size_t n, int32_t* d_normalized, int32_t* d_win,
    int32_t* d_loss, void* stream);

/* ---------------------------------------------------------------- introspection
 * Counters since create (uint64 each): what the engine keeps per thread in SearchData
 * (src/thread.h:35-79) and sums at report time.  Multi-GPU runs all-reduce these over NCCL. */
enum {
    SP_CTR_EVALS = 0,        /* positions evaluated */
    SP_CTR_FULL_REFRESH = 1, /* perspective accumulators rebuilt from scratch */
    SP_CTR_INCREMENTAL = 2,  /* perspective accumulators updated incrementally */
    SP_CTR_LAUNCHES = 3,     /* kernels launched by this context */
    SP_NUM_COUNTERS = 8
};
int sp_nnue_counters(SpNnue* ctx, uint64_t out[SP_NUM_COUNTERS]);
/* Per-kernel device time: when enabled, every launch made by this context is bracketed with CUDA
 * events on its stream; sp_nnue_profile_read waits for them, returns summed milliseconds and launch
 * counts per kernel class since the previous read, and resets the sums. */
enum {
    SP_KERNEL_FT_FULL = 0,  /* boards -> activations, both accumulators rebuilt */
    SP_KERNEL_HEAD = 1,     /* activations -> evals (L1 IMMA, L2, L3) */
    SP_KERNEL_FT_SLOTS = 2, /* slot refresh / incremental update */
    SP_KERNEL_FT_GAMES = 3, /* playout walker */
    SP_KERNEL_EXTRACT = 4,  /* boards -> row lists (split full refresh) */
    SP_KERNEL_ACCUMULATE = 5, /* row lists -> activations (split full refresh) */
    SP_KERNEL_REBUILDS = 6, /* playout walker: planning + computing the rebuilt accumulators ahead of the walk */
    SP_KERNEL_HEAD_MAIN = 7, /* sp_nnue_forward_device only: the head kernel proper, without the counting sort in front of it (SP_KERNEL_HEAD spans both) */
    SP_NUM_KERNEL_CLASSES = 8
};
int sp_nnue_profile(SpNnue* ctx, int enable);
int sp_nnue_profile_read(SpNnue* ctx, double ms[SP_NUM_KERNEL_CLASSES], uint64_t launches[SP_NUM_KERNEL_CLASSES]);
/* Debug/test access: accumulators of a slot in LOGICAL order, int16[2][1024] (black, white). */
int sp_nnue_read_slot(SpNnue* ctx, uint32_t slot, int16_t* out_acc, SpPackedBoard* out_board);

/* sp_nnue_batch with every array on the device (results too), enqueued on `stream` without waiting: the form a
 * GPU-resident driver uses (sp_selfplay_run_gpu).  Errors found on the device are reported by the next
 * sp_nnue_sync. */
int sp_nnue_batch_device(
    SpNnue* ctx, const uint32_t* d_refresh_slots, const SpPackedBoard* d_refresh_boards, size_t n_refresh, int32_t* d_refresh_out,
    const uint32_t* d_src_slots, const uint32_t* d_dst_slots, const SpPackedBoard* d_after, size_t n_update, int32_t* d_update_out,
    const uint32_t* d_eval_slots, const uint8_t* d_stm, size_t n_eval, int32_t* d_eval_out, void* stream);

/* ---------------------------------------------------------------- batched self-play (datagen)
 * Replaces the per-thread game loop of src/datagen/datagen.cpp:96-321 (`datagen::run`, :323-400): random
 * opening plies, search -> applyMove -> NnueState::applyImmediately -> (move, score) until the game is
 * decided or adjudicated, one viriformat record per game (src/datagen/viriformat.cpp:33-63).  Instead of
 * one game per thread there are `concurrency` game slots playing at once (split into contiguous ranges over
 * `threads` host threads, each with its own evaluator context on `device`); their searches are resumable
 * and every static evaluation they ask for is answered in device batches through the NnueState / EvalBatch
 * mirror.  The search itself is a stand-in (iterative-deepening alpha-beta, csrc/host/selfplay.h): the
 * reference's search is out of scope.  Slot g plays total_games / concurrency games one after the other (+ 1
 * for the first total_games % concurrency slots); every (slot, game number) has its own random stream.
 * `out` receives the records slot by slot (slot-major, game order within a slot), so the bytes do not depend
 * on `threads`; *out_len is always set to the bytes produced (SP_ERR_CAPACITY if they did not fit; out ==
 * NULL with out_capacity == 0 only counts). */
typedef struct SpSelfplayParams {
    uint32_t concurrency;    /* game slots = games in flight, all threads together */
    uint32_t total_games;    /* games to play, all slots together */
    uint32_t threads;        /* sp_selfplay_run: host threads; sp_selfplay_run_gpu: concurrent driver instances (0 = 1) */
    uint32_t depth;          /* iterative deepening stops after this depth ... */
    uint32_t nodes_per_move; /* ... or once a finished iteration has used this many nodes (datagen.cpp:76 soft limit) */
    uint32_t max_plies;      /* undecided games are drawn here (0 = 300; at most 510) */
    uint64_t seed;
    uint32_t dfrc;           /* 1: `datagen <fmt> dfrc`: every game starts from a random double-Fischer-random position */
    uint32_t reserved;       /* 0 */
} SpSelfplayParams;
typedef struct SpSelfplayStats {
    uint64_t games, positions, nodes, evals, batches, searches;
} SpSelfplayStats;
int sp_selfplay_run(
    const void* net_image, size_t len, int device, const SpSelfplayParams* params, SpSelfplayStats* stats, uint8_t* out,
    size_t out_capacity, size_t* out_len);
/* The same games played by a GPU-resident driver: one device thread per game slot runs the search state
 * machine (the board model, move generator, search and record writer of csrc/host/selfplay.h compile for the
 * device too); the host only reads four counters per round and submits one sp_nnue_batch_device.  `threads`
 * driver instances (slot ranges, own context and stream each) run concurrently.  Same slots, same random
 * streams, same record order: both entry points produce the same bytes. */
int sp_selfplay_run_gpu(
    const void* net_image, size_t len, int device, const SpSelfplayParams* params, SpSelfplayStats* stats, uint8_t* out,
    size_t out_capacity, size_t* out_len);

/* One game as a viriformat record through the driver's writer (src/datagen/viriformat.cpp:27-63): returns
 * the bytes written or -1.  outcome: 0 white loss, 1 draw, 2 white win (src/datagen/common.h:24-28). */
long sp_host_viriformat(
    const SpPackedBoard* start, const SpMove* moves, const int16_t* scores, uint32_t n, int outcome, uint8_t* out, size_t cap);
/* wdl::normalizeScore<false>(score, pos.classicalMaterial()) as the driver's adjudication uses it
 * (src/wdl.cpp:28-80, src/position.h:515-523). */
int sp_host_normalize_score(const SpPackedBoard* board, int32_t score, int32_t* material, int32_t* normalized);
/* wdl::wdlModel(povScore, pos.classicalMaterial()): win and loss per mille (src/wdl.cpp:43-50; UCI output) */
int sp_host_wdl_model(const SpPackedBoard* board, int32_t pov_score, int32_t* win, int32_t* loss);

/* ---------------------------------------------------------------- host utilities (no GPU)
 * Workload generation and CPU execution of the shared feature code, for tests and benchmarks. */
/* Random legal playouts from the standard start position: game g is seeded from (seed, g); all
 * positions including the start are written; moves[i] is the move played from boards[i] (0 at
 * the last position of a game); game_start has n_games + 1 entries.  boards/moves must hold
 * n_games * (max_plies + 1) records.  Returns the number of positions written. */
size_t sp_host_playouts(
    uint64_t seed,
    uint32_t n_games,
    uint32_t max_plies,
    int threads,
    SpPackedBoard* boards,
    SpMove* moves,
    uint32_t* game_start);
/* The logical payload (SP_NET_PAYLOAD_BYTES of raw arrays) of a network image whose header may carry the
 * kZstdCompressed flag: what sp_nnue_create uploads.  out may be NULL (size query).  Returns the payload size or -1. */
long sp_host_net_payload(const void* net_image, size_t len, void* out, size_t cap);
int sp_host_board_from_fen(const char* fen, SpPackedBoard* out);
int sp_host_board_to_fen(const SpPackedBoard* board, char* out, size_t cap);
/* Double Fischer random start position (Position::fromDfrcIndex, src/position.cpp:1215-1270): index = black * 960 + white */
int sp_host_board_from_dfrc(uint32_t index, SpPackedBoard* out);
int sp_host_legal_moves(const SpPackedBoard* board, SpMove* out /* [256] */);
int sp_host_in_check(const SpPackedBoard* board); /* 1 / 0, -1 for a malformed record */
/* adjustStatic + adjustEval through the C++ mirror's per-position form (csrc/host/nnue_state.h), on the host:
 * same arguments and results as sp_nnue_adjust */
int sp_host_adjust(const SpPackedBoard* boards, const int32_t* raw, const int32_t* correction, size_t n, const SpAdjustParams* params, int32_t* out);
int sp_host_apply_move(const SpPackedBoard* board, SpMove move, SpPackedBoard* out);
int sp_host_features(const SpPackedBoard* board, int perspective, int kind, uint32_t* out /* [512] */);
/* Workload statistics for the roofline: out[0] = PSQ rows, out[1] = threat rows, out[2] = pawn-pair
 * rows, summed over both perspectives of all n boards (what a full refresh must read). */
int sp_host_feature_counts(const SpPackedBoard* boards, size_t n, int threads, uint64_t out[3]);
/* The same for the incremental walker over a playout stream: out = {PSQ delta rows, threat delta rows,
 * PSQ rows of rebuilt perspectives, threat rows of rebuilt perspectives, updated perspectives,
 * rebuilt perspectives}, counted with the delta generator the kernels run (csrc/sp_delta.h). */
int sp_host_playout_stats(const SpPackedBoard* boards, const uint32_t* game_start, uint32_t n_games, int threads, uint64_t out[6]);
int sp_host_feature_delta(
    const SpPackedBoard* before,
    const SpPackedBoard* after,
    int perspective,
    uint32_t* psq_add, int* n_psq_add,
    uint32_t* psq_sub, int* n_psq_sub,
    uint32_t* thr_add, int* n_thr_add,
    uint32_t* thr_sub, int* n_thr_sub);

#ifdef __cplusplus
}
#endif

#endif /* SP_NNUE_H */
