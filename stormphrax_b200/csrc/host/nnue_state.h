/*
 * nnue_state.h -- host-side mirror of the reference's eval API, implemented over the C-ABI
 * (include/sp_nnue.h).  Same names, argument meaning and call protocol as
 *   eval::init / shutdown / isNetworkLoaded / getNetwork        src/eval/nnue.h:38-45
 *   eval::NnueState::{reset,push,pop,applyImmediately,evaluate,evaluateOnce}
 *                                                               src/eval/nnue_state.h:85-116
 *   eval::BoardObserver / UpdateContext                         src/eval/nnue_state.h:28-45
 *   eval::staticEval / staticEvalOnce / adjustStatic            src/eval/eval.{h,cpp}
 * so that search / datagen code written against the reference compiles against this header
 * with `Position` being anything that offers `SpPackedBoard pack() const` and `stm()`.
 *
 * What differs underneath: the accumulator stack lives in device memory (one slot per stack
 * level), and the observer records nothing -- the GPU derives add/sub feature lists from the
 * (stored predecessor board, new board) pair, however many plies apart they are
 * (csrc/sp_delta.h).  Laziness is preserved: push() is O(1) and touches no device memory;
 * work happens in evaluate(), from the nearest up-to-date ancestor (nnue_state.cpp:636-697).
 *
 * One NnueState per host thread, like the reference (thread.h:147).  For throughput, many
 * states share one EvalBatch: evaluateAsync() queues, flush() runs ONE device batch.
 */
#ifndef SP_HOST_NNUE_STATE_H
#define SP_HOST_NNUE_STATE_H

#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../../include/sp_nnue.h"
#include "position.h"

namespace sp::host::eval {

using i32 = int32_t;
using Color = int;

constexpr i32 kScoreWin = 25000; /* src/core.h:708 */

/* ---- lifecycle (src/eval/nnue.cpp:200-321).  One network copy per GPU instead of per NUMA node. */
bool init(const void* networkImage, size_t len, int device = 0);
bool initFromFile(const std::string& path, int device = 0);
void shutdown();
[[nodiscard]] bool isNetworkLoaded();
[[nodiscard]] SpNnue* getNetwork(int device = -1); /* -1: the device passed to init */
/* One more evaluator context (network copy, stream, slot store) on the device given to init, from the same image: what a
 * further host scheduler thread binds to.  The reference keeps one network copy per NUMA node and hands a thread the copy of
 * its node (nnue.cpp:267-277, 314-320; search.cpp:206); here the unit is the host thread, because a context's entry points
 * are serialised on its stream.  nullptr (and lastError()) on failure.  Contexts still alive at shutdown() are destroyed there. */
[[nodiscard]] SpNnue* createContext();
void destroyContext(SpNnue* ctx);
[[nodiscard]] const char* lastError();

/* The reference's UpdateContext carries the add/sub lists the observer collected
 * (nnue_state.h:28-31).  Here the device derives them from boards, so it is empty. */
struct UpdateContext {};

/* Same callback surface as eval::BoardObserver (nnue_state.h:33-45); every callback is a no-op. */
struct BoardObserver {
    UpdateContext& ctx;
    SP_POS_HD void prepareKingMove(Color, Square, Square) {}
    template <typename P> SP_POS_HD void pieceAdded(const P&, Piece, Square) {}
    template <typename P> SP_POS_HD void pieceRemoved(const P&, Piece, Square) {}
    template <typename P> SP_POS_HD void pieceMutated(const P&, Piece, Piece, Square) {}
    template <typename P> SP_POS_HD void pieceMoved(const P&, Piece, Square, Square) {}
    template <typename P> SP_POS_HD void piecePromoted(const P&, Piece, Square, Piece, Square) {}
    template <typename P> SP_POS_HD void finalize(const P&, const P&) {}
};

class NnueState;

/* Collects evaluations from many NnueStates (many concurrent games / search threads) and runs
 * them as one device batch.  enqueue is mutex-protected (MPSC), flush is called by one thread. */
class EvalBatch {
public:
    static constexpr uint32_t kNoUpdate = 0xFFFFFFFFu; /* srcSlot value: dstSlot is already up to date */

    explicit EvalBatch(SpNnue* network) : m_network{network} {}

    /* result is written to *out by flush().  A state may have ONE evaluation pending per flush.
     * out == nullptr: bring dstSlot up to date without evaluating it. */
    void enqueue(uint32_t srcSlot, uint32_t dstSlot, bool rebuild, const SpPackedBoard& board, Color stm, i32* out);
    [[nodiscard]] size_t pending() const { return m_out.size() + m_refresh.out.size() + m_update.out.size(); }
    /* Runs everything queued; returns an SpStatus. */
    int flush();

private:
    /* an item is refreshed or updated and evaluated in the same launch when the wanted side equals its board's
     * side to move (the normal case); otherwise (null-move style evaluation, or no update needed) it joins the
     * evaluate-only group, which runs last */
    struct Group {
        std::vector<uint32_t> src, dst;
        std::vector<SpPackedBoard> boards;
        std::vector<i32*> out; /* nullptr: result not wanted */
        std::vector<i32> results;
        void clear() { src.clear(), dst.clear(), boards.clear(), out.clear(); }
    };
    SpNnue* m_network;
    std::mutex m_mutex;
    Group m_refresh, m_update;
    std::vector<uint32_t> m_evalSlots;
    std::vector<uint8_t> m_stm;
    std::vector<i32*> m_out;
    std::vector<i32> m_results;
};

class NnueState {
public:
    static constexpr uint32_t kStackDepth = 256; /* nnue_state.h:88 */

    /* Slots [slotBase, slotBase + depth) of `network` belong to this state.  The reference's stack is 256
     * deep (kStackDepth); drivers that run tens of thousands of states with a bounded search depth pass
     * a smaller one (a slot is 4 KB of device memory). */
    NnueState() = default;
    explicit NnueState(SpNnue* network, uint32_t slotBase = 0, uint32_t depth = kStackDepth) { setNetwork(network, slotBase, depth); }

    void setNetwork(SpNnue* network, uint32_t slotBase = 0, uint32_t depth = kStackDepth);

    template <typename Position> void reset(const Position& pos) { resetPacked(pos.pack()); }

    BoardObserver push(); /* O(1): bumps the stack pointer and marks the level dirty */
    void pop();

    /* datagen form (datagen.cpp:257-260): the current level itself moves on to `pos` */
    template <typename Position> void applyImmediately(const UpdateContext&, const Position& pos) {
        applyPacked(pos.pack());
    }

    template <typename Position> [[nodiscard]] i32 evaluate(const Position& pos, Color stm) {
        return evaluatePacked(pos.pack(), stm);
    }
    /* queue instead of running: the result lands in *out when batch.flush() is called */
    template <typename Position> void evaluateAsync(EvalBatch& batch, const Position& pos, Color stm, i32* out) {
        evaluateAsyncPacked(batch, pos.pack(), stm, out);
    }

    /* applyImmediately for batched drivers: the current level moves on to a new position, but the device
     * work is left to the next evaluate / evaluateAsync at or below this level (the device derives the
     * delta from the board stored in the slot, however many plies behind it is). */
    void applyLazily() { m_stale[m_top] = 1; }
    /* reset for batched drivers: forget everything; the next evaluation rebuilds level 0 from scratch */
    void invalidate() {
        m_top = 0;
        m_clean.assign(m_clean.size(), 0);
        m_stale.assign(m_stale.size(), 0);
    }

    template <typename Position> [[nodiscard]] static i32 evaluateOnce(const Position& pos, Color stm) {
        return evaluateOncePacked(pos.pack(), stm);
    }

    void resetPacked(const SpPackedBoard& board);
    void applyPacked(const SpPackedBoard& board);
    [[nodiscard]] i32 evaluatePacked(const SpPackedBoard& board, Color stm);
    void evaluateAsyncPacked(EvalBatch& batch, const SpPackedBoard& board, Color stm, i32* out);
    [[nodiscard]] static i32 evaluateOncePacked(const SpPackedBoard& board, Color stm);

    [[nodiscard]] uint32_t depth() const { return m_top; }

private:
    [[nodiscard]] uint32_t slot(uint32_t level) const { return m_slotBase + level; }
    /* nearest level <= m_top whose slot is up to date, or -1 */
    [[nodiscard]] int cleanAncestor() const;

    SpNnue* m_network{nullptr};
    uint32_t m_slotBase{0};
    uint32_t m_top{0};
    UpdateContext m_ctx{};
    std::vector<uint8_t> m_clean = std::vector<uint8_t>(kStackDepth, 0); /* slot holds the accumulators of its stored board */
    std::vector<uint8_t> m_stale = std::vector<uint8_t>(kStackDepth, 0); /* ... but that board is an ancestor of the level's position */
};

/* ---- eval.h wrappers (src/eval/eval.cpp:25-28, 74-77, 109-112) */
struct Contempt {
    i32 value[2]{0, 0};
    SP_POS_HD i32 operator[](Color c) const { return value[c]; }
};

SP_POS_HD inline i32 adjustStatic(i32 eval, Color stm, const Contempt& contempt) {
    eval += contempt[stm];
    return eval < -kScoreWin + 1 ? -kScoreWin + 1 : (eval > kScoreWin - 1 ? kScoreWin - 1 : eval);
}

/* eval::adjustEval (src/eval/eval.cpp:31-67) for ONE position, the form the search calls per node: material
 * scaling, optimism, 50-move damping, correction, clamp.  `correction` is what the caller's
 * CorrectionHistoryTable::correction(pos, keyHistory) returned (0 = adjustEval<false>).  Batches go through
 * sp_nnue_adjust on the device; this is the same arithmetic on the host (checked against the same
 * reference-generated vectors, tests/test_host.py). */
struct Optimism {
    i32 value[2]{0, 0};
    i32 operator[](Color c) const { return value[c]; }
};

inline i32 adjustEvalPacked(const SpPackedBoard& board, const Optimism& optimism, i32 eval, i32 correction, const SpAdjustParams& tunables) {
    int material = 0, k = 0;
    for (uint64_t occ = board.occupancy; occ && k < 32; occ &= occ - 1, ++k) {
        unsigned type = (board.pieces[k / 2] >> ((k % 2) * 4)) & 7;
        if (type == 6) type = kRook; /* rook with castling rights */
        if (type < static_cast<unsigned>(kKing)) material += tunables.scaling_value[type];
    }
    const Color stm = (board.stm_ep & 0x80) ? kBlack : kWhite;
    eval = (eval * (tunables.material_scaling_base + material)
            + optimism[stm] * (tunables.optimism_base + material * tunables.optimism_material_scale / 1024))
         / 32768;
    eval = eval * (200 - static_cast<i32>(board.halfmove)) / 200;
    eval += correction / 2048;
    return eval < -kScoreWin + 1 ? -kScoreWin + 1 : (eval > kScoreWin - 1 ? kScoreWin - 1 : eval);
}

template <typename Position>
i32 adjustEval(const Position& pos, const Optimism& optimism, i32 eval, i32 correction = 0) {
    SpAdjustParams tunables;
    sp_nnue_adjust_defaults(&tunables);
    return adjustEvalPacked(pos.pack(), optimism, eval, correction, tunables);
}

template <typename Position> i32 staticEval(const Position& pos, NnueState& state, const Contempt& contempt = {}) {
    return adjustStatic(state.evaluate(pos, pos.stm()), pos.stm(), contempt);
}

template <typename Position> i32 staticEvalOnce(const Position& pos, const Contempt& contempt = {}) {
    return adjustStatic(NnueState::evaluateOnce(pos, pos.stm()), pos.stm(), contempt);
}

} // namespace sp::host::eval

#endif
