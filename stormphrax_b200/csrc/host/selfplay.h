/*
 * selfplay.h -- batched self-play driver: the caller side of the evaluator for the datagen workload
 * (BASELINE config 5).
 *
 * The reference's datagen (src/datagen/datagen.cpp:96-321) runs ONE game per thread: random opening
 * plies, then search -> applyMove -> NnueState::applyImmediately -> record (move, score) until the game
 * is decided, then Viriformat::writeAllWithOutcome.  Its search evaluates one position at a time, which is
 * exactly what a GPU cannot use.  Here thousands of games run concurrently on each host thread: every
 * game's search is a resumable state machine that runs until it needs a static evaluation, queues it
 * with NnueState::evaluateAsync and yields; when every game of the thread is waiting, ONE device batch
 * (EvalBatch::flush) answers them all.
 *
 * What is mirrored from the reference: the game loop (opening randomisation, win / draw adjudication
 * thresholds, score clamping, filtered flag, viriformat records: datagen.cpp:146-300,
 * viriformat.cpp:33-63), the evaluator call protocol (reset / push / pop / applyImmediately / evaluate
 * with lazy catch-up: nodes in check are not evaluated, search.cpp:657-664, so their children catch up
 * across two plies), adjustStatic, and wdl::normalizeScore<false> (wdl.cpp:28-80).
 * What is NOT: the reference's search itself (out of scope).  The stand-in is a
 * plain iterative-deepening negamax alpha-beta with static evaluation at every node entered (as the
 * reference does), reverse futility pruning at shallow depth and a check extension -- enough to
 * produce the reference's access pattern on the accumulator stack.
 *
 * The driver is a template over the evaluator so that its scheduling and search logic can be tested on
 * a CPU-only box with a stand-in evaluator (tests/cpp/test_selfplay_host.cpp); the library itself
 * instantiates it ONLY with DeviceEvaluator (NnueState + EvalBatch over the C-ABI): there is no CPU
 * evaluation path in libsp_nnue.so.
 */
#ifndef SP_HOST_SELFPLAY_H
#define SP_HOST_SELFPLAY_H

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../../include/sp_nnue.h"
#include "nnue_state.h"
#include "position.h"
#include "rng.h"

namespace sp::host::selfplay {

using eval::i32;

template <typename T> SP_POS_HD constexpr T hdMax(T a, T b) { return a > b ? a : b; }
template <typename T> SP_POS_HD constexpr T hdAbs(T a) { return a < 0 ? -a : a; }

constexpr i32 kScoreMate = 32766;  /* src/core.h:706 */
constexpr i32 kScoreTbWin = 30000; /* src/core.h:707 */
constexpr int kMaxPly = 24;        /* search stack depth of the stand-in search (depth + extensions) */

SP_POS_HD inline bool isDecisive(i32 score) { return hdAbs(score) > eval::kScoreWin; } /* core.h:722-724 */

/* datagen.cpp:72-90 */
constexpr i32 kWinAdjMinScore = 1250;
constexpr i32 kDrawAdjMaxScore = 10;
constexpr uint32_t kDrawAdjMinPlies = 70;
constexpr uint32_t kWinAdjPlyCount = 5;
constexpr uint32_t kDrawAdjPlyCount = 10;

enum class Outcome : uint8_t { kWhiteLoss = 0, kDraw, kWhiteWin }; /* datagen/common.h:24-28 */

/* wdl::wdlParams / normalizeScore<false>, src/wdl.cpp:28-80: a cubic in material / 58 (material clamped to
 * [17, 78]) gives the score that means "50 % win"; the normalised score is 100 * score / that. */
SP_POS_HD inline int classicalMaterial(const Position& pos) {
    return popcount64(pos.bbType(kPawn)) + 3 * popcount64(pos.bbType(kKnight)) + 3 * popcount64(pos.bbType(kBishop))
         + 5 * popcount64(pos.bbType(kRook)) + 9 * popcount64(pos.bbType(kQueen));
}
SP_POS_HD inline uint32_t plyFromStartpos(const Position& pos) { /* src/position.h:511-513 */
    return static_cast<uint32_t>(pos.fullmove() * 2 - (pos.stm() == kWhite ? 1 : 0) - 1);
}
/* No fused multiply-add anywhere in it: the device must round exactly like the host (and the reference). */
SP_POS_HD inline i32 normalizeScore(i32 score, int material) {
    if (score == 0 || isDecisive(score)) return score;
    const double m = static_cast<double>(material < 17 ? 17 : (material > 78 ? 78 : material)) / 58.0;
#if defined(__CUDA_ARCH__)
    double a = __dadd_rn(__dmul_rn(-244.97139595, m), 687.39969858);
    a = __dadd_rn(__dmul_rn(a, m), -654.38002091);
    a = __dadd_rn(__dmul_rn(a, m), 608.47087786);
    return static_cast<i32>(round(__dmul_rn(100.0, __ddiv_rn(static_cast<double>(score), a))));
#else
    volatile double a = -244.97139595 * m; /* volatile: keeps the host compiler from contracting into FMAs too */
    a = a + 687.39969858;
    a = a * m;
    a = a + -654.38002091;
    a = a * m;
    a = a + 608.47087786;
    volatile double q = static_cast<double>(score) / a;
    q = 100.0 * q;
    return static_cast<i32>(std::round(q));
#endif
}

/* wdl::wdlModel (src/wdl.cpp:43-50): win / loss per mille of a side-to-move score, the logistic of
 * (score -+ a) / b with a, b the two material cubics of wdl::wdlParams.  Host only (UCI / report output). */
inline void wdlModel(i32 povScore, int material, i32& win, i32& loss) {
    const double m = static_cast<double>(material < 17 ? 17 : (material > 78 ? 78 : material)) / 58.0;
    auto cubic = [m](double c0, double c1, double c2, double c3) { /* Horner, one rounding per operation */
        volatile double v = c0 * m;
        v = v + c1;
        v = v * m;
        v = v + c2;
        v = v * m;
        v = v + c3;
        return static_cast<double>(v);
    };
    const double a = cubic(-244.97139595, 687.39969858, -654.38002091, 608.47087786);
    const double b = cubic(68.24072080, -111.17718819, 74.50316570, 71.16566713);
    const double x = static_cast<double>(povScore);
    win = static_cast<i32>(std::round(1000.0 / (1.0 + std::exp((a - x) / b))));
    loss = static_cast<i32>(std::round(1000.0 / (1.0 + std::exp((a + x) / b))));
}

/* One game in viriformat (src/datagen/viriformat.cpp:33-63): the 32-byte start record with the outcome in
 * `wdl`, then (move, score) pairs of 4 bytes, then 4 zero bytes. */
SP_POS_HD inline uint16_t viriMove(Move m) {
    const int type = static_cast<int>(m.type()); /* standard, promotion, castling, en passant -> 0x0000, 0xC000, 0x8000, 0x4000 */
    const uint16_t flags = static_cast<uint16_t>(type == 0 ? 0x0000 : (type == 1 ? 0xC000 : (type == 2 ? 0x8000 : 0x4000)));
    const uint16_t promoIdx = m.type() == MoveType::kPromotion ? static_cast<uint16_t>(m.promo() - 1) : 0;
    return static_cast<uint16_t>(m.from() | m.to() << 6 | promoIdx << 12 | flags);
}
constexpr uint32_t kMaxRecordMoves = 512; /* a record holds at most this many (move, score) pairs: Params::maxPlies <= 510 */
struct ViriGame {
    SpPackedBoard initial{};
    uint32_t n{0};
    uint32_t moves[kMaxRecordMoves]; /* move | score << 16, i.e. the record's bytes */

    SP_POS_HD void start(const Position& pos) {
        initial = pos.pack();
        initial.eval = 0, initial.wdl = 0, initial.extra = 0;
        n = 0;
    }
    SP_POS_HD void push(Move move, i32 score) {
        if (n < kMaxRecordMoves) moves[n++] = viriMove(move) | static_cast<uint32_t>(static_cast<uint16_t>(static_cast<int16_t>(score))) << 16;
    }
    [[nodiscard]] SP_POS_HD static constexpr size_t maxBytes(uint32_t maxPlies) { return sizeof(SpPackedBoard) + 4 * (size_t{maxPlies} + 2); }
    [[nodiscard]] SP_POS_HD size_t bytes() const { return sizeof(SpPackedBoard) + 4 * (size_t{n} + 1); }
    /* writes bytes() bytes: start record with the outcome, the pairs, the 4-byte terminator */
    SP_POS_HD void serialize(uint8_t* out, Outcome outcome) {
        initial.wdl = static_cast<uint8_t>(outcome);
        const uint8_t* src = reinterpret_cast<const uint8_t*>(&initial);
        for (size_t i = 0; i < sizeof(SpPackedBoard); ++i) out[i] = src[i];
        uint8_t* q = out + sizeof(SpPackedBoard);
        for (uint32_t i = 0; i <= n; ++i) {
            const uint32_t w = i < n ? moves[i] : 0u;
            q[4 * i] = static_cast<uint8_t>(w), q[4 * i + 1] = static_cast<uint8_t>(w >> 8);
            q[4 * i + 2] = static_cast<uint8_t>(w >> 16), q[4 * i + 3] = static_cast<uint8_t>(w >> 24);
        }
    }
    /* returns the number of positions written (moves + 1, like the reference) */
    size_t writeAllWithOutcome(std::vector<uint8_t>& out, Outcome outcome) {
        const size_t at = out.size();
        out.resize(at + bytes());
        serialize(out.data() + at, outcome);
        return size_t{n} + 1;
    }
};

struct Params {
    uint32_t concurrency = 1024;  /* game slots of this driver instance: games in flight */
    uint32_t totalGames = 1024;   /* games to play: slot g plays totalGames / concurrency of them (+ 1 for the first totalGames % concurrency slots) */
    uint32_t depth = 3;           /* iterative deepening stops after this depth ... */
    uint32_t nodesPerMove = 5000; /* ... or once a finished iteration has used this many nodes (soft limit, datagen.cpp:76) */
    uint32_t maxPlies = 300;      /* games still undecided are drawn here (stands in for repetition detection); at most 510 */
    uint64_t seed = 42;
    uint32_t dfrc = 0;            /* 1: every game starts from a random double-Fischer-random position (datagen.cpp:146-148) */
    uint32_t slotBegin = 0, slotEnd = 0; /* the slots THIS driver instance plays (a host thread's share); 0, 0 = all */
    Position start;               /* the start position (Position::startpos(), built on the host) */

    [[nodiscard]] SP_POS_HD uint32_t localSlots() const { return slotEnd > slotBegin ? slotEnd - slotBegin : concurrency; }
    [[nodiscard]] SP_POS_HD uint32_t firstSlot() const { return slotEnd > slotBegin ? slotBegin : 0; }

    [[nodiscard]] SP_POS_HD uint32_t gamesOfSlot(uint32_t slot) const { return totalGames / concurrency + (slot < totalGames % concurrency ? 1 : 0); }
    /* Every (slot, game number) has its own random stream, whichever driver or thread plays it.  (slot and k
     * enter through different odd multipliers, neither of them SplitMix64's own increment: stepping one
     * generator per slot made game k of slot g the same game as game 0 of slot g + k.) */
    [[nodiscard]] SP_POS_HD uint64_t gameSeed(uint32_t slot, uint32_t k) const {
        SplitMix64 mix{seed ^ (0xD1B54A32D192ED03ULL * (uint64_t{slot} + 1)) ^ (0x8CB92BA72F3D8DD7ULL * (uint64_t{k} + 1))};
        return mix.next();
    }
};

struct Stats {
    uint64_t games{0}, positions{0}, nodes{0}, evals{0}, batches{0}, searches{0};
};

/* C-ABI parameters -> Params (host only: builds the start position) */
inline Params makeParams(const SpSelfplayParams& in) {
    Params p;
    p.concurrency = in.concurrency ? in.concurrency : 1;
    p.totalGames = in.total_games;
    p.depth = in.depth < 1 ? 1 : (in.depth > static_cast<uint32_t>(kMaxPly - 8) ? static_cast<uint32_t>(kMaxPly - 8) : in.depth);
    p.nodesPerMove = in.nodes_per_move;
    p.maxPlies = in.max_plies ? (in.max_plies > kMaxRecordMoves - 2 ? kMaxRecordMoves - 2 : in.max_plies) : 300;
    p.seed = in.seed;
    p.dfrc = in.dfrc ? 1 : 0;
    p.start = Position::startpos();
    return p;
}

/* ------------------------------------------------------------------ evaluator over the device */
class DeviceEvaluator {
public:
    DeviceEvaluator(SpNnue* network, uint32_t games) : m_batch{network} {
        sp_nnue_slots_reserve(network, size_t{games} * (kMaxPly + 1)); /* once, instead of growing state by state */
        m_states.reserve(games);
        for (uint32_t g = 0; g < games; ++g) m_states.emplace_back(network, g * (kMaxPly + 1), kMaxPly + 1);
    }
    /* NnueState::reset is synchronous (one tiny launch per game); a batched driver lets the game's first
     * evaluation do the rebuild instead */
    void reset(uint32_t game, const Position&) { m_states[game].invalidate(); }
    auto push(uint32_t game) { return m_states[game].push(); }
    void pop(uint32_t game) { m_states[game].pop(); }
    void applyImmediately(uint32_t game, const Position&) { m_states[game].applyLazily(); }
    void evaluateAsync(uint32_t game, const Position& pos, i32* out) { m_states[game].evaluateAsync(m_batch, pos, pos.stm(), out); }
    bool flush() { return m_batch.flush() == SP_OK; }

private:
    eval::EvalBatch m_batch;
    std::vector<eval::NnueState> m_states;
};

/* ------------------------------------------------------------------ one game: search + game loop as a state machine
 * Host and device code: the same class runs per host-driver slot and per GPU thread (selfplay_gpu.cu). */
template <typename Evaluator>
class Game {
public:
    enum class Status { kNeedEval, kGameOver };

    SP_POS_HD void start(uint32_t id, uint64_t seed, const Params& params, Evaluator* evaluator, Stats* stats) {
        m_id = id, m_params = params, m_eval = evaluator, m_stats = stats;
        m_rng = Jsf64{seed};
        newGame();
    }
    /* after the object was moved to another address space: rebind what it points to */
    SP_POS_HD void bind(Evaluator* evaluator, Stats* stats) { m_eval = evaluator, m_stats = stats; }

    /* Runs until a static evaluation is needed (queued with the evaluator; call again after its flush)
     * or the game is over (its record is then in `record()` / `outcome()`). */
    SP_POS_HD Status step() {
        for (;;) {
            if (!runSearch()) return Status::kNeedEval;
            /* one iteration finished */
            m_score = m_rootScore, m_best = m_rootBest;
            if (m_iterDepth < m_params.depth && m_searchNodes < m_params.nodesPerMove && !isDecisive(m_score)) {
                beginIteration(m_iterDepth + 1);
                continue;
            }
            ++m_stats->searches;
            if (playMove()) return Status::kGameOver;
            beginSearch();
        }
    }

    [[nodiscard]] SP_POS_HD ViriGame& record() { return m_record; }
    [[nodiscard]] SP_POS_HD Outcome outcome() const { return m_outcome; }
    [[nodiscard]] SP_POS_HD i32* leaf() { return &m_leaf; }

private:
    enum class NodePhase : uint8_t { kEnter, kWaitEval, kLoop };

    struct Node {
        Position pos;
        Move moves[256];
        int n, next;
        i32 alpha, beta, best;
        int depth;
        bool inCheck;
        NodePhase phase;
    };

    SP_POS_HD void newGame() {
        /* datagen.cpp:146-178: 8 or 9 random plies from the (standard or random DFRC) start position; start over on a dead end */
        for (;;) {
            m_pos = m_params.dfrc ? Position::fromDfrcIndex(m_rng.below(960u * 960u)) : m_params.start;
            const uint32_t plies = 8 + static_cast<uint32_t>(m_rng.next() >> 63);
            bool dead = false;
            Move moves[256];
            for (uint32_t i = 0; i < plies; ++i) {
                const int n = m_pos.generateLegal(moves);
                if (!n) {
                    dead = true;
                    break;
                }
                m_pos = m_pos.applyMove(moves[m_rng.below(static_cast<uint32_t>(n))]);
            }
            if (!dead && m_pos.generateLegal(moves)) break;
        }
        m_eval->reset(m_id, m_pos);
        m_record.start(m_pos);
        m_winPlies = m_lossPlies = m_drawPlies = 0;
        m_plies = 0;
        beginSearch();
    }

    SP_POS_HD void beginSearch() {
        m_searchNodes = 0;
        beginIteration(1);
    }

    SP_POS_HD void beginIteration(uint32_t depth) {
        m_iterDepth = depth;
        m_sp = 0;
        Node& root = m_stack[0];
        root.pos = m_pos;
        root.alpha = -kScoreMate, root.beta = kScoreMate;
        root.depth = static_cast<int>(depth);
        root.phase = NodePhase::kEnter;
        m_rootBest = Move{};
    }

    /* order: captures by victim value first, the previous iteration's best root move before everything.
     * Stable insertion sort by descending key (the same order std::stable_sort gives). */
    SP_POS_HD void orderMoves(Node& nd, bool root) const {
        int keys[256];
        for (int i = 0; i < nd.n; ++i) {
            const Move m = nd.moves[i];
            int k = 0;
            const Piece victim = m.type() == MoveType::kEnPassant ? kPawn << 1 : (m.type() == MoveType::kCastling ? kNoPiece : nd.pos.pieceOn(m.to()));
            if (victim != kNoPiece) k = 16 + 2 * (victim >> 1) - ((nd.pos.pieceOn(m.from()) >> 1) > (victim >> 1) ? 1 : 0);
            if (m.type() == MoveType::kPromotion) k += 8;
            if (root && m == m_best) k = 1000;
            int j = i;
            for (; j > 0 && keys[j - 1] < k; --j) keys[j] = keys[j - 1], nd.moves[j] = nd.moves[j - 1];
            keys[j] = k, nd.moves[j] = m;
        }
    }

    /* Negamax alpha-beta on an explicit stack.  Returns false when it had to queue an evaluation. */
    SP_POS_HD bool runSearch() {
        i32 ret = 0;
        for (;;) {
            Node& nd = m_stack[m_sp];
            bool done = false;
            if (nd.phase == NodePhase::kEnter) {
                ++m_stats->nodes, ++m_searchNodes;
                nd.n = nd.pos.generateLegal(nd.moves);
                nd.inCheck = nd.pos.isCheck();
                nd.next = 0, nd.best = -kScoreMate;
                if (nd.n == 0) {
                    ret = nd.inCheck ? -kScoreMate + m_sp : 0, done = true;
                } else if (m_sp > 0 && nd.pos.halfmove() >= 100) {
                    ret = 0, done = true;
                } else {
                    if (nd.inCheck && nd.depth == 0 && m_sp + 1 < kMaxPly) nd.depth = 1; /* check extension */
                    if (!nd.inCheck || nd.depth == 0) {
                        /* static evaluation, as the reference takes it at every node it enters unless in
                         * check (search.cpp:657-664) */
                        ++m_stats->evals;
                        m_eval->evaluateAsync(m_id, nd.pos, &m_leaf);
                        nd.phase = NodePhase::kWaitEval;
                        return false;
                    }
                    orderMoves(nd, m_sp == 0);
                    nd.phase = NodePhase::kLoop;
                }
            } else if (nd.phase == NodePhase::kWaitEval) {
                const i32 staticEval = eval::adjustStatic(m_leaf, nd.pos.stm(), {});
                if (nd.depth == 0) {
                    ret = staticEval, done = true;
                } else if (m_sp > 0 && nd.depth <= 2 && staticEval - 120 * nd.depth >= nd.beta) {
                    ret = staticEval, done = true; /* reverse futility pruning */
                } else {
                    orderMoves(nd, m_sp == 0);
                    nd.phase = NodePhase::kLoop;
                }
            }
            if (!done && nd.phase == NodePhase::kLoop) {
                if (nd.next == nd.n || nd.alpha >= nd.beta) {
                    ret = nd.best, done = true;
                } else {
                    const Move m = nd.moves[nd.next++];
                    Node& child = m_stack[m_sp + 1];
                    child.pos = nd.pos.applyMove(m, m_eval->push(m_id));
                    child.alpha = -nd.beta, child.beta = -nd.alpha;
                    child.depth = nd.depth - 1;
                    child.phase = NodePhase::kEnter;
                    ++m_sp;
                    continue;
                }
            }
            /* done: hand `ret` to the parent */
            if (m_sp == 0) {
                m_rootScore = ret;
                return true;
            }
            m_eval->pop(m_id);
            --m_sp;
            Node& parent = m_stack[m_sp];
            const i32 v = -ret;
            if (v > parent.best) {
                parent.best = v;
                if (m_sp == 0) m_rootBest = parent.moves[parent.next - 1];
            }
            parent.alpha = hdMax(parent.alpha, v);
        }
    }

    /* datagen.cpp:206-300.  Returns true when the game is over. */
    SP_POS_HD bool playMove() {
        const i32 score = m_score; /* side-to-move relative; the reference's datagen search reports white-relative */
        const i32 whiteScore = m_pos.stm() == kWhite ? score : -score;
        const Move move = m_best;
        bool over = false;
        if (!move) { /* cannot happen: games never start or continue from a position without moves */
            m_outcome = Outcome::kDraw;
            return true;
        }
        if (isDecisive(whiteScore)) {
            m_outcome = whiteScore > 0 ? Outcome::kWhiteWin : Outcome::kWhiteLoss, over = true;
        } else {
            const i32 norm = normalizeScore(whiteScore, classicalMaterial(m_pos));
            if (norm > kWinAdjMinScore) {
                ++m_winPlies, m_lossPlies = 0, m_drawPlies = 0;
            } else if (norm < -kWinAdjMinScore) {
                m_winPlies = 0, ++m_lossPlies, m_drawPlies = 0;
            } else if (plyFromStartpos(m_pos) >= kDrawAdjMinPlies && hdAbs(norm) < kDrawAdjMaxScore) {
                m_winPlies = 0, m_lossPlies = 0, ++m_drawPlies;
            } else {
                m_winPlies = m_lossPlies = m_drawPlies = 0;
            }
            if (m_winPlies >= kWinAdjPlyCount) m_outcome = Outcome::kWhiteWin, over = true;
            else if (m_lossPlies >= kWinAdjPlyCount) m_outcome = Outcome::kWhiteLoss, over = true;
            else if (m_drawPlies >= kDrawAdjPlyCount) m_outcome = Outcome::kDraw, over = true;
        }

        eval::UpdateContext ctx{};
        m_pos = m_pos.applyMove(move, eval::BoardObserver{ctx});
        m_eval->applyImmediately(m_id, m_pos);
        ++m_plies, ++m_stats->positions;

        Move replies[256];
        const int nReplies = m_pos.generateLegal(replies);
        const bool bareKings = popcount64(m_pos.occ()) <= 2;
        /* isDrawn stand-in, datagen.cpp:264-268.  Fifty-move rule as Position::isDrawn has it (position.cpp:621-633):
         * a side that is in check with no legal reply has been mated, not drawn. */
        const bool fiftyMoves = m_pos.halfmove() >= 100 && (nReplies || !m_pos.isCheck());
        if (fiftyMoves || bareKings || m_plies >= m_params.maxPlies) {
            m_outcome = Outcome::kDraw;
            m_record.push(move, 0);
            return true;
        }
        m_record.push(move, hdAbs(whiteScore) <= 2 ? 0 : whiteScore);
        if (over) return true;
        if (!nReplies) { /* the reference finds this at its next search (datagen.cpp:213-222) */
            m_outcome = m_pos.isCheck() ? (m_pos.stm() == kBlack ? Outcome::kWhiteWin : Outcome::kWhiteLoss) : Outcome::kDraw;
            return true;
        }
        return false;
    }

    uint32_t m_id{0};
    Params m_params;
    Evaluator* m_eval{nullptr};
    Stats* m_stats{nullptr};
    Jsf64 m_rng{0};
    Position m_pos;
    ViriGame m_record;
    Outcome m_outcome{Outcome::kDraw};
    uint32_t m_winPlies{0}, m_lossPlies{0}, m_drawPlies{0}, m_plies{0};
    /* search */
    Node m_stack[kMaxPly + 1];
    int m_sp{0};
    uint32_t m_iterDepth{0}, m_searchNodes{0};
    i32 m_leaf{0}, m_rootScore{0}, m_score{0};
    Move m_rootBest{}, m_best{};
};

/* ------------------------------------------------------------------ scheduler of one host thread */
template <typename Evaluator>
class Driver {
public:
    Driver(const Params& params, Evaluator& evaluator) : m_params{params}, m_eval{evaluator}, m_games(params.localSlots()) {}

    /* Slot g plays params.gamesOfSlot(g) games one after the other, all slots concurrently.  The viriformat
     * records are appended to `out` slot by slot (slot-major, game order within a slot): the same bytes
     * whichever driver -- this one or the GPU-resident one -- played them.  Returns false if a device
     * batch failed. */
    bool run(std::vector<uint8_t>& out, Stats& stats) {
        const uint32_t slots = m_params.localSlots(), first = m_params.firstSlot(); /* local index l = global slot first + l */
        std::vector<std::vector<uint8_t>> records(slots);
        std::vector<uint32_t> played(slots, 0), active;
        for (uint32_t g = 0; g < slots; ++g) {
            if (!m_params.gamesOfSlot(first + g)) continue;
            m_games[g].start(g, m_params.gameSeed(first + g, 0), m_params, &m_eval, &stats);
            active.push_back(g);
        }
        while (!active.empty()) {
            size_t keep = 0;
            for (size_t i = 0; i < active.size(); ++i) {
                const uint32_t g = active[i];
                bool alive = true;
                while (m_games[g].step() == Game<Evaluator>::Status::kGameOver) {
                    stats.games += 1;
                    m_games[g].record().writeAllWithOutcome(records[g], m_games[g].outcome());
                    if (++played[g] >= m_params.gamesOfSlot(first + g)) {
                        alive = false;
                        break;
                    }
                    m_games[g].start(g, m_params.gameSeed(first + g, played[g]), m_params, &m_eval, &stats); /* the slot's next game */
                }
                if (alive) active[keep++] = g;
            }
            active.resize(keep);
            if (active.empty()) break;
            ++stats.batches;
            if (!m_eval.flush()) return false;
        }
        for (const auto& r : records) out.insert(out.end(), r.begin(), r.end());
        return true;
    }

private:
    Params m_params;
    Evaluator& m_eval;
    std::vector<Game<Evaluator>> m_games;
};

} // namespace sp::host::selfplay

#endif
