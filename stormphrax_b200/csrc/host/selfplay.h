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

constexpr i32 kScoreMate = 32766;  /* src/core.h:706 */
constexpr i32 kScoreTbWin = 30000; /* src/core.h:707 */
constexpr int kMaxPly = 24;        /* search stack depth of the stand-in search (depth + extensions) */

inline bool isDecisive(i32 score) { return std::abs(score) > eval::kScoreWin; } /* core.h:722-724 */

/* datagen.cpp:72-90 */
constexpr i32 kWinAdjMinScore = 1250;
constexpr i32 kDrawAdjMaxScore = 10;
constexpr uint32_t kDrawAdjMinPlies = 70;
constexpr uint32_t kWinAdjPlyCount = 5;
constexpr uint32_t kDrawAdjPlyCount = 10;

enum class Outcome : uint8_t { kWhiteLoss = 0, kDraw, kWhiteWin }; /* datagen/common.h:24-28 */

/* wdl::wdlParams / normalizeScore<false>, src/wdl.cpp:28-80: a cubic in material / 58 (material clamped to
 * [17, 78]) gives the score that means "50 % win"; the normalised score is 100 * score / that. */
inline int classicalMaterial(const Position& pos) {
    return __builtin_popcountll(pos.bbType(kPawn)) + 3 * __builtin_popcountll(pos.bbType(kKnight))
         + 3 * __builtin_popcountll(pos.bbType(kBishop)) + 5 * __builtin_popcountll(pos.bbType(kRook))
         + 9 * __builtin_popcountll(pos.bbType(kQueen));
}
inline uint32_t plyFromStartpos(const Position& pos) { /* src/position.h:511-513 */
    return static_cast<uint32_t>(pos.fullmove() * 2 - (pos.stm() == kWhite ? 1 : 0) - 1);
}
inline i32 normalizeScore(i32 score, int material) {
    if (score == 0 || isDecisive(score)) return score;
    static constexpr double kA[4] = {-244.97139595, 687.39969858, -654.38002091, 608.47087786};
    const double m = static_cast<double>(std::clamp(material, 17, 78)) / 58.0;
    const double a = ((kA[0] * m + kA[1]) * m + kA[2]) * m + kA[3];
    return static_cast<i32>(std::round(100.0 * (static_cast<double>(score) / a)));
}

/* One game in viriformat (src/datagen/viriformat.cpp:33-63): the 32-byte start record with the outcome in
 * `wdl`, then (move, score) pairs of 4 bytes, then 4 zero bytes. */
inline uint16_t viriMove(Move m) {
    static constexpr uint16_t kTypes[4] = {0x0000, 0xC000, 0x8000, 0x4000}; /* standard, promotion, castling, en passant */
    const uint16_t promoIdx = m.type() == MoveType::kPromotion ? static_cast<uint16_t>(m.promo() - 1) : 0;
    return static_cast<uint16_t>(m.from() | m.to() << 6 | promoIdx << 12 | kTypes[static_cast<int>(m.type())]);
}
struct ViriGame {
    SpPackedBoard initial{};
    std::vector<std::pair<uint16_t, int16_t>> moves;

    void start(const Position& pos) {
        initial = pos.pack();
        initial.eval = 0, initial.wdl = 0, initial.extra = 0;
        moves.clear();
    }
    void push(Move move, i32 score) { moves.emplace_back(viriMove(move), static_cast<int16_t>(score)); }
    /* returns the number of positions written (moves + 1, like the reference) */
    size_t writeAllWithOutcome(std::vector<uint8_t>& out, Outcome outcome) {
        initial.wdl = static_cast<uint8_t>(outcome);
        const size_t at = out.size();
        out.resize(at + sizeof(SpPackedBoard) + 4 * moves.size() + 4, 0);
        std::memcpy(out.data() + at, &initial, sizeof(SpPackedBoard));
        uint8_t* p = out.data() + at + sizeof(SpPackedBoard);
        for (const auto& [mv, sc] : moves) {
            std::memcpy(p, &mv, 2), std::memcpy(p + 2, &sc, 2);
            p += 4;
        }
        return moves.size() + 1;
    }
};

struct Params {
    uint32_t concurrency = 1024;  /* games in flight per host thread */
    uint32_t totalGames = 1024;   /* games to finish per host thread */
    uint32_t depth = 3;           /* iterative deepening stops after this depth ... */
    uint32_t nodesPerMove = 5000; /* ... or once a finished iteration has used this many nodes (soft limit, datagen.cpp:76) */
    uint32_t maxPlies = 300;      /* games still undecided are drawn here (stands in for repetition detection) */
    uint64_t seed = 42;
};

struct Stats {
    uint64_t games{0}, positions{0}, nodes{0}, evals{0}, batches{0}, searches{0};
};

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

/* ------------------------------------------------------------------ one game: search + game loop as a state machine */
template <typename Evaluator>
class Game {
public:
    enum class Status { kNeedEval, kGameOver };

    void start(uint32_t id, uint64_t seed, const Params& params, Evaluator* evaluator, Stats* stats) {
        m_id = id, m_params = &params, m_eval = evaluator, m_stats = stats;
        m_rng = Jsf64{seed};
        newGame();
    }

    /* Runs until a static evaluation is needed (queued with the evaluator; call again after its flush)
     * or the game is over (its record is then in `record()` / `outcome()`). */
    Status step() {
        for (;;) {
            if (m_phase == Phase::kSearch) {
                if (!runSearch()) return Status::kNeedEval;
                /* one iteration finished */
                m_score = m_rootScore, m_best = m_rootBest;
                if (m_iterDepth < m_params->depth && m_searchNodes < m_params->nodesPerMove && !isDecisive(m_score)) {
                    beginIteration(m_iterDepth + 1);
                    continue;
                }
                ++m_stats->searches;
                if (playMove()) return Status::kGameOver;
                beginSearch();
            }
        }
    }

    [[nodiscard]] const ViriGame& record() const { return m_record; }
    [[nodiscard]] ViriGame& record() { return m_record; }
    [[nodiscard]] Outcome outcome() const { return m_outcome; }

private:
    enum class Phase { kSearch };
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

    void newGame() {
        /* datagen.cpp:146-178: 8 or 9 random plies from the start position; start over on a dead end */
        for (;;) {
            m_pos = Position::startpos();
            const uint32_t plies = 8 + static_cast<uint32_t>(m_rng.next() >> 63);
            bool dead = false;
            for (uint32_t i = 0; i < plies; ++i) {
                Move moves[256];
                const int n = m_pos.generateLegal(moves);
                if (!n) {
                    dead = true;
                    break;
                }
                m_pos = m_pos.applyMove(moves[m_rng.below(static_cast<uint32_t>(n))]);
            }
            Move moves[256];
            if (!dead && m_pos.generateLegal(moves)) break;
        }
        m_eval->reset(m_id, m_pos);
        m_record.start(m_pos);
        m_winPlies = m_lossPlies = m_drawPlies = 0;
        m_plies = 0;
        m_phase = Phase::kSearch;
        beginSearch();
    }

    void beginSearch() {
        m_searchNodes = 0;
        beginIteration(1);
    }

    void beginIteration(uint32_t depth) {
        m_iterDepth = depth;
        m_sp = 0;
        Node& root = m_stack[0];
        root.pos = m_pos;
        root.alpha = -kScoreMate, root.beta = kScoreMate;
        root.depth = static_cast<int>(depth);
        root.phase = NodePhase::kEnter;
        m_rootBest = Move{};
    }

    /* order: captures by victim value first, the previous iteration's best root move before everything */
    void orderMoves(Node& nd, bool root) const {
        auto key = [&](Move m) {
            int k = 0;
            const Piece victim = m.type() == MoveType::kEnPassant ? kPawn << 1 : (m.type() == MoveType::kCastling ? kNoPiece : nd.pos.pieceOn(m.to()));
            if (victim != kNoPiece) k = 16 + 2 * (victim >> 1) - ((nd.pos.pieceOn(m.from()) >> 1) > (victim >> 1) ? 1 : 0);
            if (m.type() == MoveType::kPromotion) k += 8;
            if (root && m == m_best) k = 1000;
            return k;
        };
        std::stable_sort(nd.moves, nd.moves + nd.n, [&](Move a, Move b) { return key(a) > key(b); });
    }

    /* Negamax alpha-beta on an explicit stack.  Returns false when it had to queue an evaluation. */
    bool runSearch() {
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
            parent.alpha = std::max(parent.alpha, v);
        }
    }

    /* datagen.cpp:206-300.  Returns true when the game is over. */
    bool playMove() {
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
            } else if (plyFromStartpos(m_pos) >= kDrawAdjMinPlies && std::abs(norm) < kDrawAdjMaxScore) {
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
        const bool bareKings = __builtin_popcountll(m_pos.occ()) <= 2;
        if (m_pos.halfmove() >= 100 || bareKings || m_plies >= m_params->maxPlies) { /* isDrawn stand-in, datagen.cpp:264-268 */
            m_outcome = Outcome::kDraw;
            m_record.push(move, 0);
            return true;
        }
        m_record.push(move, std::abs(whiteScore) <= 2 ? 0 : whiteScore);
        if (over) return true;
        if (!nReplies) { /* the reference finds this at its next search (datagen.cpp:213-222) */
            m_outcome = m_pos.isCheck() ? (m_pos.stm() == kBlack ? Outcome::kWhiteWin : Outcome::kWhiteLoss) : Outcome::kDraw;
            return true;
        }
        return false;
    }

    uint32_t m_id{0};
    const Params* m_params{nullptr};
    Evaluator* m_eval{nullptr};
    Stats* m_stats{nullptr};
    Jsf64 m_rng{0};
    Position m_pos;
    ViriGame m_record;
    Outcome m_outcome{Outcome::kDraw};
    Phase m_phase{Phase::kSearch};
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
    Driver(const Params& params, Evaluator& evaluator) : m_params{params}, m_eval{evaluator}, m_games(params.concurrency) {}

    /* Plays params.totalGames games, at most params.concurrency at a time; appends their viriformat records
     * to `out` in completion order.  Returns false if a device batch failed. */
    bool run(std::vector<uint8_t>& out, Stats& stats) {
        SplitMix64 seeds{m_params.seed};
        uint32_t started = 0;
        std::vector<uint32_t> active;
        for (uint32_t g = 0; g < m_params.concurrency && started < m_params.totalGames; ++g, ++started) {
            m_games[g].start(g, seeds.next(), m_params, &m_eval, &stats);
            active.push_back(g);
        }
        while (!active.empty()) {
            size_t keep = 0;
            for (size_t i = 0; i < active.size(); ++i) {
                const uint32_t g = active[i];
                bool alive = true;
                while (m_games[g].step() == Game<Evaluator>::Status::kGameOver) {
                    stats.games += 1;
                    m_games[g].record().writeAllWithOutcome(out, m_games[g].outcome());
                    if (started >= m_params.totalGames) {
                        alive = false;
                        break;
                    }
                    ++started;
                    m_games[g].start(g, seeds.next(), m_params, &m_eval, &stats); /* the slot starts its next game */
                }
                if (alive) active[keep++] = g;
            }
            active.resize(keep);
            if (active.empty()) break;
            ++stats.batches;
            if (!m_eval.flush()) return false;
        }
        return true;
    }

private:
    Params m_params;
    Evaluator& m_eval;
    std::vector<Game<Evaluator>> m_games;
};

} // namespace sp::host::selfplay

#endif
