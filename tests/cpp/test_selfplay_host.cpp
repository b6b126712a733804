/*
 * test_selfplay_host.cpp -- CPU-only check of the batched self-play driver's HOST logic
 * (stormphrax_b200/csrc/host/selfplay.h): the resumable search state machine, the scheduler and the
 * viriformat writer, with a stand-in evaluator (material count) in place of the device.  The stand-in
 * lives in this test only; the library instantiates the driver with DeviceEvaluator alone.
 *
 * Checks
 *   1. protocol: at most one evaluation pending per game and flush, push / pop balanced, stack depth
 *      bounded, results only delivered by flush();
 *   2. equivalence: every search the state machine finishes returns the score and move of a plain
 *      RECURSIVE implementation of the same algorithm with a synchronous evaluator;
 *   3. records: every viriformat record replays move by move through the legal move generator, ends with
 *      the 4-byte terminator and carries a valid outcome.
 *
 *   usage: test_selfplay_host   (exit code 0 = ok)
 */
#include <cstdio>
#include <cstdlib>
#include <map>
#include <string>

#include "../../stormphrax_b200/csrc/host/selfplay.h"

using namespace sp;
using namespace sp::host;
using selfplay::i32;

static int g_failures = 0;
#define EXPECT(cond, ...)                                \
    do {                                                 \
        if (!(cond)) {                                   \
            ++g_failures;                                \
            std::fprintf(stderr, "FAIL: " __VA_ARGS__);  \
            std::fprintf(stderr, "\n");                  \
        }                                                \
    } while (0)

/* side-to-move relative material + a little mobility noise from the board hash so that scores differ */
static i32 standInEval(const Position& pos) {
    static const int kValue[6] = {100, 320, 330, 500, 900, 0};
    int sum = 0;
    for (int t = 0; t < 5; ++t)
        sum += kValue[t] * (__builtin_popcountll(pos.bb(t, kWhite)) - __builtin_popcountll(pos.bb(t, kBlack)));
    const uint64_t h = (pos.occ() * 0x9E3779B97F4A7C15ULL) >> 58; /* 0..63 */
    sum += static_cast<int>(h) - 32;
    return pos.stm() == kWhite ? sum : -sum;
}

struct StandInEvaluator {
    struct Pending { Position pos; i32* out; };
    std::map<uint32_t, Pending> pending;
    std::map<uint32_t, int> depth;
    uint64_t flushes = 0, delivered = 0;

    void reset(uint32_t game, const Position&) { depth[game] = 0; }
    eval::BoardObserver push(uint32_t game) {
        EXPECT(++depth[game] <= selfplay::kMaxPly, "stack overflow in game %u", game);
        return eval::BoardObserver{ctx};
    }
    void pop(uint32_t game) { EXPECT(--depth[game] >= 0, "pop below the root in game %u", game); }
    void applyImmediately(uint32_t game, const Position&) { EXPECT(depth[game] == 0, "applyImmediately away from the root"); }
    void evaluateAsync(uint32_t game, const Position& pos, i32* out) {
        EXPECT(!pending.count(game), "two evaluations pending for game %u", game);
        *out = 0x7FFFFFFF; /* poison: must not be read before the flush */
        pending[game] = Pending{pos, out};
    }
    bool flush() {
        ++flushes;
        for (auto& [game, p] : pending) *p.out = standInEval(p.pos), ++delivered;
        pending.clear();
        return true;
    }
    eval::UpdateContext ctx{};
};

/* ---- the same search, recursively, with a synchronous evaluator */
struct Recursive {
    uint32_t nodes = 0;
    Move rootBest{}, prevBest{};

    static void order(const Position& pos, Move* moves, int n, bool root, Move best) {
        auto key = [&](Move m) {
            int k = 0;
            const Piece victim = m.type() == MoveType::kEnPassant ? kPawn << 1 : (m.type() == MoveType::kCastling ? kNoPiece : pos.pieceOn(m.to()));
            if (victim != kNoPiece) k = 16 + 2 * (victim >> 1) - ((pos.pieceOn(m.from()) >> 1) > (victim >> 1) ? 1 : 0);
            if (m.type() == MoveType::kPromotion) k += 8;
            if (root && m == best) k = 1000;
            return k;
        };
        std::stable_sort(moves, moves + n, [&](Move a, Move b) { return key(a) > key(b); });
    }

    i32 search(const Position& pos, int depth, i32 alpha, i32 beta, int ply) {
        ++nodes;
        Move moves[256];
        const int n = pos.generateLegal(moves);
        const bool inCheck = pos.isCheck();
        if (n == 0) return inCheck ? -selfplay::kScoreMate + ply : 0;
        if (ply > 0 && pos.halfmove() >= 100) return 0;
        if (inCheck && depth == 0 && ply + 1 < selfplay::kMaxPly) depth = 1;
        if (!inCheck || depth == 0) {
            const i32 staticEval = eval::adjustStatic(standInEval(pos), pos.stm(), {});
            if (depth == 0) return staticEval;
            if (ply > 0 && depth <= 2 && staticEval - 120 * depth >= beta) return staticEval;
        }
        order(pos, moves, n, ply == 0, prevBest);
        i32 best = -selfplay::kScoreMate;
        for (int i = 0; i < n && alpha < beta; ++i) {
            const i32 v = -search(pos.applyMove(moves[i]), depth - 1, -beta, -alpha, ply + 1);
            if (v > best) {
                best = v;
                if (ply == 0) rootBest = moves[i];
            }
            alpha = std::max(alpha, v);
        }
        return best;
    }

    /* iterative deepening like Game::step */
    std::pair<i32, Move> run(const Position& pos, const selfplay::Params& p) {
        nodes = 0, prevBest = Move{};
        i32 score = 0;
        for (uint32_t d = 1;; ++d) {
            rootBest = Move{};
            score = search(pos, static_cast<int>(d), -selfplay::kScoreMate, selfplay::kScoreMate, 0);
            prevBest = rootBest;
            if (!(d < p.depth && nodes < p.nodesPerMove && !selfplay::isDecisive(score))) break;
        }
        return {score, prevBest};
    }
};

static int run(const selfplay::Params& params) {
    StandInEvaluator evaluator;
    selfplay::Driver<StandInEvaluator> driver{params, evaluator};
    std::vector<uint8_t> out;
    selfplay::Stats stats;
    EXPECT(driver.run(out, stats), "driver.run failed");
    EXPECT(stats.games == params.totalGames, "games %llu", static_cast<unsigned long long>(stats.games));
    EXPECT(evaluator.delivered == stats.evals, "evals queued %llu, delivered %llu", static_cast<unsigned long long>(stats.evals),
           static_cast<unsigned long long>(evaluator.delivered));
    EXPECT(stats.batches > 0 && stats.evals / stats.batches >= params.concurrency / 4, "batches are not being filled: %llu evals in %llu batches",
           static_cast<unsigned long long>(stats.evals), static_cast<unsigned long long>(stats.batches));

    /* ---- records: replay, and re-search every recorded position with the recursive search */
    size_t at = 0, games = 0, positions = 0, compared = 0;
    std::map<std::string, int> seen; /* record bytes -> count: every game must have its own random stream */
    while (at < out.size()) {
        const size_t record_begin = at;
        SpPackedBoard initial;
        std::memcpy(&initial, out.data() + at, sizeof(initial));
        at += sizeof(initial);
        EXPECT(initial.wdl <= 2, "bad outcome byte %u", initial.wdl);
        Position pos;
        EXPECT(Position::fromPacked(initial, pos), "record %zu: start board does not unpack", games);
        for (;;) {
            uint16_t mv;
            int16_t score;
            std::memcpy(&mv, out.data() + at, 2), std::memcpy(&score, out.data() + at + 2, 2);
            at += 4;
            if (mv == 0 && score == 0) break; /* terminator */
            Move legal[256];
            const int n = pos.generateLegal(legal);
            Move played{};
            for (int i = 0; i < n; ++i)
                if (selfplay::viriMove(legal[i]) == mv) played = legal[i];
            EXPECT(static_cast<bool>(played), "record %zu: move %04x is not legal in %s", games, mv, pos.toFen().c_str());
            if (!played) return 1;
            if (games % 6 == 0) { /* the state machine's result == the recursive search's */
                Recursive ref;
                const auto [refScore, refMove] = ref.run(pos, params);
                const i32 white = pos.stm() == kWhite ? refScore : -refScore;
                const bool last = [&] { uint16_t nm; int16_t ns; std::memcpy(&nm, out.data() + at, 2), std::memcpy(&ns, out.data() + at + 2, 2); return nm == 0 && ns == 0; }();
                EXPECT(refMove == played, "record %zu: played %04x, recursive search says %04x at %s", games, mv, selfplay::viriMove(refMove), pos.toFen().c_str());
                /* the last move of a drawn game is recorded with score 0 (datagen.cpp:264-268) */
                if (!(last && score == 0))
                    EXPECT(score == (std::abs(white) <= 2 ? 0 : static_cast<int16_t>(white)), "record %zu: score %d, recursive search says %d", games, score, white);
                ++compared;
            }
            pos = pos.applyMove(played);
            ++positions;
        }
        EXPECT(++seen[std::string(out.begin() + static_cast<long>(record_begin), out.begin() + static_cast<long>(at))] == 1, "record %zu: the same game was played twice", games);
        ++games;
    }
    EXPECT(games == params.totalGames, "parsed %zu records", games);
    EXPECT(positions == stats.positions, "parsed %zu positions, driver counted %llu", positions, static_cast<unsigned long long>(stats.positions));
    std::printf("%s: %zu games, %zu positions, %llu nodes, %llu evals in %llu batches (%.1f per batch), %zu searches re-checked, %d failures\n",
                params.dfrc ? "dfrc" : "standard", games,
                positions, static_cast<unsigned long long>(stats.nodes), static_cast<unsigned long long>(stats.evals),
                static_cast<unsigned long long>(stats.batches), static_cast<double>(stats.evals) / static_cast<double>(stats.batches), compared, g_failures);
    return g_failures ? 1 : 0;
}

int main() {
    selfplay::Params params;
    params.start = Position::startpos();
    params.concurrency = 24, params.totalGames = 40, params.depth = 3, params.nodesPerMove = 400, params.maxPlies = 120, params.seed = 7;
    int rc = run(params);
    params.dfrc = 1, params.totalGames = 30, params.seed = 8; /* `datagen dfrc`: Chess960 castling through search, records and replay */
    rc |= run(params);
    return rc;
}
