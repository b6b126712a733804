/*
 * test_nnue_state.cpp -- drives the C++ mirror of the reference eval API (csrc/host/nnue_state.h)
 * the way the engine does and checks the reference's own invariant
 *     evaluate() after any push / pop / applyImmediately sequence == evaluateOnce()
 * (src/datagen/datagen.cpp:262).  Built and run by tests/test_gpu_host_mirror.py on the GPU box.
 *
 *   usage: test_nnue_state <network file>
 */
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>

#include "../../stormphrax_b200/csrc/host/nnue_state.h"

using namespace sp::host;

static uint64_t g_rng = 0x9E3779B97F4A7C15ULL;
static uint32_t rnd(uint32_t n) {
    g_rng ^= g_rng << 13, g_rng ^= g_rng >> 7, g_rng ^= g_rng << 17;
    return static_cast<uint32_t>((g_rng >> 33) % n);
}

static int g_failures = 0, g_checks = 0;
static void expect(bool ok, const char* what, const Position& pos) {
    ++g_checks;
    if (ok) return;
    ++g_failures;
    std::fprintf(stderr, "FAIL %s at %s\n", what, pos.toFen().c_str());
}

/* depth-first walk with push/pop like ThreadData::applyMove + ThreadPosGuard (thread.cpp:46-67,
 * thread.h:107-127); evaluates lazily at some nodes only */
static void search(eval::NnueState& state, const Position& pos, int depth) {
    if (rnd(3) != 0) {
        expect(state.evaluate(pos, pos.stm()) == eval::NnueState::evaluateOnce(pos, pos.stm()), "search evaluate", pos);
        if (rnd(4) == 0) /* null-move child: parent's accumulators, flipped side (thread.cpp:25-44) */
            expect(state.evaluate(pos, pos.stm() ^ 1) == eval::NnueState::evaluateOnce(pos, pos.stm() ^ 1), "null-move evaluate", pos);
    }
    if (depth == 0) return;
    Move moves[256];
    const int n = pos.generateLegal(moves);
    for (int k = 0; k < 2 && n > 0; ++k) {
        const Move m = moves[rnd(static_cast<uint32_t>(n))];
        const Position child = pos.applyMove(m, state.push());
        search(state, child, depth - 1);
        state.pop();
    }
    if (rnd(2) == 0)
        expect(state.evaluate(pos, pos.stm()) == eval::NnueState::evaluateOnce(pos, pos.stm()), "evaluate after pop", pos);
}

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    if (!eval::init(nullptr, 0)) std::fprintf(stderr, "(expected) init without a network fails: %s\n", eval::lastError());
    if (eval::isNetworkLoaded()) return 3;
    if (!eval::initFromFile(argv[1])) return 4;

    eval::NnueState state{eval::getNetwork()};
    /* datagen form: applyImmediately after every played move (datagen.cpp:257-262) */
    for (int game = 0; game < 4; ++game) {
        Position pos = Position::startpos();
        state.reset(pos);
        for (int ply = 0; ply < 60; ++ply) {
            Move moves[256];
            const int n = pos.generateLegal(moves);
            if (!n) break;
            eval::UpdateContext ctx{};
            pos = pos.applyMove(moves[rnd(static_cast<uint32_t>(n))], eval::BoardObserver{ctx});
            state.applyImmediately(ctx, pos);
            expect(eval::staticEvalOnce(pos) == eval::staticEval(pos, state), "datagen consistency", pos);
        }
        /* search form from the position reached */
        state.reset(pos);
        search(state, pos, 5);
    }

    /* batched form: 64 concurrent games share one EvalBatch, one device batch per ply */
    {
        constexpr int kGames = 64;
        std::vector<eval::NnueState> states;
        std::vector<Position> games(kGames, Position::startpos());
        for (int g = 0; g < kGames; ++g) states.emplace_back(eval::getNetwork(), static_cast<uint32_t>(g) * eval::NnueState::kStackDepth);
        eval::EvalBatch batch{eval::getNetwork()};
        std::vector<int32_t> out(kGames);
        for (int g = 0; g < kGames; ++g) states[g].reset(games[g]);
        for (int ply = 0; ply < 30; ++ply) {
            for (int g = 0; g < kGames; ++g) {
                Move moves[256];
                const int n = games[g].generateLegal(moves);
                if (n) games[g] = games[g].applyMove(moves[rnd(static_cast<uint32_t>(n))], states[g].push());
                else states[g].push();
                states[g].evaluateAsync(batch, games[g], games[g].stm(), &out[g]);
            }
            if (batch.flush() != SP_OK) return 5;
            for (int g = 0; g < kGames; ++g)
                expect(out[g] == eval::NnueState::evaluateOnce(games[g], games[g].stm()), "batched evaluate", games[g]);
        }
    }
    /* several host threads, each a scheduler with its own evaluator context (createContext): 16 games per thread through its own
     * EvalBatch, all threads submitting to the GPU at once; results against the stack-free path afterwards on this thread */
    {
        constexpr int kThreads = 4, kGames = 16, kPlies = 24;
        std::vector<SpNnue*> contexts(kThreads, eval::getNetwork());
        for (int t = 1; t < kThreads; ++t) {
            contexts[t] = eval::createContext();
            if (!contexts[t]) return 6;
        }
        std::vector<std::vector<Position>> seen(kThreads);
        std::vector<std::vector<int32_t>> values(kThreads);
        std::vector<std::thread> workers;
        for (int t = 0; t < kThreads; ++t) {
            workers.emplace_back([&, t] {
                uint64_t rng = 0xD1B54A32D192ED03ULL * static_cast<uint64_t>(t + 1);
                const auto pick = [&rng](int n) {
                    rng ^= rng << 13, rng ^= rng >> 7, rng ^= rng << 17;
                    return static_cast<int>((rng >> 33) % static_cast<uint64_t>(n));
                };
                std::vector<eval::NnueState> states;
                std::vector<Position> games(kGames, Position::startpos());
                /* thread 0 shares the first context with the states above: slots past theirs */
                const uint32_t base = t == 0 ? 64 * eval::NnueState::kStackDepth : 0;
                for (int g = 0; g < kGames; ++g)
                    states.emplace_back(contexts[t], base + static_cast<uint32_t>(g) * eval::NnueState::kStackDepth);
                eval::EvalBatch batch{contexts[t]};
                std::vector<int32_t> out(kGames);
                for (int g = 0; g < kGames; ++g) states[g].invalidate();
                for (int ply = 0; ply < kPlies; ++ply) {
                    for (int g = 0; g < kGames; ++g) {
                        Move moves[256];
                        const int n = games[g].generateLegal(moves);
                        if (n) games[g] = games[g].applyMove(moves[pick(n)], states[g].push());
                        else states[g].push();
                        states[g].evaluateAsync(batch, games[g], games[g].stm(), &out[g]);
                    }
                    if (batch.flush() != SP_OK) std::abort();
                    for (int g = 0; g < kGames; ++g) seen[t].push_back(games[g]), values[t].push_back(out[g]);
                }
            });
        }
        for (auto& w : workers) w.join();
        for (int t = 0; t < kThreads; ++t)
            for (size_t i = 0; i < seen[t].size(); ++i)
                expect(values[t][i] == eval::NnueState::evaluateOnce(seen[t][i], seen[t][i].stm()), "threaded contexts", seen[t][i]);
        for (int t = 1; t < kThreads; ++t) eval::destroyContext(contexts[t]);
    }
    eval::shutdown();
    std::printf("%d checks, %d failures\n", g_checks, g_failures);
    return g_failures ? 1 : 0;
}
