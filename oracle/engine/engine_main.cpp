/*
 * engine_main.cpp -- TEST INFRASTRUCTURE: a driver over the reference engine's own sources (search.cpp, bench.cpp,
 * datagen.cpp, thread.cpp, position.cpp ... compiled where they lie, oracle/engine/Makefile), built twice:
 *
 *   sp_engine_cpu    the stock CPU evaluation (src/eval/**), network loaded at run time
 *   sp_engine_b200   -DSP_EVAL_B200: eval/nnue.h and eval/nnue_state.h replaced by the adapter over libsp_nnue.so
 *
 *   <binary> <network.nnue> bench [depth]            the engine's own `bench` (src/bench.cpp:95-170): node-count signature
 *   <binary> <network.nnue> evalcheck <games> <seed> datagen's invariant (src/datagen/datagen.cpp:257-262) on random
 *                                                   playouts: staticEvalOnce == staticEval(nnueState) after
 *                                                   applyMove<BoardObserver> + applyImmediately; prints an eval checksum
 *   <binary> <network.nnue> datagen <dir> <seconds>  the engine's own datagen::run (one thread), interrupted after <seconds>
 *
 * Bit-exact evaluations make both binaries walk the same search trees: their bench node counts must be identical.
 */
#include <chrono>
#include <csignal>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <thread>
#include <vector>

#include "bench.h"
#include "cuckoo.h"
#include "datagen/datagen.h"
#include "eval/eval.h"
#include "movegen.h"
#include "opts.h"
#include "position.h"
#include "tunable.h"
#include "util/numa/numa.h"
#include "util/rng.h"

using namespace stormphrax;

#ifndef SP_EVAL_B200
namespace stormphrax::eval {
    bool oracleLoadNetwork(const std::byte* payload, usize size);
}
#endif

namespace {
    int evalCheck(u32 games, u64 seed) {
        opts::mutableOpts().chess960 = false;
        util::rng::SeedGenerator seeds{seed};
        eval::NnueState state{};
        state.setNetwork(eval::getNetwork(0));
        u64 checksum = 0, positions = 0, mismatches = 0;
        for (u32 g = 0; g < games; ++g) {
            util::rng::Jsf64Rng rng{seeds.nextSeed()};
            auto pos = Position::startpos();
            state.reset(pos);
            for (u32 ply = 0; ply < 80 && pos.occ().popcount() > 2; ++ply) {
                ScoredMoveList moves;
                generateAll(moves, pos);
                StaticVector<Move, 256> legal;
                for (const auto [move, score] : moves) {
                    if (pos.isLegal(move)) {
                        legal.push(move);
                    }
                }
                if (legal.empty()) {
                    break;
                }
                const auto move = legal[rng.nextU32(static_cast<u32>(legal.size()))];
                eval::UpdateContext ctx{};
                pos = pos.applyMove(move, eval::BoardObserver{ctx}); // datagen.cpp:259
                state.applyImmediately(ctx, pos);                    // datagen.cpp:260
                const auto once = eval::staticEvalOnce(pos);
                const auto incremental = eval::staticEval(pos, state);
                mismatches += once != incremental; // datagen.cpp:262
                checksum = checksum * 0x100000001B3ull + static_cast<u32>(incremental);
                ++positions;
            }
        }
        println("evalcheck: {} positions, {} mismatches, checksum {:016x}", positions, mismatches, checksum);
        return mismatches ? 1 : 0;
    }
} // namespace

int main(int argc, char** argv) {
    if (argc < 3) {
        eprintln("usage: {} <network.nnue> bench [depth] | evalcheck <games> <seed> | datagen <dir> <seconds>", argv[0]);
        return 2;
    }
    if (!numa::init()) {
        return 1;
    }
    tunable::init();
    cuckoo::init();

    std::ifstream in{argv[1], std::ios::binary};
    const std::vector<char> bytes{std::istreambuf_iterator<char>{in}, std::istreambuf_iterator<char>{}};
    if (bytes.size() < 64) {
        eprintln("cannot read network file {}", argv[1]);
        return 1;
    }
#ifdef SP_EVAL_B200
    const char* dev = std::getenv("SP_ENGINE_DEVICE");
    if (!eval::initB200(bytes.data(), bytes.size(), dev ? std::atoi(dev) : 0)) {
        return 1;
    }
#else
    if (!eval::oracleLoadNetwork(reinterpret_cast<const std::byte*>(bytes.data()) + 64, bytes.size() - 64)) {
        eprintln("failed to load network {}", argv[1]);
        return 1;
    }
#endif

    const std::string cmd = argv[2];
    int rc = 0;
    if (cmd == "bench") {
        bench::run(argc > 3 ? std::atoi(argv[3]) : bench::kDefaultBenchDepth, bench::kDefaultBenchTtSize);
    } else if (cmd == "evalcheck") {
        rc = evalCheck(argc > 3 ? static_cast<u32>(std::atoi(argv[3])) : 4, argc > 4 ? std::strtoull(argv[4], nullptr, 10) : 42);
    } else if (cmd == "datagen" && argc > 4) {
        const int seconds = std::atoi(argv[4]);
        std::thread timer{[seconds] {
            std::this_thread::sleep_for(std::chrono::seconds(seconds));
            std::raise(SIGINT); // datagen's own ctrl-c handler finishes the games in flight (datagen.cpp:50-55)
        }};
        rc = datagen::run([] {}, "viriformat", false, argv[3], 1, std::nullopt);
        timer.join();
    } else {
        eprintln("unknown command {}", cmd);
        rc = 2;
    }
    eval::shutdown();
    return rc;
}
