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
 *   <binary> <network.nnue> searches <n> <depth> <fibers 0|1> [threads [width]]
 *                                                   n independent fixed-depth searches (own Searcher each, roots = random playouts):
 *                                                   one after the other, or (sp_engine_b200, fibers = 1) as fibers of one thread whose
 *                                                   evaluations are answered in device batches; prints the node count of each search
 *   <binary> <network.nnue> games <n> <nodes> <plies> <seed> <fibers 0|1> [threads [width]]
 *                                                   BASELINE configs[4] in miniature: n concurrent self-play games, datagen's per-move
 *                                                   search (runDatagenSearch, soft node limit <nodes>), <plies> moves each; prints
 *                                                   nodes, nodes/s, a checksum of (move, score) and the batch statistics
 *                                                   [threads] (with fibers = 1): that many host threads share the searches / games, each
 *                                                   a fiber scheduler with its own evaluator context on the GPU (sp_engine_cpu: plain
 *                                                   host threads over its CPU evaluation) -- datagen's "N threads" (datagen.cpp:378-384);
 *                                                   [width]: searches / games alive per scheduler, the rest wait in a queue shared by
 *                                                   all schedulers (0 = all at once)
 *
 * Bit-exact evaluations make both binaries walk the same search trees: their bench node counts must be identical.
 */
#include <atomic>
#include <chrono>
#include <csignal>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>
#include <thread>
#include <vector>

#include <functional>
#include <memory>

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

#ifdef SP_EVAL_B200
    #include "eval/batch.h"
#else
namespace stormphrax::eval {
    bool oracleLoadNetwork(const std::byte* payload, usize size);

    namespace batch { // the CPU build has nothing to batch: same entry point, jobs run one after the other
        struct Stats {
            u64 rounds{};
            u64 evaluations{};
        };
        inline void reserveStates(u32) {}
        inline Stats runFibers(std::vector<std::function<void()>>& jobs, usize = 0, u32 threads = 1, usize = 0) {
            std::atomic<usize> nextJob{0};
            const auto work = [&jobs, &nextJob] {
                for (auto job = nextJob.fetch_add(1); job < jobs.size(); job = nextJob.fetch_add(1)) {
                    jobs[job]();
                }
            };
            std::vector<std::thread> workers;
            for (u32 t = 1; t < threads; ++t) {
                workers.emplace_back(work);
            }
            work();
            for (auto& worker : workers) {
                worker.join();
            }
            return {};
        }
    } // namespace batch
} // namespace stormphrax::eval
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
    // a random legal playout from the start position (the reference's own generator and RNG): `plies` moves
    Position randomPosition(u64 seed, u32 plies, std::vector<u64>* keys = nullptr) {
        util::rng::Jsf64Rng rng{seed};
        auto pos = Position::startpos();
        for (u32 ply = 0; ply < plies; ++ply) {
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
            if (keys) {
                keys->push_back(pos.key());
            }
            pos = pos.applyMove(legal[rng.nextU32(static_cast<u32>(legal.size()))]);
        }
        return pos;
    }

    int searches(u32 n, i32 depth, bool fibers, u32 threads, usize width) {
        opts::mutableOpts().chess960 = false;
        util::rng::SeedGenerator seeds{1234};
        eval::batch::reserveStates(n);
        std::vector<std::unique_ptr<search::Searcher>> searchers(n);
        std::vector<usize> nodes(n, 0);
        std::vector<std::function<void()>> jobs;
        for (u32 i = 0; i < n; ++i) {
            searchers[i] = std::make_unique<search::Searcher>(1);
            auto& searcher = *searchers[i];
            searcher.setSilent(true);
            searcher.setLimiter(limit::SearchLimiter{util::Instant::now()});
            searcher.setMaxDepth(depth);
            auto& thread = searcher.take();
            searcher.newGame();
            thread.rootPos = randomPosition(seeds.nextSeed(), 16 + i % 24);
            jobs.emplace_back([&searcher, &nodes, i] {
                search::BenchData data{};
                searcher.runBenchSearch(data); // src/search.cpp:241-266: reset + searchRoot, as `bench` runs it
                nodes[i] = data.nodes;
            });
        }
        const auto start = util::Instant::now();
        eval::batch::Stats stats{};
        if (fibers) {
            stats = eval::batch::runFibers(jobs, usize{4} << 20, threads, width);
        } else {
            for (auto& job : jobs) {
                job();
            }
        }
        const auto seconds = start.elapsed();
        usize total = 0;
        print("search nodes:");
        for (const auto count : nodes) {
            print(" {}", count);
            total += count;
        }
        println();
        println(
            "searches: {} searches depth {} threads {} width {} {} nodes {:.3f} seconds {} nps rounds {} evaluations {}",
            n,
            depth,
            threads,
            width,
            total,
            seconds,
            static_cast<usize>(static_cast<f64>(total) / seconds),
            stats.rounds,
            stats.evaluations
        );
        return 0;
    }

    int games(u32 n, usize softNodes, u32 plies, u64 seed, bool fibers, u32 threads, usize width) {
        opts::mutableOpts().chess960 = false;
        util::rng::SeedGenerator seeds{seed};
        eval::batch::reserveStates(n);
        std::vector<std::unique_ptr<search::Searcher>> searchers(n);
        std::vector<usize> nodes(n, 0);
        std::vector<u64> sums(n, 0);
        std::vector<std::function<void()>> jobs;
        for (u32 i = 0; i < n; ++i) {
            searchers[i] = std::make_unique<search::Searcher>(1);
            auto& searcher = *searchers[i];
            searcher.setSilent(true);
            auto& thread = searcher.take();
            thread.datagen = true;
            limit::SearchLimiter limiter{util::Instant::now()}; // datagen.cpp:113-121
            limiter.setHardNodes(softNodes * 40);
            limiter.setSoftNodes(softNodes);
            searcher.setLimiter(limiter);
            searcher.setMaxDepth(kMaxDepth);
            searcher.newGame();
            thread.search = search::SearchData{};
            thread.keyHistory.clear();
            thread.rootPos = randomPosition(seeds.nextSeed(), 8 + i % 2, &thread.keyHistory); // datagen.cpp:153-171: 8 or 9 random plies
            jobs.emplace_back([&searcher, &thread, &nodes, &sums, i, plies] {
                auto& pos = thread.rootPos;
                thread.nnueState.reset(pos); // datagen.cpp:179
                for (u32 ply = 0; ply < plies; ++ply) {
                    { // a game that has ended (mate or stalemate on the board) has no root move to search
                        ScoredMoveList moves;
                        generateAll(moves, pos);
                        bool anyLegal = false;
                        for (const auto [move, score] : moves) {
                            anyLegal = anyLegal || pos.isLegal(move);
                        }
                        if (!anyLegal) {
                            break;
                        }
                    }
                    const auto [score, normScore] = searcher.runDatagenSearch(); // datagen.cpp:206
                    nodes[i] += thread.search.loadNodes();
                    thread.search = search::SearchData{};
                    const auto move = thread.rootMoves[0].pv.moves[0];
                    if (!move) {
                        break;
                    }
                    sums[i] = sums[i] * 0x100000001B3ull + (static_cast<u64>(move.data()) << 32 | static_cast<u32>(score));
                    eval::UpdateContext ctx{};
                    thread.keyHistory.push_back(pos.key());
                    pos = pos.applyMove(move, eval::BoardObserver{ctx}); // datagen.cpp:257-260
                    thread.nnueState.applyImmediately(ctx, pos);
                }
            });
        }
        const auto start = util::Instant::now();
        eval::batch::Stats stats{};
        if (fibers) {
            stats = eval::batch::runFibers(jobs, usize{4} << 20, threads, width);
        } else {
            for (auto& job : jobs) {
                job();
            }
        }
        const auto seconds = start.elapsed();
        usize total = 0;
        u64 checksum = 0;
        for (u32 i = 0; i < n; ++i) {
            total += nodes[i];
            checksum = checksum * 0x9E3779B97F4A7C15ull + sums[i];
        }
        println(
            "games: {} games {} plies soft {} threads {} width {} nodes: {} nodes {:.3f} seconds {} nps checksum {:016x} rounds {} evaluations {}",
            n,
            plies,
            softNodes,
            threads,
            width,
            total,
            seconds,
            static_cast<usize>(static_cast<f64>(total) / seconds),
            checksum,
            stats.rounds,
            stats.evaluations
        );
        return 0;
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
    } else if (cmd == "searches" && argc > 5) {
        rc = searches(
            static_cast<u32>(std::atoi(argv[3])),
            std::atoi(argv[4]),
            std::atoi(argv[5]) != 0,
            argc > 6 ? static_cast<u32>(std::max(1, std::atoi(argv[6]))) : 1,
            argc > 7 ? static_cast<usize>(std::max(0, std::atoi(argv[7]))) : 0
        );
    } else if (cmd == "games" && argc > 7) {
        rc = games(
            static_cast<u32>(std::atoi(argv[3])),
            static_cast<usize>(std::atol(argv[4])),
            static_cast<u32>(std::atoi(argv[5])),
            std::strtoull(argv[6], nullptr, 10),
            std::atoi(argv[7]) != 0,
            argc > 8 ? static_cast<u32>(std::max(1, std::atoi(argv[8]))) : 1,
            argc > 9 ? static_cast<usize>(std::max(0, std::atoi(argv[9]))) : 0
        );
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
