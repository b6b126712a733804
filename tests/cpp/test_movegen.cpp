/*
 * test_movegen.cpp -- the pin-aware legal move generator (Position::generateLegal) against the
 * make-and-test form (generateLegalSlow): same moves in the same order on every position of many random
 * playouts (both are also held against the reference's generator in tests/test_host.py), and the published
 * perft counts of the six standard test positions.
 *   usage: test_movegen   (exit code 0 = ok)
 */
#include <cstdio>

#include "../../stormphrax_b200/csrc/host/position.h"
#include "../../stormphrax_b200/csrc/host/rng.h"

using namespace sp::host;

/* perft: the number of leaf nodes of the legal move tree -- the standard movegen cross-check (published
 * counts from the chessprogramming wiki's perft results page) */
static long perft(const Position& pos, int depth) {
    Move moves[256];
    const int n = pos.generateLegal(moves);
    if (depth == 1) return n;
    long total = 0;
    for (int i = 0; i < n; ++i) total += perft(pos.applyMove(moves[i]), depth - 1);
    return total;
}

static long check_perft() {
    struct Case { const char* fen; int depth; long nodes; };
    static const Case kCases[] = {
        {"rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", 5, 4865609},
        {"r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1", 4, 4085603},
        {"8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - - 0 1", 5, 674624},
        {"r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq - 0 1", 4, 422333},
        {"rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ - 1 8", 4, 2103487},
        {"r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - - 0 10", 4, 3894594},
    };
    long failures = 0;
    for (const Case& c : kCases) {
        Position pos;
        if (!Position::fromFen(c.fen, pos)) {
            std::fprintf(stderr, "FAIL: cannot parse %s\n", c.fen);
            ++failures;
            continue;
        }
        const long got = perft(pos, c.depth);
        if (got != c.nodes) {
            std::fprintf(stderr, "FAIL perft(%d) of %s: %ld, expected %ld\n", c.depth, c.fen, got, c.nodes);
            ++failures;
        }
    }
    std::printf("perft: %zu positions, %ld failures\n", sizeof(kCases) / sizeof(kCases[0]), failures);
    return failures;
}

int main() {
    long positions = 0, moves = 0, failures = check_perft(), checks = 0;
    for (int game = 0; game < 3000; ++game) {
        Jsf64 rng{0xC0FFEEull + static_cast<uint64_t>(game)};
        Position pos = Position::startpos();
        for (int ply = 0; ply < 120; ++ply) {
            Move fast[256], slow[256];
            const int nf = pos.generateLegal(fast), ns = pos.generateLegalSlow(slow);
            ++positions, moves += ns, checks += pos.isCheck();
            bool same = nf == ns;
            for (int i = 0; same && i < nf; ++i) same = fast[i] == slow[i];
            if (!same) {
                ++failures;
                std::fprintf(stderr, "FAIL %s: %d moves, make-and-test says %d\n", pos.toFen().c_str(), nf, ns);
            }
            if (!ns) break;
            pos = pos.applyMove(slow[rng.below(static_cast<uint32_t>(ns))]);
        }
    }
    std::printf("%ld positions (%ld in check), %ld moves, %ld failures\n", positions, checks, moves, failures);
    return failures ? 1 : 0;
}
