/*
 * test_movegen.cpp -- the pin-aware legal move generator (Position::generateLegal) against the
 * make-and-test form (generateLegalSlow): same moves in the same order on every position of many random
 * playouts (both are also held against the reference's generator in tests/test_host.py).
 *   usage: test_movegen   (exit code 0 = ok)
 */
#include <cstdio>

#include "../../stormphrax_b200/csrc/host/position.h"
#include "../../stormphrax_b200/csrc/host/rng.h"

using namespace sp::host;

int main() {
    long positions = 0, moves = 0, failures = 0, checks = 0;
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
