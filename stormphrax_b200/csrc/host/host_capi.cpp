/*
 * host_capi.cpp -- extern "C" exports of the host-side helpers (no GPU involved):
 * board conversion, legal move generation, random legal playouts (the synthetic workload of
 * BASELINE.json configs 2 and 3) and CPU evaluation of the shared feature-index code
 * (sp_features.h) so that it can be unit-tested without a GPU.
 * Declared in include/sp_nnue.h ("host utilities").
 */
#include <algorithm>
#include <array>
#include <cstring>
#include <thread>
#include <vector>

#include "../../../include/sp_nnue.h"
#include "../sp_delta.h"
#include "../sp_features.h"
#include "nnue_state.h"
#include "position.h"
#include "rng.h"

using namespace sp;
using namespace sp::host;

namespace {

const FeatureTables& tables() {
    static const FeatureTables t = [] {
        FeatureTables x;
        build_feature_tables(x);
        return x;
    }();
    return t;
}

/* One game: emits every position including the start; moves[i] is the move played from boards[i]. */
size_t play_game(uint64_t seed, uint32_t max_plies, SpPackedBoard* boards, SpMove* moves) {
    Jsf64 rng{seed};
    Position pos = Position::startpos();
    size_t n = 0;
    for (uint32_t ply = 0;; ++ply) {
        boards[n] = pos.pack();
        moves[n] = 0;
        ++n;
        if (ply >= max_plies || popcount64(pos.occ()) <= 2) break;
        Move legal[256];
        const int count = pos.generateLegal(legal);
        if (!count) break;
        const Move m = legal[rng.below(static_cast<uint32_t>(count))];
        moves[n - 1] = m.raw;
        pos = pos.applyMove(m);
    }
    return n;
}

} // namespace

extern "C" {

/* see include/sp_nnue.h */
size_t sp_host_playouts(
    uint64_t seed,
    uint32_t n_games,
    uint32_t max_plies,
    int threads,
    SpPackedBoard* boards,
    SpMove* moves,
    uint32_t* game_start
) {
    if (threads < 1) threads = 1;
    const size_t stride = static_cast<size_t>(max_plies) + 1;
    std::vector<uint64_t> seeds(n_games);
    SplitMix64 sm{seed};
    for (auto& s : seeds) s = sm.next();
    std::vector<uint32_t> len(n_games);
    /* every game writes into its own fixed-stride window, then the windows are compacted in order,
     * so the result does not depend on the thread count */
    std::vector<SpPackedBoard> tmpBoards(static_cast<size_t>(n_games) * stride);
    std::vector<SpMove> tmpMoves(static_cast<size_t>(n_games) * stride);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t] {
            for (uint32_t g = static_cast<uint32_t>(t); g < n_games; g += static_cast<uint32_t>(threads))
                len[g] = static_cast<uint32_t>(play_game(seeds[g], max_plies, &tmpBoards[g * stride], &tmpMoves[g * stride]));
        });
    }
    for (auto& th : pool) th.join();
    size_t n = 0;
    for (uint32_t g = 0; g < n_games; ++g) {
        game_start[g] = static_cast<uint32_t>(n);
        std::memcpy(&boards[n], &tmpBoards[g * stride], len[g] * sizeof(SpPackedBoard));
        std::memcpy(&moves[n], &tmpMoves[g * stride], len[g] * sizeof(SpMove));
        n += len[g];
    }
    game_start[n_games] = static_cast<uint32_t>(n);
    return n;
}

int sp_host_board_from_fen(const char* fen, SpPackedBoard* out) {
    Position p;
    if (!Position::fromFen(fen, p)) return SP_ERR_BAD_BOARD;
    *out = p.pack();
    return SP_OK;
}

int sp_host_board_to_fen(const SpPackedBoard* board, char* out, size_t cap) {
    Position p;
    if (!Position::fromPacked(*board, p)) return SP_ERR_BAD_BOARD;
    const std::string fen = p.toFen();
    if (fen.size() + 1 > cap) return SP_ERR_INVALID;
    std::memcpy(out, fen.c_str(), fen.size() + 1);
    return SP_OK;
}

int sp_host_legal_moves(const SpPackedBoard* board, SpMove* out) {
    Position p;
    if (!Position::fromPacked(*board, p)) return -1;
    Move legal[256];
    const int n = p.generateLegal(legal);
    for (int i = 0; i < n; ++i) out[i] = legal[i].raw;
    return n;
}

int sp_host_board_from_dfrc(uint32_t index, SpPackedBoard* out) {
    if (!out || index >= 960u * 960u) return SP_ERR_INVALID;
    *out = Position::fromDfrcIndex(index).pack();
    return SP_OK;
}

/* adjustStatic + adjustEval of the C++ mirror (host/nnue_state.h) on an array: the host counterpart of
 * sp_nnue_adjust, used by the tests */
int sp_host_adjust(const SpPackedBoard* boards, const int32_t* raw, const int32_t* correction, size_t n, const SpAdjustParams* params, int32_t* out) {
    if (!boards || !raw || !params || !out) return SP_ERR_INVALID;
    for (size_t i = 0; i < n; ++i) {
        const int stm = (boards[i].stm_ep & 0x80) ? kBlack : kWhite;
        eval::Contempt contempt;
        contempt.value[0] = params->contempt[0], contempt.value[1] = params->contempt[1];
        eval::Optimism optimism;
        optimism.value[0] = params->optimism[0], optimism.value[1] = params->optimism[1];
        const int32_t adjusted = eval::adjustStatic(raw[i], stm, contempt);
        out[i] = eval::adjustEvalPacked(boards[i], optimism, adjusted, correction ? correction[i] : 0, *params);
    }
    return SP_OK;
}

int sp_host_in_check(const SpPackedBoard* board) {
    Position p;
    if (!Position::fromPacked(*board, p)) return -1;
    return p.isCheck() ? 1 : 0;
}

int sp_host_apply_move(const SpPackedBoard* board, SpMove move, SpPackedBoard* out) {
    Position p;
    if (!Position::fromPacked(*board, p)) return SP_ERR_BAD_BOARD;
    *out = p.applyMove(Move{move}).pack();
    return SP_OK;
}

/* Feature index lists of perspective c computed by the SAME code the kernels run
 * (sp_features.h), on the CPU. kind 0 = PSQ, 1 = threats + pawn pairs. Returns the count. */
int sp_host_features(const SpPackedBoard* board, int c, int kind, uint32_t* out) {
    Board b;
    if (unpack_board(*board, b)) return -1;
    int n = 0;
    auto emit = [&](int persp, uint32_t idx) {
        if (persp == c) out[n++] = idx;
    };
    for (int sq = 0; sq < 64; ++sq) {
        if (kind == 0)
            square_psq_features(tables(), b, sq, emit);
        else
            square_threat_features(tables(), b, sq, emit);
    }
    return n;
}

int sp_host_feature_counts(const SpPackedBoard* boards, size_t n, int threads, uint64_t out[3]) {
    if (threads < 1) threads = 1;
    std::vector<std::array<uint64_t, 4>> part(static_cast<size_t>(threads), {0, 0, 0, 0});
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t] {
            auto& p = part[static_cast<size_t>(t)];
            for (size_t i = n * t / threads; i < n * (t + 1) / threads; ++i) {
                Board b;
                if (unpack_board(boards[i], b)) {
                    p[3] = 1;
                    continue;
                }
                for (int sq = 0; sq < 64; ++sq) {
                    square_psq_features(tables(), b, sq, [&](int, uint32_t) { ++p[0]; });
                    square_threat_features(tables(), b, sq, [&](int, uint32_t idx) { ++p[idx < SP_PP_FEATURES ? 2 : 1]; });
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    out[0] = out[1] = out[2] = 0;
    bool bad = false;
    for (const auto& p : part) {
        out[0] += p[0], out[1] += p[1], out[2] += p[2];
        bad |= p[3] != 0;
    }
    return bad ? SP_ERR_BAD_BOARD : SP_OK;
}

/* What the incremental walker must read for a playout stream (roofline bookkeeping): per perspective
 * of every board either the delta rows against the previous board of its game, or -- first board of
 * a game / king changed bucket or side / too many changed squares -- its full row lists.
 * out = {psq_delta_rows, threat_delta_rows, rebuild_psq_rows, rebuild_threat_rows,
 *        updated_perspectives, rebuilt_perspectives}. */
int sp_host_playout_stats(const SpPackedBoard* boards, const uint32_t* game_start, uint32_t n_games, int threads, uint64_t out[6]) {
    if (threads < 1) threads = 1;
    std::vector<std::array<uint64_t, 7>> part(static_cast<size_t>(threads), {0, 0, 0, 0, 0, 0, 0});
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t) {
        pool.emplace_back([&, t] {
            auto& p = part[static_cast<size_t>(t)];
            for (uint32_t g = static_cast<uint32_t>(t); g < n_games; g += static_cast<uint32_t>(threads)) {
                Board prev{}, cur{};
                for (uint32_t i = game_start[g]; i < game_start[g + 1]; ++i) {
                    if (unpack_board(boards[i], cur)) {
                        p[6] = 1;
                        continue;
                    }
                    const bool first = i == game_start[g];
                    const uint64_t changed = first ? 0 : changed_squares(prev, cur);
                    for (int c = 0; c < 2; ++c) {
                        const bool rebuild = first || popcount64(changed) > kMaxChanged || needs_refresh(tables(), prev, cur, c);
                        if (rebuild) {
                            ++p[5];
                            for (int sq = 0; sq < 64; ++sq) {
                                square_psq_features(tables(), cur, sq, [&](int pc, uint32_t) { p[2] += pc == c; });
                                square_threat_features(tables(), cur, sq, [&](int pc, uint32_t) { p[3] += pc == c; });
                            }
                        } else {
                            ++p[4];
                            delta_all(tables(), prev, cur, changed, [&](int pc, int kind, int, uint32_t) {
                                if (pc == c) ++p[kind == 0 ? 0 : 1];
                            });
                        }
                    }
                    prev = cur;
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    bool bad = false;
    for (int k = 0; k < 6; ++k) out[k] = 0;
    for (const auto& p : part) {
        for (int k = 0; k < 6; ++k) out[k] += p[static_cast<size_t>(k)];
        bad |= p[6] != 0;
    }
    return bad ? SP_ERR_BAD_BOARD : SP_OK;
}

/* Feature deltas between two boards for perspective c, computed by the shared delta generator
 * (sp_delta.h) on the CPU. Returns 0, or 1 if the perspective needs a full refresh. */
int sp_host_feature_delta(
    const SpPackedBoard* before,
    const SpPackedBoard* after,
    int c,
    uint32_t* psq_add, int* n_psq_add,
    uint32_t* psq_sub, int* n_psq_sub,
    uint32_t* thr_add, int* n_thr_add,
    uint32_t* thr_sub, int* n_thr_sub
) {
    Board bb, ba;
    if (unpack_board(*before, bb) || unpack_board(*after, ba)) return -1;
    *n_psq_add = *n_psq_sub = *n_thr_add = *n_thr_sub = 0;
    if (needs_refresh(tables(), bb, ba, c)) return 1;
    const uint64_t changed = changed_squares(bb, ba);
    auto emit = [&](int persp, int kind, int sign, uint32_t idx) {
        if (persp != c) return;
        if (kind == 0) {
            if (sign > 0) psq_add[(*n_psq_add)++] = idx; else psq_sub[(*n_psq_sub)++] = idx;
        } else {
            if (sign > 0) thr_add[(*n_thr_add)++] = idx; else thr_sub[(*n_thr_sub)++] = idx;
        }
    };
    if (popcount64(changed) > kMaxChanged) return 1;
    delta_all(tables(), bb, ba, changed, emit);
    return 0;
}

} // extern "C"
