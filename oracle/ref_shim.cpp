/*
 * ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.
 *
 * A thin extern "C" surface over the UNMODIFIED reference sources under
 * /root/reference (Stormphrax 8.0.2).  It is compiled by oracle/Makefile together
 * with the reference's own .cpp files, where they lie, into oracle/_ref/*.so.
 * Nothing here re-implements the evaluation: every number that comes out of this
 * library is produced by the reference's own code paths
 *   eval::NnueState::evaluateOnce          (src/eval/nnue_state.cpp:612-634)
 *   NnueState::reset/push/applyImmediately/evaluate (src/eval/nnue_state.cpp:539-697)
 *   Position::applyMove<BoardObserver>     (src/position.cpp:109-197)
 *   generateAll                            (src/movegen.h:37)
 *   psq::featureIndex / threats::threatFeatureIndex / ppFeatureIndex
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline / --impl reference
 * legs may load this library.  The product (stormphrax_b200/) never does.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "attacks/attacks.h"
#include "datagen/viriformat.h"
#include "eval/eval.h"
#include "eval/header.h"
#include "eval/nnue.h"
#include "eval/nnue/features/psq.h"
#include "eval/nnue/features/threats.h"
#include "eval/nnue/loader.h"
#include "eval/nnue_state.h"
#include "movegen.h"
#include "opts.h"
#include "position.h"
#include "util/align.h"
#include "util/rng.h"
#include "wdl.h"

#include "../include/sp_types.h"

using namespace stormphrax;

namespace {
    std::byte* s_netData = nullptr;
    bool s_loaded = false;

    // marlinformat nibble -> FEN char, white upper-case (src/datagen/marlinformat.h:43-68)
    std::string toFen(const SpPackedBoard& b) {
        char board[64];
        std::memset(board, 0, sizeof(board));
        u64 castleRooks = 0;
        u64 occ = b.occupancy;
        int i = 0;
        while (occ) {
            const int sq = __builtin_ctzll(occ);
            occ &= occ - 1;
            const u8 nib = (b.pieces[i / 2] >> ((i % 2) * 4)) & 0xF;
            ++i;
            u8 type = nib & 7;
            const bool black = (nib & 8) != 0;
            if (type == 6) {
                type = 3;
                castleRooks |= u64{1} << sq;
            }
            static const char kChars[] = "pnbrqk";
            char c = kChars[type];
            if (!black) {
                c = static_cast<char>(c - 'a' + 'A');
            }
            board[sq] = c;
        }
        std::string fen;
        for (int rank = 7; rank >= 0; --rank) {
            int empty = 0;
            for (int file = 0; file < 8; ++file) {
                const char c = board[rank * 8 + file];
                if (!c) {
                    ++empty;
                    continue;
                }
                if (empty) {
                    fen += static_cast<char>('0' + empty);
                    empty = 0;
                }
                fen += c;
            }
            if (empty) {
                fen += static_cast<char>('0' + empty);
            }
            if (rank) {
                fen += '/';
            }
        }
        const bool blackToMove = (b.stm_ep & 0x80) != 0;
        fen += blackToMove ? " b " : " w ";
        std::string castling;
        for (int file = 7; file >= 0; --file) { // white first, file letters (Shredder-FEN)
            if (castleRooks & (u64{1} << file)) {
                castling += static_cast<char>('A' + file);
            }
        }
        for (int file = 7; file >= 0; --file) {
            if (castleRooks & (u64{1} << (56 + file))) {
                castling += static_cast<char>('a' + file);
            }
        }
        fen += castling.empty() ? "-" : castling;
        const int ep = b.stm_ep & 0x7F;
        if (ep < 64) {
            fen += ' ';
            fen += static_cast<char>('a' + (ep & 7));
            fen += static_cast<char>('1' + (ep >> 3));
        } else {
            fen += " -";
        }
        fen += ' ';
        fen += std::to_string(b.halfmove);
        fen += ' ';
        fen += std::to_string(b.fullmove ? b.fullmove : 1);
        return fen;
    }

    bool toPosition(const SpPackedBoard& b, Position& out) {
        auto pos = Position::fromFen(toFen(b));
        if (!pos) {
            return false;
        }
        out = *pos;
        return true;
    }

    // Same record the reference's datagen writes (marlinformat.h:43-84), eval/wdl zeroed.
    SpPackedBoard pack(const Position& pos) {
        SpPackedBoard out{};
        const auto rooks = pos.castlingRooks();
        const auto occ = pos.occ();
        out.occupancy = occ;
        usize i = 0;
        for (const auto sq : occ) {
            const auto piece = pos.pieceOn(sq);
            u8 pt = piece.type().raw();
            if (piece.type() == PieceTypes::kRook
                && (sq == rooks.black().kingside || sq == rooks.black().queenside || sq == rooks.white().kingside
                    || sq == rooks.white().queenside))
            {
                pt = 6;
            }
            const u8 nib = pt | (piece.color() == Colors::kBlack ? 8 : 0);
            out.pieces[i / 2] |= static_cast<u8>(nib << ((i % 2) * 4));
            ++i;
        }
        const u8 stm = pos.stm() == Colors::kBlack ? 0x80 : 0;
        const auto ep = pos.enPassant() == Squares::kNone
                          ? Squares::kNone
                          : pos.enPassant().withRank(pos.stm() == Colors::kBlack ? kRank3 : kRank6);
        out.stm_ep = stm | ep.raw();
        out.halfmove = static_cast<u8>(std::min<u32>(pos.halfmove(), 255));
        out.fullmove = static_cast<u16>(pos.fullmove());
        return out;
    }

    Move moveFromRaw(u16 raw) {
        static_assert(sizeof(Move) == sizeof(u16));
        Move m;
        std::memcpy(&m, &raw, sizeof(m));
        return m;
    }
} // namespace

extern "C" {

// 512 when the reference's AVX-512 path was compiled in, else 256 (AVX2)
int spref_isa(void) {
#if SP_HAS_AVX512
    return 512;
#else
    return 256;
#endif
}

// Load a LOGICAL (un-permuted) CBNF network image: 64-byte header + raw arrays
// (src/eval/header.h:38-52, array order src/eval/nnue/input.h:359-361, multilayer.h:492-496).
// The reference's own loader permutes the FT for x86 pack order (network.h:58-67).
int spref_load_net(const void* bytes, size_t len) {
    opts::mutableOpts().chess960 = true; // accept Shredder-FEN castling letters everywhere
    const auto need = eval::Network::byteSize();
    if (len < sizeof(eval::NetworkHeader) + need) {
        return 1;
    }
    if (!s_netData) {
        s_netData = util::alignedAlloc<std::byte>(64, need);
    }
    std::memcpy(s_netData, static_cast<const std::byte*>(bytes) + sizeof(eval::NetworkHeader), need);
    eval::nnue::NetworkLoader loader{s_netData, need};
    auto* net = const_cast<eval::Network*>(eval::getNetwork(0));
    if (!net->loadFrom(loader, /*prePermuted=*/false)) {
        return 2;
    }
    s_loaded = true;
    return 0;
}

// evaluateOnce over n packed boards, single thread. stm taken from the record.
int spref_eval_once(const SpPackedBoard* boards, size_t n, int32_t* out) {
    if (!s_loaded) {
        return 1;
    }
    for (size_t i = 0; i < n; ++i) {
        Position pos;
        if (!toPosition(boards[i], pos)) {
            return 2;
        }
        out[i] = eval::NnueState::evaluateOnce(pos, pos.stm());
    }
    return 0;
}

// The engine's post-processed static eval of n boards, produced by the reference's own functions:
// eval::staticEvalOnce (contempt + clamp, src/eval/eval.cpp:25-28,109-112) followed by
// eval::adjustEval<false> (src/eval/eval.cpp:31-67).  raw_out (optional) receives evaluateOnce.
int spref_adjusted_eval(
    const SpPackedBoard* boards, size_t n, const int32_t contempt[2], const int32_t optimism[2], int32_t* raw_out, int32_t* out
) {
    if (!s_loaded) {
        return 1;
    }
    const eval::Contempt c{contempt[0], contempt[1]};
    const eval::Optimism o{optimism[0], optimism[1]};
    for (size_t i = 0; i < n; ++i) {
        Position pos;
        if (!toPosition(boards[i], pos)) {
            return 2;
        }
        if (raw_out) {
            raw_out[i] = eval::NnueState::evaluateOnce(pos, pos.stm());
        }
        out[i] = eval::adjustEval<false>(pos, o, {}, nullptr, eval::staticEvalOnce(pos, c));
    }
    return 0;
}

// Timed evaluateOnce: positions are parsed first (untimed), then `threads` std::threads
// each evaluate a contiguous shard `reps` times. Returns the best wall-clock seconds of one
// full pass over all n positions (max over threads per rep), or a negative value on error.
double spref_time_eval_once(const SpPackedBoard* boards, size_t n, int threads, int reps, int32_t* out) {
    if (!s_loaded || threads < 1 || reps < 1) {
        return -1.0;
    }
    std::vector<Position> positions(n);
    for (size_t i = 0; i < n; ++i) {
        if (!toPosition(boards[i], positions[i])) {
            return -2.0;
        }
    }
    double best = 1e30;
    for (int rep = 0; rep < reps; ++rep) {
        std::vector<std::thread> pool;
        std::atomic<int> ready{0};
        std::atomic<bool> go{false};
        std::vector<double> secs(threads, 0.0);
        for (int t = 0; t < threads; ++t) {
            pool.emplace_back([&, t] {
                const size_t lo = n * t / threads;
                const size_t hi = n * (t + 1) / threads;
                ++ready;
                while (!go.load(std::memory_order_acquire)) {}
                const auto t0 = std::chrono::steady_clock::now();
                for (size_t i = lo; i < hi; ++i) {
                    out[i] = eval::NnueState::evaluateOnce(positions[i], positions[i].stm());
                }
                secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            });
        }
        while (ready.load() < threads) {}
        go.store(true, std::memory_order_release);
        for (auto& th : pool) {
            th.join();
        }
        best = std::min(best, *std::max_element(secs.begin(), secs.end()));
    }
    return best;
}

// Random legal playouts with the reference's own move generator and RNG
// (movegen.h:37, util/rng.h:28-125; same scheme as SURVEY.md section 8d):
// game g uses Jsf64Rng(SeedGenerator(seed).nextSeed() #g); each ply picks
// legal[rng.nextU32(count)]. A game stops at max_plies, at (stale)mate, or at bare kings.
// Every position INCLUDING the start position is emitted; moves[i] is the move played
// FROM boards[i] (0 for the last board of a game). Returns number of boards written.
size_t spref_playouts(
    uint64_t seed,
    uint32_t n_games,
    uint32_t max_plies,
    int dfrc,
    SpPackedBoard* out_boards,
    uint16_t* out_moves,
    uint32_t* out_game_start, // n_games + 1 offsets into out_boards
    size_t cap
) {
    opts::mutableOpts().chess960 = true;
    util::rng::SeedGenerator seeds{seed};
    size_t n = 0;
    for (u32 g = 0; g < n_games; ++g) {
        util::rng::Jsf64Rng rng{seeds.nextSeed()};
        out_game_start[g] = static_cast<u32>(n);
        auto pos = dfrc ? *Position::fromDfrcIndex(rng.nextU32(960 * 960)) : Position::startpos();
        for (u32 ply = 0;; ++ply) {
            if (n >= cap) {
                out_game_start[g + 1] = static_cast<u32>(n);
                for (u32 r = g + 1; r < n_games; ++r) {
                    out_game_start[r + 1] = static_cast<u32>(n);
                }
                return n;
            }
            out_boards[n] = pack(pos);
            out_moves[n] = 0;
            ++n;
            if (ply >= max_plies || pos.occ().popcount() <= 2) {
                break;
            }
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
            out_moves[n - 1] = move.data();
            pos = pos.applyMove(move);
        }
        out_game_start[g + 1] = static_cast<u32>(n);
    }
    return n;
}

// Incremental evaluation along one playout, the way the engine does it.
//   mode 0: datagen form -- applyMove(BoardObserver{ctx}) + applyImmediately + evaluate
//           (src/datagen/datagen.cpp:257-262)
//   mode 1: search form  -- applyMove(nnueState.push()) then lazy evaluate every `stride` plies
//           (src/thread.cpp:46-67, nnue_state.cpp:636-697); plies not evaluated get INT32_MIN
// out[0] is the eval of the start board, out[i] the eval after moves[i-1]; all stm-relative.
int spref_eval_playout(
    const SpPackedBoard* start,
    const uint16_t* moves,
    uint32_t n_moves,
    int mode,
    int stride,
    int32_t* out
) {
    if (!s_loaded) {
        return 1;
    }
    Position pos;
    if (!toPosition(*start, pos)) {
        return 2;
    }
    eval::NnueState state;
    state.setNetwork(eval::getNetwork(0));
    state.reset(pos);
    out[0] = state.evaluate(pos, pos.stm());
    if (stride < 1) {
        stride = 1;
    }
    for (u32 i = 0; i < n_moves; ++i) {
        const auto move = moveFromRaw(moves[i]);
        if (mode == 0) {
            eval::UpdateContext ctx{};
            pos = pos.applyMove(move, eval::BoardObserver{ctx});
            state.applyImmediately(ctx, pos);
            out[i + 1] = state.evaluate(pos, pos.stm());
        } else {
            if (i >= 200) {
                return 3; // accumulator stack is 256 deep (nnue_state.h:88)
            }
            pos = pos.applyMove(move, state.push());
            out[i + 1] = ((i + 1) % stride == 0 || i + 1 == n_moves) ? state.evaluate(pos, pos.stm()) : INT32_MIN;
        }
    }
    return 0;
}

// Timed incremental playouts (datagen form), sharded by game over `threads` threads.
// boards/moves/game_start as produced by spref_playouts. out gets one eval per board.
double spref_time_playouts(
    const SpPackedBoard* boards,
    const uint16_t* moves,
    const uint32_t* game_start,
    uint32_t n_games,
    int threads,
    int reps,
    int32_t* out
) {
    if (!s_loaded || threads < 1 || reps < 1) {
        return -1.0;
    }
    std::vector<Position> starts(n_games);
    for (u32 g = 0; g < n_games; ++g) {
        if (!toPosition(boards[game_start[g]], starts[g])) {
            return -2.0;
        }
    }
    double best = 1e30;
    for (int rep = 0; rep < reps; ++rep) {
        std::vector<std::thread> pool;
        std::atomic<int> ready{0};
        std::atomic<bool> go{false};
        std::vector<double> secs(threads, 0.0);
        for (int t = 0; t < threads; ++t) {
            pool.emplace_back([&, t] {
                const u32 lo = static_cast<u32>(u64{n_games} * t / threads);
                const u32 hi = static_cast<u32>(u64{n_games} * (t + 1) / threads);
                eval::NnueState state;
                state.setNetwork(eval::getNetwork(0));
                ++ready;
                while (!go.load(std::memory_order_acquire)) {}
                const auto t0 = std::chrono::steady_clock::now();
                for (u32 g = lo; g < hi; ++g) {
                    auto pos = starts[g];
                    state.reset(pos);
                    u32 i = game_start[g];
                    out[i] = state.evaluate(pos, pos.stm());
                    for (; i + 1 < game_start[g + 1]; ++i) {
                        eval::UpdateContext ctx{};
                        pos = pos.applyMove(moveFromRaw(moves[i]), eval::BoardObserver{ctx});
                        state.applyImmediately(ctx, pos);
                        out[i + 1] = state.evaluate(pos, pos.stm());
                    }
                }
                secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            });
        }
        while (ready.load() < threads) {}
        go.store(true, std::memory_order_release);
        for (auto& th : pool) {
            th.join();
        }
        best = std::min(best, *std::max_element(secs.begin(), secs.end()));
    }
    return best;
}

// PSQ feature indices of perspective c, in the reference's board-iteration order
// (nnue_state.cpp:440-449, psq.h:338-365). Returns the count (<= 32) or -1.
int spref_psq_features(const SpPackedBoard* board, int c, uint32_t* out) {
    Position pos;
    if (!toPosition(*board, pos)) {
        return -1;
    }
    const auto color = c == 0 ? Colors::kBlack : Colors::kWhite;
    const auto king = pos.king(color);
    int n = 0;
    for (const auto [piece, sq] : pos) {
        out[n++] = eval::nnue::features::psq::featureIndex<eval::InputFeatureSet>(color, piece, sq, king);
    }
    return n;
}

// Threat + pawn-pair feature indices of perspective c, enumerated exactly as
// addThreatFeatures does (nnue_state.cpp:309-354). Returns the count (<= 256) or -1.
int spref_threat_features(const SpPackedBoard* board, int c, uint32_t* out) {
    using namespace eval::nnue::features::threats;
    Position pos;
    if (!toPosition(*board, pos)) {
        return -1;
    }
    const auto color = c == 0 ? Colors::kBlack : Colors::kWhite;
    const auto kingSq = pos.king(color);
    const auto occ = pos.occ();
    const auto kings = pos.bb(PieceTypes::kKing);
    int n = 0;
    for (const auto from : occ & ~kings) {
        const auto piece = pos.pieceOn(from);
        for (const auto to : occ & attacks::getAttacks(piece, from, occ) & ~kings) {
            const auto feature = threatFeatureIndex(color, kingSq, piece, from, pos.pieceOn(to), to);
            if (feature >= 0) {
                out[n++] = static_cast<u32>(feature);
            }
        }
    }
    const auto ourPawns = pos.bb(PieceTypes::kPawn, color);
    const auto theirPawns = pos.bb(PieceTypes::kPawn, color.flip());
    for (const auto [a, remaining] : ourPawns.iterWithRemaining()) {
        const auto mask = kPpMasks[a.idx()];
        for (const auto b : remaining & mask) {
            out[n++] = ppFeatureIndex(color, kingSq, color, a, color, b);
        }
        for (const auto b : theirPawns & mask) {
            out[n++] = ppFeatureIndex(color, kingSq, color, a, color.flip(), b);
        }
    }
    for (const auto [a, remaining] : theirPawns.iterWithRemaining()) {
        const auto mask = kPpMasks[a.idx()];
        for (const auto b : remaining & mask) {
            out[n++] = ppFeatureIndex(color, kingSq, color.flip(), a, color.flip(), b);
        }
    }
    return n;
}

// Raw threatFeatureIndex (threats.cpp:170-198) for table-level tests.
int32_t spref_threat_index(int c, int kingSq, int attacker, int attackerSq, int attacked, int attackedSq) {
    return eval::nnue::features::threats::threatFeatureIndex(
        c == 0 ? Colors::kBlack : Colors::kWhite,
        Square::fromRaw(static_cast<u8>(kingSq)),
        Piece::fromRaw(static_cast<u8>(attacker)),
        Square::fromRaw(static_cast<u8>(attackerSq)),
        Piece::fromRaw(static_cast<u8>(attacked)),
        Square::fromRaw(static_cast<u8>(attackedSq))
    );
}

// Legal moves of a board via generateAll + isLegal. Returns the count.
int spref_legal_moves(const SpPackedBoard* board, uint16_t* out) {
    opts::mutableOpts().chess960 = true;
    Position pos;
    if (!toPosition(*board, pos)) {
        return -1;
    }
    ScoredMoveList moves;
    generateAll(moves, pos);
    int n = 0;
    for (const auto [move, score] : moves) {
        if (pos.isLegal(move)) {
            out[n++] = move.data();
        }
    }
    return n;
}

int spref_apply_move(const SpPackedBoard* board, uint16_t move, SpPackedBoard* out) {
    opts::mutableOpts().chess960 = true;
    Position pos;
    if (!toPosition(*board, pos)) {
        return 1;
    }
    *out = pack(pos.applyMove(moveFromRaw(move)));
    return 0;
}

int spref_board_from_fen(const char* fen, SpPackedBoard* out) {
    opts::mutableOpts().chess960 = true;
    auto pos = Position::fromFen(fen);
    if (!pos) {
        return 1;
    }
    *out = pack(*pos);
    return 0;
}

// wdl::wdlModel(povScore, pos.classicalMaterial()) (src/wdl.cpp:43-50)
int spref_wdl_model(const SpPackedBoard* board, int32_t score, int32_t* win, int32_t* loss) {
    opts::mutableOpts().chess960 = true;
    Position pos;
    if (!toPosition(*board, pos)) {
        return 1;
    }
    const auto [w, l] = wdl::wdlModel(score, pos.classicalMaterial());
    *win = w;
    *loss = l;
    return 0;
}

// Position::fromDfrcIndex (src/position.cpp:1215-1270)
int spref_board_from_dfrc(uint32_t index, SpPackedBoard* out) {
    opts::mutableOpts().chess960 = true;
    const auto pos = Position::fromDfrcIndex(index);
    if (!pos) {
        return 1;
    }
    *out = pack(*pos);
    return 0;
}

// One game through the reference's own Viriformat writer (src/datagen/viriformat.cpp:27-63): start(),
// push() per (move, score), writeAllWithOutcome().  Returns the bytes written, or -1.
long spref_viriformat(
    const SpPackedBoard* start, const uint16_t* moves, const int16_t* scores, uint32_t n, int outcome, uint8_t* out, size_t cap
) {
    opts::mutableOpts().chess960 = true;
    Position pos;
    if (!toPosition(*start, pos)) {
        return -1;
    }
    datagen::Viriformat game{};
    game.start(pos);
    for (uint32_t i = 0; i < n; ++i) {
        game.push(false, moveFromRaw(moves[i]), scores[i]);
    }
    std::ostringstream stream{std::ios::binary};
    game.writeAllWithOutcome(stream, static_cast<datagen::Outcome>(outcome));
    const std::string bytes = stream.str();
    if (bytes.size() > cap) {
        return -1;
    }
    std::memcpy(out, bytes.data(), bytes.size());
    return static_cast<long>(bytes.size());
}

// wdl::normalizeScore<false>(score, pos.classicalMaterial()) (src/wdl.cpp:52-75, src/position.h:515-523)
int spref_normalize_score(const SpPackedBoard* board, int32_t score, int32_t* material, int32_t* normalized) {
    opts::mutableOpts().chess960 = true;
    Position pos;
    if (!toPosition(*board, pos)) {
        return 1;
    }
    *material = pos.classicalMaterial();
    *normalized = wdl::normalizeScore<false>(score, *material);
    return 0;
}

} // extern "C"
