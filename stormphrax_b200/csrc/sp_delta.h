/*
 * sp_delta.h -- incremental feature deltas between two boards (host + device).
 *
 * The reference generates threat deltas inside Position::applyMove through BoardObserver
 * callbacks: a 64-lane "ray permutation" of the mailbox around each touched square
 * (src/eval/nnue.cpp:339-599, src/eval/nnue/features/threats/geometry*.h), plus pawn-pair rows
 * (src/eval/nnue_state.cpp:163-307) and PSQ add/sub pairs (nnue_state.cpp:34-87).
 *
 * Accumulators are sums in Z/2^16, so ANY list of (add, sub) rows whose multiset difference
 * equals features(after) - features(before) gives bit-identical accumulators (SURVEY.md
 * appendix D).  This file produces such a list from the two boards alone -- it never needs to
 * know which move was played, so captures, castling (incl. Chess960), en passant, promotions
 * and even several plies at once are all the same code path.
 *
 * Work decomposition.  D = squares whose content differs.  For every square s in D, for each of
 * the two boards X (before: rows are subtracted, after: rows are added), 17 independent items:
 *     0..7   one ray direction from s   (the reference's 8 rays)
 *     8..15  one knight offset from s   (slot 0 of each reference ray)
 *     16     PSQ row of the piece on s, and the pawn pairs of a pawn on s
 * A GPU lane takes items round-robin; the CPU tests run them in a loop.  Rules that make every
 * (attacker, victim) pair of X that touches D appear exactly once:
 *     - attacker on a D square           -> emitted by the attacker's item (ray or knight)
 *     - victim on D, attacker not on D   -> emitted by the victim's item
 *     - both off D, but a D square that is EMPTY in X lies strictly between them on the line
 *       (x-ray through the vacated / newly blocked square: the reference's "discovered"
 *       threats, nnue.cpp:424-484) -> emitted by the ray item of the D square closest to the
 *       attacker.
 */
#ifndef SP_DELTA_H
#define SP_DELTA_H

#include "sp_features.h"

namespace sp {

constexpr int kMaxChanged = 8;         /* more changed squares than this -> full refresh */
constexpr int kItemsPerSquareBoard = 17;
constexpr int kDeltaItems = kMaxChanged * 2 * kItemsPerSquareBoard;

SP_HD uint64_t changed_squares(const Board& before, const Board& after) {
    uint64_t d = 0;
    for (int sq = 0; sq < 64; ++sq)
        if (before.mailbox[sq] != after.mailbox[sq]) d |= bit(sq);
    return d;
}

/* Perspective c must be rebuilt when its king changes input bucket or board half
 * (psq.h:264-283 refreshRequired; nnue_state.h:118-128 threat refresh is implied by the latter). */
template <typename B>
SP_HD bool needs_refresh(const FeatureTables& t, const B& before, const B& after, int c) {
    const int kb = before.king[c], ka = after.king[c];
    if (kb == ka) return false;
    return ((kb & 7) >= 4) != ((ka & 7) >= 4) || king_bucket(t, c, kb) != king_bucket(t, c, ka);
}

SP_HD bool slides_along(int piece, int dir) { /* dir 0..7 = N,NE,E,SE,S,SW,W,NW; odd = diagonal */
    const int type = piece >> 1;
    return type == kQueen || (type == kBishop && (dir & 1)) || (type == kRook && !(dir & 1));
}

/* Does `piece` standing `dist` steps away attack along direction `dir` (pointing from the piece
 * towards the target)?  Kings never count (threats.cpp:42-54 maps them to -1 anyway). */
SP_HD bool attacks_along(int piece, int dir, int dist) {
    if (piece == kNoPiece) return false;
    if (slides_along(piece, dir)) return true;
    if ((piece >> 1) == kPawn && dist == 1) {
        /* white pawns capture NE / NW, black pawns SE / SW (attacks.h:37-48) */
        return (piece & 1) == kWhite ? (dir == 1 || dir == 7) : (dir == 3 || dir == 5);
    }
    return false;
}

/* First occupied square from s along dir (exclusive of s); kNoSquare if the ray leaves the board.
 * `between` receives the empty squares walked over. */
template <typename B>
SP_HD int ray_first(const B& b, int s, int dir, uint64_t& between, int& dist) {
    const int dx = (dir >= 1 && dir <= 3) ? 1 : ((dir >= 5) ? -1 : 0);
    const int dy = (dir == 0 || dir == 1 || dir == 7) ? 1 : ((dir >= 3 && dir <= 5) ? -1 : 0);
    int f = (s & 7) + dx, r = (s >> 3) + dy;
    between = 0;
    dist = 1;
    while (f >= 0 && f < 8 && r >= 0 && r < 8) {
        const int sq = r * 8 + f;
        if ((b.occ >> sq) & 1) return sq;
        between |= bit(sq);
        f += dx;
        r += dy;
        ++dist;
    }
    return kNoSquare;
}

template <typename B, typename Emit>
SP_HD void emit_threat(const FeatureTables& t, const B& b, int sign, int attacker, int asq, int victim, int vsq, Emit&& emit) {
    const int32_t fb = threat_index(t, kBlack, b.king[kBlack], attacker, asq, victim, vsq);
    const int32_t fw = threat_index(t, kWhite, b.king[kWhite], attacker, asq, victim, vsq);
    if (fb >= 0) emit(kBlack, 1, sign, static_cast<uint32_t>(fb));
    if (fw >= 0) emit(kWhite, 1, sign, static_cast<uint32_t>(fw));
}

/* emit(perspective, kind, sign, index): kind 0 = PSQ row, 1 = threat / pawn-pair row;
 * sign +1 = add (feature of `after`), -1 = subtract (feature of `before`).
 * Callers skip perspectives that are being refreshed. */
template <typename B, typename Emit>
SP_HD void delta_item(const FeatureTables& t, const B& before, const B& after, uint64_t changed, int item, Emit&& emit) {
    const int j = item / (2 * kItemsPerSquareBoard);
    const int rem = item - j * (2 * kItemsPerSquareBoard);
    const bool is_after = rem >= kItemsPerSquareBoard;
    const int k = is_after ? rem - kItemsPerSquareBoard : rem;
    if (j >= popcount64(changed)) return;
    uint64_t rest = changed;
    for (int i = 0; i < j; ++i) rest &= rest - 1;
    const int s = lsb64(rest);
    const B& b = is_after ? after : before;
    const int sign = is_after ? 1 : -1;
    const int piece = b.mailbox[s];

    if (k == 16) {
        if (piece == kNoPiece) return;
        emit(kBlack, 0, sign, psq_index(t, kBlack, piece, s, b.king[kBlack]));
        emit(kWhite, 0, sign, psq_index(t, kWhite, piece, s, b.king[kWhite]));
        if ((piece >> 1) != kPawn) return;
        uint64_t partners = (b.pawns[0] | b.pawns[1]) & pp_mask(s) & ~bit(s);
        partners &= ~(changed & (bit(s) - 1)); /* a pair of two changed pawns belongs to the lower square */
        while (partners) {
            const int o = lsb64(partners);
            partners &= partners - 1;
            const int oc = b.mailbox[o] & 1;
            emit(kBlack, 1, sign, pp_index(kBlack, b.king[kBlack], piece & 1, s, oc, o));
            emit(kWhite, 1, sign, pp_index(kWhite, b.king[kWhite], piece & 1, s, oc, o));
        }
        return;
    }

    if (k >= 8) { /* knight offsets */
        const int kdx[8] = {1, 2, 2, 1, -1, -2, -2, -1};
        const int kdy[8] = {2, 1, -1, -2, -2, -1, 1, 2};
        if (piece == kNoPiece) return;
        const int f = (s & 7) + kdx[k - 8], r = (s >> 3) + kdy[k - 8];
        if (f < 0 || f > 7 || r < 0 || r > 7) return;
        const int o = r * 8 + f;
        const int other = b.mailbox[o];
        if (other == kNoPiece) return;
        if ((piece >> 1) == kKnight) emit_threat(t, b, sign, piece, s, other, o, emit);
        if ((other >> 1) == kKnight && !((changed >> o) & 1)) emit_threat(t, b, sign, other, o, piece, s, emit);
        return;
    }

    /* ray direction k */
    uint64_t gap_ahead;
    int dist_ahead;
    const int ahead = ray_first(b, s, k, gap_ahead, dist_ahead);
    if (piece != kNoPiece) {
        if (ahead == kNoSquare) return;
        const int other = b.mailbox[ahead];
        if (attacks_along(piece, k, dist_ahead)) emit_threat(t, b, sign, piece, s, other, ahead, emit);
        if (!((changed >> ahead) & 1) && attacks_along(other, k ^ 4, dist_ahead))
            emit_threat(t, b, sign, other, ahead, piece, s, emit);
        return;
    }
    /* s is empty in this board: sliders behind s see through it to the first piece ahead */
    if (ahead == kNoSquare || ((changed >> ahead) & 1)) return;
    uint64_t gap_behind;
    int dist_behind;
    const int behind = ray_first(b, s, k ^ 4, gap_behind, dist_behind);
    if (behind == kNoSquare || ((changed >> behind) & 1)) return;
    if (gap_behind & changed) return; /* a changed square nearer to the attacker owns this pair */
    const int attacker = b.mailbox[behind];
    if (slides_along(attacker, k)) emit_threat(t, b, sign, attacker, behind, b.mailbox[ahead], ahead, emit);
}

} // namespace sp

#endif /* SP_DELTA_H */
