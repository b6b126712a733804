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
 * Work decomposition.  D = squares whose content differs.  A UNIT is (square s in D, board X):
 * X = before emits rows to subtract, X = after rows to add.  Per unit there are
 *     8 LINE items   k = 0..7: ray direction k from s AND knight offset k from s
 *                    (the reference's 8 rays, whose slot 0 is the knight square)
 *     1 SQUARE item  PSQ row of the piece on s, and the pawn pairs of a pawn on s
 * Line items are uniform code for every (unit, k), so a GPU warp runs 32 of them in lock-step:
 * lane = unit * 8 + k.  Rules that make every (attacker, victim) pair of X that touches D
 * appear exactly once:
 *     - attacker on a D square           -> emitted by the attacker's item
 *     - victim on D, attacker not on D   -> emitted by the victim's item
 *     - both off D, but a D square that is EMPTY in X lies strictly between them on the line
 *       (x-ray through the vacated / newly blocked square: the reference's "discovered"
 *       threats, nnue.cpp:424-484) -> emitted by the item of the D square closest to the
 *       attacker.
 */
#ifndef SP_DELTA_H
#define SP_DELTA_H

#include "sp_features.h"

namespace sp {

constexpr int kMaxChanged = 8; /* more changed squares than this -> full rebuild */
constexpr int kLineItemsPerUnit = 8;

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

/* Directions 0..7 = N, NE, E, SE, S, SW, W, NW; k ^ 4 is the opposite; odd = diagonal.
 * N, NE, E and NW lead to higher square numbers. */
SP_HD bool dir_is_up(int k) { return k < 3 || k == 7; }

SP_HD bool slides_along(int piece, int dir) {
    const int type = piece >> 1;
    return type == kQueen || (type == kBishop && (dir & 1)) || (type == kRook && !(dir & 1));
}

/* Does `piece` attack along direction `dir` (pointing from the piece towards the target)?
 * `adjacent`: the target is the very next square.  Kings never count (threats.cpp:42-54 maps
 * them to -1 anyway). */
SP_HD bool attacks_along(int piece, int dir, bool adjacent) {
    if (piece == kNoPiece) return false;
    if (slides_along(piece, dir)) return true;
    if ((piece >> 1) == kPawn && adjacent) {
        /* white pawns capture NE / NW, black pawns SE / SW (attacks.h:37-48) */
        return (piece & 1) == kWhite ? (dir == 1 || dir == 7) : (dir == 3 || dir == 5);
    }
    return false;
}

SP_HD uint64_t squares_above(int s) { return ~((uint64_t{2} << s) - 1); }
SP_HD uint64_t squares_below(int s) { return bit(s) - 1; }

/* First occupied square from s along k (exclusive of s), or kNoSquare; `between` receives the
 * empty squares passed over (the whole ray if it runs off the board). */
SP_HD int ray_first(uint64_t ray, uint64_t occ, int k, uint64_t& between) { /* ray = t.rays[k][s], already fetched */
    const uint64_t hits = ray & occ;
    between = ray;
    if (!hits) return kNoSquare;
    int first;
    if (dir_is_up(k)) {
        first = lsb64(hits);
        between = ray & squares_below(first);
    } else {
        first = msb64(hits);
        between = ray & squares_above(first);
    }
    return first;
}
SP_HD int ray_first(const FeatureTables& t, uint64_t occ, int s, int k, uint64_t& between) {
    const uint64_t ray = t.rays[k][s];
    const uint64_t hits = ray & occ;
    between = ray;
    if (!hits) return kNoSquare;
    int first;
    if (dir_is_up(k)) {
        first = lsb64(hits);
        between = ray & squares_below(first);
    } else {
        first = msb64(hits);
        between = ray & squares_above(first);
    }
    return first;
}

/* emit(perspective, kind, sign, index): kind 0 = PSQ row, 1 = threat / pawn-pair row;
 * sign +1 = add (feature of `after`), -1 = subtract (feature of `before`).
 * Callers skip perspectives that are being rebuilt. */
template <typename B, typename Emit>
SP_HD void emit_threat(const FeatureTables& t, const B& b, int sign, int attacker, int asq, int victim, int vsq, Emit&& emit) {
    const int32_t fb = threat_index(t, kBlack, b.king[kBlack], attacker, asq, victim, vsq);
    const int32_t fw = threat_index(t, kWhite, b.king[kWhite], attacker, asq, victim, vsq);
    if (fb >= 0) emit(kBlack, 1, sign, static_cast<uint32_t>(fb));
    if (fw >= 0) emit(kWhite, 1, sign, static_cast<uint32_t>(fw));
}

/* A threat candidate: asq | vsq << 8 | attacker << 16 | victim << 24 (bits 20-23 and 28-31 stay free
 * for the kernels' task flags). */
SP_HD uint32_t pack_candidate(int attacker, int asq, int victim, int vsq) {
    return static_cast<uint32_t>(asq | vsq << 8 | attacker << 16 | victim << 24);
}

/* Line item k of unit (s, b): up to four (attacker, victim) candidates in fixed slots
 * (0: piece -> ahead or x-ray, 1: ahead -> piece, 2: knight out, 3: knight in); returns the mask of
 * filled slots.  Candidates may still turn out to have no feature (threat_index < 0). */
template <typename B>
SP_HD unsigned delta_line_candidates(const FeatureTables& t, const B& b, uint64_t changed, int s, int k, uint32_t (&c)[4]) {
    unsigned valid = 0;
    const int piece = b.mailbox[s];
    uint64_t gap_ahead;
    const int ahead = ray_first(t, b.occ, s, k, gap_ahead);
    if (piece != kNoPiece) {
        if (ahead != kNoSquare) {
            const int other = b.mailbox[ahead];
            const bool adjacent = gap_ahead == 0;
            if (attacks_along(piece, k, adjacent)) c[0] = pack_candidate(piece, s, other, ahead), valid |= 1;
            if (!((changed >> ahead) & 1) && attacks_along(other, k ^ 4, adjacent)) c[1] = pack_candidate(other, ahead, piece, s), valid |= 2;
        }
        /* knight offset k: dx = {1,2,2,1,-1,-2,-2,-1}, dy = {2,1,-1,-2,-2,-1,1,2}, stored +2 per nibble */
        const int fx = (s & 7) + static_cast<int>((0x10013443u >> (4 * k)) & 0xF) - 2;
        const int ry = (s >> 3) + static_cast<int>((0x43100134u >> (4 * k)) & 0xF) - 2;
        if (fx >= 0 && fx < 8 && ry >= 0 && ry < 8) {
            const int o = ry * 8 + fx;
            const int other = b.mailbox[o];
            if (other != kNoPiece) {
                if ((piece >> 1) == kKnight) c[2] = pack_candidate(piece, s, other, o), valid |= 4;
                if ((other >> 1) == kKnight && !((changed >> o) & 1)) c[3] = pack_candidate(other, o, piece, s), valid |= 8;
            }
        }
    } else if (ahead != kNoSquare && !((changed >> ahead) & 1)) {
        /* s is empty in this board: sliders behind s see through it to the first piece ahead */
        uint64_t gap_behind;
        const int behind = ray_first(t, b.occ, s, k ^ 4, gap_behind);
        /* a changed square nearer to the attacker would own this pair */
        if (behind != kNoSquare && !((changed >> behind) & 1) && !(gap_behind & changed)) {
            const int attacker = b.mailbox[behind];
            if (slides_along(attacker, k)) c[0] = pack_candidate(attacker, behind, b.mailbox[ahead], ahead), valid |= 1;
        }
    }
    return valid;
}

template <typename B, typename Emit>
SP_HD void delta_line_item(const FeatureTables& t, const B& b, int sign, uint64_t changed, int s, int k, Emit&& emit) {
    uint32_t c[4] = {0, 0, 0, 0};
    const unsigned valid = delta_line_candidates(t, b, changed, s, k, c);
    for (int i = 0; i < 4; ++i) {
        if (!((valid >> i) & 1)) continue;
        emit_threat(t, b, sign, static_cast<int>((c[i] >> 16) & 0xF), static_cast<int>(c[i] & 0xFF), static_cast<int>((c[i] >> 24) & 0xF),
                    static_cast<int>((c[i] >> 8) & 0xFF), emit);
    }
}

/* Partners of a pawn on s for the pawn-pair features of unit (s, b): every other pawn within one file,
 * except changed pawns on lower squares (the pair then belongs to that lower square's unit). */
template <typename B>
SP_HD uint64_t delta_pawn_partners(const B& b, uint64_t changed, int s) {
    return (b.pawns[0] | b.pawns[1]) & pp_mask(s) & ~bit(s) & ~(changed & squares_below(s));
}

/* Square item of unit (s, b): PSQ row and pawn pairs. */
template <typename B, typename Emit>
SP_HD void delta_square_item(const FeatureTables& t, const B& b, int sign, uint64_t changed, int s, Emit&& emit) {
    const int piece = b.mailbox[s];
    if (piece == kNoPiece) return;
    emit(kBlack, 0, sign, psq_index(t, kBlack, piece, s, b.king[kBlack]));
    emit(kWhite, 0, sign, psq_index(t, kWhite, piece, s, b.king[kWhite]));
    if ((piece >> 1) != kPawn) return;
    uint64_t partners = delta_pawn_partners(b, changed, s);
    while (partners) {
        const int o = lsb64(partners);
        partners &= partners - 1;
        const int oc = b.mailbox[o] & 1;
        emit(kBlack, 1, sign, pp_index(kBlack, b.king[kBlack], piece & 1, s, oc, o));
        emit(kWhite, 1, sign, pp_index(kWhite, b.king[kWhite], piece & 1, s, oc, o));
    }
}

/* Square s of unit u (u = 2 * index-into-D + (after ? 1 : 0)). */
SP_HD int unit_square(uint64_t changed, int unit) {
    uint64_t rest = changed;
    for (int i = 0; i < (unit >> 1); ++i) rest &= rest - 1;
    return lsb64(rest);
}

/* Everything, sequentially (host tests; the kernels spread the items over lanes). */
template <typename B, typename Emit>
SP_HD void delta_all(const FeatureTables& t, const B& before, const B& after, uint64_t changed, Emit&& emit) {
    const int units = 2 * popcount64(changed);
    for (int u = 0; u < units; ++u) {
        const int s = unit_square(changed, u);
        const B& b = (u & 1) ? after : before;
        const int sign = (u & 1) ? 1 : -1;
        for (int k = 0; k < kLineItemsPerUnit; ++k) delta_line_item(t, b, sign, changed, s, k, emit);
        delta_square_item(t, b, sign, changed, s, emit);
    }
}

} // namespace sp

#endif /* SP_DELTA_H */
