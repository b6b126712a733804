/*
 * position.h -- minimal board model for the host side of the NNUE path.
 *
 * The reference's Position / movegen (src/position.{h,cpp}, src/movegen.cpp) are OUT of scope
 * for this project; a drop-in build keeps using them.  This small stand-in exists so that the
 * eval-API mirror (nnue_state.h), the synthetic-workload generator (random legal playouts) and
 * the tests have something to drive the evaluator with.  It keeps the reference's accessor
 * names and, crucially, the observer protocol of Position::applyMove<Observer>
 * (src/position.cpp:109-197, 1306-1466): the same callbacks fire in the same order on the same
 * intermediate board states, so an observer written for the engine works unchanged.
 */
#ifndef SP_HOST_POSITION_H
#define SP_HOST_POSITION_H

#include <cstdint>
#include <string>

#include "../sp_features.h"

/* The board core (put / remove / attacks / applyMove / generateLegal / pack) is header-inline and compiles for
 * the device as well: the GPU-resident self-play driver (selfplay_gpu.cu) runs the same code per game. */
#if defined(__CUDACC__)
    #define SP_POS_HD __host__ __device__
#else
    #define SP_POS_HD
#endif

namespace sp::host {

using Color = int;  /* kBlack = 0, kWhite = 1 */
using Piece = int;  /* type << 1 | color, kNoPiece = 12 */
using Square = int; /* a1 = 0 .. h8 = 63, kNoSquare = 64 */

enum class MoveType : uint16_t { kStandard = 0, kPromotion, kCastling, kEnPassant };

/* 16-bit move, bit-compatible with src/move.h:28-121 */
struct Move {
    uint16_t raw{0};

    SP_POS_HD Square from() const { return raw >> 10; }
    SP_POS_HD Square to() const { return (raw >> 4) & 0x3F; }
    SP_POS_HD int promo() const { return ((raw >> 2) & 3) + 1; } /* piece type */
    SP_POS_HD MoveType type() const { return static_cast<MoveType>(raw & 3); }
    SP_POS_HD explicit operator bool() const { return raw != 0; }
    SP_POS_HD bool operator==(const Move& o) const { return raw == o.raw; }

    SP_POS_HD static Move standard(Square s, Square d) { return {static_cast<uint16_t>(s << 10 | d << 4)}; }
    SP_POS_HD static Move promotion(Square s, Square d, int pt) {
        return {static_cast<uint16_t>(s << 10 | d << 4 | (pt - 1) << 2 | 1)};
    }
    SP_POS_HD static Move castling(Square king, Square rook) { return {static_cast<uint16_t>(king << 10 | rook << 4 | 2)}; }
    SP_POS_HD static Move enPassant(Square s, Square d) { return {static_cast<uint16_t>(s << 10 | d << 4 | 3)}; }
};

struct NullObserver {
    SP_POS_HD void prepareKingMove(Color, Square, Square) {}
    template <typename P> SP_POS_HD void pieceAdded(const P&, Piece, Square) {}
    template <typename P> SP_POS_HD void pieceRemoved(const P&, Piece, Square) {}
    template <typename P> SP_POS_HD void pieceMutated(const P&, Piece, Piece, Square) {}
    template <typename P> SP_POS_HD void pieceMoved(const P&, Piece, Square, Square) {}
    template <typename P> SP_POS_HD void piecePromoted(const P&, Piece, Square, Piece, Square) {}
    template <typename P> SP_POS_HD void finalize(const P&, const P&) {}
};

class Position {
public:
    SP_POS_HD Position() {
        for (int i = 0; i < 64; ++i) m_mailbox[i] = kNoPiece;
    }

    static Position startpos();
    /* Double Fischer random start position n = black * 960 + white, both sides numbered by Scharnagl's
     * scheme (src/position.cpp:38-100, 1215-1270); n < 960 * 960.  518 * 960 + 518 is the standard one. */
    SP_POS_HD static Position fromDfrcIndex(uint32_t n);
    static bool fromFen(const std::string& fen, Position& out);
    static bool fromPacked(const SpPackedBoard& packed, Position& out);
    [[nodiscard]] SP_POS_HD SpPackedBoard pack() const;
    [[nodiscard]] std::string toFen() const;

    /* Moves are assumed to be legal (same contract as the reference). */
    template <typename Observer>
    [[nodiscard]] SP_POS_HD Position applyMove(Move move, Observer&& observer) const;
    [[nodiscard]] SP_POS_HD Position applyMove(Move move) const { return applyMove(move, NullObserver{}); }

    /* All legal moves; returns the count (<= 256). Castling is encoded king-takes-rook. */
    SP_POS_HD int generateLegal(Move* out) const;
    /* The same list in the same order with every pseudo-legal move made and tested (the obviously correct
     * form; tests/cpp/test_movegen.cpp holds generateLegal against it). */
    SP_POS_HD int generateLegalSlow(Move* out) const;

    [[nodiscard]] SP_POS_HD Piece pieceOn(Square sq) const { return m_mailbox[sq]; }
    [[nodiscard]] SP_POS_HD const uint8_t* mailbox() const { return m_mailbox; }
    [[nodiscard]] SP_POS_HD uint64_t occ() const { return m_color[0] | m_color[1]; }
    [[nodiscard]] SP_POS_HD uint64_t bb(Color c) const { return m_color[c]; }
    [[nodiscard]] SP_POS_HD uint64_t bbType(int type) const { return m_type[type]; }
    [[nodiscard]] SP_POS_HD uint64_t bb(int type, Color c) const { return m_type[type] & m_color[c]; }
    [[nodiscard]] SP_POS_HD Square king(Color c) const { return m_king[c]; }
    [[nodiscard]] SP_POS_HD Color stm() const { return m_stm; }
    [[nodiscard]] SP_POS_HD Square enPassant() const { return m_ep; }
    [[nodiscard]] SP_POS_HD int halfmove() const { return m_halfmove; }
    [[nodiscard]] SP_POS_HD int fullmove() const { return m_fullmove; }
    /* [color][0 = kingside, 1 = queenside] rook squares that may still castle */
    [[nodiscard]] SP_POS_HD Square castlingRook(Color c, int side) const { return m_rooks[c][side]; }

    [[nodiscard]] SP_POS_HD bool isAttacked(Square sq, Color by, uint64_t occupancy) const {
        const uint64_t them = m_color[by];
        if (pawn_attacks(sq, by ^ 1) & them & m_type[kPawn]) return true;
        if (knight_attacks(sq) & them & m_type[kKnight]) return true;
        if (king_attacks(sq) & them & m_type[kKing]) return true;
        if (bishop_attacks(sq, occupancy) & them & (m_type[kBishop] | m_type[kQueen])) return true;
        if (rook_attacks(sq, occupancy) & them & (m_type[kRook] | m_type[kQueen])) return true;
        return false;
    }
    [[nodiscard]] SP_POS_HD bool isCheck() const { return isAttacked(m_king[m_stm], m_stm ^ 1, occ()); }

    /* View for the shared feature code (sp_features.h) */
    void toBoard(Board& b) const;

private:
    SP_POS_HD int generatePseudo(Move* pseudo) const;
    SP_POS_HD void put(Piece p, Square sq) {
        m_mailbox[sq] = static_cast<uint8_t>(p);
        m_color[p & 1] |= bit(sq);
        m_type[p >> 1] |= bit(sq);
    }
    SP_POS_HD void remove(Piece p, Square sq) {
        m_mailbox[sq] = kNoPiece;
        m_color[p & 1] &= ~bit(sq);
        m_type[p >> 1] &= ~bit(sq);
    }
    SP_POS_HD void filterEp();

    uint8_t m_mailbox[64];
    uint64_t m_color[2]{};
    uint64_t m_type[6]{};
    Square m_king[2]{kNoSquare, kNoSquare};
    Square m_rooks[2][2]{{kNoSquare, kNoSquare}, {kNoSquare, kNoSquare}};
    Square m_ep{kNoSquare};
    int m_halfmove{0};
    int m_fullmove{1};
    Color m_stm{kWhite};
};

/*
 * Observer protocol, callback for callback as in src/position.cpp:1306-1466.
 */
template <typename Observer>
SP_POS_HD Position Position::applyMove(Move move, Observer&& observer) const {
    Position np = *this;
    np.m_stm = m_stm ^ 1;
    np.m_ep = kNoSquare;
    const Color us = m_stm, them = us ^ 1;
    if (us == kBlack) ++np.m_fullmove;
    if (!move) return np;

    const Square src = move.from(), dst = move.to();
    const Piece moving = pieceOn(src);
    const int movingType = moving >> 1;
    Piece captured = kNoPiece;

    switch (move.type()) {
        case MoveType::kStandard: { /* movePiece, position.cpp:1306-1341 */
            if (movingType == kKing) {
                observer.prepareKingMove(us, np.m_king[us], dst);
                np.m_king[us] = dst;
            }
            captured = np.pieceOn(dst);
            if (captured != kNoPiece) {
                np.remove(moving, src);
                observer.pieceRemoved(np, moving, src);
                np.remove(captured, dst);
                np.put(moving, dst);
                observer.pieceMutated(np, captured, moving, dst);
            } else {
                np.remove(moving, src);
                np.put(moving, dst);
                observer.pieceMoved(np, moving, src, dst);
            }
            break;
        }
        case MoveType::kPromotion: { /* promotePawn, position.cpp:1348-1385 */
            captured = np.pieceOn(dst);
            const Piece promo = move.promo() << 1 | us;
            if (captured != kNoPiece) {
                np.remove(moving, src);
                observer.pieceRemoved(np, moving, src);
                np.remove(captured, dst);
                np.put(promo, dst);
                observer.pieceMutated(np, captured, promo, dst);
            } else {
                np.remove(moving, src);
                np.put(promo, dst);
                observer.piecePromoted(np, moving, src, promo, dst);
            }
            break;
        }
        case MoveType::kCastling: { /* castle, position.cpp:1392-1435; dst is the rook's square */
            const Square rookSrc = dst;
            const bool kingside = (src & 7) < (rookSrc & 7);
            const Square kingDst = (src & 56) | (kingside ? 6 : 2);
            const Square rookDst = (src & 56) | (kingside ? 5 : 3);
            const Piece rook = kRook << 1 | us;
            observer.prepareKingMove(us, src, kingDst);
            np.m_king[us] = kingDst;
            np.remove(moving, src);
            observer.pieceRemoved(np, moving, src);
            np.remove(rook, rookSrc);
            observer.pieceRemoved(np, rook, rookSrc);
            np.put(moving, kingDst);
            observer.pieceAdded(np, moving, kingDst);
            np.put(rook, rookDst);
            observer.pieceAdded(np, rook, rookDst);
            break;
        }
        case MoveType::kEnPassant: { /* enPassant, position.cpp:1442-1466 */
            const Square capSq = dst ^ 8;
            const Piece enemyPawn = moving ^ 1;
            np.remove(enemyPawn, capSq);
            observer.pieceRemoved(np, enemyPawn, capSq);
            np.remove(moving, src);
            np.put(moving, dst);
            observer.pieceMoved(np, moving, src, dst);
            captured = enemyPawn;
            break;
        }
    }

    observer.finalize(*this, np);

    /* castling rights, clocks, en passant: position.cpp:164-186 */
    if (movingType == kRook) {
        for (int side = 0; side < 2; ++side)
            if (np.m_rooks[us][side] == src) np.m_rooks[us][side] = kNoSquare;
    } else if (movingType == kKing) {
        np.m_rooks[us][0] = np.m_rooks[us][1] = kNoSquare;
    } else if (movingType == kPawn && ((src >> 3) - (dst >> 3) == 2 || (dst >> 3) - (src >> 3) == 2)) {
        np.m_ep = dst ^ 8;
    }
    np.m_halfmove = (captured == kNoPiece && movingType != kPawn) ? m_halfmove + 1 : 0;
    if (captured != kNoPiece && (captured >> 1) == kRook && move.type() != MoveType::kEnPassant) {
        for (int side = 0; side < 2; ++side)
            if (np.m_rooks[them][side] == dst) np.m_rooks[them][side] = kNoSquare;
    }
    np.filterEp();
    return np;
}

/* marlinformat record, src/datagen/marlinformat.h:43-84 */
SP_POS_HD inline SpPackedBoard Position::pack() const {
    SpPackedBoard out{};
    out.occupancy = occ();
    int i = 0;
    for (uint64_t bbs = occ(); bbs; bbs &= bbs - 1, ++i) {
        const Square sq = lsb64(bbs);
        const Piece p = pieceOn(sq);
        unsigned pt = static_cast<unsigned>(p >> 1);
        if (pt == kRook) {
            const Color c = p & 1;
            if (m_rooks[c][0] == sq || m_rooks[c][1] == sq) pt = 6;
        }
        const unsigned nib = pt | ((p & 1) == kBlack ? 8u : 0u);
        out.pieces[i / 2] |= static_cast<uint8_t>(nib << ((i % 2) * 4));
    }
    const Square ep = m_ep == kNoSquare ? kNoSquare : ((m_ep & 7) | (m_stm == kBlack ? 2 * 8 : 5 * 8));
    out.stm_ep = static_cast<uint8_t>((m_stm == kBlack ? 0x80 : 0) | ep);
    out.halfmove = static_cast<uint8_t>(m_halfmove > 255 ? 255 : m_halfmove);
    out.fullmove = static_cast<uint16_t>(m_fullmove);
    return out;
}

/* Keep the en-passant square only if some en-passant capture is legal
 * (intent of Position::filterEp, src/position.cpp:1608-1700). */
SP_POS_HD inline void Position::filterEp() {
    if (m_ep == kNoSquare) return;
    const Color us = m_stm;
    const Piece pawn = kPawn << 1 | us;
    uint64_t candidates = pawn_attacks(m_ep, us ^ 1) & bb(kPawn, us);
    const Square capSq = m_ep ^ 8;
    if (pieceOn(capSq) != (kPawn << 1 | (us ^ 1))) candidates = 0;
    bool ok = false;
    for (; candidates && !ok; candidates &= candidates - 1) {
        const Square src = lsb64(candidates);
        Position np = *this;
        np.remove(pawn ^ 1, capSq);
        np.remove(pawn, src);
        np.put(pawn, m_ep);
        ok = !np.isAttacked(np.m_king[us], us ^ 1, np.occ());
    }
    if (!ok) m_ep = kNoSquare;
}

SP_POS_HD inline int Position::generatePseudo(Move* pseudo) const {
    int n = 0;
    const Color us = m_stm, them = us ^ 1;
    const uint64_t own = m_color[us], enemy = m_color[them], all = own | enemy;
    const int up = us == kWhite ? 8 : -8;
    const int promoRank = us == kWhite ? 7 : 0, startRank = us == kWhite ? 1 : 6;

    for (uint64_t bbs = own; bbs; bbs &= bbs - 1) {
        const Square src = lsb64(bbs);
        const Piece p = pieceOn(src);
        const int type = p >> 1;
        if (type == kPawn) {
            auto push = [&](Square dst) {
                if ((dst >> 3) == promoRank) {
                    for (int pt = kQueen; pt >= kKnight; --pt) pseudo[n++] = Move::promotion(src, dst, pt);
                } else {
                    pseudo[n++] = Move::standard(src, dst);
                }
            };
            const Square one = src + up;
            if (!(all & bit(one))) {
                push(one);
                if ((src >> 3) == startRank && !(all & bit(one + up))) pseudo[n++] = Move::standard(src, one + up);
            }
            for (uint64_t caps = pawn_attacks(src, us) & enemy; caps; caps &= caps - 1) push(lsb64(caps));
            if (m_ep != kNoSquare && (pawn_attacks(src, us) & bit(m_ep))) pseudo[n++] = Move::enPassant(src, m_ep);
        } else {
            for (uint64_t dsts = piece_attacks(p, src, all) & ~own; dsts; dsts &= dsts - 1)
                pseudo[n++] = Move::standard(src, lsb64(dsts));
        }
    }

    /* castling (Chess960 rules, src/movegen.cpp:172-196): squares the king and rook cross or land
     * on must be empty apart from the two of them; the king may not start on, cross or land on an
     * attacked square. Rook-shielded attacks on the landing square are caught by the make-and-test. */
    const Square ksq = m_king[us];
    if (!isAttacked(ksq, them, all)) {
        for (int side = 0; side < 2; ++side) {
            const Square rsq = m_rooks[us][side];
            if (rsq == kNoSquare) continue;
            const Square kingDst = (ksq & 56) | (side == 0 ? 6 : 2);
            const Square rookDst = (ksq & 56) | (side == 0 ? 5 : 3);
            auto between = [](Square a, Square b) { /* exclusive of a, inclusive of b */
                uint64_t m = 0;
                const int step = b > a ? 1 : -1;
                for (Square s = a; s != b;) {
                    s += step;
                    m |= bit(s);
                }
                return m;
            };
            const uint64_t clear = (between(ksq, kingDst) | between(ksq, rsq) | bit(kingDst) | bit(rookDst))
                                 & ~(bit(ksq) | bit(rsq));
            if (clear & all) continue;
            bool safe = true;
            for (uint64_t path = between(ksq, kingDst); path && safe; path &= path - 1)
                safe = !isAttacked(lsb64(path), them, all);
            if (safe) pseudo[n++] = Move::castling(ksq, rsq);
        }
    }
    return n;
}

/* Scharnagl's direct derivation of back rank n (0..959), files a..h: bishops from n mod 4 and (n / 4) mod 4
 * on their own square colours, the queen on the ((n / 16) mod 6)-th free file, the knights on the pair of
 * free files numbered n / 96 (0..9: 01 02 03 04 12 13 14 23 24 34), rook, king, rook on what is left. */
SP_POS_HD inline void scharnagl_backrank(uint32_t n, int (&rank)[8]) {
    for (int& t : rank) t = -1;
    auto nthFree = [&](int k) {
        for (int f = 0; f < 8; ++f)
            if (rank[f] < 0 && k-- == 0) return f;
        return 7;
    };
    rank[2 * (n % 4) + 1] = kBishop;
    n /= 4;
    rank[2 * (n % 4)] = kBishop;
    n /= 4;
    rank[nthFree(static_cast<int>(n % 6))] = kQueen;
    n /= 6;
    /* the n-th pair (a, b), a < b, of the five free files in lexicographic order */
    int a = 0, b = 1;
    for (uint32_t k = 0; k < n; ++k)
        if (++b == 5) ++a, b = a + 1;
    const int fa = nthFree(a), fb = nthFree(b);
    rank[fa] = kKnight, rank[fb] = kKnight;
    rank[nthFree(0)] = kRook;
    rank[nthFree(0)] = kKing;
    rank[nthFree(0)] = kRook;
}

SP_POS_HD inline Position Position::fromDfrcIndex(uint32_t n) {
    Position p;
    int back[2][8]; /* [color] */
    scharnagl_backrank(n % 960, back[kWhite]);
    scharnagl_backrank((n / 960) % 960, back[kBlack]);
    for (Color c = 0; c < 2; ++c) {
        const int base = c == kWhite ? 0 : 56, pawns = c == kWhite ? 8 : 48;
        bool firstRook = true;
        for (int f = 0; f < 8; ++f) {
            p.put(kPawn << 1 | c, pawns + f);
            p.put(back[c][f] << 1 | c, base + f);
            if (back[c][f] == kKing) p.m_king[c] = base + f;
            if (back[c][f] == kRook) { /* the rook on the lower file castles queenside */
                p.m_rooks[c][firstRook ? 1 : 0] = base + f;
                firstRook = false;
            }
        }
    }
    return p;
}

/* squares strictly between two aligned squares (0 if they share no rank, file or diagonal) */
SP_POS_HD inline uint64_t between_mask(Square a, Square b) {
    const uint64_t ra = rook_attacks(a, bit(b));
    if (ra & bit(b)) return ra & rook_attacks(b, bit(a));
    const uint64_t ba = bishop_attacks(a, bit(b));
    if (ba & bit(b)) return ba & bishop_attacks(b, bit(a));
    return 0;
}
/* the whole line through two aligned squares */
SP_POS_HD inline uint64_t line_mask(Square a, Square b) {
    const uint64_t ra = rook_attacks(a, 0);
    if (ra & bit(b)) return (ra & rook_attacks(b, 0)) | bit(a) | bit(b);
    const uint64_t ba = bishop_attacks(a, 0);
    if (ba & bit(b)) return (ba & bishop_attacks(b, 0)) | bit(a) | bit(b);
    return 0;
}

SP_POS_HD inline int Position::generateLegalSlow(Move* out) const {
    Move pseudo[256];
    const int n = generatePseudo(pseudo);
    const Color us = m_stm, them = us ^ 1;
    int legal = 0;
    for (int i = 0; i < n; ++i) {
        const Position np = applyMove(pseudo[i]);
        if (!np.isAttacked(np.m_king[us], them, np.occ())) out[legal++] = pseudo[i];
    }
    return legal;
}

/* Pseudo-legal generation, then: in check, make and test every move; otherwise only king moves, castling
 * and en passant need the test -- any other move is legal unless its piece is pinned and leaves the line
 * through the king and the pinner. */
SP_POS_HD inline int Position::generateLegal(Move* out) const {
    Move pseudo[256];
    const int n = generatePseudo(pseudo);
    const Color us = m_stm, them = us ^ 1;
    const Square ksq = m_king[us];
    const uint64_t own = m_color[us], enemy = m_color[them], all = own | enemy;
    const bool inCheck = isAttacked(ksq, them, all);
    uint64_t pinned = 0;
    if (!inCheck) {
        uint64_t snipers = (rook_attacks(ksq, 0) & enemy & (m_type[kRook] | m_type[kQueen]))
                         | (bishop_attacks(ksq, 0) & enemy & (m_type[kBishop] | m_type[kQueen]));
        for (; snipers; snipers &= snipers - 1) {
            const uint64_t blockers = between_mask(ksq, lsb64(snipers)) & all;
            if (blockers && !(blockers & (blockers - 1))) pinned |= blockers & own;
        }
    }
    int legal = 0;
    for (int i = 0; i < n; ++i) {
        const Move m = pseudo[i];
        const Square src = m.from();
        bool ok;
        if (inCheck || src == ksq || m.type() == MoveType::kEnPassant || m.type() == MoveType::kCastling) {
            const Position np = applyMove(m);
            ok = !np.isAttacked(np.m_king[us], them, np.occ());
        } else {
            ok = !(pinned & bit(src)) || (line_mask(ksq, src) & bit(m.to()));
        }
        if (ok) out[legal++] = m;
    }
    return legal;
}

} // namespace sp::host

#endif
