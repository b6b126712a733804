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

#include <array>
#include <cstdint>
#include <string>

#include "../sp_features.h"

namespace sp::host {

using Color = int;  /* kBlack = 0, kWhite = 1 */
using Piece = int;  /* type << 1 | color, kNoPiece = 12 */
using Square = int; /* a1 = 0 .. h8 = 63, kNoSquare = 64 */

enum class MoveType : uint16_t { kStandard = 0, kPromotion, kCastling, kEnPassant };

/* 16-bit move, bit-compatible with src/move.h:28-121 */
struct Move {
    uint16_t raw{0};

    Square from() const { return raw >> 10; }
    Square to() const { return (raw >> 4) & 0x3F; }
    int promo() const { return ((raw >> 2) & 3) + 1; } /* piece type */
    MoveType type() const { return static_cast<MoveType>(raw & 3); }
    explicit operator bool() const { return raw != 0; }
    bool operator==(const Move& o) const { return raw == o.raw; }

    static Move standard(Square s, Square d) { return {static_cast<uint16_t>(s << 10 | d << 4)}; }
    static Move promotion(Square s, Square d, int pt) {
        return {static_cast<uint16_t>(s << 10 | d << 4 | (pt - 1) << 2 | 1)};
    }
    static Move castling(Square king, Square rook) { return {static_cast<uint16_t>(king << 10 | rook << 4 | 2)}; }
    static Move enPassant(Square s, Square d) { return {static_cast<uint16_t>(s << 10 | d << 4 | 3)}; }
};

struct NullObserver {
    void prepareKingMove(Color, Square, Square) {}
    template <typename P> void pieceAdded(const P&, Piece, Square) {}
    template <typename P> void pieceRemoved(const P&, Piece, Square) {}
    template <typename P> void pieceMutated(const P&, Piece, Piece, Square) {}
    template <typename P> void pieceMoved(const P&, Piece, Square, Square) {}
    template <typename P> void piecePromoted(const P&, Piece, Square, Piece, Square) {}
    template <typename P> void finalize(const P&, const P&) {}
};

class Position {
public:
    Position();

    static Position startpos();
    static bool fromFen(const std::string& fen, Position& out);
    static bool fromPacked(const SpPackedBoard& packed, Position& out);
    [[nodiscard]] SpPackedBoard pack() const;
    [[nodiscard]] std::string toFen() const;

    /* Moves are assumed to be legal (same contract as the reference). */
    template <typename Observer>
    [[nodiscard]] Position applyMove(Move move, Observer&& observer) const;
    [[nodiscard]] Position applyMove(Move move) const { return applyMove(move, NullObserver{}); }

    /* All legal moves; returns the count (<= 256). Castling is encoded king-takes-rook. */
    int generateLegal(Move* out) const;

    [[nodiscard]] Piece pieceOn(Square sq) const { return m_mailbox[sq]; }
    [[nodiscard]] const std::array<uint8_t, 64>& mailbox() const { return m_mailbox; }
    [[nodiscard]] uint64_t occ() const { return m_color[0] | m_color[1]; }
    [[nodiscard]] uint64_t bb(Color c) const { return m_color[c]; }
    [[nodiscard]] uint64_t bbType(int type) const { return m_type[type]; }
    [[nodiscard]] uint64_t bb(int type, Color c) const { return m_type[type] & m_color[c]; }
    [[nodiscard]] Square king(Color c) const { return m_king[c]; }
    [[nodiscard]] Color stm() const { return m_stm; }
    [[nodiscard]] Square enPassant() const { return m_ep; }
    [[nodiscard]] int halfmove() const { return m_halfmove; }
    [[nodiscard]] int fullmove() const { return m_fullmove; }
    /* [color][0 = kingside, 1 = queenside] rook squares that may still castle */
    [[nodiscard]] Square castlingRook(Color c, int side) const { return m_rooks[c][side]; }

    [[nodiscard]] bool isAttacked(Square sq, Color by, uint64_t occupancy) const;
    [[nodiscard]] bool isCheck() const { return isAttacked(m_king[m_stm], m_stm ^ 1, occ()); }

    /* View for the shared feature code (sp_features.h) */
    void toBoard(Board& b) const;

private:
    void put(Piece p, Square sq);
    void remove(Piece p, Square sq);
    void filterEp();

    std::array<uint8_t, 64> m_mailbox{};
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
Position Position::applyMove(Move move, Observer&& observer) const {
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

} // namespace sp::host

#endif
