/*
 * position.cpp -- see position.h.  Legal move generation is pseudo-legal generation followed by
 * a make-and-test king-safety check: simple, obviously correct, fast enough to feed benchmarks
 * (a few million plies per second per core).  Cross-checked move-for-move against the
 * reference's generateAll + isLegal in tests/test_host_board.py.
 */
#include "position.h"

#include <cstring>
#include <sstream>

namespace sp::host {

Position::Position() { m_mailbox.fill(kNoPiece); }

void Position::put(Piece p, Square sq) {
    m_mailbox[sq] = static_cast<uint8_t>(p);
    m_color[p & 1] |= bit(sq);
    m_type[p >> 1] |= bit(sq);
}

void Position::remove(Piece p, Square sq) {
    m_mailbox[sq] = kNoPiece;
    m_color[p & 1] &= ~bit(sq);
    m_type[p >> 1] &= ~bit(sq);
}

void Position::toBoard(Board& b) const {
    std::memcpy(b.mailbox, m_mailbox.data(), 64);
    b.occ = occ();
    b.pawns[0] = bb(kPawn, kBlack);
    b.pawns[1] = bb(kPawn, kWhite);
    b.king[0] = m_king[0];
    b.king[1] = m_king[1];
    b.stm = m_stm;
}

bool Position::isAttacked(Square sq, Color by, uint64_t occupancy) const {
    const uint64_t them = m_color[by];
    if (pawn_attacks(sq, by ^ 1) & them & m_type[kPawn]) return true;
    if (knight_attacks(sq) & them & m_type[kKnight]) return true;
    if (king_attacks(sq) & them & m_type[kKing]) return true;
    if (bishop_attacks(sq, occupancy) & them & (m_type[kBishop] | m_type[kQueen])) return true;
    if (rook_attacks(sq, occupancy) & them & (m_type[kRook] | m_type[kQueen])) return true;
    return false;
}

Position Position::startpos() {
    Position p;
    fromFen("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq - 0 1", p);
    return p;
}

bool Position::fromFen(const std::string& fen, Position& out) {
    Position p;
    std::istringstream in{fen};
    std::string board, stm, castling, ep;
    int halfmove = 0, fullmove = 1;
    if (!(in >> board >> stm >> castling >> ep)) return false;
    in >> halfmove >> fullmove;
    int rank = 7, file = 0;
    int kings[2] = {0, 0};
    for (const char ch : board) {
        if (ch == '/') {
            if (file != 8) return false;
            --rank;
            file = 0;
            continue;
        }
        if (ch >= '1' && ch <= '8') {
            file += ch - '0';
            continue;
        }
        static const char kChars[] = "pnbrqk";
        const char lower = static_cast<char>(ch | 0x20);
        const char* pos = std::strchr(kChars, lower);
        if (!pos || !lower || rank < 0 || file > 7) return false;
        const int type = static_cast<int>(pos - kChars);
        const Color color = (ch == lower) ? kBlack : kWhite;
        p.put(type << 1 | color, rank * 8 + file);
        if (type == kKing) {
            p.m_king[color] = rank * 8 + file;
            ++kings[color];
        }
        ++file;
    }
    if (rank != 0 || file != 8 || kings[0] != 1 || kings[1] != 1 || popcount64(p.occ()) > 32) return false;
    if (stm != "w" && stm != "b") return false;
    p.m_stm = stm == "w" ? kWhite : kBlack;
    if (p.isAttacked(p.m_king[p.m_stm ^ 1], p.m_stm, p.occ())) return false; /* opponent must not be in check */
    if (castling != "-") {
        for (const char flag : castling) {
            const Color c = (flag >= 'A' && flag <= 'Z') ? kWhite : kBlack;
            const int base = c == kWhite ? 0 : 56;
            const int kf = p.m_king[c] & 7;
            const char lower = static_cast<char>(flag | 0x20);
            const Piece rook = kRook << 1 | c;
            int rf = -1;
            if (lower == 'k') { /* outermost... the reference takes the first rook outward from the king */
                for (int f = kf + 1; f < 8; ++f)
                    if (p.pieceOn(base + f) == rook) { rf = f; break; }
            } else if (lower == 'q') {
                for (int f = kf - 1; f >= 0; --f)
                    if (p.pieceOn(base + f) == rook) { rf = f; break; }
            } else if (lower >= 'a' && lower <= 'h') {
                rf = lower - 'a';
                if (rf == kf) return false;
            } else {
                return false;
            }
            if (rf < 0 || (p.m_king[c] >> 3) != (base >> 3)) continue;
            p.m_rooks[c][rf > kf ? 0 : 1] = base + rf;
        }
    }
    if (ep != "-") {
        if (ep.size() != 2 || ep[0] < 'a' || ep[0] > 'h' || ep[1] < '1' || ep[1] > '8') return false;
        p.m_ep = (ep[1] - '1') * 8 + (ep[0] - 'a');
    }
    p.m_halfmove = halfmove;
    p.m_fullmove = fullmove;
    p.filterEp();
    out = p;
    return true;
}

std::string Position::toFen() const {
    std::string fen;
    static const char kChars[] = "pnbrqk";
    for (int rank = 7; rank >= 0; --rank) {
        int empty = 0;
        for (int file = 0; file < 8; ++file) {
            const Piece p = pieceOn(rank * 8 + file);
            if (p == kNoPiece) {
                ++empty;
                continue;
            }
            if (empty) fen += static_cast<char>('0' + empty);
            empty = 0;
            const char c = kChars[p >> 1];
            fen += (p & 1) == kWhite ? static_cast<char>(c - 32) : c;
        }
        if (empty) fen += static_cast<char>('0' + empty);
        if (rank) fen += '/';
    }
    fen += m_stm == kWhite ? " w " : " b ";
    std::string castling;
    for (int side = 0; side < 2; ++side)
        if (m_rooks[kWhite][side] != kNoSquare) castling += static_cast<char>('A' + (m_rooks[kWhite][side] & 7));
    for (int side = 0; side < 2; ++side)
        if (m_rooks[kBlack][side] != kNoSquare) castling += static_cast<char>('a' + (m_rooks[kBlack][side] & 7));
    fen += castling.empty() ? "-" : castling;
    if (m_ep != kNoSquare) {
        fen += ' ';
        fen += static_cast<char>('a' + (m_ep & 7));
        fen += static_cast<char>('1' + (m_ep >> 3));
    } else {
        fen += " -";
    }
    fen += ' ' + std::to_string(m_halfmove) + ' ' + std::to_string(m_fullmove);
    return fen;
}

/* marlinformat record, src/datagen/marlinformat.h:43-84 */
SpPackedBoard Position::pack() const {
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

bool Position::fromPacked(const SpPackedBoard& packed, Position& out) {
    Position p;
    if (popcount64(packed.occupancy) > 32) return false;
    int kings[2] = {0, 0};
    int i = 0;
    for (uint64_t bbs = packed.occupancy; bbs; bbs &= bbs - 1, ++i) {
        const Square sq = lsb64(bbs);
        const unsigned nib = (packed.pieces[i / 2] >> ((i % 2) * 4)) & 0xF;
        unsigned type = nib & 7;
        const Color c = (nib & 8) ? kBlack : kWhite;
        bool castle = false;
        if (type == 6) {
            type = kRook;
            castle = true;
        }
        if (type > kKing) return false;
        p.put(static_cast<Piece>(type << 1 | static_cast<unsigned>(c)), sq);
        if (type == kKing) {
            p.m_king[c] = sq;
            ++kings[c];
        }
        if (castle) p.m_rooks[c][0] = sq; /* side fixed up below once the king is known */
    }
    if (kings[0] != 1 || kings[1] != 1) return false;
    /* second pass: sort castling rooks into kingside / queenside */
    i = 0;
    for (Color c = 0; c < 2; ++c) p.m_rooks[c][0] = p.m_rooks[c][1] = kNoSquare;
    for (uint64_t bbs = packed.occupancy; bbs; bbs &= bbs - 1, ++i) {
        const Square sq = lsb64(bbs);
        const unsigned nib = (packed.pieces[i / 2] >> ((i % 2) * 4)) & 0xF;
        if ((nib & 7) != 6) continue;
        const Color c = (nib & 8) ? kBlack : kWhite;
        p.m_rooks[c][(sq & 7) > (p.m_king[c] & 7) ? 0 : 1] = sq;
    }
    p.m_stm = (packed.stm_ep & 0x80) ? kBlack : kWhite;
    const int ep = packed.stm_ep & 0x7F;
    p.m_ep = ep < 64 ? ep : kNoSquare;
    p.m_halfmove = packed.halfmove;
    p.m_fullmove = packed.fullmove ? packed.fullmove : 1;
    out = p;
    return true;
}

/* Keep the en-passant square only if some en-passant capture is legal
 * (intent of Position::filterEp, src/position.cpp:1608-1700). */
void Position::filterEp() {
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

int Position::generateLegal(Move* out) const {
    Move pseudo[256];
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

    int legal = 0;
    for (int i = 0; i < n; ++i) {
        const Position np = applyMove(pseudo[i]);
        if (!np.isAttacked(np.m_king[us], them, np.occ())) out[legal++] = pseudo[i];
    }
    return legal;
}

} // namespace sp::host
