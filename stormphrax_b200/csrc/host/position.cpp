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

void Position::toBoard(Board& b) const {
    std::memcpy(b.mailbox, m_mailbox, 64);
    b.occ = occ();
    b.pawns[0] = bb(kPawn, kBlack);
    b.pawns[1] = bb(kPawn, kWhite);
    b.king[0] = m_king[0];
    b.king[1] = m_king[1];
    b.stm = m_stm;
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

} // namespace sp::host
