/*
 * sp_features.h -- input-feature indexing for the Stormphrax NNUE, written once for host and
 * device (every function is SP_HD).  The CUDA kernels call these per lane; the host-side
 * NnueState mirror and the CPU unit tests call exactly the same code.
 *
 * What the reference computes with ~100 KB of constexpr lookup tables
 * (src/eval/nnue/features/threats.cpp:31-167) is restated here so that it fits a GPU:
 *   - kPieceIndices[piece][from][to] (48 KB)  ->  popcount(pseudo[piece][from] & (bit(to) - 1))
 *   - slider attacks (PEXT / magic tables)    ->  hyperbola quintessence with __brevll
 * leaving 9 KB of tables (SpFeatureTables) that are built on the host at start-up and copied to
 * device memory.
 *
 * Feature definitions followed:
 *   PSQ       src/eval/nnue/features/psq.h:338-365, king buckets src/eval/arch.h:53-65
 *   threats   src/eval/nnue/features/threats.cpp:170-198 (+ tables :31-167)
 *   pawn pair src/eval/nnue/features/threats.cpp:200-221, masks threats.h:106-123
 *   full enumeration  src/eval/nnue_state.cpp:309-354, 440-449
 */
#ifndef SP_FEATURES_H
#define SP_FEATURES_H

#include <stdint.h>

#include "../../include/sp_types.h"

#if defined(__CUDACC__)
    #define SP_HD __host__ __device__ __forceinline__
#else
    #define SP_HD inline
#endif

namespace sp {

enum : int { kPawn = 0, kKnight, kBishop, kRook, kQueen, kKing };
enum : int { kBlack = 0, kWhite = 1 };
enum : int { kNoPiece = 12, kNoSquare = 64 };

SP_HD int popcount64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __popcll(x);
#else
    return __builtin_popcountll(x);
#endif
}
SP_HD int lsb64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)x) - 1;
#else
    return __builtin_ctzll(x);
#endif
}
SP_HD int msb64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return 63 - __clzll((long long)x);
#else
    return 63 - __builtin_clzll(x);
#endif
}
SP_HD uint64_t bitrev64(uint64_t x) {
#if defined(__CUDA_ARCH__)
    return __brevll(x);
#else
    x = ((x >> 1) & 0x5555555555555555ULL) | ((x & 0x5555555555555555ULL) << 1);
    x = ((x >> 2) & 0x3333333333333333ULL) | ((x & 0x3333333333333333ULL) << 2);
    x = ((x >> 4) & 0x0F0F0F0F0F0F0F0FULL) | ((x & 0x0F0F0F0F0F0F0F0FULL) << 4);
    return __builtin_bswap64(x);
#endif
}
SP_HD uint64_t bit(int sq) { return uint64_t{1} << sq; }

constexpr uint64_t kFileA = 0x0101010101010101ULL;
constexpr uint64_t kFileH = 0x8080808080808080ULL;

/* ------------------------------------------------------------------ attack sets */

/* Attacks along one line; `mask` is the full line through sq (sq included).
 * forward = o - 2r, backward = the same on the bit-reversed board; blockers are included. */
SP_HD uint64_t line_attacks(uint64_t occ, uint64_t mask, int sq) {
    const uint64_t r = bit(sq);
    const uint64_t o = (occ & mask) | r; /* the slider's own square counts as occupied */
    const uint64_t fwd = o - 2 * r;
    const uint64_t rev = bitrev64(bitrev64(o) - 2 * bitrev64(r));
    return (fwd ^ rev) & mask;
}
SP_HD uint64_t rank_mask(int sq) { return 0xFFULL << (sq & 56); }
SP_HD uint64_t file_mask(int sq) { return kFileA << (sq & 7); }
SP_HD uint64_t diag_mask(int sq) { /* a1-h8 direction */
    const int d = (sq & 7) - (sq >> 3);
    const uint64_t main = 0x8040201008040201ULL;
    return d >= 0 ? (main >> (8 * d)) : (main << (8 * -d));
}
SP_HD uint64_t anti_mask(int sq) { /* a8-h1 direction */
    const int d = 7 - (sq & 7) - (sq >> 3);
    const uint64_t main = 0x0102040810204080ULL;
    return d >= 0 ? (main >> (8 * d)) : (main << (8 * -d));
}
SP_HD uint64_t bishop_attacks(int sq, uint64_t occ) {
    return line_attacks(occ, diag_mask(sq), sq) | line_attacks(occ, anti_mask(sq), sq);
}
SP_HD uint64_t rook_attacks(int sq, uint64_t occ) {
    return line_attacks(occ, rank_mask(sq), sq) | line_attacks(occ, file_mask(sq), sq);
}
SP_HD uint64_t knight_attacks(int sq) {
    const uint64_t b = bit(sq);
    const uint64_t l1 = (b >> 1) & ~kFileH, l2 = (b >> 2) & 0x3F3F3F3F3F3F3F3FULL;
    const uint64_t r1 = (b << 1) & ~kFileA, r2 = (b << 2) & 0xFCFCFCFCFCFCFCFCULL;
    const uint64_t h1 = l1 | r1, h2 = l2 | r2;
    return (h1 << 16) | (h1 >> 16) | (h2 << 8) | (h2 >> 8);
}
SP_HD uint64_t king_attacks(int sq) {
    uint64_t b = bit(sq);
    const uint64_t side = ((b >> 1) & ~kFileH) | ((b << 1) & ~kFileA);
    b |= side;
    return side | (b << 8) | (b >> 8);
}
SP_HD uint64_t pawn_attacks(int sq, int color) { /* attacks.h:37-48 */
    const uint64_t b = bit(sq);
    const uint64_t side = ((b >> 1) & ~kFileH) | ((b << 1) & ~kFileA);
    return color == kWhite ? side << 8 : side >> 8;
}
/* attacks::getAttacks(piece, src, occ), src/attacks/attacks.h:130-151 */
SP_HD uint64_t piece_attacks(int piece, int sq, uint64_t occ) {
    switch (piece >> 1) {
        case kPawn: return pawn_attacks(sq, piece & 1);
        case kKnight: return knight_attacks(sq);
        case kBishop: return bishop_attacks(sq, occ);
        case kRook: return rook_attacks(sq, occ);
        case kQueen: return bishop_attacks(sq, occ) | rook_attacks(sq, occ);
        case kKing: return king_attacks(sq);
        default: return 0;
    }
}

/* ------------------------------------------------------------------ tables */

struct FeatureTables {
    uint64_t pseudo[12][64];      /* empty-board attacks of piece (own colour for pawns) */
    int32_t attack_idx[12][12][2]; /* kAttackIndices, threats.cpp:138-167; INT32_MIN = excluded */
    uint16_t offsets[12][64];      /* kOffsets.offsets, threats.cpp:108-136 */
    uint8_t half_buckets[32];      /* eval/arch.h:53-65 */
    uint64_t rays[8][64];          /* empty-board ray from sq: N, NE, E, SE, S, SW, W, NW (sq excluded) */
};

/* Host-side construction; mirrors the constexpr lambdas in threats.cpp. */
inline void build_feature_tables(FeatureTables& t) {
    /* threats.cpp:42-54: pawn-pair inputs select the map without pawn->pawn threats */
    static const int kTargetMap[6][6] = {
        {-1, 0, -1, 1, -1, -1}, {0, 1, 2, 3, 4, -1},  {0, 1, 2, 3, -1, -1},
        {0, 1, 2, 3, -1, -1},   {0, 1, 2, 3, 4, -1},  {-1, -1, -1, -1, -1, -1},
    };
    static const uint8_t kHalf[32] = {0, 1, 2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 8,  9,  10, 11,
                                      12, 12, 13, 13, 12, 12, 13, 13, 14, 14, 15, 15, 14, 14, 15, 15};
    for (int i = 0; i < 32; ++i) t.half_buckets[i] = kHalf[i];
    int target_count[6];
    for (int a = 0; a < 6; ++a) {
        int n = 0;
        for (int v = 0; v < 6; ++v) n += kTargetMap[a][v] >= 0;
        target_count[a] = 2 * n;
    }
    for (int piece = 0; piece < 12; ++piece)
        for (int sq = 0; sq < 64; ++sq) t.pseudo[piece][sq] = piece_attacks(piece, sq, 0);
    static const int kDx[8] = {0, 1, 1, 1, 0, -1, -1, -1}, kDy[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    for (int k = 0; k < 8; ++k)
        for (int sq = 0; sq < 64; ++sq) {
            uint64_t ray = 0;
            for (int f = (sq & 7) + kDx[k], r = (sq >> 3) + kDy[k]; f >= 0 && f < 8 && r >= 0 && r < 8; f += kDx[k], r += kDy[k])
                ray |= bit(r * 8 + f);
            t.rays[k][sq] = ray;
        }
    int32_t total[12], base[12];
    int32_t running = 0;
    for (int ci = 0; ci < 2; ++ci) { /* white pieces first, then black: threats.cpp:116 */
        const int color = ci == 0 ? kWhite : kBlack;
        for (int pt = 0; pt < 6; ++pt) {
            const int piece = pt << 1 | color;
            int32_t acc = 0;
            for (int sq = 0; sq < 64; ++sq) {
                t.offsets[piece][sq] = static_cast<uint16_t>(acc);
                const int rank = sq >> 3;
                if (pt != kPawn || (rank > 0 && rank < 7)) acc += popcount64(t.pseudo[piece ^ 1][sq]);
            }
            total[piece] = acc;
            base[piece] = running;
            running += target_count[pt] * acc;
        }
    }
    for (int atk = 0; atk < 12; ++atk)
        for (int vic = 0; vic < 12; ++vic) {
            const int at = atk >> 1, vt = vic >> 1;
            const bool enemy = (atk & 1) != (vic & 1);
            const int map = kTargetMap[at][vt];
            const bool semi = at == vt && (enemy || at != kPawn);
            const int vic_black = (vic & 1) == kBlack ? 1 : 0;
            const int32_t f = base[atk] + (vic_black * (target_count[at] / 2) + map) * total[atk];
            t.attack_idx[atk][vic][0] = map < 0 ? INT32_MIN : f;
            t.attack_idx[atk][vic][1] = (map < 0 || semi) ? INT32_MIN : f;
        }
}

/* ------------------------------------------------------------------ feature indices */

/* Square transform shared by all three feature families for perspective c with its king on ksq:
 * vertical flip for black, horizontal flip when the king stands on files e-h. */
SP_HD int orient_mask(int c, int ksq) { return (c == kBlack ? 56 : 0) ^ ((ksq & 7) >= 4 ? 7 : 0); }

SP_HD int king_bucket(const FeatureTables& t, int c, int ksq) { /* psq.h:241-246 */
    if (c == kBlack) ksq ^= 56;
    const int f = ksq & 7;
    return t.half_buckets[(ksq >> 3) * 4 + (f < 4 ? f : 7 - f)];
}

/* psq::featureIndex, psq.h:338-365, merged king planes */
SP_HD uint32_t psq_index(const FeatureTables& t, int c, int piece, int sq, int ksq) {
    const int type = piece >> 1;
    const uint32_t color = (type == kKing || (piece & 1) == c) ? 0u : 1u;
    return static_cast<uint32_t>(king_bucket(t, c, ksq)) * SP_PSQ_PER_BUCKET + color * 384u
         + static_cast<uint32_t>(type) * 64u + static_cast<uint32_t>(sq ^ orient_mask(c, ksq));
}

/* threats::threatFeatureIndex, threats.cpp:170-198. Negative = no such feature. */
SP_HD int32_t threat_index(const FeatureTables& t, int c, int ksq, int attacker, int asq, int attacked, int vsq) {
    const int flip = orient_mask(c, ksq);
    const int col = c == kBlack ? 1 : 0;
    attacker ^= col;
    attacked ^= col;
    asq ^= flip;
    vsq ^= flip;
    const int32_t a = t.attack_idx[attacker][attacked][asq < vsq];
    if (a == INT32_MIN) return -1;
    return SP_PP_FEATURES + a + t.offsets[attacker][asq] + popcount64(t.pseudo[attacker][asq] & (bit(vsq) - 1));
}

/* threats::ppPawnId / ppFeatureIndex, threats.cpp:200-221 */
SP_HD uint32_t pp_index(int c, int ksq, int a_color, int asq, int b_color, int bsq) {
    const int flip = orient_mask(c, ksq);
    const uint32_t ia = static_cast<uint32_t>((asq ^ flip) - 8 + (a_color != c ? 48 : 0));
    const uint32_t ib = static_cast<uint32_t>((bsq ^ flip) - 8 + (b_color != c ? 48 : 0));
    const uint32_t hi = ia > ib ? ia : ib, lo = ia > ib ? ib : ia;
    return hi * (hi - 1) / 2 + lo;
}

/* kPpMasks, threats.h:106-123: files f-1..f+1, every rank; empty for squares on ranks 1/8 */
SP_HD uint64_t pp_mask(int sq) {
    if (sq < 8 || sq >= 56) return 0;
    const uint64_t f = file_mask(sq);
    return f | ((f >> 1) & ~kFileH) | ((f << 1) & ~kFileA);
}

/* MaterialCount<8>::getBucket, src/eval/nnue/output.h:51-54 */
SP_HD int output_bucket(uint64_t occ) { return (popcount64(occ) - 2) / 4; }

/* ------------------------------------------------------------------ boards */

struct Board {
    uint8_t mailbox[64]; /* Piece = type << 1 | color, kNoPiece when empty */
    uint64_t occ;
    uint64_t pawns[2];   /* [black, white] */
    int king[2];
    int stm;
};

/* Piece on `sq` of a packed board (marlinformat.h:43-68); kNoPiece if the square is empty. */
SP_HD int packed_piece_at(uint64_t occ, const uint8_t* nibbles, int sq) {
    if (!((occ >> sq) & 1)) return kNoPiece;
    const int i = popcount64(occ & (bit(sq) - 1));
    const unsigned nib = (nibbles[i >> 1] >> ((i & 1) * 4)) & 0xF;
    unsigned type = nib & 7;
    if (type == 6) type = kRook; /* rook with castling rights */
    return static_cast<int>(type << 1 | ((nib & 8) ? kBlack : kWhite));
}

/* Returns 0 on success; rejects records without exactly one king per side or with >32 pieces. */
SP_HD int unpack_board(const SpPackedBoard& p, Board& b) {
    b.occ = p.occupancy;
    b.pawns[0] = b.pawns[1] = 0;
    b.king[0] = b.king[1] = kNoSquare;
    b.stm = (p.stm_ep & 0x80) ? kBlack : kWhite;
    if (popcount64(b.occ) > 32) return 1;
    int kings[2] = {0, 0};
    for (int sq = 0; sq < 64; ++sq) {
        const int piece = packed_piece_at(p.occupancy, p.pieces, sq);
        b.mailbox[sq] = static_cast<uint8_t>(piece);
        if (piece == kNoPiece) continue;
        if ((piece >> 1) > kKing) return 2;
        if ((piece >> 1) == kPawn) b.pawns[piece & 1] |= bit(sq);
        if ((piece >> 1) == kKing) {
            b.king[piece & 1] = sq;
            ++kings[piece & 1];
        }
    }
    return (kings[0] == 1 && kings[1] == 1) ? 0 : 3;
}

/*
 * Per-square pieces of the full enumeration (nnue_state.cpp:309-354, 440-449).  Each emits the
 * features that "belong" to one square for BOTH perspectives, so a GPU lane can own a square.
 * emit(perspective, index) is called once per existing feature; order is irrelevant because the
 * accumulators are sums in Z/2^16.
 */
template <typename B, typename Emit>
SP_HD void square_psq_features(const FeatureTables& t, const B& b, int sq, Emit&& emit) {
    const int piece = b.mailbox[sq];
    if (piece == kNoPiece) return;
    emit(kBlack, psq_index(t, kBlack, piece, sq, b.king[kBlack]));
    emit(kWhite, psq_index(t, kWhite, piece, sq, b.king[kWhite]));
}

template <typename B, typename Emit>
SP_HD void square_threat_features(const FeatureTables& t, const B& b, int sq, Emit&& emit) {
    const int piece = b.mailbox[sq];
    if (piece == kNoPiece || (piece >> 1) == kKing) return;
    const uint64_t kings = bit(b.king[0]) | bit(b.king[1]);
    uint64_t targets = piece_attacks(piece, sq, b.occ) & b.occ & ~kings;
    while (targets) {
        const int to = lsb64(targets);
        targets &= targets - 1;
        const int victim = b.mailbox[to];
        const int32_t fb = threat_index(t, kBlack, b.king[kBlack], piece, sq, victim, to);
        const int32_t fw = threat_index(t, kWhite, b.king[kWhite], piece, sq, victim, to);
        if (fb >= 0) emit(kBlack, static_cast<uint32_t>(fb));
        if (fw >= 0) emit(kWhite, static_cast<uint32_t>(fw));
    }
    if ((piece >> 1) == kPawn) {
        /* every unordered pawn pair within one file of each other, once: partner on a higher square */
        const uint64_t all = b.pawns[0] | b.pawns[1];
        uint64_t partners = all & pp_mask(sq) & ~(bit(sq) | (bit(sq) - 1));
        while (partners) {
            const int other = lsb64(partners);
            partners &= partners - 1;
            const int oc = b.mailbox[other] & 1;
            emit(kBlack, pp_index(kBlack, b.king[kBlack], piece & 1, sq, oc, other));
            emit(kWhite, pp_index(kWhite, b.king[kWhite], piece & 1, sq, oc, other));
        }
    }
}

} // namespace sp

#endif /* SP_FEATURES_H */
