/*
 * nnue_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A slow, obviously-correct, plain-C restatement of the NNUE evaluation path of
 * Stormphrax 8.0.2.  Each function cites the reference file:line it follows.  It exists so
 * that GPU results can be checked on machines where /root/reference (and therefore
 * oracle/_ref) is absent, and to dump intermediates (feature lists, accumulators, FT
 * activations) that the reference does not expose.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this file against the reference's own
 * compiled code (oracle/_ref/libsp_ref_*.so) -- feature lists element for element, evals bit
 * for bit -- and against tests/golden/ fixtures generated from it (tests/golden/make_golden.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may use it.
 *
 * All arithmetic is two's-complement with explicit widths; compile with -fwrapv.
 */
#include "nnue_oracle.h"

#include <limits.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

enum { L1 = SP_L1_SIZE, L2 = SP_L2_SIZE, L3 = SP_L3_SIZE, NB = SP_OUTPUT_BUCKETS };
enum { PAWN = 0, KNIGHT, BISHOP, ROOK, QUEEN, KING };
enum { NONE = SP_PIECE_NONE };

typedef uint64_t u64;

/* ------------------------------------------------------------------ network */

static const int16_t* g_psq_w;  /* [11264][1024] */
static const int8_t* g_thr_w;   /* [64368][1024] */
static const int16_t* g_ft_b;   /* [1024] */
static const int8_t* g_l1_w;    /* [8][(1024/4)][32][4]  multilayer.h:180-196 */
static const int32_t* g_l1_b;   /* [8][32] */
static const int32_t* g_l2_w;   /* [8][64][64] */
static const int32_t* g_l2_b;   /* [8][64] */
static const int32_t* g_l3_w;   /* [8][64] */
static const int32_t* g_l3_b;   /* [8] */
static unsigned char* g_image;

/* Array order and sizes: src/eval/nnue/input.h:359-361, multilayer.h:492-496; header.h:38-52. */
int spo_load_net(const void* image, size_t len) {
    if (len < (size_t)SP_NET_HEADER_BYTES + SP_NET_PAYLOAD_BYTES) return 1;
    if (memcmp(image, "CBNF", 4) != 0) return 2;
    free(g_image);
    g_image = (unsigned char*)malloc(SP_NET_PAYLOAD_BYTES);
    if (!g_image) return 3;
    memcpy(g_image, (const unsigned char*)image + SP_NET_HEADER_BYTES, SP_NET_PAYLOAD_BYTES);
    const unsigned char* p = g_image;
    g_psq_w = (const int16_t*)p; p += (size_t)SP_PSQ_FEATURES * L1 * 2;
    g_thr_w = (const int8_t*)p;  p += (size_t)SP_THREAT_FEATURES * L1;
    g_ft_b = (const int16_t*)p;  p += L1 * 2;
    g_l1_w = (const int8_t*)p;   p += NB * L1 * L2;
    g_l1_b = (const int32_t*)p;  p += NB * L2 * 4;
    g_l2_w = (const int32_t*)p;  p += NB * 2 * L2 * L3 * 4;
    g_l2_b = (const int32_t*)p;  p += NB * L3 * 4;
    g_l3_w = (const int32_t*)p;  p += NB * L3 * 4;
    g_l3_b = (const int32_t*)p;  p += NB * 4;
    return (size_t)(p - g_image) == SP_NET_PAYLOAD_BYTES ? 0 : 4;
}

/* ------------------------------------------------------------------ board */

typedef struct {
    uint8_t mailbox[64]; /* Piece = type << 1 | color (core.h:337-349), NONE = 12 */
    u64 occ;
    int king[2];
    int stm;
} Board;

/* marlinformat PackedBoard decode (src/datagen/marlinformat.h:43-77) */
static int decode(const SpPackedBoard* pb, Board* b) {
    memset(b->mailbox, NONE, sizeof(b->mailbox));
    b->occ = pb->occupancy;
    b->king[0] = b->king[1] = SP_SQUARE_NONE;
    b->stm = (pb->stm_ep & 0x80) ? SP_BLACK : SP_WHITE;
    u64 occ = pb->occupancy;
    int i = 0;
    while (occ) {
        if (i >= 32) return 1;
        const int sq = __builtin_ctzll(occ);
        occ &= occ - 1;
        const unsigned nib = (pb->pieces[i / 2] >> ((i % 2) * 4)) & 0xF;
        ++i;
        unsigned type = nib & 7;
        if (type == 6) type = ROOK;
        if (type > KING) return 2;
        const int color = (nib & 8) ? SP_BLACK : SP_WHITE;
        b->mailbox[sq] = (uint8_t)(type << 1 | (unsigned)color);
        if (type == KING) b->king[color] = sq;
    }
    return (b->king[0] == SP_SQUARE_NONE || b->king[1] == SP_SQUARE_NONE) ? 3 : 0;
}

/* ------------------------------------------------------------------ attacks (src/attacks/attacks.h) */

static u64 ray_attacks(int sq, u64 occ, const int (*dirs)[2], int ndirs) {
    u64 out = 0;
    for (int d = 0; d < ndirs; ++d) {
        int f = (sq & 7) + dirs[d][0], r = (sq >> 3) + dirs[d][1];
        while (f >= 0 && f < 8 && r >= 0 && r < 8) {
            const u64 bit = (u64)1 << (r * 8 + f);
            out |= bit;
            if (occ & bit) break; /* blockers are included, then the ray stops */
            f += dirs[d][0];
            r += dirs[d][1];
        }
    }
    return out;
}

static u64 step_attacks(int sq, const int (*steps)[2], int n) {
    u64 out = 0;
    for (int i = 0; i < n; ++i) {
        const int f = (sq & 7) + steps[i][0], r = (sq >> 3) + steps[i][1];
        if (f >= 0 && f < 8 && r >= 0 && r < 8) out |= (u64)1 << (r * 8 + f);
    }
    return out;
}

static const int kDiag[4][2] = {{1, 1}, {1, -1}, {-1, 1}, {-1, -1}};
static const int kOrth[4][2] = {{1, 0}, {-1, 0}, {0, 1}, {0, -1}};
static const int kKnight[8][2] = {{1, 2}, {2, 1}, {2, -1}, {1, -2}, {-1, -2}, {-2, -1}, {-2, 1}, {-1, 2}};
static const int kKing[8][2] = {{1, 0}, {1, 1}, {0, 1}, {-1, 1}, {-1, 0}, {-1, -1}, {0, -1}, {1, -1}};

/* attacks::getAttacks(piece, src, occ), attacks.h:130-151 */
static u64 get_attacks(int piece, int sq, u64 occ) {
    const int type = piece >> 1, color = piece & 1;
    switch (type) {
        case PAWN: { /* generatePawnAttacks, attacks.h:37-48: up-left/up-right relative to colour */
            const int up = color == SP_WHITE ? 1 : -1;
            const int steps[2][2] = {{-1, up}, {1, up}};
            return step_attacks(sq, steps, 2);
        }
        case KNIGHT: return step_attacks(sq, kKnight, 8);
        case BISHOP: return ray_attacks(sq, occ, kDiag, 4);
        case ROOK: return ray_attacks(sq, occ, kOrth, 4);
        case QUEEN: return ray_attacks(sq, occ, kDiag, 4) | ray_attacks(sq, occ, kOrth, 4);
        case KING: return step_attacks(sq, kKing, 8);
        default: return 0;
    }
}

/* ------------------------------------------------------------------ PSQ features */

/* eval/arch.h:53-65 (half-board, files a-d, a1 first) mirrored onto e-h by psq.h:209-225 */
static const uint8_t kHalfBuckets[32] = {0, 1, 2,  3,  4,  5,  6,  7,  8,  9,  10, 11, 8,  9,  10, 11,
                                         12, 12, 13, 13, 12, 12, 13, 13, 14, 14, 15, 15, 14, 14, 15, 15};

static int king_bucket(int c, int king_sq) { /* psq.h:241-246 */
    if (c == SP_BLACK) king_sq ^= 56;
    const int file = king_sq & 7, rank = king_sq >> 3;
    return kHalfBuckets[rank * 4 + (file < 4 ? file : 7 - file)];
}

/* psq::featureIndex<KingBucketsMergedMirrored<kAbcd,...>>, psq.h:338-365 */
static uint32_t psq_index(int c, int piece, int sq, int king_sq) {
    const int type = piece >> 1;
    const uint32_t color = (type == KING) ? 0u : ((piece & 1) == c ? 0u : 1u); /* merged kings :349-354 */
    if (c == SP_BLACK) sq ^= 56;                                               /* :356-358 */
    if ((king_sq & 7) > 3) sq ^= 7;                                            /* psq.h:227-238, kAbcd */
    return (uint32_t)king_bucket(c, king_sq) * SP_PSQ_PER_BUCKET + color * 384u + (uint32_t)type * 64u + (uint32_t)sq;
}

/* ------------------------------------------------------------------ threat features (threats.cpp) */

/* threats.cpp:42-54: with pawn-pair inputs enabled the reference selects the "NoPpThreats" map */
static const int kTargetMap[6][6] = {
    {-1, 0, -1, 1, -1, -1}, {0, 1, 2, 3, 4, -1},  {0, 1, 2, 3, -1, -1},
    {0, 1, 2, 3, -1, -1},   {0, 1, 2, 3, 4, -1},  {-1, -1, -1, -1, -1, -1},
};
static int g_target_count[6];          /* threats.cpp:56-72 */
static uint8_t g_piece_idx[12][64][64]; /* threats.cpp:74-106 */
static uint32_t g_offsets[12][64];     /* threats.cpp:108-136 */
static int32_t g_piece_total[12], g_piece_base[12];
static int32_t g_attack_idx[12][12][2]; /* threats.cpp:138-167 */
static pthread_once_t g_tables_once = PTHREAD_ONCE_INIT;

static void build_tables(void) {
    for (int t = 0; t < 6; ++t) {
        int n = 0;
        for (int v = 0; v < 6; ++v) n += kTargetMap[t][v] >= 0;
        g_target_count[t] = 2 * n;
    }
    /* kPieceIndices: number of empty-board attack squares below `to` */
    for (int piece = 0; piece < 12; ++piece)
        for (int from = 0; from < 64; ++from) {
            const u64 pseudo = get_attacks(piece, from, 0);
            for (int to = 0; to < 64; ++to)
                g_piece_idx[piece][from][to] = (uint8_t)__builtin_popcountll(pseudo & (((u64)1 << to) - 1));
        }
    /* kOffsets: white pieces first, then black (threats.cpp:116) */
    int32_t offset = 0;
    for (int ci = 0; ci < 2; ++ci) {
        const int color = ci == 0 ? SP_WHITE : SP_BLACK;
        for (int pt = 0; pt < 6; ++pt) {
            const int piece = pt << 1 | color;
            int32_t piece_offset = 0;
            for (int sq = 0; sq < 64; ++sq) {
                g_offsets[piece][sq] = (uint32_t)piece_offset;
                const int rank = sq >> 3;
                if (pt != PAWN || (rank > 0 && rank < 7))
                    piece_offset += __builtin_popcountll(get_attacks(piece ^ 1, sq, 0)); /* flipColor, :124 */
            }
            g_piece_total[piece] = piece_offset;
            g_piece_base[piece] = offset;
            offset += g_target_count[pt] * piece_offset;
        }
    }
    for (int atk = 0; atk < 12; ++atk)
        for (int vic = 0; vic < 12; ++vic) {
            const int at = atk >> 1, vt = vic >> 1;
            const int enemy = (atk & 1) != (vic & 1);
            const int map = kTargetMap[at][vt];
            const int semi = at == vt && (enemy || at != PAWN);
            const int excluded = map < 0;
            const int vic_black = (vic & 1) == SP_BLACK; /* attacked.color().flip().raw(), :155 */
            const int32_t feature = g_piece_base[atk] + (vic_black * (g_target_count[at] / 2) + map) * g_piece_total[atk];
            g_attack_idx[atk][vic][0] = excluded ? INT32_MIN : feature;
            g_attack_idx[atk][vic][1] = (excluded || semi) ? INT32_MIN : feature;
        }
}

/* threats::threatFeatureIndex, threats.cpp:170-198. Negative result = feature does not exist. */
int32_t spo_threat_index(int c, int king_sq, int attacker, int asq, int attacked, int vsq) {
    pthread_once(&g_tables_once, build_tables);
    if (c == SP_BLACK) {
        attacker ^= 1;
        attacked ^= 1;
        asq ^= 56;
        vsq ^= 56;
    }
    if ((king_sq & 7) >= 4) {
        asq ^= 7;
        vsq ^= 7;
    }
    const int forwards = asq < vsq;
    const int32_t attack_idx = g_attack_idx[attacker][attacked][forwards];
    if (attack_idx == INT32_MIN) return -1; /* the reference lets INT_MIN + small stay negative */
    return SP_PP_FEATURES + attack_idx + (int32_t)g_offsets[attacker][asq] + g_piece_idx[attacker][asq][vsq];
}

/* threats::ppPawnId / ppFeatureIndex, threats.cpp:200-221 */
static uint32_t pp_id(int c, int king_sq, int pawn_color, int sq) {
    if (c == SP_BLACK) sq ^= 56;
    if ((king_sq & 7) >= 4) sq ^= 7;
    return (uint32_t)((c != pawn_color ? 48 : 0) + sq - 8);
}
static uint32_t pp_index(int c, int king_sq, int a, int asq, int b, int bsq) {
    const uint32_t ia = pp_id(c, king_sq, a, asq), ib = pp_id(c, king_sq, b, bsq);
    const uint32_t hi = ia > ib ? ia : ib, lo = ia > ib ? ib : ia;
    return hi * (hi - 1) / 2 + lo;
}

/* kPpMasks, threats.h:106-123: own file and both neighbours, all ranks; zero on ranks 1 and 8 */
static u64 pp_mask(int sq) {
    if (sq < 8 || sq >= 56) return 0;
    const int f = sq & 7;
    u64 m = 0;
    for (int df = -1; df <= 1; ++df)
        if (f + df >= 0 && f + df < 8) m |= 0x0101010101010101ULL << (f + df);
    return m;
}

static u64 pieces_of(const Board* b, int type, int color) {
    u64 out = 0;
    for (int sq = 0; sq < 64; ++sq)
        if (b->mailbox[sq] == (type << 1 | color)) out |= (u64)1 << sq;
    return out;
}

/* Enumeration order of addThreatFeatures, nnue_state.cpp:309-354 */
static int threat_features(const Board* b, int c, uint32_t* out) {
    pthread_once(&g_tables_once, build_tables);
    const int king_sq = b->king[c];
    const u64 kings = ((u64)1 << b->king[0]) | ((u64)1 << b->king[1]);
    int n = 0;
    for (u64 from_bb = b->occ & ~kings; from_bb; from_bb &= from_bb - 1) {
        const int from = __builtin_ctzll(from_bb);
        const int piece = b->mailbox[from];
        for (u64 to_bb = b->occ & get_attacks(piece, from, b->occ) & ~kings; to_bb; to_bb &= to_bb - 1) {
            const int to = __builtin_ctzll(to_bb);
            const int32_t f = spo_threat_index(c, king_sq, piece, from, b->mailbox[to], to);
            if (f >= 0) out[n++] = (uint32_t)f;
        }
    }
    const u64 ours = pieces_of(b, PAWN, c), theirs = pieces_of(b, PAWN, c ^ 1);
    for (u64 rem = ours; rem;) { /* iterWithRemaining: `remaining` excludes a itself */
        const int a = __builtin_ctzll(rem);
        rem &= rem - 1;
        const u64 mask = pp_mask(a);
        for (u64 bb = rem & mask; bb; bb &= bb - 1) out[n++] = pp_index(c, king_sq, c, a, c, __builtin_ctzll(bb));
        for (u64 bb = theirs & mask; bb; bb &= bb - 1) out[n++] = pp_index(c, king_sq, c, a, c ^ 1, __builtin_ctzll(bb));
    }
    for (u64 rem = theirs; rem;) {
        const int a = __builtin_ctzll(rem);
        rem &= rem - 1;
        for (u64 bb = rem & pp_mask(a); bb; bb &= bb - 1)
            out[n++] = pp_index(c, king_sq, c ^ 1, a, c ^ 1, __builtin_ctzll(bb));
    }
    return n;
}

/* Board iteration order of resetPsqAccumulator, nnue_state.cpp:440-449 (ascending squares) */
static int psq_features(const Board* b, int c, uint32_t* out) {
    int n = 0;
    for (u64 bb = b->occ; bb; bb &= bb - 1) {
        const int sq = __builtin_ctzll(bb);
        out[n++] = psq_index(c, b->mailbox[sq], sq, b->king[c]);
    }
    return n;
}

int spo_psq_features(const SpPackedBoard* board, int c, uint32_t* out) {
    Board b;
    if (decode(board, &b)) return -1;
    return psq_features(&b, c, out);
}

int spo_threat_features(const SpPackedBoard* board, int c, uint32_t* out) {
    Board b;
    if (decode(board, &b)) return -1;
    return threat_features(&b, c, out);
}

/* ------------------------------------------------------------------ accumulators */

/* evaluateOnce: initBoth + resetPsqAccumulator + resetThreatAccumulator, nnue_state.cpp:612-634.
 * PSQ accumulator starts from the FT bias (input.h:72-75), threat accumulator from zero
 * (applyThreatRows<kZeroInit = true>, nnue_state.cpp:89-145,353). int16 adds wrap. */
static void accumulate(const Board* b, int16_t psq[2][L1], int16_t thr[2][L1]) {
    uint32_t idx[512];
    for (int c = 0; c < 2; ++c) {
        for (int i = 0; i < L1; ++i) {
            psq[c][i] = g_ft_b[i];
            thr[c][i] = 0;
        }
        int n = psq_features(b, c, idx);
        for (int k = 0; k < n; ++k) {
            const int16_t* row = g_psq_w + (size_t)idx[k] * L1;
            for (int i = 0; i < L1; ++i) psq[c][i] = (int16_t)(uint16_t)((uint16_t)psq[c][i] + (uint16_t)row[i]);
        }
        n = threat_features(b, c, idx);
        for (int k = 0; k < n; ++k) {
            const int8_t* row = g_thr_w + (size_t)idx[k] * L1;
            for (int i = 0; i < L1; ++i)
                thr[c][i] = (int16_t)(uint16_t)((uint16_t)thr[c][i] + (uint16_t)(int16_t)row[i]); /* widenLoadI8ToI16 */
        }
    }
}

int spo_accumulators(const SpPackedBoard* board, int16_t* psq, int16_t* thr) {
    Board b;
    if (!g_image || decode(board, &b)) return 1;
    accumulate(&b, (int16_t(*)[L1])psq, (int16_t(*)[L1])thr);
    return 0;
}

/* ------------------------------------------------------------------ forward pass */

/* activateFt, multilayer.h:92-152.  a = clamp(a, 0, 255); d = min(d, 255) (NOT floored at zero);
 * p = ((a << 7) * d) >> 16 as a signed 32-bit product (shiftLeftMulHi, avx512.h:179-182);
 * packus saturates to [0, 255] (avx512.h:188-190).  stm half first (nnue_state.cpp:405-419). */
static void activate_ft(const int16_t psq[2][L1], const int16_t thr[2][L1], int stm, uint8_t ft[L1]) {
    for (int h = 0; h < 2; ++h) {
        const int c = h == 0 ? stm : stm ^ 1;
        for (int i = 0; i < L1 / 2; ++i) {
            int32_t a = (int16_t)(uint16_t)((uint16_t)psq[c][i] + (uint16_t)thr[c][i]);
            int32_t d = (int16_t)(uint16_t)((uint16_t)psq[c][i + L1 / 2] + (uint16_t)thr[c][i + L1 / 2]);
            a = a < 0 ? 0 : (a > 255 ? 255 : a);
            d = d > 255 ? 255 : d;
            int32_t p = ((a << 7) * d) >> 16; /* arithmetic shift: floor */
            p = p < 0 ? 0 : (p > 255 ? 255 : p);
            ft[h * (L1 / 2) + i] = (uint8_t)p;
        }
    }
}

/* propagateL1 / L2 / L3 and the final scale, multilayer.h:154-490.
 * All sums wrap in 32 bits, as the SIMD ops do. */
int32_t spo_forward(const uint8_t* ft, int bucket) {
    int32_t l1o[2 * L2], l2[L3];
    const int8_t* w1 = g_l1_w + (size_t)bucket * L1 * L2;
    for (int o = 0; o < L2; ++o) {
        uint32_t s = 0;
        for (int k = 0; k < L1; ++k) /* weight at [(k/4)][o][k%4], multilayer.h:180-196 */
            s += (uint32_t)((int32_t)ft[k] * (int32_t)w1[(k >> 2) * (L2 * 4) + o * 4 + (k & 3)]);
        const int32_t x = (int32_t)((uint32_t)((int32_t)s >> 2) + (uint32_t)g_l1_b[bucket * L2 + o]); /* kShift = -2, :162,228-230 */
        const int32_t cr = x < 0 ? 0 : (x > 4096 ? 4096 : x);
        l1o[o] = (int32_t)((uint32_t)cr << 6);                   /* :234-238 */
        int32_t sq = (int32_t)((uint32_t)x * (uint32_t)x);       /* mulLo wraps BEFORE the min, :242 */
        sq = sq > 16777216 ? 16777216 : sq;
        l1o[L2 + o] = sq >> 6;                                   /* :243-244 */
    }
    for (int o = 0; o < L3; ++o) {
        uint32_t s = (uint32_t)g_l2_b[bucket * L3 + o];
        for (int i = 0; i < 2 * L2; ++i) /* :268-301, input >>= kQuantBits */
            s += (uint32_t)(l1o[i] >> 6) * (uint32_t)g_l2_w[(size_t)bucket * 2 * L2 * L3 + (size_t)i * L3 + o];
        l2[o] = (int32_t)s;
    }
    uint32_t s3 = (uint32_t)g_l3_b[bucket];
    for (int o = 0; o < L3; ++o) { /* :353-446: clamp(l2, 0, 64^3) + skipped L1 output, times weight */
        const int32_t cl = l2[o] < 0 ? 0 : (l2[o] > 262144 ? 262144 : l2[o]);
        s3 += ((uint32_t)cl + (uint32_t)l1o[o]) * (uint32_t)g_l3_w[bucket * L3 + o];
    }
    const int64_t out = (int64_t)(int32_t)s3 * 400 / 16777216; /* :484-489, C division truncates */
    return (int32_t)out;
}

int32_t spo_forward_acc(const int16_t* psq, const int16_t* thr, int stm, int bucket) {
    uint8_t ft[L1];
    activate_ft((const int16_t(*)[L1])psq, (const int16_t(*)[L1])thr, stm, ft);
    return spo_forward(ft, bucket);
}

static int bucket_of(const Board* b) { /* MaterialCount<8>::getBucket, output.h:51-54 */
    return (__builtin_popcountll(b->occ) - 2) / 4;
}

int spo_ft_activations(const SpPackedBoard* board, uint8_t* out, int* bucket) {
    Board b;
    int16_t psq[2][L1], thr[2][L1];
    if (!g_image || decode(board, &b)) return 1;
    accumulate(&b, psq, thr);
    activate_ft(psq, thr, b.stm, out);
    if (bucket) *bucket = bucket_of(&b);
    return 0;
}

static int eval_one(const SpPackedBoard* board, int32_t* out) {
    uint8_t ft[L1];
    int bucket;
    if (spo_ft_activations(board, ft, &bucket)) return 1;
    *out = spo_forward(ft, bucket);
    return 0;
}

int spo_eval_once(const SpPackedBoard* boards, size_t n, int32_t* out) {
    if (!g_image) return 1;
    for (size_t i = 0; i < n; ++i)
        if (eval_one(&boards[i], &out[i])) return 2;
    return 0;
}

/* ------------------------------------------------------------------ timing (CPU-baseline "port" leg) */

typedef struct {
    const SpPackedBoard* boards;
    int32_t* out;
    size_t lo, hi;
    double secs;
    int rc;
} Shard;

static void* shard_main(void* arg) {
    Shard* s = (Shard*)arg;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    for (size_t i = s->lo; i < s->hi; ++i) s->rc |= eval_one(&s->boards[i], &s->out[i]);
    clock_gettime(CLOCK_MONOTONIC, &t1);
    s->secs = (double)(t1.tv_sec - t0.tv_sec) + 1e-9 * (double)(t1.tv_nsec - t0.tv_nsec);
    return NULL;
}

double spo_time_eval_once(const SpPackedBoard* boards, size_t n, int threads, int reps, int32_t* out) {
    if (!g_image || threads < 1 || reps < 1) return -1.0;
    pthread_once(&g_tables_once, build_tables);
    double best = 1e30;
    pthread_t* tid = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)threads);
    Shard* sh = (Shard*)malloc(sizeof(Shard) * (size_t)threads);
    for (int rep = 0; rep < reps; ++rep) {
        double worst = 0;
        for (int t = 0; t < threads; ++t) {
            sh[t] = (Shard){boards, out, n * (size_t)t / (size_t)threads, n * (size_t)(t + 1) / (size_t)threads, 0, 0};
            pthread_create(&tid[t], NULL, shard_main, &sh[t]);
        }
        for (int t = 0; t < threads; ++t) {
            pthread_join(tid[t], NULL);
            if (sh[t].rc) best = -2.0;
            if (sh[t].secs > worst) worst = sh[t].secs;
        }
        if (best >= 0 && worst < best) best = worst;
    }
    free(tid);
    free(sh);
    return best;
}
