/*
 * sp_types.h -- wire formats shared by the C-ABI (sp_nnue.h), the CUDA kernels,
 * the host-side mirror of Stormphrax's eval API, and the test oracles.
 *
 * Encodings follow the reference (Stormphrax 8.0.2) so that records produced by
 * the engine can be handed to the library unchanged:
 *   Color      black = 0, white = 1                         (src/core.h:73-75)
 *   PieceType  P,N,B,R,Q,K = 0..5                           (src/core.h:181-186)
 *   Piece      type << 1 | color, none = 12                 (src/core.h:337-359)
 *   Square     rank * 8 + file, a1 = 0, none = 64           (src/core.h:407-415)
 *   Move       from << 10 | to << 4 | (promo-1) << 2 | type (src/move.h:28-33,99-121)
 *              type: 0 standard, 1 promotion, 2 castling (king takes own rook), 3 en passant
 */
#ifndef SP_TYPES_H
#define SP_TYPES_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/*
 * 32-byte position record. Layout-identical to the reference's marlinformat
 * `PackedBoard` (src/datagen/marlinformat.h:32-84, src/util/u4array.h:58-66):
 *   occupancy   bitboard of occupied squares
 *   pieces      32 nibbles, one per occupied square in ascending square order;
 *               nibble i lives in byte i/2, even i = low nibble;
 *               value = type | (black ? 8 : 0), type 0..5 = P,N,B,R,Q,K,
 *               6 = rook that still has castling rights
 *   stm_ep      (black to move ? 0x80 : 0) | en-passant square (64 = none),
 *               ep square normalised to rank 6 (white to move) / rank 3 (black to move)
 *   halfmove, fullmove, eval, wdl, extra: carried, not used by the NNUE path
 */
typedef struct SpPackedBoard {
    uint64_t occupancy;
    uint8_t pieces[16];
    uint8_t stm_ep;
    uint8_t halfmove;
    uint16_t fullmove;
    int16_t eval;
    uint8_t wdl;
    uint8_t extra;
} SpPackedBoard;

typedef uint16_t SpMove;

enum {
    SP_BLACK = 0,
    SP_WHITE = 1,
    SP_PIECE_NONE = 12,
    SP_SQUARE_NONE = 64,
};

/* Network architecture constants (src/eval/arch.h:36-82). */
enum {
    SP_L1_SIZE = 1024,          /* FT outputs per perspective */
    SP_L2_SIZE = 32,            /* L1 outputs (dual activation doubles this to 64) */
    SP_L3_SIZE = 64,            /* L2 outputs */
    SP_OUTPUT_BUCKETS = 8,
    SP_INPUT_BUCKETS = 16,
    SP_PSQ_PER_BUCKET = 704,    /* merged kings: 11 planes x 64 */
    SP_PSQ_FEATURES = 11264,    /* 16 x 704 */
    SP_PP_FEATURES = 4560,      /* 96 * 95 / 2 pawn-pair features */
    SP_THREAT_FEATURES = 64368, /* 59808 piece threats + 4560 pawn pairs */
    SP_MAX_THREAT_INDICES = 256, /* StaticVector<u16, 256> in nnue_state.cpp:315 */
    SP_NET_HEADER_BYTES = 64,
    SP_NET_PAYLOAD_BYTES = 89381920,
};

#ifdef __cplusplus
}
#endif

#endif /* SP_TYPES_H */
