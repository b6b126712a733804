/*
 * eval/nnue_state.h for an SP_EVAL_B200 build of the reference engine -- TEST INFRASTRUCTURE (oracle/engine/Makefile).
 *
 * Same class and member names as /root/reference/src/eval/nnue_state.h:28-116, implemented over the B200 library's
 * C++ mirror.  The observer callbacks are no-ops: the device derives add / sub feature lists from (board stored in the
 * ancestor slot, board of the evaluated position), so nothing has to be recorded while Position::applyMove runs.
 */
#pragma once

#include "../types.h"

#include <memory>

#include "../core.h"
#include "nnue.h"

namespace stormphrax::eval {
    struct UpdateContext {};

    struct BoardObserver {
        UpdateContext& ctx;

        inline void prepareKingMove(Color, Square, Square) {}

        inline void pieceAdded(const Position&, Piece, Square) {}
        inline void pieceRemoved(const Position&, Piece, Square) {}
        inline void pieceMutated(const Position&, Piece, Piece, Square) {}
        inline void pieceMoved(const Position&, Piece, Square, Square) {}
        inline void piecePromoted(const Position&, Piece, Square, Piece, Square) {}

        inline void finalize(const Position&, const Position&) {}
    };

    class NnueState {
    public:
        NnueState();
        ~NnueState();

        NnueState(const NnueState&) = delete;
        NnueState& operator=(const NnueState&) = delete;

        // claims 256 device accumulator slots (the reference's stack depth, nnue_state.h:88) the first time
        void setNetwork(const Network* network);

        void reset(const Position& pos);

        [[nodiscard]] BoardObserver push();
        void pop();

        void applyImmediately(const UpdateContext& ctx, const Position& pos);

        [[nodiscard]] i32 evaluate(const Position& pos, Color stm);

        [[nodiscard]] static i32 evaluateOnce(const Position& pos, Color stm);

    private:
        struct Impl;
        std::unique_ptr<Impl> m_impl;
        UpdateContext m_ctx{};
    };
} // namespace stormphrax::eval
