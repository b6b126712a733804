/*
 * eval/batch.h -- TEST INFRASTRUCTURE (oracle/engine): run many searches of the reference engine as fibers of one host thread and
 * answer all their static evaluations with ONE device batch per round ("host fibers first").
 *
 * The reference's Searcher is used unmodified: inside runFibers() every NnueState::evaluate() queues its request on the round's
 * EvalBatch (stormphrax_b200/csrc/host/nnue_state.h) and switches back to the scheduler; when every live fiber is waiting the
 * scheduler flushes the batch (one sp_nnue_batch submission) and resumes them.  reset() / applyImmediately() become lazy inside
 * fibers (the device derives whatever delta is due at the next evaluation).  The CPU build has the same entry point and simply
 * runs the jobs one after the other.
 */
#pragma once

#include "../types.h"

#include <functional>
#include <vector>

namespace stormphrax::eval::batch {
    struct Stats {
        u64 rounds{};      // device batches submitted
        u64 evaluations{}; // static evaluations answered through them
    };

    // runs every job to completion; stackBytes per fiber
    Stats runFibers(std::vector<std::function<void()>>& jobs, usize stackBytes = usize{4} << 20);
} // namespace stormphrax::eval::batch
