/*
 * eval/batch.h -- TEST INFRASTRUCTURE (oracle/engine): run many searches of the reference engine as fibers of one host thread and
 * answer all their static evaluations with ONE device batch per round ("host fibers first").
 *
 * The reference's Searcher is used unmodified: inside runFibers() every NnueState::evaluate() queues its request on the round's
 * EvalBatch (stormphrax_b200/csrc/host/nnue_state.h) and switches back to the scheduler; when every live fiber is waiting the
 * scheduler flushes the batch (one sp_nnue_batch submission) and resumes them.  reset() / applyImmediately() become lazy inside
 * fibers (the device derives whatever delta is due at the next evaluation).  The CPU build has the same entry point and simply
 * runs the jobs one after the other (on `threads` host threads).
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

    // runs every job to completion; stackBytes per fiber.  threads > 1: that many host threads, each a scheduler of its own over
    // every threads-th job with its own evaluator context on the device (the reference binds a search thread to the network
    // copy of its NUMA node, search.cpp:206; here the unit is the scheduler thread): datagen's "N threads" on one GPU.
    // Room for `states` more NnueStates on the context of eval::init in one step (each claims 256 accumulator slots when its Searcher is
    // made; growing the device's slot store one state at a time copies it every time).
    void reserveStates(u32 states);

    // width: fibers alive per scheduler; a fiber whose job has returned takes the next job of the shared queue, so a round keeps
    // about `width` evaluations until the queue runs dry (0: every job gets its own fiber from the start).
    Stats runFibers(std::vector<std::function<void()>>& jobs, usize stackBytes = usize{4} << 20, u32 threads = 1, usize width = 0);
} // namespace stormphrax::eval::batch
