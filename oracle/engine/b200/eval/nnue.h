/*
 * eval/nnue.h for an SP_EVAL_B200 build of the reference engine -- TEST INFRASTRUCTURE (oracle/engine/Makefile).
 *
 * Takes the place of /root/reference/src/eval/nnue.h in the shadow source tree the recipe assembles: the engine's
 * own search.cpp / thread.cpp / position.cpp / datagen.cpp / bench.cpp / uci.cpp then compile UNCHANGED against the
 * B200 library through its C++ mirror (stormphrax_b200/csrc/host/nnue_state.h).  This is the adapter INTEGRATION.md
 * section 3 describes, compiled for real.
 */
#pragma once

#include "../types.h"

#include <algorithm> // the reference's nnue.h brings these in for everything that includes eval.h
#include <array>
#include <cassert>
#include <cstring>
#include <memory>
#include <span>
#include <string_view>
#include <vector>

#include "../core.h"
#include "../position.h"

struct SpNnue;

namespace stormphrax::eval {
    using Network = SpNnue; // what getNetwork() hands to NnueState::setNetwork: the device context

    void init(); // the driver loads the network (oracle/engine/engine_main.cpp); kept for main.cpp's call shape
    void shutdown();

    [[nodiscard]] bool isNetworkLoaded();

    const Network* getNetwork(u32 numaId);

    [[nodiscard]] std::string_view defaultNetworkName();

    // driver entry: upload a LOGICAL network image to GPU `device`
    bool initB200(const void* image, usize len, int device);
} // namespace stormphrax::eval
