/*
 * ref_nnue_runtime_net.cpp -- TEST INFRASTRUCTURE: the reference's src/eval/nnue.cpp, included UNMODIFIED, plus one
 * function in the same translation unit that loads a LOGICAL network file at run time through the reference's own
 * NetworkLoader / Network::loadFrom (network.h:53-70) and sets the file-static "loaded" flag -- the stock engine can
 * only use the network incbin embeds at build time (nnue.cpp:52), and this sandbox has synthetic networks only.
 */
#include "eval/nnue.cpp"

#include <cstring>

#include "util/align.h"

namespace stormphrax::eval {
    bool oracleLoadNetwork(const std::byte* payload, usize size) {
        const auto need = Network::byteSize();
        if (size < need) {
            return false;
        }
        if (!s_loadedNetworkData) {
            s_loadedNetworkData = util::alignedAlloc<std::byte>(util::simd::kAlignment, need);
        }
        std::memcpy(s_loadedNetworkData, payload, need);
        nnue::NetworkLoader loader{s_loadedNetworkData, need};
        // prePermuted = false: the x86 FT permutation is applied here, as for a compressed network (network.h:58-67)
        if (!s_network.loadFrom(loader, false)) {
            return false;
        }
        s_networkLoaded = true;
        return true;
    }
} // namespace stormphrax::eval
