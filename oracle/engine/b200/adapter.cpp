/*
 * adapter.cpp -- TEST INFRASTRUCTURE: the reference engine's eval API (src/eval/nnue.h:38-45, nnue_state.h:85-116) on
 * top of the B200 library's C++ mirror.  Positions cross as the engine's own 32-byte marlinformat record
 * (src/datagen/marlinformat.h:43-84), which is the library's wire format.
 */
#include <eval/nnue_state.h> // angle brackets throughout: these must be found through the shadow tree, not next to this file

#include <atomic>
#include <cstring>

#include <datagen/marlinformat.h>
#include <position.h>

#include <nnue_state.h> // the mirror: stormphrax_b200/csrc/host/nnue_state.h (found through -I, after the shadow tree)

namespace mirror = sp::host::eval;

namespace stormphrax::eval {
    namespace {
        std::atomic<u32> s_nextSlot{0};

        SpPackedBoard pack(const Position& pos) {
            const auto board = datagen::marlinformat::PackedBoard::pack(pos, 0);
            SpPackedBoard out;
            static_assert(sizeof(out) == sizeof(board));
            std::memcpy(&out, &board, sizeof(out));
            return out;
        }
    } // namespace

    void init() {}

    void shutdown() {
        mirror::shutdown();
    }

    bool initB200(const void* image, usize len, int device) {
        return mirror::init(image, len, device);
    }

    bool isNetworkLoaded() {
        return mirror::isNetworkLoaded();
    }

    const Network* getNetwork(u32) {
        return mirror::getNetwork();
    }

    std::string_view defaultNetworkName() {
        return "b200";
    }

    struct NnueState::Impl {
        mirror::NnueState state{};
        bool bound{false};
    };

    NnueState::NnueState() : m_impl{std::make_unique<Impl>()} {}
    NnueState::~NnueState() = default;

    void NnueState::setNetwork(const Network* network) {
        if (m_impl->bound) {
            return;
        }
        const auto base = s_nextSlot.fetch_add(mirror::NnueState::kStackDepth);
        m_impl->state.setNetwork(const_cast<Network*>(network), base, mirror::NnueState::kStackDepth);
        m_impl->bound = true;
    }

    void NnueState::reset(const Position& pos) {
        m_impl->state.resetPacked(pack(pos));
    }

    BoardObserver NnueState::push() {
        (void)m_impl->state.push();
        return BoardObserver{m_ctx};
    }

    void NnueState::pop() {
        m_impl->state.pop();
    }

    void NnueState::applyImmediately(const UpdateContext&, const Position& pos) {
        m_impl->state.applyPacked(pack(pos));
    }

    i32 NnueState::evaluate(const Position& pos, Color stm) {
        return m_impl->state.evaluatePacked(pack(pos), stm.raw());
    }

    i32 NnueState::evaluateOnce(const Position& pos, Color stm) {
        return mirror::NnueState::evaluateOncePacked(pack(pos), stm.raw());
    }
} // namespace stormphrax::eval
