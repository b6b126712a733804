/*
 * adapter.cpp -- TEST INFRASTRUCTURE: the reference engine's eval API (src/eval/nnue.h:38-45, nnue_state.h:85-116) on
 * top of the B200 library's C++ mirror.  Positions cross as the engine's own 32-byte marlinformat record
 * (src/datagen/marlinformat.h:43-84), which is the library's wire format.
 */
#include <eval/nnue_state.h> // angle brackets throughout: these must be found through the shadow tree, not next to this file

#include <eval/batch.h>

#include <atomic>
#include <cstdlib>
#include <cstring>
#include <memory>

#include <sys/mman.h>
#include <ucontext.h>

#include <datagen/marlinformat.h>
#include <position.h>

#include <nnue_state.h> // the mirror: stormphrax_b200/csrc/host/nnue_state.h (found through -I, after the shadow tree)

namespace mirror = sp::host::eval;

namespace stormphrax::eval {
    namespace {
        std::atomic<u32> s_nextSlot{0};

        // ---- fibers: one scheduler per host thread
        struct Fiber {
            ucontext_t context{};
            void* stack{};
            usize stackBytes{};
            std::function<void()>* job{};
            bool done{false};
        };

        struct Scheduler {
            ucontext_t main{};
            Fiber* current{};
            mirror::EvalBatch* batch{};
            u64 queued{};
        };

        thread_local Scheduler* t_scheduler = nullptr;

        void fiberEntry() {
            auto* self = t_scheduler->current;
            (*self->job)();
            self->done = true;
            swapcontext(&self->context, &t_scheduler->main);
        }

        // called from NnueState::evaluate inside a fiber: back to the scheduler until the round's batch has been flushed
        void yieldToScheduler() {
            auto* self = t_scheduler->current;
            swapcontext(&self->context, &t_scheduler->main);
        }

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
        if (t_scheduler) {
            m_impl->state.invalidate(); // lazily: the next evaluation rebuilds from its own board, inside the round's batch
            return;
        }
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
        if (t_scheduler) {
            m_impl->state.applyLazily();
            return;
        }
        m_impl->state.applyPacked(pack(pos));
    }

    i32 NnueState::evaluate(const Position& pos, Color stm) {
        if (t_scheduler) {
            i32 out = 0;
            m_impl->state.evaluateAsyncPacked(*t_scheduler->batch, pack(pos), stm.raw(), &out);
            ++t_scheduler->queued;
            yieldToScheduler(); // resumed after EvalBatch::flush() has written `out`
            return out;
        }
        return m_impl->state.evaluatePacked(pack(pos), stm.raw());
    }

    i32 NnueState::evaluateOnce(const Position& pos, Color stm) {
        return mirror::NnueState::evaluateOncePacked(pack(pos), stm.raw());
    }
    namespace batch {
        Stats runFibers(std::vector<std::function<void()>>& jobs, usize stackBytes) {
            Stats stats{};
            Scheduler scheduler{};
            mirror::EvalBatch evalBatch{mirror::getNetwork()};
            scheduler.batch = &evalBatch;
            std::vector<Fiber> fibers(jobs.size());
            for (usize i = 0; i < jobs.size(); ++i) {
                auto& f = fibers[i];
                f.job = &jobs[i];
                f.stackBytes = stackBytes;
                f.stack = mmap(nullptr, stackBytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_STACK | MAP_NORESERVE, -1, 0);
                if (f.stack == MAP_FAILED) {
                    eprintln("runFibers: cannot map a fiber stack");
                    std::abort();
                }
                getcontext(&f.context);
                f.context.uc_stack.ss_sp = f.stack;
                f.context.uc_stack.ss_size = stackBytes;
                f.context.uc_link = nullptr;
                makecontext(&f.context, fiberEntry, 0);
            }
            t_scheduler = &scheduler;
            usize live = fibers.size();
            while (live > 0) {
                for (auto& f : fibers) {
                    if (f.done) {
                        continue;
                    }
                    scheduler.current = &f;
                    swapcontext(&scheduler.main, &f.context); // until it asks for an evaluation or finishes
                    if (f.done) {
                        --live;
                    }
                }
                if (evalBatch.pending() > 0) {
                    if (evalBatch.flush() != SP_OK) {
                        eprintln("runFibers: EvalBatch::flush failed: {}", mirror::lastError());
                        std::abort();
                    }
                    ++stats.rounds;
                }
            }
            stats.evaluations = scheduler.queued;
            t_scheduler = nullptr;
            for (auto& f : fibers) {
                munmap(f.stack, f.stackBytes);
            }
            return stats;
        }
    } // namespace batch
} // namespace stormphrax::eval
