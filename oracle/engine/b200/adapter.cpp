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
#include <algorithm>
#include <memory>
#include <thread>
#include <vector>

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
            bool finished{false};       // its job has returned: the scheduler hands it the next one from the queue, or retires it
            bool done{false};           // retired
            std::vector<u32> slotBases; // slot ranges its job's states claimed on this scheduler's context (handed back with the job)
        };

        struct Scheduler {
            ucontext_t main{};
            Fiber* current{};
            mirror::EvalBatch* batch{};
            u64 queued{};
            SpNnue* network{}; // this host thread's evaluator context; states met inside its fibers are (re)bound to it
            u32 nextSlot{};    // next never-used accumulator slot of that context
            std::vector<u32> freeBases; // slot ranges of finished jobs, reused before nextSlot moves on
            u64 id{};          // 0: the context of eval::init (lives on); otherwise unique per run (those contexts are destroyed after it)
        };

        thread_local Scheduler* t_scheduler = nullptr;
        std::atomic<u64> s_nextSchedulerId{1};

        void fiberEntry() {
            for (;;) { // resumed after `finished` only with a new job
                auto* self = t_scheduler->current;
                (*self->job)();
                self->finished = true;
                swapcontext(&self->context, &t_scheduler->main);
            }
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
        u64 schedulerId{}; // whose context `state`'s slots belong to (Scheduler::id)
        bool bound{false};

        // A scheduler thread other than the first evaluates on its own context: the state moves there (fresh slots, nothing
        // carried over -- only legal at stack level 0, where reset() of the search / game about to run has just put it).
        void rebind(Scheduler& scheduler) {
            if (state.depth() != 0) {
                eprintln("NnueState: first use inside a fiber of another scheduler thread must be reset()");
                std::abort();
            }
            u32 base;
            if (!scheduler.freeBases.empty()) {
                base = scheduler.freeBases.back();
                scheduler.freeBases.pop_back();
            } else if (scheduler.id == 0) { // the context of eval::init: its slots are handed out process-wide (setNetwork)
                base = s_nextSlot.fetch_add(mirror::NnueState::kStackDepth);
            } else {
                base = scheduler.nextSlot;
                scheduler.nextSlot += mirror::NnueState::kStackDepth;
            }
            state.setNetwork(scheduler.network, base, mirror::NnueState::kStackDepth);
            scheduler.current->slotBases.push_back(base);
            schedulerId = scheduler.id;
        }

        // outside fibers again after a threaded run: that run's context is gone, back to the one of eval::init (fresh slots)
        void home() {
            if (schedulerId == 0) {
                return;
            }
            state.setNetwork(mirror::getNetwork(), s_nextSlot.fetch_add(mirror::NnueState::kStackDepth), mirror::NnueState::kStackDepth);
            schedulerId = 0;
        }
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
            if (m_impl->schedulerId != t_scheduler->id) {
                m_impl->state.invalidate();
                m_impl->rebind(*t_scheduler);
            }
            m_impl->state.invalidate(); // lazily: the next evaluation rebuilds from its own board, inside the round's batch
            return;
        }
        m_impl->home();
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
        m_impl->home();
        m_impl->state.applyPacked(pack(pos));
    }

    i32 NnueState::evaluate(const Position& pos, Color stm) {
        if (t_scheduler) {
            if (m_impl->schedulerId != t_scheduler->id) {
                m_impl->rebind(*t_scheduler);
            }
            i32 out = 0;
            m_impl->state.evaluateAsyncPacked(*t_scheduler->batch, pack(pos), stm.raw(), &out);
            ++t_scheduler->queued;
            yieldToScheduler(); // resumed after EvalBatch::flush() has written `out`
            return out;
        }
        m_impl->home();
        return m_impl->state.evaluatePacked(pack(pos), stm.raw());
    }

    i32 NnueState::evaluateOnce(const Position& pos, Color stm) {
        return mirror::NnueState::evaluateOncePacked(pack(pos), stm.raw());
    }
    namespace batch {
        namespace {
            // one scheduler: the calling host thread runs jobs from the shared queue as fibers against `network`, at most `width` at a
            // time (a fiber whose job has returned takes the next one: rounds stay full until the queue is empty)
            Stats runScheduler(
                std::vector<std::function<void()>>& jobs,
                std::atomic<usize>& nextJob,
                usize width,
                usize stackBytes,
                SpNnue* network,
                u32 firstSlot
            ) {
                Stats stats{};
                Scheduler scheduler{};
                mirror::EvalBatch evalBatch{network};
                scheduler.batch = &evalBatch;
                scheduler.network = network;
                scheduler.nextSlot = firstSlot;
                scheduler.id = network == mirror::getNetwork() ? 0 : s_nextSchedulerId.fetch_add(1);
                std::vector<Fiber> fibers;
                fibers.reserve(width);
                for (usize i = 0; i < width; ++i) {
                    const auto job = nextJob.fetch_add(1);
                    if (job >= jobs.size()) {
                        break;
                    }
                    auto& f = fibers.emplace_back();
                    f.job = &jobs[job];
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
                        swapcontext(&scheduler.main, &f.context); // until it asks for an evaluation or its job returns
                        if (f.finished) {
                            scheduler.freeBases.insert(scheduler.freeBases.end(), f.slotBases.begin(), f.slotBases.end());
                            f.slotBases.clear();
                            const auto job = nextJob.fetch_add(1);
                            if (job < jobs.size()) {
                                f.job = &jobs[job]; // starts at the next pass over the fibers
                                f.finished = false;
                            } else {
                                f.done = true;
                                --live;
                            }
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
        } // namespace

        void reserveStates(u32 states) {
            const auto slots = usize{s_nextSlot.load()} + usize{states} * mirror::NnueState::kStackDepth;
            if (sp_nnue_slots_reserve(mirror::getNetwork(), slots) != SP_OK) {
                eprintln("reserveStates: {}", sp_nnue_last_error(mirror::getNetwork()));
                std::abort();
            }
        }

        Stats runFibers(std::vector<std::function<void()>>& jobs, usize stackBytes, u32 threads, usize width) {
            threads = std::max<u32>(1, std::min<u32>(threads, static_cast<u32>(std::max<usize>(jobs.size(), 1))));
            if (width == 0) {
                width = (jobs.size() + threads - 1) / threads; // everything at once
            }
            std::atomic<usize> nextJob{0};
            if (threads == 1) {
                // states were bound to the first context when their Searcher was made; slots past theirs stay free for rebinds
                return runScheduler(jobs, nextJob, width, stackBytes, mirror::getNetwork(), s_nextSlot.load());
            }
            // Scheduler thread 0 keeps the context of eval::init; every other one gets its own (network copy, stream, slot store),
            // made and sized here: nothing but the job queue is shared between schedulers afterwards.
            std::vector<SpNnue*> contexts(threads, mirror::getNetwork());
            for (u32 t = 1; t < threads; ++t) {
                contexts[t] = mirror::createContext();
                if (!contexts[t] || sp_nnue_slots_reserve(contexts[t], width * mirror::NnueState::kStackDepth) != SP_OK) {
                    eprintln("runFibers: cannot create evaluator context {}: {}", t, mirror::lastError());
                    std::abort();
                }
            }
            std::vector<Stats> perThread(threads);
            std::vector<std::thread> workers;
            for (u32 t = 1; t < threads; ++t) {
                workers.emplace_back([&, t] { perThread[t] = runScheduler(jobs, nextJob, width, stackBytes, contexts[t], 0); });
            }
            perThread[0] = runScheduler(jobs, nextJob, width, stackBytes, contexts[0], s_nextSlot.load());
            Stats stats{};
            for (auto& worker : workers) {
                worker.join();
            }
            for (u32 t = 0; t < threads; ++t) {
                stats.rounds += perThread[t].rounds;
                stats.evaluations += perThread[t].evaluations;
                if (t > 0) {
                    mirror::destroyContext(contexts[t]);
                }
            }
            return stats;
        }
    } // namespace batch
} // namespace stormphrax::eval
