/*
 * selfplay_gpu.cu -- GPU-resident batched self-play: sp_selfplay_run_gpu (include/sp_nnue.h).
 *
 * The host driver (host/selfplay.cpp) is bound by the host cores that step the games' searches (16 cores
 * feed ~9 M static evaluations per second to a device that sustains > 150 M/s through the same slot API).
 * Here the games live on the device: one thread per game slot runs the SAME state machine
 * (selfplay::Game, host/selfplay.h -- board model, move generator, search, adjudication and record writer
 * all compile for host and device) until it needs a static evaluation, appends the request to one of three
 * device lists and returns; the host then only reads four counters and submits ONE sp_nnue_batch_device
 * (refresh / update / evaluate-only groups, exactly what EvalBatch::flush submits for the host driver) and a
 * scatter of the results into the games.  No position, move or evaluation crosses PCIe.
 *
 * Determinism: a game's random stream depends only on (seed, slot, game number), evaluations are exact
 * integers, and records are emitted slot-major -- the output is byte-identical to sp_selfplay_run's
 * (tests/test_selfplay.py).
 */
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <new>
#include <thread>
#include <vector>

#include "../../include/sp_nnue.h"
#include "host/selfplay.h"

namespace {

using namespace sp;
using namespace sp::host;
using namespace sp::host::selfplay;

constexpr uint32_t kSlotsPerGame = kMaxPly + 1;
#ifndef SP_SELFPLAY_MIN_BLOCKS
#define SP_SELFPLAY_MIN_BLOCKS 4 /* 128-thread blocks per SM the step kernel is compiled for (register budget) */
#endif

/* request lists of one round; capacity = number of game slots each (a game has one request per round) */
struct Requests {
    uint32_t* refresh_slot;
    SpPackedBoard* refresh_board;
    i32** refresh_out;
    uint32_t* update_src;
    uint32_t* update_dst;
    SpPackedBoard* update_board;
    i32** update_out;
    uint32_t* eval_slot;
    uint8_t* eval_stm;
    i32** eval_out;
    uint32_t* counters; /* [0] refresh, [1] update, [2] evaluate-only requests, [3] games still running */
};

/* The accumulator-stack bookkeeping of eval::NnueState (host/nnue_state.cpp: push / pop / applyLazily /
 * evaluateAsync with lazy catch-up from the nearest clean ancestor), one bit per stack level. */
struct GpuEvaluator {
    Requests rq;
    uint32_t* top;
    uint32_t* clean; /* bit l: slot of level l holds the accumulators of its stored board */
    uint32_t* stale; /* bit l: ... but that board is an ancestor of the level's position */

    __device__ void reset(uint32_t g, const Position&) { top[g] = 0, clean[g] = 0, stale[g] = 0; }
    __device__ NullObserver push(uint32_t g) {
        const uint32_t t = ++top[g];
        clean[g] &= ~(1u << t), stale[g] &= ~(1u << t);
        return NullObserver{};
    }
    __device__ void pop(uint32_t g) { --top[g]; }
    __device__ void applyImmediately(uint32_t g, const Position&) { stale[g] |= 1u << top[g]; }
    __device__ void evaluateAsync(uint32_t g, const Position& pos, i32* out) {
        const uint32_t t = top[g], base = g * kSlotsPerGame, dst = base + t;
        const uint32_t c = clean[g], s = stale[g];
        if (((c >> t) & 1) && !((s >> t) & 1)) { /* up to date: evaluate only */
            const uint32_t i = atomicAdd(&rq.counters[2], 1u);
            rq.eval_slot[i] = dst, rq.eval_stm[i] = static_cast<uint8_t>(pos.stm()), rq.eval_out[i] = out;
            return;
        }
        const uint32_t ancestors = c & ((2u << t) - 1); /* clean levels <= t (this very level when it is stale) */
        if (!ancestors) {
            const uint32_t i = atomicAdd(&rq.counters[0], 1u);
            rq.refresh_slot[i] = dst, rq.refresh_board[i] = pos.pack(), rq.refresh_out[i] = out;
        } else {
            const uint32_t from = 31u - static_cast<uint32_t>(__clz(static_cast<int>(ancestors)));
            const uint32_t i = atomicAdd(&rq.counters[1], 1u);
            rq.update_src[i] = base + from, rq.update_dst[i] = dst, rq.update_board[i] = pos.pack(), rq.update_out[i] = out;
        }
        clean[g] = c | (1u << t), stale[g] = s & ~(1u << t);
    }
};

using DeviceGame = Game<GpuEvaluator>;

struct Slots {
    DeviceGame* games;
    Stats* stats;        /* per slot */
    uint32_t* played;    /* games finished per slot */
    uint8_t* alive;
    uint8_t* records;    /* per slot: record_stride bytes, records back to back */
    uint32_t* record_len;
    size_t record_stride;
    uint32_t n;
};

__global__ void __launch_bounds__(128) selfplay_init_kernel(Slots sl, GpuEvaluator* ev, Params params) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sl.n) return;
    sl.stats[g] = Stats{};
    sl.played[g] = 0, sl.record_len[g] = 0;
    const uint32_t slot = params.firstSlot() + g; /* g is local to this driver instance */
    sl.alive[g] = params.gamesOfSlot(slot) ? 1 : 0;
    if (!sl.alive[g]) return;
    DeviceGame* game = new (&sl.games[g]) DeviceGame();
    game->start(g, params.gameSeed(slot, 0), params, ev, &sl.stats[g]);
}

/* every running game advances until it needs a static evaluation (or its slot has played all its games) */
__global__ void __launch_bounds__(128, SP_SELFPLAY_MIN_BLOCKS) selfplay_step_kernel(Slots sl, GpuEvaluator* ev, Params params) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= sl.n || !sl.alive[g]) return;
    DeviceGame& game = sl.games[g];
    while (game.step() == DeviceGame::Status::kGameOver) {
        sl.stats[g].games += 1;
        ViriGame& rec = game.record();
        rec.serialize(sl.records + g * sl.record_stride + sl.record_len[g], game.outcome());
        sl.record_len[g] += static_cast<uint32_t>(rec.bytes());
        const uint32_t slot = params.firstSlot() + g;
        if (++sl.played[g] >= params.gamesOfSlot(slot)) {
            sl.alive[g] = 0;
            return;
        }
        game.start(g, params.gameSeed(slot, sl.played[g]), params, ev, &sl.stats[g]);
    }
    atomicAdd(&ev->rq.counters[3], 1u);
}

__global__ void selfplay_scatter_kernel(Requests rq, const i32* refresh, uint32_t n_refresh, const i32* update, uint32_t n_update, const i32* evals,
                                        uint32_t n_eval) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_refresh) *rq.refresh_out[i] = refresh[i];
    if (i < n_update) *rq.update_out[i] = update[i];
    if (i < n_eval) *rq.eval_out[i] = evals[i];
}

struct DeviceBuffers {
    std::vector<void*> all;
    template <typename T> bool alloc(T*& p, size_t count) {
        void* raw = nullptr;
        if (cudaMalloc(&raw, std::max<size_t>(count, 1) * sizeof(T)) != cudaSuccess) return false;
        all.push_back(raw);
        p = static_cast<T*>(raw);
        return true;
    }
    ~DeviceBuffers() {
        for (void* p : all) cudaFree(p);
    }
};

/* One driver instance: the game slots [params.slotBegin, params.slotEnd) with their own evaluator context,
 * stream and buffers.  Several instances (host threads) overlap one instance's search step with another's
 * evaluation batch on the device. */
int run_instance(const void* net_image, size_t len, int device, const Params& params, Stats& total, uint64_t& batches, std::vector<uint8_t>& records) {
    SpNnue* ctx = nullptr;
    int rc = sp_nnue_create(net_image, len, device, &ctx);
    if (rc != SP_OK) return rc;
    cudaSetDevice(device);
    const uint32_t n = params.localSlots();
    {
        DeviceBuffers mem;
        Slots sl{};
        Requests rq{};
        GpuEvaluator ev{};
        GpuEvaluator* d_ev = nullptr;
        i32 *d_out_refresh = nullptr, *d_out_update = nullptr, *d_out_eval = nullptr;
        sl.n = n;
        sl.record_stride = static_cast<size_t>(params.gamesOfSlot(0)) * ViriGame::maxBytes(params.maxPlies); /* slot 0 plays the most games */
        bool ok = mem.alloc(sl.games, n) && mem.alloc(sl.stats, n) && mem.alloc(sl.played, n) && mem.alloc(sl.alive, n)
               && mem.alloc(sl.records, static_cast<size_t>(n) * sl.record_stride) && mem.alloc(sl.record_len, n)
               && mem.alloc(rq.refresh_slot, n) && mem.alloc(rq.refresh_board, n) && mem.alloc(rq.refresh_out, n) && mem.alloc(rq.update_src, n)
               && mem.alloc(rq.update_dst, n) && mem.alloc(rq.update_board, n) && mem.alloc(rq.update_out, n) && mem.alloc(rq.eval_slot, n)
               && mem.alloc(rq.eval_stm, n) && mem.alloc(rq.eval_out, n) && mem.alloc(rq.counters, 4) && mem.alloc(ev.top, n) && mem.alloc(ev.clean, n)
               && mem.alloc(ev.stale, n) && mem.alloc(d_ev, 1) && mem.alloc(d_out_refresh, n) && mem.alloc(d_out_update, n) && mem.alloc(d_out_eval, n);
        if (!ok) rc = SP_ERR_CUDA;
        if (rc == SP_OK) rc = sp_nnue_slots_reserve(ctx, static_cast<size_t>(n) * kSlotsPerGame);
        cudaStream_t stream = nullptr;
        if (rc == SP_OK && cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) rc = SP_ERR_CUDA;
        uint32_t* h_counters = nullptr;
        if (rc == SP_OK && cudaMallocHost(&h_counters, 4 * sizeof(uint32_t)) != cudaSuccess) rc = SP_ERR_CUDA;
        if (rc == SP_OK) {
            ev.rq = rq;
            cudaMemcpyAsync(d_ev, &ev, sizeof(ev), cudaMemcpyHostToDevice, stream);
            const unsigned grid = (n + 127) / 128;
            cudaMemsetAsync(rq.counters, 0, 4 * sizeof(uint32_t), stream);
            selfplay_init_kernel<<<grid, 128, 0, stream>>>(sl, d_ev, params);
            for (;;) {
                selfplay_step_kernel<<<grid, 128, 0, stream>>>(sl, d_ev, params);
                cudaMemcpyAsync(h_counters, rq.counters, 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost, stream);
                cudaMemsetAsync(rq.counters, 0, 4 * sizeof(uint32_t), stream);
                if (cudaStreamSynchronize(stream) != cudaSuccess) {
                    std::fprintf(stderr, "sp_selfplay_run_gpu: %s\n", cudaGetErrorString(cudaGetLastError()));
                    rc = SP_ERR_CUDA;
                    break;
                }
                const uint32_t n_refresh = h_counters[0], n_update = h_counters[1], n_eval = h_counters[2];
                if (!h_counters[3]) break; /* every slot has played its games */
                rc = sp_nnue_batch_device(
                    ctx, rq.refresh_slot, rq.refresh_board, n_refresh, d_out_refresh, rq.update_src, rq.update_dst, rq.update_board, n_update,
                    d_out_update, rq.eval_slot, rq.eval_stm, n_eval, d_out_eval, stream);
                if (rc != SP_OK) break;
                const uint32_t most = std::max({n_refresh, n_update, n_eval});
                selfplay_scatter_kernel<<<(most + 255) / 256, 256, 0, stream>>>(rq, d_out_refresh, n_refresh, d_out_update, n_update, d_out_eval, n_eval);
                ++batches;
            }
        }
        if (rc == SP_OK) rc = sp_nnue_sync(ctx, stream);
        if (rc == SP_OK) {
            std::vector<Stats> h_stats(n);
            std::vector<uint32_t> h_len(n);
            cudaMemcpy(h_stats.data(), sl.stats, n * sizeof(Stats), cudaMemcpyDeviceToHost);
            cudaMemcpy(h_len.data(), sl.record_len, n * sizeof(uint32_t), cudaMemcpyDeviceToHost);
            std::vector<uint8_t> h_records(static_cast<size_t>(n) * sl.record_stride);
            cudaMemcpy(h_records.data(), sl.records, h_records.size(), cudaMemcpyDeviceToHost);
            for (uint32_t g = 0; g < n; ++g) { /* slot-major, like the host driver */
                total.games += h_stats[g].games, total.positions += h_stats[g].positions, total.nodes += h_stats[g].nodes;
                total.evals += h_stats[g].evals, total.searches += h_stats[g].searches;
                records.insert(records.end(), h_records.begin() + static_cast<ptrdiff_t>(g * sl.record_stride),
                               h_records.begin() + static_cast<ptrdiff_t>(g * sl.record_stride + h_len[g]));
            }
            if (cudaGetLastError() != cudaSuccess) rc = SP_ERR_CUDA;
        }
        if (h_counters) cudaFreeHost(h_counters);
        if (stream) cudaStreamDestroy(stream);
    }
    sp_nnue_destroy(ctx);
    return rc;
}

} // namespace

extern "C" int sp_selfplay_run_gpu(
    const void* net_image, size_t len, int device, const SpSelfplayParams* in, SpSelfplayStats* stats, uint8_t* out, size_t out_capacity,
    size_t* out_len) {
    if (!net_image || !in || !stats || !out_len || (!out && out_capacity)) return SP_ERR_INVALID;
    const Params params = makeParams(*in);
    const uint32_t instances = std::min<uint32_t>(std::max<uint32_t>(1, in->threads), params.concurrency);
    *stats = SpSelfplayStats{};
    *out_len = 0;
    int prev_device = 0;
    cudaGetDevice(&prev_device);
    cudaSetDevice(device);
    /* the games' search runs per thread on its own stack (move lists, board copies: 2 KB measured); the
     * caller's limit is put back afterwards */
    size_t prev_stack = 0;
    cudaDeviceGetLimit(&prev_stack, cudaLimitStackSize);
    cudaDeviceSetLimit(cudaLimitStackSize, std::max<size_t>(prev_stack, 8 * 1024));
    std::vector<Stats> totals(instances);
    std::vector<uint64_t> batches(instances, 0);
    std::vector<std::vector<uint8_t>> records(instances);
    std::vector<int> rcs(instances, SP_OK);
    std::vector<std::thread> pool;
    for (uint32_t t = 0; t < instances; ++t)
        pool.emplace_back([&, t] {
            Params p = params;
            p.slotBegin = static_cast<uint32_t>(uint64_t{p.concurrency} * t / instances);
            p.slotEnd = static_cast<uint32_t>(uint64_t{p.concurrency} * (t + 1) / instances);
            rcs[t] = run_instance(net_image, len, device, p, totals[t], batches[t], records[t]);
        });
    for (auto& th : pool) th.join();
    if (prev_stack) cudaDeviceSetLimit(cudaLimitStackSize, prev_stack);
    cudaSetDevice(prev_device);
    int rc = SP_OK;
    size_t total = 0;
    for (uint32_t t = 0; t < instances; ++t) {
        if (rcs[t] != SP_OK) rc = rcs[t];
        stats->games += totals[t].games, stats->positions += totals[t].positions, stats->nodes += totals[t].nodes;
        stats->evals += totals[t].evals, stats->searches += totals[t].searches, stats->batches += batches[t];
        total += records[t].size();
    }
    if (rc != SP_OK) return rc;
    *out_len = total;
    if (out && total > out_capacity) return SP_ERR_CAPACITY;
    if (out) {
        size_t at = 0;
        for (const auto& r : records) {
            std::copy(r.begin(), r.end(), out + at);
            at += r.size();
        }
    }
    return SP_OK;
}
