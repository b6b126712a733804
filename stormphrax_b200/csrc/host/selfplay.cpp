/* selfplay.cpp -- extern "C" entry point of the batched self-play driver (selfplay.h). */
#include "selfplay.h"

#include <atomic>
#include <string>
#include <thread>

using namespace sp::host;

extern "C" int sp_selfplay_run(
    const void* net_image, size_t len, int device, const SpSelfplayParams* params, SpSelfplayStats* stats, uint8_t* out,
    size_t out_capacity, size_t* out_len) {
    if (!net_image || !params || !stats || !out_len || (!out && out_capacity)) return SP_ERR_INVALID;
    const uint32_t threads = std::max<uint32_t>(1, params->threads);
    /* one evaluator context per host thread: its own stream, scratch and slot store, so the threads' batches
     * overlap on the device and nothing is shared on the host */
    std::vector<SpNnue*> contexts(threads, nullptr);
    int rc = SP_OK;
    for (uint32_t t = 0; t < threads && rc == SP_OK; ++t) rc = sp_nnue_create(net_image, len, device, &contexts[t]);
    std::vector<std::vector<uint8_t>> records(threads);
    std::vector<selfplay::Stats> per_thread(threads);
    std::atomic<int> failed{0};
    if (rc == SP_OK) {
        std::vector<std::thread> pool;
        for (uint32_t t = 0; t < threads; ++t)
            pool.emplace_back([&, t] {
                selfplay::Params p = selfplay::makeParams(*params);
                /* thread t plays a contiguous range of the game slots */
                p.slotBegin = static_cast<uint32_t>(uint64_t{p.concurrency} * t / threads);
                p.slotEnd = static_cast<uint32_t>(uint64_t{p.concurrency} * (t + 1) / threads);
                if (p.slotEnd == p.slotBegin) return;
                selfplay::DeviceEvaluator evaluator{contexts[t], p.localSlots()};
                selfplay::Driver<selfplay::DeviceEvaluator> driver{p, evaluator};
                if (!driver.run(records[t], per_thread[t])) failed.store(1);
            });
        for (auto& th : pool) th.join();
    }
    *stats = SpSelfplayStats{};
    size_t total = 0;
    for (uint32_t t = 0; t < threads; ++t) {
        stats->games += per_thread[t].games, stats->positions += per_thread[t].positions, stats->nodes += per_thread[t].nodes;
        stats->evals += per_thread[t].evals, stats->batches += per_thread[t].batches, stats->searches += per_thread[t].searches;
        total += records[t].size();
    }
    *out_len = total;
    if (rc == SP_OK && failed.load()) rc = SP_ERR_CUDA;
    if (rc == SP_OK && total > out_capacity) rc = out ? SP_ERR_CAPACITY : SP_OK; /* out == NULL: size query only */
    if (rc == SP_OK && out) {
        size_t at = 0;
        for (uint32_t t = 0; t < threads; ++t) {
            std::memcpy(out + at, records[t].data(), records[t].size());
            at += records[t].size();
        }
    }
    for (SpNnue* ctx : contexts) sp_nnue_destroy(ctx);
    return rc;
}

/* ---- the record writer and the score normalisation on their own (parity tests against the reference) */
extern "C" long sp_host_viriformat(
    const SpPackedBoard* start, const SpMove* moves, const int16_t* scores, uint32_t n, int outcome, uint8_t* out, size_t cap) {
    Position pos;
    if (!start || !out || outcome < 0 || outcome > 2 || !Position::fromPacked(*start, pos)) return -1;
    selfplay::ViriGame game;
    game.start(pos);
    for (uint32_t i = 0; i < n; ++i) game.push(Move{moves[i]}, scores[i]);
    std::vector<uint8_t> bytes;
    game.writeAllWithOutcome(bytes, static_cast<selfplay::Outcome>(outcome));
    if (bytes.size() > cap) return -1;
    std::memcpy(out, bytes.data(), bytes.size());
    return static_cast<long>(bytes.size());
}

extern "C" int sp_host_normalize_score(const SpPackedBoard* board, int32_t score, int32_t* material, int32_t* normalized) {
    Position pos;
    if (!board || !material || !normalized || !Position::fromPacked(*board, pos)) return SP_ERR_BAD_BOARD;
    *material = selfplay::classicalMaterial(pos);
    *normalized = selfplay::normalizeScore(score, *material);
    return SP_OK;
}

extern "C" int sp_host_wdl_model(const SpPackedBoard* board, int32_t pov_score, int32_t* win, int32_t* loss) {
    Position pos;
    if (!board || !win || !loss || !Position::fromPacked(*board, pos)) return SP_ERR_BAD_BOARD;
    selfplay::wdlModel(pov_score, selfplay::classicalMaterial(pos), *win, *loss);
    return SP_OK;
}
