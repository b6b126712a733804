/* nnue_state.cpp -- see nnue_state.h. */
#include "nnue_state.h"

#include <cstdio>
#include <fstream>
#include <iterator>

namespace sp::host::eval {

namespace {
SpNnue* g_network = nullptr;
thread_local std::string g_error; /* schedulers on several host threads report through here */
/* the image and device of init(), kept for createContext(); the extra contexts, so that shutdown() leaves none behind */
std::vector<uint8_t> g_image;
int g_device = 0;
std::vector<SpNnue*> g_extra;
std::mutex g_extraMutex;

bool report(const char* what, SpNnue* ctx, int rc) {
    if (rc == SP_OK) return true;
    g_error = std::string{what} + ": " + sp_nnue_last_error(ctx);
    std::fprintf(stderr, "%s\n", g_error.c_str()); /* the reference reports load errors on stderr too (loader.cpp:30-43) */
    return false;
}
} // namespace

bool init(const void* networkImage, size_t len, int device) {
    shutdown();
    SpNnue* ctx = nullptr;
    if (!report("eval::init", nullptr, sp_nnue_create(networkImage, len, device, &ctx))) return false;
    g_network = ctx;
    g_image.assign(static_cast<const uint8_t*>(networkImage), static_cast<const uint8_t*>(networkImage) + len);
    g_device = device;
    return true;
}

SpNnue* createContext() {
    if (!g_network) {
        g_error = "eval::createContext: no network loaded";
        return nullptr;
    }
    SpNnue* ctx = nullptr;
    std::lock_guard<std::mutex> lock{g_extraMutex};
    if (!report("eval::createContext", nullptr, sp_nnue_create(g_image.data(), g_image.size(), g_device, &ctx))) return nullptr;
    g_extra.push_back(ctx);
    return ctx;
}

void destroyContext(SpNnue* ctx) {
    if (!ctx || ctx == g_network) return;
    std::lock_guard<std::mutex> lock{g_extraMutex};
    for (size_t i = 0; i < g_extra.size(); ++i) {
        if (g_extra[i] != ctx) continue;
        g_extra.erase(g_extra.begin() + static_cast<std::ptrdiff_t>(i));
        sp_nnue_destroy(ctx);
        return;
    }
}

bool initFromFile(const std::string& path, int device) {
    std::ifstream in{path, std::ios::binary};
    if (!in) {
        g_error = "failed to open network file " + path;
        std::fprintf(stderr, "%s\n", g_error.c_str());
        return false;
    }
    const std::vector<char> bytes{std::istreambuf_iterator<char>{in}, std::istreambuf_iterator<char>{}};
    return init(bytes.data(), bytes.size(), device);
}

void shutdown() {
    {
        std::lock_guard<std::mutex> lock{g_extraMutex};
        for (SpNnue* ctx : g_extra) sp_nnue_destroy(ctx);
        g_extra.clear();
    }
    sp_nnue_destroy(g_network);
    g_network = nullptr;
    g_image.clear();
    g_image.shrink_to_fit();
}

bool isNetworkLoaded() { return g_network != nullptr; }
SpNnue* getNetwork(int) { return g_network; }
const char* lastError() { return g_error.c_str(); }

/* ------------------------------------------------------------------ EvalBatch */

void EvalBatch::enqueue(uint32_t srcSlot, uint32_t dstSlot, bool rebuild, const SpPackedBoard& board, Color stm, i32* out) {
    std::lock_guard<std::mutex> lock{m_mutex};
    const bool boardStm = stm == ((board.stm_ep & 0x80) ? kBlack : kWhite);
    if (rebuild || srcSlot != kNoUpdate) {
        Group& g = rebuild ? m_refresh : m_update;
        g.src.push_back(srcSlot);
        g.dst.push_back(dstSlot);
        g.boards.push_back(board);
        g.out.push_back(boardStm ? out : nullptr);
        if (boardStm || !out) return;
    }
    if (!out) return;
    m_evalSlots.push_back(dstSlot);
    m_stm.push_back(static_cast<uint8_t>(stm));
    m_out.push_back(out);
}

int EvalBatch::flush() {
    std::lock_guard<std::mutex> lock{m_mutex};
    m_refresh.results.resize(m_refresh.dst.size());
    m_update.results.resize(m_update.dst.size());
    m_results.resize(m_evalSlots.size());
    const int rc = sp_nnue_batch(
        m_network, m_refresh.dst.data(), m_refresh.boards.data(), m_refresh.dst.size(), m_refresh.results.data(), m_update.src.data(),
        m_update.dst.data(), m_update.boards.data(), m_update.dst.size(), m_update.results.data(), m_evalSlots.data(), m_stm.data(),
        m_evalSlots.size(), m_results.data());
    if (rc == SP_OK) {
        for (Group* g : {&m_refresh, &m_update})
            for (size_t i = 0; i < g->out.size(); ++i)
                if (g->out[i]) *g->out[i] = g->results[i];
        for (size_t i = 0; i < m_out.size(); ++i) *m_out[i] = m_results[i];
    }
    m_refresh.clear(), m_update.clear();
    m_evalSlots.clear(), m_stm.clear(), m_out.clear();
    if (rc != SP_OK) report("EvalBatch::flush", m_network, rc);
    return rc;
}

/* ------------------------------------------------------------------ NnueState */

void NnueState::setNetwork(SpNnue* network, uint32_t slotBase, uint32_t depth) {
    m_network = network;
    m_slotBase = slotBase;
    m_top = 0;
    m_clean.assign(depth, 0);
    m_stale.assign(depth, 0);
    if (network) report("NnueState::setNetwork", network, sp_nnue_slots_reserve(network, size_t{slotBase} + depth));
}

/* reset, nnue_state.cpp:539-560: rebuild both perspectives at stack level 0 */
void NnueState::resetPacked(const SpPackedBoard& board) {
    m_top = 0;
    std::fill(m_clean.begin(), m_clean.end(), 0);
    std::fill(m_stale.begin(), m_stale.end(), 0);
    const uint32_t s = slot(0);
    m_clean[0] = report("NnueState::reset", m_network, sp_nnue_refresh(m_network, &s, &board, 1));
}

/* push, nnue_state.cpp:562-570 */
BoardObserver NnueState::push() {
    ++m_top;
    m_clean[m_top] = 0;
    m_stale[m_top] = 0;
    m_ctx = {};
    return BoardObserver{m_ctx};
}

/* pop, nnue_state.cpp:593-596 */
void NnueState::pop() { --m_top; }

int NnueState::cleanAncestor() const {
    for (int level = static_cast<int>(m_top); level >= 0; --level)
        if (m_clean[static_cast<size_t>(level)]) return level;
    return -1;
}

/* applyImmediately, nnue_state.cpp:572-591: the current level advances in place */
void NnueState::applyPacked(const SpPackedBoard& board) {
    const int from = cleanAncestor();
    const uint32_t dst = slot(m_top);
    int rc;
    if (from < 0) {
        rc = sp_nnue_refresh(m_network, &dst, &board, 1);
    } else {
        const uint32_t src = slot(static_cast<uint32_t>(from));
        rc = sp_nnue_update(m_network, &src, &dst, &board, 1);
    }
    m_clean[m_top] = report("NnueState::applyImmediately", m_network, rc);
    m_stale[m_top] = 0;
}

/* evaluate, nnue_state.cpp:598-610 + ensureUpToDate :636-697.  The usual case -- the level is dirty and the wanted side
 * is the board's side to move -- is ONE submission: update (or rebuild) and evaluate together. */
i32 NnueState::evaluatePacked(const SpPackedBoard& board, Color stm) {
    const uint32_t dst = slot(m_top);
    i32 out = 0;
    const bool dirty = !m_clean[m_top] || m_stale[m_top];
    const bool boardStm = stm == ((board.stm_ep & 0x80) ? kBlack : kWhite);
    if (dirty && boardStm) {
        const int from = cleanAncestor();
        int rc;
        if (from < 0) {
            rc = sp_nnue_batch(m_network, &dst, &board, 1, &out, nullptr, nullptr, nullptr, 0, nullptr, nullptr, nullptr, 0, nullptr);
        } else {
            const uint32_t src = slot(static_cast<uint32_t>(from));
            rc = sp_nnue_update_eval(m_network, &src, &dst, &board, 1, &out);
        }
        m_clean[m_top] = report("NnueState::evaluate", m_network, rc);
        m_stale[m_top] = 0;
        return out;
    }
    if (dirty) applyPacked(board);
    const uint8_t side = static_cast<uint8_t>(stm);
    report("NnueState::evaluate", m_network, sp_nnue_eval_slots(m_network, &dst, &side, 1, &out));
    return out;
}

void NnueState::evaluateAsyncPacked(EvalBatch& batch, const SpPackedBoard& board, Color stm, i32* out) {
    const uint32_t dst = slot(m_top);
    if (m_clean[m_top] && !m_stale[m_top]) {
        batch.enqueue(EvalBatch::kNoUpdate, dst, false, board, stm, out); /* already up to date: evaluate only */
        return;
    }
    const int from = cleanAncestor(); /* may be this very level when it is stale: the slot is updated in place */
    batch.enqueue(from < 0 ? dst : slot(static_cast<uint32_t>(from)), dst, from < 0, board, stm, out);
    m_clean[m_top] = 1; /* valid once the batch is flushed, which must happen before the next dependent call */
    m_stale[m_top] = 0;
}

/* evaluateOnce, nnue_state.cpp:612-634: stack-free, from scratch */
i32 NnueState::evaluateOncePacked(const SpPackedBoard& board, Color stm) {
    SpPackedBoard b = board;
    b.stm_ep = static_cast<uint8_t>((b.stm_ep & 0x7F) | (stm == kBlack ? 0x80 : 0));
    i32 out = 0;
    report("NnueState::evaluateOnce", g_network, sp_nnue_eval_full(g_network, &b, 1, &out));
    return out;
}

} // namespace sp::host::eval
