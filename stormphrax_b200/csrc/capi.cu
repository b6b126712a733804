/*
 * capi.cu -- the extern "C" boundary of libsp_nnue.so (include/sp_nnue.h): network upload,
 * device buffer management, batching, error reporting.  No evaluation arithmetic lives here
 * and there is NO CPU fallback: without a CUDA device every entry point fails.
 */
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/sp_nnue.h"
#include "kernels.cuh"
#include "host/net_loader.h"

using namespace sp;
using namespace sp::gpu;

struct SpNnue {
    int device = -1;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    bool owns_stream = true;
    DeviceNet net{};
    uint8_t* d_net_blob = nullptr;
    FeatureTables* d_tables = nullptr;
    DeviceStatus* d_status = nullptr;
    DeviceStatus* h_status = nullptr; /* pinned */

    /* scratch for one chunk of positions, double-buffered: the head of chunk i runs on `aux` while the
     * feature transformer of chunk i + 1 runs on the caller's stream */
    size_t chunk = 262144; /* positions per launch of the full refresh: 65,536 measured 4 % slower (launch gaps and kernel tails per chunk) */
    uint32_t games_chunk = 7104; /* playout walker: games per launch (3 per resident warp) */
    uint8_t* d_act2[2] = {nullptr, nullptr};
    uint8_t* d_bucket2[2] = {nullptr, nullptr};
    uint8_t* d_act = nullptr;    /* = d_act2[0] */
    uint8_t* d_bucket = nullptr; /* = d_bucket2[0] */
    cudaStream_t aux = nullptr;
    cudaStream_t h2d = nullptr, d2h = nullptr; /* host-pointer entry points: copies overlap the kernels chunk by chunk */
    cudaEvent_t ev_ft[2] = {nullptr, nullptr}, ev_head[2] = {nullptr, nullptr}, ev_join = nullptr, ev_start = nullptr;
    std::vector<cudaEvent_t> ev_chunk; /* one "inputs of chunk i have landed" event per chunk */
    bool overlap = true;
    HeadSort head_sort{};        /* scratch of the dense head's bucket grouping, grown on demand */
    bool plan_rebuilds = true;   /* playout walker: rebuilds computed ahead by their own kernel (SP_NNUE_PLAN_REBUILDS=0: inline) */
    RebuildPlan plan{};          /* scratch of that scheme, sized for the largest stream seen */
    size_t plan_boards = 0;
    bool split = false;          /* full refresh as extract + accumulate kernels instead of the fused ft_full kernel (SP_NNUE_SPLIT=1) */
    bool group = true;           /* full refresh on the tensor cores, 16 positions per group (ft_group_kernel); SP_NNUE_FT=warp: ft_full_kernel */
    uint32_t* d_group_overflow[2] = {nullptr, nullptr}; /* per scratch buffer: groups ft_group_kernel hands to ft_full_kernel */
    /* Feedback for input whose neighbours share no rows (shuffled positions): every group launch copies its overflow count home
     * (pinned, no wait); when a launch that has completed handed more than half of its groups to ft_full_kernel, the next launches
     * go to that kernel directly (81 instead of 66 Mpos/s on shuffled positions: no detour) and the group kernel is tried again after
     * 3, 7, 15, 15, ... launches (back to 3 as soon as a launch mostly fits). */
    uint32_t* h_group_overflow = nullptr; /* [2]: count written by the copy, per scratch buffer */
    uint32_t group_launch_groups[2] = {0, 0}; /* groups of the launch whose count is in flight / has arrived (0: none, or already looked at) */
    cudaEvent_t ev_group_count[2] = {nullptr, nullptr}; /* recorded behind the copy: the count may be read once it has completed */
    int group_backoff = 0, group_backoff_len = 3; /* launches left on the warp kernel; length of the next back-off (3, 7, 15, 15, ...) */
    void* d_lists[2] = {nullptr, nullptr};
    /* whole-stream scratch of the playout walker (one activation row per board) */
    uint8_t* d_act_big = nullptr;
    uint8_t* d_bucket_big = nullptr;
    size_t act_cap = 0;
    /* staging for the host-pointer entry points */
    SpPackedBoard* d_boards = nullptr;
    int32_t* d_out = nullptr;
    uint32_t* d_ids = nullptr; /* [3][cap]: src slots, dst slots, game starts */
    uint8_t* d_stm = nullptr;
    size_t cap = 0;

    /* search-sized rounds (small_batch_kernel): one pinned host block and its device twin, inputs | error word | results */
    bool small = true;           /* SP_NNUE_SMALL=0: always take the general kernels */
    size_t small_mapped = 64;    /* SP_NNUE_SMALL_MAPPED: up to this many items (<= 64) the kernel works on the host block directly (zero-copy) */
    uint8_t* h_small = nullptr;
    uint8_t* h_small_dev = nullptr; /* device address of h_small */
    uint8_t* d_small = nullptr;

    SlotStore slots{};
    uint64_t counters[SP_NUM_COUNTERS] = {};
    std::string error;

    /* optional per-kernel timing (sp_nnue_profile): event pairs around launches, by kernel class */
    bool profiling = false;
    struct Span { cudaEvent_t a, b; int kind; };
    std::vector<Span> spans;
    std::vector<cudaEvent_t> event_pool;
    double prof_ms[SP_NUM_KERNEL_CLASSES] = {};
    uint64_t prof_launches[SP_NUM_KERNEL_CLASSES] = {};
};

namespace {

thread_local std::string g_create_error; /* create() may fail on several threads at once (sp_selfplay_run_gpu) */

int fail(SpNnue* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    (ctx ? ctx->error : g_create_error) = buf;
    return code;
}

#define SP_CUDA(ctx, call)                                                                              \
    do {                                                                                                \
        const cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess) return fail(ctx, SP_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
    } while (0)

/* Header checks in the order of `validate`, src/eval/nnue.cpp:85-185 (header layout header.h:38-52). */
const char* validate_header(const uint8_t* h) {
    static thread_local char msg[160];
    if (std::memcmp(h, "CBNF", 4) != 0) return "invalid magic bytes in network header";
    const unsigned version = h[4] | h[5] << 8, flags = h[6] | h[7] << 8;
    const unsigned arch = h[9], activation = h[10], hidden = h[11] | h[12] << 8, in_b = h[13], out_b = h[14];
    auto fmt = [&](const char* f, unsigned a, unsigned b) { snprintf(msg, sizeof(msg), f, a, b); return msg; };
    if (version != 1) return fmt("unsupported network format version %u (expected: %u)", version, 1);
    if (arch != 5) return fmt("wrong network architecture %u (expected: %u)", arch, 5);
    if (!(flags & 0x0002)) return "unmirrored network, expected horizontally mirrored";
    if (!(flags & 0x0004)) return "network does not have merged king planes, expected merged";
    if (!(flags & 0x0008)) return "network L1 does not require pairwise multiplication, expected paired";
    if (activation != 0) return fmt("wrong l1 activation function %u (expected: %u = crelu)", activation, 0);
    if (hidden != SP_L1_SIZE) return fmt("wrong number of l1 neurons %u (expected: %u)", hidden, SP_L1_SIZE);
    if (!(in_b & 0x80)) return "network does not have the expected threat inputs";
    if ((in_b & 0x7F) != SP_INPUT_BUCKETS) return fmt("wrong number of input buckets %u (expected: %u)", in_b & 0x7F, SP_INPUT_BUCKETS);
    if (out_b != SP_OUTPUT_BUCKETS) return fmt("wrong number of output buckets %u (expected: %u)", out_b, SP_OUTPUT_BUCKETS);
    (void)flags; /* kZstdCompressed (0x0001) is honoured by the loader: host/net_loader.cpp */
    return nullptr;
}

constexpr size_t align256(size_t x) { return (x + 255) & ~size_t{255}; }

/* Device layout of the network. Offsets into one allocation. */
struct NetLayout {
    size_t psq, psq_planes, thr, l1_w, l1_b, l2_w, l2_limbs, l2_frags, l2_b, l3_w, l3_b, total;
    NetLayout() {
        size_t o = 0;
        psq = o;  o = align256(o + size_t{kPsqRows} * SP_L1_SIZE * 2);
        psq_planes = o; o = align256(o + size_t{kPsqRows} * SP_L1_SIZE * 2);
        thr = o;  o = align256(o + size_t{kThrRows} * SP_L1_SIZE);
        l1_w = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * SP_L1_SIZE * SP_L2_SIZE);
        l1_b = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * SP_L2_SIZE * 4);
        l2_w = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * 2 * SP_L2_SIZE * SP_L3_SIZE * 4);
        l2_limbs = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * 2 * SP_L2_SIZE * SP_L3_SIZE * 4);
        l2_frags = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * 2 * SP_L2_SIZE * SP_L3_SIZE * 4);
        l2_b = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * SP_L3_SIZE * 4);
        l3_w = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * SP_L3_SIZE * 4);
        l3_b = o; o = align256(o + size_t{SP_OUTPUT_BUCKETS} * 4);
        total = o;
    }
};

/* logical file image -> device image (host side) */
void build_device_image(const uint8_t* payload, const NetLayout& L, uint8_t* img) {
    const int16_t* psq_w = reinterpret_cast<const int16_t*>(payload);
    const uint8_t* thr_w = payload + size_t{SP_PSQ_FEATURES} * SP_L1_SIZE * 2;
    const uint8_t* rest = thr_w + size_t{SP_THREAT_FEATURES} * SP_L1_SIZE;
    const int16_t* ft_b = reinterpret_cast<const int16_t*>(rest);
    rest += SP_L1_SIZE * 2;

    int16_t* psq = reinterpret_cast<int16_t*>(img + L.psq);
    auto permute_row = [&](const int16_t* src, int16_t* dst) {
        for (int k = 0; k < 4; ++k)
            for (int lane = 0; lane < 32; ++lane)
                for (int e = 0; e < 8; ++e) dst[(k * 32 + lane) * 8 + e] = src[lane_order_element(k, lane, e)];
    };
    for (int r = 0; r < SP_PSQ_FEATURES; ++r) permute_row(psq_w + size_t(r) * SP_L1_SIZE, psq + size_t(r) * SP_L1_SIZE);
    permute_row(ft_b, psq + size_t{kPsqBiasRow} * SP_L1_SIZE);
    std::memset(psq + size_t{kPsqZeroRow} * SP_L1_SIZE, 0, SP_L1_SIZE * 2);

    /* byte planes of the same rows (bias row and zero row included) in natural column order */
    {
        uint8_t* planes = img + L.psq_planes;
        auto split_row = [&](const int16_t* src, uint8_t* dst) {
            for (int i = 0; i < SP_L1_SIZE; ++i) {
                const uint16_t v = static_cast<uint16_t>(src[i]);
                dst[i] = static_cast<uint8_t>(v & 0xFF);
                dst[SP_L1_SIZE + i] = static_cast<uint8_t>(v >> 8);
            }
        };
        for (int r = 0; r < SP_PSQ_FEATURES; ++r) split_row(psq_w + size_t(r) * SP_L1_SIZE, planes + size_t(r) * 2 * SP_L1_SIZE);
        split_row(ft_b, planes + size_t{kPsqBiasRow} * 2 * SP_L1_SIZE);
        std::memset(planes + size_t{kPsqZeroRow} * 2 * SP_L1_SIZE, 0, 2 * SP_L1_SIZE);
    }

    uint8_t* thr = img + L.thr;
    const size_t thr_bytes = size_t{SP_THREAT_FEATURES} * SP_L1_SIZE;
    for (size_t i = 0; i < thr_bytes; ++i) thr[i] = static_cast<uint8_t>(thr_w[i] ^ 0x80); /* int8 + 128 */
    std::memset(thr + thr_bytes, 0x80, SP_L1_SIZE);

    auto take = [&](size_t off, size_t bytes) {
        std::memcpy(img + off, rest, bytes);
        rest += bytes;
    };
    take(L.l1_w, size_t{SP_OUTPUT_BUCKETS} * SP_L1_SIZE * SP_L2_SIZE);
    take(L.l1_b, size_t{SP_OUTPUT_BUCKETS} * SP_L2_SIZE * 4);
    take(L.l2_w, size_t{SP_OUTPUT_BUCKETS} * 2 * SP_L2_SIZE * SP_L3_SIZE * 4);
    /* byte limbs of the int32 L2 weights, laid out as tensor-core B fragments (kernels.cuh: l2_limb_index) */
    {
        const uint32_t* w2 = reinterpret_cast<const uint32_t*>(img + L.l2_w);
        uint8_t* limbs = img + L.l2_limbs;
        for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b)
            for (int k = 0; k < 2 * SP_L2_SIZE; ++k)
                for (int o = 0; o < SP_L3_SIZE; ++o) {
                    const uint32_t w = w2[(static_cast<size_t>(b) * 2 * SP_L2_SIZE + k) * SP_L3_SIZE + o];
                    for (int limb = 0; limb < 4; ++limb)
                        limbs[(static_cast<size_t>(b) * 4096 + l2_limb_index(limb, k >> 2, o)) * 4 + (k & 3)] = static_cast<uint8_t>(w >> (8 * limb));
                }
        uint8_t* frags = img + L.l2_frags;
        for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b)
            for (int limb = 0; limb < 4; ++limb)
                for (int nt = 0; nt < 8; ++nt)
                    for (int ks = 0; ks < 2; ++ks)
                        for (int lane = 0; lane < 32; ++lane)
                            for (int reg = 0; reg < 2; ++reg)
                                for (int m = 0; m < 4; ++m) {
                                    const int g = lane >> 2, t = lane & 3;
                                    const int k = 32 * ks + 8 * t + 4 * reg + m, o = 8 * nt + g;
                                    const uint32_t w = w2[(static_cast<size_t>(b) * 2 * SP_L2_SIZE + k) * SP_L3_SIZE + o];
                                    frags[(static_cast<size_t>(b) * 4096 + l2_fragment_index(limb, nt, ks, lane, reg)) * 4 + m] = static_cast<uint8_t>(w >> (8 * limb));
                                }
    }
    take(L.l2_b, size_t{SP_OUTPUT_BUCKETS} * SP_L3_SIZE * 4);
    take(L.l3_w, size_t{SP_OUTPUT_BUCKETS} * SP_L3_SIZE * 4);
    take(L.l3_b, size_t{SP_OUTPUT_BUCKETS} * 4);
}

int ensure_staging(SpNnue* ctx, size_t n) {
    if (n <= ctx->cap) return SP_OK;
    const size_t cap = std::max(n, ctx->cap * 2);
    cudaFree(ctx->d_boards);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_ids);
    cudaFree(ctx->d_stm);
    ctx->d_boards = nullptr, ctx->d_out = nullptr, ctx->d_ids = nullptr, ctx->d_stm = nullptr;
    ctx->cap = 0;
    SP_CUDA(ctx, cudaMalloc(&ctx->d_boards, cap * sizeof(SpPackedBoard)));
    SP_CUDA(ctx, cudaMalloc(&ctx->d_out, cap * sizeof(int32_t)));
    SP_CUDA(ctx, cudaMalloc(&ctx->d_ids, 3 * (cap + 1) * sizeof(uint32_t)));
    SP_CUDA(ctx, cudaMalloc(&ctx->d_stm, cap));
    ctx->cap = cap;
    return SP_OK;
}

/* Wait for `stream`, then translate the device status word. */
int finish(SpNnue* ctx, cudaStream_t stream) {
    SP_CUDA(ctx, cudaMemcpyAsync(ctx->h_status, ctx->d_status, sizeof(int), cudaMemcpyDeviceToHost, stream));
    SP_CUDA(ctx, cudaStreamSynchronize(stream));
    const int err = ctx->h_status->error;
    if (!err) return SP_OK;
    SP_CUDA(ctx, cudaMemsetAsync(ctx->d_status, 0, sizeof(int), stream));
    if (err & kErrBadSlot) return fail(ctx, SP_ERR_INVALID, "slot index out of range");
    if (err & kErrBadBoard) return fail(ctx, SP_ERR_BAD_BOARD, "malformed position record (affected outputs are INT32_MIN)");
    return fail(ctx, SP_ERR_CAPACITY, "threat feature list exceeded %d entries", SP_MAX_THREAT_INDICES);
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

cudaStream_t pick(SpNnue* ctx, void* stream) { return stream ? static_cast<cudaStream_t>(stream) : ctx->stream; }

/* The dense head groups a launch's positions by bucket in this scratch.  Growing it waits for the
 * device, which only happens when a call is larger than any before. */
int ensure_head_sort(SpNnue* ctx, size_t n) {
    const size_t need = n + kHeadGroupPad * SP_OUTPUT_BUCKETS;
    if (need <= ctx->head_sort.capacity) return SP_OK;
    SP_CUDA(ctx, cudaDeviceSynchronize());
    cudaFree(ctx->head_sort.order);
    ctx->head_sort.order = nullptr, ctx->head_sort.capacity = 0;
    if (!ctx->head_sort.counters) SP_CUDA(ctx, cudaMalloc(&ctx->head_sort.counters, kHeadSortCounters * sizeof(uint32_t)));
    SP_CUDA(ctx, cudaMalloc(&ctx->head_sort.order, need * sizeof(uint32_t)));
    ctx->head_sort.capacity = need;
    return SP_OK;
}

cudaEvent_t take_event(SpNnue* ctx) {
    if (!ctx->event_pool.empty()) {
        cudaEvent_t e = ctx->event_pool.back();
        ctx->event_pool.pop_back();
        return e;
    }
    cudaEvent_t e = nullptr;
    cudaEventCreate(&e);
    return e;
}

/* RAII: brackets one kernel launch with events when profiling is on */
/* Full refresh of one chunk: the group kernel, unless recent launches mostly overflowed (see SpNnue::group_backoff).  The counts are
 * read without waiting: a value that has not arrived yet only delays the decision. */
bool group_kernel_wanted(SpNnue* ctx, int buf) {
    if (ctx->group_backoff > 0) {
        --ctx->group_backoff;
        return false;
    }
    bool wanted = true;
    for (int b = 0; b < 2; ++b) {
        const uint32_t groups = ctx->group_launch_groups[b];
        if (!groups || cudaEventQuery(ctx->ev_group_count[b]) != cudaSuccess) continue; /* nothing pending, or not there yet */
        ctx->group_launch_groups[b] = 0; /* looked at */
        const uint32_t overflowed = *static_cast<volatile uint32_t*>(ctx->h_group_overflow + b);
        if (groups >= 64 && overflowed * 2 > groups) {
            ctx->group_backoff = ctx->group_backoff_len;
            ctx->group_backoff_len = std::min(2 * ctx->group_backoff_len + 1, 15);
            wanted = false;
        } else {
            ctx->group_backoff_len = 3;
        }
    }
    /* a launch on this buffer whose count has not been looked at yet would be overwritten: look at it first (rare: the host is
     * two launches ahead of the device) */
    (void)buf;
    return wanted;
}

struct Timed {
    SpNnue* ctx;
    cudaStream_t stream;
    SpNnue::Span span{};
    Timed(SpNnue* c, cudaStream_t s, int kind) : ctx{c}, stream{s} {
        if (!ctx->profiling) return;
        span = {take_event(ctx), take_event(ctx), kind};
        cudaEventRecord(span.a, stream);
    }
    ~Timed() {
        if (!ctx->profiling) return;
        cudaEventRecord(span.b, stream);
        ctx->spans.push_back(span);
    }
};

/* boards (device) -> out (device), in chunks that keep the activation scratch L2-sized.  The dense
 * head of chunk i (latency-bound, low occupancy) runs on the auxiliary stream underneath the feature
 * transformer of chunk i + 1; everything is ordered with events, the host never waits. */
/* Host side of a chunked call: where chunk i's boards come from and where its results go. */
struct HostIo {
    const SpPackedBoard* boards;
    int32_t* out;
};

int chunk_event(SpNnue* ctx, size_t i, cudaEvent_t* ev) {
    while (ctx->ev_chunk.size() <= i) {
        cudaEvent_t e = nullptr;
        SP_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->ev_chunk.push_back(e);
    }
    *ev = ctx->ev_chunk[i];
    return SP_OK;
}

int eval_full_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, size_t n, int32_t* d_out, cudaStream_t stream, const HostIo* io = nullptr) {
    if (const int rc = ensure_head_sort(ctx, std::min(n, ctx->chunk))) return rc;
    const bool overlap = ctx->overlap && n > ctx->chunk;
    if (io) {
        /* all uploads are queued up front on their own stream; chunk i's kernels wait for upload i only */
        SP_CUDA(ctx, cudaEventRecord(ctx->ev_start, stream));
        SP_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ctx->ev_start, 0));
        SP_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->ev_start, 0));
        size_t i = 0;
        for (size_t off = 0; off < n; off += ctx->chunk, ++i) {
            const size_t m = std::min(ctx->chunk, n - off);
            cudaEvent_t ev;
            if (const int rc = chunk_event(ctx, i, &ev)) return rc;
            SP_CUDA(ctx, cudaMemcpyAsync(const_cast<SpPackedBoard*>(d_boards) + off, io->boards + off, m * sizeof(SpPackedBoard), cudaMemcpyHostToDevice, ctx->h2d));
            SP_CUDA(ctx, cudaEventRecord(ev, ctx->h2d));
        }
    }
    size_t i = 0;
    for (size_t off = 0; off < n; off += ctx->chunk, ++i) {
        const size_t m = std::min(ctx->chunk, n - off);
        const int buf = overlap ? static_cast<int>(i & 1) : 0;
        if (io) SP_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_chunk[i], 0));
        if (overlap && i >= 2) SP_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_head[buf], 0)); /* scratch is free again */
        if (ctx->split) {
            {
                Timed timed{ctx, stream, SP_KERNEL_EXTRACT};
                launch_extract(ctx->net, d_boards + off, m, ctx->d_lists[buf], ctx->d_status, ctx->sm_count, stream);
            }
            {
                Timed timed{ctx, stream, SP_KERNEL_ACCUMULATE};
                launch_accumulate(ctx->net, ctx->d_lists[buf], m, ctx->d_act2[buf], ctx->d_bucket2[buf], ctx->sm_count, stream);
            }
            ctx->counters[SP_CTR_LAUNCHES] += 1;
        } else if (ctx->group && group_kernel_wanted(ctx, buf)) {
            Timed timed{ctx, stream, SP_KERNEL_FT_FULL};
            SP_CUDA(ctx, launch_ft_group(ctx->net, d_boards + off, m, ctx->d_act2[buf], ctx->d_bucket2[buf], ctx->d_status, ctx->d_group_overflow[buf], ctx->sm_count, stream));
            ctx->group_launch_groups[buf] = static_cast<uint32_t>((m + 15) / 16);
            SP_CUDA(ctx, cudaMemcpyAsync(ctx->h_group_overflow + buf, ctx->d_group_overflow[buf], sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            SP_CUDA(ctx, cudaEventRecord(ctx->ev_group_count[buf], stream));
            ctx->counters[SP_CTR_LAUNCHES] += 1;
        } else {
            Timed timed{ctx, stream, SP_KERNEL_FT_FULL};
            launch_ft_full(ctx->net, d_boards + off, m, ctx->d_act2[buf], ctx->d_bucket2[buf], ctx->d_status, ctx->sm_count, stream);
        }
        cudaStream_t hs = stream;
        if (overlap) {
            SP_CUDA(ctx, cudaEventRecord(ctx->ev_ft[buf], stream));
            SP_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_ft[buf], 0));
            hs = ctx->aux;
        }
        {
            Timed timed{ctx, hs, SP_KERNEL_HEAD};
            SP_CUDA(ctx, launch_head(ctx->net, ctx->d_act2[buf], ctx->d_bucket2[buf], m, d_out + off, nullptr, ctx->head_sort, ctx->sm_count, hs));
            ctx->counters[SP_CTR_LAUNCHES] += head_kernel_launches(m, ctx->head_sort);
        }
        if (overlap) SP_CUDA(ctx, cudaEventRecord(ctx->ev_head[buf], ctx->aux));
        if (io) { /* results of this chunk go home while the next chunk computes */
            if (!overlap) SP_CUDA(ctx, cudaEventRecord(ctx->ev_head[buf], hs));
            SP_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->ev_head[buf], 0));
            SP_CUDA(ctx, cudaMemcpyAsync(io->out + off, d_out + off, m * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->d2h));
        }
        ctx->counters[SP_CTR_LAUNCHES] += 1;
    }
    if (overlap) { /* join: later work on `stream` sees every result */
        SP_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->aux));
        SP_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_join, 0));
    }
    if (io) {
        SP_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->d2h));
        SP_CUDA(ctx, cudaStreamWaitEvent(stream, ctx->ev_join, 0));
    }
    ctx->counters[SP_CTR_EVALS] += n;
    ctx->counters[SP_CTR_FULL_REFRESH] += 2 * n;
    SP_CUDA(ctx, cudaGetLastError());
    return SP_OK;
}

} // namespace

extern "C" {

void sp_nnue_destroy(SpNnue* ctx);

static int create_impl(const void* net_image, size_t len, int device, SpNnue** out) {
    if (!net_image || len < SP_NET_HEADER_BYTES) return fail(nullptr, SP_ERR_BAD_NETWORK, "missing network header");
    const uint8_t* bytes = static_cast<const uint8_t*>(net_image);
    if (const char* why = validate_header(bytes)) return fail(nullptr, SP_ERR_BAD_NETWORK, "%s", why);
    /* raw arrays, or one zstd frame of them (eval::init, src/eval/nnue.cpp:215-263) */
    std::vector<uint8_t> inflated;
    std::string why_not;
    const uint8_t* payload = sp::host::network_payload(bytes, len, SP_NET_PAYLOAD_BYTES, inflated, why_not);
    if (!payload) return fail(nullptr, SP_ERR_BAD_NETWORK, "%s", why_not.c_str());

    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0) {
        cudaGetLastError();
        return fail(nullptr, SP_ERR_NO_DEVICE, "no CUDA device available (this library has no CPU path)");
    }
    if (device < 0 || device >= count) return fail(nullptr, SP_ERR_NO_DEVICE, "device %d out of range (%d present)", device, count);

    /* the caller destroys *out if anything below fails, releasing whatever was allocated so far */
    SpNnue* ctx = new (std::nothrow) SpNnue;
    if (!ctx) return fail(nullptr, SP_ERR_INVALID, "out of host memory");
    *out = ctx;
    ctx->device = device;
    DeviceGuard guard{device};
    cudaDeviceProp prop{};
    SP_CUDA(nullptr, cudaGetDeviceProperties(&prop, device));
    /* arch-specific ("a") targets have no forward compatibility: only compute capability 10.0 can run this image */
    if (prop.major != 10 || prop.minor != 0)
        return fail(nullptr, SP_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
    ctx->sm_count = prop.multiProcessorCount;
    if (const char* env = std::getenv("SP_NNUE_CHUNK")) {
        const long v = std::atol(env);
        if (v >= 16) ctx->chunk = static_cast<size_t>(v);
    }
    SP_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));

    const NetLayout L;
    std::vector<uint8_t> img(L.total);
    build_device_image(payload, L, img.data());
    /* buckets whose L2 weights all fit int16: the streaming head then contracts two weight limbs instead of four */
    uint32_t l2_narrow_mask = 0;
    {
        const int32_t* w2 = reinterpret_cast<const int32_t*>(img.data() + L.l2_w);
        for (int b = 0; b < SP_OUTPUT_BUCKETS; ++b) {
            bool narrow = true;
            for (int i = 0; i < 2 * SP_L2_SIZE * SP_L3_SIZE; ++i) narrow = narrow && w2[b * 2 * SP_L2_SIZE * SP_L3_SIZE + i] >= -32768 && w2[b * 2 * SP_L2_SIZE * SP_L3_SIZE + i] <= 32767;
            if (narrow) l2_narrow_mask |= 1u << b;
        }
        if (const char* env = std::getenv("SP_NNUE_L2_NARROW")) /* 0: always take the general path (tests, A/B) */
            if (std::atoi(env) == 0) l2_narrow_mask = 0;
    }
    SP_CUDA(nullptr, cudaMalloc(&ctx->d_net_blob, L.total));
    SP_CUDA(nullptr, cudaMemcpy(ctx->d_net_blob, img.data(), L.total, cudaMemcpyHostToDevice));
    FeatureTables tables;
    build_feature_tables(tables);
    SP_CUDA(nullptr, cudaMalloc(&ctx->d_tables, sizeof(FeatureTables)));
    SP_CUDA(nullptr, cudaMemcpy(ctx->d_tables, &tables, sizeof(FeatureTables), cudaMemcpyHostToDevice));
    SP_CUDA(nullptr, cudaMalloc(&ctx->d_status, sizeof(DeviceStatus)));
    SP_CUDA(nullptr, cudaMemset(ctx->d_status, 0, sizeof(DeviceStatus)));
    SP_CUDA(nullptr, cudaMallocHost(&ctx->h_status, sizeof(DeviceStatus)));
    if (const char* env = std::getenv("SP_NNUE_GAMES_CHUNK")) {
        const long v = std::atol(env);
        if (v >= 1) ctx->games_chunk = static_cast<uint32_t>(v);
    }
    if (const char* env = std::getenv("SP_NNUE_OVERLAP")) ctx->overlap = std::atoi(env) != 0;
    if (const char* env = std::getenv("SP_NNUE_SPLIT")) ctx->split = std::atoi(env) != 0;
    if (const char* env = std::getenv("SP_NNUE_SMALL")) ctx->small = std::atoi(env) != 0;
    ctx->head_sort.variant = head_variant_from_env();
    ctx->head_sort.direct_max = head_direct_max_from_env();
    if (const char* env = std::getenv("SP_NNUE_SMALL_MAPPED")) ctx->small_mapped = std::min<size_t>(64, std::strtoul(env, nullptr, 10));
    if (const char* env = std::getenv("SP_NNUE_FT")) ctx->group = std::strcmp(env, "warp") != 0; /* warp: one warp per position (ft_full_kernel) */
    if (const char* env = std::getenv("SP_NNUE_PLAN_REBUILDS")) ctx->plan_rebuilds = std::atoi(env) != 0;
    for (int b = 0; b < 2; ++b) {
        SP_CUDA(nullptr, cudaMalloc(&ctx->d_act2[b], ctx->chunk * SP_L1_SIZE));
        SP_CUDA(nullptr, cudaMalloc(&ctx->d_bucket2[b], ctx->chunk));
        if (ctx->split) SP_CUDA(nullptr, cudaMalloc(&ctx->d_lists[b], row_list_bytes(ctx->chunk)));
        SP_CUDA(nullptr, cudaMalloc(&ctx->d_group_overflow[b], ft_group_scratch_words(ctx->chunk) * sizeof(uint32_t)));
        if (!ctx->h_group_overflow) {
            SP_CUDA(nullptr, cudaMallocHost(&ctx->h_group_overflow, 2 * sizeof(uint32_t)));
            ctx->h_group_overflow[0] = ctx->h_group_overflow[1] = 0;
        }
        SP_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->ev_group_count[b], cudaEventDisableTiming));
        SP_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->ev_ft[b], cudaEventDisableTiming));
        SP_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->ev_head[b], cudaEventDisableTiming));
    }
    SP_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    SP_CUDA(nullptr, cudaEventCreateWithFlags(&ctx->ev_start, cudaEventDisableTiming));
    SP_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking));
    SP_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking));
    SP_CUDA(nullptr, cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking));
    ctx->d_act = ctx->d_act2[0];
    ctx->d_bucket = ctx->d_bucket2[0];

    uint8_t* b = ctx->d_net_blob;
    ctx->net.psq = reinterpret_cast<const uint4*>(b + L.psq);
    ctx->net.thr = reinterpret_cast<const uint4*>(b + L.thr);
    ctx->net.psq_planes = b + L.psq_planes;
    ctx->net.l1_w = reinterpret_cast<const int8_t*>(b + L.l1_w);
    ctx->net.l1_b = reinterpret_cast<const int32_t*>(b + L.l1_b);
    ctx->net.l2_w = reinterpret_cast<const int32_t*>(b + L.l2_w);
    ctx->net.l2_limbs = reinterpret_cast<const uint32_t*>(b + L.l2_limbs);
    ctx->net.l2_frags = reinterpret_cast<const uint32_t*>(b + L.l2_frags);
    ctx->net.l2_narrow = l2_narrow_mask;
    ctx->net.l2_b = reinterpret_cast<const int32_t*>(b + L.l2_b);
    ctx->net.l3_w = reinterpret_cast<const int32_t*>(b + L.l3_w);
    ctx->net.l3_b = reinterpret_cast<const int32_t*>(b + L.l3_b);
    ctx->net.tables = ctx->d_tables;
    return SP_OK;
}

int sp_nnue_create(const void* net_image, size_t len, int device, SpNnue** out) {
    if (!out) return fail(nullptr, SP_ERR_INVALID, "out is null");
    *out = nullptr;
    const int rc = create_impl(net_image, len, device, out);
    if (rc != SP_OK && *out) {
        sp_nnue_destroy(*out);
        *out = nullptr;
    }
    return rc;
}

void sp_nnue_destroy(SpNnue* ctx) {
    if (!ctx) return;
    DeviceGuard guard{ctx->device};
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    cudaFree(ctx->d_net_blob);
    cudaFree(ctx->d_tables);
    cudaFree(ctx->d_status);
    cudaFreeHost(ctx->h_status);
    if (ctx->aux) cudaStreamSynchronize(ctx->aux);
    for (int b = 0; b < 2; ++b) {
        cudaFree(ctx->d_act2[b]);
        cudaFree(ctx->d_bucket2[b]);
        cudaFree(ctx->d_lists[b]);
        cudaFree(ctx->d_group_overflow[b]);
        if (b == 0) cudaFreeHost(ctx->h_group_overflow);
        if (ctx->ev_group_count[b]) cudaEventDestroy(ctx->ev_group_count[b]);
        if (ctx->ev_ft[b]) cudaEventDestroy(ctx->ev_ft[b]);
        if (ctx->ev_head[b]) cudaEventDestroy(ctx->ev_head[b]);
    }
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->ev_start) cudaEventDestroy(ctx->ev_start);
    for (cudaEvent_t e : ctx->ev_chunk) cudaEventDestroy(e);
    if (ctx->aux) cudaStreamDestroy(ctx->aux);
    if (ctx->h2d) cudaStreamSynchronize(ctx->h2d), cudaStreamDestroy(ctx->h2d);
    if (ctx->d2h) cudaStreamSynchronize(ctx->d2h), cudaStreamDestroy(ctx->d2h);
    cudaFree(ctx->head_sort.order);
    cudaFree(ctx->head_sort.counters);
    cudaFree(ctx->plan.slot);
    cudaFree(ctx->plan.items);
    cudaFree(ctx->plan.acc);
    cudaFree(ctx->plan.counters);
    cudaFree(ctx->d_act_big);
    cudaFree(ctx->d_bucket_big);
    cudaFree(ctx->d_boards);
    cudaFree(ctx->d_out);
    cudaFree(ctx->d_ids);
    cudaFree(ctx->d_stm);
    cudaFree(ctx->slots.acc);
    cudaFree(ctx->slots.boards);
    cudaFree(ctx->d_small);
    cudaFreeHost(ctx->h_small);
    for (const auto& span : ctx->spans) cudaEventDestroy(span.a), cudaEventDestroy(span.b);
    for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
    if (ctx->stream && ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char* sp_nnue_last_error(const SpNnue* ctx) { return ctx ? ctx->error.c_str() : g_create_error.c_str(); }

int sp_nnue_device(const SpNnue* ctx) { return ctx ? ctx->device : -1; }

int sp_nnue_set_stream(SpNnue* ctx, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    DeviceGuard guard{ctx->device};
    SP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->owns_stream) cudaStreamDestroy(ctx->stream);
    ctx->owns_stream = stream == nullptr;
    if (stream) {
        ctx->stream = static_cast<cudaStream_t>(stream);
    } else {
        SP_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    }
    return SP_OK;
}

int sp_nnue_profile(SpNnue* ctx, int enable) {
    if (!ctx) return SP_ERR_INVALID;
    ctx->profiling = enable != 0;
    return SP_OK;
}

int sp_nnue_profile_read(SpNnue* ctx, double ms[SP_NUM_KERNEL_CLASSES], uint64_t launches[SP_NUM_KERNEL_CLASSES]) {
    if (!ctx || !ms || !launches) return SP_ERR_INVALID;
    DeviceGuard guard{ctx->device};
    for (const auto& span : ctx->spans) {
        SP_CUDA(ctx, cudaEventSynchronize(span.b));
        float t = 0;
        SP_CUDA(ctx, cudaEventElapsedTime(&t, span.a, span.b));
        ctx->prof_ms[span.kind] += t;
        ctx->prof_launches[span.kind] += 1;
        ctx->event_pool.push_back(span.a);
        ctx->event_pool.push_back(span.b);
    }
    ctx->spans.clear();
    for (int k = 0; k < SP_NUM_KERNEL_CLASSES; ++k) {
        ms[k] = ctx->prof_ms[k];
        launches[k] = ctx->prof_launches[k];
        ctx->prof_ms[k] = 0;
        ctx->prof_launches[k] = 0;
    }
    return SP_OK;
}

int sp_nnue_sync(SpNnue* ctx, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    DeviceGuard guard{ctx->device};
    return finish(ctx, pick(ctx, stream));
}

int sp_nnue_eval_full(SpNnue* ctx, const SpPackedBoard* boards, size_t n, int32_t* out) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!boards || !out) return fail(ctx, SP_ERR_INVALID, "null argument");
    DeviceGuard guard{ctx->device};
    if (const int rc = ensure_staging(ctx, n)) return rc;
    const HostIo io{boards, out};
    if (const int rc = eval_full_device(ctx, ctx->d_boards, n, ctx->d_out, ctx->stream, &io)) return rc;
    return finish(ctx, ctx->stream);
}

int sp_nnue_eval_full_device(SpNnue* ctx, const SpPackedBoard* d_boards, size_t n, int32_t* d_out, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!d_boards || !d_out) return fail(ctx, SP_ERR_INVALID, "null argument");
    if (reinterpret_cast<uintptr_t>(d_boards) & 15) return fail(ctx, SP_ERR_INVALID, "d_boards must be 16-byte aligned");
    DeviceGuard guard{ctx->device};
    return eval_full_device(ctx, d_boards, n, d_out, pick(ctx, stream));
}

int sp_nnue_activations_device(SpNnue* ctx, const SpPackedBoard* d_boards, size_t n, uint8_t* d_act, uint8_t* d_bucket, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!d_boards || !d_act || !d_bucket) return fail(ctx, SP_ERR_INVALID, "null argument");
    if ((reinterpret_cast<uintptr_t>(d_boards) | reinterpret_cast<uintptr_t>(d_act)) & 15)
        return fail(ctx, SP_ERR_INVALID, "d_boards and d_act must be 16-byte aligned");
    DeviceGuard guard{ctx->device};
    if (ctx->group) { /* chunk by chunk: the overflow scratch is sized for one chunk (launches are stream-ordered) */
        for (size_t off = 0; off < n; off += ctx->chunk) {
            const size_t m = std::min(ctx->chunk, n - off);
            SP_CUDA(ctx, launch_ft_group(ctx->net, d_boards + off, m, d_act + off * SP_L1_SIZE, d_bucket + off, ctx->d_status, ctx->d_group_overflow[0], ctx->sm_count, pick(ctx, stream)));
            ctx->counters[SP_CTR_LAUNCHES] += 2;
        }
    } else {
        launch_ft_full(ctx->net, d_boards, n, d_act, d_bucket, ctx->d_status, ctx->sm_count, pick(ctx, stream));
        ctx->counters[SP_CTR_LAUNCHES] += 1;
    }
    ctx->counters[SP_CTR_FULL_REFRESH] += 2 * n;
    SP_CUDA(ctx, cudaGetLastError());
    return SP_OK;
}

int sp_nnue_forward_device(SpNnue* ctx, const uint8_t* d_act, const uint8_t* d_bucket, size_t n, int32_t* d_out, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!d_act || !d_bucket || !d_out) return fail(ctx, SP_ERR_INVALID, "null argument");
    if (reinterpret_cast<uintptr_t>(d_act) & 15) return fail(ctx, SP_ERR_INVALID, "d_act must be 16-byte aligned");
    DeviceGuard guard{ctx->device};
    if (const int rc = ensure_head_sort(ctx, n)) return rc;
    {
        cudaStream_t st = pick(ctx, stream);
        Timed timed{ctx, st, SP_KERNEL_HEAD};
        HeadSort sort = ctx->head_sort;
        SpNnue::Span main{};
        if (ctx->profiling) { /* the head kernel proper, without the counting sort: what the roofline of this kernel is measured on */
            main = {take_event(ctx), take_event(ctx), SP_KERNEL_HEAD_MAIN};
            sort.ev_main_begin = main.a, sort.ev_main_end = main.b;
        }
        SP_CUDA(ctx, launch_head(ctx->net, d_act, d_bucket, n, d_out, nullptr, sort, ctx->sm_count, st));
        if (ctx->profiling) ctx->spans.push_back(main);
    }
    ctx->counters[SP_CTR_LAUNCHES] += head_kernel_launches(n, ctx->head_sort);
    ctx->counters[SP_CTR_EVALS] += n;
    SP_CUDA(ctx, cudaGetLastError());
    return SP_OK;
}

void sp_nnue_adjust_defaults(SpAdjustParams* p) {
    if (!p) return;
    *p = SpAdjustParams{{48, 442, 461, 637, 1223}, 26000, 2024, 1005, {0, 0}, {0, 0}}; /* src/tunable.h:161-169 */
}

int sp_nnue_adjust_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, const int32_t* d_raw, const int32_t* d_correction, size_t n,
    const SpAdjustParams* params, int32_t* d_out, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!d_boards || !d_raw || !d_out || !params) return fail(ctx, SP_ERR_INVALID, "null argument");
    DeviceGuard guard{ctx->device};
    launch_adjust(d_boards, d_raw, d_correction, n, *params, d_out, ctx->sm_count, pick(ctx, stream));
    ctx->counters[SP_CTR_LAUNCHES] += 1;
    SP_CUDA(ctx, cudaGetLastError());
    return SP_OK;
}

int sp_nnue_adjust(
    SpNnue* ctx, const SpPackedBoard* boards, const int32_t* raw, const int32_t* correction, size_t n,
    const SpAdjustParams* params, int32_t* out) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!boards || !raw || !out || !params) return fail(ctx, SP_ERR_INVALID, "null argument");
    DeviceGuard guard{ctx->device};
    if (const int rc = ensure_staging(ctx, n)) return rc;
    int32_t* d_raw = reinterpret_cast<int32_t*>(ctx->d_ids);              /* 3 * (cap + 1) words of scratch */
    int32_t* d_corr = reinterpret_cast<int32_t*>(ctx->d_ids) + ctx->cap + 1;
    SP_CUDA(ctx, cudaMemcpyAsync(ctx->d_boards, boards, n * sizeof(SpPackedBoard), cudaMemcpyHostToDevice, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(d_raw, raw, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    if (correction) SP_CUDA(ctx, cudaMemcpyAsync(d_corr, correction, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    launch_adjust(ctx->d_boards, d_raw, correction ? d_corr : nullptr, n, *params, ctx->d_out, ctx->sm_count, ctx->stream);
    ctx->counters[SP_CTR_LAUNCHES] += 1;
    SP_CUDA(ctx, cudaGetLastError());
    SP_CUDA(ctx, cudaMemcpyAsync(out, ctx->d_out, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    return finish(ctx, ctx->stream);
}

int sp_nnue_wdl_device(
    SpNnue* ctx, const SpPackedBoard* d_boards, const int32_t* d_scores, size_t n, int32_t* d_normalized, int32_t* d_win,
    int32_t* d_loss, void* stream) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!d_boards || !d_scores || (!d_normalized && !d_win) || (!d_win != !d_loss)) return fail(ctx, SP_ERR_INVALID, "null argument");
    DeviceGuard guard{ctx->device};
    launch_wdl(d_boards, d_scores, n, d_normalized, d_win, d_loss, ctx->sm_count, pick(ctx, stream));
    ctx->counters[SP_CTR_LAUNCHES] += 1;
    SP_CUDA(ctx, cudaGetLastError());
    return SP_OK;
}

int sp_nnue_wdl(
    SpNnue* ctx, const SpPackedBoard* boards, const int32_t* scores, size_t n, int32_t* normalized, int32_t* win, int32_t* loss) {
    if (!ctx) return SP_ERR_INVALID;
    if (!n) return SP_OK;
    if (!boards || !scores || (!normalized && !win) || (!win != !loss)) return fail(ctx, SP_ERR_INVALID, "null argument");
    DeviceGuard guard{ctx->device};
    if (const int rc = ensure_staging(ctx, n)) return rc;
    int32_t* d_scores = reinterpret_cast<int32_t*>(ctx->d_ids); /* 3 * (cap + 1) words of scratch: scores, win, loss */
    int32_t* d_win = d_scores + ctx->cap + 1;
    int32_t* d_loss = d_win + ctx->cap + 1;
    SP_CUDA(ctx, cudaMemcpyAsync(ctx->d_boards, boards, n * sizeof(SpPackedBoard), cudaMemcpyHostToDevice, ctx->stream));
    SP_CUDA(ctx, cudaMemcpyAsync(d_scores, scores, n * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    launch_wdl(ctx->d_boards, d_scores, n, normalized ? ctx->d_out : nullptr, win ? d_win : nullptr, win ? d_loss : nullptr, ctx->sm_count, ctx->stream);
    ctx->counters[SP_CTR_LAUNCHES] += 1;
    SP_CUDA(ctx, cudaGetLastError());
    if (normalized) SP_CUDA(ctx, cudaMemcpyAsync(normalized, ctx->d_out, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (win) {
        SP_CUDA(ctx, cudaMemcpyAsync(win, d_win, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
        SP_CUDA(ctx, cudaMemcpyAsync(loss, d_loss, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    }
    return finish(ctx, ctx->stream);
}

int sp_nnue_counters(SpNnue* ctx, uint64_t out[SP_NUM_COUNTERS]) {
    if (!ctx || !out) return SP_ERR_INVALID;
    std::memcpy(out, ctx->counters, sizeof(ctx->counters));
    return SP_OK;
}

} // extern "C"

/* ------------------------------------------------------------------ accumulator slots / playouts */
#include "capi_slots.inc"
