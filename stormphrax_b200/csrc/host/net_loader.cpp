/* net_loader.cpp -- see net_loader.h. */
#include "net_loader.h"

#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "../../../include/sp_nnue.h"

namespace sp::host {

namespace {

struct Zstd {
    size_t (*decompress)(void*, size_t, const void*, size_t) = nullptr;
    unsigned (*isError)(size_t) = nullptr;
    const char* (*getErrorName)(size_t) = nullptr;
    bool ok = false;
};

const Zstd& zstd() {
    static Zstd z;
    static std::once_flag once;
    std::call_once(once, [] {
        void* h = nullptr;
        for (const char* name : {"libzstd.so.1", "libzstd.so"}) {
            h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
            if (h) break;
        }
        if (!h) return;
        z.decompress = reinterpret_cast<decltype(z.decompress)>(dlsym(h, "ZSTD_decompress"));
        z.isError = reinterpret_cast<decltype(z.isError)>(dlsym(h, "ZSTD_isError"));
        z.getErrorName = reinterpret_cast<decltype(z.getErrorName)>(dlsym(h, "ZSTD_getErrorName"));
        z.ok = z.decompress && z.isError && z.getErrorName;
    });
    return z;
}

} // namespace

const uint8_t* network_payload(const uint8_t* image, size_t len, size_t payload_bytes, std::vector<uint8_t>& storage, std::string& error) {
    const unsigned flags = image[6] | image[7] << 8;
    const uint8_t* body = image + SP_NET_HEADER_BYTES;
    const size_t body_len = len - SP_NET_HEADER_BYTES;
    if (!(flags & kNetFlagZstd)) {
        if (body_len < payload_bytes) {
            error = "network too small? " + std::to_string(body_len) + " < " + std::to_string(payload_bytes);
            return nullptr;
        }
        return body;
    }
    const Zstd& z = zstd();
    if (!z.ok) {
        error = "zstd-compressed network, but libzstd.so.1 could not be loaded on this host";
        return nullptr;
    }
    storage.resize(payload_bytes);
    const size_t got = z.decompress(storage.data(), payload_bytes, body, body_len); /* nnue.cpp:230-235 */
    if (z.isError(got)) {
        error = std::string{"failed to decompress network: "} + z.getErrorName(got);
        return nullptr;
    }
    if (got < payload_bytes) {
        error = "decompressed network too small? " + std::to_string(got) + " < " + std::to_string(payload_bytes);
        return nullptr;
    }
    return storage.data();
}

} // namespace sp::host

/* Host-only form for tools and tests: the logical payload of a (possibly compressed) network image.
 * Returns the payload size, or -1 (see sp_nnue.h). */
extern "C" long sp_host_net_payload(const void* net_image, size_t len, void* out, size_t cap) {
    if (!net_image || len < SP_NET_HEADER_BYTES || std::memcmp(net_image, "CBNF", 4) != 0) return -1;
    std::vector<uint8_t> storage;
    std::string error;
    const uint8_t* p = sp::host::network_payload(static_cast<const uint8_t*>(net_image), len, SP_NET_PAYLOAD_BYTES, storage, error);
    if (!p) return -1;
    if (out) {
        if (cap < SP_NET_PAYLOAD_BYTES) return -1;
        std::memcpy(out, p, SP_NET_PAYLOAD_BYTES);
    }
    return SP_NET_PAYLOAD_BYTES;
}
