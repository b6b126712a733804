/*
 * net_loader.h -- the host half of the network loader: from a CBNF file image (src/eval/header.h:38-52) to the
 * logical payload bytes the device image is built from.  Follows eval::init, src/eval/nnue.cpp:200-263: the 64-byte
 * header is never compressed; when its kZstdCompressed flag (header.h:28-34) is set, everything after it is one
 * zstd frame that decompresses to the raw arrays (nnue.cpp:226-247), otherwise the arrays follow directly.
 *
 * The reference links a vendored copy of the zstd decoder (3rdparty/zstd/zstddeclib.c); this library binds the
 * system's libzstd.so.1 at run time (dlopen: ZSTD_decompress / ZSTD_isError / ZSTD_getErrorName, the same three
 * calls nnue.cpp makes) and reports its absence as a load error.
 */
#ifndef SP_HOST_NET_LOADER_H
#define SP_HOST_NET_LOADER_H

#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace sp::host {

constexpr unsigned kNetFlagZstd = 0x0001; /* NetworkFlags::kZstdCompressed, src/eval/header.h:28-34 */

/* Payload of `image` (len bytes, header already validated): a pointer into `image` for a raw network, or into
 * `storage` after decompression.  Returns nullptr and sets `error` (the reference's messages, nnue.cpp:238-246) when
 * the payload is shorter than `payload_bytes` or cannot be decompressed. */
const uint8_t* network_payload(const uint8_t* image, size_t len, size_t payload_bytes, std::vector<uint8_t>& storage, std::string& error);

} // namespace sp::host

#endif
