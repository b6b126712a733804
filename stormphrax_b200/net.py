"""Stormphrax network file format (``CBNF``) -- reader, validator and synthetic generator.

The reference downloads its trained network at build time (``Makefile:51-54``); no weights
exist in this sandbox, so every test and benchmark runs on a *synthetic* network written in
the reference's own on-disk format:

* 64-byte packed header, ``src/eval/header.h:38-52``; validation rules follow
  ``src/eval/nnue.cpp:85-185`` (magic, version, arch id, flags, activation, sizes).
* payload = raw little-endian arrays, each starting 64-byte aligned, in the order the
  reference's loader consumes them (``src/eval/nnue/input.h:359-361``,
  ``src/eval/nnue/arch/multilayer.h:492-496``)::

      psqW  i16[11264][1024]   thrW  i8[64368][1024]   ftBias i16[1024]
      l1W   i8[8][1024*32]     l1B   i32[8][32]
      l2W   i32[8][64][64]     l2B   i32[8][64]
      l3W   i32[8][64]         l3B   i32[8]

This is the LOGICAL (un-permuted) layout.  x86 builds of the reference permute the FT for
their ``packus`` lane order at load/build time (``multilayer.h:504-547``); the GPU library
and the oracles all take the logical file.
"""
from __future__ import annotations

import dataclasses
import struct

import numpy as np

L1 = 1024
L2 = 32
L3 = 64
OUT_BUCKETS = 8
IN_BUCKETS = 16
PSQ_FEATURES = IN_BUCKETS * 704
THREAT_FEATURES = 59808 + 96 * 95 // 2
HEADER_BYTES = 64

FLAG_ZSTD = 0x0001
FLAG_MIRRORED = 0x0002
FLAG_MERGED_KINGS = 0x0004
FLAG_PAIRWISE = 0x0008
ARCH_ID = 5  # 2 + dual activation + 2 * skip-L2  (multilayer.h:51)

# (name, dtype, shape) in file order
LAYOUT = (
    ("psq_w", np.int16, (PSQ_FEATURES, L1)),
    ("thr_w", np.int8, (THREAT_FEATURES, L1)),
    ("ft_b", np.int16, (L1,)),
    ("l1_w", np.int8, (OUT_BUCKETS, L1 * L2)),
    ("l1_b", np.int32, (OUT_BUCKETS, L2)),
    ("l2_w", np.int32, (OUT_BUCKETS, 2 * L2, L3)),
    ("l2_b", np.int32, (OUT_BUCKETS, L3)),
    ("l3_w", np.int32, (OUT_BUCKETS, L3)),
    ("l3_b", np.int32, (OUT_BUCKETS,)),
)

PAYLOAD_BYTES = sum(int(np.prod(shape)) * np.dtype(dt).itemsize for _, dt, shape in LAYOUT)
assert PAYLOAD_BYTES == 89_381_920
FILE_BYTES = HEADER_BYTES + PAYLOAD_BYTES


class NetworkFormatError(ValueError):
    pass


def make_header(name: str = "synthetic") -> bytes:
    raw = name.encode()[:48]
    hdr = struct.pack(
        "<4sHHBBBHBBB48s",
        b"CBNF",
        1,
        FLAG_MIRRORED | FLAG_MERGED_KINGS | FLAG_PAIRWISE,
        0,
        ARCH_ID,
        0,  # ClippedReLU::kId
        L1,
        0x80 | IN_BUCKETS,
        OUT_BUCKETS,
        len(raw),
        raw,
    )
    assert len(hdr) == HEADER_BYTES
    return hdr


def validate_header(buf: bytes | memoryview) -> None:
    """Same checks, same order, as ``validate`` in src/eval/nnue.cpp:85-185."""
    if len(buf) < HEADER_BYTES:
        raise NetworkFormatError("missing network header")
    magic, version, flags, _pad, arch, act, hidden, in_b, out_b, _nlen, _name = struct.unpack(
        "<4sHHBBBHBBB48s", bytes(buf[:HEADER_BYTES])
    )
    if magic != b"CBNF":
        raise NetworkFormatError("invalid magic bytes in network header")
    if version != 1:
        raise NetworkFormatError(f"unsupported network format version {version} (expected: 1)")
    if arch != ARCH_ID:
        raise NetworkFormatError(f"wrong network architecture {arch} (expected: {ARCH_ID})")
    if not flags & FLAG_MIRRORED:
        raise NetworkFormatError("unmirrored network, expected horizontally mirrored")
    if not flags & FLAG_MERGED_KINGS:
        raise NetworkFormatError("network does not have merged king planes, expected merged")
    if not flags & FLAG_PAIRWISE:
        raise NetworkFormatError("network L1 does not require pairwise multiplication, expected paired")
    if act != 0:
        raise NetworkFormatError(f"wrong l1 activation function {act} (expected: crelu)")
    if hidden != L1:
        raise NetworkFormatError(f"wrong number of l1 neurons {hidden} (expected: {L1})")
    if not in_b & 0x80:
        raise NetworkFormatError("network does not have the expected threat inputs")
    if in_b & 0x7F != IN_BUCKETS:
        raise NetworkFormatError(f"wrong number of input buckets {in_b & 0x7F} (expected: {IN_BUCKETS})")
    if out_b != OUT_BUCKETS:
        raise NetworkFormatError(f"wrong number of output buckets {out_b} (expected: {OUT_BUCKETS})")


@dataclasses.dataclass
class Network:
    """Logical network arrays (views into one contiguous file image)."""

    image: np.ndarray  # uint8[FILE_BYTES]
    psq_w: np.ndarray
    thr_w: np.ndarray
    ft_b: np.ndarray
    l1_w: np.ndarray
    l1_b: np.ndarray
    l2_w: np.ndarray
    l2_b: np.ndarray
    l3_w: np.ndarray
    l3_b: np.ndarray

    def tobytes(self) -> bytes:
        return self.image.tobytes()


def _views(image: np.ndarray) -> dict[str, np.ndarray]:
    out = {}
    off = HEADER_BYTES
    for name, dt, shape in LAYOUT:
        nbytes = int(np.prod(shape)) * np.dtype(dt).itemsize
        assert off % 64 == 0, "every parameter block must start 64-byte aligned (loader.cpp:40)"
        out[name] = image[off : off + nbytes].view(dt).reshape(shape)
        off += nbytes
    assert off == FILE_BYTES
    return out


def from_bytes(buf: bytes | np.ndarray) -> Network:
    image = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    validate_header(memoryview(image)[:HEADER_BYTES])
    if image.size < FILE_BYTES:
        raise NetworkFormatError(f"network too small? {image.size - HEADER_BYTES} < {PAYLOAD_BYTES}")
    image = image[:FILE_BYTES]
    return Network(image=image, **_views(image))


def compressed(net: "Network", level: int = 3) -> np.ndarray:
    """The same network as a zstd-flagged file image (header flag kZstdCompressed + one zstd frame of the arrays,
    what the reference's release builds embed: src/eval/nnue.cpp:215-247).  Test helper; needs pyarrow's zstd codec."""
    import pyarrow as pa

    body = pa.compress(net.image[HEADER_BYTES:].tobytes(), codec="zstd", asbytes=True)
    hdr = bytearray(net.image[:HEADER_BYTES].tobytes())
    flags = struct.unpack_from("<H", hdr, 6)[0] | FLAG_ZSTD
    struct.pack_into("<H", hdr, 6, flags)
    return np.frombuffer(bytes(hdr) + body, dtype=np.uint8)


def load(path: str) -> Network:
    return from_bytes(np.fromfile(path, dtype=np.uint8))


def synthetic(seed: int = 1234, stress: bool = False, tame: bool = False) -> Network:
    """Random network in the reference's format.

    ``stress=False`` uses the ranges from SURVEY.md section 8(d): both sides of every clamp are
    exercised and ``x*x`` occasionally wraps in 32 bits.  ``stress=True`` uses full-range FT and
    dense-layer weights so that the int16 accumulators and the int32 L2/L3 sums wrap constantly --
    a bit-exactness torture test, not a plausible network.  ``tame=True`` keeps the default ranges but shrinks the
    output layer so that evaluations stay within a few hundred centipawns: the default network's outputs reach
    +-50,000, beyond the engine's win score (25,000, src/core.h:708), which sends the reference's alpha-beta search
    into pathological trees -- the search-level tests (oracle/engine) need scores a search can work with.
    """
    rng = np.random.default_rng(seed)
    image = np.zeros(FILE_BYTES, dtype=np.uint8)
    image[:HEADER_BYTES] = np.frombuffer(make_header("stress" if stress else "synthetic"), dtype=np.uint8)
    v = _views(image)
    if not stress:
        v["psq_w"][...] = rng.integers(-40, 41, v["psq_w"].shape, dtype=np.int16)
        v["thr_w"][...] = rng.integers(-6, 7, v["thr_w"].shape, dtype=np.int8)
        v["ft_b"][...] = rng.integers(-100, 200, v["ft_b"].shape, dtype=np.int16)
        v["l1_w"][...] = rng.integers(-127, 128, v["l1_w"].shape, dtype=np.int8)
        v["l1_b"][...] = rng.integers(-2000, 2000, v["l1_b"].shape, dtype=np.int32)
        v["l2_w"][...] = rng.integers(-300, 300, v["l2_w"].shape, dtype=np.int32)
        v["l2_b"][...] = rng.integers(-100000, 100001, v["l2_b"].shape, dtype=np.int32)
        v["l3_w"][...] = rng.integers(-300, 300, v["l3_w"].shape, dtype=np.int32)
        v["l3_b"][...] = rng.integers(-1000000, 1000001, v["l3_b"].shape, dtype=np.int32)
        if tame:
            v["l3_w"][...] = rng.integers(-3, 4, v["l3_w"].shape, dtype=np.int32)
    else:
        v["psq_w"][...] = rng.integers(-32768, 32768, v["psq_w"].shape, dtype=np.int16)
        v["thr_w"][...] = rng.integers(-128, 128, v["thr_w"].shape, dtype=np.int8)
        v["ft_b"][...] = rng.integers(-32768, 32768, v["ft_b"].shape, dtype=np.int16)
        v["l1_w"][...] = rng.integers(-128, 128, v["l1_w"].shape, dtype=np.int8)
        v["l1_b"][...] = rng.integers(-60000, 60000, v["l1_b"].shape, dtype=np.int32)
        v["l2_w"][...] = rng.integers(-(2**31), 2**31, v["l2_w"].shape, dtype=np.int64).astype(np.int32)
        v["l2_b"][...] = rng.integers(-(2**31), 2**31, v["l2_b"].shape, dtype=np.int64).astype(np.int32)
        v["l3_w"][...] = rng.integers(-(2**31), 2**31, v["l3_w"].shape, dtype=np.int64).astype(np.int32)
        v["l3_b"][...] = rng.integers(-(2**31), 2**31, v["l3_b"].shape, dtype=np.int64).astype(np.int32)
    return Network(image=image, **v)
