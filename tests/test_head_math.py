"""The integer identities the dense head's tensor-core L2 rests on (kernels.cu: head_stream_kernel), checked in
numpy.  The kernels themselves are held against the oracle on the GPU (tests/test_gpu_full.py); this file pins the
arithmetic argument: an int32 x int32 product modulo 2^32 from u8 x u8 (or u8 x s8) partial products."""
import numpy as np

M32 = 1 << 32


def limbs(x):
    """unsigned byte limbs of the 32-bit two's complement of x (int64 array in int32 range)"""
    u = np.asarray(x, dtype=np.int64) & (M32 - 1)
    return [(u >> (8 * i)) & 0xFF for i in range(4)]


def test_ten_limb_products_give_the_product_modulo_2_32():
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.integers(-(1 << 19), 4097, 4000), [-(1 << 19), -1, 0, 1, 4096]])       # L2 inputs (multilayer.h:219-256)
    w = np.concatenate([rng.integers(-(1 << 31), 1 << 31, 4000), [-(1 << 31), (1 << 31) - 1, -1, 0, 1]])  # any int32 weight
    a, w = np.meshgrid(a, w[:64])
    al, wl = limbs(a), limbs(w)
    want = (a * w) & (M32 - 1)
    # sum over i + j <= 3 of (a_i w_j) << 8 (i + j): ten products, everything above bit 31 drops out
    got = np.zeros_like(want)
    for i in range(4):
        for j in range(4 - i):
            got = (got + ((al[i] * wl[j]) << (8 * (i + j)))) & (M32 - 1)
    assert (got == want).all()
    # the same by Horner's rule over the limb weight, as the kernel accumulates it (acc = (acc << 8) + level)
    acc = np.zeros_like(want)
    for level in (3, 2, 1, 0):
        acc = (acc << 8) & (M32 - 1)
        for i in range(level + 1):
            acc = (acc + al[i] * wl[level - i]) & (M32 - 1)
    assert (acc == want).all()


def test_zero_input_limbs_can_be_skipped():
    """inputs in [0, 2^16) -- the CReLU half always, the squared half unless the square wrapped -- have limbs 2, 3 = 0"""
    rng = np.random.default_rng(1)
    a = rng.integers(0, 4097, 3000)
    w = rng.integers(-(1 << 31), 1 << 31, 3000)
    al, wl = limbs(a), limbs(w)
    assert not al[2].any() and not al[3].any()
    got = np.zeros_like(a)
    for i in range(2):
        for j in range(4 - i):
            got = (got + ((al[i] * wl[j]) << (8 * (i + j)))) & (M32 - 1)
    assert (got == ((a * w) & (M32 - 1))).all()


def test_int16_weights_need_two_limbs_with_a_signed_top():
    """w = lo + 256 hi with lo = byte 0 unsigned and hi = byte 1 SIGNED (the same stored bytes, read u8 x s8):
    a w = a0 lo + 2^8 (a0 hi + a1 lo) + 2^16 a1 hi for a in [0, 2^16), exactly."""
    rng = np.random.default_rng(2)
    a = np.concatenate([rng.integers(0, 1 << 16, 3000), [0, 4096, 65535]])
    w = np.concatenate([rng.integers(-(1 << 15), 1 << 15, 3000), [-(1 << 15), (1 << 15) - 1, -1]])
    a, w = a[: len(w)], w[: len(a)]
    a0, a1 = a & 0xFF, a >> 8
    lo = w & 0xFF
    hi = ((w >> 8) & 0xFF).astype(np.int8).astype(np.int64)  # byte 1 of the two's complement, sign-interpreted
    assert (lo + 256 * hi == w).all()
    acc = a1 * hi
    acc = (acc << 8) + a0 * hi + a1 * lo
    acc = (acc << 8) + a0 * lo
    assert ((acc & (M32 - 1)) == ((a * w) & (M32 - 1))).all()
