/* rng.h -- the random streams of the synthetic workloads (playouts, self-play openings). */
#ifndef SP_HOST_RNG_H
#define SP_HOST_RNG_H

#include <cstdint>

#if defined(__CUDACC__)
    #define SP_RNG_HD __host__ __device__
#else
    #define SP_RNG_HD
#endif

namespace sp::host {

/* splitmix64 seeding + JSF64 stream + Lemire's bounded draw: public-domain generators, the same
 * family the reference uses for datagen (src/util/rng.h), re-stated so playouts are reproducible
 * from a (seed, game index) pair on any machine. */
struct SplitMix64 {
    uint64_t s;
    SP_RNG_HD uint64_t next() {
        s += 0x9E3779B97F4A7C15ULL;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
        return z ^ (z >> 31);
    }
};

struct Jsf64 {
    uint64_t a{0xF1EA5EED}, b, c, d;
    SP_RNG_HD explicit Jsf64(uint64_t seed) : b{seed}, c{seed}, d{seed} {
        for (int i = 0; i < 20; ++i) next();
    }
    SP_RNG_HD static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
    SP_RNG_HD uint64_t next() {
        const uint64_t e = a - rotl(b, 7);
        a = b ^ rotl(c, 13);
        b = c + rotl(d, 37);
        c = d + e;
        d = e + a;
        return d;
    }
    SP_RNG_HD uint32_t below(uint32_t bound) {
        if (!bound) return 0;
        uint32_t x = static_cast<uint32_t>(next() >> 32);
        uint64_t m = static_cast<uint64_t>(x) * bound;
        uint32_t l = static_cast<uint32_t>(m);
        if (l < bound) {
            const uint32_t t = (0u - bound) % bound;
            while (l < t) {
                x = static_cast<uint32_t>(next() >> 32);
                m = static_cast<uint64_t>(x) * bound;
                l = static_cast<uint32_t>(m);
            }
        }
        return static_cast<uint32_t>(m >> 32);
    }
};

} // namespace sp::host

#endif
