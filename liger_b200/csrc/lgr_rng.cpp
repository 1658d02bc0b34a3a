// liger_b200 -- R's default RNG stream (Mersenne-Twister, "Inversion"), so that the seeded
// initialisation of .runINMF.list (reference R/integration.R:412-437: set.seed(seed); runif(n, 0, 2))
// is literally the same numbers.  Algorithm: R's RNG.c scrambling of the seed (50 + 625 LCG steps,
// dummy[0] = mti = N) followed by standard MT19937 genrand_int32 scaled by 2^-32 and kept inside (0,1).
#include <stdint.h>

#include <new>

#include "../../include/liger_b200.h"

struct lgr_rng {
    uint32_t mt[624];
    int mti;
};

extern "C" lgr_rng* lgr_r_set_seed(uint32_t seed) {
    lgr_rng* r = new (std::nothrow) lgr_rng;
    if (!r) return nullptr;
    uint32_t s = seed;
    for (int j = 0; j < 50; ++j) s = 69069u * s + 1u;
    uint32_t dummy[625];
    for (int j = 0; j < 625; ++j) {
        s = 69069u * s + 1u;
        dummy[j] = s;
    }
    // FixupSeeds(MERSENNE_TWISTER): dummy[0] = 624 (mti), i.e. the state is regenerated on the first draw
    for (int j = 0; j < 624; ++j) r->mt[j] = dummy[j + 1];
    r->mti = 624;
    return r;
}

static inline uint32_t mt_next(lgr_rng* r) {
    static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
    uint32_t* mt = r->mt;
    if (r->mti >= 624) {
        int kk;
        uint32_t y;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < 623; kk++) {
            y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
            mt[kk] = mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (mt[623] & 0x80000000u) | (mt[0] & 0x7fffffffu);
        mt[623] = mt[396] ^ (y >> 1) ^ mag01[y & 1u];
        r->mti = 0;
    }
    uint32_t y = mt[r->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

extern "C" void lgr_r_runif(lgr_rng* r, int64_t n, double lo, double hi, double* out) {
    const double i2_32m1 = 2.328306437080797e-10;
    for (int64_t i = 0; i < n; ++i) {
        double u = (double)mt_next(r) * 2.3283064365386963e-10;
        if (u <= 0.0) u = 0.5 * i2_32m1;
        if (1.0 - u <= 0.0) u = 1.0 - 0.5 * i2_32m1;
        out[i] = lo + (hi - lo) * u;
    }
}

extern "C" void lgr_r_free(lgr_rng* r) { delete r; }
