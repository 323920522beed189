/*
 * ux_rng.h -- the counter-based random stream of the B200 backend (Philox4x32-10).
 *
 * Replaces the reference's per-entity std::mt19937 streams (trafficpp/utils.h:33-36 make_entity_rng,
 * traffi.h:404-411) with a stateless stream keyed on (seed, timestep, entity, purpose, draw#), so that a
 * draw does not depend on execution order: any parallel schedule of nodes / links / vehicles produces
 * the same numbers as a serial sweep.  The reference's Python engine (one global PCG64) and C++ engine
 * (per-node / per-link mt19937) are already not stream-compatible with each other, so stochastic
 * parity with the reference is statistical (BASELINE.json); CUDA-vs-oracle parity is bit-exact because
 * both include this header.
 *
 * Philox4x32-10: Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3", SC'11.
 * Known-answer vectors (Random123 kat_vectors) are checked in tests/test_rng.py.
 *
 * Usable from C99, C++ and CUDA device code.
 */
#ifndef UX_RNG_H
#define UX_RNG_H

#include <stdint.h>

#if defined(__CUDACC__)
#define UX_HD __host__ __device__ __forceinline__
#else
#define UX_HD static inline
#endif

/* purposes (counter word 2) */
#define UX_RNG_GENERATE 1u /* Node::generate route draw      : entity = node id,    draw# = attempt   */
#define UX_RNG_TRANSFER 2u /* Node::transfer shuffle + merge : entity = node id,    draw# = running   */
#define UX_RNG_VEHROUTE 3u /* Vehicle::update route draw     : entity = vehicle id, draw# = 0         */
#define UX_RNG_LINKNOISE 4u /* update_adj_time_matrix noise  : entity = link id,    draw# = 0         */

typedef struct ux_u32x4 {
    uint32_t v[4];
} ux_u32x4;

UX_HD uint32_t ux_mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * (uint64_t)b) >> 32);
#endif
}

UX_HD ux_u32x4 ux_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; r++) {
        uint32_t hi0 = ux_mulhi32(M0, c0), lo0 = M0 * c0;
        uint32_t hi1 = ux_mulhi32(M1, c2), lo1 = M1 * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    ux_u32x4 out;
    out.v[0] = c0; out.v[1] = c1; out.v[2] = c2; out.v[3] = c3;
    return out;
}

/* uniform double in [0, 1) with 53 random bits */
UX_HD double ux_rng_uniform(int64_t seed, uint32_t timestep, uint32_t entity, uint32_t purpose, uint32_t draw) {
    ux_u32x4 r = ux_philox4x32_10(timestep, entity, purpose, draw, (uint32_t)((uint64_t)seed & 0xFFFFFFFFu),
                                  (uint32_t)((uint64_t)seed >> 32));
    uint64_t a = (uint64_t)(r.v[0] >> 5), b = (uint64_t)(r.v[1] >> 6);
    return (double)(a * 67108864ull + b) * (1.0 / 9007199254740992.0);
}

/* uniform integer in [0, n) (n >= 1) */
UX_HD int ux_rng_below(int64_t seed, uint32_t timestep, uint32_t entity, uint32_t purpose, uint32_t draw, int n) {
    double u = ux_rng_uniform(seed, timestep, entity, purpose, draw);
    int k = (int)(u * (double)n);
    return k >= n ? n - 1 : k;
}

#endif /* UX_RNG_H */
