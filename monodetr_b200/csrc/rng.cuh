// rng.cuh -- counter-based hash RNG for dropout masks.  mask(seed, index) is a pure function, so the
// backward pass regenerates exactly the mask the forward pass used (no mask tensors in HBM), and the
// seed lives in device memory so a captured CUDA graph sees a fresh value every replay.
#pragma once
#include <stdint.h>

namespace mdb {

__host__ __device__ __forceinline__ uint64_t fmix64(uint64_t x) {   // MurmurHash3 finalizer
    x ^= x >> 33;
    x *= 0xff51afd7ed558ccdull;
    x ^= x >> 33;
    x *= 0xc4ceb9fe1a85ec53ull;
    x ^= x >> 33;
    return x;
}

// uniform in [0, 1) with 24 bits of resolution
__host__ __device__ __forceinline__ float rng_uniform(uint64_t seed, uint64_t idx) {
    const uint64_t h = fmix64(idx * 0x9E3779B97F4A7C15ull + seed);
    return (float)(uint32_t)(h >> 40) * (1.0f / 16777216.0f);
}

// four independent uniforms for indices 4*quad .. 4*quad+3 from one hash (16 bits each)
__host__ __device__ __forceinline__ void rng_uniform4(uint64_t seed, uint64_t quad, float (&u)[4]) {
    const uint64_t h = fmix64(quad * 0x9E3779B97F4A7C15ull + seed);
    u[0] = (float)(uint32_t)(h & 0xFFFF) * (1.0f / 65536.0f);
    u[1] = (float)(uint32_t)((h >> 16) & 0xFFFF) * (1.0f / 65536.0f);
    u[2] = (float)(uint32_t)((h >> 32) & 0xFFFF) * (1.0f / 65536.0f);
    u[3] = (float)(uint32_t)((h >> 48) & 0xFFFF) * (1.0f / 65536.0f);
}

// 32-bit variant for per-element masks inside tight loops (attention probabilities): key = one 64-bit hash per
// (seed, site, batch*head) computed once per CTA, then ~10 integer instructions per element.  idx < 2^32.
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t x) {   // MurmurHash3 32-bit finalizer
    x ^= x >> 16;
    x *= 0x85ebca6bu;
    x ^= x >> 13;
    x *= 0xc2b2ae35u;
    x ^= x >> 16;
    return x;
}
__host__ __device__ __forceinline__ uint32_t rng_key32(uint64_t seed, uint64_t stream) {
    return (uint32_t)(fmix64(stream * 0x9E3779B97F4A7C15ull + seed) >> 17);
}
// Keep decisions of the attention-probability dropout: ONE hash serves the two elements 2k and 2k+1, and a decision is a
// single unsigned compare of the hash against thr16 << 16 (i.e. of its top 16 bits against thr16), so a kernel that walks a
// row in pairs pays ~3.5 integer instructions per element (the elementwise phase of the tensor-core attention kernels is
// instruction-bound, and the previous fmix32-per-element mask doubled it).  keep(idx) is true with probability
// 1 - thr16 / 65536, thr16 = rng_thr16(drop_p).  The same pure function in every kernel (both forwards, all backwards).
__host__ __device__ __forceinline__ uint32_t rng_pair32(uint32_t key, uint32_t pair) {   // decision word of element 2 * pair
    uint32_t x = pair * 0x9E3779B1u + key;
    x ^= x >> 15;
    x *= 0x2C1B3C6Du;
    x ^= x >> 12;
    return x;
}
__host__ __device__ __forceinline__ uint32_t rng_pair32_odd(uint32_t h) { return h * 0x297A2D39u; }   // ... of element 2 * pair + 1
__host__ __device__ __forceinline__ uint32_t rng_thr16(float drop_p) {
    const float t = drop_p * 65536.f;
    return t >= 65535.f ? 65535u : (uint32_t)t;
}
__host__ __device__ __forceinline__ bool rng_keep16(uint32_t key, uint32_t idx, uint32_t thr16) {
    uint32_t h = rng_pair32(key, idx >> 1);
    if (idx & 1u) h = rng_pair32_odd(h);
    return h >= (thr16 << 16);
}

}  // namespace mdb
