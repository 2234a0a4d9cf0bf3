/*
 * pfd_synth.h -- point-evaluable synthetic terrain + steepest-descent D8, shared by the
 * CUDA library (bench / large-size tests) and the CPU oracle (small-size tests).
 *
 * This is NOT part of the reference's hot path; it is the input generator SURVEY.md §8(d)
 * specifies ("random steepest-descent on Perlin DEM"): z = sum_o 2^-o * perlin_o(x, y) plus a
 * 1e-4 * U(0,1) tie-breaker, D8 = strict steepest descent over the 8 neighbours in the fixed
 * order NW,N,NE,W,E,SW,S,SE, pit (code 0) when no neighbour is lower, optional "sea" nodata
 * (code 247) below a threshold.
 *
 * Every float operation goes through PFD_F* macros that map to the non-contractable
 * __f*_rn intrinsics on the device and to plain IEEE ops on the host (compile the host side
 * with -ffp-contract=off), so host and device produce bit-identical rasters.
 */
#ifndef PFD_SYNTH_H
#define PFD_SYNTH_H

#include <stdint.h>

#ifdef __CUDA_ARCH__
#define PFD_HD __host__ __device__ __forceinline__
#define PFD_FADD(a, b) __fadd_rn((a), (b))
#define PFD_FSUB(a, b) __fsub_rn((a), (b))
#define PFD_FMUL(a, b) __fmul_rn((a), (b))
#elif defined(__CUDACC__)
#define PFD_HD __host__ __device__ __forceinline__
#define PFD_FADD(a, b) ((a) + (b))
#define PFD_FSUB(a, b) ((a) - (b))
#define PFD_FMUL(a, b) ((a) * (b))
#else
#define PFD_HD static inline
#define PFD_FADD(a, b) ((a) + (b))
#define PFD_FSUB(a, b) ((a) - (b))
#define PFD_FMUL(a, b) ((a) * (b))
#endif

/* 32-bit avalanche hash (lowbias32-style), integer only => identical everywhere */
PFD_HD uint32_t pfd_hash_u32(uint32_t x) {
    x ^= x >> 16;
    x *= 0x7feb352dU;
    x ^= x >> 15;
    x *= 0x846ca68bU;
    x ^= x >> 16;
    return x;
}

PFD_HD uint32_t pfd_hash3(uint32_t a, uint32_t b, uint32_t c) {
    return pfd_hash_u32(a * 0x9E3779B1U ^ pfd_hash_u32(b * 0x85EBCA77U ^ pfd_hash_u32(c + 0xC2B2AE3DU)));
}

/* dot product of one of 8 fixed lattice gradients with (dx, dy) */
PFD_HD float pfd_grad_dot(uint32_t h, float dx, float dy) {
    switch (h & 7u) {
    case 0: return PFD_FADD(dx, dy);
    case 1: return PFD_FSUB(dy, dx);
    case 2: return PFD_FSUB(dx, dy);
    case 3: return PFD_FSUB(0.0f, PFD_FADD(dx, dy));
    case 4: return dx;
    case 5: return PFD_FSUB(0.0f, dx);
    case 6: return dy;
    default: return PFD_FSUB(0.0f, dy);
    }
}

PFD_HD float pfd_fade(float t) {
    /* 6t^5 - 15t^4 + 10t^3 = t*t*t*(t*(t*6-15)+10) */
    float a = PFD_FSUB(PFD_FMUL(t, 6.0f), 15.0f);
    float b = PFD_FADD(PFD_FMUL(t, a), 10.0f);
    return PFD_FMUL(PFD_FMUL(PFD_FMUL(t, t), t), b);
}

PFD_HD float pfd_lerp(float a, float b, float t) {
    return PFD_FADD(a, PFD_FMUL(t, PFD_FSUB(b, a)));
}

/* one octave of gradient noise; cs = lattice cell size in pixels (power of two >= 2) */
PFD_HD float pfd_perlin(int64_t r, int64_t c, int64_t cs, uint32_t seed) {
    int64_t iy = r / cs, ix = c / cs;
    float inv = 1.0f / (float)cs; /* cs is a power of two: exact */
    float fy = PFD_FMUL((float)(r - iy * cs), inv);
    float fx = PFD_FMUL((float)(c - ix * cs), inv);
    uint32_t x0 = (uint32_t)ix, y0 = (uint32_t)iy;
    float n00 = pfd_grad_dot(pfd_hash3(x0, y0, seed), fx, fy);
    float n10 = pfd_grad_dot(pfd_hash3(x0 + 1u, y0, seed), PFD_FSUB(fx, 1.0f), fy);
    float n01 = pfd_grad_dot(pfd_hash3(x0, y0 + 1u, seed), fx, PFD_FSUB(fy, 1.0f));
    float n11 = pfd_grad_dot(pfd_hash3(x0 + 1u, y0 + 1u, seed), PFD_FSUB(fx, 1.0f), PFD_FSUB(fy, 1.0f));
    float u = pfd_fade(fx), v = pfd_fade(fy);
    return pfd_lerp(pfd_lerp(n00, n10, u), pfd_lerp(n01, n11, u), v);
}

/*
 * Elevation at (r, c). `n` = lattice reference size in pixels (use max(nrow, ncol) rounded up to a
 * power of two), `octaves` = number of octaves (octave o has 4*2^o lattice cells per n pixels; the
 * cell size is clamped at 2 px), seed as in SURVEY.md §8(d).
 */
PFD_HD float pfd_synth_z(int64_t r, int64_t c, int64_t n, int octaves, uint32_t seed) {
    float z = 0.0f;
    float amp = 1.0f;
    int64_t cs = n / 4;
    for (int o = 0; o < octaves; ++o) {
        if (cs < 2) cs = 2;
        z = PFD_FADD(z, PFD_FMUL(amp, pfd_perlin(r, c, cs, seed * 100u + (uint32_t)o)));
        amp = PFD_FMUL(amp, 0.5f);
        cs = cs / 2;
    }
    /* tie-breaking white noise: 1e-4 * U[0,1) with 24 random bits */
    uint32_t h = pfd_hash3((uint32_t)c, (uint32_t)r, seed + 999u);
    float u = PFD_FMUL((float)(h >> 8), 5.9604644775390625e-08f); /* 2^-24 */
    return PFD_FADD(z, PFD_FMUL(u, 1.0e-4f));
}

/*
 * D8 code of the centre cell of a 3x3 elevation window z[9] (row-major, z[4] = centre).
 * valid[k] = 0 marks off-raster neighbours. sea: cells with z < sea_level are nodata (247);
 * pass sea_level = -INFINITY for none.
 */
PFD_HD uint8_t pfd_synth_d8_from_window(const float* z, const int* valid, float sea_level) {
    const uint8_t code[9] = {32, 64, 128, 16, 0, 1, 8, 4, 2};
    float z0 = z[4];
    if (z0 < sea_level) return 247;
    float best = 0.0f;
    uint8_t d8 = 0;
    for (int k = 0; k < 9; ++k) {
        if (k == 4 || !valid[k]) continue;
        float drop = PFD_FSUB(z0, z[k]);
        if ((k & 1) == 0) drop = PFD_FMUL(drop, 0.70710678118654752f); /* diagonal */
        if (drop > best) {
            best = drop;
            d8 = code[k];
        }
    }
    return d8;
}

#endif /* PFD_SYNTH_H */
