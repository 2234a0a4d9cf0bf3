// pfd_fill.cuh -- dem.fill_depressions (pyflwdir/dem.py:17-143, Wang & Liu 2006 priority flood) on the device, bit for bit:
// the depression-filled elevation AND the D8 raster the reference derives while it floods (the producer of
// the hot path's input; pyflwdir.from_dem, pyflwdir/pyflwdir.py:51-102, is this + from_array).
//
// What the reference's heap loop computes, restated without the heap. Every valid cell gets a LEVEL = the float32 key it is
// pushed with. A popped cell (key z0) visits its not-yet-done neighbours (and itself, for an initial outlet): the neighbour
// is raised to z0 if it lies lower (delv = z0 - z1), pushed with key float32(z1 + delv), marked done, and its D8 code points
// at the popped cell. So   parent(x) = the neighbour of x that is popped FIRST,   level(x) = max(float32(z_x), level(parent)).
// Keys never decrease along a parent chain, hence cells pop in ascending level, and
//   (1) level(x) = min over paths from x to an outlet of the largest float32(z) on the path (outlets: their own key) --
//       a pure min/max fixpoint, no arithmetic: solved by tile-wise chaotic relaxation in shared memory (fd_relax_kernel,
//       monotone, so any update order reaches the same fixpoint), tiles re-activated by their neighbours until nothing moves;
//   (2) parent(x) = the neighbour with the smallest level, when that neighbour is unique;
//   (3) among cells of EQUAL level the heap order decides: it pops the smallest (boundary, row, col) among the cells queued
//       so far -- a priority-first search that is sequential by nature, but confined to a connected set of equal-level
//       cells (a filled lake, a flat). Such "tie components" (equal-level neighbours, plus cells that compete as
//       smallest-level neighbours of a common cell) are labelled by min-index propagation (fd_label_kernel) and each one is
//       replayed by ONE thread with a binary heap over exactly the reference's keys (fd_simulate_kernel); all components of
//       all levels run concurrently because the set of cells queued before a level starts is known from (1).
// Raising a cell is ARITHMETIC in the raster's own type (z1 += z0 - z1; float32 rasters: float32, float64 / integer rasters:
// float64) and the pushed key is float32(z1): for float32 rasters the result can miss z0 by an ulp or, when z0 and z1 differ
// in sign or magnitude, by much more, so the keys inside a filled lake DRIFT around the pour level and so does the pop order.
// The replay therefore works with the actual keys and the actual arithmetic, and "equal level" in (3) means "within BAND of each
// other" (an absolute elevation difference; 0 = exact ties at first, 16 ulps of max |z| or more once a raise has drifted):
// everything the heap might order differently from the drift-free levels of (1) is inside one component and is replayed
// exactly; across components (and for cells in none) levels differ by more than BAND and the order follows (1). The replay
// checks that no key drifted by more than BAND / 4 from its level (1); if one did, the labelling and the replay are repeated with
// a wider band (x 32 or 8 x the drift seen; in the limit the whole raster is one component, i.e. the reference's own serial
// loop). Float64 / integer rasters reproduce z0 exactly and never leave BAND = 0.
// Why that is sound. Let every key lie within d of its cell's level (checked: d <= BAND / 4) and let u, v be cells whose levels
// differ by more than BAND > 2d, level(u) < level(v). Then the heap pops u before v: walk from u along smallest-level neighbours to
// an outlet -- levels never rise on that walk, so every cell on it has a key below key(v); the first cell of the walk (from the
// outlet side) that is not yet popped when v pops is already in the heap (an outlet, or visited when its predecessor popped)
// with a smaller key than v, which contradicts v being the minimum. So a cell's parent is among its neighbours within BAND of
// the smallest level; two or more of those share a component by construction ("connectors"), a single one needs no order. A cell
// visited from outside its component is visited by a neighbour more than BAND below it, hence never raised, hence its key is
// its own float32 elevation whatever the drift elsewhere: the set queued before a component starts, and their keys, are exact.
// max_depth >= 0 (re-opening of visited cells, dem.py:121-132) is not restated: PFD_ERR_UNSUPPORTED.
#pragma once
#include "pfd_common.cuh"
#include <chrono>

#define FD_T 64            // tile of the level relaxation
#define FD_TS (FD_T + 2)
#ifndef FD_SWEEPS
#define FD_SWEEPS 2         // in-place sweeps of the level relaxation between two barriers (one down, one up; 2 / 3 / 4 / 8 / 16: 12.0 / 12.3 / 12.9 / 15.0 / 19.1 ms at 8192^2)
#endif
#define FL_T 32            // tile of the label propagation
#define FL_TS (FL_T + 2)
#define FD_UNREACHED 0xFFFFFFFFu
#define FD_NONE 0xFFFFFFFFu
enum { FDF_VALID = 1, FDF_OUTLET = 2, FDF_TIED = 4, FDF_QUEUED = 8, FDF_DISC = 16, FDF_SRCDISC = 32 };
#define FDF_STATIC (FDF_VALID | FDF_OUTLET)  // what survives a retry with a wider band

// monotone map float32 -> uint32 (-0.0 and +0.0 share a key: the reference's tuple comparison treats them as equal)
__device__ __forceinline__ uint32_t fd_ord(float f) {
    const uint32_t u = __float_as_uint(f + 0.0f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float fd_unord(uint32_t o) { return __uint_as_float((o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o); }

// neighbour k = 3 * (dr + 1) + (dc + 1) of the 3x3 structure (dem.py:75-78,110-111); core_d8._us (core_d8.py:16)
__device__ __constant__ uint8_t fd_us[9] = {2, 4, 8, 1, 0, 16, 128, 64, 32};
#define FD_NB8 0x1EFu  // all but the centre
#define FD_NB4 0x0AAu  // N, W, E, S

template <typename T>
__device__ __forceinline__ bool fd_is_nodata(T z, double nodata, int nodata_nan) {
    return nodata_nan ? (z != z) : ((double)z == nodata);
}

__device__ __forceinline__ void fd_mark_neighbours(uint8_t* active_next, int64_t ty, int64_t tx, int64_t nty, int64_t ntx, int dy, int dx) {
    if (dy != 0 && ty + dy >= 0 && ty + dy < nty) active_next[(ty + dy) * ntx + tx] = 1;
    if (dx != 0 && tx + dx >= 0 && tx + dx < ntx) active_next[ty * ntx + tx + dx] = 1;
    if (dy != 0 && dx != 0 && ty + dy >= 0 && ty + dy < nty && tx + dx >= 0 && tx + dx < ntx) active_next[(ty + dy) * ntx + tx + dx] = 1;
}

struct FdCounters {
    unsigned long long n_outlets, minkey, n_tied, n_roots, pool_top, root_fill, err_pit, n_drift, n_unreached, max_drift, max_comp, max_abs, big_fill;
};

// dem.py:70-71,81-86 + gis_utils.get_edge (gis_utils.py:118-144): validity, initial outlets, their keys
template <typename T>
__global__ void fd_init_kernel(const T* __restrict__ elev, int64_t nrow, int64_t ncol, double nodata, int nodata_nan, uint32_t nbmask,
                               int mode, int has_elv_max, double elv_max, uint8_t* __restrict__ flags, uint32_t* __restrict__ S,
                               uint8_t* __restrict__ act0, int64_t ntx, FdCounters* cnt) {
    const int64_t n = nrow * ncol;
    float zmax = 0.0f;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const T z = elev[i];
        uint8_t f = 0;
        uint32_t s = FD_UNREACHED;
        if (!fd_is_nodata(z, nodata, nodata_nan)) {
            f = FDF_VALID;
            if (mode != 2) {
                const int64_t r = i / ncol, c = i - r * ncol;
                bool edge = r == 0 || r == nrow - 1 || c == 0 || c == ncol - 1;
                if (!edge) {
#pragma unroll
                    for (int k = 0; k < 9; ++k)
                        if (((nbmask >> k) & 1u) && fd_is_nodata(elev[i + (k / 3 - 1) * ncol + (k % 3 - 1)], nodata, nodata_nan)) edge = true;
                }
                if (edge && (!has_elv_max || (double)z <= elv_max)) {
                    f |= FDF_OUTLET;
                    s = fd_ord((float)z);
                    // the flood starts in the tiles that hold an outlet (and next door, when the outlet sits on a tile edge)
                    const int64_t ty = r / FD_T, tx = c / FD_T;
                    const int ly = (int)(r - ty * FD_T), lx = (int)(c - tx * FD_T);
                    act0[ty * ntx + tx] = 1;
                    fd_mark_neighbours(act0, ty, tx, (nrow + FD_T - 1) / FD_T, ntx, ly == 0 ? -1 : (ly == FD_T - 1 ? 1 : 0),
                                       lx == 0 ? -1 : (lx == FD_T - 1 ? 1 : 0));
                    atomicAdd(&cnt->n_outlets, 1ull);
                    if (mode == 1) atomicMin(&cnt->minkey, ((unsigned long long)s << 32) | (unsigned long long)i);
                }
            }
            const float az = fabsf((float)z);
            if (az > zmax && az < 3.0e38f) zmax = az;
        }
        flags[i] = f;
        S[i] = s;
    }
    const uint32_t zb = __reduce_max_sync(0xFFFFFFFFu, __float_as_uint(zmax));  // (non-negative floats order like their bits)
    if ((threadIdx.x & 31) == 0 && zb) atomicMax(&cnt->max_abs, (unsigned long long)zb);
}

// outlets="min" (dem.py:104-107): only the smallest (key, 1, row, col) stays an outlet
__global__ void fd_keep_min_kernel(int64_t n, uint8_t* __restrict__ flags, uint32_t* __restrict__ S, const FdCounters* cnt) {
    const int64_t keep = (int64_t)(cnt->minkey & 0xFFFFFFFFull);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if ((flags[i] & FDF_OUTLET) && i != keep) {
            flags[i] &= (uint8_t)~FDF_OUTLET;
            S[i] = FD_UNREACHED;
        }
}

// idxs_pit given (dem.py:87-90)
template <typename T>
__global__ void fd_pits_kernel(const T* __restrict__ elev, int64_t n, const int64_t* __restrict__ idxs, int64_t npit,
                               uint8_t* __restrict__ flags, uint32_t* __restrict__ S, uint8_t* __restrict__ act0, int64_t ncol, int64_t ntx,
                               FdCounters* cnt) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < npit; k += (int64_t)gridDim.x * blockDim.x) {
        int64_t i = idxs[k];
        if (i < 0) i += n;
        if (i < 0 || i >= n || !(flags[i] & FDF_VALID)) {
            atomicAdd(&cnt->err_pit, 1ull);
            continue;
        }
        flags[i] |= FDF_OUTLET;  // duplicates write the same values
        S[i] = fd_ord((float)elev[i]);
        {
            const int64_t r = i / ncol, c = i - r * ncol, ty = r / FD_T, tx = c / FD_T;
            const int ly = (int)(r - ty * FD_T), lx = (int)(c - tx * FD_T);
            act0[ty * ntx + tx] = 1;
            fd_mark_neighbours(act0, ty, tx, (n / ncol + FD_T - 1) / FD_T, ntx, ly == 0 ? -1 : (ly == FD_T - 1 ? 1 : 0),
                               lx == 0 ? -1 : (lx == FD_T - 1 ? 1 : 0));
        }
        atomicAdd(&cnt->n_outlets, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------------------
// (1) levels: S(x) <- max(key(x), min over neighbours S(y)) until nothing moves, one CTA per active 64x64 tile
// ---------------------------------------------------------------------------------------------------------


template <typename T>
__global__ void __launch_bounds__(1024) fd_relax_kernel(const T* __restrict__ elev, const uint8_t* __restrict__ flags, uint32_t* __restrict__ S,
                                                        int64_t nrow, int64_t ncol, int64_t nty, int64_t ntx, uint32_t nbmask,
                                                        const uint8_t* __restrict__ active_cur, uint8_t* __restrict__ active_next,
                                                        unsigned int* __restrict__ changed) {
    const int64_t tile = blockIdx.x;
    if (!active_cur[tile]) return;
    const int64_t ty = tile / ntx, tx = tile - ty * ntx;
    const int64_t r0 = ty * FD_T, c0 = tx * FD_T;
    __shared__ uint32_t sS[FD_TS * FD_TS];
    for (int idx = threadIdx.x; idx < FD_TS * FD_TS; idx += 1024) {
        const int sy = idx / FD_TS, sx = idx - sy * FD_TS;
        const int64_t r = r0 + sy - 1, c = c0 + sx - 1;
        sS[idx] = (r >= 0 && r < nrow && c >= 0 && c < ncol) ? S[r * ncol + c] : FD_UNREACHED;
    }
    uint32_t kz[4], orig[4];
    bool upd[4];
    int pos[4];
    // a thread owns four vertically adjacent cells and relaxes them in place top-down, then bottom-up: a level travels four rows per
    // sweep along a column instead of one; several sweeps run between two barriers (any interleaving reaches the same fixpoint)
    const int lx = threadIdx.x & 63, ly0 = (threadIdx.x >> 6) << 2;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int ly = ly0 + j;
        const int64_t r = r0 + ly, c = c0 + lx;
        pos[j] = (ly + 1) * FD_TS + lx + 1;
        upd[j] = false;
        kz[j] = 0;
        if (r < nrow && c < ncol) {
            const uint8_t f = flags[r * ncol + c];
            upd[j] = (f & FDF_VALID) && !(f & FDF_OUTLET);
            if (upd[j]) kz[j] = fd_ord((float)elev[r * ncol + c]);
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) orig[j] = sS[pos[j]];
    volatile uint32_t* vS = sS;
    for (;;) {
        bool ch = false;
#pragma unroll 1
        for (int sweep = 0; sweep < FD_SWEEPS; ++sweep) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const int j = (sweep & 1) ? 3 - jj : jj;
                if (!upd[j]) continue;
                uint32_t m = FD_UNREACHED;
#pragma unroll
                for (int k = 0; k < 9; ++k)
                    if ((nbmask >> k) & 1u) m = min(m, vS[pos[j] + (k / 3 - 1) * FD_TS + (k % 3 - 1)]);
                const uint32_t ns = max(kz[j], m);
                if (ns < vS[pos[j]]) {
                    vS[pos[j]] = ns;
                    ch = true;
                }
            }
        }
        if (!__syncthreads_or((int)ch)) break;
    }
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t v = sS[pos[j]];
        if (v != orig[j]) {
            const int ly = ly0 + j;
            S[(r0 + ly) * ncol + c0 + lx] = v;
            any = true;
            const int dy = (ly == 0) ? -1 : ((ly == FD_T - 1) ? 1 : 0), dx = (lx == 0) ? -1 : ((lx == FD_T - 1) ? 1 : 0);
            if (dy != 0 || dx != 0) fd_mark_neighbours(active_next, ty, tx, nty, ntx, dy, dx);
        }
    }
    if (any) *changed = 1u;
}

// ---------------------------------------------------------------------------------------------------------
// (3a) tie components. Levels within `band` steps of each other count as tied. label[x] = x for cells with a tied neighbour;
// link[x] != NONE marks a "connector": a cell whose smallest-level neighbours (within band of the smallest) are two or more
// -- their pop order decides its parent, so they must share a heap.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool fd_near(uint32_t a, uint32_t b, float band) { return a == b || fabsf(fd_unord(a) - fd_unord(b)) <= band; }
// b is one of the smallest levels around a cell whose smallest neighbour level is m (b >= m)
__device__ __forceinline__ bool fd_nearmin(uint32_t b, uint32_t m, float band) { return b == m || fd_unord(b) - fd_unord(m) <= band; }

__global__ void fd_tie_kernel(const uint32_t* __restrict__ S, int64_t nrow, int64_t ncol, uint32_t nbmask, float band,
                              uint32_t* __restrict__ M, uint32_t* __restrict__ label, uint32_t* __restrict__ link, uint8_t* __restrict__ act0,
                              int64_t ltx) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t s = S[i];
        uint32_t m = FD_UNREACHED, lab = FD_NONE, lk = FD_NONE;
        if (s != FD_UNREACHED) {
            const int64_t r = i / ncol, c = i - r * ncol;
            uint32_t sn[9];
            bool eq = false;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                sn[k] = FD_UNREACHED;
                if (!((nbmask >> k) & 1u)) continue;
                const int64_t rr = r + k / 3 - 1, cc = c + k % 3 - 1;
                if (rr < 0 || rr >= nrow || cc < 0 || cc >= ncol) continue;
                sn[k] = S[rr * ncol + cc];
                if (sn[k] == FD_UNREACHED) continue;
                eq |= fd_near(sn[k], s, band);
                m = min(m, sn[k]);
            }
            int nmin = 0;
            uint32_t first = FD_NONE;
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (sn[k] != FD_UNREACHED && fd_nearmin(sn[k], m, band)) {
                    if (nmin++ == 0) first = (uint32_t)((r + k / 3 - 1) * ncol + (c + k % 3 - 1));
                }
            if (eq) lab = (uint32_t)i;
            if (nmin >= 2) lk = first;  // smallest index among the candidates (ascending k = ascending index)
        }
        M[i] = m;
        label[i] = lab;
        link[i] = lk;
        if (lab != FD_NONE || lk != FD_NONE) {  // label propagation starts in the tiles that hold a tied cell or a connector
            const int64_t r = i / ncol, c = i - r * ncol;
            act0[(r / FL_T) * ltx + c / FL_T] = 1;
        }
    }
}

// the candidates of a connector join the labelled cells (even when they have no tied neighbour themselves)
__global__ void fd_tie2_kernel(const uint32_t* __restrict__ S, const uint32_t* __restrict__ M, const uint32_t* __restrict__ link,
                               int64_t nrow, int64_t ncol, uint32_t nbmask, float band, uint32_t* __restrict__ label, uint8_t* __restrict__ act0,
                               int64_t ltx) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (link[i] == FD_NONE) continue;
        const uint32_t m = M[i];
        const int64_t r = i / ncol, c = i - r * ncol;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!((nbmask >> k) & 1u)) continue;
            const int64_t rr = r + k / 3 - 1, cc = c + k % 3 - 1;
            if (rr < 0 || rr >= nrow || cc < 0 || cc >= ncol) continue;
            const int64_t y = rr * ncol + cc;
            const uint32_t sy = S[y];
            if (sy != FD_UNREACHED && fd_nearmin(sy, m, band) && label[y] == FD_NONE) {
                label[y] = (uint32_t)y;  // racing writers store the same value
                act0[(rr / FL_T) * ltx + cc / FL_T] = 1;
            }
        }
    }
}

// min-label propagation, one CTA (256 threads, 4 cells each) per active 32x32 tile:
//   labelled y : label <- min(label of tied labelled neighbours, link of adjacent connectors whose candidate y is)
//   connector x: link  <- min(label of its candidates)
__global__ void __launch_bounds__(256) fd_label_kernel(const uint32_t* __restrict__ S, const uint32_t* __restrict__ M, uint32_t* __restrict__ label,
                                                       uint32_t* __restrict__ link, int64_t nrow, int64_t ncol, int64_t nty, int64_t ntx,
                                                       uint32_t nbmask, float band, const uint8_t* __restrict__ active_cur,
                                                       uint8_t* __restrict__ active_next, unsigned int* __restrict__ changed) {
    const int64_t tile = blockIdx.x;
    if (!active_cur[tile]) return;
    const int64_t ty = tile / ntx, tx = tile - ty * ntx;
    const int64_t r0 = ty * FL_T, c0 = tx * FL_T;
    __shared__ uint32_t sS[FL_TS * FL_TS], sM[FL_TS * FL_TS], sLab[FL_TS * FL_TS], sLnk[FL_TS * FL_TS];
    for (int idx = threadIdx.x; idx < FL_TS * FL_TS; idx += 256) {
        const int sy = idx / FL_TS, sx = idx - sy * FL_TS;
        const int64_t r = r0 + sy - 1, c = c0 + sx - 1;
        const bool in = r >= 0 && r < nrow && c >= 0 && c < ncol;
        const int64_t g = r * ncol + c;
        sS[idx] = in ? S[g] : FD_UNREACHED;
        sM[idx] = in ? M[g] : FD_UNREACHED;
        sLab[idx] = in ? label[g] : FD_NONE;
        sLnk[idx] = in ? link[g] : FD_NONE;
    }
    int pos[4];
    uint32_t lab0[4], lnk0[4];
    const int lx = threadIdx.x & 31, ly0 = threadIdx.x >> 5;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        pos[j] = (ly0 + 8 * j + 1) * FL_TS + lx + 1;
        lab0[j] = sLab[pos[j]];
        lnk0[j] = sLnk[pos[j]];
    }
    volatile uint32_t* vLab = sLab;
    volatile uint32_t* vLnk = sLnk;
    for (;;) {
        bool ch = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int p = pos[j];
            const uint32_t s = sS[p], m = sM[p];
            const uint32_t lab = vLab[p], lnk = vLnk[p];
            if (lab == FD_NONE && lnk == FD_NONE) continue;
            uint32_t bl = lab, bk = lnk;
#pragma unroll
            for (int k = 0; k < 9; ++k) {
                if (!((nbmask >> k) & 1u)) continue;
                const int q = p + (k / 3 - 1) * FL_TS + (k % 3 - 1);
                const uint32_t sy = sS[q];
                if (sy == FD_UNREACHED) continue;
                if (lab != FD_NONE) {
                    if (fd_near(sy, s, band)) bl = min(bl, vLab[q]);             // FD_NONE is the largest value
                    if (fd_nearmin(s, sM[q], band)) bl = min(bl, vLnk[q]);       // this cell is a candidate of q (if q is a connector)
                }
                if (lnk != FD_NONE && fd_nearmin(sy, m, band)) bk = min(bk, vLab[q]);
            }
            if (bl < lab) {
                vLab[p] = bl;
                ch = true;
            }
            if (bk < lnk) {
                vLnk[p] = bk;
                ch = true;
            }
        }
        if (!__syncthreads_or((int)ch)) break;
    }
    bool any = false;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t lab = sLab[pos[j]], lnk = sLnk[pos[j]];
        if (lab != lab0[j] || lnk != lnk0[j]) {
            const int ly = ly0 + 8 * j;
            const int64_t g = (r0 + ly) * ncol + c0 + lx;
            if (lab != lab0[j]) label[g] = lab;
            if (lnk != lnk0[j]) link[g] = lnk;
            any = true;
            const int dy = (ly == 0) ? -1 : ((ly == FL_T - 1) ? 1 : 0), dx = (lx == 0) ? -1 : ((lx == FL_T - 1) ? 1 : 0);
            if (dy != 0 || dx != 0) fd_mark_neighbours(active_next, ty, tx, nty, ntx, dy, dx);
        }
    }
    if (any) *changed = 1u;
}

// component sizes (cnt[root]) and the number of labelled cells / components
__global__ void fd_count_kernel(const uint32_t* __restrict__ label, int64_t n, uint32_t* __restrict__ cntarr, FdCounters* cnt) {
    unsigned long long tied = 0, roots = 0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t l = label[i];
        if (l == FD_NONE) continue;
        atomicAdd(&cntarr[l], 1u);
        ++tied;
        roots += l == (uint32_t)i;
    }
    if (tied) atomicAdd(&cnt->n_tied, tied);
    if (roots) atomicAdd(&cnt->n_roots, roots);
}

#ifndef FD_BIG
#define FD_BIG 2048     // components of at least this many cells are replayed by a warp (fd_simulate_warp_kernel)
#endif
#ifndef FDW_CAP
#define FDW_CAP 24576   // heap entries that fit the warp's shared memory (192 KiB); tests build with a tiny value to exercise the spill
#endif
// every component gets a slice of the heap pool; cnt[root] becomes the fill counter of that slice
__global__ void fd_roots_kernel(const uint32_t* __restrict__ label, int64_t n, uint32_t* __restrict__ cntarr, uint32_t* __restrict__ off,
                                uint32_t* __restrict__ roots, unsigned long long nroots, FdCounters* cnt) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        if (label[i] != (uint32_t)i) continue;
        if (cntarr[i] >= FD_BIG) roots[nroots - 1 - atomicAdd(&cnt->big_fill, 1ull)] = (uint32_t)i;  // big ones from the back
        else roots[atomicAdd(&cnt->root_fill, 1ull)] = (uint32_t)i;
        off[i] = (uint32_t)atomicAdd(&cnt->pool_top, (unsigned long long)cntarr[i]);
        atomicMax(&cnt->max_comp, (unsigned long long)cntarr[i]);
        cntarr[i] = 0;
    }
}

// cells that are in the heap before their component starts to pop: initial outlets (boundary = 1) and cells visited by a
// lower neighbour OUTSIDE the component (boundary = 0, already done; never raised: that neighbour lies more than `band`
// below). Heap entry = float32 key (ordered) << 32 | boundary << 31 | linear index  ==  the tuple (key, boundary, row, col).
__global__ void fd_sources_kernel(const uint32_t* __restrict__ S, const uint32_t* __restrict__ label, int64_t nrow, int64_t ncol, uint32_t nbmask,
                                  uint8_t* __restrict__ flags, uint32_t* __restrict__ cntarr, const uint32_t* __restrict__ off,
                                  unsigned long long* __restrict__ pool) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t root = label[i];
        if (root == FD_NONE) continue;
        const uint32_t s = S[i];
        const int64_t r = i / ncol, c = i - r * ncol;
        uint32_t m = FD_UNREACHED, mlab = FD_NONE;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            if (!((nbmask >> k) & 1u)) continue;
            const int64_t rr = r + k / 3 - 1, cc = c + k % 3 - 1;
            if (rr < 0 || rr >= nrow || cc < 0 || cc >= ncol) continue;
            const uint32_t sy = S[rr * ncol + cc];
            if (sy < m) {
                m = sy;
                mlab = label[rr * ncol + cc];
            }
        }
        uint8_t f = (flags[i] & FDF_STATIC) | FDF_TIED;
        const bool lower_outside = m < s && mlab != root;
        if (lower_outside) f |= FDF_DISC | FDF_SRCDISC;
        if (lower_outside || (f & FDF_OUTLET)) {
            f |= FDF_QUEUED;
            pool[off[root] + atomicAdd(&cntarr[root], 1u)] =
                ((unsigned long long)s << 32) | ((f & FDF_OUTLET) ? 0x80000000ull : 0ull) | (unsigned long long)i;
        }
        flags[i] = f;
    }
}

// linear index -> row without a hardware division (a dependent ~20-instruction sequence on the replay's critical path):
// row = (i * magic) >> (31 + shift) is exact for i < 2^31 with magic = ceil(2^(31 + shift) / ncol), shift = ceil(log2(ncol))
struct FdDiv {
    uint32_t magic, shift;
};
static FdDiv fd_make_div(uint32_t ncol) {
    uint32_t s = 0;
    while ((1ull << s) < ncol) ++s;
    const unsigned long long num = 1ull << (31 + s);
    return FdDiv{(uint32_t)((num + ncol - 1) / ncol), 31 + s};
}
__device__ __forceinline__ uint32_t fd_row_of(uint32_t i, FdDiv dv) { return (uint32_t)(((unsigned long long)i * dv.magic) >> dv.shift); }

// (3b) one thread replays the heap loop (dem.py:112-142) of one tie component with the reference's keys and arithmetic
__device__ __forceinline__ void fd_sift_down(unsigned long long* hp, uint32_t n, uint32_t i, unsigned long long v) {
    for (;;) {
        uint32_t l = 2 * i + 1;
        if (l >= n) break;
        if (l + 1 < n && hp[l + 1] < hp[l]) ++l;
        if (hp[l] >= v) break;
        hp[i] = hp[l];
        i = l;
    }
    hp[i] = v;
}

// visit of cell y by a popped cell with key z0 (dem.py:119-120,133-135): returns the key y is pushed with, writes elevtn + delv
template <typename T, typename W>
__device__ __forceinline__ float fd_visit_z(T zy, T* __restrict__ out, uint32_t y, float z0, int int_delv);

template <typename T, typename W>
__device__ __forceinline__ float fd_visit(const T* __restrict__ elev, T* __restrict__ out, uint32_t y, float z0, int int_delv) {
    return fd_visit_z<T, W>(elev[y], out, y, z0, int_delv);
}

template <typename T, typename W>
__device__ __forceinline__ float fd_visit_z(T zy, T* __restrict__ out, uint32_t y, float z0, int int_delv) {
    const W z1 = (W)zy;
    const W dz = (W)z0 - z1;
    W delv = (W)0, z1n = z1;
    if (dz > (W)0) {
        delv = int_delv ? (W)(long long)dz : dz;
        z1n = z1 + dz;
    }
    out[y] = (T)(z1 + delv);
    return (float)z1n;
}

template <typename T, typename W>
__global__ void fd_simulate_kernel(const uint32_t* __restrict__ roots, uint32_t nroots, const uint32_t* __restrict__ off,
                                   const uint32_t* __restrict__ cntarr, unsigned long long* __restrict__ pool, const uint32_t* __restrict__ S,
                                   const uint32_t* __restrict__ label, const T* __restrict__ elev, uint8_t* flags, uint32_t* __restrict__ Tord,
                                   uint8_t* d8, T* __restrict__ out, int64_t nrow, int64_t ncol, uint32_t nbmask, int int_delv,
                                   float max_drift, FdDiv dv, FdCounters* cnt) {
    const uint32_t k0 = blockIdx.x * blockDim.x + threadIdx.x;
    if (k0 >= nroots) return;
    const uint32_t root = roots[k0];
    unsigned long long* hp = pool + off[root];
    uint32_t n = cntarr[root];
    const int32_t nr = (int32_t)nrow, nc = (int32_t)ncol;
    for (int32_t i = (int32_t)(n / 2) - 1; i >= 0; --i) fd_sift_down(hp, n, (uint32_t)i, hp[i]);
    uint32_t t = 0;
    float worst = 0.0f;
    unsigned long long ndrift = 0;
    while (n > 0) {
        const unsigned long long top = hp[0];
        --n;
        if (n > 0) fd_sift_down(hp, n, 0, hp[n]);
        const uint32_t p = (uint32_t)(top & 0x7FFFFFFFull);
        const float z0 = fd_unord((uint32_t)(top >> 32));
        Tord[p] = t++;
        const uint8_t fp = flags[p];
        if (!(fp & FDF_DISC)) {  // an initial outlet that no neighbour visited before it popped: it visits itself (d8 = 0)
            flags[p] = fp | FDF_DISC;
            d8[p] = 0;
            fd_visit<T, W>(elev, out, p, z0, int_delv);
        }
        const int32_t pr = (int32_t)fd_row_of(p, dv), pc = (int32_t)(p - (uint32_t)pr * (uint32_t)nc);
        for (int k = 0; k < 9; ++k) {
            if (!((nbmask >> k) & 1u)) continue;
            const int32_t rr = pr + k / 3 - 1, cc = pc + k % 3 - 1;
            if (rr < 0 || rr >= nr || cc < 0 || cc >= nc) continue;
            const uint32_t y = (uint32_t)rr * (uint32_t)nc + (uint32_t)cc;
            if (label[y] != root) continue;
            uint8_t fy = flags[y];
            if (fy & FDF_DISC) continue;
            fy |= FDF_DISC;
            d8[y] = fd_us[k];
            const float kf = fd_visit<T, W>(elev, out, y, z0, int_delv);
            const uint32_t ky = fd_ord(kf), sy = S[y];
            const float drift = ky == sy ? 0.0f : fabsf(kf - fd_unord(sy));
            worst = fmaxf(worst, drift);
            ndrift += !(drift <= max_drift);
            if (!(fy & FDF_QUEUED)) {
                fy |= FDF_QUEUED;
                const unsigned long long e = ((unsigned long long)ky << 32) | (unsigned long long)y;
                uint32_t i = n++;  // sift up
                while (i > 0) {
                    const uint32_t par = (i - 1) / 2;
                    if (hp[par] <= e) break;
                    hp[i] = hp[par];
                    i = par;
                }
                hp[i] = e;
            }
            flags[y] = fy;
        }
    }
    if (worst > 0.0f) atomicMax(&cnt->max_drift, (unsigned long long)__float_as_uint(worst));  // non-negative floats order like their bits
    if (ndrift) atomicAdd(&cnt->n_drift, ndrift);
}

// The same replay for a BIG component, by one warp: the heap (32-ary) lives in shared memory while it fits -- a pop is a chain of
// dependent reads, ~17 levels of global-memory latency per pop in the one-thread version --, the 32 children of a node are read by
// the 32 lanes at once, and the eight neighbours of the popped cell are visited by eight lanes at once.
__device__ __forceinline__ unsigned long long fd_shfl64(unsigned long long v, int src) { return __shfl_sync(0xFFFFFFFFu, v, src); }

#define FDW_ARY 32  // heap fan-out = warp width: one shared-memory read + two warp reductions (REDUX) per level, 3 levels for 24 k entries
__device__ __forceinline__ void fd_wsift_down(unsigned long long* hp, uint32_t n, uint32_t i, unsigned long long v, int lane) {
    for (;;) {
        const uint32_t c0 = FDW_ARY * i + 1;
        if (c0 >= n) break;
        const unsigned long long cv = (c0 + lane < n) ? hp[c0 + lane] : ~0ull;
        const uint32_t hi = (uint32_t)(cv >> 32), lo = (uint32_t)cv;
        const uint32_t mh = __reduce_min_sync(0xFFFFFFFFu, hi);
        const uint32_t ml = __reduce_min_sync(0xFFFFFFFFu, hi == mh ? lo : 0xFFFFFFFFu);
        const unsigned long long m = ((unsigned long long)mh << 32) | ml;
        if (m >= v) break;
        const uint32_t which = __ffs(__ballot_sync(0xFFFFFFFFu, cv == m)) - 1;  // entries are distinct
        if (lane == 0) hp[i] = m;
        i = c0 + which;
    }
    if (lane == 0) hp[i] = v;
    __syncwarp();
}

template <typename T, typename W>
__global__ void __launch_bounds__(32) fd_simulate_warp_kernel(const uint32_t* __restrict__ roots, const uint32_t* __restrict__ off,
                                                              const uint32_t* __restrict__ cntarr, unsigned long long* __restrict__ pool,
                                                              const uint32_t* __restrict__ S, const uint32_t* __restrict__ label,
                                                              const T* __restrict__ elev, uint8_t* flags, uint32_t* __restrict__ Tord,
                                                              uint8_t* d8, T* __restrict__ out, int64_t nrow, int64_t ncol, uint32_t nbmask,
                                                              int int_delv, float max_drift, FdDiv dv, FdCounters* cnt) {
    extern __shared__ unsigned long long fd_sheap[];
    const int lane = threadIdx.x;
    const uint32_t root = roots[blockIdx.x];
    unsigned long long* gheap = pool + off[root];
    uint32_t n = cntarr[root];
    unsigned long long* hp = gheap;
    bool in_smem = n <= FDW_CAP;
    if (in_smem) {
        for (uint32_t i = lane; i < n; i += 32) fd_sheap[i] = gheap[i];
        hp = fd_sheap;
    }
    __syncwarp();
    if (n > 1)
        for (int32_t i = (int32_t)((n - 2) / FDW_ARY); i >= 0; --i) fd_wsift_down(hp, n, (uint32_t)i, hp[i], lane);
    const int32_t nr = (int32_t)nrow, nc = (int32_t)ncol;
    const bool nb_lane = lane < 9 && ((nbmask >> lane) & 1u);
    uint32_t t = 0;
    float worst = 0.0f;
    unsigned long long ndrift = 0;
    while (n > 0) {
        const unsigned long long top = hp[0];
        --n;
        if (n > 0) fd_wsift_down(hp, n, 0, hp[n], lane);
        const uint32_t p = (uint32_t)(top & 0x7FFFFFFFull);
        const float z0 = fd_unord((uint32_t)(top >> 32));
        if (lane == 0) Tord[p] = t;
        ++t;
        const int32_t pr = (int32_t)fd_row_of(p, dv), pc = (int32_t)(p - (uint32_t)pr * (uint32_t)nc);
        bool push = false;
        unsigned long long e = 0;
        if (lane < 9) {  // lane k looks at cell k of the 3 x 3 window (k = 4: the popped cell itself); all loads issued at once
            const int32_t rr = pr + lane / 3 - 1, cc = pc + lane % 3 - 1;
            if (rr >= 0 && rr < nr && cc >= 0 && cc < nc) {
                const uint32_t y = (uint32_t)rr * (uint32_t)nc + (uint32_t)cc;
                uint8_t fy = flags[y];
                const uint32_t ly = label[y], sy = S[y];
                const T zy = elev[y];
                if (lane == 4) {
                    if (!(fy & FDF_DISC)) {  // an initial outlet that no neighbour visited before it popped: it visits itself (d8 = 0)
                        flags[y] = fy | FDF_DISC;
                        d8[y] = 0;
                        fd_visit_z<T, W>(zy, out, y, z0, int_delv);
                    }
                } else if (nb_lane && ly == root && !(fy & FDF_DISC)) {
                    fy |= FDF_DISC;
                    d8[y] = fd_us[lane];
                    const float kf = fd_visit_z<T, W>(zy, out, y, z0, int_delv);
                    const uint32_t ky = fd_ord(kf);
                    const float drift = ky == sy ? 0.0f : fabsf(kf - fd_unord(sy));
                    worst = fmaxf(worst, drift);
                    ndrift += !(drift <= max_drift);
                    if (!(fy & FDF_QUEUED)) {
                        fy |= FDF_QUEUED;
                        push = true;
                        e = ((unsigned long long)ky << 32) | (unsigned long long)y;
                    }
                    flags[y] = fy;
                }
            }
        }
        uint32_t pm = __ballot_sync(0xFFFFFFFFu, push);
        while (pm) {
            const int src = __ffs(pm) - 1;
            pm &= pm - 1;
            const unsigned long long ev = fd_shfl64(e, src);
            if (in_smem && n == FDW_CAP) {  // the frontier outgrew shared memory: continue in the component's pool slice
                __syncwarp();
                for (uint32_t i = lane; i < n; i += 32) gheap[i] = fd_sheap[i];
                hp = gheap;
                in_smem = false;
                __syncwarp();
            }
            if (lane == 0) {
                uint32_t i = n;  // sift up
                while (i > 0) {
                    const uint32_t par = (i - 1) / FDW_ARY;
                    const unsigned long long pv = hp[par];
                    if (pv <= ev) break;
                    hp[i] = pv;
                    i = par;
                }
                hp[i] = ev;
            }
            ++n;
        }
        __syncwarp();
    }
    if (worst > 0.0f) atomicMax(&cnt->max_drift, (unsigned long long)__float_as_uint(worst));
    if (ndrift) atomicAdd(&cnt->n_drift, ndrift);
}

// (2) + output for every cell the replay did not visit: its parent lies more than `band` below it (never raised), or it is an
// outlet that visits itself
template <typename T, typename W>
__global__ void fd_finalize_kernel(const T* __restrict__ elev, const uint8_t* __restrict__ flags, const uint32_t* __restrict__ S,
                                   const uint32_t* __restrict__ Tord, int64_t nrow, int64_t ncol, uint32_t nbmask, float band,
                                   int int_delv, T* __restrict__ out, uint8_t* __restrict__ d8, FdCounters* cnt) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint8_t f = flags[i];
        if ((f & FDF_TIED) && !(f & FDF_SRCDISC)) continue;  // settled by the replay of its component
        const T z = elev[i];
        if (!(f & FDF_VALID)) {
            d8[i] = 247;
            out[i] = (T)((W)z + (W)0);
            continue;
        }
        const uint32_t s = S[i];
        if (s == FD_UNREACHED) {  // no outlet reaches this cell: never visited (d8 stays 0, dem.py:71)
            d8[i] = 0;
            out[i] = (T)((W)z + (W)0);
            atomicAdd(&cnt->n_unreached, 1ull);
            continue;
        }
        const int64_t r = i / ncol, c = i - r * ncol;
        uint32_t sn[9], m = FD_UNREACHED;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            sn[k] = FD_UNREACHED;
            if (!((nbmask >> k) & 1u)) continue;
            const int64_t rr = r + k / 3 - 1, cc = c + k % 3 - 1;
            if (rr < 0 || rr >= nrow || cc < 0 || cc >= ncol) continue;
            sn[k] = S[rr * ncol + cc];
            m = min(m, sn[k]);
        }
        if (m < s) {  // visited first by the candidate (level within band of the smallest) that pops first
            uint32_t bestT = 0xFFFFFFFFu;
            int bestk = -1, ncand = 0;
#pragma unroll
            for (int k = 0; k < 9; ++k)
                if (sn[k] != FD_UNREACHED && fd_nearmin(sn[k], m, band)) {
                    ++ncand;
                    const int64_t y = (r + k / 3 - 1) * ncol + (c + k % 3 - 1);
                    const uint32_t ty = (flags[y] & FDF_TIED) ? Tord[y] : 0u;  // several candidates always share a component
                    if (bestk < 0 || ty < bestT) {
                        bestT = ty;
                        bestk = k;
                    }
                }
            d8[i] = fd_us[8 - bestk];  // the code that points from this cell at neighbour bestk
            out[i] = (T)((W)z + (W)0);
        } else {  // an outlet without a lower neighbour: it visits itself with its own key
            d8[i] = 0;
            fd_visit<T, W>(elev, out, (uint32_t)i, (float)z, int_delv);
        }
    }
}

// flags back to (valid, outlet) before a retry with a wider band
__global__ void fd_reset_flags_kernel(uint8_t* __restrict__ flags, int64_t n) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) flags[i] &= FDF_STATIC;
}

// ---------------------------------------------------------------------------------------------------------
// Host side
// ---------------------------------------------------------------------------------------------------------
struct FdScratch {  // device buffers of one call: slots of the handle that persist (and only grow) across calls
    pfd_handle* h;
    int next = 0;
    explicit FdScratch(pfd_handle* handle) : h(handle) {}
    int alloc(pfd_handle*, void** p, size_t bytes) {
        if (next >= (int)(sizeof(h->fill_bufs) / sizeof(h->fill_bufs[0]))) return pfd_fail(h, PFD_ERR_STATE, "pfd_fill_depressions: out of buffer slots");
        PFD_TRY(pfd_reserve(h, h->fill_bufs[next], bytes));
        *p = h->fill_bufs[next++].p;
        return PFD_OK;
    }
    void release(void* p) {  // the last slot handed out can be handed out again (heap pool / root list of a retry)
        if (next > 0 && h->fill_bufs[next - 1].p == p) --next;
    }
};

static inline float uint_as_float_host(uint32_t u) {
    float f;
    memcpy(&f, &u, 4);
    return f;
}
static inline uint32_t float_as_uint_host(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
}

// runs a tile kernel over the active tiles until no tile changes anything; `launch(cur, next, changed)` queues one pass. Passes are
// queued four at a time between two looks at the flags (a pass after convergence finds no active tile and costs next to nothing)
#define FD_BATCH 4
template <class Launch>
static int fd_converge(pfd_handle* h, int64_t ntiles, uint8_t* act[2], unsigned int* changed_dev, Launch launch, int* passes_out) {
    // act[0] holds the tiles to start from (set by the kernels that created the state to relax)
    int passes = 0, cur = 0;
    for (;;) {
        PFD_CUDA(h, cudaMemsetAsync(changed_dev, 0, FD_BATCH * sizeof(unsigned int), h->stream));
        for (int p = 0; p < FD_BATCH; ++p, cur ^= 1) {
            PFD_CUDA(h, cudaMemsetAsync(act[cur ^ 1], 0, (size_t)ntiles, h->stream));
            launch(act[cur], act[cur ^ 1], changed_dev + p);
            PFD_LAUNCH_CHECK(h);
        }
        unsigned int changed[FD_BATCH];
        PFD_CUDA(h, cudaMemcpyAsync(changed, changed_dev, sizeof(changed), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        bool done = false;
        for (int p = 0; p < FD_BATCH && !done; ++p) {
            ++passes;
            done = changed[p] == 0;
        }
        if (done) break;
    }
    if (passes_out) *passes_out = passes;
    return PFD_OK;
}

template <typename T, typename W>
static int fd_fill_impl(pfd_handle* h, const T* elev, int64_t nrow, int64_t ncol, int mode, const int64_t* idxs_pit, int64_t npit,
                        double nodata, int has_elv_max, double elv_max, int connectivity, int int_delv, T* out, uint8_t* d8,
                        int64_t* stats) {
    const int64_t n = nrow * ncol;
    const uint32_t nbmask = connectivity == 4 ? FD_NB4 : FD_NB8;
    const int nodata_nan = nodata != nodata;
    const FdDiv dv = fd_make_div((uint32_t)ncol);
    FdScratch sc(h);
    uint8_t* flags = nullptr;
    uint32_t *S = nullptr, *A1 = nullptr, *A2 = nullptr, *A3 = nullptr, *A4 = nullptr;
    FdCounters* cnt = nullptr;
    unsigned int* changed = nullptr;
    PFD_TRY(sc.alloc(h, (void**)&flags, (size_t)n));
    PFD_TRY(sc.alloc(h, (void**)&S, (size_t)n * 4));
    PFD_TRY(sc.alloc(h, (void**)&A1, (size_t)n * 4));
    PFD_TRY(sc.alloc(h, (void**)&A2, (size_t)n * 4));
    PFD_TRY(sc.alloc(h, (void**)&A3, (size_t)n * 4));
    PFD_TRY(sc.alloc(h, (void**)&A4, (size_t)n * 4));
    PFD_TRY(sc.alloc(h, (void**)&cnt, sizeof(FdCounters)));
    PFD_TRY(sc.alloc(h, (void**)&changed, 16 * sizeof(unsigned int)));
    FdCounters hc;
    memset(&hc, 0, sizeof(hc));
    hc.minkey = ~0ull;
    PFD_CUDA(h, cudaMemcpyAsync(cnt, &hc, sizeof(hc), cudaMemcpyHostToDevice, h->stream));
    const int grid = grid_for(n, 256, 4);
    const int64_t nty = (nrow + FD_T - 1) / FD_T, ntx = (ncol + FD_T - 1) / FD_T;
    const int64_t lty = (nrow + FL_T - 1) / FL_T, ltx = (ncol + FL_T - 1) / FL_T;
    uint8_t* act[2] = {nullptr, nullptr};
    PFD_TRY(sc.alloc(h, (void**)&act[0], (size_t)(lty * ltx)));
    PFD_TRY(sc.alloc(h, (void**)&act[1], (size_t)(lty * ltx)));
    PFD_CUDA(h, cudaMemsetAsync(act[0], 0, (size_t)(nty * ntx), h->stream));
    fd_init_kernel<T><<<grid, 256, 0, h->stream>>>(elev, nrow, ncol, nodata, nodata_nan, nbmask, mode, has_elv_max, elv_max, flags, S, act[0], ntx,
                                                   cnt);
    PFD_LAUNCH_CHECK(h);
    if (mode == 1) {
        fd_keep_min_kernel<<<grid, 256, 0, h->stream>>>(n, flags, S, cnt);
        PFD_LAUNCH_CHECK(h);
    } else if (mode == 2) {
        fd_pits_kernel<T><<<grid_for(npit > 0 ? npit : 1, 256, 1), 256, 0, h->stream>>>(elev, n, idxs_pit, npit, flags, S, act[0], ncol, ntx, cnt);
        PFD_LAUNCH_CHECK(h);
    }
    PFD_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (hc.err_pit) return pfd_fail(h, PFD_ERR_INVALID_ARG, "pfd_fill_depressions: idxs_pit holds an index outside the raster or of a nodata cell");
    if (has_elv_max && mode != 2 && hc.n_outlets == 0) return pfd_fail(h, PFD_ERR_INVALID_ARG, "No initial outlet cells found.");
    const unsigned long long n_outlets = hc.n_outlets, max_abs_bits = hc.max_abs;

    // (1) levels
    const auto t_start = std::chrono::steady_clock::now();
    int passes_levels = 0, passes_labels = 0, tries = 0;
    PFD_TRY(fd_converge(h, nty * ntx, act, changed, [&](const uint8_t* cur, uint8_t* next, unsigned int* chg) {
        fd_relax_kernel<T><<<(unsigned)(nty * ntx), 1024, 0, h->stream>>>(elev, flags, S, nrow, ncol, nty, ntx, nbmask, cur, next, chg);
    }, &passes_levels));

    const auto t_levels = std::chrono::steady_clock::now();
    // (3) tie components and their replay; the band widens until no key drifted half a band away from its level
    uint32_t *M = A1, *label = A2, *link = A3, *cntarr = A1, *off = A3, *Tord = A4;
    // "tied" = levels within `band` (absolute) of each other. The first attempt takes exact ties only (band 0): raising is exact in
    // float64, and in float32 whenever z0 - z1 fits 24 bits (elevations of one sign and similar magnitude); a raise that misses its
    // pour level is seen by the replay as drift, and the band widens to 16 ulps of max |z| or 8 x the drift seen.
    const float zmax = uint_as_float_host((uint32_t)max_abs_bits);
    float band = 0.0f;  // exact ties first: any raise that misses its pour level (possible in float32 only) shows up as drift
    for (;; ++tries) {
        const float max_drift = band * 0.25f;
        PFD_CUDA(h, cudaMemsetAsync(act[0], 0, (size_t)(lty * ltx), h->stream));
        fd_tie_kernel<<<grid, 256, 0, h->stream>>>(S, nrow, ncol, nbmask, band, M, label, link, act[0], ltx);
        PFD_LAUNCH_CHECK(h);
        fd_tie2_kernel<<<grid, 256, 0, h->stream>>>(S, M, link, nrow, ncol, nbmask, band, label, act[0], ltx);
        PFD_LAUNCH_CHECK(h);
        int lp = 0;
        PFD_TRY(fd_converge(h, lty * ltx, act, changed, [&](const uint8_t* cur, uint8_t* next, unsigned int* chg) {
            fd_label_kernel<<<(unsigned)(lty * ltx), 256, 0, h->stream>>>(S, M, label, link, nrow, ncol, lty, ltx, nbmask, band, cur, next, chg);
        }, &lp));
        passes_labels += lp;
        // heap slices (cnt = A1, off = A3)
        memset(&hc, 0, sizeof(hc));
        PFD_CUDA(h, cudaMemcpyAsync(cnt, &hc, sizeof(hc), cudaMemcpyHostToDevice, h->stream));
        PFD_CUDA(h, cudaMemsetAsync(cntarr, 0, (size_t)n * 4, h->stream));
        fd_count_kernel<<<grid, 256, 0, h->stream>>>(label, n, cntarr, cnt);
        PFD_LAUNCH_CHECK(h);
        PFD_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        unsigned long long* pool = nullptr;
        uint32_t* roots = nullptr;
        if (hc.n_tied > 0) {
            PFD_TRY(sc.alloc(h, (void**)&pool, (size_t)hc.n_tied * 8));
            PFD_TRY(sc.alloc(h, (void**)&roots, (size_t)hc.n_roots * 4));
            const unsigned long long nroots = hc.n_roots;
            fd_roots_kernel<<<grid, 256, 0, h->stream>>>(label, n, cntarr, off, roots, nroots, cnt);
            PFD_LAUNCH_CHECK(h);
            fd_sources_kernel<<<grid, 256, 0, h->stream>>>(S, label, nrow, ncol, nbmask, flags, cntarr, off, pool);
            PFD_LAUNCH_CHECK(h);
            PFD_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
            PFD_CUDA(h, cudaStreamSynchronize(h->stream));
            const unsigned long long nbig = hc.big_fill, nsmall = nroots - nbig;
            // the few long replays (one warp each) run next to the many short ones (one thread each) on a second stream
            cudaStream_t s2 = (nbig > 0 && nsmall > 0 && h->copy_stream && h->ev_copy) ? h->copy_stream : h->stream;
            if (s2 != h->stream) {
                PFD_CUDA(h, cudaEventRecord(h->ev_copy, h->stream));
                PFD_CUDA(h, cudaStreamWaitEvent(s2, h->ev_copy, 0));
            }
            if (nbig > 0) {
                if (!h->fill_attr_set[sizeof(W) == 8]) {
                    PFD_CUDA(h, cudaFuncSetAttribute(fd_simulate_warp_kernel<T, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, FDW_CAP * 8));
                    h->fill_attr_set[sizeof(W) == 8] = true;
                }
                fd_simulate_warp_kernel<T, W><<<(unsigned)nbig, 32, FDW_CAP * 8, h->stream>>>(
                    roots + nsmall, off, cntarr, pool, S, label, elev, flags, Tord, d8, out, nrow, ncol, nbmask, int_delv, max_drift, dv, cnt);
                PFD_LAUNCH_CHECK(h);
            }
            if (nsmall > 0) {
                fd_simulate_kernel<T, W><<<(unsigned)((nsmall + 31) / 32), 32, 0, s2>>>(
                    roots, (uint32_t)nsmall, off, cntarr, pool, S, label, elev, flags, Tord, d8, out, nrow, ncol, nbmask, int_delv, max_drift, dv, cnt);
                PFD_LAUNCH_CHECK(h);
            }
            if (s2 != h->stream) {
                PFD_CUDA(h, cudaEventRecord(h->ev_copy, s2));
                PFD_CUDA(h, cudaStreamWaitEvent(h->stream, h->ev_copy, 0));
            }
        }
        PFD_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        if (roots) sc.release(roots);
        if (pool) sc.release(pool);
        if (hc.n_drift == 0) break;
        if (band > 3.0e38f) return pfd_fail(h, PFD_ERR_CUDA, "pfd_fill_depressions: internal error (drift with an unbounded band)");
        // some key drifted too far from its level: everything it may have been ordered against must share its heap
        const float seen = uint_as_float_host((uint32_t)hc.max_drift);
        float nb = band > 0.0f ? band * 32.0f : zmax * 1.9073486e-06f;  // first widening: 16 ulps of max |z| (2^-19 relative)
        if (!(nb >= seen * 8.0f)) nb = seen * 8.0f;
        band = (nb > 3.0e38f || nb != nb) ? INFINITY : nb;
        fd_reset_flags_kernel<<<grid, 256, 0, h->stream>>>(flags, n);
        PFD_LAUNCH_CHECK(h);
    }
    fd_finalize_kernel<T, W><<<grid, 256, 0, h->stream>>>(elev, flags, S, Tord, nrow, ncol, nbmask, band, int_delv, out, d8, cnt);
    PFD_LAUNCH_CHECK(h);
    const unsigned long long n_tied = hc.n_tied, n_roots = hc.n_roots, max_comp = hc.max_comp, max_drift_seen = hc.max_drift;
    PFD_CUDA(h, cudaMemcpyAsync(&hc, cnt, sizeof(hc), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    if (stats) {
        stats[0] = passes_levels;
        stats[1] = passes_labels;
        stats[2] = (int64_t)n_tied;
        stats[3] = (int64_t)n_roots;
        stats[4] = (int64_t)hc.n_unreached;
        stats[5] = (int64_t)n_outlets;
        stats[6] = (int64_t)float_as_uint_host(band);                                  // float32 bits
        stats[7] = (int64_t)max_comp;
        stats[8] = (int64_t)max_drift_seen;                                           // float32 bits
        stats[9] = tries + 1;
        const auto t_end = std::chrono::steady_clock::now();
        stats[10] = (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(t_levels - t_start).count();  // (1), us
        stats[11] = (int64_t)std::chrono::duration_cast<std::chrono::microseconds>(t_end - t_levels).count();    // (3) + (2), us
    }
    return PFD_OK;
}
