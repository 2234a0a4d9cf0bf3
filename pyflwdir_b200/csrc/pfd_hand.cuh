// pfd_hand.cuh -- dem.height_above_nearest_drain (pyflwdir/dem.py:299-330) as a tile-hierarchical PATH SUM.
//
// The reference walks the sequence downstream -> upstream and sets hand[i] = hand[ds] + float64(elevtn[i] - elevtn[ds]) (the
// difference formed in elevtn's type), 0 at drain cells and pits: a left fold of float64 additions along every flow path. The
// terms are float32 (or float64) differences of neighbouring elevations, a few dozen bits wide, and their running sums stay
// far inside the 53 bits of a double -- so on real rasters every one of those additions is EXACT, and exact sums may be
// re-associated freely. This file computes the path sums with pointer doubling instead of hop by hop:
//   phase A (one CTA per 64 x 64 tile): every cell carries the summary of the path segment to its current ancestor,
//       (sum of the hop terms, "passed a drain cell" flag); summaries compose associatively (a segment that already passed a
//       drain cell ignores what lies downstream), so ancestor <- ancestor's ancestor doubles the segment per round (double-
//       buffered in shared memory, one barrier per round) until the ancestor is a pit or the cell where the path leaves the tile;
//   phase B: the same doubling over the ring nodes of all tiles (node = border cell, successor = the border cell of the
//       neighbouring tile its path enters) until every node knows its sum down to the pit -- or that it never gets there;
//   phase C: hand = own segment (+ the solved sum behind the exit cell); -9999 for nodata and for cells that reach no pit,
//       exactly the cells outside the reference's sequence.
// Then hand_check_kernel re-evaluates the reference's statement for EVERY cell from the finished values (hand[i] == hand[ds] +
// dz bit for bit, 0 at drains and pits, -9999 exactly where the downstream cell is -9999). Zero violations prove the array
// equal to the reference's (induction from the pits upstream); a single violation -- an addition that was not exact -- and
// pfd_hand discards the result and runs the hop-by-hop sweep (pfd_tilesweep.cuh / pfd_sweeps.cuh) instead.
#pragma once
#include "pfd_tiles.cuh"

#define HD_SINK 0x7FFFFFFEu     // node successor: the path ended in a pit (value final)
#define HD_INVALID 0x7FFFFFFFu  // node reaches no pit (loop, nodata)
#define HD_HIT 0x80000000u      // node / cell flag: the segment already passed a drain cell
#define HD_ROOT_NONE 0xFFFFu
#define HD_ROOT_EXIT 0x4000u
#define HD_ROOT_HIT 0x8000u

template <typename T>
__device__ __forceinline__ double hd_dz(T a, T b) { return (double)(T)(a - b); }
template <>
__device__ __forceinline__ double hd_dz<float>(float a, float b) { return (double)__fsub_rn(a, b); }
template <>
__device__ __forceinline__ double hd_dz<double>(double a, double b) { return __dsub_rn(a, b); }

struct HandTileShared {
    double D[2][TL_CELLS];
    uint16_t nx[2][TL_CELLS];  // in-tile index of the current ancestor | HD_ROOT_HIT
    double wexit[TL_RING];     // exit cells: their own hop term
    uint32_t eslot[TL_RING];   //             ring slot of the cell they drain into | HD_HIT when they are drain cells
    uint8_t kind[TL_CELLS];    // 0 inner, 1 pit, 2 exit, 3 nodata
};

// What a cell contributes to the path summaries. hit(g): the cell's value does not depend on what lies downstream (HAND: a drain
// cell; label filling: a cell that has a value of its own), hit_value(g): the segment sum frozen there, w(g, g_ds): the hop term
// of a cell that is not hit.
template <typename T>
struct HandSrc {
    const uint8_t* drain;
    const T* elev;
    __device__ __forceinline__ bool hit(long long g) const { return drain[g] == 1; }
    __device__ __forceinline__ double hit_value(long long) const { return 0.0; }
    __device__ __forceinline__ double w(long long g, long long g_ds) const { return hd_dz<T>(elev[g], elev[g_ds]); }
};
// core.fillnodata_upstream (pyflwdir/core.py:120-146; basins.basins, pyflwdir/basins.py:12-18, is this on a raster of outlet ids):
// every cell without a value takes the value of the first cell downstream that has one. No arithmetic at all: the "sum" carries
// (index of that cell + 1), exact in a double, and the result is gathered from there.
template <typename U, class Has>
struct FillSrc {
    const U* data;
    Has has;
    __device__ __forceinline__ bool hit(long long g) const { return has(data[g]); }
    __device__ __forceinline__ double hit_value(long long g) const { return (double)(g + 1); }
    __device__ __forceinline__ double w(long long, long long) const { return 0.0; }
};

// streams.stream_distance(real_length=False) (pyflwdir/streams.py:272-315): hops to the first masked cell downstream (or the pit),
// int32 -- unit hop terms, exact
struct HopSrc {
    const uint8_t* mask;  // may be null: distance to the pit
    __device__ __forceinline__ bool hit(long long g) const { return mask && mask[g]; }
    __device__ __forceinline__ double hit_value(long long) const { return 0.0; }
    __device__ __forceinline__ double w(long long, long long) const { return 1.0; }
};

template <class Src>
__global__ void __launch_bounds__(1024, 2) hand_tile_a_kernel(const uint8_t* __restrict__ dir, Src src, long long nrow, long long ncol,
                                                              long long ntx, uint16_t* __restrict__ hroot, double* __restrict__ hD,
                                                              uint32_t* __restrict__ s_nxt, double* __restrict__ s_val) {
    extern __shared__ __align__(16) unsigned char hd_smem[];
    HandTileShared& s = *reinterpret_cast<HandTileShared*>(hd_smem);
    const long long ty = blockIdx.y, tx = blockIdx.x;
    const uint32_t tile = (uint32_t)((ty + 1) * ntx + tx);
    const long long r0 = ty * TL_H, c0 = tx * TL_W;
    // loads in two waves (all directions, then everything that depends on them) so that the four cells of a thread overlap
    uint32_t dd[4];
    long long gg[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + 1024 * j;
        const long long r = r0 + (i >> 6), c = c0 + (i & 63);
        dd[j] = PFD_DIR_NODATA;
        gg[j] = 0;
        if (r < nrow && c < ncol) {
            gg[j] = r * ncol + c;
            dd[j] = dir[gg[j]];
        }
    }
    bool hitf[4];
    double wv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hitf[j] = false;
        wv[j] = 0.0;
        if (dd[j] != PFD_DIR_NODATA) {
            hitf[j] = src.hit(gg[j]);
            if (hitf[j]) wv[j] = src.hit_value(gg[j]);
            else if (dd[j] < 8u) wv[j] = src.w(gg[j], gg[j] + pfd_slot_off((int)dd[j], ncol));
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + 1024 * j;
        const int ly = i >> 6, lx = i & 63;
        const uint32_t d = dd[j];
        uint16_t nx = (uint16_t)i;
        double D = 0.0;
        uint8_t kind = 3;
        if (d != PFD_DIR_NODATA) {
            if (d >= 8u) {
                kind = 1;  // a pit: the root of its paths; one that is hit carries its value into every segment that ends there
                if (hitf[j]) {
                    nx |= HD_ROOT_HIT;
                    D = wv[j];
                }
            } else {
                const int y = ly + pfd_slot_dr((int)d), x = lx + pfd_slot_dc((int)d);
                if ((unsigned)y < (unsigned)TL_H && (unsigned)x < (unsigned)TL_W) {
                    kind = 0;
                    nx = (uint16_t)((y << 6) | x) | (hitf[j] ? HD_ROOT_HIT : 0);
                    D = wv[j];
                } else {
                    kind = 2;
                    const int rp = tl_ring_pos(ly, lx);
                    s.wexit[rp] = wv[j];
                    s.eslot[rp] = tl_exit_slot(tile, (uint32_t)ntx, ly, lx, d) | (hitf[j] ? HD_HIT : 0u);
                }
            }
        }
        s.kind[i] = kind;
        s.nx[0][i] = nx;
        s.D[0][i] = D;
    }
    __syncthreads();
    // own (ancestor, sum) live in registers; shared memory only publishes them to the cells that point here
    uint32_t own_a[4];
    double own_d[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        own_a[j] = s.nx[0][threadIdx.x + 1024 * j];
        own_d[j] = s.D[0][threadIdx.x + 1024 * j];
    }
    int fin = 0;
    for (int k = 0; k < TL_MAXROUNDS; ++k) {
        const int cur = k & 1, nb = cur ^ 1;
        bool ch = false;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = threadIdx.x + 1024 * j;
            const uint32_t n = own_a[j] & 0xFFFu;
            const uint32_t an = s.nx[cur][n];
            if (!(own_a[j] & HD_ROOT_HIT)) own_d[j] = __dadd_rn(own_d[j], s.D[cur][n]);  // (a root carries the empty segment: + 0)
            own_a[j] = (an & 0xFFFu) | ((own_a[j] | an) & HD_ROOT_HIT);
            s.D[nb][i] = own_d[j];
            s.nx[nb][i] = (uint16_t)own_a[j];
            ch |= (an & 0xFFFu) != n;
        }
        fin = nb;
        if (!__syncthreads_or((int)ch)) break;
    }
    // per-cell records for phase C
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + 1024 * j;
        const long long r = r0 + (i >> 6), c = c0 + (i & 63);
        if (r < nrow && c < ncol) {
            const uint32_t a = s.nx[fin][i];
            const uint32_t root = a & 0xFFFu;
            uint16_t rec = HD_ROOT_NONE;
            // (an in-tile loop never settles on a root; one of 2^k cells comes back to itself, which is no root either)
            // and a path that runs into a nodata cell (possible in graphs loaded from idxs_ds arrays) reaches no pit
            if (s.kind[i] != 3 && (s.kind[root] == 1 || s.kind[root] == 2) && (s.nx[fin][root] & 0xFFFu) == root)
                rec = (uint16_t)(root | (a & HD_ROOT_HIT) | (s.kind[root] == 2 ? HD_ROOT_EXIT : 0));
            hroot[r * ncol + c] = rec;
            hD[r * ncol + c] = s.D[fin][i];
        }
    }
    // ring nodes
    if (threadIdx.x < TL_RING) {
        uint32_t nxt = HD_INVALID;
        double val = 0.0;
        if (threadIdx.x < TL_NRING) {
            const int ri = tl_ring_cell(threadIdx.x);
            const uint32_t a = s.nx[fin][ri];
            const uint32_t root = a & 0xFFFu;
            if (s.kind[ri] != 3 && (s.kind[root] == 1 || s.kind[root] == 2) && (s.nx[fin][root] & 0xFFFu) == root) {
                const uint32_t hit = (a & HD_ROOT_HIT) ? HD_HIT : 0u;
                val = s.D[fin][ri];
                if (s.kind[root] == 1) {
                    nxt = HD_SINK | hit;
                } else {
                    const int rp = tl_ring_pos((int)(root >> 6), (int)(root & 63u));
                    if (!hit) val = __dadd_rn(val, s.wexit[rp]);
                    nxt = s.eslot[rp] | hit;
                }
            }
        }
        s_nxt[tile * TL_RING + threadIdx.x] = nxt;
        s_val[tile * TL_RING + threadIdx.x] = val;
    }
}

// one doubling round over the ring nodes [lo, hi): (successor, sum, flag) <- composed with the successor's
__global__ void hand_slots_round_kernel(const uint32_t* __restrict__ nxt_c, const double* __restrict__ val_c, uint32_t* __restrict__ nxt_n,
                                        double* __restrict__ val_n, long long lo, long long hi, unsigned int* __restrict__ changed) {
    bool ch = false;
    for (long long i = lo + blockIdx.x * (long long)blockDim.x + threadIdx.x; i < hi; i += (long long)gridDim.x * blockDim.x) {
        const uint32_t a = nxt_c[i];
        const uint32_t j = a & 0x7FFFFFFFu;
        if (j >= HD_SINK) {  // final: mirrored into the other buffer once
            if (nxt_n[i] != a) {
                nxt_n[i] = a;
                val_n[i] = val_c[i];
            }
            continue;
        }
        const uint32_t b = nxt_c[j];
        const double v = val_c[i];
        val_n[i] = (a & HD_HIT) ? v : __dadd_rn(v, val_c[j]);
        nxt_n[i] = (b & 0x7FFFFFFFu) | ((a | b) & HD_HIT);
        ch = true;
    }
    if (ch) *changed = 1u;
}

struct HandOut {  // hand[g] = the path sum, -9999 where no pit is reached
    double* out;
    __device__ __forceinline__ void operator()(long long g, double h) const { out[g] = h; }
};
struct HopOut {  // int32 hop counts, -9999 where no pit is reached
    int32_t* out;
    __device__ __forceinline__ void operator()(long long g, double h) const { out[g] = (int32_t)h; }
};
template <typename U>
struct FillOut {  // in place: a cell without a value takes the value of the cell the "sum" names (sources never change)
    U* data;
    __device__ __forceinline__ void operator()(long long g, double h) const {
        if (h > 0.0) {
            const long long src = (long long)h - 1;
            if (src != g) data[g] = data[src];
        }
    }
};

template <class Out>
__global__ void __launch_bounds__(1024) hand_tile_c_kernel(const uint8_t* __restrict__ dir, const uint16_t* __restrict__ hroot,
                                                           const double* __restrict__ hD, const uint32_t* __restrict__ s_nxt,
                                                           const double* __restrict__ s_val, long long nrow, long long ncol, long long ntx,
                                                           Out out) {
    __shared__ double ringval[TL_RING];
    __shared__ uint8_t ringok[TL_RING];
    const long long ty = blockIdx.y, tx = blockIdx.x;
    const uint32_t tile = (uint32_t)((ty + 1) * ntx + tx);
    const long long r0 = ty * TL_H, c0 = tx * TL_W;
    if (threadIdx.x < TL_RING) {
        const uint32_t a = s_nxt[tile * TL_RING + threadIdx.x];
        ringok[threadIdx.x] = (a & 0x7FFFFFFFu) == HD_SINK;
        ringval[threadIdx.x] = s_val[tile * TL_RING + threadIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + 1024 * j;
        const long long r = r0 + (i >> 6), c = c0 + (i & 63);
        if (r >= nrow || c >= ncol) continue;
        const long long g = r * ncol + c;
        const uint32_t rec = hroot[g];
        double h = -9999.0;
        if (rec != HD_ROOT_NONE) {
            const double D = hD[g];
            if (!(rec & HD_ROOT_EXIT)) {
                h = D;
            } else {
                const uint32_t root = rec & 0xFFFu;
                const int rp = tl_ring_pos((int)(root >> 6), (int)(root & 63u));
                if (ringok[rp]) h = (rec & HD_ROOT_HIT) ? D : __dadd_rn(D, ringval[rp]);
            }
        }
        out(g, h);
    }
}

// -9999 for every cell that reaches no pit, decided by the path structure alone (valid whether or not the sums were exact): the
// hop-by-hop tile sweep treats every drain cell as a source, also one that sits above a loop, outside the reference's sequence
__global__ void __launch_bounds__(1024) hand_mask_unreached_kernel(const uint16_t* __restrict__ hroot, const uint32_t* __restrict__ s_nxt,
                                                                   long long nrow, long long ncol, long long ntx, double* __restrict__ out) {
    __shared__ uint8_t ringok[TL_RING];
    const long long ty = blockIdx.y, tx = blockIdx.x;
    const uint32_t tile = (uint32_t)((ty + 1) * ntx + tx);
    const long long r0 = ty * TL_H, c0 = tx * TL_W;
    if (threadIdx.x < TL_RING) ringok[threadIdx.x] = (s_nxt[tile * TL_RING + threadIdx.x] & 0x7FFFFFFFu) == HD_SINK;
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int i = threadIdx.x + 1024 * j;
        const long long r = r0 + (i >> 6), c = c0 + (i & 63);
        if (r >= nrow || c >= ncol) continue;
        const uint32_t rec = hroot[r * ncol + c];
        bool reached = rec != HD_ROOT_NONE;
        if (reached && (rec & HD_ROOT_EXIT)) {
            const uint32_t root = rec & 0xFFFu;
            reached = ringok[tl_ring_pos((int)(root >> 6), (int)(root & 63u))];
        }
        if (!reached) out[r * ncol + c] = -9999.0;
    }
}

// the reference's per-cell statement, from the finished values: counts the cells that violate it (four independent cells per
// thread and trip, so that the two dependent rounds of loads -- the cell, then its downstream cell -- overlap)
template <typename T>
__global__ void __launch_bounds__(256) hand_check_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ drain,
                                                         const T* __restrict__ elev, int64_t n, int64_t ncol,
                                                         const double* __restrict__ hand, unsigned long long* __restrict__ n_bad) {
    unsigned int bad = 0;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i0 = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i0 < n; i0 += 4 * stride) {
        uint32_t d[4];
        uint8_t dr[4];
        T e[4];
        double got[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int64_t i = i0 + k * stride;
            d[k] = PFD_DIR_NODATA;
            got[k] = -9999.0;
            if (i < n) {
                d[k] = dir[i];
                dr[k] = drain[i];
                e[k] = elev[i];
                got[k] = hand[i];
            }
        }
        double hds[4];
        T eds[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (d[k] < 8u) {
                const int64_t ds = i0 + k * stride + pfd_slot_off((int)d[k], ncol);
                hds[k] = hand[ds];
                eds[k] = elev[ds];
            }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            double want = -9999.0;
            if (d[k] != PFD_DIR_NODATA) {
                if (d[k] >= 8u) want = 0.0;  // a pit: in the sequence, hand[pit] + 0
                else if (__double_as_longlong(hds[k]) != __double_as_longlong(-9999.0))  // (else: the cell drains to no pit either)
                    want = dr[k] == 1 ? 0.0 : __dadd_rn(hds[k], hd_dz<T>(e[k], eds[k]));
            }
            bad += (i0 + k * stride < n) && __double_as_longlong(got[k]) != __double_as_longlong(want);
        }
    }
    bad = __reduce_add_sync(0xFFFFFFFFu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(n_bad, (unsigned long long)bad);
}

// Host side. Phases A - C for any source / output pair; the ring-node buffers stay on the handle.
static int hd_next_slot = 0;
template <class Src>
static int hd_slot() {  // one flag per instantiation of phase A (its shared-memory opt-in is per device, hence kept per handle)
    static const int slot = hd_next_slot++;
    return slot;
}

template <class Src, class Out>
static int hd_solve(pfd_handle* h, Src src, Out outf) {
    const int attr_slot = hd_slot<Src>();
    if (attr_slot >= (int)(sizeof(h->hand_attr_set) / sizeof(h->hand_attr_set[0]))) return pfd_fail(h, PFD_ERR_STATE, "path sums: out of attribute slots");
    const long long nrow = h->nrow, ncol = h->ncol, n = h->n;
    const long long ntx = (ncol + TL_W - 1) / TL_W, nty = (nrow + TL_H - 1) / TL_H;
    const long long nslots = (nty + 2) * ntx * TL_RING;
    if (nslots >= (long long)HD_SINK) return pfd_fail(h, PFD_ERR_UNSUPPORTED, "path sums: too many ring nodes");
    const uint8_t* dir = (const uint8_t*)h->dir.p + h->dir_off;
    PFD_TRY(pfd_reserve(h, h->hand_root, (size_t)n * sizeof(uint16_t)));
    PFD_TRY(pfd_reserve(h, h->hand_sum, (size_t)n * sizeof(double)));
    PFD_TRY(pfd_reserve(h, h->hand_slots, (size_t)nslots * 2 * (sizeof(uint32_t) + sizeof(double)) + 64));
    uint16_t* hroot = (uint16_t*)h->hand_root.p;
    double* hD = (double*)h->hand_sum.p;
    double* sval[2] = {(double*)h->hand_slots.p, (double*)h->hand_slots.p + nslots};
    uint32_t* snxt[2] = {(uint32_t*)(sval[1] + nslots), (uint32_t*)(sval[1] + nslots) + nslots};
    unsigned int* changed = (unsigned int*)(snxt[1] + nslots);
    if (!h->hand_attr_set[attr_slot]) {
        PFD_CUDA(h, cudaFuncSetAttribute(hand_tile_a_kernel<Src>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(HandTileShared)));
        h->hand_attr_set[attr_slot] = true;
    }
    const dim3 grid((unsigned)ntx, (unsigned)nty);
    hand_tile_a_kernel<Src><<<grid, 1024, sizeof(HandTileShared), h->stream>>>(dir, src, nrow, ncol, ntx, hroot, hD, snxt[0], sval[0]);
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaMemsetAsync(snxt[1], 0, (size_t)nslots * sizeof(uint32_t), h->stream));
    const long long lo = ntx * TL_RING, hi = (nty + 1) * ntx * TL_RING;
    int fin = 0;
    for (int k = 0; k < 48; ++k) {
        const int cur = k & 1, nb = cur ^ 1;
        PFD_CUDA(h, cudaMemsetAsync(changed, 0, sizeof(unsigned int), h->stream));
        hand_slots_round_kernel<<<grid_for(hi - lo, 256, 2, 148 * 16), 256, 0, h->stream>>>(snxt[cur], sval[cur], snxt[nb], sval[nb], lo, hi, changed);
        PFD_LAUNCH_CHECK(h);
        fin = nb;
        unsigned int ch = 0;
        PFD_CUDA(h, cudaMemcpyAsync(&ch, changed, sizeof(ch), cudaMemcpyDeviceToHost, h->stream));
        PFD_CUDA(h, cudaStreamSynchronize(h->stream));
        if (!ch) break;
    }
    h->hand_fin = fin;
    hand_tile_c_kernel<Out><<<grid, 1024, 0, h->stream>>>(dir, hroot, hD, snxt[fin], sval[fin], nrow, ncol, ntx, outf);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}

// HAND: the path sums, then the proof; *n_bad = number of cells that violate the reference's statement (0 = the result is the
// reference's)
template <typename T>
static int hand_pathsum(pfd_handle* h, const uint8_t* drain_dev, const T* elev_dev, double* out_dev, unsigned long long* n_bad) {
    PFD_TRY((hd_solve(h, HandSrc<T>{drain_dev, elev_dev}, HandOut{out_dev})));
    const long long ntx = (h->ncol + TL_W - 1) / TL_W, nty = (h->nrow + TL_H - 1) / TL_H;
    const long long nslots = (nty + 2) * ntx * TL_RING;
    unsigned long long* bad = (unsigned long long*)((unsigned int*)((uint32_t*)((double*)h->hand_slots.p + 2 * nslots) + 2 * nslots) + 2);
    PFD_CUDA(h, cudaMemsetAsync(bad, 0, sizeof(unsigned long long), h->stream));
    hand_check_kernel<T><<<grid_for(h->n, 256, 4, 148 * 32), 256, 0, h->stream>>>((const uint8_t*)h->dir.p + h->dir_off, drain_dev, elev_dev, h->n,
                                                                                  h->ncol, out_dev, bad);
    PFD_LAUNCH_CHECK(h);
    PFD_CUDA(h, cudaMemcpyAsync(n_bad, bad, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->stream));
    PFD_CUDA(h, cudaStreamSynchronize(h->stream));
    return PFD_OK;
}

// core.fillnodata_upstream in place: `has(v)` says which cells hold a value of their own
template <typename U, class Has>
static int fill_up_paths(pfd_handle* h, U* data_dev, Has has) {
    return hd_solve(h, FillSrc<U, Has>{data_dev, has}, FillOut<U>{data_dev});
}
template <typename U>
struct HasNonZero {
    __device__ __forceinline__ bool operator()(U v) const { return v != (U)0; }
};
template <typename T>
struct HasData {
    NoData nd;
    __device__ __forceinline__ bool operator()(T v) const { return not_nodata(v, nd); }
};

// after a REJECTED path-sum attempt the hop-by-hop engine wrote `out`; the path structure of the attempt still says which cells
// reach a pit
static int hand_mask_unreached(pfd_handle* h, double* out_dev) {
    const long long nrow = h->nrow, ncol = h->ncol;
    const long long ntx = (ncol + TL_W - 1) / TL_W, nty = (nrow + TL_H - 1) / TL_H;
    const long long nslots = (nty + 2) * ntx * TL_RING;
    double* sval1 = (double*)h->hand_slots.p + nslots;
    uint32_t* snxt[2] = {(uint32_t*)(sval1 + nslots), (uint32_t*)(sval1 + nslots) + nslots};
    hand_mask_unreached_kernel<<<dim3((unsigned)ntx, (unsigned)nty), 1024, 0, h->stream>>>((const uint16_t*)h->hand_root.p, snxt[h->hand_fin], nrow, ncol,
                                                                                            ntx, out_dev);
    PFD_LAUNCH_CHECK(h);
    return PFD_OK;
}
