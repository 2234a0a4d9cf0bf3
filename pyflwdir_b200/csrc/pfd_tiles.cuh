// pfd_tiles.cuh -- tile-hierarchical solver for the order-independent INTEGER outputs of the hot path:
//   rank            (core.rank, pyflwdir/core.py:17-47)
//   basins()        (basins.basins with all pits as outlets, pyflwdir/basins.py:12-18)
//   upstream_area() (streams.accuflux over int32 ones, pyflwdir/streams.py:15-41 via pyflwdir.py:790-800)
// These three depend only on the tree structure (distance to the pit, id of the pit, subtree size), and int32
// sums wrap associatively, so they can be re-associated without changing a bit. That allows a traffic-optimal
// three-phase scheme instead of one scattered access per cell and level:
//
//  Phase A (one CTA per 64x64 tile, everything in shared memory):
//     local pointer doubling along the downstream links until every cell knows its local terminal (a pit in the
//     tile, or the cell where its path EXITS the tile) and its hop distance; a doubling accumulate
//     (A_{k+1}[anc_2^k(d)] += A_k[d] for cells whose 2^k-th ancestor exists) gives the in-tile subtree size.
//     Border ("ring") cells publish a reduced-graph node: next node = the ring cell of the neighbouring tile
//     where their path enters it, cell hops to it, or the pit they end in. Exit cells add their subtree size
//     to the inflow weight W of the entry cell they drain into.
//  Phase B (global, ~N/16 nodes): the same doubling on the reduced graph -> for every ring cell its rank, basin
//     id and total inflow from outside its tile.
//  Phase C (one CTA per tile): repeat the local solve with ring cells pre-loaded with their inflow; combine with
//     the reduced-graph solution of the exit target and write rank / basins / uparea with coalesced stores.
//
// HBM traffic: 1 B/cell read in A, 1 B/cell read + 12 B/cell written in C, ~2 B/cell for the reduced graph.
#pragma once
#include "pfd_common.cuh"

#define TL_H 64
#define TL_W 64
#define TL_CELLS (TL_H * TL_W)
#define TL_THREADS 256
#define TL_CPT (TL_CELLS / TL_THREADS)  // 16 cells per thread: fixed column, rows tid/64 + 4*it
#define TL_RING 256                     // ring slots per tile (252 used)
#define TL_MAXROUNDS 13                 // 2^12 = 4096 >= longest simple path in a tile (+1 accumulate round)

#define SLOT_INVALID 0x7FFFFFFFu  // node does not drain to a pit (loop) / unused
#define TERM_PIT 0x80000000u      // term = TERM_PIT | pit ordinal: node's local path ends in that pit

struct TileSlots {
    uint32_t* nxt[2];   // next node (self when last)
    uint32_t* rh[2];    // reduced hops to nxt, saturating doubling: min(2^k, distance to last node)
    uint32_t* ch[2];    // cell hops to nxt
    uint32_t* acc[2];   // inflow accumulate (starts as W)
    uint32_t* term;     // TERM_PIT|ord for last nodes, SLOT_INVALID for nodes on a local loop / nodata
    uint32_t* term_h;   // cell hops from a last node to its pit
    int32_t* rank;      // results: rank of the ring cell (-1 invalid)
    uint32_t* basin;    // basin id of the ring cell (0 invalid)
    unsigned int* flag; // "another round needed"
    long long nslots;
};

__host__ __device__ __forceinline__ int tl_ring_pos(int ly, int lx) {
    if (ly == 0) return lx;
    if (ly == TL_H - 1) return TL_W + lx;
    if (lx == 0) return 2 * TL_W + (ly - 1);
    return 2 * TL_W + (TL_H - 2) + (ly - 1);  // lx == TL_W-1
}

__device__ __forceinline__ bool tl_on_ring(int ly, int lx) {
    return ly == 0 || ly == TL_H - 1 || lx == 0 || lx == TL_W - 1;
}

__device__ __forceinline__ uint32_t tl_slot_of(long long r, long long c, long long ntx) {
    const long long ty = r / TL_H, tx = c / TL_W;
    return (uint32_t)((ty * ntx + tx) * TL_RING + tl_ring_pos((int)(r % TL_H), (int)(c % TL_W)));
}

// binary search of `cell` in the ascending pit list -> ordinal
__device__ __forceinline__ uint32_t tl_pit_ordinal(const cell_t* __restrict__ pits, long long npits, cell_t cell) {
    long long lo = 0, hi = npits - 1;
    while (lo < hi) {
        const long long mid = (lo + hi) >> 1;
        if (__ldg(pits + mid) < cell) lo = mid + 1;
        else hi = mid;
    }
    return (uint32_t)lo;
}

struct TileShared {
    uint16_t nxt[TL_CELLS];
    uint16_t hops[TL_CELLS];
    uint32_t A[2][TL_CELLS];
};

// Local solve shared by phases A and C. On return: s.nxt = local terminal of every cell, s.hops = hop distance to
// it, s.A[abuf] = subtree sum of the initial weights w0 inside the tile; bit `it` of inv_mask is set for owned
// cells that never reach a terminal (they sit on / above a loop inside the tile).
// Per round k (Jacobi): A_{k+1}[anc(d)] += A_k[d] for cells whose 2^k-th ancestor exists (hops == 2^k, hops being
// min(2^k, distance to the terminal)), then nxt <- nxt[nxt], hops += hops[nxt].
__device__ __forceinline__ bool tl_is_terminal(const TileShared& s, uint32_t i) { return s.nxt[i] == i && s.hops[i] == 0; }

__device__ __forceinline__ void tl_local_solve(TileShared& s, const uint32_t* dirs /*[4] packed 16 bytes*/,
                                               const uint32_t* w0 /*[16]*/, int& abuf, uint32_t& inv_mask) {
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + 4 * it;
        const int i = ly * TL_W + lx;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        int ni = i;
        if (d < 8u) {
            const int y = ly + pfd_slot_dr((int)d), x = lx + pfd_slot_dc((int)d);
            if (y >= 0 && y < TL_H && x >= 0 && x < TL_W) ni = y * TL_W + x;
        }
        s.nxt[i] = (uint16_t)ni;
        s.hops[i] = (ni != i) ? 1 : 0;
        s.A[0][i] = w0[it];
    }
    __syncthreads();
    abuf = 0;
    for (int k = 0; k < TL_MAXROUNDS; ++k) {
        uint32_t pk[TL_CPT];  // nxt[nxt] | hops[nxt] << 16
        int more = 0;
        const uint32_t two_k = 1u << k;
#pragma unroll
        for (int it = 0; it < TL_CPT; ++it) {
            const int i = (ly0 + 4 * it) * TL_W + lx;
            const uint32_t n = s.nxt[i], h = s.hops[i];
            const uint32_t n2 = s.nxt[n], h2 = s.hops[n];
            pk[it] = n2 | (h2 << 16);
            s.A[abuf ^ 1][i] = s.A[abuf][i];
            more |= (h2 != 0u) | ((h + h2) == (two_k << 1));
        }
        more = __syncthreads_or(more);
#pragma unroll
        for (int it = 0; it < TL_CPT; ++it) {
            const int i = (ly0 + 4 * it) * TL_W + lx;
            const uint32_t n = s.nxt[i], h = s.hops[i];
            if (h == two_k) atomicAdd(&s.A[abuf ^ 1][n], s.A[abuf][i]);
            s.nxt[i] = (uint16_t)(pk[it] & 0xFFFFu);
            s.hops[i] = (uint16_t)min(h + (pk[it] >> 16), 0xFFFFu);
        }
        __syncthreads();
        abuf ^= 1;
        if (!more) break;
    }
    inv_mask = 0;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int i = (ly0 + 4 * it) * TL_W + lx;
        if (!tl_is_terminal(s, s.nxt[i])) inv_mask |= 1u << it;
    }
}

// load the 16 direction bytes this thread owns (rows ly0 + 4*it, column lx) -> 4 packed words; cells outside the
// raster read as nodata
__device__ __forceinline__ void tl_load_dirs(const uint8_t* __restrict__ dir, long long nrow, long long ncol,
                                             long long r0, long long c0, uint32_t* dirs) {
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
    const long long c = c0 + lx;
    dirs[0] = dirs[1] = dirs[2] = dirs[3] = 0;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const long long r = r0 + ly0 + 4 * it;
        uint32_t d = PFD_DIR_NODATA;
        if (r < nrow && c < ncol) d = __ldg(dir + r * ncol + c);
        dirs[it >> 2] |= d << (8 * (it & 3));
    }
}

// terminal descriptor of an owned cell that is a local terminal: exit -> slot id of the target ring cell,
// pit -> TERM_PIT | ordinal. (valid, non-nodata cells only)
__device__ __forceinline__ uint32_t tl_terminal_info(uint32_t d, long long r, long long c, long long ncol, long long ntx,
                                                     const cell_t* __restrict__ pits, long long npits) {
    if (d < 8u) return tl_slot_of(r + pfd_slot_dr((int)d), c + pfd_slot_dc((int)d), ntx);
    return TERM_PIT | tl_pit_ordinal(pits, npits, (cell_t)(r * ncol + c));
}

// ---------------------------------------------------------------------------------------------------------
// Phase A
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TL_THREADS, 4) tile_phase_a_kernel(const uint8_t* __restrict__ dir, long long nrow,
                                                                     long long ncol, long long ntx,
                                                                     const cell_t* __restrict__ pits, long long npits,
                                                                     TileSlots S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared& s = *reinterpret_cast<TileShared*>(smem_raw);
    const long long tile = (long long)blockIdx.y * ntx + blockIdx.x;
    const long long r0 = (long long)blockIdx.y * TL_H, c0 = (long long)blockIdx.x * TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;

    uint32_t dirs[4], w0[TL_CPT];
    tl_load_dirs(dir, nrow, ncol, r0, c0, dirs);
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) w0[it] = (((dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu) != PFD_DIR_NODATA) ? 1u : 0u;
    int abuf;
    uint32_t inv;
    tl_local_solve(s, dirs, w0, abuf, inv);

    // terminal descriptors go into the spare accumulate buffer
    uint32_t* tinfo = s.A[abuf ^ 1];
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + 4 * it;
        const int i = ly * TL_W + lx;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        if (d != PFD_DIR_NODATA && tl_is_terminal(s, i)) {  // local terminal (exit cell or pit)
            const uint32_t ti = tl_terminal_info(d, r0 + ly, c0 + lx, ncol, ntx, pits, npits);
            tinfo[i] = ti;
            if (!(ti & TERM_PIT)) atomicAdd(S.acc[0] + ti, s.A[abuf][i]);  // inflow weight of the entry cell
        }
    }
    __syncthreads();
    // ring cells publish their reduced-graph node
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + 4 * it;
        if (!tl_on_ring(ly, lx)) continue;
        const int i = ly * TL_W + lx;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        const uint32_t slot = (uint32_t)(tile * TL_RING + tl_ring_pos(ly, lx));
        uint32_t nx = slot, rh = 0, ch = 0, term = SLOT_INVALID, th = 0;
        if (d != PFD_DIR_NODATA && !((inv >> it) & 1u)) {
            const uint32_t root = s.nxt[i];
            const uint32_t ti = tinfo[root];
            const uint32_t dist = s.hops[i];
            if (ti & TERM_PIT) {
                term = ti;
                th = dist;
            } else {
                nx = ti;
                rh = 1;
                ch = dist + 1;
                term = 0;
            }
        }
        S.nxt[0][slot] = nx;
        S.rh[0][slot] = rh;
        S.ch[0][slot] = ch;
        S.term[slot] = term;
        S.term_h[slot] = th;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Phase B: synchronous doubling rounds over the ring slots
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) slots_round_kernel(TileSlots S, int src, uint32_t two_k) {
    const int dst = src ^ 1;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S.nslots; s += (long long)gridDim.x * blockDim.x) {
        const uint32_t n = S.nxt[src][s];
        const uint32_t h = S.rh[src][s];
        const uint32_t a = S.acc[src][s];
        if (a) atomicAdd(S.acc[dst] + s, a);
        if (n == (uint32_t)s) {  // last node / unused: nothing to jump
            S.nxt[dst][s] = n;
            S.rh[dst][s] = h;
            S.ch[dst][s] = S.ch[src][s];
            continue;
        }
        if (h == two_k && a) atomicAdd(S.acc[dst] + n, a);
        const uint32_t n2 = S.nxt[src][n];
        const uint32_t h2 = S.rh[src][n];
        S.nxt[dst][s] = n2;
        S.rh[dst][s] = h + h2;  // saturating in effect: h2 == 0 once n is a last node
        S.ch[dst][s] = S.ch[src][s] + S.ch[src][n];
        if (h2 != 0u || (h + h2) == (two_k << 1)) *S.flag = 1u;
    }
}

__global__ void __launch_bounds__(256) slots_finalize_kernel(TileSlots S, int src) {
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < S.nslots; s += (long long)gridDim.x * blockDim.x) {
        const uint32_t last = S.nxt[src][s];
        const uint32_t t = S.term[last];
        int32_t rk = -1;
        uint32_t b = 0;
        if ((t & TERM_PIT) && S.nxt[src][last] == last && S.term[s] != SLOT_INVALID) {
            rk = (int32_t)(S.ch[src][s] + S.term_h[last]);
            b = (t & ~TERM_PIT) + 1u;
        }
        S.rank[s] = rk;
        S.basin[s] = b;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Phase C
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TL_THREADS, 4) tile_phase_c_kernel(const uint8_t* __restrict__ dir, long long nrow,
                                                                     long long ncol, long long ntx,
                                                                     const cell_t* __restrict__ pits, long long npits,
                                                                     TileSlots S, int src, int32_t* __restrict__ rank_out,
                                                                     uint32_t* __restrict__ basin_out,
                                                                     int32_t* __restrict__ uparea_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    TileShared& s = *reinterpret_cast<TileShared*>(smem_raw);
    const long long tile = (long long)blockIdx.y * ntx + blockIdx.x;
    const long long r0 = (long long)blockIdx.y * TL_H, c0 = (long long)blockIdx.x * TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;

    uint32_t dirs[4], w0[TL_CPT];
    tl_load_dirs(dir, nrow, ncol, r0, c0, dirs);
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + 4 * it;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        uint32_t w = (d != PFD_DIR_NODATA) ? 1u : 0u;
        if (w && tl_on_ring(ly, lx)) w += S.acc[src][tile * TL_RING + tl_ring_pos(ly, lx)];  // inflow from outside
        w0[it] = w;
    }
    int abuf;
    uint32_t inv;
    tl_local_solve(s, dirs, w0, abuf, inv);

    // per terminal: rank at the terminal first, then (second pass through the same spare buffer) its basin id
    uint32_t* tbuf = s.A[abuf ^ 1];
    uint32_t t_b[TL_CPT];
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + 4 * it;
        const int i = ly * TL_W + lx;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        t_b[it] = 0;
        if (d != PFD_DIR_NODATA && tl_is_terminal(s, i)) {
            uint32_t rk;
            if (d < 8u) {  // exit cell: one hop above the entry cell of the neighbouring tile
                const uint32_t slot = tl_slot_of(r0 + ly + pfd_slot_dr((int)d), c0 + lx + pfd_slot_dc((int)d), ntx);
                const int32_t rs = S.rank[slot];
                rk = (rs < 0) ? 0xFFFFFFFFu : (uint32_t)(rs + 1);
                t_b[it] = (rs < 0) ? 0u : S.basin[slot];
            } else {
                rk = 0;
                t_b[it] = tl_pit_ordinal(pits, npits, (cell_t)((r0 + ly) * ncol + c0 + lx)) + 1u;
            }
            tbuf[i] = rk;
        }
    }
    __syncthreads();
    int32_t rks[TL_CPT];
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int i = (ly0 + 4 * it) * TL_W + lx;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        int32_t rk = -9999;
        if (d != PFD_DIR_NODATA) {
            rk = -1;
            if (!((inv >> it) & 1u)) {
                const uint32_t tr = tbuf[s.nxt[i]];
                if (tr != 0xFFFFFFFFu) rk = (int32_t)(tr + s.hops[i]);
            }
        }
        rks[it] = rk;
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int i = (ly0 + 4 * it) * TL_W + lx;
        const uint32_t d = (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu;
        if (d != PFD_DIR_NODATA && tl_is_terminal(s, i)) tbuf[i] = t_b[it];
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + 4 * it;
        const long long r = r0 + ly, c = c0 + lx;
        if (r >= nrow || c >= ncol) continue;
        const int i = ly * TL_W + lx;
        const int32_t rk = rks[it];
        int32_t up = (rk == -9999) ? -9999 : 1;
        uint32_t b = 0;
        if (rk >= 0) {
            b = tbuf[s.nxt[i]];
            up = (int32_t)s.A[abuf][i];
        }
        const long long g = r * ncol + c;
        if (rank_out) rank_out[g] = rk;
        if (basin_out) basin_out[g] = b;
        if (uparea_out) uparea_out[g] = up;
    }
}
