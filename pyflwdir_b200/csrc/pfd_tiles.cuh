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
// Block shapes (tuned on B200, see profiles/): both phases run best with 1024 threads x 2 CTAs/SM (4 cells per
// thread keep the per-thread arrays in 32 registers without spills).
#ifndef TLA_THREADS
#define TLA_THREADS 1024
#endif
#ifndef TLA_MINBLOCKS
#define TLA_MINBLOCKS 2
#endif
#ifndef TLC_THREADS
#define TLC_THREADS 1024
#endif
#ifndef TLC_MINBLOCKS
#define TLC_MINBLOCKS 2
#endif
#define TL_RING 256                       // ring slots per tile (252 used)
#define TL_MAXROUNDS 13                   // 2^12 = 4096 >= longest simple path in a tile (+1 accumulate round)

#define SLOT_INVALID 0x7FFFFFFFu  // node does not drain to a pit (loop) / unused
#define TERM_PIT 0x80000000u      // term = TERM_PIT | pit ordinal: node's local path ends in that pit

__host__ __device__ __forceinline__ int tl_ring_pos(int ly, int lx) {
    if (ly == 0) return lx;
    if (ly == TL_H - 1) return TL_W + lx;
    if (lx == 0) return 2 * TL_W + (ly - 1);
    return 2 * TL_W + (TL_H - 2) + (ly - 1);  // lx == TL_W-1
}

__device__ __forceinline__ bool tl_on_ring(int ly, int lx) {
    return ly == 0 || ly == TL_H - 1 || lx == 0 || lx == TL_W - 1;
}

// inverse of tl_ring_pos: local cell index of ring position rp (0 .. 4*TL_W-5)
__device__ __forceinline__ int tl_ring_cell(int rp) {
    if (rp < TL_W) return rp;
    if (rp < 2 * TL_W) return (TL_H - 1) * TL_W + (rp - TL_W);
    if (rp < 2 * TL_W + TL_H - 2) return (rp - 2 * TL_W + 1) * TL_W;
    return (rp - 2 * TL_W - (TL_H - 2) + 1) * TL_W + (TL_W - 1);
}
#define TL_NRING (2 * TL_W + 2 * (TL_H - 2))

// slot of the cell one step in direction d outside the tile `tile` (slot-array tile index incl. the halo row shift)
// from local cell (ly, lx): 32-bit arithmetic only
__device__ __forceinline__ uint32_t tl_exit_slot(uint32_t tile, uint32_t ntx, int ly, int lx, uint32_t d) {
    int y = ly + pfd_slot_dr((int)d), x = lx + pfd_slot_dc((int)d);
    uint32_t t = tile;
    if (y < 0) {
        t -= ntx;
        y += TL_H;
    } else if (y >= TL_H) {
        t += ntx;
        y -= TL_H;
    }
    if (x < 0) {
        t -= 1;
        x += TL_W;
    } else if (x >= TL_W) {
        t += 1;
        x -= TL_W;
    }
    return t * TL_RING + (uint32_t)tl_ring_pos(y, x);
}

// Ring-slot arrays cover the local tile rows PLUS one halo tile row above and below (tile row index shifted by
// one): exits across the top / bottom edge of a row block (multi-GPU row tiling) land in halo slots, which act as
// terminals of the local reduced graph. r may be -1 or nrow (one row outside the block).
__device__ __forceinline__ uint32_t tl_slot_of(long long r, long long c, long long ntx) {
    const long long ty = (r + TL_H) / TL_H, tx = c / TL_W;  // = floor(r / TL_H) + 1 for r >= -TL_H
    return (uint32_t)((ty * ntx + tx) * TL_RING + tl_ring_pos((int)((r + TL_H) % TL_H), (int)(c % TL_W)));
}

// basin id (= pit ordinal + 1) of every pit cell, pre-written at the pit's own position of the basin output buffer
// so that the tile kernels find it with one load of a cell they own (no search in the sorted pit list)
__global__ void stash_pit_ids_kernel(const cell_t* __restrict__ pits, long long npits, long long cell_off,
                                     unsigned long long id_off, uint32_t* __restrict__ basin) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < npits; k += (long long)gridDim.x * blockDim.x)
        basin[(long long)pits[k] - cell_off] = (uint32_t)(id_off + (unsigned long long)k + 1ull);
}

// Shared state of one tile in phase A: P packs (next cell : 12 bits | hops to it : 20 bits), A is the accumulate
// buffer.
struct TileShared {
    uint32_t P[TL_CELLS];
    uint32_t A[TL_CELLS];
};
#define TP_PACK(n, h) ((uint32_t)(n) | ((uint32_t)(h) << 12))
#define TP_N(p) ((p) & 0xFFFu)
#define TP_H(p) ((p) >> 12)
#define TL_LOC_INVALID 0xFFFFFFFFu

__device__ __forceinline__ uint32_t tl_dir_of(const uint32_t* dirs, int it) { return (dirs[it >> 2] >> (8 * (it & 3))) & 0xFFu; }

// local index of the next cell inside the tile, or i itself when the link leaves the tile / the cell is a pit
__device__ __forceinline__ int tl_local_next(int i, uint32_t d) {
    if (d >= 8u) return i;
    const int y = (i >> 6) + pfd_slot_dr((int)d), x = (i & (TL_W - 1)) + pfd_slot_dc((int)d);
    return (y >= 0 && y < TL_H && x >= 0 && x < TL_W) ? y * TL_W + x : i;
}

// Local solve of phase A. On return own[it] (mirrored in s.P) = packed (local terminal, hop distance) of every owned
// cell and s.A[i] = subtree sum of the unit weights inside the tile.
// Round k (Jacobi): every cell whose 2^k-th ancestor exists (hops == 2^k) snapshots its A and its ancestor's P,
// then -- after a barrier -- adds the snapshot to that ancestor (A_{k+1}[anc] += A_k[d]) and jumps
// (next <- next[next], hops += hops[next]). A cell stays active only while its hop count is an exact power of two,
// i.e. its chain is not exhausted; the loop ends when no cell of the tile is active (<= 12 rounds without loops).
template <int THREADS>
__device__ __forceinline__ void tl_local_solve(TileShared& s, const uint32_t* dirs, uint32_t* own) {
    constexpr int TL_CPT = TL_CELLS / THREADS, TL_RPI = THREADS / TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
    uint32_t active = 0;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int i = (ly0 + TL_RPI * it) * TL_W + lx;
        const uint32_t d = tl_dir_of(dirs, it);
        const int ni = tl_local_next(i, d);
        own[it] = TP_PACK(ni, ni != i ? 1 : 0);
        if (ni != i) active |= 1u << it;
        s.P[i] = own[it];
        s.A[i] = (d != PFD_DIR_NODATA) ? 1u : 0u;
    }
    __syncthreads();
    for (int k = 0; k < TL_MAXROUNDS; ++k) {
        const uint32_t two_k = 1u << k;
        uint32_t pn[TL_CPT], a[TL_CPT];
#pragma unroll
        for (int it = 0; it < TL_CPT; ++it) {
            if (active & (1u << it)) {
                pn[it] = s.P[TP_N(own[it])];
                a[it] = s.A[(ly0 + TL_RPI * it) * TL_W + lx];
            }
        }
        __syncthreads();  // every snapshot is taken before any update
        uint32_t next_active = 0;
#pragma unroll
        for (int it = 0; it < TL_CPT; ++it) {
            if (active & (1u << it)) {
                atomicAdd(&s.A[TP_N(own[it])], a[it]);
                const uint32_t h = two_k + TP_H(pn[it]);
                own[it] = TP_PACK(TP_N(pn[it]), h);
                s.P[(ly0 + TL_RPI * it) * TL_W + lx] = own[it];
                if (h == (two_k << 1)) next_active |= 1u << it;
            }
        }
        active = next_active;
        if (!__syncthreads_or((int)active)) break;
    }
}

// load the direction bytes this thread owns (rows ly0 + TL_RPI*it, column lx) packed 4 per word; cells outside the
// raster read as nodata
template <int THREADS>
__device__ __forceinline__ void tl_load_dirs(const uint8_t* __restrict__ dir, long long nrow, long long ncol,
                                             long long r0, long long c0, uint32_t* dirs) {
    constexpr int TL_CPT = TL_CELLS / THREADS, TL_RPI = THREADS / TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
    const long long c = c0 + lx;
#pragma unroll
    for (int w = 0; w < (TL_CPT + 3) / 4; ++w) dirs[w] = 0;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const long long r = r0 + ly0 + TL_RPI * it;
        uint32_t d = PFD_DIR_NODATA;
        if (r < nrow && c < ncol) d = __ldg(dir + r * ncol + c);
        dirs[it >> 2] |= d << (8 * (it & 3));
    }
}

// ---------------------------------------------------------------------------------------------------------
// Fused parse (single-GPU headline path): the tile's raw D8 codes plus a one-cell halo are staged in shared
// memory (32-bit loads for the interior, out-of-raster cells read as 247), every thread derives the `dir` byte
// of its own cells exactly like parse_kernel (core_d8.py:42-67: pit / forced pit when the downstream cell is
// nodata or off the raster / nodata; illegal codes raise the invalid flag, core_d8.py:115-122) and writes it to
// global memory, so that the separate parse pass over the raster disappears from pfd_d8_flow_all.
// ---------------------------------------------------------------------------------------------------------
#define TLF_STRIDE 72   // bytes per staged row: 3 pad | left halo | 64 cells | right halo | 3 pad
#define TLF_X0 4        // byte offset of the tile's first column inside a staged row (word aligned)
struct TileCodes {
    uint8_t c[(TL_H + 2) * TLF_STRIDE];
};

template <int THREADS>
__device__ __forceinline__ void tl_stage_codes(TileCodes& sc, const uint8_t* __restrict__ d8, long long nrow, long long ncol,
                                               long long r0, long long c0, bool al4) {
    // interior: 64 rows x 16 words
    for (int w = threadIdx.x; w < TL_H * (TL_W / 4); w += THREADS) {
        const int row = w >> 4, wx = w & 15;
        const long long r = r0 + row, c = c0 + 4 * wx;
        uint32_t v = 0xF7F7F7F7u;
        if (r < nrow && c < ncol) {
            const uint8_t* p = d8 + r * ncol + c;
            if (al4) {
                v = __ldg(reinterpret_cast<const uint32_t*>(p));  // ncol % 4 == 0: the word never straddles the row end
            } else {
                v = 0;
#pragma unroll
                for (int b = 0; b < 4; ++b) v |= ((c + b < ncol) ? (uint32_t)__ldg(p + b) : 247u) << (8 * b);
            }
        }
        *reinterpret_cast<uint32_t*>(&sc.c[(row + 1) * TLF_STRIDE + TLF_X0 + 4 * wx]) = v;
    }
    // halo: top / bottom rows (66 cells each), left / right columns (64 cells each)
    for (int k = threadIdx.x; k < 2 * (TL_W + 2) + 2 * TL_H; k += THREADS) {
        int sy, sx;  // staged coordinates: row 0..65, column -1..64
        if (k < TL_W + 2) {
            sy = 0;
            sx = k - 1;
        } else if (k < 2 * (TL_W + 2)) {
            sy = TL_H + 1;
            sx = k - (TL_W + 2) - 1;
        } else if (k < 2 * (TL_W + 2) + TL_H) {
            sy = k - 2 * (TL_W + 2) + 1;
            sx = -1;
        } else {
            sy = k - 2 * (TL_W + 2) - TL_H + 1;
            sx = TL_W;
        }
        const long long r = r0 + sy - 1, c = c0 + sx;
        uint8_t v = 247;
        if (r >= 0 && r < nrow && c >= 0 && c < ncol) v = __ldg(d8 + r * ncol + c);
        sc.c[sy * TLF_STRIDE + TLF_X0 + sx] = v;
    }
}

// D8 code -> neighbour slot (codes are powers of two): log2 E0 SE1 S2 SW3 W4 NW5 N6 NE7 -> slot 4 7 6 5 3 0 1 2
__device__ __forceinline__ uint32_t tl_code_slot(uint32_t code) { return (0x21035674u >> (4 * (__ffs((int)code) - 1))) & 7u; }

// dir bytes of the thread's own cells from the staged codes (packed 4 per word like tl_load_dirs); returns false
// when one of them is not a legal D8 code
template <int THREADS>
__device__ __forceinline__ bool tl_parse_dirs(const TileCodes& sc, uint32_t* dirs) {
    constexpr int TL_CPT = TL_CELLS / THREADS, TL_RPI = THREADS / TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
    bool ok = true;
#pragma unroll
    for (int w = 0; w < (TL_CPT + 3) / 4; ++w) dirs[w] = 0;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + TL_RPI * it;
        const int at = (ly + 1) * TLF_STRIDE + TLF_X0 + lx;
        const uint32_t code = sc.c[at];
        uint32_t d;
        if (code == 247u) {
            d = PFD_DIR_NODATA;
        } else if (code == 0u || code == 255u) {
            d = PFD_DIR_PIT;
        } else if (code & (code - 1u)) {
            ok = false;
            d = PFD_DIR_NODATA;
        } else {
            const uint32_t k = tl_code_slot(code);
            const uint32_t nb = sc.c[at + pfd_slot_dr((int)k) * TLF_STRIDE + pfd_slot_dc((int)k)];
            d = (nb == 247u) ? (uint32_t)PFD_DIR_FPIT : k;
        }
        dirs[it >> 2] |= d << (8 * (it & 3));
    }
    return ok;
}

// ---------------------------------------------------------------------------------------------------------
// Phase A: local solve; per cell (local terminal, hops) -> loc[], in-tile subtree size -> cnt[]; ring nodes; W
// ---------------------------------------------------------------------------------------------------------
// FUSED = false: `dir` is the parsed direction raster, pit terminals carry the pit ordinal found in `pit_ids`.
// FUSED = true : `d8` holds raw D8 codes; the kernel derives the directions itself, writes them to `dir_out`
//                (and flags illegal codes); pit terminals carry TERM_PIT | local cell index, resolved to the pit
//                ordinal by slots_finalize_kernel once the pits have been numbered.
template <int THREADS, int MINBLOCKS, bool FUSED>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
    tile_phase_a_kernel(const uint8_t* __restrict__ dir, long long nrow, long long ncol, long long ntx,
                        const uint32_t* pit_ids, uint2* __restrict__ loccnt, uint32_t* __restrict__ W, uint32_t* __restrict__ s_nxt, uint32_t* __restrict__ s_rh,
                        uint32_t* __restrict__ s_ch, uint32_t* __restrict__ s_term, uint32_t* __restrict__ s_term_h,
                        const uint8_t* __restrict__ d8, uint8_t* __restrict__ dir_out, unsigned int* __restrict__ invalid_flag,
                        int al4) {
    constexpr int TL_CPT = TL_CELLS / THREADS, TL_RPI = THREADS / TL_W;
    __shared__ TileShared s;
    __shared__ TileCodes sc;  // FUSED only (the compiler drops it otherwise)
    const uint32_t tile = (uint32_t)(((long long)blockIdx.y + 1) * ntx + blockIdx.x);  // +1: halo tile row
    const long long r0 = (long long)blockIdx.y * TL_H, c0 = (long long)blockIdx.x * TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
    const long long g00 = r0 * ncol + c0;  // global index of the tile's first cell

    uint32_t dirs[(TL_CPT + 3) / 4], own[TL_CPT];
    if (FUSED) {
        tl_stage_codes<THREADS>(sc, d8, nrow, ncol, r0, c0, al4 != 0);
        __syncthreads();
        if (!tl_parse_dirs<THREADS>(sc, dirs)) atomicOr(invalid_flag, 1u);
        __syncthreads();  // every neighbour code has been read: the staged codes may now be replaced by the dirs
#pragma unroll
        for (int it = 0; it < TL_CPT; ++it) {
            const int ly = ly0 + TL_RPI * it;
            const uint32_t d = tl_dir_of(dirs, it);
            sc.c[(ly + 1) * TLF_STRIDE + TLF_X0 + lx] = (uint8_t)d;  // read back by the ring threads below
            if (r0 + ly < nrow && c0 + lx < ncol) dir_out[g00 + (long long)ly * ncol + lx] = (uint8_t)d;
        }
    } else {
        tl_load_dirs<THREADS>(dir, nrow, ncol, r0, c0, dirs);
    }
    tl_local_solve<THREADS>(s, dirs, own);

    // (1) per-cell results for phase C
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + TL_RPI * it;
        const int i = ly * TL_W + lx;
        const uint32_t root = TP_N(own[it]);
        const bool inv = s.P[root] != root;  // a terminal is (next = itself, hops = 0)
        if (r0 + ly < nrow && c0 + lx < ncol) {
            const long long g = g00 + (long long)ly * ncol + lx;
            loccnt[g] = make_uint2(inv ? TL_LOC_INVALID : own[it], s.A[i]);  // (local terminal | hops, in-tile count)
        }
    }
    __syncthreads();
    // (2) terminals publish their descriptor through their A slot: pits by their owner thread, exit cells (always
    //     ring cells) by one thread per ring position, which also hands the exit cell's in-tile subtree size to the
    //     entry cell it drains into
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const uint32_t d = tl_dir_of(dirs, it);
        if (d == PFD_DIR_PIT || d == PFD_DIR_FPIT) {
            const int ly = ly0 + TL_RPI * it;
            if (FUSED) s.A[ly * TL_W + lx] = TERM_PIT | (uint32_t)(ly * TL_W + lx);
            else s.A[ly * TL_W + lx] = TERM_PIT | (pit_ids ? (pit_ids[g00 + (long long)ly * ncol + lx] - 1u) : 0u);
        }
    }
    int ri = -1;
    uint32_t rd = PFD_DIR_NODATA;
    if (threadIdx.x < TL_NRING) {
        ri = tl_ring_cell(threadIdx.x);
        const int rly = ri >> 6, rlx = ri & (TL_W - 1);
        if (FUSED) rd = sc.c[(rly + 1) * TLF_STRIDE + TLF_X0 + rlx];  // out-of-raster cells were staged as nodata
        else if (r0 + rly < nrow && c0 + rlx < ncol) rd = __ldg(dir + g00 + (long long)rly * ncol + rlx);  // L1/L2 hit
        if (rd < 8u && s.P[ri] == (uint32_t)ri) {  // exit cell
            const uint32_t ti = tl_exit_slot(tile, (uint32_t)ntx, ri >> 6, ri & (TL_W - 1), rd);
            atomicAdd(W + ti, s.A[ri]);
            s.A[ri] = ti;
        }
    }
    __syncthreads();
    // (3) ring cells publish their reduced-graph node
    if (ri >= 0) {
        const uint32_t slot = tile * TL_RING + threadIdx.x;
        uint32_t nx = slot, rh = 0, ch = 0, term = SLOT_INVALID, th = 0;
        const uint32_t p = s.P[ri];
        const uint32_t root = TP_N(p);
        if (rd != PFD_DIR_NODATA && s.P[root] == root) {
            const uint32_t ti = s.A[root];
            const uint32_t dist = TP_H(p);
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
        s_nxt[slot] = nx;
        s_rh[slot] = rh;
        s_ch[slot] = ch;
        s_term[slot] = term;
        s_term_h[slot] = th;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Phase B: synchronous doubling rounds over the ring slots
// ---------------------------------------------------------------------------------------------------------
struct SlotBuf {  // one side of the double-buffered reduced-graph state
    uint32_t* nxt;
    uint32_t* rh;   // reduced hops (low 30 bits) | SLOT_DONE1 | SLOT_DONE2
    uint32_t* ch;
    uint32_t* acc;  // inflow accumulate -- ONE array shared by both sides
};
#define SLOT_DONE1 0x40000000u  // final state stored on this side only
#define SLOT_DONE2 0x80000000u  // final state stored on both sides: the node is skipped from now on
#define SLOT_HMASK 0x3FFFFFFFu

struct SlotProtect {  // ring rows on the block's edges that must stay live in the first local solve (multi-rank)
    long long per_row;   // slots per tile row
    long long nty;
    int top, bot;
};

// All doubling rounds of a reduced graph in ONE cooperative launch.
//  * prologue: ring nodes that received no inflow weight are never referenced by any chain -> made inert;
//  * round k: every live node adds its accumulate to its 2^k-th successor (into the `recv` buffer of this round's
//    parity, folded into acc by the owner at the start of the next round: Jacobi without copying acc) and jumps;
//    a node whose chain is exhausted writes its final state to both sides once and is skipped afterwards, so late
//    rounds only stream rh + recv;
//  * the number of executed rounds K is left in *rounds_out: the final state of every node is on side K & 1.
__global__ void __launch_bounds__(256, 8) slots_solve_kernel(SlotBuf b0, SlotBuf b1, uint32_t* recv0, uint32_t* recv1,
                                                             long long nslots, unsigned int* flags, int* rounds_out,
                                                             int mark_inert, SlotProtect prot) {
    cg::grid_group grid = cg::this_grid();
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const long long stride = (long long)gridDim.x * blockDim.x;
    uint32_t* const acc = b0.acc;
    if (mark_inert) {
#pragma unroll 1
        for (long long s = tid; s < nslots; s += stride) {
            if (acc[s] == 0u && b0.nxt[s] != (uint32_t)s) {
                const long long trow = s / prot.per_row;
                const int rp = (int)(s % TL_RING);
                const bool keep = (prot.top && trow == 1 && rp < TL_W) ||
                                  (prot.bot && trow == prot.nty && rp >= TL_W && rp < 2 * TL_W);
                if (!keep) {
                    b0.nxt[s] = (uint32_t)s;
                    b0.rh[s] = 0u;
                }
            }
        }
        grid.sync();
    }
    int k = 0;
    for (; k < 30; ++k) {
        const SlotBuf S = (k & 1) ? b1 : b0, D = (k & 1) ? b0 : b1;
        uint32_t* const recv_prev = (k & 1) ? recv0 : recv1;  // filled in round k-1
        uint32_t* const recv_cur = (k & 1) ? recv1 : recv0;
        const uint32_t two_k = 1u << k;
        // three rotating flags: the one zeroed here was last READ at the end of round k-2, two grid.sync()s ago
        if (tid == 0) flags[(k + 1) % 3] = 0u;
        unsigned int again = 0;
#pragma unroll 2
        for (long long s = tid; s < nslots; s += stride) {
            const uint32_t rhw = S.rh[s];
            const uint32_t r = recv_prev[s];
            if (r) {
                acc[s] += r;
                recv_prev[s] = 0u;
            }
            if (rhw & SLOT_DONE2) continue;
            const uint32_t n = S.nxt[s];
            const uint32_t h = rhw & SLOT_HMASK;
            if ((rhw & SLOT_DONE1) || n == (uint32_t)s) {  // chain exhausted: replicate the final state once
                D.nxt[s] = n;
                D.rh[s] = h | SLOT_DONE2;
                D.ch[s] = S.ch[s];
                S.rh[s] = h | SLOT_DONE2;  // only flag bits change: concurrent readers mask them off
                continue;
            }
            if (h == two_k) {
                const uint32_t a = acc[s];
                if (a) atomicAdd(recv_cur + n, a);
            }
            const uint32_t n2 = S.nxt[n];
            const uint32_t nh = h + (S.rh[n] & SLOT_HMASK);
            const bool live = nh == (two_k << 1);
            D.nxt[s] = n2;
            D.rh[s] = nh | (live ? 0u : SLOT_DONE1);
            D.ch[s] = S.ch[s] + S.ch[n];
            again |= live ? 1u : 0u;
        }
        if (again) flags[k % 3] = 1u;
        grid.sync();
        if (*((volatile unsigned int*)&flags[k % 3]) == 0u) {
            ++k;
            break;
        }
    }
    // fold what the last round delivered
    uint32_t* const recv_last = ((k - 1) & 1) ? recv1 : recv0;
#pragma unroll 1
    for (long long s = tid; s < nslots; s += stride) {
        const uint32_t r = recv_last[s];
        if (r) {
            acc[s] += r;
            recv_last[s] = 0u;
        }
    }
    if (tid == 0) *rounds_out = k;
}

// rounds: number of rounds the solve executed (device) -> the final node state is on side (*rounds & 1)
// pit_stash != null (fused-parse path): pit terminals hold TERM_PIT | local cell index of the pit inside the tile
// of slot `last`; the basin id (pit ordinal + 1) is read from the pit's own cell of the basin buffer, where
// stash_pit_ids_kernel left it.
__global__ void __launch_bounds__(256) slots_finalize_kernel(SlotBuf b0, SlotBuf b1, const int* __restrict__ rounds,
                                                             const uint32_t* __restrict__ term,
                                                             const uint32_t* __restrict__ term_h, long long nslots,
                                                             int32_t* __restrict__ rank, uint32_t* __restrict__ basin,
                                                             const uint32_t* __restrict__ pit_stash, long long ncol, long long ntx) {
    const SlotBuf cur = (*rounds & 1) ? b1 : b0;
    for (long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x; s < nslots; s += (long long)gridDim.x * blockDim.x) {
        const uint32_t last = cur.nxt[s];
        const uint32_t t = term[last];
        int32_t rk = -1;
        uint32_t b = 0;
        if ((t & TERM_PIT) && cur.nxt[last] == last && term[s] != SLOT_INVALID) {
            rk = (int32_t)(cur.ch[s] + term_h[last]);
            if (pit_stash) {
                const long long tl = (long long)(last / TL_RING);  // slot-array tile index (halo tile row included)
                const long long ty = tl / ntx - 1, tx = tl % ntx;
                const uint32_t li = t & (uint32_t)(TL_CELLS - 1);
                b = __ldg(pit_stash + (ty * TL_H + (li >> 6)) * ncol + tx * TL_W + (li & (TL_W - 1)));
            } else {
                b = (t & ~TERM_PIT) + 1u;
            }
        }
        rank[s] = rk;
        basin[s] = b;
    }
}

// ---------------------------------------------------------------------------------------------------------
// Phase C: no second solve. acc = in-tile count (phase A) + the outside inflows of the tile's entry cells walked
// down their local paths (sparse: ~90 entry cells per tile); rank / basin = hops + solution of the local terminal.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tl_cp_async4(uint32_t* smem_dst, const uint32_t* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void tl_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct TileSharedC {
    uint32_t X[TL_CELLS];   // extra inflow per cell, later basin id per terminal
    uint32_t R[TL_CELLS];   // pit cells: stashed basin id (fetched with the first loads); later rank at the terminal
    uint16_t S[TL_CELLS];   // in-tile successor, for the walkers
    uint32_t wl_cell[TL_RING];  // walker list
    uint32_t wl_w[TL_RING];
    uint32_t ring_t[TL_RING];   // per ring position: rank at the terminal if the cell is an exit cell, else TL_NOT_EXIT
    uint32_t ring_b[TL_RING];   //                    basin id behind that exit
    uint32_t ring_w[TL_RING];   //                    outside inflow of the (entry) cell
    uint32_t wl_count;
};
#define TL_NOT_EXIT 0xFFFFFFFEu

// IDXMODE: 0 = no idxs_ds output, 1 = 32-bit (int32 / uint32 share the bit pattern), 2 = int64 -- the fused-parse path
// writes idxs_ds (core_d8.from_array, core_d8.py:42-67) from here, next to the other per-cell outputs.
template <int THREADS, int MINBLOCKS, int IDXMODE>
__global__ void __launch_bounds__(THREADS, MINBLOCKS)
    tile_phase_c_kernel(const uint8_t* __restrict__ dir, long long nrow, long long ncol, long long ntx,
                        const uint2* __restrict__ loccnt, const uint32_t* __restrict__ inflow, const int32_t* __restrict__ s_rank,
                        const uint32_t* __restrict__ s_basin, int32_t* __restrict__ rank_out, uint32_t* basin_out,
                        int32_t* __restrict__ uparea_out, void* __restrict__ idxs_out) {
    constexpr int TL_CPT = TL_CELLS / THREADS, TL_RPI = THREADS / TL_W;
    __shared__ TileSharedC s;
    const uint32_t tile = (uint32_t)(((long long)blockIdx.y + 1) * ntx + blockIdx.x);  // +1: halo tile row
    const long long r0 = (long long)blockIdx.y * TL_H, c0 = (long long)blockIdx.x * TL_W;
    const int lx = threadIdx.x & (TL_W - 1);
    const int ly0 = threadIdx.x >> 6;
    const long long g00 = r0 * ncol + c0;

    // one thread per ring position: everything it needs from global memory is requested up front, together with
    // the per-cell loads below (entry inflow; for exit cells the solution of the entry cell they drain into)
    // (the results are parked in shared memory so that they do not occupy registers across the kernel)
    if (threadIdx.x < TL_NRING) {
        const int ri = tl_ring_cell(threadIdx.x);
        const int ly = ri >> 6, lxr = ri & (TL_W - 1);
        uint32_t rw = 0, rt = TL_NOT_EXIT, rb = 0;
        if (r0 + ly < nrow && c0 + lxr < ncol) {
            const long long g = g00 + (long long)ly * ncol + lxr;
            const uint32_t rd = __ldg(dir + g);
            const uint32_t rloc = __ldg(&loccnt[g].x);
            if (uparea_out && rloc != TL_LOC_INVALID) rw = __ldg(inflow + tile * TL_RING + threadIdx.x);
            if (rd < 8u && rloc == (uint32_t)ri) {  // exit cell
                const uint32_t slot = tl_exit_slot(tile, (uint32_t)ntx, ly, lxr, rd);
                const int32_t rrank = __ldg(s_rank + slot);
                rt = (rrank < 0) ? 0xFFFFFFFFu : (uint32_t)(rrank + 1);
                rb = (rrank < 0) ? 0u : __ldg(s_basin + slot);
            }
        }
        s.ring_w[threadIdx.x] = rw;
        s.ring_t[threadIdx.x] = rt;
        s.ring_b[threadIdx.x] = rb;
    }
    uint32_t dirs[(TL_CPT + 3) / 4], own[TL_CPT], up[TL_CPT];
    tl_load_dirs<THREADS>(dir, nrow, ncol, r0, c0, dirs);
    if (threadIdx.x == 0) s.wl_count = 0;
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + TL_RPI * it;
        const int i = ly * TL_W + lx;
        const bool inside = r0 + ly < nrow && c0 + lx < ncol;
        const long long g = g00 + (long long)ly * ncol + lx;
        const uint2 lc = inside ? __ldg(loccnt + g) : make_uint2(TL_LOC_INVALID, 0u);
        own[it] = lc.x;
        up[it] = lc.y;
        s.X[i] = 0;
        const uint32_t d = tl_dir_of(dirs, it);
        s.S[i] = (uint16_t)tl_local_next(i, d);  // in-tile successor, for the walkers below
        // pit cells: basin id stashed by stash_pit_ids_kernel at the pit's own cell (requested now, used much later)
        // (asynchronous 4-byte copy straight into shared memory: no register is held across the kernel)
        if ((d == PFD_DIR_PIT || d == PFD_DIR_FPIT) && basin_out) tl_cp_async4(&s.R[i], basin_out + g);
    }
    __syncthreads();
    // entry cells with outside inflow become walkers
    if (threadIdx.x < TL_NRING) {
        const uint32_t my_w = s.ring_w[threadIdx.x];
        if (my_w != 0u) {
            const uint32_t k = atomicAdd(&s.wl_count, 1u);
            s.wl_cell[k] = (uint32_t)tl_ring_cell(threadIdx.x);
            s.wl_w[k] = my_w;
        }
    }
    __syncthreads();
    if (threadIdx.x < s.wl_count) {
        int i = (int)s.wl_cell[threadIdx.x];
        const uint32_t w = s.wl_w[threadIdx.x];
        for (int step = 0; step < TL_CELLS; ++step) {
            atomicAdd(&s.X[i], w);
            const int ni = (int)s.S[i];
            if (ni == i) break;
            i = ni;
        }
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) up[it] += s.X[(ly0 + TL_RPI * it) * TL_W + lx];
    __syncthreads();
    // terminals publish (rank at the terminal, basin id): pits by their owner, exit cells by the ring threads
    tl_cp_async_wait();
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const uint32_t d = tl_dir_of(dirs, it);
        if (d == PFD_DIR_PIT || d == PFD_DIR_FPIT) {
            const int ly = ly0 + TL_RPI * it;
            const int i = ly * TL_W + lx;
            s.X[i] = basin_out ? s.R[i] : 0u;  // own asynchronous copy, completed by tl_cp_async_wait() above
            s.R[i] = 0;
        }
    }
    if (threadIdx.x < TL_NRING && s.ring_t[threadIdx.x] != TL_NOT_EXIT) {  // exit cell: one hop above the entry cell
        const int ri = tl_ring_cell(threadIdx.x);                           // of the neighbouring tile
        s.R[ri] = s.ring_t[threadIdx.x];
        s.X[ri] = s.ring_b[threadIdx.x];
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < TL_CPT; ++it) {
        const int ly = ly0 + TL_RPI * it;
        if (r0 + ly >= nrow || c0 + lx >= ncol) continue;
        const uint32_t d = tl_dir_of(dirs, it);
        int32_t rk = -9999, ua = -9999;
        uint32_t b = 0;
        if (d != PFD_DIR_NODATA) {
            rk = -1;
            ua = 1;
            if (own[it] != TL_LOC_INVALID) {
                const uint32_t root = TP_N(own[it]);
                const uint32_t tr = s.R[root];
                if (tr != 0xFFFFFFFFu) {
                    rk = (int32_t)(tr + TP_H(own[it]));
                    b = s.X[root];
                    ua = (int32_t)up[it];
                }
            }
        }
        const long long g = g00 + (long long)ly * ncol + lx;
        if (rank_out) rank_out[g] = rk;
        if (basin_out) basin_out[g] = b;
        if (uparea_out) uparea_out[g] = ua;
        if (IDXMODE == 1) {
            // wrap-around 32-bit arithmetic gives the right low word for int32 and uint32 alike
            const uint32_t g32 = (uint32_t)g;
            const uint32_t off = (uint32_t)pfd_slot_dr((int)(d & 7u)) * (uint32_t)ncol + (uint32_t)pfd_slot_dc((int)(d & 7u));
            reinterpret_cast<uint32_t*>(idxs_out)[g] = (d < 8u) ? g32 + off : ((d == PFD_DIR_NODATA) ? 0xFFFFFFFFu : g32);
        } else if (IDXMODE == 2) {
            reinterpret_cast<long long*>(idxs_out)[g] =
                (d < 8u) ? g + pfd_slot_off((int)d, ncol) : ((d == PFD_DIR_NODATA) ? -1ll : g);
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Row-tiled multi-GPU solve: boundary tables.
// Boundary j (between rank j and j+1) has two sides of ncol entries: side 0 = last row of rank j, side 1 = first
// row of rank j+1; entry index b = (2*j + side) * ncol + col. Four uint32 tables of NB = 2*(R-1)*ncol entries,
// exchanged with ONE all-reduce(sum) (every entry is written by exactly one rank, the rest contribute zeros):
//   H1  : inflow weight arriving at the entry from the neighbouring rank (what that rank's halo slot accumulated)
//   NXT : 0 invalid | 1 path ends in a pit inside the owner's block | 2 + b' path leaves the block into entry b'
//   HOP : cell hops from the entry to that pit / to entry b'
//   BAS : basin id of that pit (NXT == 1)
// ---------------------------------------------------------------------------------------------------------
struct BoundaryTables {
    uint32_t* h1;
    uint32_t* nxt;
    uint32_t* hop;
    uint32_t* bas;
};

// halo slots of the two halo tile rows: terminals of the local reduced graph
__global__ void halo_slots_init_kernel(SlotBuf b0, uint32_t* __restrict__ term, uint32_t* __restrict__ term_h,
                                       long long ntx, long long nty) {
    const long long per_row = ntx * TL_RING;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < 2 * per_row; k += (long long)gridDim.x * blockDim.x) {
        const long long s = (k < per_row) ? k : (nty + 1) * per_row + (k - per_row);
        b0.nxt[s] = (uint32_t)s;
        b0.rh[s] = 0;
        b0.ch[s] = 0;
        b0.acc[s] = 0;
        term[s] = SLOT_INVALID;
        term_h[s] = 0;
    }
}

// side_sel 0: my top boundary (boundary rank-1), 1: my bottom boundary (boundary rank)
__global__ void boundary_fill_kernel(SlotBuf b0, SlotBuf b1, const int* __restrict__ rounds, const uint32_t* __restrict__ term,
                                     const uint32_t* __restrict__ term_h, long long nrow, long long ncol, long long ntx,
                                     long long nty, int rank, int has_top, int has_bot, BoundaryTables T) {
    const SlotBuf cur = (*rounds & 1) ? b1 : b0;
    const long long per_row = ntx * TL_RING;
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < 2 * ncol; k += (long long)gridDim.x * blockDim.x) {
        const int bottom = k >= ncol;
        const long long c = bottom ? k - ncol : k;
        if (bottom ? !has_bot : !has_top) continue;
        const long long j = bottom ? rank : rank - 1;              // boundary index
        const long long own_b = (2 * j + (bottom ? 0 : 1)) * ncol + c;   // my entry on that boundary
        const long long nb_b = (2 * j + (bottom ? 1 : 0)) * ncol + c;    // the neighbour's entry (my halo slot)
        const uint32_t hs = tl_slot_of(bottom ? nrow : -1, c, ntx);      // halo slot of the neighbour's cell
        T.h1[nb_b] = cur.acc[hs];
        const uint32_t s = tl_slot_of(bottom ? nrow - 1 : 0, c, ntx);
        uint32_t nx = 0, hop = 0, bas = 0;
        if (term[s] != SLOT_INVALID) {
            const uint32_t last = cur.nxt[s];
            const uint32_t t = term[last];
            const bool in_halo = (long long)last < per_row || (long long)last >= (nty + 1) * per_row;
            if (cur.nxt[last] == last) {
                if (in_halo) {
                    // which entry of which boundary is that halo slot? top halo row -> boundary rank-1 side 0,
                    // bottom halo row -> boundary rank side 1; column from the slot's tile column + ring position
                    const bool top_halo = (long long)last < per_row;
                    const long long within = top_halo ? last : last - (nty + 1) * per_row;
                    const long long tx = within / TL_RING;
                    const int rp = (int)(within % TL_RING);
                    // adjacent ring row of the halo tile: bottom ring row (rp in [TL_W, 2*TL_W)) for the top halo,
                    // top ring row (rp < TL_W) for the bottom halo
                    const int lx = top_halo ? rp - TL_W : rp;
                    if (lx >= 0 && lx < TL_W) {
                        const long long jj = top_halo ? rank - 1 : rank;
                        nx = (uint32_t)(2 + (2 * jj + (top_halo ? 0 : 1)) * ncol + tx * TL_W + lx);
                        hop = cur.ch[s];
                    }
                } else if (t & TERM_PIT) {
                    nx = 1;
                    hop = cur.ch[s] + term_h[last];
                    bas = (t & ~TERM_PIT) + 1u;
                }
            }
        }
        T.nxt[own_b] = nx;
        T.hop[own_b] = hop;
        T.bas[own_b] = bas;
    }
}

// all-reduced tables -> reduced-graph state of the boundary graph (NB nodes), same layout as the ring slots
__global__ void boundary_build_kernel(BoundaryTables T, long long nb, SlotBuf b0, uint32_t* __restrict__ term,
                                      uint32_t* __restrict__ term_h) {
    for (long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x; b < nb; b += (long long)gridDim.x * blockDim.x) {
        const uint32_t nx = T.nxt[b];
        b0.acc[b] = T.h1[b];
        if (nx >= 2u) {
            b0.nxt[b] = nx - 2u;
            b0.rh[b] = 1;
            b0.ch[b] = T.hop[b];
            term[b] = 0;
            term_h[b] = 0;
        } else {
            b0.nxt[b] = (uint32_t)b;
            b0.rh[b] = 0;
            b0.ch[b] = 0;
            term[b] = (nx == 1u) ? (TERM_PIT | (T.bas[b] - 1u)) : SLOT_INVALID;
            term_h[b] = (nx == 1u) ? T.hop[b] : 0;
        }
    }
}

// boundary solution -> my halo slots become pit-like terminals carrying (rank, basin) of the neighbour's entry;
// my own boundary entries receive the total remote inflow X on top of their local weight
__global__ void boundary_writeback_kernel(const int32_t* __restrict__ brank, const uint32_t* __restrict__ bbasin,
                                          const uint32_t* __restrict__ bx, long long nrow, long long ncol, long long ntx,
                                          int rank, int has_top, int has_bot, uint32_t* __restrict__ w,
                                          uint32_t* __restrict__ term, uint32_t* __restrict__ term_h) {
    for (long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x; k < 2 * ncol; k += (long long)gridDim.x * blockDim.x) {
        const int bottom = k >= ncol;
        const long long c = bottom ? k - ncol : k;
        if (bottom ? !has_bot : !has_top) continue;
        const long long j = bottom ? rank : rank - 1;
        const long long own_b = (2 * j + (bottom ? 0 : 1)) * ncol + c;
        const long long nb_b = (2 * j + (bottom ? 1 : 0)) * ncol + c;
        const uint32_t hs = tl_slot_of(bottom ? nrow : -1, c, ntx);
        const int32_t rk = brank[nb_b];
        term[hs] = (rk >= 0) ? (TERM_PIT | (bbasin[nb_b] - 1u)) : SLOT_INVALID;
        term_h[hs] = (rk >= 0) ? (uint32_t)rk : 0u;
        const uint32_t s = tl_slot_of(bottom ? nrow - 1 : 0, c, ntx);
        if (bx[own_b]) atomicAdd(w + s, bx[own_b]);  // a 1-row block has the same slot on both of its boundaries
    }
}
