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

// Thread layout of the tile kernels: 1024 threads, thread t owns the FOUR horizontally adjacent cells 4t .. 4t+3
// (row t >> 4, columns 4 * (t & 15) ..), so that its own cells are one 32-bit word of direction bytes and 128-bit
// vectors in global memory.
#define TQ_ROW ((int)(threadIdx.x >> 4))
#define TQ_COL0 ((int)((threadIdx.x & 15u) << 2))
#define TQ_I0 ((int)(threadIdx.x << 2))
// Shared-memory position of local cell i = 4t + j: plane j, entry t. A warp touching cell j of its 32 quads hits 32
// consecutive words (no bank conflict), so do the typical gathers (a neighbour of every such cell), and the thread's
// own cells sit at compile-time offsets (j * 4 KiB) from one per-thread base address.
#define TPHYS(i) ((((i) & 3) << 10) | ((i) >> 2))

// Shared state of one tile in phase A. P packs (shared-memory ADDRESS of the P entry of the next cell : 18 bits |
// hops to it : 13 bits); A is the accumulate buffer, laid out right behind P so that &A[x] == &P[x] + 16 KiB.
struct TileShared {
    uint32_t P[TL_CELLS];
    uint32_t A[TL_CELLS];
};
#define TPK_MASK 0x3FFFFu
#define TPK_SHIFT 18
#define TL_A_OFF (TL_CELLS * 4)
// per-cell record handed from phase A to phase C: TPHYS position of the local terminal | hops << 12
#define TP_N(p) ((p) & 0xFFFu)
#define TP_H(p) ((p) >> 12)
#define TL_LOC_INVALID 0xFFFFFFFFu

// shared-memory accesses by 32-bit shared-window address: one LDS / STS / RED each, no address arithmetic left to
// the compiler inside the solve loop (the issue-bound part of phase A)
__device__ __forceinline__ uint32_t tl_smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int OFF>
__device__ __forceinline__ uint32_t tl_lds(uint32_t a) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1+%2];" : "=r"(v) : "r"(a), "n"(OFF) : "memory");
    return v;
}
template <int OFF>
__device__ __forceinline__ void tl_sts(uint32_t a, uint32_t v) {
    asm volatile("st.shared.u32 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(OFF) : "memory");
}
template <int OFF>
__device__ __forceinline__ void tl_red_add(uint32_t a, uint32_t v) {
    asm volatile("red.shared.add.u32 [%0+%2], %1;" ::"r"(a), "r"(v), "n"(OFF) : "memory");
}

__device__ __forceinline__ uint32_t tl_dir_of(uint32_t dirw, int j) { return (dirw >> (8 * j)) & 0xFFu; }

// local index of the next cell inside the tile, or i itself when the link leaves the tile / the cell is a pit
__device__ __forceinline__ int tl_local_next(int i, uint32_t d) {
    if (d >= 8u) return i;
    const int y = (i >> 6) + pfd_slot_dr((int)d), x = (i & (TL_W - 1)) + pfd_slot_dc((int)d);
    return (y >= 0 && y < TL_H && x >= 0 && x < TL_W) ? y * TL_W + x : i;
}

// Local solve of phase A. baseP = shared address of P[0]. On return own[j] (mirrored in P) = packed (address of the
// local terminal's P entry, hop distance) of every owned cell and A = subtree sum of the unit weights inside the tile.
// Round k (Jacobi): every cell whose 2^k-th ancestor exists (hops == 2^k) snapshots its A and its ancestor's P,
// then -- after a barrier -- adds the snapshot to that ancestor (A_{k+1}[anc] += A_k[d]) and jumps
// (next <- next[next], hops += hops[next]). A cell stays active only while its hop count is an exact power of two,
// i.e. its chain is not exhausted; the loop ends when no cell of the tile is active (<= 12 rounds without loops).
// The body is branch-free: inactive cells repeat harmless loads, only the RED / STS are predicated.
__device__ __forceinline__ void tl_local_solve(uint32_t baseP, uint32_t dirw, uint32_t* own) {
    const int i0 = TQ_I0;
    const uint32_t ownP = baseP + (threadIdx.x << 2);  // plane 0 entry of the thread; plane j at + j * 4096
    bool act[4];
#define TL_FOR4(BODY) { { constexpr int j = 0; BODY } { constexpr int j = 1; BODY } { constexpr int j = 2; BODY } { constexpr int j = 3; BODY } }
    TL_FOR4({
        const uint32_t d = tl_dir_of(dirw, j);
        const int ni = tl_local_next(i0 + j, d);
        act[j] = ni != i0 + j;
        own[j] = (baseP + ((uint32_t)TPHYS(ni) << 2)) | (act[j] ? (1u << TPK_SHIFT) : 0u);
        tl_sts<j * 4096>(ownP, own[j]);
        tl_sts<TL_A_OFF + j * 4096>(ownP, (d != PFD_DIR_NODATA) ? 1u : 0u);
    })
    __syncthreads();
    for (int k = 0; k < TL_MAXROUNDS; ++k) {
        const uint32_t two_k = 1u << k;
        uint32_t pn[4], a[4];
        TL_FOR4({
            pn[j] = tl_lds<0>(own[j] & TPK_MASK);
            a[j] = tl_lds<TL_A_OFF + j * 4096>(ownP);
        })
        __syncthreads();  // every snapshot is taken before any update
        bool any = false;
        TL_FOR4({
            if (act[j]) tl_red_add<TL_A_OFF>(own[j] & TPK_MASK, a[j]);
            const uint32_t hp = pn[j] >> TPK_SHIFT;
            const uint32_t nw = (pn[j] & TPK_MASK) | ((hp + two_k) << TPK_SHIFT);
            if (act[j]) {
                own[j] = nw;
                tl_sts<j * 4096>(ownP, nw);
            }
            act[j] = act[j] && (hp == two_k);
            any = any || act[j];
        })
        if (!__syncthreads_or((int)any)) break;
    }
}

// the direction bytes of the thread's four cells as one word; cells outside the raster read as nodata.
// al4: ncol % 4 == 0 and the base pointer is 4-byte aligned (a quad never straddles the row end)
__device__ __forceinline__ uint32_t tl_load_dirs(const uint8_t* __restrict__ dir, long long nrow, long long ncol,
                                                 long long r0, long long c0, bool al4) {
    const long long r = r0 + TQ_ROW, c = c0 + TQ_COL0;
    uint32_t w = 0xFFFFFFFFu;
    if (r < nrow && c < ncol) {
        const uint8_t* p = dir + r * ncol + c;
        if (al4) {
            w = __ldg(reinterpret_cast<const uint32_t*>(p));
        } else {
            w = 0;
#pragma unroll
            for (int b = 0; b < 4; ++b) w |= ((c + b < ncol) ? (uint32_t)__ldg(p + b) : (uint32_t)PFD_DIR_NODATA) << (8 * b);
        }
    }
    return w;
}

// ---------------------------------------------------------------------------------------------------------
// Fused parse (single-GPU headline path): the tile's raw D8 codes plus a one-cell halo are staged in shared
// memory (32-bit loads for the interior, out-of-raster cells read as 247), every thread derives the `dir` bytes
// of its four cells with the same byte-SIMD routine as parse_kernel (pfd_parse_word; core_d8.py:42-67: pit /
// forced pit when the downstream cell is nodata or off the raster / nodata; illegal codes raise the invalid flag,
// core_d8.py:115-122) and writes them to global memory, so that the separate parse pass over the raster
// disappears from pfd_d8_flow_all.
// ---------------------------------------------------------------------------------------------------------
#ifdef TL_BULK          // the 64 interior rows of an interior tile arrive by bulk asynchronous copies (TMA unit, UBLKCP): 16-byte aligned rows
#define TLF_STRIDE 80   // bytes per staged row: 15 pad | left halo | 64 cells | right halo | 15 pad
#define TLF_X0 16
#else
#define TLF_STRIDE 72   // bytes per staged row: 3 pad | left halo | 64 cells | right halo | 3 pad
#define TLF_X0 4        // byte offset of the tile's first column inside a staged row (word aligned)
#endif
struct TileCodes {
    __align__(16) uint8_t c[(TL_H + 2) * TLF_STRIDE];
#ifdef TL_BULK
    __align__(8) unsigned long long mbar;
#endif
};

// row_lo / row_hi: rows that may be READ: [0, nrow) for a whole raster; a row block of a larger raster also has the
// neighbour's edge row at -1 / nrow (they only feed the forced-pit test of the block's edge rows, core_d8.py:58-61)
template <int THREADS>
__device__ __forceinline__ void tl_stage_codes(TileCodes& sc, const uint8_t* __restrict__ d8, long long nrow, long long ncol,
                                               long long r0, long long c0, bool al4, long long row_lo, long long row_hi) {
#ifdef TL_BULK
    // interior tile of an aligned raster: one thread arms an mbarrier with the 4 KiB it expects and issues one 64-byte bulk copy
    // per row; everybody else goes on to the halo and meets the data at the mbarrier
    const bool bulk = (ncol % 16 == 0) && ((reinterpret_cast<uintptr_t>(d8) & 15) == 0) && r0 + TL_H <= nrow && c0 + TL_W <= ncol;
    const uint32_t mb = tl_smem_addr(&sc.mbar);
    if (bulk) {
        if (threadIdx.x == 0) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mb) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mb), "r"(TL_H * TL_W) : "memory");
            const uint8_t* src = d8 + r0 * ncol + c0;
            uint32_t dst = tl_smem_addr(&sc.c[TLF_STRIDE + TLF_X0]);
#pragma unroll 8
            for (int row = 0; row < TL_H; ++row, src += ncol, dst += TLF_STRIDE)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                             "r"(TL_W), "r"(mb)
                             : "memory");
        }
    } else
#endif
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
        if (r >= row_lo && r < row_hi && c >= 0 && c < ncol) v = __ldg(d8 + r * ncol + c);
        sc.c[sy * TLF_STRIDE + TLF_X0 + sx] = v;
    }
#ifdef TL_BULK
    if (bulk) {
        __syncthreads();  // the mbarrier is initialised before anybody polls it
        asm volatile(
            "{\n\t.reg .pred p;\n\tTL_BULK_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], 0;\n\t@!p bra TL_BULK_WAIT;\n\t}" ::"r"(mb)
            : "memory");
    }
#endif
}

// dir bytes of the thread's four cells from the staged codes; *legal = false when one of them is not a D8 code
__device__ __forceinline__ uint32_t tl_parse_dirs(const TileCodes& sc, bool* legal) {
    const uint32_t* row = reinterpret_cast<const uint32_t*>(&sc.c[TQ_ROW * TLF_STRIDE + TLF_X0 + TQ_COL0 - 4]);
    constexpr int RW = TLF_STRIDE / 4;
    uint32_t dirw, upw;
    const uint32_t ok = pfd_parse_word<0>(row[0], row[1], row[2], row[RW], row[RW + 1], row[RW + 2], row[2 * RW],
                                          row[2 * RW + 1], row[2 * RW + 2], dirw, upw);
    *legal = ok == 0xFFFFFFFFu;
    return dirw;
}

// ---------------------------------------------------------------------------------------------------------
// Phase A: local solve; per cell (local terminal, hops) -> loc[], in-tile subtree size -> cnt[]; ring nodes; W
// ---------------------------------------------------------------------------------------------------------
// FUSED = false: `dir` is the parsed direction raster, pit terminals carry the pit ordinal found in `pit_ids`.
// FUSED = true : the tile's raw D8 codes (+ halo) are already staged in `sc`; the directions are derived here and
//                written to `dir_out` (illegal codes raise the flag); pit terminals carry TERM_PIT | local cell index,
//                resolved to the pit ordinal by slots_finalize_kernel once the pits have been numbered.
// al4: ncol % 4 == 0 and every raster-sized buffer is 16-byte aligned -> 32 / 128-bit accesses for the quad.
struct PhaseAArgs {
    const uint8_t* dir;
    long long nrow, ncol, ntx;
    const uint32_t* pit_ids;
    uint2* loccnt;
    uint32_t *W, *s_nxt, *s_rh, *s_ch, *s_term, *s_term_h;
    const uint8_t* d8;
    uint8_t* dir_out;
    unsigned int* invalid_flag;
    int al4;
    int halo_top, halo_bot;  // FUSED, row block of a larger raster: d8 has the neighbour's edge row above / below
};

template <bool FUSED>
__device__ __forceinline__ void tl_phase_a_tile(TileShared& s, TileCodes& sc, const PhaseAArgs& A, long long ty, long long tx) {
    const uint8_t* __restrict__ dir = A.dir;
    const long long nrow = A.nrow, ncol = A.ncol, ntx = A.ntx;
    const uint32_t* pit_ids = A.pit_ids;
    uint2* __restrict__ loccnt = A.loccnt;
    uint32_t* __restrict__ W = A.W;
    const int al4 = A.al4;
    const uint32_t tile = (uint32_t)((ty + 1) * ntx + tx);  // +1: halo tile row
    const long long r0 = ty * TL_H, c0 = tx * TL_W;
    const int ly = TQ_ROW, lx0 = TQ_COL0, i0 = TQ_I0;
    const long long g00 = r0 * ncol + c0;  // global index of the tile's first cell
    const long long g0 = g00 + (long long)ly * ncol + lx0;  // global index of the thread's first cell
    const bool row_in = r0 + ly < nrow;
    const bool quad_in = row_in && c0 + lx0 + 3 < ncol;  // all four cells inside the raster

    uint32_t dirw, own[4];
    if (FUSED) {
        bool legal;
        dirw = tl_parse_dirs(sc, &legal);
        if (!legal) atomicOr(A.invalid_flag, 1u);
        __syncthreads();  // every neighbour code has been read: the staged codes may now be replaced by the dirs
        *reinterpret_cast<uint32_t*>(&sc.c[(ly + 1) * TLF_STRIDE + TLF_X0 + lx0]) = dirw;  // read back by the ring threads
        if (al4 && quad_in) {
            *reinterpret_cast<uint32_t*>(A.dir_out + g0) = dirw;
        } else if (row_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + lx0 + j < ncol) A.dir_out[g0 + j] = (uint8_t)tl_dir_of(dirw, j);
        }
    } else {
        dirw = tl_load_dirs(dir, nrow, ncol, r0, c0, al4 != 0);
    }
    const uint32_t baseP = tl_smem_addr(&s.P[0]);
    const uint32_t ownP = baseP + (threadIdx.x << 2);
    tl_local_solve(baseP, dirw, own);

    // (1) per-cell results for phase C
    {
        uint32_t loc[4], c4[4];
        TL_FOR4({
            c4[j] = tl_lds<TL_A_OFF + j * 4096>(ownP);
            const uint32_t root = own[j] & TPK_MASK;
            // a terminal is (next = itself, hops = 0)
            loc[j] = (tl_lds<0>(root) != root) ? TL_LOC_INVALID : (((root - baseP) >> 2) | ((own[j] >> TPK_SHIFT) << 12));
        })
        if (al4 && quad_in) {
            uint4* o = reinterpret_cast<uint4*>(loccnt + g0);  // (local terminal | hops, in-tile count) x 4
            o[0] = make_uint4(loc[0], c4[0], loc[1], c4[1]);
            o[1] = make_uint4(loc[2], c4[2], loc[3], c4[3]);
        } else if (row_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + lx0 + j < ncol) loccnt[g0 + j] = make_uint2(loc[j], c4[j]);
        }
    }
    __syncthreads();
    // (2) terminals publish their descriptor through their A slot: pits by their owner thread, exit cells (always
    //     ring cells) by one thread per ring position, which also hands the exit cell's in-tile subtree size to the
    //     entry cell it drains into
    TL_FOR4({
        const uint32_t d = tl_dir_of(dirw, j);
        if (d == PFD_DIR_PIT || d == PFD_DIR_FPIT) {
            if (FUSED) tl_sts<TL_A_OFF + j * 4096>(ownP, TERM_PIT | (uint32_t)(i0 + j));
            else tl_sts<TL_A_OFF + j * 4096>(ownP, TERM_PIT | (pit_ids ? (pit_ids[g0 + j] - 1u) : 0u));
        }
    })
    int ri = -1;
    uint32_t rd = PFD_DIR_NODATA, rP = 0;
    if (threadIdx.x < TL_NRING) {
        ri = tl_ring_cell(threadIdx.x);
        rP = baseP + ((uint32_t)TPHYS(ri) << 2);
        const int rly = ri >> 6, rlx = ri & (TL_W - 1);
        if (FUSED) rd = sc.c[(rly + 1) * TLF_STRIDE + TLF_X0 + rlx];  // out-of-raster cells were staged as nodata
        else if (r0 + rly < nrow && c0 + rlx < ncol) rd = __ldg(dir + g00 + (long long)rly * ncol + rlx);  // L1/L2 hit
        if (rd < 8u && tl_lds<0>(rP) == rP) {  // exit cell
            const uint32_t ti = tl_exit_slot(tile, (uint32_t)ntx, rly, rlx, rd);
            atomicAdd(W + ti, tl_lds<TL_A_OFF>(rP));
            tl_sts<TL_A_OFF>(rP, ti);
        }
    }
    __syncthreads();
    // (3) ring cells publish their reduced-graph node
    if (ri >= 0) {
        const uint32_t slot = tile * TL_RING + threadIdx.x;
        uint32_t nx = slot, rh = 0, ch = 0, term = SLOT_INVALID, th = 0;
        const uint32_t p = tl_lds<0>(rP);
        const uint32_t root = p & TPK_MASK;
        if (rd != PFD_DIR_NODATA && tl_lds<0>(root) == root) {
            const uint32_t ti = tl_lds<TL_A_OFF>(root);
            const uint32_t dist = p >> TPK_SHIFT;
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
        A.s_nxt[slot] = nx;
        A.s_rh[slot] = rh;
        A.s_ch[slot] = ch;
        A.s_term[slot] = term;
        A.s_term_h[slot] = th;
    } else if (threadIdx.x < TL_RING) {  // the 4 padding slots of the tile: inert nodes (the solve streams over every slot)
        const uint32_t slot = tile * TL_RING + threadIdx.x;
        A.s_nxt[slot] = slot;
        A.s_rh[slot] = 0u;
        A.s_ch[slot] = 0u;
        A.s_term[slot] = SLOT_INVALID;
        A.s_term_h[slot] = 0u;
    }
}

// one CTA per tile (blockIdx = tile column, tile row): any raster shape / alignment
template <int THREADS, int MINBLOCKS, bool FUSED>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) tile_phase_a_kernel(PhaseAArgs A) {
    static_assert(THREADS == TL_CELLS / 4, "one quad of cells per thread");
    __shared__ __align__(16) TileShared s;
    __shared__ TileCodes sc;  // FUSED only (the compiler drops it otherwise)
    if (FUSED) {
        tl_stage_codes<THREADS>(sc, A.d8, A.nrow, A.ncol, (long long)blockIdx.y * TL_H, (long long)blockIdx.x * TL_W, A.al4 != 0,
                                -(long long)A.halo_top, A.nrow + A.halo_bot);
        __syncthreads();
    }
    tl_phase_a_tile<FUSED>(s, sc, A, blockIdx.y, blockIdx.x);
}

template <bool FUSED>
static void tl_launch_phase_a(dim3 grid, cudaStream_t stream, const PhaseAArgs& A) {
    tile_phase_a_kernel<TLA_THREADS, TLA_MINBLOCKS, FUSED><<<grid, TLA_THREADS, 0, stream>>>(A);
}

// asynchronous global -> shared copies (LDGSTS): no register is tied up while the data is in flight
__device__ __forceinline__ void tl_cp_async4(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void tl_cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void tl_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tl_cp_async_wait_group() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tl_cp_async_wait() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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
                                                             const uint32_t* __restrict__ pit_stash, long long ncol, long long ntx,
                                                             long long nty = (1ll << 40)) {
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
                // halo slots of a row block (tile rows -1 / nty) are pit-like terminals that carry the basin id itself
                if (ty < 0 || ty >= nty) b = (t & ~TERM_PIT) + 1u;
                else b = __ldg(pit_stash + (ty * TL_H + (li >> 6)) * ncol + tx * TL_W + (li & (TL_W - 1)));
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
// The walkers follow 4-step successor records (S12 / S34: 1st..4th in-tile successor of every cell, built with two
// pointer-jumping steps), so their serial chain is one shared-memory round trip per FOUR cells.
// ---------------------------------------------------------------------------------------------------------
struct TileSharedC {
    uint32_t X[TL_CELLS];   // extra inflow per cell, later basin id per terminal
    uint32_t R[TL_CELLS];   // pit cells: stashed basin id (fetched with the first loads); later rank at the terminal
    uint32_t S12[TL_CELLS];  // 1st | 2nd << 16 in-tile successor of the cell (a terminal repeats itself)
    uint32_t S34[TL_CELLS];  // 3rd | 4th << 16
    uint32_t wl_cell[TL_RING];  // walker list
    uint32_t wl_w[TL_RING];
    int32_t ring_t[TL_RING];    // per ring position: reduced-graph rank of the entry cell an exit cell drains into, else TL_NOT_EXIT
    uint32_t ring_b[TL_RING];   //                    basin id behind that exit
    uint32_t ring_w[TL_RING];   //                    outside inflow of the (entry) cell
    uint32_t wl_count;
};
#define TL_NOT_EXIT ((int32_t)0x80000000)

struct PhaseCArgs {
    const uint8_t* dir;
    long long nrow, ncol, ntx, nty;
    const uint2* loccnt;
    const uint32_t* inflow;
    const int32_t* s_rank;
    const uint32_t* s_basin;
    int32_t* rank_out;
    uint32_t* basin_out;
    int32_t* uparea_out;
    void* idxs_out;
    long long idx_base;  // global linear index of the first owned cell (row block of a larger raster), else 0
    int al4;
};

// IDXMODE: 0 = no idxs_ds output, 1 = 32-bit (int32 / uint32 share the bit pattern), 2 = int64 -- the fused-parse path
// writes idxs_ds (core_d8.from_array, core_d8.py:42-67) from here, next to the other per-cell outputs.
// One CTA per tile (blockIdx = tile column, tile row); dynamic shared memory sizeof(TileSharedC).
// (Variants that process several tiles per CTA and stream the next tile in with cp.async behind the current one --
// persistent with a dynamic tile counter, or 4 adjacent tiles per CTA -- were measured 7-15 % SLOWER on B200 for both
// tile phases: DESIGN.md section 4.3.)
template <int THREADS, int MINBLOCKS, int IDXMODE>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) tile_phase_c_kernel(PhaseCArgs A) {
    static_assert(THREADS == TL_CELLS / 4, "one quad of cells per thread");
    extern __shared__ __align__(16) unsigned char tl_smem_raw[];
    TileSharedC& s = *reinterpret_cast<TileSharedC*>(tl_smem_raw);
    const uint8_t* __restrict__ dir = A.dir;
    const long long nrow = A.nrow, ncol = A.ncol, ntx = A.ntx;
    const uint2* __restrict__ loccnt = A.loccnt;
    int32_t* __restrict__ rank_out = A.rank_out;
    uint32_t* basin_out = A.basin_out;
    int32_t* __restrict__ uparea_out = A.uparea_out;
    const int al4 = A.al4;
    const int ly = TQ_ROW, lx0 = TQ_COL0, i0 = TQ_I0;
    {
        const long long ty = blockIdx.y, tx = blockIdx.x;
        const uint32_t tile = (uint32_t)((ty + 1) * ntx + tx);  // +1: halo tile row
        const long long r0 = ty * TL_H, c0 = tx * TL_W;
        const long long g00 = r0 * ncol + c0;
        const long long g0 = g00 + (long long)ly * ncol + lx0;
        const bool row_in = r0 + ly < nrow;
        const bool quad_in = row_in && c0 + lx0 + 3 < ncol;
        const bool vec = al4 && quad_in;

        uint32_t dirw, own[4], up[4], n1[4];
        {  // per-cell loads first: they are in flight while the ring threads chase their dependent loads
            dirw = tl_load_dirs(dir, nrow, ncol, r0, c0, al4 != 0);
            if (vec) {
                const uint4* lc = reinterpret_cast<const uint4*>(loccnt + g0);
                const uint4 v0 = __ldg(lc), v1 = __ldg(lc + 1);
                own[0] = v0.x, up[0] = v0.y, own[1] = v0.z, up[1] = v0.w;
                own[2] = v1.x, up[2] = v1.y, own[3] = v1.z, up[3] = v1.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const uint2 lc = (row_in && c0 + lx0 + j < ncol) ? __ldg(loccnt + g0 + j) : make_uint2(TL_LOC_INVALID, 0u);
                    own[j] = lc.x;
                    up[j] = lc.y;
                }
            }
        }
        // one thread per ring position: entry inflow; for exit cells the solution of the entry cell they drain into
        // (parked in shared memory so that nothing occupies registers across the kernel)
        if (threadIdx.x < TL_NRING) {
            const int ri = tl_ring_cell(threadIdx.x);
            const int rly = ri >> 6, lxr = ri & (TL_W - 1);
            uint32_t rw = 0;
            bool is_exit = false;
            if (r0 + rly < nrow && c0 + lxr < ncol) {
                const long long g = g00 + (long long)rly * ncol + lxr;
                const uint32_t rd = __ldg(dir + g);
                const uint32_t rloc = __ldg(&loccnt[g].x);
                if (uparea_out && rloc != TL_LOC_INVALID) rw = __ldg(A.inflow + tile * TL_RING + threadIdx.x);
                if (rd < 8u && rloc == (uint32_t)TPHYS(ri)) {  // exit cell (terminal = itself, 0 hops)
                    const uint32_t slot = tl_exit_slot(tile, (uint32_t)ntx, rly, lxr, rd);
                    is_exit = true;
                    // asynchronous copies: consumed after the walkers, nothing waits for them before
                    tl_cp_async4(&s.ring_t[threadIdx.x], A.s_rank + slot);
                    tl_cp_async4(&s.ring_b[threadIdx.x], A.s_basin + slot);
                }
            }
            s.ring_w[threadIdx.x] = rw;
            if (!is_exit) s.ring_t[threadIdx.x] = TL_NOT_EXIT;
        }
        if (threadIdx.x == 0) s.wl_count = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t d = tl_dir_of(dirw, j);
            n1[j] = (uint32_t)TPHYS(tl_local_next(i0 + j, d));  // positions, not cell numbers, from here on
            // pit cells: basin id stashed by stash_pit_ids_kernel at the pit's own cell (requested now, used much
            // later; asynchronous 4-byte copy straight into shared memory)
            if ((d == PFD_DIR_PIT || d == PFD_DIR_FPIT) && basin_out) tl_cp_async4(&s.R[TPHYS(i0 + j)], basin_out + g0 + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s.X[TPHYS(i0 + j)] = 0u;
            s.S12[TPHYS(i0 + j)] = n1[j];
        }
        __syncthreads();
        // entry cells with outside inflow become walkers
        if (threadIdx.x < TL_NRING) {
            const uint32_t my_w = s.ring_w[threadIdx.x];
            if (my_w != 0u) {
                const uint32_t k = atomicAdd(&s.wl_count, 1u);
                s.wl_cell[k] = (uint32_t)TPHYS(tl_ring_cell(threadIdx.x));
                s.wl_w[k] = my_w;
            }
        }
        // successor records, jump 1: 2nd successor. Concurrent readers only use the low half of S12, which does not change.
        uint32_t x12[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x12[j] = n1[j] | ((s.S12[n1[j]] & 0xFFFFu) << 16);
#pragma unroll
        for (int j = 0; j < 4; ++j) s.S12[TPHYS(i0 + j)] = x12[j];
        __syncthreads();
        // jump 2: S12 of the 2nd successor holds its 1st | 2nd successor = my 3rd | 4th
#pragma unroll
        for (int j = 0; j < 4; ++j) s.S34[TPHYS(i0 + j)] = s.S12[x12[j] >> 16];
        __syncthreads();
        if (threadIdx.x < s.wl_count) {
            uint32_t i = s.wl_cell[threadIdx.x];
            const uint32_t w = s.wl_w[threadIdx.x];
            for (int step = 0; step <= TL_CELLS / 4; ++step) {
                const uint32_t s12 = s.S12[i], s34 = s.S34[i];
                const uint32_t m1 = s12 & 0xFFFFu, m2 = s12 >> 16, m3 = s34 & 0xFFFFu, m4 = s34 >> 16;
                atomicAdd(&s.X[i], w);
                if (m1 == i) break;
                atomicAdd(&s.X[m1], w);
                if (m2 == m1) break;
                atomicAdd(&s.X[m2], w);
                if (m3 == m2) break;
                atomicAdd(&s.X[m3], w);
                if (m4 == m3) break;
                i = m4;
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) up[j] += s.X[TPHYS(i0 + j)];
        tl_cp_async_wait();  // my asynchronous copies (pit basin ids, exit solutions) have landed
        __syncthreads();
        // terminals publish (rank at the terminal, basin id): pits by their owner, exit cells by the ring threads
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t d = tl_dir_of(dirw, j);
            if (d == PFD_DIR_PIT || d == PFD_DIR_FPIT) {
                s.X[TPHYS(i0 + j)] = basin_out ? s.R[TPHYS(i0 + j)] : 0u;  // own asynchronous copy
                s.R[TPHYS(i0 + j)] = 0;
            }
        }
        if (threadIdx.x < TL_NRING) {
            const int32_t rrank = s.ring_t[threadIdx.x];  // own (asynchronous) copy
            if (rrank != TL_NOT_EXIT) {                   // exit cell: one hop above the entry cell of the neighbouring tile
                const int ri = tl_ring_cell(threadIdx.x);
                s.R[TPHYS(ri)] = (rrank < 0) ? 0xFFFFFFFFu : (uint32_t)(rrank + 1);
                s.X[TPHYS(ri)] = (rrank < 0) ? 0u : s.ring_b[threadIdx.x];
            }
        }
        __syncthreads();
        int32_t rk[4], ua[4];
        uint32_t bs[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t d = tl_dir_of(dirw, j);
            rk[j] = -9999, ua[j] = -9999, bs[j] = 0;
            if (d != PFD_DIR_NODATA) {
                rk[j] = -1;
                ua[j] = 1;
                if (own[j] != TL_LOC_INVALID) {
                    const uint32_t root = TP_N(own[j]);
                    const uint32_t tr = s.R[root];
                    if (tr != 0xFFFFFFFFu) {
                        rk[j] = (int32_t)(tr + TP_H(own[j]));
                        bs[j] = s.X[root];
                        ua[j] = (int32_t)up[j];
                    }
                }
            }
        }
        if (vec) {
            if (rank_out) *reinterpret_cast<int4*>(rank_out + g0) = make_int4(rk[0], rk[1], rk[2], rk[3]);
            if (basin_out) *reinterpret_cast<uint4*>(basin_out + g0) = make_uint4(bs[0], bs[1], bs[2], bs[3]);
            if (uparea_out) *reinterpret_cast<int4*>(uparea_out + g0) = make_int4(ua[0], ua[1], ua[2], ua[3]);
        } else if (row_in) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                if (c0 + lx0 + j >= ncol) continue;
                if (rank_out) rank_out[g0 + j] = rk[j];
                if (basin_out) basin_out[g0 + j] = bs[j];
                if (uparea_out) uparea_out[g0 + j] = ua[j];
            }
        }
        if (IDXMODE != 0) {
            // idxs_ds: wrap-around 32-bit arithmetic gives the right low word for int32 and uint32 alike
            long long ds[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const uint32_t d = tl_dir_of(dirw, j);
                const long long g = A.idx_base + g0 + j;
                ds[j] = (d < 8u) ? g + pfd_slot_off((int)d, ncol) : ((d == PFD_DIR_NODATA) ? -1ll : g);
            }
            if (IDXMODE == 1) {
                uint32_t* o = reinterpret_cast<uint32_t*>(A.idxs_out) + g0;
                if (vec) {
                    *reinterpret_cast<uint4*>(o) = make_uint4((uint32_t)ds[0], (uint32_t)ds[1], (uint32_t)ds[2], (uint32_t)ds[3]);
                } else if (row_in) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + lx0 + j < ncol) o[j] = (uint32_t)ds[j];
                }
            } else {
                long long* o = reinterpret_cast<long long*>(A.idxs_out) + g0;
                if (vec) {
                    *reinterpret_cast<longlong2*>(o) = make_longlong2(ds[0], ds[1]);
                    *reinterpret_cast<longlong2*>(o + 2) = make_longlong2(ds[2], ds[3]);
                } else if (row_in) {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (c0 + lx0 + j < ncol) o[j] = ds[j];
                }
            }
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
                                     long long nty, int rank, int has_top, int has_bot, BoundaryTables T,
                                     const uint32_t* __restrict__ pit_stash) {
    // pit_stash != null (fused-parse path): pit terminals hold TERM_PIT | local cell index inside the tile of slot `last`;
    // the basin id was stashed at the pit's own cell of the basin buffer (see slots_finalize_kernel)
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
                    if (pit_stash) {
                        const long long tl = (long long)(last / TL_RING);
                        const long long ty = tl / ntx - 1, tx = tl % ntx;
                        const uint32_t li = t & (uint32_t)(TL_CELLS - 1);
                        bas = __ldg(pit_stash + (ty * TL_H + (li >> 6)) * ncol + tx * TL_W + (li & (TL_W - 1)));
                    } else {
                        bas = (t & ~TERM_PIT) + 1u;
                    }
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
// my own boundary entries receive the total remote inflow X on top of the local inflow of the first solve
__global__ void boundary_writeback_kernel(const int32_t* __restrict__ brank, const uint32_t* __restrict__ bbasin,
                                          const uint32_t* __restrict__ bx, long long nrow, long long ncol, long long ntx,
                                          int rank, int has_top, int has_bot, uint32_t* __restrict__ w,
                                          uint32_t* __restrict__ term, uint32_t* __restrict__ term_h,
                                          const uint32_t* __restrict__ nxt0, SlotBuf b0, SlotBuf b1,
                                          const int* __restrict__ rounds, long long nslots) {
    const SlotBuf cur = (*rounds & 1) ? b1 : b0;
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
        // The remote inflow X of my own boundary entry travels down the entry's local chain: every ring node on it gets
        // + X (what a second solve of the local reduced graph would deliver). nxt0 = the one-hop links as phase A left
        // them; chains end in a terminal (pit node or halo slot: next == itself).
        const uint32_t x = bx[own_b];
        if (x) {
            uint32_t s = tl_slot_of(bottom ? nrow - 1 : 0, c, ntx);
            // only chains that END (the first solve left their terminal in cur.nxt): a chain that runs into a loop across
            // tiles has no end, and its cells drain to no pit anyway (upstream area 1, like every cell outside `seq`)
            const uint32_t last = cur.nxt[s];
            if (cur.nxt[last] == last) {
                for (long long guard = 0; guard < nslots; ++guard) {
                    atomicAdd(w + s, x);
                    const uint32_t n = nxt0[s];
                    if (n == s) break;
                    s = n;
                }
            }
        }
    }
}
