// pfd_tilesweep.cuh -- tile-dataflow sweeps: the ORDER-SENSITIVE outputs of the hot path without a cell ordering.
//   up-sweeps   streams.accuflux, any dtype (pyflwdir/streams.py:15-41), streams.strahler_order (streams.py:228-269)
//   down-sweeps dem.height_above_nearest_drain (pyflwdir/dem.py:299-330)
//
// The reference walks `seq` (BFS order from the pits) serially. What its loops compute is, per cell, a pure function
// of the FINAL values of the cell's graph neighbours:
//   up-sweep   value(c) = fold of value(u) over the upstream neighbours u of c in DESCENDING linear index (that is
//              the order in which seq[::-1] delivers them: core.py:78-83,111-115), started from the cell's own datum;
//   down-sweep value(c) = g(value(ds(c)), data(c), data(ds(c))), hop by hop (no re-association: HAND sums float64).
// So any schedule that respects the dependencies yields the reference's bits. The level-synchronous replay of `seq`
// (pfd_sweeps.cuh) respects them but touches every 32-byte sector of the value array in ~8 different levels (82 B/cell
// of DRAM traffic for 8 algorithmic, profiles/r01a_summary.md) and needs the BFS ordering first (40 ms at 32768^2).
//
// Here the raster is cut into 64x64 tiles and the dependency chains are followed INSIDE shared memory:
//   * a tile visit stages the tile's 1-byte directions (+ one-cell halo), the values and done-flags of the halo
//     cells, and the cell data (128-bit loads); the upstream / pending masks of all cells are derived with byte-SIMD
//     compares, four cells per 32-bit word; then
//       up:   every thread starts at its ready cells (no pending upstream neighbour) and walks downstream for as long
//             as it is the LAST ARRIVER at the next cell (one shared-memory atomic clears its bit in the cell's
//             pending mask) -- barrier-free dataflow, float sums bit-exact without float atomics; a lane that ends a
//             chain picks its next start cell at once, so the warp stays converged on the one-step loop body;
//       down: resolved roots (pits, drain cells, exit cells whose downstream halo cell is resolved) spread upstream:
//             a thread follows the first child itself and queues the others for the next round;
//     cells whose chain crosses the tile edge stay pending; finished cells are written once, coalesced.
//   * pass 1 visits every tile; a tile that resolves a cell on its edge ACTIVATES the neighbour tile that waits for
//     it; pass p + 1 visits the activated tiles only. One persistent cooperative kernel runs all passes (grid.sync()
//     between them, work lists in HBM); the number of passes is the largest number of tile crossings of a dependency
//     chain (9-25 on the benchmark terrain, where 90 % of the cells resolve in pass 1).
// HBM traffic per cell: dir 1 B + data + value once in pass 1, a few bytes for the revisits.
#pragma once
#include "pfd_common.cuh"
#include "pfd_sweeps.cuh"

#define TS_T 64                 // tile edge
#define TS_S 72                 // shared-memory row stride in cells: 3 pad | left halo | 64 cells | right halo | 3 pad
#define TS_X0 4                 // column of the tile's first cell inside a staged row (word aligned)
#define TS_ROWS (TS_T + 2)
#define TS_N (TS_S * TS_ROWS)   // 4752 staged cells
#define TS_BMW 128              // done-bitmap words per tile (tile-major: word = ly * 2 + (lx >> 5), bit = lx & 31)
#define TS_NCHUNK (TS_T * TS_T / 16)  // a chunk = 16 consecutive cells of one row (one 128-bit vector of bytes); thread t works on
                                     // chunks t, t + NT, ...: row ch >> 2, columns (ch & 3) * 16 .. + 15

#define TSF_DONE 1u   // staged flag plane: value final (resolved by an earlier visit)
#define TSF_SRC 4u    // staged flag plane, down-sweep: value does not depend on the downstream cell (drain cell)
#define TSF_FOREIGN 8u  // staged flag plane: the cell belongs to the neighbour rank's row block (never resolved here; its
                        // done flag and value arrive with the halo exchange)

struct TsCtl {
    unsigned int count[4];           // work-list length of pass p at [p & 3]
    unsigned long long resolved;     // cells resolved so far
    unsigned long long visits;       // tile visits so far
    unsigned int passes;             // passes executed
    unsigned int pad;
};

struct TsArgs {
    const uint8_t* dir;
    long long nrow, ncol;
    int ntx, nty;
    long long own_lo, own_hi;  // rows [own_lo, own_hi) are resolved here (whole raster: 0, nrow); a row block of a larger
                               // raster is extended by the neighbours' edge rows own_lo - 1 / own_hi ("foreign" rows)
    const uint8_t* fdone;      // [2][ncol] done flags of the two foreign rows (row-block sweeps only, else null)
    int first_pass;            // 1: fresh sweep over every tile; > 1: resume with the work list already in list[first_pass & 1]
    int max_passes;      // > 0: stop after that many passes (profiling only: the result is incomplete)
    int al16;            // ncol % 16 == 0 and every raster-sized array is 16-byte aligned: 128-bit global accesses
    uint32_t* done;      // [ntiles][TS_BMW]
    uint32_t* list[2];   // work lists (tile ids)
    uint32_t* stamp;     // last pass a tile was queued for
    TsCtl* ctl;
};

__device__ __forceinline__ int ts_si(int ly, int lx) { return (ly + 1) * TS_S + lx + TS_X0; }
// shared-memory offset of the neighbour in slot k (NW N NE W E SW S SE): -73 -72 -71 -1 1 71 72 73 as packed int8
__device__ __forceinline__ int ts_noff(int k) {
    return (int)(int8_t)(0x494847'01FF'B9B8B7ull >> (8 * k));
}
static_assert(TS_S == 72, "ts_noff is tabulated for a row stride of 72");

// CTA-wide barrier: a one-warp CTA only needs the warp to reconverge (shared memory is ordered by __syncwarp)
template <int NT>
__device__ __forceinline__ void ts_sync() {
    if (NT == 32) __syncwarp();
    else __syncthreads();
}

// halo cell k (0 .. 259): top row, bottom row (66 cells each), left column, right column (64 each) -> (hy, hx) in -1 .. 64
__device__ __forceinline__ void ts_halo_cell(int k, int& hy, int& hx) {
    if (k < TS_T + 2) {
        hy = -1;
        hx = k - 1;
    } else if (k < 2 * (TS_T + 2)) {
        hy = TS_T;
        hx = k - (TS_T + 2) - 1;
    } else if (k < 2 * (TS_T + 2) + TS_T) {
        hy = k - 2 * (TS_T + 2);
        hx = -1;
    } else {
        hy = k - 2 * (TS_T + 2) - TS_T;
        hx = TS_T;
    }
}
#define TS_NHALO (2 * (TS_T + 2) + 2 * TS_T)

__device__ __forceinline__ bool ts_in_raster(const TsArgs& A, long long r, long long c) {
    return r >= 0 && r < A.nrow && c >= 0 && c < A.ncol;
}

// done bit of raster cell (r, c) (must be inside the raster)
__device__ __forceinline__ uint32_t ts_done_bit(const TsArgs& A, long long r, long long c) {
    const long long t = (r >> 6) * A.ntx + (c >> 6);
    const int ly = (int)(r & 63), lx = (int)(c & 63);
    return (__ldcg(A.done + t * TS_BMW + ly * 2 + (lx >> 5)) >> (lx & 31)) & 1u;
}

// 16 bits -> 16 bytes of 0 / 1 (four words)
__device__ __forceinline__ uint32_t ts_spread4(uint32_t b) {  // low 4 bits -> 4 bytes
    return ((b & 1u) | ((b & 2u) << 7) | ((b & 4u) << 14) | ((b & 8u) << 21));
}

// directions of the tile and its halo (cells outside the raster read as nodata) + done flags of earlier visits (pass > 1)
// into the staged byte planes; the tile's own done bits also go to oldbm
template <int NT>
__device__ __forceinline__ void ts_stage_graph(uint8_t* sdir, uint8_t* sflag, uint32_t* oldbm, const TsArgs& A, int tile,
                                               long long r0, long long c0, int pass) {
    for (int ch = threadIdx.x; ch < TS_NCHUNK; ch += NT) {
        const int ly = ch >> 2, lx0 = (ch & 3) << 4;
        const long long r = r0 + ly, c = c0 + lx0;
        uint32_t* dw = reinterpret_cast<uint32_t*>(sdir + ts_si(ly, lx0));
        uint32_t* fw = reinterpret_cast<uint32_t*>(sflag + ts_si(ly, lx0));
        uint32_t w[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
        if (r < A.nrow && c < A.ncol) {
            const uint8_t* p = A.dir + r * A.ncol + c;
            if (A.al16 && c + 15 < A.ncol) {
                const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
                w[0] = v.x, w[1] = v.y, w[2] = v.z, w[3] = v.w;
            } else {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const uint32_t b = (c + j < A.ncol) ? (uint32_t)__ldg(p + j) : (uint32_t)PFD_DIR_NODATA;
                    w[j >> 2] = (w[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (b << (8 * (j & 3)));
                }
            }
        }
        dw[0] = w[0], dw[1] = w[1], dw[2] = w[2], dw[3] = w[3];
        uint32_t bits = 0;
        if (pass > 1) bits = __ldcg(A.done + (long long)tile * TS_BMW + (ch >> 1)) >> (lx0 & 31);
        uint32_t f4[4] = {ts_spread4(bits), ts_spread4(bits >> 4), ts_spread4(bits >> 8), ts_spread4(bits >> 12)};
        if (A.fdone && (r < A.own_lo || r >= A.own_hi) && r < A.nrow && c < A.ncol) {  // a foreign row
            const uint8_t* fd = A.fdone + (r < A.own_lo ? 0 : A.ncol) + c;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const uint32_t b = TSF_FOREIGN | ((c + j < A.ncol && __ldcg(fd + j)) ? TSF_DONE : 0u);
                f4[j >> 2] = (f4[j >> 2] & ~(0xFFu << (8 * (j & 3)))) | (b << (8 * (j & 3)));
            }
        }
        fw[0] = f4[0], fw[1] = f4[1], fw[2] = f4[2], fw[3] = f4[3];
    }
    for (int k = threadIdx.x; k < TS_BMW; k += NT) oldbm[k] = (pass > 1) ? __ldcg(A.done + (long long)tile * TS_BMW + k) : 0u;
    for (int k = threadIdx.x; k < TS_NHALO; k += NT) {
        int hy, hx;
        ts_halo_cell(k, hy, hx);
        const long long r = r0 + hy, c = c0 + hx;
        const bool in = ts_in_raster(A, r, c);
        sdir[ts_si(hy, hx)] = in ? __ldg(A.dir + r * A.ncol + c) : (uint8_t)PFD_DIR_NODATA;
        uint8_t f = (pass > 1 && in) ? (uint8_t)ts_done_bit(A, r, c) : (uint8_t)0;
        if (A.fdone && in && (r < A.own_lo || r >= A.own_hi))
            f = (uint8_t)(TSF_FOREIGN | (__ldcg(A.fdone + (r < A.own_lo ? 0 : A.ncol) + c) ? TSF_DONE : 0u));
        sflag[ts_si(hy, hx)] = f;
    }
}

// 16 consecutive values global -> shared (src == nullptr: zero fill). smem index must be a multiple of 4 elements.
template <typename V, bool CG>
__device__ __forceinline__ void ts_load16(V* sdst, const V* gsrc, bool vec, int nvalid) {
    if (!gsrc) {
        uint32_t* d = reinterpret_cast<uint32_t*>(sdst);
#pragma unroll
        for (int i = 0; i < (int)(4 * sizeof(V)); ++i) d[i] = 0u;
        return;
    }
    if (vec) {
        const uint4* g = reinterpret_cast<const uint4*>(gsrc);
        uint32_t* d = reinterpret_cast<uint32_t*>(sdst);
#pragma unroll
        for (int q = 0; q < (int)sizeof(V); ++q) {
            const uint4 v = CG ? __ldcg(g + q) : __ldg(g + q);
            d[4 * q] = v.x, d[4 * q + 1] = v.y, d[4 * q + 2] = v.z, d[4 * q + 3] = v.w;
        }
    } else {
#pragma unroll 4
        for (int j = 0; j < 16; ++j)
            if (j < nvalid) sdst[j] = CG ? ld_cg(gsrc + j) : gsrc[j];
    }
}

template <typename V>
__device__ __forceinline__ void ts_store16(V* gdst, const V* ssrc, bool vec, int nvalid, uint32_t sel) {
    if (vec && sel == 0xFFFFu) {
        uint4* g = reinterpret_cast<uint4*>(gdst);
        const uint32_t* d = reinterpret_cast<const uint32_t*>(ssrc);
#pragma unroll
        for (int q = 0; q < (int)sizeof(V); ++q) g[q] = make_uint4(d[4 * q], d[4 * q + 1], d[4 * q + 2], d[4 * q + 3]);
    } else {
#pragma unroll 4
        for (int j = 0; j < 16; ++j)
            if (j < nvalid && ((sel >> j) & 1u)) gdst[j] = ssrc[j];
    }
}

// the neighbour tiles recorded in `act` (bit (tyo + 1) * 3 + (txo + 1)) are queued for pass + 1
__device__ __forceinline__ void ts_activate(const TsArgs& A, int tile, uint32_t act, int pass) {
    if (threadIdx.x < 9 && ((act >> threadIdx.x) & 1u)) {
        const int tyo = (int)threadIdx.x / 3 - 1, txo = (int)threadIdx.x % 3 - 1;
        const int ty = tile / A.ntx + tyo, tx = tile % A.ntx + txo;
        if (ty >= 0 && ty < A.nty && tx >= 0 && tx < A.ntx) {
            const uint32_t nt = (uint32_t)(ty * A.ntx + tx);
            if (atomicExch(A.stamp + nt, (uint32_t)(pass + 1)) != (uint32_t)(pass + 1)) {
                const unsigned int pos = atomicAdd(&A.ctl->count[(pass + 1) & 3], 1u);
                A.list[(pass + 1) & 1][pos] = nt;
            }
        }
    }
}

// which neighbour tile holds the cell (y, x) in tile coordinates (a cell outside 0 .. 63)
__device__ __forceinline__ uint32_t ts_act_bit(int y, int x) {
    const int tyo = y < 0 ? 0 : (y >= TS_T ? 2 : 1), txo = x < 0 ? 0 : (x >= TS_T ? 2 : 1);
    return 1u << (tyo * 3 + txo);
}

// Byte-SIMD neighbourhood scan of the 4 cells held by word `wi` of staged row `row` (row 1 .. 64, wi 1 .. 16):
//   ups  (per byte) bit k: the neighbour in slot k drains into the cell
//   pend (per byte) bit k: ... and is not resolved yet (and, NOSRC, is not a source cell)
template <bool FLAGS, bool NOSRC>
__device__ __forceinline__ void ts_scan_word(const uint8_t* sdir, const uint8_t* sflag, int row, int wi, uint32_t& ups_out,
                                             uint32_t& pend_out) {
    const uint32_t* d = reinterpret_cast<const uint32_t*>(sdir) + row * (TS_S / 4) + wi;
    const uint32_t* f = reinterpret_cast<const uint32_t*>(sflag) + row * (TS_S / 4) + wi;
    constexpr int RW = TS_S / 4;
    constexpr uint32_t FM = (FLAGS ? 0x01010101u : 0u) | (NOSRC ? 0x0C0C0C0Cu : 0u);  // (down-sweep: sources and foreign cells are nobody's children)
    uint32_t nd[8], nf[8];
    nd[0] = __byte_perm(d[-RW - 1], d[-RW], 0x6543);
    nd[1] = d[-RW];
    nd[2] = __byte_perm(d[-RW], d[-RW + 1], 0x4321);
    nd[3] = __byte_perm(d[-1], d[0], 0x6543);
    nd[4] = __byte_perm(d[0], d[1], 0x4321);
    nd[5] = __byte_perm(d[RW - 1], d[RW], 0x6543);
    nd[6] = d[RW];
    nd[7] = __byte_perm(d[RW], d[RW + 1], 0x4321);
    if (FM) {
        nf[0] = __byte_perm(f[-RW - 1], f[-RW], 0x6543);
        nf[1] = f[-RW];
        nf[2] = __byte_perm(f[-RW], f[-RW + 1], 0x4321);
        nf[3] = __byte_perm(f[-1], f[0], 0x6543);
        nf[4] = __byte_perm(f[0], f[1], 0x4321);
        nf[5] = __byte_perm(f[RW - 1], f[RW], 0x6543);
        nf[6] = f[RW];
        nf[7] = __byte_perm(f[RW], f[RW + 1], 0x4321);
    }
    uint32_t ups = 0, pend = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t eq = __vcmpeq4(nd[k], splat4(7u - k));
        ups |= eq & splat4(1u << k);
        if (FM) pend |= eq & __vcmpeq4(nf[k] & FM, 0u) & splat4(1u << k);
    }
    ups_out = ups;
    pend_out = FM ? pend : ups;
}

// per byte 0xFF where the cell takes part in this visit: not nodata, not resolved
__device__ __forceinline__ uint32_t ts_live4(uint32_t dirw, uint32_t flagw) {
    return ~__vcmpeq4(dirw, 0xFFFFFFFFu) & __vcmpeq4(flagw & 0x09090909u, 0u);  // not nodata, not done, not foreign
}

// upstream slots that lie in the halo, per byte, for the 4 cells of word q (0 .. 3) of the chunk at (ly, lx0)
__device__ __forceinline__ uint32_t ts_halo_slots4(int ly, int lx0, int q) {
    uint32_t m = (ly == 0 ? 0x07u : 0u) | (ly == TS_T - 1 ? 0xE0u : 0u);
    uint32_t w = splat4(m);
    if (lx0 == 0 && q == 0) w |= 0x29u;                        // NW, W, SW of the first column
    if (lx0 == TS_T - 16 && q == 3) w |= 0x94u << 24;          // NE, E, SE of the last column
    return w;
}

// per-byte population count
__device__ __forceinline__ uint32_t ts_popc4(uint32_t x) {
    x = x - ((x >> 1) & 0x55555555u);
    x = (x & 0x33333333u) + ((x >> 2) & 0x33333333u);
    return (x + (x >> 4)) & 0x0F0F0F0Fu;
}

// four bytes 0x00 / 0xFF -> four bits
__device__ __forceinline__ uint32_t ts_gather4(uint32_t b) {
    return (b & 1u) | ((b >> 7) & 2u) | ((b >> 14) & 4u) | ((b >> 21) & 8u);
}

// Warp-aggregated append: every lane of the (converged) warp calls it; lanes with `want` get consecutive positions
// base + (*counter before) .. One shared-memory atomic per warp.
__device__ __forceinline__ void ts_push(uint16_t* q, uint32_t base, uint32_t* counter, bool want, int cell) {
    const uint32_t m = __ballot_sync(0xFFFFFFFFu, want);
    if (m) {
        const int lane = threadIdx.x & 31;
        uint32_t pos = 0;
        if (lane == __ffs(m) - 1) pos = atomicAdd(counter, (uint32_t)__popc(m));
        pos = __shfl_sync(0xFFFFFFFFu, pos, __ffs(m) - 1);
        if (want) q[base + pos + __popc(m & ((1u << lane) - 1u))] = (uint16_t)cell;
    }
}

// Shared state of one tile visit. The staged byte planes (directions, flags) are only needed until the per-cell records
// are built; the frontier queue takes their place afterwards.
//   up   rec = ups mask (8) | dir nibble (4) | number of pending upstream neighbours (4)
//   down rec = unresolved in-tile children mask (8) | unresolved halo children mask (8)
template <typename V, bool AUX>
struct TsShared {
    V val[TS_N];
    uint16_t rec[TS_N];
    union {
        struct {
            uint8_t dir[TS_N];
            uint8_t flag[TS_N];
        } pl;
        uint16_t q[TS_T * TS_T];  // cells (ly * 64 + lx) in the order they became ready: round r works on q[lo_r, lo_r + cnt[r % 3])
    } u;
    uint8_t aux[AUX ? TS_N : 16];  // Op-specific byte per cell (Strahler: mask)
    uint32_t oldbm[TS_BMW];        // resolved by earlier visits
    uint32_t newbm[TS_BMW];        // resolved by this visit
    uint32_t cnt[3];
    uint32_t act;
    uint32_t newly;
    int ylo, yhi;                  // tile rows [ylo, yhi) are owned (0, 64 unless the tile holds a foreign row)
};

// write back what this visit resolved (pass 1: every cell), publish the done bitmap, queue the neighbour tiles
template <int NT, typename V, bool AUX, class Op>
__device__ __forceinline__ void ts_finish(TsShared<V, AUX>& s, const TsArgs& A, const Op& op, int tile, int pass, long long r0,
                                          long long c0, bool fill_unresolved) {
    uint32_t cnt = 0;
    for (int ch = threadIdx.x; ch < TS_NCHUNK; ch += NT) {
        const int ly = ch >> 2, lx0 = (ch & 3) << 4;
        const long long r = r0 + ly, c = c0 + lx0;
        if (r < A.nrow && c < A.ncol) {
            const int nvalid = (int)min((long long)16, A.ncol - c);
            const bool vec = A.al16 && nvalid == 16;
            uint32_t sel = (s.newbm[ch >> 1] >> (lx0 & 31)) & 0xFFFFu;
            cnt += __popc(sel);
            V* sv = &s.val[ts_si(ly, lx0)];
            if (pass == 1) {
                if (fill_unresolved) {
#pragma unroll 4
                    for (int j = 0; j < 16; ++j)
                        if (!((sel >> j) & 1u)) sv[j] = op.fill();
                }
                sel = 0xFFFFu;
            }
            if (sel) ts_store16<V>(op.out + (r * A.ncol + c), sv, vec, nvalid, sel);
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s.newly, cnt);
    __threadfence();  // a tile running in the same pass may read my done bits: the values are out before the bits
    ts_sync<NT>();
    for (int k = threadIdx.x; k < TS_BMW; k += NT) A.done[(long long)tile * TS_BMW + k] = s.oldbm[k] | s.newbm[k];
    ts_activate(A, tile, s.act, pass);
    if (threadIdx.x == 16) {
        if (s.newly) atomicAdd(&A.ctl->resolved, (unsigned long long)s.newly);
        atomicAdd(&A.ctl->visits, 1ull);
    }
    ts_sync<NT>();  // shared memory is reused by the next visit
}

// ---------------------------------------------------------------------------------------------------------
// Up-sweep. Op:  typedef V; static const bool AUX;
//   const V* init_src()            array holding the value of a cell before anything was added (accuflux: data), nullptr = 0
//   const uint8_t* aux_src()       (AUX) byte staged per cell incl. halo
//   State begin(own, own_aux); step(State&, v_up, aux_up) for the upstream neighbours in DESCENDING linear index; V end(State, own_aux)
//   V* out
// ---------------------------------------------------------------------------------------------------------
// One cell of the frontier: fold the upstream values into the cell, mark it resolved, and take one off the pending
// count of the downstream cell. Returns the downstream cell when this lane was the LAST ARRIVER there (it joins the
// next round), -1 otherwise (pit, tile edge, or another lane arrives later).
template <class Op>
__device__ __forceinline__ int ts_up_step(TsShared<typename Op::V, Op::AUX>& s, const Op& op, int i) {
    const int ly = i >> 6, lx = i & (TS_T - 1);
    const int c = ts_si(ly, lx);
    const uint32_t r = s.rec[c];
    uint32_t m = r & 0xFFu;
    const uint32_t d = (r >> 8) & 0xFu;
    const uint8_t own_aux = Op::AUX ? s.aux[c] : (uint8_t)0;
    typename Op::State st = op.begin(s.val[c], own_aux);
    while (m) {  // descending slot = descending linear index
        const int k = 31 - __clz(m);
        m ^= 1u << k;
        const int u = c + ts_noff(k);
        op.step(st, s.val[u], Op::AUX ? s.aux[u] : (uint8_t)0);
    }
    s.val[c] = op.end(st, own_aux);
    atomicOr(&s.newbm[i >> 5], 1u << (i & 31));
    int next = -1;
    if (d < 8u) {  // (a pit ends the chain)
        const int y = ly + pfd_slot_dr((int)d), x = lx + pfd_slot_dc((int)d);
        if ((unsigned)(y - s.ylo) < (unsigned)(s.yhi - s.ylo) && (unsigned)x < (unsigned)TS_T) {
            const int ds = c + ts_noff((int)d);
            const unsigned sh = 16u * (ds & 1);
            const uint32_t old = atomicSub(reinterpret_cast<uint32_t*>(s.rec) + (ds >> 1), 0x1000u << sh);
            if (((old >> (sh + 12)) & 0xFu) == 1u) next = y * TS_T + x;  // last arriver (else somebody else continues)
        } else if (!((unsigned)y < (unsigned)TS_T && (unsigned)x < (unsigned)TS_T)) {
            atomicOr(&s.act, ts_act_bit(y, x));  // leaves the tile: the neighbour tile may continue in the next pass
        }  // (else: a foreign row inside the tile -- the neighbour rank continues after the halo exchange)
    }
    return next;
}

template <int NT, class Op>
__device__ __forceinline__ void ts_up_visit(TsShared<typename Op::V, Op::AUX>& s, const TsArgs& A, const Op& op, int tile, int pass) {
    typedef typename Op::V V;
    const long long r0 = (long long)(tile / A.ntx) * TS_T, c0 = (long long)(tile % A.ntx) * TS_T;
    if (threadIdx.x == 0) {
        s.act = 0;
        s.newly = 0;
        s.cnt[0] = s.cnt[1] = s.cnt[2] = 0;
        s.ylo = (int)max(0ll, A.own_lo - r0);
        s.yhi = (int)min((long long)TS_T, A.own_hi - r0);
    }
    for (int k = threadIdx.x; k < TS_BMW; k += NT) s.newbm[k] = 0u;
    ts_stage_graph<NT>(s.u.pl.dir, s.u.pl.flag, s.oldbm, A, tile, r0, c0, pass);
    // values: pass 1 starts from the cell's own datum; later passes read what earlier visits stored (final for done
    // cells, the own datum for pending ones)
    for (int ch = threadIdx.x; ch < TS_NCHUNK; ch += NT) {
        const int ly = ch >> 2, lx0 = (ch & 3) << 4;
        const long long r = r0 + ly, c = c0 + lx0;
        if (r < A.nrow && c < A.ncol) {
            const int nvalid = (int)min((long long)16, A.ncol - c);
            const bool vec = A.al16 && nvalid == 16;
            const long long g0 = r * A.ncol + c;
            const int base = ts_si(ly, lx0);
            if (pass == 1 && r >= A.own_lo && r < A.own_hi)
                ts_load16<V, false>(&s.val[base], op.init_src() ? op.init_src() + g0 : nullptr, vec, nvalid);
            else ts_load16<V, true>(&s.val[base], op.out + g0, vec, nvalid);  // (a foreign row: what the neighbour sent)
            if (Op::AUX) ts_load16<uint8_t, false>(&s.aux[base], op.aux_src() + g0, vec, nvalid);
        }
    }
    ts_sync<NT>();  // dir + flags (incl. halo) are staged
    if (pass > 1 || Op::AUX || A.fdone) {
        for (int k = threadIdx.x; k < TS_NHALO; k += NT) {
            int hy, hx;
            ts_halo_cell(k, hy, hx);
            const long long hr = r0 + hy, hc = c0 + hx;
            if (ts_in_raster(A, hr, hc)) {
                const long long g = hr * A.ncol + hc;
                if (s.u.pl.flag[ts_si(hy, hx)] & TSF_DONE) s.val[ts_si(hy, hx)] = ld_cg(op.out + g);
                if (Op::AUX) s.aux[ts_si(hy, hx)] = __ldg(op.aux_src() + g);
            }
        }
    }
    // per-cell records (4 cells per word of the planes); the ready cells are remembered in registers until the planes are dead
    uint32_t start[(TS_NCHUNK / NT + 1) / 2];
#pragma unroll
    for (int t = 0; t < (TS_NCHUNK / NT + 1) / 2; ++t) start[t] = 0;
#pragma unroll
    for (int it = 0; it < TS_NCHUNK / NT; ++it) {
        const int ch = threadIdx.x + it * NT;
        const int ly = ch >> 2, lx0 = (ch & 3) << 4;
        const int base = ts_si(ly, lx0);
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(s.u.pl.dir + base);
        const uint32_t* fw = reinterpret_cast<const uint32_t*>(s.u.pl.flag + base);
        uint32_t* rw = reinterpret_cast<uint32_t*>(s.rec + base);
        uint32_t st16 = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t ups, pend;
            if (pass == 1) ts_scan_word<false, false>(s.u.pl.dir, s.u.pl.flag, ly + 1, (TS_X0 + lx0) / 4 + q, ups, pend);
            else ts_scan_word<true, false>(s.u.pl.dir, s.u.pl.flag, ly + 1, (TS_X0 + lx0) / 4 + q, ups, pend);
            const uint32_t live = ts_live4(dw[q], fw[q]);
            ups &= live;
            pend &= live;
            const uint32_t hi = (dw[q] & 0x0F0F0F0Fu) | (ts_popc4(pend) << 4);  // dir nibble | pending count
            rw[2 * q] = __byte_perm(ups, hi, 0x5140);
            rw[2 * q + 1] = __byte_perm(ups, hi, 0x7362);
            st16 |= ts_gather4(live & __vcmpeq4(pend, 0u)) << (4 * q);
        }
        start[it >> 1] |= st16 << (16 * (it & 1));
    }
    ts_sync<NT>();  // the planes are dead: the queue takes their place
    // In-tile dataflow, level by level: round 0 holds the ready cells (no pending upstream neighbour); a lane resolves
    // ONE cell per round and, when it was the last arriver at the downstream cell, appends that cell to the next round
    // (warp-aggregated). The frontier stays compacted, so the lanes of a warp all work, and the long chains (rivers) of
    // the tile end up side by side in one warp instead of one per warp.
#pragma unroll
    for (int it = 0; it < TS_NCHUNK / NT; ++it) {
        const int ch = threadIdx.x + it * NT;
        const uint32_t st16 = (start[it >> 1] >> (16 * (it & 1))) & 0xFFFFu;
        if (__any_sync(0xFFFFFFFFu, st16 != 0u)) {
#pragma unroll 1
            for (int j = 0; j < 16; ++j) ts_push(s.u.q, 0u, &s.cnt[0], (st16 >> j) & 1u, ch * 16 + j);
        }
    }
    uint32_t lo = 0;
    for (int rd = 0;; ++rd) {
        ts_sync<NT>();
        const uint32_t n = s.cnt[rd % 3];
        if (n == 0) break;
        if (threadIdx.x == 0) s.cnt[(rd + 2) % 3] = 0;
        uint32_t* cn = &s.cnt[(rd + 1) % 3];
        for (uint32_t e0 = threadIdx.x & ~31u; e0 < n; e0 += NT) {  // warp-uniform trip count
            const uint32_t e = e0 + (threadIdx.x & 31u);
            int next = -1;
            if (e < n) next = ts_up_step<Op>(s, op, s.u.q[lo + e]);
            ts_push(s.u.q, lo + n, cn, next >= 0, next);
        }
        lo += n;
    }
    // pass 1 writes every cell of the tile (pending and nodata cells keep their own datum, like accu = data.copy()),
    // later passes only what this visit resolved
    ts_finish<NT>(s, A, op, tile, pass, r0, c0, false);
}

// All passes in one cooperative launch. grid-stride over the work list of the pass (pass 1: every tile).
template <int NT, class Op>
__global__ void __launch_bounds__(NT, 1024 / NT) tile_up_sweep_kernel(TsArgs A, Op op) {  // (at most 64 registers: occupancy)
    extern __shared__ __align__(16) unsigned char ts_smem_raw[];
    TsShared<typename Op::V, Op::AUX>& s = *reinterpret_cast<TsShared<typename Op::V, Op::AUX>*>(ts_smem_raw);
    cg::grid_group grid = cg::this_grid();
    const unsigned int ntiles = (unsigned int)(A.ntx * A.nty);
    int pass = A.first_pass;
    for (;; ++pass) {
        const unsigned int count = (pass == 1) ? ntiles : __ldcg(&A.ctl->count[pass & 3]);
        if (count == 0 || (A.max_passes > 0 && pass > A.max_passes)) break;
        if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->count[(pass + 2) & 3] = 0u;  // last read two passes ago
        const uint32_t* list = A.list[pass & 1];
        for (unsigned int w = blockIdx.x; w < count; w += gridDim.x)
            ts_up_visit<NT, Op>(s, A, op, (pass == 1) ? (int)w : (int)__ldcg(list + w), pass);
        __threadfence();
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->passes = (unsigned int)(pass - 1);
}

// streams.accuflux (up): accu = data.copy(); accu[ds] += accu[i] in seq[::-1] order, nodata-guarded on the running sum
template <typename T>
struct AccuUpTileOp {
    typedef T V;
    typedef T State;
    static const bool AUX = false;
    const T* data;
    T* out;
    NoData nd;
    __host__ __device__ __forceinline__ const T* init_src() const { return data; }
    __host__ __device__ __forceinline__ const uint8_t* aux_src() const { return nullptr; }
    __device__ __forceinline__ T init(long long g) const { return data[g]; }
    __device__ __forceinline__ T fill() const { return T(); }
    __device__ __forceinline__ T begin(T own, uint8_t) const { return own; }
    __device__ __forceinline__ void step(T& acc, T up, uint8_t) const {
        if (not_nodata(acc, nd) && not_nodata(up, nd)) acc = acc_add(acc, up);
    }
    __device__ __forceinline__ T end(T acc, uint8_t) const { return acc; }
};

// streams.strahler_order (streams.py:250-269): masked-out cells neither push nor count as headwaters
template <bool MASKED>
struct StrahlerTileOp {
    typedef uint8_t V;
    static const bool AUX = MASKED;
    const uint8_t* mask;
    uint8_t* out;
    struct State {
        uint8_t so, smax;
    };
    __host__ __device__ __forceinline__ const uint8_t* init_src() const { return nullptr; }
    __host__ __device__ __forceinline__ const uint8_t* aux_src() const { return mask; }
    __device__ __forceinline__ uint8_t init(long long) const { return 0; }
    __device__ __forceinline__ uint8_t fill() const { return 0; }
    __device__ __forceinline__ State begin(uint8_t, uint8_t) const { return State{0, 0}; }
    __device__ __forceinline__ void step(State& st, uint8_t sto, uint8_t up_aux) const {
        if (MASKED && !up_aux) return;
        if (st.so < sto) st.so = sto;
        else if (sto == st.so && st.smax == sto) st.so = (uint8_t)(st.so + 1);
        if (st.smax < sto) st.smax = sto;
    }
    __device__ __forceinline__ uint8_t end(State st, uint8_t own_aux) const {
        return ((!MASKED || own_aux) && st.so == 0) ? (uint8_t)1 : st.so;  // headwater
    }
};

// cells that drain to no pit (on / above a loop) are outside `seq`: they keep their initial value. The dataflow does
// resolve the trees hanging on a loop, so they are reset from the rank (only run when loops exist).
template <class Op>
__global__ void ts_reset_unranked_kernel(const uint8_t* __restrict__ dir, const int32_t* __restrict__ rank, int64_t n, Op op) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (dir[i] != PFD_DIR_NODATA && rank[i] < 0) op.out[i] = op.init(i);
}

// ---------------------------------------------------------------------------------------------------------
// Down-sweep. Op:  typedef V;
//   V prep(g, d, g_ds)      per-cell term staged in val[] before the sweep (HAND: (double)(elevtn[g] - elevtn[g_ds]))
//   bool source(g)          the cell's value does not depend on its downstream cell (HAND: drain cell)
//   V source_value()
//   V pit(term)             value of a pit that is not a source
//   V down(v_ds, term)      value of a cell from the value of its downstream cell
//   V fill()                value of the cells the sweep never reaches (nodata, cells draining to no pit)
//   V* out
// ---------------------------------------------------------------------------------------------------------
// One entry of the round (called by all 32 lanes of a warp): the lane's resolved cell resolves its unresolved in-tile
// children and appends them to the next round (warp-aggregated, so the frontier stays compacted).
template <class Op>
__device__ __forceinline__ void ts_down_expand(TsShared<typename Op::V, false>& s, const Op& op, uint32_t lo, uint32_t cnt, uint32_t e,
                                               uint32_t* qn) {
    typedef typename Op::V V;
    uint32_t m = 0;
    int ly = 0, lx = 0, p = 0;
    V vp = V();
    if (e < cnt) {
        const int i = s.u.q[lo + e];
        ly = i >> 6, lx = i & (TS_T - 1);
        p = ts_si(ly, lx);
        const uint32_t r = s.rec[p];
        m = r & 0xFFu;
        vp = s.val[p];
        uint32_t om = r >> 8;
        while (om) {  // a neighbour tile waits for this cell
            const int k = __ffs(om) - 1;
            om &= om - 1;
            atomicOr(&s.act, ts_act_bit(ly + pfd_slot_dr(k), lx + pfd_slot_dc(k)));
        }
    }
    while (__any_sync(0xFFFFFFFFu, m != 0u)) {
        int n = -1;
        if (m) {
            const int k = __ffs(m) - 1;
            m &= m - 1;
            const int c = p + ts_noff(k);
            s.val[c] = op.down(vp, s.val[c]);
            n = (ly + pfd_slot_dr(k)) * TS_T + lx + pfd_slot_dc(k);
            atomicOr(&s.newbm[n >> 5], 1u << (n & 31));
        }
        ts_push(s.u.q, lo + cnt, qn, n >= 0, n);
    }
}

template <int NT, class Op>
__device__ __forceinline__ void ts_down_visit(TsShared<typename Op::V, false>& s, const TsArgs& A, const Op& op, int tile, int pass) {
    typedef typename Op::V V;
    const long long r0 = (long long)(tile / A.ntx) * TS_T, c0 = (long long)(tile % A.ntx) * TS_T;
    if (threadIdx.x == 0) {
        s.act = 0;
        s.newly = 0;
        s.cnt[0] = s.cnt[1] = s.cnt[2] = 0;
    }
    for (int k = threadIdx.x; k < TS_BMW; k += NT) s.newbm[k] = 0u;
    ts_stage_graph<NT>(s.u.pl.dir, s.u.pl.flag, s.oldbm, A, tile, r0, c0, pass);
    ts_sync<NT>();
    // per-cell terms of the unresolved cells, source flags; values of the resolved halo cells
    for (int ch = threadIdx.x; ch < TS_NCHUNK; ch += NT) {
        const int ly = ch >> 2, lx0 = (ch & 3) << 4;
        const int base = ts_si(ly, lx0);
        const long long g0 = (r0 + ly) * A.ncol + c0 + lx0;
#pragma unroll 4
        for (int j = 0; j < 16; ++j) {
            const uint32_t d = s.u.pl.dir[base + j];
            if (d != PFD_DIR_NODATA && !(s.u.pl.flag[base + j] & (TSF_DONE | TSF_FOREIGN))) {  // (own, unresolved cells only)
                const long long g = g0 + j;
                s.val[base + j] = op.prep(g, d, (d < 8u) ? g + pfd_slot_off((int)d, A.ncol) : g);
                if (op.source(g)) s.u.pl.flag[base + j] |= (uint8_t)TSF_SRC;
            }
        }
    }
    if (pass > 1 || A.fdone) {
        for (int k = threadIdx.x; k < TS_NHALO; k += NT) {
            int hy, hx;
            ts_halo_cell(k, hy, hx);
            if (s.u.pl.flag[ts_si(hy, hx)] & TSF_DONE) s.val[ts_si(hy, hx)] = ld_cg(op.out + ((r0 + hy) * A.ncol + c0 + hx));
        }
        if (A.fdone) {  // resolved foreign cells inside the tile: what the neighbour sent
            for (int ch = threadIdx.x; ch < TS_NCHUNK; ch += NT) {
                const int ly = ch >> 2, lx0 = (ch & 3) << 4;
                const long long r = r0 + ly;
                if (r >= A.own_lo && r < A.own_hi) continue;
#pragma unroll 4
                for (int j = 0; j < 16; ++j)
                    if (s.u.pl.flag[ts_si(ly, lx0 + j)] & TSF_DONE) s.val[ts_si(ly, lx0 + j)] = ld_cg(op.out + (r * A.ncol + c0 + lx0 + j));
            }
        }
    }
    ts_sync<NT>();
    // records: unresolved children inside the tile / in the halo (sources are nobody's children: they are roots).
    // Roots: sources, pits, exit cells whose downstream (halo) cell is resolved -- resolved right here, queued below.
    uint32_t roots[(TS_NCHUNK / NT + 1) / 2];
#pragma unroll
    for (int t = 0; t < (TS_NCHUNK / NT + 1) / 2; ++t) roots[t] = 0;
#pragma unroll
    for (int it = 0; it < TS_NCHUNK / NT; ++it) {
        const int ch = threadIdx.x + it * NT;
        const int ly = ch >> 2, lx0 = (ch & 3) << 4;
        const int base = ts_si(ly, lx0);
        const uint32_t* dw = reinterpret_cast<const uint32_t*>(s.u.pl.dir + base);
        const uint32_t* fw = reinterpret_cast<const uint32_t*>(s.u.pl.flag + base);
        uint32_t* rw = reinterpret_cast<uint32_t*>(s.rec + base);
        uint32_t live16 = 0;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t ups, pend;
            if (pass == 1) ts_scan_word<false, true>(s.u.pl.dir, s.u.pl.flag, ly + 1, (TS_X0 + lx0) / 4 + q, ups, pend);
            else ts_scan_word<true, true>(s.u.pl.dir, s.u.pl.flag, ly + 1, (TS_X0 + lx0) / 4 + q, ups, pend);
            const uint32_t halo = ts_halo_slots4(ly, lx0, q);
            const uint32_t kids = pend & ~halo, outm = pend & halo;
            rw[2 * q] = __byte_perm(kids, outm, 0x5140);
            rw[2 * q + 1] = __byte_perm(kids, outm, 0x7362);
            live16 |= ts_gather4(ts_live4(dw[q], fw[q])) << (4 * q);
        }
        uint32_t root16 = 0;
        while (live16) {
            const int j = __ffs(live16) - 1;
            live16 &= live16 - 1;
            const int ci = base + j;
            const uint32_t d = s.u.pl.dir[ci], f = s.u.pl.flag[ci];
            bool root = false;
            V v = s.val[ci];
            if (f & TSF_SRC) {
                v = op.source_value();
                root = true;
            } else if (d >= 8u) {
                v = op.pit(v);
                root = true;
            } else {
                // the downstream cell was resolved BEFORE this visit: it lies in the halo or in a foreign row (an own cell
                // of the tile resolves its children in the visit that resolves it)
                const int ds = ci + ts_noff((int)d);
                if (s.u.pl.flag[ds] & TSF_DONE) {
                    v = op.down(s.val[ds], v);
                    root = true;
                }
            }
            if (root) {
                s.val[ci] = v;
                root16 |= 1u << j;
            }
        }
        roots[it >> 1] |= root16 << (16 * (it & 1));
    }
    ts_sync<NT>();  // the planes are dead: the queue takes their place
#pragma unroll
    for (int it = 0; it < TS_NCHUNK / NT; ++it) {
        const int ch = threadIdx.x + it * NT;
        const uint32_t r16 = (roots[it >> 1] >> (16 * (it & 1))) & 0xFFFFu;
        if (r16) atomicOr(&s.newbm[ch >> 1], r16 << ((ch & 1) * 16));
        if (__any_sync(0xFFFFFFFFu, r16 != 0u)) {
#pragma unroll 1
            for (int j = 0; j < 16; ++j) ts_push(s.u.q, 0u, &s.cnt[0], (r16 >> j) & 1u, ch * 16 + j);
        }
    }
    // rounds, level by level: a lane takes ONE resolved cell per round, resolves its unresolved in-tile children and
    // appends them to the next round (warp-aggregated, so the frontier stays compacted)
    uint32_t lo = 0;
    for (int rd = 0;; ++rd) {
        ts_sync<NT>();
        const uint32_t cnt = s.cnt[rd % 3];
        if (cnt == 0) break;
        if (threadIdx.x == 0) s.cnt[(rd + 2) % 3] = 0;
        uint32_t* qn = &s.cnt[(rd + 1) % 3];
        for (uint32_t e0 = threadIdx.x & ~31u; e0 < cnt; e0 += NT)  // warp-uniform trip count
            ts_down_expand<Op>(s, op, lo, cnt, e0 + (threadIdx.x & 31u), qn);
        lo += cnt;
    }
    // pass 1 writes every cell: what is not resolved (yet, or never) holds the fill value
    ts_finish<NT>(s, A, op, tile, pass, r0, c0, true);
}

template <int NT, class Op>
__global__ void __launch_bounds__(NT, 1024 / NT) tile_down_sweep_kernel(TsArgs A, Op op) {
    extern __shared__ __align__(16) unsigned char ts_smem_raw[];
    TsShared<typename Op::V, false>& s = *reinterpret_cast<TsShared<typename Op::V, false>*>(ts_smem_raw);
    cg::grid_group grid = cg::this_grid();
    const unsigned int ntiles = (unsigned int)(A.ntx * A.nty);
    int pass = A.first_pass;
    for (;; ++pass) {
        const unsigned int count = (pass == 1) ? ntiles : __ldcg(&A.ctl->count[pass & 3]);
        if (count == 0 || (A.max_passes > 0 && pass > A.max_passes)) break;
        if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->count[(pass + 2) & 3] = 0u;
        const uint32_t* list = A.list[pass & 1];
        for (unsigned int w = blockIdx.x; w < count; w += gridDim.x)
            ts_down_visit<NT, Op>(s, A, op, (pass == 1) ? (int)w : (int)__ldcg(list + w), pass);
        __threadfence();
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->passes = (unsigned int)(pass - 1);
}

// dem.height_above_nearest_drain (dem.py:316-329): hand[i] = hand[ds] + (elevtn[i] - elevtn[ds]); drain cells 0
template <typename T>
struct HandTileOp {
    typedef double V;
    const uint8_t* drain;
    const T* elevtn;
    double* out;
    __device__ __forceinline__ double prep(long long g, uint32_t, long long g_ds) const {
        return (double)elev_sub<T>(__ldg(elevtn + g), __ldg(elevtn + g_ds));  // difference in elevtn's dtype, sum in float64
    }
    __device__ __forceinline__ bool source(long long g) const { return __ldg(drain + g) == 1; }
    __device__ __forceinline__ double source_value() const { return 0.0; }
    __device__ __forceinline__ double pit(double dz) const { return __dadd_rn(0.0, dz); }  // reads its own initial 0 (dem.py:323)
    __device__ __forceinline__ double down(double v_ds, double dz) const { return __dadd_rn(v_ds, dz); }
    __device__ __forceinline__ double fill() const { return -9999.0; }
};
