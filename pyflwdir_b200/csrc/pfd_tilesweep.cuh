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
//     cells, and the cell data; then
//       up:   every thread starts at its ready cells (no pending upstream neighbour) and walks downstream for as long
//             as it is the LAST ARRIVER at the next cell (one shared-memory atomic clears its bit in the cell's
//             pending mask) -- barrier-free dataflow, float sums bit-exact without float atomics;
//       down: resolved roots (pits, drain cells, exit cells whose downstream halo cell is resolved) spread upstream:
//             a thread follows the first child itself and queues the others for the next round;
//     cells whose chain crosses the tile edge stay pending; finished cells are written once, coalesced.
//   * pass 1 visits every tile; a tile that resolves a cell on its edge ACTIVATES the neighbour tile that waits for
//     it; pass p + 1 visits the activated tiles only. One persistent cooperative kernel runs all passes (grid.sync()
//     between them, work lists in HBM); the number of passes is the largest number of tile crossings of a dependency
//     chain (9-12 on the benchmark terrain, where 90 % of the cells resolve in pass 1 and a tile is visited 3.7 times
//     on average).
// HBM traffic per cell: dir 1 B + data + value once in pass 1, a few bytes for the revisits.
#pragma once
#include "pfd_common.cuh"
#include "pfd_sweeps.cuh"

#define TS_T 64                 // tile edge
#define TS_S 66                 // shared-memory row stride: tile + one-cell halo
#define TS_N (TS_S * TS_S)      // 4356 staged cells
#define TS_THREADS 256
#define TS_BMW 128              // done-bitmap words per tile (tile-major: word = ly * 2 + (lx >> 5), bit = lx & 31)
#define TS_CPT (TS_T * TS_T / TS_THREADS)  // own cells per thread

#define TSF_DONE 1u   // value final
#define TSF_NEW 2u    // ... and computed in this visit
#define TSF_SRC 4u    // down-sweep: value does not depend on the downstream cell (drain cell)

struct TsCtl {
    unsigned int count[4];           // work-list length of pass p at [p & 3]
    unsigned long long resolved;     // cells resolved so far
    unsigned int passes;             // passes executed
    unsigned int pad;
};

struct TsArgs {
    const uint8_t* dir;
    long long nrow, ncol;
    int ntx, nty;
    uint32_t* done;      // [ntiles][TS_BMW]
    uint32_t* list[2];   // work lists (tile ids)
    uint32_t* stamp;     // last pass a tile was queued for
    TsCtl* ctl;
};

__device__ __forceinline__ int ts_si(int ly, int lx) { return (ly + 1) * TS_S + lx + 1; }
__device__ __forceinline__ int ts_noff(int k) { return pfd_slot_dr(k) * TS_S + pfd_slot_dc(k); }

// own cell j of thread t: consecutive threads own consecutive columns (coalesced global rows)
#define TS_OWN(j, ly, lx)                                  \
    const int ts_i__ = (int)threadIdx.x + TS_THREADS * (j); \
    const int ly = ts_i__ >> 6, lx = ts_i__ & (TS_T - 1)

// halo cell k (0 .. 259): top row, bottom row (66 cells each), left column, right column (64 each) -> (hy, hx) in -1 .. 64
__device__ __forceinline__ void ts_halo_cell(int k, int& hy, int& hx) {
    if (k < TS_S) {
        hy = -1;
        hx = k - 1;
    } else if (k < 2 * TS_S) {
        hy = TS_T;
        hx = k - TS_S - 1;
    } else if (k < 2 * TS_S + TS_T) {
        hy = k - 2 * TS_S;
        hx = -1;
    } else {
        hy = k - 2 * TS_S - TS_T;
        hx = TS_T;
    }
}
#define TS_NHALO (2 * TS_S + 2 * TS_T)

__device__ __forceinline__ bool ts_in_raster(const TsArgs& A, long long r, long long c) {
    return r >= 0 && r < A.nrow && c >= 0 && c < A.ncol;
}

// done bit of raster cell (r, c) (must be inside the raster)
__device__ __forceinline__ uint32_t ts_done_bit(const TsArgs& A, long long r, long long c) {
    const long long t = (r >> 6) * A.ntx + (c >> 6);
    const int ly = (int)(r & 63), lx = (int)(c & 63);
    return (__ldcg(A.done + t * TS_BMW + ly * 2 + (lx >> 5)) >> (lx & 31)) & 1u;
}

// directions of the tile and its halo (cells outside the raster read as nodata) + done flags (pass > 1)
__device__ __forceinline__ void ts_stage_graph(uint8_t* sdir, uint8_t* sflag, const TsArgs& A, long long r0, long long c0, int pass) {
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const long long r = r0 + ly, c = c0 + lx;
        const bool in = r < A.nrow && c < A.ncol;
        sdir[ts_si(ly, lx)] = in ? __ldg(A.dir + r * A.ncol + c) : (uint8_t)PFD_DIR_NODATA;
        sflag[ts_si(ly, lx)] = (pass > 1 && in) ? (uint8_t)ts_done_bit(A, r, c) : (uint8_t)0;
    }
    for (int k = threadIdx.x; k < TS_NHALO; k += TS_THREADS) {
        int hy, hx;
        ts_halo_cell(k, hy, hx);
        const long long r = r0 + hy, c = c0 + hx;
        const bool in = ts_in_raster(A, r, c);
        sdir[ts_si(hy, hx)] = in ? __ldg(A.dir + r * A.ncol + c) : (uint8_t)PFD_DIR_NODATA;
        sflag[ts_si(hy, hx)] = (pass > 1 && in) ? (uint8_t)ts_done_bit(A, r, c) : (uint8_t)0;
    }
}

// done bitmap of the tile from the flags; returns nothing. 128 words, one per thread t < 128.
__device__ __forceinline__ void ts_store_bitmap(const uint8_t* sflag, const TsArgs& A, int tile) {
    if (threadIdx.x < TS_BMW) {
        const int ly = threadIdx.x >> 1, x0 = (threadIdx.x & 1) * 32;
        uint32_t w = 0;
#pragma unroll 8
        for (int b = 0; b < 32; ++b) w |= (uint32_t)(sflag[ts_si(ly, x0 + b)] & TSF_DONE) << b;
        A.done[(long long)tile * TS_BMW + threadIdx.x] = w;
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

// which neighbour tile holds halo position (hy, hx)
__device__ __forceinline__ uint32_t ts_act_bit(int hy, int hx) {
    const int tyo = hy < 0 ? 0 : (hy >= TS_T ? 2 : 1), txo = hx < 0 ? 0 : (hx >= TS_T ? 2 : 1);
    return 1u << (tyo * 3 + txo);
}

template <typename V>
struct TsSharedUp {
    V val[TS_N];
    uint32_t pendw[TS_N / 4];  // per cell (byte): upstream neighbours still pending; cleared with shared-memory atomics
    uint8_t ups[TS_N];         // per cell: all upstream neighbours (bit k = slot k)
    uint8_t dir[TS_N];
    uint8_t flag[TS_N];
    uint8_t aux[TS_N];         // Op-specific byte per cell (Strahler: mask)
    uint32_t act;
    uint32_t newly;
};

// ---------------------------------------------------------------------------------------------------------
// Up-sweep. Op:  typedef V; static const bool AUX;
//   V init(g)                      value of a cell before anything was added (accuflux: data[g]; Strahler: 0)
//   uint8 aux(g)                   (AUX) byte staged per cell incl. halo
//   State begin(own, own_aux); step(State&, v_up, aux_up) for the upstream neighbours in DESCENDING linear index; V end(State, own_aux)
//   V* out
// ---------------------------------------------------------------------------------------------------------
template <class Op>
__device__ __forceinline__ void ts_up_visit(TsSharedUp<typename Op::V>& s, const TsArgs& A, const Op& op, int tile, int pass) {
    typedef typename Op::V V;
    const long long r0 = (long long)(tile / A.ntx) * TS_T, c0 = (long long)(tile % A.ntx) * TS_T;
    if (threadIdx.x == 0) {
        s.act = 0;
        s.newly = 0;
    }
    ts_stage_graph(s.dir, s.flag, A, r0, c0, pass);
    // values: pass 1 starts from the cell's own datum; later passes read what earlier visits stored (final for done
    // cells, the own datum for pending ones)
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const long long r = r0 + ly, c = c0 + lx;
        if (r < A.nrow && c < A.ncol) {
            const long long g = r * A.ncol + c;
            s.val[ts_si(ly, lx)] = (pass == 1) ? op.init(g) : ld_cg(op.out + g);
            if (Op::AUX) s.aux[ts_si(ly, lx)] = op.aux(g);
        }
    }
    __syncthreads();  // flags of the halo are staged
    if (pass > 1 || Op::AUX) {
        for (int k = threadIdx.x; k < TS_NHALO; k += TS_THREADS) {
            int hy, hx;
            ts_halo_cell(k, hy, hx);
            const long long r = r0 + hy, c = c0 + hx;
            if (ts_in_raster(A, r, c)) {
                const long long g = r * A.ncol + c;
                if (s.flag[ts_si(hy, hx)] & TSF_DONE) s.val[ts_si(hy, hx)] = ld_cg(op.out + g);
                if (Op::AUX) s.aux[ts_si(hy, hx)] = op.aux(g);
            }
        }
    }
    // upstream / pending masks of the own cells
    uint32_t start = 0;
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const int c = ts_si(ly, lx);
        uint32_t ups = 0, pend = 0;
        const bool live = s.dir[c] != PFD_DIR_NODATA && !(s.flag[c] & TSF_DONE);
        if (live) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const int n = c + ts_noff(k);
                if (s.dir[n] == (uint8_t)(7 - k)) {
                    ups |= 1u << k;
                    if (!(s.flag[n] & TSF_DONE)) pend |= 1u << k;
                }
            }
            if (pend == 0) start |= 1u << j;
        }
        s.ups[c] = (uint8_t)ups;
        reinterpret_cast<uint8_t*>(s.pendw)[c] = (uint8_t)pend;
    }
    __syncthreads();
    // dataflow walk
    for (int j = 0; j < TS_CPT; ++j) {
        if (!((start >> j) & 1u)) continue;
        TS_OWN(j, ly, lx);
        int c = ts_si(ly, lx);
        for (;;) {
            uint32_t m = s.ups[c];
            const uint8_t own_aux = Op::AUX ? s.aux[c] : (uint8_t)0;
            typename Op::State st = op.begin(s.val[c], own_aux);
            while (m) {  // descending slot = descending linear index
                const int k = 31 - __clz(m);
                m ^= 1u << k;
                const int u = c + ts_noff(k);
                op.step(st, s.val[u], Op::AUX ? s.aux[u] : (uint8_t)0);
            }
            s.val[c] = op.end(st, own_aux);
            s.flag[c] = (uint8_t)(TSF_DONE | TSF_NEW);
            const uint32_t d = s.dir[c];
            if (d >= 8u) break;  // pit
            const int ds = c + ts_noff((int)d);
            const int hy = ds / TS_S - 1, hx = ds % TS_S - 1;
            if (hy < 0 || hy >= TS_T || hx < 0 || hx >= TS_T) {  // leaves the tile: the neighbour may continue next pass
                atomicOr(&s.act, ts_act_bit(hy, hx));
                break;
            }
            __threadfence_block();  // my value is visible before my bit disappears
            const unsigned sh = 8u * (ds & 3), bit = 1u << (7u - d);
            const uint32_t old = atomicAnd(&s.pendw[ds >> 2], ~(bit << sh));
            if (((old >> sh) & 0xFFu & ~bit) != 0u) break;  // somebody else arrives later and continues
            __threadfence_block();
            c = ds;
        }
    }
    __syncthreads();
    // store: pass 1 writes every cell of the tile (pending and nodata cells keep their own datum, like accu = data.copy()),
    // later passes only what this visit resolved
    uint32_t cnt = 0;
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const long long r = r0 + ly, c = c0 + lx;
        if (r < A.nrow && c < A.ncol) {
            const uint32_t f = s.flag[ts_si(ly, lx)];
            if (pass == 1 || (f & TSF_NEW)) op.out[r * A.ncol + c] = s.val[ts_si(ly, lx)];
            cnt += (f & TSF_NEW) ? 1u : 0u;
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s.newly, cnt);
    __threadfence();  // a tile running in the same pass may read my done bits: the values are out before the bits
    __syncthreads();
    ts_store_bitmap(s.flag, A, tile);
    __syncthreads();
    ts_activate(A, tile, s.act, pass);
    if (threadIdx.x == 0 && s.newly) atomicAdd(&A.ctl->resolved, (unsigned long long)s.newly);
    __syncthreads();  // shared memory is reused by the next visit
}

// All passes in one cooperative launch. grid-stride over the work list of the pass (pass 1: every tile).
template <class Op>
__global__ void __launch_bounds__(TS_THREADS) tile_up_sweep_kernel(TsArgs A, Op op) {
    extern __shared__ __align__(16) unsigned char ts_smem_raw[];
    TsSharedUp<typename Op::V>& s = *reinterpret_cast<TsSharedUp<typename Op::V>*>(ts_smem_raw);
    cg::grid_group grid = cg::this_grid();
    const unsigned int ntiles = (unsigned int)(A.ntx * A.nty);
    int pass = 1;
    for (;; ++pass) {
        const unsigned int count = (pass == 1) ? ntiles : __ldcg(&A.ctl->count[pass & 3]);
        if (count == 0) break;
        if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->count[(pass + 2) & 3] = 0u;  // last read two passes ago
        const uint32_t* list = A.list[pass & 1];
        for (unsigned int w = blockIdx.x; w < count; w += gridDim.x)
            ts_up_visit<Op>(s, A, op, (pass == 1) ? (int)w : (int)__ldcg(list + w), pass);
        __threadfence();
        grid.sync();
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->passes = (unsigned int)(pass - 1);
}

// streams.accuflux (up): accu = data.copy(); accu[ds] += accu[i] in seq[::-1] order, nodata-guarded on the running sum
template <typename T>
struct AccuUpTileOp {
    typedef T V;
    static const bool AUX = false;
    const T* data;
    T* out;
    NoData nd;
    __device__ __forceinline__ T init(long long g) const { return __ldg(data + g); }
    __device__ __forceinline__ uint8_t aux(long long) const { return 0; }
    typedef T State;
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
    __device__ __forceinline__ uint8_t init(long long) const { return 0; }
    __device__ __forceinline__ uint8_t aux(long long g) const { return __ldg(mask + g); }
    struct State {
        uint8_t so, smax;
    };
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
template <typename V>
struct TsSharedDown {
    V val[TS_N];
    uint16_t q[2][TS_T * TS_T];
    uint8_t dir[TS_N];
    uint8_t flag[TS_N];
    uint32_t qcnt[3];
    uint32_t act;
    uint32_t newly;
};

template <class Op>
__device__ __forceinline__ void ts_down_visit(TsSharedDown<typename Op::V>& s, const TsArgs& A, const Op& op, int tile, int pass) {
    typedef typename Op::V V;
    const long long r0 = (long long)(tile / A.ntx) * TS_T, c0 = (long long)(tile % A.ntx) * TS_T;
    if (threadIdx.x == 0) {
        s.act = 0;
        s.newly = 0;
        s.qcnt[0] = s.qcnt[1] = s.qcnt[2] = 0;
    }
    ts_stage_graph(s.dir, s.flag, A, r0, c0, pass);
    __syncthreads();
    // per-cell terms; values of the resolved halo cells
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const long long r = r0 + ly, c = c0 + lx;
        const int ci = ts_si(ly, lx);
        const uint32_t d = s.dir[ci];
        if (d != PFD_DIR_NODATA && !(s.flag[ci] & TSF_DONE)) {
            const long long g = r * A.ncol + c;
            s.val[ci] = op.prep(g, d, (d < 8u) ? g + pfd_slot_off((int)d, A.ncol) : g);
            if (op.source(g)) s.flag[ci] |= (uint8_t)TSF_SRC;
        }
    }
    if (pass > 1) {
        for (int k = threadIdx.x; k < TS_NHALO; k += TS_THREADS) {
            int hy, hx;
            ts_halo_cell(k, hy, hx);
            if (s.flag[ts_si(hy, hx)] & TSF_DONE) s.val[ts_si(hy, hx)] = ld_cg(op.out + ((r0 + hy) * A.ncol + c0 + hx));
        }
    }
    __syncthreads();
    // roots: sources, pits, exit cells whose downstream (halo) cell is resolved
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const int ci = ts_si(ly, lx);
        const uint32_t d = s.dir[ci], f = s.flag[ci];
        if (d == PFD_DIR_NODATA || (f & TSF_DONE)) continue;
        bool root = false;
        V v = s.val[ci];
        if (f & TSF_SRC) {
            v = op.source_value();
            root = true;
        } else if (d >= 8u) {
            v = op.pit(v);
            root = true;
        } else {
            const int ds = ci + ts_noff((int)d);
            const int hy = ds / TS_S - 1, hx = ds % TS_S - 1;
            if ((hy < 0 || hy >= TS_T || hx < 0 || hx >= TS_T) && (s.flag[ds] & TSF_DONE)) {
                v = op.down(s.val[ds], v);
                root = true;
            }
        }
        if (root) {
            s.val[ci] = v;
            s.flag[ci] = (uint8_t)(f | TSF_DONE | TSF_NEW);
            s.q[0][atomicAdd(&s.qcnt[0], 1u)] = (uint16_t)ci;
        }
    }
    // rounds: a thread follows the first child itself and queues the others
    for (int r = 0;; ++r) {
        __syncthreads();
        const uint32_t cnt = s.qcnt[r % 3];
        if (cnt == 0) break;
        if (threadIdx.x == 0) s.qcnt[(r + 2) % 3] = 0;
        const uint16_t* qin = s.q[r & 1];
        uint16_t* qout = s.q[(r + 1) & 1];
        for (uint32_t e = threadIdx.x; e < cnt; e += TS_THREADS) {
            int c = qin[e];
            for (;;) {
                const V vc = s.val[c];
                int first = -1;
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const int n = c + ts_noff(k);
                    if (s.dir[n] != (uint8_t)(7 - k)) continue;
                    const int hy = n / TS_S - 1, hx = n % TS_S - 1;
                    if (hy < 0 || hy >= TS_T || hx < 0 || hx >= TS_T) {  // the neighbour tile waits for this cell
                        if (!(s.flag[n] & TSF_DONE)) atomicOr(&s.act, ts_act_bit(hy, hx));
                        continue;
                    }
                    if (s.flag[n] & TSF_DONE) continue;  // a source, already resolved
                    s.val[n] = op.down(vc, s.val[n]);
                    s.flag[n] = (uint8_t)(TSF_DONE | TSF_NEW);
                    if (first < 0) first = n;
                    else qout[atomicAdd(&s.qcnt[(r + 1) % 3], 1u)] = (uint16_t)n;
                }
                if (first < 0) break;
                c = first;
            }
        }
    }
    // store
    uint32_t cnt = 0;
#pragma unroll 4
    for (int j = 0; j < TS_CPT; ++j) {
        TS_OWN(j, ly, lx);
        const long long r = r0 + ly, c = c0 + lx;
        if (r < A.nrow && c < A.ncol) {
            const uint32_t f = s.flag[ts_si(ly, lx)];
            if (f & TSF_NEW) op.out[r * A.ncol + c] = s.val[ts_si(ly, lx)];
            else if (pass == 1) op.out[r * A.ncol + c] = op.fill();
            cnt += (f & TSF_NEW) ? 1u : 0u;
        }
    }
    cnt = __reduce_add_sync(0xFFFFFFFFu, cnt);
    if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&s.newly, cnt);
    __threadfence();  // a tile running in the same pass may read my done bits: the values are out before the bits
    __syncthreads();
    ts_store_bitmap(s.flag, A, tile);
    __syncthreads();
    ts_activate(A, tile, s.act, pass);
    if (threadIdx.x == 0 && s.newly) atomicAdd(&A.ctl->resolved, (unsigned long long)s.newly);
    __syncthreads();
}

template <class Op>
__global__ void __launch_bounds__(TS_THREADS) tile_down_sweep_kernel(TsArgs A, Op op) {
    extern __shared__ __align__(16) unsigned char ts_smem_raw[];
    TsSharedDown<typename Op::V>& s = *reinterpret_cast<TsSharedDown<typename Op::V>*>(ts_smem_raw);
    cg::grid_group grid = cg::this_grid();
    const unsigned int ntiles = (unsigned int)(A.ntx * A.nty);
    int pass = 1;
    for (;; ++pass) {
        const unsigned int count = (pass == 1) ? ntiles : __ldcg(&A.ctl->count[pass & 3]);
        if (count == 0) break;
        if (blockIdx.x == 0 && threadIdx.x == 0) A.ctl->count[(pass + 2) & 3] = 0u;
        const uint32_t* list = A.list[pass & 1];
        for (unsigned int w = blockIdx.x; w < count; w += gridDim.x)
            ts_down_visit<Op>(s, A, op, (pass == 1) ? (int)w : (int)__ldcg(list + w), pass);
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
