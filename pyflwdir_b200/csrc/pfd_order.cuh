// pfd_order.cuh -- ordering of the flow graph: level-synchronous BFS from the pits.
// Replaces core.idxs_seq ("walk", pyflwdir/core.py:87-117 incl. upstream_matrix :67-84) and core.rank
// (pyflwdir/core.py:17-47). The reference's queue order is reproduced EXACTLY:
//   level 0   = pits in ascending linear index,
//   level k+1 = for every cell of level k, in order, its upstream neighbours in ascending linear index,
// i.e. an exclusive scan of the in-degrees (popcount of upmask) over the level gives every parent the output
// position of its children. The concatenated levels are `seq`; rank = level number; the default basin id is
// inherited from the parent (it travels in `bseq`, aligned with seq positions, so parents read it coalesced).
//
// One persistent cooperative kernel runs all levels: big levels are split into 4096-cell chunks that are
// scanned across CTAs with a decoupled look-back (status words tagged with the level, so no reset between
// levels) followed by one grid.sync(); runs of tiny levels are walked by CTA 0 alone with __syncthreads()
// only, so the long tail of the flow-path-length distribution costs ~one L2 round trip per level.
#pragma once
#include "pfd_common.cuh"

#ifndef BFS_THREADS
#define BFS_THREADS 1024
#endif
#ifndef BFS_ITEMS
#define BFS_ITEMS 4
#endif
#define BFS_CHUNK (BFS_THREADS * BFS_ITEMS)

struct BfsState {
    unsigned long long slot_start[2];  // [level & 1] first seq position of the level
    unsigned long long slot_end[2];    // [level & 1] one past the last position
    unsigned int cur_level;            // level the kernel (re)starts from
    unsigned int stop;                 // 0 running, 1 finished, 2 level_off capacity exhausted
    unsigned long long total;          // seq length when finished
};

struct BfsParams {
    const uint8_t* upmask;
    cell_t* seq;
    uint32_t* bseq;     // basin id per seq position (BASINS)
    int32_t* rank;      // (RANK)
    uint32_t* basins;   // (BASINS)
    long long* level_off;
    long long level_cap;  // level_off has level_cap + 1 entries
    unsigned long long* status;
    BfsState* st;
    long long ncol;
};

#define BFS_FLAG_AGG 1ull
#define BFS_FLAG_INC 2ull
#define BFS_VAL_BITS 34
#define BFS_VAL_MASK ((1ull << BFS_VAL_BITS) - 1)

__device__ __forceinline__ unsigned long long bfs_pack(unsigned int tag, unsigned long long flag, unsigned long long v) {
    return ((unsigned long long)(tag & 0x0FFFFFFFu) << 36) | (flag << BFS_VAL_BITS) | (v & BFS_VAL_MASK);
}

// Expand one chunk of the level [s, e): positions p0 .. p0 + BFS_CHUNK. `prefix_fn` supplies the exclusive
// prefix of the chunk inside the level (look-back in grid mode, 0 in solo mode). Returns the chunk total to
// every thread.
template <bool RANK, bool BASINS, bool LOOKBACK>
__device__ __forceinline__ unsigned int bfs_expand_chunk(const BfsParams& P, unsigned long long s, unsigned long long e,
                                                         unsigned int lev, long long chunk, long long nchunks,
                                                         unsigned int* s_warp, unsigned long long* s_prefix) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned long long p0 = s + (unsigned long long)chunk * BFS_CHUNK + (unsigned long long)threadIdx.x * BFS_ITEMS;

    cell_t cell[BFS_ITEMS];
    uint32_t mask[BFS_ITEMS];
    uint32_t cnt = 0;
#pragma unroll
    for (int j = 0; j < BFS_ITEMS; ++j) {
        const unsigned long long p = p0 + j;
        cell[j] = (p < e) ? __ldcg(P.seq + p) : 0u;
    }
#pragma unroll
    for (int j = 0; j < BFS_ITEMS; ++j) {
        const unsigned long long p = p0 + j;
        mask[j] = (p < e) ? (uint32_t)__ldg(P.upmask + cell[j]) : 0u;
        cnt += __popc(mask[j]);
    }
    // block-wide exclusive scan of cnt
    uint32_t incl = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    uint32_t warp_base = 0, total = 0;
#pragma unroll
    for (int w = 0; w < BFS_THREADS / 32; ++w) {
        uint32_t v = s_warp[w];
        if (w < warp) warp_base += v;
        total += v;
    }
    const uint32_t excl = warp_base + incl - cnt;

    if (LOOKBACK) {
        // decoupled look-back by warp 0
        if (warp == 0) {
            const unsigned int tag = lev + 1;
            volatile unsigned long long* status = P.status;
            unsigned long long prefix = 0;
            if (chunk == 0) {
                if (lane == 0) status[0] = bfs_pack(tag, BFS_FLAG_INC, total);
            } else {
                if (lane == 0) status[chunk] = bfs_pack(tag, BFS_FLAG_AGG, total);
                long long win = chunk - 1;  // highest predecessor of the current window
                for (;;) {
                    const long long idx = win - lane;
                    unsigned long long w;
                    bool ready;
                    do {
                        w = (idx >= 0) ? status[idx] : bfs_pack(tag, BFS_FLAG_INC, 0);
                        ready = ((unsigned int)(w >> 36) == (tag & 0x0FFFFFFFu)) && (((w >> BFS_VAL_BITS) & 3ull) != 0);
                    } while (!__all_sync(0xFFFFFFFFu, ready));
                    const bool inc = ((w >> BFS_VAL_BITS) & 3ull) == BFS_FLAG_INC;
                    const unsigned int inc_mask = __ballot_sync(0xFFFFFFFFu, inc);
                    const int first_inc = inc_mask ? (__ffs(inc_mask) - 1) : 32;  // nearest predecessor with inclusive
                    unsigned long long v = (lane <= first_inc) ? (w & BFS_VAL_MASK) : 0ull;
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xFFFFFFFFu, v, d);
                    prefix += v;
                    if (inc_mask) break;
                    win -= 32;
                }
                if (lane == 0) status[chunk] = bfs_pack(tag, BFS_FLAG_INC, prefix + total);
            }
            if (lane == 0) {
                *s_prefix = prefix;
                if (chunk == nchunks - 1) {
                    // publish the next level: nobody reads slot[(lev+1)&1] before the grid.sync()
                    P.st->slot_start[(lev + 1) & 1] = e;
                    P.st->slot_end[(lev + 1) & 1] = e + prefix + total;
                    P.level_off[lev + 1] = (long long)e;
                }
            }
        }
        __syncthreads();
    }
    const unsigned long long prefix = LOOKBACK ? *s_prefix : 0ull;
    unsigned long long o = e + prefix + excl;
#pragma unroll
    for (int j = 0; j < BFS_ITEMS; ++j) {
        uint32_t m = mask[j];
        if (m) {
            uint32_t b = 0;
            if (BASINS) b = __ldcg(P.bseq + p0 + j);
            const cell_t c = cell[j];
            while (m) {
                const int k = __ffs(m) - 1;
                m &= m - 1;
                const cell_t child = (cell_t)((long long)c + pfd_slot_off(k, P.ncol));
                P.seq[o] = child;
                if (BASINS) {
                    P.bseq[o] = b;
                    P.basins[child] = b;
                }
                if (RANK) P.rank[child] = (int32_t)(lev + 1);
                ++o;
            }
        }
    }
    return total;
}

template <bool RANK, bool BASINS>
__global__ void __launch_bounds__(BFS_THREADS) bfs_kernel(BfsParams P) {
    cg::grid_group grid = cg::this_grid();
    __shared__ unsigned int s_warp[BFS_THREADS / 32];
    __shared__ unsigned long long s_prefix;

    unsigned int lev = __ldcg(&P.st->cur_level);
    for (;;) {
        const unsigned long long s = __ldcg(&P.st->slot_start[lev & 1]);
        const unsigned long long e = __ldcg(&P.st->slot_end[lev & 1]);
        const unsigned long long size = e - s;
        if (size == 0 || (long long)lev + 1 > P.level_cap) {
            if (blockIdx.x == 0 && threadIdx.x == 0) {
                P.st->cur_level = lev;
                P.st->total = e;
                P.st->stop = (size == 0) ? 1u : 2u;
            }
            break;
        }
        if (size <= BFS_CHUNK) {
            grid.sync();  // every CTA has read the state of this level
            if (blockIdx.x == 0) {
                unsigned long long ss = s, ee = e;
                unsigned int l = lev;
                while (ee - ss > 0 && ee - ss <= BFS_CHUNK && (long long)l + 1 <= P.level_cap) {
                    const unsigned int total =
                        bfs_expand_chunk<RANK, BASINS, false>(P, ss, ee, l, 0, 1, s_warp, &s_prefix);
                    if (threadIdx.x == 0) P.level_off[l + 1] = (long long)ee;
                    __syncthreads();  // children visible to the CTA, s_warp reusable
                    ss = ee;
                    ee = ee + total;
                    ++l;
                }
                if (threadIdx.x == 0) {
                    P.st->slot_start[l & 1] = ss;
                    P.st->slot_end[l & 1] = ee;
                    P.st->cur_level = l;
                }
            }
            grid.sync();
            lev = __ldcg(&P.st->cur_level);
            continue;
        }
        const long long nchunks = (long long)((size + BFS_CHUNK - 1) / BFS_CHUNK);
        for (long long c = blockIdx.x; c < nchunks; c += gridDim.x) {
            bfs_expand_chunk<RANK, BASINS, true>(P, s, e, lev, c, nchunks, s_warp, &s_prefix);
            __syncthreads();
        }
        grid.sync();
        ++lev;
    }
}

// rank / basins initialisation: nodata -> -9999, valid -> -1 (cells the BFS never reaches keep -1 = "does not
// drain to a pit", core.py:35-38)
__global__ void order_init_rank_kernel(const uint8_t* __restrict__ dir, int64_t n, int32_t* __restrict__ rank) {
    const int64_t i4 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n) {
        const uint32_t w = *reinterpret_cast<const uint32_t*>(dir + i4);
        int4 r;
        r.x = ((w & 0xFFu) == 0xFFu) ? -9999 : -1;
        r.y = (((w >> 8) & 0xFFu) == 0xFFu) ? -9999 : -1;
        r.z = (((w >> 16) & 0xFFu) == 0xFFu) ? -9999 : -1;
        r.w = (((w >> 24) & 0xFFu) == 0xFFu) ? -9999 : -1;
        *reinterpret_cast<int4*>(rank + i4) = r;
    } else {
        for (int64_t i = i4; i < n; ++i) rank[i] = (dir[i] == PFD_DIR_NODATA) ? -9999 : -1;
    }
}

__global__ void order_init_pits_kernel(const cell_t* __restrict__ pits, int64_t npits, cell_t* __restrict__ seq,
                                       uint32_t* __restrict__ bseq, int32_t* __restrict__ rank,
                                       uint32_t* __restrict__ basins) {
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < npits; k += (int64_t)gridDim.x * blockDim.x) {
        const cell_t c = pits[k];
        seq[k] = c;
        if (bseq) bseq[k] = (uint32_t)(k + 1);
        if (basins) basins[c] = (uint32_t)(k + 1);
        if (rank) rank[c] = 0;
    }
}
