// pfd_parse.cuh -- D8 raster -> device flow graph (dir + upmask) [+ idxs_ds], pit compaction, codecs.
// Replaces core_d8.from_array / check_values / to_array (pyflwdir/core_d8.py:42-67,115-122,86-102) and
// core.upstream_count / pit_indices (pyflwdir/core.py:50-61,225-232).
#pragma once
#include "pfd_common.cuh"

// ---------------------------------------------------------------------------------------------------------
// Parse kernel. One CTA = 256 threads handles a tile of PT_H rows x 128 columns. The tile plus a one-cell
// halo is staged in shared memory as 32-bit words (4 cells per word, coalesced 128 B row segments); each
// thread then processes 4 horizontally adjacent cells at once with byte-SIMD compares (__vcmpeq4):
//   for every neighbour slot k: which of my 4 cells flow there, is that neighbour nodata (forced pit),
//   does that neighbour flow into me (upstream mask bit k).
// Out-of-raster halo cells are filled with 247 (nodata), which makes "flows off the raster" and "flows into
// nodata" the same test (core_d8.py:58-61).
// ---------------------------------------------------------------------------------------------------------
#define PT_H 32
#define PT_WW 32               // words per tile row
#define PT_W (PT_WW * 4)       // 128 cells
#define PT_SW (PT_WW + 2)      // smem words per row incl. halo words

template <bool ALIGNED, int FT>
__device__ __forceinline__ uint32_t parse_load_word(const uint8_t* __restrict__ d8, int64_t nrow, int64_t ncol,
                                                    int64_t r, int64_t wc) {
    // word wc covers columns 4*wc .. 4*wc+3; cells outside the raster read as nodata
    constexpr uint32_t ND4 = pfd_nodata_code<FT>() * 0x01010101u;
    if (r < 0 || r >= nrow || wc < 0) return ND4;
    int64_t c = wc * 4;
    if (c >= ncol) return ND4;
    if (ALIGNED) {
        return __ldg(reinterpret_cast<const uint32_t*>(d8 + r * ncol + c));
    } else {
        uint32_t w = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            uint32_t v = (c + b < ncol) ? (uint32_t)__ldg(d8 + r * ncol + c + b) : pfd_nodata_code<FT>();
            w |= v << (8 * b);
        }
        return w;
    }
}

// IDXMODE: 0 = no idxs_ds output, 1 = 32-bit (int32 / uint32 share the bit pattern), 2 = int64
// FT: 0 = D8 codes (core_d8.py:14-19), 1 = PCRaster LDD codes (core_ldd.py:12-17; same algorithm, core_ldd.py:41-66)
template <bool ALIGNED, int IDXMODE, int FT>
__global__ void __launch_bounds__(256) parse_kernel(const uint8_t* __restrict__ d8, int64_t nrow, int64_t ncol,
                                                    uint8_t* __restrict__ dir, uint8_t* __restrict__ upmask,
                                                    void* __restrict__ idxs_out, unsigned int* __restrict__ invalid_flag,
                                                    int64_t out_row0, int64_t out_nrow, int64_t glob_row0) {
    __shared__ uint32_t tile[PT_H + 2][PT_SW];
    const int64_t r0 = (int64_t)blockIdx.y * PT_H;
    const int64_t wc0 = (int64_t)blockIdx.x * PT_WW;  // first word column of the tile

    for (int i = threadIdx.x; i < (PT_H + 2) * PT_SW; i += 256) {
        int tr = i / PT_SW, tw = i % PT_SW;
        tile[tr][tw] = parse_load_word<ALIGNED, FT>(d8, nrow, ncol, r0 - 1 + tr, wc0 - 1 + tw);
    }
    __syncthreads();

    const int tw = threadIdx.x & 31;  // word within the tile row
    const int wrow = threadIdx.x >> 5;
    bool bad = false;

#pragma unroll 1
    for (int it = 0; it < PT_H / 8; ++it) {
        const int tr = wrow + it * 8;
        const int64_t r = r0 + tr;
        const int64_t c = (wc0 + tw) * 4;
        if (r >= nrow || c >= ncol) continue;
        // 3x3 words around mine
        const uint32_t a0 = tile[tr][tw], a1 = tile[tr][tw + 1], a2 = tile[tr][tw + 2];
        const uint32_t b0 = tile[tr + 1][tw], w = tile[tr + 1][tw + 1], b2 = tile[tr + 1][tw + 2];
        const uint32_t c0 = tile[tr + 2][tw], c1 = tile[tr + 2][tw + 1], c2 = tile[tr + 2][tw + 2];
        uint32_t dirw, upw;
        if (pfd_parse_word<FT>(a0, a1, a2, b0, w, b2, c0, c1, c2, dirw, upw) != 0xFFFFFFFFu) bad = true;

        const int64_t i0 = r * ncol + c;
        if (ALIGNED) {
            *reinterpret_cast<uint32_t*>(dir + i0) = dirw;
            *reinterpret_cast<uint32_t*>(upmask + i0) = upw;
        } else {
#pragma unroll
            for (int b = 0; b < 4; ++b)
                if (c + b < ncol) {
                    dir[i0 + b] = (uint8_t)(dirw >> (8 * b));
                    upmask[i0 + b] = (uint8_t)(upw >> (8 * b));
                }
        }
        // idxs_ds only for the owned rows [out_row0, out_row0 + out_nrow) of a row block, as GLOBAL linear indices
        // (single GPU: the whole raster, offset 0)
        if (IDXMODE != 0 && r >= out_row0 && r < out_row0 + out_nrow) {
            const int64_t o0 = (r - out_row0) * ncol + c;       // position in the output block
            const int64_t g0 = (r - out_row0 + glob_row0) * ncol + c;  // global index of my first cell
            if (IDXMODE == 1) {
                // 32-bit outputs: wrap-around 32-bit arithmetic gives the right low word for int32 and uint32 alike
                const uint32_t g32 = (uint32_t)g0, nc32 = (uint32_t)ncol;
                uint32_t d32[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t d = (dirw >> (8 * b)) & 0xFFu;
                    const uint32_t off = (uint32_t)pfd_slot_dr((int)(d & 7u)) * nc32 + (uint32_t)pfd_slot_dc((int)(d & 7u));
                    d32[b] = (d < 8u) ? g32 + b + off : ((d == PFD_DIR_NODATA) ? 0xFFFFFFFFu : g32 + b);
                }
                uint32_t* o = reinterpret_cast<uint32_t*>(idxs_out) + o0;
                if (ALIGNED) {
                    *reinterpret_cast<uint4*>(o) = make_uint4(d32[0], d32[1], d32[2], d32[3]);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (c + b < ncol) o[b] = d32[b];
                }
            } else {
                int64_t ds[4];
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t d = (dirw >> (8 * b)) & 0xFFu;
                    const int64_t i = g0 + b;
                    ds[b] = (d < 8u) ? i + pfd_slot_off((int)d, ncol) : ((d == PFD_DIR_NODATA) ? (int64_t)-1 : i);
                }
                int64_t* o = reinterpret_cast<int64_t*>(idxs_out) + o0;
                if (ALIGNED) {
                    *reinterpret_cast<longlong2*>(o) = make_longlong2(ds[0], ds[1]);
                    *reinterpret_cast<longlong2*>(o + 2) = make_longlong2(ds[2], ds[3]);
                } else {
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if (c + b < ncol) o[b] = ds[b];
                }
            }
        }
    }
    if (bad) atomicOr(invalid_flag, 1u);
}

// ---------------------------------------------------------------------------------------------------------
// idxs_ds -> D8 code (core_d8.to_array semantics, core_d8.py:86-102) so that an index array given to the
// FlwdirRaster constructor goes through the same parse kernel. flag bit 1: link outside the 8 neighbours.
// ---------------------------------------------------------------------------------------------------------
template <typename IDX>
__global__ void idxs_to_d8_kernel(const IDX* __restrict__ idxs, int64_t n, int64_t ncol, uint8_t* __restrict__ d8,
                                  unsigned int* __restrict__ flag) {
    const IDX mv = (IDX)-1;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        IDX v = idxs[i];
        uint8_t code = 247;
        if (v != mv) {
            int64_t ds = (int64_t)v;
            if (ds < 0 || ds >= n) {
                atomicOr(flag, 2u);
            } else {
                int64_t dr = ds / ncol - i / ncol, dc = ds % ncol - i % ncol;
                if (dr < -1 || dr > 1 || dc < -1 || dc > 1) {
                    atomicOr(flag, 2u);
                } else if (dr == 0 && dc == 0) {
                    code = 0;
                } else {
                    int k = (int)((dr + 1) * 3 + (dc + 1));
                    k = k > 4 ? k - 1 : k;  // skip the centre
                    code = (uint8_t)pfd_slot_code(k);
                }
            }
        }
        d8[i] = code;
    }
}

// dir -> idxs_ds in the caller's dtype / D8 codes / int8 upstream count
template <typename IDX>
__global__ void dir_to_idxs_kernel(const uint8_t* __restrict__ dir, int64_t n, int64_t ncol, IDX* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t d = dir[i];
        int64_t ds = (d < 8u) ? i + pfd_slot_off((int)d, ncol) : ((d == PFD_DIR_NODATA) ? (int64_t)-1 : i);
        out[i] = (IDX)ds;
    }
}

// core_d8.to_array (core_d8.py:86-102) / core_ldd.to_array: pits become 0 (D8) / 5 (LDD)
template <int FT>
__global__ void dir_to_codes_kernel(const uint8_t* __restrict__ dir, int64_t n, uint8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        uint32_t d = dir[i];
        out[i] = (d < 8u) ? (uint8_t)pfd_code<FT>((int)d)
                          : ((d == PFD_DIR_NODATA) ? (uint8_t)pfd_nodata_code<FT>() : (uint8_t)(FT == 0 ? 0 : 5));
    }
}

__global__ void upstream_count_kernel(const uint8_t* __restrict__ dir, const uint8_t* __restrict__ upmask, int64_t n,
                                      int8_t* __restrict__ out) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        out[i] = (dir[i] == PFD_DIR_NODATA) ? (int8_t)-9 : (int8_t)__popc((unsigned)upmask[i]);
}

// ---------------------------------------------------------------------------------------------------------
// Pit compaction in ascending linear index (core_d8.py:48-62 appends pits in scan order).
// Pass 1 counts per 16 KiB chunk, a single-CTA scan turns counts into offsets, pass 2 writes.
// `dir` is padded with 255 up to a multiple of PC_CHUNK so all loads are 16-byte vectors.
// ---------------------------------------------------------------------------------------------------------
#define PC_CHUNK 16384  // cells per CTA: 8 warps x 4 iterations x 512 B

__device__ __forceinline__ uint32_t pit_mask16(const uint4& v, uint32_t* valid, uint32_t* outlets) {
    // returns a 16-bit mask of pit cells among the 16 bytes; accumulates valid / outlet counts
    uint32_t m = 0;
    const uint32_t ws[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        uint32_t p8 = __vcmpeq4(ws[j], splat4(PFD_DIR_PIT));
        uint32_t p9 = __vcmpeq4(ws[j], splat4(PFD_DIR_FPIT));
        uint32_t nd = __vcmpeq4(ws[j], splat4(PFD_DIR_NODATA));
        uint32_t p = p8 | p9;
        // compress 0xFF bytes to bits
        uint32_t bits = ((p & 0x00000080u) >> 7) | ((p & 0x00008000u) >> 14) | ((p & 0x00800000u) >> 21) |
                        ((p & 0x80000000u) >> 28);
        m |= bits << (4 * j);
        *valid += 4 - (__popc(nd) >> 3);
        *outlets += __popc(p8) >> 3;
    }
    return m;
}

__global__ void __launch_bounds__(256) pit_count_kernel(const uint8_t* __restrict__ dir, uint32_t* __restrict__ blk_pits,
                                                        unsigned long long* __restrict__ counters) {
    __shared__ uint32_t s_p[8], s_v[8], s_o[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint4* base = reinterpret_cast<const uint4*>(dir + (int64_t)blockIdx.x * PC_CHUNK + warp * 2048);
    uint32_t np = 0, nv = 0, no = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        uint4 v = __ldg(base + it * 32 + lane);
        np += __popc(pit_mask16(v, &nv, &no));
    }
    np = __reduce_add_sync(0xFFFFFFFFu, np);
    nv = __reduce_add_sync(0xFFFFFFFFu, nv);
    no = __reduce_add_sync(0xFFFFFFFFu, no);
    if (lane == 0) {
        s_p[warp] = np;
        s_v[warp] = nv;
        s_o[warp] = no;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t tp = 0, tv = 0, to = 0;
        for (int w = 0; w < 8; ++w) {
            tp += s_p[w];
            tv += s_v[w];
            to += s_o[w];
        }
        blk_pits[blockIdx.x] = tp;
        if (tv) atomicAdd(&counters[0], (unsigned long long)tv);
        if (tp) atomicAdd(&counters[1], (unsigned long long)tp);
        if (to) atomicAdd(&counters[2], (unsigned long long)to);
    }
}

// exclusive scan of `n` uint32 counts into uint64 offsets; single CTA of 1024 threads
__global__ void __launch_bounds__(1024) scan_counts_kernel(const uint32_t* __restrict__ counts, int64_t n,
                                                           unsigned long long* __restrict__ offsets) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t per = (n + 1023) / 1024;
    const int64_t b = threadIdx.x * per, e = min(n, b + per);
    unsigned long long sum = 0;
    for (int64_t i = b; i < e; ++i) sum += counts[i];
    unsigned long long incl = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = s_warp[lane], inc2 = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long t = __shfl_up_sync(0xFFFFFFFFu, inc2, d);
            if (lane >= d) inc2 += t;
        }
        s_warp[lane] = inc2 - v;
        if (lane == 31) s_carry = inc2;
    }
    __syncthreads();
    unsigned long long run = s_warp[warp] + (incl - sum);
    for (int64_t i = b; i < e; ++i) {
        offsets[i] = run;
        run += counts[i];
    }
    if (threadIdx.x == 0) offsets[n] = s_carry;
}

__global__ void __launch_bounds__(256) pit_scatter_kernel(const uint8_t* __restrict__ dir,
                                                          const unsigned long long* __restrict__ blk_off,
                                                          cell_t* __restrict__ pits, uint8_t* __restrict__ pit_outlet) {
    __shared__ uint32_t s_w[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t cell0 = (int64_t)blockIdx.x * PC_CHUNK + warp * 2048;
    const uint4* base = reinterpret_cast<const uint4*>(dir + cell0);
    uint4 v[4];
    uint32_t m[4];
    uint32_t cnt[4];
    uint32_t dummy_v = 0, dummy_o = 0, total = 0;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        v[it] = __ldg(base + it * 32 + lane);
        m[it] = pit_mask16(v[it], &dummy_v, &dummy_o);
        cnt[it] = __popc(m[it]);
        total += cnt[it];
    }
    // order inside the warp's 2048 cells: iteration-major, then lane, then byte
    uint32_t iter_tot[4], lane_excl[4];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        uint32_t incl = cnt[it];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += t;
        }
        lane_excl[it] = incl - cnt[it];
        iter_tot[it] = __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) s_w[warp] = iter_tot[0] + iter_tot[1] + iter_tot[2] + iter_tot[3];
    __syncthreads();
    unsigned long long basep = blk_off[blockIdx.x];
    for (int w = 0; w < warp; ++w) basep += s_w[w];
#pragma unroll
    for (int it = 0; it < 4; ++it) {
        unsigned long long o = basep + lane_excl[it];
        uint32_t mm = m[it];
        const uint32_t ws[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
        while (mm) {
            int b = __ffs(mm) - 1;
            mm &= mm - 1;
            uint32_t d = (ws[b >> 2] >> (8 * (b & 3))) & 0xFFu;
            pits[o] = (cell_t)(cell0 + it * 512 + lane * 16 + b);
            pit_outlet[o] = (d == PFD_DIR_PIT) ? 1 : 0;
            ++o;
        }
        basep += iter_tot[it];
    }
    (void)total;
}

// ---------------------------------------------------------------------------------------------------------
// core_nextxy.from_array / to_array (pyflwdir/core_nextxy.py:24-83): one-based (column, row) of the downstream cell in
// two int32 planes; -9 / -10 = pit (river mouth / inland), -9999 = nodata. Element-wise; the only gather is the
// "downstream cell is nodata" test. flag bit 1: core_nextxy.isvalid (:86-103) fails (only when check != 0);
// flag2 bit 2: a link that the 1-byte dir graph cannot hold.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool nextxy_ispit(int32_t v) { return v == -9 || v == -10; }

template <int IDXMODE>
__global__ void nextxy_parse_kernel(const int32_t* __restrict__ nextx, const int32_t* __restrict__ nexty, int64_t nrow, int64_t ncol,
                                    int check, uint8_t* __restrict__ dir, void* __restrict__ idxs, unsigned int* __restrict__ flag,
                                    unsigned int* __restrict__ flag2) {
    const int64_t n = nrow * ncol;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t x = nextx[i], y = nexty[i];
        if (check) {
            const bool m = x == -9999 || nextxy_ispit(x);
            if (m ? (x != y) : (x < 0)) atomicOr(flag, 1u);
        }
        uint32_t d;
        int64_t ds = -1;
        if (x == -9999) {
            d = PFD_DIR_NODATA;
        } else {
            const bool pit = nextxy_ispit(x) || nextxy_ispit(y);
            const int64_t r_ds = (int64_t)y - 1, c_ds = (int64_t)x - 1;
            const bool outside = r_ds >= nrow || c_ds >= ncol || r_ds < 0 || c_ds < 0;
            const int64_t ids = c_ds + r_ds * ncol;
            if (pit || outside || nextx[ids] == -9999) {
                d = nextxy_ispit(x) ? PFD_DIR_PIT : PFD_DIR_FPIT;  // outlet <=> nextx in (-9, -10) (pyflwdir.py:193)
                ds = i;
            } else {
                const int64_t r = i / ncol, c = i - r * ncol;
                const int64_t dr = r_ds - r, dc = c_ds - c;
                if (dr < -1 || dr > 1 || dc < -1 || dc > 1 || (dr == 0 && dc == 0)) {
                    atomicOr(flag2, 2u);
                    d = PFD_DIR_FPIT;
                    ds = i;
                } else {
                    const int k3 = (int)(dr + 1) * 3 + (int)(dc + 1);
                    d = (uint32_t)(k3 < 4 ? k3 : k3 - 1);
                    ds = ids;
                }
            }
        }
        dir[i] = (uint8_t)d;
        if (IDXMODE == 1) ((uint32_t*)idxs)[i] = (uint32_t)ds;
        else if (IDXMODE == 2) ((int64_t*)idxs)[i] = ds;
    }
}

__global__ void dir_to_nextxy_kernel(const uint8_t* __restrict__ dir, int64_t n, int64_t ncol, int32_t* __restrict__ nextx,
                                     int32_t* __restrict__ nexty) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const uint32_t d = dir[i];
        int32_t x = -9999, y = -9999;
        if (d < 8u) {
            const int64_t ds = i + pfd_slot_off((int)d, ncol);
            x = (int32_t)(ds % ncol + 1);
            y = (int32_t)(ds / ncol + 1);
        } else if (d != PFD_DIR_NODATA) {
            x = y = -9;  // core_nextxy._pv[0] for every pit
        }
        nextx[i] = x;
        nexty[i] = y;
    }
}
