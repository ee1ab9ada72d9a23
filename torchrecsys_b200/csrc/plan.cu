// plan.cu -- the sort half of coalesce() (torch: optim/_functional.py:44) for a whole epoch:
// for every step and id space, a STABLE sort of that step's lookups by row id, so the training
// kernel can reduce duplicate rows deterministically (in lookup order) before the non-linear
// optimizer update.  Batched LSD radix sort, 8-bit digits, one segment per step.
//
// Lookups of a step with B_s samples:   user space: j in [0,B_s)      -> user[j]
//                                       item space: j in [0,2B_s)     -> j<B_s ? pos[j] : neg[j-B_s]
//                                       meta f    : j in [0,2B_s)     -> pos_meta[j,f] / neg_meta[j-B_s,f]
#include <stdlib.h>

#include "plan.cuh"

namespace trs {

constexpr int SORT_THREADS = 256;
constexpr int SORT_ROWS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ROWS;

// Lanes of the warp holding the same 8-bit digit as this lane (invalid lanes match nobody that is valid): one
// ballot per digit bit.  (__match_any_sync costs a round per DISTINCT value in the warp -- ~28 with random digits.)
__device__ __forceinline__ uint32_t match_digit(uint32_t dg, bool valid) {
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int bit = 0; bit < 8; ++bit) {
        const bool on = (dg >> bit) & 1u;
        const uint32_t b = __ballot_sync(0xffffffffu, on);
        peers &= on ? b : ~b;
    }
    return peers;
}

struct SortSrc {
    // first pass: ids
    const int64_t* a;
    const int64_t* b;
    int stride, off;
    // later passes: pairs
    const uint32_t* key;
    const uint32_t* val;
};

template <bool FIRST>
__device__ __forceinline__ void load_pair(const SortSrc& s, int64_t step, int B, int Bs, int mult,
                                          int j, uint32_t& key, uint32_t& val) {
    if (FIRST) {
        const int64_t sample0 = step * (int64_t)B;
        const int64_t id = (j < Bs) ? s.a[(sample0 + j) * s.stride + s.off]
                                    : s.b[(sample0 + j - Bs) * s.stride + s.off];
        key = (uint32_t)id;
        val = (uint32_t)j;
    } else {
        const int64_t base = (int64_t)mult * step * B;
        key = s.key[base + j];
        val = s.val[base + j];
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
sort_hist_kernel(SortSrc src, uint32_t* __restrict__ hist, int tiles, int64_t n_samples, int B,
                 int mult, int shift, const uint32_t* __restrict__ len_arr, int len_stride) {
    __shared__ uint32_t s_hist[256];
    const int64_t step = blockIdx.y;
    const int tile = blockIdx.x;
    const int Bs = (int)min((int64_t)B, n_samples - step * B);
    // len_arr (pairs only): the segment of step s holds len_arr[s * len_stride] pairs instead of mult * B_s
    const int len = len_arr ? (int)len_arr[step * len_stride] : mult * Bs;
    if (tile * SORT_TILE >= len) return;  // a tile beyond the segment: the scatter pass never reads its counts
    s_hist[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < SORT_ROWS; ++r) {
        int j = tile * SORT_TILE + r * SORT_THREADS + threadIdx.x;
        if (j < len) {
            uint32_t key, val;
            load_pair<FIRST>(src, step, B, Bs, mult, j, key, val);
            atomicAdd(&s_hist[(key >> shift) & 255u], 1u);
        }
    }
    __syncthreads();
    hist[((size_t)step * 256 + threadIdx.x) * tiles + tile] = s_hist[threadIdx.x];
}

template <bool FIRST>
__global__ void __launch_bounds__(SORT_THREADS)
sort_scatter_kernel(SortSrc src, uint32_t* __restrict__ dst_key, uint32_t* __restrict__ dst_val,
                    const uint32_t* __restrict__ hist, int tiles, int64_t n_samples, int B, int mult,
                    int shift, const uint32_t* __restrict__ len_arr, int len_stride) {
    __shared__ uint32_t s_base[256];
    __shared__ uint32_t s_run[256];
    __shared__ uint32_t s_wcnt[SORT_THREADS / 32][256];
    __shared__ uint32_t s_wtot[SORT_THREADS / 32];
    const int64_t step = blockIdx.y;
    const int tile = blockIdx.x;
    const int Bs = (int)min((int64_t)B, n_samples - step * B);
    const int len = len_arr ? (int)len_arr[step * len_stride] : mult * Bs;
    if (tile * SORT_TILE >= len) return;
    const int tiles_used = (len + SORT_TILE - 1) / SORT_TILE;  // the histogram pass filled only these
    const int64_t base = (int64_t)mult * step * B;
    const int d = threadIdx.x, lane = d & 31, warp = d >> 5;

    // global base of digit d for this tile: sum_{d'<d} total[d'] + sum_{t<tile} hist[d][t]
    {
        const uint32_t* h = hist + ((size_t)step * 256 + d) * tiles;
        uint32_t total = 0, before = 0;
        for (int t = 0; t < tiles_used; ++t) {
            uint32_t c = h[t];
            total += c;
            if (t < tile) before += c;
        }
        uint32_t inc = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += y;
        }
        if (lane == 31) s_wtot[warp] = inc;
        __syncthreads();
        uint32_t wbase = 0;
        for (int w = 0; w < warp; ++w) wbase += s_wtot[w];
        s_base[d] = wbase + inc - total + before;
        s_run[d] = 0;
    }
    __syncthreads();

    for (int r = 0; r < SORT_ROWS; ++r) {
        const int j = tile * SORT_TILE + r * SORT_THREADS + d;
        const bool valid = j < len;
        uint32_t key = 0, val = 0;
        if (valid) load_pair<FIRST>(src, step, B, Bs, mult, j, key, val);
        const uint32_t dg = valid ? ((key >> shift) & 255u) : 256u;
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) s_wcnt[w][d] = 0;
        __syncthreads();
        const uint32_t peers = match_digit(dg, valid);
        const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) s_wcnt[warp][dg] = __popc(peers);
        __syncthreads();
        {
            uint32_t run = s_run[d];
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) {
                uint32_t c = s_wcnt[w][d];
                s_wcnt[w][d] = run;
                run += c;
            }
            s_run[d] = run;
        }
        __syncthreads();
        if (valid) {
            const uint32_t p = s_base[dg] + s_wcnt[warp][dg] + rank;
            dst_key[base + p] = key;
            dst_val[base + p] = val;
        }
        __syncthreads();
    }
}

static int bits_for(int64_t n_rows) {
    int b = 1;
    while (((int64_t)1 << b) < n_rows) ++b;
    return b;
}

// Sort one id space for all steps.  Final (key, perm) land in out_key/out_val.
int sort_space(const int64_t* a, const int64_t* b, int stride, int off, int mult,
                      int64_t n_rows, const trs_epoch* ep, uint32_t* out_key, uint32_t* out_val,
                      uint32_t* tmp_key, uint32_t* tmp_val, uint32_t* hist, cudaStream_t st) {
    const int64_t steps = n_steps_of(ep);
    const int tiles = (int)(((int64_t)mult * ep->batch + SORT_TILE - 1) / SORT_TILE);
    const int npass = (bits_for(n_rows) + 7) / 8;
    dim3 grid(tiles, (unsigned)steps);
    for (int p = 0; p < npass; ++p) {
        const bool to_out = ((npass - 1 - p) % 2) == 0;
        uint32_t* dk = to_out ? out_key : tmp_key;
        uint32_t* dv = to_out ? out_val : tmp_val;
        SortSrc src = {a, b, stride, off, to_out ? tmp_key : out_key, to_out ? tmp_val : out_val};
        if (p == 0) {
            sort_hist_kernel<true><<<grid, SORT_THREADS, 0, st>>>(src, hist, tiles, ep->n_samples, ep->batch, mult, 0, nullptr, 0);
            sort_scatter_kernel<true><<<grid, SORT_THREADS, 0, st>>>(src, dk, dv, hist, tiles, ep->n_samples, ep->batch, mult, 0, nullptr, 0);
        } else {
            sort_hist_kernel<false><<<grid, SORT_THREADS, 0, st>>>(src, hist, tiles, ep->n_samples, ep->batch, mult, 8 * p, nullptr, 0);
            sort_scatter_kernel<false><<<grid, SORT_THREADS, 0, st>>>(src, dk, dv, hist, tiles, ep->n_samples, ep->batch, mult, 8 * p, nullptr, 0);
        }
    }
    TRS_CUDA(cudaGetLastError());
    return TRS_OK;
}

// Stable sort of (key, value) PAIRS that are already in memory: the segment of step s starts at mult * s * B and
// holds len_arr[s * len_stride] pairs (the row-sharded plan: only the lookups a rank owns).  Pass p reads
// buf[p & 1] and writes buf[(p + 1) & 1]; returns the number of passes so the caller knows where the result is.
int sort_pairs(uint32_t* key0, uint32_t* val0, uint32_t* key1, uint32_t* val1, int mult, int64_t n_rows,
               const trs_epoch* ep, const uint32_t* len_arr, int len_stride, uint32_t* hist, cudaStream_t st) {
    const int64_t steps = n_steps_of(ep);
    const int tiles = (int)(((int64_t)mult * ep->batch + SORT_TILE - 1) / SORT_TILE);
    const int npass = (bits_for(n_rows) + 7) / 8;
    dim3 grid(tiles, (unsigned)steps);
    for (int p = 0; p < npass; ++p) {
        SortSrc src = {nullptr, nullptr, 0, 0, (p & 1) ? key1 : key0, (p & 1) ? val1 : val0};
        uint32_t* dk = (p & 1) ? key0 : key1;
        uint32_t* dv = (p & 1) ? val0 : val1;
        sort_hist_kernel<false><<<grid, SORT_THREADS, 0, st>>>(src, hist, tiles, ep->n_samples, ep->batch, mult, 8 * p,
                                                              len_arr, len_stride);
        sort_scatter_kernel<false><<<grid, SORT_THREADS, 0, st>>>(src, dk, dv, hist, tiles, ep->n_samples, ep->batch, mult,
                                                                 8 * p, len_arr, len_stride);
    }
    return npass;
}
int sort_passes(int64_t n_rows) { return (bits_for(n_rows) + 7) / 8; }

// After the sort: turn every segment (run of equal row ids) into work items.  Sorted keys make
// "longer than T" one probe (K[k+T] == K[k]) and the exact length a binary search.
__global__ void __launch_bounds__(256)
build_items_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ perm, int mult,
                   int space, int64_t n_samples, int B, uint32_t* __restrict__ item_cnt,
                   uint4* __restrict__ items, int item_cap, uint32_t* __restrict__ long_cnt,
                   uint4* __restrict__ long_segs, int long_cap, uint8_t* __restrict__ single) {
    const int64_t step = blockIdx.y;
    const int Bs = (int)min((int64_t)B, n_samples - step * B);
    const int len = mult * Bs;
    const uint32_t* K = keys + (int64_t)mult * step * B;
    const uint32_t* P = perm + (int64_t)mult * step * B;
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t key = 0;
    bool head = false, is_long = false;
    int c = 0;
    if (k < len) {
        key = K[k];
        head = (k == 0) || (K[k - 1] != key);
        if (head) {
            is_long = (k + LONG_SEG_T < len) && (K[k + LONG_SEG_T] == key);
            if (!is_long) {
                c = 1;
                while (k + c < len && K[k + c] == key) ++c;
            }
        }
    }
    uint4* out = items + (size_t)step * item_cap;
    // a row looked up exactly once in this step: the sample's own row group updates it in phase A
    // (`single` is null for id spaces / nets whose rows are always reduced in phase B)
    if (head && !is_long && c == 1 && single) {
        single[(int64_t)mult * step * B + P[k]] = 1;
        head = false;
    }
    // short segments: one warp-aggregated append
    const bool is_short = head && !is_long;
    const uint32_t m = __ballot_sync(0xffffffffu, is_short);
    if (m) {
        uint32_t base = 0;
        const int leader = __ffs(m) - 1;
        if (lane == leader) base = atomicAdd(&item_cnt[step], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (is_short) {
            const uint32_t slot = base + __popc(m & ((1u << lane) - 1u));
            out[slot] = make_uint4((uint32_t)space | ((uint32_t)c << 8), (uint32_t)k, key, P[k]);
        }
    }
    if (is_long) {
        int lo = k + LONG_SEG_T, hi = len;  // K[lo] == key, K[hi] != key (hi == len: sentinel)
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (K[mid] == key) lo = mid; else hi = mid;
        }
        const int cl = hi - k;
        const uint32_t seg = atomicAdd(&long_cnt[step], 1u);
        if (seg < (uint32_t)long_cap)
            long_segs[(size_t)step * long_cap + seg] = make_uint4((uint32_t)space, (uint32_t)k, (uint32_t)cl, key);
    }
}

// Bit 1 of a lookup's flag byte: its row is looked up -- hence updated -- in the PREVIOUS step, so a copy
// fetched while that step runs may be stale: the training kernel then reads the row after the step barrier
// instead of taking it from its early shared-memory prefetch.
__global__ void __launch_bounds__(256)
dirty_flags_kernel(const uint32_t* __restrict__ keys, const int64_t* __restrict__ a, const int64_t* __restrict__ b,
                   int mult, int64_t n_samples, int B, uint8_t* __restrict__ flags) {
    const int64_t step = (int64_t)blockIdx.y + 1;  // lookups of step s+1 against the sorted rows of step s
    const int Bs = (int)min((int64_t)B, n_samples - step * B);
    const int len = mult * Bs;
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    const int64_t sample0 = step * (int64_t)B;
    const uint32_t row = (uint32_t)((j < Bs) ? a[sample0 + j] : b[sample0 + j - Bs]);
    const uint32_t* K = keys + (int64_t)mult * (step - 1) * B;
    const int plen = mult * B;  // every step but the last is full
    int lo = 0, hi = plen;      // lower bound of row in K
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (K[mid] < row) lo = mid + 1; else hi = mid;
    }
    if (lo < plen && K[lo] == row) flags[(int64_t)mult * step * B + j] |= 2;
}

// ---- single-CTA plan of one (step, id space) ---------------------------------------------------
// A step's lookups of one id space (<= FS_MAX of them) fit in shared memory, so one CTA does the whole job
// there: load the ids, every radix pass (ping-pong between two shared buffers), the copy-out of the sorted
// (row, lookup) pairs and the work items -- one launch for the whole epoch instead of two per radix pass and
// id space, and the lookups cross HBM twice (ids in, plan out).  Larger batches take the tiled path above.
constexpr int FS_THREADS = 1024;
constexpr int FS_KPT = 16;                     // lookups per thread
constexpr int FS_MAX = FS_THREADS * FS_KPT;    // lookups per (step, id space) held in shared memory
constexpr size_t FS_SMEM = (size_t)FS_MAX * (2 * 4 + 2 * 2) + 32 * 256 * 2;

// flag bytes are set by two CTAs (the step's own: bit 0, the previous step's: bit 1): word-wide atomic OR
__device__ __forceinline__ void flag_or(uint8_t* flags, int64_t idx, uint32_t bits) {
    atomicOr(reinterpret_cast<unsigned*>(flags) + (idx >> 2), bits << (8 * (int)(idx & 3)));
}

__device__ __forceinline__ uint32_t row_hash(uint32_t row) { return (row * 2654435761u) >> 13; }  // 19 bits

struct FusedSpace {
    const int64_t* a;
    const int64_t* b;
    int stride, off, mult, npass, space;
    uint32_t* out_key;
    uint32_t* out_val;
    uint8_t* single;
};
struct FusedArgs {
    FusedSpace sp[2 + TRS_MAX_META];
};

__global__ void __launch_bounds__(FS_THREADS, 1)
plan_fused_kernel(const __grid_constant__ FusedArgs A, int64_t n_samples, int B, uint32_t* __restrict__ item_cnt,
                  uint4* __restrict__ items, int item_cap, uint32_t* __restrict__ long_cnt,
                  uint4* __restrict__ long_segs, int long_cap) {
    extern __shared__ __align__(16) unsigned char fs_smem[];
    __shared__ uint32_t s_dbase[256];
    __shared__ uint32_t s_wtot[8];
    __shared__ uint32_t s_q[4][256];
    __shared__ uint32_t s_cnt[2], s_gbase[2];
    uint32_t* ks = reinterpret_cast<uint32_t*>(fs_smem);            // keys: source / destination of a pass
    uint32_t* kd = ks + FS_MAX;
    uint16_t* vs = reinterpret_cast<uint16_t*>(kd + FS_MAX);        // lookup ids (< FS_MAX: 16 bits)
    uint16_t* vd = vs + FS_MAX;
    uint16_t* whist = vd + FS_MAX;                                  // [warp][digit] counts, then offsets

    const FusedSpace& S = A.sp[blockIdx.x];
    const int64_t step = blockIdx.y;
    const int Bs = (int)min((int64_t)B, n_samples - step * B);
    const int len = S.mult * Bs;
    const int64_t base = (int64_t)S.mult * step * B;
    const int64_t sample0 = step * (int64_t)B;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t lt = (1u << lane) - 1u;

    // ids are non-negative and < 2^32 (trs_validate_ids / check above): the low word is the row number.
    // All of a thread's loads are issued before the first is used.
    auto id_ptr = [&](int64_t smp0, int bs, int j) {
        const int64_t* q = (j < bs) ? S.a + ((smp0 + j) * S.stride + S.off) : S.b + ((smp0 + j - bs) * S.stride + S.off);
        return reinterpret_cast<const uint32_t*>(q);
    };
    {
        uint32_t id[FS_KPT];
#pragma unroll
        for (int it = 0; it < FS_KPT; ++it) {
            const int j = it * FS_THREADS + tid;
            id[it] = j < len ? __ldg(id_ptr(sample0, Bs, j)) : 0u;
        }
#pragma unroll
        for (int it = 0; it < FS_KPT; ++it) {
            const int j = it * FS_THREADS + tid;
            if (j < len) {
                ks[j] = id[it];
                vs[j] = (uint16_t)j;
            }
        }
    }
    if (tid == 0) s_cnt[0] = s_cnt[1] = 0;
    // every warp ranks a contiguous run of lookups, 32 at a time (stable: runs and rounds are in lookup order)
    const int chunk = (((len + 31) / 32) + 31) & ~31;
    const int rounds = chunk / 32;  // <= FS_KPT
    const int wstart = warp * chunk;
    for (int p = 0; p < S.npass; ++p) {
        const int shift = 8 * p;
        for (int i = tid; i < 32 * 256 / 2; i += FS_THREADS) reinterpret_cast<uint32_t*>(whist)[i] = 0u;
        __syncthreads();  // ids loaded / previous pass scattered, counters cleared
        uint32_t lr[FS_KPT / 2];  // rank of my lookup of round r among its warp's equal digits, 16 bits each
#pragma unroll
        for (int r = 0; r < FS_KPT; ++r) {
            uint32_t v = 0;
            if (r < rounds) {
                const int j = wstart + r * 32 + lane;
                const bool valid = j < len;
                const uint32_t dg = valid ? ((ks[j] >> shift) & 255u) : 256u;
                const uint32_t peers = match_digit(dg, valid);
                const uint32_t rank = __popc(peers & lt);
                const uint32_t off = valid ? whist[warp * 256 + dg] : 0u;
                __syncwarp();
                if (valid && rank == 0) whist[warp * 256 + dg] = (uint16_t)(off + __popc(peers));
                __syncwarp();
                v = off + rank;
            }
            if (r & 1) lr[r / 2] |= v << 16; else lr[r / 2] = v;
        }
        __syncthreads();
        // digit-major exclusive scan over (digit, warp): thread (q, d) takes warps 8q..8q+7 of digit d
        uint32_t total = 0, inc = 0;
        {
            const int d = tid & 255, q = tid >> 8;
            uint32_t run = 0;
#pragma unroll
            for (int w = 8 * q; w < 8 * q + 8; ++w) {
                const uint32_t c = whist[w * 256 + d];
                whist[w * 256 + d] = (uint16_t)run;
                run += c;
            }
            s_q[q][d] = run;
            __syncthreads();
            uint32_t before = 0;
#pragma unroll
            for (int qq = 0; qq < 4; ++qq) {
                const uint32_t t = s_q[qq][d];
                total += t;
                if (qq < q) before += t;
            }
            if (before) {
#pragma unroll
                for (int w = 8 * q; w < 8 * q + 8; ++w) whist[w * 256 + d] += (uint16_t)before;
            }
        }
        if (tid < 256) {
            inc = total;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += y;
            }
            if (lane == 31) s_wtot[warp] = inc;
        }
        __syncthreads();
        if (tid < 256) {
            uint32_t wbase = 0;
            for (int w = 0; w < warp; ++w) wbase += s_wtot[w];
            s_dbase[tid] = wbase + inc - total;
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < FS_KPT; ++r) {
            if (r < rounds) {
                const int j = wstart + r * 32 + lane;
                if (j < len) {
                    const uint32_t key = ks[j];
                    const uint32_t dg = (key >> shift) & 255u;
                    const uint32_t pos = s_dbase[dg] + whist[warp * 256 + dg] + ((lr[r / 2] >> (16 * (r & 1))) & 0xffffu);
                    kd[pos] = key;
                    vd[pos] = vs[j];
                }
            }
        }
        __syncthreads();
        uint32_t* tk = ks; ks = kd; kd = tk;
        uint16_t* tv = vs; vs = vd; vd = tv;
    }
    // the key buffer the last pass read from is free now: a FS_MAX * 32-bit hash bitmap of the rows looked up
    // in this step (for the next step's flag bit 1, below)
    uint32_t* bitmap = kd;
    const bool want_dirty = S.single && (step + 1) * (int64_t)B < n_samples;
    if (want_dirty) {
#pragma unroll
        for (int it = 0; it < FS_KPT; ++it) bitmap[it * FS_THREADS + tid] = 0u;
    }
    __syncthreads();

    for (int k = tid; k < len; k += FS_THREADS) {
        S.out_key[base + k] = ks[k];
        S.out_val[base + k] = vs[k];
    }
    // work items (see build_items_kernel): slots come from shared counters, one global reservation per CTA
    constexpr uint32_t SHORT = 0x40000000u, LONG = 0x80000000u;
    uint32_t aux[FS_KPT];
#pragma unroll
    for (int it = 0; it < FS_KPT; ++it) {
        const int k = it * FS_THREADS + tid;
        uint32_t a = 0;
        if (k < len) {
            const uint32_t key = ks[k];
            if (k == 0 || ks[k - 1] != key) {
                int c = LONG_SEG_T + 1;
                if (k + LONG_SEG_T < len && ks[k + LONG_SEG_T] == key) {
                    a = LONG;
                } else {
                    c = 1;
                    while (k + c < len && ks[k + c] == key) ++c;
                    if (c == 1 && S.single) flag_or(S.single, base + vs[k], 1u);
                    else a = SHORT | ((uint32_t)c << 16);
                }
                if (want_dirty) {
                    const uint32_t h = row_hash(key);
                    atomicOr(&bitmap[h >> 5], 1u << (h & 31u));
                }
            }
        }
        const uint32_t ms = __ballot_sync(0xffffffffu, (a & SHORT) != 0u);
        const uint32_t ml = __ballot_sync(0xffffffffu, (a & LONG) != 0u);
        if (ms) {
            const int leader = __ffs(ms) - 1;
            uint32_t b0 = 0;
            if (lane == leader) b0 = atomicAdd(&s_cnt[0], (uint32_t)__popc(ms));
            b0 = __shfl_sync(0xffffffffu, b0, leader);
            if (a & SHORT) a |= b0 + __popc(ms & lt);
        }
        if (ml) {
            const int leader = __ffs(ml) - 1;
            uint32_t b0 = 0;
            if (lane == leader) b0 = atomicAdd(&s_cnt[1], (uint32_t)__popc(ml));
            b0 = __shfl_sync(0xffffffffu, b0, leader);
            if (a & LONG) a |= b0 + __popc(ml & lt);
        }
        aux[it] = a;
    }
    __syncthreads();
    if (tid == 0) {
        s_gbase[0] = s_cnt[0] ? atomicAdd(&item_cnt[step], s_cnt[0]) : 0u;
        s_gbase[1] = s_cnt[1] ? atomicAdd(&long_cnt[step], s_cnt[1]) : 0u;
    }
    __syncthreads();
    // flag bit 1 of the NEXT step's lookups (see dirty_flags_kernel): is the row looked up in this step?  A probe
    // of the bitmap, confirmed by a binary search in the sorted keys held in shared memory.
    if (want_dirty) {
        const int64_t sample1 = sample0 + B;
        const int Bs1 = (int)min((int64_t)B, n_samples - sample1);
        const int len1 = S.mult * Bs1;
        uint32_t row[FS_KPT];
#pragma unroll
        for (int it = 0; it < FS_KPT; ++it) {
            const int j = it * FS_THREADS + tid;
            row[it] = j < len1 ? __ldg(id_ptr(sample1, Bs1, j)) : 0u;
        }
#pragma unroll
        for (int it = 0; it < FS_KPT; ++it) {
            const int j = it * FS_THREADS + tid;
            const uint32_t h = row_hash(row[it]);
            if (j < len1 && ((bitmap[h >> 5] >> (h & 31u)) & 1u)) {  // rarely true: confirm by binary search
                int lo = 0, hi = len;  // lower bound of row[it] in ks
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    if (ks[mid] < row[it]) lo = mid + 1; else hi = mid;
                }
                if (lo < len && ks[lo] == row[it]) flag_or(S.single, base + (int64_t)S.mult * B + j, 2u);
            }
        }
    }
    uint4* out = items + (size_t)step * item_cap;
#pragma unroll
    for (int it = 0; it < FS_KPT; ++it) {
        const int k = it * FS_THREADS + tid;
        const uint32_t a = aux[it];
        if (a & SHORT) {
            const uint32_t c = (a >> 16) & 0xffu;
            out[s_gbase[0] + (a & 0xffffu)] = make_uint4((uint32_t)S.space | (c << 8), (uint32_t)k, ks[k], (uint32_t)vs[k]);
        } else if (a & LONG) {
            const uint32_t key = ks[k];
            int lo = k + LONG_SEG_T, hi = len;  // ks[lo] == key, ks[hi] != key (hi == len: sentinel)
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (ks[mid] == key) lo = mid; else hi = mid;
            }
            const uint32_t seg = s_gbase[1] + (a & 0xffffu);
            if (seg < (uint32_t)long_cap)
                long_segs[(size_t)step * long_cap + seg] = make_uint4((uint32_t)S.space, (uint32_t)k, (uint32_t)(hi - k), key);
        }
    }
}

PlanLayout plan_layout(int64_t n_samples, int batch, int n_meta) {
    PlanLayout L;
    size_t off = 0;
    auto take = [&](size_t n_elems) {
        size_t o = off;
        off += (n_elems * sizeof(uint32_t) + 255) / 256 * 256;
        return o;
    };
    L.user_key = take(n_samples);
    L.user_perm = take(n_samples);
    L.item_key = take(2 * n_samples);
    L.item_perm = take(2 * n_samples);
    for (int f = 0; f < TRS_MAX_META; ++f) {
        L.meta_key[f] = f < n_meta ? take(2 * n_samples) : 0;
        L.meta_perm[f] = f < n_meta ? take(2 * n_samples) : 0;
    }
    const int64_t steps = (n_samples + batch - 1) / batch;
    const int64_t lookups = (int64_t)batch * (3 + 2 * n_meta);
    L.item_cap = (int)lookups;                            // every item covers >= 1 lookup
    L.long_cap = (int)(lookups / (LONG_SEG_T + 1) + 1);   // a long segment has > T lookups
    L.item_cnt = take((size_t)steps);
    L.long_cnt = take((size_t)steps);
    L.single_user = take((size_t)(n_samples + 3) / 4);        // one byte per lookup
    L.single_item = take((size_t)(2 * n_samples + 3) / 4);
    L.items = take((size_t)steps * L.item_cap * 4);
    L.long_segs = take((size_t)steps * L.long_cap * 4);
    L.total = off;
    return L;
}

size_t hist_bytes(const trs_epoch* ep) {
    const int tiles = (int)((2ll * ep->batch + SORT_TILE - 1) / SORT_TILE);
    return (size_t)n_steps_of(ep) * 256 * tiles * sizeof(uint32_t);
}

}  // namespace trs

using namespace trs;

extern "C" size_t trs_plan_bytes(const trs_model* model, const trs_epoch* epoch) {
    if (!model || !epoch) return 0;
    if (epoch->batch <= 0) return 0;
    return plan_layout(epoch->n_samples, epoch->batch, model->n_meta).total;
}

extern "C" size_t trs_plan_tmp_bytes(const trs_model* model, const trs_epoch* epoch) {
    if (!model || !epoch || epoch->batch <= 0) return 0;
    // one (key, perm) pair buffer of the widest id space + the per-tile digit histograms
    return ((size_t)4 * epoch->n_samples * sizeof(uint32_t) + 511) / 256 * 256 + hist_bytes(epoch);
}

extern "C" int trs_plan_build(const trs_model* model, const trs_epoch* ep, void* plan,
                              size_t plan_bytes, void* tmp, size_t tmp_bytes, trs_stream_t stream) {
    RowShape shape;
    int rc = check_model(model, &shape);
    if (rc) return rc;
    TRS_REQUIRE(ep && ep->user && ep->pos && ep->neg, "epoch ids are NULL");
    TRS_REQUIRE(ep->batch > 0 && ep->batch <= (1 << 30), "batch out of range");
    TRS_REQUIRE(model->n_meta == 0 || (ep->pos_meta && ep->neg_meta), "metadata ids are NULL");
    TRS_REQUIRE(plan && tmp, "plan/tmp is NULL");
    TRS_REQUIRE(model->user.n_rows > 0 && model->user.n_rows <= 0xFFFFFFFFll &&
                model->item.n_rows > 0 && model->item.n_rows <= 0xFFFFFFFFll, "n_rows out of range");
    if (ep->n_samples == 0) return TRS_OK;
    TRS_REQUIRE(n_steps_of(ep) <= 65535, "plan_build: %lld steps in one call (limit 65535): pass the epoch in runs of whole steps",
                (long long)n_steps_of(ep));
    const PlanLayout L = plan_layout(ep->n_samples, ep->batch, model->n_meta);
    if (plan_bytes < L.total || tmp_bytes < trs_plan_tmp_bytes(model, ep)) {
        set_error("plan workspace too small: plan %zu < %zu or tmp %zu < %zu", plan_bytes, L.total,
                  tmp_bytes, trs_plan_tmp_bytes(model, ep));
        return TRS_ERR_WORKSPACE;
    }
    char* P = (char*)plan;
    uint32_t* tmp_key = (uint32_t*)tmp;
    uint32_t* tmp_val = tmp_key + 2 * ep->n_samples;
    uint32_t* hist = (uint32_t*)((char*)tmp + ((size_t)4 * ep->n_samples * sizeof(uint32_t) + 511) / 256 * 256);
    for (int f = 0; f < model->n_meta; ++f)
        TRS_REQUIRE(model->meta[f].n_rows > 0 && model->meta[f].n_rows <= 0xFFFFFFFFll, "meta n_rows out of range");
    const int64_t steps = n_steps_of(ep);
    uint32_t* item_cnt = (uint32_t*)(P + L.item_cnt);
    uint32_t* long_cnt = (uint32_t*)(P + L.long_cnt);
    // counters and singleton flags are adjacent (take() order): one memset
    TRS_CUDA(cudaMemsetAsync(item_cnt, 0, L.items - L.item_cnt, stream));
    // the MLP tower stages every gradient row itself (mlp.cu): nothing is updated in a phase A there
    const bool fuse_single = model->net != TRS_NET_MLP;
    auto single_of = [&](int space) -> uint8_t* {
        return !fuse_single ? nullptr
               : space == 0 ? (uint8_t*)(P + L.single_user)
               : space == 1 ? (uint8_t*)(P + L.single_item) : nullptr;
    };
    // TRS_PLAN_TILED=1 (tests): take the tiled multi-launch path at every batch size
    const char* tiled_env = getenv("TRS_PLAN_TILED");
    const bool fused = 2ll * ep->batch <= FS_MAX && !(tiled_env && atoi(tiled_env));
    if (fused) {
        FusedArgs A = {};
        auto space_of = [&](const int64_t* a, const int64_t* b, int stride, int off, int mult, int64_t n_rows,
                            int space, size_t key_off, size_t perm_off) {
            FusedSpace s = {a, b, stride, off, mult, (bits_for(n_rows) + 7) / 8, space,
                            (uint32_t*)(P + key_off), (uint32_t*)(P + perm_off), single_of(space)};
            return s;
        };
        A.sp[0] = space_of(ep->user, ep->user, 1, 0, 1, model->user.n_rows, 0, L.user_key, L.user_perm);
        A.sp[1] = space_of(ep->pos, ep->neg, 1, 0, 2, model->item.n_rows, 1, L.item_key, L.item_perm);
        for (int f = 0; f < model->n_meta; ++f)
            A.sp[2 + f] = space_of(ep->pos_meta, ep->neg_meta, model->n_meta, f, 2, model->meta[f].n_rows, 2 + f,
                                   L.meta_key[f], L.meta_perm[f]);
        TRS_CUDA(cudaFuncSetAttribute(plan_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FS_SMEM));
        dim3 grid((unsigned)(2 + model->n_meta), (unsigned)steps);
        plan_fused_kernel<<<grid, FS_THREADS, FS_SMEM, stream>>>(A, ep->n_samples, ep->batch, item_cnt,
                                                                 (uint4*)(P + L.items), L.item_cap, long_cnt,
                                                                 (uint4*)(P + L.long_segs), L.long_cap);
    } else {
        rc = sort_space(ep->user, ep->user, 1, 0, 1, model->user.n_rows, ep, (uint32_t*)(P + L.user_key),
                        (uint32_t*)(P + L.user_perm), tmp_key, tmp_val, hist, stream);
        if (rc) return rc;
        rc = sort_space(ep->pos, ep->neg, 1, 0, 2, model->item.n_rows, ep, (uint32_t*)(P + L.item_key),
                        (uint32_t*)(P + L.item_perm), tmp_key, tmp_val, hist, stream);
        if (rc) return rc;
        for (int f = 0; f < model->n_meta; ++f) {
            rc = sort_space(ep->pos_meta, ep->neg_meta, model->n_meta, f, 2, model->meta[f].n_rows, ep,
                            (uint32_t*)(P + L.meta_key[f]), (uint32_t*)(P + L.meta_perm[f]), tmp_key,
                            tmp_val, hist, stream);
            if (rc) return rc;
        }
        // work items of every id space
        auto build_items = [&](size_t key_off, size_t perm_off, int mult, int space) {
            dim3 grid((unsigned)(((int64_t)mult * ep->batch + 255) / 256), (unsigned)steps);
            build_items_kernel<<<grid, 256, 0, stream>>>(
                (const uint32_t*)(P + key_off), (const uint32_t*)(P + perm_off), mult, space, ep->n_samples,
                ep->batch, item_cnt, (uint4*)(P + L.items), L.item_cap, long_cnt, (uint4*)(P + L.long_segs),
                L.long_cap, single_of(space));
        };
        build_items(L.user_key, L.user_perm, 1, 0);
        build_items(L.item_key, L.item_perm, 2, 1);
        for (int f = 0; f < model->n_meta; ++f) build_items(L.meta_key[f], L.meta_perm[f], 2, 2 + f);
    }
    if (fuse_single && steps > 1 && !fused) {
        dim3 gu((unsigned)((ep->batch + 255) / 256), (unsigned)(steps - 1));
        dirty_flags_kernel<<<gu, 256, 0, stream>>>((const uint32_t*)(P + L.user_key), ep->user, ep->user, 1,
                                                   ep->n_samples, ep->batch, (uint8_t*)(P + L.single_user));
        dim3 gi((unsigned)((2ll * ep->batch + 255) / 256), (unsigned)(steps - 1));
        dirty_flags_kernel<<<gi, 256, 0, stream>>>((const uint32_t*)(P + L.item_key), ep->pos, ep->neg, 2,
                                                   ep->n_samples, ep->batch, (uint8_t*)(P + L.single_item));
    }
    TRS_CUDA(cudaGetLastError());
    return TRS_OK;
}
