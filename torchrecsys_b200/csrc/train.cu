// train.cu -- fused forward + hinge + backward + segmented reduce + row-wise optimizer for the
// Linear and FM scorers, as ONE persistent cooperative kernel over many steps.
//
// Per step (reference model.py:274-284):
//   phase A  one row group per sample: gather u, v+, v- (+ metadata rows), both scores, hinge,
//            closed-form gradient rows -> L2-resident staging, one row per lookup.
//            While it waits it prefetches into L2 the optimizer-state rows phase B will need.
//   grid barrier (every score of the step is computed from pre-update parameters)
//   phase B  work items from the plan (plan.cu): a short segment = one touched row whose few
//            lookups are summed in lookup order, then param+state are read once, updated
//            (SGD / Adagrad / SparseAdam) and written once; a chunk = 32 lookups of a hot row,
//            summed to a partial, the last chunk to arrive adds the partials in chunk order and
//            updates.  First thing in phase B, the rows of the NEXT step's samples are prefetched
//            into L2, so that phase A's gathers hit L2 as well.
//   grid barrier
// HBM traffic per step is ids + (param+state read, param+state write) per unique touched row;
// staging, plan and partials are served from L2.  Nothing depends on the order in which CTAs or
// row groups run: every floating-point sum has a fixed association (deterministic results).
#include <stdlib.h>

#include "plan.cuh"
#include "scorer.cuh"
#include "train.cuh"

namespace trs {

constexpr int MF = 2;  // metadata features whose rows are kept in registers between fwd and bwd

struct PlanPtrs {
    const uint32_t *user_key, *user_perm, *item_key, *item_perm;
    const uint32_t* meta_key[TRS_MAX_META];
    const uint32_t* meta_perm[TRS_MAX_META];
    const uint32_t* item_cnt;
    const uint32_t* chunk_cnt;
    const uint4* items;
    const uint4* chunks;
    const uint4* long_segs;
    int item_cap, long_cap, chunk_cap;
};

struct Stage {
    float* gU;                 // [B, dim]
    float* gI;                 // [2B, dim]
    float* gM[TRS_MAX_META];   // FM only: [2B, dim]
    float* gbU;                // FM only: [B]
    float* gbI;                // [2B]  (FM: also the gradient of linear_metadata, d w_k = delta)
    float* loss_part;          // [n_steps, gridDim.x]
    float* partials;           // [chunk_cap, dim]   partial sums of long-segment chunks
    float* partials_lin;       // [chunk_cap]
    unsigned* seg_arrive;      // [long_cap]  finished chunks per long segment (self-resetting)
    unsigned* barrier;         // [1] monotonically increasing arrival counter
    unsigned long long* trace; // debug (TRS_DEBUG_SKIP & 64): [n_steps, gridDim.x, 4] globaltimer stamps
};


struct StageLayout {
    size_t gU, gI, gM[TRS_MAX_META], gbU, gbI, loss_part, partials, partials_lin, sync_words, trace, total;
    size_t sync_bytes;
};

static StageLayout stage_layout(const trs_model* m, const trs_epoch* ep, int grid) {
    StageLayout L;
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        size_t o = off;
        off += (n_floats * sizeof(float) + 255) / 256 * 256;
        return o;
    };
    const PlanLayout PL = plan_layout(ep->n_samples, ep->batch, m->n_meta);
    const size_t B = (size_t)ep->batch, D = (size_t)m->dim;
    L.gU = take(B * D);
    L.gI = take(2 * B * D);
    for (int f = 0; f < TRS_MAX_META; ++f) L.gM[f] = (m->net != TRS_NET_LINEAR && f < m->n_meta) ? take(2 * B * D) : 0;
    L.gbU = take(B);
    L.gbI = take(2 * B);
    L.loss_part = take((size_t)n_steps_of(ep) * grid);
    L.partials = take((size_t)PL.chunk_cap * D);
    L.partials_lin = take((size_t)PL.chunk_cap);
    L.sync_words = off;  // seg_arrive[long_cap], pad, barrier[1]; zeroed before every launch
    L.sync_bytes = ((size_t)PL.long_cap + 64) * sizeof(unsigned);
    off += (L.sync_bytes + 255) / 256 * 256;
    L.trace = take((size_t)n_steps_of(ep) * grid * 16 * 2);
    L.total = off;
    return L;
}

// ---- small device helpers -------------------------------------------------------------------
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)::"memory");
    return t;
}
__device__ __forceinline__ void prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
template <int V, int G, int IT>
__device__ __forceinline__ void prefetch_row(const float* base, int nch, int gl) {
#pragma unroll
    for (int i = 0; i < IT; ++i) {
        const int c = gl + i * G;
        if (c < nch) prefetch_l2(base + (size_t)c * V);
    }
}
template <int V, int IT>
__device__ __forceinline__ void row_zero(Row<V, IT>& r) {
#pragma unroll
    for (int i = 0; i < IT; ++i) r.c[i] = Vec<V>::zero();
}
template <int V, int IT>
__device__ __forceinline__ void row_acc(Row<V, IT>& a, const Row<V, IT>& b) {  // a += b, unfused
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) a.c[i][k] = __fadd_rn(a.c[i][k], b.c[i][k]);
}
template <int V, int IT>
__device__ __forceinline__ Row<V, IT> row_scaled_diff(float a, const Row<V, IT>& x,
                                                      const Row<V, IT>& y) {  // a*(x-y)
    Row<V, IT> r;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) r.c[i][k] = a * (x.c[i][k] - y.c[i][k]);
    return r;
}
template <int V, int IT>
__device__ __forceinline__ Row<V, IT> row_scaled(float a, const Row<V, IT>& x) {
    Row<V, IT> r;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) r.c[i][k] = a * x.c[i][k];
    return r;
}

// Grid-wide barrier for a cooperative launch (all CTAs resident).  `counter` only ever grows:
// after the g-th barrier it holds g * gridDim.x.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& target) {
    __syncthreads();
    target += gridDim.x;
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned seen;
        do {  // relaxed polling; the fence below orders everything after the barrier
            asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
        } while ((int)(seen - target) < 0);
        __threadfence();
    }
    __syncthreads();
}

// Which arrays an id space (0 user, 1 item, 2+f metadata f) reduces from / updates.
struct SpaceRef {
    const trs_table* t;
    const uint32_t* K;
    const uint32_t* P;
    const float* stage;
    const float* stage_lin;
};

template <int NET>
__device__ __forceinline__ SpaceRef resolve_space(int space, const trs_model& m, const PlanPtrs& plan,
                                                  const Stage& st, int64_t lo) {
    SpaceRef r;
    if (space == 0) {
        r = {&m.user, plan.user_key + lo, plan.user_perm + lo, st.gU, NET == TRS_NET_FM ? st.gbU : nullptr};
    } else if (space == 1) {
        r = {&m.item, plan.item_key + 2 * lo, plan.item_perm + 2 * lo, st.gI, st.gbI};
    } else {
        const int f = space - 2;
        r = {&m.meta[f], plan.meta_key[f] + 2 * lo, plan.meta_perm[f] + 2 * lo,
             NET != TRS_NET_LINEAR ? st.gM[f] : st.gI, NET == TRS_NET_FM ? st.gbI : nullptr};
    }
    if (!r.t->lin) r.stage_lin = nullptr;
    return r;
}

// The row being updated: parameter + optimizer state, and its width-1 companion.
template <int V, int IT>
struct RowState {
    Row<V, IT> p, s0, s1;
    float pl, l0, l1;
};

template <int V, int G, int IT>
__device__ __forceinline__ RowState<V, IT> load_state(const trs_table& t, uint32_t key, int dim, int nch,
                                                      int gl, int kind, bool lin) {
    RowState<V, IT> r;
    const size_t roff = (size_t)key * dim;
    r.p = load_row_cg<V, G, IT>(t.emb + roff, nch, gl);
    row_zero(r.s0);
    row_zero(r.s1);
    if (kind != TRS_OPT_SGD) r.s0 = load_row_cg<V, G, IT>(t.emb_s0 + roff, nch, gl);
    if (kind == TRS_OPT_SPARSE_ADAM) r.s1 = load_row_cg<V, G, IT>(t.emb_s1 + roff, nch, gl);
    r.pl = r.l0 = r.l1 = 0.f;
    if (lin) {
        r.pl = __ldcg(t.lin + key);
        if (kind != TRS_OPT_SGD) r.l0 = __ldcg(t.lin_s0 + key);
        if (kind == TRS_OPT_SPARSE_ADAM) r.l1 = __ldcg(t.lin_s1 + key);
    }
    return r;
}

template <int V, int G, int IT>
__device__ __forceinline__ void apply_and_store(const trs_table& t, uint32_t key, int dim, int nch, int gl,
                                                const OptScalars& o, float scale, RowState<V, IT>& r,
                                                const Row<V, IT>& g, float g_lin, bool lin) {
#pragma unroll
    for (int a = 0; a < IT; ++a)
#pragma unroll
        for (int b = 0; b < V; ++b) opt_update(o, scale, g.c[a][b], r.p.c[a][b], r.s0.c[a][b], r.s1.c[a][b]);
    const size_t roff = (size_t)key * dim;
    store_row<V, G, IT>(t.emb + roff, nch, gl, r.p);
    if (o.kind != TRS_OPT_SGD) store_row<V, G, IT>(t.emb_s0 + roff, nch, gl, r.s0);
    if (o.kind == TRS_OPT_SPARSE_ADAM) store_row<V, G, IT>(t.emb_s1 + roff, nch, gl, r.s1);
    if (lin && gl == 0) {
        opt_update(o, scale, g_lin, r.pl, r.l0, r.l1);
        t.lin[key] = r.pl;
        if (o.kind != TRS_OPT_SGD) t.lin_s0[key] = r.l0;
        if (o.kind == TRS_OPT_SPARSE_ADAM) t.lin_s1[key] = r.l1;
    }
}

// ---- cp.async ring: per-thread prefetch slots in shared memory (V == 4 only) -------------------
// Every lane copies only the 16-byte chunks it will itself consume, so no cross-thread
// synchronisation is needed: cp.async.wait_group makes a thread's own copies visible to it.
constexpr int RING = 4;
template <int V, int IT>
#ifndef TRS_TRAIN_T1
#define TRS_TRAIN_T1 512  // threads per CTA for rows of <= 32 chunks (tuning hook: -DTRS_TRAIN_T1=256|384|512)
#endif
constexpr int train_threads() { return V == 1 ? 256 : (IT == 1 ? TRS_TRAIN_T1 : (IT == 2 ? 256 : 128)); }
template <int V, int IT>
constexpr size_t train_smem_bytes() {
    return V == 1 ? 0 : (size_t)train_threads<V, IT>() * (2 * RING * 16 + RING * 4 * IT * 16);
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, int IT>
struct Ring {
    uint4* desc;   // [2*RING][NT]
    float4* rows;  // [RING][4 fields: p, s0, s1, grad][IT][NT]
    __device__ __forceinline__ uint4* d(int n) const { return desc + (n % (2 * RING)) * NT + threadIdx.x; }
    __device__ __forceinline__ float4* r(int n, int f, int i) const {
        return rows + (((n % RING) * 4 + f) * IT + i) * NT + threadIdx.x;
    }
};

// A chunk of a long segment: sum its (<= LONG_CHUNK) staged rows in lookup order into a partial;
// the group whose chunk completes the segment adds the partials in chunk order and updates the row.
// All lookup ids of the chunk are fetched in one go, then all its rows (one L2 round trip each).
template <int NET, int V, int G, int IT>
__device__ __noinline__ void chunk_item(const uint4 it, int ci, const trs_model& m, const PlanPtrs& plan,
                                           const Stage& st, int64_t s, int64_t lo, int dim, int nch, int gl,
                                           const OptScalars& o, float scale) {
    constexpr int NB = (LONG_CHUNK / 2) / IT > 0 ? (LONG_CHUNK / 2) / IT : 1;  // rows in flight per batch (register budget)
    const SpaceRef sp = resolve_space<NET>((int)(it.x & 0xffu), m, plan, st, lo);
    const int cc = (int)((it.x >> 8) & 0xffu);
    const uint32_t* P = sp.P + it.y;
    const uint4 seg = plan.long_segs[(size_t)s * plan.long_cap + it.z];
    uint32_t j[LONG_CHUNK];
#pragma unroll
    for (int q = 0; q < LONG_CHUNK; ++q) j[q] = P[q < cc ? q : 0];
    const uint32_t key = sp.K[seg.y];
    if (it.w == 0) {  // first chunk of the segment: pull the row's param + state towards L2 now
        const size_t roff = (size_t)key * dim;
        prefetch_row<V, G, IT>(sp.t->emb + roff, nch, gl);
        if (o.kind != TRS_OPT_SGD) prefetch_row<V, G, IT>(sp.t->emb_s0 + roff, nch, gl);
        if (o.kind == TRS_OPT_SPARSE_ADAM) prefetch_row<V, G, IT>(sp.t->emb_s1 + roff, nch, gl);
    }
    Row<V, IT> acc;
    row_zero(acc);
    float accl = 0.f;
#pragma unroll
    for (int b0 = 0; b0 < LONG_CHUNK; b0 += NB) {
        if (b0 < cc) {
            Row<V, IT> r[NB];
            float l[NB];
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                if (b0 + q < cc) {
                    r[q] = load_row_cg<V, G, IT>(sp.stage + (size_t)j[b0 + q] * dim, nch, gl);
                    l[q] = sp.stage_lin ? __ldcg(sp.stage_lin + j[b0 + q]) : 0.f;
                }
            }
#pragma unroll
            for (int q = 0; q < NB; ++q) {
                if (b0 + q < cc) {
                    row_acc(acc, r[q]);
                    accl = __fadd_rn(accl, l[q]);
                }
            }
        }
    }
    const uint32_t first = seg.w;  // index of the segment's first chunk == its first partial slot
    const int n_chunks = ((int)seg.z + LONG_CHUNK - 1) / LONG_CHUNK;
    store_row<V, G, IT>(st.partials + (size_t)ci * dim, nch, gl, acc);
    if (gl == 0) st.partials_lin[ci] = accl;
    __threadfence();
    constexpr unsigned GM = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    const unsigned gmask = GM << ((threadIdx.x & 31) & ~(G - 1));
    __syncwarp(gmask);
    unsigned old = 0;
    if (gl == 0) old = atomicAdd(&st.seg_arrive[it.z], 1u);
    old = __shfl_sync(gmask, old, 0, G);
    if ((int)old != n_chunks - 1) return;
    __threadfence();
    if (gl == 0) st.seg_arrive[it.z] = 0u;  // ready for the next step
    RowState<V, IT> rs = load_state<V, G, IT>(*sp.t, key, dim, nch, gl, o.kind, sp.stage_lin != nullptr);
    Row<V, IT> g;
    row_zero(g);
    float g_lin = 0.f;
    constexpr int PB = 8 / (IT > 2 ? 2 : 1);
    for (int q = 0; q < n_chunks; q += PB) {
        Row<V, IT> r[PB];
        float l[PB];
#pragma unroll
        for (int z = 0; z < PB; ++z) {
            if (q + z < n_chunks) {
                r[z] = load_row_cg<V, G, IT>(st.partials + (size_t)(first + q + z) * dim, nch, gl);
                l[z] = __ldcg(st.partials_lin + first + q + z);
            }
        }
#pragma unroll
        for (int z = 0; z < PB; ++z) {
            if (q + z < n_chunks) {
                row_acc(g, r[z]);
                g_lin = __fadd_rn(g_lin, l[z]);
            }
        }
    }
    apply_and_store<V, G, IT>(*sp.t, key, dim, nch, gl, o, scale, rs, g, g_lin, sp.stage_lin != nullptr);
}

// ids of one sample, as 32-bit row numbers (every table has < 2^32 rows, checked by plan_build)
struct SampleIds {
    uint32_t u, ip, in, pm[MF], nm[MF];
};
struct SampleIds3 {
    uint32_t u, ip, in;
};
__device__ __forceinline__ SampleIds3 load_ids3(const trs_epoch& ep, int64_t smp) {
    SampleIds3 r;
    r.u = (uint32_t)ep.user[smp];
    r.ip = (uint32_t)ep.pos[smp];
    r.in = (uint32_t)ep.neg[smp];
    return r;
}
__device__ __forceinline__ SampleIds with_meta(const trs_epoch& ep, const SampleIds3& a, int64_t smp, int F) {
    SampleIds r;
    r.u = a.u;
    r.ip = a.ip;
    r.in = a.in;
#pragma unroll
    for (int f = 0; f < MF; ++f) {
        r.pm[f] = f < F ? (uint32_t)ep.pos_meta[smp * F + f] : 0u;
        r.nm[f] = f < F ? (uint32_t)ep.neg_meta[smp * F + f] : 0u;
    }
    return r;
}
__device__ __forceinline__ SampleIds load_ids(const trs_epoch& ep, int64_t smp, int F) {
    SampleIds r;
    r.u = (uint32_t)ep.user[smp];
    r.ip = (uint32_t)ep.pos[smp];
    r.in = (uint32_t)ep.neg[smp];
#pragma unroll
    for (int f = 0; f < MF; ++f) {
        r.pm[f] = f < F ? (uint32_t)ep.pos_meta[smp * F + f] : 0u;
        r.nm[f] = f < F ? (uint32_t)ep.neg_meta[smp * F + f] : 0u;
    }
    return r;
}

// ---- the kernel -------------------------------------------------------------------------------
template <int NET, int V, int G, int IT>
__global__ void __launch_bounds__((train_threads<V, IT>()), 1)
train_kernel(const __grid_constant__ trs_model m, const __grid_constant__ trs_epoch ep,
             const __grid_constant__ OptScalars opt, const __grid_constant__ PlanPtrs plan,
             const __grid_constant__ Stage st, const int first_step, const int n_steps,
             float* __restrict__ loss_out, const int dbg) {
    constexpr int NT = train_threads<V, IT>();
    constexpr bool RINGED = (V == 4);
    __shared__ float s_loss[NT / 32];
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Ring<NT, IT> ring;
    ring.desc = reinterpret_cast<uint4*>(smem_raw);
    ring.rows = reinterpret_cast<float4*>(smem_raw + (size_t)2 * RING * NT * sizeof(uint4));

    const int dim = m.dim, nch = dim / V, F = m.n_meta;
    const int gl = threadIdx.x % G;
    constexpr int GPW = 32 / G;                        // groups per warp
    constexpr int GPB = NT / G;                        // groups per block
    const int gid = blockIdx.x * GPB + threadIdx.x / G;
    const int ngroups = gridDim.x * GPB;
    const int gid_warp0 = gid - (gid % GPW);           // first group of my warp
    const int kind = opt.kind;
    unsigned bar_target = 0;

    // ---- ring stages (phase B), item n of this group = global item n*ngroups + gid ----
    auto issue_desc = [&](const uint4* items, int n_items, int n) {
        const int i = n * ngroups + gid;
        if (RINGED && i < n_items) cp_async16(ring.d(n), items + i);
    };
    auto issue_state = [&](int n_items, int n, int64_t lo) {   // needs desc n landed
        const int i = n * ngroups + gid;
        if (!RINGED || i >= n_items) return;
        const uint4 it = *ring.d(n);
        const SpaceRef sp = resolve_space<NET>((int)(it.x & 0xffu), m, plan, st, lo);
        const size_t roff = (size_t)it.z * dim;
#pragma unroll
        for (int a = 0; a < IT; ++a) {
            const int c = gl + a * G;
            if (c < nch) {
                cp_async16(ring.r(n, 0, a), sp.t->emb + roff + (size_t)c * 4);
                if (kind != TRS_OPT_SGD) cp_async16(ring.r(n, 1, a), sp.t->emb_s0 + roff + (size_t)c * 4);
                if (kind == TRS_OPT_SPARSE_ADAM) cp_async16(ring.r(n, 2, a), sp.t->emb_s1 + roff + (size_t)c * 4);
            }
        }
    };
    auto issue_grad = [&](int n_items, int n, int64_t lo) {    // after the barrier: staged gradient row
        const int i = n * ngroups + gid;
        if (!RINGED || i >= n_items) return;
        const uint4 it = *ring.d(n);
        const SpaceRef sp = resolve_space<NET>((int)(it.x & 0xffu), m, plan, st, lo);
#pragma unroll
        for (int a = 0; a < IT; ++a) {
            const int c = gl + a * G;
            if (c < nch) cp_async16(ring.r(n, 3, a), sp.stage + (size_t)it.w * dim + (size_t)c * 4);
        }
    };
    auto consume = [&](const uint4* items, int n_items, int n, int64_t lo, float scale) {
        const int i = n * ngroups + gid;
        if (i >= n_items) return;
        uint4 it;
        Row<V, IT> p, s0, s1, g;
        row_zero(s0);
        row_zero(s1);
        if (RINGED) {
            it = *ring.d(n);
#pragma unroll
            for (int a = 0; a < IT; ++a) {
                const int c = gl + a * G;
                if (c < nch) {
                    const float4 vp = *ring.r(n, 0, a), vg = *ring.r(n, 3, a);
                    p.c[a][0] = vp.x; p.c[a][V > 1 ? 1 : 0] = vp.y; p.c[a][V > 2 ? 2 : 0] = vp.z; p.c[a][V > 3 ? 3 : 0] = vp.w;
                    g.c[a][0] = vg.x; g.c[a][V > 1 ? 1 : 0] = vg.y; g.c[a][V > 2 ? 2 : 0] = vg.z; g.c[a][V > 3 ? 3 : 0] = vg.w;
                    if (kind != TRS_OPT_SGD) {
                        const float4 v0 = *ring.r(n, 1, a);
                        s0.c[a][0] = v0.x; s0.c[a][V > 1 ? 1 : 0] = v0.y; s0.c[a][V > 2 ? 2 : 0] = v0.z; s0.c[a][V > 3 ? 3 : 0] = v0.w;
                    }
                    if (kind == TRS_OPT_SPARSE_ADAM) {
                        const float4 v1 = *ring.r(n, 2, a);
                        s1.c[a][0] = v1.x; s1.c[a][V > 1 ? 1 : 0] = v1.y; s1.c[a][V > 2 ? 2 : 0] = v1.z; s1.c[a][V > 3 ? 3 : 0] = v1.w;
                    }
                } else {
                    p.c[a] = Vec<V>::zero();
                    g.c[a] = Vec<V>::zero();
                }
            }
        } else {
            it = items[i];
        }
        const SpaceRef sp = resolve_space<NET>((int)(it.x & 0xffu), m, plan, st, lo);
        const size_t roff = (size_t)it.z * dim;
        if (!RINGED) {
            g = load_row_cg<V, G, IT>(sp.stage + (size_t)it.w * dim, nch, gl);
            p = load_row_cg<V, G, IT>(sp.t->emb + roff, nch, gl);
            if (kind != TRS_OPT_SGD) s0 = load_row_cg<V, G, IT>(sp.t->emb_s0 + roff, nch, gl);
            if (kind == TRS_OPT_SPARSE_ADAM) s1 = load_row_cg<V, G, IT>(sp.t->emb_s1 + roff, nch, gl);
        }
        const int c = (int)((it.x >> 8) & 0xffu);
        for (int q = 1; q < c; ++q) {  // duplicates, in lookup order
            const uint32_t j = sp.P[it.y + q];
            row_acc(g, load_row_cg<V, G, IT>(sp.stage + (size_t)j * dim, nch, gl));
        }
#pragma unroll
        for (int a = 0; a < IT; ++a)
#pragma unroll
            for (int b = 0; b < V; ++b) opt_update(opt, scale, g.c[a][b], p.c[a][b], s0.c[a][b], s1.c[a][b]);
        store_row<V, G, IT>(sp.t->emb + roff, nch, gl, p);
        if (kind != TRS_OPT_SGD) store_row<V, G, IT>(sp.t->emb_s0 + roff, nch, gl, s0);
        if (kind == TRS_OPT_SPARSE_ADAM) store_row<V, G, IT>(sp.t->emb_s1 + roff, nch, gl, s1);
    };

    // ---- prologue: what step `first_step` needs before its phase A ----
    int n_items_cur = 0, n_chunks_cur = 0;
    SampleIds ids0;   // ids of this group's first sample of the current step
    SampleIds3 ids1;  // and (user, pos, neg) of its second
    {
        const int64_t s0i = first_step;
        const int64_t lo0 = s0i * (int64_t)ep.batch;
        const int Bs0 = (int)min((int64_t)ep.batch, ep.n_samples - lo0);
        n_items_cur = min((int)plan.item_cnt[s0i], plan.item_cap);
        n_chunks_cur = min((int)plan.chunk_cnt[s0i], plan.chunk_cap);
        ids0 = load_ids(ep, lo0 + min(gid, Bs0 - 1), F);
        ids1 = load_ids3(ep, lo0 + min(gid + ngroups, Bs0 - 1));
        const uint4* items0 = plan.items + (size_t)s0i * plan.item_cap;
        for (int n = 0; n < 2 * RING; ++n) issue_desc(items0, n_items_cur, n);
        cp_async_commit();
    }

    for (int si = 0; si < n_steps; ++si) {
        const int64_t s = first_step + si;
        const int64_t lo = s * (int64_t)ep.batch;
        const int Bs = (int)min((int64_t)ep.batch, ep.n_samples - lo);
        const float invB = 1.0f / (float)Bs;
        const uint4* items = plan.items + (size_t)s * plan.item_cap;
        const int n_items = n_items_cur, n_chunks = n_chunks_cur;
        const float scale = opt.step_scale[s];
        // what the next step will need (loaded now, used after two barriers)
        int n_items_next = 0, n_chunks_next = 0;
        SampleIds nid0 = ids0;  // ids of my first two samples of the NEXT step
        SampleIds3 nid1 = ids1;
        const int64_t lo2 = lo + ep.batch;
        const int Bs2 = (si + 1 < n_steps) ? (int)min((int64_t)ep.batch, ep.n_samples - lo2) : 0;
        if (si + 1 < n_steps) {
            n_items_next = min((int)plan.item_cnt[s + 1], plan.item_cap);
            n_chunks_next = min((int)plan.chunk_cnt[s + 1], plan.chunk_cap);
            nid0 = load_ids(ep, lo2 + min(gid, Bs2 - 1), F);
            nid1 = load_ids3(ep, lo2 + min(gid + ngroups, Bs2 - 1));
        }
        // width-1 companion work of this thread (phase B): descriptor fetched now
        const int lin_i = threadIdx.x * gridDim.x + blockIdx.x;  // every CTA gets every gridDim-th item
        uint4 lin_it = make_uint4(0, 0, 0, 0);
        if (lin_i < n_items) lin_it = items[lin_i];

        // debug trace: slots 0-7 by thread 0 (a warp that also owns chunks), 8-15 by thread NT/2
        const bool tracing = (dbg & 64) && (threadIdx.x == 0 || threadIdx.x == NT / 2);
        unsigned long long* tr = st.trace + ((size_t)si * gridDim.x + blockIdx.x) * 16 + (threadIdx.x ? 8 : 0);
        if (tracing) tr[0] = global_ns();
        // ------------------------------ phase A ------------------------------------------
        float hsum = 0.f;
        // NET == TRS_NET_MLP: the gradient rows were staged by the tower's backward (mlp.cu); this
        // kernel only runs the reduce + update phase
        if constexpr (NET != TRS_NET_MLP)
        if (!(dbg & 1))
        for (int b0 = gid_warp0; b0 < Bs; b0 += ngroups) {   // warp-uniform trip count
            const int b_raw = b0 + (gid - gid_warp0);
            const bool valid = b_raw < Bs;
            const int b = valid ? b_raw : Bs - 1;
            const int64_t smp = lo + b;
            const SampleIds id = (b0 == gid_warp0) ? ids0 : ((b0 == gid_warp0 + ngroups) ? with_meta(ep, ids1, smp, F) : load_ids(ep, smp, F));
            const int64_t* pm = F ? ep.pos_meta + smp * F : nullptr;
            const int64_t* nm = F ? ep.neg_meta + smp * F : nullptr;
            // every gather of the sample issued back to back (L2 only: rows are rewritten by other SMs)
            const Row<V, IT> ru = load_row_cg<V, G, IT>(m.user.emb + (size_t)id.u * dim, nch, gl);
            const Row<V, IT> rp = load_row_cg<V, G, IT>(m.item.emb + (size_t)id.ip * dim, nch, gl);
            const Row<V, IT> rn = load_row_cg<V, G, IT>(m.item.emb + (size_t)id.in * dim, nch, gl);
            Row<V, IT> mp[MF], mn[MF];
            float wp = 0.f, wn = 0.f;  // FM: sum of the metadata first-order weights
#pragma unroll
            for (int f = 0; f < MF; ++f) {
                if (f < F) {
                    mp[f] = load_row_cg<V, G, IT>(m.meta[f].emb + (size_t)id.pm[f] * dim, nch, gl);
                    mn[f] = load_row_cg<V, G, IT>(m.meta[f].emb + (size_t)id.nm[f] * dim, nch, gl);
                    if (NET == TRS_NET_FM && m.meta[f].lin) {
                        wp += __ldcg(m.meta[f].lin + id.pm[f]);
                        wn += __ldcg(m.meta[f].lin + id.nm[f]);
                    }
                } else {
                    row_zero(mp[f]);
                    row_zero(mn[f]);
                }
            }
            const float bu = m.user.lin ? __ldcg(m.user.lin + id.u) : 0.f;
            const float bip = m.item.lin ? __ldcg(m.item.lin + id.ip) : 0.f;
            const float bin = m.item.lin ? __ldcg(m.item.lin + id.in) : 0.f;
            // optimizer state of the rows this sample touches -> L2, for phase B's ring refills
            if (kind != TRS_OPT_SGD && valid && !(dbg & 16)) {
                prefetch_row<V, G, IT>(m.user.emb_s0 + (size_t)id.u * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb_s0 + (size_t)id.ip * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb_s0 + (size_t)id.in * dim, nch, gl);
                if (kind == TRS_OPT_SPARSE_ADAM) {
                    prefetch_row<V, G, IT>(m.user.emb_s1 + (size_t)id.u * dim, nch, gl);
                    prefetch_row<V, G, IT>(m.item.emb_s1 + (size_t)id.ip * dim, nch, gl);
                    prefetch_row<V, G, IT>(m.item.emb_s1 + (size_t)id.in * dim, nch, gl);
                }
            }
            // pooled sums over the fields beyond MF (rare): Sx = sum of rows, Qx = sum of squares
            Row<V, IT> Sp_x, Sn_x, Qp_x, Qn_x;
            row_zero(Sp_x); row_zero(Sn_x); row_zero(Qp_x); row_zero(Qn_x);
            for (int f = MF; f < F; ++f) {
                const Row<V, IT> a = load_row_cg<V, G, IT>(m.meta[f].emb + (size_t)pm[f] * dim, nch, gl);
                const Row<V, IT> c = load_row_cg<V, G, IT>(m.meta[f].emb + (size_t)nm[f] * dim, nch, gl);
#pragma unroll
                for (int i = 0; i < IT; ++i)
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        Sp_x.c[i][k] += a.c[i][k];
                        Sn_x.c[i][k] += c.c[i][k];
                        Qp_x.c[i][k] = fmaf(a.c[i][k], a.c[i][k], Qp_x.c[i][k]);
                        Qn_x.c[i][k] = fmaf(c.c[i][k], c.c[i][k], Qn_x.c[i][k]);
                    }
                if (NET == TRS_NET_FM && m.meta[f].lin) {
                    wp += __ldcg(m.meta[f].lin + pm[f]);
                    wn += __ldcg(m.meta[f].lin + nm[f]);
                }
            }

            if (NET == TRS_NET_LINEAR) {
                // v = item + sum_f meta_f (feature order), s = <u, v> + b_u + b_i
                Row<V, IT> vp = rp, vn = rn;
#pragma unroll
                for (int f = 0; f < MF; ++f) {
                    row_add(vp, mp[f]);
                    row_add(vn, mn[f]);
                }
                row_add(vp, Sp_x);
                row_add(vn, Sn_x);
                const float sp = (group_sum<G>(row_dot_partial(ru, vp)) + bu) + bip;
                const float sn = (group_sum<G>(row_dot_partial(ru, vn)) + bu) + bin;
                const float h = __fadd_rn(__fsub_rn(sn, sp), 1.0f);
                const float g = (h >= 0.f) ? invB : 0.f;
                if (valid) {
                    if (gl == 0) hsum += fmaxf(h, 0.f);
                    store_row<V, G, IT>(st.gU + (size_t)b * dim, nch, gl, row_scaled_diff(g, vn, vp));
                    store_row<V, G, IT>(st.gI + (size_t)b * dim, nch, gl, row_scaled(-g, ru));
                    store_row<V, G, IT>(st.gI + (size_t)(Bs + b) * dim, nch, gl, row_scaled(g, ru));
                    if (gl == 0) {
                        st.gbI[b] = -g;
                        st.gbI[Bs + b] = g;
                    }
                }
            } else {
                // S = sum_k e_k, Q = sum_k e_k^2 over fields user, item, meta_f
                Row<V, IT> Sp, Sn;
                float pp = 0.f, pn = 0.f;
#pragma unroll
                for (int i = 0; i < IT; ++i)
#pragma unroll
                    for (int k = 0; k < V; ++k) {
                        const float a = ru.c[i][k], x = rp.c[i][k], y = rn.c[i][k];
                        float sp_ = a + x, sn_ = a + y;
                        float qp_ = a * a + x * x, qn_ = a * a + y * y;
#pragma unroll
                        for (int f = 0; f < MF; ++f) {
                            sp_ += mp[f].c[i][k];
                            sn_ += mn[f].c[i][k];
                            qp_ = fmaf(mp[f].c[i][k], mp[f].c[i][k], qp_);
                            qn_ = fmaf(mn[f].c[i][k], mn[f].c[i][k], qn_);
                        }
                        sp_ += Sp_x.c[i][k];
                        sn_ += Sn_x.c[i][k];
                        qp_ += Qp_x.c[i][k];
                        qn_ += Qn_x.c[i][k];
                        Sp.c[i][k] = sp_;
                        Sn.c[i][k] = sn_;
                        pp += sp_ * sp_ - qp_;
                        pn += sn_ * sn_ - qn_;
                    }
                const float zp = ((bu + bip) + wp) + group_sum<G>(pp) * 0.5f;
                const float zn = ((bu + bin) + wn) + group_sum<G>(pn) * 0.5f;
                const float sp = sigmoidf_acc(zp), sn = sigmoidf_acc(zn);
                const float h = __fadd_rn(__fsub_rn(sn, sp), 1.0f);
                const float g = (h >= 0.f) ? invB : 0.f;
                const float dp = -g * sp * (1.0f - sp);
                const float dn = g * sn * (1.0f - sn);
                if (valid) {
                    if (gl == 0) hsum += fmaxf(h, 0.f);
                    Row<V, IT> gu;
#pragma unroll
                    for (int i = 0; i < IT; ++i)
#pragma unroll
                        for (int k = 0; k < V; ++k)
                            // two separately rounded products, like the reference's two lookups: when
                            // pos == neg they cancel to an exact 0 (an FMA would leave a ~1e-10 residue
                            // that Adagrad/Adam's g/(|g|+eps) blows up into a step of ~lr)
                            gu.c[i][k] = __fadd_rn(__fmul_rn(dp, Sp.c[i][k] - ru.c[i][k]),
                                                   __fmul_rn(dn, Sn.c[i][k] - ru.c[i][k]));
                    store_row<V, G, IT>(st.gU + (size_t)b * dim, nch, gl, gu);
                    store_row<V, G, IT>(st.gI + (size_t)b * dim, nch, gl, row_scaled_diff(dp, Sp, rp));
                    store_row<V, G, IT>(st.gI + (size_t)(Bs + b) * dim, nch, gl, row_scaled_diff(dn, Sn, rn));
#pragma unroll
                    for (int f = 0; f < MF; ++f) {
                        if (f < F) {
                            store_row<V, G, IT>(st.gM[f] + (size_t)b * dim, nch, gl, row_scaled_diff(dp, Sp, mp[f]));
                            store_row<V, G, IT>(st.gM[f] + (size_t)(Bs + b) * dim, nch, gl, row_scaled_diff(dn, Sn, mn[f]));
                        }
                    }
                    for (int f = MF; f < F; ++f) {
                        Row<V, IT> r = load_row_cg<V, G, IT>(m.meta[f].emb + (size_t)pm[f] * dim, nch, gl);
                        store_row<V, G, IT>(st.gM[f] + (size_t)b * dim, nch, gl, row_scaled_diff(dp, Sp, r));
                        r = load_row_cg<V, G, IT>(m.meta[f].emb + (size_t)nm[f] * dim, nch, gl);
                        store_row<V, G, IT>(st.gM[f] + (size_t)(Bs + b) * dim, nch, gl, row_scaled_diff(dn, Sn, r));
                    }
                    if (gl == 0) {
                        st.gbU[b] = dp + dn;
                        st.gbI[b] = dp;
                        st.gbI[Bs + b] = dn;
                    }
                }
            }
        }
        // the step's first work items: their descriptors landed long ago; fetch param + state rows
        // now (nothing in phase A writes them), so only the staged gradients wait for the barrier
        cp_async_wait<0>();
        if (!(dbg & 4)) {
            for (int n = 0; n < RING; ++n) issue_state(n_items, n, lo);
        }
        cp_async_commit();

        hsum = warp_sum(hsum);
        if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = hsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float H = 0.f;
#pragma unroll
            for (int w = 0; w < NT / 32; ++w) H += s_loss[w];
            st.loss_part[(size_t)si * gridDim.x + blockIdx.x] = H;
        }
        if (tracing) tr[1] = global_ns();
        if (!(dbg & 8)) grid_barrier(st.barrier, bar_target);
        if (tracing) tr[2] = global_ns();

        // ------------------------------ phase B ------------------------------------------
        if (!(dbg & 4)) {
            for (int n = 0; n < RING; ++n) issue_grad(n_items, n, lo);
        }
        cp_async_commit();
        // rows of the next step's samples -> L2 (coherent: a row updated below is still read right
        // after the next barrier); the ids were loaded a phase ago
        ids0 = nid0;
        ids1 = nid1;
        if (si + 1 < n_steps && !(dbg & 16)) {
            if (gid < Bs2) {
                prefetch_row<V, G, IT>(m.user.emb + (size_t)ids0.u * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb + (size_t)ids0.ip * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb + (size_t)ids0.in * dim, nch, gl);
            }
            if (gid + ngroups < Bs2) {
                prefetch_row<V, G, IT>(m.user.emb + (size_t)ids1.u * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb + (size_t)ids1.ip * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb + (size_t)ids1.in * dim, nch, gl);
            }
            for (int b = gid + 2 * ngroups; b < Bs2; b += ngroups) {
                const int64_t u = ep.user[lo2 + b], ip = ep.pos[lo2 + b], in = ep.neg[lo2 + b];
                prefetch_row<V, G, IT>(m.user.emb + (size_t)u * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb + (size_t)ip * dim, nch, gl);
                prefetch_row<V, G, IT>(m.item.emb + (size_t)in * dim, nch, gl);
            }
        }
        if (tracing) tr[4] = global_ns();
        // width-1 companions (biases / first-order weights): one THREAD per short segment; the
        // loads are issued here and consumed after the chunk work below
        float lin_p = 0.f, lin_0 = 0.f, lin_1 = 0.f, lin_g = 0.f;
        bool lin_live = false;
        SpaceRef lin_sp = {};
        if (!(dbg & 32) && lin_i < n_items) {
            lin_sp = resolve_space<NET>((int)(lin_it.x & 0xffu), m, plan, st, lo);
            lin_live = lin_sp.stage_lin != nullptr;
            if (lin_live) {
                const trs_table& t = *lin_sp.t;
                lin_p = __ldcg(t.lin + lin_it.z);
                if (kind != TRS_OPT_SGD) lin_0 = __ldcg(t.lin_s0 + lin_it.z);
                if (kind == TRS_OPT_SPARSE_ADAM) lin_1 = __ldcg(t.lin_s1 + lin_it.z);
                lin_g = __ldcg(lin_sp.stage_lin + lin_it.w);
            }
        }
        if (tracing) tr[5] = global_ns();
        // chunks of long segments, spread over the CTAs and over the warps of a CTA
        if (!(dbg & 2)) {
            const uint4* chunks = plan.chunks + (size_t)s * plan.chunk_cap;
            const int g_in_cta = threadIdx.x / G;
            // chunk ci -> CTA ci % grid, local index lc = ci / grid -> group (lc * 4) mod GPB (+ carry)
            for (int lc = 0; lc * (int)gridDim.x + (int)blockIdx.x < n_chunks; ++lc) {
                const int owner = (GPB >= 4) ? ((lc * 4) % GPB + ((lc * 4) / GPB) % 4) % GPB : lc % GPB;
                if (owner == g_in_cta) {
                    const int ci = lc * gridDim.x + blockIdx.x;
                    chunk_item<NET, V, G, IT>(chunks[ci], ci, m, plan, st, s, lo, dim, nch, gl, opt, scale);
                }
            }
        }
        // finish the width-1 companion items
        if (lin_live) {
            const trs_table& t = *lin_sp.t;
            const int c = (int)((lin_it.x >> 8) & 0xffu);
            for (int q = 1; q < c; ++q) lin_g = __fadd_rn(lin_g, __ldcg(lin_sp.stage_lin + lin_sp.P[lin_it.y + q]));
            opt_update(opt, scale, lin_g, lin_p, lin_0, lin_1);
            t.lin[lin_it.z] = lin_p;
            if (kind != TRS_OPT_SGD) t.lin_s0[lin_it.z] = lin_0;
            if (kind == TRS_OPT_SPARSE_ADAM) t.lin_s1[lin_it.z] = lin_1;
        }
        if (!(dbg & 32)) {  // more short segments than threads: the rest, one at a time
            for (int i = lin_i + NT * (int)gridDim.x; i < n_items; i += NT * (int)gridDim.x) {
                const uint4 it = items[i];
                const SpaceRef sp = resolve_space<NET>((int)(it.x & 0xffu), m, plan, st, lo);
                if (!sp.stage_lin) continue;
                const trs_table& t = *sp.t;
                const uint32_t key = it.z;
                float pl = __ldcg(t.lin + key), l0 = 0.f, l1 = 0.f;
                if (kind != TRS_OPT_SGD) l0 = __ldcg(t.lin_s0 + key);
                if (kind == TRS_OPT_SPARSE_ADAM) l1 = __ldcg(t.lin_s1 + key);
                float g_lin = __ldcg(sp.stage_lin + it.w);
                const int c = (int)((it.x >> 8) & 0xffu);
                for (int q = 1; q < c; ++q) g_lin = __fadd_rn(g_lin, __ldcg(sp.stage_lin + sp.P[it.y + q]));
                opt_update(opt, scale, g_lin, pl, l0, l1);
                t.lin[key] = pl;
                if (kind != TRS_OPT_SGD) t.lin_s0[key] = l0;
                if (kind == TRS_OPT_SPARSE_ADAM) t.lin_s1[key] = l1;
            }
        }
        if (tracing) tr[6] = global_ns();
        // short segments through the ring
        if (!(dbg & 4)) {
            cp_async_wait<0>();
            const int my_items = n_items > gid ? (n_items - gid + ngroups - 1) / ngroups : 0;
            for (int n = 0; n < my_items; ++n) {
                if (n >= RING) cp_async_wait<RING - 1>();
                consume(items, n_items, n, lo, scale);
                issue_state(n_items, n + RING, lo);
                issue_grad(n_items, n + RING, lo);
                issue_desc(items, n_items, n + 2 * RING);
                cp_async_commit();
            }
            cp_async_wait<0>();
        }
        if (tracing) tr[7] = global_ns();
        // descriptors of the next step's first work items
        if (si + 1 < n_steps) {
            const uint4* items2 = plan.items + (size_t)(s + 1) * plan.item_cap;
            for (int n = 0; n < 2 * RING; ++n) issue_desc(items2, n_items_next, n);
        }
        cp_async_commit();
        n_items_cur = n_items_next;
        n_chunks_cur = n_chunks_next;
        if (tracing) tr[3] = global_ns();
        if (!(dbg & 8)) grid_barrier(st.barrier, bar_target);
    }
    cp_async_wait<0>();

    // batch-mean hinge per step, summed over CTAs in a fixed order (deterministic)
    if (NET != TRS_NET_MLP && blockIdx.x == 0) {
        for (int si = threadIdx.x; si < n_steps; si += NT) {
            const int64_t lo = (first_step + (int64_t)si) * ep.batch;
            const int Bs = (int)min((int64_t)ep.batch, ep.n_samples - lo);
            float H = 0.f;
            for (unsigned c = 0; c < gridDim.x; ++c) H += __ldcg(st.loss_part + (size_t)si * gridDim.x + c);
            loss_out[si] = H / (float)Bs;
        }
    }
}

// one CTA per SM: the ring takes most of the SM's shared memory
template <int NET, int V, int G, int IT>
static int train_grid_size() {
    return device_props().sm_count;
}

template <int V, int G, int IT>
static void query_grid(int net, int* grid) {
    *grid = train_grid_size<TRS_NET_LINEAR, V, G, IT>();
    (void)net;
}

template <int V, int G, int IT>
static void launch_train(const trs_model* m, const trs_epoch* ep, const OptScalars* opt,
                         const PlanPtrs* plan, const Stage* st, int first_step, int n_steps,
                         float* loss, int grid, cudaStream_t stream, cudaError_t* err) {
    // TRS_DEBUG_SKIP (timing experiments only, results are wrong): 1 phase A, 2 long-segment chunks,
    // 4 phase B, 8 grid barriers, 16 L2 prefetches
    const char* dbg_env = getenv("TRS_DEBUG_SKIP");
    int dbg = dbg_env ? atoi(dbg_env) : 0;
    void* args[] = {(void*)m, (void*)ep, (void*)opt, (void*)plan, (void*)st,
                    (void*)&first_step, (void*)&n_steps, (void*)&loss, (void*)&dbg};
    const void* fn = m->net == TRS_NET_LINEAR ? (const void*)train_kernel<TRS_NET_LINEAR, V, G, IT>
                     : m->net == TRS_NET_FM   ? (const void*)train_kernel<TRS_NET_FM, V, G, IT>
                                              : (const void*)train_kernel<TRS_NET_MLP, V, G, IT>;
    const size_t smem = train_smem_bytes<V, IT>();
    if (smem > 48 * 1024) {
        *err = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (*err != cudaSuccess) return;
    }
    *err = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(train_threads<V, IT>()), args, smem, stream);
}

static int train_grid_for(const trs_model* m, const RowShape& shape) {
    int grid = 0;
    TRS_DISPATCH_ROW_SHAPE(shape, query_grid, m->net, &grid);
    return grid;
}

}  // namespace trs

using namespace trs;

extern "C" int trs_device_info(int* sm_count_host, int* train_grid_host, int* train_block_host) {
    if (sm_count_host) *sm_count_host = device_props().sm_count;
    if (train_block_host) *train_block_host = train_threads<4, 1>();
    if (train_grid_host) {
        trs_model m = {};
        m.net = TRS_NET_FM;
        m.dim = 64;
        RowShape shape;
        pick_row_shape(m.dim, &shape);
        *train_grid_host = train_grid_for(&m, shape);
    }
    TRS_CUDA(cudaGetLastError());
    return TRS_OK;
}

// debug hook (not part of trs.h): byte offset of the phase trace inside the train workspace
extern "C" size_t trs_debug_trace_offset(const trs_model* model, const trs_epoch* epoch) {
    RowShape shape;
    if (!model || !epoch || epoch->batch <= 0 || check_model(model, &shape)) return 0;
    return stage_layout(model, epoch, train_grid_for(model, shape)).trace;
}

extern "C" size_t trs_train_workspace_bytes(const trs_model* model, const trs_epoch* epoch) {
    RowShape shape;
    if (!model || !epoch || epoch->batch <= 0 || check_model(model, &shape)) return 0;
    return stage_layout(model, epoch, train_grid_for(model, shape)).total;
}

namespace trs {
StagePtrs stage_pointers(const trs_model* model, const trs_epoch* ep, void* workspace) {
    RowShape shape;
    StagePtrs r = {};
    if (!pick_row_shape(model->dim, &shape)) return r;
    const StageLayout SL = stage_layout(model, ep, train_grid_for(model, shape));
    char* W = (char*)workspace;
    r.gU = (float*)(W + SL.gU);
    r.gI = (float*)(W + SL.gI);
    for (int f = 0; f < model->n_meta; ++f) r.gM[f] = (float*)(W + SL.gM[f]);
    return r;
}
}  // namespace trs

extern "C" int trs_train_steps(const trs_model* model, const trs_epoch* ep, const trs_optim* optim,
                               const void* plan, void* workspace, size_t workspace_bytes,
                               int first_step, int n_steps, float* loss, trs_stream_t stream) {
    TRS_REQUIRE(model && model->net != TRS_NET_MLP, "trs_train_steps: use trs_mlp_train_steps for net_type mlp");
    return trs::run_train_steps(model, ep, optim, plan, workspace, workspace_bytes, first_step, n_steps, loss,
                                (cudaStream_t)stream);
}

int trs::run_train_steps(const trs_model* model, const trs_epoch* ep, const trs_optim* optim,
                         const void* plan, void* workspace, size_t workspace_bytes,
                         int first_step, int n_steps, float* loss, cudaStream_t stream) {
    RowShape shape;
    int rc = check_model(model, &shape);
    if (rc) return rc;
    TRS_REQUIRE(ep && ep->user && ep->pos && ep->neg, "epoch ids are NULL");
    TRS_REQUIRE(ep->batch > 0, "batch must be positive");
    TRS_REQUIRE(model->n_meta == 0 || (ep->pos_meta && ep->neg_meta), "metadata ids are NULL");
    TRS_REQUIRE(optim && optim->step_scale, "optimizer / step_scale is NULL");
    TRS_REQUIRE(optim->kind >= TRS_OPT_SGD && optim->kind <= TRS_OPT_SPARSE_ADAM, "unknown optimizer kind %d", optim->kind);
    TRS_REQUIRE(plan && workspace && (loss || model->net == TRS_NET_MLP), "plan / workspace / loss is NULL");
    const int64_t steps = n_steps_of(ep);
    TRS_REQUIRE(first_step >= 0 && n_steps >= 0 && first_step + (int64_t)n_steps <= steps,
                "steps [%d, %d) outside the epoch's %lld steps", first_step, first_step + n_steps, (long long)steps);
    if (n_steps == 0) return TRS_OK;

    auto need_state = [&](const trs_table& t, const char* name) -> int {
        const bool lin_learns = t.lin && !(model->net == TRS_NET_LINEAR && &t == &model->user);
        if (optim->kind != TRS_OPT_SGD) {
            TRS_REQUIRE(t.emb_s0 && (!lin_learns || t.lin_s0), "%s: optimizer state s0 is NULL", name);
        }
        if (optim->kind == TRS_OPT_SPARSE_ADAM) {
            TRS_REQUIRE(t.emb_s1 && (!lin_learns || t.lin_s1), "%s: optimizer state s1 is NULL", name);
        }
        return TRS_OK;
    };
    if ((rc = need_state(model->user, "user"))) return rc;
    if ((rc = need_state(model->item, "item"))) return rc;
    for (int f = 0; f < model->n_meta; ++f)
        if ((rc = need_state(model->meta[f], "metadata"))) return rc;

    const int grid = train_grid_for(model, shape);
    TRS_REQUIRE(grid > 0, "no launch configuration for n_factors %d", model->dim);
    const StageLayout SL = stage_layout(model, ep, grid);
    if (workspace_bytes < SL.total) {
        set_error("train workspace too small: %zu < %zu", workspace_bytes, SL.total);
        return TRS_ERR_WORKSPACE;
    }
    const PlanLayout PL = plan_layout(ep->n_samples, ep->batch, model->n_meta);
    char* W = (char*)workspace;
    Stage st = {};
    st.gU = (float*)(W + SL.gU);
    st.gI = (float*)(W + SL.gI);
    for (int f = 0; f < model->n_meta; ++f) st.gM[f] = (float*)(W + SL.gM[f]);
    st.gbU = (float*)(W + SL.gbU);
    st.gbI = (float*)(W + SL.gbI);
    st.loss_part = (float*)(W + SL.loss_part);
    st.partials = (float*)(W + SL.partials);
    st.partials_lin = (float*)(W + SL.partials_lin);
    st.seg_arrive = (unsigned*)(W + SL.sync_words);
    st.barrier = st.seg_arrive + PL.long_cap + 32;
    st.trace = (unsigned long long*)(W + SL.trace);
    TRS_CUDA(cudaMemsetAsync(W + SL.sync_words, 0, SL.sync_bytes, stream));

    const char* P = (const char*)plan;
    PlanPtrs pp = {};
    pp.user_key = (const uint32_t*)(P + PL.user_key);
    pp.user_perm = (const uint32_t*)(P + PL.user_perm);
    pp.item_key = (const uint32_t*)(P + PL.item_key);
    pp.item_perm = (const uint32_t*)(P + PL.item_perm);
    for (int f = 0; f < model->n_meta; ++f) {
        pp.meta_key[f] = (const uint32_t*)(P + PL.meta_key[f]);
        pp.meta_perm[f] = (const uint32_t*)(P + PL.meta_perm[f]);
    }
    pp.item_cnt = (const uint32_t*)(P + PL.item_cnt);
    pp.chunk_cnt = (const uint32_t*)(P + PL.chunk_cnt);
    pp.chunks = (const uint4*)(P + PL.chunks);
    pp.items = (const uint4*)(P + PL.items);
    pp.long_segs = (const uint4*)(P + PL.long_segs);
    pp.item_cap = PL.item_cap;
    pp.long_cap = PL.long_cap;
    pp.chunk_cap = PL.chunk_cap;

    OptScalars os;
    os.kind = optim->kind;
    os.omb1 = (float)(1.0 - optim->beta1);
    os.omb2 = (float)(1.0 - optim->beta2);
    os.eps = (float)optim->eps;
    os.step_scale = optim->step_scale;

    cudaError_t err = cudaSuccess;
    TRS_DISPATCH_ROW_SHAPE(shape, launch_train, model, ep, &os, &pp, &st, first_step, n_steps, loss,
                           grid, stream, &err);
    TRS_CUDA(err);
    return TRS_OK;
}
