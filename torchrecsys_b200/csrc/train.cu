// train.cu -- fused forward + hinge + backward + segmented reduce + row-wise optimizer for the
// Linear and FM scorers, as ONE persistent cooperative kernel over many steps.
//
// Per step (reference model.py:274-284):
//   phase A  one row group per sample: gather u, v+, v- (+ metadata rows), both scores, hinge,
//            closed-form gradient rows -> L2-resident staging, one row per lookup.
//   grid barrier (every score of the step is computed from pre-update parameters)
//   phase B  one row group per touched row: sum the staged rows of its lookups in lookup order
//            (the plan from plan.cu gives the stable sort coalesce() would do), then read
//            param+state once, apply SGD / Adagrad / SparseAdam, write param+state once.
//   grid barrier
// HBM traffic per step is therefore ids + (param+state read, param+state write) per unique
// touched row; staging and plan reads are served from L2.
#include <cooperative_groups.h>

#include "plan.cuh"
#include "scorer.cuh"

namespace cg = cooperative_groups;

namespace trs {

constexpr int TRAIN_THREADS = 256;

struct PlanPtrs {
    const uint32_t *user_key, *user_perm, *item_key, *item_perm;
    const uint32_t* meta_key[TRS_MAX_META];
    const uint32_t* meta_perm[TRS_MAX_META];
};

struct Stage {
    float* gU;                 // [B, dim]
    float* gI;                 // [2B, dim]
    float* gM[TRS_MAX_META];   // FM only: [2B, dim]
    float* gbU;                // FM only: [B]
    float* gbI;                // [2B]  (FM: also the gradient of linear_metadata, d w_k = delta)
    float* loss_part;          // [n_steps, gridDim.x]
};

struct OptScalars {
    int kind;
    float omb1, omb2, eps;  // 1-beta1, 1-beta2, eps rounded to fp32 as torch's scalar ops do
    const float* step_scale;
};

struct StageLayout {
    size_t gU, gI, gM[TRS_MAX_META], gbU, gbI, loss_part, total;
};

static StageLayout stage_layout(const trs_model* m, const trs_epoch* ep, int grid) {
    StageLayout L;
    size_t off = 0;
    auto take = [&](size_t n_floats) {
        size_t o = off;
        off += (n_floats * sizeof(float) + 255) / 256 * 256;
        return o;
    };
    const size_t B = (size_t)ep->batch, D = (size_t)m->dim;
    L.gU = take(B * D);
    L.gI = take(2 * B * D);
    for (int f = 0; f < TRS_MAX_META; ++f) L.gM[f] = (m->net == TRS_NET_FM && f < m->n_meta) ? take(2 * B * D) : 0;
    L.gbU = take(B);
    L.gbI = take(2 * B);
    L.loss_part = take((size_t)n_steps_of(ep) * grid);
    L.total = off;
    return L;
}

// ---- row-wise optimizers (torch: optim/_functional.py:65-84, optim/adagrad.py:363-373, sgd) ----
// Explicit _rn intrinsics keep nvcc from contracting mul+add into FMA where torch runs two ops.
__device__ __forceinline__ void opt_update(const OptScalars& o, float scale, float g, float& p,
                                           float& s0, float& s1) {
    if (o.kind == TRS_OPT_SPARSE_ADAM) {
        const float um = __fmul_rn(__fsub_rn(g, s0), o.omb1);
        const float uv = __fmul_rn(__fsub_rn(__fmul_rn(g, g), s1), o.omb2);
        s0 = __fadd_rn(s0, um);
        s1 = __fadd_rn(s1, uv);
        const float denom = __fadd_rn(__fsqrt_rn(s1), o.eps);
        p = __fadd_rn(p, __fmul_rn(-scale, __fdiv_rn(s0, denom)));
    } else if (o.kind == TRS_OPT_ADAGRAD) {
        s0 = __fadd_rn(s0, __fmul_rn(g, g));
        const float stdv = __fadd_rn(__fsqrt_rn(s0), o.eps);
        p = __fadd_rn(p, __fmul_rn(-scale, __fdiv_rn(g, stdv)));
    } else {
        p = __fadd_rn(p, __fmul_rn(-scale, g));
    }
}

// One touched row: segment [k, k+c) of the sorted lookups of an id space.
template <int V, int G, int IT>
__device__ __forceinline__ void reduce_and_update(const trs_table& t, int dim, int nch, int gl,
                                                  const uint32_t* __restrict__ K,
                                                  const uint32_t* __restrict__ P, int len, int k,
                                                  const float* __restrict__ stage,
                                                  const float* __restrict__ stage_lin,
                                                  const OptScalars& o, float scale) {
    const uint32_t key = K[k];
    if (k > 0 && K[k - 1] == key) return;  // not the head of its segment
    int c = 1;
    while (k + c < len && K[k + c] == key) ++c;

    Row<V, IT> acc;
#pragma unroll
    for (int i = 0; i < IT; ++i) acc.c[i] = Vec<V>::zero();
    float accl = 0.f;
    for (int i = 0; i < c; ++i) {
        const uint32_t j = P[k + i];
        Row<V, IT> r = load_row<V, G, IT>(stage + (size_t)j * dim, nch, gl);
#pragma unroll
        for (int a = 0; a < IT; ++a)
#pragma unroll
            for (int b = 0; b < V; ++b) acc.c[a][b] = __fadd_rn(acc.c[a][b], r.c[a][b]);
        if (stage_lin) accl = __fadd_rn(accl, stage_lin[j]);
    }

    const size_t roff = (size_t)key * dim;
    Row<V, IT> p = load_row<V, G, IT>(t.emb + roff, nch, gl);
    Row<V, IT> s0, s1;
#pragma unroll
    for (int i = 0; i < IT; ++i) s0.c[i] = s1.c[i] = Vec<V>::zero();
    if (o.kind != TRS_OPT_SGD) s0 = load_row<V, G, IT>(t.emb_s0 + roff, nch, gl);
    if (o.kind == TRS_OPT_SPARSE_ADAM) s1 = load_row<V, G, IT>(t.emb_s1 + roff, nch, gl);
#pragma unroll
    for (int a = 0; a < IT; ++a)
#pragma unroll
        for (int b = 0; b < V; ++b) opt_update(o, scale, acc.c[a][b], p.c[a][b], s0.c[a][b], s1.c[a][b]);
    store_row<V, G, IT>(t.emb + roff, nch, gl, p);
    if (o.kind != TRS_OPT_SGD) store_row<V, G, IT>(t.emb_s0 + roff, nch, gl, s0);
    if (o.kind == TRS_OPT_SPARSE_ADAM) store_row<V, G, IT>(t.emb_s1 + roff, nch, gl, s1);

    if (stage_lin && t.lin && gl == 0) {
        float pl = t.lin[key], l0 = 0.f, l1 = 0.f;
        if (o.kind != TRS_OPT_SGD) l0 = t.lin_s0[key];
        if (o.kind == TRS_OPT_SPARSE_ADAM) l1 = t.lin_s1[key];
        opt_update(o, scale, accl, pl, l0, l1);
        t.lin[key] = pl;
        if (o.kind != TRS_OPT_SGD) t.lin_s0[key] = l0;
        if (o.kind == TRS_OPT_SPARSE_ADAM) t.lin_s1[key] = l1;
    }
}

template <int V, int IT>
__device__ __forceinline__ Row<V, IT> row_axpby(float a, const Row<V, IT>& x, float b,
                                                const Row<V, IT>& y) {  // a*x + b*y
    Row<V, IT> r;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) r.c[i][k] = a * x.c[i][k] + b * y.c[i][k];
    return r;
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

template <int NET, int V, int G, int IT>
__global__ void __launch_bounds__(TRAIN_THREADS)
train_kernel(const trs_model m, const trs_epoch ep, const OptScalars opt, const PlanPtrs plan,
             const Stage st, const int first_step, const int n_steps, float* __restrict__ loss_out) {
    cg::grid_group grid = cg::this_grid();
    __shared__ float s_loss[TRAIN_THREADS / 32];

    const int dim = m.dim, nch = dim / V, F = m.n_meta;
    const int gl = threadIdx.x % G;
    constexpr int GPW = 32 / G;                        // groups per warp
    const int gpb = TRAIN_THREADS / G;                 // groups per block
    const int gid = blockIdx.x * gpb + threadIdx.x / G;
    const int ngroups = gridDim.x * gpb;
    const int gid_warp0 = gid - (gid % GPW);           // first group of my warp

    for (int si = 0; si < n_steps; ++si) {
        const int64_t s = first_step + si;
        const int64_t lo = s * (int64_t)ep.batch;
        const int Bs = (int)min((int64_t)ep.batch, ep.n_samples - lo);
        const float invB = 1.0f / (float)Bs;

        // ------------------------------ phase A ------------------------------------------
        float hsum = 0.f;
        for (int b0 = gid_warp0; b0 < Bs; b0 += ngroups) {   // warp-uniform trip count
            const int b_raw = b0 + (gid - gid_warp0);
            const bool valid = b_raw < Bs;
            const int b = valid ? b_raw : Bs - 1;
            const int64_t smp = lo + b;
            const int64_t u = ep.user[smp], ip = ep.pos[smp], in = ep.neg[smp];
            const int64_t* pm = F ? ep.pos_meta + smp * F : nullptr;
            const int64_t* nm = F ? ep.neg_meta + smp * F : nullptr;
            const Row<V, IT> ru = load_row<V, G, IT>(m.user.emb + (size_t)u * dim, nch, gl);
            const float bu = m.user.lin ? m.user.lin[u] : 0.f;
            if (NET == TRS_NET_LINEAR) {
                Row<V, IT> vp, vn;
                const float sp = linear_score<V, G, IT>(m, nch, gl, ru, bu, ip, pm, vp);
                const float sn = linear_score<V, G, IT>(m, nch, gl, ru, bu, in, nm, vn);
                const float h = __fadd_rn(__fsub_rn(sn, sp), 1.0f);
                const float g = (h >= 0.f) ? invB : 0.f;
                if (valid) {
                    if (gl == 0) hsum += fmaxf(h, 0.f);
                    store_row<V, G, IT>(st.gU + (size_t)b * dim, nch, gl, row_scaled_diff(g, vn, vp));
                    store_row<V, G, IT>(st.gI + (size_t)b * dim, nch, gl, row_axpby(-g, ru, 0.f, ru));
                    store_row<V, G, IT>(st.gI + (size_t)(Bs + b) * dim, nch, gl, row_axpby(g, ru, 0.f, ru));
                    if (gl == 0) {
                        st.gbI[b] = -g;
                        st.gbI[Bs + b] = g;
                    }
                }
            } else {
                Row<V, IT> Sp, Sn, rp, rn;
                const float sp = sigmoidf_acc(fm_logit<V, G, IT>(m, nch, gl, ru, bu, ip, pm, Sp, rp));
                const float sn = sigmoidf_acc(fm_logit<V, G, IT>(m, nch, gl, ru, bu, in, nm, Sn, rn));
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
                            gu.c[i][k] = dp * (Sp.c[i][k] - ru.c[i][k]) + dn * (Sn.c[i][k] - ru.c[i][k]);
                    store_row<V, G, IT>(st.gU + (size_t)b * dim, nch, gl, gu);
                    store_row<V, G, IT>(st.gI + (size_t)b * dim, nch, gl, row_scaled_diff(dp, Sp, rp));
                    store_row<V, G, IT>(st.gI + (size_t)(Bs + b) * dim, nch, gl, row_scaled_diff(dn, Sn, rn));
                    for (int f = 0; f < F; ++f) {
                        Row<V, IT> r = load_row<V, G, IT>(m.meta[f].emb + (size_t)pm[f] * dim, nch, gl);
                        store_row<V, G, IT>(st.gM[f] + (size_t)b * dim, nch, gl, row_scaled_diff(dp, Sp, r));
                        r = load_row<V, G, IT>(m.meta[f].emb + (size_t)nm[f] * dim, nch, gl);
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
        hsum = warp_sum(hsum);
        if ((threadIdx.x & 31) == 0) s_loss[threadIdx.x >> 5] = hsum;
        __syncthreads();
        if (threadIdx.x == 0) {
            float H = 0.f;
#pragma unroll
            for (int w = 0; w < TRAIN_THREADS / 32; ++w) H += s_loss[w];
            st.loss_part[(size_t)si * gridDim.x + blockIdx.x] = H;
        }
        grid.sync();

        // ------------------------------ phase B ------------------------------------------
        const float scale = opt.step_scale[s];
        const int nU = Bs, nI = 2 * Bs;
        const int total = nU + nI + F * nI;
        for (int w = gid; w < total; w += ngroups) {
            if (w < nU) {
                reduce_and_update<V, G, IT>(m.user, dim, nch, gl, plan.user_key + lo, plan.user_perm + lo,
                                            nU, w, st.gU, NET == TRS_NET_FM ? st.gbU : nullptr, opt, scale);
            } else if (w < nU + nI) {
                reduce_and_update<V, G, IT>(m.item, dim, nch, gl, plan.item_key + 2 * lo,
                                            plan.item_perm + 2 * lo, nI, w - nU, st.gI, st.gbI, opt, scale);
            } else {
                const int f = (w - nU - nI) / nI;
                const int k = (w - nU - nI) - f * nI;
                reduce_and_update<V, G, IT>(m.meta[f], dim, nch, gl, plan.meta_key[f] + 2 * lo,
                                            plan.meta_perm[f] + 2 * lo, nI, k,
                                            NET == TRS_NET_FM ? st.gM[f] : st.gI,
                                            NET == TRS_NET_FM ? st.gbI : nullptr, opt, scale);
            }
        }
        grid.sync();
    }

    // batch-mean hinge per step, summed over CTAs in a fixed order (deterministic)
    if (blockIdx.x == 0) {
        for (int si = threadIdx.x; si < n_steps; si += TRAIN_THREADS) {
            const int64_t lo = (first_step + (int64_t)si) * ep.batch;
            const int Bs = (int)min((int64_t)ep.batch, ep.n_samples - lo);
            float H = 0.f;
            for (unsigned c = 0; c < gridDim.x; ++c) H += st.loss_part[(size_t)si * gridDim.x + c];
            loss_out[si] = H / (float)Bs;
        }
    }
}

template <int NET, int V, int G, int IT>
static int train_grid_size() {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, train_kernel<NET, V, G, IT>, TRAIN_THREADS, 0);
    if (occ < 1) occ = 1;
    if (occ > 4) occ = 4;
    return occ * device_props().sm_count;
}

template <int V, int G, int IT>
static void query_grid(int net, int* grid) {
    *grid = net == TRS_NET_LINEAR ? train_grid_size<TRS_NET_LINEAR, V, G, IT>()
                                  : train_grid_size<TRS_NET_FM, V, G, IT>();
}

template <int V, int G, int IT>
static void launch_train(const trs_model* m, const trs_epoch* ep, const OptScalars* opt,
                         const PlanPtrs* plan, const Stage* st, int first_step, int n_steps,
                         float* loss, int grid, cudaStream_t stream, cudaError_t* err) {
    void* args[] = {(void*)m, (void*)ep, (void*)opt, (void*)plan, (void*)st,
                    (void*)&first_step, (void*)&n_steps, (void*)&loss};
    const void* fn = m->net == TRS_NET_LINEAR ? (const void*)train_kernel<TRS_NET_LINEAR, V, G, IT>
                                              : (const void*)train_kernel<TRS_NET_FM, V, G, IT>;
    *err = cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(TRAIN_THREADS), args, 0, stream);
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
    if (train_block_host) *train_block_host = TRAIN_THREADS;
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

extern "C" size_t trs_train_workspace_bytes(const trs_model* model, const trs_epoch* epoch) {
    RowShape shape;
    if (!model || !epoch || epoch->batch <= 0 || check_model(model, &shape)) return 0;
    return stage_layout(model, epoch, train_grid_for(model, shape)).total;
}

extern "C" int trs_train_steps(const trs_model* model, const trs_epoch* ep, const trs_optim* optim,
                               const void* plan, void* workspace, size_t workspace_bytes,
                               int first_step, int n_steps, float* loss, trs_stream_t stream) {
    RowShape shape;
    int rc = check_model(model, &shape);
    if (rc) return rc;
    TRS_REQUIRE(ep && ep->user && ep->pos && ep->neg, "epoch ids are NULL");
    TRS_REQUIRE(ep->batch > 0, "batch must be positive");
    TRS_REQUIRE(model->n_meta == 0 || (ep->pos_meta && ep->neg_meta), "metadata ids are NULL");
    TRS_REQUIRE(optim && optim->step_scale, "optimizer / step_scale is NULL");
    TRS_REQUIRE(optim->kind >= TRS_OPT_SGD && optim->kind <= TRS_OPT_SPARSE_ADAM, "unknown optimizer kind %d", optim->kind);
    TRS_REQUIRE(plan && workspace && loss, "plan / workspace / loss is NULL");
    const int64_t steps = n_steps_of(ep);
    TRS_REQUIRE(first_step >= 0 && n_steps >= 0 && first_step + (int64_t)n_steps <= steps,
                "steps [%d, %d) outside the epoch's %lld steps", first_step, first_step + n_steps, (long long)steps);
    if (n_steps == 0) return TRS_OK;

    auto need_state = [&](const trs_table& t, const char* name) -> int {
        if (optim->kind != TRS_OPT_SGD) {
            TRS_REQUIRE(t.emb_s0 && (!t.lin || t.lin_s0 || (model->net == TRS_NET_LINEAR && &t == &model->user)),
                        "%s: optimizer state s0 is NULL", name);
        }
        if (optim->kind == TRS_OPT_SPARSE_ADAM) {
            TRS_REQUIRE(t.emb_s1 && (!t.lin || t.lin_s1 || (model->net == TRS_NET_LINEAR && &t == &model->user)),
                        "%s: optimizer state s1 is NULL", name);
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
    char* W = (char*)workspace;
    Stage st = {};
    st.gU = (float*)(W + SL.gU);
    st.gI = (float*)(W + SL.gI);
    for (int f = 0; f < model->n_meta; ++f) st.gM[f] = (float*)(W + SL.gM[f]);
    st.gbU = (float*)(W + SL.gbU);
    st.gbI = (float*)(W + SL.gbI);
    st.loss_part = (float*)(W + SL.loss_part);

    const PlanLayout PL = plan_layout(ep->n_samples, model->n_meta);
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
