// scorer.cuh -- per-sample forward of the Linear and FM scorers, one row group per sample.
// Reference: collaborative/linear.py:54-80, collaborative/fm.py:60-101.
#pragma once
#include "common.cuh"

namespace trs {

template <int V, int IT>
__device__ __forceinline__ void row_add(Row<V, IT>& a, const Row<V, IT>& b) {
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) a.c[i][k] += b.c[i][k];
}

template <int V, int IT>
__device__ __forceinline__ float row_dot_partial(const Row<V, IT>& a, const Row<V, IT>& b) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) s = fmaf(a.c[i][k], b.c[i][k], s);
    return s;
}

__device__ __forceinline__ float sigmoidf_acc(float z) { return 1.0f / (1.0f + expf(-z)); }

// Linear: s = <u, v + sum_f m_f> + b_u + b_i.   v_out = the pooled item vector.
template <int V, int G, int IT>
__device__ __forceinline__ float linear_score(const trs_model& m, int nch, int gl,
                                              const Row<V, IT>& ru, float bu, int64_t item,
                                              const int64_t* __restrict__ meta,
                                              Row<V, IT>& v_out) {
    const size_t dim = (size_t)m.dim;
    Row<V, IT> v = load_row<V, G, IT>(m.item.emb + (size_t)item * dim, nch, gl);
    for (int f = 0; f < m.n_meta; ++f) {
        Row<V, IT> r = load_row<V, G, IT>(m.meta[f].emb + (size_t)meta[f] * dim, nch, gl);
        row_add(v, r);
    }
    float dot = group_sum<G>(row_dot_partial(ru, v));
    float bi = m.item.lin ? m.item.lin[item] : 0.f;
    v_out = v;
    return (dot + bu) + bi;
}

// FM: z = sum_k w_k + 0.5 * sum_d [(sum_k e_kd)^2 - sum_k e_kd^2];  fields k = user, item, meta_f.
// S_out = sum_k e_k, ri_out = the item row.
template <int V, int G, int IT>
__device__ __forceinline__ float fm_logit(const trs_model& m, int nch, int gl,
                                          const Row<V, IT>& ru, float wu, int64_t item,
                                          const int64_t* __restrict__ meta, Row<V, IT>& S_out,
                                          Row<V, IT>& ri_out) {
    const size_t dim = (size_t)m.dim;
    Row<V, IT> ri = load_row<V, G, IT>(m.item.emb + (size_t)item * dim, nch, gl);
    Row<V, IT> S, Q;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) {
            float a = ru.c[i][k], b = ri.c[i][k];
            S.c[i][k] = a + b;
            Q.c[i][k] = a * a + b * b;
        }
    float lin = wu + (m.item.lin ? m.item.lin[item] : 0.f);
    for (int f = 0; f < m.n_meta; ++f) {
        Row<V, IT> r = load_row<V, G, IT>(m.meta[f].emb + (size_t)meta[f] * dim, nch, gl);
#pragma unroll
        for (int i = 0; i < IT; ++i)
#pragma unroll
            for (int k = 0; k < V; ++k) {
                S.c[i][k] += r.c[i][k];
                Q.c[i][k] = fmaf(r.c[i][k], r.c[i][k], Q.c[i][k]);
            }
        lin += m.meta[f].lin ? m.meta[f].lin[meta[f]] : 0.f;
    }
    float p = 0.f;
#pragma unroll
    for (int i = 0; i < IT; ++i)
#pragma unroll
        for (int k = 0; k < V; ++k) p += S.c[i][k] * S.c[i][k] - Q.c[i][k];
    float pair = group_sum<G>(p) * 0.5f;
    S_out = S;
    ri_out = ri;
    return lin + pair;
}

// one (user, item[, meta]) score exactly as trs_scores computes it (Linear: raw score; FM: sigmoid)
template <int NET, int V, int G, int IT>
__device__ __forceinline__ float score_one(const trs_model& m, int nch, int gl, int64_t u,
                                           int64_t it, const int64_t* meta) {
    Row<V, IT> ru = load_row<V, G, IT>(m.user.emb + (size_t)u * m.dim, nch, gl);
    float bu = m.user.lin ? m.user.lin[u] : 0.f;
    Row<V, IT> a, b;
    if (NET == TRS_NET_LINEAR) return linear_score<V, G, IT>(m, nch, gl, ru, bu, it, meta, a);
    return sigmoidf_acc(fm_logit<V, G, IT>(m, nch, gl, ru, bu, it, meta, a, b));
}

}  // namespace trs
