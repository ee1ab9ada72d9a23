// mlp.cu -- the MLP tower (reference collaborative/mlp.py:88-115) around the tcgen05 GEMM of gemm.cu:
// gather+concat, BatchNorm1d (per-pass batch statistics) + ReLU, the output GEMV + pairwise hinge, the
// backward of all of it, the dense SGD / Adagrad update of the tower and the hand-over of the input
// gradient to the deterministic segmented reduce + row-wise update of train.cu.
//
// One training step (model.py:274-284) stacks the positive and the negative pass into ONE activation
// matrix of R = 2*Bp rows (Bp = batch rounded up to 128; row h*Bp + b = pass h, sample b; padding rows
// are zero) so every layer is one GEMM, while BatchNorm keeps the reference's per-pass statistics
// (batch statistics and running-stat updates separately for the pos and the neg pass).
//
//   forward   X0 = concat[u, v, meta...] (bf16)                                    gather_concat_kernel
//             Z_l = X_l W_l^T + b_l (bf16) + per-128-row-tile column sums           gemm (epilogue)
//             mean/rstd per pass, running stats                                     bn_finalize_kernel
//             X_{l+1} = relu(gamma * (Z_l - mean) * rstd + beta) (bf16)             bn_relu_kernel
//             s = X_n w_out + b_out, hinge, ds                                      out_hinge_kernel
//   backward  per layer, last to first:
//             column sums of dy and dy*xhat (dy = dX_{l+1} masked by relu)          bn_bwd_reduce_kernel
//             per-pass totals, dgamma, dbeta (, dw_out)                             bn_bwd_finalize_kernel
//             dZ_l (bf16) + column sums for db_l                                    bn_bwd_apply_kernel
//             dW_l = dZ_l^T X_l   (split-K over the rows, MN-major operands)        gemm + reduce_partials
//             dX_l = dZ_l W_l     (W_l as stored: MN-major B operand)               gemm
//   update    tower parameters (dense SGD / Adagrad)                                dense_update_kernel
//             embedding rows: dX_0 split per table -> staged rows -> train.cu       stage_grads_kernel
// Every floating-point reduction has a fixed association (no atomics): results are run-to-run identical.
#include "plan.cuh"
#include "tc.cuh"
#include "train.cuh"

namespace trs {

typedef __nv_bfloat16 bf16;
constexpr float BN_EPS = 1e-5f, BN_MOMENTUM = 0.1f;
constexpr int EW_THREADS = 256;

struct F8 {
    float v[8];
};
__device__ __forceinline__ F8 ld_bf8(const bf16* p) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    F8 r;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        r.v[2 * i] = __uint_as_float(w[i] << 16);
        r.v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
    return r;
}
__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ void st_bf8(bf16* p, const F8& x) {
    uint4 u;
    u.x = pack_bf2(x.v[0], x.v[1]);
    u.y = pack_bf2(x.v[2], x.v[3]);
    u.z = pack_bf2(x.v[4], x.v[5]);
    u.w = pack_bf2(x.v[6], x.v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}

// the 8 per-column parameters a thread needs for its 16-byte chunk, as two 128-bit loads
__device__ __forceinline__ F8 ld_f8(const float* p) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    F8 r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
}

// Sum over tiles [t0, t0+nt) of P[t*N + n] for the 32 columns n = blockIdx.x*32 + lane of a 1024-thread
// block: warp w adds tiles w, w+32, ... (independent loads, one memory round trip for <= 128 tiles); the
// 32 partials are then added in warp order (fixed association).  The result is valid in warp 0.
constexpr int TS_WARPS = 32;
__device__ __forceinline__ float block_tile_sum(const float* __restrict__ P, int t0, int nt, int N, int n,
                                                float (*s)[33]) {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float a = 0.f;
    if (n < N) {
#pragma unroll 4
        for (int t = w; t < nt; t += TS_WARPS) a += P[(size_t)(t0 + t) * N + n];
    }
    s[w][lane] = a;
    __syncthreads();
    float r = 0.f;
    if (w == 0) {
#pragma unroll
        for (int i = 0; i < TS_WARPS; ++i) r += s[i][lane];
    }
    __syncthreads();
    return r;
}

// ---- fp32 -> bf16 copies of the layer weights (they change every step) ----------------------------
struct CvtJobs {
    const float* src[TRS_MAX_LAYERS];
    bf16* dst[TRS_MAX_LAYERS];
    long long n[TRS_MAX_LAYERS];
};
__global__ void __launch_bounds__(EW_THREADS) cvt_kernel(const __grid_constant__ CvtJobs j) {
    const int job = blockIdx.y;
    const long long n = j.n[job];
    for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * EW_THREADS)
        j.dst[job][i] = __float2bfloat16_rn(j.src[job][i]);
}

// ---- X0[h*Bp + b, :] = concat[user[u_b], item[i_hb], meta_f[m_hbf] ...] as bf16; padding rows zero ----
__global__ void __launch_bounds__(EW_THREADS)
gather_concat_kernel(const __grid_constant__ trs_model m, const int64_t* __restrict__ user,
                     const int64_t* __restrict__ item0, const int64_t* __restrict__ item1,
                     const int64_t* __restrict__ meta0, const int64_t* __restrict__ meta1, int B, int Bp,
                     int halves, bf16* __restrict__ X) {
    const int D = m.dim, F = m.n_meta, Din = D * (2 + F), cpr = Din / 4;
    const long long total = (long long)halves * Bp * cpr;
    for (long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * EW_THREADS) {
        const int r = (int)(idx / cpr), c = (int)(idx % cpr) * 4;
        const int h = r / Bp, b = r % Bp;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (b < B) {
            const int field = c / D, off = c % D;
            const float* row;
            if (field == 0) row = m.user.emb + (size_t)user[b] * D;
            else if (field == 1) row = m.item.emb + (size_t)(h ? item1 : item0)[b] * D;
            else row = m.meta[field - 2].emb + (size_t)(h ? meta1 : meta0)[(size_t)b * F + field - 2] * D;
            v = *reinterpret_cast<const float4*>(row + off);
        }
        uint2 o;
        o.x = pack_bf2(v.x, v.y);
        o.y = pack_bf2(v.z, v.w);
        *reinterpret_cast<uint2*>(X + (size_t)r * Din + c) = o;
    }
}

// ---- BatchNorm statistics from the GEMM epilogue's per-tile column sums -----------------------------
// One thread per column; passes in order (pos, then neg), so the running statistics take the two
// momentum updates in the reference's order (model.py:173-183 calls net.forward twice).
__global__ void __launch_bounds__(32 * TS_WARPS)
bn_finalize_kernel(const float* __restrict__ psum, const float* __restrict__ psq, int tiles_per_half, int halves,
                   int N, int n_valid, float* __restrict__ mean, float* __restrict__ rstd,
                   float* __restrict__ running_mean, float* __restrict__ running_var) {
    __shared__ float s_red[TS_WARPS][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    for (int h = 0; h < halves; ++h) {
        const double S = block_tile_sum(psum, h * tiles_per_half, tiles_per_half, N, n, s_red);
        const double SS = block_tile_sum(psq, h * tiles_per_half, tiles_per_half, N, n, s_red);
        if (threadIdx.x >= 32 || n >= N) continue;
        const double mu = S / n_valid;
        double var = SS / n_valid - mu * mu;
        var = var > 0.0 ? var : 0.0;
        mean[h * N + n] = (float)mu;
        rstd[h * N + n] = 1.0f / sqrtf((float)var + BN_EPS);
        if (running_mean) {
            const float unbiased = (float)(var * ((double)n_valid / (double)(n_valid > 1 ? n_valid - 1 : 1)));
            running_mean[n] = (1.0f - BN_MOMENTUM) * running_mean[n] + BN_MOMENTUM * (float)mu;
            running_var[n] = (1.0f - BN_MOMENTUM) * running_var[n] + BN_MOMENTUM * unbiased;
        }
    }
}
__global__ void __launch_bounds__(128)
bn_eval_stats_kernel(const float* __restrict__ running_mean, const float* __restrict__ running_var, int N,
                     float* __restrict__ mean, float* __restrict__ rstd) {
    const int n = blockIdx.x * 128 + threadIdx.x;
    if (n >= N) return;
    mean[n] = running_mean[n];
    rstd[n] = 1.0f / sqrtf(running_var[n] + BN_EPS);
}

// ---- X_{l+1} = relu(BN(Z_l)) --------------------------------------------------------------------------
// CTA = 128 rows (one pass: Bp is a multiple of 128) x 64 columns; thread = 8 columns x 4 rows, its column
// parameters held in registers.
__global__ void __launch_bounds__(EW_THREADS)
bn_relu_kernel(const bf16* __restrict__ Z, bf16* __restrict__ A, const float* __restrict__ mean,
               const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
               int R, int N, int Bp, int B, int use_bn) {
    const int c8 = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const int c = blockIdx.y * 64 + c8 * 8;
    if (c >= N) return;
    const int h = (blockIdx.x * 128) / Bp;
    F8 mu, rs, ga, be;
    if (use_bn) {
        mu = ld_f8(mean + h * N + c);
        rs = ld_f8(rstd + h * N + c);
        ga = ld_f8(gamma + c);
        be = ld_f8(beta + c);
    }
    F8 y[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = blockIdx.x * 128 + rg + 32 * i;
        if (r < R && (r % Bp) < B) y[i] = ld_bf8(Z + (size_t)r * N + c);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int r = blockIdx.x * 128 + rg + 32 * i;
        if (r >= R) continue;
        if ((r % Bp) < B) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = y[i].v[j];
                if (use_bn) t = (t - mu.v[j]) * rs.v[j] * ga.v[j] + be.v[j];
                y[i].v[j] = fmaxf(t, 0.f);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) y[i].v[j] = 0.f;
        }
        st_bf8(A + (size_t)r * N + c, y[i]);
    }
}

// ---- output layer + pairwise hinge (mlp.py:113, helper/loss.py:5-9): one warp per sample ----------------
__device__ __forceinline__ float row_dot_bf16(const bf16* __restrict__ a, const float* __restrict__ w, int N, int lane) {
    float s = 0.f;
    for (int c = lane * 8; c < N; c += 256) {
        const F8 x = ld_bf8(a + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(x.v[j], __ldg(w + c + j), s);
    }
    return warp_sum(s);
}
__global__ void __launch_bounds__(EW_THREADS)
out_hinge_kernel(const bf16* __restrict__ A, const float* __restrict__ w, const float* __restrict__ bo, int N,
                 int Bp, int B, float* __restrict__ s, float* __restrict__ ds, float* __restrict__ loss_part) {
    __shared__ float s_h[EW_THREADS / 32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * (EW_THREADS / 32) + warp;
    float hinge = 0.f;
    if (b < B) {
        const float sp = row_dot_bf16(A + (size_t)b * N, w, N, lane) + bo[0];
        const float sn = row_dot_bf16(A + (size_t)(Bp + b) * N, w, N, lane) + bo[0];
        const float h = __fadd_rn(__fsub_rn(sn, sp), 1.0f);
        const float g = (h >= 0.f) ? 1.0f / (float)B : 0.f;
        hinge = fmaxf(h, 0.f);
        if (lane == 0) {
            s[b] = sp;
            s[Bp + b] = sn;
            ds[b] = -g;
            ds[Bp + b] = g;
        }
    }
    if (lane == 0) s_h[warp] = hinge;
    __syncthreads();
    if (threadIdx.x == 0) {
        float H = 0.f;
#pragma unroll
        for (int i = 0; i < EW_THREADS / 32; ++i) H += s_h[i];
        loss_part[blockIdx.x] = H;
    }
}
__global__ void __launch_bounds__(EW_THREADS)
out_scores_kernel(const bf16* __restrict__ A, const float* __restrict__ w, const float* __restrict__ bo, int N,
                  int n_rows, float* __restrict__ s) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x * (EW_THREADS / 32) + warp;
    if (r >= n_rows) return;
    const float v = row_dot_bf16(A + (size_t)r * N, w, N, lane) + bo[0];
    if (lane == 0) s[r] = v;
}
__global__ void __launch_bounds__(32) loss_finalize_kernel(const float* __restrict__ part, int n, int B, float* out) {
    float a = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) a += part[i];
    a = warp_sum(a);
    if (threadIdx.x == 0) out[0] = a / (float)B;
}

// ---- BatchNorm + ReLU backward ---------------------------------------------------------------------------
// Tiling of both kernels: CTA = 128 rows x 64 columns; thread = 8 columns (one 16-byte access) x 4 rows.
struct BwdIn {
    const bf16* dA;      // gradient w.r.t. the layer's output X_{l+1} (null for the last layer ...)
    const float* ds;     // ... where it is ds[r] * w_out[n])
    const float* w_out;
    const bf16* A;       // X_{l+1} = relu(y): the relu mask, and the input of the output GEMV
    const bf16* Z;       // pre-BN activations
    const float* mean;   // [2, N] per pass
    const float* rstd;
    int R, N, Bp, B, use_bn;
};
// column parameters of a thread's 8 columns (pass h): loaded once, before its rows
struct BwdCols {
    F8 mu, rs, wo;
};
template <bool FROM_DS>
__device__ __forceinline__ BwdCols bwd_cols(const BwdIn& in, int c, int h) {
    BwdCols k;
    if (in.use_bn) {
        k.mu = ld_f8(in.mean + h * in.N + c);
        k.rs = ld_f8(in.rstd + h * in.N + c);
    }
    if (FROM_DS) k.wo = ld_f8(in.w_out + c);
    return k;
}
template <bool FROM_DS>
__device__ __forceinline__ void bwd_load(const BwdIn& in, const BwdCols& k, int r, int c, F8& dy, F8& xhat, F8& a,
                                         float& dsr) {
    a = ld_bf8(in.A + (size_t)r * in.N + c);
    F8 da;
    if (FROM_DS) {
        dsr = in.ds[r];
#pragma unroll
        for (int j = 0; j < 8; ++j) da.v[j] = dsr * k.wo.v[j];
    } else {
        da = ld_bf8(in.dA + (size_t)r * in.N + c);
    }
    if (in.use_bn) {
        const F8 z = ld_bf8(in.Z + (size_t)r * in.N + c);
#pragma unroll
        for (int j = 0; j < 8; ++j) xhat.v[j] = (z.v[j] - k.mu.v[j]) * k.rs.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) dy.v[j] = a.v[j] > 0.f ? da.v[j] : 0.f;
}

template <bool FROM_DS>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_reduce_kernel(const __grid_constant__ BwdIn in, float* __restrict__ P1, float* __restrict__ P2,
                     float* __restrict__ P3) {
    __shared__ float s_red[3][32][64 + 1];
    const int c8 = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const int c = blockIdx.y * 64 + c8 * 8;
    const int h = (blockIdx.x * 128) / in.Bp;
    float a1[8], a2[8], a3[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a1[j] = a2[j] = a3[j] = 0.f;
    if (c < in.N) {
        const BwdCols cols = bwd_cols<FROM_DS>(in, c, h);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = blockIdx.x * 128 + rg + 32 * i;
            if (r < in.R && (r % in.Bp) < in.B) {
                F8 dy, xhat, a;
                float dsr = 0.f;
                bwd_load<FROM_DS>(in, cols, r, c, dy, xhat, a, dsr);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    a1[j] += dy.v[j];
                    if (in.use_bn) a2[j] = fmaf(dy.v[j], xhat.v[j], a2[j]);
                    if (FROM_DS) a3[j] = fmaf(dsr, a.v[j], a3[j]);
                }
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        s_red[0][rg][c8 * 8 + j] = a1[j];
        s_red[1][rg][c8 * 8 + j] = a2[j];
        s_red[2][rg][c8 * 8 + j] = a3[j];
    }
    __syncthreads();
    if (threadIdx.x < 192) {
        const int k = threadIdx.x / 64, col = threadIdx.x % 64;
        float* P = k == 0 ? P1 : (k == 1 ? P2 : P3);
        if (P && blockIdx.y * 64 + col < in.N) {
            float t = 0.f;
#pragma unroll
            for (int g = 0; g < 32; ++g) t += s_red[k][g][col];
            P[(size_t)blockIdx.x * in.N + blockIdx.y * 64 + col] = t;
        }
    }
}

__global__ void __launch_bounds__(32 * TS_WARPS)
bn_bwd_finalize_kernel(const float* __restrict__ P1, const float* __restrict__ P2, const float* __restrict__ P3,
                       int tiles_per_half, int N, float* __restrict__ S1, float* __restrict__ S2,
                       float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dw_out) {
    __shared__ float s_red[TS_WARPS][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    float s1[2] = {0.f, 0.f}, s2[2] = {0.f, 0.f}, s3 = 0.f;
    for (int h = 0; h < 2; ++h) {
        s1[h] = block_tile_sum(P1, h * tiles_per_half, tiles_per_half, N, n, s_red);
        if (P2) s2[h] = block_tile_sum(P2, h * tiles_per_half, tiles_per_half, N, n, s_red);
    }
    if (P3) s3 = block_tile_sum(P3, 0, 2 * tiles_per_half, N, n, s_red);
    if (threadIdx.x >= 32 || n >= N) return;
    S1[n] = s1[0];
    S1[N + n] = s1[1];
    S2[n] = s2[0];
    S2[N + n] = s2[1];
    if (dbeta) dbeta[n] = s1[0] + s1[1];
    if (dgamma) dgamma[n] = s2[0] + s2[1];
    if (dw_out) dw_out[n] = s3;
}

template <bool FROM_DS>
__global__ void __launch_bounds__(EW_THREADS)
bn_bwd_apply_kernel(const __grid_constant__ BwdIn in, const float* __restrict__ gamma, const float* __restrict__ S1,
                    const float* __restrict__ S2, bf16* __restrict__ dZ, float* __restrict__ Pdb) {
    __shared__ float s_red[32][64 + 1];
    const int c8 = threadIdx.x & 7, rg = threadIdx.x >> 3;
    const int c = blockIdx.y * 64 + c8 * 8;
    const int h = (blockIdx.x * 128) / in.Bp;
    const float invB = 1.0f / (float)in.B;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    if (c < in.N) {
        const BwdCols cols = bwd_cols<FROM_DS>(in, c, h);
        F8 ga, m1, m2;  // gamma*rstd, mean(dy), mean(dy*xhat) of the thread's columns
        if (in.use_bn) {
            ga = ld_f8(gamma + c);
            m1 = ld_f8(S1 + h * in.N + c);
            m2 = ld_f8(S2 + h * in.N + c);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                ga.v[j] *= cols.rs.v[j];
                m1.v[j] *= invB;
                m2.v[j] *= invB;
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int r = blockIdx.x * 128 + rg + 32 * i;
            if (r >= in.R) continue;
            F8 dz;
            if ((r % in.Bp) < in.B) {
                F8 dy, xhat, a;
                float dsr;
                bwd_load<FROM_DS>(in, cols, r, c, dy, xhat, a, dsr);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = dy.v[j];
                    if (in.use_bn) t = ga.v[j] * (t - m1.v[j] - xhat.v[j] * m2.v[j]);
                    dz.v[j] = t;
                    acc[j] += t;
                }
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) dz.v[j] = 0.f;
            }
            st_bf8(dZ + (size_t)r * in.N + c, dz);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s_red[rg][c8 * 8 + j] = acc[j];
    __syncthreads();
    if (threadIdx.x < 64 && blockIdx.y * 64 + threadIdx.x < in.N) {
        float t = 0.f;
#pragma unroll
        for (int g = 0; g < 32; ++g) t += s_red[g][threadIdx.x];
        Pdb[(size_t)blockIdx.x * in.N + blockIdx.y * 64 + threadIdx.x] = t;
    }
}

// out[n] = sum over the T row tiles of P[t, n] (db)
__global__ void __launch_bounds__(32 * TS_WARPS)
colsum_tiles_kernel(const float* __restrict__ P, int T, int N, float* __restrict__ out) {
    __shared__ float s_red[TS_WARPS][33];
    const int n = blockIdx.x * 32 + (threadIdx.x & 31);
    const float r = block_tile_sum(P, 0, T, N, n, s_red);
    if (threadIdx.x < 32 && n < N) out[n] = r;
}

// out[i] = sum_p part[p*stride + i], p in order (split-K partials of dW; per-tile partials of db)
__global__ void __launch_bounds__(EW_THREADS)
reduce_partials_kernel(const float* __restrict__ part, int nparts, long long stride, long long n, float* __restrict__ out) {
    for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * EW_THREADS) {
        float t = 0.f;
        for (int p = 0; p < nparts; ++p) t += part[(size_t)p * stride + i];
        out[i] = t;
    }
}

// ---- dX_0 -> one gradient row per lookup, in the layout train.cu reduces from ------------------------------
__global__ void __launch_bounds__(EW_THREADS)
stage_grads_kernel(const float* __restrict__ dX, int Din, int D, int B, int Bp, const __grid_constant__ StagePtrs st) {
    const int cpr = Din / 4;
    const long long total = (long long)B * cpr;
    for (long long idx = (long long)blockIdx.x * EW_THREADS + threadIdx.x; idx < total;
         idx += (long long)gridDim.x * EW_THREADS) {
        const int b = (int)(idx / cpr), c = (int)(idx % cpr) * 4;
        const int field = c / D, off = c % D;
        const float4 p = *reinterpret_cast<const float4*>(dX + (size_t)b * Din + c);
        const float4 q = *reinterpret_cast<const float4*>(dX + (size_t)(Bp + b) * Din + c);
        if (field == 0) {
            // the user row is looked up by both passes (model.py:173-183): one staged row = their sum
            *reinterpret_cast<float4*>(st.gU + (size_t)b * D + off) =
                make_float4(__fadd_rn(p.x, q.x), __fadd_rn(p.y, q.y), __fadd_rn(p.z, q.z), __fadd_rn(p.w, q.w));
        } else {
            float* g = field == 1 ? st.gI : st.gM[field - 2];
            *reinterpret_cast<float4*>(g + (size_t)b * D + off) = p;
            *reinterpret_cast<float4*>(g + (size_t)(B + b) * D + off) = q;
        }
    }
}

// ---- dense SGD / Adagrad on the tower (torch: optim/sgd.py, optim/adagrad.py:380-385) ------------------------
constexpr int MAX_DENSE = 4 * TRS_MAX_LAYERS + 2;
struct DenseJobs {
    float* p[MAX_DENSE];
    const float* g[MAX_DENSE];
    float* s0[MAX_DENSE];
    long long n[MAX_DENSE];
};
__global__ void __launch_bounds__(EW_THREADS)
dense_update_kernel(const __grid_constant__ DenseJobs j, int kind, float eps, const float* __restrict__ step_scale,
                    int step) {
    const int job = blockIdx.y;
    const float scale = step_scale[step];
    float* p = j.p[job];
    const float* g = j.g[job];
    float* s0 = j.s0[job];
    for (long long i = (long long)blockIdx.x * EW_THREADS + threadIdx.x; i < j.n[job]; i += (long long)gridDim.x * EW_THREADS) {
        const float gi = g[i];
        if (kind == TRS_OPT_ADAGRAD) {
            const float acc = __fadd_rn(s0[i], __fmul_rn(gi, gi));
            s0[i] = acc;
            p[i] = __fadd_rn(p[i], __fmul_rn(-scale, __fdiv_rn(gi, __fadd_rn(__fsqrt_rn(acc), eps))));
        } else {
            p[i] = __fadd_rn(p[i], __fmul_rn(-scale, gi));
        }
    }
}

// ---- workspace layout ---------------------------------------------------------------------------------------
struct MlpLayout {
    int Din, R, Bp, T;  // T = R / 128 row tiles
    int in[TRS_MAX_LAYERS], out[TRS_MAX_LAYERS], maxN;
    size_t Wb[TRS_MAX_LAYERS], X0, Z[TRS_MAX_LAYERS], A[TRS_MAX_LAYERS], G[TRS_MAX_LAYERS];
    size_t mean[TRS_MAX_LAYERS], rstd[TRS_MAX_LAYERS];
    size_t psum, psq, P1, P2, P3, Pdb, S1, S2, wpart, dX, s, ds, loss_part, train_ws, total;
    int splits[TRS_MAX_LAYERS];
    size_t train_ws_bytes;
};

static int wgrad_splits(int out, int in, int R) {
    const int tiles = ((out + 127) / 128) * ((in + (in > 64 ? 127 : 63)) / (in > 64 ? 128 : 64));
    const int nkb = R / 64;
    int s = 296 / tiles;
    if (s > nkb / 4) s = nkb / 4;
    return s < 1 ? 1 : s;
}

static MlpLayout mlp_layout(const trs_model* m, const trs_mlp* mlp, int64_t rows_per_half, int halves, bool train,
                            const trs_epoch* ep) {
    MlpLayout L = {};
    size_t off = 0;
    auto take = [&](size_t bytes) {
        size_t o = off;
        off += (bytes + 255) / 256 * 256;
        return o;
    };
    L.Din = m->dim * (2 + m->n_meta);
    L.Bp = (int)((rows_per_half + 127) / 128 * 128);
    L.R = halves * L.Bp;
    L.T = L.R / 128;
    int width = L.Din;
    for (int l = 0; l < mlp->n_layers; ++l) {
        L.in[l] = width;
        L.out[l] = mlp->hidden[l];
        width = mlp->hidden[l];
        if (L.out[l] > L.maxN) L.maxN = L.out[l];
    }
    for (int l = 0; l < mlp->n_layers; ++l) L.Wb[l] = take((size_t)L.out[l] * L.in[l] * 2);
    L.X0 = take((size_t)L.R * L.Din * 2);
    for (int l = 0; l < mlp->n_layers; ++l) {
        L.Z[l] = take((size_t)L.R * L.out[l] * 2);
        L.A[l] = take((size_t)L.R * L.out[l] * 2);
        L.mean[l] = take((size_t)2 * L.out[l] * 4);
        L.rstd[l] = take((size_t)2 * L.out[l] * 4);
        if (train) L.G[l] = take((size_t)L.R * L.out[l] * 2);
    }
    L.psum = take((size_t)L.T * L.maxN * 4);
    L.psq = take((size_t)L.T * L.maxN * 4);
    L.s = take((size_t)L.R * 4);
    if (train) {
        L.P1 = take((size_t)L.T * L.maxN * 4);
        L.P2 = take((size_t)L.T * L.maxN * 4);
        L.P3 = take((size_t)L.T * L.maxN * 4);
        L.Pdb = take((size_t)L.T * L.maxN * 4);
        L.S1 = take((size_t)2 * L.maxN * 4);
        L.S2 = take((size_t)2 * L.maxN * 4);
        size_t wmax = 0;
        for (int l = 0; l < mlp->n_layers; ++l) {
            L.splits[l] = wgrad_splits(L.out[l], L.in[l], L.R);
            const size_t w = (size_t)L.splits[l] * L.out[l] * L.in[l] * 4;
            if (w > wmax) wmax = w;
        }
        L.wpart = take(wmax);
        L.dX = take((size_t)L.R * L.Din * 4);
        L.ds = take((size_t)L.R * 4);
        L.loss_part = take((size_t)((rows_per_half + 7) / 8) * 4);
        L.train_ws_bytes = trs_train_workspace_bytes(m, ep);
        L.train_ws = take(L.train_ws_bytes);
    }
    L.total = off;
    return L;
}

static int check_mlp(const trs_model* m, const trs_mlp* mlp, bool train) {
    TRS_REQUIRE(m && mlp, "mlp: NULL model");
    TRS_REQUIRE(m->net == TRS_NET_MLP, "mlp: trs_model.net must be TRS_NET_MLP");
    TRS_REQUIRE(m->dim % 4 == 0 && (m->dim * (2 + m->n_meta)) % 8 == 0,
                "mlp: n_factors must be a multiple of 4 and the concatenated input a multiple of 8 (got %d x %d)",
                m->dim, 2 + m->n_meta);
    TRS_REQUIRE(mlp->n_layers >= 1 && mlp->n_layers <= TRS_MAX_LAYERS, "mlp: 1..%d hidden layers", TRS_MAX_LAYERS);
    TRS_REQUIRE(mlp->w_out && mlp->b_out, "mlp: output layer is NULL");
    for (int l = 0; l < mlp->n_layers; ++l) {
        TRS_REQUIRE(mlp->hidden[l] > 0 && mlp->hidden[l] % 8 == 0,
                    "mlp: hidden layer sizes must be multiples of 8 (layer %d has %d)", l, mlp->hidden[l]);
        TRS_REQUIRE(mlp->W[l] && mlp->b[l], "mlp: fcs.%d is NULL", l);
        if (mlp->use_bn)
            TRS_REQUIRE(mlp->gamma[l] && mlp->beta[l] && mlp->running_mean[l] && mlp->running_var[l],
                        "mlp: bns.%d is NULL", l);
        if (train) {
            TRS_REQUIRE(mlp->dW[l] && mlp->db[l], "mlp: gradient buffers of fcs.%d are NULL", l);
            if (mlp->use_bn) TRS_REQUIRE(mlp->dgamma[l] && mlp->dbeta[l], "mlp: gradient buffers of bns.%d are NULL", l);
        }
    }
    if (train) TRS_REQUIRE(mlp->dw_out && mlp->db_out, "mlp: gradient buffers of the output layer are NULL");
    return TRS_OK;
}

static inline int ew_grid(long long work_items) {
    long long g = (work_items + EW_THREADS - 1) / EW_THREADS;
    const long long cap = (long long)device_props().sm_count * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// forward of the stacked rows already gathered in X0: fills Z, A (, mean / rstd) and returns the last A
static int mlp_forward_layers(const trs_mlp* mlp, const MlpLayout& L, char* W, int halves, int n_valid, bool batch_stats,
                              bool update_running, cudaStream_t st) {
    CvtJobs cj = {};
    long long cmax = 0;
    for (int l = 0; l < mlp->n_layers; ++l) {
        cj.src[l] = mlp->W[l];
        cj.dst[l] = (bf16*)(W + L.Wb[l]);
        cj.n[l] = (long long)L.out[l] * L.in[l];
        if (cj.n[l] > cmax) cmax = cj.n[l];
    }
    cvt_kernel<<<dim3(ew_grid(cmax), mlp->n_layers), EW_THREADS, 0, st>>>(cj);
    for (int l = 0; l < mlp->n_layers; ++l) {
        const int N = L.out[l];
        trs_gemm_args g = {};
        g.a = l == 0 ? (W + L.X0) : (W + L.A[l - 1]);
        g.b = W + L.Wb[l];
        g.lda = g.ldb = g.k = L.in[l];
        g.m = L.R;
        g.n = N;
        g.out = W + L.Z[l];
        g.ldc = N;
        g.out_bf16 = 1;
        g.splits = 1;
        g.bias = mlp->b[l];
        const bool stats = mlp->use_bn && batch_stats;
        if (stats) {
            g.col_sum = (float*)(W + L.psum);
            g.col_sumsq = (float*)(W + L.psq);
        }
        g.rows_per_half = L.Bp;
        g.rows_valid = n_valid;
        int rc = trs_gemm_bf16_tn(&g, (trs_stream_t)st);
        if (rc) return rc;
        float* mean = (float*)(W + L.mean[l]);
        float* rstd = (float*)(W + L.rstd[l]);
        if (stats) {
            bn_finalize_kernel<<<(N + 31) / 32, 32 * TS_WARPS, 0, st>>>(g.col_sum, g.col_sumsq, L.T / halves, halves, N, n_valid,
                                                                mean, rstd, update_running ? mlp->running_mean[l] : nullptr,
                                                                update_running ? mlp->running_var[l] : nullptr);
        } else if (mlp->use_bn) {
            bn_eval_stats_kernel<<<(N + 127) / 128, 128, 0, st>>>(mlp->running_mean[l], mlp->running_var[l], N, mean, rstd);
        }
        // eval mode: one (mean, rstd) row shared by all rows -> present it as a single "pass" of R rows
        bn_relu_kernel<<<dim3(L.T, (N + 63) / 64), EW_THREADS, 0, st>>>(
            (const bf16*)(W + L.Z[l]), (bf16*)(W + L.A[l]), mean, rstd, mlp->gamma[l], mlp->beta[l], L.R, N,
            stats ? L.Bp : L.R, stats ? n_valid : L.R, mlp->use_bn);
    }
    TRS_CUDA(cudaGetLastError());
    return TRS_OK;
}

}  // namespace trs

using namespace trs;

extern "C" size_t trs_mlp_forward_workspace_bytes(const trs_model* model, const trs_mlp* mlp, int64_t n) {
    if (check_mlp(model, mlp, false) || n <= 0) return 0;
    return mlp_layout(model, mlp, n, 1, false, nullptr).total;
}

extern "C" int trs_mlp_forward(const trs_model* model, const trs_mlp* mlp, const int64_t* user, const int64_t* item,
                               const int64_t* meta, int64_t n, int batch_stats, float* out, void* workspace,
                               size_t workspace_bytes, trs_stream_t stream) {
    int rc = check_mlp(model, mlp, false);
    if (rc) return rc;
    if (n == 0) return TRS_OK;
    TRS_REQUIRE(user && item && out && workspace, "mlp forward: NULL pointer");
    TRS_REQUIRE(model->n_meta == 0 || meta, "mlp forward: model has metadata tables but meta ids are NULL");
    TRS_REQUIRE(n < (1ll << 30), "mlp forward: too many rows in one call");
    const MlpLayout L = mlp_layout(model, mlp, n, 1, false, nullptr);
    if (workspace_bytes < L.total) {
        set_error("mlp forward workspace too small: %zu < %zu", workspace_bytes, L.total);
        return TRS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* W = (char*)workspace;
    gather_concat_kernel<<<ew_grid((long long)L.R * L.Din / 4), EW_THREADS, 0, st>>>(
        *model, user, item, item, meta, meta, (int)n, L.Bp, 1, (bf16*)(W + L.X0));
    rc = mlp_forward_layers(mlp, L, W, 1, (int)n, batch_stats != 0, batch_stats != 0, st);
    if (rc) return rc;
    const int last = mlp->n_layers - 1;
    out_scores_kernel<<<(int)((n + 7) / 8), EW_THREADS, 0, st>>>((const bf16*)(W + L.A[last]), mlp->w_out, mlp->b_out,
                                                                 L.out[last], (int)n, out);
    TRS_CUDA(cudaGetLastError());
    return TRS_OK;
}

extern "C" size_t trs_mlp_train_workspace_bytes(const trs_model* model, const trs_mlp* mlp, const trs_epoch* epoch) {
    if (check_mlp(model, mlp, false) || !epoch || epoch->batch <= 0) return 0;
    return mlp_layout(model, mlp, epoch->batch, 2, true, epoch).total;
}

extern "C" int trs_mlp_train_steps(const trs_model* model, const trs_mlp* mlp, const trs_epoch* ep,
                                   const trs_optim* optim, const void* plan, void* workspace, size_t workspace_bytes,
                                   int first_step, int n_steps, float* loss, trs_stream_t stream) {
    int rc = check_mlp(model, mlp, true);
    if (rc) return rc;
    TRS_REQUIRE(ep && ep->user && ep->pos && ep->neg && ep->batch > 0, "mlp train: bad epoch");
    TRS_REQUIRE(model->n_meta == 0 || (ep->pos_meta && ep->neg_meta), "mlp train: metadata ids are NULL");
    TRS_REQUIRE(optim && optim->step_scale && plan && workspace && loss, "mlp train: NULL pointer");
    TRS_REQUIRE(optim->kind == TRS_OPT_SGD || optim->kind == TRS_OPT_ADAGRAD,
                "mlp train: the tower's dense parameters take SGD or Adagrad (SparseAdam rejects dense gradients)");
    const int64_t steps = n_steps_of(ep);
    TRS_REQUIRE(first_step >= 0 && n_steps >= 0 && first_step + (int64_t)n_steps <= steps, "mlp train: steps out of range");
    const MlpLayout L = mlp_layout(model, mlp, ep->batch, 2, true, ep);
    if (workspace_bytes < L.total) {
        set_error("mlp train workspace too small: %zu < %zu", workspace_bytes, L.total);
        return TRS_ERR_WORKSPACE;
    }
    cudaStream_t st = (cudaStream_t)stream;
    char* W = (char*)workspace;
    const int D = model->dim, F = model->n_meta, NL = mlp->n_layers, last = NL - 1;
    const StagePtrs stage = stage_pointers(model, ep, W + L.train_ws);

    DenseJobs dj = {};
    int nd = 0;
    long long dmax = 0;
    auto add_dense = [&](float* p, const float* g, float* s0, long long n) {
        dj.p[nd] = p;
        dj.g[nd] = g;
        dj.s0[nd] = s0;
        dj.n[nd] = n;
        if (n > dmax) dmax = n;
        ++nd;
    };
    for (int l = 0; l < NL; ++l) {
        add_dense(mlp->W[l], mlp->dW[l], mlp->s0W[l], (long long)L.out[l] * L.in[l]);
        add_dense(mlp->b[l], mlp->db[l], mlp->s0b[l], L.out[l]);
        if (mlp->use_bn) {
            add_dense(mlp->gamma[l], mlp->dgamma[l], mlp->s0gamma[l], L.out[l]);
            add_dense(mlp->beta[l], mlp->dbeta[l], mlp->s0beta[l], L.out[l]);
        }
    }
    add_dense(mlp->w_out, mlp->dw_out, mlp->s0w_out, L.out[last]);
    add_dense(mlp->b_out, mlp->db_out, mlp->s0b_out, 1);
    if (optim->kind == TRS_OPT_ADAGRAD)
        for (int i = 0; i < nd; ++i) TRS_REQUIRE(dj.s0[i], "mlp train: Adagrad state of a tower parameter is NULL");

    for (int si = 0; si < n_steps; ++si) {
        const int64_t s = first_step + si;
        const int64_t lo = s * (int64_t)ep->batch;
        const int B = (int)((ep->n_samples - lo) < ep->batch ? (ep->n_samples - lo) : ep->batch);
        // ---------------- forward ----------------
        gather_concat_kernel<<<ew_grid((long long)L.R * L.Din / 4), EW_THREADS, 0, st>>>(
            *model, ep->user + lo, ep->pos + lo, ep->neg + lo, F ? ep->pos_meta + lo * F : nullptr,
            F ? ep->neg_meta + lo * F : nullptr, B, L.Bp, 2, (bf16*)(W + L.X0));
        rc = mlp_forward_layers(mlp, L, W, 2, B, true, true, st);
        if (rc) return rc;
        const int nparts = (B + 7) / 8;
        out_hinge_kernel<<<nparts, EW_THREADS, 0, st>>>((const bf16*)(W + L.A[last]), mlp->w_out, mlp->b_out, L.out[last],
                                                       L.Bp, B, (float*)(W + L.s), (float*)(W + L.ds),
                                                       (float*)(W + L.loss_part));
        loss_finalize_kernel<<<1, 32, 0, st>>>((const float*)(W + L.loss_part), nparts, B, loss + si);
        // ---------------- backward ----------------
        for (int l = last; l >= 0; --l) {
            const int N = L.out[l];
            BwdIn in = {};
            in.dA = l == last ? nullptr : (const bf16*)(W + L.G[l]);
            in.ds = (const float*)(W + L.ds);
            in.w_out = mlp->w_out;
            in.A = (const bf16*)(W + L.A[l]);
            in.Z = (const bf16*)(W + L.Z[l]);
            in.mean = (const float*)(W + L.mean[l]);
            in.rstd = (const float*)(W + L.rstd[l]);
            in.R = L.R;
            in.N = N;
            in.Bp = L.Bp;
            in.B = B;
            in.use_bn = mlp->use_bn;
            float *P1 = (float*)(W + L.P1), *P2 = (float*)(W + L.P2), *P3 = (float*)(W + L.P3);
            float *S1 = (float*)(W + L.S1), *S2 = (float*)(W + L.S2), *Pdb = (float*)(W + L.Pdb);
            const dim3 tg(L.T, (N + 63) / 64);
            if (l == last) {
                bn_bwd_reduce_kernel<true><<<tg, EW_THREADS, 0, st>>>(in, P1, mlp->use_bn ? P2 : nullptr, P3);
            } else if (mlp->use_bn) {
                bn_bwd_reduce_kernel<false><<<tg, EW_THREADS, 0, st>>>(in, P1, P2, nullptr);
            }
            if (l == last || mlp->use_bn)
                bn_bwd_finalize_kernel<<<(N + 31) / 32, 32 * TS_WARPS, 0, st>>>(
                    P1, mlp->use_bn ? P2 : nullptr, l == last ? P3 : nullptr, L.T / 2, N, S1, S2,
                    mlp->use_bn ? mlp->dgamma[l] : nullptr, mlp->use_bn ? mlp->dbeta[l] : nullptr,
                    l == last ? mlp->dw_out : nullptr);
            if (l == last)
                bn_bwd_apply_kernel<true><<<tg, EW_THREADS, 0, st>>>(in, mlp->gamma[l], S1, S2, (bf16*)(W + L.G[l]), Pdb);
            else
                bn_bwd_apply_kernel<false><<<tg, EW_THREADS, 0, st>>>(in, mlp->gamma[l], S1, S2, (bf16*)(W + L.G[l]), Pdb);
            colsum_tiles_kernel<<<(N + 31) / 32, 32 * TS_WARPS, 0, st>>>(Pdb, L.T, N, mlp->db[l]);
            // wgrad: dW_l[out, in] = sum_r dZ_l[r, out] * X_l[r, in]
            trs_gemm_args g = {};
            g.a = W + L.G[l];
            g.lda = N;
            g.a_mn = 1;
            g.b = l == 0 ? (W + L.X0) : (W + L.A[l - 1]);
            g.ldb = L.in[l];
            g.b_mn = 1;
            g.m = N;
            g.n = L.in[l];
            g.k = L.R;
            g.out = W + L.wpart;
            g.ldc = L.in[l];
            g.splits = L.splits[l];
            g.split_stride = (int64_t)N * L.in[l];
            if ((rc = trs_gemm_bf16_tn(&g, (trs_stream_t)st))) return rc;
            reduce_partials_kernel<<<ew_grid((long long)N * L.in[l]), EW_THREADS, 0, st>>>(
                (const float*)(W + L.wpart), L.splits[l], (long long)N * L.in[l], (long long)N * L.in[l], mlp->dW[l]);
            // dgrad: dX_l[r, in] = sum_n dZ_l[r, n] * W_l[n, in]
            trs_gemm_args d = {};
            d.a = W + L.G[l];
            d.lda = N;
            d.b = W + L.Wb[l];
            d.ldb = L.in[l];
            d.b_mn = 1;
            d.m = L.R;
            d.n = L.in[l];
            d.k = N;
            d.splits = 1;
            d.ldc = L.in[l];
            if (l == 0) {
                d.out = W + L.dX;
            } else {
                d.out = W + L.G[l - 1];
                d.out_bf16 = 1;
            }
            if ((rc = trs_gemm_bf16_tn(&d, (trs_stream_t)st))) return rc;
        }
        // db_out = sum ds = sum_b (g_b - g_b): the two passes cancel exactly, as in the reference
        TRS_CUDA(cudaMemsetAsync(mlp->db_out, 0, sizeof(float), st));
        // ---------------- updates ----------------
        dense_update_kernel<<<dim3(ew_grid(dmax), nd), EW_THREADS, 0, st>>>(dj, optim->kind, (float)optim->eps,
                                                                            optim->step_scale, (int)s);
        stage_grads_kernel<<<ew_grid((long long)B * L.Din / 4), EW_THREADS, 0, st>>>((const float*)(W + L.dX), L.Din, D, B,
                                                                                    L.Bp, stage);
        TRS_CUDA(cudaGetLastError());
        rc = run_train_steps(model, ep, optim, plan, W + L.train_ws, L.train_ws_bytes, (int)s, 1, nullptr, st);
        if (rc) return rc;
    }
    return TRS_OK;
}
