// train.cuh -- what mlp.cu needs from train.cu: the reduce + update phase of the fused training kernel,
// run on gradient rows the MLP backward staged itself.
#pragma once
#include "common.cuh"

namespace trs {

// where the per-lookup gradient rows of one step live inside the train workspace
// (gU [B, dim]: one row per sample; gI / gM[f] [2B, dim]: positives then negatives)
struct StagePtrs {
    float* gU;
    float* gI;
    float* gM[TRS_MAX_META];
};
StagePtrs stage_pointers(const trs_model* model, const trs_epoch* ep, void* workspace);

struct OptScalars {
    int kind;
    float omb1, omb2, eps;  // 1-beta1, 1-beta2, eps rounded to fp32 as torch's scalar ops do
    const float* step_scale;
};

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

// fills OptScalars from the C-ABI optimizer description (beta / eps rounded to fp32 as torch's scalar ops do)
inline OptScalars make_opt_scalars(const trs_optim* optim) {
    OptScalars os;
    os.kind = optim->kind;
    os.omb1 = (float)(1.0 - optim->beta1);
    os.omb2 = (float)(1.0 - optim->beta2);
    os.eps = (float)optim->eps;
    os.step_scale = optim->step_scale;
    return os;
}

int run_train_steps(const trs_model* model, const trs_epoch* ep, const trs_optim* optim, const void* plan,
                    void* workspace, size_t workspace_bytes, int first_step, int n_steps, float* loss,
                    cudaStream_t stream);

}  // namespace trs
