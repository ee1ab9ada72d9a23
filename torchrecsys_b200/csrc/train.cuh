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

int run_train_steps(const trs_model* model, const trs_epoch* ep, const trs_optim* optim, const void* plan,
                    void* workspace, size_t workspace_bytes, int first_step, int n_steps, float* loss,
                    cudaStream_t stream);

}  // namespace trs
