// plan.cuh -- byte layout of the caller-allocated sort plan shared by plan.cu and train.cu.
#pragma once
#include "common.cuh"

namespace trs {

// offsets in bytes into the plan buffer; every array is uint32.
//   *_key[k]  : row id at sorted position k of the step's segment
//   *_perm[k] : lookup id (0..mult*B_s-1, local to the step) at sorted position k
// user arrays hold n_samples entries (segment of step s starts at s*batch); item/meta arrays
// hold 2*n_samples entries (segment of step s starts at 2*s*batch).
struct PlanLayout {
    size_t user_key, user_perm, item_key, item_perm;
    size_t meta_key[TRS_MAX_META], meta_perm[TRS_MAX_META];
    size_t total;
};
PlanLayout plan_layout(int64_t n_samples, int n_meta);

}  // namespace trs
