// plan.cuh -- byte layout of the caller-allocated sort plan shared by plan.cu and train.cu.
#pragma once
#include "common.cuh"

namespace trs {

// offsets in bytes into the plan buffer; every array is uint32.
//   *_key[k]  : row id at sorted position k of the step's segment
//   *_perm[k] : lookup id (0..mult*B_s-1, local to the step) at sorted position k
// user arrays hold n_samples entries (segment of step s starts at s*batch); item/meta arrays
// hold 2*n_samples entries (segment of step s starts at 2*s*batch).
//
// On top of the sorted pairs the plan holds, per step, what the training kernel does with every touched row
// (built once per epoch, off the training kernel's critical path):
//   single_user[lookup], single_item[lookup] (uint8)   1 = this lookup's row is looked up exactly ONCE in the
//        step: no other sample reads or writes it, so the sample's own row group applies the optimizer update
//        right in phase A (no staging, no barrier).  Indexed like the perm values: step*B + j / 2*step*B + j.
//   item_cnt[s], items[s][i] (uint4)   a SHORT segment: one row with 2..LONG_SEG_T lookups (metadata spaces
//        and the MLP tower's plans also list their single-lookup rows here)
//        x = id space (0 user, 1 item, 2+f metadata f) | count << 8
//        y = start position k in the step's sorted arrays, z = row id, w = first lookup id (perm[k])
//   long_cnt[s], long_segs[s][i] (uint4)  a LONG segment: {id space, start k, length c, row id}
// A row looked up more than LONG_SEG_T times in a step (always the case for the small metadata tables, and
// for hot rows under skew) is reduced by a whole CTA: its row groups sum strided subsets of the lookups, the
// partial sums are added in group order, one group applies the update.
constexpr int LONG_SEG_T = 8;

struct PlanLayout {
    size_t user_key, user_perm, item_key, item_perm;
    size_t meta_key[TRS_MAX_META], meta_perm[TRS_MAX_META];
    size_t item_cnt, long_cnt, single_user, single_item, items, long_segs;
    int item_cap;   // entries per step in items
    int long_cap;   // entries per step in long_segs
    size_t total;
};
PlanLayout plan_layout(int64_t n_samples, int batch, int n_meta);

// Stable LSD radix sort of one id space for every step of `ep` (ids of step s, lookup j: j < B_s ? a[...] :
// b[...], see plan.cu); final (row, lookup) pairs land in out_key / out_val.  tmp_key / tmp_val hold
// mult * n_samples entries each, hist hist_bytes(ep) bytes.
int sort_space(const int64_t* a, const int64_t* b, int stride, int off, int mult, int64_t n_rows,
               const trs_epoch* ep, uint32_t* out_key, uint32_t* out_val, uint32_t* tmp_key, uint32_t* tmp_val,
               uint32_t* hist, cudaStream_t st);
size_t hist_bytes(const trs_epoch* ep);
// the same sort on (key, value) pairs already in memory, with per-step segment lengths read from the device
int sort_pairs(uint32_t* key0, uint32_t* val0, uint32_t* key1, uint32_t* val1, int mult, int64_t n_rows,
               const trs_epoch* ep, const uint32_t* len_arr, int len_stride, uint32_t* hist, cudaStream_t st);
int sort_passes(int64_t n_rows);

}  // namespace trs
