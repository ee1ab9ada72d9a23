// api.cu -- library plumbing: version, error string, device properties.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace trs {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

const DeviceProps& device_props() {
    static thread_local DeviceProps p = {0, 0};
    static thread_local int cached_dev = -1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev != cached_dev) {
        cudaDeviceGetAttribute(&p.sm_count, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&p.max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        cached_dev = dev;
    }
    return p;
}

int check_model(const trs_model* m, RowShape* shape) {
    TRS_REQUIRE(m != nullptr, "model is NULL");
    TRS_REQUIRE(m->net == TRS_NET_LINEAR || m->net == TRS_NET_FM || m->net == TRS_NET_MLP, "unknown net %d", m->net);
    TRS_REQUIRE(m->n_meta >= 0 && m->n_meta <= TRS_MAX_META, "n_meta %d out of range", m->n_meta);
    TRS_REQUIRE(pick_row_shape(m->dim, shape), "unsupported n_factors %d (need 1..512, or a multiple of 4)", m->dim);
    TRS_REQUIRE(m->user.emb && m->item.emb, "user/item embedding pointer is NULL");
    for (int f = 0; f < m->n_meta; ++f) TRS_REQUIRE(m->meta[f].emb, "metadata table %d is NULL", f);
    return TRS_OK;
}

}  // namespace trs

extern "C" int trs_abi_version(void) { return TRS_ABI_VERSION; }
extern "C" const char* trs_last_error(void) { return trs::g_err; }
