// p2p_probe.cu -- what does NVLink give a KERNEL that reads / writes random 512-byte rows of a peer's HBM?
// (the access pattern of trs::shard_train_kernel's phase A).  One process, two devices, peer access enabled;
// device 0 runs the kernels, the table lives on device `peer` (0 = local reference numbers).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/p2p_probe tools/p2p_probe.cu && /tmp/p2p_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

constexpr int ROW_F4 = 32;  // 512-byte rows: one float4 per lane

__device__ __forceinline__ void cp_async16(void* smem, const void* g) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(g) : "memory");
}

// U rows in flight per warp, plain 128-bit loads (ld.global.cg)
template <int U>
__global__ void __launch_bounds__(512) gather_ldg(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n,
                                                  float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = warp * U; i < n; i += nwarps * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = (i + u < n) ? idx[i + u] : idx[i];
            v[u] = __ldcg(tab + (size_t)r * ROW_F4 + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// U rows in flight per warp through cp.async into thread-private shared memory slots
template <int U>
__global__ void __launch_bounds__(512) gather_cpasync(const float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n,
                                                      float4* __restrict__ out) {
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = warp * U; i < n; i += nwarps * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = (i + u < n) ? idx[i + u] : idx[i];
            cp_async16(sm + u * blockDim.x + threadIdx.x, tab + (size_t)r * ROW_F4 + lane);
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
#pragma unroll
        for (int u = 0; u < U; ++u) { float4 v = sm[u * blockDim.x + threadIdx.x]; acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// random row WRITES (posted): U rows per warp iteration
template <int U>
__global__ void __launch_bounds__(512) scatter_st(float4* __restrict__ tab, const uint32_t* __restrict__ idx, int n) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    const float4 v = make_float4(lane, 1, 2, 3);
    for (int i = warp * U; i < n; i += nwarps * U) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (i + u < n) tab[(size_t)idx[i + u] * ROW_F4 + lane] = v;
        }
    }
}

// both at once: every warp alternates U row reads and U row writes (different rows)
template <int U>
__global__ void __launch_bounds__(512) gather_scatter(const float4* __restrict__ rd, float4* __restrict__ wr,
                                                      const uint32_t* __restrict__ idx, int n, float4* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
    float4 acc = make_float4(0, 0, 0, 0);
    for (int i = warp * U; i < n; i += nwarps * U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = (i + u < n) ? idx[i + u] : idx[i];
            v[u] = __ldcg(rd + (size_t)r * ROW_F4 + lane);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc.x += v[u].x;
            if (i + u < n) wr[(size_t)idx[n - 1 - (i + u)] * ROW_F4 + lane] = v[u];
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <typename F>
static float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    f();
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(a));
        f();
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (ms < best) best = ms;
    }
    CK(cudaGetLastError());
    return best;
}

int main(int argc, char** argv) {
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    // default 3.2 GB of 512-byte rows (a C4 user shard at 8 ranks); argv[2] = rows (e.g. 80000000 = 41 GB)
    const size_t rows = argc > 2 ? (size_t)atoll(argv[2]) : 6250000;
    const int n = argc > 1 ? atoi(argv[1]) : 1 << 20;
    printf("devices %d, table %zu rows x 512 B, %d random row accesses per launch (%.1f MB)\n", ndev, rows, n, n * 512.0 / 1e6);
    uint32_t* h = (uint32_t*)malloc(n * 4);
    srand(1);
    for (int i = 0; i < n; ++i) h[i] = (uint32_t)((((uint64_t)rand() << 31) ^ (uint64_t)rand() * 2654435761ull) % rows);
    for (int peer = 0; peer < (ndev > 1 ? 2 : 1); ++peer) {
        CK(cudaSetDevice(peer));
        float4* tab; float4* tab2;
        CK(cudaMalloc(&tab, rows * 512));
        tab2 = tab;  // reads and writes hit the same table (different rows)
        CK(cudaMemset(tab, 0, rows * 512));
        CK(cudaSetDevice(0));
        if (peer) {
            int can = 0;
            CK(cudaDeviceCanAccessPeer(&can, 0, peer));
            if (!can) { printf("no peer access\n"); return 0; }
            cudaError_t e = cudaDeviceEnablePeerAccess(peer, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { printf("enable peer: %s\n", cudaGetErrorString(e)); return 1; }
            cudaGetLastError();
        }
        uint32_t* idx; float4* out;
        CK(cudaMalloc(&idx, n * 4));
        CK(cudaMalloc(&out, (size_t)592 * 512 * 16));
        CK(cudaMemcpy(idx, h, n * 4, cudaMemcpyHostToDevice));
        const double gb = n * 512.0 / 1e9;
        printf("---- table on device %d (%s) ----\n", peer, peer ? "PEER over NVLink" : "local HBM");
#define RUN(name, expr, bytes) { float ms = time_ms([&] { expr; }); printf("  %-34s %8.1f us  %7.1f GB/s\n", name, ms * 1e3, (bytes) / (ms * 1e-3)); }
        RUN("gather ldg U=1  148x512", (gather_ldg<1><<<148, 512>>>(tab, idx, n, out)), gb);
        RUN("gather ldg U=2  148x512", (gather_ldg<2><<<148, 512>>>(tab, idx, n, out)), gb);
        RUN("gather ldg U=4  148x512", (gather_ldg<4><<<148, 512>>>(tab, idx, n, out)), gb);
        RUN("gather ldg U=8  148x512", (gather_ldg<8><<<148, 512>>>(tab, idx, n, out)), gb);
        RUN("gather ldg U=8  296x512", (gather_ldg<8><<<296, 512>>>(tab, idx, n, out)), gb);
        RUN("gather ldg U=8  592x512", (gather_ldg<8><<<592, 512>>>(tab, idx, n, out)), gb);
        CK(cudaFuncSetAttribute(gather_cpasync<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 512 * 16));
        CK(cudaFuncSetAttribute(gather_cpasync<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 512 * 16));
        RUN("gather cp.async U=4  148x512", (gather_cpasync<4><<<148, 512, 4 * 512 * 16>>>(tab, idx, n, out)), gb);
        RUN("gather cp.async U=8  148x512", (gather_cpasync<8><<<148, 512, 8 * 512 * 16>>>(tab, idx, n, out)), gb);
        RUN("gather cp.async U=16 148x512", (gather_cpasync<16><<<148, 512, 16 * 512 * 16>>>(tab, idx, n, out)), gb);
        RUN("scatter st U=1  148x512", (scatter_st<1><<<148, 512>>>(tab2, idx, n)), gb);
        RUN("scatter st U=4  148x512", (scatter_st<4><<<148, 512>>>(tab2, idx, n)), gb);
        RUN("scatter st U=4  592x512", (scatter_st<4><<<592, 512>>>(tab2, idx, n)), gb);
        RUN("gather+scatter U=4 148x512 (each)", (gather_scatter<4><<<148, 512>>>(tab, tab2, idx, n, out)), gb);
        RUN("gather+scatter U=8 296x512 (each)", (gather_scatter<8><<<296, 512>>>(tab, tab2, idx, n, out)), gb);
        // small batches: the kernel's phase A moves ~16k rows per step per direction
        for (int m : {16384, 65536}) {
            char nm[64];
            snprintf(nm, 64, "gather ldg U=8, only %d rows", m);
            RUN(nm, (gather_ldg<8><<<148, 512>>>(tab, idx, m, out)), m * 512.0 / 1e9);
            snprintf(nm, 64, "scatter st U=4, only %d rows", m);
            RUN(nm, (scatter_st<4><<<148, 512>>>(tab2, idx, m)), m * 512.0 / 1e9);
        }
        CK(cudaFree(idx)); CK(cudaFree(out));
        CK(cudaSetDevice(peer));
        CK(cudaFree(tab));
    }
    return 0;
}
