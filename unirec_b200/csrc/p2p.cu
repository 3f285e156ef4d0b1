// Peer-memory plumbing for the row-sharded tables (one process per GPU on one NVSwitch box): CUDA IPC handles let every rank
// map its peers' table shards and gradient buffers, so that kernels read the rows they need straight over NVLink
// (`shard_ptrs[id % W] + (id / W) * d`) instead of staging them through zero-filled reduce-scatter / all-gather buffers.
#include <cuda.h>
#include <string.h>
#include "common.cuh"

extern "C" {

// handle_out: 64 host bytes (cudaIpcMemHandle_t) of the allocation that contains dev_ptr; offset_out: dev_ptr - allocation base
int ur_ipc_export(const void* dev_ptr, void* handle_out, int64_t* offset_out) {
    typedef CUresult (*GetRangeFn)(CUdeviceptr*, size_t*, CUdeviceptr);
    static GetRangeFn get_range = nullptr;
    if (!get_range) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q) != cudaSuccess || !sym) return UR_ERR_UNSUPPORTED;
        get_range = (GetRangeFn)sym;
    }
    CUdeviceptr base = 0;
    size_t size = 0;
    if (get_range(&base, &size, (CUdeviceptr)dev_ptr) != CUDA_SUCCESS) return UR_ERR_BAD_ARG;
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, (void*)base);
    if (e != cudaSuccess) { cudaGetLastError(); return -(1000 + (int)e); }
    memcpy(handle_out, &h, sizeof(h));
    *offset_out = (int64_t)((CUdeviceptr)dev_ptr - base);
    return UR_OK;
}

// maps a peer allocation into this process (kept mapped for the life of the process) and returns base + offset
int ur_ipc_open(const void* handle, int64_t offset, int64_t* ptr_out) {
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof(h));
    void* p = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) { cudaGetLastError(); return -(1000 + (int)e); }
    *ptr_out = (int64_t)((char*)p + offset);
    return UR_OK;
}

}  // extern "C"
