// The one collective of the path: all-gather of the per-rank outputs (logits, or the per-point
// part logits of SV_DGCNN_PSEG) over NCCL / NVLink -- replaces nn.DataParallel's gather
// (reference main_partseg_dgcnn.py:116, main_cls_dgcnn.py:125).  The library does not link against NCCL: the symbol is
// taken from the NCCL the process already has loaded (PyTorch's), else from libnccl.so.2, at the first call.
#include "common.cuh"
#include <dlfcn.h>

namespace {
typedef int (*nccl_allgather_fn)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef const char* (*nccl_errstr_fn)(int);
constexpr int NCCL_FLOAT32 = 7;      // ncclFloat32 in nccl.h's ncclDataType_t

nccl_allgather_fn g_allgather = nullptr;
nccl_errstr_fn g_errstr = nullptr;

bool resolve()
{
    if (g_allgather) return true;
    void* sym = dlsym(RTLD_DEFAULT, "ncclAllGather");
    void* h = nullptr;
    if (!sym) {
        h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (h) sym = dlsym(h, "ncclAllGather");
    }
    if (!sym) return false;
    g_allgather = reinterpret_cast<nccl_allgather_fn>(sym);
    void* es = dlsym(h ? h : RTLD_DEFAULT, "ncclGetErrorString");
    g_errstr = reinterpret_cast<nccl_errstr_fn>(es);
    return true;
}
}  // namespace

extern "C" int svnet_allgather_logits(void* nccl_comm, const float* local, float* all, size_t count_per_rank, void* stream)
{
    SV_REQUIRE(nccl_comm && local && all, "svnet_allgather_logits: null pointer");
    if (count_per_rank == 0) return SVNET_OK;
    if (!resolve()) {
        svnet_set_error("svnet_allgather_logits: no NCCL in this process and libnccl.so.2 not found");
        return SVNET_ERR_CUDA;
    }
    const int rc = g_allgather(local, all, count_per_rank, NCCL_FLOAT32, nccl_comm, sv_stream(stream));
    if (rc != 0) {
        svnet_set_error("svnet_allgather_logits: ncclAllGather failed (%d): %s", rc, g_errstr ? g_errstr(rc) : "?");
        return SVNET_ERR_CUDA;
    }
    return SVNET_OK;
}
