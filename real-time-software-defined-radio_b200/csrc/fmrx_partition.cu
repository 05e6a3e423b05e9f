// Spatial partition of the SMs between the latency-bound PLL kernel and the throughput-bound FIR kernels.
//
// The device-resident pipeline overlaps step k's PLLs with step k+1's front end and step k-1's back end.  On shared
// SMs that overlap buys nothing: a PLL warp (one long dependency chain, ready ~1 cycle in 4) that has to win the issue
// slot against always-ready FIR warps runs at a third of its speed (measured with fmrx_batch_timeline).  A CUDA green
// context gives the PLL stream SMs of its own: 8192 loops = 256 warps on e.g. 32 SMs is two warps per scheduler,
// which still hides nothing from the chain, and the FIR grids keep the other 116 SMs to themselves.
//
// The driver entry points are resolved through cudaGetDriverEntryPoint, so libfmrx.so keeps no link-time dependency on
// libcuda (it must load on a host without a driver: the C-ABI export test runs there).
#include <cstdlib>

#include <cuda.h>
#include <cuda_runtime.h>

#include "fmrx_internal.h"

namespace fmrx {

namespace {
template <class F>
bool entry(const char *name, F &fn) {
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    fn = reinterpret_cast<F>(p);
    return true;
}
}  // namespace

struct SmPartition {
    CUgreenCtx small = nullptr, big = nullptr;
    int sms_small = 0, sms_big = 0;
};

void partition_destroy(SmPartition *p) {
    if (!p) return;
    CUresult (*destroy)(CUgreenCtx) = nullptr;
    if (entry("cuGreenCtxDestroy", destroy)) {
        if (p->small) destroy(p->small);
        if (p->big) destroy(p->big);
    }
    delete p;
}

// Splits `device` into a partition of >= want_small SMs and the rest; creates one stream in the small one and
// n_big streams in the large one.  Returns nullptr (and creates nothing) when the driver cannot do it.
SmPartition *partition_create(int device, int want_small, int prio_small, fmrx_stream_t *s_small, int n_big, const int *prio_big, fmrx_stream_t *s_big) {
    CUresult (*devGet)(CUdevice *, int) = nullptr;
    CUresult (*getRes)(CUdevice, CUdevResource *, CUdevResourceType) = nullptr;
    CUresult (*split)(CUdevResource *, unsigned int *, const CUdevResource *, CUdevResource *, unsigned int, unsigned int) = nullptr;
    CUresult (*genDesc)(CUdevResourceDesc *, CUdevResource *, unsigned int) = nullptr;
    CUresult (*ctxCreate)(CUgreenCtx *, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*streamCreate)(CUstream *, CUgreenCtx, unsigned int, int) = nullptr;
    if (!entry("cuDeviceGet", devGet) || !entry("cuDeviceGetDevResource", getRes) || !entry("cuDevSmResourceSplitByCount", split) ||
        !entry("cuDevResourceGenerateDesc", genDesc) || !entry("cuGreenCtxCreate", ctxCreate) || !entry("cuGreenCtxStreamCreate", streamCreate))
        return nullptr;
    if (cudaSetDevice(device) != cudaSuccess || cudaFree(nullptr) != cudaSuccess) return nullptr;  // primary context must exist
    CUdevice dev;
    CUdevResource all, part, rest;
    unsigned int groups = 1;
    if (devGet(&dev, device) != CUDA_SUCCESS || getRes(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return nullptr;
    if (want_small <= 0 || (unsigned)want_small + 8 > all.sm.smCount) return nullptr;
    // without the flag the split granularity is 8 SMs on this architecture; IGNORE_SM_COSCHEDULING (clusters are not used
    // here) brings it down to 2, which lets the two sides be balanced more finely
    unsigned flags = CU_DEV_SM_RESOURCE_SPLIT_IGNORE_SM_COSCHEDULING;
    if (const char *e = getenv("FMRX_PLL_SPLIT_FLAGS")) flags = (unsigned)atoi(e);
    if (split(&part, &groups, &all, &rest, flags, (unsigned)want_small) != CUDA_SUCCESS || groups != 1 || rest.sm.smCount == 0) {
        groups = 1;
        if (split(&part, &groups, &all, &rest, 0, (unsigned)want_small) != CUDA_SUCCESS || groups != 1 || rest.sm.smCount == 0) return nullptr;
    }
    CUdevResourceDesc d_small, d_big;
    if (genDesc(&d_small, &part, 1) != CUDA_SUCCESS || genDesc(&d_big, &rest, 1) != CUDA_SUCCESS) return nullptr;
    SmPartition *p = new SmPartition();
    p->sms_small = (int)part.sm.smCount;
    p->sms_big = (int)rest.sm.smCount;
    if (ctxCreate(&p->small, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS || ctxCreate(&p->big, d_big, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        partition_destroy(p);
        return nullptr;
    }
    bool ok = streamCreate(reinterpret_cast<CUstream *>(s_small), p->small, CU_STREAM_NON_BLOCKING, prio_small) == CUDA_SUCCESS;
    for (int i = 0; ok && i < n_big; ++i) ok = streamCreate(reinterpret_cast<CUstream *>(&s_big[i]), p->big, CU_STREAM_NON_BLOCKING, prio_big[i]) == CUDA_SUCCESS;
    if (!ok) {
        if (*s_small) cudaStreamDestroy(*s_small);
        for (int i = 0; i < n_big; ++i) if (s_big[i]) cudaStreamDestroy(s_big[i]);
        *s_small = nullptr;
        for (int i = 0; i < n_big; ++i) s_big[i] = nullptr;
        partition_destroy(p);
        return nullptr;
    }
    return p;
}

void partition_sizes(const SmPartition *p, int *small, int *big) {
    *small = p ? p->sms_small : 0;
    *big = p ? p->sms_big : 0;
}

}  // namespace fmrx
