// rz_sort.cu — key/index radix sort between the stages of the staged K1 (cub::DeviceRadixSort: library plumbing,
// 16-bit keys => two 8-bit passes over 6 bytes per entry), plain or wrapped in a CUDA graph that sizes it on the device.
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>
#include <algorithm>

extern "C" size_t rz_sort_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned short *)nullptr, (unsigned short *)nullptr, (const uint32_t *)nullptr,
                                    (uint32_t *)nullptr, (int)n, 0, 16);
    return bytes;
}

// keys_in[i] with value i -> idx_out in key order.  Unused slots carry 0xffff: the largest key, and since the sort is
// stable and the live entries occupy the lowest indices, a live entry with that key still precedes every unused slot.
extern "C" cudaError_t rz_sort_keys(const unsigned short *keys_in, unsigned short *keys_out, const uint32_t *iota, uint32_t *idx_out, uint32_t n,
                                    void *temp, size_t temp_bytes, cudaStream_t stream) {
    return cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, iota, idx_out, (int)n, 0, 16, stream);
}

__global__ void rz_iota_kernel(uint32_t *p, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = i;
}

extern "C" cudaError_t rz_iota(uint32_t *p, uint32_t n, cudaStream_t stream) {
    if (n == 0) return cudaSuccess;
    rz_iota_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>(p, n);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------ sort sized on the device
// The number of live entries of a queue is only known on the device, and cub takes its item count from the host.  Sorting
// every slot of a pass costs 1.0 ms per sort at 2^27-slot passes although 66 % / 40 % / 25 % of the slots are live after the
// first three segments.  So the sort is a CUDA graph with a SWITCH conditional node: a one-thread kernel reads the live count
// and selects the body whose (captured) cub sort covers the next 1/RZ_SORT_BUCKETS of the buffer above it.  No host round trip, and the
// pass loop stays one uninterrupted stream of launches.  Slots between the count and the sorted size carry the unused key
// 0xffff (the caller clears the whole buffer), so the order of the live entries is the one a full sort gives.
#define RZ_SORT_BUCKETS 32

struct RzSortGraph {
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
};

__global__ void rz_sort_select_kernel(cudaGraphConditionalHandle h, const unsigned int *count, uint32_t cap) {
    const unsigned long long c = min(*count, cap);
    unsigned int j = c == 0ull ? 0u : (unsigned int)((c * RZ_SORT_BUCKETS + cap - 1ull) / cap) - 1u;
    cudaGraphSetConditional(h, min(j, (unsigned int)RZ_SORT_BUCKETS - 1u));
}

extern "C" void rz_sort_graph_destroy(RzSortGraph *g) {
    if (!g) return;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
}

// count: fixed device address the caller copies the queue's live count to before every launch.  cap: slots of the buffers.
extern "C" cudaError_t rz_sort_graph_create(RzSortGraph **out, const unsigned short *keys_in, unsigned short *keys_out, const uint32_t *iota,
                                            uint32_t *idx_out, uint32_t cap, void *temp, size_t temp_bytes, const unsigned int *count) {
    *out = nullptr;
    RzSortGraph *g = new RzSortGraph;
    cudaStream_t cs = nullptr;
    cudaError_t e = cudaSuccess;
    auto fail = [&](cudaError_t err) {
        if (cs) {
            cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(cs, &st) == cudaSuccess && st != cudaStreamCaptureStatusNone) { cudaGraph_t junk = nullptr; cudaStreamEndCapture(cs, &junk); }
            cudaStreamDestroy(cs);
        }
        rz_sort_graph_destroy(g);
        cudaGetLastError();
        return err;
    };
    if ((e = cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking)) != cudaSuccess) return fail(e);
    if ((e = cudaGraphCreate(&g->graph, 0)) != cudaSuccess) return fail(e);
    cudaGraphConditionalHandle h;
    if ((e = cudaGraphConditionalHandleCreate(&h, g->graph, RZ_SORT_BUCKETS - 1, cudaGraphCondAssignDefault)) != cudaSuccess) return fail(e);
    cudaGraphNode_t sel = nullptr;
    {
        void *args[3] = {(void *)&h, (void *)&count, (void *)&cap};
        cudaKernelNodeParams kp = {};
        kp.func = (void *)rz_sort_select_kernel;
        kp.gridDim = dim3(1); kp.blockDim = dim3(1); kp.sharedMemBytes = 0; kp.kernelParams = args; kp.extra = nullptr;
        if ((e = cudaGraphAddKernelNode(&sel, g->graph, nullptr, 0, &kp)) != cudaSuccess) return fail(e);
    }
    cudaGraphNodeParams cp = {};
    cp.type = cudaGraphNodeTypeConditional;
    cp.conditional.handle = h;
    cp.conditional.type = cudaGraphCondTypeSwitch;
    cp.conditional.size = RZ_SORT_BUCKETS;
    cudaGraphNode_t sw = nullptr;
    if ((e = cudaGraphAddNode(&sw, g->graph, &sel, 1, &cp)) != cudaSuccess) return fail(e);
    for (unsigned int j = 0; j < RZ_SORT_BUCKETS; j++) {
        const uint32_t n = (uint32_t)std::min<unsigned long long>(cap, ((unsigned long long)cap * (j + 1) + RZ_SORT_BUCKETS - 1) / RZ_SORT_BUCKETS);
        if ((e = cudaStreamBeginCaptureToGraph(cs, cp.conditional.phGraph_out[j], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed)) != cudaSuccess) return fail(e);
        e = cub::DeviceRadixSort::SortPairs(temp, temp_bytes, keys_in, keys_out, iota, idx_out, (int)n, 0, 16, cs);
        cudaGraph_t body = nullptr;
        const cudaError_t e2 = cudaStreamEndCapture(cs, &body);
        if (e != cudaSuccess) return fail(e);
        if (e2 != cudaSuccess) return fail(e2);
    }
    if ((e = cudaGraphInstantiate(&g->exec, g->graph, 0)) != cudaSuccess) return fail(e);
    cudaStreamDestroy(cs);
    *out = g;
    return cudaSuccess;
}

extern "C" cudaError_t rz_sort_graph_launch(RzSortGraph *g, cudaStream_t stream) { return cudaGraphLaunch(g->exec, stream); }
