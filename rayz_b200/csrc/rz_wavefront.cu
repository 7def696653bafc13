// rz_wavefront.cu — K2: wavefront variant (placeholder until the staged kernels land).
#include "rz_device.cuh"

extern "C" cudaError_t rz_wavefront_render(const RzPathArgs *, int, cudaStream_t, void **, size_t *, uint32_t *) {
    return cudaErrorNotSupported;
}
