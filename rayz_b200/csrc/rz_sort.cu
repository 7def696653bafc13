// rz_sort.cu — key/index radix sort between the primary and the second-segment kernel of the staged K1
// (cub::DeviceRadixSort: library plumbing, 16-bit keys => two 8-bit passes over 6 bytes per entry).
#include <cub/device/device_radix_sort.cuh>
#include <stdint.h>

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
