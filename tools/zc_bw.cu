// Device -> host bandwidth by GPU-initiated stores into host-mapped pinned memory (zero copy) vs the copy engine,
// run concurrently on several GPUs of one box (one process per GPU, started together):
//   nvcc -O2 -gencode arch=compute_100a,code=sm_100a tools/zc_bw.cu -o build/zc_bw
//   for g in 0 1; do build/zc_bw $g $START_EPOCH_S & done; wait
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include <unistd.h>

__global__ void zc_store(uint4 *dst, const uint4 *src, size_t n) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int main(int argc, char **argv) {
    const int dev = argc > 1 ? atoi(argv[1]) : 0;
    const long start_at = argc > 2 ? atol(argv[2]) : 0;
    const size_t bytes = 256u << 20, n = bytes / 16;
    cudaSetDevice(dev);
    cudaSetDeviceFlags(cudaDeviceMapHost);
    void *h, *hd, *d;
    cudaHostAlloc(&h, bytes, cudaHostAllocMapped);
    cudaHostGetDevicePointer(&hd, h, 0);
    cudaMalloc(&d, bytes);
    cudaMemset(d, 1, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    zc_store<<<296, 256>>>((uint4 *)hd, (const uint4 *)d, n);
    cudaMemcpy(h, d, bytes, cudaMemcpyDeviceToHost);
    cudaDeviceSynchronize();
    while (start_at && time(nullptr) < start_at) usleep(1000);
    float ms;
    for (int grid = 74; grid <= 1184; grid *= 4) {
        cudaEventRecord(a);
        for (int r = 0; r < 8; ++r) zc_store<<<grid, 256>>>((uint4 *)hd, (const uint4 *)d, n);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        cudaEventElapsedTime(&ms, a, b);
        printf("gpu %d zero-copy stores, grid %4d: %.1f GB/s\n", dev, grid, 8.0 * bytes / ms / 1e6);
    }
    cudaEventRecord(a);
    for (int r = 0; r < 8; ++r) cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, 0);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    cudaEventElapsedTime(&ms, a, b);
    printf("gpu %d copy engine: %.1f GB/s\n", dev, 8.0 * bytes / ms / 1e6);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("gpu %d error %s\n", dev, cudaGetErrorString(e));
    return 0;
}
