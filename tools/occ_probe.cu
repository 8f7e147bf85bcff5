// occ_probe.cu -- does a tcgen05.alloc in a kernel limit the resident CTAs per SM?  (build-time question for engine 2)
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int COLS>
__global__ void __launch_bounds__(192, 2) k(int* out, long long* t) {
    __shared__ uint32_t s_t;
    if (COLS > 0 && threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_t)), "n"(COLS > 0 ? COLS : 32));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    __syncthreads();
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    long long t0 = clock64();
    while (clock64() - t0 < 2000000) {}
    if (threadIdx.x == 0) { out[blockIdx.x] = (int)smid; t[blockIdx.x] = t0; }
    __syncthreads();
    if (COLS > 0 && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_t), "n"(COLS > 0 ? COLS : 32));
}
template <int COLS>
void run(const char* name) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<COLS>, 192, 0);
    int* d; long long* t;
    cudaMalloc(&d, 4 * 296); cudaMalloc(&t, 8 * 296);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<COLS><<<296, 192>>>(d, t);
    cudaEventRecord(e1);
    cudaError_t e = cudaDeviceSynchronize();
    float ms = 0; cudaEventElapsedTime(&ms, e0, e1);
    printf("[occ_probe] %s: occupancy API %d, 296 CTAs x ~1.02 ms spin took %.2f ms (%s)\n", name, occ, ms, cudaGetErrorString(e));
}
int main() { run<0>("no tcgen05"); run<128>("alloc 128 cols"); run<256>("alloc 256 cols"); run<512>("alloc 512 cols"); return 0; }
