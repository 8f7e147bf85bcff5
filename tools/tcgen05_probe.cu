// tcgen05_probe.cu -- stand-alone check that the arena layout of the fused kernel (4-channel
// planes, 16 B per pixel, zero halo) IS a valid no-swizzle K-major UMMA operand, including the tap shift as a plain
// start-address offset.  Not part of the product: it de-risks the tcgen05 conv path planned for N >= 16 layers.
//
//   D[128 pixels, 16 couts] = A[128 pixels, 8 cins] * B[16 couts, 8 cins]^T      (kind::tf32, fp32 accumulate in TMEM)
//
// A lives in shared memory exactly like an activation tensor: plane 0 = channels 0-3, plane 1 = channels 4-7, pixel
// p of a plane at byte 16*p; the MMA reads pixels [shift, shift + 128).  Descriptor (K-major, SWIZZLE_NONE): core
// matrix = 8 pixels x 16 B = 128 contiguous bytes, SBO = 128 B (next 8 pixels), LBO = plane size (next 4 channels).
// B (weights) uses the same canonical layout: 8 couts x 16 B core matrices, SBO = 128 B, LBO = 256 B.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tcgen05_probe tools/tcgen05_probe.cu && ./tcgen05_probe
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define PPS 160          // pixels per plane in the probe (>= shift + 128)
#define NCOUT 16

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// shared-memory matrix descriptor, K-major, no swizzle (bit layout as in CUTLASS cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);              // start address, 16-byte units
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;    // leading-dimension byte offset (between the K core matrices)
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;    // stride byte offset (between 8-row groups)
    d |= (uint64_t)1 << 46;                              // descriptor version (Blackwell)
    return d;                                            // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE (0)
}

__global__ void __launch_bounds__(128, 1) probe(const float* A, const float* B, float* D, int shift, int* status) {
    __shared__ __align__(128) float sA[2 * PPS * 4];     // two planes, [plane][pixel][4 channels]
    __shared__ __align__(128) float sB[2 * NCOUT * 4];   // [k half][cout][4 cins]
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < 2 * PPS * 4; i += 128) sA[i] = A[i];
    for (int i = tid; i < 2 * NCOUT * 4; i += 128) sB[i] = B[i];
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {   // one warp allocates 32 TMEM columns
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 32;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    // generic-proxy writes to shared memory must be visible to the async proxy (the tensor core)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base;

    if (tid == 0) {
        const uint64_t adesc = make_desc(smem_u32(sA) + (uint32_t)shift * 16u, PPS * 16u, 128u);
        const uint64_t bdesc = make_desc(smem_u32(sB), NCOUT * 16u, 128u);
        // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major,
        // N >> 3 at bit 17, M >> 4 at bit 24
        const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(NCOUT >> 3) << 17) | ((128u >> 4) << 24);
        asm volatile(
            "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem),
            "l"(adesc), "l"(bdesc), "r"(idesc), "r"(0));
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar))
                     : "memory");
    }
    // bounded wait (a wrong descriptor must not hang the box)
    bool done = false;
    for (int spin = 0; spin < (1 << 22) && !done; spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(smem_u32(&bar))
            : "memory");
        done = ok != 0;
    }
    if (!done) {
        if (tid == 0) *status = -1;
    } else {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t r[16];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);   // lane in bits 31:16, column in 15:0
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
            : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
              "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int n = 0; n < NCOUT; n++) D[tid * NCOUT + n] = __uint_as_float(r[n]);
        if (tid == 0) *status = 1;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 32;" ::"r"(tmem));
}

int main() {
    const int nA = 2 * PPS * 4, nB = 2 * NCOUT * 4;
    float *hA = (float*)malloc(nA * 4), *hB = (float*)malloc(nB * 4), *hD = (float*)malloc(128 * NCOUT * 4);
    srand(1);
    for (int i = 0; i < nA; i++) hA[i] = (float)((rand() % 33) - 16) / 8.f;     // exactly representable in TF32
    for (int i = 0; i < nB; i++) hB[i] = (float)((rand() % 17) - 8) / 4.f;
    float *dA, *dB, *dD;
    int* dS;
    cudaMalloc(&dA, nA * 4); cudaMalloc(&dB, nB * 4); cudaMalloc(&dD, 128 * NCOUT * 4); cudaMalloc(&dS, 4);
    cudaMemcpy(dA, hA, nA * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB, nB * 4, cudaMemcpyHostToDevice);
    int bad_total = 0;
    for (int shift = 0; shift <= 19; shift += 19) {     // 0 and a tap-like shift of (1 row of 18 pixels + 1)
        int st = 0;
        cudaMemcpy(dS, &st, 4, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, 128 * NCOUT * 4);
        probe<<<1, 128>>>(dA, dB, dD, shift, dS);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(hD, dD, 128 * NCOUT * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        double maxerr = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < NCOUT; n++) {
                double ref = 0;
                for (int k = 0; k < 8; k++) {
                    const float a = hA[((k >> 2) * PPS + (m + shift)) * 4 + (k & 3)];
                    const float b = hB[((k >> 2) * NCOUT + n) * 4 + (k & 3)];
                    ref += (double)a * b;
                }
                const double err = fabs(ref - hD[m * NCOUT + n]);
                if (err > maxerr) maxerr = err;
                if (err > 1e-5) bad++;
            }
        printf("[tcgen05_probe] shift %2d: cuda=%s status=%d mismatches=%d max|err|=%.3g\n", shift, cudaGetErrorString(e),
               st, bad, maxerr);
        bad_total += bad + (st != 1) + (e != cudaSuccess);
    }
    printf("[tcgen05_probe] %s\n", bad_total ? "FAILED" : "OK: the arena layout is a valid no-swizzle K-major UMMA operand");
    return bad_total ? 1 : 0;
}
