// umma_rate.cu -- micro-benchmark: issue rate of small tcgen05.mma (kind::f16, M128, K16) on B200 as a function of N,
// the shared-memory layout (SWIZZLE_NONE / 32B / 64B / 128B K-major), accumulator rotation and M; plus the L2 / L1 /
// local-memory load-to-use latencies seen by a lone warp.  Timing only (operands are garbage).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}
__global__ void __launch_bounds__(128, 1) rate(int N, int M, int layout, int lbo, int sbo, int nrot, int shift, int iters, long long* out) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ __align__(8) uint64_t bar;
    __shared__ uint32_t tb;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x3c003c00u;   // 1.0 halfs
    if (tid == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bar))); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(s32(&tb)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tb;
    if (warp == 1) {
        const uint32_t idesc = (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
        const uint64_t hi = ((uint64_t)((sbo >> 4) & 0x3FFF) | (1ull << 14) | ((uint64_t)layout << 29)) << 32;
        const uint32_t a0 = (s32(sm) >> 4), b0 = (s32(sm + 48 * 1024) >> 4);
        const uint32_t lo_a = ((uint32_t)(lbo >> 4) & 0x3FFF) << 16, lo_b = lo_a;
        long long t0 = clock64();
        if (elect_one()) {
#pragma unroll 4
            for (int i = 0; i < iters; i++) {
                const uint32_t td = tmem + (uint32_t)(i % nrot) * 64u;
                const uint64_t ad = hi | (uint64_t)(lo_a | ((a0 + (uint32_t)((i % 9) * shift)) & 0x3FFF));
                const uint64_t bd = hi | (uint64_t)(lo_b | (b0 & 0x3FFF));
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(td), "l"(ad), "l"(bd), "r"(idesc), "r"(1u) : "memory");
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(s32(&bar)) : "memory");
        }
        __syncwarp();
        long long t1 = clock64();
        uint32_t ok = 0;
        while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(ok) : "r"(s32(&bar)) : "memory");
        long long t2 = clock64();
        if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
}
__global__ void chase(const int* p, int n, long long* out, int mode) {
    __shared__ int sidx[1024];
    int local_arr[64];
    for (int i = 0; i < 64; i++) local_arr[i] = (i * 17 + 1) & 63;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) sidx[i] = (i * 33 + 7) & 1023;
    __syncthreads();
    if (threadIdx.x == 0) {
        int j = 0;
        long long t0 = clock64();
        if (mode == 0) for (int i = 0; i < n; i++) j = __ldcg(p + j);            // L2 (bypass L1)
        else if (mode == 1) for (int i = 0; i < n; i++) j = __ldca(p + j);       // L1 allowed
        else if (mode == 2) for (int i = 0; i < n; i++) j = sidx[j];             // shared
        else for (int i = 0; i < n; i++) j = local_arr[j & 63];                  // local
        long long t1 = clock64();
        out[0] = (t1 - t0) / n; out[1] = j;
    }
}
__global__ void st_ld(int* buf, long long* out) {
    // 192 threads: everybody stores 16 B, barrier, everybody loads a neighbour's 16 B (ld.cg), barrier: one op boundary
    const int tid = threadIdx.x;
    int4* b4 = reinterpret_cast<int4*>(buf) + blockIdx.x * 4096;
    __syncthreads();
    long long t0 = clock64();
    int acc = 0;
    for (int it = 0; it < 200; it++) {
        b4[(tid + it) & 4095] = make_int4(it, tid, acc, 1);
        __syncthreads();
        const int4 v = __ldcg(b4 + ((tid + it + 37) & 4095));
        acc += v.x + v.y;
        __syncthreads();
    }
    long long t1 = clock64();
    if (tid == 0) { out[0] = (t1 - t0) / 200; out[1] = acc; }
}
int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
    cudaFuncSetAttribute(rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    struct { const char* name; int N, M, layout, lbo, sbo, nrot, shift; } cfg[] = {
        {"N16 none same-acc  (engine 2 today)", 16, 128, 0, 2624, 128, 1, 1},
        {"N16 none rotate-4", 16, 128, 0, 2624, 128, 4, 1},
        {"N32 none same-acc", 32, 128, 0, 2624, 128, 1, 1},
        {"N64 none same-acc", 64, 128, 0, 2624, 128, 1, 1},
        {"N64 none rotate-4", 64, 128, 0, 2624, 128, 4, 1},
        {"N16 none LBO=128 SBO=256 (interleaved 32B rows)", 16, 128, 0, 128, 256, 1, 2},
        {"N16 sw32  SBO=256", 16, 128, 6, 16, 256, 1, 2},
        {"N16 sw64  SBO=512", 16, 128, 4, 16, 512, 1, 4},
        {"N16 sw128 SBO=1024", 16, 128, 2, 16, 1024, 1, 8},
        {"N64 sw128 SBO=1024", 64, 128, 2, 16, 1024, 1, 8},
        {"N16 none M64", 16, 64, 0, 2624, 128, 1, 1},
        {"N128 sw128", 128, 128, 2, 16, 1024, 1, 8},
    };
    for (auto& c : cfg) {
        const int iters = 2000;
        rate<<<1, 128, 64 * 1024>>>(c.N, c.M, c.layout, c.lbo, c.sbo, c.nrot, c.shift, iters, d);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("[umma_rate] %-50s issue %6.1f clk/MMA, complete %6.1f clk/MMA (%s)\n", c.name, (double)h[0] / iters, (double)h[1] / iters, cudaGetErrorString(e));
    }
    int n = 1 << 22, *hp = (int*)malloc(4 * n), *dp;
    for (int i = 0; i < n; i++) hp[i] = (int)(((long long)i * 40503 + 12345) % n);
    cudaMalloc(&dp, 4 * n); cudaMemcpy(dp, hp, 4 * n, cudaMemcpyHostToDevice);
    const char* names[4] = {"L2 (ld.cg, 16 MB footprint)", "global via L1 (ld.ca)", "shared", "local array"};
    for (int mode = 0; mode < 4; mode++) {
        chase<<<1, 32>>>(dp, 20000, d, mode);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("[umma_rate] dependent-load latency, %-30s: %lld clk\n", names[mode], h[0]);
    }
    int* buf; cudaMalloc(&buf, 296 * 4096 * 16);
    for (int grid : {1, 148, 296}) {
        st_ld<<<grid, 192>>>(buf, d);
        cudaDeviceSynchronize();
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
        printf("[umma_rate] store -> barrier -> L2 load -> barrier round trip, %3d CTAs: %lld clk\n", grid, h[0]);
    }
    return 0;
}
