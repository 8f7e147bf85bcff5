// sbc2_kernel.cuh -- engine 2: persistent annealed-Langevin kernel with the conv contractions on tcgen05 / TMEM.
//
// One CTA (192 threads, two CTAs per SM) owns a GROUP of S channel realisations and walks them in lock-step through
// the layer program of NCSNv2Deepest.forward (reference ncsnv2/models/ncsnv2.py:269-300) and then through the
// data-consistency gradient, Langevin update, Philox noise and NMSE of reference test_score.py:157-170, for every
// (sigma level, inner step) of the requested range.  Activations live in a per-CTA arena in global memory that stays
// L2 resident (layouts: sbc2_plan.h); they are never touched by another SM and never returned to the host.
//
// A conv op is a warp-specialised pipeline over 128-pixel M tiles of the S stacked images:
//   warp 4 (one lane): cp.async.bulk (TMA bulk copy) of the tile's input window, every 16-byte sub-plane, global ->
//                      shared staging ring, completion on an mbarrier; also prefetches the NEXT conv's parameter
//                      segment (UMMA list | bias | B tiles) into the other weight buffer
//   warp 5 (one lane): tcgen05.mma kind::f16 (M128 x N x K16, fp32 accumulate in TMEM) per (tap, channel chunk) straight
//                      off the staged window -- a tap is a start-address shift of the no-swizzle K-major descriptor --
//                      then tcgen05.commit to free the stage and to publish the accumulator slot
//   warps 0-3        : epilogue: tcgen05.ld of the accumulator (lane = pixel), hi + lo column groups summed, bias,
//                      residual accumulate, ELU, fp16 hi/lo split, coalesced 16-byte stores
// TMEM: 256 columns per CTA = 4 accumulator slots of 64 columns, so the MMAs of up to 4 tiles run ahead of the epilogue.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "sbc2_plan.h"
#include "sbc_ops.h"

#define SBC2_NTHR 192
#define SBC2_NACC 4
#define SBC2_TMEM_COLS 256
#define SBC2_MAXS 8
#define SBC2_NBARS 24
#define SBC2_PART_FLOATS 1024      // shared scratch of the norm statistics: 2 x 64 units x 8 channels

#define SBC2_MAX_OPS 1536
#define SBC2_MAX_MMA 4096
// Per-op constants of the conv pipeline and every conv's UMMA descriptor list, in constant memory (uploaded by
// sbc2_launch): indexed by warp-uniform values only, so the MMA issue loop runs on the uniform datapath (ULDC -> UIADD ->
// UTCHMMA) without vector -> uniform register moves.  x = first list entry (-1: the list is read from the parameter
// segment in shared memory instead), y = entries, z = instruction descriptor, w = staging ring depth.
__constant__ int4 sbc2_c_conv[SBC2_MAX_OPS];
__constant__ sbc2::Geo sbc2_c_geo[sbc2::MAX_LEVELS];     // level geometries of the plan in flight (dynamically indexed)
__constant__ uint2 sbc2_c_mma[SBC2_MAX_MMA];

struct Sbc2Launch {
    const sbc2::Op* ops;          // device copy of the layer program (records are prefetched into shared memory)
    int n_ops;
    const uint8_t* blob;
    uint8_t* gws;                 // [gridDim.x] group arenas
    long long arena_bytes;
    int S, B;
    int x_off, out_off, post_off;
    int first_w_off, first_w_len;
    int wmax, stage_bytes;        // shared-memory carve-up (bytes): two weight buffers of wmax, one staging ring
    int Nt, Nr, channels;
    int mode;                     // 0 = forward, 1 = annealed Langevin
    // forward mode
    const float* fx;
    long long fxs[4];
    const long long* labels;
    float* fout;
    const float* sigmas;
    int n_sigmas;
    // ALD mode (same meaning as engine 1 / include/sbc.h)
    int Np, level_begin, level_end, steps_each;
    const float* P;
    const float* Y;
    float* X;
    const float* Hor;
    const float* noise_var;
    const float* alpha_step;
    const float* beta;
    double sigma_end;
    float* nmse_log;
    unsigned long long seed;
    const unsigned long long* sample_ids;
    const float* ext_noise;
    const float* dc_boost;
    const int* stop_step;
    long long* prof;              // optional [n_ops + 2] clock64 stamps of CTA 0, first group, second step; when trace_op >= 0
                                  // it is followed by [3 roles][64 tiles][4] intra-conv stamps of that op
    int trace_op;
    int dbg;                      // timing experiments only (env SBC2_DBG; results are garbage): 1 no MMAs, 2 no bulk
                                  // copies, 4 no epilogue loads / stores, 8 no non-conv op bodies
};

namespace sbc2k {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a pipeline bug must surface as a CUDA error (trap), never as a hung GPU.  try_wait suspends the
// thread in hardware for a bounded time per call, so the limit below is tens of seconds.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    for (uint32_t spin = 0;; spin++) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 24)) __trap();
    }
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
// one lane of a converged warp (the pattern the compiler needs to keep descriptors in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// shared-memory matrix descriptor: K-major, SWIZZLE_NONE, version 1 (sm_100).  Core matrix = 8 rows x 16 bytes
// (128 contiguous bytes); SBO = byte distance between 8-row groups (M / N direction), LBO = between the two
// 16-byte K halves of a K = 16 (fp16) slice.  Validated on hardware by tools/tcgen05_probe.cu.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    return d;
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ELU (layers.py:13).  exp via MUFU.EX2: absolute error ~1e-7 = one ulp of the 1.0 that is subtracted (fp32 level).  A
// cancellation-free polynomial branch for small |v| was measured and dropped: it costs 11 % of engine 1's step time (the
// kernels are bound by exactly such dependent chains) and does not change the forward error (2.4e-6 either way).
__device__ __forceinline__ float elu(float v) { return v > 0.f ? v : __expf(v) - 1.f; }

struct alignas(16) H8 { __half2 a, b, c, d; };
// v[0..7] -> fp16 hi / lo pair (hi = rn(v), lo = rn(v - hi)); |v| is clamped below the fp16 overflow threshold
__device__ __forceinline__ void split8(const float (&v)[8], uint4& hi, uint4& lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const float x0 = fminf(fmaxf(v[2 * i], -65000.f), 65000.f), x1 = fminf(fmaxf(v[2 * i + 1], -65000.f), 65000.f);
        const __half2 hh = __floats2half2_rn(x0, x1);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
        h[i] = *reinterpret_cast<const uint32_t*>(&hh);
        l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}
__device__ __forceinline__ void store_sp16(uint8_t* base, int slot, int oct, int q, const float (&v)[8]) {
    uint4 hi, lo;
    split8(v, hi, lo);
    uint8_t* p = base + (size_t)(2 * oct) * slot + (size_t)q * 16;
    *reinterpret_cast<uint4*>(p) = hi;
    *reinterpret_cast<uint4*>(p + slot) = lo;
}
__device__ __forceinline__ void load_f32x8(const uint8_t* base, int slot, int oct, int q, float (&v)[8]) {
    const uint8_t* p = base + (size_t)(2 * oct) * slot + (size_t)q * 16;
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + slot);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void store_f32x8(uint8_t* base, int slot, int oct, int q, const float (&v)[8]) {
    uint8_t* p = base + (size_t)(2 * oct) * slot + (size_t)q * 16;
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + slot) = make_float4(v[4], v[5], v[6], v[7]);
}

// (sample, y, x) decomposition of a flat interior-pixel index; shifts when the sizes are powers of two
struct Dec {
    int w, hw, lw, lhw;
    __device__ __forceinline__ Dec(const sbc2::Geo& G) : w(G.w), hw(G.hw) {
        lw = (w & (w - 1)) ? -1 : (31 - __clz(w));
        lhw = (hw & (hw - 1)) ? -1 : (31 - __clz(hw));
    }
    __device__ __forceinline__ void yx(int e, int& y, int& x) const {
        y = lw >= 0 ? (e >> lw) : (e / w);
        x = e - y * w;
    }
    __device__ __forceinline__ void sex(int i, int& s, int& e) const {
        s = lhw >= 0 ? (i >> lhw) : (i / hw);
        e = i - s * hw;
    }
};
__device__ __forceinline__ int qof(const sbc2::Geo& G, int s, int y, int x) { return G.lead + s * G.pps + y * G.wp + x; }

__device__ __forceinline__ float inv_hw0(int hw) { return 1.f / (float)hw; }
// Sum over the warp of 8 per-lane values by a transposing butterfly: 7 + 2 shuffles instead of 40.  On return lane l holds
// in v[0] the warp total of value (l & 7) (lanes 0-7 are the ones the callers read).
__device__ __forceinline__ void warp_sum8(float (&v)[8]) {
    const int lane = threadIdx.x & 31;
    {   // step 1 (xor 4): lanes with bit 2 clear keep values 0-3, the others 4-7
        const bool up = lane & 4;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const float send = up ? v[i] : v[i + 4], keep = up ? v[i + 4] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
    }
    {   // step 2 (xor 2): keep 2 of the 4
        const bool up = lane & 2;
#pragma unroll
        for (int i = 0; i < 2; i++) {
            const float send = up ? v[i] : v[i + 2], keep = up ? v[i + 2] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, 2);
        }
    }
    {   // step 3 (xor 1): keep 1 of the 2
        const bool up = lane & 1;
        const float send = up ? v[0] : v[1], keep = up ? v[1] : v[0];
        v[0] = keep + __shfl_xor_sync(0xffffffffu, send, 1);
    }
    // lane l now holds the partial sum of value (l & 7) over the 4 lanes {l ^ 0, ...} of its group of 8; finish over 8, 16
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 8);
    v[0] += __shfl_xor_sync(0xffffffffu, v[0], 16);
}

// interior-pixel number s*hw + y*w + x of padded pixel q, or -1 for pads (exact magic-number divisions)
__device__ __forceinline__ int pix_of(const sbc2::Geo& G, int S, int q) {
    const uint32_t n = (uint32_t)(q - G.lead);
    const uint32_t s = __umulhi(n, G.mg_pps), r = n - s * (uint32_t)G.pps;
    const uint32_t y = G.wp == 1 ? r : __umulhi(r, G.mg_wp), x = r - y * (uint32_t)G.wp;     // ceil(2^32 / 1) does not fit 32 bits
    return (q >= G.lead && s < (uint32_t)S && y < (uint32_t)G.h && x < (uint32_t)G.w) ? (int)(s * G.hw + y * G.w + x) : -1;
}

// ---------------------------------------------------------------------------------------------------------
// non-conv ops: all 192 threads, items = (sample, channel octet, pixel), pixel fastest (coalesced 16-byte accesses).
// Activations come from L2 (~700 clk): every loop first issues the loads of U items, then computes and stores.
// ---------------------------------------------------------------------------------------------------------
struct Item { int oct, s, y, x, q; bool ok; };
__device__ __forceinline__ Item item_of(const sbc2::Geo& G, const Dec& D, int i, int n, int per) {
    Item it;
    it.ok = i < n;
    const int ii = it.ok ? i : 0;
    it.oct = ii / per;
    const int r = ii - it.oct * per;
    int e;
    D.sex(r, it.s, e);
    D.yx(e, it.y, it.x);
    it.q = qof(G, it.s, it.y, it.x);
    return it;
}

__device__ __forceinline__ void op_affine(const sbc2::Op& op, const int S_, uint8_t* arena, int tid) {
    const sbc2::Geo& G = sbc2_c_geo[0];
    const Dec D(G);
    const float2* xin = reinterpret_cast<const float2*>(arena + op.src0);
    const int n = S_ * G.hw;
    constexpr int U = 2;
    for (int i0 = tid; i0 < n; i0 += SBC2_NTHR * U) {
        float2 c[U];
#pragma unroll
        for (int u = 0; u < U; u++) { const int i = i0 + u * SBC2_NTHR; c[u] = xin[i < n ? i : 0]; }
#pragma unroll
        for (int u = 0; u < U; u++) {
            const int i = i0 + u * SBC2_NTHR;
            if (i >= n) continue;
            int s, e, y, x;
            D.sex(i, s, e);
            D.yx(e, y, x);
            float v[8] = {2.f * c[u].x - 1.f, 2.f * c[u].y - 1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // ncsnv2.py:270-271
            store_sp16(arena + op.raw16, G.slot, 0, qof(G, s, y, x), v);
        }
    }
}

__device__ __forceinline__ void op_elu(const sbc2::Op& op, const int S_, uint8_t* arena, int tid) {
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const Dec D(G);
    const int noct = op.cin >> 3, per = S_ * G.hw, n = noct * per;
    constexpr int U = 2;
    for (int i0 = tid; i0 < n; i0 += SBC2_NTHR * U) {
        Item it[U];
        float v[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
            it[u] = item_of(G, D, i0 + u * SBC2_NTHR, n, per);
            load_f32x8(arena + op.src0, G.slot, it[u].oct, it[u].q, v[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!it[u].ok) continue;
#pragma unroll
            for (int k = 0; k < 8; k++) v[u][k] = elu(v[u][k]);
            store_sp16(arena + op.elu16, G.slot, it[u].oct, it[u].q, v[u]);
        }
    }
}

// MaxPool2d(5, 1, 2) with -inf padding (layers.py:70): F32 in, SP16 out.  Clamped (replicated) coordinates give the
// same maximum as -inf padding and make all 25 taps unconditional, independent loads.
__device__ __forceinline__ void op_maxpool5(const sbc2::Op& op, const int S_, uint8_t* arena, int tid) {
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const Dec D(G);
    const int noct = op.cin >> 3, per = S_ * G.hw, n = noct * per;
    const uint8_t* src = arena + op.src0;
    for (int i = tid; i < n; i += SBC2_NTHR) {
        const Item it = item_of(G, D, i, n, per);
        int xo[5], yo[5];
#pragma unroll
        for (int d = 0; d < 5; d++) {
            xo[d] = min(max(it.x + d - 2, 0), G.w - 1);
            yo[d] = G.lead + it.s * G.pps + min(max(it.y + d - 2, 0), G.h - 1) * G.wp;
        }
        float m[8];
#pragma unroll
        for (int k = 0; k < 8; k++) m[k] = -INFINITY;
#pragma unroll
        for (int dy = 0; dy < 5; dy++) {
            float v[5][8];
#pragma unroll
            for (int dx = 0; dx < 5; dx++) load_f32x8(src, G.slot, it.oct, yo[dy] + xo[dx], v[dx]);
#pragma unroll
            for (int dx = 0; dx < 5; dx++)
#pragma unroll
                for (int k = 0; k < 8; k++) m[k] = fmaxf(m[k], v[dx][k]);
        }
        store_sp16(arena + op.raw16, G.slot, it.oct, it.q, m);
    }
}

// acc += bilinear(src, align_corners=True) (layers.py:182-183); optional elu32 = ELU(acc)
__device__ __forceinline__ void op_upacc(const sbc2::Op& op, const int S_, uint8_t* arena, int tid) {
    const sbc2::Geo& GS = sbc2_c_geo[op.gs];
    const sbc2::Geo& GD = sbc2_c_geo[op.gd];
    const Dec D(GD);
    const int H = GS.h, W = GS.w, OH = GD.h, OW = GD.w;
    const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    const int noct = op.cin >> 3, per = S_ * GD.hw, n = noct * per;
    constexpr int U = 1;
    for (int i0 = tid; i0 < n; i0 += SBC2_NTHR * U) {
        Item it[U];
        float p00[U][8], p01[U][8], p10[U][8], p11[U][8], a[U][8], ly[U], lx[U];
#pragma unroll
        for (int u = 0; u < U; u++) {
            it[u] = item_of(GD, D, i0 + u * SBC2_NTHR, n, per);
            const float fy = sy * (float)it[u].y, fx = sx * (float)it[u].x;
            const int y0 = (int)fy, x0 = (int)fx;
            const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
            ly[u] = fy - (float)y0; lx[u] = fx - (float)x0;
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, qof(GS, it[u].s, y0, x0), p00[u]);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, qof(GS, it[u].s, y0, x1), p01[u]);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, qof(GS, it[u].s, y1, x0), p10[u]);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, qof(GS, it[u].s, y1, x1), p11[u]);
            load_f32x8(arena + op.acc32, GD.slot, it[u].oct, it[u].q, a[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!it[u].ok) continue;
            const float hy = 1.f - ly[u], hx = 1.f - lx[u];
#pragma unroll
            for (int k = 0; k < 8; k++)
                a[u][k] += hy * (hx * p00[u][k] + lx[u] * p01[u][k]) + ly[u] * (hx * p10[u][k] + lx[u] * p11[u][k]);
            store_f32x8(arena + op.acc32, GD.slot, it[u].oct, it[u].q, a[u]);
            if (op.elu32 >= 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) a[u][k] = elu(a[u][k]);
                store_f32x8(arena + op.elu32, GD.slot, it[u].oct, it[u].q, a[u]);
            }
        }
    }
}

// epilogue of a conv whose input channels were split over several ops (wide nets): the finished sum `src0` gets the
// outputs a single conv op would have produced (dst32 = v; v += acc32; raw16 = split(v); elu16 / elu32 = ELU(v))
__device__ __forceinline__ void op_epilogue(const sbc2::Op& op, const int S_, uint8_t* arena, int tid) {
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const Dec D(G);
    const int noct = op.cin >> 3, per = S_ * G.hw, n = noct * per;
    for (int i = tid; i < n; i += SBC2_NTHR) {
        const Item it = item_of(G, D, i, n, per);
        float v[8];
        load_f32x8(arena + op.src0, G.slot, it.oct, it.q, v);
        if (op.dst32 >= 0) store_f32x8(arena + op.dst32, G.slot, it.oct, it.q, v);
        if (op.acc32 >= 0) {
            float o[8];
            load_f32x8(arena + op.acc32, G.slot, it.oct, it.q, o);
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] += o[k];
            store_f32x8(arena + op.acc32, G.slot, it.oct, it.q, v);
        }
        if (op.raw16 >= 0) store_sp16(arena + op.raw16, G.slot, it.oct, it.q, v);
        if (op.elu16 >= 0 || op.elu32 >= 0) {
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = elu(v[k]);
            if (op.elu16 >= 0) store_sp16(arena + op.elu16, G.slot, it.oct, it.q, v);
            if (op.elu32 >= 0) store_f32x8(arena + op.elu32, G.slot, it.oct, it.q, v);
        }
    }
}

// 2x2 mean-pool of ConvMeanPool (layers.py:309-313): F32 level g -> F32 level g+1 (+ optional raw SP16)
__device__ __forceinline__ void op_pool2(const sbc2::Op& op, const int S_, uint8_t* arena, int tid) {
    const sbc2::Geo& GS = sbc2_c_geo[op.gs];
    const sbc2::Geo& GD = sbc2_c_geo[op.gd];
    const Dec D(GD);
    const int noct = op.cin >> 3, per = S_ * GD.hw, n = noct * per;
    constexpr int U = 1;
    for (int i0 = tid; i0 < n; i0 += SBC2_NTHR * U) {
        Item it[U];
        float a[U][8], b[U][8], c[U][8], d[U][8];
#pragma unroll
        for (int u = 0; u < U; u++) {
            it[u] = item_of(GD, D, i0 + u * SBC2_NTHR, n, per);
            const int q0 = qof(GS, it[u].s, 2 * it[u].y, 2 * it[u].x);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, q0, a[u]);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, q0 + GS.wp, b[u]);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, q0 + 1, c[u]);
            load_f32x8(arena + op.src0, GS.slot, it[u].oct, q0 + GS.wp + 1, d[u]);
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!it[u].ok) continue;
            float v[8];
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = (a[u][k] + b[u][k] + c[u][k] + d[u][k]) * 0.25f;   // order of layers.py:311-312
            store_f32x8(arena + op.dst32, GD.slot, it[u].oct, it[u].q, v);
            if (op.raw16 >= 0) store_sp16(arena + op.raw16, GD.slot, it[u].oct, it[u].q, v);
        }
    }
}

// InstanceNorm2dPlus + ELU (normalization.py:163-176): F32 in, SP16 out.  Statistics: work unit = (sample, octet, chunk
// of <= 256 pixels), one warp per unit, all 8 loads of a lane in flight at once; two passes (mean, centred squares),
// partial sums exchanged through shared memory.
// general InstanceNorm++ path: three sweeps over the tensor (statistics exchanged through shared / global scratch)
// InstanceNorm2dPlus + ELU (normalization.py:163-176): F32 in, SP16 out.  Statistics: work unit = (sample, octet, chunk
// of <= 256 pixels), one warp per unit, all 8 loads of a lane in flight at once; two passes (mean, centred squares),
// partial sums exchanged through shared memory.
// general InstanceNorm++ path: three sweeps over the tensor (statistics exchanged through shared / global scratch)
__device__ __noinline__ void op_norm_elu_general(const sbc2::Op& op, const int S_, const uint8_t* blob, uint8_t* arena, float* spart, int tid) {
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const Dec D(G);
    const int C = op.cin, noct = C >> 3, S = S_, hw = G.hw;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = SBC2_NTHR / 32;
    const float* wseg = reinterpret_cast<const float*>(blob + op.w_off);
    const int n_items = S * noct;
    // shared scratch (1024 floats): small problems keep the statistics and coefficients on chip as well
    const int nchunk = (hw + 255) >> 8;
    const bool small = S * C <= 128 && n_items * nchunk <= 32;
    float* stats = small ? spart + 512 : reinterpret_cast<float*>(arena + op.scratch);       // [S][C][2] = (mean, M2)
    float* coef = small ? spart + 768 : stats + (size_t)S * C * 2;                            // [S][C][2] = (scale, shift)
    const int cap = small ? 32 : 64;
    const int batch_items = max(1, min(n_items, cap / nchunk));
    const float inv_hw = 1.f / (float)hw;
    float* part1 = spart;                 // [cap units][8]
    float* part2 = spart + (small ? 256 : 512);
    for (int it0 = 0; it0 < n_items; it0 += batch_items) {
        const int nit = min(batch_items, n_items - it0), nunits = nit * nchunk;
        for (int pass = 0; pass < 2; pass++) {
            for (int u = warp; u < nunits; u += NW) {
                const int li = u / nchunk, ch = u - li * nchunk, item = it0 + li;
                const int s = item / noct, oct = item - s * noct;
                const int e0 = ch << 8, e1 = min(hw, e0 + 256);
                float mean[8];
#pragma unroll
                for (int k = 0; k < 8; k++) mean[k] = 0.f;
                if (pass == 1) {
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        float t = 0.f;
                        for (int c2 = 0; c2 < nchunk; c2++) t += part1[(li * nchunk + c2) * 8 + k];
                        mean[k] = t * inv_hw;
                    }
                }
                float v[8][8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const int e = min(e0 + lane + 32 * j, e1 - 1);
                    int y, x;
                    D.yx(e, y, x);
                    load_f32x8(arena + op.src0, G.slot, oct, qof(G, s, y, x), v[j]);
                }
                float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const bool ok = e0 + lane + 32 * j < e1;
#pragma unroll
                    for (int k = 0; k < 8; k++) {
                        const float d = v[j][k] - mean[k];
                        const float t = pass == 0 ? d : d * d;
                        a[k] += ok ? t : 0.f;
                    }
                }
                warp_sum8(a);
                if (lane < 8) (pass == 0 ? part1 : part2)[u * 8 + lane] = a[0];
            }
            __syncthreads();
        }
        // per (item, channel): stats[s][c] = (mean, M2)
        for (int i = tid; i < nit * 8; i += SBC2_NTHR) {
            const int li = i >> 3, k = i & 7, item = it0 + li;
            const int s = item / noct, oct = item - s * noct;
            float t1 = 0.f, t2 = 0.f;
            for (int c2 = 0; c2 < nchunk; c2++) { t1 += part1[(li * nchunk + c2) * 8 + k]; t2 += part2[(li * nchunk + c2) * 8 + k]; }
            stats[(s * C + oct * 8 + k) * 2] = t1 * inv_hw;
            stats[(s * C + oct * 8 + k) * 2 + 1] = t2;
        }
        __syncthreads();
    }
    // cross-channel statistics of the per-channel means (torch.mean / unbiased torch.var over C) and the affine:
    // out = ELU(x * cs + csh)
    for (int j = tid; j < S * C; j += SBC2_NTHR) {
        const int s = j / C, c = j - s * C;
        float m = 0.f;
        for (int k = 0; k < C; k++) m += stats[(s * C + k) * 2];
        m /= (float)C;
        float v = 0.f;
        for (int k = 0; k < C; k++) { const float d = stats[(s * C + k) * 2] - m; v = fmaf(d, d, v); }
        v /= (float)(C - 1);
        const float mean = stats[j * 2], m2 = stats[j * 2 + 1];
        const float al = wseg[c], ga = wseg[C + c], be = wseg[2 * C + c];
        const float cs = ga * rsqrtf(m2 * inv_hw + 1e-5f);
        const float csh = fmaf(ga, (mean - m) * rsqrtf(v + 1e-5f) * al, be);
        coef[j * 2] = cs;
        coef[j * 2 + 1] = fmaf(-mean, cs, csh);
    }
    __syncthreads();
    const int per = S * hw, n = noct * per;
    constexpr int U = 4;
    for (int i0 = tid; i0 < n; i0 += SBC2_NTHR * U) {
        Item it[U];
        float v[U][8];
        float4 cf[U][4];
#pragma unroll
        for (int u = 0; u < U; u++) {
            it[u] = item_of(G, D, i0 + u * SBC2_NTHR, n, per);
            load_f32x8(arena + op.src0, G.slot, it[u].oct, it[u].q, v[u]);
            const float4* cp = reinterpret_cast<const float4*>(coef + (it[u].s * C + it[u].oct * 8) * 2);
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) cf[u][k2] = cp[k2];
        }
#pragma unroll
        for (int u = 0; u < U; u++) {
            if (!it[u].ok) continue;
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) {
                v[u][2 * k2] = elu(fmaf(v[u][2 * k2], cf[u][k2].x, cf[u][k2].y));
                v[u][2 * k2 + 1] = elu(fmaf(v[u][2 * k2 + 1], cf[u][k2].z, cf[u][k2].w));
            }
            store_sp16(arena + op.elu16, G.slot, it[u].oct, it[u].q, v[u]);
        }
    }
}


#define SBC2_NT(k) do { if (ntr) ntr[k] = clock64(); } while (0)
__device__ __forceinline__ void op_norm_elu(const sbc2::Op& op, const int S_, const uint8_t* blob, uint8_t* arena, float* spart, int tid, long long* ntr) {
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const Dec D(G);
    const int C = op.cin, noct = C >> 3, S = S_, hw = G.hw;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = SBC2_NTHR / 32;
    const float* wseg = reinterpret_cast<const float*>(blob + op.w_off);
    const int n_items = S * noct;
    // single-sweep layout: the warps are dealt evenly over the (sample, octet) items, each warp keeps its chunk of pixels
    // (<= JMAX per lane) in registers across the two statistics phases and the normalisation
    constexpr int JMAX = 6;
    const int wpi = n_items <= NW ? NW / n_items : 0;                          // warps per item
    const int nck = wpi > 0 ? min(wpi, (hw + 31) >> 5) : 0;                    // chunks per item actually used
    const int csz = nck > 0 ? (((hw + nck - 1) / nck + 31) & ~31) : 0;         // pixels per chunk (multiple of 32)
    if (nck > 0 && csz <= 32 * JMAX && S * C <= 128) {
        float* part1 = spart;              // [units][8] sums
        float* part2 = spart + 64;         // [units][8] centred squares
        float* coefs = spart + 128;        // [S][C][2]
        const int nchunk = nck;
        const int nunits = n_items * nchunk;
        const bool act = warp < nunits;
        const int li = act ? warp / nchunk : 0, ch = act ? warp - li * nchunk : 0;
        const int s = li / noct, oct = li - s * noct;
        const int e0 = ch * csz, e1 = min(hw, e0 + csz);
        float al = 0.f, ga = 0.f, be = 0.f;            // affine parameters of channel tid (loaded early: L2 latency)
        if (tid < S * C) { const int c = tid % C; al = wseg[c]; ga = wseg[C + c]; be = wseg[2 * C + c]; }
        SBC2_NT(0);
        float v[JMAX][8];
        int qq[JMAX];
        if (act) {
#pragma unroll
            for (int j = 0; j < JMAX; j++) {
                const int e = min(e0 + lane + 32 * j, e1 - 1);
                int y, x;
                D.yx(e, y, x);
                qq[j] = qof(G, s, y, x);
                load_f32x8(arena + op.src0, G.slot, oct, qq[j], v[j]);
            }
            float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < JMAX; j++) {
                const bool ok = e0 + lane + 32 * j < e1;
#pragma unroll
                for (int k = 0; k < 8; k++) a[k] += ok ? v[j][k] : 0.f;
            }
            SBC2_NT(1);
            warp_sum8(a);
            SBC2_NT(2);
            if (lane < 8) part1[warp * 8 + lane] = a[0];
        }
        __syncthreads();
        SBC2_NT(3);
        float* cmean = spart + 384;        // [S][C] per-channel means
        if (tid < S * C) {
            const int it = tid >> 3, k = tid & 7;       // item = (sample, octet): tid = (s * noct + oct) * 8 + k = s * C + c
            float t = 0.f;
            for (int c2 = 0; c2 < nchunk; c2++) t += part1[(it * nchunk + c2) * 8 + k];
            cmean[tid] = t * inv_hw0(hw);
        }
        __syncthreads();
        SBC2_NT(4);
        if (act) {
            float mean[8];
#pragma unroll
            for (int k = 0; k < 8; k++) mean[k] = cmean[li * 8 + k];
            float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int j = 0; j < JMAX; j++) {
                const bool ok = e0 + lane + 32 * j < e1;
#pragma unroll
                for (int k = 0; k < 8; k++) { const float d = v[j][k] - mean[k]; a[k] += ok ? d * d : 0.f; }
            }
            warp_sum8(a);
            if (lane < 8) part2[warp * 8 + lane] = a[0];
        }
        __syncthreads();
        SBC2_NT(5);
        if (tid < S * C) {     // cross-channel statistics of the per-channel means + the affine: out = ELU(x * cs + csh)
            const int ss = tid / C;
            const float ih = inv_hw0(hw);
            const float* cm = cmean + ss * C;
            float m0 = 0.f, m1 = 0.f, m2s = 0.f, m3 = 0.f;
            for (int k = 0; k < C; k += 4) { m0 += cm[k]; m1 += cm[k + 1]; m2s += cm[k + 2]; m3 += cm[k + 3]; }
            const float m = ((m0 + m1) + (m2s + m3)) / (float)C;
            float v0 = 0.f, v1 = 0.f, v2 = 0.f, v3 = 0.f;
            for (int k = 0; k < C; k += 4) {
                const float d0 = cm[k] - m, d1 = cm[k + 1] - m, d2 = cm[k + 2] - m, d3 = cm[k + 3] - m;
                v0 = fmaf(d0, d0, v0); v1 = fmaf(d1, d1, v1); v2 = fmaf(d2, d2, v2); v3 = fmaf(d3, d3, v3);
            }
            const float vv = ((v0 + v1) + (v2 + v3)) / (float)(C - 1);
            const float mean = cmean[tid];
            const int it = tid >> 3, k = tid & 7;
            float m2 = 0.f;
            for (int c2 = 0; c2 < nchunk; c2++) m2 += part2[(it * nchunk + c2) * 8 + k];
            const float cs = ga * rsqrtf(m2 * ih + 1e-5f);
            const float csh = fmaf(ga, (mean - m) * rsqrtf(vv + 1e-5f) * al, be);
            coefs[tid * 2] = cs;
            coefs[tid * 2 + 1] = fmaf(-mean, cs, csh);
        }
        __syncthreads();
        SBC2_NT(6);
        if (act) {
            float4 cf[4];
            const float4* cp = reinterpret_cast<const float4*>(coefs + (s * C + oct * 8) * 2);
#pragma unroll
            for (int k2 = 0; k2 < 4; k2++) cf[k2] = cp[k2];
#pragma unroll
            for (int j = 0; j < JMAX; j++) {
                if (e0 + lane + 32 * j >= e1) continue;
                float o[8];
#pragma unroll
                for (int k2 = 0; k2 < 4; k2++) {
                    o[2 * k2] = elu(fmaf(v[j][2 * k2], cf[k2].x, cf[k2].y));
                    o[2 * k2 + 1] = elu(fmaf(v[j][2 * k2 + 1], cf[k2].z, cf[k2].w));
                }
                store_sp16(arena + op.elu16, G.slot, oct, qq[j], o);
            }
        }
        SBC2_NT(7);
        return;
    }
    op_norm_elu_general(op, S_, blob, arena, spart, tid);
}

// ---------------------------------------------------------------------------------------------------------
// conv op (warp-specialised; see the header comment)
// ---------------------------------------------------------------------------------------------------------
struct Pipe {
    uint32_t sfull_k, sempty_k;     // next parity per staging barrier (bit b)
    uint32_t tfull_k, tempty_k;     // next parity per accumulator barrier (bit a)
    uint32_t conv_n;                // convs executed so far (weight buffer = conv_n & 1)
};

#define SBC2_TR(role, t, k) do { if (tr && (t) < 64) tr[((role) * 64 + (t)) * 4 + (k)] = clock64(); } while (0)

// epilogue of one conv for the 4 epilogue warps, NCH = cout8 / 8 channel octets per pixel (compile-time: no runtime
// indexed register arrays).  Warp w owns TMEM lanes 32w .. 32w+31 = pixels m0 + 32w + lane.
template <int NCH>
__device__ __forceinline__ uint32_t conv_epilogue(const sbc2::Op& op, const int S_, const int dbg, uint8_t* arena, const uint8_t* wb,
                                              uint64_t* tfull, uint64_t* tempty, uint32_t tmem, uint32_t tfull_k, int T,
                                              int npart, int span, int warp, int lane, long long* tr) {
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const float* bias = op.bias_rel >= 0 ? reinterpret_cast<const float*>(wb + op.bias_rel) : nullptr;
    const int slot = G.slot, cout8 = NCH * 8;
    const float us = op.unscale;
    const int flags = op.flags, dst32 = op.dst32, acc32 = op.acc32, raw16 = op.raw16, elu16 = op.elu16, elu32 = op.elu32;
    const bool has_acc = acc32 >= 0 && !(dbg & 4);
    const int qbase = G.lead + warp * 32 + lane;
    float accn[NCH][8];                           // residual values of the tile ahead (issued one tile early: L2 latency)
    if (has_acc) {
#pragma unroll
        for (int c = 0; c < NCH; c++) load_f32x8(arena + acc32, slot, c, qbase, accn[c]);
    }
    for (int t = 0; t < T; t++) {
        const uint32_t a = (uint32_t)(t * span) & (SBC2_NACC - 1);
        const int q = qbase + t * sbc2::TILE_M;
        const int px = pix_of(G, S_, q);
        if (warp == 0) SBC2_TR(2, t, 0);
        mbar_wait(&tfull[a], (tfull_k >> a) & 1u);
        tfull_k ^= 1u << a;
        if (warp == 0) SBC2_TR(2, t, 1);
        tc_fence_after();
        const uint32_t taddr = tmem + a * 64u + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c = 0; c < NCH; c++) {
            float hi[8], lo[8], v[8];
            tmem_ld8(taddr + (uint32_t)(c * 8), hi);
            tmem_ld8(taddr + (uint32_t)(cout8 + c * 8), lo);
            tmem_ld_wait();
            for (int p = 1; p < npart; p++) {     // split-K partial accumulators (independent MMA chains)
                float h2[8], l2[8];
                tmem_ld8(taddr + (uint32_t)(p * 2 * cout8 + c * 8), h2);
                tmem_ld8(taddr + (uint32_t)(p * 2 * cout8 + cout8 + c * 8), l2);
                tmem_ld_wait();
#pragma unroll
                for (int k = 0; k < 8; k++) { hi[k] += h2[k]; lo[k] += l2[k]; }
            }
            if (c == NCH - 1) {                   // every column of this slot is in registers: hand it back to the MMA warp
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[a]);
            }
            if (dbg & 4) continue;
#pragma unroll
            for (int k = 0; k < 8; k++) v[k] = fmaf(hi[k] + lo[k], us, bias ? bias[c * 8 + k] : 0.f);
            if (flags & sbc2::F_COMPACT) {     // network output: couts (0,1) = (re, im) of element px
                if (px >= 0) reinterpret_cast<float2*>(arena + dst32)[px] = make_float2(v[0], v[1]);
                continue;
            }
            if (px < 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = 0.f;      // pads of every output stay zero
            }
            if (dst32 >= 0) store_f32x8(arena + dst32, slot, c, q, v);     // the raw conv result
            if (acc32 >= 0) {
                if (px >= 0) {
#pragma unroll
                    for (int k = 0; k < 8; k++) v[k] += accn[c][k];
                }
                store_f32x8(arena + acc32, slot, c, q, v);
            }
            if (raw16 >= 0) store_sp16(arena + raw16, slot, c, q, v);
            if (elu16 >= 0 || elu32 >= 0) {
#pragma unroll
                for (int k = 0; k < 8; k++) v[k] = elu(v[k]);
                if (elu16 >= 0) store_sp16(arena + elu16, slot, c, q, v);
                if (elu32 >= 0) store_f32x8(arena + elu32, slot, c, q, v);
            }
        }
        if (has_acc && t + 1 < T) {
#pragma unroll
            for (int c = 0; c < NCH; c++) load_f32x8(arena + acc32, slot, c, q + sbc2::TILE_M, accn[c]);
        }
        if (warp == 0) SBC2_TR(2, t, 2);
    }
    return tfull_k;
}

// i = op index (warp-uniform: the per-op constants in sbc2_c_conv are read through it)
struct ConvEnv { const uint8_t* blob; int wmax, dbg, S; };     // the launch constants a conv needs
// The pipeline state crosses the call boundary packed in ONE register (structs travel through the stack, and local
// memory is an L2 round trip here): bits 0-3 sfull_k, 4-7 sempty_k, 8-11 tfull_k, 12-15 tempty_k, 16-31 conv_n.
__device__ __forceinline__ uint32_t pipe_pack(const Pipe& P) {
    return (P.sfull_k & 15u) | ((P.sempty_k & 15u) << 4) | ((P.tfull_k & 15u) << 8) | ((P.tempty_k & 15u) << 12) | (P.conv_n << 16);
}
__device__ __forceinline__ Pipe pipe_unpack(uint32_t v) {
    Pipe P;
    P.sfull_k = v & 15u; P.sempty_k = (v >> 4) & 15u; P.tfull_k = (v >> 8) & 15u; P.tempty_k = (v >> 12) & 15u; P.conv_n = v >> 16;
    return P;
}
__device__ __forceinline__ uint32_t op_conv(const int i, const sbc2::Op& op, const uint8_t* blob, const int wmax, const int dbg_, const int S_,
                                         uint8_t* arena, uint8_t* smem, uint64_t* bars, uint32_t tmem, uint32_t pipe, int tid,
                                         bool prefetch_next, long long* trace) {
    Pipe P = pipe_unpack(pipe);
    const ConvEnv L{blob, wmax, dbg_, S_};
    long long* tr = (trace && (tid & 31) == 0) ? trace : nullptr;
    uint64_t* sfull = bars;            // [4]
    uint64_t* sempty = bars + 4;       // [4]
    uint64_t* tfull = bars + 8;        // [4]
    uint64_t* tempty = bars + 12;      // [4]
    uint64_t* wfull = bars + 16;       // [2]
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t wslot = P.conv_n & 1u, wpar = (P.conv_n >> 1) & 1u;
    uint8_t* wb = smem + (size_t)wslot * L.wmax;
    uint8_t* stage = smem + 2 * (size_t)L.wmax;
    const sbc2::Geo& G = sbc2_c_geo[op.gs];
    const int T = G.T;                                  // == op.T
    const int4 cc = sbc2_c_conv[i];                     // uniform: list base, entries, idesc, ring depth
    const int ns = cc.w & 0xFF, npart = (cc.w >> 8) & 0xFF, span = (cc.w >> 16) & 0xFF;
    const int nsub0 = op.nsub0, nsub1 = op.nsub1, sps = op.sps;
    const uint32_t stage_bytes = (uint32_t)(nsub0 + nsub1) * (uint32_t)sps;

    if (warp == 4) {
        // ---------------- producer: the whole warp runs the loop (converged), one elected lane issues ----------------
        if (prefetch_next) {   // the other buffer held the previous conv's segment: that conv is complete
            uint64_t* wn = &wfull[wslot ^ 1u];
            if (elect_one()) {
                mbar_expect_tx(wn, (uint32_t)op.nw_len);
                bulk_g2s(smem + (size_t)(wslot ^ 1u) * L.wmax, L.blob + op.nw_off, (uint32_t)op.nw_len, wn);
            }
        }
        const uint8_t* g0 = arena + op.src0 + (size_t)(G.lead - op.halo) * 16;
        const uint8_t* g1 = arena + op.src1 + (size_t)(G.lead - op.halo) * 16;
        const int slot = G.slot;
        for (int t = 0; t < T; t++) {
            const int b = t & (ns - 1);
            SBC2_TR(0, t, 0);
            mbar_wait(&sempty[b], (P.sempty_k >> b) & 1u);
            SBC2_TR(0, t, 1);
            P.sempty_k ^= 1u << b;
            uint8_t* sb = stage + (size_t)b * stage_bytes;
            const size_t toff = (size_t)t * (sbc2::TILE_M * 16);
            if (L.dbg & 2) {
                if (elect_one()) mbar_arrive(&sfull[b]);
            } else if (elect_one()) {
                mbar_expect_tx(&sfull[b], stage_bytes);
                for (int j = 0; j < nsub0; j++) bulk_g2s(sb + (size_t)j * sps, g0 + toff + (size_t)j * slot, (uint32_t)sps, &sfull[b]);
                for (int j = 0; j < nsub1; j++)
                    bulk_g2s(sb + (size_t)(nsub0 + j) * sps, g1 + toff + (size_t)j * slot, (uint32_t)sps, &sfull[b]);
            }
            __syncwarp();
            SBC2_TR(0, t, 2);
        }
    } else if (warp == 5) {
        // ---------------- MMA issuer: converged warp, uniform-datapath descriptors, one elected lane issues ----------------
        mbar_wait(&wfull[wslot], wpar);
        const uint32_t wb16 = smem_u32(wb) >> 4;
        const uint32_t idesc = (uint32_t)cc.z, pmask = (uint32_t)npart - 1u, ncol = (uint32_t)op.N;
        const int n_mma = (L.dbg & 1) ? 0 : cc.y;
        const uint32_t st16 = smem_u32(stage) >> 4, sbytes16 = stage_bytes >> 4;
        for (int t = 0; t < T; t++) {
            const int b = t & (ns - 1);
            const uint32_t a = (uint32_t)(t * span) & (SBC2_NACC - 1);
            SBC2_TR(1, t, 0);
            mbar_wait(&sfull[b], (P.sfull_k >> b) & 1u);
            SBC2_TR(1, t, 1);
            P.sfull_k ^= 1u << b;
            mbar_wait(&tempty[a], (P.tempty_k >> a) & 1u);
            P.tempty_k ^= 1u << a;
            SBC2_TR(1, t, 2);
            tc_fence_after();
            const uint32_t sb16 = st16 + (uint32_t)b * sbytes16;
            const uint32_t td = tmem + a * 64u;
            if (elect_one()) {
                if (cc.x >= 0) {
                    const uint2* cl = sbc2_c_mma + cc.x;
#pragma unroll 4
                    for (int k = 0; k < n_mma; k++) {
                        const uint2 e = cl[k];
                        umma_f16(td + ((uint32_t)k & pmask) * ncol, ((uint64_t)0x4008u << 32) | (uint64_t)(e.x + sb16),
                                 ((uint64_t)0x4008u << 32) | (uint64_t)(e.y + wb16), idesc, k >= npart ? 1u : 0u);
                    }
                } else {
                    const uint2* list = reinterpret_cast<const uint2*>(wb + op.mma_rel);
#pragma unroll 2
                    for (int k = 0; k < n_mma; k++) {
                        const uint2 e = list[k];
                        umma_f16(td + ((uint32_t)k & pmask) * ncol, ((uint64_t)0x4008u << 32) | (uint64_t)(e.x + sb16),
                                 ((uint64_t)0x4008u << 32) | (uint64_t)(e.y + wb16), idesc, k >= npart ? 1u : 0u);
                    }
                }
                umma_commit(&sempty[b]);     // the stage may be refilled once these MMAs have read it
                umma_commit(&tfull[a]);      // accumulator complete
            }
            __syncwarp();
            SBC2_TR(1, t, 3);
        }
    } else {
        mbar_wait(&wfull[wslot], wpar);      // the bias lives in the parameter segment
        const int nch = op.cout8 >> 3;
        if (nch == 1) P.tfull_k = conv_epilogue<1>(op, L.S, L.dbg, arena, wb, tfull, tempty, tmem, P.tfull_k, T, npart, span, warp, lane, tr);
        else if (nch == 2) P.tfull_k = conv_epilogue<2>(op, L.S, L.dbg, arena, wb, tfull, tempty, tmem, P.tfull_k, T, npart, span, warp, lane, tr);
        else P.tfull_k = conv_epilogue<4>(op, L.S, L.dbg, arena, wb, tfull, tempty, tmem, P.tfull_k, T, npart, span, warp, lane, tr);
    }
    P.conv_n = (P.conv_n + 1u) & 0xFFFFu;      // only its two low bits matter (weight buffer, barrier parity)
    return pipe_pack(P);
}

__device__ __forceinline__ float block_sum(float v, float* red, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();                      // protects `red` against the previous use
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < SBC2_NTHR / 32; i++) s += red[i];
    return s;
}

}  // namespace sbc2k

// Everything that follows one network evaluation of a group: the score output (forward mode) or the data-consistency
// gradient, Langevin update and NMSE (test_score.py:157-170).  Out of line: the main loop stays small and spill free.
__device__ __forceinline__ uint32_t after_network(const Sbc2Launch& L, uint8_t* arena, uint8_t* smem, uint64_t* bars, const SbcStepScalars* s_sc,
                                               const float* s_hnorm, float* s_red, const int b0, const int gs, const int lvl, const int nsteps,
                                               const int conv_par, uint32_t tail_par, const int tid) {
    using namespace sbc2k;
    const int S = L.S, Nt = L.Nt, Nr = L.Nr, ne = Nt * Nr;
    if (L.mode == 0) {
        for (int s = 0; s < S && b0 + s < L.B; s++) {   // score = net / sigmas[y]   (ncsnv2.py:295-298)
            const int b = b0 + s;
            long long lab = L.labels ? L.labels[b] : 0;
            if (lab < 0) lab = 0;
            if (lab >= L.n_sigmas) lab = L.n_sigmas - 1;
            const float sg = L.sigmas[lab];
            const float2* net = reinterpret_cast<const float2*>(arena + L.out_off) + (size_t)s * ne;
            float* o = L.fout + (size_t)b * L.channels * ne;
            for (int e = tid; e < ne; e += SBC2_NTHR) {
                const float2 v = net[e];
                o[e] = v.x / sg;
                o[ne + e] = v.y / sg;
            }
        }
    } else {
        // data-consistency residual P x - y, gradient, Langevin update and NMSE, sample by sample.  The state x, the
        // residual, Y (staging ring, idle now) and P (the weight buffer that is not holding the prefetched first
        // conv) are staged in shared memory: the two complex matrix products read each operand ~Nt / ~Np times.
        const size_t pbytes = (size_t)L.Np * Nt * 8, ybytes = (size_t)L.Np * Nr * 8;
        const bool staged = pbytes <= (size_t)L.wmax && (size_t)ne * 8 + 2 * ybytes <= (size_t)L.stage_bytes &&
                            ((reinterpret_cast<uintptr_t>(L.P) & 15) == 0);
        float* xs = reinterpret_cast<float*>(smem + 2 * (size_t)L.wmax);
        float* rs = xs + (size_t)ne * 2;
        float* ys = rs + (size_t)L.Np * Nr * 2;
        float* ps = reinterpret_cast<float*>(smem + (size_t)((uint32_t)conv_par ^ 1u) * L.wmax);
        for (int s = 0; s < S && b0 + s < L.B; s++) {
            const int b = b0 + s;
            int st_last = nsteps - 1;
            if (L.stop_step) { const int st = L.stop_step[b]; st_last = min(max(st, 0), nsteps - 1); }
            if (gs > st_last) continue;                        // this sample stopped early
            float* ax = reinterpret_cast<float*>(arena + L.x_off) + (size_t)s * ne * 2;
            const float* net = reinterpret_cast<const float*>(arena + L.out_off) + (size_t)s * ne * 2;
            const float* Pm = L.P + (size_t)b * L.Np * Nt * 2;
            const float* Ym = L.Y + (size_t)b * L.Np * Nr * 2;
            const float* Hc = L.Hor ? L.Hor + (size_t)b * ne * 2 : nullptr;
            const float* en = L.ext_noise ? L.ext_noise + ((size_t)gs * L.B + b) * ne * 2 : nullptr;
            const unsigned long long sid = L.sample_ids ? L.sample_ids[b] : (unsigned long long)b;
            const uint32_t gstep = (uint32_t)(lvl * L.steps_each + gs % L.steps_each);
            float part;
            if (staged) {
                if (tid == 0) {
                    mbar_expect_tx(&bars[18], (uint32_t)pbytes);
                    bulk_g2s(ps, Pm, (uint32_t)pbytes, &bars[18]);
                }
                for (int e = tid; e < ne; e += SBC2_NTHR) reinterpret_cast<float2*>(xs)[e] = reinterpret_cast<const float2*>(ax)[e];
                for (int e = tid; e < L.Np * Nr; e += SBC2_NTHR) reinterpret_cast<float2*>(ys)[e] = reinterpret_cast<const float2*>(Ym)[e];
                mbar_wait(&bars[18], tail_par);
                tail_par ^= 1u;
                __syncthreads();
                sbc_dc_residual(xs, rs, ps, ys, Nt, Nr, L.Np, tid, SBC2_NTHR);
                __syncthreads();
                part = sbc_langevin_update(xs, net, rs, ps, Hc, en, s_sc[s], L.seed, sid, gstep, Nt, Nr, L.Np, tid, SBC2_NTHR);
                for (int e = tid; e < ne; e += SBC2_NTHR) reinterpret_cast<float2*>(ax)[e] = reinterpret_cast<const float2*>(xs)[e];   // own elements
            } else {
                float* res = reinterpret_cast<float*>(arena + L.post_off) + (size_t)s * ne * 2;
                sbc_dc_residual(ax, res, Pm, Ym, Nt, Nr, L.Np, tid, SBC2_NTHR);
                __syncthreads();
                part = sbc_langevin_update(ax, net, res, Pm, Hc, en, s_sc[s], L.seed, sid, gstep, Nt, Nr, L.Np, tid, SBC2_NTHR);
            }
            if (L.nmse_log && Hc) {
                const float tot = block_sum(part, s_red, tid);
                if (tid == 0) L.nmse_log[(size_t)gs * L.B + b] = tot / s_hnorm[s];
            } else {
                __syncthreads();
            }
        }
    }
    return tail_par;
}

// ---------------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SBC2_NTHR, 2) sbc2_ald_kernel(const __grid_constant__ Sbc2Launch L) {
    using namespace sbc2k;
    extern __shared__ __align__(128) uint8_t sbc2_smem[];
    uint8_t* smem = sbc2_smem;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * (size_t)L.wmax + (size_t)L.stage_bytes);
    float* spart = reinterpret_cast<float*>(bars + SBC2_NBARS);
    __shared__ uint32_t s_tmem;
    // op records are prefetched global -> shared one op ahead (first 40 threads, one word each; the word loaded at the
    // start of op i is stored at its end), so decoding an op never waits on L2
    __shared__ __align__(16) sbc2::Op s_ops[2];
    __shared__ __align__(16) Sbc2Launch s_L;        // copy of the launch record that out-of-line functions can take by reference
    __shared__ SbcStepScalars s_sc[SBC2_MAXS];
    __shared__ float s_hnorm[SBC2_MAXS];
    __shared__ float s_red[SBC2_NTHR / 32];
    const int tid = threadIdx.x, warp = tid >> 5;
    const int S = L.S;
    if (tid == 0) s_L = L;

    if (tid == 0) {
        for (int i = 0; i < 4; i++) { mbar_init(&bars[i], 1); mbar_init(&bars[4 + i], 1); mbar_init(&bars[8 + i], 1); mbar_init(&bars[12 + i], 4); }
        mbar_init(&bars[16], 1);
        mbar_init(&bars[17], 1);
        mbar_init(&bars[18], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(SBC2_TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    uint8_t* arena = L.gws + (size_t)blockIdx.x * (size_t)L.arena_bytes;
    {   // the arena starts out all zero: SP16 pads and guards are never written with anything else
        uint4* a4 = reinterpret_cast<uint4*>(arena);
        const size_t n16 = (size_t)L.arena_bytes / 16;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (size_t i = tid; i < n16; i += SBC2_NTHR) a4[i] = z;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem;

    Pipe P;
    P.sfull_k = 0u; P.sempty_k = 0xFu; P.tfull_k = 0u; P.tempty_k = 0xFu; P.conv_n = 0u;
    uint32_t tail_par = 0u;   // parity of the tail's bulk-copy barrier (bars[18])
    bool w_pending = false;   // a parameter-segment prefetch is in flight (or landed) for conv number P.conv_n
    const int n_groups = (L.B + S - 1) / S;
    const int Nt = L.Nt, Nr = L.Nr, ne = Nt * Nr;
    const int nsteps = (L.mode == 1) ? (L.level_end - L.level_begin) * L.steps_each : 1;
    int last_conv = -1;
    for (int i = 0; i < L.n_ops; i++)
        if (sbc2_c_conv[i].y > 0) last_conv = i;
    if (tid < 40) reinterpret_cast<int*>(&s_ops[0])[tid] = reinterpret_cast<const int*>(L.ops)[tid];
    uint32_t opc = 0;                         // ops executed so far (parity = ring slot)
    __syncthreads();

    for (int g = blockIdx.x; g < n_groups; g += gridDim.x) {
        const bool last_group = (g + (int)gridDim.x >= n_groups);
        const int b0 = g * S;
        // ---------------- load the group's state ----------------
        for (int s = 0; s < S; s++) {
            const int b = b0 + s;
            float2* ax = reinterpret_cast<float2*>(arena + L.x_off) + (size_t)s * ne;
            if (b >= L.B) {
                for (int e = tid; e < ne; e += SBC2_NTHR) ax[e] = make_float2(0.f, 0.f);
                continue;
            }
            if (L.mode == 1) {
                const float2* X = reinterpret_cast<const float2*>(L.X) + (size_t)b * ne;
                for (int e = tid; e < ne; e += SBC2_NTHR) ax[e] = X[e];
                if (L.Hor) {   // ||H||_F^2 once per sample (test_score.py:169)
                    const float2* Hc = reinterpret_cast<const float2*>(L.Hor) + (size_t)b * ne;
                    float part = 0.f;
                    for (int e = tid; e < ne; e += SBC2_NTHR) { const float2 v = Hc[e]; part += v.x * v.x + v.y * v.y; }
                    const float tot = block_sum(part, s_red, tid);
                    if (tid == 0) s_hnorm[s] = tot;
                }
            } else {
                const float* fx = L.fx + (size_t)b * L.fxs[0];
                for (int e = tid; e < ne; e += SBC2_NTHR) {
                    const int t = e / Nr, r = e - t * Nr;
                    const float* qp = fx + t * L.fxs[2] + r * L.fxs[3];
                    ax[e] = make_float2(qp[0], qp[L.fxs[1]]);
                }
            }
        }
        // steps this group runs = the longest of its samples (early stop, test_mmse.py:260-263)
        int nsteps_g = nsteps;
        if (L.mode == 1 && L.stop_step) {
            int mx = 1;
            for (int s = 0; s < S && b0 + s < L.B; s++) {
                const int st = L.stop_step[b0 + s] + 1;
                mx = max(mx, min(max(st, 1), nsteps));
            }
            nsteps_g = mx;
        }
        __syncthreads();

        for (int gs = 0; gs < nsteps_g; gs++) {
            const bool last_step = (gs + 1 == nsteps_g);
            int lvl = 0;
            if (L.mode == 1) {
                lvl = L.level_begin + gs / L.steps_each;
                if (tid < S && b0 + tid < L.B && (gs % L.steps_each) == 0) {   // per-level scalars, in double like the reference
                    const int b = b0 + tid;
                    const double sigma = (double)L.sigmas[lvl];
                    const double ratio = sigma / L.sigma_end;
                    const double alpha = (double)L.alpha_step[b] * ratio * ratio;
                    s_sc[tid].sigma = L.sigmas[lvl];
                    s_sc[tid].alpha = (float)alpha;
                    s_sc[tid].den = (float)((double)L.noise_var[b] / 2. + sigma * sigma) / (L.dc_boost ? L.dc_boost[b] : 1.f);
                    s_sc[tid].nscale = (float)sqrt(2. * alpha * (double)L.beta[b]);
                }
            }
            const bool do_prof_all = (L.prof != nullptr) && blockIdx.x == 0 && g == 0 && gs == (nsteps > 1 ? 1 : 0);
            const bool do_prof = do_prof_all && tid == 0;

            // ---------------- the network ----------------
            for (int i = 0; i < L.n_ops; i++, opc++) {
                if (do_prof) L.prof[i] = clock64();
                const sbc2::Op& op = s_ops[opc & 1u];
                int pend = 0;
                if (tid < 40) pend = reinterpret_cast<const int*>(L.ops + (i + 1 < L.n_ops ? i + 1 : 0))[tid];
                if (op.kind == sbc2::K_CONV) {
                    if (!w_pending) {      // very first conv of the launch: nobody prefetched its segment
                        if (tid == 4 * 32) {
                            mbar_expect_tx(&bars[16 + (P.conv_n & 1u)], (uint32_t)op.w_len);
                            bulk_g2s(smem + (size_t)(P.conv_n & 1u) * L.wmax, L.blob + op.w_off, (uint32_t)op.w_len, &bars[16 + (P.conv_n & 1u)]);
                        }
                    }
                    // the last conv of the launch must not leave a copy in flight
                    const bool pf = !(i == last_conv && last_step && last_group);
#ifndef SBC2_X_NOCONV
                    P = pipe_unpack(op_conv(i, op, L.blob, L.wmax, L.dbg, S, arena, smem, bars, tmem, pipe_pack(P), tid, pf, (do_prof_all && i == L.trace_op) ? L.prof + L.n_ops + 2 : nullptr));
#endif
                    w_pending = pf;
                } else if (L.dbg & 8) {
                } else if (op.kind == sbc2::K_NORM_ELU) {
#ifndef SBC2_X_NONORM
                    op_norm_elu(op, S, L.blob, arena, spart, tid, (do_prof && i == L.trace_op) ? L.prof + L.n_ops + 2 : nullptr);
#endif
                } else if (op.kind == sbc2::K_MAXPOOL5) {
#ifndef SBC2_X_NOMISC
                    op_maxpool5(op, S, arena, tid);
#endif
                } else if (op.kind == sbc2::K_ELU) {
                    op_elu(op, S, arena, tid);
                } else if (op.kind == sbc2::K_UPACC) {
#ifndef SBC2_X_NOMISC
                    op_upacc(op, S, arena, tid);
#endif
                } else if (op.kind == sbc2::K_POOL2) {
                    op_pool2(op, S, arena, tid);
                } else if (op.kind == sbc2::K_EPILOGUE) {
                    op_epilogue(op, S, arena, tid);
                } else if (op.kind == sbc2::K_AFFINE) {
                    op_affine(op, S, arena, tid);
                }
                if (op.fence_after) fence_proxy_async();      // generic-proxy stores -> visible to the next conv's bulk copies
                if (tid < 40) reinterpret_cast<int*>(&s_ops[(opc + 1u) & 1u])[tid] = pend;
                __syncthreads();
            }
            if (do_prof) L.prof[L.n_ops] = clock64();

            // ---------------- after the network ----------------
#ifndef SBC2_X_NOTAIL
            tail_par = after_network(s_L, arena, smem, bars, s_sc, s_hnorm, s_red, b0, gs, lvl, nsteps, (int)(P.conv_n & 1u), tail_par, tid);
#endif
            __syncthreads();
            if (do_prof) L.prof[L.n_ops + 1] = clock64();
        }

        if (L.mode == 1) {   // write the final estimates back (interleaved complex64)
            for (int s = 0; s < S && b0 + s < L.B; s++) {
                float2* X = reinterpret_cast<float2*>(L.X) + (size_t)(b0 + s) * ne;
                const float2* ax = reinterpret_cast<const float2*>(arena + L.x_off) + (size_t)s * ne;
                for (int e = tid; e < ne; e += SBC2_NTHR) X[e] = ax[e];
            }
        }
        __syncthreads();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(SBC2_TMEM_COLS));
}
