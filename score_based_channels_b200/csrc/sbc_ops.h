// sbc_ops.h -- per-thread bodies of the non-convolution ops of the fused NCSNv2Deepest forward and of the
// annealed-Langevin update.  All functions take (tid, nthr) explicitly and contain no barriers or warp
// intrinsics, so the same source is compiled by nvcc for the kernel (sbc_kernel.cuh) and by g++ for the
// CPU thread-emulation harness (tests/emu/emu.cpp) that checks indexing before any GPU time is spent.
//
// Layout: see SbcGeo (sbc_program.h): channel-interleaved by 4 (one pixel of one plane = one float4 "quad" of
// channels), zero halo.  Every op body writes the interior only; because arena regions are recycled
// between tensors, the CALLER re-zeroes the halo of a fresh output (op.cout channels at op.dst / op.edst)
// whenever the offline planner cannot prove that it is still zero (flags SBC_F_ZH_DST / SBC_F_ZH_EDST,
// program.py:_halo_analysis): sbc_zero_halo here (CPU emulation, global-arena path), a list-driven version in
// sbc_kernel.cuh.  The sampler
// state x and the raw network output are kept *compact*: interleaved (re, im) pairs [Nt*Nr], exactly the
// complex64 layout of `current` in reference test_score.py:126.
//
// The kernel is instruction-issue bound (profiles/): index arithmetic uses shifts whenever the row width is a
// power of two (SbcGeo.lw >= 0) and falls back to a division otherwise.
#pragma once
#include <math.h>
#include <stdint.h>

#include "sbc_program.h"

#if defined(__CUDACC__)
#define SBC_HD __host__ __device__ __forceinline__
#define SBC_HD_OUTLINE __host__ __device__ __noinline__   // one copy: the kernel's instruction footprint matters
#else
#define SBC_HD static inline
#define SBC_HD_OUTLINE static
#endif

struct alignas(16) SbcF4 {
    float x, y, z, w;
};
struct alignas(8) SbcF2 {
    float x, y;
};

// nn.ELU(alpha=1)  (reference ncsnv2/models/layers.py:13).  Device: exp via MUFU.EX2 (absolute error
// ~1e-7 = one ulp of the 1.0 that is subtracted, i.e. fp32-level); host emulation: expm1f.
SBC_HD float sbc_elu(float v) {
#if defined(__CUDA_ARCH__)
    return v > 0.f ? v : __expf(v) - 1.f;
#else
    return v > 0.f ? v : expm1f(v);
#endif
}
SBC_HD SbcF4 sbc_elu4(SbcF4 v) { return SbcF4{sbc_elu(v.x), sbc_elu(v.y), sbc_elu(v.z), sbc_elu(v.w)}; }

// i / d with d = 2^ld when ld >= 0; the general division is kept out of line (it is ~30 instructions and
// would otherwise be replicated at every call site of a kernel whose instruction-cache footprint matters)
SBC_HD_OUTLINE int sbc_div_slow(int i, int d) { return i / d; }
SBC_HD int sbc_div(int i, int d, int ld) { return ld >= 0 ? (i >> ld) : sbc_div_slow(i, d); }
SBC_HD int sbc_ilog2(int v) {   // log2(v) if v is a power of two, else -1
    if (v <= 0 || (v & (v - 1))) return -1;
#if defined(__CUDA_ARCH__)
    return 31 - __clz(v);
#else
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
#endif
}
// padded pixel index of interior pixel i = y * w + x
SBC_HD int sbc_pix(const SbcGeo& G, int i) {
    const int y = sbc_div(i, G.w, G.lw), x = i - y * G.w;
    return G.org + y * G.wp + x;
}
// visit the interior pixels s, s+T, s+2T, ... (T a power of two): f(padded pixel index).  When the row width is
// a power of two <= T the walk is one column with a constant address step (no per-pixel index arithmetic).
template <class F>
SBC_HD void sbc_for_pixels(const SbcGeo& G, int s, int T, F&& f) {
    if (G.lw >= 0 && T >= G.w) {
        const int dy = T >> G.lw, ps = dy * G.wp;
        int p = G.org + (s >> G.lw) * G.wp + (s & (G.w - 1));
        for (int y = s >> G.lw; y < G.h; y += dy, p += ps) f(p);
    } else {
        const int HW = G.h * G.w;
        for (int i = s; i < HW; i += T) f(sbc_pix(G, i));
    }
}
// 1 / sqrt(x): MUFU.RSQ on the device (2 ulp), exact division on the host emulation
SBC_HD float sbc_rsqrt(float x) {
#if defined(__CUDA_ARCH__)
    return rsqrtf(x);
#else
    return 1.f / sqrtf(x);
#endif
}
// float4 quad q (channels 4q .. 4q+3) of padded pixel p
SBC_HD SbcF4* sbc_q4(float* base, const SbcGeo& G, int q, int p) {
    return reinterpret_cast<SbcF4*>(base + ((size_t)q * G.pps + p) * 4);
}
SBC_HD const SbcF4* sbc_q4(const float* base, const SbcGeo& G, int q, int p) {
    return reinterpret_cast<const SbcF4*>(base + ((size_t)q * G.pps + p) * 4);
}

// zero the halo cells of a tensor with `c` channels: one item = one padded row of one plane.  Out of line
// (scalar arguments, so that the caller's structs stay in registers).
SBC_HD_OUTLINE void sbc_zero_halo_impl(float* t, int h, int w, int hy, int hx, int wp, int pps, int c, int tid,
                                       int nthr) {
    const int np = (c + 3) >> 2;
    const int rows = h + 2 * hy;
    const SbcF4 z{0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < np * rows; i += nthr) {
        const int pl = i / rows, row = i - pl * rows;
        SbcF4* r = reinterpret_cast<SbcF4*>(t + (size_t)(pl * pps + row * wp) * 4);
        if (row < hy || row >= hy + h) {
            for (int k = 0; k < wp; k++) r[k] = z;
        } else {
            for (int k = 0; k < hx; k++) { r[k] = z; r[hx + w + k] = z; }
        }
    }
}
SBC_HD void sbc_zero_halo(float* t, const SbcGeo& G, int c, int tid, int nthr) {
    sbc_zero_halo_impl(t, G.h, G.w, G.hy, G.hx, G.wp, G.pps, c, tid, nthr);
}

// ----------------------------------------------------------------------------------------------
// InstanceNorm2dPlus + ELU (reference normalization.py:163-176).  One float4 item covers the 4 channels of a
// quad; thread (q, s) owns pixels s, s+T, s+2T, ... of quad q (T = threads per quad, a power of two >= 32 so
// that warps never straddle quads).  Statistics are two-pass; the per-thread partials below are combined
// across the T threads by shuffles + a shared-memory exchange on the device and by a plain loop in the
// emulation.  C must be a multiple of 8.
// ----------------------------------------------------------------------------------------------
SBC_HD int sbc_norm_T(const SbcOp& op, int nthr) {
    (void)nthr;
    return op.MT;   // threads per quad: largest power of two <= nthr / (C/4), at least 32 (program.py)
}
SBC_HD float sbc_bits_f(int32_t b) {
    union { int32_t i; float f; } v;
    v.i = b;
    return v.f;
}
SBC_HD SbcF4 sbc_norm_partial_sum(const SbcOp& op, const SbcGeo& G, const float* arena, int q, int s, int T) {
    SbcF4 a{0.f, 0.f, 0.f, 0.f};
    const float* src = arena + op.src + (size_t)q * G.pps * 4;
    sbc_for_pixels(G, s, T, [&](int p) {
        const SbcF4 v = *reinterpret_cast<const SbcF4*>(src + p * 4);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    });
    return a;
}
SBC_HD SbcF4 sbc_norm_partial_m2(const SbcOp& op, const SbcGeo& G, const float* arena, int q, int s, int T,
                                 SbcF4 mean) {
    SbcF4 a{0.f, 0.f, 0.f, 0.f};
    const float* src = arena + op.src + (size_t)q * G.pps * 4;
    sbc_for_pixels(G, s, T, [&](int p) {
        const SbcF4 v = *reinterpret_cast<const SbcF4*>(src + p * 4);
        const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
        a.x = fmaf(dx, dx, a.x); a.y = fmaf(dy, dy, a.y); a.z = fmaf(dz, dz, a.z); a.w = fmaf(dw, dw, a.w);
    });
    return a;
}
// `mu` [C] per-channel means (16-byte aligned); mean4 / m2_4: the statistics of this thread's quad
SBC_HD void sbc_norm_apply(const SbcOp& op, const SbcGeo& G, float* arena, const float* wseg, const float* mu, int q,
                           int s, int T, SbcF4 mean4, SbcF4 m2_4) {
    const int C = op.cin, nq = C >> 2;
    // cross-channel statistics of the per-channel means: torch.mean / torch.var (unbiased) over C
    // (four independent partial sums: the dependent-add chain is what costs on a latency-bound SM)
    const SbcF4* mu4 = reinterpret_cast<const SbcF4*>(mu);
    SbcF4 ms{0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nq; k++) { const SbcF4 u = mu4[k]; ms.x += u.x; ms.y += u.y; ms.z += u.z; ms.w += u.w; }
    const float m = ((ms.x + ms.y) + (ms.z + ms.w)) * sbc_bits_f(op.low);        // * 1/C
    SbcF4 vs{0.f, 0.f, 0.f, 0.f};
    for (int k = 0; k < nq; k++) {
        const SbcF4 u = mu4[k];
        const float d0 = u.x - m, d1 = u.y - m, d2 = u.z - m, d3 = u.w - m;
        vs.x = fmaf(d0, d0, vs.x); vs.y = fmaf(d1, d1, vs.y); vs.z = fmaf(d2, d2, vs.z); vs.w = fmaf(d3, d3, vs.w);
    }
    const float v = ((vs.x + vs.y) + (vs.z + vs.w)) * sbc_bits_f(op.tapmask);    // * 1/(C-1)
    const float rv = sbc_rsqrt(v + 1e-5f);
    const SbcF4 al = *reinterpret_cast<const SbcF4*>(wseg + 4 * q);
    const SbcF4 ga = *reinterpret_cast<const SbcF4*>(wseg + C + 4 * q);
    const SbcF4 be = *reinterpret_cast<const SbcF4*>(wseg + 2 * C + 4 * q);
    const float inv = sbc_bits_f(op.frag_rel);                                  // 1/HW
    // nn.InstanceNorm2d: biased variance
    const float a0 = ga.x * sbc_rsqrt(m2_4.x * inv + 1e-5f), a1 = ga.y * sbc_rsqrt(m2_4.y * inv + 1e-5f);
    const float a2 = ga.z * sbc_rsqrt(m2_4.z * inv + 1e-5f), a3 = ga.w * sbc_rsqrt(m2_4.w * inv + 1e-5f);
    const float b0 = fmaf(ga.x, (mean4.x - m) * rv * al.x, be.x), b1 = fmaf(ga.y, (mean4.y - m) * rv * al.y, be.y);
    const float b2 = fmaf(ga.z, (mean4.z - m) * rv * al.z, be.z), b3 = fmaf(ga.w, (mean4.w - m) * rv * al.w, be.w);
    const size_t qo = (size_t)q * G.pps * 4;
    const float* src = arena + op.src + qo;
    float* dst = arena + op.dst + qo;
    sbc_for_pixels(G, s, T, [&](int p) {
        const SbcF4 x = *reinterpret_cast<const SbcF4*>(src + p * 4);
        SbcF4 o;
        o.x = sbc_elu(fmaf(x.x - mean4.x, a0, b0));
        o.y = sbc_elu(fmaf(x.y - mean4.y, a1, b1));
        o.z = sbc_elu(fmaf(x.z - mean4.z, a2, b2));
        o.w = sbc_elu(fmaf(x.w - mean4.w, a3, b3));
        *reinterpret_cast<SbcF4*>(dst + p * 4) = o;
    });
}

// ----------------------------------------------------------------------------------------------
// element-wise ops (one float4 quad of one pixel per item)
// ----------------------------------------------------------------------------------------------
SBC_HD void sbc_elu_op(const SbcOp& op, const SbcGeo& G, float* arena, int tid, int nthr) {
    const int nq = op.cin >> 2, T = op.MT, lT = op.NT, gpp = nthr >> lT, s = tid & (T - 1);
    for (int q = tid >> lT; q < nq; q += gpp) {
        const size_t qo = (size_t)q * G.pps * 4;
        const float* src = arena + op.src + qo;
        float* dst = arena + op.dst + qo;
        sbc_for_pixels(G, s, T, [&](int p) {
            *reinterpret_cast<SbcF4*>(dst + p * 4) = sbc_elu4(*reinterpret_cast<const SbcF4*>(src + p * 4));
        });
    }
}
// SPILL / FILL (two-CTAs-per-SM plans): a whole tensor -- halo cells included, so the zero-halo invariant travels
// with it -- is parked in the CTA's slice of a global-memory (L2-resident) area while shared memory is needed for
// something else.  MT = float4 items.  FILL issues its loads in batches so that one L2 latency covers several.
SBC_HD void sbc_spill_op(const SbcOp& op, const float* arena, float* park, int tid, int nthr) {
    const SbcF4* s = reinterpret_cast<const SbcF4*>(arena + op.src);
    SbcF4* d = reinterpret_cast<SbcF4*>(park + op.dst);
    for (int i = tid; i < op.MT; i += nthr) d[i] = s[i];
}
SBC_HD void sbc_fill_op(const SbcOp& op, float* arena, const float* park, int tid, int nthr) {
    const SbcF4* s = reinterpret_cast<const SbcF4*>(park + op.src);
    SbcF4* d = reinterpret_cast<SbcF4*>(arena + op.dst);
    const int n = op.MT;
    int i = tid;
    for (; i + 4 * nthr < n; i += 5 * nthr) {
        const SbcF4 a = s[i], b = s[i + nthr], c = s[i + 2 * nthr], e = s[i + 3 * nthr], f = s[i + 4 * nthr];
        d[i] = a; d[i + nthr] = b; d[i + 2 * nthr] = c; d[i + 3 * nthr] = e; d[i + 4 * nthr] = f;
    }
    for (; i < n; i += nthr) d[i] = s[i];
}
// dst (8 stored channels) = 2*x - 1 on channels 0,1 read from the compact state; channels 2..7 zero
// (ncsnv2.py:270-271; begin_conv then contracts over one chunk of 8 input channels)
SBC_HD void sbc_affine_op(const SbcOp& op, const SbcGeo& G, float* arena, int tid, int nthr) {
    const int HW = G.h * G.w;
    const SbcF2* xc = reinterpret_cast<const SbcF2*>(arena + op.src);
    for (int i = tid; i < 2 * HW; i += nthr) {
        const int e = i >> 1, p = sbc_pix(G, e);
        SbcF4 o{0.f, 0.f, 0.f, 0.f};
        if (!(i & 1)) {
            const SbcF2 v = xc[e];
            o.x = 2.f * v.x - 1.f;
            o.y = 2.f * v.y - 1.f;
        }
        *sbc_q4(arena + op.dst, G, i & 1, p) = o;
    }
}

// MaxPool2d(5, 1, 2) with implicit -inf padding (reference layers.py:70).  One item = one quad of a column
// segment of 4 output rows: the horizontal maxima of the 8 input rows it touches are formed once and the
// four vertical windows slide over them (10 loads per output instead of 25).  H must be a multiple of 4.
// Maps with at most one output pixel-quad per thread use the direct form instead.
SBC_HD SbcF4 sbc_max4(SbcF4 a, SbcF4 b) {
    return SbcF4{fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w)};
}
SBC_HD void sbc_maxpool5_op(const SbcOp& op, const SbcGeo& G, float* arena, int tid, int nthr) {
    const int H = G.h, W = G.w, nq = op.cin >> 2;
    const SbcF4 ninf{-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    if (nq * H * W <= nthr || (H & 3)) {
        // small maps: one output pixel per thread keeps the per-thread instruction chain short (the SM is
        // latency bound on such ops, not load bound)
        const int HW = H * W, lhw = sbc_ilog2(HW);
        for (int i = tid; i < nq * HW; i += nthr) {
            const int q = sbc_div(i, HW, lhw), r = i - q * HW;
            const int y = sbc_div(r, W, G.lw), x = r - y * W;
            const int y0 = y - 2 < 0 ? 0 : y - 2, y1 = y + 2 >= H ? H - 1 : y + 2;
            const int x0 = x - 2 < 0 ? 0 : x - 2, x1 = x + 2 >= W ? W - 1 : x + 2;
            SbcF4 m = ninf;
            for (int yy = y0; yy <= y1; yy++) {
                const int p = G.org + yy * G.wp;
                for (int xx = x0; xx <= x1; xx++) m = sbc_max4(m, *sbc_q4(arena + op.src, G, q, p + xx));
            }
            *sbc_q4(arena + op.dst, G, q, G.org + y * G.wp + x) = m;
        }
        return;
    }
    const int HB = H >> 2, per = HB * W, n = nq * per, lper = sbc_ilog2(per);
    for (int i = tid; i < n; i += nthr) {
        const int q = sbc_div(i, per, lper), r = i - q * per;
        const int yb = sbc_div(r, W, G.lw), x = r - yb * W, y0 = yb << 2;
        const int x0 = x - 2 < 0 ? 0 : x - 2, x1 = x + 2 >= W ? W - 1 : x + 2;
        SbcF4 hm[8];
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int yy = y0 - 2 + k;
            SbcF4 m = ninf;
            if (yy >= 0 && yy < H) {
                const int p = G.org + yy * G.wp;
                for (int xx = x0; xx <= x1; xx++) m = sbc_max4(m, *sbc_q4(arena + op.src, G, q, p + xx));
            }
            hm[k] = m;
        }
#pragma unroll
        for (int o = 0; o < 4; o++) {
            const SbcF4 m = sbc_max4(sbc_max4(sbc_max4(hm[o], hm[o + 1]), sbc_max4(hm[o + 2], hm[o + 3])), hm[o + 4]);
            *sbc_q4(arena + op.dst, G, q, G.org + (y0 + o) * G.wp + x) = m;
        }
    }
}

// acc += bilinear(src, size=(oh,ow), align_corners=True); optional edst = ELU(acc)  (layers.py:182-183)
SBC_HD void sbc_upacc_op(const SbcOp& op, const SbcGeo& GS, const SbcGeo& GD, float* arena, int tid, int nthr) {
    const int H = GS.h, W = GS.w, OH = GD.h, OW = GD.w, OHW = OH * OW, lohw = sbc_ilog2(OHW);
    const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    const int n = (op.cin >> 2) * OHW;
    for (int i = tid; i < n; i += nthr) {
        const int q = sbc_div(i, OHW, lohw), rem = i - q * OHW;
        const int y = sbc_div(rem, OW, GD.lw), x = rem - y * OW;
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        const int r0 = GS.org + y0 * GS.wp, r1 = GS.org + y1 * GS.wp;
        const SbcF4 p00 = *sbc_q4(arena + op.src, GS, q, r0 + x0), p01 = *sbc_q4(arena + op.src, GS, q, r0 + x1);
        const SbcF4 p10 = *sbc_q4(arena + op.src, GS, q, r1 + x0), p11 = *sbc_q4(arena + op.src, GS, q, r1 + x1);
        const int pd = GD.org + y * GD.wp + x;
        SbcF4* a = sbc_q4(arena + op.acc, GD, q, pd);
        SbcF4 v = *a;
        v.x += hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
        v.y += hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
        v.z += hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
        v.w += hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w);
        *a = v;
        if (op.edst >= 0) *sbc_q4(arena + op.edst, GD, q, pd) = sbc_elu4(v);
    }
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller: the project's RNG contract (include/sbc.h), identical to
// oracle/sbc_oracle.c:orc_noise.  counter = (element/2, step, sample_id lo, hi), key = seed.
// ----------------------------------------------------------------------------------------------
SBC_HD void sbc_philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
// CN(0,1) sample (re, im ~ N(0,1/2)) for complex element e
SBC_HD void sbc_noise_cn01(uint64_t seed, uint64_t sid, uint32_t step, int e, float& re, float& im) {
    uint32_t c[4] = {(uint32_t)(e >> 1), step, (uint32_t)sid, (uint32_t)(sid >> 32)};
    sbc_philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t a = (e & 1) ? c[2] : c[0], b = (e & 1) ? c[3] : c[1];
    const float u1 = ((float)(a >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float u2 = ((float)(b >> 8) + 0.5f) * (1.0f / 16777216.0f);
    const float rad = sqrtf(-2.0f * logf(u1));
    const float ang = 6.28318530717958647692f * u2;
    re = rad * cosf(ang) * 0.70710678118654752440f;
    im = rad * sinf(ang) * 0.70710678118654752440f;
}

// ----------------------------------------------------------------------------------------------
// Annealed-Langevin step around the network (reference test_score.py:157-170).
// x (compact, [Nt*Nr] (re, im)) lives in the arena at in_off, the raw network output at out_off.
// P [Np][Nt], Y [Np][Nr], Hor [Nt][Nr] are interleaved complex64 in global memory.
// ----------------------------------------------------------------------------------------------
struct SbcStepScalars {
    float sigma;      // sigmas[level]
    float alpha;      // alpha_step * (sigma / sigma_end)^2        (test_score.py:143-144)
    float den;        // noise_var / 2 + sigma^2                   (test_score.py:165)
    float nscale;     // sqrt(2 * alpha * beta)                    (test_score.py:160)
};

// phase 1: res = P @ x - y   (test_score.py:157-158, inner product), planar [2][Np*Nr] into `res`
SBC_HD void sbc_dc_residual(const float* xc, float* res, const float* P, const float* Y, int Nt, int Nr, int Np,
                            int tid, int nthr) {
    const SbcF2* x2 = reinterpret_cast<const SbcF2*>(xc);
    const SbcF2* P2 = reinterpret_cast<const SbcF2*>(P);
    const int lnr = sbc_ilog2(Nr);
    for (int o = tid; o < Np * Nr; o += nthr) {
        const int p = sbc_div(o, Nr, lnr), r = o - p * Nr;
        float sr = 0.f, si = 0.f;
#pragma unroll 8
        for (int t = 0; t < Nt; t++) {
            const SbcF2 pv = P2[p * Nt + t];
            const SbcF2 c = x2[t * Nr + r];
            sr += pv.x * c.x - pv.y * c.y;
            si += pv.x * c.y + pv.y * c.x;
        }
        res[o] = sr - Y[2 * o];
        res[Np * Nr + o] = si - Y[2 * o + 1];
    }
}
// phase 2: g = P^H res;  x += alpha*(net/sigma - g/den) + nscale*eps;  per-thread |x-H|^2 partial
SBC_HD float sbc_langevin_update(float* xc, const float* net, const float* res, const float* P, const float* Hor,
                                 const float* ext_noise, const SbcStepScalars& sc, uint64_t seed, uint64_t sid,
                                 uint32_t gstep, int Nt, int Nr, int Np, int tid, int nthr) {
    const int ne = Nt * Nr, lnr = sbc_ilog2(Nr);
    SbcF2* x2 = reinterpret_cast<SbcF2*>(xc);
    const SbcF2* n2 = reinterpret_cast<const SbcF2*>(net);
    const SbcF2* P2 = reinterpret_cast<const SbcF2*>(P);
    float part = 0.f;
    for (int e = tid; e < ne; e += nthr) {
        const int t = sbc_div(e, Nr, lnr), r = e - t * Nr;
        float gr = 0.f, gi = 0.f;
#pragma unroll 8
        for (int p = 0; p < Np; p++) {   // conj(P[p,t]) * res[p,r]
            const SbcF2 pv = P2[p * Nt + t];
            const float pr = pv.x, pi = -pv.y;
            const float cr = res[p * Nr + r], ci = res[Np * Nr + p * Nr + r];
            gr += pr * cr - pi * ci;
            gi += pr * ci + pi * cr;
        }
        float nr_, ni_;
        if (ext_noise) { nr_ = ext_noise[2 * e]; ni_ = ext_noise[2 * e + 1]; }
        else sbc_noise_cn01(seed, sid, gstep, e, nr_, ni_);
        const SbcF2 nv = n2[e];
        SbcF2 xv = x2[e];
        const float sr = nv.x / sc.sigma, si = nv.y / sc.sigma;   // ncsnv2.py:295-298
        xv.x = xv.x + sc.alpha * (sr - gr / sc.den) + sc.nscale * nr_;
        xv.y = xv.y + sc.alpha * (si - gi / sc.den) + sc.nscale * ni_;
        x2[e] = xv;
        if (Hor) {
            const float dr = xv.x - Hor[2 * e], di = xv.y - Hor[2 * e + 1];
            part += dr * dr + di * di;
        }
    }
    return part;
}
