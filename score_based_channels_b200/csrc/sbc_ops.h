// sbc_ops.h -- per-thread bodies of the non-convolution ops of the fused NCSNv2Deepest forward and of the
// annealed-Langevin update.  All functions take (tid, nthr) explicitly and contain no barriers or warp
// intrinsics, so the same source is compiled by nvcc for the kernel (sbc_kernel.cuh) and by g++ for the
// CPU thread-emulation harness (tests/emu/emu.cpp) that checks indexing before any GPU time is spent.
//
// Layout: see SbcGeo (sbc_program.h): channel-interleaved by 4 (one float4 = 4 channels of a pixel),
// zero halo.  Every op that writes a tensor writes the interior only and re-zeroes the halo of its
// fresh outputs (sbc_zero_halo), because arena regions are recycled between tensors.
#pragma once
#include <math.h>
#include <stdint.h>

#include "sbc_program.h"

#if defined(__CUDACC__)
#define SBC_HD __host__ __device__ __forceinline__
#else
#define SBC_HD static inline
#endif

struct alignas(16) SbcF4 {
    float x, y, z, w;
};

// nn.ELU(alpha=1)  (reference ncsnv2/models/layers.py:13).  Device: exp via MUFU.EX2 (absolute error
// ~1e-7, far inside the parity tolerance); host emulation: expm1f.
SBC_HD float sbc_elu(float v) {
#if defined(__CUDA_ARCH__)
    return v > 0.f ? v : __expf(v) - 1.f;
#else
    return v > 0.f ? v : expm1f(v);
#endif
}
SBC_HD SbcF4 sbc_elu4(SbcF4 v) { return SbcF4{sbc_elu(v.x), sbc_elu(v.y), sbc_elu(v.z), sbc_elu(v.w)}; }

SBC_HD SbcF4* sbc_px(float* base, const SbcGeo& G, int cg, int y, int x) {
    return reinterpret_cast<SbcF4*>(base) + (cg * G.pps + G.org + y * G.wp + x);
}
SBC_HD const SbcF4* sbc_px(const float* base, const SbcGeo& G, int cg, int y, int x) {
    return reinterpret_cast<const SbcF4*>(base) + (cg * G.pps + G.org + y * G.wp + x);
}

// zero the halo cells of a tensor with `c` channels (cg = ceil(c/4) channel groups)
SBC_HD void sbc_zero_halo(float* t, const SbcGeo& G, int c, int tid, int nthr) {
    const int ncg = (c + 3) >> 2;
    const int rows = G.h + 2 * G.hy;
    const int top = G.hy * G.wp;                  // cells in the top (and bottom) band
    const int side = 2 * G.hx;                    // halo cells per interior row
    const int per = 2 * top + G.h * side;
    const SbcF4 z{0.f, 0.f, 0.f, 0.f};
    for (int i = tid; i < ncg * per; i += nthr) {
        const int cg = i / per;
        int r = i - cg * per;
        int cell;
        if (r < top) cell = r;
        else if (r < 2 * top) cell = (rows - G.hy) * G.wp + (r - top);
        else {
            r -= 2 * top;
            const int row = r / side, k = r - row * side;
            cell = (G.hy + row) * G.wp + (k < G.hx ? k : G.w + k);
        }
        reinterpret_cast<SbcF4*>(t)[cg * G.pps + cell] = z;
    }
}

// ----------------------------------------------------------------------------------------------
// InstanceNorm2dPlus + ELU (reference normalization.py:163-176).  One float4 lane-item covers the 4
// channels of a channel group; thread (cg, s) owns pixels s, s+T, s+2T, ... of channel group cg
// (T = nthr / ncg threads per group).  Statistics are two-pass; the per-thread partials below are
// combined across the T threads by shuffles + a shared-memory exchange on the device and by a plain
// loop in the emulation.
// ----------------------------------------------------------------------------------------------
SBC_HD int sbc_norm_T(const SbcOp& op, int nthr) {
    const int ncg = (op.cin + 3) >> 2;
    int T = nthr / ncg;
    int p = 32;                       // a power of two >= 32 so that warps never straddle channel groups
    while (p * 2 <= T) p *= 2;
    return p;
}
SBC_HD SbcF4 sbc_norm_partial_sum(const SbcOp& op, const SbcGeo& G, const float* arena, int cg, int s, int T) {
    SbcF4 a{0.f, 0.f, 0.f, 0.f};
    const int HW = G.h * G.w;
#pragma unroll 4
    for (int i = s; i < HW; i += T) {
        const int y = i / G.w, x = i - y * G.w;
        const SbcF4 v = *sbc_px(arena + op.src, G, cg, y, x);
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    return a;
}
SBC_HD SbcF4 sbc_norm_partial_m2(const SbcOp& op, const SbcGeo& G, const float* arena, int cg, int s, int T,
                                 SbcF4 mean) {
    SbcF4 a{0.f, 0.f, 0.f, 0.f};
    const int HW = G.h * G.w;
#pragma unroll 4
    for (int i = s; i < HW; i += T) {
        const int y = i / G.w, x = i - y * G.w;
        const SbcF4 v = *sbc_px(arena + op.src, G, cg, y, x);
        const float dx = v.x - mean.x, dy = v.y - mean.y, dz = v.z - mean.z, dw = v.w - mean.w;
        a.x = fmaf(dx, dx, a.x); a.y = fmaf(dy, dy, a.y); a.z = fmaf(dz, dz, a.z); a.w = fmaf(dw, dw, a.w);
    }
    return a;
}
// `mu` [C] per-channel means (any addressable memory), mean4 / m2_4: this thread's channel-group statistics
SBC_HD void sbc_norm_apply(const SbcOp& op, const SbcGeo& G, float* arena, const float* wseg, const float* mu, int cg,
                           int s, int T, SbcF4 mean4, SbcF4 m2_4) {
    const int C = op.cin, HW = G.h * G.w;
    // cross-channel statistics of the per-channel means: torch.mean / torch.var (unbiased) over C
    float m = 0.f;
    for (int c = 0; c < C; c++) m += mu[c];
    m /= (float)C;
    float v = 0.f;
    for (int c = 0; c < C; c++) { const float d = mu[c] - m; v = fmaf(d, d, v); }
    v /= (float)(C - 1);
    const float rv = 1.f / sqrtf(v + 1e-5f);
    const float *alpha = wseg + 4 * cg, *gamma = wseg + C + 4 * cg, *beta = wseg + 2 * C + 4 * cg;
    const float inv = 1.f / (float)HW;
    const float mean[4] = {mean4.x, mean4.y, mean4.z, mean4.w};
    const float m2[4] = {m2_4.x, m2_4.y, m2_4.z, m2_4.w};
    float a[4], b[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const bool real = 4 * cg + j < C;
        const float rstd = 1.f / sqrtf(m2[j] * inv + 1e-5f);       // nn.InstanceNorm2d: biased variance
        a[j] = real ? gamma[j] * rstd : 0.f;
        b[j] = real ? fmaf(gamma[j], (mean[j] - m) * rv * alpha[j], beta[j]) : 0.f;
    }
#pragma unroll 4
    for (int i = s; i < HW; i += T) {
        const int y = i / G.w, x = i - y * G.w;
        const SbcF4 q = *sbc_px(arena + op.src, G, cg, y, x);
        SbcF4 o;
        o.x = sbc_elu(fmaf(q.x - mean[0], a[0], b[0]));
        o.y = sbc_elu(fmaf(q.y - mean[1], a[1], b[1]));
        o.z = sbc_elu(fmaf(q.z - mean[2], a[2], b[2]));
        o.w = sbc_elu(fmaf(q.w - mean[3], a[3], b[3]));
        *sbc_px(arena + op.dst, G, cg, y, x) = o;
    }
}

// ----------------------------------------------------------------------------------------------
// element-wise ops (one float4 = 4 channels of one pixel per item)
// ----------------------------------------------------------------------------------------------
SBC_HD void sbc_elu_op(const SbcOp& op, const SbcGeo& G, float* arena, int tid, int nthr) {
    const int HW = G.h * G.w, n = ((op.cin + 3) >> 2) * HW;
#pragma unroll 4
    for (int i = tid; i < n; i += nthr) {
        const int cg = i / HW, r = i - cg * HW, y = r / G.w, x = r - y * G.w;
        *sbc_px(arena + op.dst, G, cg, y, x) = sbc_elu4(*sbc_px(arena + op.src, G, cg, y, x));
    }
    sbc_zero_halo(arena + op.dst, G, op.cin, tid, nthr);
}
// dst = 2*src - 1 on channels < cin; channels cin .. cout-1 of dst are zero (ncsnv2.py:270-271)
SBC_HD void sbc_affine_op(const SbcOp& op, const SbcGeo& G, float* arena, int tid, int nthr) {
    const int HW = G.h * G.w, n = ((op.cout + 3) >> 2) * HW;
    for (int i = tid; i < n; i += nthr) {
        const int cg = i / HW, r = i - cg * HW, y = r / G.w, x = r - y * G.w;
        SbcF4 o{0.f, 0.f, 0.f, 0.f};
        if (4 * cg < op.cin) {
            const SbcF4 v = *sbc_px(arena + op.src, G, cg, y, x);
            const int c = 4 * cg;
            o.x = (c < op.cin) ? 2.f * v.x - 1.f : 0.f;
            o.y = (c + 1 < op.cin) ? 2.f * v.y - 1.f : 0.f;
            o.z = (c + 2 < op.cin) ? 2.f * v.z - 1.f : 0.f;
            o.w = (c + 3 < op.cin) ? 2.f * v.w - 1.f : 0.f;
        }
        *sbc_px(arena + op.dst, G, cg, y, x) = o;
    }
    sbc_zero_halo(arena + op.dst, G, op.cout, tid, nthr);
}

// MaxPool2d(5, 1, 2) with implicit -inf padding (reference layers.py:70); 4 channels per item
SBC_HD void sbc_maxpool5_op(const SbcOp& op, const SbcGeo& G, float* arena, int tid, int nthr) {
    const int H = G.h, W = G.w, HW = H * W, n = ((op.cin + 3) >> 2) * HW;
    for (int i = tid; i < n; i += nthr) {
        const int cg = i / HW, r = i - cg * HW, y = r / W, x = r - y * W;
        SbcF4 m{-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        const int y0 = y - 2 < 0 ? 0 : y - 2, y1 = y + 2 >= H ? H - 1 : y + 2;
        const int x0 = x - 2 < 0 ? 0 : x - 2, x1 = x + 2 >= W ? W - 1 : x + 2;
        for (int yy = y0; yy <= y1; yy++) {
            const SbcF4* row = sbc_px(arena + op.src, G, cg, yy, 0);
            for (int xx = x0; xx <= x1; xx++) {
                const SbcF4 v = row[xx];
                m.x = fmaxf(m.x, v.x); m.y = fmaxf(m.y, v.y); m.z = fmaxf(m.z, v.z); m.w = fmaxf(m.w, v.w);
            }
        }
        *sbc_px(arena + op.dst, G, cg, y, x) = m;
    }
    sbc_zero_halo(arena + op.dst, G, op.cin, tid, nthr);
}

// acc += bilinear(src, size=(oh,ow), align_corners=True); optional edst = ELU(acc)  (layers.py:182-183)
SBC_HD void sbc_upacc_op(const SbcOp& op, const SbcGeo& GS, const SbcGeo& GD, float* arena, int tid, int nthr) {
    const int H = GS.h, W = GS.w, OH = GD.h, OW = GD.w;
    const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    const int n = ((op.cin + 3) >> 2) * OH * OW;
    for (int i = tid; i < n; i += nthr) {
        const int cg = i / (OH * OW), rem = i - cg * (OH * OW);
        const int y = rem / OW, x = rem - y * OW;
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        const SbcF4 p00 = *sbc_px(arena + op.src, GS, cg, y0, x0), p01 = *sbc_px(arena + op.src, GS, cg, y0, x1);
        const SbcF4 p10 = *sbc_px(arena + op.src, GS, cg, y1, x0), p11 = *sbc_px(arena + op.src, GS, cg, y1, x1);
        SbcF4* a = sbc_px(arena + op.acc, GD, cg, y, x);
        SbcF4 v = *a;
        v.x += hy * (hx * p00.x + lx * p01.x) + ly * (hx * p10.x + lx * p11.x);
        v.y += hy * (hx * p00.y + lx * p01.y) + ly * (hx * p10.y + lx * p11.y);
        v.z += hy * (hx * p00.z + lx * p01.z) + ly * (hx * p10.z + lx * p11.z);
        v.w += hy * (hx * p00.w + lx * p01.w) + ly * (hx * p10.w + lx * p11.w);
        *a = v;
        if (op.edst >= 0) *sbc_px(arena + op.edst, GD, cg, y, x) = sbc_elu4(v);
    }
    if (op.edst >= 0) sbc_zero_halo(arena + op.edst, GD, op.cin, tid, nthr);
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
// x lives in the arena at in_off as a 2-channel tensor of geometry G (channel 0 = re, 1 = im of
// x[t][r], pixel (y=t, x=r)).  P [Np][Nt], Y [Np][Nr], Hor [Nt][Nr] are interleaved complex64 in global
// memory.
// ----------------------------------------------------------------------------------------------
struct SbcStepScalars {
    float sigma;      // sigmas[level]
    float alpha;      // alpha_step * (sigma / sigma_end)^2        (test_score.py:143-144)
    float den;        // noise_var / 2 + sigma^2                   (test_score.py:165)
    float nscale;     // sqrt(2 * alpha * beta)                    (test_score.py:160)
};

// phase 1: res = P @ x - y   (test_score.py:157-158, inner product), planar [2][Np*Nr] into `res`
SBC_HD void sbc_dc_residual(const float* ax, const SbcGeo& G, float* res, const float* P, const float* Y, int Nt,
                            int Nr, int Np, int tid, int nthr) {
    for (int o = tid; o < Np * Nr; o += nthr) {
        const int p = o / Nr, r = o - p * Nr;
        float sr = 0.f, si = 0.f;
#pragma unroll 4
        for (int t = 0; t < Nt; t++) {
            const float pr = P[2 * (p * Nt + t)], pi = P[2 * (p * Nt + t) + 1];
            const SbcF4 c = *sbc_px(ax, G, 0, t, r);
            sr += pr * c.x - pi * c.y;
            si += pr * c.y + pi * c.x;
        }
        res[o] = sr - Y[2 * o];
        res[Np * Nr + o] = si - Y[2 * o + 1];
    }
}
// phase 2: g = P^H res;  x += alpha*(net/sigma - g/den) + nscale*eps;  per-thread |x-H|^2 partial
SBC_HD float sbc_langevin_update(float* ax, const float* net, const SbcGeo& G, const float* res, const float* P,
                                 const float* Hor, const float* ext_noise, const SbcStepScalars& sc, uint64_t seed,
                                 uint64_t sid, uint32_t gstep, int Nt, int Nr, int Np, int tid, int nthr) {
    const int ne = Nt * Nr;
    float part = 0.f;
    for (int e = tid; e < ne; e += nthr) {
        const int t = e / Nr, r = e - t * Nr;
        float gr = 0.f, gi = 0.f;
#pragma unroll 2
        for (int p = 0; p < Np; p++) {   // conj(P[p,t]) * res[p,r]
            const float pr = P[2 * (p * Nt + t)], pi = -P[2 * (p * Nt + t) + 1];
            const float cr = res[p * Nr + r], ci = res[Np * Nr + p * Nr + r];
            gr += pr * cr - pi * ci;
            gi += pr * ci + pi * cr;
        }
        float nr_, ni_;
        if (ext_noise) { nr_ = ext_noise[2 * e]; ni_ = ext_noise[2 * e + 1]; }
        else sbc_noise_cn01(seed, sid, gstep, e, nr_, ni_);
        SbcF4* xp = sbc_px(ax, G, 0, t, r);
        const SbcF4 nv = *sbc_px(net, G, 0, t, r);
        SbcF4 xv = *xp;
        const float sr = nv.x / sc.sigma, si = nv.y / sc.sigma;   // ncsnv2.py:295-298
        xv.x = xv.x + sc.alpha * (sr - gr / sc.den) + sc.nscale * nr_;
        xv.y = xv.y + sc.alpha * (si - gi / sc.den) + sc.nscale * ni_;
        *xp = xv;
        if (Hor) {
            const float dr = xv.x - Hor[2 * e], di = xv.y - Hor[2 * e + 1];
            part += dr * dr + di * di;
        }
    }
    return part;
}
