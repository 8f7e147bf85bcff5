// sbc_ops.h -- per-thread bodies of every op of the fused NCSNv2Deepest forward and of the
// annealed-Langevin update.  All functions take (tid, nthr) explicitly and contain no barriers or
// warp intrinsics, so the same source is compiled by nvcc for the kernel (sbc_kernel.cuh) and by
// g++ for the CPU thread-emulation harness (tests/emu/emu.cpp) that checks indexing before any GPU
// time is spent.  Ops that need a block-wide dependency are split into phases; the caller puts a
// barrier between phases.
//
// Layout: every activation is planar fp32 [C][H][W] at a float offset inside one per-sample arena.
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

// nn.ELU(alpha=1)  (reference ncsnv2/models/layers.py:13).  Device: exp via MUFU.EX2 (absolute
// error ~1e-7, far inside the parity tolerance); host emulation: expm1f.
SBC_HD float sbc_elu(float v) {
#if defined(__CUDA_ARCH__)
    return v > 0.f ? v : __expf(v) - 1.f;
#else
    return v > 0.f ? v : expm1f(v);
#endif
}

// ----------------------------------------------------------------------------------------------
// Convolution (reference layers.py:28-60; ConvMeanPool layers.py:309-313)
//   item  = (cout block cbi, output row Y, strip of PX output pixels starting at X0)
//   kpart = which 1/ks slice of the input channels this thread accumulates
// Packed weights: [cout/CB][cin][k*k][CB] (program.py:conv), bias (if any) at wseg + b_rel.
// ----------------------------------------------------------------------------------------------
template <int PX, int CB>
SBC_HD void sbc_conv_partial(const SbcOp& op, const float* arena, const float* wseg, int item, int kpart,
                             float (&acc)[PX * CB]) {
    const int h = op.h, w = op.w, oh = op.oh, ow = op.ow, cin = op.cin, K = op.ksize, dil = op.dil;
    const int spr = ow / PX;          // strips per output row
    const int nsp = oh * spr;         // spatial items
    const int cbi = item / nsp;
    const int sp = item - cbi * nsp;
    const int Y = sp / spr;
    const int X0 = (sp - Y * spr) * PX;
    const int cper = cin / op.ks;
    const int c0 = kpart * cper, c1 = c0 + cper;
    const int KK = K * K;
    const float* src = arena + op.src;
    const int ps = SBC_PS(h, w);
    const bool pool = (op.flags & SBC_F_POOL) != 0;
#pragma unroll
    for (int i = 0; i < PX * CB; i++) acc[i] = 0.f;

    if (K == 3 && dil == 1 && !pool) {
        // hot path: contiguous (PX+2)-wide window per tap row
        for (int ci = c0; ci < c1; ci++) {
            const float* pl = src + ci * ps;
            const float* wp = wseg + (size_t)((cbi * cin + ci) * 9) * CB;
#pragma unroll
            for (int ky = 0; ky < 3; ky++) {
                const int yy = Y + ky - 1;
                if (yy < 0 || yy >= h) continue;
                const float* row = pl + yy * w + X0;
                float v[PX + 2];
                v[0] = (X0 > 0) ? row[-1] : 0.f;
                if (PX == 4) {
                    const SbcF4 q = *reinterpret_cast<const SbcF4*>(row);
                    v[1] = q.x; v[2] = q.y; v[3] = q.z; v[4] = q.w;
                } else {
#pragma unroll
                    for (int j = 0; j < PX; j++) v[j + 1] = row[j];
                }
                v[PX + 1] = (X0 + PX < w) ? row[PX] : 0.f;
#pragma unroll
                for (int kx = 0; kx < 3; kx++) {
                    float wv[CB];
                    const float* wt = wp + (ky * 3 + kx) * CB;
                    if (CB % 4 == 0) {
#pragma unroll
                        for (int c4 = 0; c4 < CB / 4; c4++) {
                            const SbcF4 q = *reinterpret_cast<const SbcF4*>(wt + 4 * c4);
                            wv[4 * c4] = q.x; wv[4 * c4 + 1] = q.y; wv[4 * c4 + 2] = q.z; wv[4 * c4 + 3] = q.w;
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < CB; c++) wv[c] = wt[c];
                    }
#pragma unroll
                    for (int p = 0; p < PX; p++)
#pragma unroll
                        for (int c = 0; c < CB; c++) acc[p * CB + c] = fmaf(v[p + kx], wv[c], acc[p * CB + c]);
                }
            }
        }
        return;
    }

    // generic path: 1x1, dilated, or pooled (conv over the 2x2 box-summed input at stride 2; weights
    // already carry the 1/4)
    const int r = K / 2;
    for (int ci = c0; ci < c1; ci++) {
        const float* pl = src + ci * ps;
        const float* wp = wseg + (size_t)((cbi * cin + ci) * KK) * CB;
        for (int ky = 0; ky < K; ky++) {
            const int oy = (ky - r) * dil;
            for (int kx = 0; kx < K; kx++) {
                const int ox = (kx - r) * dil;
                float v[PX];
                bool any = false;
                if (!pool) {
                    const int yy = Y + oy;
                    if (yy < 0 || yy >= h) continue;
#pragma unroll
                    for (int p = 0; p < PX; p++) {
                        const int xx = X0 + p + ox;
                        const bool in = (xx >= 0 && xx < w);
                        v[p] = in ? pl[yy * w + xx] : 0.f;
                        any |= in;
                    }
                } else {
#pragma unroll
                    for (int p = 0; p < PX; p++) {
                        float s = 0.f;
#pragma unroll
                        for (int dy = 0; dy < 2; dy++) {
                            const int yy = 2 * Y + dy + oy;
                            if (yy < 0 || yy >= h) continue;
#pragma unroll
                            for (int dx = 0; dx < 2; dx++) {
                                const int xx = 2 * (X0 + p) + dx + ox;
                                if (xx >= 0 && xx < w) { s += pl[yy * w + xx]; any = true; }
                            }
                        }
                        v[p] = s;
                    }
                }
                if (!any) continue;
                const float* wt = wp + (ky * K + kx) * CB;
#pragma unroll
                for (int c = 0; c < CB; c++) {
                    const float wc = wt[c];
#pragma unroll
                    for (int p = 0; p < PX; p++) acc[p * CB + c] = fmaf(v[p], wc, acc[p * CB + c]);
                }
            }
        }
    }
}

// Epilogue of one item (after the K-split partials have been summed):
//   v = acc + bias;  dst <- v;  accbuf <- (v += accbuf);  edst <- ELU(v)
template <int PX, int CB>
SBC_HD void sbc_conv_epilogue(const SbcOp& op, float* arena, const float* wseg, int item,
                              const float (&acc)[PX * CB]) {
    const int oh = op.oh, ow = op.ow;
    const int spr = ow / PX, nsp = oh * spr;
    const int cbi = item / nsp;
    const int sp = item - cbi * nsp;
    const int Y = sp / spr;
    const int X0 = (sp - Y * spr) * PX;
    const int plane = SBC_PS(oh, ow);
#pragma unroll
    for (int c = 0; c < CB; c++) {
        const int co = cbi * CB + c;
        const float b = (op.b_rel >= 0) ? wseg[op.b_rel + co] : 0.f;
        const int base = co * plane + Y * ow + X0;
        float v[PX];
#pragma unroll
        for (int p = 0; p < PX; p++) v[p] = acc[p * CB + c] + b;
        if (op.dst >= 0) {
#pragma unroll
            for (int p = 0; p < PX; p++) arena[op.dst + base + p] = v[p];
        }
        if (op.acc >= 0) {
#pragma unroll
            for (int p = 0; p < PX; p++) {
                v[p] += arena[op.acc + base + p];
                arena[op.acc + base + p] = v[p];
            }
        }
        if (op.edst >= 0) {
#pragma unroll
            for (int p = 0; p < PX; p++) arena[op.edst + base + p] = sbc_elu(v[p]);
        }
    }
}

SBC_HD int sbc_conv_items(const SbcOp& op) { return op.oh * (op.ow / op.px) * (op.cout / op.cb); }

// ----------------------------------------------------------------------------------------------
// InstanceNorm2dPlus + ELU (reference normalization.py:163-176).
//   S = power of two <= 32 lanes cooperate on one channel; lane (c, s) owns elements s, s+S, ... of
//   channel c.  Statistics are two-pass (mean, then centred second moment); the S partial sums are
//   combined with warp shuffles on the device (sbc_kernel.cuh) and by a plain loop in the CPU
//   emulation.  scratch: chan[2*C] = (mean, rstd) per channel.
// ----------------------------------------------------------------------------------------------
SBC_HD int sbc_norm_S(const SbcOp& op, int nthr) {
    int S = 32;
    while (S > 1 && op.cin * S > nthr) S >>= 1;
    return S;
}
SBC_HD float sbc_norm_partial_sum(const SbcOp& op, const float* arena, int c, int s, int S) {
    const int HW = op.h * op.w;
    const float* x = arena + op.src + c * SBC_PS(op.h, op.w);
    float sum = 0.f;
#pragma unroll 8
    for (int i = s; i < HW; i += S) sum += x[i];
    return sum;
}
SBC_HD float sbc_norm_partial_m2(const SbcOp& op, const float* arena, int c, int s, int S, float mean) {
    const int HW = op.h * op.w;
    const float* x = arena + op.src + c * SBC_PS(op.h, op.w);
    float m2 = 0.f;
#pragma unroll 8
    for (int i = s; i < HW; i += S) { const float d = x[i] - mean; m2 = fmaf(d, d, m2); }
    return m2;
}
SBC_HD void sbc_norm_store_stats(const SbcOp& op, float* arena, int c, float mean, float m2) {
    float* chan = arena + op.scratch;
    chan[c] = mean;
    chan[op.cin + c] = 1.f / sqrtf(m2 / (float)(op.h * op.w) + 1e-5f);   // nn.InstanceNorm2d: biased variance
}
SBC_HD void sbc_norm_apply(const SbcOp& op, float* arena, const float* wseg, int tid, int nthr) {
    const int C = op.cin, HW = op.h * op.w, S = sbc_norm_S(op, nthr), ps = SBC_PS(op.h, op.w);
    const float* chan = arena + op.scratch;
    // cross-channel statistics of the per-channel means: torch.mean / torch.var (unbiased) over C
    float m = 0.f;
    for (int c = 0; c < C; c++) m += chan[c];
    m /= (float)C;
    float v = 0.f;
    for (int c = 0; c < C; c++) { const float d = chan[c] - m; v = fmaf(d, d, v); }
    v /= (float)(C - 1);
    const float rv = 1.f / sqrtf(v + 1e-5f);
    const float *alpha = wseg, *gamma = wseg + C, *beta = wseg + 2 * C;
    for (int t = tid; t < C * S; t += nthr) {
        const int c = t / S, s = t - c * S;
        const float mu = chan[c];
        const float a = gamma[c] * chan[C + c];
        const float b = fmaf(gamma[c], (mu - m) * rv * alpha[c], beta[c]);
        const float* x = arena + op.src + c * ps;
        float* o = arena + op.dst + c * ps;
#pragma unroll 8
        for (int i = s; i < HW; i += S) o[i] = sbc_elu(fmaf(x[i] - mu, a, b));
    }
}

// ----------------------------------------------------------------------------------------------
// element-wise ops
// ----------------------------------------------------------------------------------------------
SBC_HD void sbc_elu_op(const SbcOp& op, float* arena, int tid, int nthr) {
    const int HW = op.h * op.w, n = op.cin * HW, ps = SBC_PS(op.h, op.w);
#pragma unroll 4
    for (int i = tid; i < n; i += nthr) {
        const int c = i / HW, j = c * ps + (i - c * HW);
        arena[op.dst + j] = sbc_elu(arena[op.src + j]);
    }
}
SBC_HD void sbc_affine_op(const SbcOp& op, float* arena, int tid, int nthr) {
    const int HW = op.h * op.w, n = op.cin * HW, ps = SBC_PS(op.h, op.w);
    for (int i = tid; i < n; i += nthr) {
        const int c = i / HW, j = c * ps + (i - c * HW);
        arena[op.dst + j] = 2.f * arena[op.src + j] - 1.f;
    }
}

// MaxPool2d(5, 1, 2) with implicit -inf padding (reference layers.py:70): strip of up to 4 pixels/thread
SBC_HD void sbc_maxpool5_op(const SbcOp& op, float* arena, int tid, int nthr) {
    const int C = op.cin, H = op.h, W = op.w;
    const int PX = (W % 4 == 0) ? 4 : ((W % 2 == 0) ? 2 : 1);
    const int spr = W / PX;
    const int n = C * H * spr;
    for (int it = tid; it < n; it += nthr) {
        const int c = it / (H * spr);
        const int rem = it - c * (H * spr);
        const int y = rem / spr, x0 = (rem - y * spr) * PX;
        const float* pl = arena + op.src + c * SBC_PS(H, W);
        float best[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
        for (int dy = -2; dy <= 2; dy++) {
            const int yy = y + dy;
            if (yy < 0 || yy >= H) continue;
            const float* row = pl + yy * W;
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int xx = x0 - 2 + j;
                v[j] = (j < PX + 4 && xx >= 0 && xx < W) ? row[xx] : -INFINITY;
            }
#pragma unroll
            for (int p = 0; p < 4; p++) {
                if (p < PX) {
                    const float mrow = fmaxf(fmaxf(fmaxf(v[p], v[p + 1]), fmaxf(v[p + 2], v[p + 3])), v[p + 4]);
                    best[p] = fmaxf(best[p], mrow);
                }
            }
        }
        float* o = arena + op.dst + c * SBC_PS(H, W) + y * W + x0;
#pragma unroll
        for (int p = 0; p < 4; p++)
            if (p < PX) o[p] = best[p];
    }
}

// acc += bilinear(src, size=(oh,ow), align_corners=True); optional edst = ELU(acc)  (layers.py:182-183)
SBC_HD void sbc_upacc_op(const SbcOp& op, float* arena, int tid, int nthr) {
    const int C = op.cin, H = op.h, W = op.w, OH = op.oh, OW = op.ow;
    const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f;
    const float sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
    const int n = C * OH * OW;
    for (int i = tid; i < n; i += nthr) {
        const int c = i / (OH * OW);
        const int rem = i - c * (OH * OW);
        const int y = rem / OW, x = rem - y * OW;
        const float fy = sy * (float)y, fx = sx * (float)x;
        const int y0 = (int)fy, x0 = (int)fx;
        const int y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
        const float* p = arena + op.src + c * SBC_PS(H, W);
        const float val = hy * (hx * p[y0 * W + x0] + lx * p[y0 * W + x1]) +
                          ly * (hx * p[y1 * W + x0] + lx * p[y1 * W + x1]);
        const int j = c * SBC_PS(OH, OW) + rem;
        const float v = arena[op.acc + j] + val;
        arena[op.acc + j] = v;
        if (op.edst >= 0) arena[op.edst + j] = sbc_elu(v);
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
// x lives planar in the arena at in_off: xr[t*Nr+r], xi = xr + SBC_PS(Nt,Nr).  P [Np][Nt], Y [Np][Nr],
// Hor [Nt][Nr] are interleaved complex64 in global memory.
// ----------------------------------------------------------------------------------------------
struct SbcStepScalars {
    float sigma;      // sigmas[level]
    float alpha;      // alpha_step * (sigma / sigma_end)^2        (test_score.py:143-144)
    float den;        // noise_var / 2 + sigma^2                   (test_score.py:165)
    float nscale;     // sqrt(2 * alpha * beta)                    (test_score.py:160)
};

// phase 1: res = P @ x - y   (test_score.py:157-158, inner product), planar into arena[post ..]
SBC_HD void sbc_dc_residual(const float* arena_x, float* res, const float* P, const float* Y, int Nt, int Nr, int Np,
                            int tid, int nthr) {
    const float* xr = arena_x;
    const float* xi = arena_x + SBC_PS(Nt, Nr);
    for (int o = tid; o < Np * Nr; o += nthr) {
        const int p = o / Nr, r = o - p * Nr;
        float sr = 0.f, si = 0.f;
        for (int t = 0; t < Nt; t++) {
            const float pr = P[2 * (p * Nt + t)], pi = P[2 * (p * Nt + t) + 1];
            const float cr = xr[t * Nr + r], ci = xi[t * Nr + r];
            sr += pr * cr - pi * ci;
            si += pr * ci + pi * cr;
        }
        res[o] = sr - Y[2 * o];
        res[Np * Nr + o] = si - Y[2 * o + 1];
    }
}
// phase 2: g = P^H res;  x += alpha*(net/sigma - g/den) + nscale*eps;  per-thread |x-H|^2 partial
SBC_HD float sbc_langevin_update(float* arena_x, const float* net_out, const float* res, const float* P,
                                 const float* Hor, const float* ext_noise, const SbcStepScalars& sc, uint64_t seed,
                                 uint64_t sid, uint32_t gstep, int Nt, int Nr, int Np, int tid, int nthr) {
    float* xr = arena_x;
    float* xi = arena_x + SBC_PS(Nt, Nr);
    const int ne = Nt * Nr;
    const float* net_im = net_out + SBC_PS(Nt, Nr);
    float part = 0.f;
    for (int e = tid; e < ne; e += nthr) {
        const int t = e / Nr, r = e - t * Nr;
        float gr = 0.f, gi = 0.f;
        for (int p = 0; p < Np; p++) {   // conj(P[p,t]) * res[p,r]
            const float pr = P[2 * (p * Nt + t)], pi = -P[2 * (p * Nt + t) + 1];
            const float cr = res[p * Nr + r], ci = res[Np * Nr + p * Nr + r];
            gr += pr * cr - pi * ci;
            gi += pr * ci + pi * cr;
        }
        float nr_, ni_;
        if (ext_noise) { nr_ = ext_noise[2 * e]; ni_ = ext_noise[2 * e + 1]; }
        else sbc_noise_cn01(seed, sid, gstep, e, nr_, ni_);
        const float sr = net_out[e] / sc.sigma, si = net_im[e] / sc.sigma;   // ncsnv2.py:295-298
        const float vr = xr[e] + sc.alpha * (sr - gr / sc.den) + sc.nscale * nr_;
        const float vi = xi[e] + sc.alpha * (si - gi / sc.den) + sc.nscale * ni_;
        xr[e] = vr; xi[e] = vi;
        if (Hor) {
            const float dr = vr - Hor[2 * e], di = vi - Hor[2 * e + 1];
            part += dr * dr + di * di;
        }
    }
    return part;
}
