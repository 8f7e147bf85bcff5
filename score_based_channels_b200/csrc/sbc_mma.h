// sbc_mma.h -- per-lane pieces of the tensor-core convolution (SBC_OP_CONV_MMA): geometry, the
// A-fragment gather (implicit im2col straight from the planar arena), and the fused epilogue.
// Host/device portable like sbc_ops.h so that tests/emu/emu.cpp can emulate a warp lane by lane.
//
// Implicit GEMM per (live tap, chunk of 8 input channels):
//     D[16 output pixels, 8 couts] += A[16 pixels, 8 cins] * B[8 cins, 8 couts]
// with the fragment layout of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32
//     lane = 4*g + t :  A: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)      (row = pixel, col = cin)
//                       B: b0=(k=t,n=g) b1=(k=t+4,n=g)      (packed by program.py: hi0,hi1[,lo0,lo1] per lane)
//                       C: c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)    (row = pixel, col = cout)
// A tile is 16 consecutive output pixels in row-major order.  ConvMeanPool (SBC_F_POOL) runs four
// accumulations per tile -- one per position of the 2x2 pooling window, input pixel
// (2Y+qy, 2X+qx) -- and sums them in the epilogue (weights carry the 1/4).
#pragma once
#include "sbc_ops.h"

struct SbcMmaGeom {
    int P;        // output pixels
    int MT, NT;   // 16-pixel tiles, 8-cout tiles
    int KC;       // chunks of 8 input channels
    int ntaps, S; // live taps, K steps = ntaps * KC
    int stride;   // 1, or 2 for pooled convs
    int nq;       // accumulations per tile (1, or 4 for pooled convs)
};

SBC_HD void sbc_mma_geom(const SbcOp& op, SbcMmaGeom& G) {
    G.stride = (op.flags & SBC_F_POOL) ? 2 : 1;
    G.nq = (op.flags & SBC_F_POOL) ? 4 : 1;
    G.P = op.oh * op.ow;
    G.MT = (G.P + 15) >> 4;
    G.NT = (op.cout + 7) >> 3;
    G.KC = (op.cin + 7) >> 3;
    int n = 0;
    for (int tap = 0; tap < op.ksize * op.ksize; tap++) n += (op.tapmask >> tap) & 1;
    G.ntaps = n;
    G.S = n * G.KC;
}

// fp32 -> TF32 operand (cvt.rna.tf32.f32: nearest, ties away from zero), returned as fp32 bits
SBC_HD float sbc_tf32(float x) {
#if defined(__CUDA_ARCH__)
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
#else
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u = (v.u + 0x1000u) & 0xFFFFE000u;
    return v.f;
#endif
}

// input coordinates (before the tap offset) of tile row m (0..15) of tile mt, pooling position quad
SBC_HD void sbc_mma_row(const SbcOp& op, const SbcMmaGeom& G, int mt, int quad, int m, int& iy, int& ix, bool& ok) {
    const int q = mt * 16 + m;
    ok = q < G.P;
    const int Y = q / op.ow, X = q - Y * op.ow;
    iy = Y * G.stride + (quad >> 1);
    ix = X * G.stride + (quad & 1);
}

// A fragment of one lane: rows g (coords 0) and g+8 (coords 1), channels kc*8 + t and + 4
SBC_HD void sbc_mma_a_frag(const SbcOp& op, const float* arena, int iy0, int ix0, bool ok0, int iy1, int ix1, bool ok1,
                           int dy, int dx, int kc, int lane, float (&a)[4]) {
    const int t = lane & 3;
    const int h = op.h, w = op.w, ps = SBC_PS(h, w);
    const int y0 = iy0 + dy, x0 = ix0 + dx, y1 = iy1 + dy, x1 = ix1 + dx;
    const bool v0 = ok0 && y0 >= 0 && y0 < h && x0 >= 0 && x0 < w;
    const bool v1 = ok1 && y1 >= 0 && y1 < h && x1 >= 0 && x1 < w;
    const int c0 = kc * 8 + t, c1 = c0 + 4;
    const float* s0 = arena + op.src + c0 * ps;
    const float* s1 = s0 + 4 * ps;
    const bool k0 = c0 < op.cin, k1 = c1 < op.cin;
    a[0] = (v0 && k0) ? s0[y0 * w + x0] : 0.f;
    a[1] = (v1 && k0) ? s0[y1 * w + x1] : 0.f;
    a[2] = (v0 && k1) ? s1[y0 * w + x0] : 0.f;
    a[3] = (v1 && k1) ? s1[y1 * w + x1] : 0.f;
}

// Epilogue of one lane for tile (mt, nt):  v = c + bias;  dst <- v;  acc <- (v += acc);  edst <- ELU(v)
SBC_HD void sbc_mma_epilogue(const SbcOp& op, float* arena, const float* wseg, int mt, int nt, int lane,
                             const float (&c)[4]) {
    const int g = lane >> 2, t = lane & 3;
    const int P = op.oh * op.ow, ps = SBC_PS(op.oh, op.ow);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int q = mt * 16 + g + (j >> 1) * 8;
        const int co = nt * 8 + 2 * t + (j & 1);
        if (q < P && co < op.cout) {
            const int idx = co * ps + q;
            float v = c[j] + ((op.b_rel >= 0) ? wseg[op.b_rel + co] : 0.f);
            if (op.dst >= 0) arena[op.dst + idx] = v;
            if (op.acc >= 0) {
                v += arena[op.acc + idx];
                arena[op.acc + idx] = v;
            }
            if (op.edst >= 0) arena[op.edst + idx] = sbc_elu(v);
        }
    }
}
