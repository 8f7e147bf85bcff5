// sbc_mma.h -- per-lane pieces of the tensor-core convolution (SBC_OP_CONV_MMA): geometry, the
// A-fragment gather (implicit im2col straight from the arena), and the fused epilogue.  Host/device
// portable like sbc_ops.h so that tests/emu/emu.cpp can emulate a warp lane by lane.
//
// Implicit GEMM per (live tap, chunk of 8 input channels):
//     D[16 output pixels, 8 couts] += A[16 pixels, 8 cins] * B[8 cins, 8 couts]
// with the fragment layout of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32
//     lane = 4*g + t :  A: a0=(g,t) a1=(g+8,t) a2=(g,t+4) a3=(g+8,t+4)      (row = pixel, col = cin)
//                       B: b0=(k=t,n=g) b1=(k=t+4,n=g)      (packed by program.py: hi0,hi1[,lo0,lo1] per lane)
//                       C: c0=(g,2t) c1=(g,2t+1) c2=(g+8,2t) c3=(g+8,2t+1)    (row = pixel, col = cout)
// A tile is 16 consecutive output pixels in row-major order.  Thanks to the zero halo of the layout
// (SbcGeo) a tap is a constant address offset and the gather has no bounds checks; 8 consecutive pixels x
// 4 channels are 128 contiguous bytes, so each of the four loads of a fragment is one conflict-free
// shared-memory wavefront.  ConvMeanPool (SBC_F_POOL) runs four accumulations per tile -- one per
// position of the 2x2 pooling window, input pixel (2Y+qy, 2X+qx) -- summed in the epilogue (the packed
// weights carry the 1/4).
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

SBC_HD void sbc_mma_geom(const SbcOp& op, SbcMmaGeom& M) {
    M.stride = (op.flags & SBC_F_POOL) ? 2 : 1;
    M.nq = (op.flags & SBC_F_POOL) ? 4 : 1;
    M.P = op.oh * op.ow;
    M.MT = (M.P + 15) >> 4;
    M.NT = (op.cout + 7) >> 3;
    M.KC = (op.cin + 7) >> 3;
    int n = 0;
    for (int tap = 0; tap < op.ksize * op.ksize; tap++) n += (op.tapmask >> tap) & 1;
    M.ntaps = n;
    M.S = n * M.KC;
}

// fp32 -> TF32 operand, round to nearest (ties away): (bits + 0x1000) & ~0x1fff   [finite inputs]
SBC_HD float sbc_tf32_rn(float x) {
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u = (v.u + 0x1000u) & 0xFFFFE000u;
    return v.f;
}
// fp32 -> TF32 operand by truncation (what the tensor core does with the low 13 mantissa bits)
SBC_HD float sbc_tf32_rz(float x) {
    union { float f; uint32_t u; } v;
    v.f = x;
    v.u &= 0xFFFFE000u;
    return v.f;
}

// float offset (relative to channel-group 0 of the source tensor, lane-in-group 0) of the input pixel that
// tile row m (0..15) of tile mt reads for pooling position `quad`, before the tap offset.  Rows past the
// last output pixel are clamped onto it (their results are discarded by the epilogue).
SBC_HD int sbc_mma_row_off(const SbcOp& op, const SbcMmaGeom& M, const SbcGeo& GS, int mt, int quad, int m) {
    int q = mt * 16 + m;
    if (q >= M.P) q = M.P - 1;
    const int Y = q / op.ow, X = q - Y * op.ow;
    const int iy = Y * M.stride + (quad >> 1), ix = X * M.stride + (quad & 1);
    return (GS.org + iy * GS.wp + ix) * 4;
}

// A fragment of one lane for tap offset (dy, dx) and input-channel chunk kc; po0 / po1 from sbc_mma_row_off
// for tile rows g and g + 8
SBC_HD void sbc_mma_a_frag(const SbcOp& op, const SbcGeo& GS, const float* arena, int po0, int po1, int dy, int dx,
                           int kc, int lane, float (&a)[4]) {
    const int t = lane & 3;
    const float* bp = arena + op.src + (2 * kc * GS.pps + dy * GS.wp + dx) * 4 + t;
    const int cg1 = GS.pps * 4;
    a[0] = bp[po0];
    a[1] = bp[po1];
    a[2] = bp[cg1 + po0];
    a[3] = bp[cg1 + po1];
}

// Epilogue of one lane for tile (mt, nt):  v = c + bias;  dst <- v;  acc <- (v += acc);  edst <- ELU(v)
// (c0,c1) and (c2,c3) are two adjacent output channels of one pixel: 8-byte stores.
SBC_HD void sbc_mma_epilogue(const SbcOp& op, const SbcGeo& GD, float* arena, const float* wseg, int mt, int nt,
                             int lane, const float (&c)[4]) {
    const int g = lane >> 2, t = lane & 3;
    const int P = op.oh * op.ow;
    const int co = nt * 8 + 2 * t;
    if (co >= op.cout) return;
    const bool two = co + 1 < op.cout;
    const float b0 = (op.b_rel >= 0) ? wseg[op.b_rel + co] : 0.f;
    const float b1 = (op.b_rel >= 0 && two) ? wseg[op.b_rel + co + 1] : 0.f;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int q = mt * 16 + g + half * 8;
        if (q >= P) continue;
        const int Y = q / op.ow, X = q - Y * op.ow;
        const int idx = ((co >> 2) * GD.pps + GD.org + Y * GD.wp + X) * 4 + (co & 3);
        float v0 = c[2 * half] + b0, v1 = c[2 * half + 1] + b1;
        if (op.dst >= 0) {
            arena[op.dst + idx] = v0;
            if (two) arena[op.dst + idx + 1] = v1;
        }
        if (op.acc >= 0) {
            v0 += arena[op.acc + idx];
            arena[op.acc + idx] = v0;
            if (two) { v1 += arena[op.acc + idx + 1]; arena[op.acc + idx + 1] = v1; }
        }
        if (op.edst >= 0) {
            arena[op.edst + idx] = sbc_elu(v0);
            if (two) arena[op.edst + idx + 1] = sbc_elu(v1);
        }
    }
}
