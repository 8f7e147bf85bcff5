// sbc_mma.h -- per-lane pieces of the tensor-core convolution (SBC_OP_CONV_MMA): the A-fragment gather
// (implicit im2col straight from the arena) and the fused epilogue.  Host/device portable like sbc_ops.h so
// that tests/emu/emu.cpp can emulate a warp lane by lane.
//
// Implicit GEMM per K step s = (live tap, chunk of 8 input channels):
//     D[16 output pixels, 8 couts] += A[16 pixels, 8 cins] * B[8 cins, 8 couts]
// with the fragment layout of mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32
//     lane = 4*g + t :  A: a0=(row g, k=t) a1=(row g+8, k=t) a2=(row g, k=t+4) a3=(row g+8, k=t+4)
//                       B: b0 = W[cout g][cin t], b1 = W[cout g][cin t+4]       (packed by program.py)
//                       C: (c0,c1) = couts (2t, 2t+1) of row g, (c2,c3) = same of row g+8
// rows = output pixels, k = input channel within the chunk.  One pixel of one plane is 4 channels = 16
// contiguous bytes = one row of the 8x8 b16 matrices ldmatrix moves (8 consecutive pixels = 128 contiguous bytes:
// conflict free), so with the arena in shared memory ONE ldmatrix.x4 per lane delivers the whole A fragment
// already in register order (matrix 0/1 = plane 2*kc (channels 0-3) of pixel rows 0-7 / 8-15, matrix 2/3 = plane
// 2*kc+1 (channels 4-7); lane l supplies the address of pixel row l & 15 in plane l >> 4).  Thanks to the zero halo (SbcGeo) a K step is a constant address offset --
// read from the per-op table at the head of the parameter segment (program.py) -- and the gather has no bounds
// checks.  ConvMeanPool (SBC_F_POOL) runs four accumulations per tile, one per position of the 2x2 pooling
// window (input pixel (2Y+qy, 2X+qx)), summed before the epilogue (the packed weights carry the 1/4).
#pragma once
#include "sbc_ops.h"

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

// float offset (relative to plane 0 of the source tensor) of the input pixel that tile row m (0..15) of tile
// mt reads for pooling position `quad`, before the K-step offset.  Rows past the last output pixel are
// clamped onto it (their results are discarded by the epilogue).
SBC_HD int sbc_mma_row_off(const SbcOp& op, const SbcGeo& GS, int mt, int quad, int m) {
    const int P = op.oh * op.ow;
    int q = mt * 16 + m;
    if (q >= P) q = P - 1;
    const int Y = sbc_div(q, op.ow, op.low), X = q - Y * op.ow;
    const int sh = (op.flags & SBC_F_POOL) ? 1 : 0;
    const int iy = (Y << sh) + (quad >> 1), ix = (X << sh) + (quad & 1);
    return (GS.org + iy * GS.wp + ix) * 4;
}

// A fragment of one lane from plain loads (global-memory arena, CPU emulation): `asrc` = arena + op.src + t,
// `off` = K-step offset from the op's table, po0 / po1 = sbc_mma_row_off of tile rows g / g + 8, pl4 = floats per
// plane of the source geometry (channels t+4 live in the next plane)
SBC_HD void sbc_mma_a_frag(const float* asrc, int off, int po0, int po1, int pl4, float (&a)[4]) {
    a[0] = asrc[off + po0];
    a[1] = asrc[off + po1];
    a[2] = asrc[off + po0 + pl4];
    a[3] = asrc[off + po1 + pl4];
}

// what the epilogue needs from the op
// (accb = base the accumulator offset is relative to: the arena, or the park area when SBC_F_ACC_G is set)
struct SbcEpi {
    int dst, acc, edst, flags, cout, b_rel, pps4;
    float* accb;
};
SBC_HD SbcEpi sbc_epi(const SbcOp& op, const SbcGeo& GD, float* arena, float* park) {
    return SbcEpi{op.dst, op.acc, op.edst, op.flags, op.cout, op.b_rel, GD.pps * 4,
                  (op.flags & SBC_F_ACC_G) ? park : arena};
}

// Epilogue of one lane for cout tile nt of one pixel tile:  v = c + bias;  dst <- v;  acc <- (v += acc);
// edst <- ELU(v).  pd0 / pd1 = float offset (org + Y*wp + X) * 4 of the output pixel of tile rows g / g + 8 in
// the destination geometry, or -1 when that row is past the last pixel; q0 = index of the pixel of row g.
// (c0,c1) and (c2,c3) are two adjacent output channels of one pixel: 8-byte accesses.  cout is even.
SBC_HD void sbc_mma_epilogue(SbcEpi e, float* arena, const float* wseg, int pd0, int pd1, int q0, int nt,
                                     int lane, float c0, float c1, float c2, float c3) {
    const int t = lane & 3;
    const int co = nt * 8 + 2 * t;
    if (co >= e.cout) return;
    float b0 = 0.f, b1 = 0.f;
    if (e.b_rel >= 0) { b0 = wseg[e.b_rel + co]; b1 = wseg[e.b_rel + co + 1]; }
    const int cofs = (co >> 2) * e.pps4 + (co & 3);
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int pd = half ? pd1 : pd0;
        if (pd < 0) continue;
        SbcF2 v{(half ? c2 : c0) + b0, (half ? c3 : c1) + b1};
        if (e.flags & SBC_F_COMPACT) {   // network output: couts (0,1) = (re, im) of element q
            reinterpret_cast<SbcF2*>(arena + e.dst)[q0 + 8 * half] = v;
            continue;
        }
        const int idx = pd + cofs;
        if (e.dst >= 0) *reinterpret_cast<SbcF2*>(arena + e.dst + idx) = v;
        if (e.acc >= 0) {
            SbcF2* ap = reinterpret_cast<SbcF2*>(e.accb + e.acc + idx);
            const SbcF2 o = *ap;
            v.x += o.x; v.y += o.y;
            *ap = v;
        }
        if (e.edst >= 0) *reinterpret_cast<SbcF2*>(arena + e.edst + idx) = SbcF2{sbc_elu(v.x), sbc_elu(v.y)};
    }
}
// destination pixel offsets of one lane for tile mt (see sbc_mma_epilogue)
SBC_HD void sbc_mma_dst_off(const SbcOp& op, const SbcGeo& GD, int mt, int g, int (&pd)[2]) {
    const int P = op.oh * op.ow;
#pragma unroll
    for (int half = 0; half < 2; half++) {
        const int q = mt * 16 + g + 8 * half;
        if (q >= P) { pd[half] = -1; continue; }
        const int Y = sbc_div(q, op.ow, op.low), X = q - Y * op.ow;
        pd[half] = (GD.org + Y * GD.wp + X) * 4;
    }
}
