// sbc_program.h -- op table + tensor geometry shared by the host program builder (program.py) and the
// kernel.  One SbcOp = 32 int32 words, in the order of program.py:OP_FIELDS.  Everything the device code
// would otherwise derive per op (tile counts, K steps, log2 of the row width, ...) is computed once on the
// host: the kernel is instruction-issue bound, so no integer division or table walk may sit on its hot path.
#pragma once
#include <stdint.h>

enum SbcOpKind : int32_t {
    SBC_OP_AFFINE = 0,    // dst = 2*x - 1 from the compact state buffer (reference ncsnv2/models/ncsnv2.py:270-271)
    SBC_OP_NORM_ELU = 2,  // dst = ELU(InstanceNorm2dPlus(src))   (normalization.py:163-176, layers.py:13)
    SBC_OP_ELU = 3,       // dst = ELU(src)
    SBC_OP_MAXPOOL5 = 4,  // MaxPool2d(5, stride 1, pad 2)        (layers.py:70)
    SBC_OP_UPACC = 5,     // acc += bilinear(src -> oh x ow, align_corners=True)   (layers.py:182-183)
    SBC_OP_CONV_MMA = 6,  // Conv2d k in {1,3}, stride 1, pad = dil*(k/2) (layers.py:28-60) + fused epilogue,
                          // contraction on tensor cores (mma.sync m16n8k8 TF32)
    SBC_OP_SPILL = 7,     // park[dst .. dst + 4*MT) = arena[src ..]: a whole tensor (halo included) leaves shared memory
    SBC_OP_FILL = 8,      // arena[dst .. dst + 4*MT) = park[src ..]: ... and comes back (two-CTAs-per-SM plans, program.py)
    SBC_OP_LAST = 8,
};

enum SbcOpFlags : int32_t {
    SBC_F_POOL = 1,       // conv followed by the 2x2 mean-pool of ConvMeanPool (layers.py:309-313)
    SBC_F_X3 = 2,         // 3xTF32 error-compensated product (fp32-equivalent accuracy)
    SBC_F_COMPACT = 4,    // conv epilogue writes couts 0,1 as interleaved (re, im) pairs [h*w] (the network output)
    SBC_F_ZH_DST = 8,     // the halo of the fresh tensor dst must be re-zeroed (decided offline by the planner:
    SBC_F_ZH_EDST = 16,   // program.py:_halo_analysis); same for edst
    SBC_F_UNIT = 32,      // conv with fewer (pixel tile, cout tile) units than warps: unit u is owned by the `ks`
                          // warps u*ks .. u*ks+ks-1 (ks = 1: one warp, no K split)
    SBC_F_LATEW = 128,    // the parameter segment is NOT prefetched during the previous parameterised op (its staging buffer
                          // would not fit next to that op's): the op issues its own bulk copy and waits for it
    SBC_F_ACC_G = 64,     // conv: `acc` is an offset into the CTA's park area in global memory (L2), not into the arena:
                          // a residual stream that only the epilogues read-modify-write needs no shared memory
};

// Geometry of every tensor of one resolution: channel-interleaved by 4 (one pixel of one plane = 4 channels
// = 16 bytes = one ldmatrix row = one row of a no-swizzle K-major UMMA core matrix; two planes = the K chunk
// of one m16n8k8 MMA), zero halo of (hy, hx) pixels.
//   addr(c, y, x) = base + ((c >> 2) * pps + org + y * wp + x) * 4 + (c & 3)        [floats]
struct SbcGeo {
    int32_t h, w;        // interior size
    int32_t hy, hx;      // halo (covers every live conv tap at this resolution)
    int32_t wp;          // padded row pitch in pixels = w + 2*hx
    int32_t pps;         // padded plane size in pixels = (h + 2*hy) * wp
    int32_t org;         // pixel index of (0, 0) = hy * wp + hx
    int32_t lw;          // log2(w) if w is a power of two, else -1 (division fallback)
};
#define SBC_MAX_GEO 8

struct SbcOp {
    int32_t kind, flags;
    int32_t src, dst, acc, edst;   // arena float offsets, -1 = unused
    int32_t cin, cout;             // (AFFINE: cin = real channels, cout = stored channels, extra ones zeroed)
    int32_t h, w;                  // input spatial size
    int32_t ksize, dil;
    int32_t w_off, w_len, b_rel;   // parameter segment (floats) in the blob; bias offset inside it
    int32_t sgeo, dgeo;            // geometry index of the input / output tensors
    int32_t ks;                    // conv: warps that split the K steps of one (pixel tile, cout tile) unit
    int32_t scratch;               // arena offset of op scratch (norm statistics, K-split partials)
    int32_t oh, ow;                // output spatial size
    int32_t next_w;                // filled by sbc_model_create: index of the next op with parameters
    int32_t tapmask;               // conv: live taps of the k x k window (bit = ky*k + kx)
    int32_t wbuf;                  // arena offset where the parameter segment is staged
    // ---- host-derived constants.  conv: as named.  NORM_ELU / ELU reuse the words: MT = T (threads per channel
    // quad, a power of two >= 32), NT = log2 T, S = number of passes over the quads; NORM_ELU also: frag_rel, low,
    // tapmask = float bits of 1/(h*w), 1/C, 1/(C-1) ----
    int32_t MT, NT;                // 16-pixel output tiles, 8-cout tiles
    int32_t S;                     // K steps = live taps * cin chunks; segment starts with S int32 A offsets
    int32_t frag_rel;              // offset of the B fragments inside the segment
    int32_t low;                   // log2(ow) or -1
    // ---- filled by sbc_model_create: parameter segment of the NEXT parameterised op (wrapping to the first),
    // so that issuing its cp.async.bulk prefetch needs no dependent global load
    int32_t nw_off, nw_len, nw_buf;
};
static_assert(sizeof(SbcOp) == 128, "SbcOp must be 32 int32 words");
static_assert(sizeof(SbcGeo) == 32, "SbcGeo must be 8 int32 words");
