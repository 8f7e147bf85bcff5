// sbc_program.h -- op table shared by the host program builder (program.py) and the kernel.
// One SbcOp = 24 int32 words, in the order of program.py:OP_FIELDS.
#pragma once
#include <stdint.h>

enum SbcOpKind : int32_t {
    SBC_OP_AFFINE = 0,    // dst = 2*src - 1                     (reference ncsnv2/models/ncsnv2.py:270-271)
    SBC_OP_CONV = 1,      // Conv2d k in {1,3}, stride 1, pad = dil*(k/2)      (layers.py:28-60)
    SBC_OP_NORM_ELU = 2,  // dst = ELU(InstanceNorm2dPlus(src))   (normalization.py:163-176, layers.py:13)
    SBC_OP_ELU = 3,       // dst = ELU(src)
    SBC_OP_MAXPOOL5 = 4,  // MaxPool2d(5, stride 1, pad 2)        (layers.py:70)
    SBC_OP_UPACC = 5,     // acc += bilinear(src -> oh x ow, align_corners=True)   (layers.py:182-183)
    SBC_OP_CONV_MMA = 6,  // SBC_OP_CONV contract, contraction on tensor cores (mma.sync m16n8k8 TF32)
    SBC_OP_LAST = 6,
};

// Every activation plane [h][w] is stored with stride h*w + SBC_PLANE_PAD floats (= 8 mod 16), so
// that the (channel, pixel) gather of an MMA A-fragment touches 32 distinct banks.
#define SBC_PLANE_PAD 8
#define SBC_PS(h, w) ((h) * (w) + SBC_PLANE_PAD)

enum SbcOpFlags : int32_t {
    SBC_F_POOL = 1,       // conv followed by the 2x2 mean-pool of ConvMeanPool (layers.py:309-313)
    SBC_F_X3 = 2,         // CONV_MMA: 3xTF32 error-compensated product (fp32-equivalent accuracy)
};

struct SbcOp {
    int32_t kind, flags;
    int32_t src, dst, acc, edst;   // arena float offsets, -1 = unused
    int32_t cin, cout;
    int32_t h, w;                  // input spatial size
    int32_t ksize, dil;
    int32_t w_off, w_len, b_rel;   // parameter segment (floats) in the blob; bias offset inside it
    int32_t px, cb, ks;            // conv tiling (pixels / couts per thread, Cin split; MMA: K split over warps)
    int32_t scratch;               // arena offset of op scratch (norm statistics, MMA K-split partials)
    int32_t oh, ow;                // output spatial size
    int32_t pad0;                  // filled by sbc_model_create: index of the next op with staged parameters
    int32_t tapmask;               // CONV_MMA: live taps of the k x k window (bit = ky*k + kx)
    int32_t pad2;
};
static_assert(sizeof(SbcOp) == 96, "SbcOp must be 24 int32 words");
