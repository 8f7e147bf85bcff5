// emu.cpp -- CPU thread-emulation of the kernel's per-thread op bodies (csrc/sbc_ops.h).
// TEST-ONLY: compiled by tests/test_emulation.py with g++; runs every "thread" of a 256-thread CTA
// sequentially, phase by phase, exactly as sbc_kernel.cuh sequences them between barriers, so that
// indexing / tiling mistakes in the shared device code are caught without a GPU.  The K-split
// shuffle reduction of the kernel is replaced by an explicit sum over the ks partials.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../score_based_channels_b200/csrc/sbc_ops.h"

template <int PX, int CB>
static void conv_op(const SbcOp& op, float* arena, const float* wseg) {
    const int items = sbc_conv_items(op);
    for (int item = 0; item < items; item++) {
        float tot[PX * CB];
        for (int i = 0; i < PX * CB; i++) tot[i] = 0.f;
        for (int kp = 0; kp < op.ks; kp++) {
            float acc[PX * CB];
            sbc_conv_partial<PX, CB>(op, arena, wseg, item, kp, acc);
            for (int i = 0; i < PX * CB; i++) tot[i] += acc[i];
        }
        sbc_conv_epilogue<PX, CB>(op, arena, wseg, item, tot);
    }
}

static int conv_dispatch(const SbcOp& op, float* arena, const float* wseg) {
#define CASE(P, C) if (op.px == P && op.cb == C) { conv_op<P, C>(op, arena, wseg); return 0; }
    CASE(4, 8) CASE(2, 8) CASE(1, 8) CASE(4, 4) CASE(2, 4) CASE(1, 4)
    CASE(4, 2) CASE(2, 2) CASE(1, 2) CASE(4, 1) CASE(2, 1) CASE(1, 1)
#undef CASE
    return -1;
}

extern "C" int emu_run_program(const int32_t* op_table, int n_ops, const float* blob, float* arena, int nthr,
                               int stop_op) {
    const SbcOp* ops = reinterpret_cast<const SbcOp*>(op_table);
    for (int i = 0; i < n_ops; i++) {
        if (stop_op >= 0 && i == stop_op) break;
        const SbcOp& op = ops[i];
        const float* wseg = blob + op.w_off;
        switch (op.kind) {
            case SBC_OP_CONV:
                if (conv_dispatch(op, arena, wseg)) return -1;
                break;
            case SBC_OP_NORM_ELU:
                for (int t = 0; t < nthr; t++) sbc_norm_phaseA(op, arena, t, nthr);
                for (int t = 0; t < nthr; t++) sbc_norm_phaseB(op, arena, t, nthr);
                for (int t = 0; t < nthr; t++) sbc_norm_phaseC(op, arena, wseg, t, nthr);
                break;
            case SBC_OP_ELU:
                for (int t = 0; t < nthr; t++) sbc_elu_op(op, arena, t, nthr);
                break;
            case SBC_OP_AFFINE:
                for (int t = 0; t < nthr; t++) sbc_affine_op(op, arena, t, nthr);
                break;
            case SBC_OP_MAXPOOL5:
                for (int t = 0; t < nthr; t++) sbc_maxpool5_op(op, arena, t, nthr);
                break;
            case SBC_OP_UPACC:
                for (int t = 0; t < nthr; t++) sbc_upacc_op(op, arena, t, nthr);
                break;
            default:
                return -2;
        }
    }
    return 0;
}

// one Langevin step after the network has run (net_out already in the arena)
extern "C" float emu_langevin_step(float* arena, int in_off, int out_off, int post_off, const float* P, const float* Y,
                                   const float* Hor, const float* ext_noise, float sigma, float alpha, float den,
                                   float nscale, uint64_t seed, uint64_t sid, uint32_t gstep, int Nt, int Nr, int Np,
                                   int nthr) {
    SbcStepScalars sc{sigma, alpha, den, nscale};
    float* res = arena + post_off;
    for (int t = 0; t < nthr; t++) sbc_dc_residual(arena + in_off, res, P, Y, Nt, Nr, Np, t, nthr);
    float tot = 0.f;
    for (int t = 0; t < nthr; t++)
        tot += sbc_langevin_update(arena + in_off, arena + out_off, res, P, Hor, ext_noise, sc, seed, sid, gstep, Nt, Nr,
                                   Np, t, nthr);
    return tot;
}

extern "C" void emu_noise(uint64_t seed, uint64_t sid, uint32_t step, int n, float* out) {
    for (int e = 0; e < n; e++) sbc_noise_cn01(seed, sid, step, e, out[2 * e], out[2 * e + 1]);
}
