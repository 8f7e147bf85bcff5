// emu.cpp -- CPU thread-emulation of the kernel's per-thread op bodies (csrc/sbc_ops.h).
// TEST-ONLY: compiled by tests/test_emulation.py with g++; runs every "thread" of a 256-thread CTA
// sequentially, phase by phase, exactly as sbc_kernel.cuh sequences them between barriers, so that
// indexing / tiling mistakes in the shared device code are caught without a GPU.  The K-split
// shuffle reduction of the kernel is replaced by an explicit sum over the ks partials.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../score_based_channels_b200/csrc/sbc_mma.h"
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

// warp-level emulation of the tensor-core conv: every lane's fragments are gathered with the shared
// per-lane helpers (csrc/sbc_mma.h), the m16n8k8 product is done as plain matrices (operands rounded
// to TF32 like the hardware path; 3-term split when SBC_F_X3), then the per-lane epilogue runs.
static void conv_mma(const SbcOp& op, float* arena, const float* wseg) {
    SbcMmaGeom G;
    sbc_mma_geom(op, G);
    const bool x3 = (op.flags & SBC_F_X3) != 0;
    const float* wfrag = wseg;
    const int E = x3 ? 4 : 2;
    const int k = op.ksize, r = k / 2;
    for (int mt = 0; mt < G.MT; mt++)
        for (int nt = 0; nt < G.NT; nt++) {
            float D[16][8];
            memset(D, 0, sizeof D);
            for (int quad = 0; quad < G.nq; quad++) {
                int s = 0;
                for (int tap = 0; tap < k * k; tap++) {
                    if (!((op.tapmask >> tap) & 1)) continue;
                    const int dy = (tap / k - r) * op.dil, dx = (tap % k - r) * op.dil;
                    for (int kc = 0; kc < G.KC; kc++, s++) {
                        float Ah[16][8], Al[16][8], Bh[8][8], Bl[8][8];
                        for (int lane = 0; lane < 32; lane++) {
                            const int g = lane >> 2, t = lane & 3;
                            int iy0, ix0, iy1, ix1;
                            bool ok0, ok1;
                            sbc_mma_row(op, G, mt, quad, g, iy0, ix0, ok0);
                            sbc_mma_row(op, G, mt, quad, g + 8, iy1, ix1, ok1);
                            float a[4];
                            sbc_mma_a_frag(op, arena, iy0, ix0, ok0, iy1, ix1, ok1, dy, dx, kc, lane, a);
                            const int rr[4] = {g, g + 8, g, g + 8}, cc[4] = {t, t, t + 4, t + 4};
                            for (int i = 0; i < 4; i++) {
                                Ah[rr[i]][cc[i]] = sbc_tf32(a[i]);
                                Al[rr[i]][cc[i]] = sbc_tf32(a[i] - Ah[rr[i]][cc[i]]);
                            }
                            const float* b = wfrag + ((size_t)(s * G.NT + nt) * 32 + lane) * E;
                            Bh[t][g] = b[0]; Bh[t + 4][g] = b[1];
                            Bl[t][g] = x3 ? b[2] : 0.f; Bl[t + 4][g] = x3 ? b[3] : 0.f;
                        }
                        for (int m = 0; m < 16; m++)
                            for (int n = 0; n < 8; n++) {
                                float d = D[m][n];
                                if (x3) {
                                    for (int kk = 0; kk < 8; kk++) d += Al[m][kk] * Bh[kk][n];
                                    for (int kk = 0; kk < 8; kk++) d += Ah[m][kk] * Bl[kk][n];
                                }
                                for (int kk = 0; kk < 8; kk++) d += Ah[m][kk] * Bh[kk][n];
                                D[m][n] = d;
                            }
                    }
                }
            }
            for (int lane = 0; lane < 32; lane++) {
                const int g = lane >> 2, t = lane & 3;
                const float c[4] = {D[g][2 * t], D[g][2 * t + 1], D[g + 8][2 * t], D[g + 8][2 * t + 1]};
                sbc_mma_epilogue(op, arena, wseg, mt, nt, lane, c);
            }
        }
}

static void norm_op(const SbcOp& op, float* arena, const float* wseg, int nthr) {
    const int S = sbc_norm_S(op, nthr);
    for (int c = 0; c < op.cin; c++) {
        float sum = 0.f;
        for (int s = 0; s < S; s++) sum += sbc_norm_partial_sum(op, arena, c, s, S);
        const float mean = sum * (1.f / (float)(op.h * op.w));
        float m2 = 0.f;
        for (int s = 0; s < S; s++) m2 += sbc_norm_partial_m2(op, arena, c, s, S, mean);
        sbc_norm_store_stats(op, arena, c, mean, m2);
    }
    for (int t = 0; t < nthr; t++) sbc_norm_apply(op, arena, wseg, t, nthr);
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
            case SBC_OP_CONV_MMA:
                conv_mma(op, arena, wseg);
                break;
            case SBC_OP_NORM_ELU:
                norm_op(op, arena, wseg, nthr);
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
