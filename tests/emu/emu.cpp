// emu.cpp -- CPU thread-emulation of the kernel's per-thread / per-lane op bodies (csrc/sbc_ops.h,
// csrc/sbc_mma.h).  TEST-ONLY: compiled by tests/test_emulation.py with g++; runs every "thread" of a CTA
// sequentially, phase by phase, as sbc_kernel.cuh sequences them between barriers, so that indexing /
// layout mistakes in the shared device code are caught without a GPU.  Warp-collective steps (mma.sync,
// shuffle reductions) are replaced by explicit loops over the lanes.
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../score_based_channels_b200/csrc/sbc_mma.h"
#include "../../score_based_channels_b200/csrc/sbc_ops.h"

// warp-level emulation of the tensor-core conv: every lane's fragments are gathered with the shared
// per-lane helpers (K-step offset table), the m16n8k8 product is done as plain matrices
// with the operand rounding of the device path (3xTF32: a = trunc(a) + (a - trunc(a)), the low parts
// truncated by the tensor core, weights split the same way; TF32: rna on activations, weights pre-rounded),
// then the per-lane epilogue runs.
static void conv_mma(const SbcOp& op, const SbcGeo& GS, const SbcGeo& GD, float* arena, float* park, const float* wseg, int nthr) {
    const bool x3 = (op.flags & SBC_F_X3) != 0;
    const int nq = (op.flags & SBC_F_POOL) ? 4 : 1;
    const int* steptab = reinterpret_cast<const int*>(wseg);
    for (int mt = 0; mt < op.MT; mt++)
        for (int nt = 0; nt < op.NT; nt++) {
            float D[16][8];
            memset(D, 0, sizeof D);
            for (int quad = 0; quad < nq; quad++) {
                for (int s = 0; s < op.S; s++) {
                    // MMA-index matrices: column kk of A / row kk of B is MMA K index kk
                    float Ah[16][8], Al[16][8], Bh[8][8], Bl[8][8];
                    for (int lane = 0; lane < 32; lane++) {
                        const int g = lane >> 2, t = lane & 3;
                        const int po0 = sbc_mma_row_off(op, GS, mt, quad, g);
                        const int po1 = sbc_mma_row_off(op, GS, mt, quad, g + 8);
                        float a[4];
                        sbc_mma_a_frag(arena + op.src + t, steptab[s], po0, po1, GS.pps * 4, a);
                        const int rr[4] = {g, g + 8, g, g + 8}, cc[4] = {t, t, t + 4, t + 4};
                        for (int i = 0; i < 4; i++) {
                            if (x3) {
                                Ah[rr[i]][cc[i]] = sbc_tf32_rz(a[i]);
                                Al[rr[i]][cc[i]] = sbc_tf32_rz(a[i] - Ah[rr[i]][cc[i]]);
                            } else {
                                Ah[rr[i]][cc[i]] = sbc_tf32_rn(a[i]);
                                Al[rr[i]][cc[i]] = 0.f;
                            }
                        }
                        const float* b = wseg + op.frag_rel + ((size_t)(s * op.NT + nt) * 32 + lane) * 2;
                        for (int i = 0; i < 2; i++) {
                            const int kk = t + 4 * i;
                            if (x3) {
                                Bh[kk][g] = sbc_tf32_rz(b[i]);
                                Bl[kk][g] = sbc_tf32_rz(b[i] - Bh[kk][g]);
                            } else {
                                Bh[kk][g] = b[i];
                                Bl[kk][g] = 0.f;
                            }
                        }
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
            for (int lane = 0; lane < 32; lane++) {
                const int g = lane >> 2, t = lane & 3;
                const float c[4] = {D[g][2 * t], D[g][2 * t + 1], D[g + 8][2 * t], D[g + 8][2 * t + 1]};
                int pd[2];
                sbc_mma_dst_off(op, GD, mt, g, pd);
                sbc_mma_epilogue(sbc_epi(op, GD, arena, park), arena, wseg, pd[0], pd[1], mt * 16 + g, nt, lane, c[0], c[1], c[2], c[3]);
            }
        }
}

static void norm_op(const SbcOp& op, const SbcGeo& G, float* arena, const float* wseg, int nthr) {
    const int C = op.cin, nq = C >> 2, T = sbc_norm_T(op, nthr);
    const float inv = 1.f / (float)(G.h * G.w);
    std::vector<SbcF4> mu(nq), m2s(nq);
    for (int q = 0; q < nq; q++) {
        SbcF4 sum{0, 0, 0, 0};
        for (int s = 0; s < T; s++) {
            const SbcF4 v = sbc_norm_partial_sum(op, G, arena, q, s, T);
            sum.x += v.x; sum.y += v.y; sum.z += v.z; sum.w += v.w;
        }
        SbcF4 mean{sum.x * inv, sum.y * inv, sum.z * inv, sum.w * inv};
        SbcF4 tot{0, 0, 0, 0};
        for (int s = 0; s < T; s++) {
            const SbcF4 v = sbc_norm_partial_m2(op, G, arena, q, s, T, mean);
            tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w;
        }
        mu[q] = mean; m2s[q] = tot;
    }
    for (int q = 0; q < nq; q++)
        for (int s = 0; s < T; s++)
            sbc_norm_apply(op, G, arena, wseg, reinterpret_cast<const float*>(mu.data()), q, s, T, mu[q], m2s[q]);
}

extern "C" int emu_run_program(const int32_t* op_table, int n_ops, const int32_t* geo_table, const float* blob,
                               float* arena, float* park, int nthr, int stop_op) {
    const SbcOp* ops = reinterpret_cast<const SbcOp*>(op_table);
    const SbcGeo* geo = reinterpret_cast<const SbcGeo*>(geo_table);
    for (int i = 0; i < n_ops; i++) {
        if (stop_op >= 0 && i == stop_op) break;
        const SbcOp& op = ops[i];
        const float* wseg = blob + op.w_off;
        const SbcGeo& GS = geo[op.sgeo];
        const SbcGeo& GD = geo[op.dgeo];
        switch (op.kind) {
            case SBC_OP_CONV_MMA:
                conv_mma(op, GS, GD, arena, park, wseg, nthr);
                break;
            case SBC_OP_NORM_ELU:
                norm_op(op, GS, arena, wseg, nthr);
                break;
            case SBC_OP_ELU:
                for (int t = 0; t < nthr; t++) sbc_elu_op(op, GS, arena, t, nthr);
                break;
            case SBC_OP_AFFINE:
                for (int t = 0; t < nthr; t++) sbc_affine_op(op, GS, arena, t, nthr);
                break;
            case SBC_OP_MAXPOOL5:
                for (int t = 0; t < nthr; t++) sbc_maxpool5_op(op, GS, arena, t, nthr);
                break;
            case SBC_OP_UPACC:
                for (int t = 0; t < nthr; t++) sbc_upacc_op(op, GS, GD, arena, t, nthr);
                break;
            case SBC_OP_SPILL:
                for (int t = 0; t < nthr; t++) sbc_spill_op(op, arena, park, t, nthr);
                break;
            case SBC_OP_FILL:
                for (int t = 0; t < nthr; t++) sbc_fill_op(op, arena, park, t, nthr);
                break;
            default:
                return -2;
        }
        // the caller of an op body re-zeroes the halos the planner flagged (see sbc_ops.h)
        for (int t = 0; t < nthr; t++) {
            if (op.flags & SBC_F_ZH_DST) sbc_zero_halo(arena + op.dst, GD, op.cout, t, nthr);
            if (op.flags & SBC_F_ZH_EDST) sbc_zero_halo(arena + op.edst, GD, op.cout, t, nthr);
        }
    }
    return 0;
}

// one Langevin step after the network has run (net_out already in the arena)
extern "C" float emu_langevin_step(float* arena, int in_off, int out_off, int post_off,
                                   const float* P, const float* Y, const float* Hor, const float* ext_noise,
                                   float sigma, float alpha, float den, float nscale, uint64_t seed, uint64_t sid,
                                   uint32_t gstep, int Nt, int Nr, int Np, int nthr) {
    SbcStepScalars sc{sigma, alpha, den, nscale};
    float* res = arena + post_off;
    for (int t = 0; t < nthr; t++) sbc_dc_residual(arena + in_off, res, P, Y, Nt, Nr, Np, t, nthr);
    float tot = 0.f;
    for (int t = 0; t < nthr; t++)
        tot += sbc_langevin_update(arena + in_off, arena + out_off, res, P, Hor, ext_noise, sc, seed, sid, gstep, Nt,
                                   Nr, Np, t, nthr);
    return tot;
}

extern "C" void emu_noise(uint64_t seed, uint64_t sid, uint32_t step, int n, float* out) {
    for (int e = 0; e < n; e++) sbc_noise_cn01(seed, sid, step, e, out[2 * e], out[2 * e + 1]);
}
