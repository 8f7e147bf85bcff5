// sbc_kernel.cuh -- persistent fused kernel: one CTA owns one channel realisation at a time and
// runs, entirely out of shared memory, the whole NCSNv2Deepest forward (161-op layer program,
// reference ncsnv2/models/ncsnv2.py:269-300) followed by the data-consistency gradient, the
// Langevin update, the Philox noise draw and the per-step NMSE (reference test_score.py:135-171)
// for every (sigma level, inner step) of the requested range.  Nothing returns to the host inside
// the loop; the state x never leaves the SM.
//
// Per-op parameters (B fragments, biases, norm affine) are streamed global -> shared one op ahead with
// cp.async.bulk (TMA bulk copy, SASS UBLKCP) completing on an mbarrier; their staging buffers are planned
// into the same arena as the activations (program.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sbc_mma.h"
#include "sbc_ops.h"

#define SBC_NTHREADS 512

struct SbcLaunch {
    // layer program
    const SbcOp* ops;
    int n_ops;
    int first_w;             // index of the first op with parameters
    const float* blob;       // packed parameters
    SbcGeo geo[SBC_MAX_GEO]; // tensor geometries (geo[0] = network input / output resolution)
    int arena_floats, in_off, out_off, post_off;
    int Nt, Nr, channels, max_w_len;
    int mode;                // 0 = forward (NCSNv2Deepest.forward), 1 = annealed Langevin
    int B;
    // forward mode
    const float* fx;         // [B,2,Nt,Nr] with element strides fxs
    long long fxs[4];
    const long long* labels; // [B]
    float* fout;             // [B,2,Nt,Nr] contiguous
    const float* sigmas;     // [n_sigmas]
    int n_sigmas;
    // ALD mode
    int Np, level_begin, level_end, steps_each;
    const float* P;          // [B,Np,Nt] complex64
    const float* Y;          // [B,Np,Nr] complex64
    float* X;                // [B,Nt,Nr] complex64 in/out
    const float* Hor;        // [B,Nt,Nr] complex64 or null
    const float* noise_var;  // [B]
    const float* alpha_step; // [B]
    const float* beta;       // [B]
    double sigma_end;
    float* nmse_log;         // [steps,B] or null
    unsigned long long seed;
    const unsigned long long* sample_ids;  // [B] or null
    const float* ext_noise;  // [steps,B,Nt,Nr] complex64 or null
    // execution
    float* gws;              // global arena workspace (when the arena does not fit in shared memory)
    int stage_weights;       // 1: cp.async.bulk double buffering, 0: read parameters from global/L2
    int debug_stop;          // >=0: stop sample 0 / step 0 before op `debug_stop`, dump the arena
    float* debug_out;
    long long* prof;         // optional [n_ops+2] clock64() stamps of CTA 0, first sample, first step
};

// ---------------------------------------------------------------------------------------------
// mbarrier + bulk async copy (PTX)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sbc_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sbc_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sbc_smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void sbc_fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void sbc_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sbc_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void sbc_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "SBC_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra SBC_DONE_%=;\n"
        "bra SBC_WAIT_%=;\n"
        "SBC_DONE_%=:\n"
        "}\n" ::"r"(sbc_smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void sbc_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     sbc_smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(sbc_smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------
// tensor-core conv (SBC_OP_CONV_MMA): warp-level implicit GEMM on mma.sync m16n8k8 TF32.
//   X3 = true : 3xTF32 split  (a = a_hi + a_lo, b = b_hi + b_lo;  D += a_lo b_hi + a_hi b_lo + a_hi b_hi)
//               -> fp32-equivalent accuracy (the parity mode);  X3 = false: plain TF32 operands.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sbc_mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
          "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}

#define SBC_MMA_SLOTS 4   // (pixel tile, pooling position) accumulators per warp per pass
#define SBC_MMA_NTW 2     // cout tiles that share one A gather

__device__ __forceinline__ float sbc_dev_tf32_rn(float x) {
    return __uint_as_float((__float_as_uint(x) + 0x1000u) & 0xFFFFE000u);
}
__device__ __forceinline__ float sbc_dev_tf32_rz(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }

// Accumulate K steps [s0, s1) for up to 4 slots (slot j = tile mt0 + (j / nq) * mt_stride, pooling position
// j % nq) and `ntw` (1 or 2) cout tiles nt0, nt0+1 that share each gathered A fragment.  B fragments come from
// the staged parameter segment.  A K step costs per slot: 2 address adds + 4 LDS + the operand split.
template <bool X3>
__device__ __forceinline__ void sbc_mma_pass(const SbcOp& op, const SbcMmaGeom& M, const SbcGeo& GS,
                                             const float* arena, const float* wseg, int mt0, int mt_stride,
                                             int nslots, int nt0, int ntw, int s0, int s1, int lane,
                                             float (&acc)[SBC_MMA_SLOTS][SBC_MMA_NTW][4]) {
    const int g = lane >> 2, t = lane & 3;
    int po0[SBC_MMA_SLOTS], po1[SBC_MMA_SLOTS];
#pragma unroll
    for (int j = 0; j < SBC_MMA_SLOTS; j++) {
        po0[j] = po1[j] = 0;
        if (j < nslots) {
            const int mt = mt0 + (j / M.nq) * mt_stride, quad = j % M.nq;
            po0[j] = sbc_mma_row_off(op, M, GS, mt, quad, g);
            po1[j] = sbc_mma_row_off(op, M, GS, mt, quad, g + 8);
        }
    }
    constexpr int E = X3 ? 4 : 2;
    const int k = op.ksize, r = k >> 1, dil = op.dil;
    const int cg1 = GS.pps * 4;
    const float* src = arena + op.src + t;
    int s = 0;
    for (int tap = 0; tap < k * k; tap++) {
        if (!((op.tapmask >> tap) & 1)) continue;
        const int ky = tap / k, kx = tap - ky * k;
        const int toff = ((ky - r) * dil * GS.wp + (kx - r) * dil) * 4;
        for (int kc = 0; kc < M.KC; kc++, s++) {
            if (s < s0 || s >= s1) continue;
            const float* bp = src + 2 * kc * cg1 + toff;
            float bh[SBC_MMA_NTW][2], bl[SBC_MMA_NTW][2];
#pragma unroll
            for (int n = 0; n < SBC_MMA_NTW; n++) {
                bh[n][0] = bh[n][1] = bl[n][0] = bl[n][1] = 0.f;
                if (n < ntw) {
                    const float* wp = wseg + ((size_t)(s * M.NT + nt0 + n) * 32 + lane) * E;
                    if (X3) {
                        const float4 b = *reinterpret_cast<const float4*>(wp);
                        bh[n][0] = b.x; bh[n][1] = b.y; bl[n][0] = b.z; bl[n][1] = b.w;
                    } else {
                        const float2 b = *reinterpret_cast<const float2*>(wp);
                        bh[n][0] = b.x; bh[n][1] = b.y;
                    }
                }
            }
#pragma unroll
            for (int j = 0; j < SBC_MMA_SLOTS; j++) {
                if (j < nslots) {
                    float a[4];
                    a[0] = bp[po0[j]];
                    a[1] = bp[po1[j]];
                    a[2] = bp[cg1 + po0[j]];
                    a[3] = bp[cg1 + po1[j]];
                    float ah[4], al[4];
                    if (X3) {
#pragma unroll
                        for (int i = 0; i < 4; i++) {
                            ah[i] = sbc_dev_tf32_rz(a[i]);                 // a = ah + (a - ah) exactly
                            al[i] = sbc_dev_tf32_rz(a[i] - ah[i]);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 4; i++) ah[i] = sbc_dev_tf32_rn(a[i]);
                    }
#pragma unroll
                    for (int n = 0; n < SBC_MMA_NTW; n++) {
                        if (n < ntw) {
                            if (X3) {
                                sbc_mma_tf32(acc[j][n], al, bh[n][0], bh[n][1]);   // small terms first
                                sbc_mma_tf32(acc[j][n], ah, bl[n][0], bl[n][1]);
                            }
                            sbc_mma_tf32(acc[j][n], ah, bh[n][0], bh[n][1]);
                        }
                    }
                }
            }
        }
    }
}

template <bool X3>
__device__ __forceinline__ void sbc_conv_mma(const SbcOp& op, const SbcGeo& GS, const SbcGeo& GD, float* arena,
                                             const float* wseg, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = SBC_NTHREADS / 32;
    SbcMmaGeom M;
    sbc_mma_geom(op, M);
    float acc[SBC_MMA_SLOTS][SBC_MMA_NTW][4];
    // fresh outputs get their halo re-zeroed (the arena region may have held another tensor)
    if (op.dst >= 0) sbc_zero_halo(arena + op.dst, GD, op.cout, tid, SBC_NTHREADS);
    if (op.edst >= 0) sbc_zero_halo(arena + op.edst, GD, op.cout, tid, SBC_NTHREADS);

    if (op.ks > 1) {
        // fewer (pixel tile, cout tile) units than warps: `ks` warps split the K steps of one unit and the
        // partial accumulators are combined through shared memory
        const int ks = op.ks, units = M.MT * M.NT;
        const int u = warp / ks, kp = warp - u * ks;
        const int mt = u / M.NT, nt = u - mt * M.NT;
        float4* part = reinterpret_cast<float4*>(arena + op.scratch);
        if (u < units) {
#pragma unroll
            for (int j = 0; j < SBC_MMA_SLOTS; j++)
#pragma unroll
                for (int n = 0; n < SBC_MMA_NTW; n++) acc[j][n][0] = acc[j][n][1] = acc[j][n][2] = acc[j][n][3] = 0.f;
            sbc_mma_pass<X3>(op, M, GS, arena, wseg, mt, 0, M.nq, nt, 1, (M.S * kp) / ks, (M.S * (kp + 1)) / ks, lane,
                             acc);
            float4 c = make_float4(acc[0][0][0], acc[0][0][1], acc[0][0][2], acc[0][0][3]);
            if (M.nq == 4) {
#pragma unroll
                for (int j = 1; j < 4; j++) {
                    c.x += acc[j][0][0]; c.y += acc[j][0][1]; c.z += acc[j][0][2]; c.w += acc[j][0][3];
                }
            }
            part[warp * 32 + lane] = c;
        }
        __syncthreads();
        if (u < units && kp == 0) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i = 0; i < ks; i++) {
                const float4 p = part[(warp + i) * 32 + lane];
                c[0] += p.x; c[1] += p.y; c[2] += p.z; c[3] += p.w;
            }
            sbc_mma_epilogue(op, GD, arena, wseg, mt, nt, lane, c);
        }
        return;
    }

    // enough pixel tiles for every warp: warp w owns tiles w, w+NW, ...; cout tiles are processed in pairs that
    // share the gathered A fragments
    const int tpp = SBC_MMA_SLOTS / M.nq;                    // tiles per pass
    for (int mt0 = warp; mt0 < M.MT; mt0 += tpp * NW) {
        int ntile = (M.MT - mt0 + NW - 1) / NW;
        if (ntile > tpp) ntile = tpp;
        for (int nt0 = 0; nt0 < M.NT; nt0 += SBC_MMA_NTW) {
            const int ntw = (M.NT - nt0 < SBC_MMA_NTW) ? M.NT - nt0 : SBC_MMA_NTW;
#pragma unroll
            for (int j = 0; j < SBC_MMA_SLOTS; j++)
#pragma unroll
                for (int n = 0; n < SBC_MMA_NTW; n++) acc[j][n][0] = acc[j][n][1] = acc[j][n][2] = acc[j][n][3] = 0.f;
            sbc_mma_pass<X3>(op, M, GS, arena, wseg, mt0, NW, ntile * M.nq, nt0, ntw, 0, M.S, lane, acc);
#pragma unroll
            for (int n = 0; n < SBC_MMA_NTW; n++) {
                if (n >= ntw) continue;
                if (M.nq == 1) {
#pragma unroll
                    for (int j = 0; j < SBC_MMA_SLOTS; j++)
                        if (j < ntile) sbc_mma_epilogue(op, GD, arena, wseg, mt0 + j * NW, nt0 + n, lane, acc[j][n]);
                } else if (ntile > 0) {
                    float c[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) c[i] = acc[0][n][i] + acc[1][n][i] + acc[2][n][i] + acc[3][n][i];
                    sbc_mma_epilogue(op, GD, arena, wseg, mt0, nt0 + n, lane, c);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm++ + ELU: T threads per channel group, two statistics passes (warp shuffles + one
// shared-memory exchange each), then the fused normalise / affine / ELU pass.
// scratch: red[2][NW] float4 | mu[C] | m2[C]
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ SbcF4 sbc_warp_sum4(SbcF4 v) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        v.x += __shfl_xor_sync(0xffffffffu, v.x, off);
        v.y += __shfl_xor_sync(0xffffffffu, v.y, off);
        v.z += __shfl_xor_sync(0xffffffffu, v.z, off);
        v.w += __shfl_xor_sync(0xffffffffu, v.w, off);
    }
    return v;
}

__device__ __forceinline__ void sbc_norm_op(const SbcOp& op, const SbcGeo& G, float* arena, const float* wseg,
                                            int tid) {
    constexpr int NW = SBC_NTHREADS / 32;
    const int C = op.cin, ncg = (C + 3) >> 2;
    const int T = sbc_norm_T(op, SBC_NTHREADS), gpp = SBC_NTHREADS / T, wpg = T / 32;   // groups / pass, warps / group
    const int warp = tid >> 5, lane = tid & 31;
    SbcF4* red = reinterpret_cast<SbcF4*>(arena + op.scratch);
    float* mu = arena + op.scratch + 2 * NW * 4;
    float* m2s = mu + C;
    const float inv = 1.f / (float)(G.h * G.w);
    const int s = tid % T;
    for (int cg0 = 0; cg0 < ncg; cg0 += gpp) {
        const int cg = cg0 + tid / T;
        const bool active = cg < ncg;
        SbcF4 z{0.f, 0.f, 0.f, 0.f};
        SbcF4 sum = active ? sbc_norm_partial_sum(op, G, arena, cg, s, T) : z;
        sum = sbc_warp_sum4(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        SbcF4 mean = z;
        const int w0 = (tid / T) * wpg;
        for (int i = 0; i < wpg; i++) { const SbcF4 v = red[w0 + i]; mean.x += v.x; mean.y += v.y; mean.z += v.z; mean.w += v.w; }
        mean.x *= inv; mean.y *= inv; mean.z *= inv; mean.w *= inv;
        SbcF4 m2 = active ? sbc_norm_partial_m2(op, G, arena, cg, s, T, mean) : z;
        m2 = sbc_warp_sum4(m2);
        if (lane == 0) red[NW + warp] = m2;
        __syncthreads();
        if (active && s == 0) {
            SbcF4 tot = z;
            for (int i = 0; i < wpg; i++) { const SbcF4 v = red[NW + w0 + i]; tot.x += v.x; tot.y += v.y; tot.z += v.z; tot.w += v.w; }
            const float mm[4] = {mean.x, mean.y, mean.z, mean.w}, tt[4] = {tot.x, tot.y, tot.z, tot.w};
            for (int j = 0; j < 4; j++)
                if (4 * cg + j < C) { mu[4 * cg + j] = mm[j]; m2s[4 * cg + j] = tt[j]; }
        }
        __syncthreads();
    }
    for (int cg0 = 0; cg0 < ncg; cg0 += gpp) {
        const int cg = cg0 + tid / T;
        if (cg < ncg) {
            SbcF4 mean, m2;
            mean.x = mu[4 * cg]; m2.x = m2s[4 * cg];
            mean.y = (4 * cg + 1 < C) ? mu[4 * cg + 1] : 0.f; m2.y = (4 * cg + 1 < C) ? m2s[4 * cg + 1] : 0.f;
            mean.z = (4 * cg + 2 < C) ? mu[4 * cg + 2] : 0.f; m2.z = (4 * cg + 2 < C) ? m2s[4 * cg + 2] : 0.f;
            mean.w = (4 * cg + 3 < C) ? mu[4 * cg + 3] : 0.f; m2.w = (4 * cg + 3 < C) ? m2s[4 * cg + 3] : 0.f;
            sbc_norm_apply(op, G, arena, wseg, mu, cg, s, T, mean, m2);
        }
    }
    sbc_zero_halo(arena + op.dst, G, C, tid, SBC_NTHREADS);
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
template <bool SMEM_ARENA>
__global__ void __launch_bounds__(SBC_NTHREADS, 1) sbc_ald_kernel(const __grid_constant__ SbcLaunch L) {
    extern __shared__ __align__(128) unsigned char sbc_smem_raw[];
    float* smem_f = reinterpret_cast<float*>(sbc_smem_raw);
    const int tid = threadIdx.x;

    float* arena;
    size_t off = 0;
    if (SMEM_ARENA) {
        arena = smem_f;
        off = (size_t)L.arena_floats;
    } else {
        arena = L.gws + (size_t)blockIdx.x * (size_t)L.arena_floats;
    }
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_f + off);   // off is a multiple of 4 floats

    __shared__ float s_hnorm;
    __shared__ SbcStepScalars s_sc;

    // parameter segments are staged into the (shared-memory) arena with cp.async.bulk one op ahead
    const bool stage = SMEM_ARENA && L.stage_weights != 0;
    if (stage && tid == 0) {
        sbc_mbar_init(&bars[0], 1);
        sbc_mbar_init(&bars[1], 1);
        sbc_fence_barrier_init();
    }
    __syncthreads();

    const int Nt = L.Nt, Nr = L.Nr, ne = Nt * Nr;
    const SbcGeo& G0 = L.geo[0];
    float* ax = arena + L.in_off;             // x: 2-channel tensor (re, im) of geometry 0
    uint32_t wcount = 0;                      // parameter segments consumed so far (same in every thread)

    if (stage && tid == 0 && (int)blockIdx.x < L.B && L.first_w >= 0) {
        const SbcOp& o = L.ops[L.first_w];
        sbc_mbar_expect_tx(&bars[0], (uint32_t)o.w_len * 4u);
        sbc_bulk_g2s(arena + o.wbuf, L.blob + o.w_off, (uint32_t)o.w_len * 4u, &bars[0]);
    }

    const int nsteps = (L.mode == 1) ? (L.level_end - L.level_begin) * L.steps_each : 1;

    for (int b = blockIdx.x; b < L.B; b += gridDim.x) {
        const bool last_sample = (b + (int)gridDim.x >= L.B);
        // ---------------- load the sample state into the arena ----------------
        if (L.mode == 1) {
            const float* X = L.X + (size_t)b * ne * 2;
            for (int e = tid; e < ne; e += SBC_NTHREADS) {
                const float2 v = reinterpret_cast<const float2*>(X)[e];
                const int t = e / Nr, r = e - t * Nr;
                *sbc_px(ax, G0, 0, t, r) = SbcF4{v.x, v.y, 0.f, 0.f};
            }
            if (L.Hor) {   // ||H||_F^2 once per sample (test_score.py:169)
                const float* Hc = L.Hor + (size_t)b * ne * 2;
                float part = 0.f;
                for (int e = tid; e < ne; e += SBC_NTHREADS) {
                    const float2 v = reinterpret_cast<const float2*>(Hc)[e];
                    part += v.x * v.x + v.y * v.y;
                }
                float* red = arena + L.post_off + 2 * ne;
                red[tid] = part;
                __syncthreads();
                if (tid == 0) {
                    float sum = 0.f;
                    for (int i = 0; i < SBC_NTHREADS; i++) sum += red[i];
                    s_hnorm = sum;
                }
            }
        } else {
            const float* fx = L.fx + (size_t)b * L.fxs[0];
            for (int e = tid; e < ne; e += SBC_NTHREADS) {
                const int t = e / Nr, r = e - t * Nr;
                const float* q = fx + t * L.fxs[2] + r * L.fxs[3];
                *sbc_px(ax, G0, 0, t, r) = SbcF4{q[0], q[L.fxs[1]], 0.f, 0.f};
            }
        }
        __syncthreads();

        for (int gs = 0; gs < nsteps; gs++) {
            const bool last_step = (gs + 1 == nsteps);
            int lvl = 0;
            if (L.mode == 1) {
                lvl = L.level_begin + gs / L.steps_each;
                if (tid == 0 && (gs % L.steps_each) == 0) {   // per-level scalars, in double like the reference
                    const double sigma = (double)L.sigmas[lvl];
                    const double ratio = sigma / L.sigma_end;
                    const double alpha = (double)L.alpha_step[b] * ratio * ratio;
                    s_sc.sigma = L.sigmas[lvl];
                    s_sc.alpha = (float)alpha;
                    s_sc.den = (float)((double)L.noise_var[b] / 2. + sigma * sigma);
                    s_sc.nscale = (float)sqrt(2. * alpha * (double)L.beta[b]);
                }
            }

            // ---------------- the network: walk the layer program ----------------
            const bool do_prof = (L.prof != nullptr) && blockIdx.x == 0 && b == 0 && gs == 0 && tid == 0;
            for (int i = 0; i < L.n_ops; i++) {
                if (L.debug_stop >= 0 && i == L.debug_stop) break;
                if (do_prof) L.prof[i] = clock64();
                const SbcOp op = L.ops[i];
                const float* wseg = L.blob + op.w_off;
                if (op.w_len > 0 && stage) {
                    const uint32_t slot = wcount & 1u;
                    if (tid == 0) {   // prefetch the next parameter segment into its staging buffer
                        int j = op.pad0;                                        // next op with parameters
                        if (j < 0 && !(last_step && last_sample)) j = L.first_w;   // wraps into the next forward
                        if (j >= 0) {
                            const SbcOp& o = L.ops[j];
                            sbc_mbar_expect_tx(&bars[slot ^ 1u], (uint32_t)o.w_len * 4u);
                            sbc_bulk_g2s(arena + o.wbuf, L.blob + o.w_off, (uint32_t)o.w_len * 4u, &bars[slot ^ 1u]);
                        }
                    }
                    sbc_mbar_wait(&bars[slot], (wcount >> 1) & 1u);
                    wseg = arena + op.wbuf;
                    wcount++;
                }
                const SbcGeo& GS = L.geo[op.sgeo];
                const SbcGeo& GD = L.geo[op.dgeo];
                switch (op.kind) {
                    case SBC_OP_CONV_MMA:
                        if (op.flags & SBC_F_X3) sbc_conv_mma<true>(op, GS, GD, arena, wseg, tid);
                        else sbc_conv_mma<false>(op, GS, GD, arena, wseg, tid);
                        break;
                    case SBC_OP_NORM_ELU:
                        sbc_norm_op(op, GS, arena, wseg, tid);
                        break;
                    case SBC_OP_ELU:
                        sbc_elu_op(op, GS, arena, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_AFFINE:
                        sbc_affine_op(op, GS, arena, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_MAXPOOL5:
                        sbc_maxpool5_op(op, GS, arena, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_UPACC:
                        sbc_upacc_op(op, GS, GD, arena, tid, SBC_NTHREADS);
                        break;
                    default:
                        break;
                }
                __syncthreads();
            }
            if (L.debug_stop >= 0) {   // debugging aid: dump the arena of sample 0 and stop
                if (b == 0)
                    for (int i = tid; i < L.arena_floats; i += SBC_NTHREADS) L.debug_out[i] = arena[i];
                return;
            }

            if (do_prof) L.prof[L.n_ops] = clock64();
            const float* net = arena + L.out_off;
            if (L.mode == 0) {
                // score = net / sigmas[y]   (ncsnv2.py:295-298)
                long long lab = L.labels[b];
                if (lab < 0) lab = 0;
                if (lab >= L.n_sigmas) lab = L.n_sigmas - 1;
                const float sg = L.sigmas[lab];
                float* o = L.fout + (size_t)b * L.channels * ne;
                for (int e = tid; e < ne; e += SBC_NTHREADS) {
                    const int t = e / Nr, r = e - t * Nr;
                    const SbcF4 v = *sbc_px(net, G0, 0, t, r);
                    o[e] = v.x / sg;
                    o[ne + e] = v.y / sg;
                }
            } else {
                // ---------------- data-consistency gradient, Langevin update, NMSE ----------------
                const float* Pm = L.P + (size_t)b * L.Np * Nt * 2;
                const float* Ym = L.Y + (size_t)b * L.Np * Nr * 2;
                const float* Hc = L.Hor ? L.Hor + (size_t)b * ne * 2 : nullptr;
                const float* en = L.ext_noise ? L.ext_noise + ((size_t)gs * L.B + b) * ne * 2 : nullptr;
                float* res = arena + L.post_off;
                float* red = res + 2 * ne;
                sbc_dc_residual(ax, G0, res, Pm, Ym, Nt, Nr, L.Np, tid, SBC_NTHREADS);
                __syncthreads();
                const unsigned long long sid = L.sample_ids ? L.sample_ids[b] : (unsigned long long)b;
                const uint32_t gstep = (uint32_t)(lvl * L.steps_each + gs % L.steps_each);
                const float part = sbc_langevin_update(ax, net, G0, res, Pm, Hc, en, s_sc, L.seed, sid, gstep, Nt, Nr,
                                                       L.Np, tid, SBC_NTHREADS);
                if (L.nmse_log && Hc) {
                    red[tid] = part;
                    __syncthreads();
                    if (tid == 0) {
                        float sum = 0.f;
                        for (int i = 0; i < SBC_NTHREADS; i++) sum += red[i];
                        L.nmse_log[(size_t)gs * L.B + b] = sum / s_hnorm;
                    }
                }
            }
            __syncthreads();
            if (do_prof) L.prof[L.n_ops + 1] = clock64();
        }

        if (L.mode == 1) {   // write the final estimate back (interleaved complex64)
            float* X = L.X + (size_t)b * ne * 2;
            for (int e = tid; e < ne; e += SBC_NTHREADS) {
                const int t = e / Nr, r = e - t * Nr;
                const SbcF4 v = *sbc_px(ax, G0, 0, t, r);
                reinterpret_cast<float2*>(X)[e] = make_float2(v.x, v.y);
            }
        }
        __syncthreads();
    }
}
