// sbc_kernel.cuh -- persistent fused kernel: one CTA owns one channel realisation at a time and
// runs, entirely out of shared memory, the whole NCSNv2Deepest forward (161-op layer program,
// reference ncsnv2/models/ncsnv2.py:269-300) followed by the data-consistency gradient, the
// Langevin update, the Philox noise draw and the per-step NMSE (reference test_score.py:135-171)
// for every (sigma level, inner step) of the requested range.  Nothing returns to the host inside
// the loop; the state x never leaves the SM.
//
// Per-op parameters are streamed global -> shared one op ahead with cp.async.bulk (TMA bulk copy,
// SASS UBLKCP) completing on an mbarrier, double buffered.
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
// conv op: K-split partials summed with warp shuffles across `ks` adjacent lanes
// ---------------------------------------------------------------------------------------------
template <int PX, int CB>
__device__ __forceinline__ void sbc_conv_op(const SbcOp& op, float* arena, const float* wseg, int tid) {
    const int ks = op.ks;
    const int total = sbc_conv_items(op) * ks;
    for (int base = 0; base < total; base += SBC_NTHREADS) {   // uniform trip count: shuffles are warp-wide
        const int t = base + tid;
        const bool valid = t < total;
        const int item = t / ks, kpart = t - item * ks;
        float acc[PX * CB];
        if (valid) {
            sbc_conv_partial<PX, CB>(op, arena, wseg, item, kpart, acc);
        } else {
#pragma unroll
            for (int i = 0; i < PX * CB; i++) acc[i] = 0.f;
        }
        for (int off = 1; off < ks; off <<= 1) {
#pragma unroll
            for (int i = 0; i < PX * CB; i++) acc[i] += __shfl_xor_sync(0xffffffffu, acc[i], off);
        }
        if (valid && kpart == 0) sbc_conv_epilogue<PX, CB>(op, arena, wseg, item, acc);
    }
}

__device__ __forceinline__ void sbc_conv_dispatch(const SbcOp& op, float* arena, const float* wseg, int tid) {
#define SBC_CASE(PXv, CBv) \
    if (op.px == PXv && op.cb == CBv) { sbc_conv_op<PXv, CBv>(op, arena, wseg, tid); return; }
    SBC_CASE(4, 8) SBC_CASE(2, 8) SBC_CASE(1, 8)
    SBC_CASE(4, 4) SBC_CASE(2, 4) SBC_CASE(1, 4)
    SBC_CASE(4, 2) SBC_CASE(2, 2) SBC_CASE(1, 2)
    SBC_CASE(4, 1) SBC_CASE(2, 1) SBC_CASE(1, 1)
#undef SBC_CASE
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

#define SBC_MMA_SLOTS 8

// Accumulate K steps [s0, s1) for up to 8 slots (slot j = tile mt0 + (j / nq) * mt_stride, pooling
// position j % nq) sharing one cout tile nt.  B fragments come from the staged parameter segment.
// Per slot the gather keeps two base offsets and a 12-bit validity word (3 row bits + 3 column bits
// for each of the two tile rows a lane feeds), so a K step costs a few integer ops per MMA.
template <bool X3>
__device__ __forceinline__ void sbc_mma_pass(const SbcOp& op, const SbcMmaGeom& G, const float* arena,
                                             const float* wseg, int mt0, int mt_stride, int nslots, int nt, int s0,
                                             int s1, int lane, float (&acc)[SBC_MMA_SLOTS][4]) {
    const int g = lane >> 2, t = lane & 3;
    const int h = op.h, w = op.w, k = op.ksize, r = k >> 1, dil = op.dil, cin = op.cin;
    const int ps = SBC_PS(h, w);
    int off0[SBC_MMA_SLOTS], off1[SBC_MMA_SLOTS];
    unsigned vm[SBC_MMA_SLOTS];
#pragma unroll
    for (int j = 0; j < SBC_MMA_SLOTS; j++) {
        off0[j] = off1[j] = 0;
        vm[j] = 0;
        if (j < nslots) {
            const int mt = mt0 + (j / G.nq) * mt_stride, quad = j % G.nq;
            int iy0, ix0, iy1, ix1;
            bool a, b;
            sbc_mma_row(op, G, mt, quad, g, iy0, ix0, a);
            sbc_mma_row(op, G, mt, quad, g + 8, iy1, ix1, b);
            off0[j] = iy0 * w + ix0;
            off1[j] = iy1 * w + ix1;
            unsigned m = 0;
#pragma unroll
            for (int q = 0; q < 3; q++) {
                const int d = (q - r) * dil;
                m |= (unsigned)(a && iy0 + d >= 0 && iy0 + d < h) << q;
                m |= (unsigned)(a && ix0 + d >= 0 && ix0 + d < w) << (3 + q);
                m |= (unsigned)(b && iy1 + d >= 0 && iy1 + d < h) << (6 + q);
                m |= (unsigned)(b && ix1 + d >= 0 && ix1 + d < w) << (9 + q);
            }
            vm[j] = m;
        }
    }
    constexpr int E = X3 ? 4 : 2;
    const float* src = arena + op.src;
    int s = 0;
    for (int tap = 0; tap < k * k; tap++) {
        if (!((op.tapmask >> tap) & 1)) continue;
        const int ky = tap / k, kx = tap - ky * k;
        const int doff = (ky - r) * dil * w + (kx - r) * dil;
        for (int kc = 0; kc < G.KC; kc++, s++) {
            if (s < s0 || s >= s1) continue;
            const float* bp = wseg + ((size_t)(s * G.NT + nt) * 32 + lane) * E;
            float bh0, bh1, bl0 = 0.f, bl1 = 0.f;
            if (X3) {
                const float4 b = *reinterpret_cast<const float4*>(bp);
                bh0 = b.x; bh1 = b.y; bl0 = b.z; bl1 = b.w;
            } else {
                const float2 b = *reinterpret_cast<const float2*>(bp);
                bh0 = b.x; bh1 = b.y;
            }
            const int c0 = kc * 8 + t;
            const bool k0 = c0 < cin, k1 = c0 + 4 < cin;
            const float* p0 = src + c0 * ps + doff;
            const float* p1 = p0 + 4 * ps;
#pragma unroll
            for (int j = 0; j < SBC_MMA_SLOTS; j++) {
                if (j < nslots) {
                    const unsigned m = vm[j];
                    const bool v0 = ((m >> ky) & (m >> (3 + kx)) & 1u) != 0;
                    const bool v1 = ((m >> (6 + ky)) & (m >> (9 + kx)) & 1u) != 0;
                    float a[4];
                    a[0] = (v0 && k0) ? p0[off0[j]] : 0.f;
                    a[1] = (v1 && k0) ? p0[off1[j]] : 0.f;
                    a[2] = (v0 && k1) ? p1[off0[j]] : 0.f;
                    a[3] = (v1 && k1) ? p1[off1[j]] : 0.f;
                    float ah[4];
#pragma unroll
                    for (int i = 0; i < 4; i++) ah[i] = sbc_tf32(a[i]);
                    if (X3) {
                        float al[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) al[i] = sbc_tf32(a[i] - ah[i]);
                        sbc_mma_tf32(acc[j], al, bh0, bh1);   // small terms first
                        sbc_mma_tf32(acc[j], ah, bl0, bl1);
                    }
                    sbc_mma_tf32(acc[j], ah, bh0, bh1);
                }
            }
        }
    }
}

template <bool X3>
__device__ __forceinline__ void sbc_conv_mma(const SbcOp& op, float* arena, const float* wseg, int tid) {
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = SBC_NTHREADS / 32;
    SbcMmaGeom G;
    sbc_mma_geom(op, G);
    float acc[SBC_MMA_SLOTS][4];

    if (op.ks > 1) {
        // fewer (tile, cout-tile) units than warps: `ks` warps split the K steps of one unit and the
        // partial accumulators are combined through shared memory
        const int ks = op.ks, units = G.MT * G.NT;
        const int u = warp / ks, kp = warp - u * ks;
        const int mt = u / G.NT, nt = u - mt * G.NT;
        float4* part = reinterpret_cast<float4*>(arena + op.scratch);
        if (u < units) {
#pragma unroll
            for (int j = 0; j < SBC_MMA_SLOTS; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            sbc_mma_pass<X3>(op, G, arena, wseg, mt, 0, G.nq, nt, (G.S * kp) / ks, (G.S * (kp + 1)) / ks, lane, acc);
            float4 c = make_float4(acc[0][0], acc[0][1], acc[0][2], acc[0][3]);
            if (G.nq == 4) {
#pragma unroll
                for (int j = 1; j < 4; j++) { c.x += acc[j][0]; c.y += acc[j][1]; c.z += acc[j][2]; c.w += acc[j][3]; }
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
            sbc_mma_epilogue(op, arena, wseg, mt, nt, lane, c);
        }
        return;
    }

    // enough units: each warp owns a fixed cout tile (so one B fragment feeds all its pixel tiles)
    int ntg = 1;
    while (ntg * 2 <= G.NT && ntg * 2 <= NW) ntg *= 2;
    const int mt_first = warp / ntg, mt_stride = NW / ntg;
    const int tpp = SBC_MMA_SLOTS / G.nq;                    // tiles per pass
    for (int nt = warp % ntg; nt < G.NT; nt += ntg) {
        for (int mt0 = mt_first; mt0 < G.MT; mt0 += tpp * mt_stride) {
            int ntile = (G.MT - mt0 + mt_stride - 1) / mt_stride;
            if (ntile > tpp) ntile = tpp;
#pragma unroll
            for (int j = 0; j < SBC_MMA_SLOTS; j++) acc[j][0] = acc[j][1] = acc[j][2] = acc[j][3] = 0.f;
            sbc_mma_pass<X3>(op, G, arena, wseg, mt0, mt_stride, ntile * G.nq, nt, 0, G.S, lane, acc);
            if (G.nq == 1) {
#pragma unroll
                for (int j = 0; j < SBC_MMA_SLOTS; j++)
                    if (j < ntile) sbc_mma_epilogue(op, arena, wseg, mt0 + j * mt_stride, nt, lane, acc[j]);
            } else {
#pragma unroll
                for (int j = 0; j < SBC_MMA_SLOTS / 4; j++)
                    if (j < ntile) {
                        float c[4];
#pragma unroll
                        for (int i = 0; i < 4; i++) c[i] = acc[4 * j][i] + acc[4 * j + 1][i] + acc[4 * j + 2][i] + acc[4 * j + 3][i];
                        sbc_mma_epilogue(op, arena, wseg, mt0 + j * mt_stride, nt, lane, c);
                    }
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm++ statistics: S lanes per channel, two passes, shuffle reductions (no barrier)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sbc_norm_stats(const SbcOp& op, float* arena, int tid) {
    const int S = sbc_norm_S(op, SBC_NTHREADS), C = op.cin;
    const float inv = 1.f / (float)(op.h * op.w);
    for (int base = 0; base < C * S; base += SBC_NTHREADS) {   // uniform trip count (shuffles are warp-wide)
        const int t = base + tid;
        const bool valid = t < C * S;
        const int c = valid ? t / S : 0, s = t - (t / S) * S;
        float sum = valid ? sbc_norm_partial_sum(op, arena, c, s, S) : 0.f;
        for (int off = 1; off < S; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
        const float mean = sum * inv;
        float m2 = valid ? sbc_norm_partial_m2(op, arena, c, s, S, mean) : 0.f;
        for (int off = 1; off < S; off <<= 1) m2 += __shfl_xor_sync(0xffffffffu, m2, off);
        if (valid && s == 0) sbc_norm_store_stats(op, arena, c, mean, m2);
    }
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
    float* wslot[2] = {smem_f + off, smem_f + off + L.max_w_len};
    if (L.stage_weights) off += 2 * (size_t)L.max_w_len;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_f + off);   // off is a multiple of 4 floats

    __shared__ float s_hnorm;
    __shared__ SbcStepScalars s_sc;

    const bool stage = L.stage_weights != 0;
    if (stage && tid == 0) {
        sbc_mbar_init(&bars[0], 1);
        sbc_mbar_init(&bars[1], 1);
        sbc_fence_barrier_init();
    }
    __syncthreads();

    const int Nt = L.Nt, Nr = L.Nr, ne = Nt * Nr, xps = SBC_PS(Nt, Nr);
    float* ax = arena + L.in_off;             // planar x: re plane, im plane (padded plane stride xps)
    uint32_t wcount = 0;                      // parameter segments consumed so far (same in every thread)

    if (stage && tid == 0 && (int)blockIdx.x < L.B && L.first_w >= 0) {
        const SbcOp& o = L.ops[L.first_w];
        sbc_mbar_expect_tx(&bars[0], (uint32_t)o.w_len * 4u);
        sbc_bulk_g2s(wslot[0], L.blob + o.w_off, (uint32_t)o.w_len * 4u, &bars[0]);
    }

    const int nsteps = (L.mode == 1) ? (L.level_end - L.level_begin) * L.steps_each : 1;

    for (int b = blockIdx.x; b < L.B; b += gridDim.x) {
        const bool last_sample = (b + (int)gridDim.x >= L.B);
        // ---------------- load the sample state into the arena (planar re/im) ----------------
        if (L.mode == 1) {
            const float* X = L.X + (size_t)b * ne * 2;
            for (int e = tid; e < ne; e += SBC_NTHREADS) {
                const float2 v = reinterpret_cast<const float2*>(X)[e];
                ax[e] = v.x;
                ax[xps + e] = v.y;
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
                    float s = 0.f;
                    for (int i = 0; i < SBC_NTHREADS; i++) s += red[i];
                    s_hnorm = s;
                }
            }
        } else {
            const float* fx = L.fx + (size_t)b * L.fxs[0];
            for (int i = tid; i < L.channels * ne; i += SBC_NTHREADS) {
                const int c = i / ne, e = i - c * ne;
                const int t = e / Nr, r = e - t * Nr;
                ax[c * xps + e] = fx[c * L.fxs[1] + t * L.fxs[2] + r * L.fxs[3]];
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
                    if (tid == 0) {   // prefetch the next parameter segment into the other slot
                        int j = op.pad0;                                        // next op with parameters
                        if (j < 0 && !(last_step && last_sample)) j = L.first_w;   // wraps into the next forward
                        if (j >= 0) {
                            const SbcOp& o = L.ops[j];
                            sbc_mbar_expect_tx(&bars[slot ^ 1u], (uint32_t)o.w_len * 4u);
                            sbc_bulk_g2s(wslot[slot ^ 1u], L.blob + o.w_off, (uint32_t)o.w_len * 4u, &bars[slot ^ 1u]);
                        }
                    }
                    sbc_mbar_wait(&bars[slot], (wcount >> 1) & 1u);
                    wseg = wslot[slot];
                    wcount++;
                }
                switch (op.kind) {
                    case SBC_OP_CONV:
                        sbc_conv_dispatch(op, arena, wseg, tid);
                        break;
                    case SBC_OP_CONV_MMA:
                        if (op.flags & SBC_F_X3) sbc_conv_mma<true>(op, arena, wseg, tid);
                        else sbc_conv_mma<false>(op, arena, wseg, tid);
                        break;
                    case SBC_OP_NORM_ELU:
                        sbc_norm_stats(op, arena, tid);
                        __syncthreads();
                        sbc_norm_apply(op, arena, wseg, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_ELU:
                        sbc_elu_op(op, arena, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_AFFINE:
                        sbc_affine_op(op, arena, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_MAXPOOL5:
                        sbc_maxpool5_op(op, arena, tid, SBC_NTHREADS);
                        break;
                    case SBC_OP_UPACC:
                        sbc_upacc_op(op, arena, tid, SBC_NTHREADS);
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
                for (int i = tid; i < L.channels * ne; i += SBC_NTHREADS) {
                    const int c = i / ne;
                    o[i] = net[c * xps + (i - c * ne)] / sg;
                }
            } else {
                // ---------------- data-consistency gradient, Langevin update, NMSE ----------------
                const float* Pm = L.P + (size_t)b * L.Np * Nt * 2;
                const float* Ym = L.Y + (size_t)b * L.Np * Nr * 2;
                const float* Hc = L.Hor ? L.Hor + (size_t)b * ne * 2 : nullptr;
                const float* en = L.ext_noise ? L.ext_noise + ((size_t)gs * L.B + b) * ne * 2 : nullptr;
                float* res = arena + L.post_off;
                float* red = res + 2 * ne;
                sbc_dc_residual(ax, res, Pm, Ym, Nt, Nr, L.Np, tid, SBC_NTHREADS);
                __syncthreads();
                const unsigned long long sid = L.sample_ids ? L.sample_ids[b] : (unsigned long long)b;
                const uint32_t gstep = (uint32_t)(lvl * L.steps_each + gs % L.steps_each);
                const float part = sbc_langevin_update(ax, net, res, Pm, Hc, en, s_sc, L.seed, sid, gstep, Nt, Nr,
                                                       L.Np, tid, SBC_NTHREADS);
                if (L.nmse_log && Hc) {
                    red[tid] = part;
                    __syncthreads();
                    if (tid == 0) {
                        float s = 0.f;
                        for (int i = 0; i < SBC_NTHREADS; i++) s += red[i];
                        L.nmse_log[(size_t)gs * L.B + b] = s / s_hnorm;
                    }
                }
            }
            __syncthreads();
            if (do_prof) L.prof[L.n_ops + 1] = clock64();
        }

        if (L.mode == 1) {   // write the final estimate back (interleaved complex64)
            float* X = L.X + (size_t)b * ne * 2;
            for (int e = tid; e < ne; e += SBC_NTHREADS)
                reinterpret_cast<float2*>(X)[e] = make_float2(ax[e], ax[xps + e]);
        }
        __syncthreads();
    }
}
