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

#include "sbc_ops.h"

#define SBC_NTHREADS 256

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

    const int Nt = L.Nt, Nr = L.Nr, ne = Nt * Nr;
    float* ax = arena + L.in_off;             // planar x: re plane, im plane
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
                ax[ne + e] = v.y;
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
                ax[i] = fx[c * L.fxs[1] + t * L.fxs[2] + r * L.fxs[3]];
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
                    case SBC_OP_NORM_ELU:
                        sbc_norm_phaseA(op, arena, tid, SBC_NTHREADS);
                        __syncthreads();
                        sbc_norm_phaseB(op, arena, tid, SBC_NTHREADS);
                        __syncthreads();
                        sbc_norm_phaseC(op, arena, wseg, tid, SBC_NTHREADS);
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
                for (int i = tid; i < L.channels * ne; i += SBC_NTHREADS) o[i] = net[i] / sg;
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
                reinterpret_cast<float2*>(X)[e] = make_float2(ax[e], ax[ne + e]);
        }
        __syncthreads();
    }
}
