// sbc_kernel.cuh -- persistent fused kernel: one CTA owns one channel realisation at a time and
// runs, entirely out of shared memory, the whole NCSNv2Deepest forward (152-op layer program,
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

// 256 threads and (at least) two CTAs per SM: the kernel is a chain of short latency-bound ops -- measured on B200, one
// sample runs only 6 % slower on 8 warps than on 16 -- so two channel realisations in flight per SM, each on its own 8
// warps, nearly double the throughput of one on 16.  128 registers per thread either way.
#ifndef SBC_NTHREADS
#define SBC_NTHREADS 256
#endif
// pixel tiles one warp accumulates per pass (register budget: 128 regs/thread at 512 threads, 64 at 1024)
#ifndef SBC_MINCTAS
#define SBC_MINCTAS 2
#endif
#ifndef SBC_MAXNS
#define SBC_MAXNS (SBC_NTHREADS * SBC_MINCTAS > 512 ? 2 : 4)
#endif

struct SbcLaunch {
    // layer program
    const SbcOp* ops;
    int n_ops;
    int first_w;             // index of the first op with parameters
    const float* blob;       // packed parameters
    SbcGeo geo[SBC_MAX_GEO]; // tensor geometries (geo[0] = network input / output resolution)
    int n_geo;
    int halo_off[SBC_MAX_GEO];   // start (uint16 units) of geometry g's halo-pixel list in the shared misc region
    int arena_floats, in_off, out_off, post_off;
    int Nt, Nr, channels, max_w_len;
    int mode;                // 0 = forward (NCSNv2Deepest.forward), 1 = annealed Langevin, 2 = denoising-score-matching loss
    int B;
    // forward mode
    const float* fx;         // [B,2,Nt,Nr] with element strides fxs
    long long fxs[4];
    const long long* labels; // [B]
    float* fout;             // [B,2,Nt,Nr] contiguous
    const float* sigmas;     // [n_sigmas]
    int n_sigmas;
    // DSM-loss mode (ncsnv2/losses/dsm.py:6-32): fx = clean samples, dsm_z = the randn draw [B,2,Nt,Nr] contiguous,
    // labels as in forward mode; dsm_out[b] = 1/2 * sum((score - target)^2) * sigma^anneal_power
    const float* dsm_z;
    float* dsm_out;
    float anneal_power;
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
    const float* dc_boost;   // [B] or null (= 1)
    const int* stop_step;    // [B] or null: last step index executed by a sample
    // execution
    float* gws;              // global arena workspace (when the arena does not fit in shared memory)
    float* gpark;            // park area: park_floats per CTA (SBC_OP_SPILL / SBC_OP_FILL / SBC_F_ACC_G), or null
    int park_floats;
    int stage_weights;       // 1: cp.async.bulk double buffering, 0: read parameters from global/L2
    int debug_stop;          // >=0: stop sample 0 / step 0 before op `debug_stop`, dump the arena
    float* debug_out;
    int dbg;                 // timing experiments only (env SBC_DBG): 1 skip K loops, 2 skip conv epilogues,
                             // 4 skip non-conv op bodies, 16 skip every op body (results are garbage)
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
//               -> fp32-equivalent accuracy (the parity mode);  X3 = false: TF32 operands (cvt.rna).
// The K loop is the hot loop of the whole sampler.  Per K step and 16-pixel tile it issues one ldmatrix.x4 (the
// whole A fragment, shared-memory arena), the operand split (8 ALU ops in X3 mode, 4 cvt otherwise) and the
// MMAs -- the tensor pipe (mma.sync TF32: one m16n8k8 per 8 cycles per SM sub-partition) is the binding unit.
// The tensor core ignores the low 13 mantissa bits of a TF32 operand, so a_hi / b_hi are passed unmasked.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void sbc_mma_tf32(float (&d)[4], const float (&a)[4], float b0, float b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(__float_as_uint(a[0])), "r"(__float_as_uint(a[1])), "r"(__float_as_uint(a[2])),
          "r"(__float_as_uint(a[3])), "r"(__float_as_uint(b0)), "r"(__float_as_uint(b1)));
}
__device__ __forceinline__ float sbc_cvt_tf32(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void sbc_ldmatrix_x4(float (&a)[4], uint32_t saddr) {
    uint32_t r0, r1, r2, r3;
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
                 : "r"(saddr));
    a[0] = __uint_as_float(r0); a[1] = __uint_as_float(r1); a[2] = __uint_as_float(r2); a[3] = __uint_as_float(r3);
}

// Per-lane A addressing of one warp for one conv op.
//   SMEM: sbase = shared address of src plane lane>>4 (of the chunk's two planes); po[j][0] = byte offset of
//         pixel row lane & 15 of slot j (ldmatrix row address)
//   else: gsrc = arena + op.src + t; po[j][0/1] = float offsets of pixel rows g / g + 8
template <bool SMEM>
struct SbcALane {
    uint32_t sbase;
    const float* gsrc;
    int pl4;   // floats per plane of the source geometry
    __device__ __forceinline__ SbcALane(const SbcOp& op, const SbcGeo& GS, float* arena, int lane) {
        pl4 = GS.pps * 4;
        if (SMEM) { sbase = sbc_smem_u32(arena + op.src) + (uint32_t)(lane >> 4) * (uint32_t)(pl4 * 4); gsrc = nullptr; }
        else { sbase = 0; gsrc = arena + op.src + (lane & 3); }
    }
    __device__ __forceinline__ void rows(const SbcOp& op, const SbcGeo& GS, int mt, int quad, int lane, int (&po)[2]) const {
        if (SMEM) {
            po[0] = sbc_mma_row_off(op, GS, mt, quad, lane & 15) * 4;
            po[1] = 0;
        } else {
            po[0] = sbc_mma_row_off(op, GS, mt, quad, lane >> 2);
            po[1] = sbc_mma_row_off(op, GS, mt, quad, (lane >> 2) + 8);
        }
    }
    __device__ __forceinline__ void frag(int off, const int (&po)[2], float (&a)[4]) const {
        if (SMEM) sbc_ldmatrix_x4(a, sbase + (uint32_t)(off * 4 + po[0]));
        else sbc_mma_a_frag(gsrc, off, po[0], po[1], pl4, a);
    }
};

// Accumulate K steps [s0, s1) for NS pixel-tile slots (row offsets po[j]) and NN cout tiles that share each
// gathered A fragment.  bfrag = first B fragment of this lane for cout tile nt0; bstride = floats per K step in
// the fragment array.  A dependent mma.sync chain advances one link per ~HMMA latency, far slower than the pipe
// rate, so when a warp owns fewer than 4 (tile, cout tile) accumulators the three 3xTF32 terms (resp. even / odd
// K steps in TF32 mode) go to separate accumulator copies that are summed at the end.
// Operands of one K step as they come out of memory: the raw A fragments of the NS tiles and the raw B fragments of
// the NN cout tiles (split into hi / lo right before the MMAs that consume them).
template <int NS, int NN>
struct SbcKStep {
    float a[NS][4];
    float2 b[NN];
};
template <bool SMEM, int NS, int NN>
__device__ __forceinline__ void sbc_kstep_load(const SbcALane<SMEM>& A, const int (&po)[NS][2], const int* tp,
                                               const float* bp, SbcKStep<NS, NN>& k) {
    const int off = *tp;
#pragma unroll
    for (int n = 0; n < NN; n++) k.b[n] = *reinterpret_cast<const float2*>(bp + n * 64);
#pragma unroll
    for (int j = 0; j < NS; j++) A.frag(off, po[j], k.a[j]);
}
// the MMAs of one K step.  ODD selects the accumulator copy in the SPLIT TF32 mode (see sbc_mma_pass).
template <bool X3, bool SPLIT, bool ODD, int NS, int NN, int NCX>
__device__ __forceinline__ void sbc_kstep_mma(SbcKStep<NS, NN>& k, float (&acc)[NS][NN][4], float (&accx)[NCX][NS][NN][4]) {
    float bh[NN][2], bl[NN][2];
#pragma unroll
    for (int n = 0; n < NN; n++) {
        bh[n][0] = k.b[n].x; bh[n][1] = k.b[n].y;   // X3: plain fp32 (hi part = what the tensor core reads of it)
        if (X3) {                                    // w = hi + lo exactly; the tensor core truncates lo to 11 bits
            bl[n][0] = k.b[n].x - __uint_as_float(__float_as_uint(k.b[n].x) & 0xFFFFE000u);
            bl[n][1] = k.b[n].y - __uint_as_float(__float_as_uint(k.b[n].y) & 0xFFFFE000u);
        } else {
            bl[n][0] = bl[n][1] = 0.f;
        }
    }
    float al[NS][4];
#pragma unroll
    for (int j = 0; j < NS; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (X3) al[j][i] = k.a[j][i] - __uint_as_float(__float_as_uint(k.a[j][i]) & 0xFFFFE000u);
            else k.a[j][i] = sbc_cvt_tf32(k.a[j][i]);
        }
    if (X3) {   // small terms first; term-major order keeps dependent MMAs NS*NN apart
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++) sbc_mma_tf32(SPLIT ? accx[0][j][n] : acc[j][n], al[j], bh[n][0], bh[n][1]);
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++) sbc_mma_tf32(SPLIT ? accx[NCX - 1][j][n] : acc[j][n], k.a[j], bl[n][0], bl[n][1]);
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++) sbc_mma_tf32(acc[j][n], k.a[j], bh[n][0], bh[n][1]);
    } else {
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++) sbc_mma_tf32((SPLIT && ODD) ? accx[0][j][n] : acc[j][n], k.a[j], bh[n][0], bh[n][1]);
    }
}

// Accumulate K steps [s0, s1) for NS pixel-tile slots (row offsets po[j]) and NN cout tiles that share each
// gathered A fragment.  bfrag = first B fragment of this lane for cout tile nt0; bstride = floats per K step in
// the fragment array.  A dependent mma.sync chain advances one link per ~HMMA latency, far slower than the pipe
// rate, so when a warp owns fewer than 4 (tile, cout tile) accumulators the three 3xTF32 terms (resp. even / odd
// K steps in TF32 mode) go to separate accumulator copies that are summed at the end.
// The loop is software-pipelined by hand, two K steps per trip with ping-pong operand registers: the loads (step
// offset, B fragments, ldmatrix) of step s+1 are issued BEFORE the MMAs of step s, so that their latency is covered by
// tensor-pipe time instead of by the few other warps of a 25 %-occupancy kernel.
template <bool X3, bool SMEM, int NS, int NN>
__device__ __forceinline__ void sbc_mma_pass(const SbcALane<SMEM>& A, const int (&po)[NS][2],
                                             const int* __restrict__ steptab, const float* __restrict__ bfrag,
                                             int bstride, int s0, int s1, float (&acc)[NS][NN][4]) {
    constexpr bool SPLIT = NS * NN < 4;
    constexpr int NC = SPLIT ? (X3 ? 3 : 2) : 1;
    constexpr int NCX = NC > 1 ? NC - 1 : 1;
    float accx[NCX][NS][NN][4];
    if (SPLIT) {
#pragma unroll
        for (int c = 0; c < NCX; c++)
#pragma unroll
            for (int j = 0; j < NS; j++)
#pragma unroll
                for (int n = 0; n < NN; n++) accx[c][j][n][0] = accx[c][j][n][1] = accx[c][j][n][2] = accx[c][j][n][3] = 0.f;
    }
    if (s0 >= s1) return;
    const int* tp = steptab + s0;                 // running pointers: no per-step 64-bit address arithmetic
    const float* bp = bfrag + (size_t)s0 * bstride;
    SbcKStep<NS, NN> k0, k1;
    sbc_kstep_load<SMEM, NS, NN>(A, po, tp, bp, k0);
    int left = s1 - s0;                           // steps not yet multiplied; k0 holds the first of them
#pragma unroll 1
    while (left >= 2) {
        sbc_kstep_load<SMEM, NS, NN>(A, po, tp + 1, bp + bstride, k1);
        sbc_kstep_mma<X3, SPLIT, false, NS, NN, NCX>(k0, acc, accx);
        tp += 2; bp += 2 * bstride; left -= 2;
        if (left > 0) sbc_kstep_load<SMEM, NS, NN>(A, po, tp, bp, k0);
        sbc_kstep_mma<X3, SPLIT, true, NS, NN, NCX>(k1, acc, accx);
    }
    if (left > 0) sbc_kstep_mma<X3, SPLIT, false, NS, NN, NCX>(k0, acc, accx);
    if (SPLIT) {
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    float small = accx[0][j][n][i];
                    if (NC > 2) small += accx[NC - 2][j][n][i];
                    acc[j][n][i] += small;
                }
    }
}

// Fused epilogue of one lane for NS tiles x NN cout tiles, phase-batched: all bias / accumulator loads are
// issued before any is consumed and every (uniform) branch is taken once per pass instead of once per tile, so
// the load and MUFU latencies of the tiles overlap.  Semantics = sbc_mma_epilogue (sbc_mma.h) per tile.
template <int NS, int NN>
__device__ __forceinline__ void sbc_epilogue_batch(const SbcEpi& e, float* arena, const float* wseg,
                                                   const int (&pd)[NS][2], const int (&q0)[NS], int nt0, int lane,
                                                   float (&acc)[NS][NN][4]) {
    const int t = lane & 3;
    int cofs[NN];
    bool live[NN];
#pragma unroll
    for (int n = 0; n < NN; n++) {
        const int co = (nt0 + n) * 8 + 2 * t;
        live[n] = co < e.cout;
        cofs[n] = (co >> 2) * e.pps4 + (co & 3);
        if (e.b_rel >= 0 && live[n]) {
            const float2 b = *reinterpret_cast<const float2*>(wseg + e.b_rel + co);   // co is even, b_rel % 4 == 0
#pragma unroll
            for (int j = 0; j < NS; j++) { acc[j][n][0] += b.x; acc[j][n][1] += b.y; acc[j][n][2] += b.x; acc[j][n][3] += b.y; }
        }
    }
    if (e.flags & SBC_F_COMPACT) {   // network output: couts (0,1) = (re, im) of element q
        if (live[0] && t == 0) {
#pragma unroll
            for (int j = 0; j < NS; j++)
#pragma unroll
                for (int half = 0; half < 2; half++)
                    if (pd[j][half] >= 0)
                        reinterpret_cast<float2*>(arena + e.dst)[q0[j] + 8 * half] = make_float2(acc[j][0][2 * half], acc[j][0][2 * half + 1]);
        }
        return;
    }
    if (e.dst >= 0) {
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int half = 0; half < 2; half++)
                    if (live[n] && pd[j][half] >= 0)
                        *reinterpret_cast<float2*>(arena + e.dst + pd[j][half] + cofs[n]) = make_float2(acc[j][n][2 * half], acc[j][n][2 * half + 1]);
    }
    if (e.acc >= 0) {
        float2 old[NS][NN][2];
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int half = 0; half < 2; half++)
                    old[j][n][half] = (live[n] && pd[j][half] >= 0)
                                          ? *reinterpret_cast<const float2*>(e.accb + e.acc + pd[j][half] + cofs[n])
                                          : make_float2(0.f, 0.f);
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int half = 0; half < 2; half++) {
                    acc[j][n][2 * half] += old[j][n][half].x;
                    acc[j][n][2 * half + 1] += old[j][n][half].y;
                    if (live[n] && pd[j][half] >= 0)
                        *reinterpret_cast<float2*>(e.accb + e.acc + pd[j][half] + cofs[n]) = make_float2(acc[j][n][2 * half], acc[j][n][2 * half + 1]);
                }
    }
    if (e.edst >= 0) {
#pragma unroll
        for (int j = 0; j < NS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++)
#pragma unroll
                for (int half = 0; half < 2; half++)
                    if (live[n] && pd[j][half] >= 0)
                        *reinterpret_cast<float2*>(arena + e.edst + pd[j][half] + cofs[n]) =
                            make_float2(sbc_elu(acc[j][n][2 * half]), sbc_elu(acc[j][n][2 * half + 1]));
    }
}

// one pass of a warp over NS tiles (mt0, mt0 + mstride, ...; only the first `ntile` are real) x NN cout tiles
template <bool X3, bool SMEM, int NS, int NN>
__device__ __forceinline__ void sbc_conv_tiles(const SbcOp& op, const SbcGeo& GS, const SbcGeo& GD, float* arena,
                                               float* park, const float* wseg, const SbcALane<SMEM>& A, const float* bfrag,
                                               int bstride, int mt0, int mstride, int ntile, int nt0, int lane,
                                               long long* stamp) {
    int po[NS][2];
    A.rows(op, GS, mt0, 0, lane, po[0]);
    if (NS > 1) {
        // tiles mt0 + j*mstride: when the tile stride is a whole number of image rows and no row is clamped, the
        // pixel offsets advance by a constant
        const int pstep = mstride * 16;
        if (op.low >= 0 && (pstep & (op.ow - 1)) == 0 && (mt0 + (NS - 1) * mstride) * 16 + 16 <= op.oh * op.ow) {
            const int step = (pstep >> op.low) * GS.wp * (SMEM ? 16 : 4);
#pragma unroll
            for (int j = 1; j < NS; j++) { po[j][0] = po[0][0] + j * step; po[j][1] = po[0][1] + j * step; }
        } else {
#pragma unroll
            for (int j = 1; j < NS; j++) A.rows(op, GS, mt0 + (j < ntile ? j : 0) * mstride, 0, lane, po[j]);
        }
    }
    float acc[NS][NN][4];
#pragma unroll
    for (int j = 0; j < NS; j++)
#pragma unroll
        for (int n = 0; n < NN; n++) acc[j][n][0] = acc[j][n][1] = acc[j][n][2] = acc[j][n][3] = 0.f;
    if (stamp) stamp[1] = clock64();
    sbc_mma_pass<X3, SMEM, NS, NN>(A, po, reinterpret_cast<const int*>(wseg), bfrag, bstride, 0, op.S, acc);
    if (stamp) stamp[2] = clock64();
    const SbcEpi e = sbc_epi(op, GD, arena, park);
    int pd[NS][2], q0[NS];
#pragma unroll
    for (int j = 0; j < NS; j++) {
        const int mt = mt0 + j * mstride;
        q0[j] = mt * 16 + (lane >> 2);
        if (j < ntile) sbc_mma_dst_off(op, GD, mt, lane >> 2, pd[j]);
        else pd[j][0] = pd[j][1] = -1;
    }
    sbc_epilogue_batch<NS, NN>(e, arena, wseg, pd, q0, nt0, lane, acc);
}

// ConvMeanPool: one output tile, the four pooling positions are slots (SBC_MAXNS at a time)
template <bool X3, bool SMEM, int NN>
__device__ __forceinline__ void sbc_conv_pooled(const SbcOp& op, const SbcGeo& GS, const float* wseg,
                                                const SbcALane<SMEM>& A, const float* bfrag, int bstride, int mt,
                                                int s0, int s1, int lane, float (&c)[NN][4]) {
    constexpr int QS = SBC_MAXNS < 4 ? SBC_MAXNS : 4;
#pragma unroll
    for (int n = 0; n < NN; n++) c[n][0] = c[n][1] = c[n][2] = c[n][3] = 0.f;
#pragma unroll 1
    for (int q0 = 0; q0 < 4; q0 += QS) {
        int po[QS][2];
#pragma unroll
        for (int j = 0; j < QS; j++) A.rows(op, GS, mt, q0 + j, lane, po[j]);
        float acc[QS][NN][4];
#pragma unroll
        for (int j = 0; j < QS; j++)
#pragma unroll
            for (int n = 0; n < NN; n++) acc[j][n][0] = acc[j][n][1] = acc[j][n][2] = acc[j][n][3] = 0.f;
        sbc_mma_pass<X3, SMEM, QS, NN>(A, po, reinterpret_cast<const int*>(wseg), bfrag, bstride, s0, s1, acc);
#pragma unroll
        for (int n = 0; n < NN; n++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                float t = acc[0][n][i];
#pragma unroll
                for (int j = 1; j < QS; j++) t += acc[j][n][i];
                c[n][i] += t;
            }
    }
}

template <bool X3, bool SMEM>
__device__ __forceinline__ void sbc_conv_mma(const SbcOp& op_, const SbcGeo& GS, const SbcGeo& GD, float* arena,
                                             float* park, const float* wseg, int tid, long long* stamp, int dbg) {
    SbcOp op = op_;
    if (dbg & 1) op.S = 0;                                      // timing experiment: no K steps
    if (dbg & 2) { op.dst = op.acc = op.edst = -1; op.flags &= ~SBC_F_COMPACT; }   // timing experiment: no stores
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int NW = SBC_NTHREADS / 32;
    constexpr int E = 2;   // floats per lane per B fragment
    const bool pool = (op.flags & SBC_F_POOL) != 0;

    const int MT = op.MT, NT = op.NT, S = op.S;
    const SbcALane<SMEM> A(op, GS, arena, lane);
    const int bstride = NT * 32 * E;
    const float* bf0 = wseg + op.frag_rel + lane * E;
    if (stamp) stamp[0] = clock64();

    if (op.flags & SBC_F_UNIT) {
        // fewer (pixel tile, cout tile) units than warps: `ks` (a power of two) warps split the K steps of one
        // unit and the partial accumulators are combined through shared memory (ks = 1: no split, no exchange)
        const int ks = op.ks, units = MT * NT;
        const int lks = 31 - __clz(ks);
        const int u = warp >> lks, kp = warp & (ks - 1);
        const int mt = sbc_div(u, NT, sbc_ilog2(NT)), nt = u - mt * NT;
        float4* part = reinterpret_cast<float4*>(arena + op.scratch);
        if (u < units) {
            const int s0 = (S * kp) >> lks, s1 = (S * (kp + 1)) >> lks;
            float c[1][4];
            if (pool) {
                sbc_conv_pooled<X3, SMEM, 1>(op, GS, wseg, A, bf0 + nt * 32 * E, bstride, mt, s0, s1, lane, c);
            } else {
                int po[1][2];
                A.rows(op, GS, mt, 0, lane, po[0]);
                float acc[1][1][4] = {{{0.f, 0.f, 0.f, 0.f}}};
                sbc_mma_pass<X3, SMEM, 1, 1>(A, po, reinterpret_cast<const int*>(wseg), bf0 + nt * 32 * E, bstride, s0, s1, acc);
                c[0][0] = acc[0][0][0]; c[0][1] = acc[0][0][1]; c[0][2] = acc[0][0][2]; c[0][3] = acc[0][0][3];
            }
            if (ks == 1) {
                int pd[2];
                sbc_mma_dst_off(op, GD, mt, lane >> 2, pd);
                sbc_mma_epilogue(sbc_epi(op, GD, arena, park), arena, wseg, pd[0], pd[1], mt * 16 + (lane >> 2), nt, lane, c[0][0], c[0][1], c[0][2], c[0][3]);
            } else {
                part[warp * 32 + lane] = make_float4(c[0][0], c[0][1], c[0][2], c[0][3]);
            }
        }
        if (ks == 1) return;
        __syncthreads();
        if (u < units && kp == 0) {
            float c[4] = {0.f, 0.f, 0.f, 0.f};
            for (int i = 0; i < ks; i++) {
                const float4 p = part[(warp + i) * 32 + lane];
                c[0] += p.x; c[1] += p.y; c[2] += p.z; c[3] += p.w;
            }
            int pd[2];
            sbc_mma_dst_off(op, GD, mt, lane >> 2, pd);
            sbc_mma_epilogue(sbc_epi(op, GD, arena, park), arena, wseg, pd[0], pd[1], mt * 16 + (lane >> 2), nt, lane, c[0], c[1], c[2], c[3]);
        }
        return;
    }

    if (pool) {
        const SbcEpi e = sbc_epi(op, GD, arena, park);
        for (int mt = warp; mt < MT; mt += NW) {
            int pd[2];
            sbc_mma_dst_off(op, GD, mt, lane >> 2, pd);
            const int q0 = mt * 16 + (lane >> 2);
            for (int nt0 = 0; nt0 < NT; nt0 += 2) {
                if (NT - nt0 >= 2) {
                    float c[2][4];
                    sbc_conv_pooled<X3, SMEM, 2>(op, GS, wseg, A, bf0 + nt0 * 32 * E, bstride, mt, 0, S, lane, c);
                    sbc_mma_epilogue(e, arena, wseg, pd[0], pd[1], q0, nt0, lane, c[0][0], c[0][1], c[0][2], c[0][3]);
                    sbc_mma_epilogue(e, arena, wseg, pd[0], pd[1], q0, nt0 + 1, lane, c[1][0], c[1][1], c[1][2], c[1][3]);
                } else {
                    float c[1][4];
                    sbc_conv_pooled<X3, SMEM, 1>(op, GS, wseg, A, bf0 + nt0 * 32 * E, bstride, mt, 0, S, lane, c);
                    sbc_mma_epilogue(e, arena, wseg, pd[0], pd[1], q0, nt0, lane, c[0][0], c[0][1], c[0][2], c[0][3]);
                }
            }
        }
        return;
    }

    // enough pixel tiles for every warp: warp w owns tiles w, w+NW, ... (up to SBC_MAXNS per pass); cout tiles
    // are processed in pairs that share the gathered A fragments
    for (int mt0 = warp; mt0 < MT; mt0 += SBC_MAXNS * NW) {
        int ntile = (MT - mt0 + NW - 1) / NW;
        if (ntile > SBC_MAXNS) ntile = SBC_MAXNS;
        for (int nt0 = 0; nt0 < NT; nt0 += 2) {
            const float* bf = bf0 + nt0 * 32 * E;
            const bool two = NT - nt0 >= 2;
            if (ntile == 1) {
                if (two) sbc_conv_tiles<X3, SMEM, 1, 2>(op, GS, GD, arena, park, wseg, A, bf, bstride, mt0, NW, ntile, nt0, lane, stamp);
                else sbc_conv_tiles<X3, SMEM, 1, 1>(op, GS, GD, arena, park, wseg, A, bf, bstride, mt0, NW, ntile, nt0, lane, stamp);
            } else if (ntile == 2 || SBC_MAXNS == 2) {
                if (two) sbc_conv_tiles<X3, SMEM, 2, 2>(op, GS, GD, arena, park, wseg, A, bf, bstride, mt0, NW, ntile, nt0, lane, stamp);
                else sbc_conv_tiles<X3, SMEM, 2, 1>(op, GS, GD, arena, park, wseg, A, bf, bstride, mt0, NW, ntile, nt0, lane, stamp);
            } else {
                if (two) sbc_conv_tiles<X3, SMEM, SBC_MAXNS, 2>(op, GS, GD, arena, park, wseg, A, bf, bstride, mt0, NW, ntile, nt0, lane, stamp);
                else sbc_conv_tiles<X3, SMEM, SBC_MAXNS, 1>(op, GS, GD, arena, park, wseg, A, bf, bstride, mt0, NW, ntile, nt0, lane, stamp);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// InstanceNorm++ + ELU: T threads per channel quad, two statistics passes (warp shuffles + one
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
    const int C = op.cin, nq = C >> 2;
    const int T = op.MT, lT = op.NT, npass = op.S;      // host-derived (program.py:norm_elu)
    const int gpp = SBC_NTHREADS >> lT, wpg = T >> 5;   // quads / pass, warps / quad
    const int warp = tid >> 5, lane = tid & 31;
    SbcF4* red = reinterpret_cast<SbcF4*>(arena + op.scratch);
    SbcF4* mu4 = red + 2 * NW;
    SbcF4* m24 = mu4 + nq;
    const float inv = __int_as_float(op.frag_rel);      // 1 / (h*w)
    const int s = tid & (T - 1), w0 = (tid >> lT) * wpg;
    const SbcF4 z{0.f, 0.f, 0.f, 0.f};
    SbcF4 mean = z, m2 = z;
    if (wpg == 1 && npass == 1) {
        // one warp per quad (small maps): both statistics come straight out of the shuffle reductions; the only
        // exchange is the per-channel means every thread needs for the cross-channel term
        const int q = tid >> 5;
        const bool active = q < nq;
        SbcF4 sum = active ? sbc_norm_partial_sum(op, G, arena, q, lane, 32) : z;
        sum = sbc_warp_sum4(sum);
        mean.x = sum.x * inv; mean.y = sum.y * inv; mean.z = sum.z * inv; mean.w = sum.w * inv;
        m2 = sbc_warp_sum4(active ? sbc_norm_partial_m2(op, G, arena, q, lane, 32, mean) : z);
        if (active && lane == 0) mu4[q] = mean;
        __syncthreads();
        if (active) sbc_norm_apply(op, G, arena, wseg, reinterpret_cast<const float*>(mu4), q, lane, 32, mean, m2);
        return;
    }
    for (int pass = 0; pass < npass; pass++) {
        const int q = pass * gpp + (tid >> lT);
        const bool active = q < nq;
        SbcF4 sum = active ? sbc_norm_partial_sum(op, G, arena, q, s, T) : z;
        sum = sbc_warp_sum4(sum);
        if (lane == 0) red[warp] = sum;
        __syncthreads();
        mean = z;
        for (int i = 0; i < wpg; i++) { const SbcF4 v = red[w0 + i]; mean.x += v.x; mean.y += v.y; mean.z += v.z; mean.w += v.w; }
        mean.x *= inv; mean.y *= inv; mean.z *= inv; mean.w *= inv;
        if (active && s == 0) mu4[q] = mean;
        SbcF4 part = active ? sbc_norm_partial_m2(op, G, arena, q, s, T, mean) : z;
        part = sbc_warp_sum4(part);
        if (lane == 0) red[NW + warp] = part;
        __syncthreads();
        m2 = z;
        for (int i = 0; i < wpg; i++) { const SbcF4 v = red[NW + w0 + i]; m2.x += v.x; m2.y += v.y; m2.z += v.z; m2.w += v.w; }
        if (npass > 1) {   // several passes: statistics go through shared memory, `red` is reused
            if (active && s == 0) m24[q] = m2;
            __syncthreads();
        }
    }
    for (int pass = 0; pass < npass; pass++) {
        const int q = pass * gpp + (tid >> lT);
        if (q < nq) {
            if (npass > 1) { mean = mu4[q]; m2 = m24[q]; }
            sbc_norm_apply(op, G, arena, wseg, reinterpret_cast<const float*>(mu4), q, s, T, mean, m2);
        }
    }
}

// Halo zeroing from a per-geometry list of halo pixel indices kept in shared memory (built once per launch):
// one float4 store per item, no index arithmetic.  hl = list of geometry g, nh = its length.
__device__ __forceinline__ void sbc_zero_halo_list(float* t, const SbcGeo& G, const uint16_t* hl, int nh, int c,
                                                   int tid) {
    const int np = (c + 3) >> 2;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int pl = 0; pl < np; pl++) {
        float* base = t + (size_t)pl * G.pps * 4;
        for (int k = tid; k < nh; k += SBC_NTHREADS) *reinterpret_cast<float4*>(base + (int)hl[k] * 4) = z;
    }
}

// block-wide sum of one float per thread; the total is returned to thread 0 only.  red: NW floats.
__device__ __forceinline__ float sbc_block_sum(float v, float* red, int tid) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    if ((tid & 31) == 0) red[tid >> 5] = v;
    __syncthreads();
    float s = 0.f;
    if (tid == 0)
        for (int i = 0; i < SBC_NTHREADS / 32; i++) s += red[i];
    return s;
}

// ---------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------
#define SBC_MISC_BARS 64      // bytes reserved for the two mbarriers at the head of the shared misc region

// INSTR = true: the instrumented build of the same kernel (per-op clock stamps, debug arena dump, timing
// experiments); the production instantiation carries none of that code in its op loop.
template <bool SMEM_ARENA, bool X3, bool INSTR>
__global__ void __launch_bounds__(SBC_NTHREADS, SBC_MINCTAS) sbc_ald_kernel(const __grid_constant__ SbcLaunch L) {
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
    float* park = L.gpark ? L.gpark + (size_t)blockIdx.x * (size_t)L.park_floats : nullptr;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_f + off);   // off is a multiple of 4 floats
    uint16_t* halo = reinterpret_cast<uint16_t*>(reinterpret_cast<unsigned char*>(bars) + SBC_MISC_BARS);

    __shared__ float s_hnorm;
    __shared__ SbcStepScalars s_sc;
    __shared__ float s_red[SBC_NTHREADS / 32];
    __shared__ int s_hcnt[SBC_MAX_GEO];
    // op records are prefetched global -> shared TWO ops ahead (warp 0, one word per lane; the word loaded
    // during op i is parked in a register and stored at the start of op i+1), so that neither decoding an op
    // nor the barrier that ends a short op ever waits on global memory
    __shared__ __align__(16) SbcOp s_ops[4];

    // parameter segments are staged into the (shared-memory) arena with cp.async.bulk one op ahead
    const bool stage = SMEM_ARENA && L.stage_weights != 0;
    if (stage && tid == 0) {
        sbc_mbar_init(&bars[0], 1);
        sbc_mbar_init(&bars[1], 1);
        sbc_fence_barrier_init();
    }
    int pend = 0;                             // warp 0: word `tid` of the op record two ops ahead
    if (tid < 32) {
        const int* o32 = reinterpret_cast<const int*>(L.ops);
        reinterpret_cast<int*>(&s_ops[0])[tid] = o32[tid];
        reinterpret_cast<int*>(&s_ops[1])[tid] = o32[(L.n_ops > 1 ? 32 : 0) + tid];
        pend = o32[(2 % L.n_ops) * 32 + tid];
    }
    if (tid < SBC_MAX_GEO) s_hcnt[tid] = 0;
    __syncthreads();
    // halo-pixel lists, one per geometry (order is irrelevant): used by sbc_zero_halo_list
    for (int g = 0; g < L.n_geo; g++) {
        const SbcGeo& G = L.geo[g];
        for (int p = tid; p < G.pps; p += SBC_NTHREADS) {
            const int row = p / G.wp, col = p - row * G.wp;
            if (row < G.hy || row >= G.hy + G.h || col < G.hx || col >= G.hx + G.w)
                halo[L.halo_off[g] + atomicAdd(&s_hcnt[g], 1)] = (uint16_t)p;
        }
    }
    __syncthreads();

    const int Nt = L.Nt, Nr = L.Nr, ne = Nt * Nr;
    float* ax = arena + L.in_off;             // x: compact (re, im) pairs [Nt*Nr]
    uint32_t wcount = 0;                      // parameter segments consumed so far (same in every thread)
    uint32_t opc = 0;                         // ops executed so far (parity selects the op record buffer)

    if (stage && tid == 0 && (int)blockIdx.x < L.B && L.first_w >= 0) {
        const SbcOp& o = L.ops[L.first_w];
        sbc_mbar_expect_tx(&bars[0], (uint32_t)o.w_len * 4u);
        sbc_bulk_g2s(arena + o.wbuf, L.blob + o.w_off, (uint32_t)o.w_len * 4u, &bars[0]);
    }

    const int nsteps = (L.mode == 1) ? (L.level_end - L.level_begin) * L.steps_each : 1;
    const long long t_cta0 = INSTR ? clock64() : 0;

    for (int b = blockIdx.x; b < L.B; b += gridDim.x) {
        const bool last_sample = (b + (int)gridDim.x >= L.B);
        // ---------------- load the sample state into the arena ----------------
        if (L.mode == 1) {
            const float2* X = reinterpret_cast<const float2*>(L.X) + (size_t)b * ne;
            for (int e = tid; e < ne; e += SBC_NTHREADS) reinterpret_cast<float2*>(ax)[e] = X[e];
            if (L.Hor) {   // ||H||_F^2 once per sample (test_score.py:169)
                const float2* Hc = reinterpret_cast<const float2*>(L.Hor) + (size_t)b * ne;
                float part = 0.f;
                for (int e = tid; e < ne; e += SBC_NTHREADS) {
                    const float2 v = Hc[e];
                    part += v.x * v.x + v.y * v.y;
                }
                const float tot = sbc_block_sum(part, s_red, tid);
                if (tid == 0) s_hnorm = tot;
            }
        } else {
            const float* fx = L.fx + (size_t)b * L.fxs[0];
            float sg = 0.f;
            const float* z = nullptr;
            if (L.mode == 2) {   // perturbed_samples = samples + randn * sigma   (dsm.py:15-17)
                long long lab = L.labels[b];
                lab = lab < 0 ? 0 : (lab >= L.n_sigmas ? L.n_sigmas - 1 : lab);
                sg = L.sigmas[lab];
                z = L.dsm_z + (size_t)b * L.channels * ne;
            }
            for (int e = tid; e < ne; e += SBC_NTHREADS) {
                const int t = e / Nr, r = e - t * Nr;
                const float* q = fx + t * L.fxs[2] + r * L.fxs[3];
                float2 v = make_float2(q[0], q[L.fxs[1]]);
                if (z) { v.x += z[e] * sg; v.y += z[ne + e] * sg; }
                reinterpret_cast<float2*>(ax)[e] = v;
            }
        }
        __syncthreads();

        // early stop (test_mmse.py:260-263): the sample executes steps 0 .. stop_step[b] only
        int nsteps_b = nsteps;
        if (L.mode == 1 && L.stop_step) {
            const int st = L.stop_step[b] + 1;
            nsteps_b = st < 1 ? 1 : (st < nsteps ? st : nsteps);   // at least step 0, like the reference loop
        }
        for (int gs = 0; gs < nsteps_b; gs++) {
            const bool last_step = (gs + 1 == nsteps_b);
            int lvl = 0;
            if (L.mode == 1) {
                lvl = L.level_begin + gs / L.steps_each;
                if (tid == 0 && (gs % L.steps_each) == 0) {   // per-level scalars, in double like the reference
                    const double sigma = (double)L.sigmas[lvl];
                    const double ratio = sigma / L.sigma_end;
                    const double alpha = (double)L.alpha_step[b] * ratio * ratio;
                    s_sc.sigma = L.sigmas[lvl];
                    s_sc.alpha = (float)alpha;
                    // data-consistency weight 1 / (local_noise/2 + sigma^2), optionally boosted (test_mmse.py:246)
                    s_sc.den = (float)((double)L.noise_var[b] / 2. + sigma * sigma) / (L.dc_boost ? L.dc_boost[b] : 1.f);
                    s_sc.nscale = (float)sqrt(2. * alpha * (double)L.beta[b]);
                }
            }

            // ---------------- the network: walk the layer program ----------------
            // INSTR: stamps of CTA 0, first sample, second step when there is one (warm instruction cache)
            const bool do_prof = INSTR && (L.prof != nullptr) && blockIdx.x == 0 && b == 0 &&
                                 gs == (nsteps > 1 ? 1 : 0) && tid == 0;
            for (int i = 0; i < L.n_ops; i++, opc++) {
                if (INSTR && L.debug_stop >= 0 && i == L.debug_stop) break;
                if (INSTR && do_prof) L.prof[i] = clock64();
                const SbcOp op = s_ops[opc & 3u];
                if (tid < 32) {   // park op i+2 (loaded during the previous op), start loading op i+3 (wrapping)
                    reinterpret_cast<int*>(&s_ops[(opc + 2u) & 3u])[tid] = pend;
                    int j = i + 3;
                    if (j >= L.n_ops) j -= L.n_ops;
                    if (j >= L.n_ops) j = 0;
                    pend = reinterpret_cast<const int*>(L.ops + j)[tid];
                }
                const float* wseg = L.blob + op.w_off;
                if (op.w_len > 0 && stage) {
                    const uint32_t slot = wcount & 1u;
                    if (tid == 0) {   // prefetch the next parameter segment into its staging buffer
                        // the staging buffers recycle arena space that generic-proxy stores wrote before
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                        if (op.flags & SBC_F_LATEW) {   // own segment first: nobody prefetched it
                            sbc_mbar_expect_tx(&bars[slot], (uint32_t)op.w_len * 4u);
                            sbc_bulk_g2s(arena + op.wbuf, L.blob + op.w_off, (uint32_t)op.w_len * 4u, &bars[slot]);
                        }
                        // next op with parameters; the last one wraps into the next forward
                        if (op.nw_len > 0 && (op.next_w >= 0 || !(last_step && last_sample))) {
                            sbc_mbar_expect_tx(&bars[slot ^ 1u], (uint32_t)op.nw_len * 4u);
                            sbc_bulk_g2s(arena + op.nw_buf, L.blob + op.nw_off, (uint32_t)op.nw_len * 4u, &bars[slot ^ 1u]);
                        }
                    }
                    sbc_mbar_wait(&bars[slot], (wcount >> 1) & 1u);
                    wseg = arena + op.wbuf;
                    wcount++;
                }
                const long long t_wait = (INSTR && do_prof) ? clock64() : 0;
                const SbcGeo& GS = L.geo[op.sgeo];
                const SbcGeo& GD = L.geo[op.dgeo];
                long long* sub = nullptr;
                int kind = op.kind, dbg = 0;
                if (INSTR) {
                    sub = do_prof ? L.prof + L.n_ops + 2 + 4 * i : nullptr;   // intra-op stamps (thread 0)
                    if (sub) { sub[0] = sub[1] = sub[2] = sub[3] = 0; L.prof[5 * L.n_ops + 2 + i] = t_wait; }
                    dbg = L.dbg;
                    if ((dbg & 16) || ((dbg & 4) && kind != SBC_OP_CONV_MMA)) kind = -1;   // timing experiments
                }
                if (kind == SBC_OP_CONV_MMA) {
                    sbc_conv_mma<X3, SMEM_ARENA>(op, GS, GD, arena, park, wseg, tid, sub, dbg);
                } else if (kind == SBC_OP_NORM_ELU) {
                    sbc_norm_op(op, GS, arena, wseg, tid);
                } else if (kind == SBC_OP_MAXPOOL5) {
                    sbc_maxpool5_op(op, GS, arena, tid, SBC_NTHREADS);
                } else if (kind == SBC_OP_ELU) {
                    sbc_elu_op(op, GS, arena, tid, SBC_NTHREADS);
                } else if (kind == SBC_OP_UPACC) {
                    sbc_upacc_op(op, GS, GD, arena, tid, SBC_NTHREADS);
                } else if (kind == SBC_OP_AFFINE) {
                    sbc_affine_op(op, GS, arena, tid, SBC_NTHREADS);
                } else if (kind == SBC_OP_SPILL) {
                    sbc_spill_op(op, arena, park, tid, SBC_NTHREADS);
                } else if (kind == SBC_OP_FILL) {
                    sbc_fill_op(op, arena, park, tid, SBC_NTHREADS);
                }
                // fresh outputs whose halo the planner could not prove clean (interior and halo cells are disjoint)
                if (op.flags & (SBC_F_ZH_DST | SBC_F_ZH_EDST)) {
                    const uint16_t* hl = halo + L.halo_off[op.dgeo];
                    const int nh = GD.pps - GD.h * GD.w;
                    if (op.flags & SBC_F_ZH_DST) sbc_zero_halo_list(arena + op.dst, GD, hl, nh, op.cout, tid);
                    if (op.flags & SBC_F_ZH_EDST) sbc_zero_halo_list(arena + op.edst, GD, hl, nh, op.cout, tid);
                }
                if (INSTR && sub) sub[3] = clock64();
                __syncthreads();
            }
            if (INSTR && L.debug_stop >= 0) {   // debugging aid: dump the arena of sample 0 and stop
                if (b == 0) {   // arena, then the park area (debug_out holds arena_floats + park_floats)
                    for (int i = tid; i < L.arena_floats; i += SBC_NTHREADS) L.debug_out[i] = arena[i];
                    if (park)
                        for (int i = tid; i < L.park_floats; i += SBC_NTHREADS) L.debug_out[L.arena_floats + i] = park[i];
                }
                return;
            }

            if (INSTR && do_prof) L.prof[L.n_ops] = clock64();
            const float* net = arena + L.out_off;   // compact (re, im) pairs
            if (L.mode == 2) {
                // target = -1/sigma^2 * noise;  loss = 1/2 * sum((score - target)^2) * sigma^anneal_power   (dsm.py:19-31)
                long long lab = L.labels[b];
                lab = lab < 0 ? 0 : (lab >= L.n_sigmas ? L.n_sigmas - 1 : lab);
                const float sg = L.sigmas[lab], isg2 = 1.f / (sg * sg);
                const float* z = L.dsm_z + (size_t)b * L.channels * ne;
                float part = 0.f;
                for (int e = tid; e < ne; e += SBC_NTHREADS) {
                    const float2 v = reinterpret_cast<const float2*>(net)[e];
                    const float dx = v.x / sg + isg2 * (z[e] * sg), dy = v.y / sg + isg2 * (z[ne + e] * sg);
                    part += dx * dx + dy * dy;
                }
                const float tot = sbc_block_sum(part, s_red, tid);
                if (tid == 0) L.dsm_out[b] = 0.5f * tot * powf(sg, L.anneal_power);
            } else if (L.mode == 0) {
                // score = net / sigmas[y]   (ncsnv2.py:295-298)
                long long lab = L.labels[b];
                if (lab < 0) lab = 0;
                if (lab >= L.n_sigmas) lab = L.n_sigmas - 1;
                const float sg = L.sigmas[lab];
                float* o = L.fout + (size_t)b * L.channels * ne;
                for (int e = tid; e < ne; e += SBC_NTHREADS) {
                    const float2 v = reinterpret_cast<const float2*>(net)[e];
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
                sbc_dc_residual(ax, res, Pm, Ym, Nt, Nr, L.Np, tid, SBC_NTHREADS);
                __syncthreads();
                const unsigned long long sid = L.sample_ids ? L.sample_ids[b] : (unsigned long long)b;
                const uint32_t gstep = (uint32_t)(lvl * L.steps_each + gs % L.steps_each);
                const float part = sbc_langevin_update(ax, net, res, Pm, Hc, en, s_sc, L.seed, sid, gstep, Nt, Nr,
                                                       L.Np, tid, SBC_NTHREADS);
                if (L.nmse_log && Hc) {
                    const float tot = sbc_block_sum(part, s_red, tid);
                    if (tid == 0) L.nmse_log[(size_t)gs * L.B + b] = tot / s_hnorm;
                }
            }
            __syncthreads();
            if (INSTR && do_prof) L.prof[L.n_ops + 1] = clock64();
        }

        if (L.mode == 1) {   // write the final estimate back (interleaved complex64)
            float2* X = reinterpret_cast<float2*>(L.X) + (size_t)b * ne;
            for (int e = tid; e < ne; e += SBC_NTHREADS) X[e] = reinterpret_cast<const float2*>(ax)[e];
        }
        __syncthreads();
    }
    if (INSTR && L.prof != nullptr && tid == 0) {   // per-CTA totals after the per-op stamps: (cycles, SM id) per CTA
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        L.prof[6 * L.n_ops + 2 + 2 * blockIdx.x] = clock64() - t_cta0;
        L.prof[6 * L.n_ops + 3 + 2 * blockIdx.x] = (long long)smid;
    }
}
