/*
 * sbc.h -- C ABI of the B200 annealed-Langevin MIMO channel-estimation library (libsbc_b200.so).
 *
 * The reference (utcsilab/score-based-channels) is pure Python over torch and has no FFI of its
 * own; each entry point below replaces the span of reference code cited next to it and is what a
 * maintainer would bind with ctypes.  Plain pointers and sizes only -- no
 * torch types.  All functions return 0 on success or a negative SBC_E_* code; the message of the
 * last failure on the calling thread is available from sbc_last_error().  Nothing here falls back
 * to a CPU implementation: without a CUDA device every compute entry point fails with SBC_E_CUDA.
 *
 * Threading / streams: every compute call only enqueues work on the caller's stream (a
 * cudaStream_t passed as void*, NULL = legacy default stream); no hidden synchronisation, no
 * allocation (workspaces are created in sbc_model_create).  A handle may be used from several
 * streams as long as calls that share the handle's global-arena workspace (large Nt x Nr only, see
 * sbc_info.arena_in_smem) are not concurrent.
 *
 * RNG contract (noise when ext_noise == NULL): Philox4x32-10, key = (seed lo, seed hi), counter =
 * (element >> 1, level * steps_each + inner_step, sample_id lo, sample_id hi); the four 32-bit
 * outputs give two Box-Muller pairs (u = ((r >> 8) + 0.5) * 2^-24), element 2k uses outputs 0,1 and
 * element 2k+1 outputs 2,3; (re, im) = (r cos, r sin) / sqrt(2), i.e. CN(0,1) like
 * torch.randn_like(complex) at reference test_score.py:161.  Results therefore depend only on
 * (seed, sample_id), never on batch composition, launch geometry or GPU count.
 */
#ifndef SBC_H_
#define SBC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SBC_VERSION 210   /* 0.2.1: engine 1 plans for two CTAs per SM (park area: sbc_model_desc.park_floats appended);
                           * 0.2.0: engine 2 (tcgen05): sbc_model_create_from_state, sbc_info gained engine / ctas_per_sm /
                           * group_size / n_ops (appended), debug views sbc_debug_plan / sbc_debug_run */

enum {
    SBC_OK = 0,
    SBC_E_ARG = -1,      /* invalid argument (shape constraint, null pointer, ...) */
    SBC_E_CUDA = -2,     /* CUDA runtime failure (message has the CUDA error string) */
    SBC_E_NOMEM = -3,
    SBC_E_UNSUPPORTED = -4
};

/* Packed model: the layer program + parameter blob produced by the host-side packer
 * (score_based_channels_b200/program.py) from a reference state_dict.  Replaces
 * `NCSNv2Deepest(config)` + `load_state_dict` + `.cuda()` (reference test_score.py:59-63). */
typedef struct sbc_model_desc {
    int32_t ngf;            /* config.model.ngf */
    int32_t Nt, Nr;         /* network input is [B, channels, Nt, Nr] (reference test_score.py:149) */
    int32_t channels;       /* 2 (re, im) */
    const int32_t* op_table;/* host, [n_ops][32] int32 (csrc/sbc_program.h: SbcOp) */
    int32_t n_ops;
    const int32_t* geo_table;/* host, [n_geo][8] int32 (csrc/sbc_program.h: SbcGeo), entry 0 = Nt x Nr */
    int32_t n_geo;
    const float* blob;      /* host, packed parameters */
    int64_t blob_floats;
    int32_t arena_floats;   /* per-sample activation arena size */
    int32_t in_off, out_off, post_off;
    int32_t max_w_len;      /* largest per-op parameter segment (floats) */
    const float* sigmas;    /* host, [n_sigmas] fp32 noise schedule (reference ncsnv2/models/__init__.py:4-8) */
    int32_t n_sigmas;
    int64_t conv_flops;     /* dense conv FLOP / forward / sample (reported by sbc_query) */
    int32_t nthreads;       /* threads per CTA the program was planned for: must equal sbc_threads_per_cta() */
    int32_t park_floats;    /* per-CTA park area in global memory (SPILL / FILL ops, SBC_F_ACC_G); 0 = none */
} sbc_model_desc;

typedef struct sbc_info {
    int32_t version;
    int32_t device;
    int32_t num_sms;
    int32_t threads_per_cta;
    int32_t arena_in_smem;      /* 1: activations never leave shared memory */
    int32_t weights_staged;     /* 1: per-op parameters streamed into shared memory with cp.async.bulk */
    int64_t smem_bytes_per_cta;
    int64_t arena_bytes;
    int64_t conv_flops_per_forward;
    int64_t kernel_launches;    /* launches issued through this handle so far */
    int32_t engine;             /* 1: fused shared-memory arena on mma.sync; 2: tcgen05 / TMEM, L2-resident arena */
    int32_t ctas_per_sm;        /* resident CTAs per SM (engine 2: 2) */
    int32_t group_size;         /* engine 2: samples per CTA group (S) of the last launch */
    int32_t n_ops;              /* ops of the layer program */
} sbc_info;

/* One tensor of a reference state dict (same key names as NCSNv2Deepest.state_dict(), reference
 * ncsnv2/models/ncsnv2.py:198-262; 'sigmas' carries the noise schedule).  fp32, C-contiguous, host memory. */
typedef struct sbc_state_entry {
    const char* name;
    const float* data;
    const int64_t* shape;
    int32_t ndim;
} sbc_state_entry;

/* Debug view of one activation tensor of an engine-2 plan (sbc_debug_plan). */
typedef struct sbc_tensor_info {
    char name[48];
    int32_t fmt;                /* 0: F32 [C/4][npx][4], 1: SP16 [C/8][hi,lo][npx][8 halfs], 2: raw */
    int32_t level, C;
    int64_t off, bytes;         /* inside the group arena */
    int32_t born, died;         /* live range in op indices */
} sbc_tensor_info;

/* One annealed-Langevin run over levels [level_begin, level_end) x steps_each for B independent
 * channel realisations.  Replaces the loop body reference test_score.py:135-171 (duplicated at
 * tune_hparams_score.py:112-148).  Per-sample noise_var / alpha_step / beta let one launch mix SNR
 * points and hyper-parameter cells. */
typedef struct sbc_ald_args {
    int32_t B, Nt, Nr, Np;
    int32_t level_begin, level_end, steps_each;
    const void* P;            /* [B,Np,Nt] complex64: `forward` = val_P        (test_score.py:128) */
    const void* Y;            /* [B,Np,Nr] complex64: `y` = val_Y              (test_score.py:122-127) */
    void* X;                  /* [B,Nt,Nr] complex64 in/out: `current`         (test_score.py:126,164) */
    const void* H_oracle;     /* [B,Nt,Nr] complex64 ground truth or NULL      (test_score.py:131) */
    const float* noise_var;   /* [B] local_noise = 10^(-snr/10)*Nt             (test_score.py:75) */
    const float* alpha_step;  /* [B]                                           (test_score.py:46-48) */
    const float* beta;        /* [B] beta_noise                                (test_score.py:46-48) */
    double sigma_end;         /* config.model.sigma_end                        (test_score.py:144) */
    float* nmse_log;          /* [steps,B] per-step NMSE or NULL               (test_score.py:168-170) */
    uint64_t seed;
    const uint64_t* sample_ids; /* [B] global sample ids (NULL: 0..B-1) */
    const void* ext_noise;    /* [steps,B,Nt,Nr] complex64 unit-power noise replacing the RNG, or NULL */
    const float* dc_boost;    /* [B] multiplier of the data-consistency term or NULL (= 1)   (test_mmse.py:25,246) */
    const int32_t* stop_step; /* [B] last step index a sample executes (early stop, `target_stop` of
                               * test_mmse.py:173,260-263) or NULL (all steps); later NMSE-log rows of that
                               * sample are left untouched */
} sbc_ald_args;

int sbc_version(void);
int sbc_threads_per_cta(void);   /* compile-time CTA size of the fused kernel (the host planner needs it) */
const char* sbc_last_error(void);

int sbc_model_create(const sbc_model_desc* desc, int device, void** handle_out);
/* Engine 2 (tcgen05): builds the layer program, plans the activation arena and packs the parameters inside the
 * library from a plain state dict -- no Python planner involved.  Replaces `NCSNv2Deepest(config)` +
 * `load_state_dict` + `.cuda()` (reference test_score.py:59-63).  Conv arithmetic: fp16 hi/lo split operands
 * ("fp16x2", 22 significant bits, fp32 accumulate in TMEM): fp32-equivalent. */
int sbc_model_create_from_state(const sbc_state_entry* entries, int32_t n_entries, int32_t ngf, int32_t Nt, int32_t Nr,
                                int32_t channels, int device, void** handle_out);
/* The same with the engine chosen by `precision`: SBC_PREC_FP16X2 = the tcgen05 engine (as above); SBC_PREC_TF32X3 /
 * SBC_PREC_TF32 = engine 1 (fused shared-memory arena on mma.sync; two channel realisations per SM at 64x16), planned
 * and packed inside the library by the C++ twin of program.py (csrc/sbc1_plan.h) -- no Python needed for either. */
enum { SBC_PREC_FP16X2 = 0, SBC_PREC_TF32X3 = 1, SBC_PREC_TF32 = 2 };
int sbc_model_create_from_state_ex(const sbc_state_entry* entries, int32_t n_entries, int32_t ngf, int32_t Nt,
                                   int32_t Nr, int32_t channels, int device, int32_t precision, void** handle_out);
int sbc_model_free(void* handle);
int sbc_query(void* handle, sbc_info* out);

/* Host-only view of an engine-1 plan built by the in-library planner (no CUDA call: usable without a GPU; the CPU
 * tests compare it word for word with program.py).  park: -1 = automatic choice, 0 / 1 = force.  The view points into
 * the plan object: valid until sbc_plan1_free(plan). */
typedef struct sbc_plan1_view {
    const int32_t* op_table;    /* [n_ops][32] */
    int32_t n_ops;
    const int32_t* geo_table;   /* [n_geo][8] */
    int32_t n_geo;
    const float* blob;
    int64_t blob_floats;
    int32_t arena_floats, in_off, out_off, post_off, max_w_len, park_floats, nthreads;
    int64_t conv_flops;
} sbc_plan1_view;
int sbc_plan1_build(const sbc_state_entry* entries, int32_t n_entries, int32_t ngf, int32_t Nt, int32_t Nr,
                    int32_t channels, int32_t nthreads, int32_t precision, int32_t park, void** plan_out,
                    sbc_plan1_view* view);
int sbc_plan1_free(void* plan);

/* NCSNv2Deepest.forward(x, y) (reference ncsnv2/models/ncsnv2.py:269-300).  Device pointers.
 * x: fp32 [B,channels,Nt,Nr] with element strides x_strides (any layout, e.g. the permuted
 * view_as_real view of test_score.py:149); labels: int64 [B]; out: fp32 contiguous. */
int sbc_forward(void* handle, const float* x, const int64_t x_strides[4], const int64_t* labels, float* out,
                int32_t B, void* stream);

/* Denoising-score-matching loss, forward only: anneal_dsm_score_estimation (reference ncsnv2/losses/dsm.py:6-32) as it is
 * evaluated under torch.no_grad() for the validation loss of train_score.py:170-185.  Device pointers; samples and z
 * (the torch.randn_like draw) are fp32 [B,channels,Nt,Nr] contiguous, labels int64 [B]; one launch perturbs, runs the
 * score network and reduces  loss_out[b] = 1/2 * sum((score - target)^2) * sigmas[labels[b]]^anneal_power  (the caller
 * takes the mean).  Engine-1 models.  The backward pass / optimiser step of the training loop is not part of this library. */
int sbc_dsm_loss(void* handle, const float* samples, const int64_t* labels, const float* z, float anneal_power,
                 float* loss_out, int32_t B, void* stream);

/* Device-pointer ALD run (all arrays in sbc_ald_args are device memory). */
int sbc_ald_run(void* handle, const sbc_ald_args* args, void* stream);

/* Host-buffer variants: same semantics, all arrays are host memory; copies in, runs, copies out and
 * synchronises the stream before returning (the end-to-end path for non-torch callers). */
int sbc_forward_host(void* handle, const float* x, const int64_t* labels, float* out, int32_t B);
int sbc_ald_run_host(void* handle, const sbc_ald_args* args);

/* Debug aid: run the layer program of sample x (device fp32 [channels,Nt,Nr], contiguous) up to but
 * excluding op `stop_op` (n_ops = all) and copy the whole arena, followed by the CTA's park area, to arena_out
 * (device, arena_floats + park_floats). */
int sbc_debug_arena(void* handle, const float* x, int32_t stop_op, float* arena_out, void* stream);

/* Engine-2 debug views.  sbc_debug_plan: tensor table (up to `cap` entries), arena size and the four level
 * geometries ([4][14] int32: h, w, hy, hx, wp, rps, pps, lead, npx, T, slot, hw, 2 magic words) of the plan for group size S
 * (reuse = 0: every tensor gets its own region, so a whole forward can be inspected afterwards).
 * sbc_debug_run: one forward of S samples x (device fp32 [S,channels,Nt,Nr] contiguous) as ONE group, then the
 * group arena is copied to arena_out (device).  sbc_op_name / sbc_op_kind: the layer program. */
int sbc_debug_plan(void* handle, int32_t S, int32_t reuse, sbc_tensor_info* out, int32_t cap, int32_t* n_out,
                   int64_t* arena_bytes, int32_t* geo_out);
int sbc_debug_run(void* handle, const float* x, int32_t S, int32_t reuse, void* arena_out, void* stream);
const char* sbc_op_name(void* handle, int32_t i);
int sbc_op_kind(void* handle, int32_t i);

/* Profiling aid: subsequent launches of this handle make CTA 0 record clock64() at every op
 * boundary of its first sample / second step (first if there is only one) into dev_stamps (device int64 [6 * n_ops + 2]: op starts,
 * end of network, end of the Langevin update, then 4 intra-op stamps per op: conv = {prologue done, K loop
 * entered, K loop done, epilogue done}, then per op the time at which its record and parameters were in hand).
 * NULL switches it off. */
int sbc_set_profile_buffer(void* handle, int64_t* dev_stamps);

#ifdef __cplusplus
}
#endif
#endif /* SBC_H_ */
