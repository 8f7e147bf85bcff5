// sbc_api.cu -- C ABI (include/sbc.h) over the fused ALD kernel.  Pure CUDA runtime; no torch types.
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <memory>
#include <vector>

#include "../../include/sbc.h"
#include "sbc_kernel.cuh"
#include "sbc2_host.cuh"
#include "sbc1_plan.h"

static thread_local char g_err[512] = "";

static int sbc_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define SBC_CUDA(call)                                                                            \
    do {                                                                                          \
        cudaError_t e_ = (call);                                                                  \
        if (e_ != cudaSuccess)                                                                    \
            return sbc_fail(SBC_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct SbcModel {
    int engine = 1;                // 1: fused shared-memory arena, mma.sync (sbc_kernel.cuh); 2: tcgen05 (sbc2_kernel.cuh)
    Sbc2Model* e2 = nullptr;
    int device = 0;
    int num_sms = 0;
    int smem_optin = 0;
    sbc_model_desc d{};
    SbcOp* d_ops = nullptr;
    float* d_blob = nullptr;
    float* d_sigmas = nullptr;
    float* d_gws = nullptr;
    float* d_park = nullptr;       // park area of two-CTAs-per-SM plans: park_floats per resident CTA
    int ctas_per_sm = 1;
    int first_w = -1;
    bool arena_in_smem = false;
    bool stage = false;
    bool x3 = false;               // every conv of the program uses the 3xTF32 product
    size_t smem_bytes = 0;
    long long launches = 0;
    long long* d_prof = nullptr;   // optional per-op clock stamps (sbc_set_profile_buffer)
    SbcGeo geo[SBC_MAX_GEO];
    int n_geo = 0;
    int halo_off[SBC_MAX_GEO] = {0};
    // grow-only device buffers of the host-buffer entry points (sbc_*_host): allocated once, reused by every call
    void* hws[16] = {nullptr};
    size_t hws_cap[16] = {0};
    void* host_ws(int slot, size_t bytes) {
        if (bytes > hws_cap[slot]) {
            cudaFree(hws[slot]);
            hws[slot] = nullptr; hws_cap[slot] = 0;
            const size_t want = bytes + bytes / 4 + 256;
            if (cudaMalloc(&hws[slot], want) != cudaSuccess) return nullptr;
            hws_cap[slot] = want;
        }
        return hws[slot];
    }
    ~SbcModel() {   // owns its device buffers: every early return of sbc_model_create releases them
        cudaFree(d_ops); cudaFree(d_blob); cudaFree(d_sigmas); cudaFree(d_gws); cudaFree(d_park);
        for (auto p : hws) cudaFree(p);
        delete e2;
    }
};

extern "C" int sbc_version(void) { return SBC_VERSION; }
extern "C" int sbc_threads_per_cta(void) { return SBC_NTHREADS; }
extern "C" const char* sbc_last_error(void) { return g_err; }

static int env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

extern "C" int sbc_model_create(const sbc_model_desc* desc, int device, void** handle_out) {
    if (!desc || !handle_out) return sbc_fail(SBC_E_ARG, "sbc_model_create: null argument");
    if (!desc->op_table || desc->n_ops <= 0 || !desc->blob || desc->blob_floats <= 0 || !desc->sigmas ||
        desc->n_sigmas <= 0)
        return sbc_fail(SBC_E_ARG, "sbc_model_create: empty program / blob / sigmas");
    if (desc->Nt % 8 || desc->Nr % 8 || desc->Nt <= 0 || desc->Nr <= 0)
        return sbc_fail(SBC_E_ARG, "sbc_model_create: Nt (%d) and Nr (%d) must be positive multiples of 8", desc->Nt,
                        desc->Nr);
    if (desc->channels != 2) return sbc_fail(SBC_E_ARG, "sbc_model_create: channels must be 2 (re, im)");
    if (desc->nthreads != SBC_NTHREADS)
        return sbc_fail(SBC_E_ARG, "sbc_model_create: program planned for %d threads per CTA, library built for %d",
                        desc->nthreads, SBC_NTHREADS);
    if (desc->arena_floats % 4 || desc->max_w_len % 4)
        return sbc_fail(SBC_E_ARG, "sbc_model_create: arena_floats and max_w_len must be multiples of 4");
    int ndev = 0;
    SBC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return sbc_fail(SBC_E_ARG, "sbc_model_create: no CUDA device %d", device);
    SBC_CUDA(cudaSetDevice(device));
    std::unique_ptr<SbcModel> mh(new SbcModel());
    SbcModel* m = mh.get();
    m->device = device;
    m->d = *desc;
    if (!desc->geo_table || desc->n_geo <= 0 || desc->n_geo > SBC_MAX_GEO) { return sbc_fail(SBC_E_ARG, "sbc_model_create: bad geometry table"); }
    memset(m->geo, 0, sizeof m->geo);
    memcpy(m->geo, desc->geo_table, sizeof(SbcGeo) * (size_t)desc->n_geo);
    m->n_geo = desc->n_geo;
    int halo_total = 0;   // uint16 entries of the per-geometry halo-pixel lists (shared misc region)
    for (int g = 0; g < desc->n_geo; g++) {
        if (m->geo[g].pps > 65535) { return sbc_fail(SBC_E_UNSUPPORTED, "geometry %d: plane too large", g); }
        m->halo_off[g] = halo_total;
        halo_total += m->geo[g].pps - m->geo[g].h * m->geo[g].w;
    }
    if (m->geo[0].h != desc->Nt || m->geo[0].w != desc->Nr) { return sbc_fail(SBC_E_ARG, "sbc_model_create: geometry 0 must be Nt x Nr"); }
    SBC_CUDA(cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device));
    SBC_CUDA(cudaDeviceGetAttribute(&m->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));

    // validate + annotate the op table (pad0 := index of the next op that has parameters)
    std::vector<SbcOp> ops(desc->n_ops);
    memcpy(ops.data(), desc->op_table, sizeof(SbcOp) * (size_t)desc->n_ops);
    int next = -1, nconv = 0;
    for (int i = desc->n_ops - 1; i >= 0; i--) {
        SbcOp& o = ops[i];
        o.next_w = next;
        if (o.w_len > 0) next = i;
        if (o.kind < 0 || o.kind > SBC_OP_LAST) { return sbc_fail(SBC_E_ARG, "op %d: bad kind %d", i, o.kind); }
        if (o.w_len < 0 || o.w_len % 4 || o.w_off % 4 || (long long)o.w_off + o.w_len > desc->blob_floats ||
            o.w_len > desc->max_w_len) { return sbc_fail(SBC_E_ARG, "op %d: bad parameter segment", i); }
        if (o.sgeo < 0 || o.sgeo >= SBC_MAX_GEO || o.dgeo < 0 || o.dgeo >= SBC_MAX_GEO) {
            return sbc_fail(SBC_E_ARG, "op %d: bad geometry index", i);
        }
        {   // every arena offset an op touches must lie inside the arena (a malformed table must not write out of bounds)
            if (o.kind == SBC_OP_SPILL || o.kind == SBC_OP_FILL) {   // whole-tensor copies between the arena and the park area
                const int a = o.kind == SBC_OP_SPILL ? o.src : o.dst, p = o.kind == SBC_OP_SPILL ? o.dst : o.src;
                const long long n = 4ll * o.MT;
                if (o.MT <= 0 || a < 0 || a % 4 || p < 0 || p % 4 || a + n > desc->arena_floats || p + n > desc->park_floats) {
                    return sbc_fail(SBC_E_ARG, "op %d: spill / fill outside the arena or the park area", i);
                }
                continue;
            }
            const bool acc_g = (o.flags & SBC_F_ACC_G) != 0;
            if (acc_g && (o.kind != SBC_OP_CONV_MMA || o.acc < 0)) { return sbc_fail(SBC_E_ARG, "op %d: ACC_G without a conv accumulator", i); }
            const int offs[5] = {o.src, o.dst, acc_g ? -1 : o.acc, o.edst, o.scratch};
            for (int k = 0; k < 5; k++)
                if (offs[k] < -1 || offs[k] >= desc->arena_floats) { return sbc_fail(SBC_E_ARG, "op %d: arena offset out of range", i); }
            const SbcGeo& gs = m->geo[o.sgeo]; const SbcGeo& gd = m->geo[o.dgeo];
            const long long in_ext = (long long)((o.cin + 3) / 4) * gs.pps * 4, out_ext = (long long)((o.cout + 3) / 4) * gd.pps * 4;
            const bool compact_in = o.kind == SBC_OP_AFFINE, compact_out = (o.kind == SBC_OP_CONV_MMA) && (o.flags & SBC_F_COMPACT);
            if (o.src >= 0 && !compact_in && o.src + in_ext > desc->arena_floats) { return sbc_fail(SBC_E_ARG, "op %d: input tensor exceeds the arena", i); }
            if (o.dst >= 0 && !compact_out && o.dst + out_ext > desc->arena_floats) { return sbc_fail(SBC_E_ARG, "op %d: output tensor exceeds the arena", i); }
            if (o.acc >= 0 && o.acc + out_ext > (acc_g ? desc->park_floats : desc->arena_floats)) { return sbc_fail(SBC_E_ARG, "op %d: accumulator tensor exceeds the arena", i); }
            if (o.edst >= 0 && o.edst + out_ext > desc->arena_floats) { return sbc_fail(SBC_E_ARG, "op %d: ELU output tensor exceeds the arena", i); }
        }
        if (o.w_len > 0 && (o.wbuf < 0 || o.wbuf % 4 || o.wbuf + o.w_len > desc->arena_floats)) {
            return sbc_fail(SBC_E_ARG, "op %d: bad parameter staging buffer", i);
        }
        if (o.kind == SBC_OP_CONV_MMA) {
            const int E = 2;
            const bool ok = o.ks >= 1 && o.ks <= SBC_NTHREADS / 32 && (o.ks & (o.ks - 1)) == 0 &&
                            (o.ksize == 1 || o.ksize == 3) && o.tapmask != 0 && (o.ks == 1 || o.scratch >= 0) &&
                            (o.ks == 1 || (o.flags & SBC_F_UNIT)) &&
                            (!(o.flags & SBC_F_UNIT) || o.MT * o.NT * o.ks <= SBC_NTHREADS / 32) &&
                            o.cout % 2 == 0 && o.MT == (o.oh * o.ow + 15) / 16 && o.NT == (o.cout + 7) / 8 &&
                            o.S >= 1 && o.frag_rel >= o.S && o.frag_rel % 4 == 0 &&
                            o.frag_rel + o.S * o.NT * 32 * E <= o.w_len;
            if (!ok) { return sbc_fail(SBC_E_ARG, "op %d: unsupported tensor-core conv", i); }
            const bool ox3 = (o.flags & SBC_F_X3) != 0;
            if (nconv++ == 0) m->x3 = ox3;
            else if (m->x3 != ox3) { return sbc_fail(SBC_E_ARG, "op %d: mixed conv precisions", i); }
        }
        if ((o.kind == SBC_OP_NORM_ELU || o.kind == SBC_OP_ELU || o.kind == SBC_OP_MAXPOOL5 || o.kind == SBC_OP_UPACC) &&
            o.cin % 8) { return sbc_fail(SBC_E_ARG, "op %d: channel count must be a multiple of 8", i); }
    }
    // parameter segment of the next parameterised op (the last one wraps to the first)
    for (int i = 0; i < desc->n_ops; i++) {
        SbcOp& o = ops[i];
        const int j = o.next_w >= 0 ? o.next_w : next;
        o.nw_off = j >= 0 ? ops[j].w_off : 0;
        o.nw_len = (j >= 0 && !(ops[j].flags & SBC_F_LATEW)) ? ops[j].w_len : 0;   // late segments load themselves
        o.nw_buf = j >= 0 ? ops[j].wbuf : 0;
    }
    m->first_w = next;
    if (next >= 0 && (ops[next].flags & SBC_F_LATEW)) { return sbc_fail(SBC_E_ARG, "the first parameterised op must not be a late-loading one"); }

    // where do activations live, and are parameters staged through shared memory?
    cudaFuncAttributes fa{};
    SBC_CUDA(cudaFuncGetAttributes(&fa, sbc_ald_kernel<true, true, true>));
    const int dyn_max = m->smem_optin - (int)fa.sharedSizeBytes;   // opt-in limit covers static + dynamic
    const size_t arena_bytes = (size_t)desc->arena_floats * 4;
    const size_t misc = ((size_t)SBC_MISC_BARS + (size_t)halo_total * 2 + 15) / 16 * 16;
    m->arena_in_smem = !env_int("SBC_FORCE_GLOBAL_ARENA", 0) && arena_bytes + misc <= (size_t)dyn_max;
    // parameter staging buffers live inside the arena: staging needs the arena in shared memory
    m->stage = m->arena_in_smem && env_int("SBC_STAGE_WEIGHTS", 1) != 0;
    m->smem_bytes = (m->arena_in_smem ? arena_bytes : 0) + misc;

    SBC_CUDA(cudaMalloc(&m->d_ops, sizeof(SbcOp) * (size_t)desc->n_ops));
    SBC_CUDA(cudaMemcpy(m->d_ops, ops.data(), sizeof(SbcOp) * (size_t)desc->n_ops, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMalloc(&m->d_blob, sizeof(float) * (size_t)desc->blob_floats));
    SBC_CUDA(cudaMemcpy(m->d_blob, desc->blob, sizeof(float) * (size_t)desc->blob_floats, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMalloc(&m->d_sigmas, sizeof(float) * (size_t)desc->n_sigmas));
    SBC_CUDA(cudaMemcpy(m->d_sigmas, desc->sigmas, sizeof(float) * (size_t)desc->n_sigmas, cudaMemcpyHostToDevice));
    m->d.op_table = nullptr; m->d.blob = nullptr; m->d.sigmas = nullptr; m->d.geo_table = nullptr;   // host pointers are not retained

    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<true, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<false, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<true, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    SBC_CUDA(cudaFuncSetAttribute(sbc_ald_kernel<false, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max));
    // resident CTAs per SM: plans made for two CTAs per SM (program.py, park mode) keep their arena under half of the
    // shared memory of an SM; the grid is sized to fill every slot (any grid is correct: CTAs stride over the batch)
    m->ctas_per_sm = 1;
    {   // (a global-memory arena needs no shared memory beyond the misc region: registers bound the residency)
        int nb = 0;
        if (m->arena_in_smem) {
            if (m->x3) SBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sbc_ald_kernel<true, true, false>, SBC_NTHREADS, m->smem_bytes));
            else SBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sbc_ald_kernel<true, false, false>, SBC_NTHREADS, m->smem_bytes));
        } else {
            if (m->x3) SBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sbc_ald_kernel<false, true, false>, SBC_NTHREADS, m->smem_bytes));
            else SBC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, sbc_ald_kernel<false, false, false>, SBC_NTHREADS, m->smem_bytes));
        }
        const int cap = env_int("SBC_CTAS", 0);
        if (cap > 0 && nb > cap) nb = cap;
        m->ctas_per_sm = nb < 1 ? 1 : nb;
    }
    if (desc->park_floats < 0 || desc->park_floats % 4) { return sbc_fail(SBC_E_ARG, "sbc_model_create: bad park_floats"); }
    if (!m->arena_in_smem) SBC_CUDA(cudaMalloc(&m->d_gws, arena_bytes * (size_t)m->num_sms * (size_t)m->ctas_per_sm));
    if (desc->park_floats > 0)
        SBC_CUDA(cudaMalloc(&m->d_park, sizeof(float) * (size_t)desc->park_floats * (size_t)m->num_sms * (size_t)m->ctas_per_sm));
    *handle_out = mh.release();
    return SBC_OK;
}

extern "C" int sbc_model_free(void* handle) {
    if (!handle) return SBC_OK;
    SbcModel* m = (SbcModel*)handle;
    cudaSetDevice(m->engine == 2 ? m->e2->device : m->device);
    delete m;
    return SBC_OK;
}

extern "C" int sbc_query(void* handle, sbc_info* out) {
    if (!handle || !out) return sbc_fail(SBC_E_ARG, "sbc_query: null argument");
    SbcModel* m = (SbcModel*)handle;
    if (m->engine == 2) {
        const Sbc2Model* e = m->e2;
        memset(out, 0, sizeof *out);
        out->version = SBC_VERSION; out->device = e->device; out->num_sms = e->num_sms; out->threads_per_cta = SBC2_NTHR;
        out->arena_in_smem = 0; out->weights_staged = 1; out->smem_bytes_per_cta = (int64_t)e->smem_bytes;
        out->arena_bytes = e->last_grid > 0 ? (int64_t)(e->gws_bytes / (size_t)e->last_grid) : 0;
        out->conv_flops_per_forward = e->builder->conv_flops; out->kernel_launches = e->launches;
        out->engine = 2; out->ctas_per_sm = e->ctas_per_sm; out->group_size = e->last_S; out->n_ops = e->builder->n_ops();
        return SBC_OK;
    }
    out->engine = 1; out->ctas_per_sm = m->ctas_per_sm; out->group_size = 1; out->n_ops = m->d.n_ops;
    out->version = SBC_VERSION;
    out->device = m->device;
    out->num_sms = m->num_sms;
    out->threads_per_cta = SBC_NTHREADS;
    out->arena_in_smem = m->arena_in_smem;
    out->weights_staged = m->stage;
    out->smem_bytes_per_cta = (int64_t)m->smem_bytes;
    out->arena_bytes = (int64_t)m->d.arena_floats * 4;
    out->conv_flops_per_forward = m->d.conv_flops;
    out->kernel_launches = m->launches;
    return SBC_OK;
}

static void fill_common(const SbcModel* m, SbcLaunch& L) {
    memset(&L, 0, sizeof L);
    L.ops = m->d_ops; L.n_ops = m->d.n_ops; L.first_w = m->first_w; L.blob = m->d_blob;
    memcpy(L.geo, m->geo, sizeof L.geo);
    L.n_geo = m->n_geo;
    memcpy(L.halo_off, m->halo_off, sizeof L.halo_off);
    L.arena_floats = m->d.arena_floats; L.in_off = m->d.in_off; L.out_off = m->d.out_off; L.post_off = m->d.post_off;
    L.Nt = m->d.Nt; L.Nr = m->d.Nr; L.channels = m->d.channels; L.max_w_len = m->d.max_w_len;
    L.sigmas = m->d_sigmas; L.n_sigmas = m->d.n_sigmas;
    L.gpark = m->d_park; L.park_floats = m->d.park_floats;
    L.gws = m->d_gws; L.stage_weights = m->stage ? 1 : 0; L.debug_stop = -1;
    L.prof = m->d_prof;
    L.dbg = env_int("SBC_DBG", 0);
}

static int launch(SbcModel* m, SbcLaunch& L, cudaStream_t st) {
    SBC_CUDA(cudaSetDevice(m->device));
    const int slots = m->num_sms * m->ctas_per_sm;
    const int grid = L.B < slots ? L.B : slots;
    size_t smem = m->smem_bytes;
    if (L.debug_stop >= 0) L.stage_weights = 0;   // debug runs read parameters straight from global memory
    const bool instr = L.prof != nullptr || L.debug_stop >= 0 || L.dbg != 0;
#define SBC_LAUNCH(A, X, I) sbc_ald_kernel<A, X, I><<<grid, SBC_NTHREADS, smem, st>>>(L)
    if (m->arena_in_smem) {
        if (m->x3) { if (instr) SBC_LAUNCH(true, true, true); else SBC_LAUNCH(true, true, false); }
        else { if (instr) SBC_LAUNCH(true, false, true); else SBC_LAUNCH(true, false, false); }
    } else {
        if (m->x3) { if (instr) SBC_LAUNCH(false, true, true); else SBC_LAUNCH(false, true, false); }
        else { if (instr) SBC_LAUNCH(false, false, true); else SBC_LAUNCH(false, false, false); }
    }
#undef SBC_LAUNCH
    SBC_CUDA(cudaGetLastError());
    m->launches++;
    return SBC_OK;
}

extern "C" int sbc_forward(void* handle, const float* x, const int64_t x_strides[4], const int64_t* labels,
                           float* out, int32_t B, void* stream) {
    if (!handle || !x || !x_strides || !labels || !out) return sbc_fail(SBC_E_ARG, "sbc_forward: null argument");
    if (B < 0) return sbc_fail(SBC_E_ARG, "sbc_forward: negative batch");
    if (B == 0) return SBC_OK;
    SbcModel* m = (SbcModel*)handle;
    if (m->engine == 2) {
        SBC_CUDA(cudaSetDevice(m->e2->device));
        Sbc2Launch L2;
        int grid = 0;
        const std::string err = sbc2_prepare(m->e2, B, sbc2_pick_S(m->e2, B), true, L2, grid);
        if (!err.empty()) return sbc_fail(SBC_E_CUDA, "sbc_forward: %s", err.c_str());
        L2.mode = 0; L2.fx = x; L2.labels = (const long long*)labels; L2.fout = out;
        for (int i = 0; i < 4; i++) L2.fxs[i] = x_strides[i];
        SBC_CUDA(sbc2_launch(m->e2, L2, grid, (cudaStream_t)stream));
        return SBC_OK;
    }
    SbcLaunch L;
    fill_common(m, L);
    L.mode = 0; L.B = B; L.fx = x; L.labels = (const long long*)labels; L.fout = out;
    for (int i = 0; i < 4; i++) L.fxs[i] = x_strides[i];
    return launch(m, L, (cudaStream_t)stream);
}

extern "C" int sbc_dsm_loss(void* handle, const float* samples, const int64_t* labels, const float* z, float anneal_power,
                            float* loss_out, int32_t B, void* stream) {
    if (!handle || !samples || !labels || !z || !loss_out) return sbc_fail(SBC_E_ARG, "sbc_dsm_loss: null argument");
    if (B < 0) return sbc_fail(SBC_E_ARG, "sbc_dsm_loss: negative batch");
    if (B == 0) return SBC_OK;
    SbcModel* m = (SbcModel*)handle;
    if (m->engine == 2) return sbc_fail(SBC_E_UNSUPPORTED, "sbc_dsm_loss: engine-1 models only (precision tf32x3 / tf32)");
    SbcLaunch L;
    fill_common(m, L);
    L.mode = 2; L.B = B; L.fx = samples; L.labels = (const long long*)labels; L.dsm_z = z; L.dsm_out = loss_out;
    L.anneal_power = anneal_power;
    L.fxs[0] = (long long)m->d.channels * m->d.Nt * m->d.Nr; L.fxs[1] = (long long)m->d.Nt * m->d.Nr;
    L.fxs[2] = m->d.Nr; L.fxs[3] = 1;
    return launch(m, L, (cudaStream_t)stream);
}

static int check_ald(const SbcModel* m, const sbc_ald_args* a) {
    if (!a) return sbc_fail(SBC_E_ARG, "sbc_ald_run: null args");
    if (a->B < 0) return sbc_fail(SBC_E_ARG, "sbc_ald_run: negative batch");
    const int mNt = m->engine == 2 ? m->e2->Nt : m->d.Nt, mNr = m->engine == 2 ? m->e2->Nr : m->d.Nr;
    const int mns = m->engine == 2 ? m->e2->n_sigmas : m->d.n_sigmas;
    if (a->Nt != mNt || a->Nr != mNr)
        return sbc_fail(SBC_E_ARG, "sbc_ald_run: model packed for %dx%d, got Nt=%d Nr=%d", mNt, mNr, a->Nt, a->Nr);
    if (a->Np <= 0 || a->Np > a->Nt) return sbc_fail(SBC_E_ARG, "sbc_ald_run: need 0 < Np <= Nt (Np=%d)", a->Np);
    if (a->level_begin < 0 || a->level_end > mns || a->level_begin > a->level_end)
        return sbc_fail(SBC_E_ARG, "sbc_ald_run: level range [%d,%d) outside [0,%d]", a->level_begin, a->level_end, mns);
    if (a->steps_each <= 0) return sbc_fail(SBC_E_ARG, "sbc_ald_run: steps_each must be positive");
    if (a->B > 0 && (!a->P || !a->Y || !a->X || !a->noise_var || !a->alpha_step || !a->beta))
        return sbc_fail(SBC_E_ARG, "sbc_ald_run: null array");
    if (!(a->sigma_end > 0.)) return sbc_fail(SBC_E_ARG, "sbc_ald_run: sigma_end must be positive");
    return SBC_OK;
}

static void fill_ald(SbcLaunch& L, const sbc_ald_args* a) {
    L.mode = 1; L.B = a->B; L.Np = a->Np;
    L.level_begin = a->level_begin; L.level_end = a->level_end; L.steps_each = a->steps_each;
    L.P = (const float*)a->P; L.Y = (const float*)a->Y; L.X = (float*)a->X; L.Hor = (const float*)a->H_oracle;
    L.noise_var = a->noise_var; L.alpha_step = a->alpha_step; L.beta = a->beta; L.sigma_end = a->sigma_end;
    L.nmse_log = a->nmse_log; L.seed = a->seed; L.sample_ids = (const unsigned long long*)a->sample_ids;
    L.ext_noise = (const float*)a->ext_noise;
    L.dc_boost = a->dc_boost; L.stop_step = a->stop_step;
}

extern "C" int sbc_ald_run(void* handle, const sbc_ald_args* a, void* stream) {
    if (!handle) return sbc_fail(SBC_E_ARG, "sbc_ald_run: null handle");
    SbcModel* m = (SbcModel*)handle;
    int rc = check_ald(m, a);
    if (rc) return rc;
    if (a->B == 0 || a->level_begin == a->level_end) return SBC_OK;
    if (m->engine == 2) {
        SBC_CUDA(cudaSetDevice(m->e2->device));
        Sbc2Launch L2;
        int grid = 0;
        const std::string err = sbc2_prepare(m->e2, a->B, sbc2_pick_S(m->e2, a->B), true, L2, grid);
        if (!err.empty()) return sbc_fail(SBC_E_CUDA, "sbc_ald_run: %s", err.c_str());
        L2.mode = 1; L2.Np = a->Np;
        L2.level_begin = a->level_begin; L2.level_end = a->level_end; L2.steps_each = a->steps_each;
        L2.P = (const float*)a->P; L2.Y = (const float*)a->Y; L2.X = (float*)a->X; L2.Hor = (const float*)a->H_oracle;
        L2.noise_var = a->noise_var; L2.alpha_step = a->alpha_step; L2.beta = a->beta; L2.sigma_end = a->sigma_end;
        L2.nmse_log = a->nmse_log; L2.seed = a->seed; L2.sample_ids = (const unsigned long long*)a->sample_ids;
        L2.ext_noise = (const float*)a->ext_noise; L2.dc_boost = a->dc_boost; L2.stop_step = a->stop_step;
        SBC_CUDA(sbc2_launch(m->e2, L2, grid, (cudaStream_t)stream));
        return SBC_OK;
    }
    SbcLaunch L;
    fill_common(m, L);
    fill_ald(L, a);
    return launch(m, L, (cudaStream_t)stream);
}

// ---- host-buffer variants ------------------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { if (p) cudaFree(p); }
    int alloc(size_t n) { return cudaMalloc(&p, n ? n : 1) == cudaSuccess ? 0 : -1; }
};

extern "C" int sbc_forward_host(void* handle, const float* x, const int64_t* labels, float* out, int32_t B) {
    if (!handle || !x || !labels || !out) return sbc_fail(SBC_E_ARG, "sbc_forward_host: null argument");
    if (B <= 0) return B == 0 ? SBC_OK : sbc_fail(SBC_E_ARG, "sbc_forward_host: negative batch");
    SbcModel* m = (SbcModel*)handle;
    if (m->engine == 2) { m->d.channels = m->e2->channels; m->d.Nt = m->e2->Nt; m->d.Nr = m->e2->Nr; m->device = m->e2->device; }
    SBC_CUDA(cudaSetDevice(m->device));
    const size_t n = (size_t)B * m->d.channels * m->d.Nt * m->d.Nr;
    struct { void* p; } dx{m->host_ws(12, n * 4)}, dl{m->host_ws(13, (size_t)B * 8)}, dout{m->host_ws(14, n * 4)};
    if (!dx.p || !dl.p || !dout.p) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
    SBC_CUDA(cudaMemcpy(dx.p, x, n * 4, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(dl.p, labels, (size_t)B * 8, cudaMemcpyHostToDevice));
    const int64_t st[4] = {(int64_t)m->d.channels * m->d.Nt * m->d.Nr, (int64_t)m->d.Nt * m->d.Nr, m->d.Nr, 1};
    int rc = sbc_forward(handle, (const float*)dx.p, st, (const int64_t*)dl.p, (float*)dout.p, B, nullptr);
    if (rc) return rc;
    SBC_CUDA(cudaMemcpy(out, dout.p, n * 4, cudaMemcpyDeviceToHost));
    return SBC_OK;
}

extern "C" int sbc_ald_run_host(void* handle, const sbc_ald_args* a) {
    if (!handle) return sbc_fail(SBC_E_ARG, "sbc_ald_run_host: null handle");
    SbcModel* m = (SbcModel*)handle;
    int rc = check_ald(m, a);
    if (rc) return rc;
    if (a->B == 0 || a->level_begin == a->level_end) return SBC_OK;
    SBC_CUDA(cudaSetDevice(m->engine == 2 ? m->e2->device : m->device));
    const size_t B = a->B, ne = (size_t)a->Nt * a->Nr, steps = (size_t)(a->level_end - a->level_begin) * a->steps_each;
    const size_t nP = B * a->Np * a->Nt * 8, nY = B * a->Np * a->Nr * 8, nX = B * ne * 8;
    // device buffers come from the handle's grow-only workspace: no cudaMalloc / cudaFree per call
    struct Slot { void* p = nullptr; SbcModel* m; int id; int alloc(size_t n) { p = m->host_ws(id, n ? n : 1); return p ? 0 : -1; } };
    Slot dP{nullptr, m, 0}, dY{nullptr, m, 1}, dX{nullptr, m, 2}, dH{nullptr, m, 3}, dnv{nullptr, m, 4}, dal{nullptr, m, 5}, dbe{nullptr, m, 6},
        dlog{nullptr, m, 7}, dids{nullptr, m, 8}, dn{nullptr, m, 9}, ddb{nullptr, m, 10}, dst{nullptr, m, 11};
    if (dP.alloc(nP) || dY.alloc(nY) || dX.alloc(nX) || dnv.alloc(B * 4) || dal.alloc(B * 4) || dbe.alloc(B * 4))
        return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
    SBC_CUDA(cudaMemcpy(dP.p, a->P, nP, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(dY.p, a->Y, nY, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(dX.p, a->X, nX, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(dnv.p, a->noise_var, B * 4, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(dal.p, a->alpha_step, B * 4, cudaMemcpyHostToDevice));
    SBC_CUDA(cudaMemcpy(dbe.p, a->beta, B * 4, cudaMemcpyHostToDevice));
    sbc_ald_args d = *a;
    d.P = dP.p; d.Y = dY.p; d.X = dX.p; d.noise_var = (const float*)dnv.p; d.alpha_step = (const float*)dal.p;
    d.beta = (const float*)dbe.p;
    if (a->H_oracle) {
        if (dH.alloc(nX)) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
        SBC_CUDA(cudaMemcpy(dH.p, a->H_oracle, nX, cudaMemcpyHostToDevice));
        d.H_oracle = dH.p;
    }
    if (a->nmse_log) {
        if (dlog.alloc(steps * B * 4)) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
        d.nmse_log = (float*)dlog.p;
    }
    if (a->sample_ids) {
        if (dids.alloc(B * 8)) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
        SBC_CUDA(cudaMemcpy(dids.p, a->sample_ids, B * 8, cudaMemcpyHostToDevice));
        d.sample_ids = (const uint64_t*)dids.p;
    }
    if (a->ext_noise) {
        if (dn.alloc(steps * nX)) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
        SBC_CUDA(cudaMemcpy(dn.p, a->ext_noise, steps * nX, cudaMemcpyHostToDevice));
        d.ext_noise = dn.p;
    }
    if (a->dc_boost) {
        if (ddb.alloc(B * 4)) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
        SBC_CUDA(cudaMemcpy(ddb.p, a->dc_boost, B * 4, cudaMemcpyHostToDevice));
        d.dc_boost = (const float*)ddb.p;
    }
    if (a->stop_step) {
        if (dst.alloc(B * 4)) return sbc_fail(SBC_E_NOMEM, "cudaMalloc failed");
        SBC_CUDA(cudaMemcpy(dst.p, a->stop_step, B * 4, cudaMemcpyHostToDevice));
        d.stop_step = (const int32_t*)dst.p;
        if (a->nmse_log) SBC_CUDA(cudaMemcpy(dlog.p, a->nmse_log, steps * B * 4, cudaMemcpyHostToDevice));   // untouched rows
    }
    rc = sbc_ald_run(handle, &d, nullptr);
    if (rc) return rc;
    SBC_CUDA(cudaMemcpy(a->X, dX.p, nX, cudaMemcpyDeviceToHost));
    if (a->nmse_log) SBC_CUDA(cudaMemcpy(a->nmse_log, dlog.p, steps * B * 4, cudaMemcpyDeviceToHost));
    SBC_CUDA(cudaDeviceSynchronize());
    return SBC_OK;
}

extern "C" int sbc_set_profile_buffer(void* handle, int64_t* dev_stamps) {
    if (!handle) return sbc_fail(SBC_E_ARG, "sbc_set_profile_buffer: null handle");
    ((SbcModel*)handle)->d_prof = (long long*)dev_stamps;
    if (((SbcModel*)handle)->engine == 2) ((SbcModel*)handle)->e2->d_prof = (long long*)dev_stamps;
    return SBC_OK;
}

extern "C" int sbc_debug_arena(void* handle, const float* x, int32_t stop_op, float* arena_out, void* stream) {
    if (!handle || !x || !arena_out) return sbc_fail(SBC_E_ARG, "sbc_debug_arena: null argument");
    SbcModel* m = (SbcModel*)handle;
    if (m->engine == 2) return sbc_fail(SBC_E_UNSUPPORTED, "sbc_debug_arena: engine 2 models use sbc_debug_run");
    if (stop_op < 0 || stop_op > m->d.n_ops) return sbc_fail(SBC_E_ARG, "sbc_debug_arena: stop_op out of range");
    static const long long zero_label = 0;
    (void)zero_label;
    SbcLaunch L;
    fill_common(m, L);
    L.mode = 0; L.B = 1; L.fx = x; L.labels = nullptr; L.fout = nullptr;
    L.fxs[0] = (long long)m->d.channels * m->d.Nt * m->d.Nr; L.fxs[1] = (long long)m->d.Nt * m->d.Nr;
    L.fxs[2] = m->d.Nr; L.fxs[3] = 1;
    L.debug_stop = stop_op; L.debug_out = arena_out;
    return launch(m, L, (cudaStream_t)stream);
}


// ---- engine 2: self-contained model creation from a plain state dict + debug views ---------------------------
extern "C" int sbc_model_create_from_state(const sbc_state_entry* entries, int32_t n_entries, int32_t ngf, int32_t Nt,
                                           int32_t Nr, int32_t channels, int device, void** handle_out) {
    if (!entries || n_entries <= 0 || !handle_out) return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state: null / empty argument");
    int ndev = 0;
    SBC_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state: no CUDA device %d", device);
    SBC_CUDA(cudaSetDevice(device));
    sbc2::StateDict sd;
    for (int i = 0; i < n_entries; i++) {
        const sbc_state_entry& e = entries[i];
        if (!e.name || !e.data || e.ndim < 0 || e.ndim > 8 || (e.ndim > 0 && !e.shape))
            return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state: bad entry %d", i);
        sbc2::TensorArg t;
        t.data = e.data;
        for (int k = 0; k < e.ndim; k++) t.shape.push_back(e.shape[k]);
        sd[e.name] = t;
    }
    if (ngf <= 0 || Nt <= 0 || Nr <= 0) return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state: ngf, Nt, Nr must be positive");
    std::unique_ptr<SbcModel> mh(new SbcModel());
    mh->engine = 2;
    mh->e2 = new Sbc2Model();
    mh->device = device;
    const std::string err = sbc2_create(mh->e2, sd, ngf, Nt, Nr, channels, device);
    if (!err.empty()) return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state: %s", err.c_str());
    *handle_out = mh.release();
    return SBC_OK;
}

extern "C" int sbc_debug_plan(void* handle, int32_t S, int32_t reuse, sbc_tensor_info* out, int32_t cap, int32_t* n_out,
                              int64_t* arena_bytes, int32_t* geo_out /* [4][14] */) {
    if (!handle) return sbc_fail(SBC_E_ARG, "sbc_debug_plan: null handle");
    SbcModel* m = (SbcModel*)handle;
    if (m->engine != 2) return sbc_fail(SBC_E_UNSUPPORTED, "sbc_debug_plan: engine 2 only");
    SBC_CUDA(cudaSetDevice(m->e2->device));
    if (S < 1 || S > SBC2_MAXS) return sbc_fail(SBC_E_ARG, "sbc_debug_plan: group size must be in 1..%d", SBC2_MAXS);
    std::string err;
    Sbc2PlanDev* pd = sbc2_plan(m->e2, S, reuse != 0, err);
    if (!pd) return sbc_fail(SBC_E_CUDA, "sbc_debug_plan: %s", err.c_str());
    static_assert(sizeof(sbc_tensor_info) == sizeof(sbc2::TensorInfo), "tensor info layout");
    const int n = (int)pd->plan.tensors.size();
    if (n_out) *n_out = n;
    if (out) memcpy(out, pd->plan.tensors.data(), sizeof(sbc2::TensorInfo) * (size_t)(n < cap ? n : cap));
    if (arena_bytes) *arena_bytes = pd->plan.arena_bytes;
    if (geo_out) memcpy(geo_out, pd->plan.geo, sizeof(sbc2::Geo) * sbc2::MAX_LEVELS);
    return SBC_OK;
}

extern "C" int sbc_debug_run(void* handle, const float* x, int32_t S, int32_t reuse, void* arena_out, void* stream) {
    if (!handle || !x) return sbc_fail(SBC_E_ARG, "sbc_debug_run: null argument");
    SbcModel* m = (SbcModel*)handle;
    if (m->engine != 2) return sbc_fail(SBC_E_UNSUPPORTED, "sbc_debug_run: engine 2 only");
    Sbc2Model* e = m->e2;
    if (S < 1 || S > SBC2_MAXS) return sbc_fail(SBC_E_ARG, "sbc_debug_run: group size must be in 1..%d", SBC2_MAXS);
    SBC_CUDA(cudaSetDevice(e->device));
    Sbc2Launch L2;
    int grid = 0;
    const std::string err = sbc2_prepare(e, S, S, reuse != 0, L2, grid);     // one group of S samples
    if (!err.empty()) return sbc_fail(SBC_E_CUDA, "sbc_debug_run: %s", err.c_str());
    L2.mode = 0; L2.fx = x; L2.labels = nullptr; L2.fout = nullptr;
    L2.fxs[0] = (long long)e->channels * e->Nt * e->Nr; L2.fxs[1] = (long long)e->Nt * e->Nr; L2.fxs[2] = e->Nr; L2.fxs[3] = 1;
    float* scratch = nullptr;
    SBC_CUDA(cudaMalloc(&scratch, sizeof(float) * (size_t)S * e->channels * e->Nt * e->Nr));
    L2.fout = scratch;
    cudaError_t ce = sbc2_launch(e, L2, grid, (cudaStream_t)stream);
    if (ce == cudaSuccess && arena_out) ce = cudaMemcpyAsync(arena_out, e->d_gws, (size_t)L2.arena_bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize((cudaStream_t)stream);
    cudaFree(scratch);
    if (ce != cudaSuccess) return sbc_fail(SBC_E_CUDA, "sbc_debug_run: %s", cudaGetErrorString(ce));
    return SBC_OK;
}

// ---- engine 1 planned inside the library (csrc/sbc1_plan.h) -----------------------------------------------------
static int state_from_entries(const sbc_state_entry* entries, int32_t n_entries, sbc2::StateDict& sd, const char* who) {
    if (!entries || n_entries <= 0) return sbc_fail(SBC_E_ARG, "%s: null / empty state", who);
    for (int i = 0; i < n_entries; i++) {
        const sbc_state_entry& e = entries[i];
        if (!e.name || !e.data || e.ndim < 0 || e.ndim > 8 || (e.ndim > 0 && !e.shape)) return sbc_fail(SBC_E_ARG, "%s: bad entry %d", who, i);
        sbc2::TensorArg t;
        t.data = e.data;
        for (int k = 0; k < e.ndim; k++) t.shape.push_back(e.shape[k]);
        sd[e.name] = t;
    }
    return SBC_OK;
}

static void plan1_view(const sbc1::Plan& P, sbc_plan1_view* v) {
    v->op_table = reinterpret_cast<const int32_t*>(P.ops.data()); v->n_ops = (int32_t)P.ops.size();
    v->geo_table = reinterpret_cast<const int32_t*>(P.geos.data()); v->n_geo = (int32_t)P.geos.size();
    v->blob = P.blob.data(); v->blob_floats = (int64_t)P.blob.size();
    v->arena_floats = P.arena_floats; v->in_off = P.in_off; v->out_off = P.out_off; v->post_off = P.post_off;
    v->max_w_len = P.max_w_len; v->park_floats = P.park_floats; v->nthreads = P.nthreads; v->conv_flops = P.conv_flops;
}

extern "C" int sbc_plan1_build(const sbc_state_entry* entries, int32_t n_entries, int32_t ngf, int32_t Nt, int32_t Nr,
                               int32_t channels, int32_t nthreads, int32_t precision, int32_t park, void** plan_out,
                               sbc_plan1_view* view) {
    if (!plan_out || !view) return sbc_fail(SBC_E_ARG, "sbc_plan1_build: null argument");
    if (precision != SBC_PREC_TF32X3 && precision != SBC_PREC_TF32) return sbc_fail(SBC_E_ARG, "sbc_plan1_build: engine-1 precisions are SBC_PREC_TF32X3 / SBC_PREC_TF32");
    if (nthreads < 32 || nthreads > 1024 || (nthreads & (nthreads - 1))) return sbc_fail(SBC_E_ARG, "sbc_plan1_build: nthreads must be a power of two in [32, 1024]");
    sbc2::StateDict sd;
    int rc = state_from_entries(entries, n_entries, sd, "sbc_plan1_build");
    if (rc) return rc;
    try {
        std::unique_ptr<sbc1::Plan> P(new sbc1::Plan(sbc1::build_auto(sd, ngf, Nt, Nr, channels, nthreads, precision == SBC_PREC_TF32X3, park)));
        plan1_view(*P, view);
        *plan_out = P.release();
    } catch (const std::exception& ex) {
        return sbc_fail(SBC_E_ARG, "sbc_plan1_build: %s", ex.what());
    }
    return SBC_OK;
}

extern "C" int sbc_plan1_free(void* plan) {
    delete (sbc1::Plan*)plan;
    return SBC_OK;
}

extern "C" int sbc_model_create_from_state_ex(const sbc_state_entry* entries, int32_t n_entries, int32_t ngf, int32_t Nt,
                                              int32_t Nr, int32_t channels, int device, int32_t precision, void** handle_out) {
    if (precision == SBC_PREC_FP16X2) return sbc_model_create_from_state(entries, n_entries, ngf, Nt, Nr, channels, device, handle_out);
    if (!handle_out) return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state_ex: null argument");
    void* plan = nullptr;
    sbc_plan1_view v;
    int rc = sbc_plan1_build(entries, n_entries, ngf, Nt, Nr, channels, SBC_NTHREADS, precision, -1, &plan, &v);
    if (rc) return rc;
    std::unique_ptr<sbc1::Plan> P((sbc1::Plan*)plan);
    const float* sigmas = nullptr;
    int n_sigmas = 0;
    for (int i = 0; i < n_entries; i++)
        if (strcmp(entries[i].name, "sigmas") == 0) {
            sigmas = entries[i].data;
            n_sigmas = entries[i].ndim > 0 ? (int)entries[i].shape[0] : 1;
        }
    if (!sigmas || n_sigmas <= 0) return sbc_fail(SBC_E_ARG, "sbc_model_create_from_state_ex: the state has no 'sigmas' (noise schedule)");
    int32_t geo[SBC_MAX_GEO][8];
    memset(geo, 0, sizeof geo);
    memcpy(geo, v.geo_table, sizeof(SbcGeo) * (size_t)v.n_geo);
    sbc_model_desc d;
    memset(&d, 0, sizeof d);
    d.ngf = ngf; d.Nt = Nt; d.Nr = Nr; d.channels = channels;
    d.op_table = v.op_table; d.n_ops = v.n_ops; d.geo_table = &geo[0][0]; d.n_geo = v.n_geo;
    d.blob = v.blob; d.blob_floats = v.blob_floats; d.arena_floats = v.arena_floats;
    d.in_off = v.in_off; d.out_off = v.out_off; d.post_off = v.post_off; d.max_w_len = v.max_w_len;
    d.sigmas = sigmas; d.n_sigmas = n_sigmas; d.conv_flops = v.conv_flops; d.nthreads = v.nthreads; d.park_floats = v.park_floats;
    return sbc_model_create(&d, device, handle_out);
}

extern "C" const char* sbc_op_name(void* handle, int32_t i) {
    if (!handle) return "";
    SbcModel* m = (SbcModel*)handle;
    if (m->engine != 2 || i < 0 || i >= m->e2->builder->n_ops()) return "";
    return m->e2->builder->op_names()[(size_t)i].c_str();
}

extern "C" int sbc_op_kind(void* handle, int32_t i) {
    if (!handle) return -1;
    SbcModel* m = (SbcModel*)handle;
    if (m->engine != 2 || i < 0 || i >= m->e2->builder->n_ops()) return -1;
    std::string err;
    Sbc2PlanDev* pd = sbc2_plan(m->e2, 1, true, err);
    return pd ? pd->plan.ops[(size_t)i].kind : -1;
}
