// sbc2_host.cuh -- host side of engine 2 (tcgen05): model object, per-group-size plan cache, launch.
// Included by sbc_api.cu, which exposes it through the C ABI of include/sbc.h.
#pragma once
#include <cuda_runtime.h>

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "sbc2_kernel.cuh"
#include "sbc2_plan.h"

struct Sbc2PlanDev {
    sbc2::Plan plan;
    sbc2::Op* d_ops = nullptr;
    int32_t* d_pix[sbc2::MAX_LEVELS] = {nullptr, nullptr, nullptr, nullptr};
    ~Sbc2PlanDev() {
        cudaFree(d_ops);
        for (auto p : d_pix) cudaFree(p);
    }
};

struct Sbc2Model {
    int device = 0, num_sms = 0, smem_optin = 0;
    int ngf = 0, Nt = 0, Nr = 0, channels = 2;
    std::unique_ptr<sbc2::Builder> builder;
    // the builder keeps a reference to the state dict: own copies of the tensors
    sbc2::StateDict sd;
    std::vector<std::vector<float>> storage;
    uint8_t* d_blob = nullptr;
    float* d_sigmas = nullptr;
    int n_sigmas = 0;
    std::map<std::pair<int, int>, std::unique_ptr<Sbc2PlanDev>> plans;   // (S, reuse) -> device plan
    uint8_t* d_gws = nullptr;
    size_t gws_bytes = 0;
    long long* d_prof = nullptr;
    long long launches = 0;
    int wmax = 0, stage_bytes = 0;
    size_t smem_bytes = 0;
    int ctas_per_sm = 2;
    int last_S = 0, last_grid = 0, last_reuse = 1;
    ~Sbc2Model() {
        cudaFree(d_blob); cudaFree(d_sigmas); cudaFree(d_gws);
    }
};

static const int SBC2_STAGE_CAP = 32 * 1024;

static int sbc2_env_int(const char* name, int dflt) {
    const char* s = getenv(name);
    return s ? atoi(s) : dflt;
}

// returns an error string ("" = ok)
static std::string sbc2_create(Sbc2Model* m, const sbc2::StateDict& sd_in, int ngf, int Nt, int Nr, int channels, int device) {
    m->device = device; m->ngf = ngf; m->Nt = Nt; m->Nr = Nr; m->channels = channels;
    for (auto& kv : sd_in) {
        int64_t n = 1;
        for (auto d : kv.second.shape) n *= d;
        m->storage.emplace_back(kv.second.data, kv.second.data + n);
        m->sd[kv.first] = sbc2::TensorArg{m->storage.back().data(), kv.second.shape};
    }
    auto it = m->sd.find("sigmas");
    if (it == m->sd.end()) return "state dict has no 'sigmas' entry";
    if (it->second.shape.size() != 1 || it->second.shape[0] < 1) return "'sigmas' must be a non-empty 1-D tensor";
    m->n_sigmas = (int)it->second.shape[0];
    try {
        m->builder.reset(new sbc2::Builder(m->sd, ngf, Nt, Nr, channels, SBC2_STAGE_CAP));
    } catch (const std::exception& e) {
        return std::string("planner: ") + e.what();
    }
    if (cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return "cudaDeviceGetAttribute failed";
    cudaDeviceGetAttribute(&m->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device);
    const sbc2::Builder& b = *m->builder;
    if (b.max_stage > SBC2_STAGE_CAP) return "a conv input is too wide for the staging ring of engine 2 (" + std::to_string(b.max_stage) + " bytes per tile)";
    m->wmax = (b.max_seg + 127) / 128 * 128;
    m->stage_bytes = SBC2_STAGE_CAP;      // the ring the planner sized every op's stage count against (Op::nstage)
    m->smem_bytes = 2 * (size_t)m->wmax + (size_t)m->stage_bytes + SBC2_NBARS * 8 + SBC2_PART_FLOATS * 4;
    cudaFuncAttributes fa{};
    if (cudaFuncGetAttributes(&fa, sbc2_ald_kernel) != cudaSuccess) return std::string("cudaFuncGetAttributes: ") + cudaGetErrorString(cudaGetLastError());
    if (m->smem_bytes + fa.sharedSizeBytes > (size_t)m->smem_optin)
        return "model needs " + std::to_string(m->smem_bytes) + " bytes of shared memory per CTA (weights of the widest conv twice + staging): too wide for engine 2";
    if (cudaFuncSetAttribute(sbc2_ald_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes) != cudaSuccess)
        return std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(cudaGetLastError());
    cudaFuncSetAttribute(sbc2_ald_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    // Resident CTAs per SM.  cudaOccupancyMaxActiveBlocksPerMultiprocessor answers 1 for ANY kernel that contains a
    // tcgen05.alloc, yet the hardware co-schedules two such CTAs when each allocates <= 256 TMEM columns (measured with
    // tools/occ_probe.cu: 296 CTAs of 192 threads run in one wave).  So: shared memory and registers decide, capped at 2.
    int occ = 0, smpm = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, sbc2_ald_kernel, SBC2_NTHR, m->smem_bytes);
    cudaDeviceGetAttribute(&smpm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, device);
    cudaFuncGetAttributes(&fa, sbc2_ald_kernel);
    const size_t per_cta_smem = m->smem_bytes + fa.sharedSizeBytes + 1024;
    const int regs_per_cta = ((fa.numRegs + 7) / 8 * 8) * 32 * (SBC2_NTHR / 32);
    const int by_smem = (int)((size_t)smpm / per_cta_smem), by_regs = 65536 / regs_per_cta;
    int mine = by_smem < by_regs ? by_smem : by_regs;
    mine = mine < 1 ? 1 : (mine > 2 ? 2 : mine);
    m->ctas_per_sm = mine;
    if (sbc2_env_int("SBC2_VERBOSE", 0))
        fprintf(stderr, "[sbc2] regs %d static smem %zu dyn smem %zu (wmax %d stage %d) smem/SM %d: occupancy API %d, used %d\n",
                fa.numRegs, fa.sharedSizeBytes, m->smem_bytes, m->wmax, m->stage_bytes, smpm, occ, mine);
    if (sbc2_env_int("SBC2_CTAS", 0) > 0) m->ctas_per_sm = sbc2_env_int("SBC2_CTAS", 0);
    if (cudaMalloc(&m->d_blob, b.blob.size() ? b.blob.size() : 16) != cudaSuccess) return "cudaMalloc(blob) failed";
    cudaMemcpy(m->d_blob, b.blob.data(), b.blob.size(), cudaMemcpyHostToDevice);
    if (cudaMalloc(&m->d_sigmas, sizeof(float) * (size_t)m->n_sigmas) != cudaSuccess) return "cudaMalloc(sigmas) failed";
    cudaMemcpy(m->d_sigmas, it->second.data, sizeof(float) * (size_t)m->n_sigmas, cudaMemcpyHostToDevice);
    return "";
}

static Sbc2PlanDev* sbc2_plan(Sbc2Model* m, int S, bool reuse, std::string& err) {
    auto key = std::make_pair(S, reuse ? 1 : 0);
    auto it = m->plans.find(key);
    if (it != m->plans.end()) return it->second.get();
    std::unique_ptr<Sbc2PlanDev> pd(new Sbc2PlanDev());
    try {
        pd->plan = m->builder->layout(S, reuse);
    } catch (const std::exception& e) {
        err = std::string("planner: ") + e.what();
        return nullptr;
    }
    const auto& ops = pd->plan.ops;
    if (cudaMalloc(&pd->d_ops, sizeof(sbc2::Op) * ops.size()) != cudaSuccess) { err = "cudaMalloc(ops) failed"; return nullptr; }
    cudaMemcpy(pd->d_ops, ops.data(), sizeof(sbc2::Op) * ops.size(), cudaMemcpyHostToDevice);
    for (int l = 0; l < sbc2::MAX_LEVELS; l++) {
        const auto& pm = pd->plan.pix[l];
        if (cudaMalloc(&pd->d_pix[l], pm.size() * 4) != cudaSuccess) { err = "cudaMalloc(pix) failed"; return nullptr; }
        cudaMemcpy(pd->d_pix[l], pm.data(), pm.size() * 4, cudaMemcpyHostToDevice);
    }
    Sbc2PlanDev* r = pd.get();
    m->plans[key] = std::move(pd);
    return r;
}

// group size for a batch: one sample per CTA slot while the batch fits in one wave, then grow the groups
static int sbc2_pick_S(const Sbc2Model* m, int B) {
    const int forced = sbc2_env_int("SBC2_S", 0);
    if (forced > 0) return forced > SBC2_MAXS ? SBC2_MAXS : forced;
    const int slots = m->num_sms * m->ctas_per_sm;
    int S = (B + slots - 1) / slots;
    if (S < 1) S = 1;
    if (S > SBC2_MAXS) S = SBC2_MAXS;
    return S;
}

// fills the common part of the launch record and makes sure the workspace is large enough
static std::string sbc2_prepare(Sbc2Model* m, int B, int S, bool reuse, Sbc2Launch& L, int& grid) {
    std::string err;
    Sbc2PlanDev* pd = sbc2_plan(m, S, reuse, err);
    if (!pd) return err;
    const int n_groups = (B + S - 1) / S;
    const int slots = m->num_sms * m->ctas_per_sm;
    grid = n_groups < slots ? n_groups : slots;
    const size_t need = (size_t)grid * (size_t)pd->plan.arena_bytes;
    if (need > m->gws_bytes) {
        cudaFree(m->d_gws);
        m->d_gws = nullptr; m->gws_bytes = 0;
        if (cudaMalloc(&m->d_gws, need) != cudaSuccess) return "cudaMalloc(workspace, " + std::to_string(need) + " bytes) failed";
        m->gws_bytes = need;
    }
    memset(&L, 0, sizeof L);
    L.ops = pd->d_ops; L.n_ops = (int)pd->plan.ops.size(); L.blob = m->d_blob;
    L.gws = m->d_gws; L.arena_bytes = pd->plan.arena_bytes;
    L.S = S; L.B = B;
    L.x_off = pd->plan.x_off; L.out_off = pd->plan.out_off; L.post_off = pd->plan.post_off;
    L.wmax = m->wmax; L.stage_bytes = m->stage_bytes;
    L.Nt = m->Nt; L.Nr = m->Nr; L.channels = m->channels;
    L.sigmas = m->d_sigmas; L.n_sigmas = m->n_sigmas;
    L.prof = m->d_prof;
    L.dbg = sbc2_env_int("SBC2_DBG", 0);
    L.trace_op = sbc2_env_int("SBC2_TRACE_OP", -1);
    m->last_S = S; m->last_grid = grid; m->last_reuse = reuse ? 1 : 0;
    if ((int)pd->plan.ops.size() > SBC2_MAX_OPS) return "layer program too long for the constant-memory op table";
    return "";
}

static cudaError_t sbc2_launch(Sbc2Model* m, const Sbc2Launch& L, int grid, cudaStream_t st) {
    // the layer program travels through constant memory: stream-ordered upload in front of every launch (tens of KB,
    // negligible against a launch that runs the whole schedule).  Launches of engine-2 models on DIFFERENT streams of
    // one device must therefore not overlap.
    Sbc2PlanDev* pd = m->plans[std::make_pair(L.S, m->last_reuse)].get();
    const auto& ops = pd->plan.ops;
    std::vector<int4> cc(ops.size());
    const bool in_const = m->builder->all_mma.size() <= (size_t)SBC2_MAX_MMA;
    for (size_t i = 0; i < ops.size(); i++) {
        const sbc2::Op& o = ops[i];
        if (o.kind != sbc2::K_CONV) { cc[i] = make_int4(-1, 0, 0, 1 | (1 << 8) | (1 << 16)); continue; }
        // Split-K partial accumulators (SBC2_NPART=2|4): independent accumulation chains, summed by the epilogue.  Measured
        // on B200 (tools/umma_rate.cu): a small tcgen05.mma costs ~54 clk whatever N <= 64, the layout or the accumulator
        // it targets -- the chains do not serialise on the TMEM read-modify-write, so the default is ONE accumulator.
        int npart = 1;
        const int want = sbc2_env_int("SBC2_NPART", 1);
        while (npart * 2 <= want && npart * 2 <= 4 && npart * 2 <= o.n_mma) npart *= 2;
        int span = (npart * o.N + 63) / 64;
        while (span > (o.T > 1 ? 2 : 4) && npart > 1) { npart /= 2; span = (npart * o.N + 63) / 64; }
        span = span <= 1 ? 1 : (span <= 2 ? 2 : 4);
        cc[i] = make_int4(in_const ? o.mma_idx : -1, o.n_mma, o.idesc, o.nstage | (npart << 8) | (span << 16));
    }
    cudaError_t ce = cudaMemcpyToSymbolAsync(sbc2_c_geo, pd->plan.geo, sizeof(sbc2::Geo) * sbc2::MAX_LEVELS, 0, cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) return ce;
    ce = cudaMemcpyToSymbolAsync(sbc2_c_conv, cc.data(), sizeof(int4) * cc.size(), 0, cudaMemcpyHostToDevice, st);
    if (ce != cudaSuccess) return ce;
    if (in_const) {
        ce = cudaMemcpyToSymbolAsync(sbc2_c_mma, m->builder->all_mma.data(), sizeof(sbc2::MmaEntry) * m->builder->all_mma.size(), 0,
                                     cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) return ce;
    }
    sbc2_ald_kernel<<<grid, SBC2_NTHR, m->smem_bytes, st>>>(L);
    m->launches++;
    return cudaGetLastError();
}
