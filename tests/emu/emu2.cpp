// emu2.cpp -- CPU emulation of an engine-2 plan (score_based_channels_b200/csrc/sbc2_plan.h).  TEST-ONLY, g++.
//
// Executes the op list on the planned byte layout of one group arena: every tcgen05.mma of a conv is replayed from
// its packed instruction record exactly as the hardware decodes it (staged window copy, K-major no-swizzle core
// matrices, LBO / SBO strides, fp16 operands, fp32 accumulation), the epilogue / norm / pool / bilinear ops follow the
// layout contract of sbc2_plan.h.  It validates the planner, the weight packing and the address arithmetic against
// the oracle without a GPU; the CUDA kernel (sbc2_kernel.cuh) is then compared tensor by tensor with this arena.
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <vector>

#include "../../score_based_channels_b200/csrc/sbc2_plan.h"

using namespace sbc2;

static float elu_ref(float v) { return v > 0.f ? v : expm1f(v); }

struct Emu {
    std::unique_ptr<Builder> b;
    StateDict sd;
    std::vector<std::vector<float>> storage;
};

static void split8(const float* v, uint16_t* hi, uint16_t* lo) {
    for (int k = 0; k < 8; k++) {
        float x = fminf(fmaxf(v[k], -65000.f), 65000.f);
        hi[k] = f32_to_f16(x);
        lo[k] = f32_to_f16(x - f16_to_f32(hi[k]));
    }
}
static void store_sp16(uint8_t* base, int slot, int oct, int q, const float* v) {
    uint8_t* p = base + (size_t)(2 * oct) * slot + (size_t)q * 16;
    split8(v, reinterpret_cast<uint16_t*>(p), reinterpret_cast<uint16_t*>(p + slot));
}
static void load_f32x8(const uint8_t* base, int slot, int oct, int q, float* v) {
    const uint8_t* p = base + (size_t)(2 * oct) * slot + (size_t)q * 16;
    memcpy(v, p, 16);
    memcpy(v + 4, p + slot, 16);
}
static void store_f32x8(uint8_t* base, int slot, int oct, int q, const float* v) {
    uint8_t* p = base + (size_t)(2 * oct) * slot + (size_t)q * 16;
    memcpy(p, v, 16);
    memcpy(p + slot, v + 4, 16);
}
static int qof(const Geo& G, int s, int y, int x) { return G.lead + s * G.pps + y * G.wp + x; }

static void run_conv(const Op& op, const Plan& P, const uint8_t* blob, uint8_t* arena) {
    const Geo& G = P.geo[op.gs];
    const uint8_t* seg = blob + op.w_off;
    const MmaEntry* list = reinterpret_cast<const MmaEntry*>(seg + op.mma_rel);
    const float* bias = op.bias_rel >= 0 ? reinterpret_cast<const float*>(seg + op.bias_rel) : nullptr;
    const int nsub = op.nsub0 + op.nsub1, N = op.N;
    std::vector<uint8_t> stage((size_t)nsub * op.sps);
    std::vector<float> D((size_t)TILE_M * N);
    const std::vector<int32_t>& pix = P.pix[op.gd];
    for (int t = 0; t < op.T; t++) {
        const size_t goff = (size_t)(G.lead + t * TILE_M - op.halo) * 16;
        for (int j = 0; j < op.nsub0; j++) memcpy(&stage[(size_t)j * op.sps], arena + op.src0 + goff + (size_t)j * G.slot, op.sps);
        for (int j = 0; j < op.nsub1; j++) memcpy(&stage[(size_t)(op.nsub0 + j) * op.sps], arena + op.src1 + goff + (size_t)j * G.slot, op.sps);
        std::fill(D.begin(), D.end(), 0.f);
        for (int i = 0; i < op.n_mma; i++) {
            const uint32_t a_off = desc_off(list[i].a_lo), a_lbo = desc_lbo(list[i].a_lo);
            const uint32_t b_off = desc_off(list[i].b_lo), nlbo = desc_lbo(list[i].b_lo);
            for (int m = 0; m < TILE_M; m++) {
                float a[16];
                for (int k = 0; k < 16; k++) {
                    const size_t off = a_off + (size_t)(k >> 3) * a_lbo + (size_t)(m >> 3) * 128 + (size_t)(m & 7) * 16 + (size_t)(k & 7) * 2;
                    uint16_t h;
                    memcpy(&h, &stage[off], 2);
                    a[k] = f16_to_f32(h);
                }
                for (int n = 0; n < N; n++) {
                    float acc = 0.f;
                    for (int k = 0; k < 16; k++) {
                        const size_t off = b_off + (size_t)(k >> 3) * nlbo + (size_t)(n >> 3) * 128 + (size_t)(n & 7) * 16 + (size_t)(k & 7) * 2;
                        uint16_t h;
                        memcpy(&h, seg + off, 2);
                        acc += a[k] * f16_to_f32(h);
                    }
                    D[(size_t)m * N + n] += acc;
                }
            }
        }
        for (int m = 0; m < TILE_M; m++) {
            const int q = G.lead + t * TILE_M + m;
            const int px = pix[(size_t)q];
            for (int c = 0; c < op.cout8 / 8; c++) {
                float v[8];
                for (int k = 0; k < 8; k++) {
                    v[k] = (D[(size_t)m * N + c * 8 + k] + D[(size_t)m * N + op.cout8 + c * 8 + k]) * op.unscale;
                    if (bias) v[k] += bias[c * 8 + k];
                }
                if (op.flags & F_COMPACT) {
                    if (px >= 0) {
                        float* o = reinterpret_cast<float*>(arena + op.dst32) + (size_t)px * 2;
                        o[0] = v[0]; o[1] = v[1];
                    }
                    continue;
                }
                if (px < 0)
                    for (int k = 0; k < 8; k++) v[k] = 0.f;
                if (op.dst32 >= 0) store_f32x8(arena + op.dst32, G.slot, c, q, v);
                if (op.acc32 >= 0) {
                    if (px >= 0) {
                        float o[8];
                        load_f32x8(arena + op.acc32, G.slot, c, q, o);
                        for (int k = 0; k < 8; k++) v[k] += o[k];
                    }
                    store_f32x8(arena + op.acc32, G.slot, c, q, v);
                }
                if (op.raw16 >= 0) store_sp16(arena + op.raw16, G.slot, c, q, v);
                if (op.elu16 >= 0 || op.elu32 >= 0) {
                    for (int k = 0; k < 8; k++) v[k] = elu_ref(v[k]);
                    if (op.elu16 >= 0) store_sp16(arena + op.elu16, G.slot, c, q, v);
                    if (op.elu32 >= 0) store_f32x8(arena + op.elu32, G.slot, c, q, v);
                }
            }
        }
    }
}

static void run_op(const Op& op, const Plan& P, const uint8_t* blob, uint8_t* arena) {
    const int S = P.S;
    const Geo& GS = P.geo[op.gs];
    const Geo& GD = P.geo[op.gd];
    switch (op.kind) {
    case K_CONV: run_conv(op, P, blob, arena); break;
    case K_AFFINE: {
        const float* xin = reinterpret_cast<const float*>(arena + op.src0);
        for (int s = 0; s < S; s++)
            for (int y = 0; y < GS.h; y++)
                for (int x = 0; x < GS.w; x++) {
                    const float* c = xin + ((size_t)s * GS.hw + y * GS.w + x) * 2;
                    float v[8] = {2.f * c[0] - 1.f, 2.f * c[1] - 1.f, 0, 0, 0, 0, 0, 0};
                    store_sp16(arena + op.raw16, GS.slot, 0, qof(GS, s, y, x), v);
                }
        break;
    }
    case K_ELU:
        for (int oct = 0; oct < op.cin / 8; oct++)
            for (int s = 0; s < S; s++)
                for (int y = 0; y < GS.h; y++)
                    for (int x = 0; x < GS.w; x++) {
                        float v[8];
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, y, x), v);
                        for (int k = 0; k < 8; k++) v[k] = elu_ref(v[k]);
                        store_sp16(arena + op.elu16, GS.slot, oct, qof(GS, s, y, x), v);
                    }
        break;
    case K_MAXPOOL5:
        for (int oct = 0; oct < op.cin / 8; oct++)
            for (int s = 0; s < S; s++)
                for (int y = 0; y < GS.h; y++)
                    for (int x = 0; x < GS.w; x++) {
                        float m[8];
                        for (int k = 0; k < 8; k++) m[k] = -INFINITY;
                        for (int yy = std::max(0, y - 2); yy <= std::min(GS.h - 1, y + 2); yy++)
                            for (int xx = std::max(0, x - 2); xx <= std::min(GS.w - 1, x + 2); xx++) {
                                float v[8];
                                load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, yy, xx), v);
                                for (int k = 0; k < 8; k++) m[k] = fmaxf(m[k], v[k]);
                            }
                        store_sp16(arena + op.raw16, GS.slot, oct, qof(GS, s, y, x), m);
                    }
        break;
    case K_UPACC: {
        const int H = GS.h, W = GS.w, OH = GD.h, OW = GD.w;
        const float sy = OH > 1 ? (float)(H - 1) / (float)(OH - 1) : 0.f, sx = OW > 1 ? (float)(W - 1) / (float)(OW - 1) : 0.f;
        for (int oct = 0; oct < op.cin / 8; oct++)
            for (int s = 0; s < S; s++)
                for (int y = 0; y < OH; y++)
                    for (int x = 0; x < OW; x++) {
                        const float fy = sy * (float)y, fx = sx * (float)x;
                        const int y0 = (int)fy, x0 = (int)fx, y1 = y0 + (y0 < H - 1 ? 1 : 0), x1 = x0 + (x0 < W - 1 ? 1 : 0);
                        const float ly = fy - (float)y0, lx = fx - (float)x0, hy = 1.f - ly, hx = 1.f - lx;
                        float p00[8], p01[8], p10[8], p11[8], a[8];
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, y0, x0), p00);
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, y0, x1), p01);
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, y1, x0), p10);
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, y1, x1), p11);
                        const int q = qof(GD, s, y, x);
                        load_f32x8(arena + op.acc32, GD.slot, oct, q, a);
                        for (int k = 0; k < 8; k++) a[k] += hy * (hx * p00[k] + lx * p01[k]) + ly * (hx * p10[k] + lx * p11[k]);
                        store_f32x8(arena + op.acc32, GD.slot, oct, q, a);
                        if (op.elu32 >= 0) {
                            for (int k = 0; k < 8; k++) a[k] = elu_ref(a[k]);
                            store_f32x8(arena + op.elu32, GD.slot, oct, q, a);
                        }
                    }
        break;
    }
    case K_EPILOGUE:
        for (int oct = 0; oct < op.cin / 8; oct++)
            for (int s = 0; s < S; s++)
                for (int y = 0; y < GS.h; y++)
                    for (int x = 0; x < GS.w; x++) {
                        const int q = qof(GS, s, y, x);
                        float v[8];
                        load_f32x8(arena + op.src0, GS.slot, oct, q, v);
                        if (op.dst32 >= 0) store_f32x8(arena + op.dst32, GS.slot, oct, q, v);
                        if (op.acc32 >= 0) {
                            float o[8];
                            load_f32x8(arena + op.acc32, GS.slot, oct, q, o);
                            for (int k = 0; k < 8; k++) v[k] += o[k];
                            store_f32x8(arena + op.acc32, GS.slot, oct, q, v);
                        }
                        if (op.raw16 >= 0) store_sp16(arena + op.raw16, GS.slot, oct, q, v);
                        if (op.elu16 >= 0 || op.elu32 >= 0) {
                            for (int k = 0; k < 8; k++) v[k] = elu_ref(v[k]);
                            if (op.elu16 >= 0) store_sp16(arena + op.elu16, GS.slot, oct, q, v);
                            if (op.elu32 >= 0) store_f32x8(arena + op.elu32, GS.slot, oct, q, v);
                        }
                    }
        break;
    case K_POOL2:
        for (int oct = 0; oct < op.cin / 8; oct++)
            for (int s = 0; s < S; s++)
                for (int y = 0; y < GD.h; y++)
                    for (int x = 0; x < GD.w; x++) {
                        float a[8], b[8], c[8], d[8], v[8];
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, 2 * y, 2 * x), a);
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, 2 * y + 1, 2 * x), b);
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, 2 * y, 2 * x + 1), c);
                        load_f32x8(arena + op.src0, GS.slot, oct, qof(GS, s, 2 * y + 1, 2 * x + 1), d);
                        for (int k = 0; k < 8; k++) v[k] = (a[k] + b[k] + c[k] + d[k]) * 0.25f;
                        store_f32x8(arena + op.dst32, GD.slot, oct, qof(GD, s, y, x), v);
                        if (op.raw16 >= 0) store_sp16(arena + op.raw16, GD.slot, oct, qof(GD, s, y, x), v);
                    }
        break;
    case K_NORM_ELU: {
        const int C = op.cin, hw = GS.hw;
        const float* w = reinterpret_cast<const float*>(blob + op.w_off);
        for (int s = 0; s < S; s++) {
            std::vector<double> mean(C, 0.0), m2(C, 0.0);
            for (int c = 0; c < C; c++) {
                for (int y = 0; y < GS.h; y++)
                    for (int x = 0; x < GS.w; x++) {
                        float v[8];
                        load_f32x8(arena + op.src0, GS.slot, c / 8, qof(GS, s, y, x), v);
                        mean[c] += v[c % 8];
                    }
                mean[c] /= hw;
                for (int y = 0; y < GS.h; y++)
                    for (int x = 0; x < GS.w; x++) {
                        float v[8];
                        load_f32x8(arena + op.src0, GS.slot, c / 8, qof(GS, s, y, x), v);
                        m2[c] += (v[c % 8] - mean[c]) * (v[c % 8] - mean[c]);
                    }
            }
            double m = 0, vv = 0;
            for (int c = 0; c < C; c++) m += mean[c];
            m /= C;
            for (int c = 0; c < C; c++) vv += (mean[c] - m) * (mean[c] - m);
            vv /= (C - 1);
            for (int oct = 0; oct < C / 8; oct++)
                for (int y = 0; y < GS.h; y++)
                    for (int x = 0; x < GS.w; x++) {
                        float v[8];
                        const int q = qof(GS, s, y, x);
                        load_f32x8(arena + op.src0, GS.slot, oct, q, v);
                        for (int k = 0; k < 8; k++) {
                            const int c = oct * 8 + k;
                            const double hn = (v[k] - mean[c]) / sqrt(m2[c] / hw + 1e-5);
                            const double mh = (mean[c] - m) / sqrt(vv + 1e-5);
                            v[k] = elu_ref((float)(w[C + c] * (hn + mh * w[c]) + w[2 * C + c]));
                        }
                        store_sp16(arena + op.elu16, GS.slot, oct, q, v);
                    }
        }
        break;
    }
    default: break;
    }
}

extern "C" {

struct emu2_entry { const char* name; const float* data; const int64_t* shape; int32_t ndim; };

void* emu2_create(const emu2_entry* e, int n, int ngf, int H, int W, int channels, int stage_cap, char* err, int errlen) {
    Emu* m = new Emu();
    for (int i = 0; i < n; i++) {
        int64_t cnt = 1;
        std::vector<int64_t> shp;
        for (int k = 0; k < e[i].ndim; k++) { shp.push_back(e[i].shape[k]); cnt *= e[i].shape[k]; }
        m->storage.emplace_back(e[i].data, e[i].data + cnt);
        m->sd[e[i].name] = TensorArg{m->storage.back().data(), shp};
    }
    try {
        m->b.reset(new Builder(m->sd, ngf, H, W, channels, stage_cap));
    } catch (const std::exception& ex) {
        snprintf(err, errlen, "%s", ex.what());
        delete m;
        return nullptr;
    }
    return m;
}
void emu2_free(void* h) { delete (Emu*)h; }
int emu2_n_ops(void* h) { return ((Emu*)h)->b->n_ops(); }
const char* emu2_op_name(void* h, int i) { return ((Emu*)h)->b->op_names()[(size_t)i].c_str(); }
long long emu2_conv_flops(void* h) { return ((Emu*)h)->b->conv_flops; }
int emu2_max_seg(void* h) { return ((Emu*)h)->b->max_seg; }
int emu2_max_stage(void* h) { return ((Emu*)h)->b->max_stage; }

// plan summary for group size S: arena bytes; tensor table copied into `out` (cap entries); geo[4][12]
long long emu2_plan(void* h, int S, int reuse, TensorInfo* out, int cap, int* n_out, int32_t* geo_out, int32_t* ops_out /* [n_ops][40] */) {
    Plan P = ((Emu*)h)->b->layout(S, reuse != 0);
    const int n = (int)P.tensors.size();
    if (n_out) *n_out = n;
    if (out) memcpy(out, P.tensors.data(), sizeof(TensorInfo) * (size_t)std::min(n, cap));
    if (geo_out) memcpy(geo_out, P.geo, sizeof(Geo) * MAX_LEVELS);
    if (ops_out) memcpy(ops_out, P.ops.data(), sizeof(Op) * P.ops.size());
    return P.arena_bytes;
}

// run the program on S samples x [S][channels][H][W]; net_out [S][channels][H][W] = raw network output (before /sigma);
// arena_out (optional) receives the whole group arena; stop_op < 0: all ops
int emu2_forward(void* h, int S, int reuse, const float* x, float* net_out, uint8_t* arena_out, int stop_op) {
    Emu* m = (Emu*)h;
    const Builder& b = *m->b;
    Plan P = b.layout(S, reuse != 0);
    std::vector<uint8_t> arena((size_t)P.arena_bytes, 0);
    const int HW = b.H * b.W;
    float* xin = reinterpret_cast<float*>(arena.data() + P.x_off);
    for (int s = 0; s < S; s++)
        for (int e = 0; e < HW; e++) {
            xin[((size_t)s * HW + e) * 2] = x[((size_t)s * 2) * HW + e];
            xin[((size_t)s * HW + e) * 2 + 1] = x[((size_t)s * 2 + 1) * HW + e];
        }
    const int n = stop_op < 0 ? (int)P.ops.size() : std::min(stop_op, (int)P.ops.size());
    for (int i = 0; i < n; i++) run_op(P.ops[(size_t)i], P, b.blob.data(), arena.data());
    const float* o = reinterpret_cast<const float*>(arena.data() + P.out_off);
    for (int s = 0; s < S; s++)
        for (int e = 0; e < HW; e++) {
            net_out[((size_t)s * 2) * HW + e] = o[((size_t)s * HW + e) * 2];
            net_out[((size_t)s * 2 + 1) * HW + e] = o[((size_t)s * HW + e) * 2 + 1];
        }
    if (arena_out) memcpy(arena_out, arena.data(), arena.size());
    return 0;
}
}
