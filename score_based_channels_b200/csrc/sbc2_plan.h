// sbc2_plan.h -- host-side planner / packer of the tcgen05 engine ("engine 2").  Pure C++17, no CUDA: compiled by
// nvcc into libsbc_b200.so (sbc_api.cu) and by g++ into the CPU plan-emulation harness (tests/emu/emu2.cpp).
//
// It turns a plain state dict (reference key names, ncsnv2/models/ncsnv2.py:198-262) into
//   * a flat op list restating NCSNv2Deepest.forward (ncsnv2.py:269-300; blocks: layers.py:443-456 ResidualBlock,
//     309-313 ConvMeanPool, 234-249 RefineBlock, 126-134 RCUBlock, 76-83 CRPBlock, 178-184 MSFBlock,
//     normalization.py:163-176 InstanceNorm2dPlus),
//   * a parameter blob (per conv: UMMA instruction list | bias | B tiles in fp16 hi/lo split), and
//   * for a given group size S (samples a CTA walks through the program in lock-step) the layout of the per-group
//     activation arena in global memory (L2 resident) with liveness-based reuse.
//
// Arithmetic of a conv ("fp16x2", fp32-equivalent): every operand is carried as an fp16 pair x = hi + lo
// (hi = rn16(x), lo = rn16(x - hi): 22 significant bits).  One tcgen05.mma kind::f16 instruction (M128, K16) contracts
// A' = [a_hi(8 ch) | a_lo(8 ch)] against B' rows [w_hi | w_hi] (accumulator columns 0..C-1) and [w_lo | w_lo] (columns
// C..2C-1): (a_hi + a_lo) * (w_hi + w_lo) summed in fp32 in TMEM, the epilogue adds the two column groups.  For
// Cin % 16 == 0 the K = 16 slice is 16 channels of hi, then 16 channels of lo, against one shared B tile.
//
// Activation layouts (per geometry level g: h x w image, S samples stacked):
//   padded pixel index q(s, y, x) = lead + s*pps + y*wp + x,   wp = w + hx, pps = (h + hy) * wp, lead = hy*wp + hx
//   (one shared zero column between rows, hy shared zero rows between samples; zero guards of `lead` pixels at both
//   ends) so that a conv tap is a constant shift of q and 128 consecutive q are one UMMA M tile.
//   F32  tensor: [C/4 quads][npx][4 floats]                      16 B per pixel per quad
//   SP16 tensor: [C/8 octets][hi, lo][npx][8 halfs]              16 B per pixel per sub-plane (= one UMMA core row)
//   Both take C/4 "slots" of npx*16 bytes.  SP16 pads are kept zero by every writer.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace sbc2 {

enum : int32_t { K_AFFINE = 0, K_CONV = 1, K_NORM_ELU = 2, K_ELU = 3, K_MAXPOOL5 = 4, K_UPACC = 5, K_POOL2 = 6, K_EPILOGUE = 7 };
enum : int32_t { F_COMPACT = 1 };      // conv: couts 0,1 go to the compact (re, im) output buffer

constexpr int TILE_M = 128;            // UMMA M
constexpr int MAX_LEVELS = 4;

struct Geo {                           // 14 words
    int32_t h, w, hy, hx, wp, rps, pps, lead;
    int32_t npx;                       // allocated pixels per slot = lead + 128*T + lead (multiple of 8)
    int32_t T;                         // 128-pixel tiles covering [lead, lead + S*pps)
    int32_t slot;                      // bytes per slot = npx * 16
    int32_t hw;                        // h * w
    uint32_t mg_pps, mg_wp;            // ceil(2^32 / pps), ceil(2^32 / wp): n / d = umulhi(n, magic) for n * d < 2^32
};

// one op = 40 int32 words = 160 bytes.  Offsets are bytes relative to the group arena (or to the blob); -1 = unused.
struct Op {
    int32_t kind, flags;
    int32_t gs, gd;                    // geometry level of the source(s) / of the outputs
    int32_t cin, cout;                 // channels (conv: of branch 0 / real output channels)
    int32_t src0, src1;                // conv: SP16 source of branch 0 / 1.  others: src0 = the input tensor
    int32_t nsub0, nsub1;              // conv: sub-planes (16 B / pixel) staged per branch
    int32_t dst32, acc32, raw16, elu16, elu32;   // dst32 = v; then v += acc32, acc32 = v; raw16 = split(v); elu16 / elu32 = ELU(v)
    int32_t w_off, w_len;              // parameter segment in the blob (bytes, multiples of 16).  norm: 3*C floats
    int32_t n_mma, mma_rel, bias_rel, btile_rel;   // conv: inside the segment (bytes); bias_rel = -1: none
    int32_t halo;                      // conv: largest |tap shift| in pixels; staged window = 128 + 2*halo pixels
    int32_t N;                         // conv: accumulator columns = 2 * cout8
    int32_t cout8;                     // conv: output channels rounded up to 8
    int32_t idesc;                     // conv: tcgen05 instruction descriptor (kind::f16, M128, N)
    float unscale;                     // conv: 1 / weight scale
    int32_t T;                         // conv: tiles
    int32_t nstage;                    // conv: staging ring depth used by this op (1, 2 or 4)
    int32_t sps;                       // conv: bytes per staged sub-plane = (128 + 2*halo) * 16
    int32_t scratch;                   // norm: byte offset of the statistics scratch (RAW region)
    int32_t nw_off, nw_len;            // conv: segment of the NEXT conv (wraps to the first) for the prefetch
    int32_t wslot;                     // conv: which of the two shared-memory weight buffers holds this segment
    int32_t mma_idx;                   // conv: first entry of this op's UMMA list in the model-wide list (constant memory), or -1
    int32_t fence_after;               // 1: the next op (wrapping) is a conv, i.e. reads through the async proxy
    int32_t pad[5];
};
static_assert(sizeof(Op) == 160, "Op must be 40 words");

// one tcgen05.mma, as the low words of its two shared-memory matrix descriptors relative to the staging slot / to the
// parameter segment: bits 0-13 = byte offset >> 4, bits 16-29 = LBO >> 4.  The kernel adds (base address >> 4); the high
// word (SBO = 128 B, descriptor version 1) is the constant 0x4008.
struct MmaEntry { uint32_t a_lo, b_lo; };
static inline uint32_t desc_lo(uint32_t off, uint32_t lbo) { return ((off >> 4) & 0x3FFFu) | (((lbo >> 4) & 0x3FFFu) << 16); }
static inline uint32_t desc_off(uint32_t lo) { return (lo & 0x3FFFu) << 4; }
static inline uint32_t desc_lbo(uint32_t lo) { return ((lo >> 16) & 0x3FFFu) << 4; }

struct TensorInfo {                    // debug / test view of the arena
    char name[48];
    int32_t fmt;                       // 0 F32, 1 SP16, 2 RAW
    int32_t level, C;
    int64_t off, bytes;
    int32_t born, died;
};

struct Plan {                          // everything that depends on the group size S
    int S = 0;
    Geo geo[MAX_LEVELS];
    std::vector<Op> ops;
    std::vector<std::vector<int32_t>> pix;   // per level: q -> s*hw + y*w + x, or -1 for pads
    int64_t arena_bytes = 0;
    int32_t x_off = 0, out_off = 0, post_off = 0;
    std::vector<TensorInfo> tensors;
    int32_t first_conv = -1;
};

// ---------------------------------------------------------------------------------------------------------
// fp16 helpers (round to nearest even), host side
// ---------------------------------------------------------------------------------------------------------
static inline uint16_t f32_to_f16(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7FFFFFFFu;
    if (x >= 0x7F800000u) return (uint16_t)(sign | 0x7C00u | ((x > 0x7F800000u) ? 0x200u : 0));
    if (x >= 0x477FF000u) return (uint16_t)(sign | 0x7C00u);                 // rounds to >= 65520 -> inf
    if (x < 0x33000001u) return (uint16_t)sign;                               // < 2^-25 (or == 2^-25: ties to even 0)
    int32_t e = (int32_t)(x >> 23) - 127;
    uint32_t m = (x & 0x7FFFFFu) | 0x800000u;
    int shift;
    uint32_t base;
    if (e < -14) { shift = 13 + (-14 - e); base = 0; }                        // subnormal half
    else { shift = 13; base = (uint32_t)(e + 15) << 10; m &= 0x7FFFFFu; }
    uint32_t q = m >> shift, r = m & ((1u << shift) - 1), half = 1u << (shift - 1);
    if (r > half || (r == half && (q & 1))) q++;
    return (uint16_t)(sign | (base + q));                                     // mantissa carry rolls into the exponent
}
static inline float f16_to_f32(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    uint32_t e = (h >> 10) & 0x1F, m = h & 0x3FF, x;
    if (e == 0) {
        if (m == 0) x = sign;
        else {
            int s = 0;
            while (!(m & 0x400)) { m <<= 1; s++; }
            x = sign | ((uint32_t)(127 - 15 - s + 1) << 23) | ((m & 0x3FF) << 13);
        }
    } else if (e == 31) x = sign | 0x7F800000u | (m << 13);
    else x = sign | ((e + 112) << 23) | (m << 13);
    float f;
    memcpy(&f, &x, 4);
    return f;
}

struct TensorArg { const float* data; std::vector<int64_t> shape; };
typedef std::map<std::string, TensorArg> StateDict;

// ---------------------------------------------------------------------------------------------------------
class Builder {
  public:
    int ngf, H, W, channels;
    int arch = 2, npools = 3;            // 0 NCSNv2, 1 NCSNv2Deeper, 2 NCSNv2Deepest
    int64_t conv_flops = 0;
    std::vector<uint8_t> blob;
    std::vector<MmaEntry> all_mma;       // every conv's UMMA list back to back (Op::mma_idx), for the constant-memory copy
    int32_t max_seg = 0, max_stage = 0;
    int stage_cap;                       // bytes available for the A staging ring in shared memory
    std::string error;

    Builder(const StateDict& sd_, int ngf_, int H_, int W_, int channels_, int stage_cap_ = 35 * 1024)
        : ngf(ngf_), H(H_), W(W_), channels(channels_), stage_cap(stage_cap_), sd(sd_) {
        // the architecture is read off the state dict: NCSNv2Deepest has res31 (ncsnv2.py:198-262), NCSNv2Deeper res5 but no
        // res31 (ncsnv2.py:104-156), NCSNv2 neither (ncsnv2.py:11-68)
        arch = has("res31.0.conv1.weight") ? 2 : (has("res5.0.conv1.weight") ? 1 : 0);
        npools = arch == 2 ? 3 : (arch == 1 ? 2 : 1);
        const int div = 1 << npools;
        if (H <= 0 || W <= 0 || H % div || W % div || H % 8 || W % 8) throw std::runtime_error("Nt and Nr must be positive multiples of 8");
        if (ngf <= 0 || ngf % 8) throw std::runtime_error("ngf must be a positive multiple of 8");
        if (channels != 2) throw std::runtime_error("channels must be 2 (re, im)");
        for (int l = 0; l < MAX_LEVELS; l++) {
            Geo& g = base[l];
            memset(&g, 0, sizeof g);
            g.h = std::max(1, H >> l); g.w = std::max(1, W >> l);
            const int dils[3] = {1, 2, 4};
            const int nd = l < npools ? 1 : 3;    // the dilated stages run at the lowest resolution (ncsnv2.py:240-254)
            g.hy = g.hx = 0;
            for (int i = 0; i < nd; i++) {
                if (dils[i] < g.h) g.hy = std::max(g.hy, dils[i]);
                if (dils[i] < g.w) g.hx = std::max(g.hx, dils[i]);
            }
            g.wp = g.w + g.hx; g.rps = g.h + g.hy; g.pps = g.rps * g.wp; g.lead = g.hy * g.wp + g.hx; g.hw = g.h * g.w;
            g.mg_pps = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)g.pps - 1) / (uint64_t)g.pps);
            g.mg_wp = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)g.wp - 1) / (uint64_t)g.wp);
        }
        build();
    }

    // ---- S-dependent part ------------------------------------------------------------------------------
    Plan layout(int S, bool reuse = true) const {
        Plan P;
        P.S = S;
        for (int l = 0; l < MAX_LEVELS; l++) {
            Geo g = base[l];
            g.T = (S * g.pps + TILE_M - 1) / TILE_M;
            g.npx = (g.lead + TILE_M * g.T + g.lead + 7) / 8 * 8;
            g.slot = g.npx * 16;
            if ((uint64_t)g.npx * (uint64_t)g.pps >= ((uint64_t)1 << 32)) throw std::runtime_error("group too large for the index arithmetic");
            P.geo[l] = g;
            std::vector<int32_t> pm((size_t)g.npx, -1);
            for (int s = 0; s < S; s++)
                for (int y = 0; y < g.h; y++)
                    for (int x = 0; x < g.w; x++) pm[(size_t)g.lead + (size_t)s * g.pps + y * g.wp + x] = s * g.hw + y * g.w + x;
            P.pix.push_back(pm);
        }
        // RAW tensors first, never reused
        std::vector<int64_t> toff(tens.size(), -1), tbytes(tens.size(), 0);
        int64_t cur = 0;
        for (size_t i = 0; i < tens.size(); i++) {
            const T_& t = tens[i];
            if (t.fmt != 2) continue;
            toff[i] = cur;
            tbytes[i] = ((int64_t)t.raw_per_sample * S + 127) / 128 * 128;
            cur += tbytes[i];
        }
        // one region per (format, level): slots are interchangeable inside a region, so SP16 pads stay aligned
        for (int fmt = 0; fmt < 2; fmt++)
            for (int l = 0; l < MAX_LEVELS; l++) {
                std::vector<int> ids;
                for (size_t i = 0; i < tens.size(); i++)
                    if (tens[i].fmt == fmt && tens[i].level == l) ids.push_back((int)i);
                if (ids.empty()) continue;
                std::vector<int> start(tens.size(), 0);
                int peak = 0;
                if (reuse) {
                    std::sort(ids.begin(), ids.end(), [&](int a, int b) {
                        if (tens[a].C != tens[b].C) return tens[a].C > tens[b].C;
                        return tens[a].born < tens[b].born;
                    });
                    std::vector<int> placed;
                    for (int id : ids) {
                        const int n = tens[id].C / 4;
                        std::vector<std::pair<int, int>> busy;
                        for (int m : placed)
                            if (tens[m].born < tens[id].died && tens[id].born < tens[m].died) busy.push_back({start[m], tens[m].C / 4});
                        std::sort(busy.begin(), busy.end());
                        int pos = 0;
                        for (auto& b : busy) {
                            if (b.first - pos >= n) break;
                            pos = std::max(pos, b.first + b.second);
                        }
                        start[id] = pos;
                        placed.push_back(id);
                        peak = std::max(peak, pos + n);
                    }
                } else {
                    for (int id : ids) { start[id] = peak; peak += tens[id].C / 4; }
                }
                for (int id : ids) {
                    toff[id] = cur + (int64_t)start[id] * P.geo[l].slot;
                    tbytes[id] = (int64_t)(tens[id].C / 4) * P.geo[l].slot;
                }
                cur += (int64_t)peak * P.geo[l].slot;
            }
        P.arena_bytes = (cur + 255) / 256 * 256;
        if (P.arena_bytes >= ((int64_t)1 << 31)) throw std::runtime_error("group arena exceeds 2 GiB: lower the group size");
        for (size_t i = 0; i < tens.size(); i++) {
            TensorInfo ti;
            memset(&ti, 0, sizeof ti);
            snprintf(ti.name, sizeof ti.name, "%s", tens[i].name.c_str());
            ti.fmt = tens[i].fmt; ti.level = tens[i].level; ti.C = tens[i].C; ti.off = toff[i]; ti.bytes = tbytes[i];
            ti.born = tens[i].born; ti.died = tens[i].died;
            P.tensors.push_back(ti);
        }
        auto res = [&](int id, int chan_off = 0) -> int32_t {
            if (id < 0) return -1;
            return (int32_t)(toff[id] + (int64_t)(chan_off / 4) * P.geo[tens[id].level].slot);
        };
        P.ops = ops;
        int wslot = 0;
        for (size_t i = 0; i < P.ops.size(); i++) {
            Op& o = P.ops[i];
            const R_& r = refs[i];
            o.src0 = res(r.src0, 8 * r.sp0); o.src1 = res(r.src1, 8 * r.sp1);     // first staged plane of each branch
            o.dst32 = (o.flags & F_COMPACT) ? res(r.dst32) : res(r.dst32, r.co0);
            o.acc32 = res(r.acc32, r.co0); o.raw16 = res(r.raw16, r.co0); o.elu16 = res(r.elu16, r.co0);
            o.elu32 = res(r.elu32, r.co0); o.scratch = res(r.scratch);
            if (o.kind == K_CONV) {
                o.T = P.geo[o.gs].T;
                const int stage_bytes = (o.nsub0 + o.nsub1) * o.sps;
                int ns = stage_cap / stage_bytes;
                ns = ns >= 4 ? 4 : (ns >= 2 ? 2 : 1);
                while (ns > 1 && ns > o.T) ns >>= 1;
                o.nstage = ns;
                o.wslot = wslot;
                wslot ^= 1;
                if (P.first_conv < 0) P.first_conv = (int)i;
            }
        }
        for (size_t i = 0; i < P.ops.size(); i++) P.ops[i].fence_after = P.ops[(i + 1) % P.ops.size()].kind == K_CONV ? 1 : 0;
        // prefetch chain: every conv knows the segment of the next conv (the last wraps to the first)
        int next = P.first_conv;
        for (int i = (int)P.ops.size() - 1; i >= 0; i--) {
            Op& o = P.ops[i];
            if (o.kind != K_CONV) continue;
            o.nw_off = P.ops[next].w_off; o.nw_len = P.ops[next].w_len;
            next = i;
        }
        P.x_off = (int32_t)toff[t_xin]; P.out_off = (int32_t)toff[t_out]; P.post_off = (int32_t)toff[t_post];
        return P;
    }

    int n_ops() const { return (int)ops.size(); }
    const std::vector<std::string>& op_names() const { return names; }
    const Geo& level(int l) const { return base[l]; }

  private:
    struct T_ { std::string name; int fmt, level, C, born, died; int64_t raw_per_sample; };
    struct R_ { int src0 = -1, src1 = -1, dst32 = -1, acc32 = -1, raw16 = -1, elu16 = -1, elu32 = -1, scratch = -1, co0 = 0, sp0 = 0, sp1 = 0; };
    struct LT { int f32 = -1, s16 = -1; int C = 0, level = 0; };   // a logical tensor: optional F32 and SP16 copies

    const StateDict& sd;
    Geo base[MAX_LEVELS];
    std::vector<T_> tens;
    std::vector<Op> ops;
    std::vector<R_> refs;
    std::vector<std::string> names;
    int t_xin = -1, t_out = -1, t_post = -1;
    int tmpc = 0;

    const TensorArg& P_(const std::string& k) const {
        auto it = sd.find(k);
        if (it == sd.end()) throw std::runtime_error("state dict has no entry '" + k + "'");
        return it->second;
    }
    static int64_t numel(const TensorArg& t) {
        int64_t n = 1;
        for (auto d : t.shape) n *= d;
        return n;
    }
    // conv weight [cout, cin, k, k] with a matching optional bias [cout]: shapes are validated before anything is indexed
    const TensorArg& W4(const std::string& prefix) const {
        const TensorArg& wt = P_(prefix + ".weight");
        if (wt.shape.size() != 4 || wt.shape[0] <= 0 || wt.shape[1] <= 0 || wt.shape[2] != wt.shape[3] || (wt.shape[2] != 1 && wt.shape[2] != 3))
            throw std::runtime_error("'" + prefix + ".weight' must be a [cout, cin, k, k] tensor with k in {1, 3}");
        if (has(prefix + ".bias") && numel(P_(prefix + ".bias")) != wt.shape[0])
            throw std::runtime_error("'" + prefix + ".bias' does not match the weight's output channels");
        return wt;
    }
    bool has(const std::string& k) const { return sd.find(k) != sd.end(); }

    int newt(int fmt, int C, int level, const std::string& tag) {
        if (fmt != 2 && C % 8) throw std::runtime_error("tensor channels must be a multiple of 8");
        T_ t{tag + std::to_string(tmpc++), fmt, level, C, (int)ops.size(), 1 << 30, 0};
        tens.push_back(t);
        return (int)tens.size() - 1;
    }
    int new32(int C, int level, const char* tag = "f") { return newt(0, C, level, tag); }
    int new16(int C, int level, const char* tag = "h") { return newt(1, C, level, tag); }
    int newraw(int64_t bytes_per_sample, const char* tag) {
        int id = newt(2, 0, 0, tag);
        tens[id].raw_per_sample = bytes_per_sample;
        tens[id].born = 0;
        return id;
    }
    void rel(int id) { if (id >= 0) tens[id].died = (int)ops.size(); }

    void push(const Op& o, const R_& r, const std::string& name) { ops.push_back(o); refs.push_back(r); names.push_back(name); }
    static Op blank(int kind) { Op o; memset(&o, 0, sizeof o); o.kind = kind; o.src0 = o.src1 = o.dst32 = o.acc32 = o.raw16 = o.elu16 = o.elu32 = o.scratch = -1; o.bias_rel = -1; return o; }

    size_t blob_align(size_t a) { while (blob.size() % a) blob.push_back(0); return blob.size(); }

    // ---- conv --------------------------------------------------------------------------------------------
    struct Branch { std::string prefix; int src; int dil; int p0 = 0, p1 = -1; bool bias = true; };   // planes [p0, p1) of src (-1: all)
    struct Outs { int dst32 = -1, acc32 = -1, raw16 = -1, elu16 = -1, elu32 = -1; bool compact = false; };

    void conv(std::vector<Branch> br, Outs out) {
        // stage footprint of the merged op; fall back to one op per branch when it exceeds the staging ring
        bool wide = false;
        for (auto& b : br) wide = wide || tens[b.src].C / 8 > KMAX_PLANES;
        if (wide) { conv1(br, out); return; }
        if (br.size() > 1) {
            int nsub = 0, halo = 0;
            for (auto& b : br) {
                nsub += tens[b.src].C / 4;
                halo = std::max(halo, branch_halo(b));
            }
            if (nsub * (TILE_M + 2 * halo) * 16 > stage_cap) {
                // first branch alone into an F32 tensor (the requested dst32, else a temporary); the rest accumulate
                // on top of it in place and apply the remaining outputs
                if (out.acc32 >= 0 || out.compact) throw std::runtime_error("unsupported sibling split");
                const auto& w0 = W4(br[0].prefix);
                const int cout = (int)w0.shape[0], lvl = tens[br[0].src].level;
                int tmp = -1;
                if (out.dst32 < 0) tmp = new32((cout + 7) / 8 * 8, lvl, "sib");
                Outs o1; o1.dst32 = out.dst32 >= 0 ? out.dst32 : tmp;
                conv1({br[0]}, o1);
                std::vector<Branch> rest(br.begin() + 1, br.end());
                Outs o2 = out;
                o2.dst32 = -1;
                o2.acc32 = o1.dst32;
                conv1(rest, o2);
                rel(tmp);
                return;
            }
        }
        conv1(br, out);
    }
    int branch_halo(const Branch& b) const {
        const auto& wt = W4(b.prefix);
        const int k = (int)wt.shape[2], r = k / 2, lvl = tens[b.src].level;
        const Geo& g = base[lvl];
        int halo = 0;
        for (int tap = 0; tap < k * k; tap++) {
            const int dy = (tap / k - r) * b.dil, dx = (tap % k - r) * b.dil;
            if (abs(dy) < g.h && abs(dx) < g.w) halo = std::max(halo, abs(dy * g.wp + dx));
        }
        return halo;
    }

    static constexpr int KMAX_PLANES = 4;      // input channels per conv op = 32: bounds the staged window and the B tiles

    void conv1(const std::vector<Branch>& br, const Outs& out) {
        bool wide = false;
        for (auto& b : br) wide = wide || tens[b.src].C / 8 > KMAX_PLANES;
        if (!wide) { conv_chunks(br, out); return; }
        // wide inputs: one op per group of 32 input channels, summed in an F32 temporary; a K_EPILOGUE op then applies the
        // requested outputs to the finished sum
        if (out.compact) throw std::runtime_error("compact conv with a wide input is not supported");
        const auto& w0 = W4(br[0].prefix);
        const int cout = (int)w0.shape[0], lvl = tens[br[0].src].level;
        const int T = new32((cout + 7) / 8 * 8, lvl, "ks");
        bool first = true;
        for (auto& b : br) {
            const int planes = tens[b.src].C / 8;
            for (int p0 = 0; p0 < planes; p0 += KMAX_PLANES) {
                Branch piece = b;
                piece.p0 = p0; piece.p1 = std::min(planes, p0 + KMAX_PLANES); piece.bias = (p0 == 0);
                Outs o;
                if (first) o.dst32 = T; else o.acc32 = T;
                first = false;
                conv_chunks({piece}, o);
            }
        }
        Op o = blank(K_EPILOGUE);
        o.gs = o.gd = lvl; o.cin = o.cout = (cout + 7) / 8 * 8;
        R_ r; r.src0 = T; r.dst32 = out.dst32; r.acc32 = out.acc32; r.raw16 = out.raw16; r.elu16 = out.elu16; r.elu32 = out.elu32;
        push(o, r, "epilogue(" + br[0].prefix + ")");
        rel(T);
    }

    void conv_chunks(const std::vector<Branch>& br, const Outs& out) {
        const auto& w0 = W4(br[0].prefix);
        const int cout = (int)w0.shape[0];
        const int cout8 = std::max(8, (cout + 7) / 8 * 8);
        const int N = 2 * cout8;
        if (N > 64) {   // accumulator slot = 64 TMEM columns: split the output channels
            if (out.compact) throw std::runtime_error("compact conv too wide");
            for (int co0 = 0; co0 < cout8; co0 += 32) conv_emit(br, out, co0, std::min(32, cout8 - co0), cout);
        } else {
            conv_emit(br, out, 0, cout8, cout);
        }
    }

    void conv_emit(const std::vector<Branch>& br, const Outs& out, int co0, int c8, int cout_real) {
        const int lvl = tens[br[0].src].level;
        const Geo& g = base[lvl];
        const int N = 2 * c8;
        int halo = 0;
        for (auto& b : br) halo = std::max(halo, branch_halo(b));
        if (halo > g.lead) throw std::runtime_error("halo exceeds the guard band");
        const int sps = (TILE_M + 2 * halo) * 16;
        // weight scale: largest power of two that keeps every |w| * ws below 2^14 (lo parts stay normal halfs)
        float wmax = 0.f;
        for (auto& b : br) {
            const auto& wt = P_(b.prefix + ".weight");
            int64_t n = 1;
            for (auto d : wt.shape) n *= d;
            for (int64_t i = 0; i < n; i++) wmax = std::max(wmax, fabsf(wt.data[i]));
        }
        float ws = 1.f;
        if (wmax > 0.f) ws = exp2f(floorf(log2f(16384.f / wmax)));
        if (!(ws > 0.f) || !std::isfinite(ws)) ws = 1.f;
        ws = std::min(ws, 1.8446744e19f);   // 2^64

        std::vector<MmaEntry> list;
        std::vector<uint8_t> tiles;
        std::vector<float> bias((size_t)c8, 0.f);
        bool any_bias = false;
        int subbase = 0, nsub[2] = {0, 0};
        if (br.size() > 2) throw std::runtime_error("at most two summed convs per op");
        for (size_t bi = 0; bi < br.size(); bi++) {
            const Branch& b = br[bi];
            const auto& wt = W4(b.prefix);
            const int bco = (int)wt.shape[0], cin = (int)wt.shape[1], k = (int)wt.shape[2], r = k / 2;
            if (bco != cout_real || wt.shape[3] != k || (k != 1 && k != 3)) throw std::runtime_error("bad conv weight " + b.prefix);
            const T_& st = tens[b.src];
            if (st.fmt != 1 || st.level != lvl) throw std::runtime_error("conv source must be an SP16 tensor of the op's level: " + b.prefix);
            if (cin > st.C) throw std::runtime_error("conv source has too few channels: " + b.prefix);
            const int pb = b.p0, pe = b.p1 < 0 ? st.C / 8 : b.p1;     // plane range of this piece
            if (co0 == 0 && pb == 0) conv_flops += 2LL * g.h * g.w * cin * k * k * bco;   // dense count, reference convention
            const int planes = pe - pb;
            nsub[bi] = 2 * planes;
            if (has(b.prefix + ".bias") && b.bias) {
                const auto& bt = P_(b.prefix + ".bias");
                for (int c = 0; c < c8; c++)
                    if (co0 + c < bco) bias[c] += bt.data[co0 + c];
                any_bias = true;
            }
            auto W_ = [&](int co, int ci, int tap) -> float {
                if (co >= bco || ci >= cin) return 0.f;
                return wt.data[((size_t)co * cin + ci) * k * k + tap] * ws;
            };
            for (int tap = 0; tap < k * k; tap++) {
                const int dy = (tap / k - r) * b.dil, dx = (tap % k - r) * b.dil;
                if (!(abs(dy) < g.h && abs(dx) < g.w)) continue;   // the tap only ever reads zero padding
                const int shift = dy * g.wp + dx;
                int p = 0;
                while (p < planes) {
                    const bool pair = (p + 1 < planes);
                    // B tile: [kh 0..1][n/8][n%8][8 halfs]; rows 0..c8-1 = hi weights, c8..2*c8-1 = lo weights
                    const uint32_t b_off = (uint32_t)tiles.size();
                    tiles.resize(tiles.size() + (size_t)N * 32, 0);
                    uint16_t* bt = reinterpret_cast<uint16_t*>(tiles.data() + b_off);
                    for (int n = 0; n < c8; n++)
                        for (int kk = 0; kk < 16; kk++) {
                            const int ci = pair ? (8 * (pb + p) + kk) : (8 * (pb + p) + (kk & 7));
                            const float wv = W_(co0 + n, ci, tap);
                            const uint16_t hi = f32_to_f16(wv);
                            const uint16_t lo = f32_to_f16(wv - f16_to_f32(hi));
                            const size_t kh = kk >> 3, ke = kk & 7;
                            auto at = [&](int row) { return kh * ((size_t)N / 8 * 64) + (size_t)(row >> 3) * 64 + (size_t)(row & 7) * 8 + ke; };
                            bt[at(n)] = hi;
                            bt[at(c8 + n)] = lo;
                        }
                    const uint32_t a_base = (uint32_t)((halo + shift) * 16);
                    const uint32_t blo = desc_lo(b_off, (uint32_t)N * 16u);       // rebased to the segment below
                    if (pair) {      // 16 channels of hi, then 16 channels of lo, same B tile
                        list.push_back({desc_lo((uint32_t)(subbase + 2 * p) * sps + a_base, (uint32_t)(2 * sps)), blo});
                        list.push_back({desc_lo((uint32_t)(subbase + 2 * p + 1) * sps + a_base, (uint32_t)(2 * sps)), blo});
                        p += 2;
                    } else {         // [hi | lo] of 8 channels against [w | w]
                        list.push_back({desc_lo((uint32_t)(subbase + 2 * p) * sps + a_base, (uint32_t)sps), blo});
                        p += 1;
                    }
                }
            }
            subbase += nsub[bi];
        }
        if (list.empty()) throw std::runtime_error("conv without live taps");
        // ---- segment: [mma list (padded to 16 B)][bias c8 floats][B tiles] ----
        const size_t n_real = list.size();
        if (n_real & 1) list.push_back(list.back());           // the kernel reads entry pairs; n_mma stays the real count
        const size_t seg0 = blob_align(128);
        const size_t mma_rel = 0, bias_rel = list.size() * sizeof(MmaEntry);
        size_t btile_rel = bias_rel + (size_t)c8 * 4;
        btile_rel = (btile_rel + 127) / 128 * 128;
        blob.resize(seg0 + btile_rel + tiles.size(), 0);
        memcpy(blob.data() + seg0 + mma_rel, list.data(), list.size() * sizeof(MmaEntry));
        memcpy(blob.data() + seg0 + bias_rel, bias.data(), (size_t)c8 * 4);
        memcpy(blob.data() + seg0 + btile_rel, tiles.data(), tiles.size());
        blob_align(16);
        // b_off of the entries is relative to the tile array: rebase to the segment
        MmaEntry* le = reinterpret_cast<MmaEntry*>(blob.data() + seg0);
        for (size_t i = 0; i < list.size(); i++) le[i].b_lo += (uint32_t)(btile_rel >> 4);
        const int32_t mma_idx = (int32_t)all_mma.size();
        for (size_t i = 0; i < n_real; i++) all_mma.push_back(le[i]);

        Op o = blank(K_CONV);
        o.flags = out.compact ? F_COMPACT : 0;
        o.gs = o.gd = lvl;
        o.cin = tens[br[0].src].C; o.cout = std::min(c8, std::max(0, cout_real - co0));
        o.nsub0 = nsub[0]; o.nsub1 = nsub[1];
        o.w_off = (int32_t)seg0; o.w_len = (int32_t)(blob.size() - seg0);
        o.n_mma = (int32_t)n_real; o.mma_rel = (int32_t)mma_rel;
        o.bias_rel = any_bias ? (int32_t)bias_rel : -1;
        o.btile_rel = (int32_t)btile_rel;
        o.halo = halo; o.N = N; o.cout8 = c8; o.sps = sps;
        o.idesc = (int32_t)((1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(TILE_M >> 4) << 24));   // F32 accum, F16 x F16, K-major
        o.unscale = 1.0f / ws;
        o.mma_idx = mma_idx;
        max_seg = std::max(max_seg, o.w_len);
        max_stage = std::max(max_stage, (nsub[0] + nsub[1]) * sps);
        R_ r;
        r.src0 = br[0].src; r.src1 = br.size() > 1 ? br[1].src : -1;
        r.sp0 = br[0].p0; r.sp1 = br.size() > 1 ? br[1].p0 : 0;
        r.dst32 = out.dst32; r.acc32 = out.acc32; r.raw16 = out.raw16; r.elu16 = out.elu16; r.elu32 = out.elu32; r.co0 = co0;
        std::string nm = br[0].prefix;
        for (size_t i = 1; i < br.size(); i++) nm += "+" + br[i].prefix;
        if (co0 || c8 < (cout_real + 7) / 8 * 8) nm += "[co" + std::to_string(co0) + "]";
        if (br[0].p1 >= 0) nm += "[ci" + std::to_string(8 * br[0].p0) + "]";
        push(o, r, nm);
    }

    // ---- non-conv ops -------------------------------------------------------------------------------------
    void norm_elu(const std::string& p, int src32, int dst16) {
        const int C = tens[src32].C, lvl = tens[src32].level;
        const auto &al = P_(p + ".alpha"), &ga = P_(p + ".gamma"), &be = P_(p + ".beta");
        if (numel(al) != C || numel(ga) != C || numel(be) != C) throw std::runtime_error("'" + p + "' alpha / gamma / beta must have " + std::to_string(C) + " entries");
        const size_t off = blob_align(16);
        blob.resize(off + (size_t)3 * C * 4);
        memcpy(blob.data() + off, al.data, (size_t)C * 4);
        memcpy(blob.data() + off + (size_t)C * 4, ga.data, (size_t)C * 4);
        memcpy(blob.data() + off + (size_t)2 * C * 4, be.data, (size_t)C * 4);
        Op o = blank(K_NORM_ELU);
        o.gs = o.gd = lvl; o.cin = o.cout = C; o.w_off = (int32_t)off; o.w_len = 3 * C * 4;
        R_ r; r.src0 = src32; r.elu16 = dst16;
        push(o, r, p);
    }
    void elu(int src32, int dst16) {
        Op o = blank(K_ELU);
        o.gs = o.gd = tens[src32].level; o.cin = o.cout = tens[src32].C;
        R_ r; r.src0 = src32; r.elu16 = dst16;
        push(o, r, "elu");
    }
    void maxpool5(int src32, int dst16) {
        Op o = blank(K_MAXPOOL5);
        o.gs = o.gd = tens[src32].level; o.cin = o.cout = tens[src32].C;
        R_ r; r.src0 = src32; r.raw16 = dst16;
        push(o, r, "maxpool5");
    }
    void upacc(int src32, int acc32, int elu32) {
        Op o = blank(K_UPACC);
        o.gs = tens[src32].level; o.gd = tens[acc32].level; o.cin = o.cout = tens[src32].C;
        R_ r; r.src0 = src32; r.acc32 = acc32; r.elu32 = elu32;
        push(o, r, "upacc");
    }
    void pool2(int src32, int dst32, int raw16) {
        Op o = blank(K_POOL2);
        o.gs = tens[src32].level; o.gd = tens[dst32].level; o.cin = o.cout = tens[src32].C;
        R_ r; r.src0 = src32; r.dst32 = dst32; r.raw16 = raw16;
        push(o, r, "pool2");
    }
    void affine(int src_raw, int dst16) {
        Op o = blank(K_AFFINE);
        o.gs = o.gd = 0; o.cin = channels; o.cout = 8;
        R_ r; r.src0 = src_raw; r.raw16 = dst16;
        push(o, r, "2x-1");
    }

    // ---- blocks (mirroring program.py of engine 1) ------------------------------------------------------------
    // ResidualBlock without resampling (layers.py:443-456): x += conv2(ELU(norm2(conv1(ELU(norm1(x))))))
    void residual_same(const std::string& p, LT& x, int dil, int elu16_out, bool want_raw16) {
        const int C = x.C, l = x.level;
        int t = new16(C, l);
        norm_elu(p + ".normalize1", x.f32, t);
        int t2 = new32(C, l);
        { Outs o; o.dst32 = t2; conv({{p + ".conv1", t, dil}}, o); }
        rel(t);
        int t3 = new16(C, l);
        norm_elu(p + ".normalize2", t2, t3);
        rel(t2);
        if (want_raw16 && x.s16 < 0) x.s16 = new16(C, l, "r");
        { Outs o; o.acc32 = x.f32; o.elu16 = elu16_out; o.raw16 = want_raw16 ? x.s16 : -1; conv({{p + ".conv2", t3, dil}}, o); }
        rel(t3);
    }
    // first ('down') block of a stage; skip (F32 + raw SP16) must survive.  Returns the output (F32 only).
    LT residual_down(const std::string& p, const LT& skip, int cout, int dil /*0: mean-pool*/) {
        const int C = skip.C, l = skip.level, d = dil ? dil : 1;
        int t = new16(C, l);
        norm_elu(p + ".normalize1", skip.f32, t);
        int t2 = new32(C, l);
        { Outs o; o.dst32 = t2; conv({{p + ".conv1", t, d}}, o); }
        rel(t);
        int t3 = new16(C, l);
        norm_elu(p + ".normalize2", t2, t3);
        rel(t2);
        LT out;
        out.C = cout;
        if (!dil) {   // ConvMeanPool 3x3 + ConvMeanPool 1x1 shortcut (layers.py:416-420): pool(conv2 + shortcut)
            out.level = l + 1;
            int pre = new32(cout, l, "pre");
            { Outs o; o.dst32 = pre; conv({{p + ".conv2.conv", t3, 1}, {p + ".shortcut.conv", skip.s16, 1}}, o); }
            rel(t3);
            out.f32 = new32(cout, l + 1, "o");
            pool2(pre, out.f32, -1);
            rel(pre);
        } else {
            out.level = l;
            out.f32 = new32(cout, l, "o");
            { Outs o; o.dst32 = out.f32; conv({{p + ".conv2", t3, d}, {p + ".shortcut", skip.s16, d}}, o); }
            rel(t3);
        }
        return out;
    }
    // RCUBlock (layers.py:126-134), in place on x.f32.  e16 = SP16 ELU(x) if already available (consumed).
    // Returns SP16 ELU(x_out) when want_elu16; want_raw16 makes x.s16 = split(x_out); want_elu32 returns F32 ELU(x_out)
    void rcu(const std::string& p, LT& x, int nblocks, int e16, bool want_elu16, int* elu16_out, bool want_raw16,
             bool want_elu32, int* elu32_out) {
        const int C = x.C, l = x.level;
        int e = e16;
        for (int i = 0; i < nblocks; i++) {
            if (e < 0) { e = new16(C, l, "e"); elu(x.f32, e); }
            int u = new16(C, l, "u");
            { Outs o; o.elu16 = u; conv({{p + "." + std::to_string(i + 1) + "_1_conv", e, 1}}, o); }
            const bool last = (i == nblocks - 1);
            Outs o;
            o.acc32 = x.f32;
            if (!last || want_elu16) o.elu16 = e;          // e is rewritten in place: its reader (conv1) is done
            else { rel(e); e = -1; }
            if (last && want_raw16) { if (x.s16 < 0) x.s16 = new16(C, l, "r"); o.raw16 = x.s16; }
            if (last && want_elu32) { *elu32_out = new32(C, l, "E"); o.elu32 = *elu32_out; }
            conv({{p + "." + std::to_string(i + 1) + "_2_conv", u, 1}}, o);
            rel(u);
        }
        if (elu16_out) *elu16_out = e;
    }
    // CRPBlock (layers.py:76-83): e32 = ELU(x) is the running sum.  Returns SP16 ELU(sum).
    int crp(const std::string& p, int e32) {
        const int C = tens[e32].C, l = tens[e32].level;
        int m = new16(C, l, "m");
        maxpool5(e32, m);
        int path = new32(C, l, "p");
        { Outs o; o.dst32 = path; o.acc32 = e32; conv({{p + ".convs.0", m, 1}}, o); }
        maxpool5(path, m);
        rel(path);
        int e2 = new16(C, l, "e");
        { Outs o; o.acc32 = e32; o.elu16 = e2; conv({{p + ".convs.1", m, 1}}, o); }
        rel(m);
        return e2;
    }
    // RefineBlock (layers.py:234-249).  xs[i].f32 are consumed; es[i] = SP16 ELU(xs[i]) or -1.
    LT refine(const std::string& p, std::vector<LT> xs, std::vector<int> es, int features, bool end, int* elu16_out) {
        int e32 = -1;
        const int l0 = xs[0].level;
        if (xs.size() == 1) {
            rcu(p + ".adapt_convs.0", xs[0], 2, es[0], false, nullptr, false, true, &e32);
            rel(xs[0].f32); rel(xs[0].s16);
        } else {
            for (size_t i = 0; i < xs.size(); i++) rcu(p + ".adapt_convs." + std::to_string(i), xs[i], 2, es[i], false, nullptr, true, false, nullptr);
            e32 = new32(features, l0, "E");
            if (xs[1].level == l0) {   // same-size bilinear (align_corners=True) is the identity: one summed conv
                Outs o; o.elu32 = e32;
                conv({{p + ".msf.convs.0", xs[0].s16, 1}, {p + ".msf.convs.1", xs[1].s16, 1}}, o);
            } else {
                int s = new32(features, l0, "s");
                { Outs o; o.dst32 = s; conv({{p + ".msf.convs.0", xs[0].s16, 1}}, o); }
                int lo = new32(features, xs[1].level, "lo");
                { Outs o; o.dst32 = lo; conv({{p + ".msf.convs.1", xs[1].s16, 1}}, o); }
                upacc(lo, s, e32);
                rel(lo); rel(s);
            }
            for (auto& x : xs) { rel(x.f32); rel(x.s16); }
        }
        int e2 = crp(p + ".crp", e32);
        LT h; h.C = features; h.level = l0; h.f32 = e32;
        rcu(p + ".output_convs", h, end ? 3 : 1, e2, !end, elu16_out, false, false, nullptr);
        return h;
    }
    // two ResidualBlocks, the first 'down'.  Returns the stage output (F32 + raw SP16 when wanted) and SP16 ELU(out)
    LT stage(const std::string& p, const LT& skip, int cout, int dil, bool want_elu16, int* elu16, bool want_raw16) {
        LT out = residual_down(p + ".0", skip, cout, dil);
        int e = -1;
        if (want_elu16) e = new16(cout, out.level, "e");
        residual_same(p + ".1", out, dil ? dil : 1, e, want_raw16);
        if (elu16) *elu16 = e;
        return out;
    }

    void build() {
        t_xin = newraw((int64_t)channels * H * W * 4, "x_in");
        t_out = newraw((int64_t)channels * H * W * 4, "net_out");
        t_post = newraw((int64_t)2 * H * W * 4, "post");
        int a = new16(8, 0, "a");
        affine(t_xin, a);
        LT o; o.C = ngf; o.level = 0; o.f32 = new32(ngf, 0, "o");
        { Outs oo; oo.dst32 = o.f32; conv({{"begin_conv", a, 1}}, oo); }
        rel(a);
        residual_same("res1.0", o, 1, -1, false);
        residual_same("res1.1", o, 1, -1, true);          // l1: raw SP16 feeds res2.0's shortcut conv
        // encoder stages after res1 (name, width multiplier, dilation; 0 = ConvMeanPool) and decoder blocks (name, width
        // multiplier), per architecture (ncsnv2.py:26-58, 118-150, 218-262)
        struct St { const char* name; int mult, dil; };
        struct Rf { const char* name; int mult; };
        std::vector<St> stages;
        std::vector<Rf> rfs;
        if (arch == 2) {
            stages = {{"res2", 2, 0}, {"res3", 2, 0}, {"res31", 2, 0}, {"res4", 4, 2}, {"res5", 4, 4}};
            rfs = {{"refine1", 4}, {"refine2", 2}, {"refine31", 2}, {"refine3", 2}, {"refine4", 1}, {"refine5", 1}};
        } else if (arch == 1) {
            stages = {{"res2", 2, 0}, {"res3", 2, 0}, {"res4", 4, 2}, {"res5", 4, 4}};
            rfs = {{"refine1", 4}, {"refine2", 2}, {"refine3", 2}, {"refine4", 1}, {"refine5", 1}};
        } else {
            stages = {{"res2", 2, 0}, {"res3", 2, 2}, {"res4", 2, 4}};
            rfs = {{"refine1", 2}, {"refine2", 2}, {"refine3", 1}, {"refine4", 1}};
        }
        std::vector<LT> skips = {o};            // l1, l2, ...
        std::vector<int> elus = {-1};           // SP16 ELU(skip) where the stage's last conv emitted it (dilated stages)
        for (size_t i = 0; i < stages.size(); i++) {
            int e = -1;
            const bool dilated = stages[i].dil > 0;
            // the raw SP16 copy of a skip only feeds the next stage's shortcut conv: dead right after that stage
            LT out = stage(stages[i].name, skips.back(), stages[i].mult * ngf, stages[i].dil, dilated, &e, i + 1 < stages.size());
            rel(skips.back().s16); skips.back().s16 = -1;
            skips.push_back(out);
            elus.push_back(e);
        }
        const int n = (int)skips.size();
        int e_prev = -1;
        LT r5 = refine(rfs[0].name, {skips[n - 1]}, {elus[n - 1]}, rfs[0].mult * ngf, false, &e_prev);
        for (int j = 1; j < n; j++) {
            int e = -1;
            const bool end = (j == n - 1);
            r5 = refine(rfs[j].name, {skips[n - 1 - j], r5}, {elus[n - 1 - j], e_prev}, rfs[j].mult * ngf, end, end ? nullptr : &e);
            e_prev = e;
        }
        int t = new16(ngf, 0);
        norm_elu("normalizer", r5.f32, t);
        rel(r5.f32);
        { Outs oo; oo.dst32 = t_out; oo.compact = true; conv({{"end_conv", t, 1}}, oo); }
        rel(t);
        // every tensor still marked live dies at the end
        for (auto& tt : tens)
            if (tt.died == (1 << 30)) tt.died = (int)ops.size() + 1;
        // norm statistics scratch: per sample 4*C floats (mean, m2, a, b), one RAW buffer shared by all norms
        int maxC = 0;
        for (auto& oo : ops)
            if (oo.kind == K_NORM_ELU) maxC = std::max(maxC, oo.cin);
        int sc = newraw((int64_t)4 * maxC * 4, "nstat");
        tens[sc].died = (int)ops.size() + 1;
        for (size_t i = 0; i < ops.size(); i++)
            if (ops[i].kind == K_NORM_ELU) refs[i].scratch = sc;
    }
};

}  // namespace sbc2
