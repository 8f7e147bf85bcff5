// sbc1_plan.h -- host-side planner / packer of engine 1 (the fused shared-memory-arena kernel, sbc_kernel.cuh) in C++.
// Pure C++17, no CUDA.  It is the in-library twin of score_based_channels_b200/program.py: from a plain state dict
// (reference key names, ncsnv2/models/ncsnv2.py:198-262) it produces the SAME op table, geometry table and parameter
// blob, word for word (tests/test_host.py compares the two), so that a C / Julia / MATLAB caller can create an engine-1
// model through sbc_model_create_from_state_ex without any Python.  The Python planner stays the readable
// specification (it carries the schedule simulator the CPU tests pin against the reference modules).
//
// Schedule restated: NCSNv2Deepest.forward (ncsnv2.py:269-300) with ResidualBlock (layers.py:443-456), ConvMeanPool
// (309-313), RefineBlock (234-249), RCUBlock (126-134), CRPBlock (76-83), MSFBlock (178-184), InstanceNorm2dPlus
// (normalization.py:163-176).  `park` plans for two resident CTAs per SM (see ProgramBuilder in program.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "sbc2_plan.h"      // sbc2::StateDict / TensorArg
#include "sbc_program.h"

namespace sbc1 {

constexpr int SMEM_PER_SM = 233472, SMEM_RESERVED_PER_CTA = 1024, SMEM_STATIC_BOUND = 1024;   // program.py
inline int smem_budget_bytes(int ctas_per_sm) { return SMEM_PER_SM / ctas_per_sm - SMEM_RESERVED_PER_CTA - SMEM_STATIC_BOUND; }

struct Geo {
    int h, w, hy, hx;
    int wp() const { return w + 2 * hx; }
    int pps() const { return (h + 2 * hy) * wp(); }
    int org() const { return hy * wp() + hx; }
    int floats(int c) const { return ((c + 3) / 4) * pps() * 4; }
};
inline int ilog2(int v) {
    if (v <= 0 || (v & (v - 1))) return -1;
    int l = 0;
    while ((1 << l) < v) l++;
    return l;
}

// symbolic tensor reference until the arena plan is solved
struct Ref {
    enum Kind { NONE, ARENA, PARK } kind = NONE;
    int id = -1;       // ARENA: planner tensor id; PARK: park slot id
    int shift = 0;
    static Ref none() { return Ref(); }
    static Ref arena(int id, int shift = 0) { Ref r; r.kind = ARENA; r.id = id; r.shift = shift; return r; }
    static Ref park(int id, int shift = 0) { Ref r; r.kind = PARK; r.id = id; r.shift = shift; return r; }
};

struct Branch { int src; int cin, k, dil, KC, step0; std::vector<int> live; };

struct SymOp {
    SbcOp op;
    Ref src, dst, acc, edst, scratch, wbuf;
    std::vector<Branch> branches;
    std::string name;
};

struct Plan {
    std::vector<SbcOp> ops;
    std::vector<std::string> names;
    std::vector<SbcGeo> geos;
    std::vector<float> blob;
    int arena_floats = 0, in_off = 0, out_off = 0, post_off = 0, max_w_len = 0, park_floats = 0, nthreads = 0;
    long long conv_flops = 0;
    bool x3 = true;
    int misc_bytes() const {
        int halo = 0;
        for (const SbcGeo& g : geos) halo += g.pps - g.h * g.w;
        return (64 + 2 * halo + 15) / 16 * 16;
    }
    int smem_bytes() const { return 4 * arena_floats + misc_bytes(); }
};

class Builder {
  public:
    Builder(const sbc2::StateDict& sd_, int ngf_, int H_, int W_, int channels_, int nthreads_, bool x3_, bool park_)
        : sd(sd_), ngf(ngf_), H(H_), W(W_), channels(channels_), nthreads(nthreads_), x3(x3_), park(park_) {
        if (H <= 0 || W <= 0 || H % 8 || W % 8) throw std::runtime_error("Nt and Nr must be positive multiples of 8");
        if (ngf % 8) throw std::runtime_error("ngf must be a multiple of 8");
        late_floats = park ? 4700 : (1 << 30);
        for (int lvl = 0; lvl < 4; lvl++) {
            const int h = H >> lvl, w = W >> lvl;
            const int dils[3] = {1, 2, 4};
            const int nd = lvl < 3 ? 1 : 3;
            int hy = 0, hx = 0;
            for (int i = 0; i < nd; i++) {
                if (dils[i] < h) hy = std::max(hy, dils[i]);
                if (dils[i] < w) hx = std::max(hx, dils[i]);
            }
            if (w == 2) hx = std::max(hx, 2);
            geos.push_back(Geo{h, w, hy, hx});
        }
    }

    Plan build() {
        const int nx = channels * H * W;
        const int xin = new_raw(nx);
        pin(xin, 0);
        int a = tmp(8, H, W);
        affine(xin, a);
        int xin_slot = -1;
        if (park) {
            xin_slot = gnew(nx);
            copy_raw(SBC_OP_SPILL, Ref::arena(xin), Ref::park(xin_slot), nx, "spill x_in");
            free_(xin);
        }
        int o = tmp(ngf, H, W);
        conv("begin_conv", a, o, -1, -1);
        free_(a);
        int l1 = residual("res1.1", residual("res1.0", o, ngf, -1), ngf, -1);
        int l2, l3, l31, l4, l5, e_;
        int el3 = -1, el31 = -1, el4 = -1, el5 = -1;
        stage("res2", l1, 2 * ngf, 0, false, l2, e_);
        stage("res3", l2, 2 * ngf, 0, false, l3, el3);
        keep(l2);
        stage("res31", l3, 2 * ngf, 0, false, l31, el31);
        keep(l3); keep(el3);
        stage("res4", l31, 4 * ngf, 2, true, l4, el4);
        keep(l31); keep(el31);
        stage("res5", l4, 4 * ngf, 4, true, l5, el5);
        keep(l4); keep(el4);
        int r1, e1, r2, e2, r31, e31, r3, e3, r4, e4, r5, e5;
        refine("refine1", {l5}, {el5}, 4 * ngf, false, r1, e1);
        { const int b0 = back(l4), b1 = back(el4); refine("refine2", {b0, r1}, {b1, e1}, 2 * ngf, false, r2, e2); }
        { const int b0 = back(l31), b1 = back(el31); refine("refine31", {b0, r2}, {b1, e2}, 2 * ngf, false, r31, e31); }
        { const int b0 = back(l3), b1 = back(el3); refine("refine3", {b0, r31}, {b1, e31}, 2 * ngf, false, r3, e3); }
        { const int b0 = back(l2); refine("refine4", {b0, r3}, {-1, e3}, ngf, false, r4, e4); }
        refine("refine5", {l1, r4}, {-1, e4}, ngf, true, r5, e5);
        int t = tmp(ngf, H, W);
        const int r5s = smem_of(r5);
        norm_elu("normalizer", r5s, t);
        free_(r5s);
        const int out = new_raw(nx);
        conv("end_conv", t, out, -1, -1, 1, false, true);
        free_(t);
        if (park) {
            const int xin2 = new_raw(nx);
            pin(xin2, 0);
            copy_raw(SBC_OP_FILL, Ref::park(xin_slot), Ref::arena(xin2), nx, "fill x_in");
        }
        const int post = new_raw(2 * H * W);
        int max_w_len = 0;
        for (const SymOp& so : ops) max_w_len = std::max(max_w_len, so.op.w_len);
        // staging buffers of the parameter segments (program.py:build)
        const int end = (int)ops.size() + 1;
        int prev = -1;
        for (int i = 0; i < (int)ops.size(); i++) {
            SbcOp& op = ops[i].op;
            if (op.w_len > 0) {
                int id;
                if (prev < 0) id = alloc(op.w_len, 0, end);
                else if (op.w_len > late_floats) { op.flags |= SBC_F_LATEW; id = alloc(op.w_len, i, i + 1); }
                else id = alloc(op.w_len, prev, i + 1);
                ops[i].wbuf = Ref::arena(id);
                prev = i;
            }
        }
        solve(end);
        Plan P;
        for (SymOp& so : ops) {
            if (!so.branches.empty()) {   // sibling convs: fold the distance between the source tensors into the K-step offsets
                const int base = T[so.branches[0].src].off;
                int32_t* tab = reinterpret_cast<int32_t*>(blob.data() + so.op.w_off);
                for (const Branch& m : so.branches) {
                    const int n = (int)m.live.size() * m.KC;
                    for (int k = 0; k < n; k++) tab[m.step0 + k] += T[m.src].off - base;
                }
            }
            so.op.src = resolve(so.src); so.op.dst = resolve(so.dst); so.op.acc = resolve(so.acc);
            so.op.edst = resolve(so.edst); so.op.scratch = resolve(so.scratch); so.op.wbuf = resolve(so.wbuf);
        }
        halo_analysis(peak, {{T[xin].off, nx}, {T[post].off, 2 * H * W}});
        for (const SymOp& so : ops) { P.ops.push_back(so.op); P.names.push_back(so.name); }
        for (const Geo& g : geos) P.geos.push_back(SbcGeo{g.h, g.w, g.hy, g.hx, g.wp(), g.pps(), g.org(), ilog2(g.w)});
        P.blob = blob;
        P.arena_floats = peak; P.in_off = T[xin].off; P.out_off = T[out].off; P.post_off = T[post].off;
        P.max_w_len = max_w_len; P.park_floats = gtop; P.nthreads = nthreads; P.conv_flops = flops; P.x3 = x3;
        return P;
    }

  private:
    // ---- arena planner (program.py:_Planner) ----
    struct Tn { int size, born, died, off, pinned; int c, h, w; };
    std::vector<Tn> T;
    std::vector<int> pin_order;
    int peak = 0;
    int alloc(int n, int now, int died = -1) {
        T.push_back(Tn{(n + 3) / 4 * 4, now, died, -1, -1, 0, 0, 0});
        return (int)T.size() - 1;
    }
    void pin(int id, int off) { T[id].pinned = off; pin_order.push_back(id); }
    void free_(int id) {
        if (T[id].died >= 0) throw std::runtime_error("planner: double free");
        T[id].died = (int)ops.size();
    }
    bool live(int id) const { return id >= 0 && T[id].died < 0; }
    void solve(int end) {
        for (Tn& t : T) if (t.died < 0) t.died = end;
        std::vector<int> order;
        for (int i = 0; i < (int)T.size(); i++) if (T[i].pinned < 0) order.push_back(i);
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            if (T[a].size != T[b].size) return T[a].size > T[b].size;
            return T[a].born < T[b].born;
        });
        std::vector<int> placed;
        for (int id : pin_order) { T[id].off = T[id].pinned; placed.push_back(id); peak = std::max(peak, T[id].off + T[id].size); }
        for (int n : order) {
            std::vector<std::pair<int, int>> busy;
            for (int m : placed)
                if (T[m].born < T[n].died && T[n].born < T[m].died) busy.push_back({T[m].off, T[m].size});
            std::sort(busy.begin(), busy.end());
            int pos = 0;
            for (auto& ol : busy) {
                if (ol.first - pos >= T[n].size) break;
                pos = std::max(pos, ol.first + ol.second);
            }
            T[n].off = pos;
            placed.push_back(n);
            peak = std::max(peak, pos + T[n].size);
        }
    }
    int resolve(const Ref& r) const {
        if (r.kind == Ref::ARENA) return T[r.id].off + r.shift;
        if (r.kind == Ref::PARK) return goff[r.id] + r.shift;
        return -1;
    }

    // ---- tensors ----
    int gi(int h, int w) const {
        for (int i = 0; i < (int)geos.size(); i++) if (geos[i].h == h && geos[i].w == w) return i;
        throw std::runtime_error("no geometry for this size");
    }
    int tmp(int c, int h, int w) {
        const int id = alloc(geos[gi(h, w)].floats(c), (int)ops.size());
        T[id].c = c; T[id].h = h; T[id].w = w;
        return id;
    }
    int new_raw(int n) { return alloc(n, (int)ops.size()); }
    // ---- park area ----
    std::vector<int> goff;              // park slot -> float offset
    std::map<int, int> gslot;           // logical tensor id -> park slot; the logical id stays the handle of the stream
    std::map<int, int> alias;           // fill copies: arena tensor id -> logical id it mirrors (shape lookup only)
    int gtop = 0;
    int gnew(int n) { goff.push_back(gtop); gtop += n; return (int)goff.size() - 1; }
    bool is_big(int id) const { return park && geos[gi(T[id].h, T[id].w)].floats(T[id].c) * 4 >= 24 * 1024; }
    bool parked(int id) const { return gslot.count(id) != 0; }
    void spill(int id) {
        if (!live(id)) throw std::runtime_error("spill of a dead tensor");
        if (!parked(id)) gslot[id] = gnew(geos[gi(T[id].h, T[id].w)].floats(T[id].c));
        copy_op(SBC_OP_SPILL, Ref::arena(id), Ref::park(gslot[id]), id, "spill");
    }
    void park_out(int id) { spill(id); free_(id); }
    int fill(int id) {
        const int t = tmp(T[id].c, T[id].h, T[id].w);
        copy_op(SBC_OP_FILL, Ref::park(gslot.at(id)), Ref::arena(t), id, "fill");
        return t;
    }
    int smem_of(int id) { return live(id) ? id : fill(id); }
    void release(int id) { if (live(id)) free_(id); }
    void keep(int id) { if (park && id >= 0 && live(id)) park_out(id); }
    int back(int id) { return (park && id >= 0 && !live(id)) ? fill(id) : id; }

    SymOp blank(int kind, const std::string& name) {
        SymOp so;
        memset(&so.op, 0, sizeof so.op);
        so.op.kind = kind; so.op.src = so.op.dst = so.op.acc = so.op.edst = -1; so.op.dil = 1; so.op.b_rel = -1;
        so.op.ks = 1; so.op.scratch = -1; so.op.wbuf = -1; so.op.low = -1;
        so.name = name;
        return so;
    }
    void copy_op(int kind, Ref src, Ref dst, int like, const char* label) {
        const int c = T[like].c, h = T[like].h, w = T[like].w, g = gi(h, w);
        SymOp so = blank(kind, std::string(label));
        so.src = src; so.dst = dst;
        so.op.cin = so.op.cout = c; so.op.h = so.op.oh = h; so.op.w = so.op.ow = w; so.op.sgeo = so.op.dgeo = g;
        so.op.MT = geos[g].floats(c) / 4;
        ops.push_back(so);
    }
    void copy_raw(int kind, Ref src, Ref dst, int n, const char* label) {
        SymOp so = blank(kind, label);
        so.src = src; so.dst = dst; so.op.MT = n / 4;
        ops.push_back(so);
    }

    // ---- parameter blob ----
    const sbc2::TensorArg& param(const std::string& k) const {
        auto it = sd.find(k);
        if (it == sd.end()) throw std::runtime_error("state dict has no '" + k + "'");
        return it->second;
    }
    bool has(const std::string& k) const { return sd.find(k) != sd.end(); }
    // append arrays, each padded to a multiple of 4 floats; returns (offset, length); rels = start of each array
    void push(const std::vector<std::vector<float>>& arrs, int& off, int& len, std::vector<int>& rels) {
        off = (int)blob.size();
        int cur = 0;
        rels.clear();
        for (const auto& a : arrs) {
            rels.push_back(cur);
            blob.insert(blob.end(), a.begin(), a.end());
            const int pad = (4 - (int)(a.size() % 4)) % 4;
            blob.insert(blob.end(), pad, 0.f);
            cur += (int)a.size() + pad;
        }
        len = cur;
    }
    static float tf32_rna(float x) {
        uint32_t b;
        memcpy(&b, &x, 4);
        b = (b + 0x1000u) & 0xFFFFE000u;
        memcpy(&x, &b, 4);
        return x;
    }
    static int32_t fbits(double v) { const float f = (float)v; int32_t b; memcpy(&b, &f, 4); return b; }

    // ---- ops ----
    struct Sib { std::string prefix; int src; int dil; };
    void conv(const std::string& prefix, int src, int dst, int acc, int edst, int dil = 1, bool pool = false,
              bool compact = false, const std::vector<Sib>& siblings = {}, bool acc_g = false) {
        std::vector<Sib> br{{prefix, src, dil}};
        br.insert(br.end(), siblings.begin(), siblings.end());
        const int E = 2;
        const int h = T[src].h, w = T[src].w;
        const Geo& sg = geos[gi(h, w)];
        const int oh = pool ? h / 2 : h, ow = pool ? w / 2 : w;
        const sbc2::TensorArg& w0t = param(prefix + ".weight");
        const int cout = (int)w0t.shape[0];
        if (cout % 2) throw std::runtime_error("odd cout");
        const int NT = (cout + 7) / 8;
        std::vector<float> bias;
        std::vector<Branch> meta;
        std::vector<std::vector<int32_t>> aoffs;
        std::vector<std::vector<float>> wpads;   // [NT*8][KC*8][k*k]
        std::vector<int> kk2s, KC8s;
        int steps = 0;
        for (const Sib& b : br) {
            const sbc2::TensorArg& wt = param(b.prefix + ".weight");
            const int bco = (int)wt.shape[0], cin = (int)wt.shape[1], k = (int)wt.shape[2];
            if (wt.shape.size() != 4 || wt.shape[3] != k || (k != 1 && k != 3) || bco != cout) throw std::runtime_error("bad conv weight " + b.prefix);
            const int c = T[b.src].c;
            if (T[b.src].h != h || T[b.src].w != w) throw std::runtime_error("sibling geometry mismatch");
            if (!(c == cin || (c == 8 && cin < 8))) throw std::runtime_error("conv channel mismatch " + b.prefix);
            flops += 2LL * h * w * cin * k * k * cout;
            const int r = k / 2;
            Branch m;
            m.src = b.src; m.cin = cin; m.k = k; m.dil = b.dil; m.KC = (cin + 7) / 8; m.step0 = steps;
            for (int tap = 0; tap < k * k; tap++) {
                const int dy = (tap / k - r) * b.dil, dx = (tap % k - r) * b.dil;
                if (std::abs(dy) < h && std::abs(dx) < w) {
                    if (std::abs(dy) > sg.hy || std::abs(dx) > sg.hx) throw std::runtime_error("halo too small");
                    m.live.push_back(tap);
                }
            }
            std::vector<int32_t> ao(m.live.size() * m.KC);
            for (size_t i = 0; i < m.live.size(); i++) {
                const int tap = m.live[i];
                const int dy = (tap / k - r) * b.dil, dx = (tap % k - r) * b.dil;
                for (int kc = 0; kc < m.KC; kc++) ao[i * m.KC + kc] = (2 * kc * sg.pps() + dy * sg.wp() + dx) * 4;
            }
            const int kk = k * k, KC8 = m.KC * 8;
            std::vector<float> wp_((size_t)NT * 8 * KC8 * kk, 0.f);
            const float sc = pool ? 0.25f : 1.0f;
            for (int co = 0; co < cout; co++)
                for (int ci = 0; ci < cin; ci++)
                    for (int tp = 0; tp < kk; tp++)
                        wp_[((size_t)co * KC8 + ci) * kk + tp] = wt.data[((size_t)co * cin + ci) * kk + tp] * sc;
            if (has(b.prefix + ".bias")) {
                const float* bb = param(b.prefix + ".bias").data;
                if (bias.empty()) bias.assign(bb, bb + cout);
                else for (int i = 0; i < cout; i++) bias[i] = bias[i] + bb[i];
            }
            steps += (int)ao.size();
            meta.push_back(m); aoffs.push_back(ao); wpads.push_back(wp_); kk2s.push_back(kk); KC8s.push_back(KC8);
        }
        const int S = steps, S4 = (S + 3) / 4 * 4;
        std::vector<float> aoff(S4, 0.f);
        {
            int k = 0;
            for (const auto& ao : aoffs)
                for (int32_t v : ao) { memcpy(&aoff[k], &v, 4); k++; }
        }
        const int per_nt = S * 32 * E;
        int nt_chunk = NT;
        while (nt_chunk > 1 && S4 + per_nt * nt_chunk + 8 * nt_chunk > slot_floats) nt_chunk /= 2;
        const int nwarps = nthreads / 32;
        const int dgi = gi(oh, ow);
        const Geo& dgeo = geos[dgi];
        int tapmask = 0;
        for (int tp : meta[0].live) tapmask |= 1 << tp;
        for (int nt0 = 0; nt0 < NT; nt0 += nt_chunk) {
            const int ntc = std::min(nt_chunk, NT - nt0);
            const int co0 = nt0 * 8, co1 = std::min(cout, (nt0 + ntc) * 8);
            std::vector<float> frag((size_t)S * ntc * 32 * E, 0.f);
            for (size_t bi = 0; bi < meta.size(); bi++) {
                const Branch& m = meta[bi];
                const std::vector<float>& wp_ = wpads[bi];
                const int kk = kk2s[bi], KC8 = KC8s[bi];
                for (size_t i = 0; i < m.live.size(); i++)
                    for (int kc = 0; kc < m.KC; kc++)
                        for (int nt = 0; nt < ntc; nt++)
                            for (int lane = 0; lane < 32; lane++) {
                                const int g = lane >> 2, t = lane & 3, tap = m.live[i];
                                float w0 = wp_[((size_t)((nt0 + nt) * 8 + g) * KC8 + kc * 8 + t) * kk + tap];
                                float w1 = wp_[((size_t)((nt0 + nt) * 8 + g) * KC8 + kc * 8 + t + 4) * kk + tap];
                                if (!x3) { w0 = tf32_rna(w0); w1 = tf32_rna(w1); }
                                const int st = m.step0 + (int)i * m.KC + kc;
                                float* f = &frag[(((size_t)st * ntc + nt) * 32 + lane) * E];
                                f[0] = w0; f[1] = w1;
                            }
            }
            std::vector<std::vector<float>> arrs{aoff, frag};
            if (!bias.empty()) arrs.push_back(std::vector<float>(bias.begin() + co0, bias.begin() + co1));
            int w_off, w_len;
            std::vector<int> rels;
            push(arrs, w_off, w_len, rels);
            const int MT = (oh * ow + 15) / 16, units = MT * ntc;
            int ks = 1;
            const bool unit = units * 2 <= nwarps;
            if (unit)
                while (units * ks * 2 <= std::min(nwarps, 16) && ks * 2 <= S) ks *= 2;
            int scratch = -1;
            if (ks > 1) scratch = new_raw(units * ks * 32 * 4);
            int flags = (pool ? SBC_F_POOL : 0) | (x3 ? SBC_F_X3 : 0) | (compact ? SBC_F_COMPACT : 0) | (unit ? SBC_F_UNIT : 0);
            if (acc_g) {
                if (acc < 0 || !parked(acc) || live(acc)) throw std::runtime_error("acc_g: the stream must live in the park area only");
                flags |= SBC_F_ACC_G;
            }
            if (compact && !(nt_chunk == NT && cout == 2 && acc < 0 && edst < 0)) throw std::runtime_error("bad compact conv");
            const int sh = compact ? 0 : (co0 / 4) * dgeo.pps() * 4;
            std::string nm;
            for (size_t bi = 0; bi < br.size(); bi++) nm += (bi ? "+" : "") + br[bi].prefix;
            if (nt_chunk != NT) nm += "[co" + std::to_string(co0) + ":" + std::to_string(co1) + "]";
            SymOp so = blank(SBC_OP_CONV_MMA, nm);
            so.op.flags = flags;
            so.src = Ref::arena(src);
            so.dst = dst >= 0 ? Ref::arena(dst, sh) : Ref::none();
            so.acc = acc >= 0 ? (acc_g ? Ref::park(gslot.at(acc), sh) : Ref::arena(acc, sh)) : Ref::none();
            so.edst = edst >= 0 ? Ref::arena(edst, sh) : Ref::none();
            so.op.cin = meta[0].cin; so.op.cout = co1 - co0; so.op.h = h; so.op.w = w; so.op.ksize = meta[0].k; so.op.dil = dil;
            so.op.w_off = w_off; so.op.w_len = w_len; so.op.b_rel = bias.empty() ? -1 : rels[2];
            so.op.sgeo = gi(h, w); so.op.dgeo = dgi; so.op.ks = ks;
            so.scratch = scratch >= 0 ? Ref::arena(scratch) : Ref::none();
            so.op.oh = oh; so.op.ow = ow; so.op.tapmask = tapmask; so.op.MT = MT; so.op.NT = ntc; so.op.S = S;
            so.op.frag_rel = rels[1]; so.op.low = ilog2(ow);
            so.branches = meta;
            ops.push_back(so);
            if (scratch >= 0) free_(scratch);
        }
    }

    void quad_threads(int c, int hw, int& Tq, int& lT, int& npass) const {
        const int nq = c / 4;
        Tq = 32;
        while (Tq * 2 <= nthreads / nq && hw > 64) Tq *= 2;
        const int gpp = nthreads / Tq;
        lT = ilog2(Tq);
        npass = (nq + gpp - 1) / gpp;
    }
    void norm_elu(const std::string& prefix, int src, int dst) {
        const int c = T[src].c, h = T[src].h, w = T[src].w;
        std::vector<float> p;
        for (const char* k : {".alpha", ".gamma", ".beta"}) {
            const sbc2::TensorArg& a = param(prefix + k);
            p.insert(p.end(), a.data, a.data + c);
        }
        int w_off, w_len;
        std::vector<int> rels;
        push({p}, w_off, w_len, rels);
        const int scratch = new_raw(2 * (nthreads / 32) * 4 + 2 * c);
        int Tq, lT, npass;
        quad_threads(c, h * w, Tq, lT, npass);
        SymOp so = blank(SBC_OP_NORM_ELU, prefix);
        so.src = Ref::arena(src); so.dst = Ref::arena(dst);
        so.op.cin = so.op.cout = c; so.op.h = so.op.oh = h; so.op.w = so.op.ow = w; so.op.w_off = w_off; so.op.w_len = w_len;
        so.op.sgeo = so.op.dgeo = gi(h, w); so.scratch = Ref::arena(scratch);
        so.op.MT = Tq; so.op.NT = lT; so.op.S = npass;
        so.op.frag_rel = fbits(1.0 / (h * w)); so.op.low = fbits(1.0 / c); so.op.tapmask = fbits(1.0 / (c - 1));
        ops.push_back(so);
        free_(scratch);
    }
    void elu(int src, int dst) {
        const int c = T[src].c, h = T[src].h, w = T[src].w;
        int Tq, lT, npass;
        quad_threads(c, 1 << 30, Tq, lT, npass);
        SymOp so = blank(SBC_OP_ELU, "elu");
        so.src = Ref::arena(src); so.dst = Ref::arena(dst);
        so.op.cin = so.op.cout = c; so.op.h = so.op.oh = h; so.op.w = so.op.ow = w; so.op.sgeo = so.op.dgeo = gi(h, w);
        so.op.MT = Tq; so.op.NT = lT; so.op.S = npass;
        ops.push_back(so);
    }
    void affine(int src_raw, int dst) {
        const int c = T[dst].c, h = T[dst].h, w = T[dst].w;
        if (c != 8 || channels != 2) throw std::runtime_error("affine expects 2 real channels in an 8-channel chunk");
        SymOp so = blank(SBC_OP_AFFINE, "2x-1");
        so.src = Ref::arena(src_raw); so.dst = Ref::arena(dst);
        so.op.cin = channels; so.op.cout = c; so.op.h = so.op.oh = h; so.op.w = so.op.ow = w; so.op.sgeo = so.op.dgeo = gi(h, w);
        ops.push_back(so);
    }
    void maxpool5(int src, int dst) {
        const int c = T[src].c, h = T[src].h, w = T[src].w;
        SymOp so = blank(SBC_OP_MAXPOOL5, "maxpool5");
        so.src = Ref::arena(src); so.dst = Ref::arena(dst);
        so.op.cin = so.op.cout = c; so.op.h = so.op.oh = h; so.op.w = so.op.ow = w; so.op.sgeo = so.op.dgeo = gi(h, w);
        ops.push_back(so);
    }
    void upacc(int src, int acc, int edst) {
        const int c = T[src].c, h = T[src].h, w = T[src].w, oh = T[acc].h, ow = T[acc].w;
        SymOp so = blank(SBC_OP_UPACC, "upacc");
        so.src = Ref::arena(src); so.acc = Ref::arena(acc); so.edst = edst >= 0 ? Ref::arena(edst) : Ref::none();
        so.op.cin = so.op.cout = c; so.op.h = h; so.op.w = w; so.op.sgeo = gi(h, w); so.op.dgeo = gi(oh, ow);
        so.op.oh = oh; so.op.ow = ow;
        ops.push_back(so);
    }

    // ---- blocks (program.py: residual / rcu / crp / refine / _stage) ----
    int residual(const std::string& p, int x, int cout, int elu_out, int dil = 0) {
        const int cin = T[x].c, h = T[x].h, w = T[x].w, d = dil ? dil : 1;
        if (cout != cin) throw std::runtime_error("residual: only same-width blocks are reached by this network");
        if (is_big(x)) {
            if (!parked(x)) spill(x);
            const int s_ = smem_of(x);
            const int t = tmp(cin, h, w);
            norm_elu(p + ".normalize1", s_, t);
            free_(s_);
            const int t2 = tmp(cout, h, w);
            conv(p + ".conv1", t, t2, -1, -1, d);
            free_(t);
            const int t3 = tmp(cout, h, w);
            norm_elu(p + ".normalize2", t2, t3);
            free_(t2);
            conv(p + ".conv2", t3, -1, x, elu_out, d, false, false, {}, true);
            free_(t3);
            return x;
        }
        const int t = tmp(cin, h, w);
        norm_elu(p + ".normalize1", x, t);
        const int t2 = tmp(cout, h, w);
        conv(p + ".conv1", t, t2, -1, -1, d);
        free_(t);
        const int t3 = tmp(cout, h, w);
        norm_elu(p + ".normalize2", t2, t3);
        free_(t2);
        conv(p + ".conv2", t3, -1, x, elu_out, d);
        free_(t3);
        return x;
    }
    void stage(const std::string& p, int skip, int cout, int dil, bool want_elu, int& out_, int& e_out) {
        const int cin = T[skip].c, h = T[skip].h, w = T[skip].w, d = dil ? dil : 1;
        const bool big = is_big(skip);
        if (big && !parked(skip)) spill(skip);
        int s_ = big ? smem_of(skip) : skip;
        const int t = tmp(cin, h, w);
        norm_elu(p + ".0.normalize1", s_, t);
        if (big) free_(s_);
        const int t2 = tmp(cin, h, w);
        conv(p + ".0.conv1", t, t2, -1, -1, d);
        free_(t);
        const int t3 = tmp(cin, h, w);
        norm_elu(p + ".0.normalize2", t2, t3);
        free_(t2);
        s_ = big ? smem_of(skip) : skip;
        int out;
        if (!dil) {
            out = tmp(cout, h / 2, w / 2);
            conv(p + ".0.conv2.conv", t3, out, -1, -1, 1, true, false, {{p + ".0.shortcut.conv", s_, 1}});
        } else {
            out = tmp(cout, h, w);
            conv(p + ".0.conv2", t3, out, -1, -1, d, false, false, {{p + ".0.shortcut", s_, d}});
        }
        if (big) free_(s_);
        free_(t3);
        int e = -1;
        if (want_elu) e = tmp(T[out].c, T[out].h, T[out].w);
        out_ = residual(p + ".1", out, cout, e, dil);
        e_out = e;
    }
    void rcu(const std::string& p, int x, int n_blocks, int e_in, bool want_elu_out, int& x_out, int& e_out) {
        const int c = T[x].c, h = T[x].h, w = T[x].w;
        int e = e_in;
        const bool big = is_big(x);
        if (big) {
            if (!parked(x)) spill(x);
            if (e >= 0) release(x);
        }
        for (int i = 0; i < n_blocks; i++) {
            if (e < 0) {
                e = tmp(c, h, w);
                if (big) { const int s_ = smem_of(x); elu(s_, e); free_(s_); }
                else elu(x, e);
            }
            const int u = tmp(c, h, w);
            const std::string b = p + "." + std::to_string(i + 1);
            conv(b + "_1_conv", e, -1, -1, u);
            const bool last = (i == n_blocks - 1);
            if (last && !want_elu_out) {
                free_(e);
                e = -1;
                conv(b + "_2_conv", u, -1, x, -1, 1, false, false, {}, big);
            } else {
                conv(b + "_2_conv", u, -1, x, e, 1, false, false, {}, big);
            }
            free_(u);
        }
        x_out = x; e_out = e;
    }
    void crp(const std::string& p, int x, int e, int& sum_out, int& e2_out) {
        const int c = T[e].c, h = T[e].h, w = T[e].w;
        release(x);
        const bool big = is_big(e);
        const int m = tmp(c, h, w);
        maxpool5(e, m);
        if (big) park_out(e);
        const int path = tmp(c, h, w);
        conv(p + ".convs.0", m, path, e, -1, 1, false, false, {}, big);
        maxpool5(path, m);
        free_(path);
        const int e2 = tmp(c, h, w);
        conv(p + ".convs.1", m, -1, e, e2, 1, false, false, {}, big);
        free_(m);
        sum_out = e; e2_out = e2;
    }
    void refine(const std::string& p, std::vector<int> xs, std::vector<int> es, int features, bool end, int& h_out, int& e_out) {
        std::vector<int> hs;
        int e_single = -1;
        for (size_t i = 0; i < xs.size(); i++) {
            int h_, e_;
            rcu(p + ".adapt_convs." + std::to_string(i), xs[i], 2, es[i], xs.size() == 1, h_, e_);
            hs.push_back(h_);
            e_single = e_;
        }
        const int oh = T[hs[0]].h, ow = T[hs[0]].w;
        int s, e;
        if (hs.size() > 1) {
            s = tmp(features, oh, ow);
            const bool same = T[hs[1]].h == oh && T[hs[1]].w == ow;
            if (same) {
                e = tmp(features, oh, ow);
                const int a0 = smem_of(hs[0]), a1 = smem_of(hs[1]);
                conv(p + ".msf.convs.0", a0, s, -1, e, 1, false, false, {{p + ".msf.convs.1", a1, 1}});
                free_(a0);
                free_(a1);
            } else {
                const int a0 = smem_of(hs[0]);
                conv(p + ".msf.convs.0", a0, s, -1, -1);
                free_(a0);
                const int lo = tmp(features, T[hs[1]].h, T[hs[1]].w);
                const int a1 = smem_of(hs[1]);
                conv(p + ".msf.convs.1", a1, lo, -1, -1);
                free_(a1);
                e = tmp(features, oh, ow);
                upacc(lo, s, e);
                free_(lo);
            }
        } else {
            s = hs[0];
            e = e_single;
        }
        int h_, eh;
        crp(p + ".crp", s, e, h_, eh);
        rcu(p + ".output_convs", h_, end ? 3 : 1, eh, !end, h_out, e_out);
    }

    // ---- program.py:_halo_analysis ----
    void halo_analysis(int arena_floats, const std::vector<std::pair<int, int>>& raw_regions) {
        std::vector<int64_t> tags((size_t)arena_floats, -1);
        auto raw = [&](int off, int n) {
            if (off >= 0 && n > 0) std::fill(tags.begin() + off, tags.begin() + off + n, (int64_t)-2);
        };
        auto fresh = [&](int off, int g, int c) {
            const Geo& G = geos[g];
            bool need = false;
            for (int pl = 0; pl < (c + 3) / 4; pl++) {
                const int a = off + pl * G.pps() * 4;
                const int64_t key = ((int64_t)g << 40) | (int64_t)a;
                for (int i = a; i < a + G.pps() * 4; i++) {
                    if (tags[i] != key) need = true;
                    tags[i] = key;
                }
            }
            return need;
        };
        for (auto& r : raw_regions) raw(r.first, r.second);
        const int nwarps = nthreads / 32;
        for (SymOp& so : ops) {
            SbcOp& op = so.op;
            raw(op.wbuf, op.w_len);
            switch (op.kind) {
                case SBC_OP_CONV_MMA:
                    raw(op.scratch, op.ks > 1 ? op.MT * op.NT * op.ks * 32 * 4 : 0);
                    if (op.flags & SBC_F_COMPACT) raw(op.dst, channels * op.oh * op.ow);
                    else if (op.dst >= 0 && fresh(op.dst, op.dgeo, op.cout)) op.flags |= SBC_F_ZH_DST;
                    if (op.edst >= 0 && fresh(op.edst, op.dgeo, op.cout)) op.flags |= SBC_F_ZH_EDST;
                    break;
                case SBC_OP_NORM_ELU:
                    raw(op.scratch, 2 * nwarps * 4 + 2 * op.cin);
                    if (fresh(op.dst, op.dgeo, op.cin)) op.flags |= SBC_F_ZH_DST;
                    break;
                case SBC_OP_ELU:
                case SBC_OP_MAXPOOL5:
                    if (fresh(op.dst, op.dgeo, op.cin)) op.flags |= SBC_F_ZH_DST;
                    break;
                case SBC_OP_AFFINE:
                    if (fresh(op.dst, op.dgeo, op.cout)) op.flags |= SBC_F_ZH_DST;
                    break;
                case SBC_OP_UPACC:
                    if (op.edst >= 0 && fresh(op.edst, op.dgeo, op.cin)) op.flags |= SBC_F_ZH_EDST;
                    break;
                case SBC_OP_SPILL:
                    break;
                case SBC_OP_FILL:
                    if (op.cin == 0) raw(op.dst, 4 * op.MT);
                    else fresh(op.dst, op.dgeo, op.cin);
                    break;
                default:
                    throw std::runtime_error("halo analysis: unknown op kind");
            }
        }
    }

    const sbc2::StateDict& sd;
    int ngf, H, W, channels, nthreads;
    bool x3, park;
    int late_floats = 1 << 30;
    int slot_floats = 9300;
    std::vector<Geo> geos;
    std::vector<SymOp> ops;
    std::vector<float> blob;
    long long flops = 0;
};

// program.py:build_program with park=None: the two-CTAs-per-SM plan when it fits half an SM, else the plain plan when it
// fits one SM, else the park plan when THAT fits one SM, else the plain plan (global-memory arena).
inline Plan build_auto(const sbc2::StateDict& sd, int ngf, int H, int W, int channels, int nthreads, bool x3, int park = -1) {
    if (park >= 0) return Builder(sd, ngf, H, W, channels, nthreads, x3, park != 0).build();
    Plan pp = Builder(sd, ngf, H, W, channels, nthreads, x3, true).build();
    if (pp.smem_bytes() <= smem_budget_bytes(2)) return pp;
    Plan p1 = Builder(sd, ngf, H, W, channels, nthreads, x3, false).build();
    if (p1.smem_bytes() > smem_budget_bytes(1) && pp.smem_bytes() <= smem_budget_bytes(1)) return pp;
    return p1;
}

}  // namespace sbc1
