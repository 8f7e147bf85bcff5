"""Host-side "layer program" for the fused NCSNv2Deepest forward.

The CUDA kernel (``csrc/sbc_kernel.cuh``) keeps every activation of one sample
in a single on-chip arena and walks a flat list of ops.  This module builds that
list, plans the arena (which tensor lives at which float offset, with liveness
based reuse) and packs the checkpoint tensors into the parameter blob the
kernel streams per op.

The schedule restates ``NCSNv2Deepest.forward`` (reference
``ncsnv2/models/ncsnv2.py:269-300``) with its building blocks
``ResidualBlock`` (``ncsnv2/models/layers.py:443-456``), ``ConvMeanPool``
(``layers.py:309-313``), ``RefineBlock`` (``layers.py:234-249``), ``RCUBlock``
(``layers.py:126-134``), ``CRPBlock`` (``layers.py:76-83``), ``MSFBlock``
(``layers.py:178-184``) and ``InstanceNorm2dPlus``
(``ncsnv2/models/normalization.py:163-176``).  Fusions applied (all exact
re-associations of the reference arithmetic):

* ``ELU`` is folded into the epilogue of whichever op produces its input
  (``edst``), so no stand-alone activation pass remains except after skips;
* residual / multi-scale sums are epilogue accumulations (``acc``);
* ``ConvMeanPool`` = conv then 2x2 mean-pool is evaluated as a conv over the
  2x2 box-summed input sampled at stride 2, times 1/4 (same linear map, 4x
  fewer MACs) -- ``F_POOL``.

Nothing here touches a GPU; ``simulate`` is a torch (CPU) interpreter of the
program used by the CPU test-suite to pin the schedule against the reference
module before any CUDA code runs.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import numpy as np

# ---- op kinds (mirrored in csrc/sbc_program.h) ---------------------------
OP_AFFINE = 0     # dst = 2*src - 1                       (ncsnv2.py:270-271)
OP_CONV = 1       # k in {1,3} conv, stride 1, pad = dil*(k//2)
OP_NORM_ELU = 2   # dst = ELU(InstanceNorm++(src))
OP_ELU = 3        # dst = ELU(src)
OP_MAXPOOL5 = 4   # dst = maxpool 5x5 stride 1 pad 2
OP_UPACC = 5      # acc += bilinear(src -> (oh,ow), align_corners=True)
OP_CONV_MMA = 6   # same contract as OP_CONV, contraction on tensor cores (mma.sync m16n8k8 TF32)
N_OP_KINDS = 7

PLANE_PAD = 8     # every activation plane is stored with stride h*w + 8 floats: with the stride
                  # = 8 (mod 16) the (channel t, pixel g) gather of an MMA A-fragment hits 32 banks


def PS(h: int, w: int) -> int:
    return h * w + PLANE_PAD


PRECISIONS = ("fp32", "tf32x3", "tf32")

# ---- flags ----------------------------------------------------------------
F_POOL = 1        # conv followed by 2x2 mean-pool (ConvMeanPool)
F_X3 = 2          # OP_CONV_MMA: 3xTF32 error-compensated product (fp32-equivalent accuracy)

OP_FIELDS = ("kind", "flags", "src", "dst", "acc", "edst", "cin", "cout",
             "h", "w", "ksize", "dil", "w_off", "w_len", "b_rel", "px",
             "cb", "ks", "scratch", "oh", "ow", "pad0", "tapmask", "pad2")
OP_WORDS = len(OP_FIELDS)            # 24 int32 = 96 bytes per op
assert OP_WORDS == 24


@dataclass
class Op:
    kind: int
    flags: int = 0
    src: int = -1       # float offset into the arena
    dst: int = -1       # raw result store           (-1: none)
    acc: int = -1       # v += acc[i]; acc[i] = v    (-1: none)
    edst: int = -1      # edst[i] = ELU(v)           (-1: none)
    cin: int = 0
    cout: int = 0
    h: int = 0          # INPUT spatial size
    w: int = 0
    ksize: int = 0
    dil: int = 1
    w_off: int = 0      # float offset of this op's parameter segment in the blob
    w_len: int = 0      # floats (multiple of 4) streamed for this op
    b_rel: int = -1     # bias offset inside the segment (-1: no bias)
    px: int = 1         # conv tiling: output pixels per thread along W
    cb: int = 1         # conv tiling: output channels per thread
    ks: int = 1         # conv tiling: split of the Cin loop across threads
    scratch: int = -1   # arena offset of the per-op scratch (K-split partials / norm stats)
    oh: int = 0         # OUTPUT spatial size
    ow: int = 0
    tapmask: int = 0    # OP_CONV_MMA: bit i set = tap i (row-major in the k x k window) can touch the image
    name: str = ""

    def words(self) -> List[int]:
        vals = [getattr(self, f) if not f.startswith("pad") else 0 for f in OP_FIELDS]
        return [int(v) for v in vals]


@dataclass
class Program:
    ops: List[Op]
    arena_floats: int
    blob: np.ndarray                 # float32 parameter blob
    in_off: int                      # arena offset of the [2,H,W] network input (x, planar re/im)
    out_off: int                     # arena offset of the [2,H,W] raw network output (before /sigma)
    H: int
    W: int
    ngf: int
    channels: int
    nthreads: int
    max_w_len: int
    conv_flops: int                  # dense conv FLOPs / forward / sample (reference convention)
    precision: str = "fp32"
    post_off: int = 0                # arena offset of the post-network scratch (2*H*W + 4*nthreads floats)

    def op_table(self) -> np.ndarray:
        t = np.zeros((len(self.ops), OP_WORDS), dtype=np.int32)
        for i, op in enumerate(self.ops):
            t[i] = op.words()
        return t


class _Planner:
    """Offline arena planner.

    Tensors are registered with their live interval in units of emitted ops
    ([born, died)); ``solve`` places them greedily by decreasing size at the lowest
    offset that does not collide with an already placed tensor whose interval overlaps."""

    def __init__(self, align: int = 4):
        self.align = align
        self.born: Dict[str, int] = {}
        self.died: Dict[str, int] = {}
        self.size: Dict[str, int] = {}
        self.offs: Dict[str, int] = {}
        self.peak = 0

    def alloc(self, name: str, n: int, now: int) -> None:
        assert name not in self.size, name
        self.size[name] = (n + self.align - 1) // self.align * self.align
        self.born[name] = now

    def free(self, name: str, now: int) -> None:
        assert name in self.size and name not in self.died, name
        self.died[name] = now

    def solve(self, end: int) -> None:
        for n in self.size:
            self.died.setdefault(n, end)
        order = sorted(self.size, key=lambda n: (-self.size[n], self.born[n]))
        placed: List[str] = []
        for n in order:
            busy = sorted((self.offs[m], self.size[m]) for m in placed
                          if self.born[m] < self.died[n] and self.born[n] < self.died[m])
            pos = 0
            for (o, l) in busy:
                if o - pos >= self.size[n]:
                    break
                pos = max(pos, o + l)
            self.offs[n] = pos
            placed.append(n)
            self.peak = max(self.peak, pos + self.size[n])


class ProgramBuilder:
    """Builds the op list for one (ngf, H, W) instance of NCSNv2Deepest."""

    def __init__(self, state: Dict[str, np.ndarray], ngf: int, H: int, W: int,
                 channels: int = 2, nthreads: int = 512, precision: str = "fp32"):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (PRECISIONS,))
        self.precision = precision
        if H <= 0 or W <= 0 or H % 8 or W % 8:
            raise ValueError("Nt and Nr must be positive multiples of 8 (three 2x mean-pools), got %dx%d" % (H, W))
        self.sd = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
        self.ngf, self.H, self.W, self.channels = ngf, H, W, channels
        self.nthreads = nthreads
        self.ar = _Planner()
        self.ops: List[Op] = []
        self.blob: List[np.ndarray] = []
        self.blob_len = 0
        self.shape: Dict[str, Tuple[int, int, int]] = {}
        self.flops = 0
        self._tmp = 0

    # -- tensors ----------------------------------------------------------
    def new(self, name: str, c: int, h: int, w: int) -> str:
        self.ar.alloc(name, c * PS(h, w), len(self.ops))
        self.shape[name] = (c, h, w)
        return name

    def tmp(self, c: int, h: int, w: int, tag: str = "t") -> str:
        self._tmp += 1
        return self.new("%s%d" % (tag, self._tmp), c, h, w)

    def free(self, *names: str) -> None:
        for n in names:
            self.ar.free(n, len(self.ops))

    def off(self, name: Optional[str]):
        """Tensor reference; resolved to an arena offset by ``build`` once the plan is solved."""
        return name

    # -- parameter blob ---------------------------------------------------
    def _push(self, arrs: List[np.ndarray]) -> Tuple[int, int, List[int]]:
        off = self.blob_len
        rels, cur = [], 0
        for a in arrs:
            a = np.ascontiguousarray(a, dtype=np.float32).ravel()
            rels.append(cur)
            pad = (-a.size) % 4
            self.blob.append(a)
            if pad:
                self.blob.append(np.zeros(pad, np.float32))
            cur += a.size + pad
        self.blob_len += cur
        return off, cur, rels

    # -- ops --------------------------------------------------------------
    def conv(self, prefix: str, src: str, dst: Optional[str] = None, acc: Optional[str] = None,
             edst: Optional[str] = None, dil: int = 1, pool: bool = False) -> None:
        """One Conv2d (reference ``layers.py:28-60``) + fused epilogue.

        v = conv(src) + bias;  dst <- v;  acc <- (v += acc);  edst <- ELU(v)."""
        wt = self.sd[prefix + ".weight"]
        bias = self.sd.get(prefix + ".bias")
        cout, cin, k, k2 = wt.shape
        assert k == k2 and k in (1, 3)
        c, h, w = self.shape[src]
        assert c == cin, (prefix, c, cin)
        oh, ow = (h // 2, w // 2) if pool else (h, w)
        for t in (dst, acc, edst):
            if t is not None:
                assert self.shape[t] == (cout, oh, ow), (prefix, t, self.shape[t], (cout, oh, ow))
        # dense FLOP count, reference convention (conv evaluated at input resolution)
        self.flops += 2 * h * w * cin * k * k * cout
        if self.precision != "fp32":
            self._conv_mma(prefix, wt, bias, src, dst, acc, edst, dil, pool, (cin, h, w), (cout, oh, ow))
            return
        # ---- tiling choice (see csrc/sbc_ops.h: conv_partial) ----
        cb = 8 if cout % 8 == 0 else (4 if cout % 4 == 0 else (2 if cout % 2 == 0 else 1))
        px = 4 if ow % 4 == 0 else (2 if ow % 2 == 0 else 1)
        items = oh * (ow // px) * (cout // cb)
        ks = 1
        while items * ks * 2 <= self.nthreads and cin % (ks * 2) == 0 and ks < 32:
            ks *= 2
        # pack weights [cout/cb][cin][k*k][cb]
        wp = wt.reshape(cout // cb, cb, cin, k * k).transpose(0, 2, 3, 1)
        if pool:
            wp = wp * np.float32(0.25)
        arrs = [wp] + ([bias] if bias is not None else [])
        w_off, w_len, rels = self._push(arrs)
        # K-split partial sums are combined with warp shuffles (ks adjacent lanes): no scratch
        self.ops.append(Op(OP_CONV, F_POOL if pool else 0, self.off(src), self.off(dst), self.off(acc),
                           self.off(edst), cin, cout, h, w, k, dil, w_off, w_len,
                           rels[1] if bias is not None else -1, px, cb, ks,
                           -1, oh, ow, name=prefix))

    # largest parameter segment (floats) that one slot of the shared-memory ring holds; MMA convs whose
    # fragment array is bigger are split into cout chunks (each chunk is its own op)
    SLOT_FLOATS = 9248

    def _conv_mma(self, prefix, wt, bias, src, dst, acc, edst, dil, pool, ishape, oshape) -> None:
        """Tensor-core conv: implicit GEMM  D[16 pixels, 8 couts] += A[16 pixels, 8 cins] . B[8 cins, 8 couts]
        per (tap, cin chunk) with mma.sync.m16n8k8 TF32 (csrc/sbc_mma.h).  The weights are packed in
        B-fragment order: frag[step][ntile][lane] = (hi0, hi1[, lo0, lo1]) where lane = 4*g + t holds
        B[k = t (+4)][n = g], step = live_tap_index * KC + cin_chunk; hi = TF32(w), lo = TF32(w - hi)
        (lo only in the 3xTF32 mode)."""
        cin, h, w = ishape
        cout, oh, ow = oshape
        k = wt.shape[2]
        r = k // 2
        x3 = self.precision == "tf32x3"
        E = 4 if x3 else 2                          # floats per lane per fragment
        live = []
        for tap in range(k * k):
            dy, dx = (tap // k - r) * dil, (tap % k - r) * dil
            if abs(dy) < h and abs(dx) < w:          # otherwise the tap only ever reads zero padding
                live.append(tap)
        tapmask = sum(1 << t for t in live)
        KC, NT = (cin + 7) // 8, (cout + 7) // 8
        per_nt = len(live) * KC * 32 * E
        nt_chunk = NT
        while nt_chunk > 1 and per_nt * nt_chunk + 8 * nt_chunk > self.SLOT_FLOATS:
            nt_chunk //= 2
        wpad = np.zeros((NT * 8, KC * 8, k * k), np.float32)
        wpad[:cout, :cin] = wt.reshape(cout, cin, k * k) * (np.float32(0.25) if pool else np.float32(1.0))
        g, t = np.arange(32) >> 2, np.arange(32) & 3
        nwarps = self.nthreads // 32
        ps_out = PS(oh, ow)
        for nt0 in range(0, NT, nt_chunk):
            ntc = min(nt_chunk, NT - nt0)
            co0, co1 = nt0 * 8, min(cout, (nt0 + ntc) * 8)
            frag = np.zeros((len(live), KC, ntc, 32, E), np.float32)
            for i, tap in enumerate(live):
                for kc in range(KC):
                    for nt in range(ntc):
                        w0 = wpad[(nt0 + nt) * 8 + g, kc * 8 + t, tap]
                        w1 = wpad[(nt0 + nt) * 8 + g, kc * 8 + t + 4, tap]
                        h0, h1 = tf32_rna(w0), tf32_rna(w1)
                        frag[i, kc, nt, :, 0], frag[i, kc, nt, :, 1] = h0, h1
                        if x3:
                            frag[i, kc, nt, :, 2], frag[i, kc, nt, :, 3] = tf32_rna(w0 - h0), tf32_rna(w1 - h1)
            arrs = [frag] + ([bias[co0:co1]] if bias is not None else [])
            w_off, w_len, rels = self._push(arrs)
            MT = (oh * ow + 15) // 16
            units, S = MT * ntc, len(live) * KC
            ks = 1
            while units * ks * 2 <= nwarps and ks * 2 <= S:
                ks *= 2
            scratch = self.tmp(1, 1, nwarps * 32 * 4, "ksp") if ks > 1 else None
            flags = (F_POOL if pool else 0) | (F_X3 if x3 else 0)
            shift = lambda name: None if name is None else (name, co0 * ps_out)
            self.ops.append(Op(OP_CONV_MMA, flags, self.off(src), shift(dst), shift(acc), shift(edst),
                               cin, co1 - co0, h, w, k, dil, w_off, w_len, rels[1] if bias is not None else -1,
                               0, 0, ks, self.off(scratch), oh, ow, tapmask=tapmask,
                               name=prefix + ("" if nt_chunk == NT else "[co%d:%d]" % (co0, co1))))
            if scratch is not None:
                self.free(scratch)

    def norm_elu(self, prefix: str, src: str, dst: str) -> None:
        """dst = ELU(InstanceNorm2dPlus(src))  (``normalization.py:163-176`` + ``layers.py:13``)."""
        c, h, w = self.shape[src]
        assert self.shape[dst] == (c, h, w)
        w_off, w_len, _ = self._push([np.concatenate([self.sd[prefix + ".alpha"],
                                                      self.sd[prefix + ".gamma"],
                                                      self.sd[prefix + ".beta"]])])
        # scratch: per-channel (mean, rstd)
        scratch = self.tmp(1, 1, 2 * c, "nsc")
        self.ops.append(Op(OP_NORM_ELU, 0, self.off(src), self.off(dst), cin=c, cout=c, h=h, w=w,
                           w_off=w_off, w_len=w_len, scratch=self.off(scratch), oh=h, ow=w, name=prefix))
        self.free(scratch)

    def elu(self, src: str, dst: str) -> None:
        c, h, w = self.shape[src]
        self.ops.append(Op(OP_ELU, 0, self.off(src), self.off(dst), cin=c, cout=c, h=h, w=w, oh=h, ow=w,
                           name="elu"))

    def affine(self, src: str, dst: str) -> None:
        c, h, w = self.shape[src]
        self.ops.append(Op(OP_AFFINE, 0, self.off(src), self.off(dst), cin=c, cout=c, h=h, w=w, oh=h, ow=w,
                           name="2x-1"))

    def maxpool5(self, src: str, dst: str) -> None:
        c, h, w = self.shape[src]
        self.ops.append(Op(OP_MAXPOOL5, 0, self.off(src), self.off(dst), cin=c, cout=c, h=h, w=w, oh=h, ow=w,
                           name="maxpool5"))

    def upacc(self, src: str, acc: str, edst: Optional[str] = None) -> None:
        """acc += bilinear(src, size=acc.shape, align_corners=True) (``layers.py:182-183``)."""
        c, h, w = self.shape[src]
        c2, oh, ow = self.shape[acc]
        assert c == c2
        self.ops.append(Op(OP_UPACC, 0, self.off(src), -1, self.off(acc), self.off(edst), cin=c, cout=c,
                           h=h, w=w, oh=oh, ow=ow, name="upacc"))

    # -- blocks -----------------------------------------------------------
    def residual(self, p: str, x: str, cout: int, down: bool, dil: Optional[int]) -> str:
        """ResidualBlock.forward (``layers.py:443-456``). Consumes ``x``; returns the output tensor."""
        cin, h, w = self.shape[x]
        d = dil or 1
        t = self.tmp(cin, h, w)
        self.norm_elu(p + ".normalize1", x, t)
        t2 = self.tmp(cin if down else cout, h, w)
        self.conv(p + ".conv1", t, dst=t2, dil=d)
        self.free(t)
        c1 = self.shape[t2][0]
        t3 = self.tmp(c1, h, w)
        self.norm_elu(p + ".normalize2", t2, t3)
        self.free(t2)
        if cout == cin and not down:
            self.conv(p + ".conv2", t3, acc=x, dil=d)           # shortcut = x
            self.free(t3)
            return x
        if down and dil is None:
            out = self.tmp(cout, h // 2, w // 2, "o")
            self.conv(p + ".conv2.conv", t3, dst=out, pool=True)  # ConvMeanPool 3x3
            self.free(t3)
            self.conv(p + ".shortcut.conv", x, acc=out, pool=True)  # ConvMeanPool 1x1
        else:
            out = self.tmp(cout, h, w, "o")
            self.conv(p + ".conv2", t3, dst=out, dil=d)
            self.free(t3)
            self.conv(p + ".shortcut", x, acc=out, dil=d)
        self.free(x)
        return out

    def rcu(self, p: str, x: str, n_blocks: int, e_in: Optional[str] = None,
            want_elu_out: bool = False) -> Tuple[str, Optional[str]]:
        """RCUBlock.forward (``layers.py:126-134``), in place on ``x``.

        ``e_in`` (optional) already holds ELU(x).  Returns (x, ELU(x) or None)."""
        c, h, w = self.shape[x]
        e = e_in
        for i in range(n_blocks):
            if e is None:
                e = self.tmp(c, h, w)
                self.elu(x, e)
            u = self.tmp(c, h, w)
            self.conv("%s.%d_1_conv" % (p, i + 1), e, edst=u)      # u = ELU(conv1(ELU(x)))
            last = (i == n_blocks - 1)
            if last and not want_elu_out:
                self.free(e)
                e = None
                self.conv("%s.%d_2_conv" % (p, i + 1), u, acc=x)
            else:
                self.conv("%s.%d_2_conv" % (p, i + 1), u, acc=x, edst=e)  # x += conv2(u); e = ELU(x)
            self.free(u)
        return x, e

    def crp(self, p: str, x: str, e: str) -> Tuple[str, str]:
        """CRPBlock.forward (``layers.py:76-83``).  ``e`` = ELU(x) becomes the running sum; x is dead.

        Returns (sum, ELU(sum)) -- the ELU feeds the output RCU that always follows."""
        c, h, w = self.shape[e]
        self.free(x)
        m = self.tmp(c, h, w)
        self.maxpool5(e, m)
        path = self.tmp(c, h, w)
        self.conv(p + ".convs.0", m, dst=path, acc=e)
        self.maxpool5(path, m)
        self.free(path)
        e2 = self.tmp(c, h, w)
        self.conv(p + ".convs.1", m, acc=e, edst=e2)
        self.free(m)
        return e, e2

    def refine(self, p: str, xs: List[str], es: List[Optional[str]], features: int,
               end: bool = False) -> Tuple[str, Optional[str]]:
        """RefineBlock.forward (``layers.py:234-249``).  ``es[i]`` optionally holds ELU(xs[i]).

        Returns (h, ELU(h)) (ELU(h) is None for the last block)."""
        hs = []
        for i, x in enumerate(xs):
            h_, _ = self.rcu("%s.adapt_convs.%d" % (p, i), x, 2, e_in=es[i])
            hs.append(h_)
        c0, oh, ow = self.shape[hs[0]]
        if len(hs) > 1:
            s = self.tmp(features, oh, ow, "s")
            e = self.tmp(features, oh, ow, "e")
            same = self.shape[hs[1]][1:] == (oh, ow)
            self.conv(p + ".msf.convs.0", hs[0], dst=s)
            self.free(hs[0])
            if same:   # bilinear to the same size with align_corners=True is the identity
                self.conv(p + ".msf.convs.1", hs[1], acc=s, edst=e)
                self.free(hs[1])
            else:
                _, lh, lw = self.shape[hs[1]]
                lo = self.tmp(features, lh, lw, "lo")
                self.conv(p + ".msf.convs.1", hs[1], dst=lo)
                self.free(hs[1])
                self.upacc(lo, s, edst=e)
                self.free(lo)
        else:
            s = hs[0]
            e = self.tmp(c0, oh, ow, "e")
            self.elu(s, e)
        h_, eh = self.crp(p + ".crp", s, e)
        return self.rcu(p + ".output_convs", h_, 3 if end else 1, e_in=eh, want_elu_out=not end)

    # -- whole network ----------------------------------------------------
    def build(self) -> Program:
        ngf, H, W = self.ngf, self.H, self.W
        xin = self.new("x_in", self.channels, H, W)
        a = self.tmp(self.channels, H, W)
        self.affine(xin, a)
        o = self.tmp(ngf, H, W, "o")
        self.conv("begin_conv", a, dst=o)
        self.free(a)
        l1 = self.residual("res1.1", self.residual("res1.0", o, ngf, False, None), ngf, False, None)
        # skip tensors stay live; each stage works on a copy-free continuation:
        # the first block of the next stage reads the skip (norm1 + shortcut) without modifying it.
        l2 = self._stage("res2", l1, 2 * ngf, None)
        l3 = self._stage("res3", l2, 2 * ngf, None)
        l31 = self._stage("res31", l3, 2 * ngf, None)
        l4 = self._stage("res4", l31, 4 * ngf, 2)
        l5 = self._stage("res5", l4, 4 * ngf, 4)
        r1, e1 = self.refine("refine1", [l5], [None], 4 * ngf)
        r2, e2 = self.refine("refine2", [l4, r1], [None, e1], 2 * ngf)
        r31, e31 = self.refine("refine31", [l31, r2], [None, e2], 2 * ngf)
        r3, e3 = self.refine("refine3", [l3, r31], [None, e31], 2 * ngf)
        r4, e4 = self.refine("refine4", [l2, r3], [None, e3], ngf)
        r5, _ = self.refine("refine5", [l1, r4], [None, e4], ngf, end=True)
        t = self.tmp(ngf, H, W)
        self.norm_elu("normalizer", r5, t)
        self.free(r5)
        out = self.new("net_out", self.channels, H, W)
        self.conv("end_conv", t, dst=out)
        self.free(t)
        # scratch for the sampler phases that follow the network (residual P*x-y, reductions)
        post = self.new("post", 1, 1, 2 * H * W + 4 * self.nthreads)   # one plane of 2*H*W + 4*nthr (+pad)
        blob = np.concatenate(self.blob) if self.blob else np.zeros(0, np.float32)
        assert blob.size == self.blob_len
        max_w_len = max(op.w_len for op in self.ops)
        # ---- solve the arena plan and resolve tensor names to float offsets ----
        self.ar.solve(len(self.ops) + 1)
        for op in self.ops:
            for f in ("src", "dst", "acc", "edst", "scratch"):
                v = getattr(op, f)
                if isinstance(v, tuple):
                    setattr(op, f, self.ar.offs[v[0]] + v[1])
                else:
                    setattr(op, f, -1 if (v is None or v == -1) else self.ar.offs[v])
        return Program(self.ops, self.ar.peak, blob, self.ar.offs[xin], self.ar.offs[out], H, W, ngf,
                       self.channels, self.nthreads, max_w_len, self.flops, precision=self.precision,
                       post_off=self.ar.offs[post])

    def _stage(self, p: str, skip: str, cout: int, dil: Optional[int]) -> str:
        """Two ResidualBlocks, the first 'down'; ``skip`` must survive (it feeds a RefineBlock later)."""
        cin, h, w = self.shape[skip]
        d = dil or 1
        # first block, written out so that it does not consume `skip`
        t = self.tmp(cin, h, w)
        self.norm_elu(p + ".0.normalize1", skip, t)
        t2 = self.tmp(cin, h, w)
        self.conv(p + ".0.conv1", t, dst=t2, dil=d)
        self.free(t)
        t3 = self.tmp(cin, h, w)
        self.norm_elu(p + ".0.normalize2", t2, t3)
        self.free(t2)
        if dil is None:
            out = self.tmp(cout, h // 2, w // 2, "o")
            self.conv(p + ".0.conv2.conv", t3, dst=out, pool=True)
            self.free(t3)
            self.conv(p + ".0.shortcut.conv", skip, acc=out, pool=True)
        else:
            out = self.tmp(cout, h, w, "o")
            self.conv(p + ".0.conv2", t3, dst=out, dil=d)
            self.free(t3)
            self.conv(p + ".0.shortcut", skip, acc=out, dil=d)
        return self.residual(p + ".1", out, cout, False, dil)


def tf32_rna(x: np.ndarray) -> np.ndarray:
    """fp32 -> TF32 (10-bit mantissa), round to nearest with ties away from zero: cvt.rna.tf32.f32."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def build_program(state: Dict[str, np.ndarray], ngf: int, H: int, W: int, channels: int = 2,
                  nthreads: int = 512, precision: str = "fp32") -> Program:
    return ProgramBuilder(state, ngf, H, W, channels, nthreads, precision).build()


# ---------------------------------------------------------------------------
# torch (CPU) interpreter of a Program -- host-side check of the schedule only
# ---------------------------------------------------------------------------
def tensor_view(arena, off: int, c: int, h: int, w: int):
    """[c,h,w] view of a planar tensor stored at float offset ``off`` with padded plane stride
    (works for numpy arrays and torch tensors)."""
    ps = PS(h, w)
    if isinstance(arena, np.ndarray):
        return np.lib.stride_tricks.as_strided(arena[off:], (c, h, w), (4 * ps, 4 * w, 4))
    return arena.as_strided((c, h, w), (ps, w, 1), off)


def conv_weights(prog: Program, op: Op):
    """Decode (weight [cout,cin,k,k], bias or None) of a conv op back from the packed blob."""
    import torch
    blob = torch.from_numpy(prog.blob)
    k = op.ksize
    if op.kind == OP_CONV:
        cb = op.cb
        nw = op.cout * op.cin * k * k
        wp = blob[op.w_off:op.w_off + nw].view(op.cout // cb, op.cin, k * k, cb)
        wt = wp.permute(0, 3, 1, 2).reshape(op.cout, op.cin, k, k)
    else:
        live = [t for t in range(k * k) if (op.tapmask >> t) & 1]
        KC, NT = (op.cin + 7) // 8, (op.cout + 7) // 8
        E = 4 if (op.flags & F_X3) else 2
        n = len(live) * KC * NT * 32 * E
        frag = blob[op.w_off:op.w_off + n].view(len(live), KC, NT, 32, E)
        full = torch.zeros(NT * 8, KC * 8, k * k)
        g, t = torch.arange(32) >> 2, torch.arange(32) & 3
        for i, tap in enumerate(live):
            for kc in range(KC):
                for nt in range(NT):
                    lo0 = frag[i, kc, nt, :, 2] if E == 4 else 0.0
                    lo1 = frag[i, kc, nt, :, 3] if E == 4 else 0.0
                    full[nt * 8 + g, kc * 8 + t, tap] = frag[i, kc, nt, :, 0] + lo0
                    full[nt * 8 + g, kc * 8 + t + 4, tap] = frag[i, kc, nt, :, 1] + lo1
        wt = full[:op.cout, :op.cin].reshape(op.cout, op.cin, k, k).contiguous()
    bias = blob[op.w_off + op.b_rel:op.w_off + op.b_rel + op.cout] if op.b_rel >= 0 else None
    return wt, bias


def simulate(prog: Program, x, upto: Optional[int] = None):
    """Run the program on one sample ``x`` ([channels,H,W] float32 torch tensor) with torch CPU ops
    (fp32 arithmetic whatever the program's precision mode).

    Returns (raw network output [channels,H,W] (before the /sigma of ncsnv2.py:295-298), arena)."""
    import torch
    import torch.nn.functional as F

    arena = torch.zeros(prog.arena_floats, dtype=torch.float32)
    blob = torch.from_numpy(prog.blob)

    def view(off, c, h, w):
        return tensor_view(arena, off, c, h, w)

    view(prog.in_off, prog.channels, prog.H, prog.W).copy_(x.float())
    for i, op in enumerate(prog.ops):
        if upto is not None and i >= upto:
            break
        if op.kind == OP_AFFINE:
            view(op.dst, op.cin, op.h, op.w).copy_(2 * view(op.src, op.cin, op.h, op.w) - 1.0)
        elif op.kind == OP_ELU:
            view(op.dst, op.cin, op.h, op.w).copy_(F.elu(view(op.src, op.cin, op.h, op.w)))
        elif op.kind == OP_MAXPOOL5:
            view(op.dst, op.cin, op.h, op.w).copy_(
                F.max_pool2d(view(op.src, op.cin, op.h, op.w)[None], 5, 1, 2)[0])
        elif op.kind == OP_NORM_ELU:
            c = op.cin
            xs = view(op.src, c, op.h, op.w)
            al, ga, be = blob[op.w_off:op.w_off + 3 * c].view(3, c)
            mu = xs.mean(dim=(1, 2))
            m, v = mu.mean(), mu.var()
            mh = (mu - m) / torch.sqrt(v + 1e-5)
            hn = (xs - mu[:, None, None]) / torch.sqrt(xs.var(dim=(1, 2), unbiased=False)[:, None, None] + 1e-5)
            o = ga[:, None, None] * (hn + (mh * al)[:, None, None]) + be[:, None, None]
            view(op.dst, c, op.h, op.w).copy_(F.elu(o))
        elif op.kind == OP_UPACC:
            s = view(op.src, op.cin, op.h, op.w)
            a = view(op.acc, op.cin, op.oh, op.ow)
            a += F.interpolate(s[None], size=(op.oh, op.ow), mode="bilinear", align_corners=True)[0]
            if op.edst >= 0:
                view(op.edst, op.cin, op.oh, op.ow).copy_(F.elu(a))
        elif op.kind in (OP_CONV, OP_CONV_MMA):
            k = op.ksize
            wt, bias = conv_weights(prog, op)
            s = view(op.src, op.cin, op.h, op.w)[None]
            if op.flags & F_POOL:
                # packed weights already carry the 1/4; conv at full res then 2x2 SUM
                v = F.conv2d(s, wt, None, 1, op.dil * (k // 2), op.dil)
                v = F.avg_pool2d(v, 2) * 4.0
                if bias is not None:
                    v = v + bias[None, :, None, None]
            else:
                v = F.conv2d(s, wt, bias, 1, op.dil * (k // 2), op.dil)
            v = v[0]
            if op.dst >= 0:
                view(op.dst, op.cout, op.oh, op.ow).copy_(v)
            if op.acc >= 0:
                a = view(op.acc, op.cout, op.oh, op.ow)
                a += v
                v = a
            if op.edst >= 0:
                view(op.edst, op.cout, op.oh, op.ow).copy_(F.elu(v))
        else:
            raise ValueError(op.kind)
    out = view(prog.out_off, prog.channels, prog.H, prog.W).clone()
    return out, arena
