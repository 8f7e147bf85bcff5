"""Host-side "layer program" for the fused NCSNv2Deepest forward.

The CUDA kernel (``csrc/sbc_kernel.cuh``) keeps every activation of one sample in a single on-chip
arena and walks a flat list of ops.  This module builds that list, plans the arena (which tensor lives
at which float offset, with liveness based reuse) and packs the checkpoint tensors into the parameter
blob the kernel streams per op.

The schedule restates ``NCSNv2Deepest.forward`` (reference ``ncsnv2/models/ncsnv2.py:269-300``) with
its building blocks ``ResidualBlock`` (``ncsnv2/models/layers.py:443-456``), ``ConvMeanPool``
(``layers.py:309-313``), ``RefineBlock`` (``layers.py:234-249``), ``RCUBlock`` (``layers.py:126-134``),
``CRPBlock`` (``layers.py:76-83``), ``MSFBlock`` (``layers.py:178-184``) and ``InstanceNorm2dPlus``
(``ncsnv2/models/normalization.py:163-176``).  Fusions applied (all exact re-associations of the
reference arithmetic):

* ``ELU`` is folded into the epilogue of whichever op produces its input (``edst``);
* residual / multi-scale sums are epilogue accumulations (``acc``);
* ``ConvMeanPool`` = conv then 2x2 mean-pool: four accumulations (one per pooling position) summed in the
  epilogue, weights pre-scaled by 1/4 -- ``F_POOL``.

Activation layout (``Geo``): a [C, h, w] tensor is stored channel-interleaved by 4 with a zero halo,
``addr(c, y, x) = base + ((c // 4) * pps + org + y * wp + x) * 4 + c % 4`` with ``wp = w + 2*hx``,
``pps = (h + 2*hy) * wp``, ``org = hy * wp + hx``.  The halo (hy, hx) covers every live convolution tap
at that resolution, so the implicit-im2col gather of a K step (tap, chunk of 8 input channels) is the
*same* address pattern shifted by a constant -- tabulated per op at the head of its parameter segment --
and needs no bounds checks; one pixel of one plane (4 channels, 16 bytes) is one ldmatrix row, 8 consecutive
pixels are 128 contiguous bytes (conflict free), one ldmatrix.x4 per lane loads a whole m16n8k8 A fragment (two
planes) in register order, and the same bytes are a no-swizzle K-major UMMA core matrix (tcgen05-ready).
The sampler state ``x`` and the raw network output are *compact* (re, im) pair arrays [h*w].

Nothing here touches a GPU; ``simulate`` is a torch (CPU) interpreter of the program used by the CPU
test-suite to pin the schedule against the reference module before any CUDA code runs.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np

# ---- op kinds (mirrored in csrc/sbc_program.h) ---------------------------
OP_AFFINE = 0     # dst = 2*src - 1                       (ncsnv2.py:270-271)
OP_NORM_ELU = 2   # dst = ELU(InstanceNorm++(src))
OP_ELU = 3        # dst = ELU(src)
OP_MAXPOOL5 = 4   # dst = maxpool 5x5 stride 1 pad 2
OP_UPACC = 5      # acc += bilinear(src -> (oh,ow), align_corners=True)
OP_CONV_MMA = 6   # Conv2d k in {1,3} (+ fused epilogue) on tensor cores (mma.sync m16n8k8 TF32)
OP_SPILL = 7      # park[dst:dst+4*MT] = arena[src:...]   (whole tensor incl. halo leaves shared memory)
OP_FILL = 8       # arena[dst:dst+4*MT] = park[src:...]

# ---- flags ----------------------------------------------------------------
F_POOL = 1        # conv followed by 2x2 mean-pool (ConvMeanPool)
F_X3 = 2          # 3xTF32 error-compensated product (fp32-equivalent accuracy)

PRECISIONS = ("tf32x3", "tf32")

F_COMPACT = 4     # conv epilogue writes couts (0,1) as compact (re, im) pairs: the network output
F_ZH_DST = 8      # the halo of the fresh tensor `dst` must be re-zeroed (set by ProgramBuilder._halo_analysis)
F_ZH_EDST = 16    # same for `edst`
F_UNIT = 32       # conv: unit (pixel tile, cout tile) u is owned by warps u*ks .. u*ks+ks-1 (ks = 1: no K split)
F_LATEW = 128     # the op loads its own parameter segment (single staging buffer) instead of being prefetched
F_ACC_G = 64      # conv: `acc` is an offset into the CTA's park area (global memory / L2) instead of the arena

# Shared memory of one B200 SM that CTAs can share (cudaDevAttrMaxSharedMemoryPerMultiprocessor), the per-CTA
# reservation and an upper bound of the kernel's static shared memory: a plan whose arena + misc region is at most
# smem_budget_bytes(2) runs with TWO CTAs (= two channel realisations in flight) per SM.
SMEM_PER_SM = 233472
SMEM_RESERVED_PER_CTA = 1024
SMEM_STATIC_BOUND = 1024


def smem_budget_bytes(ctas_per_sm: int) -> int:
    return SMEM_PER_SM // ctas_per_sm - SMEM_RESERVED_PER_CTA - SMEM_STATIC_BOUND

OP_FIELDS = ("kind", "flags", "src", "dst", "acc", "edst", "cin", "cout", "h", "w", "ksize", "dil",
             "w_off", "w_len", "b_rel", "sgeo", "dgeo", "ks", "scratch", "oh", "ow", "next_w", "tapmask",
             "wbuf", "MT", "NT", "S", "frag_rel", "low", "nw_off", "nw_len", "nw_buf")
DEVICE_FILLED = ("next_w", "nw_off", "nw_len", "nw_buf")     # written by sbc_model_create
OP_WORDS = len(OP_FIELDS)            # 32 int32 = 128 bytes per op
assert OP_WORDS == 32
MAX_GEO = 8
GEO_WORDS = 8                        # h, w, hy, hx, wp, pps, org, lw


def ilog2(v: int) -> int:
    """log2(v) if v is a power of two, else -1 (the device code then divides instead of shifting)."""
    return v.bit_length() - 1 if v > 0 and not (v & (v - 1)) else -1


@dataclass(frozen=True)
class Geo:
    h: int
    w: int
    hy: int
    hx: int

    @property
    def wp(self) -> int:
        return self.w + 2 * self.hx

    @property
    def pps(self) -> int:
        return (self.h + 2 * self.hy) * self.wp

    @property
    def org(self) -> int:
        return self.hy * self.wp + self.hx

    def floats(self, c: int) -> int:
        return ((c + 3) // 4) * self.pps * 4

    def words(self) -> List[int]:
        return [self.h, self.w, self.hy, self.hx, self.wp, self.pps, self.org, ilog2(self.w)]


@dataclass
class Op:
    kind: int
    flags: int = 0
    src: object = -1    # float offset into the arena (tensor name until the plan is solved)
    dst: object = -1    # raw result store           (-1: none)
    acc: object = -1    # v += acc[i]; acc[i] = v    (-1: none)
    edst: object = -1   # edst[i] = ELU(v)           (-1: none)
    cin: int = 0
    cout: int = 0
    h: int = 0          # INPUT spatial size
    w: int = 0
    ksize: int = 0
    dil: int = 1
    w_off: int = 0      # float offset of this op's parameter segment in the blob
    w_len: int = 0      # floats (multiple of 4) streamed for this op
    b_rel: int = -1     # bias offset inside the segment (-1: no bias)
    sgeo: int = 0       # geometry index of the input tensor
    dgeo: int = 0       # geometry index of the output tensors (dst / acc / edst)
    ks: int = 1         # conv: number of warps that split the K steps of one (tile, cout-tile) unit
    scratch: object = -1  # arena offset of the per-op scratch (K-split partials / norm statistics)
    oh: int = 0         # OUTPUT spatial size
    ow: int = 0
    tapmask: int = 0    # conv: bit i set = tap i (row-major in the k x k window) can touch the image
    wbuf: object = -1   # arena offset where this op's parameter segment is staged (cp.async.bulk)
    MT: int = 0         # conv: 16-pixel output tiles
    NT: int = 0         # conv: 8-cout tiles
    S: int = 0          # conv: K steps (live taps x cin chunks); the segment starts with S int32 A offsets
    frag_rel: int = 0   # conv: offset of the B fragments inside the segment
    low: int = -1       # log2(ow) or -1
    name: str = ""
    branches: object = None   # conv, host side only: per summed conv {prefix, src, cin, k, dil, live, KC, step0}

    def words(self) -> List[int]:
        vals = [0 if f in DEVICE_FILLED else getattr(self, f) for f in OP_FIELDS]
        return [int(v) for v in vals]


@dataclass
class Program:
    ops: List[Op]
    geos: List[Geo]
    arena_floats: int
    blob: np.ndarray                 # float32 parameter blob
    in_off: int                      # arena offset of the network input x: compact (re, im) pairs [H*W]
    out_off: int                     # arena offset of the raw network output (before /sigma), compact pairs
    post_off: int                    # scratch for the sampler phases that follow the network
    H: int
    W: int
    ngf: int
    channels: int
    nthreads: int
    max_w_len: int
    conv_flops: int                  # dense conv FLOPs / forward / sample (reference convention)
    precision: str
    park_floats: int = 0             # per-CTA park area in global memory (SPILL / FILL / F_ACC_G); 0: none

    def misc_bytes(self) -> int:
        """Shared-memory misc region next to the arena (csrc/sbc_api.cu): two mbarriers + the halo-pixel lists."""
        halo = sum(g.pps - g.h * g.w for g in self.geos)
        return (64 + 2 * halo + 15) // 16 * 16

    def smem_bytes(self) -> int:
        return 4 * self.arena_floats + self.misc_bytes()

    def op_table(self) -> np.ndarray:
        t = np.zeros((len(self.ops), OP_WORDS), dtype=np.int32)
        for i, op in enumerate(self.ops):
            t[i] = op.words()
        return t

    def geo_table(self) -> np.ndarray:
        t = np.zeros((MAX_GEO, GEO_WORDS), dtype=np.int32)
        for i, g in enumerate(self.geos):
            t[i] = g.words()
        return t

    def geo_of(self, h: int, w: int) -> Geo:
        for g in self.geos:
            if (g.h, g.w) == (h, w):
                return g
        raise KeyError((h, w))

    # ---- arena access helpers (numpy arrays or torch tensors) ----
    def _index(self, off: int, c: int, h: int, w: int) -> np.ndarray:
        g = self.geo_of(h, w)
        cc, yy, xx = np.meshgrid(np.arange(c), np.arange(h), np.arange(w), indexing="ij")
        return off + ((cc // 4) * g.pps + g.org + yy * g.wp + xx) * 4 + cc % 4

    def read(self, arena, off: int, c: int, h: int, w: int):
        """[c,h,w] copy of the tensor stored at float offset ``off``."""
        idx = self._index(off, c, h, w)
        if isinstance(arena, np.ndarray):
            return arena[idx]
        import torch
        return arena[torch.from_numpy(idx)]

    def write(self, arena, off: int, value, cstore: Optional[int] = None) -> None:
        """Store a [c,h,w] tensor (zero halo, zero padding channels)."""
        c, h, w = value.shape
        g = self.geo_of(h, w)
        n = g.floats(cstore or c)
        arena[off:off + n] = 0
        idx = self._index(off, c, h, w)
        if isinstance(arena, np.ndarray):
            arena[idx] = value
        else:
            import torch
            arena[torch.from_numpy(idx)] = value

    def _compact_index(self, off: int) -> np.ndarray:
        cc, ee = np.meshgrid(np.arange(self.channels), np.arange(self.H * self.W), indexing="ij")
        return (off + ee * self.channels + cc).reshape(self.channels, self.H, self.W)

    def write_input(self, arena, x) -> None:
        """Store the network input x [channels,H,W] in its compact form (interleaved (re, im) pairs)."""
        idx = self._compact_index(self.in_off)
        if isinstance(arena, np.ndarray):
            arena[idx] = np.asarray(x, np.float32)
        else:
            import torch
            arena[torch.from_numpy(idx)] = x

    def read_output(self, arena):
        """[channels,H,W] copy of the raw network output (before the /sigma of ncsnv2.py:295-298)."""
        idx = self._compact_index(self.out_off)
        if isinstance(arena, np.ndarray):
            return arena[idx]
        import torch
        return arena[torch.from_numpy(idx)]


class _Planner:
    """Offline arena planner.

    Tensors are registered with their live interval in units of emitted ops ([born, died)); ``solve``
    places them greedily by decreasing size at the lowest offset that does not collide with an already
    placed tensor whose interval overlaps."""

    def __init__(self, align: int = 4):
        self.align = align
        self.born: Dict[str, int] = {}
        self.died: Dict[str, int] = {}
        self.size: Dict[str, int] = {}
        self.offs: Dict[str, int] = {}
        self.pinned: Dict[str, int] = {}     # name -> fixed offset (placed before everything else)
        self.peak = 0

    def alloc(self, name: str, n: int, now: int, died: Optional[int] = None) -> None:
        assert name not in self.size, name
        self.size[name] = (n + self.align - 1) // self.align * self.align
        self.born[name] = now
        if died is not None:
            self.died[name] = died

    def free(self, name: str, now: int) -> None:
        assert name in self.size and name not in self.died, name
        self.died[name] = now

    def solve(self, end: int) -> None:
        for n in self.size:
            self.died.setdefault(n, end)
        order = sorted((n for n in self.size if n not in self.pinned), key=lambda n: (-self.size[n], self.born[n]))
        placed: List[str] = []
        for n, off in self.pinned.items():
            self.offs[n] = off
            placed.append(n)
            self.peak = max(self.peak, off + self.size[n])
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


def tf32_rna(x: np.ndarray) -> np.ndarray:
    """fp32 -> TF32 (10-bit mantissa), round to nearest with ties away from zero: cvt.rna.tf32.f32."""
    b = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((b + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


class _G:
    """Reference to a tensor in the per-CTA park area (global memory): name + float shift."""

    def __init__(self, name: str, shift: int = 0):
        self.name, self.shift = name, shift


class ProgramBuilder:
    """Builds the op list for one (ngf, H, W) instance of NCSNv2Deepest.

    ``park=True`` plans for TWO resident CTAs per SM (two channel realisations in flight per SM):
    the arena must stay under half of an SM's shared memory, so (a) the residual / skip streams of the largest
    resolution live in a per-CTA *park area* in global memory (L2 resident) -- the conv epilogues read-modify-write them
    there (``F_ACC_G``) and an ``OP_FILL`` brings a copy into the arena for the ops that gather from it; (b) the skip
    tensors of the other resolutions are spilled between their producer and their RefineBlock; (c) parameter segments
    are capped at half the size (more cout chunks)."""

    BIG_BYTES = 24 * 1024        # park mode: tensors at least this large are "big" (their streams live in the park area)

    # largest parameter segment (floats) one slot of the shared-memory ring holds; convs whose fragment
    # array is bigger are split into cout chunks (each chunk is its own op)
    SLOT_FLOATS = 9300

    def __init__(self, state: Dict[str, np.ndarray], ngf: int, H: int, W: int,
                 channels: int = 2, nthreads: int = 512, precision: str = "tf32x3", park: bool = False):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % (PRECISIONS,))
        if H <= 0 or W <= 0 or H % 8 or W % 8:
            raise ValueError("Nt and Nr must be positive multiples of 8 (three 2x mean-pools), got %dx%d" % (H, W))
        self.precision = precision
        self.sd = {k: np.asarray(v, dtype=np.float32) for k, v in state.items()}
        self.ngf, self.H, self.W, self.channels = ngf, H, W, channels
        self.nthreads = nthreads
        self.park = park
        # park mode: a segment above this size gets a single staging buffer, filled at the start of its own op
        # (F_LATEW): two of them side by side (double buffering) would not fit half an SM
        self.LATE_FLOATS = 4700 if park else 1 << 30
        self.goff: Dict[str, int] = {}          # park area: tensor name -> float offset (bump allocated, never reused)
        self.gtop = 0
        self.ar = _Planner()
        self.ops: List[Op] = []
        self.blob: List[np.ndarray] = []
        self.blob_len = 0
        self.shape: Dict[str, Tuple[int, int, int]] = {}
        self.flops = 0
        self._tmp = 0
        # geometries: one per resolution; the halo covers every live tap used at that resolution.  The
        # dilated stages (dilation 2 and 4) run at the lowest resolution (ncsnv2.py:240-254).
        self.geos: List[Geo] = []
        for lvl in range(4):
            h, w = H >> lvl, W >> lvl
            dils = [1] if lvl < 3 else [1, 2, 4]
            hy = max([d for d in dils if d < h] + [0])
            hx = max([d for d in dils if d < w] + [0])
            if w == 2:
                # the 8 rows of one ldmatrix matrix are the pixels (Y,0) (Y,1) ... (Y+3,1), 16 B each: with a row
                # pitch of 6 pixels (96 B) they fall into disjoint shared-memory banks
                hx = max(hx, 2)
            self.geos.append(Geo(h, w, hy, hx))

    def gi(self, h: int, w: int) -> int:
        for i, g in enumerate(self.geos):
            if (g.h, g.w) == (h, w):
                return i
        raise KeyError((h, w))

    # -- tensors ----------------------------------------------------------
    def new(self, name: str, c: int, h: int, w: int, cstore: Optional[int] = None) -> str:
        self.ar.alloc(name, self.geos[self.gi(h, w)].floats(cstore or c), len(self.ops))
        self.shape[name] = (c, h, w)
        return name

    def new_raw(self, name: str, nfloats: int) -> str:
        self.ar.alloc(name, nfloats, len(self.ops))
        return name

    def tmp(self, c: int, h: int, w: int, tag: str = "t", cstore: Optional[int] = None) -> str:
        self._tmp += 1
        return self.new("%s%d" % (tag, self._tmp), c, h, w, cstore)

    def tmp_raw(self, nfloats: int, tag: str) -> str:
        self._tmp += 1
        return self.new_raw("%s%d" % (tag, self._tmp), nfloats)

    def free(self, *names: str) -> None:
        for n in names:
            self.ar.free(n, len(self.ops))

    # -- park area (park mode) --------------------------------------------
    def live(self, name: str) -> bool:
        return name in self.ar.size and name not in self.ar.died

    def release(self, name: str) -> None:
        if self.live(name):
            self.free(name)

    def is_big(self, name: str) -> bool:
        c, h, w = self.shape[name]
        return self.park and self.geos[self.gi(h, w)].floats(c) * 4 >= self.BIG_BYTES

    def _copy_op(self, kind: int, src, dst, name: str, label: str) -> None:
        c, h, w = self.shape[name]
        g = self.gi(h, w)
        self.ops.append(Op(kind, 0, src, dst, cin=c, cout=c, h=h, w=w, sgeo=g, dgeo=g, oh=h, ow=w,
                           MT=self.geos[g].floats(c) // 4, name="%s %s" % (label, name)))

    def spill(self, name: str) -> None:
        """Copy the (live) arena tensor ``name`` into its park slot; the arena copy stays live until freed."""
        assert self.live(name)
        if name not in self.goff:
            c, h, w = self.shape[name]
            self.goff[name] = self.gtop
            self.gtop += self.geos[self.gi(h, w)].floats(c)
        self._copy_op(OP_SPILL, name, _G(name), name, "spill")

    def park_out(self, name: str) -> None:
        self.spill(name)
        self.free(name)

    def fill(self, gname: str) -> str:
        """Bring the parked tensor back into a fresh arena tensor (returned); the park copy stays valid."""
        c, h, w = self.shape[gname]
        t = self.tmp(c, h, w, "f")
        self._copy_op(OP_FILL, _G(gname), t, gname, "fill")
        return t

    def _copy_raw(self, kind: int, src, dst, nfloats: int, label: str) -> None:
        assert nfloats % 4 == 0
        self.ops.append(Op(kind, 0, src, dst, MT=nfloats // 4, name=label))

    def smem_of(self, name: str) -> str:
        return name if self.live(name) else self.fill(name)

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
             edst: Optional[str] = None, dil: int = 1, pool: bool = False, compact: bool = False,
             siblings=(), acc_g: bool = False) -> None:
        """One Conv2d (reference ``layers.py:28-60``) + fused epilogue, as an implicit GEMM on tensor cores:

            D[16 pixels, 8 couts] += A[16 pixels, 8 cins] . B[8 cins, 8 couts]   per K step (live tap, cin chunk)

        v = conv(src) + bias;  dst <- v;  acc <- (v += acc);  edst <- ELU(v).
        Parameter segment: [S int32 A offsets | B fragments | bias].  ``aoff[s]`` is the float offset
        (2 * cin_chunk * pps + dy * wp + dx) * 4 that K step s adds to every gathered address.  The weights are in
        mma.sync m16n8k8 B-fragment order (csrc/sbc_mma.h):
        frag[step][ntile][lane] = (w0, w1) where lane = 4*g + t holds W[cout g][cin t (+4)]: TF32-rounded (rna)
        in the "tf32" mode, plain fp32 in the 3xTF32 mode (the kernel splits w = hi + lo in registers).

        ``siblings``: further (prefix, src, dil) convs whose outputs are SUMMED with this one (the shortcut of a
        'down' ResidualBlock, ``layers.py:416-420,455-456``; the second branch of a same-size MSFBlock,
        ``layers.py:181-183``).  Because a K step is just an address offset, their K steps are appended to the table
        with the distance between the two source tensors folded into the offsets (patched in once the arena plan
        is known), their fragments to the fragment array, and their biases are added: one op, one pass."""
        branches = [(prefix, src, dil)] + list(siblings)
        x3 = self.precision == "tf32x3"
        E = 2                                              # floats per lane per fragment
        c0, h, w = self.shape[src]
        sg = self.geos[self.gi(h, w)]
        oh, ow = (h // 2, w // 2) if pool else (h, w)
        cout = self.sd[prefix + ".weight"].shape[0]
        assert cout % 2 == 0, "the epilogue stores adjacent cout pairs"
        for t in (acc, edst) + (() if compact else (dst,)):
            if t is not None:
                assert self.shape[t] == (cout, oh, ow), (prefix, t, self.shape[t], (cout, oh, ow))
        NT = (cout + 7) // 8
        bias = None
        meta, aoffs, wpads = [], [], []                    # per branch
        for (bp, bsrc, bdil) in branches:
            wt = self.sd[bp + ".weight"]
            bco, cin, k, k2 = wt.shape
            assert k == k2 and k in (1, 3) and bco == cout
            c, bh, bw = self.shape[bsrc]
            assert (bh, bw) == (h, w), "sibling convs must read tensors of the same geometry"
            assert c == cin or (c == 8 and cin < 8), (bp, c, cin)   # begin_conv reads a zero-padded chunk
            self.flops += 2 * h * w * cin * k * k * cout   # dense count, reference convention
            r = k // 2
            live = []
            for tap in range(k * k):
                dy, dx = (tap // k - r) * bdil, (tap % k - r) * bdil
                if abs(dy) < h and abs(dx) < w:            # otherwise the tap only ever reads zero padding
                    assert abs(dy) <= sg.hy and abs(dx) <= sg.hx, "halo too small"
                    live.append(tap)
            KC = (cin + 7) // 8
            ao = np.zeros(len(live) * KC, np.int32)
            for i, tap in enumerate(live):
                dy, dx = (tap // k - r) * bdil, (tap % k - r) * bdil
                for kc in range(KC):
                    ao[i * KC + kc] = (2 * kc * sg.pps + dy * sg.wp + dx) * 4
            wp_ = np.zeros((NT * 8, KC * 8, k * k), np.float32)
            wp_[:cout, :cin] = wt.reshape(cout, cin, k * k) * (np.float32(0.25) if pool else np.float32(1.0))
            b = self.sd.get(bp + ".bias")
            if b is not None:
                bias = b.astype(np.float32) if bias is None else (bias + b).astype(np.float32)
            meta.append(dict(prefix=bp, src=bsrc, cin=cin, k=k, dil=bdil, live=live, KC=KC, step0=sum(len(a_) for a_ in aoffs)))
            aoffs.append(ao)
            wpads.append(wp_)
        S = sum(len(a_) for a_ in aoffs)
        S4 = (S + 3) // 4 * 4
        aoff = np.zeros(S4, np.int32)
        aoff[:S] = np.concatenate(aoffs)
        per_nt = S * 32 * E
        nt_chunk = NT
        while nt_chunk > 1 and S4 + per_nt * nt_chunk + 8 * nt_chunk > self.SLOT_FLOATS:
            nt_chunk //= 2
        g, t = np.arange(32) >> 2, np.arange(32) & 3
        nwarps = self.nthreads // 32
        dgeo = self.geos[self.gi(oh, ow)]
        k0, cin0 = meta[0]["k"], meta[0]["cin"]
        tapmask = sum(1 << tp for tp in meta[0]["live"])
        for nt0 in range(0, NT, nt_chunk):
            ntc = min(nt_chunk, NT - nt0)
            co0, co1 = nt0 * 8, min(cout, (nt0 + ntc) * 8)
            frag = np.zeros((S, ntc, 32, E), np.float32)
            for m, wp_ in zip(meta, wpads):
                for i, tap in enumerate(m["live"]):
                    for kc in range(m["KC"]):
                        for nt in range(ntc):
                            w0 = wp_[(nt0 + nt) * 8 + g, kc * 8 + t, tap]
                            w1 = wp_[(nt0 + nt) * 8 + g, kc * 8 + t + 4, tap]
                            if not x3:
                                w0, w1 = tf32_rna(w0), tf32_rna(w1)
                            st = m["step0"] + i * m["KC"] + kc
                            frag[st, nt, :, 0], frag[st, nt, :, 1] = w0, w1
            arrs = [aoff.view(np.float32), frag] + ([bias[co0:co1]] if bias is not None else [])
            w_off, w_len, rels = self._push(arrs)
            MT = (oh * ow + 15) // 16
            units = MT * ntc
            # fewer units than warps: `ks` warps split the K steps of a unit (measured: even in the 3xTF32 mode the
            # serial K chain of a single warp costs more than the shared-memory exchange + barrier of the split:
            # 36.3 est/s without, 40.5 with)
            ks, unit = 1, units * 2 <= nwarps
            if unit:
                while units * ks * 2 <= min(nwarps, 16) and ks * 2 <= S:
                    ks *= 2
            scratch = self.tmp_raw(units * ks * 32 * 4, "ksp") if ks > 1 else None
            flags = (F_POOL if pool else 0) | (F_X3 if x3 else 0) | (F_COMPACT if compact else 0) | (F_UNIT if unit else 0)
            if acc_g:
                assert acc is not None and acc in self.goff and not self.live(acc), "acc_g: the stream must live in the park area only"
                flags |= F_ACC_G
            if compact:
                assert nt_chunk == NT and cout == 2 and acc is None and edst is None
                shift = lambda name: -1 if name is None else name
            else:
                shift = lambda name: -1 if name is None else (name, (co0 // 4) * dgeo.pps * 4)
            acc_ref = _G(acc, (co0 // 4) * dgeo.pps * 4) if acc_g else shift(acc)
            op = Op(OP_CONV_MMA, flags, src, shift(dst), acc_ref, shift(edst), cin0, co1 - co0, h, w, k0,
                    dil, w_off, w_len, rels[2] if bias is not None else -1, self.gi(h, w), self.gi(oh, ow),
                    ks, scratch if scratch is not None else -1, oh, ow, tapmask=tapmask, MT=MT, NT=ntc,
                    S=S, frag_rel=rels[1], low=ilog2(ow),
                    name="+".join(m["prefix"] for m in meta) + ("" if nt_chunk == NT else "[co%d:%d]" % (co0, co1)))
            op.branches = [dict(m) for m in meta]          # host-side only (simulator, offset fix-ups)
            self.ops.append(op)
            if scratch is not None:
                self.free(scratch)

    def norm_elu(self, prefix: str, src: str, dst: str) -> None:
        """dst = ELU(InstanceNorm2dPlus(src))  (``normalization.py:163-176`` + ``layers.py:13``)."""
        c, h, w = self.shape[src]
        assert self.shape[dst] == (c, h, w)
        w_off, w_len, _ = self._push([np.concatenate([self.sd[prefix + ".alpha"],
                                                      self.sd[prefix + ".gamma"],
                                                      self.sd[prefix + ".beta"]])])
        # scratch: per-warp partial sums of the two statistics passes (4 channels each) + per-channel (mean, M2)
        nwarps = self.nthreads // 32
        scratch = self.tmp_raw(2 * nwarps * 4 + 2 * c, "nsc")
        T, lT, npass = self._quad_threads(c, h * w)
        fb = lambda v: int(np.float32(v).view(np.int32))
        self.ops.append(Op(OP_NORM_ELU, 0, src, dst, cin=c, cout=c, h=h, w=w, w_off=w_off, w_len=w_len,
                           sgeo=self.gi(h, w), dgeo=self.gi(h, w), scratch=scratch, oh=h, ow=w, MT=T, NT=lT, S=npass,
                           frag_rel=fb(1.0 / (h * w)), low=fb(1.0 / c), tapmask=fb(1.0 / (c - 1)), name=prefix))
        self.free(scratch)

    def _quad_threads(self, c: int, hw: int = 1 << 30) -> Tuple[int, int, int]:
        """(T, log2 T, passes): T = threads per channel quad = largest power of two <= nthreads / (c/4), >= 32.
        Small maps (<= 64 pixels) get one warp per quad: the statistics then need no cross-warp exchange."""
        assert c % 8 == 0
        nq = c // 4
        T = 32
        while T * 2 <= self.nthreads // nq and hw > 64:
            T *= 2
        gpp = self.nthreads // T
        return T, T.bit_length() - 1, (nq + gpp - 1) // gpp

    def elu(self, src: str, dst: str) -> None:
        c, h, w = self.shape[src]
        T, lT, npass = self._quad_threads(c)
        self.ops.append(Op(OP_ELU, 0, src, dst, cin=c, cout=c, h=h, w=w, sgeo=self.gi(h, w), dgeo=self.gi(h, w),
                           oh=h, ow=w, MT=T, NT=lT, S=npass, name="elu"))

    def affine(self, src: str, dst: str) -> None:
        """dst (8 stored channels) = 2*x - 1 on the real channels read from the compact state buffer ``src``;
        the padding channels of dst are written as zeros."""
        c, h, w = self.shape[dst]
        assert c == 8 and self.channels == 2
        self.ops.append(Op(OP_AFFINE, 0, src, dst, cin=self.channels, cout=c, h=h, w=w, sgeo=self.gi(h, w),
                           dgeo=self.gi(h, w), oh=h, ow=w, name="2x-1"))

    def maxpool5(self, src: str, dst: str) -> None:
        c, h, w = self.shape[src]
        self.ops.append(Op(OP_MAXPOOL5, 0, src, dst, cin=c, cout=c, h=h, w=w, sgeo=self.gi(h, w), dgeo=self.gi(h, w),
                           oh=h, ow=w, name="maxpool5"))

    def upacc(self, src: str, acc: str, edst: Optional[str] = None) -> None:
        """acc += bilinear(src, size=acc.shape, align_corners=True) (``layers.py:182-183``)."""
        c, h, w = self.shape[src]
        c2, oh, ow = self.shape[acc]
        assert c == c2
        self.ops.append(Op(OP_UPACC, 0, src, -1, acc, edst if edst is not None else -1, cin=c, cout=c, h=h, w=w,
                           sgeo=self.gi(h, w), dgeo=self.gi(oh, ow), oh=oh, ow=ow, name="upacc"))

    # -- blocks -----------------------------------------------------------
    def residual(self, p: str, x: str, cout: int, down: bool, dil: Optional[int], elu_out: Optional[str] = None) -> str:
        """ResidualBlock.forward (``layers.py:443-456``). Consumes ``x``; returns the output tensor.  ``elu_out``
        (same-shape blocks only): tensor that additionally receives ELU(output) from the last conv's epilogue."""
        cin, h, w = self.shape[x]
        d = dil or 1
        if self.is_big(x) and cout == cin and not down:
            # park mode: the stream x lives in the park area; only its normalised copy and conv1's output are in the arena
            if x not in self.goff:
                self.spill(x)
            s_ = self.smem_of(x)
            t = self.tmp(cin, h, w)
            self.norm_elu(p + ".normalize1", s_, t)
            self.free(s_)
            t2 = self.tmp(cout, h, w)
            self.conv(p + ".conv1", t, dst=t2, dil=d)
            self.free(t)
            t3 = self.tmp(cout, h, w)
            self.norm_elu(p + ".normalize2", t2, t3)
            self.free(t2)
            self.conv(p + ".conv2", t3, acc=x, edst=elu_out, dil=d, acc_g=True)
            self.free(t3)
            return x
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
            self.conv(p + ".conv2", t3, acc=x, edst=elu_out, dil=d)   # shortcut = x
            self.free(t3)
            return x
        assert elu_out is None
        if down and dil is None:
            out = self.tmp(cout, h // 2, w // 2, "o")
            # ConvMeanPool 3x3 + ConvMeanPool 1x1 shortcut, one op
            self.conv(p + ".conv2.conv", t3, dst=out, pool=True, siblings=[(p + ".shortcut.conv", x, 1)])
        else:
            out = self.tmp(cout, h, w, "o")
            self.conv(p + ".conv2", t3, dst=out, dil=d, siblings=[(p + ".shortcut", x, d)])
        self.free(t3)
        self.free(x)
        return out

    def rcu(self, p: str, x: str, n_blocks: int, e_in: Optional[str] = None,
            want_elu_out: bool = False) -> Tuple[str, Optional[str]]:
        """RCUBlock.forward (``layers.py:126-134``), in place on ``x``.

        ``e_in`` (optional) already holds ELU(x).  Returns (x, ELU(x) or None)."""
        c, h, w = self.shape[x]
        e = e_in
        big = self.is_big(x)       # park mode: x is only ever read-modify-written by the conv epilogues, in the park area
        if big:
            if x not in self.goff:
                self.spill(x)
            if e is not None:
                self.release(x)
        for i in range(n_blocks):
            if e is None:
                e = self.tmp(c, h, w)
                if big:
                    s_ = self.smem_of(x)
                    self.elu(s_, e)
                    self.free(s_)
                else:
                    self.elu(x, e)
            u = self.tmp(c, h, w)
            self.conv("%s.%d_1_conv" % (p, i + 1), e, edst=u)      # u = ELU(conv1(ELU(x)))
            last = (i == n_blocks - 1)
            if last and not want_elu_out:
                self.free(e)
                e = None
                self.conv("%s.%d_2_conv" % (p, i + 1), u, acc=x, acc_g=big)
            else:
                self.conv("%s.%d_2_conv" % (p, i + 1), u, acc=x, edst=e, acc_g=big)  # x += conv2(u); e = ELU(x)
            self.free(u)
        return x, e

    def crp(self, p: str, x: str, e: str) -> Tuple[str, str]:
        """CRPBlock.forward (``layers.py:76-83``).  ``e`` = ELU(x) becomes the running sum; x is dead.

        Returns (sum, ELU(sum)) -- the ELU feeds the output RCU that always follows."""
        c, h, w = self.shape[e]
        self.release(x)
        big = self.is_big(e)
        m = self.tmp(c, h, w)
        self.maxpool5(e, m)
        if big:                    # park mode: the running sum moves to the park area
            self.park_out(e)
        path = self.tmp(c, h, w)
        self.conv(p + ".convs.0", m, dst=path, acc=e, acc_g=big)
        self.maxpool5(path, m)
        self.free(path)
        e2 = self.tmp(c, h, w)
        self.conv(p + ".convs.1", m, acc=e, edst=e2, acc_g=big)
        self.free(m)
        return e, e2

    def refine(self, p: str, xs: List[str], es: List[Optional[str]], features: int,
               end: bool = False) -> Tuple[str, Optional[str]]:
        """RefineBlock.forward (``layers.py:234-249``).  ``es[i]`` optionally holds ELU(xs[i]).

        Returns (h, ELU(h)) (ELU(h) is None for the last block)."""
        hs, e_single = [], None
        for i, x in enumerate(xs):
            # a single-input block applies ELU to the adapted tensor right away (CRP): take it from the epilogue
            h_, e_single = self.rcu("%s.adapt_convs.%d" % (p, i), x, 2, e_in=es[i], want_elu_out=(len(xs) == 1))
            hs.append(h_)
        c0, oh, ow = self.shape[hs[0]]
        if len(hs) > 1:
            s = self.tmp(features, oh, ow, "s")
            same = self.shape[hs[1]][1:] == (oh, ow)
            if same:
                e = self.tmp(features, oh, ow, "e")
            if same:   # bilinear to the same size with align_corners=True is the identity: one summed conv
                a0, a1 = self.smem_of(hs[0]), self.smem_of(hs[1])
                self.conv(p + ".msf.convs.0", a0, dst=s, edst=e, siblings=[(p + ".msf.convs.1", a1, 1)])
                self.free(a0)
                self.free(a1)
            else:
                a0 = self.smem_of(hs[0])   # (park mode: the adapted stream comes back from the park area)
                self.conv(p + ".msf.convs.0", a0, dst=s)
                self.free(a0)
                _, lh, lw = self.shape[hs[1]]
                lo = self.tmp(features, lh, lw, "lo")
                a1 = self.smem_of(hs[1])
                self.conv(p + ".msf.convs.1", a1, dst=lo)
                self.free(a1)
                e = self.tmp(features, oh, ow, "e")
                self.upacc(lo, s, edst=e)
                self.free(lo)
        else:
            s, e = hs[0], e_single
        h_, eh = self.crp(p + ".crp", s, e)
        return self.rcu(p + ".output_convs", h_, 3 if end else 1, e_in=eh, want_elu_out=not end)

    # -- whole network ----------------------------------------------------
    def build(self) -> Program:
        ngf, H, W = self.ngf, self.H, self.W
        assert ngf % 8 == 0, "ngf must be a multiple of 8 (one MMA K chunk = 8 input channels = two planes)"
        nx = self.channels * H * W
        xin = self.new_raw("x_in", nx)                          # compact (re, im) pairs
        self.ar.pinned[xin] = 0
        a = self.tmp(8, H, W)                                   # begin_conv reads one chunk of 8 input channels
        self.affine(xin, a)
        if self.park:   # the sampler state waits in the park area while the network runs; it returns to the SAME offset
            self.goff["x_in"] = self.gtop
            self.gtop += nx
            self._copy_raw(OP_SPILL, xin, _G("x_in"), nx, "spill x_in")
            self.free(xin)
        o = self.tmp(ngf, H, W, "o")
        self.conv("begin_conv", a, dst=o)
        self.free(a)
        l1 = self.residual("res1.1", self.residual("res1.0", o, ngf, False, None), ngf, False, None)
        pk = self.park
        keep = lambda name: self.park_out(name) if (pk and name is not None and self.live(name)) else None
        back = lambda name: self.fill(name) if (pk and name is not None and not self.live(name)) else name
        # l4 and l5 also emit ELU(skip) from the epilogue of their last conv (the first thing their RefineBlock does
        # with them); for the others a stand-alone ELU op later is cheaper than arena held across the phase where
        # the two largest parameter staging buffers are live (it would push the plan past the 227 KB of one SM)
        # park mode: every skip tensor waits for its RefineBlock in the park area (keep / back are no-ops otherwise)
        l2, _ = self._stage("res2", l1, 2 * ngf, None)
        l3, el3 = self._stage("res3", l2, 2 * ngf, None)
        keep(l2)
        l31, el31 = self._stage("res31", l3, 2 * ngf, None)
        keep(l3); keep(el3)
        l4, el4 = self._stage("res4", l31, 4 * ngf, 2, want_elu=True)
        keep(l31); keep(el31)
        l5, el5 = self._stage("res5", l4, 4 * ngf, 4, want_elu=True)
        keep(l4); keep(el4)
        r1, e1 = self.refine("refine1", [l5], [el5], 4 * ngf)
        r2, e2 = self.refine("refine2", [back(l4), r1], [back(el4), e1], 2 * ngf)
        r31, e31 = self.refine("refine31", [back(l31), r2], [back(el31), e2], 2 * ngf)
        r3, e3 = self.refine("refine3", [back(l3), r31], [back(el3), e31], 2 * ngf)
        r4, e4 = self.refine("refine4", [back(l2), r3], [None, e3], ngf)
        r5, _ = self.refine("refine5", [l1, r4], [None, e4], ngf, end=True)
        t = self.tmp(ngf, H, W)
        r5s = self.smem_of(r5)
        self.norm_elu("normalizer", r5s, t)
        self.free(r5s)
        out = self.new_raw("net_out", self.channels * H * W)    # compact (re, im) pairs
        self.conv("end_conv", t, dst=out, compact=True)
        self.free(t)
        if self.park:
            xin2 = self.new_raw("x_in.back", nx)
            self.ar.pinned[xin2] = 0
            self._copy_raw(OP_FILL, _G("x_in"), xin2, nx, "fill x_in")
        # scratch for the sampler phases that follow the network (residual P*x-y, reductions)
        post = self.new_raw("post", 2 * H * W)
        blob = np.concatenate(self.blob) if self.blob else np.zeros(0, np.float32)
        assert blob.size == self.blob_len
        max_w_len = max(op.w_len for op in self.ops)
        # ---- staging buffers for the per-op parameter segments: the segment of op i is copied (cp.async.bulk)
        # while the previous parameterised op runs, so its buffer is live over [prev, i]; the first one is
        # refilled at the end of every forward for the next one and simply stays allocated.
        end = len(self.ops) + 1
        prev = None
        for i, op in enumerate(self.ops):
            if op.w_len > 0:
                name = "wbuf%d" % i
                if prev is None:
                    self.ar.alloc(name, op.w_len, 0, end)
                elif op.w_len > self.LATE_FLOATS:
                    op.flags |= F_LATEW
                    self.ar.alloc(name, op.w_len, i, i + 1)
                else:
                    self.ar.alloc(name, op.w_len, prev, i + 1)
                op.wbuf = name
                prev = i
        # ---- solve the arena plan and resolve tensor names to float offsets ----
        self.ar.solve(end)
        for op in self.ops:
            if op.branches:   # sibling convs: fold the distance between the source tensors into their K-step offsets
                base = self.ar.offs[op.branches[0]["src"]]
                tab = blob[op.w_off:op.w_off + op.S].view(np.int32)
                for m in op.branches:
                    m["src_off"] = self.ar.offs[m["src"]]
                    n = len(m["live"]) * m["KC"]
                    tab[m["step0"]:m["step0"] + n] += m["src_off"] - base
            for f in ("src", "dst", "acc", "edst", "scratch", "wbuf"):
                v = getattr(op, f)
                if isinstance(v, _G):
                    setattr(op, f, self.goff[v.name] + v.shift)
                elif isinstance(v, tuple):
                    setattr(op, f, self.ar.offs[v[0]] + v[1])
                elif isinstance(v, str):
                    setattr(op, f, self.ar.offs[v])
                else:
                    setattr(op, f, -1 if v is None else int(v))
        self._halo_analysis(self.ar.peak, [(self.ar.offs[xin], self.channels * H * W),
                                           (self.ar.offs[post], 2 * H * W)])
        return Program(self.ops, self.geos, self.ar.peak, blob, self.ar.offs[xin], self.ar.offs[out],
                       self.ar.offs[post], H, W, ngf, self.channels, self.nthreads, max_w_len, self.flops,
                       self.precision, self.gtop)

    def _halo_analysis(self, arena_floats: int, raw_regions) -> None:
        """Decide which ops must re-zero the halo of the tensors they create (flags F_ZH_DST / F_ZH_EDST).

        Ops write the interior of their outputs only; a halo is known to be zero when every cell of the plane
        was last owned by a plane of the SAME geometry at the SAME address (its halo cells coincide and were
        zero by induction).  The walk assumes an arbitrary arena at the start of every forward pass, and treats
        raw buffers (parameter staging, scratch, the compact state / output, the sampler scratch) as dirt."""
        tags = np.full(arena_floats, -1, np.int64)

        def raw(off: int, n: int) -> None:
            if off >= 0 and n > 0:
                tags[off:off + n] = -2

        def fresh(off: int, gi: int, c: int) -> bool:
            g = self.geos[gi]
            need = False
            for pl in range((c + 3) // 4):
                a = off + pl * g.pps * 4
                key = (gi << 40) | a
                seg = tags[a:a + g.pps * 4]
                if not bool((seg == key).all()):
                    need = True
                seg[:] = key
            return need

        for off, n in raw_regions:
            raw(off, n)
        nwarps = self.nthreads // 32
        for op in self.ops:
            raw(op.wbuf, op.w_len)
            if op.kind == OP_CONV_MMA:
                raw(op.scratch, op.MT * op.NT * op.ks * 32 * 4 if op.ks > 1 else 0)
                if op.flags & F_COMPACT:
                    raw(op.dst, self.channels * op.oh * op.ow)
                elif op.dst >= 0 and fresh(op.dst, op.dgeo, op.cout):
                    op.flags |= F_ZH_DST
                if op.edst >= 0 and fresh(op.edst, op.dgeo, op.cout):
                    op.flags |= F_ZH_EDST
            elif op.kind == OP_NORM_ELU:
                raw(op.scratch, 2 * nwarps * 4 + 2 * op.cin)
                if fresh(op.dst, op.dgeo, op.cin):
                    op.flags |= F_ZH_DST
            elif op.kind in (OP_ELU, OP_MAXPOOL5):
                if fresh(op.dst, op.dgeo, op.cin):
                    op.flags |= F_ZH_DST
            elif op.kind == OP_AFFINE:
                if fresh(op.dst, op.dgeo, op.cout):
                    op.flags |= F_ZH_DST
            elif op.kind == OP_UPACC:
                if op.edst >= 0 and fresh(op.edst, op.dgeo, op.cin):
                    op.flags |= F_ZH_EDST
            elif op.kind == OP_SPILL:
                pass
            elif op.kind == OP_FILL:
                if op.cin == 0:
                    raw(op.dst, 4 * op.MT)         # raw buffer (the sampler state)
                else:
                    fresh(op.dst, op.dgeo, op.cin) # the copy brings its (zero) halo along: tag the planes, no re-zeroing
            else:
                raise ValueError(op.kind)

    def _stage(self, p: str, skip: str, cout: int, dil: Optional[int], want_elu: bool = False):
        """Two ResidualBlocks, the first 'down'; ``skip`` must survive (it feeds a RefineBlock later).
        Returns (output, ELU(output) or None)."""
        cin, h, w = self.shape[skip]
        d = dil or 1
        big = self.is_big(skip)    # park mode: the skip stream lives in the park area, arena copies come and go
        if big and skip not in self.goff:
            self.spill(skip)
        s_ = self.smem_of(skip) if big else skip
        t = self.tmp(cin, h, w)
        self.norm_elu(p + ".0.normalize1", s_, t)
        if big:
            self.free(s_)
        t2 = self.tmp(cin, h, w)
        self.conv(p + ".0.conv1", t, dst=t2, dil=d)
        self.free(t)
        t3 = self.tmp(cin, h, w)
        self.norm_elu(p + ".0.normalize2", t2, t3)
        self.free(t2)
        s_ = self.smem_of(skip) if big else skip
        if dil is None:
            out = self.tmp(cout, h // 2, w // 2, "o")
            self.conv(p + ".0.conv2.conv", t3, dst=out, pool=True, siblings=[(p + ".0.shortcut.conv", s_, 1)])
        else:
            out = self.tmp(cout, h, w, "o")
            self.conv(p + ".0.conv2", t3, dst=out, dil=d, siblings=[(p + ".0.shortcut", s_, d)])
        if big:
            self.free(s_)
        self.free(t3)
        e = None
        if want_elu:
            c, h2, w2 = self.shape[out]
            e = self.tmp(c, h2, w2, "e")
        return self.residual(p + ".1", out, cout, False, dil, elu_out=e), e


def build_program(state: Dict[str, np.ndarray], ngf: int, H: int, W: int, channels: int = 2,
                  nthreads: int = 256, precision: str = "tf32x3", park: Optional[bool] = None) -> Program:
    """``park=None``: plan for two resident CTAs per SM when that plan fits (``smem_budget_bytes(2)``), else the
    single-CTA plan (which itself falls back to a global-memory arena in the library when it exceeds one SM)."""
    if park is None:
        pp = ProgramBuilder(state, ngf, H, W, channels, nthreads, precision, park=True).build()
        if pp.smem_bytes() <= smem_budget_bytes(2):
            return pp
        p1 = ProgramBuilder(state, ngf, H, W, channels, nthreads, precision, park=False).build()
        # one CTA per SM: the plain plan when it fits an SM, else the park plan when THAT fits (shared-memory resident
        # beats the global-memory arena), else the plain plan from a global-memory arena
        if p1.smem_bytes() > smem_budget_bytes(1) and pp.smem_bytes() <= smem_budget_bytes(1):
            return pp
        return p1
    return ProgramBuilder(state, ngf, H, W, channels, nthreads, precision, park=park).build()


# ---------------------------------------------------------------------------
# torch (CPU) interpreter of a Program -- host-side check of the schedule only
# ---------------------------------------------------------------------------
def conv_weights(prog: Program, op: Op, branch: int = 0):
    """Decode (weight [cout,cin,k,k], bias or None) of (one summed branch of) a conv op back from the packed blob;
    the bias (already summed over the branches) is returned with branch 0 only."""
    import torch
    blob = torch.from_numpy(prog.blob)
    m = op.branches[branch]
    k, live, KC, NT = m["k"], m["live"], m["KC"], op.NT
    E = 2
    f0 = op.w_off + op.frag_rel + m["step0"] * NT * 32 * E
    n = len(live) * KC * NT * 32 * E
    frag = blob[f0:f0 + n].view(len(live), KC, NT, 32, E)
    full = torch.zeros(NT * 8, KC * 8, k * k)
    g, t = torch.arange(32) >> 2, torch.arange(32) & 3
    for i, tap in enumerate(live):
        for kc in range(KC):
            for nt in range(NT):
                full[nt * 8 + g, kc * 8 + t, tap] = frag[i, kc, nt, :, 0]
                full[nt * 8 + g, kc * 8 + t + 4, tap] = frag[i, kc, nt, :, 1]
    wt = full[:op.cout, :m["cin"]].reshape(op.cout, m["cin"], k, k).contiguous()
    bias = blob[op.w_off + op.b_rel:op.w_off + op.b_rel + op.cout] if (op.b_rel >= 0 and branch == 0) else None
    return wt, bias


def simulate(prog: Program, x, upto: Optional[int] = None):
    """Run the program on one sample ``x`` ([channels,H,W] float32 torch tensor) with torch CPU ops (fp32
    arithmetic; in "tf32" mode the decoded weights carry their TF32 rounding).

    Returns (raw network output [channels,H,W] (before the /sigma of ncsnv2.py:295-298), arena); the park area of a
    two-CTAs-per-SM plan is attached to the arena tensor as ``arena.park``."""
    import torch
    import torch.nn.functional as F

    arena = torch.zeros(prog.arena_floats, dtype=torch.float32)
    park = torch.zeros(max(prog.park_floats, 1), dtype=torch.float32)
    blob = torch.from_numpy(prog.blob)
    rd = lambda off, c, h, w: prog.read(arena, off, c, h, w)
    prog.write_input(arena, x.float())
    for i, op in enumerate(prog.ops):
        if upto is not None and i >= upto:
            break
        if op.kind == OP_SPILL:
            park[op.dst:op.dst + 4 * op.MT] = arena[op.src:op.src + 4 * op.MT]
        elif op.kind == OP_FILL:
            arena[op.dst:op.dst + 4 * op.MT] = park[op.src:op.src + 4 * op.MT]
        elif op.kind == OP_AFFINE:
            xin = arena[op.src:op.src + op.cin * op.h * op.w].view(op.h, op.w, op.cin).permute(2, 0, 1)
            prog.write(arena, op.dst, 2 * xin - 1.0, cstore=op.cout)
        elif op.kind == OP_ELU:
            prog.write(arena, op.dst, F.elu(rd(op.src, op.cin, op.h, op.w)))
        elif op.kind == OP_MAXPOOL5:
            prog.write(arena, op.dst, F.max_pool2d(rd(op.src, op.cin, op.h, op.w)[None], 5, 1, 2)[0])
        elif op.kind == OP_NORM_ELU:
            c = op.cin
            xs = rd(op.src, c, op.h, op.w)
            al, ga, be = blob[op.w_off:op.w_off + 3 * c].view(3, c)
            mu = xs.mean(dim=(1, 2))
            m, v = mu.mean(), mu.var()
            mh = (mu - m) / torch.sqrt(v + 1e-5)
            hn = (xs - mu[:, None, None]) / torch.sqrt(xs.var(dim=(1, 2), unbiased=False)[:, None, None] + 1e-5)
            o = ga[:, None, None] * (hn + (mh * al)[:, None, None]) + be[:, None, None]
            prog.write(arena, op.dst, F.elu(o))
        elif op.kind == OP_UPACC:
            s = rd(op.src, op.cin, op.h, op.w)
            a = rd(op.acc, op.cin, op.oh, op.ow)
            a = a + F.interpolate(s[None], size=(op.oh, op.ow), mode="bilinear", align_corners=True)[0]
            prog.write(arena, op.acc, a)
            if op.edst >= 0:
                prog.write(arena, op.edst, F.elu(a))
        elif op.kind == OP_CONV_MMA:
            v = None
            for bi, m in enumerate(op.branches):
                k = m["k"]
                wt, bias = conv_weights(prog, op, bi)
                s = rd(m["src_off"], m["cin"], op.h, op.w)[None]     # begin_conv: the real channels of the padded chunk
                vb = F.conv2d(s, wt, None, 1, m["dil"] * (k // 2), m["dil"])
                if op.flags & F_POOL:                                # packed weights already carry the 1/4: 2x2 SUM
                    vb = F.avg_pool2d(vb, 2) * 4.0
                if bias is not None:
                    vb = vb + bias[None, :, None, None]
                v = vb if v is None else v + vb
            v = v[0]
            if op.flags & F_COMPACT:
                arena[op.dst:op.dst + op.cout * op.oh * op.ow] = v.permute(1, 2, 0).reshape(-1)
            elif op.dst >= 0:
                prog.write(arena, op.dst, v)
            if op.acc >= 0:
                accb = park if (op.flags & F_ACC_G) else arena
                v = prog.read(accb, op.acc, op.cout, op.oh, op.ow) + v
                prog.write(accb, op.acc, v)
            if op.edst >= 0:
                prog.write(arena, op.edst, F.elu(v))
        else:
            raise ValueError(op.kind)
    out = prog.read_output(arena).clone()
    arena.park = park
    return out, arena
