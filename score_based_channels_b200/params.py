"""Parameter catalogue of ``NCSNv2Deepest`` (reference ``ncsnv2/models/ncsnv2.py:198-262``).

``param_shapes`` lists every ``state_dict`` entry of the reference module (same key names, same
shapes -- SURVEY.md Appendix A) so that checkpoints written by the reference load unchanged, and
``random_state`` draws a reproducible synthetic parameter set (numpy ``default_rng``; identical on
every machine) for tests and benchmarks that must not depend on the shipped checkpoint.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np


def get_sigmas_np(sigma_begin: float, sigma_end: float, num_classes: int) -> np.ndarray:
    """Geometric schedule, computed in float64 then cast to float32
    (reference ``ncsnv2/models/__init__.py:4-8``)."""
    return np.exp(np.linspace(np.log(sigma_begin), np.log(sigma_end), num_classes)).astype(np.float32)


ARCHS = ("deepest", "deeper", "ncsnv2")     # NCSNv2Deepest / NCSNv2Deeper / NCSNv2 (reference ncsnv2/models/ncsnv2.py)


def param_shapes(ngf: int, channels: int = 2, num_classes: int = 2311, arch: str = "deepest") -> "OrderedDict[str, Tuple[int, ...]]":
    if arch not in ARCHS:
        raise ValueError("arch must be one of %s" % (ARCHS,))
    d: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    d["sigmas"] = (num_classes,)

    def conv(p, cin, cout, k=3, bias=True):
        d[p + ".weight"] = (cout, cin, k, k)
        if bias:
            d[p + ".bias"] = (cout,)

    def norm(p, c):
        for leaf in ("alpha", "gamma", "beta"):
            d["%s.%s" % (p, leaf)] = (c,)

    conv("begin_conv", channels, ngf)
    norm("normalizer", ngf)
    conv("end_conv", ngf, channels)

    # encoder: ResidualBlock parameter sets (layers.py:401-441)
    def resblock(p, cin, cout, down, dil):
        c1 = cin if down else cout
        conv(p + ".conv1", cin, c1)
        norm(p + ".normalize2", c1)
        if down and dil is None:
            conv(p + ".conv2.conv", c1, cout)
            conv(p + ".shortcut.conv", cin, cout, k=1)
        else:
            conv(p + ".conv2", c1, cout)
            if cout != cin or down:
                conv(p + ".shortcut", cin, cout)
        norm(p + ".normalize1", cin)

    # stage plans per architecture (ncsnv2.py:26-58 NCSNv2, 118-150 NCSNv2Deeper, 218-262 NCSNv2Deepest)
    if arch == "deepest":
        plan = [("res1", ngf, ngf, False, None), ("res2", ngf, 2 * ngf, True, None),
                ("res3", 2 * ngf, 2 * ngf, True, None), ("res31", 2 * ngf, 2 * ngf, True, None),
                ("res4", 2 * ngf, 4 * ngf, True, 2), ("res5", 4 * ngf, 4 * ngf, True, 4)]
    elif arch == "deeper":
        plan = [("res1", ngf, ngf, False, None), ("res2", ngf, 2 * ngf, True, None),
                ("res3", 2 * ngf, 2 * ngf, True, None), ("res4", 2 * ngf, 4 * ngf, True, 2),
                ("res5", 4 * ngf, 4 * ngf, True, 4)]
    else:
        plan = [("res1", ngf, ngf, False, None), ("res2", ngf, 2 * ngf, True, None),
                ("res3", 2 * ngf, 2 * ngf, True, 2), ("res4", 2 * ngf, 2 * ngf, True, 4)]
    for p, cin, cout, down, dil in plan:
        resblock(p + ".0", cin, cout, down, dil)
        resblock(p + ".1", cout, cout, False, dil)

    # decoder: RefineBlock parameter sets (layers.py:215-232)
    def refine(p, in_planes, features, start=False, end=False):
        for i, c in enumerate(in_planes):
            for b in (1, 2):
                for s in (1, 2):
                    conv("%s.adapt_convs.%d.%d_%d_conv" % (p, i, b, s), c, c, bias=False)
        for b in range(1, (3 if end else 1) + 1):
            for s in (1, 2):
                conv("%s.output_convs.%d_%d_conv" % (p, b, s), features, features, bias=False)
        if not start:
            for i, c in enumerate(in_planes):
                conv("%s.msf.convs.%d" % (p, i), c, features)
        for i in range(2):
            conv("%s.crp.convs.%d" % (p, i), features, features, bias=False)

    if arch == "deepest":
        refine("refine1", [4 * ngf], 4 * ngf, start=True)
        refine("refine2", [4 * ngf, 4 * ngf], 2 * ngf)
        refine("refine3", [2 * ngf, 2 * ngf], 2 * ngf)
        refine("refine31", [2 * ngf, 2 * ngf], 2 * ngf)
        refine("refine4", [2 * ngf, 2 * ngf], ngf)
        refine("refine5", [ngf, ngf], ngf, end=True)
    elif arch == "deeper":
        refine("refine1", [4 * ngf], 4 * ngf, start=True)
        refine("refine2", [4 * ngf, 4 * ngf], 2 * ngf)
        refine("refine3", [2 * ngf, 2 * ngf], 2 * ngf)
        refine("refine4", [2 * ngf, 2 * ngf], ngf)
        refine("refine5", [ngf, ngf], ngf, end=True)
    else:
        refine("refine1", [2 * ngf], 2 * ngf, start=True)
        refine("refine2", [2 * ngf, 2 * ngf], 2 * ngf)
        refine("refine3", [2 * ngf, 2 * ngf], ngf)
        refine("refine4", [ngf, ngf], ngf, end=True)
    return d


def random_state(ngf: int = 8, channels: int = 2, num_classes: int = 2311, sigma_begin: float = 27.77,
                 sigma_end: float = 2.599515446446343e-4, seed: int = 0, arch: str = "deepest") -> Dict[str, np.ndarray]:
    """Synthetic, machine-independent parameter set with every bias / affine term non-trivial."""
    rng = np.random.default_rng(seed)
    out: Dict[str, np.ndarray] = OrderedDict()
    for k, shp in param_shapes(ngf, channels, num_classes, arch).items():
        if k == "sigmas":
            out[k] = get_sigmas_np(sigma_begin, sigma_end, num_classes)
        elif k.endswith(".weight"):
            fan_in = shp[1] * shp[2] * shp[3]
            out[k] = (rng.standard_normal(shp) * (1.0 / np.sqrt(fan_in))).astype(np.float32)
        elif k.endswith(".bias"):
            out[k] = (rng.standard_normal(shp) * 0.05).astype(np.float32)
        elif k.endswith(".alpha") or k.endswith(".gamma"):
            out[k] = (1.0 + 0.05 * rng.standard_normal(shp)).astype(np.float32)
        elif k.endswith(".beta"):
            out[k] = (0.05 * rng.standard_normal(shp)).astype(np.float32)
        else:
            raise KeyError(k)
    return out
