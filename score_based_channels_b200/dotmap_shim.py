"""Minimal stand-in for ``dotmap.DotMap`` (dotmap==1.3.30 is a dependency of the
reference, ``pyproject.toml:9``, but is not installed in this image).

Only the behaviour the hot path relies on is provided:

* attribute and item access over an ordered mapping;
* *dynamic* children: reading a missing key returns a fresh, empty (falsy)
  ``DotMap`` and stores it.  ``NCSNv2Deepest.__init__`` reads
  ``config.data.logit_transform`` / ``config.data.rescaled`` this way
  (reference ``ncsnv2/models/ncsnv2.py:201-202``) and the resulting falsy
  values are what makes ``h = 2*x - 1`` active at ``ncsnv2.py:270-271``;
* pickle compatibility with the state the real class writes
  (``{'_map': OrderedDict, '_dynamic': True, '_prevent_method_masking': False}``)
  so that ``final_model.pt`` checkpoints written by the reference
  (``train_score.py:211-216``) can be ``torch.load``-ed.

Use :func:`install` to register this module as ``dotmap`` in ``sys.modules``
before unpickling a reference checkpoint.
"""
from __future__ import annotations

import sys
import types
from collections import OrderedDict


class DotMap(object):
    def __init__(self, *args, **kwargs):
        object.__setattr__(self, "_map", OrderedDict())
        object.__setattr__(self, "_dynamic", kwargs.pop("_dynamic", True))
        object.__setattr__(self, "_prevent_method_masking",
                           kwargs.pop("_prevent_method_masking", False))
        for a in args:
            if isinstance(a, DotMap):
                a = a._map
            if isinstance(a, dict):
                for k, v in a.items():
                    self[k] = v
        for k, v in kwargs.items():
            self[k] = v

    # -- mapping protocol -------------------------------------------------
    @staticmethod
    def _wrap(v):
        if isinstance(v, dict) and not isinstance(v, DotMap):
            return DotMap(v)
        return v

    def __setitem__(self, k, v):
        self._map[k] = self._wrap(v)

    def __getitem__(self, k):
        if k not in self._map and self._dynamic and k != "_ipython_canary_method_should_not_exist_":
            self._map[k] = DotMap()
        return self._map[k]

    def __setattr__(self, k, v):
        if k in ("_map", "_dynamic", "_prevent_method_masking"):
            object.__setattr__(self, k, v)
        else:
            self[k] = v

    def __getattr__(self, k):
        if k.startswith("__") and k.endswith("__"):
            raise AttributeError(k)
        if k in ("_map", "_dynamic", "_prevent_method_masking"):
            raise AttributeError(k)
        return self[k]

    def __delattr__(self, k):
        del self._map[k]

    def __contains__(self, k):
        return k in self._map

    def __len__(self):
        return len(self._map)

    def __bool__(self):
        return len(self._map) > 0

    def __iter__(self):
        return iter(self._map)

    def keys(self):
        return self._map.keys()

    def values(self):
        return self._map.values()

    def items(self):
        return self._map.items()

    def get(self, k, default=None):
        return self._map.get(k, default)

    def toDict(self):
        out = {}
        for k, v in self._map.items():
            out[k] = v.toDict() if isinstance(v, DotMap) else v
        return out

    def __repr__(self):
        return "DotMap(%s)" % ", ".join("%s=%r" % kv for kv in self._map.items())

    # -- copy / pickle ----------------------------------------------------
    def __getstate__(self):
        return {"_map": self._map, "_dynamic": self._dynamic,
                "_prevent_method_masking": self._prevent_method_masking}

    def __setstate__(self, d):
        object.__setattr__(self, "_map", d.get("_map", OrderedDict()))
        object.__setattr__(self, "_dynamic", d.get("_dynamic", True))
        object.__setattr__(self, "_prevent_method_masking",
                           d.get("_prevent_method_masking", False))

    def __deepcopy__(self, memo):
        import copy
        out = DotMap(_dynamic=self._dynamic)
        for k, v in self._map.items():
            out._map[k] = copy.deepcopy(v, memo)
        return out


def install():
    """Register this shim as the importable module ``dotmap`` (idempotent)."""
    if "dotmap" in sys.modules and getattr(sys.modules["dotmap"], "DotMap", None) is not None:
        return sys.modules["dotmap"]
    m = types.ModuleType("dotmap")
    m.DotMap = DotMap
    DotMap.__module__ = "dotmap"
    sys.modules["dotmap"] = m
    return m
