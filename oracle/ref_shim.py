"""Loader for the UNMODIFIED reference helpers.

TEST INFRASTRUCTURE ONLY.  Works only where /root/reference exists (the build container); it is
used by tests/golden/make_golden.py to generate the committed golden vectors and by the optional
`-m "not gpu"` differential tests, which skip when the reference tree is absent (the GPU box).

The reference cannot be imported as a package here (matplotlib / tensorflow missing, and sklearn
>= 1.1 renamed algorithm="full" to "lloyd"), so utility.py is loaded by path behind two stubs.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REF_UTILITY = "/root/reference/neural_network_compression/common/utility.py"


def available() -> bool:
    return os.path.exists(REF_UTILITY)


_ref = None


def load():
    global _ref
    if _ref is not None:
        return _ref
    import sklearn.cluster

    if "matplotlib" not in sys.modules:
        mp, pp = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
        mp.pyplot = pp
        sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mp, pp
    spec = importlib.util.spec_from_file_location("ref_utility", REF_UTILITY)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    _KM = sklearn.cluster.KMeans

    def _kmeans(*a, **k):
        if k.get("algorithm") == "full":
            k["algorithm"] = "lloyd"
        return _KM(*a, **k)

    ref.KMeans = _kmeans
    _ref = ref
    return ref
