"""The C-ABI library loads and exports every symbol include/nnc.h declares (no compute without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "nnc.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nnc_[a-z0-9_]+)\s*\(", src)) - {"nnc_allreduce_i64_fn"})


def test_library_exports_every_declared_symbol():
    from neural_network_compression_b200 import _native as N
    from neural_network_compression_b200 import build

    build.build()
    L = ctypes.CDLL(N.LIB_PATH)
    names = _declared()
    assert len(names) >= 20
    for name in names:
        assert hasattr(L, name), name
    assert sorted(N.EXPORTS) == names
    assert N.lib().nnc_version() == 100


def test_struct_layouts_match_the_header(tmp_path):
    """The ctypes mirrors of nnc_kmeans_info and nnc_tensor_job have the size and field offsets a C compiler gives the
    declarations in include/nnc.h (the binding a maintainer of the reference would write is exactly this)."""
    import subprocess

    from neural_network_compression_b200 import _native as N

    fields = ["w", "n", "threshold", "prune", "mask", "centers", "centred", "packed", "hist", "info", "thr", "n_pruned", "k",
              "code_bits", "status", "error"]
    src = tmp_path / "layout.c"
    src.write_text(
        '#include <stdio.h>\n#include <stddef.h>\n#include "nnc.h"\nint main(void) {\n'
        '  printf("%zu %zu\\n", sizeof(nnc_kmeans_info), sizeof(nnc_tensor_job));\n'
        + "".join('  printf("%%zu\\n", offsetof(nnc_tensor_job, %s));\n' % f for f in fields)
        + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = subprocess.check_output([str(exe)]).decode().split()
    assert int(out[0]) == ctypes.sizeof(N.KMeansInfo)
    assert int(out[1]) == ctypes.sizeof(N.TensorJob)
    for f, off in zip(fields, out[2:]):
        assert getattr(N.TensorJob, f).offset == int(off), f


def test_no_cpu_fallback_without_device():
    import torch

    from neural_network_compression_b200 import _native as N

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(N.NncError) as e:
        N.Context(0)
    assert e.value.code == N.NNC_ERR_CUDA and "no CPU fallback" in str(e.value)
    import numpy as np

    from neural_network_compression_b200.common import utility

    with pytest.raises(N.NncError):
        utility.prune_weigth(np.ones(10, dtype=np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "neural_network_compression_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text and "nnc_oracle" not in text, f


def test_reference_argument_errors():
    import numpy as np

    from neural_network_compression_b200.common import utility

    w = np.ones(100, dtype=np.float32)
    with pytest.raises(Exception, match="error mode not found"):
        utility.get_quantized_weight(w, 2, "unknown")
    with pytest.raises(Exception, match="error mode not found"):
        utility.get_quantized_weight(w, 2, "density", None)
    small = np.ones(4, dtype=np.float32)
    out, km = utility.get_quantized_weight(small, 2, "linear")  # n < 2^bits + 1 (utility.py:202-204)
    assert out is small and km is None
    # the whole-model calls validate their arguments before touching a device
    with pytest.raises(ValueError, match="one threshold per tensor"):
        utility.compress_model([w, w], [0.5], True, 4, "linear")
    with pytest.raises(Exception, match="error mode not found"):
        utility.compress_model([w], [0.5], True, 4, "forgy")
    with pytest.raises(ValueError, match="all Python floats or all numpy.float64"):
        utility.compress_model([w, w], [0.5, np.float64(0.5)], True, 4, "linear")
    with pytest.raises(TypeError, match="float32"):
        utility.compress_model([w.astype(np.float64)], [0.5], True, 4, "linear")
    assert utility.compress_model([], None) == []
