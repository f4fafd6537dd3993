"""No-GPU checks of the boundary: the C-ABI library loads, exports every symbol the header declares,
refuses to run without a B200, and the weight container round-trips."""
import ctypes as C
import os

import numpy as np
import pytest

from ocr_inference_b200 import _lib, weights


def test_library_exports_every_header_symbol():
    lib = _lib.load()
    syms = _lib.header_symbols()
    assert len(syms) >= 30
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing
    assert set(syms) == set(_lib._SIGS), set(syms) ^ set(_lib._SIGS)
    assert b"sm_100a" in lib.bdrc_version()


def test_header_cites_reference_lines():
    src = open(_lib.HEADER_PATH).read()
    for cite in ("inference.py:132-139", "inference.py:152-160", "inference.py:175-181", "line_detection.py:330-359",
                 "inference.py:313-357", "inference.py:359-370", "inference.py:372-384", "line_detection.py:62-76"):
        assert cite in src, cite
    # no torch types in any signature: only C scalars, void*/uint8_t*/float*/int32_t*/int64_t*/double*
    import re
    code = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    assert "torch::" not in code and "at::" not in code and "Tensor" not in code and "#include <torch" not in src


@pytest.mark.skipif(os.path.exists("/dev/nvidia0"), reason="GPU present")
def test_no_cpu_fallback_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.bdrc_create(0, C.byref(h)) != 0 and not h.value
    with pytest.raises(_lib.BdrcError):
        _lib.Context(0)
    from ocr_inference_b200.engine import Engine
    with pytest.raises(_lib.BdrcError):
        Engine(0)


def test_null_ctx_is_an_error_not_a_crash():
    lib = _lib.load()
    assert lib.bdrc_seg_preprocess(None, None, 1, 1, 512, None, None) != 0
    assert lib.bdrc_ctc_greedy(None, None, 1, 1, 1, 0, 0, None, None, None) != 0
    assert lib.bdrc_launch_count(None) == 0


def test_weight_container_roundtrip():
    rng = np.random.default_rng(0)
    t = {"a.weight": rng.standard_normal((4, 3, 7, 7)).astype(np.float32), "a.bias": rng.standard_normal(4).astype(np.float32),
         "fc.weight": rng.standard_normal((5, 9)).astype(np.float32)}
    blob = weights.write_blob(t)
    assert blob[:8] == weights.MAGIC
    back = weights.read_blob(blob)
    assert list(back) == list(t)
    for k in t:
        assert np.array_equal(back[k], t[k])
    with pytest.raises(ValueError):
        weights.read_blob(b"not a blob at all")


def test_product_package_never_imports_the_oracle():
    root = os.path.dirname(_lib.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_host_gather_packs_ragged_ranges():
    """bdrc_host_gather (host-only entry point, runs without a GPU): ragged byte ranges land at their offsets,
    whatever the thread count; empty ranges and n = 0 are fine; bad arguments are refused."""
    import numpy as np

    from ocr_inference_b200 import _lib

    lib = _lib.load()
    rng = np.random.default_rng(3)
    srcs = [rng.integers(0, 256, int(n), dtype=np.uint8) for n in (3_000_001, 0, 17, 5_242_880, 1, 2_000_000)]
    nb = np.array([s.size for s in srcs], dtype=np.int64)
    off = np.zeros(len(srcs), dtype=np.int64)
    off[1:] = np.cumsum((nb[:-1] + 15) // 16 * 16)
    ptrs = np.array([s.ctypes.data for s in srcs], dtype=np.uint64)
    for threads in (1, 3, 8, 64):
        dst = np.full(int(off[-1] + nb[-1]) + 32, 0xAB, np.uint8)
        assert lib.bdrc_host_gather(ptrs.ctypes.data, nb.ctypes.data, off.ctypes.data, len(srcs), dst.ctypes.data,
                                    threads) == 0
        for s, o in zip(srcs, off):
            assert np.array_equal(dst[o:o + s.size], s)
        assert (dst[-32:] == 0xAB).all()
    assert lib.bdrc_host_gather(None, None, None, 0, None, 4) == 0
    bad = nb.copy()
    bad[2] = -1
    assert lib.bdrc_host_gather(ptrs.ctypes.data, bad.ctypes.data, off.ctypes.data, len(srcs), dst.ctypes.data, 2) != 0
