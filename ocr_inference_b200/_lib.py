"""ctypes binding of libbdrc_b200.so (the C ABI declared in include/bdrc_b200.h).

There is no CPU fallback: if the library is missing or no B200 is visible, creating a
context raises.  PyTorch is used by callers only to own device memory and streams.
"""
from __future__ import annotations

import ctypes as C
import os
import re

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbdrc_b200.so")
HEADER_PATH = os.path.join(_HERE, "..", "include", "bdrc_b200.h")

c_void_p, c_int, c_size_t, c_float, c_double = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_double
c_u64 = C.c_uint64

_SIGS = {
    # name: (restype, [argtypes])
    "bdrc_create": (c_int, [c_int, C.POINTER(c_void_p)]),
    "bdrc_destroy": (None, [c_void_p]),
    "bdrc_last_error": (C.c_char_p, [c_void_p]),
    "bdrc_version": (C.c_char_p, []),
    "bdrc_launch_count": (c_u64, [c_void_p]),
    "bdrc_simt_launch_count": (c_u64, [c_void_p]),
    "bdrc_set_blocking_sync": (c_int, [c_void_p, c_int]),
    "bdrc_set_dec_tail": (c_int, [c_void_p, c_int]),
    "bdrc_set_wide_conv": (c_int, [c_void_p, c_int]),
    "bdrc_conv_plans_built": (c_u64, [c_void_p]),
    "bdrc_set_conv_engine": (c_int, [c_void_p, c_int]),
    "bdrc_profile_enable": (c_int, [c_void_p, c_int]),
    "bdrc_profile_reset": (c_int, [c_void_p]),
    "bdrc_profile_count": (c_int, [c_void_p]),
    "bdrc_profile_get": (c_int, [c_void_p, c_int, C.POINTER(C.c_char_p), C.POINTER(c_float),
                                 C.POINTER(c_double)]),
    "bdrc_host_gather": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int]),
    "bdrc_host_copy": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int]),
    "bdrc_load_weights": (c_int, [c_void_p, c_int, c_void_p, c_size_t, c_int, C.POINTER(c_void_p)]),
    "bdrc_free_model": (None, [c_void_p]),
    "bdrc_seg_preprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p]),
    "bdrc_seg_preprocess_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p]),
    "bdrc_seg_tile_rgb": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                  c_void_p]),
    "bdrc_seg_workspace_bytes": (c_size_t, [c_void_p, c_int]),
    "bdrc_seg_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "bdrc_seg_forward_lowres": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                                        c_void_p, c_size_t, c_void_p]),
    "bdrc_seg_postprocess": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int,
                                     c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "bdrc_seg_postprocess_lowres": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                            c_int, c_int, c_float, c_void_p, c_void_p]),
    "bdrc_seg_postprocess_lowres_batch": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int,
                                                  c_int, c_int, c_float, c_void_p, c_void_p]),
    "bdrc_seg_postprocess_torch": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                           c_int, c_int, c_int, c_float, c_void_p, c_void_p]),
    "bdrc_resize_bilinear_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                        c_int, c_void_p]),
    "bdrc_resize_bilinear_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_int,
                                         c_int, c_void_p]),
    "bdrc_warp_affine_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                    c_void_p]),
    "bdrc_remap_bilinear_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p]),
    "bdrc_tps_workspace_bytes": (c_size_t, []),
    "bdrc_tps_remap_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                  c_void_p, c_void_p, c_size_t, c_void_p]),
    "bdrc_spline3_coef_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bdrc_spline3_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bdrc_spline3_prefilter_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t,
                                          c_void_p]),
    "bdrc_remap_cubic_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "bdrc_tps_remap_cubic_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int,
                                        c_void_p, c_void_p, c_size_t, c_void_p]),
    "bdrc_extract_lines_workspace_bytes": (c_size_t, [c_int, c_int, c_int, C.c_int64]),
    "bdrc_line_crop_heights": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                       c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_size_t, c_void_p]),
    "bdrc_extract_lines": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p,
                                   c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p,
                                   c_void_p, c_void_p, c_size_t, c_void_p]),
    "bdrc_ocr_prepare_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bdrc_ocr_prepare_lines": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int,
                                       c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t,
                                       c_void_p]),
    "bdrc_ocr_workspace_bytes": (c_size_t, [c_void_p, c_int, c_int]),
    "bdrc_ocr_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                 c_size_t, c_void_p]),
    "bdrc_mask_census_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "bdrc_census_max_centers": (c_int, []),
    "bdrc_mask_census": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t,
                                 c_void_p]),
    "bdrc_mask_components_workspace_bytes": (c_size_t, [c_int, c_int]),
    "bdrc_mask_components": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_size_t,
                                     c_void_p]),
    "bdrc_ctc_topk": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                              c_void_p]),
    "bdrc_ctc_topk_width": (c_int, []),
    "bdrc_ctc_greedy": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                c_void_p, c_void_p]),
    "bdrc_depthwise3x3_test": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "bdrc_conv2d_test": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p,
                                 c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p, c_int, c_int, c_void_p, c_void_p]),
}

_lib = None


def header_symbols() -> list[str]:
    """Every function name declared in include/bdrc_b200.h."""
    src = open(HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bdrc_[a-z0-9_]+)\s*\(", src)))


def load() -> C.CDLL:
    """Load the shared library (building is __graft_entry__.build()'s job). Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: run `python -m ocr_inference_b200.build` "
            "(there is no CPU fallback for the B200 hot path)")
    lib = C.CDLL(LIB_PATH)
    missing = []
    for name, (res, args) in _SIGS.items():
        fn = getattr(lib, name, None)
        if fn is None:
            missing.append(name)
            continue
        fn.restype = res
        fn.argtypes = args
    if missing and os.environ.get("BDRC_ALLOW_PARTIAL_LIB") != "1":
        raise RuntimeError(f"{LIB_PATH} is stale: missing symbols {missing}; rebuild it")
    _lib = lib
    return lib


class BdrcError(RuntimeError):
    pass


class Context:
    """One bdrc_ctx per GPU. Raises BdrcError on any non-zero status."""

    def __init__(self, device: int = 0):
        self.lib = load()
        h = c_void_p()
        rc = self.lib.bdrc_create(device, C.byref(h))
        if rc != 0 or not h.value:
            raise BdrcError(
                f"bdrc_create(device={device}) failed with code {rc}: a B200 (sm_100a) GPU is "
                "required; this package has no CPU fallback")
        self.h = h
        self.device = device

    def check(self, rc: int, what: str = ""):
        if rc != 0:
            msg = self.lib.bdrc_last_error(self.h)
            raise BdrcError(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")

    def launch_count(self) -> int:
        return int(self.lib.bdrc_launch_count(self.h))

    def simt_launch_count(self) -> int:
        return int(self.lib.bdrc_simt_launch_count(self.h))

    def conv_plans_built(self) -> int:
        return int(self.lib.bdrc_conv_plans_built(self.h))

    def close(self):
        if getattr(self, "h", None) is not None and self.h.value:
            self.lib.bdrc_destroy(self.h)
            self.h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
