"""ctypes binding of libgimb200.so (include/gimb200.h).  There is NO fallback: if the CUDA library is
missing or fails to load, importing the matcher raises."""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_size_t, c_uint64, c_uint8, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgimb200.so")


class LoftrCfg(Structure):
    _fields_ = [("thr", c_float), ("border_rm", c_int32), ("dsmax_temperature", c_float), ("fine_window", c_int32)]


class LoftrOut(Structure):
    _fields_ = [("capacity", c_int64), ("b_ids", c_void_p), ("i_ids", c_void_p), ("j_ids", c_void_p),
                ("mconf", c_void_p), ("mkpts0_c", c_void_p), ("mkpts1_c", c_void_p), ("mkpts0_f", c_void_p),
                ("mkpts1_f", c_void_p), ("expec_f", c_void_p)]


class LoftrTaps(Structure):
    _fields_ = [(k, c_void_p) for k in ("feat_c_backbone0", "feat_c_backbone1", "feat_f0", "feat_f1", "feat_c0",
                                        "feat_c1", "fine_win0", "fine_win1", "conf_matrix")]


EXPORTS = {
    # name: (restype, argtypes)
    "gimb_last_error": (c_char_p, []),
    "gimb_abi_version": (c_int, []),
    "gimb_loftr_create": (c_int, [c_void_p, c_size_t, POINTER(LoftrCfg), c_int, POINTER(c_void_p)]),
    "gimb_loftr_destroy": (None, [c_void_p]),
    "gimb_loftr_workspace_bytes": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "gimb_loftr_set_pe": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "gimb_loftr_forward": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, POINTER(LoftrOut),
                                   POINTER(LoftrTaps), POINTER(c_int64), c_void_p]),
    "gimb_loftr_forward_host": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                        c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_size_t,
                                        POINTER(LoftrOut), POINTER(LoftrOut), POINTER(c_int64), POINTER(c_uint64),
                                        POINTER(c_uint64), c_void_p]),
    "gimb_loftr_host_staging_bytes": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_size_t)]),
    "gimb_loftr_host_u8_staging_bytes": (c_int, [c_int] * 10 + [POINTER(c_size_t)]),
    "gimb_loftr_forward_host_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                           c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p, c_size_t,
                                           POINTER(LoftrOut), POINTER(LoftrOut), POINTER(c_int64), POINTER(c_uint64),
                                           POINTER(c_uint64), c_void_p]),
    "gimb_loftr_stage_host_u8": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_void_p,
                                         c_int, c_int, c_int, c_int, c_int, c_void_p, c_size_t, POINTER(c_uint64), c_void_p]),
    "gimb_loftr_forward_staged_u8": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                             c_void_p, c_size_t, c_void_p, c_size_t, POINTER(LoftrOut), POINTER(LoftrOut),
                                             POINTER(c_int64), POINTER(c_uint64), c_void_p]),
    "gimb_loftr_launch_count": (c_uint64, [c_void_p]),
    "gimb_loftr_corr_fallbacks": (c_uint64, [c_void_p]),
    "gimb_loftr_set_profiling": (c_int, [c_void_p, c_int]),
    "gimb_loftr_set_engine": (c_int, [c_void_p, c_int]),
    "gimb_loftr_last_profile": (c_int, [c_void_p, POINTER(c_char_p), POINTER(c_float), POINTER(c_int)]),
    # ---- gim_dkm (include/gimb200.h, csrc/dkm_api.cu)
    "gimb_dkm_create": (c_int, [c_char_p, c_size_t, c_int, POINTER(c_void_p)]),
    "gimb_dkm_destroy": (None, [c_void_p]),
    "gimb_dkm_set_engine": (c_int, [c_void_p, c_int]),
    "gimb_dkm_launch_count": (c_uint64, [c_void_p]),
    "gimb_dkm_workspace_bytes": (c_int, [c_void_p] + [c_int] * 9 + [POINTER(c_size_t)]),
    "gimb_kde_density": (c_int, [c_void_p, c_int, c_float, c_void_p, c_void_p]),
    "gimb_dkm_match": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                               c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p]),
}

# libgimb200_test.so = the product library + the layer-level test / measurement hooks (include/gimb200_test.h)
TEST_LIB_PATH = os.path.join(HERE, "libgimb200_test.so")
TEST_LIB_PATH = os.environ.get("GIMB_TEST_LIB", TEST_LIB_PATH)  # measurement only: A/B two builds
TEST_EXPORTS = {
    "gimb_bench_layer": (c_int, [c_int] * 11 + [POINTER(c_float), c_void_p]),
    "gimb_probe_tma": (c_int, [c_int, c_int, POINTER(c_float), c_void_p]),
    "gimb_test_conv": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                               c_void_p, c_void_p, c_size_t, c_void_p]),
}

_lib = None
_test_lib = None


def load():
    """Load (once) and return the ctypes handle with prototypes set."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m gim_b200.build` (nvcc, sm_100a). "
            "gim_b200 has no CPU or PyTorch fallback.")
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in EXPORTS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export it
        fn.restype, fn.argtypes = res, args
    if lib.gimb_abi_version() != 1:
        raise RuntimeError("libgimb200 ABI version mismatch")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise RuntimeError("libgimb200: " + load().gimb_last_error().decode(errors="replace"))


def load_test():
    """The test-hook superset library (tests/test_umma_gpu.py, tools/*): product symbols + TEST_EXPORTS."""
    global _test_lib
    if _test_lib is not None:
        return _test_lib
    if not os.path.isfile(TEST_LIB_PATH):
        raise RuntimeError(f"{TEST_LIB_PATH} is missing: build it with `python -m gim_b200.build`")
    lib = ctypes.CDLL(TEST_LIB_PATH)
    for name, (res, args) in {**EXPORTS, **TEST_EXPORTS}.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _test_lib = lib
    return lib


def check_test(rc):
    if rc != 0:
        raise RuntimeError("libgimb200_test: " + load_test().gimb_last_error().decode(errors="replace"))
