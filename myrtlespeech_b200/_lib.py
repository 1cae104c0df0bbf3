"""ctypes binding of the C-ABI shared library (``include/rnnt_b200.h``).

The library is the product: there is no CPU or eager fallback.  If it has not
been built (``python -c "import __graft_entry__ as g; g.build()"`` or
``python -m myrtlespeech_b200.build``) every entry point raises.
"""
import ctypes
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RNNT_LIB_PATH") or os.path.join(_HERE, "lib", "librnnt_b200.so")   # override: bring-up builds

_c_int = ctypes.c_int
_c_size_t = ctypes.c_size_t
_vp = ctypes.c_void_p

#: symbol -> (restype, argtypes); mirrors include/rnnt_b200.h one to one
SIGNATURES = {
    "rnnt_abi_version": (_c_int, []),
    "rnnt_last_error": (ctypes.c_char_p, []),
    "rnnt_fused_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "rnnt_fused_state_bytes": (_c_size_t, [_c_int] * 5),
    "rnnt_fused_forward": (_c_int, [_vp] * 7 + [_c_int] * 6 + [_vp, _vp, _c_size_t, _vp]),
    "rnnt_fused_backward": (_c_int, [_vp] * 7 + [_c_int] * 6 + [_vp] * 5 + [_vp, _c_size_t, _vp]),
    "rnnt_fused_kept_bytes": (_c_size_t, [_c_int] * 5),
    "rnnt_fused_forward_keep": (_c_int, [_vp] * 7 + [_c_int] * 6 + [_vp, _vp, _c_size_t, _vp, _c_size_t, _vp]),
    "rnnt_fused_backward_kept": (_c_int, [_vp] * 7 + [_c_int] * 6 + [_vp] * 5 + [_vp, _c_size_t, _vp, _c_size_t, _vp]),
    "rnnt_lattice_workspace_bytes": (_c_size_t, [_c_int] * 3),
    "rnnt_lattice_forward": (_c_int, [_vp] * 4 + [_c_int] * 3 + [_vp] * 3 + [_vp, _c_size_t, _vp]),
    "rnnt_greedy_joint_argmax": (_c_int, [_vp] * 6 + [_c_int] * 4 + [_vp]),
    "rnnt_greedy_step": (_c_int, [_vp] * 9 + [_c_int] + [_vp] * 3 + [_c_int] * 6 + [_vp]),
    "rnnt_greedy_decode_workspace_bytes": (_c_size_t, [_c_int] * 4),
    "rnnt_greedy_decode_lstm": (_c_int, [_vp] * 8 + [_c_int] * 7 + [_vp, _c_int, _vp, _vp, _c_size_t, _vp]),
    "rnnt_greedy_decode_stack_workspace_bytes": (_c_size_t, [_c_int] * 5),
    "rnnt_greedy_decode_lstm_stack": (_c_int, [_vp] * 6 + [_c_int] + [_vp] * 4 + [_c_int] * 7 + [_vp, _c_int, _vp, _vp, _c_size_t, _vp]),
    "rnnt_greedy_decode_gru_stack": (_c_int, [_vp] * 6 + [_c_int] + [_vp] * 4 + [_c_int] * 7 + [_vp, _c_int, _vp, _vp, _c_size_t, _vp]),
    "rnnt_debug_copy_stats": (_c_int, [_vp] + [_c_int] * 5 + [_vp] * 5 + [_vp]),
    "rnnt_debug_set": (None, [ctypes.c_char_p, _c_int]),
    "rnnt_debug_get": (ctypes.c_longlong, [ctypes.c_char_p]),
    "rnnt_debug_kernel_times": (_c_int, [_vp, _vp, _c_int]),
    "rnnt_debug_read_prof": (_c_int, [_vp, _c_int]),
    "rnnt_debug_read_prof3": (_c_int, [_vp, _c_int, _c_int]),
    "rnnt_debug_read_active_tiles": (_c_int, [_vp] + [_c_int] * 5 + [_vp]),
    "rnnt_debug_decode_prof": (_c_int, [_vp, _c_int]),
}

_lib: Optional[ctypes.CDLL] = None


class RNNTLibraryError(RuntimeError):
    """Raised when the CUDA library is missing or a C-ABI call returns non-zero."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RNNTLibraryError(
                f"{LIB_PATH} not found: the CUDA extension is not built and there is no fallback "
                "(run `python -m myrtlespeech_b200.build`)"
            )
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        msg = load().rnnt_last_error().decode()
        if rc in (1, 4):
            raise ValueError(msg)
        raise RNNTLibraryError(f"rnnt_b200 error {rc}: {msg}")
