"""ctypes binding of libsparch_b200.so (the C ABI declared in include/sparch_b200.h).

This is the only place Python touches the native library.  There is no CPU or eager
fallback: if the library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libsparch_b200.so")

_P, _I, _L, _F = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_float

# name -> argument type string: p pointer, i int, l int64, f float (stream is a pointer)
_PROTOS = {
    "sparch_abi_version": "",
    "sparch_boxcar_fwd": "pplp",
    "sparch_boxcar_bwd": "ppplp",
    "sparch_col_stats": "plippp",
    "sparch_col_dot": "pppplipppp",
    "sparch_bn_bwd_apply_f16": "pppppppli" "ppppplpp",
    "sparch_bn_fold_train": "pplppffppppppip",
    "sparch_col_dot_bidir": "pppplii" + "lpppp",
    "sparch_bn_bwd_apply_bidir": "pppppppliip",
    "sparch_bn_bwd_apply_f16_bidir": "pppppppli" "i" "ppppplpp",
    "sparch_bn_bwd_apply": "pppppppli" "pp",
    "sparch_layernorm_fwd": "pppfliPPPp".replace("P", "p"),
    "sparch_layernorm_bwd_workspace": "li",
    "sparch_layernorm_bwd": "ppppplipppp" "p",
    "sparch_cell_fwd": "i" + "p" * 10 + "f" + "ppp" + "iii" + "p",
    "sparch_cell_step_fwd": "ii" + "p" * 11 + "f" + "ppp" + "iii" + "p",
    "sparch_cell_bwd": "i" + "p" * 10 + "f" + "p" * 5 + "iii" + "p",
    "sparch_cell_step_bwd": "ii" + "p" * 11 + "f" + "p" * 7 + "iii" + "p",
    "sparch_split_bf16": "pliiifppplp",
    "sparch_split_bf16_transpose": "piiiiifppplp",
    "sparch_gemm_workspace": "iii",
    "sparch_gemm_bf16": "pipill" + "iii" + "ppiiiifppl" + "pppp",
    "sparch_absmax": "pllipp",
    "sparch_split_f16": "pliiifpipplp",
    "sparch_gemm_terms": "ipippipll" + "iii" + "ppiiiifppl" + "pppp",
    "sparch_recur_padded": "i",
    "sparch_recur_prepare": "pipppp",
    "sparch_recur_sync_words": "i",
    "sparch_recur_debug_clocks": "p",
    "sparch_recur_fwd": "i" + "p" * 13 + "f" + "pppp" + "iiii" + "p",
    "sparch_recur_fwd_tc_max_h": "",
    "sparch_recur_fwd_tc_image_bytes": "i",
    "sparch_recur_fwd_tc_bits_bytes": "iii",
    "sparch_recur_prepare_fwd_tc": "pipp",
    "sparch_recur_fwd_tc": "i" + "p" * 12 + "f" + "pppp" + "iiii" + "p",
    "sparch_recur_fwd_tc_bidir": "i" + "p" * 12 + "f" + "pppp" + "iiiiii" + "p",
    "sparch_recur_bwd_workspace": "ii",
    "sparch_recur_bwd": "i" + "p" * 12 + "f" + "p" * 7 + "iiii" + "p",
    "sparch_recur_tc_padded": "i",
    "sparch_recur_bwd_tc_image_bytes": "i",
    "sparch_recur_bwd_tc_workspace": "iii",
    "sparch_recur_prepare_tc": "pippp",
    "sparch_recur_bwd_tc": "i" + "p" * 12 + "f" + "p" * 6 + "iiii" + "pp",
    "sparch_recur_bwd_tc_ck": "i" + "p" * 12 + "f" + "p" * 6 + "iiii" + "pip",
    "sparch_spike_post_fwd": "plifppppipp",
    "sparch_spike_post_fwd_bits": "piiifppppippp",
    "sparch_spike_post_bwd": "plifpppp",
    "sparch_spike_post_fwd_bits_bidir": "piiifppppipp" + "ip",
    "sparch_spike_post_bwd_bidir": "piiifpppp",
    "sparch_neuron_params": "pppppiipp",
    "sparch_param_grads": "ppppppiiipp",
    "sparch_small_gemm": "pliplpliiiip",
    "sparch_recur_v0": "pipp",
    "sparch_zero_diag": "pip",
    "sparch_dv_boundary": "ppiiipp",
    "sparch_adam_step": "ippppppp" "p",
    "sparch_readout_fwd": "p" * 7 + "iii" + "p",
    "sparch_readout_bwd": "p" * 6 + "iii" + "p",
    "sparch_ce_fwd": "ppiippp",
    "sparch_ce_bwd": "ppppiipp",
    "sparch_events_to_dense": "ppppiiilppp",
    "sparch_mt19937_uniform": "pilppp",
}
_CT = {"p": _P, "i": _I, "l": _L, "f": _F}

_lock = threading.Lock()
_lib = None
launches = 0  # number of native entry-point calls made by this process (bench.py reads it)


def lib():
    """Load the library once; raise loudly if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m sparch_b200.build` "
                "(needs nvcc; sparch_b200 has no CPU or eager fallback)")
        h = ctypes.CDLL(LIB_PATH)
        h.sparch_last_error.restype = ctypes.c_char_p
        h.sparch_last_error.argtypes = []
        for name, sig in _PROTOS.items():
            fn = getattr(h, name)
            fn.restype = ctypes.c_size_t if name.endswith(("_workspace", "_bytes")) else _I
            fn.argtypes = [_CT[c] for c in sig]
        _lib = h
    return _lib


def ptr(t):
    """Device pointer of a tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def call(name, *args):
    """Call an entry point; raise RuntimeError with the library's message on failure."""
    global launches
    h = lib()
    rc = getattr(h, name)(*args)
    launches += 1
    if rc != 0:
        msg = h.sparch_last_error()
        raise RuntimeError(f"{name} failed ({rc}): {msg.decode() if msg else '?'}")
    return rc
