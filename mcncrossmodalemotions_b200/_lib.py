"""ctypes binding of libxemo.so (include/xemo.h).  No torch types cross this boundary: plain
pointers and sizes only.  There is no CPU fallback -- if the library or an sm_100 device is
missing, importing the binding or creating a context fails loudly."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libxemo.so")

c_int, c_float, c_size_t, c_void_p, c_int64, c_uint64 = C.c_int, C.c_float, C.c_size_t, C.c_void_p, C.c_int64, C.c_uint64
P = C.POINTER


class XemoArray(C.Structure):
    """xemo_array: single-precision H x W x C x N column-major (MATLAB) array."""

    _fields_ = [("data", c_void_p), ("h", c_int64), ("w", c_int64), ("c", c_int64), ("n", c_int64)]


class XemoError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("xemo error %d: %s" % (code, msg))
        self.code = code


I4 = c_int * 4
I2 = c_int * 2
_AP = P(XemoArray)

# name -> (restype, argtypes); every symbol declared in include/xemo.h
SIGNATURES = {
    "xemo_version": (c_int, []),
    "xemo_current_device": (c_int, [P(c_int)]),
    "xemo_create": (c_int, [c_int, c_void_p, P(c_void_p)]),
    "xemo_destroy": (None, [c_void_p]),
    "xemo_last_error": (C.c_char_p, [c_void_p]),
    "xemo_sync": (c_int, [c_void_p]),
    "xemo_trim": (c_int, [c_void_p]),
    "xemo_num_sms": (c_int, [c_void_p]),
    "xemo_launch_count": (c_uint64, [c_void_p]),
    "xemo_h2d": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "xemo_d2h": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t]),
    "xemo_memset": (c_int, [c_void_p, c_void_p, c_int, c_size_t]),
    "xemo_set_stream": (c_int, [c_void_p, c_void_p]),
    "xemo_stream_wait": (c_int, [c_void_p, c_void_p, c_void_p]),
    "xemo_capture_begin": (c_int, [c_void_p]),
    "xemo_capture_end": (c_int, [c_void_p, P(c_void_p)]),
    "xemo_graph_launch": (c_int, [c_void_p, c_void_p]),
    "xemo_graph_num_kernels": (c_int, [c_void_p]),
    "xemo_graph_destroy": (None, [c_void_p]),
    "xemo_out_size": (c_int, [c_int64, c_int64, c_int, c_int, P(c_int), P(c_int), P(c_int64), P(c_int64)]),
    "xemo_vl_nnconv": (c_int, [c_void_p, _AP, _AP, _AP, _AP, P(c_int), P(c_int), _AP, _AP, _AP, _AP]),
    "xemo_vl_nnpool": (c_int, [c_void_p, _AP, P(c_int), _AP, P(c_int), P(c_int), c_int, _AP, c_void_p]),
    "xemo_vl_nnbnorm": (c_int, [c_void_p, _AP, c_void_p, c_void_p, _AP, c_float, c_void_p, _AP, c_void_p, c_void_p, c_void_p]),
    "xemo_vl_nnrelu": (c_int, [c_void_p, _AP, _AP, c_float, _AP]),
    "xemo_vl_nnsigmoid": (c_int, [c_void_p, _AP, _AP, _AP]),
    "xemo_vl_nnsoftmaxt": (c_int, [c_void_p, _AP, c_float, _AP]),
    "xemo_vl_nnsoftmaxceloss": (c_int, [c_void_p, _AP, _AP, c_void_p, c_float, c_int, c_void_p, c_void_p, _AP]),
    "xemo_vl_nnloss_classerror": (c_int, [c_void_p, _AP, c_void_p, c_void_p]),
    "xemo_vl_nnglobalpool": (c_int, [c_void_p, _AP, _AP, _AP]),
    "xemo_vl_nnaxpy": (c_int, [c_void_p, _AP, _AP, _AP, _AP]),
    "xemo_op_hwcn_to_nhwc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int]),
    "xemo_op_nhwc_to_hwcn": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_filters_to_krsc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int]),
    "xemo_op_face_rows_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_face_u8_rows_im2col": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_spec_s2d": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_spectrogram": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_float, c_int, c_void_p]),
    "xemo_op_spec_rownorm": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int]),
    "xemo_op_conv_fwd": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                 c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int]),
    "xemo_dgrad_pack_elems": (c_size_t, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "xemo_op_pack_dgrad_filters": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_conv_dgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_pack_dgrad_filters_fullheight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "xemo_op_conv_dgrad_fullheight": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]),
    "xemo_op_conv_wgrad": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int,
                                   c_int, c_int, c_int, c_int, c_int, c_void_p, c_float]),
    "xemo_debug_conv_plan": (c_int, [c_int] * 14 + [P(c_int)]),
    "xemo_debug_set_conv_pair_mode": (c_int, [c_int]),
    "xemo_debug_conv_plan2": (c_int, [c_int] * 14 + [P(c_int)]),
    "xemo_debug_wgrad_plan": (c_int, [c_int] * 15 + [P(c_int)]),
    "xemo_debug_se_gate_plan": (c_int, [c_int] * 6 + [P(c_int)]),
    "xemo_debug_fixed_channel_grid": (c_int, [C.c_longlong, c_int, c_int, c_int, c_int]),
    "xemo_op_colsum": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_int, c_float, c_void_p]),
    "xemo_op_maxpool_fwd": (c_int, [c_void_p, c_void_p] + [c_int] * 12 + [c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_maxpool_fwd_win": (c_int, [c_void_p, c_void_p] + [c_int] * 12 + [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "xemo_op_maxpool_bwd_ld": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 12 + [c_void_p, c_int]),
    "xemo_stem_ws_doubles": (c_size_t, []),
    "xemo_op_stem_pair_filter": (c_int, [c_void_p, c_void_p, c_int, c_void_p]),
    "xemo_op_tile_f32": (c_int, [c_void_p, c_void_p, c_int, c_int, c_float, c_void_p]),
    "xemo_op_stem_autocorr": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_stem_bn_train": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_float,
                                      c_void_p, c_void_p, c_void_p]),
    "xemo_op_stem_pool_bn_reduce": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_stem_wgrad_finalize": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p,
                                            c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_maxpool_bwd": (c_int, [c_void_p, c_void_p, c_void_p] + [c_int] * 12 + [c_void_p]),
    "xemo_op_avgpool_fwd": (c_int, [c_void_p, c_void_p] + [c_int] * 12 + [c_void_p]),
    "xemo_op_avgpool_bwd": (c_int, [c_void_p, c_void_p] + [c_int] * 12 + [c_void_p]),
    "xemo_op_bn_train": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_bn_test": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_affine_act": (c_int, [c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_int, c_void_p]),
    "xemo_op_bn_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_void_p,
                               c_void_p, c_void_p, c_void_p, c_void_p, c_float]),
    "xemo_op_bn_bwd_pool": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p] + [c_int] * 12 + [c_void_p] * 8 + [c_float]),
    "xemo_op_relu_bwd": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xemo_op_add_act": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "xemo_op_se_squeeze": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "xemo_op_se_gate": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_se_gate_lin": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int] + [c_void_p] * 9),
    "xemo_op_conv_fwd_nc": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p] + [c_int] * 9 + [c_void_p, c_void_p, c_void_p,
                                                                                                          c_int, c_void_p]),
    "xemo_op_se_excite": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "xemo_op_logit_aggregate": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "xemo_op_softmaxce": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_float, c_int, c_float,
                                  c_float, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xemo_op_loss": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_float, c_int, c_float,
                             c_float, c_void_p, c_int, c_int, c_void_p, c_void_p, c_void_p]),
    "xemo_set_deterministic": (c_int, [c_void_p, c_int]),
    "xemo_get_deterministic": (c_int, [c_void_p]),
    "xemo_set_conv_precision": (c_int, [c_void_p, c_int]),
    "xemo_get_conv_precision": (c_int, [c_void_p]),
    "xemo_op_sgd_momentum": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_float, c_float, c_float, c_void_p]),
    "xemo_op_grad_guard": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "xemo_op_sgd_momentum_guarded": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_float, c_float, c_float,
                                             c_void_p, c_void_p]),
    "xemo_op_moments_average_guarded": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_void_p]),
    "xemo_op_moments_average": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_float]),
    "xemo_op_cast_f32_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "xemo_op_cast_f16_f32": (c_int, [c_void_p, c_void_p, c_size_t, c_float, c_void_p]),
    "xemo_op_fill_strided_f32": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_size_t, c_int, c_float]),
    # (C) graph-level entry points
    "xemo_net_create": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_int, P(c_void_p)]),
    "xemo_net_destroy": (None, [c_void_p]),
    "xemo_net_batch": (c_int, [c_void_p]),
    "xemo_net_num_params": (c_int, [c_void_p]),
    "xemo_net_param_name": (C.c_char_p, [c_void_p, c_int]),
    "xemo_net_param_dims": (c_int, [c_void_p, C.c_char_p, P(c_int64)]),
    "xemo_net_set_param": (c_int, [c_void_p, C.c_char_p, c_void_p, c_size_t]),
    "xemo_net_finalize": (c_int, [c_void_p]),
    "xemo_net_get_tensor": (c_int, [c_void_p, c_int, C.c_char_p, c_void_p, c_size_t]),
    "xemo_net_set_momentum": (c_int, [c_void_p, C.c_char_p, c_void_p, c_size_t]),
    "xemo_net_input_bytes": (c_int, [c_void_p, P(c_size_t)]),
    "xemo_net_set_input": (c_int, [c_void_p, c_void_p, c_size_t]),
    "xemo_net_set_target": (c_int, [c_void_p, c_void_p, c_void_p]),
    "xemo_net_set_loss": (c_int, [c_void_p, c_int, c_float, c_float]),
    "xemo_net_set_hyper": (c_int, [c_void_p, c_float, c_float, c_float, c_int]),
    "xemo_net_buffer": (c_void_p, [c_void_p, C.c_char_p]),
    "xemo_net_grad_elems": (c_size_t, [c_void_p]),
    "xemo_net_num_kernels": (c_int, [c_void_p]),
    "xemo_teacher_forward": (c_int, [c_void_p, c_void_p]),
    "xemo_student_forward": (c_int, [c_void_p, c_int, c_void_p]),
    "xemo_student_train_step": (c_int, [c_void_p, c_void_p]),
    "xemo_sgd_step": (c_int, [c_void_p, c_float, c_float, c_float, c_int]),
    "xemo_allreduce_grads": (c_int, [c_void_p, c_void_p]),
    "xemo_distill_set_windows": (c_int, [c_void_p, c_void_p, c_void_p]),
    "xemo_distill_couple": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int]),
    "xemo_distill_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_float, c_int]),
    "xemo_net_set_overlap": (c_int, [c_void_p, c_int]),
    "xemo_net_reset_metrics": (c_int, [c_void_p]),
    "xemo_net_metrics": (c_int, [c_void_p, c_void_p, c_int]),
    "xemo_comm_unique_id": (c_int, [c_void_p]),
    "xemo_comm_create": (c_int, [c_void_p, c_void_p, c_int, c_int, P(c_void_p)]),
    "xemo_comm_destroy": (None, [c_void_p]),
    "xemo_comm_allreduce_f32": (c_int, [c_void_p, c_void_p, c_size_t]),
}

_lib = None


def load_library():
    """dlopen libxemo.so (built in-tree by mcncrossmodalemotions_b200.build) and type every symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -m mcncrossmodalemotions_b200.build` "
            "(there is no CPU fallback)" % LIB_PATH
        )
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class Context:
    """xemo_ctx bound to one device and one stream.  Every xemo_* entry point is reachable as a
    method without the `xemo_` prefix and with the context argument supplied; non-zero status codes
    raise XemoError carrying xemo_last_error()."""

    def __init__(self, device=0, stream=None):
        self.lib = load_library()
        h = c_void_p()
        rc = self.lib.xemo_create(int(device), c_void_p(stream or 0), C.byref(h))
        if rc != 0:
            raise XemoError(rc, "xemo_create failed (no sm_100 device, or TMA driver entry points unavailable); no CPU fallback exists")
        self.handle = h
        self.device = device
        self.profiler = None  # optional object with before(name, args) / after(name, args) hooks (bench.py)
        self.num_sms = self.lib.xemo_num_sms(h)
        self.children = []    # weak references to objects living in this context (networks, communicators): closed first

    def adopt(self, obj):
        import weakref

        self.children.append(weakref.ref(obj))

    def close(self):
        if getattr(self, "handle", None):
            # networks before communicators (ncclCommDestroy waits for the graphs that captured it), both before the context
            kids = [r() for r in getattr(self, "children", [])]
            for k in sorted((k for k in kids if k is not None), key=lambda k: k.__class__.__name__ == "Comm"):
                k.close()
            self.lib.xemo_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def call(self, name, *args):
        fn = getattr(self.lib, "xemo_" + name)
        prof = self.profiler
        if prof is not None:
            prof.before(name, args)
        rc = fn(self.handle, *args)
        if prof is not None:
            prof.after(name, args)
        if rc != 0:
            raise XemoError(rc, self.lib.xemo_last_error(self.handle).decode())

    def __getattr__(self, name):
        if name.startswith(("op_", "vl_")) or name in ("sync", "trim", "h2d", "d2h", "memset", "capture_begin", "graph_launch", "set_stream", "stream_wait",
                                                     "set_conv_precision", "set_deterministic"):
            return lambda *a: self.call(name, *a)
        raise AttributeError(name)

    def launch_count(self):
        return int(self.lib.xemo_launch_count(self.handle))

    def capture_end(self):
        g = c_void_p()
        self.call("capture_end", C.byref(g))
        return Graph(self, g)


class Graph:
    def __init__(self, ctx, handle):
        self.ctx, self.handle = ctx, handle
        self.num_kernels = ctx.lib.xemo_graph_num_kernels(handle)

    def launch(self):
        self.ctx.call("graph_launch", self.handle)

    def destroy(self):
        if self.handle:
            self.ctx.lib.xemo_graph_destroy(self.handle)
            self.handle = None
