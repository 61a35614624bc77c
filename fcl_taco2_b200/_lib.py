"""ctypes binding of the C ABI in include/fcl_taco2.h (libfcl_taco2.so).

This is the thin custom-op layer: PyTorch only supplies device memory
(`tensor.data_ptr()`) and the current stream. There is no CPU fallback -- if the
shared library is missing, loading raises with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# FCL_TACO2_LIB selects another build of the same library (tools: the -DFCL_DEC_PROF profiling build)
LIB_PATH = os.environ.get("FCL_TACO2_LIB") or os.path.join(_HERE, "lib", "libfcl_taco2.so")
ABI_VERSION = 29

i32, i64, u64, f32 = C.c_int32, C.c_int64, C.c_uint64, C.c_float
ptr = C.c_void_p


class LenRegParams(C.Structure):
    _fields_ = [("n_rows", i32), ("n_utts", i32), ("dur", ptr), ("utt_off", ptr), ("frame_off", ptr),
                ("utt_frame_off", ptr), ("order", ptr), ("totals", ptr), ("ws", ptr)]


class FrameMapParams(C.Structure):
    _fields_ = [("n_rows", i32), ("n_utts", i32), ("n_frames", i32), ("frame_off", ptr), ("utt_frame_off", ptr),
                ("frame_row", ptr), ("frame_step", ptr), ("frame_seg_lo", ptr), ("frame_seg_hi", ptr),
                ("position", ptr)]


class ConvGemmParams(C.Structure):
    _fields_ = [("rows", i32), ("cin", i32), ("cout", i32), ("taps", i32), ("a", ptr), ("lda", i32),
                ("gather", ptr), ("seg_lo", ptr), ("seg_hi", ptr), ("w", ptr), ("bias", ptr), ("residual", ptr),
                ("ldr", i32), ("out", ptr), ("ldo", i32), ("act", i32)]


class ConvGemmBf16Params(C.Structure):
    _fields_ = [("rows", i32), ("cin", i32), ("cout", i32), ("taps", i32), ("a", ptr), ("lda", i32),
                ("gather", ptr), ("row_gather", ptr), ("tile_src", ptr), ("tile_dst", ptr), ("n_tiles_dev", ptr),
                ("n_tiles", i32), ("map_halo", i32), ("w_packed", ptr), ("ntile", i32), ("kstage", i32),
                ("bias", ptr), ("residual", ptr), ("ldr", i32), ("out", ptr), ("ldo", i32), ("act", i32),
                ("out_bf16", i32)]


class ConvTilesParams(C.Structure):
    _fields_ = [("n_segs", i32), ("max_tiles", i32), ("halo", i32), ("seg_off", ptr), ("seg_first_tile", ptr),
                ("tile_src", ptr), ("tile_dst", ptr), ("n_tiles", ptr)]


MAX_STACK_LAYERS = 5


class ConvLayer(C.Structure):
    _fields_ = [("cin", i32), ("cout", i32), ("kstage", i32), ("act", i32), ("w_packed", ptr), ("bias", ptr)]


class ConvStackTilesParams(C.Structure):
    _fields_ = [("n_segs", i32), ("max_tiles", i32), ("stride", i32), ("seg_off", ptr), ("tiles", ptr), ("n_tiles", ptr)]


class ConvStackParams(C.Structure):
    _fields_ = [("n_layers", i32), ("taps", i32), ("layers", ConvLayer * MAX_STACK_LAYERS), ("in_", ptr), ("ld_in", i32),
                ("in_channels", i32), ("b_stages", i32), ("gather", ptr), ("tiles", ptr), ("n_tiles_dev", ptr), ("n_tiles", i32), ("residual", ptr),
                ("ldr", i32), ("out", ptr), ("ldo", i32)]


class LayerNormParams(C.Structure):
    _fields_ = [("rows", i32), ("chans", i32), ("x", ptr), ("gamma", ptr), ("beta", ptr), ("y", ptr),
                ("head_w", ptr), ("head_b", f32), ("head_out", ptr), ("dur_out", ptr)]


class EmbedAddParams(C.Structure):
    _fields_ = [("rows", i32), ("chans", i32), ("taps", i32), ("h", ptr), ("pitch", ptr), ("energy", ptr),
                ("seg_lo", ptr), ("seg_hi", ptr), ("wp", ptr), ("bp", ptr), ("we", ptr), ("be", ptr), ("hn", ptr),
                ("order", ptr), ("img", ptr)]


class BiLstmParams(C.Structure):
    _fields_ = [("n_utts", i32), ("hidden", i32), ("utt_off", ptr), ("gx", ptr), ("whh", ptr), ("out", ptr),
                ("group", i32)]


class DecoderParams(C.Structure):
    _fields_ = [("n_rows", i32), ("eunits", i32), ("dunits", i32), ("prenet_units", i32), ("odim", i32),
                ("order", ptr), ("dur", ptr), ("frame_off", ptr), ("row_utt", ptr), ("row_phone", ptr),
                ("g0h", ptr), ("y0h", ptr), ("wp0", ptr), ("bp0", ptr), ("wp1", ptr), ("bp1", ptr),
                ("w0", ptr), ("wpos", ptr), ("w1", ptr), ("b1", ptr), ("wf", ptr), ("cstate", ptr), ("before", ptr),
                ("zoneout", f32), ("dropout_p", f32), ("dropout_seed", u64), ("tile_rows", i32), ("tf_y", ptr)]


class DecoderBf16Params(C.Structure):
    _fields_ = [("n_rows", i32), ("n_tiles", i32), ("n_slots", i32), ("eunits", i32), ("dunits", i32),
                ("prenet_units", i32), ("odim", i32), ("order", ptr), ("dur", ptr), ("frame_off", ptr),
                ("row_utt", ptr), ("row_phone", ptr), ("hn_img", ptr), ("w_stream", ptr), ("bp0", ptr), ("bp1", ptr),
                ("wpos", ptr), ("b0", ptr), ("b1", ptr), ("group", i32), ("act_priv", ptr), ("act_shared", ptr),
                ("c_ws", ptr), ("group_sync", ptr), ("before", ptr),
                ("zoneout", f32), ("dropout_p", f32), ("dropout_seed", u64), ("tile_slot", ptr), ("tile_rank", ptr),
                ("trace", ptr), ("trace_cap", i32), ("inflight", i32), ("tf_x1", ptr)]


class BiLstmBf16Params(C.Structure):
    _fields_ = [("n_utts", i32), ("hidden", i32), ("tile_utts", i32), ("utt_off", ptr), ("gx", ptr),
                ("whh_packed", ptr), ("c_ws", ptr), ("out", ptr), ("gx_blk", ptr), ("prow_off", ptr), ("gx_rows", i32),
                ("gx_blk_half", i32)]


class Prenet0TfParams(C.Structure):
    _fields_ = [("n_frames", i32), ("odim", i32), ("prenet_units", i32), ("y", ptr), ("frame_row", ptr), ("frame_step", ptr),
                ("row_utt", ptr), ("row_phone", ptr), ("wp0", ptr), ("bp0", ptr), ("dropout_p", f32), ("dropout_seed", u64),
                ("x1", ptr)]


class PadRowsParams(C.Structure):
    _fields_ = [("n_utts", i32), ("n_rows", i32), ("gap", i32), ("utt_off", ptr), ("n_tiles", i32), ("prow_src", ptr),
                ("prow_off", ptr)]


class RowsToImageParams(C.Structure):
    _fields_ = [("n_tiles", i32), ("chans", i32), ("src", ptr), ("ld", i32), ("gather", ptr), ("prow_src", ptr), ("img", ptr),
                ("src_chans", i32)]


class ConvImgParams(C.Structure):
    _fields_ = [("n_tiles", i32), ("cin", i32), ("cout", i32), ("taps", i32), ("nb", i32), ("act", i32), ("epi", i32),
                ("in_img", ptr), ("w_packed", ptr), ("bias", ptr), ("prow_src", ptr), ("out_img", ptr), ("out_blk", ptr),
                ("gamma", ptr), ("beta", ptr), ("head_w", ptr), ("head_b", f32), ("head_out", ptr), ("dur_out", ptr),
                ("n_pairs", i32), ("trace", ptr), ("trace_cap", i32), ("out_rows", ptr), ("ldo", i32), ("out_chans", i32),
                ("residual", ptr), ("ldr", i32)]


class DecoderScheduleParams(C.Structure):
    _fields_ = [("n_rows", i32), ("n_tiles", i32), ("n_slots", i32), ("unit_rows", i32), ("order", ptr), ("dur", ptr),
                ("tile_slot", ptr), ("tile_rank", ptr)]


class PackRowsParams(C.Structure):
    _fields_ = [("n_rows", i32), ("cols", i32), ("src", ptr), ("ld", i32), ("order", ptr), ("dst", ptr)]


STRUCTS = [LenRegParams, FrameMapParams, ConvGemmParams, LayerNormParams, EmbedAddParams, BiLstmParams,
           DecoderParams, ConvGemmBf16Params, DecoderBf16Params, PackRowsParams, BiLstmBf16Params, ConvTilesParams, DecoderScheduleParams, ConvStackTilesParams,
           ConvStackParams, PadRowsParams, RowsToImageParams, ConvImgParams, Prenet0TfParams]

ENTRY_POINTS = {
    "fcl_len_reg_scan": LenRegParams,
    "fcl_len_reg_frame_map": FrameMapParams,
    "fcl_conv_gemm_f32": ConvGemmParams,
    "fcl_layernorm_f32": LayerNormParams,
    "fcl_embed_add_f32": EmbedAddParams,
    "fcl_bilstm_f32": BiLstmParams,
    "fcl_decoder_f32": DecoderParams,
    "fcl_conv_gemm_bf16": ConvGemmBf16Params,
    "fcl_decoder_bf16": DecoderBf16Params,
    "fcl_pack_rows_bf16": PackRowsParams,
    "fcl_bilstm_bf16": BiLstmBf16Params,
    "fcl_conv_tiles": ConvTilesParams,
    "fcl_decoder_schedule": DecoderScheduleParams,
    "fcl_decoder_bf16_pair": DecoderBf16Params,
    "fcl_decoder_bf16_pair_v1": DecoderBf16Params,
    "fcl_conv_stack_tiles": ConvStackTilesParams,
    "fcl_conv_stack_bf16": ConvStackParams,
    "fcl_pad_rows": PadRowsParams,
    "fcl_rows_to_image": RowsToImageParams,
    "fcl_conv_img_bf16": ConvImgParams,
    "fcl_prenet0_tf": Prenet0TfParams,
}
PLAIN_SYMBOLS = ["fcl_abi_version", "fcl_operand_format", "fcl_last_error", "fcl_sm_count", "fcl_struct_size",
                 "fcl_decoder_bf16_workspace", "fcl_len_reg_ws_ints", "fcl_peer_alloc", "fcl_peer_free", "fcl_ipc_export", "fcl_ipc_open",
                 "fcl_ipc_close", "fcl_copy_async", "fcl_wait_flags", "fcl_write_flags"]

ACT_NONE, ACT_RELU, ACT_TANH = 0, 1, 2
EPI_IMAGE, EPI_LN_IMAGE, EPI_LN_HEAD, EPI_BLOCKED_F32, EPI_BLOCKED_F16, EPI_ROWS_F32 = 0, 1, 2, 3, 4, 5
PAD_GAP = 2            # zero rows between utterances in the padded row space (halo of the k <= 5 convolutions)
MAX_DURATION = 1023


class FclError(RuntimeError):
    pass


_lib = None


def load():
    """Load libfcl_taco2.so once; verify ABI version and struct layouts."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FclError(
            f"{LIB_PATH} not found: the B200 path has no CPU fallback. Build it with "
            "`python -c 'import __graft_entry__ as g; g.build()'` (needs nvcc, sm_100a).")
    lib = C.CDLL(LIB_PATH)
    lib.fcl_abi_version.restype = C.c_int
    lib.fcl_last_error.restype = C.c_char_p
    lib.fcl_sm_count.restype = C.c_int
    lib.fcl_operand_format.restype = C.c_int
    lib.fcl_struct_size.restype = C.c_int
    lib.fcl_struct_size.argtypes = [C.c_int]
    if lib.fcl_abi_version() != ABI_VERSION:
        raise FclError(f"ABI mismatch: library {lib.fcl_abi_version()} != binding {ABI_VERSION}; rebuild")
    for i, st in enumerate(STRUCTS):
        if lib.fcl_struct_size(i) != C.sizeof(st):
            raise FclError(f"struct layout mismatch for {st.__name__}: C {lib.fcl_struct_size(i)} != ctypes {C.sizeof(st)}")
    for name, st in ENTRY_POINTS.items():
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = [C.POINTER(st), C.c_void_p]
    lib.fcl_len_reg_ws_ints.restype = C.c_int
    lib.fcl_len_reg_ws_ints.argtypes = [C.c_int32]
    lib.fcl_decoder_bf16_workspace.restype = C.c_int
    lib.fcl_decoder_bf16_workspace.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                               C.POINTER(C.c_int64)]
    for name, argtypes in (("fcl_peer_alloc", [C.c_int64, C.POINTER(C.c_void_p)]), ("fcl_peer_free", [C.c_void_p]),
                           ("fcl_ipc_export", [C.c_void_p, C.c_char_p]), ("fcl_ipc_open", [C.c_char_p, C.POINTER(C.c_void_p)]),
                           ("fcl_ipc_close", [C.c_void_p]),
                           ("fcl_copy_async", [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
                           ("fcl_wait_flags", [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]),
                           ("fcl_write_flags", [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p])):
        fn = getattr(lib, name)
        fn.restype = C.c_int
        fn.argtypes = argtypes
    _lib = lib
    return lib


def plain(name: str, *args):
    """Invoke one of the non-struct entry points (peer-memory plumbing); raise FclError on failure."""
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise FclError(f"{name} failed ({rc}): {lib.fcl_last_error().decode()}")


def call(name: str, params, stream: int):
    """Invoke an entry point; raise FclError with the library's message on failure."""
    lib = load()
    rc = getattr(lib, name)(C.byref(params), C.c_void_p(stream))
    if rc != 0:
        raise FclError(f"{name} failed ({rc}): {lib.fcl_last_error().decode()}")


def operand_format() -> str:
    """'fp16' or 'bf16': element format of the 16-bit tensor-core operands this build of the library uses."""
    return "bf16" if load().fcl_operand_format() == 1 else "fp16"


def dptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    return None if t is None else t.data_ptr()


def decoder_bf16_workspace(prenet_units: int, dunits: int):
    """-> (private scratch bytes per CTA, shared scratch bytes per group, cell-state floats per CTA)."""
    a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
    rc = load().fcl_decoder_bf16_workspace(prenet_units, dunits, C.byref(a), C.byref(b), C.byref(c))
    if rc != 0:
        raise FclError("fcl_decoder_bf16_workspace failed")
    return a.value, b.value, c.value
