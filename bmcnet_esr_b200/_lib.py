"""ctypes binding of libbmc_b200.so (include/bmc_b200.h).  No fallback: if the library is not
built, importing a symbol raises with the build command."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# BMC_B200_LIB selects another build (e.g. the bf16 variant made by `build.py --bf16`)
LIB_PATH = os.environ.get('BMC_B200_LIB') or os.path.join(_HERE, 'libbmc_b200.so')

ENC_FLIP_Y, ENC_MUTATE, ENC_NO_QUIRKS, ENC_TNORM, ENC_BILINEAR, ENC_SKIP_ZERO_ENDS, ENC_DETERMINISTIC = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20, 0x40
ENC_SPLIT_BINS = 0x80
MODEL_BMCNET, MODEL_BMCNET_PLAIN = 0, 1

_vp, _i, _i64, _f, _sz, _u = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_size_t, C.c_uint


class GemmJob(C.Structure):
    """bmc_gemm_job_t"""
    _fields_ = [('n_seg', _i), ('a', _vp * 3), ('a_rows', _i * 3), ('a_ch', _i * 3), ('a_row_base', _i * 3),
                ('w', _vp), ('w_rows', _i), ('w_k', _i), ('w_row_base', _i), ('w_img_stride', _i),
                ('bias', _vp), ('residual', _vp), ('res_row_base', _i), ('out_act16', _vp),
                ('out_row_base', _i), ('out_f32', _vp), ('relu', _i),
                ('ln_gamma', _vp), ('ln_beta', _vp), ('ln_eps', _f)]


# name -> (restype, argtypes); every symbol declared in include/bmc_b200.h
SIGNATURES = {
    'bmc_abi_version': (_i, []),
    'bmc_last_error': (C.c_char_p, []),
    'bmc_act_dtype': (C.c_char_p, []),
    'bmc_encode_workspace_bytes': (_sz, [_i64]),
    'bmc_encode_channels': (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _sz, _u, _vp]),
    'bmc_encode_channels_windows': (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _u, _vp]),
    'bmc_encode_channels_windows_raw': (_i, [_vp, _vp, _vp, _i64, _i64, _i64, _i, _i, _i, _vp, _u, _vp]),
    'bmc_format_events': (_i, [_vp, _vp, _vp, _vp, _i64, _vp, _vp]),
    'bmc_encode_image': (_i, [_vp, _vp, _vp, _i64, _i, _i, _vp, _vp, _sz, _u, _vp]),
    'bmc_encode_voxel': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _vp, _vp, _sz, _u, _vp]),
    'bmc_encode_stack_flag_offset': (_sz, [_i64]),
    'bmc_encode_stack': (_i, [_vp, _vp, _vp, _vp, _i64, _i, _i, _i, _i, _vp, _vp, _sz, _u, _vp]),
    'bmc_encode_stack_shard': (_i, [_vp, _vp, _vp, _i64, _vp, _i64, _i64, _i, _i, _i, _i, _vp, _vp, _sz, _u, _vp]),
    'bmc_model_create': (_vp, [_i, _i, _i, _i, _i]),
    'bmc_model_destroy': (None, [_vp]),
    'bmc_model_weight_bytes': (_sz, [_vp]),
    'bmc_model_load_state_dict': (_i, [_vp, C.POINTER(C.c_char_p), C.POINTER(_vp), C.POINTER(_i64), _i,
                                       _vp, _sz, _vp]),
    'bmc_model_configure': (_i, [_vp, _i, _i, _i]),
    'bmc_model_workspace_bytes': (_sz, [_vp]),
    'bmc_model_bind_workspace': (_i, [_vp, _vp, _sz]),
    'bmc_model_forward': (_i, [_vp, _vp, C.POINTER(_i64), _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    'bmc_model_step': (_i, [_vp, _vp, C.POINTER(_i64), _i, _vp, _vp]),
    'bmc_model_launches_per_step': (_i, [_vp]),
    'bmc_model_set_debug_simt': (_i, [_vp, _i]),
    'bmc_conv_gemm': (_i, [C.POINTER(GemmJob), _i, _i, _i, _i, _i, _i, _i, _vp]),
    'bmc_attention_weights': (_i, [_vp, _vp, _i, _i, _i, _f, _vp, _i, _vp, _i, _vp]),
    'bmc_layernorm_rows': (_i, [_vp, _vp, _vp, _f, _i64, _vp, _vp]),
    'bmc_pack_nchw': (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp]),
    'bmc_unpack_nchw': (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    'bmc_stack_to_events_workspace_bytes': (_sz, [_i, _i64, _i64]),
    'bmc_stack_event_counts': (_i, [_vp, _i, _i64, _vp, _sz, _vp, _vp, _vp]),
    'bmc_stack_to_events': (_i, [_vp, _i, _i, _i, _i, _i, _i64, _vp, _vp, _vp, _sz, _vp]),
    'bmc_stack2cnt': (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    'bmc_sr_metrics': (_i, [_vp, _i, _i, _i, _i, _vp, _i, _i, _vp, _i, _i, _vp, _vp]),
    'bmc_conv_wgrad_workspace_bytes': (_sz, [_i, _i, _i]),
    'bmc_conv_wgrad': (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _f, _vp, _vp, _vp, _sz, _i, _vp]),
    'bmc_relu_backward': (_i, [_vp, _vp, _i64, _vp, _vp]),
    'bmc_layernorm_rows_backward_workspace_bytes': (_sz, []),
    'bmc_layernorm_rows_backward': (_i, [_vp, _vp, _vp, _f, _i64, _vp, _f, _vp, _vp, _vp, _sz, _vp]),
    'bmc_adam_amsgrad_step': (_i, [_vp, _vp, _vp, _vp, _vp, _i64, _i, _f, _f, _f, _f, _f, _vp]),
}

_lib = None


class BmcError(RuntimeError):
    pass


def lib():
    """The loaded shared library (loaded once).  Raises if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BmcError('%s is missing: build it with `python -m bmcnet_esr_b200.build` '
                           '(there is no CPU / PyTorch fallback)' % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the header and the .so disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def act_dtype():
    """torch dtype of the 16-bit activation / weight tensors of the loaded build."""
    import torch
    return {'f16': torch.float16, 'bf16': torch.bfloat16}[lib().bmc_act_dtype().decode()]


def check(rc):
    if rc != 0:
        raise BmcError('libbmc_b200 error %d: %s' % (rc, lib().bmc_last_error().decode()))


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
