"""ctypes signatures of the tensor-core entry points (include/vd_b200.h, second half)."""
from ctypes import POINTER, c_int, c_int64, c_void_p


def declare(lib):
    from ._lib import TcPlan, EXPORTED
    P = c_void_p
    sig = {
        'vd_tc_plan_make': (c_int, [POINTER(TcPlan), c_int, c_int, c_int]),
        'vd_tc_pack_video': (c_int, [P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_pack_video_u8': (c_int, [P, P, P, POINTER(TcPlan), c_int, P, P, P]),
        'vd_tc_pack_weights': (c_int, [P, P, P, P, P, P, P]),
        'vd_tc_conv_layer': (c_int, [c_int, P, P, P, P, P, c_int, POINTER(TcPlan), P, c_int, c_int, P]),
        'vd_tc_x3_sizes': (c_int, [POINTER(TcPlan), POINTER(c_int64)]),
        'vd_tc_x3_pack_video': (c_int, [P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_x3_pack_video_u8': (c_int, [P, P, P, POINTER(TcPlan), c_int, P, P, P]),
        'vd_tc_x3_pack_weights': (c_int, [P, P, P, P, P, P, P]),
        'vd_tc_x3_conv_layer': (c_int, [c_int, P, P, P, P, P, c_int, POINTER(TcPlan), P, c_int, P]),
        'vd_tc_x3_conv_layer_ex': (c_int, [c_int, P, P, P, P, P, c_int, POINTER(TcPlan), P, c_int, c_int, P]),
        'vd_tc_x3_pack_act': (c_int, [c_int, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_x3_conv_plain': (c_int, [c_int, P, P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_x3_pack_video_hi': (c_int, [P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_x3_pack_video_hi_u8': (c_int, [P, P, P, POINTER(TcPlan), c_int, P, P, P]),
        'vd_tc_pack_weights_bwd': (c_int, [P, P, P, P, P, P, P]),
        'vd_tc_bwd_emb': (c_int, [P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_bwd_gemm': (c_int, [c_int, P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_bwd_col2im': (c_int, [c_int, P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_pack_act': (c_int, [c_int, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_pack_video_ncdhw': (c_int, [P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_pack_weights_part': (c_int, [P, P, P, P, P, P, c_int, P]),
        'vd_tc_pack_dy': (c_int, [c_int, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_bwd_col2im_plain': (c_int, [c_int, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_bwd_gemm_ex': (c_int, [c_int, P, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_wgrad_plan': (c_int, [c_int, POINTER(TcPlan), c_int, POINTER(c_int64)]),
        'vd_tc_wgrad_kt_mode': (c_int, [c_int]),
        'vd_tc_wgrad_pack': (c_int, [c_int, P, P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_wgrad_pack_parts': (c_int, [c_int, P, c_int, P, c_int, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_wgrad_gemm': (c_int, [c_int, P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_wgrad_reduce': (c_int, [c_int, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_dgrad1_sizes': (c_int, [POINTER(TcPlan), POINTER(c_int64)]),
        'vd_tc_pack_dgrad1_weights': (c_int, [P, P, P, POINTER(TcPlan), P]),
        'vd_tc_pack_dyp1': (c_int, [P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_bwd_col2im_ex': (c_int, [c_int, P, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_dgrad1': (c_int, [P, P, P, P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_dgrad1_ex': (c_int, [P, P, P, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_dgrad0_sizes': (c_int, [POINTER(TcPlan), POINTER(c_int64)]),
        'vd_tc_pack_dgrad0_weights': (c_int, [P, P, P]),
        'vd_tc_pack_dyp0': (c_int, [P, P, POINTER(TcPlan), c_int, P]),
        'vd_tc_dgrad0': (c_int, [P, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_pack_dy_part': (c_int, [c_int, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_pack_dyp1_part': (c_int, [P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_pack_dyp0_part': (c_int, [P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_dgrad1_plain': (c_int, [P, P, P, P, POINTER(TcPlan), c_int, c_int, P]),
        'vd_tc_dgrad0_ex': (c_int, [P, P, P, POINTER(TcPlan), c_int, c_int, c_int, P]),
        'vd_tc_debug_params': (c_int, [c_int, POINTER(TcPlan), c_int, POINTER(c_int64), c_int]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    EXPORTED.update(sig)
