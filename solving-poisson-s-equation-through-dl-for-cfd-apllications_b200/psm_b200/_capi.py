"""ctypes binding of ``include/psm_b200.h`` (the C-ABI shared library ``libpsm_b200.so``).

The library is built in-tree by ``psm_b200.build`` (nvcc, sm_100a).  There is no Python or
CPU fallback: if the library is missing, or no B200 is visible, the product path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libpsm_b200.so')

PSM_OK, PSM_SKIPPED = 0, 1
PSM_ERR_INVALID, PSM_ERR_CUDA, PSM_ERR_GEOMETRY, PSM_ERR_STATE, PSM_ERR_COMM = -1, -2, -3, -4, -5
PSM_DELTAU_TO_DELTAP, PSM_U_TO_GRADP, PSM_THESIS_U_TO_P = 0, 1, 2
PSM_STD, PSM_MAX_ABS, PSM_MIN_MAX = 0, 1, 2
GEMM_TC_3XTF32, GEMM_TC_TF32, GEMM_FP32_SIMT = 0, 1, 2
(STAGE_GRID, STAGE_XINPUT, STAGE_MLPOUT, STAGE_BLOCKS, STAGE_OFFSETS, STAGE_FIELD, STAGE_SCALARS,
 STAGE_MEANS, STAGE_XU) = range(9)
UNIQUE_ID_BYTES = 128
N_TIMINGS = 12
TIMING_NAMES = ('h2d', 'prep', 'gather', 'extract', 'pca_project', 'mlp', 'pca_inverse', 'strip_means',
                'offsets', 'place', 'back_gather', 'd2h')

c_double_p = C.POINTER(C.c_double)
c_float_p = C.POINTER(C.c_float)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
c_uint8_p = C.POINTER(C.c_uint8)


class PsmConfig(C.Structure):
    _fields_ = [('variant', C.c_int32), ('device', C.c_int32), ('delta', C.c_double), ('shape', C.c_int32),
                ('overlap', C.c_int32), ('input_cols', C.c_int32), ('additive', C.c_int32),
                ('ref_bc', C.c_double), ('skip_threshold', C.c_double), ('near_wall_sdf', C.c_double),
                ('enable_timings', C.c_int32), ('gemm_mode', C.c_int32), ('filter_sigma', C.c_double)]


class PsmParams(C.Structure):
    _fields_ = [('maxs', C.c_double * 5), ('pc_in', C.c_int32), ('pc_p', C.c_int32),
                ('n_out_channels', C.c_int32), ('standardization', C.c_int32),
                ('pca_in_components', c_double_p), ('pca_in_mean', c_double_p),
                ('pca_out_components', c_double_p), ('pca_out_mean', c_double_p),
                ('mean_in', c_double_p), ('std_in', c_double_p), ('mean_out', c_double_p), ('std_out', c_double_p),
                ('max_abs_input_PCA', C.c_double), ('max_abs_output_PCA', C.c_double),
                ('n_dense', C.c_int32), ('reserved', C.c_int32), ('layer_dims', c_int32_p),
                ('dense_kernels', C.POINTER(c_float_p)), ('dense_biases', C.POINTER(c_float_p))]


class PsmTables(C.Structure):
    _fields_ = [('n_cells', C.c_int64), ('grid_h', C.c_int32), ('grid_w', C.c_int32),
                ('vert', c_int32_p), ('weights', c_double_p), ('vert_back', c_int32_p),
                ('weights_back', c_double_p), ('indices', c_int64_p), ('sdfunct', c_double_p)]


class PsmGeometry(C.Structure):
    _fields_ = [('grid_h', C.c_int32), ('grid_w', C.c_int32), ('shape', C.c_int32), ('overlap', C.c_int32),
                ('n_x', C.c_int32), ('n_y', C.c_int32), ('p_i', C.c_int32), ('p_j', C.c_int32),
                ('n_blocks', C.c_int32), ('n_fields', C.c_int32), ('n_cells', C.c_int64),
                ('n_tasks', C.c_int32), ('peer_memory_exchange', C.c_int32),
                ('row0', C.c_int32), ('row1', C.c_int32), ('ext_rows', C.c_int32), ('first_block', C.c_int32),
                ('n_local_blocks', C.c_int32), ('world', C.c_int32),
                ('n_ghost_cells', C.c_int64), ('n_ghost_pix', C.c_int64)]


class PsmMesh(C.Structure):
    _fields_ = [('n_cells', C.c_int64), ('cells_xy', c_double_p), ('xy_stride', C.c_int32), ('back_closed_form', C.c_int32),
                ('top', c_double_p), ('n_top', C.c_int64), ('obst', c_double_p), ('n_obst', C.c_int64), ('probe', c_double_p),
                ('vert', c_int32_p), ('weights', c_double_p), ('vert_back', c_int32_p), ('weights_back', c_double_p),
                ('cache_dir', C.c_char_p)]


class PsmRoute(C.Structure):
    _fields_ = [('n_local', C.c_int64), ('dest_rank', c_int32_p), ('dest_index', c_int32_p)]


class PsmIntegrateGeometry(C.Structure):
    _fields_ = [('min_x', C.c_double), ('max_x', C.c_double), ('min_y', C.c_double), ('max_y', C.c_double),
                ('x0_min', C.c_double), ('center_row', C.c_int32), ('reserved', C.c_int32)]


class PsmShard(C.Structure):
    _fields_ = [('rank', C.c_int32), ('world', C.c_int32), ('grid_h', C.c_int32), ('grid_w', C.c_int32),
                ('row0', C.c_int32), ('row1', C.c_int32), ('ext_rows', C.c_int32), ('send_rows', C.c_int32),
                ('blk_row0', C.c_int32), ('blk_row1', C.c_int32), ('local_ext_rows', C.c_int32),
                ('mask_global', c_uint8_p),
                ('n_owned', C.c_int64), ('n_ghost', C.c_int64), ('n_ghost_pix', C.c_int64),
                ('vert', c_int32_p), ('weights', c_double_p), ('sdfunct', c_double_p),
                ('vert_back', c_int32_p), ('weights_back', c_double_p),
                ('cell_send_ptr', c_int64_p), ('cell_send_idx', c_int32_p), ('cell_recv_ptr', c_int64_p),
                ('pix_send_ptr', c_int64_p), ('pix_send_idx', c_int32_p), ('pix_recv_ptr', c_int64_p),
                ('ghost_pix', c_int64_p)]


# every symbol include/psm_b200.h declares: name -> (restype, argtypes)
SYMBOLS = {
    'psm_api_version': (C.c_int, []),
    'psm_create': (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(PsmConfig)]),
    'psm_load_params': (C.c_int, [C.c_void_p, C.POINTER(PsmParams)]),
    'psm_init_with_tables': (C.c_int, [C.c_void_p, C.POINTER(PsmTables)]),
    'psm_init_sharded': (C.c_int, [C.c_void_p, C.POINTER(PsmShard)]),
    'psm_init_mesh': (C.c_int, [C.c_void_p, C.POINTER(PsmMesh)]),
    'psm_shard_build': (C.c_int, [C.POINTER(PsmTables), c_double_p, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_double, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]),
    'psm_shard_view': (C.POINTER(PsmShard), [C.c_void_p]),
    'psm_shard_cells': (C.c_int, [C.c_void_p, C.POINTER(c_int64_p), C.POINTER(c_int32_p)]),
    'psm_shard_free': (C.c_int, [C.c_void_p]),
    'psm_route_init': (C.c_int, [C.c_void_p, C.POINTER(PsmRoute)]),
    'psm_predict_routed': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'psm_mesh_hash': (C.c_int, [C.c_int32, C.c_double, c_double_p, C.c_int32, C.c_int64, c_double_p, C.c_int64, c_double_p, C.c_int64,
                                c_double_p, C.c_char_p]),
    'psm_mesh_grid': (C.c_int, [C.c_int32, C.c_double, c_double_p, C.c_int32, C.c_int64, c_double_p, c_int32_p, c_int32_p]),
    'psm_back_tables_closed_form': (C.c_int, [c_double_p, C.c_int32, C.c_int64, c_double_p, C.c_int32, c_double_p, C.c_int32, c_int32_p,
                                              c_double_p]),
    'psm_comm_get_unique_id': (C.c_int, [C.c_void_p]),
    'psm_comm_init': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32]),
    'psm_init_from_file': (C.c_int, [C.c_void_p, C.c_char_p]),
    'psm_load_params_file': (C.c_int, [C.c_void_p, C.c_char_p]),
    'psm_save_tables': (C.c_int, [C.POINTER(PsmTables), C.c_char_p]),
    'psm_save_params': (C.c_int, [C.POINTER(PsmParams), C.c_int32, C.c_char_p]),
    'psm_destroy': (C.c_int, [C.c_void_p]),
    'psm_predict': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'psm_predict_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    'psm_predict_fields': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]),
    'psm_predict_fields_device': (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int32]),
    'psm_synchronize': (C.c_int, [C.c_void_p]),
    'psm_get_stream': (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    'psm_register_host_buffer': (C.c_int, [C.c_void_p, C.c_int64]),
    'psm_unregister_host_buffer': (C.c_int, [C.c_void_p]),
    'psm_last_error': (C.c_char_p, [C.c_void_p]),
    'psm_get_geometry': (C.c_int, [C.c_void_p, C.POINTER(PsmGeometry)]),
    'psm_get_plan': (C.c_int, [C.c_void_p, c_int32_p, c_int32_p]),
    'psm_get_owner_map': (C.c_int, [C.c_void_p, c_int32_p]),
    'psm_get_forward_table': (C.c_int, [C.c_void_p, c_int32_p, c_float_p]),
    'psm_get_stage': (C.c_int, [C.c_void_p, C.c_int32, C.c_void_p, C.c_int64]),
    'psm_integrate_gradp': (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    'psm_get_timings': (C.c_int, [C.c_void_p, c_float_p, C.c_int32]),
    'psm_set_timings': (C.c_int, [C.c_void_p, C.c_int32]),
    'psm_get_launch_count': (C.c_int, [C.c_void_p]),
    'psm_get_wait_ns': (C.c_int, [C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32), C.c_int32]),
    'psm_debug_gemm': (C.c_int, [C.c_int32] * 5 + [c_float_p, c_float_p, c_float_p, C.c_int32]),
    'psm_debug_dense_stack': (C.c_int, [C.c_int32] * 4 + [c_int32_p, C.POINTER(c_float_p), C.POINTER(c_float_p), c_float_p, c_float_p,
                                        C.c_int32]),
    'psm_grid_operand_plan': (C.c_int, [C.c_int32, c_int32_p, c_int32_p, C.c_int32, c_int32_p, c_int32_p, c_int32_p, c_int32_p, c_int32_p,
                                        c_int32_p, C.c_int32]),
    'psm_send_map_build': (C.c_int, [C.c_int64, C.c_int32, C.POINTER(C.c_int64), c_int32_p, C.POINTER(C.c_uint32), c_int32_p,
                                     C.POINTER(C.c_int64)]),
    'psm_plan_sizes': (C.c_int, [C.c_int32] * 5 + [c_uint8_p, c_int32_p, c_int32_p, c_int32_p]),
    'psm_plan_compile': (C.c_int, [C.c_int32] * 5 + [c_uint8_p] + [c_int32_p] * 5),
    'psm_plan_shift_lines': (C.c_int, [C.c_int32] * 5 + [c_uint8_p, c_int32_p, c_int32_p]),
}

_lib = None


class PsmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__('psm_b200 error %d: %s' % (code, msg))
        self.code = code


def load():
    """Load ``libpsm_b200.so``; raises if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s not found: build it with `python -m psm_b200.build` "
                          "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(handle, rc):
    if rc < 0:
        msg = load().psm_last_error(handle)
        raise PsmError(rc, (msg or b'').decode('utf-8', 'replace'))
    return rc
