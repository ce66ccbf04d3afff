"""ctypes binding of libvmadpm.so (declared in include/vmadpm.h).

There is no CPU fallback: importing this module without the built library, or creating a
context without a CUDA device, raises.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, byref, c_char_p, c_double, c_int, c_int32, c_int64, c_uint64,
                    c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, 'libvmadpm.so')


class VpmError(RuntimeError):
    pass


class vpm_mesh(Structure):
    _fields_ = [
        ('n', c_int64 * 3),
        ('scale', c_double * 3),
        ('start0', c_int64),
        ('nloc0', c_int64),
        ('stride0', c_int64),
        ('stride1', c_int64),
        ('ghost', c_int32),
        ('reserved', c_int32),
    ]


def _load():
    if not os.path.exists(LIB_PATH):
        raise VpmError(
            "libvmadpm.so is not built (%s). Run `python -m vmad_b200.build` (needs nvcc). "
            "vmad_b200 has no CPU path." % LIB_PATH)
    return ctypes.CDLL(LIB_PATH)


lib = _load()

_P = c_void_p  # device pointers travel as integers
_I3 = c_int64 * 3

# name -> (restype, argtypes); every name here must be declared in include/vmadpm.h
class vpm_peer_sync(Structure):
    _fields_ = [('peer_flags', POINTER(c_uint64)), ('my_flags', c_void_p), ('rank', c_int32),
                ('mode', c_int32), ('buf', c_int32), ('reserved', c_int32)]


PEER_RAISE_ACK, PEER_WAIT_ACK, PEER_SIGNAL, PEER_WAIT_DATA = 1, 2, 4, 8
PEER_FLAG_WORDS = 128
ROW_PADDING = -2147483648


SIGNATURES = {
    'vpm_ctx_create': (c_int, [POINTER(c_void_p), c_int, c_void_p]),
    'vpm_ctx_destroy': (c_int, [c_void_p]),
    'vpm_ctx_set_stream': (c_int, [c_void_p, c_void_p]),
    'vpm_ctx_push_stream': (c_int, [c_void_p, c_void_p]),
    'vpm_ctx_pop_stream': (c_int, [c_void_p]),
    'vpm_sync': (c_int, [c_void_p]),
    'vpm_last_error': (c_char_p, []),
    'vpm_version': (c_int, []),
    'vpm_launch_count': (c_int64, []),
    'vpm_device_sm_count': (c_int, [c_void_p, POINTER(c_int)]),
    'vpm_profile_start': (c_int, []),
    'vpm_profile_stop': (c_int, []),
    'vpm_profile_read': (c_int64, [ctypes.c_char_p, c_int64]),
    'vpm_cell_keys': (c_int, [c_void_p, POINTER(vpm_mesh), _P, c_int64, _P]),
    'vpm_layout_ncell': (c_int64, [POINTER(vpm_mesh)]),
    'vpm_layout_build': (c_int, [c_void_p, POINTER(vpm_mesh), _P, c_int64, _P, _P, _P]),
    'vpm_route_build': (c_int, [c_void_p, POINTER(vpm_mesh), c_int, _P, c_int64, _P, c_int64,
                                POINTER(c_int32)]),
    'vpm_route_build_static': (c_int, [c_void_p, POINTER(vpm_mesh), c_int, _P, c_int64, POINTER(c_int32), _P, _P]),
    'vpm_ctx_set_fill_row': (c_int, [c_void_p, POINTER(c_double)]),
    'vpm_route_mark_padding': (c_int, [c_void_p, _P, c_int64, _P, c_int]),
    'vpm_take_rows': (c_int, [c_void_p, _P, _P, c_int64, c_int, _P]),
    'vpm_take_rows_scaled': (c_int, [c_void_p, _P, _P, c_int64, c_int, c_double, _P]),
    'vpm_scatter_rows': (c_int, [c_void_p, _P, _P, c_int64, c_int, _P]),
    'vpm_scatter_add_rows': (c_int, [c_void_p, _P, _P, c_int64, c_int, _P]),
    'vpm_take_rows2': (c_int, [c_void_p, _P, _P, _P, c_int64, c_int, _P]),
    'vpm_take_rows2_scaled': (c_int, [c_void_p, _P, _P, _P, c_int64, c_int, c_double, _P]),
    'vpm_scatter_add_rows_scaled': (c_int, [c_void_p, _P, _P, c_int64, c_int, c_double, _P]),
    'vpm_route_compose': (c_int, [c_void_p, _P, c_int64, _P, c_int64, c_int64, c_int64, c_int64, _P, _P, _P]),
    'vpm_paint_sorted': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_double, c_int64, _P, _P]),
    'vpm_paint_sorted_cols3': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_int, c_int64, _P, _P, c_int64]),
    'vpm_paint_sorted_col': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_int, c_int, c_int64, _P, _P]),
    'vpm_paint_set_algorithm': (c_int, [c_int]),
    'vpm_readout_set_variant': (c_int, [c_int]),
    'vpm_paint_set_tile': (c_int, [c_int, c_int, c_int]),
    'vpm_paint_sorted_rows': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_double, c_int64, _P, _P]),
    'vpm_paint_jvp_sorted': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_double, _P, _P,
                                     c_int64, _P, _P]),
    'vpm_paint_atomic': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_double, c_int64, _P]),
    'vpm_readout': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, c_int64, _P]),
    'vpm_readout3': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, _P, _P, c_int64, _P, _P, _P, c_double, _P]),
    'vpm_readout3_home': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, _P, _P, c_int64, _P, _P, _P, c_double, _P]),
    'vpm_readout_grad': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, _P, c_double, c_int64,
                                 _P, _P]),
    'vpm_readout_grad4': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P, c_double, _P]),
    'vpm_readout_grad4_home': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, _P, _P, _P, _P, c_int64, _P, _P, _P, _P]),
    'vpm_readout_jvp': (c_int, [c_void_p, POINTER(vpm_mesh), _P, _P, _P, _P, c_int64, _P]),
    'vpm_r2c': (c_int, [c_void_p, POINTER(c_int64), _P, _P, c_double]),
    'vpm_c2r': (c_int, [c_void_p, POINTER(c_int64), _P, _P, c_double]),
    'vpm_c2r_inplace': (c_int, [c_void_p, POINTER(c_int64), _P, _P, c_double]),
    'vpm_fft_r2c_planes': (c_int, [c_void_p, c_int64, c_int64, c_int64, _P, _P]),
    'vpm_fft_c2r_planes': (c_int, [c_void_p, c_int64, c_int64, c_int64, _P, _P, c_double]),
    'vpm_fft_c2c_axis0': (c_int, [c_void_p, c_int64, c_int64, _P, c_int, c_double]),
    'vpm_fft_c2c_axis0_oop': (c_int, [c_void_p, c_int64, c_int64, _P, _P, c_int, c_double]),
    'vpm_peer_alloc': (c_int, [c_void_p, c_int64, POINTER(c_void_p), ctypes.c_char_p]),
    'vpm_peer_open': (c_int, [c_void_p, ctypes.c_char_p, POINTER(c_void_p)]),
    'vpm_peer_close': (c_int, [c_void_p, c_void_p]),
    'vpm_peer_free': (c_int, [c_void_p, c_void_p]),
    'vpm_peer_put': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, _P, POINTER(c_uint64), c_int64,
                             POINTER(vpm_peer_sync)]),
    'vpm_peer_put_strided': (c_int, [c_void_p, c_int, c_int64, c_int64, c_int64, c_int64, c_int64, _P, POINTER(c_uint64),
                                     c_int64, POINTER(vpm_peer_sync)]),
    'vpm_peer_put_to': (c_int, [c_void_p, c_int, ctypes.c_uint32, c_int64, _P, POINTER(c_uint64), c_int64,
                                POINTER(vpm_peer_sync)]),
    'vpm_peer_put_blocks': (c_int, [c_void_p, c_int, c_int, c_int64, _P, POINTER(c_int32), POINTER(c_int64),
                                    POINTER(c_int64), POINTER(c_uint64), POINTER(vpm_peer_sync)]),
    'vpm_copy2': (c_int, [c_void_p, c_int64, _P, _P, _P, _P]),
    'vpm_peer_put_rows': (c_int, [c_void_p, c_int, c_int64, c_int, _P, _P, c_int, c_int, POINTER(c_int32),
                                  POINTER(c_int64), _P, _P, POINTER(c_double), POINTER(c_uint64),
                                  POINTER(vpm_peer_sync)]),
    'vpm_peer_read_mail': (c_int, [c_void_p, c_int, _P, c_int, _P]),
    'vpm_peer_set_spin_limit': (c_int, [c_void_p, c_uint64]),
    'vpm_swap_leading_axes': (c_int, [c_void_p, c_int64, c_int64, c_int64, _P, _P]),
    'vpm_transfer': (c_int, [c_void_p, POINTER(c_int64), _P, _P, c_double, c_int, _P, _P, _P,
                             _P, _P, _P, c_int]),
    'vpm_transfer_table': (c_int, [c_void_p, POINTER(c_int64), _P, _P, _P, POINTER(c_int64),
                                   c_int, c_int]),
    'vpm_digitized_mul': (c_int, [c_void_p, c_int64, _P, _P, _P, c_int, _P, c_double, c_int, _P]),
    'vpm_digitized_bin': (c_int, [c_void_p, POINTER(c_int64), c_int64, _P, _P, _P, _P, c_int, _P, _P]),
    'vpm_force_kspace': (c_int, [c_void_p, POINTER(c_int64), _P, c_double, _P, _P, _P, _P, _P,
                                 _P, _P, _P, _P, _P]),
    'vpm_force_kspace_vjp': (c_int, [c_void_p, POINTER(c_int64), _P, _P, _P, _P, c_double, _P, _P, _P,
                                     _P, _P, _P, _P]),
    'vpm_fd4_neg_gradient': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), _P, _P, _P, _P]),
    'vpm_fd4_neg_gradient_adjoint': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), _P, _P, _P, _P]),
    'vpm_fd4_set_algorithm': (c_int, [c_int]),
    'vpm_fd4_neg_gradient_slab': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), c_int, _P, _P, _P, _P]),
    'vpm_fd4_neg_gradient_adjoint_slab': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), c_int, _P, _P, _P, _P]),
    'vpm_lpt2_source': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), _P, _P, _P, _P, _P, _P, c_double, _P]),
    'vpm_lpt2_source_adjoint': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), _P, _P, _P, _P, _P, _P, _P, _P]),
    'vpm_lpt2_source_slab': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), c_int, _P, _P, _P, _P, _P, _P, c_double, _P]),
    'vpm_lpt2_adjoint_w_slab': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), c_int, _P, _P, _P, _P, _P]),
    'vpm_lpt2_adjoint_f_slab': (c_int, [c_void_p, POINTER(c_int64), POINTER(c_double), c_int, _P, _P, _P, _P]),
    'vpm_axpby': (c_int, [c_void_p, c_int64, c_double, _P, c_double, _P, c_double, _P]),
    'vpm_drift_pos': (c_int, [c_void_p, c_int64, _P, _P, c_double, _P, _P, _P]),
    'vpm_mul': (c_int, [c_void_p, c_int64, _P, _P, _P]),
    'vpm_div': (c_int, [c_void_p, c_int64, _P, _P, _P]),
    'vpm_mul_rows': (c_int, [c_void_p, c_int64, c_int, _P, _P, _P]),
    'vpm_cmul': (c_int, [c_void_p, c_int64, _P, _P, c_int, _P]),
    'vpm_fill': (c_int, [c_void_p, c_int64, c_double, _P]),
    'vpm_stack': (c_int, [c_void_p, c_int64, c_int, POINTER(c_void_p), _P]),
    'vpm_take_col': (c_int, [c_void_p, c_int64, c_int, c_int, _P, _P]),
    'vpm_reduce_sum': (c_int, [c_void_p, c_int64, _P, POINTER(c_double)]),
    'vpm_reduce_dot': (c_int, [c_void_p, c_int64, _P, _P, POINTER(c_double)]),
    'vpm_reduce_cdot': (c_int, [c_void_p, POINTER(c_int64), c_int64, _P, _P, POINTER(c_double)]),
}

for _name, (_res, _args) in SIGNATURES.items():
    _f = getattr(lib, _name)  # AttributeError here means header and library disagree
    _f.restype = _res
    _f.argtypes = _args


def last_error():
    msg = lib.vpm_last_error()
    return msg.decode() if msg else ''


def check(status):
    if status != 0:
        raise VpmError(last_error() or ('vmadpm call failed with status %d' % status))


def launch_count():
    return int(lib.vpm_launch_count())


class profile(object):
    """``with profile() as p: ...`` -- every launch of the library inside the block is timed in
    place with CUDA events on the launching stream; afterwards ``p.records`` is a list of
    ``(group, kernel, ms, algorithmic_bytes, group_bytes)``."""

    def __enter__(self):
        check(lib.vpm_profile_start())
        self.records = []
        return self

    def __exit__(self, *exc):
        lib.vpm_profile_stop()
        need = int(lib.vpm_profile_read(None, 0))
        if need < 0:
            raise VpmError(last_error())
        buf = ctypes.create_string_buffer(need)
        if int(lib.vpm_profile_read(buf, need)) < 0:
            raise VpmError(last_error())
        for line in buf.value.decode().splitlines():
            g, k, ms, b, gb = line.split('\t')
            k = k.strip()
            if k.startswith('(') and k.endswith(')'):
                k = k[1:-1]
            self.records.append((g, k, float(ms), float(b), float(gb)))
        return False


def i3(values):
    return _I3(*[int(v) for v in values])
