"""ctypes binding of ``libeae_b200.so`` (C ABI declared in ``include/eae_b200.h``).

The library is the product: there is no Python or CPU fallback behind any of these calls. Importing
this module never touches the GPU; ``lib()`` loads the shared object (raising ``RuntimeError`` if it
has not been built) and every compute entry point returns ``EAE_ERR_CUDA`` when no CUDA device is
present, which ``check()`` turns into ``RuntimeError``.
"""
import ctypes
import os

import numpy

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, 'libeae_b200.so')

EAE_NB_MAPS = 128
MATH_FP32_SIMT = 0
MATH_TF32X3 = 1
MATH_TF32 = 2
MATH_MIXED = 3
MATH_NAMES = {'fp32': MATH_FP32_SIMT, 'tf32x3': MATH_TF32X3, 'tf32': MATH_TF32, 'mixed': MATH_MIXED}

ERR_NULL = -1
ERR_UNARY_LENGTH = -2
ERR_ARGUMENT = -3
ERR_INT16_RANGE = -4
ERR_ROUND_TRIP = -5
ERR_CUDA = -10

c_void_p = ctypes.c_void_p
c_int = ctypes.c_int
u8 = ctypes.c_uint8
u32 = ctypes.c_uint32
u64 = ctypes.c_uint64
P = ctypes.POINTER


class Weights(ctypes.Structure):
    """``eae_weights_t``: host pointers to the TF-layout variables."""
    FIELDS = ['weights_1', 'biases_1', 'gamma_1', 'beta_1', 'weights_2', 'biases_2', 'gamma_2', 'beta_2',
              'weights_3', 'biases_3', 'gamma_3', 'beta_3', 'gamma_4', 'beta_4', 'weights_4', 'biases_4',
              'gamma_5', 'beta_5', 'weights_5', 'biases_5', 'gamma_6', 'beta_6', 'weights_6']
    _fields_ = [(name, c_void_p) for name in FIELDS]


class CodingParams(ctypes.Structure):
    """``eae_coding_params_t``."""
    _fields_ = [('map_mean', c_void_p), ('bin_widths', c_void_p), ('table', c_void_p),
                ('truncated_unary_length', u32)]


class BatchStats(ctypes.Structure):
    """``eae_batch_stats_t``."""
    _fields_ = [('bits_per_map', u64*EAE_NB_MAPS), ('total_bits', u64), ('nb_dead_maps', u64)]


class CodecStatus(ctypes.Structure):
    """``eae_codec_status_t``."""
    _fields_ = [('int16_overflow', u32), ('container_overflow', u32), ('container_invalid', u32), ('coder_error', u32),
                ('tensor_timeout_mask', u32)]


# name -> (restype, argtypes). Every symbol declared in include/eae_b200.h is listed here and
# tests/test_abi.py checks the two against each other.
PROTOTYPES = {
    'eae_last_error': (ctypes.c_char_p, []),
    'eae_abi_version': (c_int, []),
    'eae_device_count': (c_int, []),
    'eae_device_info': (c_int, [c_int, P(c_int), P(c_int), P(c_int), P(ctypes.c_size_t)]),
    'eae_launch_count': (u64, []),
    'eae_set_device': (c_int, [c_int]),
    'eae_set_blocking_sync': (c_int, [c_int]),
    'eae_profile_enable': (c_int, [c_int]),
    'eae_profile_reset': (c_int, []),
    'eae_profile_read': (c_int, [c_int, P(u64), P(ctypes.c_double)]),
    'eae_profile_name': (ctypes.c_char_p, [c_int]),
    'eae_host_alloc': (c_void_p, [ctypes.c_size_t]),
    'eae_host_free': (None, [c_void_p]),
    'eae_device_alloc': (c_void_p, [ctypes.c_size_t]),
    'eae_device_free': (None, [c_void_p]),
    'eae_memcpy_h2d': (c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    'eae_memcpy_d2h': (c_int, [c_void_p, c_void_p, ctypes.c_size_t, c_void_p]),
    'eae_stream_create': (c_int, [P(c_void_p)]),
    'eae_stream_destroy': (c_int, [c_void_p]),
    'eae_stream_synchronize': (c_int, [c_void_p]),
    'eae_event_create': (c_int, [P(c_void_p)]),
    'eae_event_destroy': (c_int, [c_void_p]),
    'eae_event_record': (c_int, [c_void_p, c_void_p]),
    'eae_event_elapsed_ms': (c_int, [c_void_p, c_void_p, P(ctypes.c_float)]),
    'eae_compress_lossless': (c_int, [u32, c_void_p, c_void_p, u8, c_void_p, P(u32)]),
    'eae_encode_map_host': (c_int, [u32, c_void_p, u8, c_void_p, c_void_p, P(u32), c_void_p, P(u32)]),
    'eae_decode_map_host': (c_int, [u32, c_void_p, u8, c_void_p, c_void_p, u32, c_void_p, u32]),
    'eae_coder_capacity_bytes': (u32, [u32, u32]),
    'eae_compress_lossless_maps_host': (c_int, [c_void_p, u32, u32, u32, c_void_p, u32, c_void_p, c_void_p,
                                                c_void_p, c_void_p]),
    'eae_rescale_compress_lossless_maps_host': (c_int, [c_void_p, u32, u32, u32, c_void_p, c_void_p, u32,
                                                        c_void_p, c_void_p, c_void_p, c_void_p]),
    'eae_histogram_maps_host': (c_int, [c_void_p, u32, u32, u32, u32, c_int, c_void_p, c_void_p, c_void_p,
                                        u32, P(u32), c_void_p, c_void_p]),
    'eae_histogram_streams_dev': (c_int, [c_void_p, u32, u32, u32, c_int, c_void_p, c_void_p, c_void_p, c_void_p, u32,
                                            c_void_p]),
    'eae_coder_slot_bytes': (u32, [u32, u32]),
    'eae_nhwc_to_planar_i16_dev': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p]),
    'eae_planar_to_nhwc_i16_dev': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p]),
    'eae_encode_streams_dev': (c_int, [c_void_p, u32, u32, c_void_p, u32, u32, c_void_p, c_void_p, c_void_p,
                                       u32, c_void_p, c_void_p, c_void_p, c_void_p]),
    'eae_decode_streams_dev': (c_int, [c_void_p, u32, u32, c_void_p, u32, u32, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    'eae_quantize_per_map_host': (c_int, [c_void_p, c_void_p, u64, u32, c_void_p, c_void_p]),
    'eae_cast_float_to_int16_host': (c_int, [c_void_p, c_void_p, u64, c_void_p]),
    'eae_cast_bt601_host': (c_int, [c_void_p, c_void_p, u64, c_void_p]),
    'eae_sum_squared_error_u8_host': (c_int, [c_void_p, c_void_p, u64, P(u64), c_void_p]),
    'eae_count_nb_deads_host': (c_int, [c_void_p, u32, u64, u32, c_void_p, c_void_p]),
    'eae_latent_statistics_host': (c_int, [c_void_p, u64, u32, c_void_p, c_void_p, c_void_p, c_void_p, u32, P(u32),
                                           c_void_p, c_void_p, u32, c_void_p, c_void_p]),
    'eae_codec_create': (c_int, [P(c_void_p), P(Weights), c_int, c_int]),
    'eae_codec_destroy': (c_int, [c_void_p]),
    'eae_codec_set_math': (c_int, [c_void_p, c_int]),
    'eae_codec_get_math': (c_int, [c_void_p]),
    'eae_codec_set_coder_lanes': (c_int, [c_void_p, u32]),
    'eae_encode_host': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p, c_void_p]),
    'eae_encode_dev': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p, c_void_p]),
    'eae_decode_host': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p, c_void_p]),
    'eae_decode_dev': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p, c_void_p]),
    'eae_decode_float_host': (c_int, [c_void_p, c_void_p, u32, u32, u32, c_void_p, c_void_p]),
    'eae_container_bound': (u64, [u32, u32, u32, u32]),
    'eae_compress_host': (c_int, [c_void_p, P(CodingParams), c_void_p, u32, u32, u32, c_void_p, u64, P(u64),
                                  P(BatchStats), c_void_p]),
    'eae_decompress_host': (c_int, [c_void_p, P(CodingParams), c_void_p, u64, c_void_p, u64, c_void_p]),
    'eae_compress_dev': (c_int, [c_void_p, P(CodingParams), c_void_p, u32, u32, u32, c_void_p, u64, c_void_p,
                                 c_void_p, c_void_p]),
    'eae_decompress_dev': (c_int, [c_void_p, P(CodingParams), c_void_p, u64, u32, u32, u32, c_void_p, c_void_p]),
    'eae_codec_set_stats_accumulator': (c_int, [c_void_p, c_void_p]),
    'eae_codec_poll_status': (c_int, [c_void_p, c_void_p, P(CodecStatus)]),
    'eae_debug_check_norm_arithmetic': (c_int, [u64, P(u64), P(u64)]),
    'eae_last_indices_host': (c_int, [c_void_p, c_void_p, u64]),
}

_lib = None


def lib():
    """Loads ``libeae_b200.so`` once. Raises ``RuntimeError`` when it has not been built."""
    global _lib
    if _lib is None:
        path = os.environ.get('EAE_LIB_PATH') or LIB_PATH      # (experiments: a variant build of the same sources)
        if not os.path.isfile(path):
            raise RuntimeError('{} is missing: run `python -m autoencoder_based_image_compression_b200.build` '
                               '(or __graft_entry__.build()). There is no CPU fallback.'.format(path))
        # Every kernel of the library is loaded when its module is (CUDA's default loads a kernel at its first launch,
        # which here would happen in the middle of twelve concurrent pipeline slots); no effect if the process has
        # already initialised CUDA.
        os.environ.setdefault('CUDA_MODULE_LOADING', 'EAGER')
        handle = ctypes.CDLL(path)
        for (name, (restype, argtypes)) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def last_error():
    return lib().eae_last_error().decode('utf-8', 'replace')


class CoderError(RuntimeError):
    """The reference's ``std::runtime_error("Error of type N ...")`` (compression.cpp:32-62), which
    Cython's ``except +`` surfaces as ``RuntimeError``. ``code`` is the reference's ``error_code``."""

    def __init__(self, code, message):
        RuntimeError.__init__(self, message)
        self.code = code


def check(code):
    """Maps a C-ABI return code to the exception the reference raises in the same situation."""
    if code == 0:
        return
    msg = last_error()
    if 1 <= code <= 4:
        raise CoderError(code, msg or 'Error of type {} during the encoding.'.format(code))
    if code == ERR_NULL:
        raise ValueError(msg or 'One of the pointers is NULL.')          # std::invalid_argument
    if code == ERR_UNARY_LENGTH:
        raise IndexError(msg or 'truncated unary length is 0')           # std::out_of_range
    if code == ERR_ARGUMENT:
        raise ValueError(msg)
    if code in (ERR_INT16_RANGE, ERR_ROUND_TRIP):
        raise AssertionError(msg)
    raise RuntimeError('libeae_b200: {} (code {})'.format(msg, code))


def ptr(array):
    """Host pointer of a C-contiguous numpy array (or None)."""
    if array is None:
        return None
    if not array.flags['C_CONTIGUOUS']:
        raise ValueError('array is not C-contiguous')
    return array.ctypes.data_as(c_void_p)


def device_count():
    return lib().eae_device_count()


def require_gpu():
    if device_count() <= 0:
        raise RuntimeError('libeae_b200: no CUDA device available; the product path has no CPU fallback.')


def pinned_empty(shape, dtype):
    """numpy array over pinned host memory (``cudaHostAlloc``) for the end-to-end path."""
    dtype = numpy.dtype(dtype)
    nbytes = int(numpy.prod(shape))*dtype.itemsize
    raw = lib().eae_host_alloc(max(nbytes, 16))
    if not raw:
        raise RuntimeError('libeae_b200: ' + last_error())
    buf = (ctypes.c_uint8*max(nbytes, 16)).from_address(raw)
    array = numpy.frombuffer(buf, dtype=dtype, count=int(numpy.prod(shape))).reshape(shape)
    _PINNED[array.ctypes.data] = raw
    return array


_PINNED = {}


def pinned_free(array):
    raw = _PINNED.pop(array.ctypes.data, None)
    if raw:
        lib().eae_host_free(raw)
