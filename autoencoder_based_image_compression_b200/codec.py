"""Python face of the native codec handle (``eae_codec_t``) and of the fused pipeline.

``Session`` takes the place of ``tf.Session`` in the reference's call sites
(kodak_tensorflow/reconstructing_eae_kodak.py:142,179): an opaque device / stream handle passed in the
same positional slot. ``Codec`` owns the device copy of the weights and the workspaces;
``compress`` / ``decompress`` run encode -> quantize -> lossless code -> container and back entirely
on the GPU (one H2D of the uint8 images, one D2H of the container).
"""
import ctypes

import numpy

from autoencoder_based_image_compression_b200 import _native
from autoencoder_based_image_compression_b200 import weights as wts


class Session(object):
    """Device / stream handle standing in for ``tf.Session``. Usable as a context manager."""

    def __init__(self, device=0, math='fp32'):
        self.device = device
        self.math = math

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def close(self):
        pass


class CodingParams(object):
    """One operating point (reconstructing_eae_kodak.py:170-217): per-map means, test bin widths and
    the binary probability table ``float64 [128, L]``."""

    def __init__(self, bin_widths, table, map_mean=None):
        self.bin_widths = numpy.ascontiguousarray(bin_widths, dtype=numpy.float32)
        self.table = numpy.ascontiguousarray(table, dtype=numpy.float64)
        self.map_mean = None if map_mean is None else numpy.ascontiguousarray(map_mean, dtype=numpy.float32)
        if self.bin_widths.shape != (_native.EAE_NB_MAPS,):
            raise ValueError('`bin_widths` must have shape (128,).')
        if self.table.ndim != 2:
            raise ValueError('`binary_probabilities.ndim` is not equal to 2.')
        if self.table.shape[0] != _native.EAE_NB_MAPS:
            raise ValueError('`binary_probabilities.shape[0]` is not equal to 128.')
        if self.map_mean is not None and self.map_mean.shape != (_native.EAE_NB_MAPS,):
            raise ValueError('`map_mean` must have shape (128,).')

    @property
    def truncated_unary_length(self):
        return self.table.shape[1]

    def native(self):
        p = _native.CodingParams()
        p.map_mean = None if self.map_mean is None else self.map_mean.ctypes.data
        p.bin_widths = self.bin_widths.ctypes.data
        p.table = self.table.ctypes.data
        p.truncated_unary_length = self.table.shape[1]
        return p


class Codec(object):
    """Both transforms of one EAE on one GPU."""

    def __init__(self, weights, are_bin_widths_learned, device=0, math='fp32', own_stream=False):
        """``own_stream``: run this codec's work on its own non-blocking CUDA stream, so that several
        codecs (pipeline slots) overlap on one GPU; by default the legacy default stream is used."""
        wts.validate(weights, are_bin_widths_learned)
        self.are_bin_widths_learned = bool(are_bin_widths_learned)
        self.device = device
        self._keep = {}
        native_w = _native.Weights()
        for field in _native.Weights.FIELDS:
            scope = 'encoder/' if int(field[-1]) <= 3 else 'decoder/'
            key = scope + field
            if key in weights:
                a = numpy.ascontiguousarray(weights[key], dtype=numpy.float32)
                self._keep[key] = a
                setattr(native_w, field, a.ctypes.data)
            else:
                setattr(native_w, field, None)
        self._handle = ctypes.c_void_p()
        _native.check(_native.lib().eae_codec_create(ctypes.byref(self._handle), ctypes.byref(native_w),
                                                     int(self.are_bin_widths_learned), device))
        self.set_math(math)
        self.stream = None
        if own_stream:
            stream = ctypes.c_void_p()
            _native.check(_native.lib().eae_stream_create(ctypes.byref(stream)))
            self.stream = stream

    def close(self):
        if getattr(self, '_handle', None):
            _native.lib().eae_codec_destroy(self._handle)
            self._handle = None
        if getattr(self, 'stream', None):
            _native.lib().eae_stream_destroy(self.stream)
            self.stream = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        if not self._handle:
            raise RuntimeError('codec is closed')
        return self._handle

    def set_math(self, math):
        mode = _native.MATH_NAMES[math] if isinstance(math, str) else int(math)
        _native.check(_native.lib().eae_codec_set_math(self.handle, mode))

    def set_coder_lanes(self, lanes):
        """GPU threads per coded stream (0 = auto / lowest latency; 1, 2, 4 = throughput-oriented packing)."""
        _native.check(_native.lib().eae_codec_set_coder_lanes(self.handle, int(lanes)))

    # ---- transforms, host arrays (the reference's sess.run boundaries) ----
    def encode(self, luminances_uint8):
        """uint8 [n, h, w, 1] (or [n, h, w]) -> float32 [n, h/16, w/16, 128] (components.py:86-142)."""
        x = numpy.ascontiguousarray(luminances_uint8)
        (n, h, w) = x.shape[:3]
        y = numpy.empty((n, h//16, w//16, 128), dtype=numpy.float32)
        _native.check(_native.lib().eae_encode_host(self.handle, _native.ptr(x), n, h, w, _native.ptr(y), self.stream))
        return y

    def decode(self, quantized_y_float32, h_in=None, w_in=None):
        """float32 [n, h/16, w/16, 128] -> uint8 [n, h, w, 1] (components.py:11-84 + cast_bt601)."""
        q = numpy.ascontiguousarray(quantized_y_float32, dtype=numpy.float32)
        (n, hl, wl, _) = q.shape
        (h, w) = (hl*16, wl*16)
        out = numpy.empty((n, h, w, 1), dtype=numpy.uint8)
        _native.check(_native.lib().eae_decode_host(self.handle, _native.ptr(q), n, h, w, _native.ptr(out), self.stream))
        return out

    def decode_float(self, quantized_y_float32):
        """The decoder's float32 output ``node_reconstruction`` before ``cast_bt601`` (components.py:79-84)."""
        q = numpy.ascontiguousarray(quantized_y_float32, dtype=numpy.float32)
        (n, hl, wl, _) = q.shape
        out = numpy.empty((n, hl*16, wl*16, 1), dtype=numpy.float32)
        _native.check(_native.lib().eae_decode_float_host(self.handle, _native.ptr(q), n, hl*16, wl*16,
                                                          _native.ptr(out), self.stream))
        return out

    # ---- fused pipeline ----
    def compress(self, luminances_uint8, params, container=None, return_stats=False):
        """uint8 [n, h, w] -> container bytes (numpy uint8 view). See include/eae_b200.h for the layout."""
        x = numpy.ascontiguousarray(luminances_uint8)
        if x.dtype != numpy.uint8:
            raise TypeError('`luminances_uint8.dtype` is not equal to `numpy.uint8`.')
        if x.ndim == 4:
            x = x[:, :, :, 0] if x.shape[3] == 1 else x
        (n, h, w) = x.shape
        x = numpy.ascontiguousarray(x)
        bound = _native.lib().eae_container_bound(n, h, w, params.truncated_unary_length)
        if container is None:
            # A container is almost always far below the worst case; start at a quarter of the
            # bound and retry at the bound if the library says it does not fit.
            container = numpy.empty(max(int(bound)//4, 4096), dtype=numpy.uint8)
        nbytes = ctypes.c_uint64(0)
        stats = _native.BatchStats()
        native_p = params.native()
        code = _native.lib().eae_compress_host(self.handle, ctypes.byref(native_p), _native.ptr(x), n, h, w,
                                               _native.ptr(container), container.size, ctypes.byref(nbytes),
                                               ctypes.byref(stats), self.stream)
        if code == _native.ERR_ARGUMENT and container.size < bound and 'container needs' in _native.last_error():
            container = numpy.empty(int(bound), dtype=numpy.uint8)
            code = _native.lib().eae_compress_host(self.handle, ctypes.byref(native_p), _native.ptr(x), n, h, w,
                                                   _native.ptr(container), container.size, ctypes.byref(nbytes),
                                                   ctypes.byref(stats), self.stream)
        _native.check(code)
        out = container[:nbytes.value]
        if return_stats:
            return (out, {'bits_per_map': numpy.array(list(stats.bits_per_map), dtype=numpy.uint64),
                          'total_bits': int(stats.total_bits), 'nb_dead_maps': int(stats.nb_dead_maps)})
        return out

    def decompress(self, container, params, out=None):
        """Container bytes -> uint8 [n, h, w]."""
        c = numpy.ascontiguousarray(container, dtype=numpy.uint8)
        if c.size < 32:
            raise ValueError('container shorter than its header')
        hdr = c[:32].view(numpy.uint32)
        (n, h, w) = (int(hdr[2]), int(hdr[3]), int(hdr[4]))
        if out is None:
            out = numpy.empty((n, h, w), dtype=numpy.uint8)
        native_p = params.native()
        _native.check(_native.lib().eae_decompress_host(self.handle, ctypes.byref(native_p), _native.ptr(c), c.size,
                                                        _native.ptr(out), out.size, self.stream))
        return out

    def poll_status(self, raise_on_error=True):
        """What the device-resident steps (``eae_compress_dev`` / ``eae_decompress_dev``) of this codec recorded since the
        previous poll; waits for the codec's stream. Returns a dict; raises like the host entry points would have."""
        status = _native.CodecStatus()
        code = _native.lib().eae_codec_poll_status(self.handle, self.stream, ctypes.byref(status))
        out = {name: int(getattr(status, name)) for (name, _) in _native.CodecStatus._fields_}
        out['code'] = int(code)
        if raise_on_error:
            _native.check(code)
        return out

    def last_indices(self, n, h, w):
        """int16 [n, 128, h/16 * w/16] produced by the last compress / decompress (parity hook)."""
        out = numpy.empty((n, 128, (h//16)*(w//16)), dtype=numpy.int16)
        _native.check(_native.lib().eae_last_indices_host(self.handle, _native.ptr(out), out.size))
        return out


def parse_container(container):
    """Splits a container into its header dict and per-stream (bac_bits, bypass_bits, bac, bypass)."""
    c = numpy.ascontiguousarray(container, dtype=numpy.uint8)
    hdr = c[:32].view(numpy.uint32)
    if hdr[0] != 0x42454145 or hdr[1] != 1:
        raise ValueError('not an EAEB v1 container')
    (n, h, w, nb_maps, L) = (int(hdr[2]), int(hdr[3]), int(hdr[4]), int(hdr[5]), int(hdr[6]))
    n_streams = n*nb_maps
    table = c[32:32 + 8*n_streams].view(numpy.uint32).reshape(n_streams, 2)
    streams = []
    at = 32 + 8*n_streams
    for s in range(n_streams):
        (bb, rb) = (int(table[s, 0]), int(table[s, 1]))
        (nb, nr) = ((bb + 7)//8, (rb + 7)//8)
        streams.append((bb, rb, c[at:at + nb], c[at + nb:at + nb + nr]))
        at += nb + nr
    return ({'n': n, 'h': h, 'w': w, 'nb_maps': nb_maps, 'L': L, 'bytes': at}, streams)
