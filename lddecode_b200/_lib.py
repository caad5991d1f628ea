"""ctypes binding of libldd_b200.so (C ABI declared in include/ldd_b200.h).

There is no CPU fallback: if the shared library is missing or no CUDA device is visible the
package raises.  (tests/emu builds a CPU emulation of the same sources for GPU-less unit tests
and injects it through the private `path=` argument; the package itself never looks for it.)
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PATH = os.path.join(HERE, "libldd_b200.so")

ABI_VERSION = 2
SYSTEM = {"NTSC": 0, "PAL": 1}
FMT_U8, FMT_S16, FMT_U16, FMT_R30, FMT_LDS40 = range(5)
F_RFVIDEO, F_VIDEO, F_VIDEO05, F_BURST, F_PILOT, F_AUDIO_L, F_AUDIO_R, F_AUDIO_LPF2, F_MTF = range(9)
P_DEMOD, P_DEMOD05, P_SYNC, P_BURST, P_PILOT = range(5)
PREC_F64, PREC_F32, PREC_MIXED = 0, 1, 2
OK, EINVAL, ESHORT, ECUDA, ENOMEM, ECAP = 0, -1, -2, -3, -4, -5


class LddError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("ldd_b200 error %d: %s" % (code, msg))
        self.code = code


class Config(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int), ("device", C.c_int), ("system", C.c_int), ("blocklen", C.c_int),
        ("blockcut", C.c_int), ("blockcut_end", C.c_int), ("f05_offset", C.c_int), ("precision", C.c_int),
        ("decode_analog_audio", C.c_int), ("audio_slice_lo", C.c_int), ("audio_slice_hi", C.c_int),
        ("linelen", C.c_int), ("outlinelen", C.c_int),
        ("freq_hz", C.c_double), ("freq_arf", C.c_double), ("audio_lowfreq", C.c_double),
        ("ire0", C.c_double), ("hz_ire", C.c_double), ("vsync_ire", C.c_double),
        ("sync_lo_hz", C.c_double), ("sync_hi_hz", C.c_double),
        ("fpsync_b0", C.c_double), ("fpsync_b1", C.c_double), ("fpsync_a1", C.c_double),
    ]


class Range(C.Structure):
    _fields_ = [("first_sample", C.c_longlong), ("nblocks", C.c_longlong), ("total_out", C.c_longlong),
                ("audio1_len", C.c_longlong), ("audio2_len", C.c_longlong), ("last_needed", C.c_longlong)]


# name -> (restype, argtypes); every symbol include/ldd_b200.h declares
SIGNATURES = {
    "ldd_abi_version": (C.c_int, []),
    "ldd_device_count": (C.c_int, []),
    "ldd_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "ldd_destroy": (None, [C.c_void_p]),
    "ldd_last_error": (C.c_char_p, [C.c_void_p]),
    "ldd_set_filter": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_int]),
    "ldd_unpack_r30_ddunpack": (C.c_int, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ldd_unpack_raw": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ldd_unpack_f32": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "ldd_demod_range_query": (C.c_int, [C.c_void_p, C.c_longlong, C.c_longlong, C.POINTER(Range)]),
    "ldd_demod_range": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong,
                                  C.c_longlong, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_demod_blocks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong,
                                   C.c_longlong, C.c_longlong, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p,
                                   C.c_longlong, C.c_void_p]),
    "ldd_demodblock": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.POINTER(C.c_void_p), C.c_void_p,
                                 C.c_void_p, C.c_void_p]),
    "ldd_mixed_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "ldd_audio_phase2": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_sync_peaks": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int,
                                 C.c_void_p, C.c_void_p]),
    "ldd_window_peaks_from_global": (C.c_int, [C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong, C.c_int,
                                              C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ldd_copy_small": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ldd_peaks_to_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_sync_peaks_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int,
                                      C.POINTER(C.c_int)]),
    "ldd_tbc_fields": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                                 C.c_double, C.c_void_p, C.c_void_p]),
    "ldd_field_locate": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_int]),
    "ldd_field_chain": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong,
                                  C.c_longlong, C.c_longlong, C.c_longlong, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ldd_refine_hsync": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_refine_burst": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_refine_pilot": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                   C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
}


class FieldInfo(C.Structure):
    _fields_ = [("stage", C.c_int), ("istop", C.c_int), ("linecount", C.c_int), ("npeaks", C.c_int), ("nvsyncs", C.c_int),
                ("vsyncs", (C.c_int * 3) * 4), ("nextfieldoffset", C.c_longlong), ("tbcstart", C.c_longlong),
                ("med_hsync", C.c_double), ("hsync_tolerance", C.c_double)]


class PipeBufs(C.Structure):
    _fields_ = [("planes", C.c_void_p * 5), ("plane_cap", C.c_longlong),
                ("audio1_l", C.c_void_p), ("audio1_r", C.c_void_p), ("audio1_cap", C.c_longlong),
                ("audio2_l", C.c_void_p), ("audio2_r", C.c_void_p),
                ("peaks", C.c_void_p), ("peak_vals", C.c_void_p), ("peak_count", C.c_void_p), ("peak_cap", C.c_int),
                ("field_tables", C.c_void_p), ("tables_bytes", C.c_longlong),
                ("h_peaks", C.c_void_p), ("h_peak_vals", C.c_void_p), ("h_peak_count", C.c_void_p),
                ("h_tables", C.c_void_p), ("h_tables_bytes", C.c_longlong),
                ("h_prefix", C.c_void_p), ("prefix_cap", C.c_longlong)]


class PipeResult(C.Structure):
    _fields_ = [("nwindows", C.c_int), ("nowned", C.c_int), ("nlocated", C.c_int), ("npeaks", C.c_int), ("nframes", C.c_int),
                ("prefix_windows", C.c_int), ("ll_stride", C.c_int),
                ("plane_origin", C.c_longlong), ("plane_len", C.c_longlong), ("walk_start", C.c_longlong),
                ("audio1_len", C.c_longlong), ("audio2_len", C.c_longlong), ("lineloc_add", C.c_double),
                ("fields", C.POINTER(FieldInfo)), ("base", C.POINTER(C.c_longlong)), ("winlen", C.POINTER(C.c_longlong)),
                ("readsample", C.POINTER(C.c_longlong)), ("linelocs1", C.POINTER(C.c_double)), ("linebad", C.POINTER(C.c_ubyte)),
                ("owned", C.POINTER(C.c_int)), ("located", C.POINTER(C.c_int)), ("frame_of", C.POINTER(C.c_int)),
                ("gpeaks", C.POINTER(C.c_longlong)), ("gvals", C.POINTER(C.c_double)),
                ("d_base", C.c_void_p), ("d_winlen", C.c_void_p), ("d_linecount", C.c_void_p), ("d_linelocs1", C.c_void_p),
                ("d_linelocs2", C.c_void_p), ("d_linebad2", C.c_void_p), ("d_linelocs3", C.c_void_p), ("d_linelocs4", C.c_void_p),
                ("d_burstlevel", C.c_void_p), ("d_final", C.c_void_p), ("d_vbi", C.c_void_p)]


SIGNATURES.update({
    "ldd_set_mtf_level": (C.c_int, [C.c_void_p, C.c_double, C.c_void_p]),
    "ldd_set_mtf_ramp": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double]),
    "ldd_downscale_audio": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_double,
                                      C.c_double, C.c_double, C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_tbc_fields_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                    C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                                    C.c_longlong, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "ldd_tbc_long_lines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_double, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_longlong, C.c_void_p,
                                     C.c_longlong, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]),
    "ldd_pipe_long_lines": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "ldd_vbi_decode": (C.c_int, [C.c_void_p, C.c_void_p, C.c_longlong, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                 C.POINTER(C.c_int), C.c_int, C.c_void_p, C.c_void_p]),
    "ldd_peer_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p), C.c_void_p]),
    "ldd_peer_open": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "ldd_peer_close": (C.c_int, [C.c_void_p]),
    "ldd_peer_free": (C.c_int, [C.c_void_p]),
    "ldd_peer_read": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "ldd_peer_copy": (C.c_int, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "ldd_peer_signal": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p]),
    "ldd_peer_wait": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ldd_pipe_table_bytes": (C.c_int, [C.c_int, C.POINTER(C.c_longlong), C.POINTER(C.c_longlong)]),
    "ldd_pipe_create": (C.c_int, [C.c_void_p, C.POINTER(PipeBufs), C.c_int, C.c_longlong, C.POINTER(C.c_void_p)]),
    "ldd_pipe_destroy": (None, [C.c_void_p]),
    "ldd_pipe_launch": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_longlong, C.c_longlong, C.c_longlong,
                                  C.c_longlong, C.c_longlong, C.c_double, C.c_int, C.c_void_p]),
    "ldd_pipe_finish": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.POINTER(PipeResult)]),
    "ldd_field_vote": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_longlong, C.c_double, C.c_double, C.c_int,
                                 C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "ldd_pcm_chain": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double),
                                C.POINTER(C.c_int), C.c_void_p]),
    "ldd_pipe_pcm": (C.c_int, [C.c_void_p, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, C.c_int,
                               C.POINTER(C.c_double), C.POINTER(C.c_int), C.c_void_p, C.c_longlong, C.POINTER(C.c_longlong),
                               C.c_void_p, C.c_void_p]),
})
PCM_CHAIN_FIELDS, PCM_CHAIN_FRAMER = 0, 1
ST_LINE_LONG, ST_LINE_BAD = 32, 64

WINDOW_PEAKS_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_longlong, C.c_longlong, C.POINTER(C.c_void_p),
                              C.POINTER(C.c_void_p), C.POINTER(C.c_int))

FIELD_NOVSYNC, FIELD_SHORT, FIELD_LOCATED, FIELD_BADLINES, FIELD_CRASH = range(5)

_cache = {}


def load(path=None):
    """Loads the shared library and declares every entry point.  Raises if it is missing."""
    path = os.path.abspath(path or DEFAULT_PATH)
    if path in _cache:
        return _cache[path]
    if not os.path.exists(path):
        raise ImportError(
            "%s not found: build it with `python lddecode_b200/csrc/build.py` (nvcc, sm_100a). "
            "lddecode_b200 has no CPU fallback." % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the library does not export it
        fn.restype = res
        fn.argtypes = args
    if lib.ldd_abi_version() != ABI_VERSION:
        raise ImportError("ABI version mismatch in " + path)
    _cache[path] = lib
    return lib
