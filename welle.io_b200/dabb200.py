"""ctypes binding of libdab_b200.so (include/dab_b200.h) — the same C ABI a cgo/JNI/N-API stub would bind.

No computation happens in Python and there is no CPU fallback: if the shared library is missing or no CUDA device
is present, construction fails loudly.
"""
import ctypes as C
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DABB_LIB") or os.path.join(HERE, "libdab_b200.so")   # DABB_LIB: build-variant experiments only

L, K, TU, TS, TG, TNULL, TF = 76, 1536, 2048, 2552, 504, 2656, 196608
SOFT_PER_FRAME = 75 * 3072
MAX_SUBCH = 4
FFT_EXACT, FFT_FMA = 0, 1
NCO_EXACT, NCO_FAST = 0, 1
IQ_CF32, IQ_U8, IQ_S8, IQ_S16LE, IQ_S16BE = 0, 1, 2, 3, 4
FRAME_DECODED, FRAME_NEED_SAMPLES, FRAME_NO_SYNC, FRAME_ACQUIRING = 0, 1, 2, 3


class DabbError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("device", C.c_int32), ("n_streams", C.c_int32), ("transmission_mode", C.c_int32),
                ("fft_mode", C.c_int32), ("disable_coarse", C.c_int32), ("keep_taps", C.c_int32), ("n_subch_slots", C.c_int32),
                ("max_subch_cu", C.c_int32), ("ofdm_groups", C.c_int32), ("fft_placement", C.c_int32), ("freqsync_method", C.c_int32),
                ("nco_mode", C.c_int32), ("ofdm_tail_split", C.c_int32), ("reserved", C.c_int32 * 2)]


class Options(C.Structure):
    _fields_ = [("disable_coarse", C.c_int32), ("fft_placement", C.c_int32), ("freqsync_method", C.c_int32), ("decode_tii", C.c_int32), ("reserved", C.c_int32 * 4)]


PLACEMENT_THRESHOLD_BEFORE_PEAK, PLACEMENT_STRONGEST_PEAK, PLACEMENT_EARLIEST_PEAK_WITH_BINNING = 0, 1, 2
FREQSYNC_PATTERN_OF_ZEROS, FREQSYNC_GET_MIDDLE, FREQSYNC_CORRELATE_PRS = 0, 1, 2


class Subchannel(C.Structure):
    _fields_ = [("subch_id", C.c_int32), ("start_cu", C.c_int32), ("length_cu", C.c_int32), ("bitrate", C.c_int32),
                ("short_form", C.c_int32), ("uep_level", C.c_int32), ("eep_profile_a", C.c_int32), ("eep_level", C.c_int32),
                ("dabplus", C.c_int32)]


class FrameResult(C.Structure):
    _fields_ = [("status", C.c_int32), ("start_index", C.c_int32), ("fine_corr", C.c_int32), ("coarse_corr", C.c_int32),
                ("snr_raw", C.c_int32), ("fib_crc_mask", C.c_int32), ("fic_ratio", C.c_int32),
                ("n_logical", C.c_int32 * MAX_SUBCH), ("n_rs_events", C.c_int32 * MAX_SUBCH), ("rs_uncorr_mask", C.c_int32 * MAX_SUBCH),
                ("rs_corr", (C.c_int32 * 4) * MAX_SUBCH), ("sf_ready", C.c_int32 * MAX_SUBCH), ("sf_au_count", C.c_int32 * MAX_SUBCH),
                ("sf_au_crc_mask", C.c_int32 * MAX_SUBCH), ("next_pos", C.c_int64), ("freq_corr_re", C.c_float), ("freq_corr_im", C.c_float),
                ("slevel", C.c_float), ("acq_failed", C.c_int32), ("reserved", C.c_int32 * 2)]


RESULT_DTYPE = np.dtype([("status", "<i4"), ("start_index", "<i4"), ("fine_corr", "<i4"), ("coarse_corr", "<i4"), ("snr_raw", "<i4"),
                         ("fib_crc_mask", "<i4"), ("fic_ratio", "<i4"), ("n_logical", "<i4", (4,)), ("n_rs_events", "<i4", (4,)),
                         ("rs_uncorr_mask", "<i4", (4,)), ("rs_corr", "<i4", (4, 4)), ("sf_ready", "<i4", (4,)), ("sf_au_count", "<i4", (4,)),
                         ("sf_au_crc_mask", "<i4", (4,)), ("next_pos", "<i8"), ("freq_corr_re", "<f4"), ("freq_corr_im", "<f4"),
                         ("slevel", "<f4"), ("acq_failed", "<i4"), ("reserved", "<i4", (2,))], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(FrameResult), (RESULT_DTYPE.itemsize, C.sizeof(FrameResult))


class IO(C.Structure):
    _fields_ = [("iq", C.c_void_p), ("iq_is_host", C.c_int32), ("stride_samples", C.c_int64), ("buf_start", C.c_void_p), ("buf_len", C.c_int64),
                ("results", C.c_void_p), ("fibs", C.c_void_p), ("msc", C.c_void_p), ("msc_stride", C.c_int32), ("sf", C.c_void_p), ("sf_stride", C.c_int32),
                ("iq_format", C.c_int32), ("carry_samples", C.c_int32)]


EXPORTS = ["dabb_create", "dabb_destroy", "dabb_last_error", "dabb_abi_version", "dabb_stream_reset", "dabb_set_options", "dabb_get_info", "dabb_select_subchannel",
           "dabb_remove_subchannel", "dabb_process", "dabb_process_async", "dabb_sync", "dabb_join_lanes", "dabb_submit", "dabb_collect", "dabb_cuda_stream", "dabb_kernel_launches",
           "dabb_read_tap", "dabb_profile", "dabb_profile_read", "dabb_ofdm_demod", "dabb_find_index", "dabb_find_index_ex", "dabb_coarse_estimate", "dabb_viterbi", "dabb_fic_decode", "dabb_msc_decode",
           "dabb_rs_superframes", "dabb_dev_alloc", "dabb_dev_free", "dabb_memcpy_h2d", "dabb_memcpy_d2h"]


def build():
    """Compile the CUDA sources for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-s", "-j8", "-C", os.path.join(HERE, "csrc")])


def load_library():
    if not os.path.exists(LIB_PATH):
        raise DabbError(f"{LIB_PATH} is missing: run welle.io_b200.build() (nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    lib.dabb_last_error.restype = C.c_char_p
    lib.dabb_last_error.argtypes = [C.c_void_p]
    lib.dabb_cuda_stream.restype = C.c_void_p
    lib.dabb_cuda_stream.argtypes = [C.c_void_p]
    lib.dabb_kernel_launches.restype = C.c_int64
    lib.dabb_kernel_launches.argtypes = [C.c_void_p]
    lib.dabb_destroy.argtypes = [C.c_void_p]
    lib.dabb_destroy.restype = None
    return lib


def _addr(x):
    """integer address of a numpy array / torch tensor / DevBuf / raw int"""
    if isinstance(x, (int, np.integer)):
        return int(x)
    if isinstance(x, np.ndarray):
        return x.ctypes.data
    if hasattr(x, "ptr"):
        return x.ptr
    if hasattr(x, "data_ptr"):
        return x.data_ptr()
    raise TypeError(type(x))


def _vp(x):
    if x is None:
        return None
    if isinstance(x, np.ndarray):
        return x.ctypes.data_as(C.c_void_p)
    if isinstance(x, int):
        return C.c_void_p(x)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    return x


class DevBuf:
    """Device memory owned through the ABI (so that tests need neither torch nor cuda-python)."""

    def __init__(self, ctx, nbytes):
        self.ctx, self.nbytes = ctx, int(nbytes)
        p = C.c_void_p()
        ctx._ck(ctx.lib.dabb_dev_alloc(ctx.h, C.c_size_t(self.nbytes), C.byref(p)))
        self.ptr = p.value

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes <= self.nbytes
        self.ctx._ck(self.ctx.lib.dabb_memcpy_h2d(self.ctx.h, C.c_void_p(self.ptr), _vp(arr), C.c_size_t(arr.nbytes)))
        return self

    def download(self, dtype, count=None):
        dt = np.dtype(dtype)
        n = self.nbytes // dt.itemsize if count is None else count
        out = np.empty(n, dt)
        self.ctx._ck(self.ctx.lib.dabb_memcpy_d2h(self.ctx.h, _vp(out), C.c_void_p(self.ptr), C.c_size_t(out.nbytes)))
        return out

    def free(self):
        if self.ptr:
            self.ctx.lib.dabb_dev_free(self.ctx.h, C.c_void_p(self.ptr))
            self.ptr = None


class Context:
    def __init__(self, n_streams=1, device=0, fft_mode=FFT_EXACT, disable_coarse=True, keep_taps=False, n_subch_slots=1,
                 max_subch_cu=0, ofdm_groups=0, fft_placement=0, freqsync_method=0, nco_mode=NCO_EXACT, ofdm_tail_split=0):
        self.lib = load_library()
        cfg = Config()
        cfg.abi_version = self.lib.dabb_abi_version()
        cfg.device, cfg.n_streams, cfg.transmission_mode = device, n_streams, 1
        cfg.fft_mode, cfg.disable_coarse, cfg.keep_taps = fft_mode, int(disable_coarse), int(keep_taps)
        cfg.n_subch_slots, cfg.max_subch_cu, cfg.ofdm_groups = n_subch_slots, max_subch_cu, ofdm_groups
        cfg.fft_placement, cfg.freqsync_method = int(fft_placement), int(freqsync_method)
        cfg.nco_mode, cfg.ofdm_tail_split = int(nco_mode), int(ofdm_tail_split)
        h = C.c_void_p()
        rc = self.lib.dabb_create(C.byref(cfg), C.byref(h))
        if rc != 0:
            raise DabbError(f"dabb_create failed ({rc}): {self.lib.dabb_last_error(None).decode()}")
        self.h, self.n_streams, self.n_slots = h, n_streams, max(1, n_subch_slots)
        self._bufs = []

    def _ck(self, rc):
        if rc != 0:
            raise DabbError(f"libdab_b200 error {rc}: {self.lib.dabb_last_error(self.h).decode()}")

    def close(self):
        if getattr(self, "h", None):
            self.lib.dabb_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- memory helpers
    def dev(self, arr_or_bytes):
        if isinstance(arr_or_bytes, (int, np.integer)):
            return DevBuf(self, arr_or_bytes)
        a = np.ascontiguousarray(arr_or_bytes)
        return DevBuf(self, a.nbytes).upload(a)

    def cuda_stream(self):
        return self.lib.dabb_cuda_stream(self.h)

    def kernel_launches(self):
        return self.lib.dabb_kernel_launches(self.h)

    def sync(self):
        self._ck(self.lib.dabb_sync(self.h))

    def join_lanes(self):
        self._ck(self.lib.dabb_join_lanes(self.h))

    def profile(self, enable):
        self._ck(self.lib.dabb_profile(self.h, int(enable)))

    def profile_read(self):
        import json
        buf = C.create_string_buffer(16384)
        self._ck(self.lib.dabb_profile_read(self.h, buf, C.c_size_t(16384)))
        return json.loads(buf.value.decode())

    # ---- receiver control
    def reset(self, first=0, count=None, pos=0):
        self._ck(self.lib.dabb_stream_reset(self.h, first, self.n_streams if count is None else count, C.c_int64(pos)))

    def select_subchannel(self, start_cu, length_cu, bitrate, eep_profile_a=True, eep_level=3, short_form=False, uep_level=0,
                          dabplus=True, slot=0, first=0, count=None, subch_id=0):
        sc = Subchannel(subch_id, start_cu, length_cu, bitrate, int(short_form), uep_level, int(eep_profile_a), eep_level, int(dabplus))
        self._ck(self.lib.dabb_select_subchannel(self.h, first, self.n_streams if count is None else count, slot, C.byref(sc)))

    def remove_subchannel(self, slot=0, first=0, count=None):
        self._ck(self.lib.dabb_remove_subchannel(self.h, first, self.n_streams if count is None else count, slot))

    def process(self, iq, stride, buf_start, buf_len, iq_is_host=False, msc_stride=0, sf_stride=0, want=("results", "fibs"), iq_format=0):
        S = self.n_streams
        bs = np.ascontiguousarray(buf_start, np.int64)
        assert bs.size == S
        io = IO()
        io.iq = _addr(iq)
        io.iq_is_host, io.stride_samples, io.buf_len = int(iq_is_host), stride, buf_len
        io.iq_format = iq_format
        io.buf_start = bs.ctypes.data
        out = {}
        if "results" in want:
            out["results"] = np.zeros(S, RESULT_DTYPE); io.results = out["results"].ctypes.data
        if "fibs" in want:
            out["fibs"] = np.zeros((S, 12, 32), np.uint8); io.fibs = out["fibs"].ctypes.data
        if msc_stride:
            out["msc"] = np.zeros((S, MAX_SUBCH, 4, msc_stride), np.uint8); io.msc = out["msc"].ctypes.data; io.msc_stride = msc_stride
        if sf_stride:
            out["sf"] = np.zeros((S, MAX_SUBCH, sf_stride), np.uint8); io.sf = out["sf"].ctypes.data; io.sf_stride = sf_stride
        self._ck(self.lib.dabb_process(self.h, C.byref(io)))
        return out

    def submit(self, iq, stride, buf_start, buf_len, msc_stride=0, sf_stride=0, want=("results", "fibs"), iq_format=0, out=None, carry=0):
        """pipelined host-buffer step (dabb_submit): returns immediately; collect() hands back the output dict of the oldest step.
        `out`: a dict returned by an earlier collect() to be reused for this step's results (saves allocating the arrays again).
        `carry`: the first `carry` samples of every window repeat the end of the previous submit's window and are not copied again"""
        S = self.n_streams
        bs = np.ascontiguousarray(buf_start, np.int64)
        io = IO()
        io.iq = _addr(iq)
        io.iq_is_host, io.stride_samples, io.buf_len, io.iq_format = 1, stride, buf_len, iq_format
        io.carry_samples = int(carry)
        io.buf_start = bs.ctypes.data
        if out is not None:
            io.results = out["results"].ctypes.data if "results" in out else None
            io.fibs = out["fibs"].ctypes.data if "fibs" in out else None
            if "msc" in out:
                io.msc = out["msc"].ctypes.data; io.msc_stride = out["msc"].shape[-1]
            if "sf" in out:
                io.sf = out["sf"].ctypes.data; io.sf_stride = out["sf"].shape[-1]
            self._ck(self.lib.dabb_submit(self.h, C.byref(io)))
            self._pending = getattr(self, "_pending", [])
            self._pending.append((out, bs, iq))
            return None
        out = {}
        if "results" in want:
            out["results"] = np.zeros(S, RESULT_DTYPE); io.results = out["results"].ctypes.data
        if "fibs" in want:
            out["fibs"] = np.zeros((S, 12, 32), np.uint8); io.fibs = out["fibs"].ctypes.data
        if msc_stride:
            out["msc"] = np.zeros((S, MAX_SUBCH, 4, msc_stride), np.uint8); io.msc = out["msc"].ctypes.data; io.msc_stride = msc_stride
        if sf_stride:
            out["sf"] = np.zeros((S, MAX_SUBCH, sf_stride), np.uint8); io.sf = out["sf"].ctypes.data; io.sf_stride = sf_stride
        self._ck(self.lib.dabb_submit(self.h, C.byref(io)))
        self._pending = getattr(self, "_pending", [])
        self._pending.append((out, bs, iq))          # keep the buffers alive until collected
        return None

    def collect(self):
        self._ck(self.lib.dabb_collect(self.h))
        out, _, _ = self._pending.pop(0)
        return out

    def process_async(self, iq_ptr, stride, buf_start, buf_len):
        bs = np.ascontiguousarray(buf_start, np.int64)
        io = IO()
        io.iq, io.iq_is_host, io.stride_samples, io.buf_len, io.buf_start = iq_ptr, 0, stride, buf_len, bs.ctypes.data
        self._ck(self.lib.dabb_process_async(self.h, C.byref(io)))

    def get_info(self, what):
        v = C.c_int64()
        self._ck(self.lib.dabb_get_info(self.h, int(what), C.byref(v)))
        return v.value

    def set_options(self, disable_coarse=True, fft_placement=0, freqsync_method=0, decode_tii=False):
        o = Options(); o.disable_coarse, o.fft_placement, o.freqsync_method, o.decode_tii = int(disable_coarse), int(fft_placement), int(freqsync_method), int(decode_tii)
        self._ck(self.lib.dabb_set_options(self.h, C.byref(o)))

    def read_tap(self, what):
        """0 softbits, 1 CIR, 2 constellation points (75 x 16 per stream), 3 null symbol (2656 samples, NCO applied)"""
        if what == 0:
            out = np.zeros((self.n_streams, 75, 3072), np.int8)
        elif what == 2:
            out = np.zeros((self.n_streams, 75, 16), np.complex64)
        elif what == 3:
            out = np.zeros((self.n_streams, 2656), np.complex64)
        elif what == 4:
            out = np.zeros((self.n_streams, 2, 2048), np.complex64)
        else:
            out = np.zeros((self.n_streams, TU), np.float32)
        self._ck(self.lib.dabb_read_tap(self.h, what, _vp(out), C.c_size_t(out.nbytes)))
        return out

    # ---- stage-level (numpy in / numpy out; device staging through the ABI)
    def ofdm_demod(self, iq_frames, prs_start, nco=None, want_r1=False, want_fc=False):
        """iq_frames: [n, stride] complex64; prs_start: [n] sample offsets of the first useful PRS sample"""
        iq = np.ascontiguousarray(iq_frames, np.complex64); n, stride = iq.shape
        d_iq = self.dev(iq); d_ps = self.dev(np.ascontiguousarray(prs_start, np.int64))
        d_soft = self.dev(n * SOFT_PER_FRAME)
        d_r1 = self.dev(n * 75 * K * 8) if want_r1 else None
        d_fc = self.dev(n * 8) if want_fc else None
        d_nco = self.dev(np.ascontiguousarray(nco, np.int32)) if nco is not None else None
        self._ck(self.lib.dabb_ofdm_demod(self.h, C.c_void_p(d_iq.ptr), C.c_int64(stride), C.c_void_p(d_ps.ptr), n,
                                          C.c_void_p(d_nco.ptr) if d_nco else None, C.c_void_p(d_soft.ptr),
                                          C.c_void_p(d_r1.ptr) if d_r1 else None, C.c_void_p(d_fc.ptr) if d_fc else None))
        self.sync()
        out = [d_soft.download(np.int8).reshape(n, 75, 3072)]
        if want_r1:
            out.append(d_r1.download(np.complex64).reshape(n, 75, K))
        if want_fc:
            out.append(d_fc.download(np.complex64))
        for b in (d_iq, d_ps, d_soft, d_r1, d_fc, d_nco):
            if b:
                b.free()
        return out[0] if len(out) == 1 else tuple(out)

    def coarse_estimate(self, iq_frames, prs_start, method=0):
        """OFDMProcessor::processPRS on aligned phase reference symbols; method = FREQSYNC_*"""
        iq = np.ascontiguousarray(iq_frames, np.complex64); n, stride = iq.shape
        d_iq = self.dev(iq); d_ps = self.dev(np.ascontiguousarray(prs_start, np.int64)); d_out = self.dev(4 * n)
        self._ck(self.lib.dabb_coarse_estimate(self.h, C.c_void_p(d_iq.ptr), C.c_int64(stride), C.c_void_p(d_ps.ptr), n, int(method), C.c_void_p(d_out.ptr)))
        self.sync()
        out = d_out.download(np.int32)
        for b in (d_iq, d_ps, d_out):
            b.free()
        return out

    def find_index(self, iq_frames, win_start, want_cir=False, placement=0):
        iq = np.ascontiguousarray(iq_frames, np.complex64); n, stride = iq.shape
        d_iq = self.dev(iq); d_ws = self.dev(np.ascontiguousarray(win_start, np.int64)); d_idx = self.dev(4 * n)
        d_cir = self.dev(4 * n * TU) if want_cir else None
        self._ck(self.lib.dabb_find_index_ex(self.h, C.c_void_p(d_iq.ptr), C.c_int64(stride), C.c_void_p(d_ws.ptr), n, int(placement), C.c_void_p(d_idx.ptr),
                                             C.c_void_p(d_cir.ptr) if d_cir else None))
        self.sync()
        idx = d_idx.download(np.int32)
        cir = d_cir.download(np.float32).reshape(n, TU) if want_cir else None
        for b in (d_iq, d_ws, d_idx, d_cir):
            if b:
                b.free()
        return (idx, cir) if want_cir else idx

    def viterbi(self, soft, nbits):
        soft = np.ascontiguousarray(soft, np.int8); n = soft.shape[0]
        assert soft.shape[1] == (nbits + 6) * 4
        d_s = self.dev(soft); d_o = self.dev(n * nbits)
        self._ck(self.lib.dabb_viterbi(self.h, C.c_void_p(d_s.ptr), n, nbits, C.c_void_p(d_o.ptr)))
        out = d_o.download(np.uint8).reshape(n, nbits)
        d_s.free(); d_o.free()
        return out

    def fic_decode(self, soft):
        soft = np.ascontiguousarray(soft, np.int8).reshape(-1, 9216); n = soft.shape[0]
        d_s = self.dev(soft); d_f = self.dev(n * 12 * 32); d_c = self.dev(4 * n)
        self._ck(self.lib.dabb_fic_decode(self.h, C.c_void_p(d_s.ptr), n, C.c_void_p(d_f.ptr), C.c_void_p(d_c.ptr)))
        fibs = d_f.download(np.uint8).reshape(n, 12, 32); crc = d_c.download(np.int32)
        for b in (d_s, d_f, d_c):
            b.free()
        return fibs, crc

    def msc_decode(self, soft, length_cu, bitrate, eep_profile_a=True, eep_level=3, short_form=False, uep_level=0):
        soft = np.ascontiguousarray(soft, np.int8).reshape(-1, length_cu * 64); n = soft.shape[0]
        sc = Subchannel(0, 0, length_cu, bitrate, int(short_form), uep_level, int(eep_profile_a), eep_level, 1)
        d_s = self.dev(soft); d_o = self.dev(n * 3 * bitrate)
        self._ck(self.lib.dabb_msc_decode(self.h, C.byref(sc), C.c_void_p(d_s.ptr), n, C.c_void_p(d_o.ptr)))
        out = d_o.download(np.uint8).reshape(n, 3 * bitrate)
        d_s.free(); d_o.free()
        return out

    def rs_superframes(self, sfs):
        sfs = np.ascontiguousarray(sfs, np.uint8); n, sf_len = sfs.shape
        d_s = self.dev(sfs); d_i = self.dev(16 * n)
        self._ck(self.lib.dabb_rs_superframes(self.h, C.c_void_p(d_s.ptr), n, sf_len, C.c_void_p(d_i.ptr)))
        self.sync()
        out = d_s.download(np.uint8).reshape(n, sf_len); info = d_i.download(np.int32).reshape(n, 4)
        d_s.free(); d_i.free()
        return out, info
