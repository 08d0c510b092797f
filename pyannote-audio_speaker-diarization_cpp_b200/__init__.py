"""ctypes binding of libsdb200.so -- the B200 (sm_100a) hot path of the C++ pyannote diarization pipeline.

The product is the C-ABI library (include/sdb200.h) plus the C++ host shim (host/sdb200_host.hpp) that keeps the
reference's function-level API.  This module is the thin Python face used by tests/ and bench.py; method names
follow the reference functions they stand in for.  There is no CPU fallback: if the library or a CUDA device is
missing, construction fails loudly.

The directory name contains '-', so import it by path (see __graft_entry__.load_package()).
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SDB200_LIB") or os.path.join(HERE, "libsdb200.so")  # SDB200_LIB: an experiment build

c_fp = C.POINTER(C.c_float)
c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int32)
c_lp = C.POINTER(C.c_int64)
c_bp = C.POINTER(C.c_uint8)

SD_OK, SD_ERR_INVALID, SD_ERR_CUDA, SD_ERR_ZERO_MAGNITUDE, SD_ERR_UNSUPPORTED, SD_ERR_NOMEM, SD_ERR_CAPACITY = range(7)

SD_OPT_FORCE_EXACT_LINKAGE = 1

# constants of the reference pipeline (speakerDiarizer.cpp:1335-1340, 2049-2050, 2429-2432)
FRAME_STEP = 0.016875
FRAME_DURATION = 0.016875
ONSET = 0.4442333667381752
EPS = float(np.finfo(np.float64).eps)


class Window(C.Structure):
    """SlidingWindow POD (speakerDiarizer.cpp:1029-1036)."""
    _fields_ = [("start", C.c_double), ("step", C.c_double), ("duration", C.c_double), ("num_samples", C.c_int64)]

    def astuple(self):
        return (self.start, self.step, self.duration, self.num_samples)


class StftParams(C.Structure):
    _fields_ = [("n_fft", C.c_int), ("hop", C.c_int), ("window_kind", C.c_int), ("window", c_fp),
                ("preemph", C.c_float), ("pad_batch_to", C.c_int), ("frame_mode", C.c_int),
                ("remove_dc_offset", C.c_int)]


class FbankParams(C.Structure):
    _fields_ = [("stft", StftParams), ("n_mels", C.c_int), ("f_min", C.c_float), ("f_max", C.c_float),
                ("sample_rate", C.c_int), ("top_db", C.c_float), ("amin", C.c_float), ("mean_norm", C.c_int),
                ("mel_kind", C.c_int), ("log_kind", C.c_int)]


class ClusterParams(C.Structure):
    _fields_ = [("threshold", C.c_float), ("min_cluster_size", C.c_int), ("num_clusters", C.c_int),
                ("min_clusters", C.c_int), ("max_clusters", C.c_int), ("pdist_mode", C.c_int)]


class SdFile(C.Structure):
    """sd_file of include/sdb200.h: one file of a batch (geometry, input / output pointers, results)."""
    _fields_ = [("C", C.c_int), ("F", C.c_int), ("S", C.c_int), ("L", C.c_int), ("D", C.c_int), ("Kd", C.c_int),
                ("onset", C.c_double), ("chunks", Window), ("frames", Window),
                ("wav_items", C.c_void_p), ("segmentations", C.c_void_p), ("embeddings", C.c_void_p),
                ("diar_scores", C.c_void_p),
                ("stft", C.c_void_p), ("binarized", C.c_void_p), ("count", C.c_void_p), ("count_cap", C.c_int64),
                ("hard", C.c_void_p), ("diar", C.c_void_p),
                ("n_count", C.c_int64), ("n_diar", C.c_int64), ("num_clusters", C.c_int), ("status", C.c_int),
                ("count_frames", Window)]


SD_BATCH_HOST, SD_BATCH_DEVICE = 0, 1


class SdError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sdb200 error %d: %s" % (code, msg))
        self.code = code


EXPORTS = [
    "sd_version", "sd_ctx_create", "sd_ctx_destroy", "sd_last_error", "sd_ctx_set_stream", "sd_ctx_stream", "sd_sync",
    "sd_malloc", "sd_free", "sd_host_alloc", "sd_host_free", "sd_memcpy_h2d", "sd_memcpy_d2h", "sd_memset",
    "sd_timer_start", "sd_timer_stop", "sd_timer_elapsed_ms", "sd_launch_count", "sd_ctx_set_option", "sd_debug_counters", "sd_flush_l2",
    "sd_stft_default_params", "sd_stft_num_frames", "sd_stft", "sd_stft_dev", "sd_pack_wav_lens",
    "sd_fbank_default_params", "sd_fbank", "sd_fbank_dev", "sd_np_rint", "sd_closest_frame", "sd_aggregate_num_frames",
    "sd_aggregate", "sd_aggregate_dev", "sd_binarize", "sd_binarize_dev", "sd_binarize_rows", "sd_trim_num_frames",
    "sd_trim", "sd_speaker_count", "sd_speaker_count_dev", "sd_clean_segmentations", "sd_normalize", "sd_pdist",
    "sd_linkage", "sd_linkage_dev", "sd_fcluster", "sd_cluster", "sd_cosine_cdist", "sd_cluster_default_params",
    "sd_cluster_labels", "sd_clustering", "sd_clustering_dev", "sd_mask_compact", "sd_select_masks_dev",
    "sd_mask_compact_file_dev", "sd_reconstruct_rows", "sd_reconstruct", "sd_reconstruct_dev", "sd_to_annotation",
    "sd_to_annotation_dev", "sd_ingest_pcm16", "sd_ingest_pcm16_dev", "sd_slide_geometry", "sd_crop_chunks",
    "sd_crop_chunks_dev", "sd_clustering_async_dev", "sd_status_reset", "sd_status_check",
    "sd_clustering_ex", "sd_binarize_rows_stages", "sd_trim_sum", "sd_mask_interpolate", "sd_clustered_segmentations",
    "sd_to_diarization", "sd_stft_kaldi_params", "sd_stft_num_frames_mode", "sd_fbank_kaldi_params",
    "sd_batch_create", "sd_batch_create_ex", "sd_batch_get_config", "sd_batch_destroy", "sd_batch_set_params", "sd_batch_workers", "sd_batch_stream",
    "sd_batch_submit", "sd_batch_wait", "sd_batch_last_error", "sd_batch_launch_count", "sd_linkage_stage_ms",
]

_lib = None


def lib():
    """Load libsdb200.so (built by __graft_entry__.build() / `make` in this directory)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("libsdb200.so is missing (%s): build it with `make -C %s`; there is no fallback path"
                           % (LIB_PATH, HERE))
    L = C.CDLL(LIB_PATH)
    vp, i, d, i64, sz = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_size_t
    W = C.POINTER(Window)
    sig = {
        "sd_version": (i, []),
        "sd_ctx_create": (i, [i, C.POINTER(vp)]),
        "sd_ctx_destroy": (None, [vp]),
        "sd_last_error": (C.c_char_p, [vp]),
        "sd_ctx_set_stream": (i, [vp, vp]),
        "sd_ctx_stream": (vp, [vp]),
        "sd_sync": (i, [vp]),
        "sd_malloc": (i, [vp, sz, C.POINTER(vp)]),
        "sd_free": (i, [vp, vp]),
        "sd_host_alloc": (i, [vp, sz, C.POINTER(vp)]),
        "sd_host_free": (i, [vp, vp]),
        "sd_memcpy_h2d": (i, [vp, vp, vp, sz]),
        "sd_memcpy_d2h": (i, [vp, vp, vp, sz]),
        "sd_memset": (i, [vp, vp, i, sz]),
        "sd_timer_start": (i, [vp, i]),
        "sd_timer_stop": (i, [vp, i]),
        "sd_timer_elapsed_ms": (i, [vp, i, c_fp]),
        "sd_launch_count": (i64, [vp]),
        "sd_flush_l2": (i, [vp]),
        "sd_debug_counters": (i, [vp, c_lp, i]),
        "sd_ctx_set_option": (i, [vp, i, i]),
        "sd_stft_default_params": (None, [C.POINTER(StftParams)]),
        "sd_stft_num_frames": (i64, [i, i]),
        "sd_stft": (i, [vp, vp, i, i, C.POINTER(StftParams), vp]),
        "sd_stft_dev": (i, [vp, vp, i, i, C.POINTER(StftParams), vp]),
        "sd_pack_wav_lens": (i, [c_fp, i, i, c_fp]),
        "sd_fbank_default_params": (None, [C.POINTER(FbankParams)]),
        "sd_fbank": (i, [vp, vp, i, i, vp, C.POINTER(FbankParams), vp]),
        "sd_fbank_dev": (i, [vp, vp, i, i, vp, C.POINTER(FbankParams), vp]),
        "sd_np_rint": (i, [d]),
        "sd_closest_frame": (i64, [W, d]),
        "sd_aggregate_num_frames": (i64, [i, W, W]),
        "sd_aggregate": (i, [vp, vp, i, i, i, W, W, i, d, i, d, vp, i64, c_lp, W, vp, vp]),
        "sd_aggregate_dev": (i, [vp, vp, i, i, i, W, W, i, d, i, d, vp, i64, c_lp, W, vp, vp]),
        "sd_binarize": (i, [vp, vp, i, i, i, d, i, vp]),
        "sd_binarize_dev": (i, [vp, vp, i, i, i, d, i, vp]),
        "sd_binarize_rows": (i, [vp, vp, i, i, d, i, vp]),
        "sd_trim_num_frames": (i64, [i, d, d]),
        "sd_trim": (i, [vp, vp, i, i, i, d, d, W, vp, W]),
        "sd_speaker_count": (i, [vp, vp, i, i, i, W, W, vp, i64, c_lp, W]),
        "sd_speaker_count_dev": (i, [vp, vp, i, i, i, W, W, vp, i64, c_lp, W]),
        "sd_clean_segmentations": (i, [vp, vp, i, i, i, vp]),
        "sd_normalize": (i, [vp, vp, i, i]),
        "sd_pdist": (i, [vp, vp, i, i, i, vp]),
        "sd_linkage": (i, [vp, vp, i, i, vp]),
        "sd_linkage_dev": (i, [vp, vp, i, i, vp]),
        "sd_fcluster": (i, [vp, vp, i, d, vp]),
        "sd_cluster": (i, [vp, vp, i, i, d, vp]),
        "sd_cosine_cdist": (i, [vp, vp, i, vp, i, i, vp]),
        "sd_cluster_default_params": (None, [C.POINTER(ClusterParams)]),
        "sd_cluster_labels": (i, [vp, vp, i, i, C.POINTER(ClusterParams), vp]),
        "sd_clustering": (i, [vp, vp, i, i, i, C.POINTER(ClusterParams), vp, i, vp, vp, i, c_ip]),
        "sd_clustering_dev": (i, [vp, vp, i, i, i, C.POINTER(ClusterParams), vp, i, vp, vp, i, c_ip]),
        "sd_mask_compact": (i, [vp, vp, vp, i, i, i, i, vp, vp, vp, c_ip]),
        "sd_select_masks_dev": (i, [vp, vp, i, i, i, d, vp]),
        "sd_mask_compact_file_dev": (i, [vp, vp, i64, i, i, i, i, vp, i, i, i, vp, vp, vp, vp]),
        "sd_reconstruct_rows": (i, [i, W, i64, W, c_lp, W]),
        "sd_reconstruct": (i, [vp, vp, i, i, i, W, vp, vp, i64, W, vp, i64, c_lp, c_ip, W]),
        "sd_reconstruct_dev": (i, [vp, vp, i, i, i, W, vp, i, vp, i64, W, vp, i64, c_lp, W]),
        "sd_to_annotation": (i, [vp, vp, i64, i, W, d, d, d, d, vp, vp, i64, c_lp]),
        "sd_to_annotation_dev": (i, [vp, vp, i64, i, W, d, d, d, d, vp, vp, i64, c_lp]),
        "sd_ingest_pcm16": (i, [vp, vp, i64, vp]),
        "sd_ingest_pcm16_dev": (i, [vp, vp, i64, vp]),
        "sd_slide_geometry": (i, [i64, d, d, c_lp, c_lp, c_lp]),
        "sd_crop_chunks": (i, [vp, vp, i64, vp, i, d, i, vp]),
        "sd_crop_chunks_dev": (i, [vp, vp, i64, vp, i, d, i, vp]),
        "sd_clustering_async_dev": (i, [vp, vp, i, i, i, C.POINTER(ClusterParams), vp, i, vp, i, vp, vp, i, vp]),
        "sd_status_reset": (i, [vp]),
        "sd_status_check": (i, [vp]),
        "sd_stft_kaldi_params": (None, [C.POINTER(StftParams), i]),
        "sd_stft_num_frames_mode": (i64, [i, i, i, i]),
        "sd_fbank_kaldi_params": (None, [C.POINTER(FbankParams), i]),
        "sd_clustering_ex": (i, [vp, vp, i, i, i, C.POINTER(ClusterParams), vp, i, vp, vp, vp, i, c_ip]),
        "sd_binarize_rows_stages": (i, [vp, vp, i, i, d, vp, vp, vp, c_ip]),
        "sd_trim_sum": (i, [vp, vp, i, i, i, d, d, vp]),
        "sd_mask_interpolate": (i, [vp, vp, i, i, i, C.c_float, vp, vp]),
        "sd_clustered_segmentations": (i, [vp, vp, i, i, i, vp, i, vp]),
        "sd_to_diarization": (i, [vp, vp, i64, i, W, vp, i64, W, vp, i64, c_lp, W, vp, c_lp]),
        "sd_batch_create": (i, [i, i, C.POINTER(vp)]),
        "sd_batch_create_ex": (i, [i, i, C.POINTER(SdBatchConfig), C.POINTER(vp)]),
        "sd_batch_get_config": (i, [vp, C.POINTER(SdBatchConfig)]),
        "sd_batch_destroy": (None, [vp]),
        "sd_batch_set_params": (i, [vp, C.POINTER(StftParams), C.POINTER(ClusterParams)]),
        "sd_batch_workers": (i, [vp]),
        "sd_batch_stream": (vp, [vp, i]),
        "sd_batch_submit": (i, [vp, C.POINTER(SdFile), i, i]),
        "sd_batch_wait": (i, [vp]),
        "sd_batch_last_error": (C.c_char_p, [vp]),
        "sd_batch_launch_count": (i64, [vp]),
        "sd_linkage_stage_ms": (i, [vp, c_fp, c_fp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _win(w):
    if isinstance(w, Window):
        return w
    return Window(float(w[0]), float(w[1]), float(w[2]), int(w[3]) if len(w) > 3 else 0)


FRAMES = (0.0, FRAME_STEP, FRAME_DURATION, 0)


class SdBatchConfig(C.Structure):
    _fields_ = [("stft_chain", C.c_int), ("linkage_cluster", C.c_int), ("narrow_sms", C.c_int)]


class Batch:
    """sd_batch: `workers` files in flight on one GPU, driven by library-owned host threads."""

    def __init__(self, device=0, workers=8, stft_chain=-1, linkage_cluster=-1, narrow_sms=-1):
        """-1 = automatic (sd_batch_config, include/sdb200.h)."""
        self.L = lib()
        h = C.c_void_p()
        cfg = SdBatchConfig(stft_chain, linkage_cluster, narrow_sms)
        rc = self.L.sd_batch_create_ex(device, workers, C.byref(cfg), C.byref(h))
        if rc != SD_OK:
            raise SdError(rc, "sd_batch_create(device=%d) failed: no usable CUDA device (no CPU fallback exists)" % device)
        self.h = h
        self.workers = workers

    def config(self):
        """The configuration in effect (narrow_sms as granted by the driver, 0 = no partition)."""
        cfg = SdBatchConfig()
        self.L.sd_batch_get_config(self.h, C.byref(cfg))
        return dict(stft_chain=cfg.stft_chain, linkage_cluster=cfg.linkage_cluster, narrow_sms=cfg.narrow_sms)

    def close(self):
        if self.h:
            self.L.sd_batch_destroy(self.h)
            self.h = None

    def submit(self, files, pointers=SD_BATCH_HOST):
        """files: a ctypes array of SdFile (kept alive by the caller until wait() returns)."""
        rc = self.L.sd_batch_submit(self.h, files, len(files), pointers)
        if rc:
            raise SdError(rc, "sd_batch_submit: invalid arguments")

    def wait(self):
        rc = self.L.sd_batch_wait(self.h)
        if rc:
            raise SdError(rc, (self.L.sd_batch_last_error(self.h) or b"").decode())

    def run(self, files, pointers=SD_BATCH_HOST):
        self.submit(files, pointers)
        self.wait()

    def stream(self, worker):
        return self.L.sd_batch_stream(self.h, worker)

    def launch_count(self):
        return int(self.L.sd_batch_launch_count(self.h))


def make_file(geo_C, F, S, L, D, chunks, frames=None, onset=None, Kd=0, **ptrs):
    """Fill an SdFile; pointer arguments are numpy arrays (host mode) or integer device addresses."""
    f = SdFile()
    f.C, f.F, f.S, f.L, f.D, f.Kd = geo_C, F, S, L, D, Kd
    f.onset = ONSET if onset is None else onset
    f.chunks = _win(chunks)
    f.frames = _win(frames if frames is not None else FRAMES)
    for k, v in ptrs.items():
        if k == "count_cap":
            f.count_cap = int(v)
        elif v is None:
            setattr(f, k, None)
        elif isinstance(v, np.ndarray):
            setattr(f, k, v.ctypes.data)
        else:
            setattr(f, k, int(v))
    return f


class Context:
    """One sd_ctx (one GPU, one stream).  Host-array methods mirror the reference's function names."""

    def __init__(self, device=0):
        self.L = lib()
        h = C.c_void_p()
        rc = self.L.sd_ctx_create(device, C.byref(h))
        if rc != SD_OK:
            raise SdError(rc, "sd_ctx_create(device=%d) failed: no usable CUDA device (no CPU fallback exists)" % device)
        self.h = h

    def close(self):
        if getattr(self, "h", None):
            self.L.sd_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != SD_OK:
            raise SdError(rc, self.L.sd_last_error(self.h).decode())

    # ---- plumbing
    def sync(self):
        self._check(self.L.sd_sync(self.h))

    def set_stream(self, cuda_stream_ptr):
        self._check(self.L.sd_ctx_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def malloc(self, nbytes):
        p = C.c_void_p()
        self._check(self.L.sd_malloc(self.h, nbytes, C.byref(p)))
        return p.value

    def free(self, dptr):
        self._check(self.L.sd_free(self.h, C.c_void_p(dptr)))

    def host_alloc(self, shape, dtype):
        """Pinned host array."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape)) * dtype.itemsize
        p = C.c_void_p()
        self._check(self.L.sd_host_alloc(self.h, n, C.byref(p)))
        buf = (C.c_char * n).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype).reshape(shape)
        return arr

    def h2d(self, dptr, arr):
        arr = np.ascontiguousarray(arr)
        self._check(self.L.sd_memcpy_h2d(self.h, C.c_void_p(dptr), _ptr(arr), arr.nbytes))
        return arr  # keep alive until sync

    def d2h(self, arr, dptr):
        self._check(self.L.sd_memcpy_d2h(self.h, _ptr(arr), C.c_void_p(dptr), arr.nbytes))

    def to_device(self, arr):
        arr = np.ascontiguousarray(arr)
        d = self.malloc(arr.nbytes)
        self.h2d(d, arr)
        self.sync()
        return d

    def timer_start(self, slot=0):
        self._check(self.L.sd_timer_start(self.h, slot))

    def timer_stop(self, slot=0):
        self._check(self.L.sd_timer_stop(self.h, slot))

    def timer_ms(self, slot=0):
        ms = C.c_float()
        self._check(self.L.sd_timer_elapsed_ms(self.h, slot, C.byref(ms)))
        return ms.value

    def launch_count(self):
        return self.L.sd_launch_count(self.h)

    def set_option(self, option, value):
        self._check(self.L.sd_ctx_set_option(self.h, int(option), int(value)))

    def debug_counters(self, reset=True):
        out = np.zeros(8, np.int64)
        self._check(self.L.sd_debug_counters(self.h, out.ctypes.data_as(c_lp), int(reset)))
        return out

    def flush_l2(self):
        self._check(self.L.sd_flush_l2(self.h))

    # ---- a1/a2
    def stft_params(self, pad_batch_to=0, window=None, window_kind=0):
        p = StftParams()
        self.L.sd_stft_default_params(C.byref(p))
        p.pad_batch_to = pad_batch_to
        p.window_kind = window_kind
        if window is not None:
            self._win_keep = np.ascontiguousarray(window, np.float32)
            p.window_kind = 2
            p.window = self._win_keep.ctypes.data_as(c_fp)
        return p

    def linkage_stage_ms(self):
        """(pdist ms, merge-loop ms) of the last linkage on this context."""
        a, b = C.c_float(), C.c_float()
        self._check(self.L.sd_linkage_stage_ms(self.h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def stft_kaldi_params(self, snip_edges=False, preemph=0.97, remove_dc_offset=True):
        """Kaldi framing: povey window, per-frame pre-emphasis / DC removal, snip_edges as given."""
        p = StftParams()
        self.L.sd_stft_kaldi_params(C.byref(p), int(snip_edges))
        p.preemph = preemph
        p.remove_dc_offset = int(remove_dc_offset)
        return p

    def stft(self, wav, pad_batch_to=0, window=None, window_kind=0, params=None):
        """EmbeddingModel1::infer up to the ORT input: [max(B,pad)][T][201][2] fp32."""
        wav = np.ascontiguousarray(wav, np.float32)
        B, Ls = wav.shape
        p = params or self.stft_params(pad_batch_to, window, window_kind)
        T = self.L.sd_stft_num_frames_mode(Ls, p.n_fft, p.hop, p.frame_mode)
        out = np.empty((max(B, pad_batch_to), T, p.n_fft // 2 + 1, 2), np.float32)
        self._check(self.L.sd_stft(self.h, _ptr(wav), B, Ls, C.byref(p), _ptr(out)))
        return out

    def stft_dev(self, d_wav, B, Ls, d_out, p=None):
        if p is None:
            p = self.stft_params()
        self._check(self.L.sd_stft_dev(self.h, C.c_void_p(d_wav), B, Ls, C.byref(p), C.c_void_p(d_out)))

    def pack_wav_lens(self, lens, batch=32):
        lens = np.ascontiguousarray(lens, np.float32)
        out = np.empty(batch, np.float32)
        rc = self.L.sd_pack_wav_lens(lens.ctypes.data_as(c_fp), lens.shape[0], batch, out.ctypes.data_as(c_fp))
        if rc:
            raise SdError(rc, "sd_pack_wav_lens")
        return out

    def fbank_params(self):
        p = FbankParams()
        self.L.sd_fbank_default_params(C.byref(p))
        return p

    def fbank_kaldi_params(self, snip_edges=False, preemph=0.97, remove_dc_offset=True, n_mels=80):
        """torchaudio.compliance.kaldi.fbank defaults on a 400-point transform (round_to_power_of_two=False)."""
        p = FbankParams()
        self.L.sd_fbank_kaldi_params(C.byref(p), int(snip_edges))
        p.stft.preemph = preemph
        p.stft.remove_dc_offset = int(remove_dc_offset)
        p.n_mels = n_mels
        return p

    def fbank(self, wav, wav_lens=None, params=None):
        wav = np.ascontiguousarray(wav, np.float32)
        B, Ls = wav.shape
        wav_lens = np.ascontiguousarray(np.ones(B) if wav_lens is None else wav_lens, np.float32)
        p = params or self.fbank_params()
        T = self.L.sd_stft_num_frames_mode(Ls, p.stft.n_fft, p.stft.hop, p.stft.frame_mode)
        out = np.empty((B, T, p.n_mels), np.float32)
        self._check(self.L.sd_fbank(self.h, _ptr(wav), B, Ls, _ptr(wav_lens), C.byref(p), _ptr(out)))
        return out

    # ---- a4/a5
    def np_rint(self, v):
        return self.L.sd_np_rint(float(v))

    def closest_frame(self, t, window=FRAMES):
        w = _win(window)
        return self.L.sd_closest_frame(C.byref(w), float(t))

    def aggregate(self, scores, chunks, frames=FRAMES, hamming=False, missing=np.nan, skip_average=False, epsilon=EPS,
                  want_aux=False):
        """PipelineHelper::aggregate.  chunks = (start, step, duration, num_samples)."""
        scores = np.ascontiguousarray(scores, np.float64)
        Cn, F, K = scores.shape
        cw, fw = _win(chunks), _win(frames)
        NF = self.L.sd_aggregate_num_frames(Cn, C.byref(cw), C.byref(fw))
        out = np.empty((NF, K), np.float64)
        cnt = np.empty((NF, K), np.float64) if want_aux else None
        msk = np.empty((NF, K), np.float64) if want_aux else None
        n = C.c_int64()
        post = Window()
        self._check(self.L.sd_aggregate(self.h, _ptr(scores), Cn, F, K, C.byref(cw), C.byref(fw), int(hamming),
                                        float(missing), int(skip_average), float(epsilon), _ptr(out), NF, C.byref(n),
                                        C.byref(post), _ptr(cnt), _ptr(msk)))
        assert n.value == NF
        if want_aux:
            return out, post, cnt, msk
        return out, post

    # ---- a6/a7
    def binarize_swf(self, scores, onset=ONSET, initial_state=False):
        scores = np.ascontiguousarray(scores, np.float32)
        Cn, F, K = scores.shape
        out = np.empty((Cn, F, K), np.float64)
        self._check(self.L.sd_binarize(self.h, _ptr(scores), Cn, F, K, float(onset), int(initial_state), _ptr(out)))
        return out

    def binarize_ndarray(self, scores, onset=0.5, initial_state=False):
        scores = np.ascontiguousarray(scores, np.float64)
        R, F = scores.shape
        out = np.empty((R, F), np.uint8)
        self._check(self.L.sd_binarize_rows(self.h, _ptr(scores), R, F, float(onset), int(initial_state), _ptr(out)))
        return out

    def trim(self, binarized, left=0.1, right=0.1, before=(0.0, 0.5, 5.0, 0)):
        b = np.ascontiguousarray(binarized, np.float64)
        Cn, F, K = b.shape
        Ft = self.L.sd_trim_num_frames(F, left, right)
        out = np.empty((Cn, Ft, K), np.float64)
        bw = _win(before)
        tw = Window()
        self._check(self.L.sd_trim(self.h, _ptr(b), Cn, F, K, left, right, C.byref(bw), _ptr(out), C.byref(tw)))
        return out, tw

    def speaker_count(self, binarized, chunks=(0.0, 0.5, 5.0, 1), frames=FRAMES):
        b = np.ascontiguousarray(binarized, np.float64)
        Cn, F, K = b.shape
        cw, fw = _win(chunks), _win(frames)
        cap = int((Cn * cw.step + cw.duration) / fw.step) + F + 64
        out = np.empty(cap, np.int32)
        n = C.c_int64()
        cf = Window()
        self._check(self.L.sd_speaker_count(self.h, _ptr(b), Cn, F, K, C.byref(cw), C.byref(fw), _ptr(out), cap,
                                            C.byref(n), C.byref(cf)))
        return out[:n.value].copy(), cf

    def clean_segmentations(self, binarized):
        b = np.ascontiguousarray(binarized, np.float64)
        out = np.empty_like(b)
        self._check(self.L.sd_clean_segmentations(self.h, _ptr(b), b.shape[0], b.shape[1], b.shape[2], _ptr(out)))
        return out

    # ---- a9-a12
    def normalize_embeddings(self, x):
        x = np.ascontiguousarray(x, np.float64).copy()
        self._check(self.L.sd_normalize(self.h, _ptr(x), x.shape[0], x.shape[1]))
        return x

    def pdist(self, x, mode=0):
        x = np.ascontiguousarray(x, np.float64)
        N = x.shape[0]
        out = np.empty(N * (N - 1) // 2, np.float64)
        self._check(self.L.sd_pdist(self.h, _ptr(x), N, x.shape[1], mode, _ptr(out)))
        return out

    def linkage(self, x):
        x = np.ascontiguousarray(x, np.float64)
        Z = np.empty((x.shape[0] - 1, 4), np.float64)
        self._check(self.L.sd_linkage(self.h, _ptr(x), x.shape[0], x.shape[1], _ptr(Z)))
        return Z

    def fcluster(self, Z, cutoff):
        Z = np.ascontiguousarray(Z, np.float64)
        N = Z.shape[0] + 1
        T = np.empty(N, np.int32)
        self._check(self.L.sd_fcluster(self.h, _ptr(Z), N, float(cutoff), _ptr(T)))
        return T

    def cluster(self, x, cutoff):
        """Clustering::cluster (linkage + fcluster)."""
        x = np.ascontiguousarray(x, np.float64)
        T = np.empty(x.shape[0], np.int32)
        self._check(self.L.sd_cluster(self.h, _ptr(x), x.shape[0], x.shape[1], float(cutoff), _ptr(T)))
        return T

    def cosine_cdist(self, a, b):
        a = np.ascontiguousarray(a, np.float64)
        b = np.ascontiguousarray(b, np.float64)
        out = np.empty((a.shape[0], b.shape[0]), np.float64)
        self._check(self.L.sd_cosine_cdist(self.h, _ptr(a), a.shape[0], _ptr(b), b.shape[0], a.shape[1], _ptr(out)))
        return out

    # ---- a8/a13-a15
    def cluster_params(self, **kw):
        p = ClusterParams()
        self.L.sd_cluster_default_params(C.byref(p))
        for k, v in kw.items():
            setattr(p, k, v)
        return p

    def cluster_labels(self, x, params=None):
        """Cluster::cluster on filtered embeddings."""
        x = np.ascontiguousarray(x, np.float64)
        p = params or self.cluster_params()
        lab = np.empty(x.shape[0], np.int32)
        self._check(self.L.sd_cluster_labels(self.h, _ptr(x), x.shape[0], x.shape[1], C.byref(p), _ptr(lab)))
        return lab

    def clustering(self, embeddings, binarized=None, params=None, soft_k_cap=0):
        """Cluster::clustering (+ inactive-speaker mask when `binarized` is given) -> hard[C][S]."""
        e = np.ascontiguousarray(embeddings, np.float64)
        Cn, S, D = e.shape
        p = params or self.cluster_params()
        hard = np.empty((Cn, S), np.int32)
        soft = np.empty((Cn, S, soft_k_cap), np.float64) if soft_k_cap else None
        F = 0
        if binarized is not None:
            binarized = np.ascontiguousarray(binarized, np.float64)
            F = binarized.shape[1]
        k = C.c_int(0)
        self._check(self.L.sd_clustering(self.h, _ptr(e), Cn, S, D, C.byref(p), _ptr(binarized), F, _ptr(hard),
                                         _ptr(soft), soft_k_cap, C.byref(k)))
        if soft_k_cap:
            return hard, soft, k.value
        return hard, k.value

    # ---- next rows
    def mask_compact(self, wav, masks, min_num_samples=640):
        """Helper::interpolate + padSequence + wav_lens logic of getEmbedding for one batch."""
        wav = np.ascontiguousarray(wav, np.float32)
        masks = np.ascontiguousarray(masks, np.float32)
        B, Ls = wav.shape
        sig = np.empty((B, Ls), np.float32)
        lens = np.empty(B, np.float32)
        ts = np.zeros(B, np.uint8)
        inv = C.c_int(0)
        self._check(self.L.sd_mask_compact(self.h, _ptr(wav), _ptr(masks), B, Ls, masks.shape[1], min_num_samples,
                                           _ptr(sig), _ptr(lens), _ptr(ts), C.byref(inv)))
        return inv.value, sig, lens, ts

    def select_masks(self, binarized, min_num_frames):
        """(chunk, speaker) mask choice of speakerDiarization() (SD:3056-3078): [C][F][K] fp64 -> [C*K][F] fp32."""
        b = np.ascontiguousarray(binarized, np.float64)
        Cn, F, K = b.shape
        d_b = self.to_device(b)
        d_m = self.malloc(4 * Cn * K * F)
        try:
            self._check(self.L.sd_select_masks_dev(self.h, d_b, Cn, F, K, float(min_num_frames), d_m))
            out = np.empty((Cn * K, F), np.float32)
            self.d2h(out, d_m)
        finally:
            self.free(d_b)
            self.free(d_m)
        return out

    def mask_compact_file(self, wave, masks, C_, K, Ls, step_samples, batch=32, min_num_samples=640):
        """Whole-file masking: chunk c = wave[c*step : c*step + Ls] (zero padded); masks [C*K][F]."""
        wave = np.ascontiguousarray(wave, np.float32)
        masks = np.ascontiguousarray(masks, np.float32)
        R, F = masks.shape
        assert R == C_ * K
        ng = (R + batch - 1) // batch
        d_w, d_m = self.to_device(wave), self.to_device(masks)
        d_s, d_l, d_t, d_i = self.malloc(4 * R * Ls), self.malloc(4 * R), self.malloc(R), self.malloc(ng)
        try:
            self._check(self.L.sd_mask_compact_file_dev(self.h, d_w, wave.size, C_, K, Ls, step_samples, d_m, F, batch,
                                                        min_num_samples, d_s, d_l, d_t, d_i))
            sig, lens = np.empty((R, Ls), np.float32), np.empty(R, np.float32)
            ts, inv = np.empty(R, np.uint8), np.empty(ng, np.uint8)
            self.d2h(sig, d_s); self.d2h(lens, d_l); self.d2h(ts, d_t); self.d2h(inv, d_i)
        finally:
            for p in (d_w, d_m, d_s, d_l, d_t, d_i):
                self.free(p)
        return sig, lens, ts, inv

    def reconstruct(self, segmentations, chunks, hard, count, count_frames):
        """reconstruct + to_diarization: -> (discrete diarization [rows][cols] fp64, Window of its rows)."""
        seg = np.ascontiguousarray(segmentations, np.float32)
        hard = np.ascontiguousarray(hard, np.int32)
        count = np.ascontiguousarray(count, np.int32)
        Cn, F, K = seg.shape
        cw, cf = _win(chunks), _win(count_frames)
        rows = C.c_int64()
        self._check(self.L.sd_reconstruct_rows(Cn, C.byref(cw), count.shape[0], C.byref(cf), C.byref(rows), None))
        kc = max(int(hard.max()), 0) + 1
        out = np.empty((max(rows.value, 0), kc), np.float64)
        cols = C.c_int(0)
        fr = Window()
        self._check(self.L.sd_reconstruct(self.h, _ptr(seg), Cn, F, K, C.byref(cw), _ptr(hard), _ptr(count),
                                          count.shape[0], C.byref(cf), _ptr(out), out.size, C.byref(rows),
                                          C.byref(cols), C.byref(fr)))
        assert cols.value == kc and rows.value == out.shape[0]
        return out, fr

    def to_annotation(self, scores, frames, onset=0.5, offset=0.5, min_duration_on=0.0,
                      min_duration_off=float(np.float32(0.5817029604921046)), cap=None):
        s = np.ascontiguousarray(scores, np.float64)
        rows, cols = s.shape
        cap = (rows // 2 + 2) * cols if cap is None else cap
        seg = np.empty((max(cap, 1), 2), np.float64)
        lab = np.empty(max(cap, 1), np.int32)
        n = C.c_int64()
        fw = _win(frames)
        self._check(self.L.sd_to_annotation(self.h, _ptr(s), rows, cols, C.byref(fw), onset, offset, min_duration_on,
                                            min_duration_off, _ptr(seg), _ptr(lab), cap, C.byref(n)))
        return seg[:n.value].copy(), lab[:n.value].copy()

    def ingest_pcm16(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        out = np.empty(pcm.shape, np.float32)
        self._check(self.L.sd_ingest_pcm16(self.h, _ptr(pcm), pcm.size, _ptr(out)))
        return out

    def slide_geometry(self, num_samples, duration=5.0, step=0.5):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        rc = self.L.sd_slide_geometry(int(num_samples), duration, step, C.byref(a), C.byref(b), C.byref(c))
        if rc:
            raise SdError(rc, "sd_slide_geometry: invalid arguments")
        return a.value, b.value, c.value

    def crop_chunks(self, wave, starts_s, duration=5.0, sample_rate=16000):
        wave = np.ascontiguousarray(wave, np.float32)
        starts = np.ascontiguousarray(starts_s, np.float64)
        Ls = int(np.floor(duration * sample_rate))
        out = np.empty((starts.size, Ls), np.float32)
        self._check(self.L.sd_crop_chunks(self.h, _ptr(wave), wave.size, _ptr(starts), starts.size, duration, sample_rate,
                                          _ptr(out)))
        return out

    def clustering_async(self, embeddings, binarized=None, params=None):
        """sd_clustering_async_dev + sd_status_check: same results as clustering(), nothing read back in between."""
        emb = np.ascontiguousarray(embeddings, np.float64)
        Cn, S, D = emb.shape
        keep = np.flatnonzero(~np.isnan(emb.reshape(Cn * S, D)[:, 0])).astype(np.int32)
        p = params if params is not None else self.cluster_params()
        d_e = self.to_device(emb)
        d_b, F = None, 0
        if binarized is not None:
            b = np.ascontiguousarray(binarized, np.float64)
            F = b.shape[1]
            d_b = self.to_device(b)
        d_h, d_k = self.malloc(4 * Cn * S), self.malloc(4)
        try:
            self._check(self.L.sd_status_reset(self.h))
            self._check(self.L.sd_clustering_async_dev(self.h, d_e, Cn, S, D, C.byref(p), _ptr(keep), keep.size, d_b, F,
                                                       d_h, None, 0, d_k))
            self._check(self.L.sd_status_check(self.h))
            hard, k = np.empty((Cn, S), np.int32), np.empty(1, np.int32)
            self.d2h(hard, d_h)
            self.d2h(k, d_k)
        finally:
            for q in (d_e, d_b, d_h, d_k):
                if q:
                    self.free(q)
        return hard, int(k[0])
