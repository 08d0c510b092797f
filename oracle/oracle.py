"""TEST INFRASTRUCTURE ONLY -- ctypes bindings for the two CPU checkers.

* ``Oracle``  -> oracle/liboracle.so   (plain-C restatement, sd_oracle.c)
* ``Ref``     -> oracle/_ref/libsdref.so (the reference's own sources, compiled unmodified)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

c_dp = C.POINTER(C.c_double)
c_fp = C.POINTER(C.c_float)
c_ip = C.POINTER(C.c_int)
c_bp = C.POINTER(C.c_ubyte)
c_lp = C.POINTER(C.c_int64)


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def build(ref=True):
    """Compile liboracle.so and, when /root/reference is present, _ref/libsdref.so."""
    subprocess.check_call(["make", "-s", "-C", HERE, "liboracle.so"])
    if ref:
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


# frame grid of the reference pipeline (speakerDiarizer.cpp:2429-2432, 1335-1340)
FRAME_STEP = 0.016875
FRAME_DURATION = 0.016875
ONSET = 0.4442333667381752
THRESHOLD_F32 = float(np.float32(0.7153814381597874))
MIN_CLUSTER_SIZE = 15


class Oracle:
    def __init__(self):
        path = os.path.join(HERE, "liboracle.so")
        if not os.path.exists(path):
            build(ref=False)
        self.lib = L = C.CDLL(path)
        L.sdo_np_rint.argtypes = [C.c_double]
        L.sdo_np_rint.restype = C.c_int
        L.sdo_closest_frame.argtypes = [C.c_double] * 4
        L.sdo_closest_frame.restype = C.c_long
        L.sdo_aggregate.restype = C.c_long
        L.sdo_aggregate.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_long,
                                    C.c_double, C.c_double, C.c_int, C.c_double, C.c_int, C.c_double, c_dp, C.c_long,
                                    c_dp, c_dp, c_dp]
        L.sdo_trim.restype = C.c_long
        L.sdo_trim.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                               C.c_double, c_dp, c_dp]
        L.sdo_speaker_count.restype = C.c_long
        L.sdo_speaker_count.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_double, c_ip, C.c_long, c_dp]
        L.sdo_binarize.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_int, c_dp]
        L.sdo_stft.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp]
        L.sdo_stft_frames.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, c_fp, C.c_int, C.c_int, c_fp]
        L.sdo_mel_matrix.argtypes = [C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, c_fp]
        L.sdo_mel_matrix.restype = None
        L.sdo_fbank_tail.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_fp]
        L.sdo_hamming_window_f32.argtypes = [C.c_int, c_fp]
        L.sdo_hamming_window_f32.restype = None
        L.sdo_clean_segmentations.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, c_dp]
        L.sdo_normalize_embeddings.argtypes = [c_dp, C.c_int, C.c_int]
        L.sdo_pdist.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        L.sdo_linkage.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        L.sdo_linkage_condensed.argtypes = [c_dp, C.c_int, c_dp]
        L.sdo_fcluster.argtypes = [c_dp, C.c_int, C.c_double, c_ip]
        L.sdo_cosine_cdist.argtypes = [c_dp, C.c_int, c_dp, C.c_int, C.c_int, c_dp]
        L.sdo_cluster_labels.argtypes = [c_dp, C.c_int, C.c_int, C.c_float, C.c_int, c_ip]
        L.sdo_clustering_stage.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, c_dp, C.c_int, c_ip,
                                           c_dp, C.c_int, c_ip]
        L.sdo_mask_compact.argtypes = [c_fp, c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_bp]
        L.sdo_reconstruct.restype = C.c_long
        L.sdo_reconstruct.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_long,
                                      c_ip, c_ip, C.c_long, C.c_double, C.c_double, C.c_double, C.c_long, c_dp,
                                      C.c_long, c_ip, c_dp]
        L.sdo_ingest_pcm16.restype = None
        L.sdo_ingest_pcm16.argtypes = [C.POINTER(C.c_short), C.c_long, c_fp]
        L.sdo_crop.restype = C.c_long
        L.sdo_crop.argtypes = [c_fp, C.c_long, C.c_double, C.c_double, C.c_int, c_fp]
        L.sdo_slide_geometry.restype = None
        L.sdo_slide_geometry.argtypes = [C.c_long, C.c_double, C.c_double, C.POINTER(C.c_long), C.POINTER(C.c_long),
                                         C.POINTER(C.c_long)]
        L.sdo_to_annotation.restype = C.c_long
        L.sdo_to_annotation.argtypes = [c_dp, C.c_long, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, c_dp, c_ip, C.c_long]

    # -- scalars
    def np_rint(self, v):
        return self.lib.sdo_np_rint(float(v))

    def closest_frame(self, t, start=0.0, step=FRAME_STEP, duration=FRAME_DURATION):
        return self.lib.sdo_closest_frame(start, step, duration, float(t))

    # -- stft / fbank
    def hamming_window(self, n=400):
        w = np.empty(n, np.float32)
        self.lib.sdo_hamming_window_f32(n, _p(w, c_fp))
        return w

    def stft(self, wav, n_fft=400, hop=160, window=None):
        wav = _f32(wav)
        B, L = wav.shape
        if window is None:
            window = self.hamming_window(n_fft)
        window = _f32(window)
        T = 1 + L // hop
        out = np.empty((B, T, n_fft // 2 + 1, 2), np.float32)
        self.lib.sdo_stft(_p(wav, c_fp), B, L, n_fft, hop, _p(window, c_fp), _p(out, c_fp))
        return out

    def stft_frames(self, item, t0, t1, n_fft=400, hop=160, window=None):
        item = _f32(item)
        if window is None:
            window = self.hamming_window(n_fft)
        window = _f32(window)
        out = np.empty((t1 - t0, n_fft // 2 + 1, 2), np.float32)
        self.lib.sdo_stft_frames(_p(item, c_fp), item.shape[0], n_fft, hop, _p(window, c_fp), t0, t1, _p(out, c_fp))
        return out

    def mel_matrix(self, n_bins=201, n_mels=80, f_min=0.0, f_max=8000.0, sample_rate=16000):
        W = np.empty((n_bins, n_mels), np.float32)
        self.lib.sdo_mel_matrix(n_bins, n_mels, f_min, f_max, sample_rate, _p(W, c_fp))
        return W

    def fbank_tail(self, stft, wav_lens, W=None):
        stft = _f32(stft)
        B, T, nb, _ = stft.shape
        if W is None:
            W = self.mel_matrix(nb)
        W = _f32(W)
        wav_lens = _f32(wav_lens)
        out = np.empty((B, T, W.shape[1]), np.float32)
        self.lib.sdo_fbank_tail(_p(stft, c_fp), B, T, nb, W.shape[1], _p(W, c_fp), _p(wav_lens, c_fp), _p(out, c_fp))
        return out

    # -- aggregation
    def num_frames(self, C_, sf, pf_step=FRAME_STEP, pf_duration=FRAME_DURATION):
        target = sf[0] + sf[2] + float(C_ - 1) * sf[1]
        return self.lib.sdo_closest_frame(sf[0], pf_step, pf_duration, target) + 1

    def aggregate(self, scores, sf, pf_step=FRAME_STEP, pf_duration=FRAME_DURATION, hamming=False, missing=np.nan,
                  skip_average=False, epsilon=np.finfo(np.float64).eps, want_aux=False):
        """sf = (start, step, duration, num_samples) of the chunk window."""
        scores = _f64(scores)
        Cn, F, K = scores.shape
        NF = self.num_frames(Cn, sf, pf_step, pf_duration)
        out = np.empty((NF, K), np.float64)
        cnt = np.empty((NF, K), np.float64) if want_aux else None
        msk = np.empty((NF, K), np.float64) if want_aux else None
        post = np.empty(4, np.float64)
        r = self.lib.sdo_aggregate(_p(scores, c_dp), Cn, F, K, sf[0], sf[1], sf[2], int(sf[3]), pf_step, pf_duration,
                                   int(hamming), missing, int(skip_average), epsilon, _p(out, c_dp), NF,
                                   _p(cnt, c_dp), _p(msk, c_dp), _p(post, c_dp))
        assert r == NF
        if want_aux:
            return out, post, cnt, msk
        return out, post

    def binarize(self, scores, onset=ONSET, initial_state=False):
        scores = _f32(scores)
        Cn, F, K = scores.shape
        out = np.empty((Cn, F, K), np.float64)
        self.lib.sdo_binarize(_p(scores, c_fp), Cn, F, K, onset, int(initial_state), _p(out, c_dp))
        return out

    def trim(self, binarized, left=0.1, right=0.1, before=(0.0, 0.5, 5.0)):
        b = _f64(binarized)
        Cn, F, K = b.shape
        tw = np.empty(4, np.float64)
        Ft = self.lib.sdo_trim(_p(b, c_dp), Cn, F, K, left, right, before[0], before[1], before[2], None, _p(tw, c_dp))
        out = np.empty((Cn, Ft, K), np.float64)
        self.lib.sdo_trim(_p(b, c_dp), Cn, F, K, left, right, before[0], before[1], before[2], _p(out, c_dp),
                          _p(tw, c_dp))
        return out, tw

    def speaker_count(self, binarized, chunk_step=0.5, chunk_duration=5.0, pf_step=FRAME_STEP,
                      pf_duration=FRAME_DURATION):
        b = _f64(binarized)
        Cn, F, K = b.shape
        cap = int((Cn * chunk_step + chunk_duration) / pf_step) + F + 16
        out = np.empty(cap, np.int32)
        cf = np.empty(4, np.float64)
        n = self.lib.sdo_speaker_count(_p(b, c_dp), Cn, F, K, chunk_step, chunk_duration, pf_step, pf_duration,
                                       _p(out, c_ip), cap, _p(cf, c_dp))
        assert n >= 0
        return out[:n].copy(), cf

    def clean_segmentations(self, binarized):
        b = _f64(binarized)
        out = np.empty_like(b)
        self.lib.sdo_clean_segmentations(_p(b, c_dp), b.shape[0], b.shape[1], b.shape[2], _p(out, c_dp))
        return out

    # -- clustering
    def normalize(self, x):
        x = _f64(x).copy()
        self.lib.sdo_normalize_embeddings(_p(x, c_dp), x.shape[0], x.shape[1])
        return x

    def pdist(self, x):
        x = _f64(x)
        N = x.shape[0]
        out = np.empty(N * (N - 1) // 2, np.float64)
        self.lib.sdo_pdist(_p(x, c_dp), N, x.shape[1], _p(out, c_dp))
        return out

    def linkage(self, x):
        x = _f64(x)
        Z = np.zeros((x.shape[0] - 1, 4), np.float64)
        self.lib.sdo_linkage(_p(x, c_dp), x.shape[0], x.shape[1], _p(Z, c_dp))
        return Z

    def linkage_condensed(self, d, N):
        d = _f64(d)
        Z = np.zeros((N - 1, 4), np.float64)
        self.lib.sdo_linkage_condensed(_p(d, c_dp), N, _p(Z, c_dp))
        return Z

    def fcluster(self, Z, cutoff):
        Z = _f64(Z)
        N = Z.shape[0] + 1
        T = np.zeros(N, np.int32)
        self.lib.sdo_fcluster(_p(Z, c_dp), N, float(cutoff), _p(T, c_ip))
        return T

    def cosine_cdist(self, a, b):
        a, b = _f64(a), _f64(b)
        out = np.empty((a.shape[0], b.shape[0]), np.float64)
        rc = self.lib.sdo_cosine_cdist(_p(a, c_dp), a.shape[0], _p(b, c_dp), b.shape[0], a.shape[1], _p(out, c_dp))
        return rc, out

    def cluster_labels(self, x, threshold=THRESHOLD_F32, min_cluster_size=MIN_CLUSTER_SIZE):
        x = _f64(x)
        lab = np.zeros(x.shape[0], np.int32)
        rc = self.lib.sdo_cluster_labels(_p(x, c_dp), x.shape[0], x.shape[1], threshold, min_cluster_size,
                                         _p(lab, c_ip))
        return rc, lab

    def clustering_stage(self, emb, binarized=None, threshold=THRESHOLD_F32, min_cluster_size=MIN_CLUSTER_SIZE,
                         soft_k_cap=0):
        emb = _f64(emb)
        Cn, S, D = emb.shape
        hard = np.zeros((Cn, S), np.int32)
        soft = np.full((Cn, S, soft_k_cap), np.nan, np.float64) if soft_k_cap else None
        kc = C.c_int(0)
        F = 0
        if binarized is not None:
            binarized = _f64(binarized)
            F = binarized.shape[1]
        rc = self.lib.sdo_clustering_stage(_p(emb, c_dp), Cn, S, D, threshold, min_cluster_size, _p(binarized, c_dp),
                                           F, _p(hard, c_ip), _p(soft, c_dp), soft_k_cap, C.byref(kc))
        return rc, hard, soft, kc.value

    # -- next rows
    def mask_compact(self, wav, masks, min_num_samples=640):
        wav, masks = _f32(wav), _f32(masks)
        B, L = wav.shape
        sig = np.empty((B, L), np.float32)
        lens = np.empty(B, np.float32)
        ts = np.zeros(B, np.uint8)
        rc = self.lib.sdo_mask_compact(_p(wav, c_fp), _p(masks, c_fp), B, L, masks.shape[1], min_num_samples,
                                       _p(sig, c_fp), _p(lens, c_fp), _p(ts, c_bp))
        return rc, sig, lens, ts

    def reconstruct(self, segmentations, sf, hard, count, cf):
        """sf/cf = (start, step, duration, num_samples) of chunk window / count frames."""
        seg = _f32(segmentations)
        Cn, F, K = seg.shape
        hard, count = _i32(hard), _i32(count)
        kc = max(int(hard.max()), 0) + 1
        nf = self.num_frames(Cn, sf, cf[1], cf[2])
        out = np.empty(nf * kc, np.float64)
        cols = C.c_int(0)
        fr = np.zeros(4, np.float64)
        rows = self.lib.sdo_reconstruct(_p(seg, c_fp), Cn, F, K, sf[0], sf[1], sf[2], int(sf[3]), _p(hard, c_ip),
                                        _p(count, c_ip), count.shape[0], cf[0], cf[1], cf[2], int(cf[3]),
                                        _p(out, c_dp), out.size, C.byref(cols), _p(fr, c_dp))
        assert rows >= 0
        return out[:rows * cols.value].reshape(rows, cols.value).copy(), fr[:3].copy()

    def to_annotation(self, scores, frames, onset=0.5, offset=0.5, min_duration_on=0.0,
                      min_duration_off=float(np.float32(0.5817029604921046))):
        s = _f64(scores)
        rows, cols = s.shape
        cap = rows * cols + 8
        seg = np.empty((cap, 2), np.float64)
        lab = np.empty(cap, np.int32)
        n = self.lib.sdo_to_annotation(_p(s, c_dp), rows, cols, frames[0], frames[1], frames[2], onset, offset,
                                       min_duration_on, min_duration_off, _p(seg, c_dp), _p(lab, c_ip), cap)
        assert n >= 0
        return seg[:n].copy(), lab[:n].copy()


    def ingest_pcm16(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        out = np.empty(pcm.shape, np.float32)
        self.lib.sdo_ingest_pcm16(_p(pcm, C.POINTER(C.c_short)), pcm.size, _p(out, c_fp))
        return out

    def crop(self, wave, start, duration=5.0, sample_rate=16000):
        wave = _f32(wave)
        out = np.empty(int(duration * sample_rate) + 8, np.float32)
        n = self.lib.sdo_crop(_p(wave, c_fp), wave.size, float(start), float(duration), sample_rate, _p(out, c_fp))
        return out[:n].copy()

    def slide_geometry(self, num_samples, duration=5.0, step=0.5):
        a, b, c = C.c_long(), C.c_long(), C.c_long()
        self.lib.sdo_slide_geometry(num_samples, duration, step, C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value


class Ref:
    """The reference itself (oracle/_ref/libsdref.so).  ``Ref.available()`` is False when it has not been built
    (e.g. /root/reference absent and no prebuilt library travelled with the snapshot)."""

    PATH = os.path.join(HERE, "_ref", "libsdref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        import torch  # noqa: F401  (libsdref.so links libtorch from the venv; importing resolves its deps)
        self.lib = L = C.CDLL(self.PATH)
        L.ref_np_rint.argtypes = [C.c_double]
        L.ref_closest_frame.argtypes = [C.c_double] * 4
        L.ref_closest_frame.restype = C.c_long
        L.ref_stft.argtypes = [c_fp, C.c_int, C.c_int, c_fp, C.c_int, c_fp, c_fp, c_lp]
        L.ref_stft_fft_only.argtypes = [c_fp, C.c_int, C.c_int, c_fp]
        L.ref_aggregate.restype = C.c_long
        L.ref_aggregate.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_long,
                                    C.c_double, C.c_double, C.c_double, C.c_int, c_dp, C.c_long, c_dp]
        L.ref_binarize_swf.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, C.c_int, c_dp]
        L.ref_binarize_ndarray.argtypes = [c_dp, C.c_int, C.c_int, C.c_double, C.c_int, c_bp]
        L.ref_trim.restype = C.c_long
        L.ref_trim.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                               C.c_double, c_dp, c_dp]
        L.ref_speaker_count.restype = C.c_long
        L.ref_speaker_count.argtypes = [c_fp, c_dp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double,
                                        C.c_int, c_ip, C.c_long, c_dp]
        L.ref_clean_segmentations.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, c_dp]
        L.ref_normalize_embeddings.argtypes = [c_dp, C.c_int, C.c_int]
        L.ref_pdist.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        L.ref_linkage.argtypes = [c_dp, C.c_int, C.c_int, c_dp]
        L.ref_fcluster.argtypes = [c_dp, C.c_int, C.c_double, c_ip]
        L.ref_clustering_cluster.argtypes = [c_dp, C.c_int, C.c_int, C.c_double, c_ip]
        L.ref_cosine_cdist.argtypes = [c_dp, C.c_int, c_dp, C.c_int, C.c_int, c_dp]
        L.ref_cluster_labels.argtypes = [c_dp, C.c_int, C.c_int, c_ip]
        L.ref_clustering_stage.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, c_dp, C.c_int, c_ip]
        L.ref_mask_compact.argtypes = [c_fp, c_fp, C.c_int, C.c_int, C.c_int, c_fp, c_fp, c_bp]
        L.ref_reconstruct.restype = C.c_long
        L.ref_reconstruct.argtypes = [c_fp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_double, C.c_long,
                                      c_ip, c_ip, C.c_long, C.c_double, C.c_double, C.c_double, C.c_long, c_dp,
                                      C.c_long, c_ip, c_dp]
        L.ref_to_annotation.restype = C.c_long
        L.ref_to_annotation.argtypes = [c_dp, C.c_long, C.c_int, C.c_double, C.c_double, C.c_double, C.c_double,
                                        C.c_double, C.c_double, C.c_double, c_dp, c_ip, C.c_long]

    def np_rint(self, v):
        return self.lib.ref_np_rint(float(v))

    def closest_frame(self, t, start=0.0, step=FRAME_STEP, duration=FRAME_DURATION):
        return self.lib.ref_closest_frame(start, step, duration, float(t))

    def stft(self, wav, lens=None):
        """EmbeddingModel1::infer as written -> the [32, T, 201, 2] tensor handed to emd4.onnx, and wav_lens[32]."""
        wav = _f32(wav)
        B, L = wav.shape
        lens = _f32(np.ones(B) if lens is None else lens)
        T = 1 + L // 160
        out = np.zeros((32, T, 201, 2), np.float32)
        wl = np.zeros(32, np.float32)
        shape = np.zeros(4, np.int64)
        rc = self.lib.ref_stft(_p(wav, c_fp), B, L, _p(lens, c_fp), lens.shape[0], _p(out, c_fp), _p(wl, c_fp),
                               _p(shape, c_lp))
        assert rc == 0 and tuple(shape) == out.shape, (rc, shape)
        return out, wl

    def stft_fft_only(self, wav):
        wav = _f32(wav)
        B, L = wav.shape
        out = np.empty((B, 1 + L // 160, 201, 2), np.float32)
        self.lib.ref_stft_fft_only(_p(wav, c_fp), B, L, _p(out, c_fp))
        return out

    def aggregate(self, scores, sf, pf_step=FRAME_STEP, pf_duration=FRAME_DURATION, missing=np.nan,
                  skip_average=False):
        scores = _f64(scores)
        Cn, F, K = scores.shape
        cap = int((Cn * sf[1] + sf[2]) / pf_step) + F + 16
        out = np.empty((cap, K), np.float64)
        post = np.empty(4, np.float64)
        n = self.lib.ref_aggregate(_p(scores, c_dp), Cn, F, K, sf[0], sf[1], sf[2], int(sf[3]), pf_step, pf_duration,
                                   missing, int(skip_average), _p(out, c_dp), cap, _p(post, c_dp))
        assert n >= 0
        return out[:n].copy(), post

    def binarize(self, scores, initial_state=False):
        scores = _f32(scores)
        Cn, F, K = scores.shape
        out = np.empty((Cn, F, K), np.float64)
        self.lib.ref_binarize_swf(_p(scores, c_fp), Cn, F, K, int(initial_state), _p(out, c_dp))
        return out

    def binarize_ndarray(self, scores, onset=0.5, initial_state=False):
        s = _f64(scores)
        out = np.empty(s.shape, np.uint8)
        self.lib.ref_binarize_ndarray(_p(s, c_dp), s.shape[0], s.shape[1], onset, int(initial_state), _p(out, c_bp))
        return out

    def trim(self, binarized, left=0.1, right=0.1, before=(0.0, 0.5, 5.0)):
        b = _f64(binarized)
        Cn, F, K = b.shape
        out = np.empty((Cn, F, K), np.float64)
        tw = np.empty(4, np.float64)
        Ft = self.lib.ref_trim(_p(b, c_dp), Cn, F, K, left, right, before[0], before[1], before[2], _p(out, c_dp),
                               _p(tw, c_dp))
        return out.reshape(-1)[:Cn * Ft * K].reshape(Cn, Ft, K).copy(), tw

    def speaker_count(self, binarized, pf_step=FRAME_STEP, pf_duration=FRAME_DURATION, num_samples=16000):
        b = _f64(binarized)
        Cn, F, K = b.shape
        seg = np.zeros((Cn, F, K), np.float32)
        cap = int((Cn * 0.5 + 5.0) / pf_step) + F + 16
        out = np.empty(cap, np.int32)
        cf = np.empty(4, np.float64)
        n = self.lib.ref_speaker_count(_p(seg, c_fp), _p(b, c_dp), Cn, F, K, 0.0, pf_step, pf_duration, num_samples,
                                       _p(out, c_ip), cap, _p(cf, c_dp))
        assert n >= 0
        return out[:n].copy(), cf

    def clean_segmentations(self, binarized):
        b = _f64(binarized)
        out = np.empty_like(b)
        self.lib.ref_clean_segmentations(_p(b, c_dp), b.shape[0], b.shape[1], b.shape[2], _p(out, c_dp))
        return out

    def normalize(self, x):
        x = _f64(x).copy()
        self.lib.ref_normalize_embeddings(_p(x, c_dp), x.shape[0], x.shape[1])
        return x

    def pdist(self, x):
        x = _f64(x)
        N = x.shape[0]
        out = np.empty(N * (N - 1) // 2, np.float64)
        self.lib.ref_pdist(_p(x, c_dp), N, x.shape[1], _p(out, c_dp))
        return out

    def linkage(self, x):
        x = _f64(x)
        Z = np.zeros((x.shape[0] - 1, 4), np.float64)
        self.lib.ref_linkage(_p(x, c_dp), x.shape[0], x.shape[1], _p(Z, c_dp))
        return Z

    def fcluster(self, Z, cutoff):
        Z = _f64(Z)
        N = Z.shape[0] + 1
        T = np.zeros(N, np.int32)
        self.lib.ref_fcluster(_p(Z, c_dp), N, float(cutoff), _p(T, c_ip))
        return T

    def clustering_cluster(self, x, cutoff):
        x = _f64(x)
        T = np.zeros(x.shape[0], np.int32)
        self.lib.ref_clustering_cluster(_p(x, c_dp), x.shape[0], x.shape[1], float(cutoff), _p(T, c_ip))
        return T

    def cosine_cdist(self, a, b):
        a, b = _f64(a), _f64(b)
        out = np.empty((a.shape[0], b.shape[0]), np.float64)
        rc = self.lib.ref_cosine_cdist(_p(a, c_dp), a.shape[0], _p(b, c_dp), b.shape[0], a.shape[1], _p(out, c_dp))
        return rc, out

    def cluster_labels(self, x):
        x = _f64(x)
        lab = np.zeros(x.shape[0], np.int32)
        rc = self.lib.ref_cluster_labels(_p(x, c_dp), x.shape[0], x.shape[1], _p(lab, c_ip))
        return rc, lab

    def clustering_stage(self, emb, binarized=None):
        emb = _f64(emb)
        Cn, S, D = emb.shape
        hard = np.zeros((Cn, S), np.int32)
        F = 0
        if binarized is not None:
            binarized = _f64(binarized)
            F = binarized.shape[1]
        rc = self.lib.ref_clustering_stage(_p(emb, c_dp), Cn, S, D, _p(binarized, c_dp), F, _p(hard, c_ip))
        return rc, hard

    def mask_compact(self, wav, masks):
        wav, masks = _f32(wav), _f32(masks)
        B, L = wav.shape
        sig = np.empty((B, L), np.float32)
        lens = np.empty(B, np.float32)
        ts = np.zeros(B, np.uint8)
        rc = self.lib.ref_mask_compact(_p(wav, c_fp), _p(masks, c_fp), B, L, masks.shape[1], _p(sig, c_fp),
                                       _p(lens, c_fp), _p(ts, c_bp))
        return rc, sig, lens, ts

    def reconstruct(self, segmentations, sf, hard, count, cf):
        seg = _f32(segmentations)
        Cn, F, K = seg.shape
        hard, count = _i32(hard), _i32(count)
        kc = max(int(hard.max()), 0) + 1
        cap = (int((Cn * sf[1] + sf[2]) / cf[1]) + F + 16) * kc
        out = np.empty(cap, np.float64)
        cols = C.c_int(0)
        fr = np.zeros(4, np.float64)
        rows = self.lib.ref_reconstruct(_p(seg, c_fp), Cn, F, K, sf[0], sf[1], sf[2], int(sf[3]), _p(hard, c_ip),
                                        _p(count, c_ip), count.shape[0], cf[0], cf[1], cf[2], int(cf[3]),
                                        _p(out, c_dp), cap, C.byref(cols), _p(fr, c_dp))
        assert rows >= 0
        return out[:rows * cols.value].reshape(rows, cols.value).copy(), fr[:3].copy()

    def to_annotation(self, scores, frames, onset=0.5, offset=0.5, min_duration_on=0.0,
                      min_duration_off=float(np.float32(0.5817029604921046))):
        s = _f64(scores)
        rows, cols = s.shape
        cap = rows * cols + 8
        seg = np.empty((cap, 2), np.float64)
        lab = np.empty(cap, np.int32)
        n = self.lib.ref_to_annotation(_p(s, c_dp), rows, cols, frames[0], frames[1], frames[2], onset, offset,
                                       min_duration_on, min_duration_off, _p(seg, c_dp), _p(lab, c_ip), cap)
        assert n >= 0
        return seg[:n].copy(), lab[:n].copy()

    def wav_load(self, path, cap=1 << 24):
        self.lib.ref_wav_load.restype = C.c_long
        self.lib.ref_wav_load.argtypes = [C.c_char_p, c_fp, C.c_long, C.POINTER(C.c_int)]
        out = np.empty(cap, np.float32)
        meta = (C.c_int * 3)()
        n = self.lib.ref_wav_load(path.encode(), _p(out, c_fp), cap, meta)
        assert n >= 0
        return out[:n].copy(), tuple(meta)

    def crop(self, wave, start):
        self.lib.ref_crop.restype = C.c_long
        self.lib.ref_crop.argtypes = [c_fp, C.c_long, C.c_double, c_fp]
        wave = _f32(wave)
        out = np.empty(80000 + 8, np.float32)
        n = self.lib.ref_crop(_p(wave, c_fp), wave.size, float(start), _p(out, c_fp))
        return out[:n].copy()


# ---------------------------------------------------------------------------------------------------------------
# Kaldi-compatible front-end (BASELINE north_star bullet 1).  Not on the reference's path -- the reference frames
# with torch::stft / Hamming -- so there is no reference code to restate: this follows kaldi::ProcessWindow /
# kaldi::MelBanks as published (feature-window.cc, mel-computations.cc) in the form torchaudio.compliance.kaldi gives
# them, in fp64 numpy, and is pinned against torchaudio.compliance.kaldi.fbank itself (tests/test_oracle_golden.py,
# tests/golden/kaldi_fbank.npz generated by oracle/make_golden_kaldi.py).
def kaldi_frames(x, snip_edges=False, preemph=0.97, remove_dc_offset=True, n_fft=400, hop=160):
    """Conditioned, povey-windowed frames [T, n_fft] (fp64) of a 1-D signal."""
    x = np.asarray(x, np.float64)
    L = x.shape[0]
    if snip_edges:
        T = 0 if L < n_fft else 1 + (L - n_fft) // hop
        idx = hop * np.arange(T)[:, None] + np.arange(n_fft)[None, :]
    else:
        T = (L + hop // 2) // hop
        idx = hop * np.arange(T)[:, None] - (n_fft // 2 - hop // 2) + np.arange(n_fft)[None, :]
        idx = np.where(idx < 0, -idx - 1, idx)
        idx = np.where(idx >= L, 2 * L - 1 - idx, idx)
    fr = x[idx]
    if remove_dc_offset:
        fr = fr - fr.mean(axis=1, keepdims=True)
    if preemph != 0.0:
        prev = np.concatenate([fr[:, :1], fr[:, :-1]], axis=1)
        fr = fr - preemph * prev
    n = np.arange(n_fft)
    window = (0.5 - 0.5 * np.cos(2.0 * np.pi * n / (n_fft - 1))) ** 0.85
    return fr * window


def kaldi_stft(x, **kw):
    """[T, 201, 2] fp64 spectrum of the conditioned frames."""
    X = np.fft.rfft(kaldi_frames(x, **kw), axis=1)
    return np.stack([X.real, X.imag], axis=-1)


def kaldi_mel_banks(n_mels=80, n_fft=400, sample_rate=16000, low_freq=20.0, high_freq=0.0):
    """[n_mels, n_fft/2 + 1] (last column zero: the Nyquist bin gets no weight)."""
    nyq = 0.5 * sample_rate
    high = high_freq + nyq if high_freq <= 0 else high_freq
    mel = lambda f: 1127.0 * np.log(1.0 + np.asarray(f, np.float64) / 700.0)
    lo, hi = mel(low_freq), mel(high)
    delta = (hi - lo) / (n_mels + 1)
    m = np.arange(n_mels)[:, None]
    left, center, right = lo + m * delta, lo + (m + 1) * delta, lo + (m + 2) * delta
    mf = mel(sample_rate / n_fft * np.arange(n_fft // 2))[None, :]
    w = np.maximum(0.0, np.minimum((mf - left) / (center - left), (right - mf) / (right - center)))
    return np.concatenate([w, np.zeros((n_mels, 1))], axis=1)


def kaldi_fbank(x, n_mels=80, snip_edges=False, preemph=0.97, remove_dc_offset=True, subtract_mean=False):
    """log mel energies [T, n_mels] like torchaudio.compliance.kaldi.fbank(dither=0, round_to_power_of_two=False,
    window_type="povey", use_power=True, use_log_fbank=True, low_freq=20, high_freq=0)."""
    X = np.fft.rfft(kaldi_frames(x, snip_edges, preemph, remove_dc_offset), axis=1)
    e = (np.abs(X) ** 2) @ kaldi_mel_banks(n_mels).T
    out = np.log(np.maximum(e, float(np.finfo(np.float32).eps)))
    if subtract_mean:
        out = out - out.mean(axis=0, keepdims=True)
    return out
