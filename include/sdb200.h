/* sdb200.h -- C-ABI of libsdb200.so: the B200 (sm_100a) implementation of the data-parallel hot path of
 * leohuang2013/pyannote-audio_speaker-diarization_cpp.
 *
 * The reference has no plugin / FFI layer: its boundary is the C++ function-level API of
 * pipeline/src/speakerDiarizer.cpp (SD) and pipeline/src/clustering/clustering.h (CL).  Each entry point
 * below names the reference function it replaces (file:line relative to the reference checkout).  The C++
 * host shim `host/sdb200_host.hpp` rebuilds the reference's nested-std::vector signatures on top of these,
 * so speakerDiarization() (SD:2937) can call them unchanged.
 *
 * Conventions
 *  - plain pointers and sizes, caller-owned row-major buffers, no exceptions across the ABI;
 *  - every function returns an sd_status (0 = OK); sd_last_error() gives the message;
 *  - `*_dev` variants take DEVICE pointers and only enqueue work on the context's stream (no sync);
 *    the un-suffixed variants take HOST pointers and include H2D, kernels, D2H and a stream sync;
 *  - one sd_ctx per host thread and GPU (the reference is single-threaded and non-re-entrant);
 *  - there is no CPU fallback: without a CUDA device sd_ctx_create fails with SD_ERR_CUDA.
 */
#ifndef SDB200_H_
#define SDB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDB200_VERSION 100

typedef struct sd_ctx sd_ctx;

typedef enum sd_status {
    SD_OK = 0,
    SD_ERR_INVALID = 1,        /* bad argument (NULL, non-positive size, ...) */
    SD_ERR_CUDA = 2,           /* CUDA runtime / driver failure, or no device */
    SD_ERR_ZERO_MAGNITUDE = 3, /* reference throws std::runtime_error("Vectors have zero magnitude.") SD:493-495 */
    SD_ERR_UNSUPPORTED = 4,    /* parameter combination the reference itself does not implement */
    SD_ERR_NOMEM = 5,
    SD_ERR_CAPACITY = 6        /* caller buffer too small; required size reported through the out-param */
} sd_status;

/* SlidingWindow POD, SD:1029-1036 */
typedef struct sd_window {
    double start;
    double step;
    double duration;
    int64_t num_samples;
} sd_window;

/* ---- context / memory / timing ------------------------------------------------------------------ */
int sd_version(void);
int sd_ctx_create(int device, sd_ctx** out);
void sd_ctx_destroy(sd_ctx* ctx);
const char* sd_last_error(const sd_ctx* ctx);
/* Adopt an existing cudaStream_t (e.g. the framework's current stream); NULL restores the private stream. */
int sd_ctx_set_stream(sd_ctx* ctx, void* cuda_stream);
void* sd_ctx_stream(sd_ctx* ctx);
int sd_sync(sd_ctx* ctx);
int sd_malloc(sd_ctx* ctx, size_t bytes, void** dptr);
int sd_free(sd_ctx* ctx, void* dptr);
int sd_host_alloc(sd_ctx* ctx, size_t bytes, void** hptr); /* pinned */
int sd_host_free(sd_ctx* ctx, void* hptr);
int sd_memcpy_h2d(sd_ctx* ctx, void* dst, const void* src, size_t bytes); /* async on the ctx stream */
int sd_memcpy_d2h(sd_ctx* ctx, void* dst, const void* src, size_t bytes);
int sd_memset(sd_ctx* ctx, void* dst, int value, size_t bytes);
/* CUDA-event timers recorded on the ctx stream (slot 0..15). */
int sd_timer_start(sd_ctx* ctx, int slot);
int sd_timer_stop(sd_ctx* ctx, int slot);
int sd_timer_elapsed_ms(sd_ctx* ctx, int slot, float* ms); /* synchronises on the stop event */
/* Number of kernels of this library launched on this ctx since creation. */
int64_t sd_launch_count(const sd_ctx* ctx);
/* Options.  SD_OPT_FORCE_EXACT_LINKAGE: always run the heap-driven linkage kernel (normally it only runs when
 * the heap-free kernel meets a tied minimum); results are identical either way. */
typedef enum sd_option { SD_OPT_FORCE_EXACT_LINKAGE = 1, SD_OPT_LINKAGE_THREADS = 2 /* 0 auto, 512, 1024 */, SD_OPT_STFT_VARIANT = 3 /* tuning: other builds of the front-end kernels, results within the same tolerance -- 0 default (register Hamming window, 4 CTAs/SM), 1 / 2 window table at 3 / 4 CTAs/SM, 3 five CTAs/SM, 5 8-frame tiles, 6 two sample buffers, 8 fbank mel projection by (frame, part) threads */, SD_OPT_LINKAGE_CLUSTER = 4 /* 1 (default): 8-CTA cluster merge loop, 0: one CTA */, SD_OPT_STFT_WAVES = 6 /* CTAs per resident slot of the STFT grid: 1 = persistent CTAs (default), n > 1 = slots are handed back n times per launch */, SD_OPT_LINKAGE_WIDE = 5 /* whole-GPU merge loop (one CTA per SM, cooperative launch, global-memory mailbox): 0 never (default), 1 from 32768 rows, 2 always */ } sd_option;
int sd_ctx_set_option(sd_ctx* ctx, int option, int value);
/* Diagnostic counters accumulated by the kernels: [0] stale nearest-neighbour revalidations in linkage,
 * [1] heap updates replayed after Lance-Williams sweeps, [2] problems handed from the heap-free linkage kernel to
 * the heap-driven one, [3] cycles popping/publishing, [4] cycles revalidating, [5] cycles sweeping, [6] cycles at
 * the request barrier (all as seen by the control warp of the heap-free kernel), [7] merges.  Synchronises. */
int sd_debug_counters(sd_ctx* ctx, int64_t* out8, int reset);
/* Write `bytes` of zeros into a scratch buffer larger than L2 (benchmark hygiene). */
int sd_flush_l2(sd_ctx* ctx);

/* ---- a1/a2: STFT front-end of the embedding stage ------------------------------------------------
 * Replaces EmbeddingModel1::infer (SD:1977-2036) up to the tensor handed to emd4.onnx, and the packing
 * of EmbeddingModel1::_infer (SD:1889-1917): centre zero padding n_fft/2, periodic Hamming window,
 * n_fft-point real DFT, one-sided, un-normalised, layout [B][T][n_fft/2+1][{re,im}] fp32, T = 1 + L/hop. */
typedef enum sd_window_kind { SD_WINDOW_HAMMING_PERIODIC = 0, SD_WINDOW_POVEY = 1, SD_WINDOW_CUSTOM = 2 } sd_window_kind;
/* Where frame t starts and what lies beyond the ends of the signal:
 *  SD_FRAMES_CENTER_ZERO   t*hop - n_fft/2, zeros: torch::stft(center=true, "constant"), the reference (SD:2008);
 *                          T = 1 + L/hop
 *  SD_FRAMES_KALDI_REFLECT t*hop - (n_fft/2 - hop/2), the signal mirrored about its ends: Kaldi snip_edges=false;
 *                          T = (L + hop/2) / hop
 *  SD_FRAMES_KALDI_SNIP    t*hop, whole frames only: Kaldi snip_edges=true; T = 1 + (L - n_fft)/hop */
typedef enum sd_frame_mode { SD_FRAMES_CENTER_ZERO = 0, SD_FRAMES_KALDI_REFLECT = 1, SD_FRAMES_KALDI_SNIP = 2 } sd_frame_mode;

typedef struct sd_stft_params {
    int n_fft;          /* 400 (the only size with a kernel today) */
    int hop;            /* 160 */
    int window_kind;    /* sd_window_kind */
    const float* window; /* HOST pointer to n_fft floats when SD_WINDOW_CUSTOM, else NULL */
    float preemph;      /* 0 = off (reference); 0.97 = Kaldi per-frame pre-emphasis y[n] = x[n] - c x[n-1], y[0] = x[0] - c x[0] */
    int pad_batch_to;   /* 32 reproduces _infer's fixed batch (rows >= B are zero); 0 = no padding */
    int frame_mode;     /* sd_frame_mode; 0 = the reference */
    int remove_dc_offset; /* Kaldi: subtract the mean of every frame before pre-emphasis; 0 = off (reference) */
} sd_stft_params;

void sd_stft_default_params(sd_stft_params* p);
/* Kaldi-compatible framing (kaldi::FrameExtractionOptions defaults at 16 kHz / 25 ms / 10 ms with a 400-point
 * transform, i.e. round_to_power_of_two=false): povey window, pre-emphasis 0.97, DC removal, snip_edges as given. */
void sd_stft_kaldi_params(sd_stft_params* p, int snip_edges);
int64_t sd_stft_num_frames(int L, int hop); /* SD_FRAMES_CENTER_ZERO */
int64_t sd_stft_num_frames_mode(int L, int n_fft, int hop, int frame_mode);
/* out must hold max(B, pad_batch_to) * T * (n_fft/2+1) * 2 floats. */
int sd_stft(sd_ctx* ctx, const float* wav, int B, int L, const sd_stft_params* p, float* out);
int sd_stft_dev(sd_ctx* ctx, const float* d_wav, int B, int L, const sd_stft_params* p, float* d_out);
/* wav_lens packing of _infer (SD:1899-1900): copy n_lens values, remaining rows 1.0 (host only, trivial). */
int sd_pack_wav_lens(const float* lens, int n_lens, int batch, float* out);

/* ---- a3: fused |X|^2 -> mel -> dB -> top_db -> mean-norm (executes inside emd4.onnx in the reference;
 * spec: embeddings/threeModel.py:212-221, 333-396).  out[B][T][n_mels] fp32.  Optional product. */
typedef struct sd_fbank_params {
    sd_stft_params stft;
    int n_mels;        /* 80 */
    float f_min;       /* 0 */
    float f_max;       /* 8000 */
    int sample_rate;   /* 16000 */
    float top_db;      /* 80 */
    float amin;        /* 1e-10 */
    int mean_norm;     /* 1 */
    int mel_kind;      /* 0: speechbrain Filterbank (triangles in Hz, mel = 2595 log10(1 + f/700));
                          1: Kaldi mel banks (triangles in mel, mel = 1127 ln(1 + f/700), bins below n_fft/2 only) */
    int log_kind;      /* 0: 10 log10 (dB) clamped to max - top_db;  1: natural log, floor amin, no clamp (Kaldi) */
} sd_fbank_params;
void sd_fbank_default_params(sd_fbank_params* p);
/* torchaudio.compliance.kaldi.fbank / kaldi::FbankOptions defaults with an 80-bin bank and a 400-point transform:
 * low_freq 20, high_freq Nyquist, natural log floored at FLT_EPSILON, no mean normalisation; framing as
 * sd_stft_kaldi_params. */
void sd_fbank_kaldi_params(sd_fbank_params* p, int snip_edges);
int sd_fbank(sd_ctx* ctx, const float* wav, int B, int L, const float* wav_lens, const sd_fbank_params* p, float* out);
int sd_fbank_dev(sd_ctx* ctx, const float* d_wav, int B, int L, const float* d_wav_lens, const sd_fbank_params* p,
                 float* d_out);

/* ---- a4/a5: sliding-window aggregation ------------------------------------------------------------
 * Replaces PipelineHelper::aggregate (SD:1167-1311), SlidingWindow::closest_frame (SD:1084-1090) and
 * Helper::np_rint (SD:260-272). */
int sd_np_rint(double v);
int64_t sd_closest_frame(const sd_window* w, double t);
/* number of output frames for C chunks: closest_frame(start + duration + (C-1)*step) + 1, SD:1232-1234 */
int64_t sd_aggregate_num_frames(int C, const sd_window* chunks, const sd_window* frames);
/* scores[C][F][K] fp64 (NaN = missing) -> out[num_frames][K].  hamming != 0 weights each chunk with
 * np.hamming(F) (the reference asserts "not implemented", SD:1214; pyannote semantics).
 * count_out / mask_out (optional, [num_frames][K]) are overlapping_chunk_count / aggregated_mask (SD:1241-1245).
 * cap_rows = rows available in out; *num_frames receives the row count (SD_ERR_CAPACITY if too small). */
int sd_aggregate(sd_ctx* ctx, const double* scores, int C, int F, int K, const sd_window* chunks,
                 const sd_window* frames, int hamming, double missing, int skip_average, double epsilon, double* out,
                 int64_t cap_rows, int64_t* num_frames, sd_window* post_frames, double* count_out, double* mask_out);
int sd_aggregate_dev(sd_ctx* ctx, const double* d_scores, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int hamming, double missing, int skip_average, double epsilon,
                     double* d_out, int64_t cap_rows, int64_t* num_frames, sd_window* post_frames, double* d_count_out,
                     double* d_mask_out);

/* ---- a6/a7: binarisation, trim, speaker count, clean ---------------------------------------------
 * sd_binarize        : SegmentModel::binarize_swf (SD:1506-1563)  scores[C][F][K] fp32 -> out[C][F][K] fp64 {0,1}
 * sd_binarize_rows   : SegmentModel::binarize_ndarray (SD:1565-1639) scores[R][F] fp64 -> out[R][F] uint8
 * sd_trim            : SegmentModel::trim (SD:1742-1782)
 * sd_speaker_count   : SegmentModel::speaker_count (SD:1665-1738): trim 10%/10% -> sum_k -> aggregate -> rint
 * sd_clean_segmentations : Helper::cleanSegmentations (SD:710-743) */
int sd_binarize(sd_ctx* ctx, const float* scores, int C, int F, int K, double onset, int initial_state, double* out);
int sd_binarize_dev(sd_ctx* ctx, const float* d_scores, int C, int F, int K, double onset, int initial_state,
                    double* d_out);
int sd_binarize_rows(sd_ctx* ctx, const double* scores, int R, int F, double onset, int initial_state, uint8_t* out);
int64_t sd_trim_num_frames(int F, double left, double right);
int sd_trim(sd_ctx* ctx, const double* binarized, int C, int F, int K, double left, double right,
            const sd_window* before, double* out, sd_window* trimmed_frames);
int sd_speaker_count(sd_ctx* ctx, const double* binarized, int C, int F, int K, const sd_window* chunks,
                     const sd_window* frames, int32_t* out, int64_t cap, int64_t* n_out, sd_window* count_frames);
int sd_speaker_count_dev(sd_ctx* ctx, const double* d_binarized, int C, int F, int K, const sd_window* chunks,
                         const sd_window* frames, int32_t* d_out, int64_t cap, int64_t* n_out, sd_window* count_frames);
int sd_clean_segmentations(sd_ctx* ctx, const double* binarized, int C, int F, int K, double* out);

/* ---- a9-a12: clustering library --------------------------------------------------------------------
 * sd_normalize   : Helper::normalizeEmbeddings (SD:330-357)   in place, x[N][D]
 * sd_pdist       : pdist of Clustering::linkage (CL:408-431)  condensed fp64, N(N-1)/2
 * sd_linkage     : Clustering::linkage (CL:417-440) = pdist + fast_linkage (CL:289-406); Z[(N-1)][4]
 * sd_fcluster    : Clustering::fcluster (CL:442-457), criterion "distance"; T[N] labels 1..K
 * sd_cluster     : Clustering::cluster (CL:459-468)
 * sd_cosine_cdist: Helper::cosineSimilarity (SD:502-516), out[na][nb] */
typedef enum sd_pdist_mode {
    SD_PDIST_EXACT_F64 = 0, /* parity mode: fp64, sequential k, no FMA -> bit-identical to the reference */
    SD_PDIST_GEMM_TF32X3 = 1 /* tcgen05 Gram GEMM with 3xTF32 split precision; |err| evidenced in DESIGN.md */
} sd_pdist_mode;
int sd_normalize(sd_ctx* ctx, double* x, int N, int D);
int sd_pdist(sd_ctx* ctx, const double* x, int N, int D, int mode, double* condensed);
int sd_linkage(sd_ctx* ctx, const double* x, int N, int D, double* Z);
int sd_linkage_dev(sd_ctx* ctx, const double* d_x, int N, int D, double* d_Z);
/* CUDA-event times of the last linkage on this context: the distance matrix and the N-1 merges (synchronises). */
int sd_linkage_stage_ms(sd_ctx* ctx, float* pdist_ms, float* merges_ms);
int sd_fcluster(sd_ctx* ctx, const double* Z, int N, double cutoff, int32_t* T);
int sd_cluster(sd_ctx* ctx, const double* x, int N, int D, double cutoff, int32_t* T);
int sd_cosine_cdist(sd_ctx* ctx, const double* a, int na, const double* b, int nb, int D, double* out);

/* ---- a8/a13/a14/a15: clustering driver -----------------------------------------------------------
 * sd_cluster_labels : Cluster::cluster (SD:2300-2422) on already-filtered embeddings x[N][D] -> labels[N]
 * sd_clustering     : Cluster::clustering (SD:2063-2116) = filter_embeddings (SD:2214) + set_num_clusters
 *                     (SD:2261) + cluster + assign_embeddings (SD:2120-2212), then, when binarized != NULL,
 *                     the inactive-speaker mask of SD:3166-3191 (hard = -2).
 *   embeddings[C][S][D] fp64, a row whose first element is NaN is absent;  hard[C][S];
 *   soft (optional) [C][S][soft_k_cap] receives 2 - cosine distance to each centroid. */
typedef struct sd_cluster_params {
    float threshold;       /* Cluster::m_threshold, a float: 0.7153814381597874f (SD:2049) */
    int min_cluster_size;  /* 15 (SD:2050) */
    int num_clusters;      /* -1; anything else is "not implemented" in the reference (SD:2368-2369) */
    int min_clusters;      /* -1 */
    int max_clusters;      /* -1 */
    int pdist_mode;        /* sd_pdist_mode */
} sd_cluster_params;
void sd_cluster_default_params(sd_cluster_params* p);
int sd_cluster_labels(sd_ctx* ctx, const double* x, int N, int D, const sd_cluster_params* p, int32_t* labels);
int sd_clustering(sd_ctx* ctx, const double* embeddings, int C, int S, int D, const sd_cluster_params* p,
                  const double* binarized, int F, int32_t* hard, double* soft, int soft_k_cap, int* num_clusters);
int sd_clustering_dev(sd_ctx* ctx, const double* d_embeddings, int C, int S, int D, const sd_cluster_params* p,
                      const double* d_binarized, int F, int32_t* d_hard, double* d_soft, int soft_k_cap,
                      int* num_clusters);

/* sd_clustering with one more optional output: dist[C][S][soft_k_cap] = the cosine distances to the centroids
 * themselves (Helper::cosineSimilarity inside assign_embeddings, SD:2183; soft = 2 - dist), NaN beyond the cluster
 * count.  This is what the reference dumps as cpp_dist (SD:2186). */
int sd_clustering_ex(sd_ctx* ctx, const double* embeddings, int C, int S, int D, const sd_cluster_params* p,
                     const double* binarized, int F, int32_t* hard, double* soft, double* dist, int soft_k_cap,
                     int* num_clusters);

/* Asynchronous Cluster::clustering: like sd_clustering_dev, but nothing is read back and the call only enqueues, so
 * that a stream never idles between the kernels of one file (which matters when several files are in flight on
 * different streams).  The caller supplies what the synchronous call reads back at its start: keep_rows (host array,
 * ascending) = the rows of d_embeddings[C*S][D] whose first element is not NaN (filter_embeddings, SD:2214-2259 --
 * in the pipeline these are the items that were not "too short").  The final cluster count is written to
 * *d_num_clusters on the device.  Errors found on the device (zero-magnitude vectors, more than 1 024 raw clusters)
 * are latched in the context's status word: sd_status_reset before the first call of a sequence, sd_status_check
 * (synchronises) after the last. */
int sd_clustering_async_dev(sd_ctx* ctx, const double* d_embeddings, int C, int S, int D, const sd_cluster_params* p,
                            const int32_t* keep_rows, int n_keep, const double* d_binarized, int F, int32_t* d_hard,
                            double* d_soft, int soft_k_cap, int32_t* d_num_clusters);
int sd_status_reset(sd_ctx* ctx);
int sd_status_check(sd_ctx* ctx);

/* ---- "next" rows --------------------------------------------------------------------------
 * f1: pre-embedding masking.
 * sd_mask_compact      : Helper::interpolate (SD:746) + Helper::padSequence (SD:770) + the wav_lens / too-short logic
 *                        of getEmbedding (SD:2466-2510) for one batch: wav[B][L], masks[B][F] ->
 *                        signals[B][L] (compacted, zero tail), wav_lens[B], too_short[B]; *all_too_short = 1 when the
 *                        longest item is shorter than min_num_samples (the reference then returns NaN embeddings).
 * sd_select_masks_dev  : per (chunk, speaker) choice between the clean and the raw mask (SD:3056-3078);
 *                        binarized[C][F][K] fp64 -> masks[C*K][F] fp32.
 * sd_mask_compact_file_dev : whole file, device resident: chunk c of the waveform is samples [c*step, c*step + L)
 *                        (zero beyond num_samples, SegmentModel::crop SD:1641); items are (chunk, speaker) pairs in
 *                        chunk-major order, wav_lens are normalised per consecutive `batch` items (32 in the reference),
 *                        batch_invalid[ceil(C*K/batch)] flags batches whose longest item is too short. */
int sd_mask_compact(sd_ctx* ctx, const float* wav, const float* masks, int B, int L, int F, int min_num_samples,
                    float* signals, float* wav_lens, uint8_t* too_short, int* all_too_short);
int sd_select_masks_dev(sd_ctx* ctx, const double* d_binarized, int C, int F, int K, double min_num_frames,
                        float* d_masks);
int sd_mask_compact_file_dev(sd_ctx* ctx, const float* d_wave, int64_t num_samples, int C, int K, int L,
                             int step_samples, const float* d_masks, int F, int batch, int min_num_samples,
                             float* d_signals, float* d_wav_lens, uint8_t* d_too_short, uint8_t* d_batch_invalid);

/* f2: reconstruct (SD:2789-2848) + to_diarization (SD:2638-2764) + crop_segment (SD:2568-2635).
 * segmentations[C][F][K] fp32, hard_clusters[C][K] (-2 = inactive), count[n_count] from sd_speaker_count and its
 * count_frames window -> binary discrete diarization out[rows][cols] fp64 (cols = max label + 1) and the window of
 * its rows.  sd_reconstruct_rows gives `rows` from the geometry alone (to size `out`). */
int sd_reconstruct_rows(int C, const sd_window* chunks, int64_t n_count, const sd_window* count_frames,
                        int64_t* rows, sd_window* frames_out);
int sd_reconstruct(sd_ctx* ctx, const float* segmentations, int C, int F, int K, const sd_window* chunks,
                   const int32_t* hard_clusters, const int32_t* count, int64_t n_count,
                   const sd_window* count_frames, double* out, int64_t cap_elems, int64_t* rows_out, int* cols_out,
                   sd_window* frames_out);
/* device pointers; cols (= max label + 1, at least 1) is supplied by the caller */
int sd_reconstruct_dev(sd_ctx* ctx, const float* d_segmentations, int C, int F, int K, const sd_window* chunks,
                       const int32_t* d_hard_clusters, int cols, const int32_t* d_count, int64_t n_count,
                       const sd_window* count_frames, double* d_out, int64_t cap_elems, int64_t* rows_out,
                       sd_window* frames_out);

/* f3: to_annotation (SD:2852-2935) + Track::support (SD:911-941) + Track::removeShort (SD:943-953) +
 * finalResult (SD:962-978).  scores[rows][cols] fp64 -> segments[n][2] (start, end seconds) ordered by start,
 * labels[n].  Segments with equal start are ordered by label (the reference's std::sort leaves it unspecified). */
int sd_to_annotation(sd_ctx* ctx, const double* scores, int64_t rows, int cols, const sd_window* frames,
                     double onset, double offset, double min_duration_on, double min_duration_off,
                     double* segments, int32_t* labels, int64_t cap, int64_t* n_out);
int sd_to_annotation_dev(sd_ctx* ctx, const double* d_scores, int64_t rows, int cols, const sd_window* frames,
                         double onset, double offset, double min_duration_on, double min_duration_off,
                         double* d_segments, int32_t* d_labels, int64_t cap, int64_t* n_out);

/* f4: ingest.
 * sd_ingest_pcm16     : int16 PCM -> float (frontend/wav.h:98-104) scaled by 1/32768 (SD:2948-2951).
 * sd_slide_geometry   : the chunk loop of SegmentModel::slide (SD:1407-1470): number of full windows and the start /
 *                       length (samples) of the shorter tail chunk (tail_start = -1 when there is none).
 * sd_crop_chunks(_dev): SegmentModel::crop (SD:1641-1662) for n_chunks windows starting at starts_s[] seconds (host
 *                       array): out[n_chunks][floor(duration*sample_rate)], zero padded outside the file. */
int sd_ingest_pcm16(sd_ctx* ctx, const int16_t* pcm, int64_t n, float* out);
int sd_ingest_pcm16_dev(sd_ctx* ctx, const int16_t* d_pcm, int64_t n, float* d_out);
int sd_slide_geometry(int64_t num_samples, double duration, double step, int64_t* full_chunks, int64_t* tail_start,
                      int64_t* tail_len);
int sd_crop_chunks(sd_ctx* ctx, const float* wave, int64_t num_samples, const double* starts_s, int n_chunks,
                   double duration, int sample_rate, float* out);
int sd_crop_chunks_dev(sd_ctx* ctx, const float* d_wave, int64_t num_samples, const double* starts_s, int n_chunks,
                       double duration, int sample_rate, float* d_out);

/* ---- batches of files ------------------------------------------------------------------
 * Files are independent units: a batch keeps `workers` files in flight on one GPU, each on its own sd_ctx (stream +
 * scratch) driven by a library-owned host thread that runs the per-file sequence of speakerDiarization()
 * (SD:2937-3234) over the hot path: STFT of the C*S embedding items -> binarize -> speaker_count -> clustering (with
 * the inactive-speaker mask) -> skip-average aggregate of the diarization scores.  A stage runs when both its input
 * and its output pointer are set (so one long file can be split by chunk range: STFT + binarize per range, then one
 * sd_file with `binarized` as an input for the stages that need every chunk; shard.split_chunk_range).  One thread submits; file i of a batch runs on worker i % workers, in order.
 * Pointers are HOST (pinned recommended; H2D/D2H inside, overlapped across the files in flight) or DEVICE.
 * Multi-GPU: one sd_batch per GPU (one process per GPU, or one batch per device in a process); files are assigned
 * to GPUs by the caller (shard.assign_files: longest-processing-time first). */
typedef struct sd_batch sd_batch;
typedef enum sd_batch_pointers { SD_BATCH_HOST = 0, SD_BATCH_DEVICE = 1 } sd_batch_pointers;
typedef struct sd_file {
    /* geometry */
    int C, F, S;       /* chunks, frames per chunk, local speakers */
    int L;             /* samples per embedding item */
    int D;             /* embedding dimension */
    int Kd;            /* columns of diar_scores (0 = none) */
    double onset;      /* binarisation threshold (SegmentModel: 0.4442333667381752, SD:1339) */
    sd_window chunks;  /* chunk window: start, step, duration, num_samples of the file */
    sd_window frames;  /* model frame window (step = duration = 0.016875, SD:2430-2432) */
    /* inputs */
    const float* wav_items;      /* [C*S][L] masked chunk signals */
    const float* segmentations;  /* [C][F][S] */
    const double* embeddings;    /* [C][S][D], NaN row = absent */
    const double* diar_scores;   /* [C][F][Kd], NaN = absent cluster */
    /* outputs */
    float* stft;       /* [C*S][T][201][2] */
    double* binarized; /* [C][F][S] */
    int32_t* count;    /* [count_cap] */
    int64_t count_cap;
    int32_t* hard;     /* [C][S] */
    double* diar;      /* [sd_aggregate_num_frames(C, chunks, frames)][Kd] */
    /* results, valid after sd_batch_wait */
    int64_t n_count;
    int64_t n_diar;
    int num_clusters;
    int status;        /* sd_status of this file */
    sd_window count_frames;
} sd_file;
int sd_batch_create(int device, int workers, sd_batch** out); /* = sd_batch_create_ex with cfg == NULL */
/* How the files in flight share the GPU (every field: -1 = automatic).
 *   stft_chain      k > 0: the STFT launch of a file waits (on the device, through an event) for the k-th most recent
 *                   STFT launch of the batch, so the persistent STFT grids run first in, first out instead of sharing the
 *                   SMs and finishing together; the merge loops that follow then start staggered and run under the other
 *                   files' STFTs.  0 = unordered.  Automatic: 1.
 *   linkage_cluster 1: merge loop on an 8-CTA cluster (lowest latency for one file: 8.7 ms at 1 683 embeddings);
 *                   0: merge loop in one CTA (11.6 ms, but it holds one SM's registers instead of eight, which is what
 *                   the bandwidth-bound STFTs of the other files lose: 413 k against 378 k audio-s/s at 16 files in
 *                   flight).  Automatic: 1 up to 8 workers, 0 beyond.
 *   narrow_sms      n > 0: partition the GPU with CUDA green contexts -- the STFT runs on (SMs - n) SMs, every other
 *                   stage of a file on the remaining n (a multiple of 8 on sm_100).  Isolates the stages from each other
 *                   (the STFT then runs at its stand-alone rate per SM); measured slower overall than sharing all SMs
 *                   for the benchmark mix, so automatic = 0 (no partition).  Falls back to 0 if the driver refuses.
 * Results are identical whatever the configuration.  The environment variables SDB_BATCH_STFT_CHAIN,
 * SDB_BATCH_LINKAGE_CLUSTER and SDB_BATCH_NARROW_SMS override the structure (experiments); SDB_BATCH_TRACE=<path>
 * writes a device-side timeline of every file (one %globaltimer stamp per stage) when the batch is destroyed. */
typedef struct sd_batch_config {
    int stft_chain;
    int linkage_cluster;
    int narrow_sms;
} sd_batch_config;
int sd_batch_create_ex(int device, int workers, const sd_batch_config* cfg, sd_batch** out);
int sd_batch_get_config(const sd_batch* b, sd_batch_config* out); /* the values in effect (narrow_sms as granted) */
void sd_batch_destroy(sd_batch* b);
int sd_batch_set_params(sd_batch* b, const sd_stft_params* stft, const sd_cluster_params* cluster); /* NULL = keep */
int sd_batch_workers(const sd_batch* b);
int64_t sd_batch_launch_count(const sd_batch* b); /* kernels of this library launched by all workers so far */
void* sd_batch_stream(sd_batch* b, int worker); /* cudaStream_t of a worker (to order the caller's own work) */
int sd_batch_submit(sd_batch* b, sd_file* files, int n, int pointers); /* returns at once; files must stay alive */
int sd_batch_wait(sd_batch* b); /* all submitted files done (streams synchronised); first error, 0 if none */
const char* sd_batch_last_error(const sd_batch* b);

/* ---- stage intermediates (WRITE_DATA builds of the reference) ------------------------------------------
 * The bodies this library replaces write their intermediates to /tmp/cpp_<stage>.txt when the reference is built with
 * WRITE_DATA (consumer: pipeline/script/verifyEveryStepResult.py:6-17).  The fused kernels never materialise them;
 * these entry points compute each one on the device so that the host shim can re-emit the dumps, following the
 * reference's own call structure (host pointers, synchronous).
 * sd_binarize_rows_stages    : `on`, `same_as`, `well_defined_idx` of binarize_ndarray (SD:1574-1633).  All [R][F];
 *                              well_defined_idx rows are -1 padded, *idx_cols = the width the reference writes
 *                              (longest row, Helper::wellDefinedIndex SD:623-651).
 * sd_trim_sum                : np.sum(trimmed, axis=-1) of speaker_count (SD:1701-1714) -> out[C][F - floor(F*left) -
 *                              floor(F*right)].
 * sd_mask_interpolate        : Helper::interpolate (SD:746-767): masks[B][F] -> imasks[B][L] (0/1) and counts[B] =
 *                              imasks.sum(dim=1), the wav_lens before normalisation (SD:2466-2476).
 * sd_clustered_segmentations : clusteredSegmentations of reconstruct (SD:2815-2838): out[C][F][cols], NaN where no
 *                              local speaker of the chunk belongs to the cluster.
 * sd_to_diarization          : to_diarization after its aggregate (SD:2672-2764): activations[n_frames][cols] on
 *                              window act_frames (the aggregate's post_frames, num_samples included) + count ->
 *                              out[rows][cols]; optional sorted_speakers[rows][cols] (SD:2720-2730) and
 *                              crop4 = {first activation row, rows, first count row, count rows} (crop_segment). */
int sd_binarize_rows_stages(sd_ctx* ctx, const double* scores, int R, int F, double onset, uint8_t* on,
                            int32_t* same_as, int32_t* well_defined_idx, int* idx_cols);
int sd_trim_sum(sd_ctx* ctx, const double* binarized, int C, int F, int K, double left, double right, double* out);
int sd_mask_interpolate(sd_ctx* ctx, const float* masks, int B, int F, int L, float threshold, uint8_t* imasks,
                        int32_t* counts);
int sd_clustered_segmentations(sd_ctx* ctx, const float* segmentations, int C, int F, int K,
                               const int32_t* hard_clusters, int cols, double* out);
int sd_to_diarization(sd_ctx* ctx, const double* activations, int64_t n_frames, int cols, const sd_window* act_frames,
                      const int32_t* count, int64_t n_count, const sd_window* count_frames, double* out,
                      int64_t cap_elems, int64_t* rows_out, sd_window* frames_out, int32_t* sorted_speakers,
                      int64_t* crop4);

#ifdef __cplusplus
}
#endif
#endif /* SDB200_H_ */
