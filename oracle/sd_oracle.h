/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product library (libsdb200.so) never
 * links, imports or calls anything in oracle/.
 *
 * Every function cites the reference lines (relative to /root/reference) it
 * restates.  SD = pipeline/src/speakerDiarizer.cpp,
 * CL = pipeline/src/clustering/clustering.cpp.
 *
 * Parity pinning (see oracle/README.md, tests/test_oracle_*.py):
 *   - closest_frame / np_rint  : pipeline/src/test/closest_frame.txt (10 000 rows)
 *   - linkage / fcluster       : the 12-point toy of pipeline/src/clustering/cluster.cpp:8-13,
 *                                scipy 1.18.1, and oracle/_ref (the reference itself, compiled)
 *   - everything else          : oracle/_ref outputs + fixtures in tests/golden/
 *   - sdo_fbank_tail (a3)      : PARITY UNPINNED -- speechbrain 0.5.14 is not vendored in the
 *                                reference and emd4.onnx is missing; restated from
 *                                embeddings/threeModel.py:212-221,333-396 + published semantics.
 */
#ifndef SD_ORACLE_H_
#define SD_ORACLE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

int sdo_np_rint(double v);                                                        /* SD:260-272 */
long sdo_closest_frame(double sw_start, double sw_step, double sw_duration, double t); /* SD:1084-1090 */

void sdo_hamming_window_f32(int n, float* w); /* at::hamming_window(n) periodic, fp32; SD:2007 */

/* SD:1977-2036 + 1889-1917: out[B][T][n_fft/2+1][2] fp32, T = 1 + L/hop, centre zero padding. */
int sdo_stft(const float* wav, int B, int L, int n_fft, int hop, const float* window, float* out);
/* same, but only frames t0 <= t < t1 of item b (for spot checks at full size): out[(t1-t0)][bins][2] */
int sdo_stft_frames(const float* wav_item, int L, int n_fft, int hop, const float* window, int t0, int t1, float* out);

/* a3, embeddings/threeModel.py:212-221: |X|^2 -> mel(80) -> dB -> top_db clamp -> mean-norm. PARITY UNPINNED. */
void sdo_mel_matrix(int n_bins, int n_mels, double f_min, double f_max, int sample_rate, float* W /*[n_bins][n_mels]*/);
int sdo_fbank_tail(const float* stft, int B, int T, int n_bins, int n_mels, const float* W, const float* wav_lens,
                   float* out /*[B][T][n_mels]*/);

/* SD:1167-1311.  out[num_frames][K]; count_out / mask_out optional ([num_frames][K]).
 * post_out = {start, step, duration, num_samples}.  Returns num_frames or -1 (cap). */
long sdo_aggregate(const double* scores, int C, int F, int K, double sf_start, double sf_step, double sf_duration,
                   long sf_num_samples, double pf_step, double pf_duration, int hamming, double missing,
                   int skip_average, double epsilon, double* out, long cap_rows, double* count_out, double* mask_out,
                   double* post_out);

/* SD:1506-1639 (binarize_swf -> binarize_ndarray).  scores[C][F][K] fp32 -> out[C][F][K] in {0,1}. */
int sdo_binarize(const float* scores, int C, int F, int K, double onset, int initial_state, double* out);

/* SD:1742-1782.  Returns F'.  tw_out = {start, step, duration, num_samples}. */
long sdo_trim(const double* binarized, int C, int F, int K, double left, double right, double bt_start, double bt_step,
              double bt_duration, double* out, double* tw_out);

/* SD:1665-1738 (chunk_step / chunk_duration are SegmentModel::m_step / m_duration, 0.5 / 5.0). */
long sdo_speaker_count(const double* binarized, int C, int F, int K, double chunk_step, double chunk_duration,
                       double pf_step, double pf_duration, int* out, long cap, double* count_frames_out);

int sdo_clean_segmentations(const double* binarized, int C, int F, int K, double* out); /* SD:710-743 */

int sdo_normalize_embeddings(double* x, int N, int D); /* SD:330-357 */
int sdo_pdist(const double* x, int N, int D, double* out); /* CL:408-431 */
int sdo_linkage_condensed(const double* dists, int N, double* Z); /* CL:289-406 */
int sdo_linkage(const double* x, int N, int D, double* Z);       /* CL:417-440 */
int sdo_fcluster(const double* Z, int N, double cutoff, int* T);  /* CL:442-457, 121-232 */

/* SD:476-516.  Returns 2 where the reference throws (zero magnitude). */
int sdo_cosine_cdist(const double* a, int na, const double* b, int nb, int D, double* out);

/* SD:2300-2422 on filtered embeddings; threshold is the float member m_threshold (SD:2049). */
int sdo_cluster_labels(const double* x, int N, int D, float threshold, int min_cluster_size, int* labels);

/* SD:2063-2212 (+3166-3191 when binarized != NULL).  emb[C][S][D], NaN rows absent.
 * soft_out optional [C][S][num_clusters_cap]; returns status, *num_clusters_out = K. */
int sdo_clustering_stage(const double* emb, int C, int S, int D, float threshold, int min_cluster_size,
                         const double* binarized, int F, int* hard, double* soft_out, int soft_k_cap,
                         int* num_clusters_out);

/* ---- "next" rows ---- */
/* SD:746-797 + 2466-2510 */
int sdo_mask_compact(const float* wav, const float* masks, int B, int L, int F, int min_num_samples, float* signals,
                     float* wav_lens, unsigned char* too_short);
/* SD:2789-2848 + 2638-2764 + 2568-2635 */
long sdo_reconstruct(const float* segmentations, int C, int F, int K, double sf_start, double sf_step,
                     double sf_duration, long sf_num_samples, const int* hard, const int* count, long n_count,
                     double cf_start, double cf_step, double cf_duration, long cf_num_samples, double* out,
                     long cap_elems, int* cols_out, double* frames_out);
/* SD:2852-2935 + 911-941 + 962-978 */
long sdo_to_annotation(const double* scores, long rows, int cols, double f_start, double f_step, double f_duration,
                       double onset, double offset, double min_duration_on, double min_duration_off, double* seg_out,
                       int* label_out, long cap);

/* f4: ingest.  wav.h:98-104 (int16 -> float) + SD:2948-2951 (x * 1.0f / 32768.0, stored as float) */
void sdo_ingest_pcm16(const short* pcm, long n, float* out);
/* SegmentModel::crop, SD:1641-1662: window [floor(start*sr), +floor(duration*sr)) of the waveform, zero padded on
 * both sides; returns the chunk length */
long sdo_crop(const float* wave, long n, double start, double duration, int sample_rate, float* out);
/* the chunk loop of SegmentModel::slide, SD:1407-1470: number of full windows and the zero-padded tail chunk */
void sdo_slide_geometry(long num_samples, double duration, double step, long* full_chunks, long* tail_start,
                        long* tail_len);

#ifdef __cplusplus
}
#endif
#endif
