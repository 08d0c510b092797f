// Batches of files (SURVEY 8b `sd_batch_run`, 8e): files are independent units, so a batch is a set of files in
// flight on one GPU -- each worker owns an sd_ctx (own stream and scratch buffers) and a host thread that runs the
// per-file sequence of the hot path
//     STFT of every (chunk, speaker) item -> binarize -> speaker_count -> clustering (+ inactive mask) -> diarization
//     aggregate
// exactly as speakerDiarization() orders it (speakerDiarizer.cpp:2937-3234), through the same entry points a
// single-file caller uses.  Why threads and not one stream-ordered submission: the merge loop of the clustering is
// latency-bound on 8 SMs and the calls around it read two small results back (row validity, cluster count); with one
// file per worker those waits overlap with the other files' bandwidth-bound kernels and PCIe copies, which is what
// fills the GPU.  File i always runs on worker i % workers, in submission order, so the same
// sd_file may be submitted again (next step of a benchmark, next pass of a stream) without racing with itself.
#include "common.cuh"

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>

struct sd_batch {
    struct Worker {
        sd_ctx* ctx = nullptr;
        std::thread th;
        std::deque<std::pair<sd_file*, int>> q;  // (file, pointers)
    };
    int device = 0;
    std::vector<Worker> workers;
    std::mutex mu;
    std::condition_variable cv_work, cv_done;
    long pending = 0;
    bool stop = false;
    int first_error = SD_OK;
    std::string err;
    sd_stft_params sp;
    sd_cluster_params cp;
    long next = 0;  // files submitted so far (round-robin position)
};

namespace {

#define SD_TRY(call)                \
    do {                            \
        const int rc__ = (call);    \
        if (rc__ != SD_OK) return rc__; \
    } while (0)

int run_file(sd_batch* b, sd_ctx* ctx, sd_file* f, int pointers) {
    const bool dev = pointers == SD_BATCH_DEVICE;
    const int items = f->C * f->S;
    if (f->C <= 0 || f->S <= 0) return ctx->fail(SD_ERR_INVALID, "sd_batch: file with C = %d, S = %d", f->C, f->S);
    if (f->wav_items && f->stft) {
        sd_stft_params sp = b->sp;
        sp.pad_batch_to = 0;
        SD_TRY(dev ? sd_stft_dev(ctx, f->wav_items, items, f->L, &sp, f->stft)
                   : sd_stft(ctx, f->wav_items, items, f->L, &sp, f->stft));
    }
    if (f->segmentations && f->binarized)
        SD_TRY(dev ? sd_binarize_dev(ctx, f->segmentations, f->C, f->F, f->S, f->onset, 0, f->binarized)
                   : sd_binarize(ctx, f->segmentations, f->C, f->F, f->S, f->onset, 0, f->binarized));
    if (f->binarized && f->count)
        SD_TRY(dev ? sd_speaker_count_dev(ctx, f->binarized, f->C, f->F, f->S, &f->chunks, &f->frames, f->count,
                                          f->count_cap, &f->n_count, &f->count_frames)
                   : sd_speaker_count(ctx, f->binarized, f->C, f->F, f->S, &f->chunks, &f->frames, f->count, f->count_cap,
                                      &f->n_count, &f->count_frames));
    if (f->embeddings && f->hard) {
        const double* bin = f->binarized;  // computed above or supplied by the caller (chunk-range splits); may be NULL
        SD_TRY(dev ? sd_clustering_dev(ctx, f->embeddings, f->C, f->S, f->D, &b->cp, bin, f->F, f->hard, nullptr, 0,
                                       &f->num_clusters)
                   : sd_clustering(ctx, f->embeddings, f->C, f->S, f->D, &b->cp, bin, f->F, f->hard, nullptr, 0,
                                   &f->num_clusters));
    }
    if (f->diar_scores && f->diar && f->Kd > 0) {
        sd_window post;
        int64_t nf = 0;
        const int64_t cap = sd_aggregate_num_frames(f->C, &f->chunks, &f->frames);
        SD_TRY(dev ? sd_aggregate_dev(ctx, f->diar_scores, f->C, f->F, f->Kd, &f->chunks, &f->frames, 0, 0.0, 1,
                                      2.220446049250313e-16, f->diar, cap, &nf, &post, nullptr, nullptr)
                   : sd_aggregate(ctx, f->diar_scores, f->C, f->F, f->Kd, &f->chunks, &f->frames, 0, 0.0, 1,
                                  2.220446049250313e-16, f->diar, cap, &nf, &post, nullptr, nullptr));
        f->n_diar = nf;
    }
    return sd_sync(ctx);
}

void worker_loop(sd_batch* b, int index) {
    cudaSetDevice(b->device);
    sd_batch::Worker& w = b->workers[index];
    for (;;) {
        std::pair<sd_file*, int> job;
        {
            std::unique_lock<std::mutex> lock(b->mu);
            b->cv_work.wait(lock, [&] { return b->stop || !w.q.empty(); });
            if (w.q.empty()) return;  // stop requested and nothing left
            job = w.q.front();
            w.q.pop_front();
        }
        const int rc = run_file(b, w.ctx, job.first, job.second);
        job.first->status = rc;
        {
            std::lock_guard<std::mutex> lock(b->mu);
            if (rc != SD_OK && b->first_error == SD_OK) {
                b->first_error = rc;
                b->err = sd_last_error(w.ctx);
            }
            if (--b->pending == 0) b->cv_done.notify_all();
        }
    }
}

}  // namespace

extern "C" {

int sd_batch_create(int device, int workers, sd_batch** out) {
    if (!out || workers < 1 || workers > 64) return SD_ERR_INVALID;
    *out = nullptr;
    // Streams beyond the number of hardware work queues (8 by default) share a queue and serialise behind each
    // other's 9 ms merge loops.  Takes effect only if CUDA has not been initialised in this process yet; otherwise
    // the caller must export it.  Never overrides a value the user has set.
    if (workers > 8) setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    sd_batch* b = new sd_batch();
    b->device = device;
    sd_stft_default_params(&b->sp);
    sd_cluster_default_params(&b->cp);
    b->workers.resize((size_t)workers);
    for (int i = 0; i < workers; ++i) {
        const int rc = sd_ctx_create(device, &b->workers[i].ctx);
        if (rc != SD_OK) {
            for (int j = 0; j < i; ++j) sd_ctx_destroy(b->workers[j].ctx);
            delete b;
            return rc;  // no device: no CPU fallback
        }
    }
    // experiment hooks: SDB_BATCH_OPTS="option=value,..." is applied to every worker context (sd_ctx_set_option)
    if (const char* opts = std::getenv("SDB_BATCH_OPTS")) {
        std::string s(opts);
        size_t pos = 0;
        while (pos < s.size()) {
            const size_t comma = s.find(',', pos), eq = s.find('=', pos);
            const size_t end = comma == std::string::npos ? s.size() : comma;
            if (eq != std::string::npos && eq < end) {
                const int o = std::atoi(s.substr(pos, eq - pos).c_str()), v = std::atoi(s.substr(eq + 1, end - eq - 1).c_str());
                for (auto& w : b->workers) sd_ctx_set_option(w.ctx, o, v);
            }
            pos = end + 1;
        }
    }
    for (int i = 0; i < workers; ++i) b->workers[i].th = std::thread(worker_loop, b, i);
    *out = b;
    return SD_OK;
}

void sd_batch_destroy(sd_batch* b) {
    if (!b) return;
    {
        std::lock_guard<std::mutex> lock(b->mu);
        b->stop = true;
    }
    b->cv_work.notify_all();
    for (auto& w : b->workers)
        if (w.th.joinable()) w.th.join();
    for (auto& w : b->workers) sd_ctx_destroy(w.ctx);
    delete b;
}

int sd_batch_set_params(sd_batch* b, const sd_stft_params* sp, const sd_cluster_params* cp) {
    if (!b) return SD_ERR_INVALID;
    std::lock_guard<std::mutex> lock(b->mu);
    if (b->pending) return SD_ERR_INVALID;  // only between batches
    if (sp) b->sp = *sp;
    if (cp) b->cp = *cp;
    return SD_OK;
}

int sd_batch_workers(const sd_batch* b) { return b ? (int)b->workers.size() : 0; }

int64_t sd_batch_launch_count(const sd_batch* b) {
    int64_t n = 0;
    if (b)
        for (const auto& w : b->workers) n += sd_launch_count(w.ctx);
    return n;
}

void* sd_batch_stream(sd_batch* b, int worker) {
    if (!b || worker < 0 || worker >= (int)b->workers.size()) return nullptr;
    return sd_ctx_stream(b->workers[(size_t)worker].ctx);
}

int sd_batch_submit(sd_batch* b, sd_file* files, int n, int pointers) {
    if (!b || (n > 0 && !files) || n < 0 || (pointers != SD_BATCH_HOST && pointers != SD_BATCH_DEVICE))
        return SD_ERR_INVALID;
    {
        std::lock_guard<std::mutex> lock(b->mu);
        for (int i = 0; i < n; ++i) {
            files[i].status = SD_OK;
            b->workers[(size_t)((b->next + i) % (long)b->workers.size())].q.emplace_back(&files[i], pointers);
        }
        b->next += n;
        b->pending += n;
    }
    b->cv_work.notify_all();
    return SD_OK;
}

int sd_batch_wait(sd_batch* b) {
    if (!b) return SD_ERR_INVALID;
    std::unique_lock<std::mutex> lock(b->mu);
    b->cv_done.wait(lock, [&] { return b->pending == 0; });
    const int rc = b->first_error;
    b->first_error = SD_OK;
    b->next = 0;  // the next batch starts at worker 0 again: file i of every batch runs on worker i % workers
    return rc;
}

const char* sd_batch_last_error(const sd_batch* b) { return b ? b->err.c_str() : "null batch"; }

}  // extern "C"
