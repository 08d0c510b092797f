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

#include <cuda.h>

#include <atomic>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <deque>
#include <mutex>
#include <thread>

struct sd_batch {
    struct Worker {
        sd_ctx* ctx = nullptr;
        cudaEvent_t ev_stft = nullptr;  // this worker's last STFT launch
        cudaEvent_t ev_hop = nullptr;   // wide -> narrow hand-over within a file
        cudaStream_t s_wide = nullptr, s_narrow = nullptr;  // streams of the two SM partitions (nullptr: no partition)
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
    // SDB_BATCH_TRACE=<path>: device-side timeline of every file (events on the worker's stream), written at destroy
    std::string trace_path;
    // STFT launches of the files in flight are chained on the device in submission order (stft_chain = k: launch n
    // waits for launch n - k): without it the persistent STFT grids of all files share the SMs, finish together, and
    // the latency-bound merge loops that follow run in a wave of their own instead of under the other files' STFTs
    int stft_chain = 1;
    int linkage_cluster = 1;  // merge-loop kernel of the worker contexts (sd_batch_config)
    // Spatial partition of the GPU (CUDA green contexts): the bandwidth-bound STFT runs on `wide_sms` SMs, everything
    // else of a file -- the latency-bound merge loop and the ~20 small kernels around it -- on the remaining
    // `narrow_sms`.  Without it the persistent STFT grids hold every SM's registers, so the small kernels of the other
    // files wait for an STFT to end before each of their launches (measured: 7 ms per file), and the merge-loop CTAs
    // take a resident slot away from the STFT on the SMs they land on.
    CUgreenCtx g_wide = nullptr, g_narrow = nullptr;
    int wide_sms = 0, narrow_sms = 0;
    std::mutex stft_mu;
    std::deque<cudaEvent_t> stft_recent;  // events of the last stft_chain STFT launches (guarded by stft_mu)
};

namespace sdb {
__global__ void trace_stamp_kernel(unsigned long long* slot) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    *slot = t;
}
void trace_stamp(sd_ctx* ctx, int tag) {
    if (!ctx->d_trace || (int)ctx->trace_tags.size() >= ctx->trace_cap) return;
    trace_stamp_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_trace + ctx->trace_tags.size());
    ctx->trace_tags.push_back(tag);
}
}  // namespace sdb


namespace {

// Driver entry points through the runtime (the library does not link libcuda: it must load on machines without a
// driver, where every entry point fails with SD_ERR_CUDA instead)
template <typename F>
bool driver_fn(const char* name, F* fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult st;
    if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess || !p) {
        cudaGetLastError();
        return false;
    }
    *fn = reinterpret_cast<F>(p);
    return true;
}

struct GreenApi {
    CUresult (*DeviceGet)(CUdevice*, int) = nullptr;
    CUresult (*DeviceGetDevResource)(CUdevice, CUdevResource*, CUdevResourceType) = nullptr;
    CUresult (*DevSmResourceSplitByCount)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int,
                                          unsigned int) = nullptr;
    CUresult (*DevResourceGenerateDesc)(CUdevResourceDesc*, CUdevResource*, unsigned int) = nullptr;
    CUresult (*GreenCtxCreate)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int) = nullptr;
    CUresult (*GreenCtxDestroy)(CUgreenCtx) = nullptr;
    CUresult (*GreenCtxStreamCreate)(CUstream*, CUgreenCtx, unsigned int, int) = nullptr;
    bool load() {
        return driver_fn("cuDeviceGet", &DeviceGet) && driver_fn("cuDeviceGetDevResource", &DeviceGetDevResource) &&
               driver_fn("cuDevSmResourceSplitByCount", &DevSmResourceSplitByCount) &&
               driver_fn("cuDevResourceGenerateDesc", &DevResourceGenerateDesc) &&
               driver_fn("cuGreenCtxCreate", &GreenCtxCreate) && driver_fn("cuGreenCtxDestroy", &GreenCtxDestroy) &&
               driver_fn("cuGreenCtxStreamCreate", &GreenCtxStreamCreate);
    }
};

// Splits the device's SMs into a narrow group of (at least) `narrow` SMs and the rest, one green context each, and
// gives every worker a stream in both.  Returns false (and leaves the batch unpartitioned) if anything is missing.
bool make_partition(sd_batch* b, int narrow) {
    GreenApi api;
    if (!api.load()) return false;
    CUdevice dev;
    CUdevResource all, small, rest;
    unsigned int groups = 1;
    if (api.DeviceGet(&dev, b->device) != CUDA_SUCCESS) return false;
    if (api.DeviceGetDevResource(dev, &all, CU_DEV_RESOURCE_TYPE_SM) != CUDA_SUCCESS) return false;
    if (api.DevSmResourceSplitByCount(&small, &groups, &all, &rest, 0, (unsigned)narrow) != CUDA_SUCCESS || groups != 1)
        return false;
    if (rest.sm.smCount < 8) return false;
    CUdevResourceDesc d_small, d_rest;
    if (api.DevResourceGenerateDesc(&d_small, &small, 1) != CUDA_SUCCESS) return false;
    if (api.DevResourceGenerateDesc(&d_rest, &rest, 1) != CUDA_SUCCESS) return false;
    if (api.GreenCtxCreate(&b->g_narrow, d_small, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) return false;
    if (api.GreenCtxCreate(&b->g_wide, d_rest, dev, CU_GREEN_CTX_DEFAULT_STREAM) != CUDA_SUCCESS) {
        api.GreenCtxDestroy(b->g_narrow);
        b->g_narrow = nullptr;
        return false;
    }
    b->narrow_sms = (int)small.sm.smCount;
    b->wide_sms = (int)rest.sm.smCount;
    for (auto& w : b->workers) {
        CUstream sw = nullptr, sn = nullptr;
        if (api.GreenCtxStreamCreate(&sw, b->g_wide, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS ||
            api.GreenCtxStreamCreate(&sn, b->g_narrow, CU_STREAM_NON_BLOCKING, 0) != CUDA_SUCCESS) {
            for (auto& v : b->workers) v.s_wide = v.s_narrow = nullptr;  // streams die with the green contexts
            api.GreenCtxDestroy(b->g_narrow);
            api.GreenCtxDestroy(b->g_wide);
            b->g_narrow = b->g_wide = nullptr;
            return false;
        }
        w.s_wide = (cudaStream_t)sw;
        w.s_narrow = (cudaStream_t)sn;
    }
    return true;
}

void destroy_partition(sd_batch* b) {
    if (!b->g_wide) return;
    GreenApi api;
    if (!api.load()) return;
    for (auto& w : b->workers) {
        if (w.s_wide) cudaStreamDestroy(w.s_wide);
        if (w.s_narrow) cudaStreamDestroy(w.s_narrow);
    }
    api.GreenCtxDestroy(b->g_wide);
    api.GreenCtxDestroy(b->g_narrow);
}

}  // namespace

namespace {

#define SD_TRY(call)                \
    do {                            \
        const int rc__ = (call);    \
        if (rc__ != SD_OK) return rc__; \
    } while (0)

int run_file(sd_batch* b, sd_batch::Worker& wk, sd_file* f, int pointers) {
    sd_ctx* ctx = wk.ctx;
    const bool dev = pointers == SD_BATCH_DEVICE;
    const int items = f->C * f->S;
    if (f->C <= 0 || f->S <= 0) return ctx->fail(SD_ERR_INVALID, "sd_batch: file with C = %d, S = %d", f->C, f->S);
    // timeline tags: 0 start, 1 STFT done, 2 binarize + count done, 3 / 4 merge loop begin / end (cluster.cu),
    // 5 clustering done, 6 end
    auto mark = [&](int tag) { sdb::trace_stamp(ctx, tag); };
    const bool split = wk.s_wide != nullptr;
    const int all_sms = ctx->num_sms;
    auto on_wide = [&] {
        if (split) {
            ctx->stream = wk.s_wide;
            ctx->num_sms = b->wide_sms;
        }
    };
    auto on_narrow = [&] {  // continues the file on the narrow partition, ordered after what the wide one was given
        if (split) {
            cudaEventRecord(wk.ev_hop, wk.s_wide);
            cudaStreamWaitEvent(wk.s_narrow, wk.ev_hop, 0);
            ctx->stream = wk.s_narrow;
            ctx->num_sms = b->narrow_sms;
        }
    };
    struct Restore {  // whatever happens, the context leaves with its own geometry
        sd_ctx* c;
        int sms;
        cudaStream_t st;
        ~Restore() {
            c->num_sms = sms;
            c->stream = st;
        }
    } restore{ctx, all_sms, ctx->stream};
    on_wide();
    mark(0);
    if (f->wav_items && f->stft) {
        sd_stft_params sp = b->sp;
        sp.pad_batch_to = 0;
        if (dev && b->stft_chain > 0 && wk.ev_stft) {
            std::lock_guard<std::mutex> lock(b->stft_mu);
            if ((int)b->stft_recent.size() >= b->stft_chain)
                cudaStreamWaitEvent(ctx->stream, b->stft_recent[b->stft_recent.size() - (size_t)b->stft_chain], 0);
            const int rc = sd_stft_dev(ctx, f->wav_items, items, f->L, &sp, f->stft);
            if (rc != SD_OK) return rc;
            cudaEventRecord(wk.ev_stft, ctx->stream);
            b->stft_recent.push_back(wk.ev_stft);
            while ((int)b->stft_recent.size() > b->stft_chain) b->stft_recent.pop_front();
        } else {
            // host pointers: the call pipelines its own copies and is bound by PCIe, not chained
            SD_TRY(dev ? sd_stft_dev(ctx, f->wav_items, items, f->L, &sp, f->stft)
                       : sd_stft(ctx, f->wav_items, items, f->L, &sp, f->stft));
        }
    }
    mark(1);
    on_narrow();
    if (f->segmentations && f->binarized)
        SD_TRY(dev ? sd_binarize_dev(ctx, f->segmentations, f->C, f->F, f->S, f->onset, 0, f->binarized)
                   : sd_binarize(ctx, f->segmentations, f->C, f->F, f->S, f->onset, 0, f->binarized));
    if (f->binarized && f->count)
        SD_TRY(dev ? sd_speaker_count_dev(ctx, f->binarized, f->C, f->F, f->S, &f->chunks, &f->frames, f->count,
                                          f->count_cap, &f->n_count, &f->count_frames)
                   : sd_speaker_count(ctx, f->binarized, f->C, f->F, f->S, &f->chunks, &f->frames, f->count, f->count_cap,
                                      &f->n_count, &f->count_frames));
    mark(2);
    if (f->embeddings && f->hard) {
        const double* bin = f->binarized;  // computed above or supplied by the caller (chunk-range splits); may be NULL
        SD_TRY(dev ? sd_clustering_dev(ctx, f->embeddings, f->C, f->S, f->D, &b->cp, bin, f->F, f->hard, nullptr, 0,
                                       &f->num_clusters)
                   : sd_clustering(ctx, f->embeddings, f->C, f->S, f->D, &b->cp, bin, f->F, f->hard, nullptr, 0,
                                   &f->num_clusters));
    }
    mark(5);
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
    mark(6);
    return sd_sync(ctx);  // the narrow stream: everything of this file precedes it
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
        const int rc = run_file(b, w, job.first, job.second);
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

int sd_batch_create(int device, int workers, sd_batch** out) { return sd_batch_create_ex(device, workers, nullptr, out); }

int sd_batch_create_ex(int device, int workers, const sd_batch_config* cfg, sd_batch** out) {
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
    sd_batch_config c{-1, -1, -1};
    if (cfg) c = *cfg;
    auto env_int = [](const char* name, int* v) {
        if (const char* e = std::getenv(name)) *v = std::atoi(e);
    };
    env_int("SDB_BATCH_STFT_CHAIN", &c.stft_chain);
    env_int("SDB_BATCH_LINKAGE_CLUSTER", &c.linkage_cluster);
    env_int("SDB_BATCH_NARROW_SMS", &c.narrow_sms);
    b->stft_chain = c.stft_chain < 0 ? 1 : c.stft_chain;
    b->linkage_cluster = c.linkage_cluster < 0 ? (workers <= 8 ? 1 : 0) : (c.linkage_cluster != 0);
    int narrow = c.narrow_sms < 0 ? 0 : c.narrow_sms;
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
    for (auto& w : b->workers) w.ctx->linkage_cluster = b->linkage_cluster;
    cudaSetDevice(device);
    for (auto& w : b->workers) {
        if (cudaEventCreateWithFlags(&w.ev_stft, cudaEventDisableTiming) != cudaSuccess) w.ev_stft = nullptr;
        if (cudaEventCreateWithFlags(&w.ev_hop, cudaEventDisableTiming) != cudaSuccess) w.ev_hop = nullptr;
    }
    if (narrow > 0 && workers > 1) {
        bool ok = true;
        for (auto& w : b->workers) ok = ok && w.ev_hop;
        if (!ok || !make_partition(b, narrow))
            std::fprintf(stderr, "sd_batch: SM partition (%d narrow SMs) unavailable, running unpartitioned\n", narrow);
    }
    if (const char* tp = std::getenv("SDB_BATCH_TRACE")) {
        b->trace_path = tp;
        cudaSetDevice(device);
        for (auto& w : b->workers) {
            w.ctx->trace_cap = 1 << 16;
            if (cudaMalloc(&w.ctx->d_trace, sizeof(unsigned long long) * w.ctx->trace_cap) != cudaSuccess) {
                w.ctx->d_trace = nullptr;
                w.ctx->trace_cap = 0;
            }
            w.ctx->trace_tags.reserve(w.ctx->trace_cap);
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
    if (!b->trace_path.empty()) {  // one line per file: worker + the seven stamps in ms since the first stamp
        cudaSetDevice(b->device);
        std::vector<std::vector<unsigned long long>> stamps;
        unsigned long long t0 = ~0ull;
        for (auto& w : b->workers) {
            std::vector<unsigned long long> v(w.ctx->trace_tags.size());
            if (!v.empty()) cudaMemcpy(v.data(), w.ctx->d_trace, v.size() * sizeof(v[0]), cudaMemcpyDeviceToHost);
            for (auto x : v) t0 = x < t0 ? x : t0;
            stamps.push_back(std::move(v));
        }
        if (FILE* fp = std::fopen(b->trace_path.c_str(), "w")) {
            std::fprintf(fp, "worker,start,stft_done,count_done,merge_begin,merge_end,cluster_done,end\n");
            for (size_t k = 0; k < b->workers.size(); ++k) {
                const auto& tags = b->workers[k].ctx->trace_tags;
                double t[7];
                for (size_t i = 0; i < tags.size(); ++i) {
                    if (tags[i] == 0)
                        for (auto& x : t) x = -1.0;
                    t[tags[i]] = (double)(stamps[k][i] - t0) * 1e-6;
                    if (tags[i] == 6)
                        std::fprintf(fp, "%d,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f,%.4f\n", (int)k, t[0], t[1], t[2], t[3], t[4], t[5],
                                     t[6]);
                }
            }
            std::fclose(fp);
        }
        for (auto& w : b->workers) {
            cudaFree(w.ctx->d_trace);
            w.ctx->d_trace = nullptr;
        }
    }
    destroy_partition(b);
    for (auto& w : b->workers) {
        if (w.ev_stft) cudaEventDestroy(w.ev_stft);
        if (w.ev_hop) cudaEventDestroy(w.ev_hop);
        sd_ctx_destroy(w.ctx);
    }
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

int sd_batch_get_config(const sd_batch* b, sd_batch_config* out) {
    if (!b || !out) return SD_ERR_INVALID;
    out->stft_chain = b->stft_chain;
    out->linkage_cluster = b->linkage_cluster;
    out->narrow_sms = b->narrow_sms;
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
