// TEST INFRASTRUCTURE ONLY -- not part of the product.
//
// Minimal stand-in for the ONNX Runtime C++ API surface that the reference
// translation unit (pipeline/src/speakerDiarizer.cpp + onnxModel/onnx_model.cc)
// touches.  onnxruntime is not installed in this image and the model blobs are
// missing from the reference checkout, so the network forward passes cannot run;
// this stub lets the *unmodified* reference sources compile so that the hot-path
// functions around the forward passes can be executed as the parity oracle.
//
// Session::Run does not run a model.  It records the input tensors it was
// handed (that is exactly what the reference would feed to emd4.onnx /
// segment2.onnx) into ort_stub::captured and returns zero tensors of the
// documented output shapes -- or, when ort_stub::config().deterministic is set,
// the input-independent stand-in outputs of model_standins.h, which lets the
// whole pipeline run end to end (tests/dropin).
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "model_standins.h"

struct OrtStatus;
struct OrtSessionOptions {};

enum OrtLoggingLevel { ORT_LOGGING_LEVEL_VERBOSE, ORT_LOGGING_LEVEL_INFO, ORT_LOGGING_LEVEL_WARNING };
enum OrtAllocatorType { OrtInvalidAllocator = -1, OrtDeviceAllocator = 0, OrtArenaAllocator = 1 };
enum OrtMemType { OrtMemTypeCPUInput = -2, OrtMemTypeCPUOutput = -1, OrtMemTypeCPU = OrtMemTypeCPUOutput, OrtMemTypeDefault = 0 };

namespace ort_stub {
struct Captured {
    std::vector<int64_t> shape;
    std::vector<float> data;
};
// one entry per input of the most recent Session::Run call
inline std::vector<Captured>& captured() {
    static std::vector<Captured> c;
    return c;
}
// Whole-pipeline mode (tests/dropin): set by the harness before the pipeline runs.
struct Config {
    bool deterministic = false;   // answer with model_standins.h instead of zeros
    std::vector<int> seg_rows;    // rows the caller keeps from each segmentation call (SegmentModel::infer always
                                  // submits 32); gives every row its chunk number
    std::string capture_dir;      // when set, every embedding-model input goes to <dir>/emb_input_NN.f32 (+ lens)
    int seg_calls = 0, emb_calls = 0;
};
inline Config& config() {
    static Config c;
    return c;
}
inline void dump_f32(const std::string& path, const float* p, size_t n) {
    if (FILE* f = std::fopen(path.c_str(), "wb")) {
        std::fwrite(p, sizeof(float), n, f);
        std::fclose(f);
    }
}
}  // namespace ort_stub

namespace Ort {

struct Env {
    Env(OrtLoggingLevel, const char*) {}
};

struct SessionOptions {
    OrtSessionOptions raw;
    void SetIntraOpNumThreads(int) {}
    operator OrtSessionOptions*() { return &raw; }
};

struct MemoryInfo {
    static MemoryInfo CreateCpu(OrtAllocatorType, OrtMemType) { return MemoryInfo(); }
};

struct RunOptions {
    RunOptions(std::nullptr_t) {}
};

struct TensorTypeAndShapeInfo {
    std::vector<int64_t> shape;
    std::vector<int64_t> GetShape() const { return shape; }
};

struct Value {
    std::vector<int64_t> shape;
    const float* ext = nullptr;   // borrowed (inputs)
    std::vector<float> own;       // owned (outputs)
    size_t count = 0;

    template <typename T>
    static Value CreateTensor(const MemoryInfo&, T* p, size_t n, const int64_t* dims, size_t ndims) {
        Value v;
        v.ext = reinterpret_cast<const float*>(p);
        v.count = n;
        v.shape.assign(dims, dims + ndims);
        return v;
    }
    template <typename T>
    const T* GetTensorData() const {
        return reinterpret_cast<const T*>(ext ? ext : own.data());
    }
    TensorTypeAndShapeInfo GetTensorTypeAndShapeInfo() const { return TensorTypeAndShapeInfo{shape}; }
};

struct AllocatorWithDefaultOptions {};

using AllocatedStringPtr = std::unique_ptr<char[]>;

struct Session {
    std::string path;
    Session(Env&, const char* model_path, SessionOptions&) : path(model_path) {}
    // The embedding model (emd4.onnx) takes (feats, wav_lens); the segmentation
    // model takes one waveform tensor.  The stub decides by file name.
    bool is_embedding() const { return path.find("emd") != std::string::npos || path.find("emb") != std::string::npos; }
    size_t GetInputCount() const { return is_embedding() ? 2 : 1; }
    size_t GetOutputCount() const { return 1; }
    static AllocatedStringPtr dup(const char* s) {
        AllocatedStringPtr p(new char[strlen(s) + 1]);
        strcpy(p.get(), s);
        return p;
    }
    AllocatedStringPtr GetInputNameAllocated(size_t i, AllocatorWithDefaultOptions&) const {
        return dup(i == 0 ? "input0" : "input1");
    }
    AllocatedStringPtr GetOutputNameAllocated(size_t, AllocatorWithDefaultOptions&) const { return dup("output0"); }

    std::vector<Value> Run(const RunOptions&, const char* const*, const Value* inputs, size_t n_in,
                           const char* const*, size_t) {
        auto& cap = ort_stub::captured();
        cap.clear();
        for (size_t i = 0; i < n_in; ++i) {
            ort_stub::Captured c;
            c.shape = inputs[i].shape;
            const float* p = inputs[i].GetTensorData<float>();
            c.data.assign(p, p + inputs[i].count);
            cap.push_back(std::move(c));
        }
        Value out;
        int64_t batch = n_in ? inputs[0].shape[0] : 1;
        if (is_embedding())
            out.shape = {batch, 1, 192};
        else
            out.shape = {batch, 293, 3};
        size_t n = 1;
        for (auto d : out.shape) n *= static_cast<size_t>(d);
        out.own.assign(n, 0.0f);
        out.count = n;
        ort_stub::Config& cfg = ort_stub::config();
        if (is_embedding()) {
            if (!cfg.capture_dir.empty() && n_in == 2) {
                char name[64];
                std::snprintf(name, sizeof(name), "/emb_input_%02d.f32", cfg.emb_calls);
                ort_stub::dump_f32(cfg.capture_dir + name, cap[0].data.data(), cap[0].data.size());
                std::snprintf(name, sizeof(name), "/emb_lens_%02d.f32", cfg.emb_calls);
                ort_stub::dump_f32(cfg.capture_dir + name, cap[1].data.data(), cap[1].data.size());
            }
            if (cfg.deterministic)
                for (int64_t r = 0; r < batch; ++r)
                    ort_stub::standin::embedding(cfg.emb_calls * static_cast<int>(batch) + static_cast<int>(r),
                                                 out.own.data() + r * 192);
            cfg.emb_calls++;
        } else {
            if (cfg.deterministic) {
                int first = 0;  // chunk number of row 0 = rows kept from the earlier calls
                for (int i = 0; i < cfg.seg_calls; ++i)
                    first += i < static_cast<int>(cfg.seg_rows.size()) ? cfg.seg_rows[i] : static_cast<int>(batch);
                for (int64_t r = 0; r < batch; ++r)
                    ort_stub::standin::segmentation(first + static_cast<int>(r), out.own.data() + r * 293 * 3);
            }
            cfg.seg_calls++;
        }
        std::vector<Value> res;
        res.push_back(std::move(out));
        return res;
    }
};

inline std::vector<std::string> GetAvailableProviders() { return {"CPUExecutionProvider"}; }

}  // namespace Ort

inline OrtStatus* OrtSessionOptionsAppendExecutionProvider_CUDA(OrtSessionOptions*, int) { return nullptr; }
