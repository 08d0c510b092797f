// TEST INFRASTRUCTURE ONLY -- see onnxruntime_cxx_api.h in this directory.
#pragma once
struct OrtStatus;
struct OrtSessionOptions;
inline OrtStatus* OrtSessionOptionsAppendExecutionProvider_CPU(OrtSessionOptions*, int) { return nullptr; }
