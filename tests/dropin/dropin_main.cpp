// TEST INFRASTRUCTURE.  Runs the reference pipeline -- speakerDiarization() (SD:2937-3234), the function the
// reference's main() calls -- on a wav file and prints the speaker segments the way that main() does (SD:3434-3441),
// with the two ONNX forward passes answered by the deterministic stand-ins of
// oracle/ref_harness/ort_stub/model_standins.h.  (The reference's main() itself cannot be called under another name:
// it has no return statement, SD:3418-3442, which is only legal for the real main; renamed, gcc -O2 lets it run off
// its end into whatever function follows.)
//
// Two binaries are built from this file (tests/dropin/Makefile):
//   dropin_reference  DROPIN_TU = the reference translation unit as it lies in /root/reference, unmodified
//   dropin_sdb200     DROPIN_TU = the same file after tests/dropin/patch_reference.py swapped the hot-path bodies
//                     for sdb200::* calls; links libsdb200.so, no libtorch, no clustering.cpp
// Both are compiled with -DWRITE_DATA, so every stage writes its /tmp/cpp_<stage>.txt dump
// (pipeline/script/verifyEveryStepResult.py:6-17); tests/dropin/compare.py checks that the two sets are identical.
//
//   dropin_xxx <wav> [capture_dir]
#include <malloc.h>

#define main reference_main
#include DROPIN_TU
#undef main

int main(int argc, char** argv) {
    if (argc < 2) {
        std::fprintf(stderr, "usage: %s <wav> [capture_dir]\n", argv[0]);
        return 2;
    }
    // SegmentModel::infer copies 32 rows out of `waveform` even when fewer chunks were passed (SD:1356-1364): on the
    // remainder and tail batches it reads vectors that were already freed.  Keep freed blocks mapped so that this
    // (harmless here: the stand-in ignores the values) read cannot fault.
    mallopt(M_MMAP_THRESHOLD, 1 << 30);
    mallopt(M_TRIM_THRESHOLD, 1 << 30);

    // chunk plan of SegmentModel::slide (SD:1407-1470) for this file: full 5 s windows every 0.5 s in batches of 32,
    // then one call for the tail chunk
    wav::WavReader probe(argv[1]);
    const long n = probe.num_samples(), window = 80000, step = 8000;
    long full = 0;
    for (long i = 0; i + window < n; i += step) ++full;
    ort_stub::Config& cfg = ort_stub::config();
    cfg.deterministic = true;
    for (long left = full; left > 0; left -= 32) cfg.seg_rows.push_back(left >= 32 ? 32 : (int)left);
    cfg.seg_rows.push_back(1);
    if (argc > 2) cfg.capture_dir = argv[2];

    auto res = speakerDiarization(argv[1], "segment2.onnx", "emd4.onnx");
    std::cout << "----------------------------------------------------" << std::endl;
    for (const auto& dr : res.finalResult())  // SD:3437-3440
        std::cout << "[" << dr.start << " -- " << dr.end << "]" << " --> Speaker_" << dr.label << std::endl;
    std::cout << "----------------------------------------------------" << std::endl;
    return 0;
}
