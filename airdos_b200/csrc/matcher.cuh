// matcher.cuh -- the matcher handle (shared by match.cu and search.cu).
#pragma once
#include "common.cuh"

struct adb_matcher {
    int device = 0;
    cudaStream_t stream = nullptr;
    // growable staging for the guided searches: one pinned host block and one device block per call
    uint8_t* d_scratch = nullptr;
    uint8_t* h_scratch = nullptr;
    size_t scratch_bytes = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    float last_ms = 0.f;
};
