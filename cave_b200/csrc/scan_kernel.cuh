#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cave {

struct ScanParams {
    const float* A;
    const int* m_rows;      // nullable
    int B, m_max, d;
    int R, stages;          // filled by launch_scan
    size_t stage_stride;    // filled by launch_scan
    int64_t dpad;
    int *nvalid, *navg, *ngen, *gennnz, *nsingc;
    int2* gen;
    unsigned char* ctype;
    float* avg;
};

cudaError_t launch_scan(const ScanParams& p, cudaStream_t stream);

}  // namespace cave
