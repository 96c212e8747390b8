// api_util.cuh -- error reporting, launch accounting and per-kernel event timing shared by
// the C ABI translation units.
#pragma once
#include <cuda_runtime.h>

int b200r_fail(int code, const char* fmt, ...);            // records message, returns code
int b200r_cuda_fail(cudaError_t e, const char* what);      // records message, returns (int)e
void b200r_count_launch(void);

#include "../../include/b200raster.h"  // kernel ids B200R_K_* for b200r_profile_read
#define B200R_K_COUNT 24

// When profiling is enabled (b200r_profile_enable(1)) each launch is bracketed by a pair of
// CUDA events recorded on the launch stream; otherwise these are no-ops.
void b200r_prof_begin(int kernel, cudaStream_t st);
void b200r_prof_end(int kernel, cudaStream_t st);

struct B200rProfScope {
    int k; cudaStream_t st;
    B200rProfScope(int kernel, cudaStream_t s) : k(kernel), st(s) { b200r_prof_begin(k, st); }
    ~B200rProfScope() { b200r_prof_end(k, st); b200r_count_launch(); }
};
