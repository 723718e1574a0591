// tcgen05 / TMA fused transform path (sc_fast.cu): interface seen by the orchestration in sc_api.cu.
#pragma once
#include "sc_plan.h"

#define SC_STR_(x) #x
#define SC_STR(x) SC_STR_(x)

namespace sc {

bool fast_plan_init(Plan* p);       // builds the fast-path tables when the shape qualifies; false only on CUDA errors
void fast_plan_destroy(Plan* p);
bool fast_can_analyze(const Plan* p, bool adjoint);
bool fast_can_synthesize(const Plan* p, bool adjoint);
bool fast_analyze(const Plan* p, const float* images, int64_t n_images, float2* modes_out, bool adjoint,
                  cudaStream_t st);
bool fast_synthesize(const Plan* p, const float2* modes_in, int64_t n_images, int n_channels, const float* bias,
                     float* images_out, bool adjoint, cudaStream_t st);

bool umma_selftest(const float* A, const float* B, float* D, int N, int K, cudaStream_t st);

}  // namespace sc
