// tcgen05 / TMA engine (placeholder until the tensor-core kernel lands in this file).
#pragma once
#include "common.cuh"

namespace vqb200 {

inline bool tc_supported(const RowLayout&, int, int) { return false; }
inline int tc_prepare_codebook(const CodebookImage&, int, int, cudaStream_t) { return 0; }
inline int tc_forward(const float*, const RowLayout&, int, int, const CodebookImage&, float*, int64_t*,
                      const ForwardScratch&, float*, float*, cudaStream_t) { return 1; }

}  // namespace vqb200
