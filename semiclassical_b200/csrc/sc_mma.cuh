// sc_mma.cuh -- FP64 tensor-core (DMMA.8x8x4) variant of the fused Herman-Kluk step for large d.
#pragma once
#include "sc_kernels.cuh"
namespace sc {
static bool mma_supported(int) { return false; }
static void mma_leading_dims(int d, int &ldu, int &ldh) { ldu = 2 * d; ldh = d; }
static int mma_threads(int) { return 320; }
static cudaError_t launch_mma(int, int, size_t, const EngDev &, const PotDev &, double, int, double *, const SmemLayout &,
                              cudaStream_t) { return cudaErrorNotSupported; }
}  // namespace sc
