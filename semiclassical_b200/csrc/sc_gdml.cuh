// sc_gdml.cuh -- sGDML energy / gradient / Hessian (gdml_predictor.py:140-250).  Placeholder until the
// dedicated kernel lands: reports "unsupported".
#pragma once
#include "sc_device.cuh"
namespace sc {
static int launch_gdml_eval(const PotDev &, int, const double *, double *, double *, double *, cudaStream_t) { return 1; }
}  // namespace sc
