// sc_wm.cuh -- Walton-Manolopoulos prefactor pieces and contributions (propagators.py:1132-1389, 1577-1719).
#pragma once
#include "sc_device.cuh"
struct DevPool;
namespace sc {
struct WMState {
  double *signA = nullptr, *signM = nullptr;
  double *scratch5 = nullptr;
};
}  // namespace sc
static int wm_setup(sc::WMState &, const sc_engine_config &, DevPool &) { return 1; }
static void wm_set_nac(sc::WMState &, const double *, int) {}
static int wm_alloc(sc::WMState &, DevPool &, const sc::EngDev &, cudaStream_t) { return 1; }
static int wm_prefactor_launch(sc::WMState &, const sc::EngDev &, int, cudaStream_t) { return 1; }
static int wm_corr_launch(sc::WMState &, const sc::EngDev &, const sc::PotDev &, double, double *, double *, cudaStream_t) { return 1; }
