"""
semiclassical_b200 -- B200-native engine for the `dynamics` task of humeniuka/semiclassical.

    propagators   HermanKlukPropagator, WaltonManolopoulosPropagator (interface of semiclassical/propagators.py)
    potentials    NonHarmonicPotential, MorsePotential, MolecularHarmonicPotential, MolecularGDMLPotential
                  (interface of semiclassical/potentials.py)
    dynamics      run_semiclassical_dynamics(task): the JSON task of `semi dynamics` on fused launches
    distributed   sharding of an ensemble over the GPUs of a node and the all-reduce of the correlation functions
    workloads     the models of BASELINE.json (5-mode AS fixture, synthetic 60-mode AS, sGDML, harmonic molecule)

The numerics live in semiclassical_b200/lib/libsemiclassical_b200.so (C ABI: include/semiclassical_b200.h, sources:
semiclassical_b200/csrc); there is no CPU fallback.  Sub-modules are imported on demand (importing this package does
not load the CUDA library).
"""
__version__ = "0.1.0"
__all__ = ["propagators", "potentials", "dynamics", "distributed", "workloads", "units"]
