"""physical constants in atomic units; values are those of the reference (semiclassical/units.py:8-18),
parity of the time grid and of the fchk masses depends on them"""
hbar = 1.0
hartree_to_ev = 27.211396132
hartree_to_wavenumbers = 219474.63
bohr_to_angs = 0.529177249
autime_to_fs = 0.02418884326505
amu_to_aumass = 1822.888486192
