"""
ctypes binding of the C ABI (include/semiclassical_b200.h -> lib/libsemiclassical_b200.so).

There is no CPU fallback: importing the symbols fails loudly when the CUDA library is missing and cannot be
built, and every compute entry point needs a CUDA device.
"""
import ctypes
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB_PATH = os.path.join(HERE, "lib", "libsemiclassical_b200.so")
CSRC = os.path.join(HERE, "csrc")
HEADER = os.path.join(ROOT, "include", "semiclassical_b200.h")

SC_OK, SC_ERR_INVALID, SC_ERR_CUDA, SC_ERR_UNSUPPORTED = 0, 1, 2, 3

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC"]
# development only: SC_FAST_BUILD=1 adds -split-compile 0 (51 s instead of 155 s), which changes ptxas' register allocation --
# k_rk4_wcols then spills 360 bytes and the headline drops by 2-10 %; never used for measured builds
if os.environ.get("SC_FAST_BUILD") == "1":
    NVCC_FLAGS += ["-split-compile", "0"]

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p


class EngineConfig(ctypes.Structure):
    """mirror of sc_engine_config"""
    _fields_ = [("d", ctypes.c_int), ("dr", ctypes.c_int), ("wm", ctypes.c_int),
                ("L1", _dp), ("L2", _dp), ("R1", _dp), ("R2", _dp), ("U", _dp), ("q0", _dp), ("p0", _dp),
                ("oi0_A", _dp), ("oi0_B", _dp), ("oi0_C", _dp), ("oi0_fac", ctypes.c_double),
                ("ot0_A", _dp), ("ot0_B", _dp), ("ot0_C", _dp), ("ot0_fac", ctypes.c_double),
                ("Gamma_0", _dp), ("Gamma_i", _dp), ("Gamma_t", _dp), ("iGi0", _dp),
                ("alpha", ctypes.c_double), ("beta", ctypes.c_double), ("iGamma_0", _dp),
                ("detG0", ctypes.c_double), ("detGi", ctypes.c_double), ("detGt", ctypes.c_double),
                ("detGi0", ctypes.c_double)]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))) + [HEADER]


def needs_build():
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(s) > t for s in sources())


def build(force=False, verbose=False):
    """compile the CUDA library for sm_100a in-tree (nvcc cross-compiles without a GPU)"""
    if not (force or needs_build()):
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    # one builder at a time (torchrun starts every rank at once); the library is written next to its final place and
    # renamed into it, so that no process can dlopen a half-written file
    import fcntl
    with open(LIB_PATH + ".lock", "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)
        try:
            if not (force or needs_build()):      # another process built it while this one waited
                return LIB_PATH
            tmp = "%s.tmp.%d" % (LIB_PATH, os.getpid())
            cmd = [nvcc] + NVCC_FLAGS + ["-o", tmp, os.path.join(CSRC, "sc_engine.cu")]
            if verbose:
                print(" ".join(cmd))
            try:
                subprocess.check_call(cmd)
                os.replace(tmp, LIB_PATH)
            finally:
                if os.path.exists(tmp):
                    os.remove(tmp)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None

_SIGNATURES = {
    "sc_abi_version": (ctypes.c_int, []),
    "sc_last_error": (ctypes.c_char_p, []),
    "sc_potential_create_morse": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, _dp, _dp, _dp, ctypes.c_int, _dp]),
    "sc_potential_create_rotated_morse": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, _dp, _dp, _dp, ctypes.c_int, _dp, _dp]),
    "sc_potential_create_nonharmonic": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, _dp, _dp]),
    "sc_potential_create_harmonic": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, _dp, ctypes.c_double, _dp, _dp, _dp, _dp]),
    "sc_potential_create_gdml": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, _dp, _dp,
                                                ctypes.c_double, ctypes.c_double, ctypes.c_double, _dp, _dp]),
    "sc_potential_set_origin": (ctypes.c_int, [_vp, ctypes.c_double]),
    "sc_potential_dimensions": (ctypes.c_int, [_vp]),
    "sc_potential_destroy": (ctypes.c_int, [_vp]),
    "sc_potential_eval": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp]),
    "sc_engine_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.POINTER(EngineConfig)]),
    "sc_engine_destroy": (ctypes.c_int, [_vp]),
    "sc_engine_set_ensemble": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_longlong, _vp, _vp, _vp]),
    "sc_engine_sample_ensemble": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_longlong, ctypes.c_ulonglong, _dp, _dp, ctypes.c_double,
                                                 _vp, _vp, _vp]),
    "sc_engine_set_ensemble_host": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_longlong, _vp, _vp, _vp]),
    "sc_engine_step": (ctypes.c_int, [_vp, _vp, ctypes.c_double, ctypes.c_int, _vp, _vp]),
    "sc_engine_step_dev": (ctypes.c_int, [_vp, _vp, ctypes.c_double, ctypes.c_int, _vp, _vp]),
    "sc_engine_correlations": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "sc_engine_correlations_n1": (ctypes.c_int, [_vp, _vp, _vp, _vp]),
    "sc_engine_correlations_general": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sc_engine_stage_positions": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, _vp, _vp]),
    "sc_engine_stage_apply": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_double, _vp, _vp, _vp, _vp, _vp, _vp]),
    "sc_engine_stage_finish": (ctypes.c_int, [_vp, ctypes.c_double, _vp]),
    "sc_engine_get_state": (ctypes.c_int, [_vp, _vp, _vp]),
    "sc_engine_set_state": (ctypes.c_int, [_vp, _vp, _vp]),
    "sc_engine_get_prefactor": (ctypes.c_int, [_vp, _vp, _vp, _vp, _vp]),
    "sc_engine_num_trajectories": (ctypes.c_int, [_vp]),
    "sc_engine_coefficients": (ctypes.c_int, [_vp, _vp, _vp]),
    "sc_engine_norm": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_double, _vp, _vp]),
    "sc_engine_norm_pack_size": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(ctypes.c_longlong)]),
    "sc_engine_norm_pack": (ctypes.c_int, [_vp, _vp, _vp, _vp, ctypes.c_int, _vp, _vp]),
    "sc_engine_norm_block": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_int, _vp, _vp, ctypes.c_double, _vp, _vp]),
    "sc_engine_wavefunction": (ctypes.c_int, [_vp, _vp, ctypes.c_double, ctypes.c_int, _vp, _vp, _vp]),
    "sc_engine_launch_count": (ctypes.c_longlong, [_vp]),
    "sc_engine_kernel_name": (ctypes.c_char_p, [_vp]),
    "sc_measure_fp64_peak": (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    "sc_engine_set_option": (ctypes.c_int, [_vp, ctypes.c_char_p, ctypes.c_int]),
    "sc_engine_set_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "sc_engine_get_timing": (ctypes.c_int, [_vp, _vp]),
    "sc_engine_get_timing_slots": (ctypes.c_int, [_vp, _vp, ctypes.c_int]),
}


def exported_symbols():
    return sorted(_SIGNATURES)


def lib():
    """load (building if necessary) the native library; raises if that is impossible"""
    global _lib
    if _lib is None:
        if needs_build():
            try:
                build()
            except Exception as err:  # no silent fallback
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError("semiclassical_b200: the CUDA library %s is missing and could not be built (%s); "
                                       "there is no CPU fallback" % (LIB_PATH, err))
                # a library built from OLDER sources exists: using it silently would hide the failed build
                if os.environ.get("SC_ALLOW_STALE_LIBRARY") != "1":
                    raise RuntimeError("semiclassical_b200: %s is older than its sources and the rebuild failed (%s); "
                                       "fix the build or set SC_ALLOW_STALE_LIBRARY=1 to load the stale library" % (LIB_PATH, err))
                import warnings
                warnings.warn("semiclassical_b200: loading a STALE CUDA library (rebuild failed: %s)" % (err,), RuntimeWarning)
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        if L.sc_abi_version() != 1:
            raise RuntimeError("semiclassical_b200: ABI version mismatch")
        _lib = L
    return _lib


def check(rc):
    """map status codes to the exception types of the reference ('Errors')"""
    if rc == SC_OK:
        return
    msg = lib().sc_last_error().decode()
    if rc == SC_ERR_INVALID:
        raise AssertionError(msg)
    if rc == SC_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RuntimeError(msg)
