"""ctypes binding of libmjb.so (the C-ABI in include/mjb.h and include/mjb_modelio.h).

There is no fallback: if the CUDA library is missing or cannot be loaded, importing the product
API raises. Nothing here touches oracle/.
"""
import ctypes
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MJB_LIB") or os.path.join(_PKG, "lib", "libmjb.so")

_lib = None

c_void_p, c_int, c_char_p = ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p
c_ll, c_uint, c_double = ctypes.c_longlong, ctypes.c_uint, ctypes.c_double

# name -> (restype, argtypes); mirrors include/mjb.h and include/mjb_modelio.h one to one
SIGNATURES = {
    "mjb_makeData": (c_void_p, [c_void_p, c_int, c_int, c_uint, c_int, c_int, c_char_p, c_int]),
    "mjb_makeDataMulti": (c_void_p, [c_void_p, c_int, ctypes.POINTER(c_int), c_int, c_uint, c_int, c_int,
                                     c_char_p, c_int]),
    "mjb_ndevice": (c_int, [c_void_p]),
    "mjb_deleteData": (None, [c_void_p]),
    "mjb_setStream": (None, [c_void_p, c_void_p]),
    "mjb_setState": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mjb_setMocap": (c_int, [c_void_p, c_int, c_void_p, c_void_p]),
    "mjb_setXfrcApplied": (c_int, [c_void_p, c_int, c_void_p]),
    "mjb_setEqActive": (c_int, [c_void_p, c_int, c_void_p]),
    "mjb_setStateDevice": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_ll]),
    "mjb_inverse": (c_int, [c_void_p, c_void_p, c_int]),
    "mjb_inverseAsync": (c_int, [c_void_p, c_void_p, c_int]),
    "mjb_inverseFD": (c_int, [c_void_p, c_void_p, c_int, c_double, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mjb_inverseFDSensor": (c_int, [c_void_p, c_void_p, c_int, c_double, c_int] + [c_void_p] * 7),
    "mjb_compareFwdInv": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mjb_inverseSkip": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int]),
    "mjb_inverseHost": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mjb_get": (c_int, [c_void_p, c_int, c_void_p]),
    "mjb_lastBatch": (c_int, [c_void_p]),
    "mjb_getQfrcInverse": (c_int, [c_void_p, c_void_p]),
    "mjb_devicePtr": (c_void_p, [c_void_p, c_int]),
    "mjb_fieldRows": (c_int, [c_void_p, c_int]),
    "mjb_stride": (c_ll, [c_void_p]),
    "mjb_internalSlot": (c_int, [c_void_p, c_char_p, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "mjb_internalSize": (c_int, [c_void_p]),
    "mjb_ncandidate": (c_int, [c_void_p]),
    "mjb_candidate": (None, [c_void_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int),
                             ctypes.POINTER(c_int)]),
    "mjb_lastError": (c_char_p, [c_void_p]),
    "mjb_kernelLaunches": (c_ll, [c_void_p]),
    "mjb_debugQueue": (c_int, [c_void_p, ctypes.POINTER(c_int)]),
    "mjb_phaseTiming": (None, [c_void_p, c_int]),
    "mjb_phaseTimes": (c_int, [c_void_p, ctypes.POINTER(c_double), c_int]),
    "mjb_specialize": (c_int, [c_void_p, c_char_p, c_int]),
    "mjb_specialized": (c_int, [c_void_p]),
    "mjb_specializeInfo": (c_int, [c_void_p, c_char_p, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_double)]),
    "mjb_precompile": (c_int, [c_void_p, c_char_p, c_int]),
    "mjb_synchronize": (c_int, [c_void_p]),
    "mjb_fp64PeakTflops": (c_double, [c_int]),
    "mjb_loadModel": (c_void_p, [c_char_p, c_char_p, c_int]),
    "mjb_loadModelBuffer": (c_void_p, [c_void_p, c_ll, c_char_p, c_int]),
    "mjb_freeModel": (None, [c_void_p]),
    "mjb_modelInt": (c_int, [c_void_p, c_char_p, ctypes.POINTER(c_ll)]),
    "mjb_modelArray": (c_int, [c_void_p, c_char_p, ctypes.POINTER(c_void_p), ctypes.POINTER(c_int),
                               ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "mjb_modelOptInt": (ctypes.POINTER(c_int), [c_void_p, c_char_p]),
    "mjb_modelOptNum": (ctypes.POINTER(c_double), [c_void_p, c_char_p, ctypes.POINTER(c_int)]),
}


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} not found: build it with `python -m mujoco_inversedynamicstest_b200.build` "
                "(libmjb has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)       # AttributeError here means the library is stale
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
