"""mjb: batched MuJoCo 3.3.1 mj_inverse on NVIDIA B200 (hand-written sm_100a CUDA behind a C-ABI).

The product is csrc/ (kernels + C-ABI, built into lib/libmjb.so); this package is the thin Python
host mirror used by the tests and the benchmark. Importing it requires the built CUDA library:
there is no CPU path.
"""
from .batch import (BatchData, MjbError, Model, fp64_peak_tflops, OUT_QFRC, OUT_COUNTS, OUT_CONTACT,
                    OUT_EFC, OUT_INERTIA, OUT_INTERNAL, OUT_RNEPOST, OUT_CAMLIGHT, OUT_TRANSMISSION, F_QFRC_INVERSE, F_QFRC_CONSTRAINT,
                    F_QFRC_PASSIVE, F_COUNTS, F_STATUS, F_CONTACT_GEOM, F_CONTACT_INFO, F_CONTACT_NUM,
                    F_EFC_INT, F_EFC_NUM, F_QM, F_QLD, F_QLDIAGINV, F_INTERNAL, F_CACC, F_CFRC_INT,
                    F_CFRC_EXT, F_SENSORDATA, F_QFRC_BIAS, F_ENERGY, F_CAM_XPOS, F_CAM_XMAT, F_LIGHT_XPOS, F_LIGHT_XDIR,
                    F_ACTUATOR_LENGTH, F_ACTUATOR_MOMENT, F_ACTUATOR_VELOCITY, STATUS_BADQPOS, STATUS_BADQVEL, STATUS_BADQACC,
                    STATUS_CONTACTFULL, STATUS_CNSTRFULL)
from .states import SEED, generate_states

__all__ = [n for n in dir() if not n.startswith("_")]
