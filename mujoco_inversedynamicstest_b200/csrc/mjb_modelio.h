// Model ingestion helpers exported by libmjb (see mjb_modelio.cc); declared for C callers in
// include/mjb_modelio.h.
#ifndef MJB_MODELIO_INTERNAL_H_
#define MJB_MODELIO_INTERNAL_H_
#include "../../include/mjb_modelio.h"
#endif
