/* Minimal stand-in for libccd's <ccd/ccd.h>; see vec3.h in this directory. */
#ifndef ORACLE_STUB_CCD_H_
#define ORACLE_STUB_CCD_H_
#include <ccd/vec3.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef void (*ccd_support_fn)(const void* obj, const ccd_vec3_t* dir, ccd_vec3_t* vec);
typedef void (*ccd_first_dir_fn)(const void* obj1, const void* obj2, ccd_vec3_t* dir);
typedef void (*ccd_center_fn)(const void* obj1, ccd_vec3_t* center);
struct _ccd_t {
  ccd_first_dir_fn first_dir;
  ccd_support_fn support1;
  ccd_support_fn support2;
  ccd_center_fn center1;
  ccd_center_fn center2;
  unsigned long max_iterations;
  ccd_real_t epa_tolerance;
  ccd_real_t mpr_tolerance;
  ccd_real_t dist_tolerance;
};
typedef struct _ccd_t ccd_t;
void ccdFirstDirDefault(const void* o1, const void* o2, ccd_vec3_t* dir);
int ccdMPRPenetration(const void* obj1, const void* obj2, const ccd_t* ccd,
                      ccd_real_t* depth, ccd_vec3_t* dir, ccd_vec3_t* pos);
#define CCD_INIT(ccd) do { \
    (ccd)->first_dir = ccdFirstDirDefault; (ccd)->support1 = 0; (ccd)->support2 = 0; \
    (ccd)->center1 = 0; (ccd)->center2 = 0; (ccd)->max_iterations = (unsigned long)-1; \
    (ccd)->epa_tolerance = 0.0001; (ccd)->mpr_tolerance = 0.0001; (ccd)->dist_tolerance = 1e-6; \
  } while (0)
#ifdef __cplusplus
}
#endif
#endif
