/* Minimal stand-in for libccd's <ccd/vec3.h> (libccd v2.1 is not in this image).
 * Test infrastructure: only the types/inlines the reference engine names.
 * libccd is reached only for hfields or with mjDSBL_NATIVECCD (engine_collision_convex.c:37,794,822-836),
 * neither of which is on the mj_inverse path we check. */
#ifndef ORACLE_STUB_CCD_VEC3_H_
#define ORACLE_STUB_CCD_VEC3_H_
#ifdef __cplusplus
extern "C" {
#endif
typedef double ccd_real_t;
typedef struct { ccd_real_t v[3]; } ccd_vec3_t;
extern ccd_vec3_t* ccd_vec3_origin;
static inline void ccdVec3Set(ccd_vec3_t* a, ccd_real_t x, ccd_real_t y, ccd_real_t z) {
  a->v[0] = x; a->v[1] = y; a->v[2] = z;
}
static inline int ccdVec3Eq(const ccd_vec3_t* a, const ccd_vec3_t* b) {
  return a->v[0] == b->v[0] && a->v[1] == b->v[1] && a->v[2] == b->v[2];
}
#ifdef __cplusplus
}
#endif
#endif
