/* Link-time stand-ins for the three libccd symbols the reference engine leaves unresolved.
 * Reaching any of them means a test model left the supported path: abort loudly. */
#include <stdio.h>
#include <stdlib.h>
#include <ccd/ccd.h>
static ccd_vec3_t origin_ = {{0, 0, 0}};
ccd_vec3_t* ccd_vec3_origin = &origin_;
void ccdFirstDirDefault(const void* o1, const void* o2, ccd_vec3_t* dir) {
  (void)o1; (void)o2; ccdVec3Set(dir, 1, 0, 0);
}
int ccdMPRPenetration(const void* obj1, const void* obj2, const ccd_t* ccd,
                      ccd_real_t* depth, ccd_vec3_t* dir, ccd_vec3_t* pos) {
  (void)obj1; (void)obj2; (void)ccd; (void)depth; (void)dir; (void)pos;
  fprintf(stderr, "oracle/_ref: libccd MPR is not available in this build\n");
  abort();
}
