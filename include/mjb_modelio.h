/* mjb_modelio.h -- model ingestion for libmjb without the XML compiler in-process.
 *
 * Replaces, for hosts that only hold a serialized model, the reference's
 *   mjModel* mj_loadModel(const char* filename, const mjVFS* vfs);   include/mujoco/mujoco.h
 *   (format written by mj_saveModel, src/engine/engine_io.c:720-772)
 * and gives name-based read access to mjModel arrays for tools written in languages that do not
 * parse the C struct (the Python host mirror uses it to generate synthetic states).
 */
#ifndef MJB_MODELIO_H_
#define MJB_MODELIO_H_

#include <mujoco/mujoco.h>

#if defined(__cplusplus)
extern "C" {
#endif

#ifndef MJB_API
#define MJB_API __attribute__((visibility("default")))
#endif

/* read an MJB file saved by MuJoCo 3.3.1's mj_saveModel; NULL + message in err on failure */
MJB_API mjModel* mjb_loadModel(const char* path, char* err, int err_sz);
/* same from memory (mj_loadModelBuffer, src/engine/engine_io.c:776) */
MJB_API mjModel* mjb_loadModelBuffer(const void* buffer, long long buffer_sz, char* err, int err_sz);
/* release a model returned by mjb_loadModel (NOT for models owned by libmujoco) */
MJB_API void mjb_freeModel(mjModel* m);

/* size fields by name ("nq", "nv", ...): 0 on success */
MJB_API int mjb_modelInt(const mjModel* m, const char* name, long long* out);
/* array fields by name: pointer, rows, cols and element code (0 double, 1 int, 2 byte, 3 float) */
MJB_API int mjb_modelArray(const mjModel* m, const char* name, const void** ptr, int* nr, int* nc,
                           int* code);
/* addresses of mjOption members by name (NULL if unknown), for toggling flags before mjb_makeData */
MJB_API int* mjb_modelOptInt(mjModel* m, const char* name);
MJB_API double* mjb_modelOptNum(mjModel* m, const char* name, int* n);

#if defined(__cplusplus)
}
#endif

#endif  /* MJB_MODELIO_H_ */
