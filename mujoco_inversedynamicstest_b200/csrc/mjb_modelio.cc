// Model ingestion without the XML compiler in-process: read the reference's binary MJB format
// (written by mj_saveModel, src/engine/engine_io.c:720-772; read by mj_loadModelBuffer :776-893)
// into a caller-owned mjModel, and give name-based access to its arrays for host-side tools.
//
// The struct layout and the field order come from the reference's own headers
// (<mujoco/mjmodel.h>, <mujoco/mjxmacro.h>), so a file saved by the user's libmujoco round-trips.
// A model made here is valid for every read-only use (mjb_makeData, inspection); it must be
// released with mjb_freeModel, not mj_deleteModel.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <mujoco/mujoco.h>
#include <mujoco/mjxmacro.h>

#include "../../include/mjb.h"
#include "mjb_modelio.h"

namespace {

const int kHeaderInts = 5;       // NHEADER (engine_io.c:302)
const int kMjbId = 54321;        // ID (engine_io.c:298)

int countInts() {
  int n = 0;
#define X(name) n += (sizeof(((mjModel*)0)->name) == sizeof(int)) ? 1 : 0;
  MJMODEL_INTS
#undef X
  return n;
}

int countSizes() {
  int n = 0;
#define X(name) n += (sizeof(((mjModel*)0)->name) == sizeof(int)) ? 0 : 1;
  MJMODEL_INTS
#undef X
  return n;
}

int countPointers() {
  int n = 0;
#define X(type, name, nr, nc) n++;
  MJMODEL_POINTERS
#undef X
  return n;
}

void fail(char* err, int err_sz, const std::string& msg) {
  if (err && err_sz > 0) std::snprintf(err, err_sz, "%s", msg.c_str());
}

enum { CODE_DOUBLE = 0, CODE_INT = 1, CODE_BYTE = 2, CODE_FLOAT = 3 };
template <typename T> struct Code { static const int v = CODE_BYTE; };
template <> struct Code<double> { static const int v = CODE_DOUBLE; };
template <> struct Code<int> { static const int v = CODE_INT; };
template <> struct Code<float> { static const int v = CODE_FLOAT; };

}  // namespace

extern "C" {

mjModel* mjb_loadModel(const char* path, char* err, int err_sz) {
  FILE* fp = std::fopen(path, "rb");
  if (!fp) { fail(err, err_sz, std::string("cannot open ") + path); return nullptr; }
  std::fseek(fp, 0, SEEK_END);
  const long sz = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  std::vector<unsigned char> buf((size_t)(sz > 0 ? sz : 0));
  if (sz <= 0 || std::fread(buf.data(), 1, (size_t)sz, fp) != (size_t)sz) {
    std::fclose(fp);
    fail(err, err_sz, "cannot read model file");
    return nullptr;
  }
  std::fclose(fp);
  return mjb_loadModelBuffer(buf.data(), (long long)buf.size(), err, err_sz);
}

mjModel* mjb_loadModelBuffer(const void* buffer, long long buffer_sz, char* err, int err_sz) {
  const unsigned char* bytes = (const unsigned char*)buffer;
  const size_t nbytes = (size_t)(buffer_sz > 0 ? buffer_sz : 0);
  size_t p = 0;
  auto take = [&](void* dst, size_t n) -> bool {
    if (p + n > nbytes) return false;
    std::memcpy(dst, bytes + p, n);
    p += n;
    return true;
  };

  int header[kHeaderInts];
  if (!take(header, sizeof(header))) { fail(err, err_sz, "incomplete MJB header"); return nullptr; }
  const int expect[kHeaderInts] = {kMjbId, (int)sizeof(mjtNum), countInts(), countSizes(), countPointers()};
  for (int i = 0; i < kHeaderInts; i++) {
    if (header[i] != expect[i]) {
      char msg[160];
      std::snprintf(msg, sizeof(msg), "MJB header field %d is %d, this build expects %d "
                    "(file from a different MuJoCo version?)", i, header[i], expect[i]);
      fail(err, err_sz, msg);
      return nullptr;
    }
  }

  mjModel* m = (mjModel*)std::calloc(1, sizeof(mjModel));
  if (!m) { fail(err, err_sz, "out of memory"); return nullptr; }
  bool ok = true;
#define X(name) ok = ok && take(&m->name, sizeof(m->name));
  MJMODEL_INTS
#undef X
  ok = ok && take(&m->opt, sizeof(mjOption));
  ok = ok && take(&m->vis, sizeof(mjVisual));
  ok = ok && take(&m->stat, sizeof(mjStatistic));
  if (!ok) { std::free(m); fail(err, err_sz, "truncated MJB file (sizes)"); return nullptr; }

  // one allocation for all arrays, each 64-byte aligned like mj_makeModel's buffer
  size_t total = 0;
  {
    MJMODEL_POINTERS_PREAMBLE(m)
#define X(type, name, nr, nc) total += ((sizeof(type) * (size_t)(m->nr) * (size_t)(nc)) + 63) & ~(size_t)63;
    MJMODEL_POINTERS
#undef X
  }
  unsigned char* base = (unsigned char*)std::aligned_alloc(64, total ? total : 64);
  if (!base) { std::free(m); fail(err, err_sz, "out of memory"); return nullptr; }
  m->buffer = base;
  size_t off = 0;
  {
    MJMODEL_POINTERS_PREAMBLE(m)
#define X(type, name, nr, nc)                                            \
    {                                                                    \
      const size_t bytes = sizeof(type) * (size_t)(m->nr) * (size_t)(nc); \
      m->name = (type*)(base + off);                                     \
      ok = ok && take(m->name, bytes);                                   \
      off += (bytes + 63) & ~(size_t)63;                                 \
    }
    MJMODEL_POINTERS
#undef X
  }
  if (!ok || p != nbytes) {
    std::free(base);
    std::free(m);
    fail(err, err_sz, ok ? "MJB file has trailing data" : "truncated MJB file (arrays)");
    return nullptr;
  }
  return m;
}

void mjb_freeModel(mjModel* m) {
  if (m) {
    std::free(m->buffer);
    std::free(m);
  }
}

int mjb_modelInt(const mjModel* m, const char* name, long long* out) {
#define X(field) if (!std::strcmp(name, #field)) { *out = (long long)m->field; return 0; }
  MJMODEL_INTS
#undef X
  return -1;
}

int mjb_modelArray(const mjModel* m, const char* name, const void** ptr, int* nr, int* nc, int* code) {
  MJMODEL_POINTERS_PREAMBLE(m)
  (void)nuser_body; (void)nuser_jnt; (void)nuser_geom; (void)nuser_site; (void)nuser_cam;
  (void)nuser_tendon; (void)nuser_actuator; (void)nuser_sensor; (void)nq; (void)nv; (void)na;
  (void)nu; (void)nmocap;
#define X(type, field, r, c)                                                                 \
  if (!std::strcmp(name, #field)) {                                                          \
    *ptr = m->field; *nr = (int)m->r; *nc = (int)(c); *code = Code<type>::v; return 0;       \
  }
  MJMODEL_POINTERS
#undef X
  return -1;
}

int* mjb_modelOptInt(mjModel* m, const char* name) {
#define X(type, field) if (!std::strcmp(name, #field)) return &m->opt.field;
  MJOPTION_INTS
#undef X
  return nullptr;
}

double* mjb_modelOptNum(mjModel* m, const char* name, int* n) {
#define X(type, field) if (!std::strcmp(name, #field)) { *n = 1; return &m->opt.field; }
  MJOPTION_FLOATS
#undef X
#define X(field, cnt) if (!std::strcmp(name, #field)) { *n = (cnt); return m->opt.field; }
  MJOPTION_VECTORS
#undef X
  return nullptr;
}

}  // extern "C"
