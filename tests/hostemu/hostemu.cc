// TEST-ONLY host compilation of the per-state device pipeline (csrc/mjb_pipeline.h).
//
// This container has no GPU, so the CPU test-suite (`pytest -m "not gpu"`) compiles the very same
// per-thread functions the CUDA kernel runs as plain C++ and checks them against the reference
// engine (oracle/_ref). It exists to debug the algorithmic restructuring (Jacobian-free constraint
// rows, static candidate pairs) before spending GPU time. It is NOT part of libmjb.so, is not
// importable from the package, and nothing in the product path can reach it: mjb_inverse() has no
// CPU fallback.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>


#include "mjb_pipeline.h"
#include "mjb_upload.h"

extern "C" {

// out: mjb::Outputs with HOST pointers laid out [rows][nbatch]; null members are skipped
__attribute__((visibility("default")))
int hostemu_inverse(const mjModel_* m, int nbatch, const double* qpos_soa, const double* qvel_soa,
                    const double* qacc_soa, int nconmax, int njmax, const mjb::Outputs* out,
                    char* err, int err_sz) {
  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) {
    if (err && err_sz > 0) std::snprintf(err, err_sz, "%s", msg.c_str());
    return -1;
  }
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(blob.data());
  std::vector<double> scratch((size_t)H->nscratch + 1);
  std::vector<int> iscratch_v((size_t)mjb::isc_rows(*H));
  int* iscratch = iscratch_v.data();
  std::vector<double> qacc_discrete(H->discrete_acc ? (size_t)H->nv * nbatch : 1);
  for (int s = 0; s < nbatch; s++) {
    mjb::Ctx c;
    c.H = H;
    c.I = reinterpret_cast<const int*>(blob.data() + H->int_section);
    c.D = reinterpret_cast<const double*>(blob.data() + H->num_section);
    double carry_slots[MJB_SM_SLOTS];
    c.sm = carry_slots;
    c.sc = scratch.data();
    c.isc = iscratch;
    c.qpos = qpos_soa + s;
    c.qvel = qvel_soa + s;
    c.qacc = qacc_soa + s;
    c.N = nbatch;
    c.s = s;
    c.nconmax = nconmax;
    c.njmax = njmax;
    c.out = *out;
    c.lci = nullptr; c.lcd = nullptr; c.lbody0 = 0; c.ldof0 = 0;
    mjb::inverse_one_state(c, qacc_discrete.data());
  }
  return 0;
}

__attribute__((visibility("default")))
int hostemu_nscratch(const mjModel_* m) {
  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) return -1;
  return reinterpret_cast<const mjbHdr*>(blob.data())->nscratch;
}

// size of the model blob in bytes (what a CTA stages into shared memory), or -1
__attribute__((visibility("default")))
int hostemu_model_bytes(const mjModel_* m) {
  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) return -1;
  return (int)blob.size();
}

__attribute__((visibility("default")))
int hostemu_slot(const mjModel_* m, const char* name, int* offset, int* size) {
  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) return -1;
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(blob.data());
  for (int s = 0; s < MJB_SC_COUNT; s++) {
    if (!std::strcmp(mjb::scratchSlotName(s), name)) {
      *offset = H->scoff[s];
      *size = (s + 1 < MJB_SC_COUNT ? H->scoff[s + 1] : H->nscratch) - H->scoff[s];
      return 0;
    }
  }
  return -1;
}

// candidate geom pairs (g1, g2, func) in contact order; returns the count (or -1)
__attribute__((visibility("default")))
int hostemu_candidates(const mjModel_* m, int* out, int max_pairs, char* err, int err_sz) {
  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) {
    if (err && err_sz > 0) std::snprintf(err, err_sz, "%s", msg.c_str());
    return -1;
  }
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(blob.data());
  const int* ci = reinterpret_cast<const int*>(blob.data() + H->int_section) + H->ioff[MJB_I_cand_int];
  for (int i = 0; i < H->ncand && i < max_pairs; i++) {
    out[3*i] = ci[MJB_CAND_NI*i + MJB_CI_G1];
    out[3*i + 1] = ci[MJB_CAND_NI*i + MJB_CI_G2];
    out[3*i + 2] = ci[MJB_CAND_NI*i + MJB_CI_FUNC];
  }
  return H->ncand;
}

}  // extern "C"
