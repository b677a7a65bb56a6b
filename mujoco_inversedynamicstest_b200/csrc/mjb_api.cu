// C-ABI of libmjb (declared in include/mjb.h): owns device memory, moves states across the host
// boundary and launches the kernels. No torch types, no C++ types in any signature.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>
#include <dlfcn.h>

#include "../../include/mjb.h"
#include "mjb_kernels.cuh"
#include "mjb_jit.h"
#include "mjb_upload.h"

struct mjbData_ {
  int device = 0;
  cudaStream_t stream = 0;
  unsigned outmask = 0;
  int nbatch_max = 0;
  long long stride = 0;        // row stride of the internal SoA buffers
  int nconmax = 0, njmax = 0;
  int last_nbatch = 0;
  mjbHdr hdr;                  // host copy of the model header
  std::vector<unsigned char> blob;   // host copy of the model blob (input of mjb_specialize)
  mjb::SpecKernels spec;       // kernels compiled for this model (mjb_specialize), if any
  bool spec_on = false;
  bool spec_wanted = false;    // the caller asked for specialised kernels (re-specialise after a model refresh)
  unsigned long long opt_hash = 0;   // hash of m->opt when the blob was built (mj_inverse honours m->opt per call)
  std::vector<int> cand;       // host copy of candidate (g1, g2, func)
  unsigned char* d_model = nullptr;
  int model_bytes = 0;
  int model_in_smem = 0;
  double* d_scratch = nullptr;   // [chunk_stride/32][nscratch][32]
  int* d_iscratch = nullptr;      // [MJB_ISC_MASK + ceil(ncand/32)][chunk_stride]
  long long chunk_stride = 0;     // states per chunk (intermediates are allocated per chunk)
  long long kernel_launches = 0;  // phase kernels launched so far (reported by the benchmark)
  // optional per-kernel timing (mjb_phaseTiming): event pool and (phase, begin, end) marks
  bool phase_timing = false;
  std::vector<cudaEvent_t> ev_pool;
  size_t ev_used = 0;
  struct Mark { int phase; cudaEvent_t b, e; };
  std::vector<Mark> marks;
  // inputs: internal SoA buffers and the views currently in use
  double *d_qpos = nullptr, *d_qvel = nullptr, *d_qacc = nullptr;
  // item-parallel contact phase: global lists of one chunk (null: pooled kernel only)
  mjb::ContactQueue* d_cq = nullptr;
  mjb::ContactItem* d_items = nullptr;
  mjb::ItemCon* d_item_con = nullptr;
  mjb::ContactRec* d_contacts = nullptr;
  mjb::SlotRec* d_slot_rec = nullptr;
  int* d_scan_buf = nullptr;
  int* d_cmask = nullptr;      // per-state survivor masks of the warp-per-state scans (one state's words contiguous)
  int* d_pair_ci = nullptr;    // geom pair -> candidate index (ngeom x ngeom), long candidate lists only   // per-warp candidate buffers of the warp-per-state scan (long candidate lists)
  int items_cap = 0, contacts_cap = 0;
  double* d_eq_active = nullptr;       // per-state d->eq_active [neq][stride] as 0 / 1 (mjb_setEqActive)
  double* d_xfrc_applied = nullptr;    // per-state d->xfrc_applied [nbody*6][stride] (mjb_setXfrcApplied)
  double* d_qacc_discrete = nullptr;   // continuous-time qacc when mjENBL_INVDISCRETE converts it
  const double *in_qpos = nullptr, *in_qvel = nullptr, *in_qacc = nullptr;
  long long in_stride = 0;
  // AoS staging (host boundary), grown on demand
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  int* d_counter = nullptr;
  // host pipeline of mjb_inverseHost: copy-in / copy-out streams, double-buffered staging, events
  cudaStream_t s_in = nullptr, s_out = nullptr, s_comp = nullptr;
  cudaEvent_t ev_start = nullptr;
  // work has been queued on the caller's stream since the last mjb_inverseHost: the pipeline's
  // compute stream has to be ordered after it once
  bool stream_dirty = true;
  void* pipe_in[2] = {nullptr, nullptr};
  void* pipe_out[2] = {nullptr, nullptr};
  size_t pipe_piece = 0;
  unsigned pipe_seq = 0;                 // pieces sent so far: piece k uses staging buffer k & 1, across calls
  bool pipe_used[2] = {false, false};   // staging buffer b has been through the pipeline (its events are valid)
  cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_tr[2] = {nullptr, nullptr},
              ev_comp[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
  mjb::Outputs out;            // device SoA outputs
  void* field_ptr[mjbF_COUNT];
  int field_rows[mjbF_COUNT];
  int field_isint[mjbF_COUNT];
  bool field_set[mjbF_COUNT];   // requested in outmask (a field of a model with no rows of it has no buffer)
  // mjb_inverseFD: inner batch of perturbed states and the device buffer of the differences
  mjbData* fd = nullptr;
  int fd_tile = 0;
  // mjb_compareFwdInv: forward-pass quantities (SoA) and the result, allocated on first use
  double *d_fwd_qforce = nullptr, *d_fwd_xfrc = nullptr, *d_fwd_qc = nullptr, *d_fwdinv = nullptr;
  bool own_qfrc_constraint = false;
  double *d_mocap_pos = nullptr, *d_mocap_quat = nullptr;   // per-state mocap poses (mjb_setMocap)
  int skip_sensors = 0;        // mj_inverseSkip(skipsensor = 1) / inner batches of mjb_inverseFD
  double* d_fd_out = nullptr;
  size_t fd_out_doubles = 0;
  std::string error;
  // multi-device front (mjb_makeDataMulti): this object owns no device memory itself; shard g is a
  // complete mjbData on devices[g] and evaluates the contiguous range [g*per, (g+1)*per) of a batch,
  // per = ceil(nbatch / ndevice). No collective: states are independent.
  std::vector<mjbData*> shards;
};

namespace {

unsigned long long optHash(const mjModel* m) {
  const unsigned char* p = reinterpret_cast<const unsigned char*>(&m->opt);
  unsigned long long h = 0xcbf29ce484222325ull;
  for (size_t i = 0; i < sizeof(m->opt); i++) h = (h ^ p[i]) * 0x100000001b3ull;
  return h;
}

// The reference consults global callbacks on this path (mjcb_passive engine_passive.c:497-499,
// mjcb_contactfilter engine_collision_driver.c:1474-1478, mjcb_sensor, mjcb_control is not on it): a
// device path cannot call back into host code, so a process whose libmujoco has one of them set is
// refused. The symbols are looked up in the process image (present when the caller links libmujoco).
const char* activeCallback() {
  static const char* const names[] = {"mjcb_passive", "mjcb_contactfilter", "mjcb_sensor", nullptr};
  for (int i = 0; names[i]; i++) {
    void** p = reinterpret_cast<void**>(dlsym(RTLD_DEFAULT, names[i]));
    if (p && *p) return names[i];
  }
  return nullptr;
}

bool check(mjbData* d, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  d->error = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

template <typename T>
bool devAlloc(mjbData* d, T** p, size_t count, const char* what) {
  *p = nullptr;
  if (count == 0) return true;
  return check(d, cudaMalloc((void**)p, count * sizeof(T)), what);
}

bool ensureStage(mjbData* d, size_t bytes) {
  if (bytes <= d->stage_bytes) return true;
  if (d->d_stage) cudaFree(d->d_stage);
  d->d_stage = nullptr;
  d->stage_bytes = 0;
  if (!check(d, cudaMalloc(&d->d_stage, bytes), "cudaMalloc(staging)")) return false;
  d->stage_bytes = bytes;
  return true;
}

void setField(mjbData* d, int f, void* p, int rows, int isint) {
  d->field_ptr[f] = p;
  d->field_rows[f] = rows;
  d->field_isint[f] = isint;
  d->field_set[f] = true;
}

}  // namespace


namespace {

struct ShardRange { long long first; int n; };

ShardRange shardRange(const mjbData* d, int g, int nbatch) {
  const long long G = (long long)d->shards.size();
  const long long per = (nbatch + G - 1) / G;
  long long first = per * g, end = first + per;
  if (first > nbatch) first = nbatch;
  if (end > nbatch) end = nbatch;
  return {first, (int)(end - first)};
}

// run fn(shard, range) for every shard, one host thread per device; returns false if any failed
template <typename F>
bool forShards(mjbData* d, int nbatch, bool threaded, F fn) {
  const int G = (int)d->shards.size();
  std::vector<int> rc(G, 0);
  auto work = [&](int g) { rc[g] = fn(d->shards[g], shardRange(d, g, nbatch)); };
  if (threaded && G > 1) {
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++) th.emplace_back(work, g);
    for (auto& t : th) t.join();
  } else {
    for (int g = 0; g < G; g++) work(g);
  }
  for (int g = 0; g < G; g++) {
    if (rc[g] < 0) { d->error = "device " + std::to_string(d->shards[g]->device) + ": " + d->shards[g]->error; return false; }
  }
  return true;
}

}  // namespace

extern "C" {

mjbData* mjb_makeData(const mjModel* m, int nbatch_max, int device, unsigned outmask, int nconmax,
                      int njmax, char* err, int err_sz) {
  auto fail = [&](const std::string& msg) -> mjbData* {
    if (err && err_sz > 0) std::snprintf(err, err_sz, "%s", msg.c_str());
    return nullptr;
  };
  if (!m) return fail("mjb_makeData: null model");
  if (nbatch_max <= 0) return fail("mjb_makeData: nbatch_max must be positive");
  if (const char* cb = activeCallback()) {
    return fail(std::string("mjb_makeData: the global callback ") + cb +
                " is set; the batched device path cannot call host callbacks");
  }

  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) return fail("mjb_makeData: " + msg);
  {
    const mjbHdr* H0 = reinterpret_cast<const mjbHdr*>(blob.data());
    if (H0->sensor_post) outmask |= mjbOUT_RNEPOST;   // accelerometer / force / torque / frame*acc sensors
    // touch sensors walk the contact list and the contact rows' forces: those outputs are their input
    if (H0->sensor_touch) outmask |= mjbOUT_COUNTS | mjbOUT_CONTACT | mjbOUT_EFC;
    if (H0->sensor_camlight) outmask |= mjbOUT_CAMLIGHT;           // camprojection sensors read the camera poses
    if (H0->sensor_trn) outmask |= mjbOUT_TRANSMISSION;       // actuatorpos / actuatorvel sensors
    if (H0->discrete_trn) outmask |= mjbOUT_TRANSMISSION;     // implicitfast mj_discreteAcc reads the moment rows
  }
  if (outmask & mjbOUT_RNEPOST) {
    // constraint forces of spatial tendons travel as body wrenches here but are not part of the
    // reference's cfrc_ext; a massless tree has its frame origin at xipos, which is not kept
    const mjbHdr* H0 = reinterpret_cast<const mjbHdr*>(blob.data());
    if (H0->has_spatial) {
      return fail("mjb_makeData: mjbOUT_RNEPOST is not available for models whose spatial tendons carry forces");
    }
    for (int b = 1; b < m->nbody; b++) {
      if (m->body_parentid[b] == 0 && m->body_subtreemass[b] < mjMINVAL) {
        return fail("mjb_makeData: mjbOUT_RNEPOST needs every kinematic tree to have mass");
      }
    }
  }

  if (outmask & mjbOUT_TRANSMISSION) {
    // adhesion actuators take their moment from the contact normals (engine_core_smooth.c:1222-1330): the
    // contact list is an input of the transmission stage
    for (int i = 0; i < m->nu; i++) {
      if (m->actuator_trntype[i] == mjTRN_BODY) outmask |= mjbOUT_COUNTS | mjbOUT_CONTACT;
    }
  }

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    return fail("mjb_makeData: no CUDA device available (libmjb has no CPU path)");
  }
  if (device < 0 || device >= ndev) return fail("mjb_makeData: invalid device index");

  mjbData* d = new mjbData_;
  std::memset(&d->out, 0, sizeof(d->out));
  for (int f = 0; f < mjbF_COUNT; f++) { setField(d, f, nullptr, 0, 0); d->field_set[f] = false; }
  d->device = device;
  d->outmask = outmask;
  d->nbatch_max = nbatch_max;
  d->stride = ((long long)nbatch_max + 31) & ~31LL;   // keep every row 256-byte aligned
  std::memcpy(&d->hdr, blob.data(), sizeof(mjbHdr));
  d->blob = blob;
  d->opt_hash = optHash(m);
  const mjbHdr& H = d->hdr;
  d->nconmax = nconmax > 0 ? nconmax : 64;
  d->njmax = njmax > 0 ? njmax : 256;
  {
    const int* ci = reinterpret_cast<const int*>(blob.data() + H.int_section) + H.ioff[MJB_I_cand_int];
    for (int i = 0; i < H.ncand; i++) {
      d->cand.push_back(ci[MJB_CAND_NI*i + MJB_CI_G1]);
      d->cand.push_back(ci[MJB_CAND_NI*i + MJB_CI_G2]);
      d->cand.push_back(ci[MJB_CAND_NI*i + MJB_CI_FUNC]);
    }
  }

  bool ok = check(d, cudaSetDevice(device), "cudaSetDevice");
  d->model_bytes = H.staged_bytes;      // what the kernels stage; the whole blob (H.bytes) is on the device
  d->model_in_smem = H.staged_bytes <= 64*1024;
  ok = ok && devAlloc(d, &d->d_model, (size_t)H.bytes, "cudaMalloc(model)");
  ok = ok && check(d, cudaMemcpy(d->d_model, blob.data(), (size_t)H.bytes, cudaMemcpyHostToDevice),
                   "cudaMemcpy(model)");
  // intermediates live per chunk of states: at most 2^20 states and at most ~20 GB (2^20 humanoid states
  // are one chunk of 16.3 GB)
  {
    const double bytes_per_state = 8.0 * H.nscratch + 4.0 * mjb::isc_rows(H);
    long long chunk = (long long)(20.0e9 / bytes_per_state);
    if (chunk > (1LL << 20)) chunk = 1LL << 20;
    if (chunk > d->stride) chunk = d->stride;
    chunk &= ~127LL;
    if (chunk < 128) chunk = 128;
    d->chunk_stride = chunk;
  }
  // Item-parallel contact lists, sized for the common case (<= 48 surviving pairs and <= 16 contacts
  // per state on average over a chunk, at most ~4 GB); denser chunks fall back to the pooled
  // kernel on the device. MJB_CONTACT_PATH=pooled disables the item path (A/B measurements).
  {
    const char* path = std::getenv("MJB_CONTACT_PATH");
    const bool want = H.ncand > 0 && !(H.disableflags & (MJB_DSBL_CONSTRAINT | MJB_DSBL_CONTACT)) &&
                      !(path && !std::strcmp(path, "pooled"));
    if (want) {
      // MJB_ITEMS_PER_STATE / MJB_CONTACTS_PER_STATE override the list sizing (tests, measurements)
      // sized per model: a single articulated figure has <= 48 survivors and <= 16 contacts per state
      // on average (humanoid: 15.6 and 6.1); scenes with many geoms (22 humanoids in a pile: 1,210
      // survivors and 322 contacts per state, 419 geoms; 100 humanoids: ~4,300 contacts, 1901 geoms) get lists
      // proportional to the geom count
      long long per_items = H.ngeom > 64 ? 8LL * H.ngeom : 48, per_contacts = H.ngeom > 64 ? 3LL * H.ngeom : 16;
      if (per_items > H.ncand) per_items = H.ncand;
      if (const char* env = std::getenv("MJB_ITEMS_PER_STATE")) { if (std::atol(env) > 0) per_items = std::atol(env); }
      if (const char* env = std::getenv("MJB_CONTACTS_PER_STATE")) { if (std::atol(env) > 0) per_contacts = std::atol(env); }
      long long ni = d->chunk_stride * per_items, nc = d->chunk_stride * per_contacts;
      const long long budget = 4LL << 30;
      const long long bytes = ni * (long long)(sizeof(mjb::ContactItem) + sizeof(mjb::ItemCon)) +
                              nc * (long long)(sizeof(mjb::ContactRec) + sizeof(mjb::SlotRec));
      if (bytes > budget) { ni = ni * budget / bytes; nc = nc * budget / bytes; }
      if (ni > 0x7ffffff0LL) ni = 0x7ffffff0LL;
      if (nc > 0x7ffffff0LL) nc = 0x7ffffff0LL;
      d->items_cap = (int)ni; d->contacts_cap = (int)nc;
      ok = ok && devAlloc(d, &d->d_cq, 1, "cudaMalloc(contact queue)");
      ok = ok && devAlloc(d, &d->d_items, (size_t)ni, "cudaMalloc(contact items)");
      ok = ok && devAlloc(d, &d->d_item_con, (size_t)ni, "cudaMalloc(item contacts)");
      ok = ok && devAlloc(d, &d->d_contacts, (size_t)nc, "cudaMalloc(contact records)");
      ok = ok && devAlloc(d, &d->d_slot_rec, (size_t)nc, "cudaMalloc(contact slots)");
    }
    if (H.ncand > 0 && mjb::scan_wide_states(H.ncand, H.ngeom) > 0) {
      ok = ok && devAlloc(d, &d->d_scan_buf, (size_t)mjb::scan_wide_buf_ints(H.ncand, H.ngeom), "cudaMalloc(scan buffers)");
      ok = ok && devAlloc(d, &d->d_cmask, (size_t)d->chunk_stride * (size_t)((H.ncand + 31) / 32), "cudaMalloc(scan masks)");
      // geom pair -> candidate: the pair-organised scan needs every pair to map to at most one candidate
      // and geom ids that fit 16 bits; MJB_SCAN=list keeps the list-driven kernel (A/B measurements)
      const char* sm = std::getenv("MJB_SCAN");
      if (ok && H.ngeom < 65536 && !(sm && !std::strcmp(sm, "list"))) {
        const int* si = reinterpret_cast<const int*>(blob.data() + H.int_section) + H.ioff[MJB_I_scan_int];
        std::vector<int> pc((size_t)H.ngeom * H.ngeom, -1);
        bool unique = true;
        for (int i = 0; i < H.ncand && unique; i++) {
          const size_t g1 = (size_t)(si[2*i] & 0xfffffff), g2 = (size_t)si[2*i + 1];
          if (g1 == g2 || pc[g1 * H.ngeom + g2] >= 0) { unique = false; break; }
          pc[g1 * H.ngeom + g2] = i; pc[g2 * H.ngeom + g1] = i;
        }
        if (unique) {
          ok = ok && devAlloc(d, &d->d_pair_ci, pc.size(), "cudaMalloc(pair table)");
          ok = ok && check(d, cudaMemcpy(d->d_pair_ci, pc.data(), pc.size() * sizeof(int), cudaMemcpyHostToDevice),
                           "cudaMemcpy(pair table)");
        }
      }
    }
  }
  ok = ok && devAlloc(d, &d->d_scratch, (size_t)H.nscratch * (size_t)d->chunk_stride, "cudaMalloc(scratch)");
  ok = ok && devAlloc(d, &d->d_iscratch, (size_t)mjb::isc_rows(H) * (size_t)d->chunk_stride, "cudaMalloc(iscratch)");
  const size_t S = (size_t)d->stride;
  ok = ok && devAlloc(d, &d->d_qpos, (size_t)H.nq * S, "cudaMalloc(qpos)");
  ok = ok && devAlloc(d, &d->d_qvel, (size_t)H.nv * S, "cudaMalloc(qvel)");
  ok = ok && devAlloc(d, &d->d_qacc, (size_t)H.nv * S, "cudaMalloc(qacc)");
  ok = ok && devAlloc(d, &d->d_counter, 1, "cudaMalloc(counter)");
  d->in_qpos = d->d_qpos; d->in_qvel = d->d_qvel; d->in_qacc = d->d_qacc; d->in_stride = d->stride;

  mjb::Outputs& o = d->out;
  ok = ok && devAlloc(d, &o.qfrc_inverse, (size_t)H.nv * S, "cudaMalloc(qfrc_inverse)");
  ok = ok && devAlloc(d, &o.status, S, "cudaMalloc(status)");
  setField(d, mjbF_QFRC_INVERSE, o.qfrc_inverse, H.nv, 0);
  setField(d, mjbF_STATUS, o.status, 1, 1);
  if (outmask & mjbOUT_QFRC) {
    ok = ok && devAlloc(d, &o.qfrc_constraint, (size_t)H.nv * S, "cudaMalloc(qfrc_constraint)");
    ok = ok && devAlloc(d, &o.qfrc_passive, (size_t)H.nv * S, "cudaMalloc(qfrc_passive)");
    setField(d, mjbF_QFRC_CONSTRAINT, o.qfrc_constraint, H.nv, 0);
    setField(d, mjbF_QFRC_PASSIVE, o.qfrc_passive, H.nv, 0);
    ok = ok && devAlloc(d, &o.qfrc_bias, (size_t)H.nv * S, "cudaMalloc(qfrc_bias)");
    setField(d, mjbF_QFRC_BIAS, o.qfrc_bias, H.nv, 0);
  }
  if (outmask & mjbOUT_COUNTS) {
    ok = ok && devAlloc(d, &o.counts, 5 * S, "cudaMalloc(counts)");
    setField(d, mjbF_COUNTS, o.counts, 5, 1);
  }
  if (outmask & mjbOUT_CONTACT) {
    const size_t nc = (size_t)d->nconmax;
    ok = ok && devAlloc(d, &o.contact_geom, 2 * nc * S, "cudaMalloc(contact_geom)");
    ok = ok && devAlloc(d, &o.contact_info, 3 * nc * S, "cudaMalloc(contact_info)");
    ok = ok && devAlloc(d, &o.contact_num, 13 * nc * S, "cudaMalloc(contact_num)");
    setField(d, mjbF_CONTACT_GEOM, o.contact_geom, 2 * d->nconmax, 1);
    setField(d, mjbF_CONTACT_INFO, o.contact_info, 3 * d->nconmax, 1);
    setField(d, mjbF_CONTACT_NUM, o.contact_num, 13 * d->nconmax, 0);
  }
  if (outmask & mjbOUT_EFC) {
    const size_t nj = (size_t)d->njmax;
    ok = ok && devAlloc(d, &o.efc_int, 3 * nj * S, "cudaMalloc(efc_int)");
    ok = ok && devAlloc(d, &o.efc_num, 8 * nj * S, "cudaMalloc(efc_num)");
    setField(d, mjbF_EFC_INT, o.efc_int, 3 * d->njmax, 1);
    setField(d, mjbF_EFC_NUM, o.efc_num, 8 * d->njmax, 0);
  }
  if (H.discrete_acc) {
    ok = ok && devAlloc(d, &d->d_qacc_discrete, (size_t)H.nv * S, "cudaMalloc(qacc_discrete)");
  }
  if ((outmask & mjbOUT_INERTIA) || H.discrete_acc) {   // mj_discreteAcc solves with the factors
    ok = ok && devAlloc(d, &o.qM, (size_t)H.nM * S, "cudaMalloc(qM)");
    ok = ok && devAlloc(d, &o.qLD, (size_t)H.nC * S, "cudaMalloc(qLD)");
    ok = ok && devAlloc(d, &o.qLDiagInv, (size_t)H.nv * S, "cudaMalloc(qLDiagInv)");
    setField(d, mjbF_QM, o.qM, H.nM, 0);
    setField(d, mjbF_QLD, o.qLD, H.nC, 0);
    setField(d, mjbF_QLDIAGINV, o.qLDiagInv, H.nv, 0);
  }
  if (outmask & mjbOUT_INTERNAL) {
    ok = ok && devAlloc(d, &o.scratch_dump, (size_t)H.nscratch * S, "cudaMalloc(internal)");
    setField(d, mjbF_INTERNAL, o.scratch_dump, H.nscratch, 0);
  }
  if (outmask & mjbOUT_RNEPOST) {
    // mj_rnePostConstraint: cacc, cfrc_int, cfrc_ext (engine_core_smooth.c:2027-2181)
    const size_t nb6 = 6 * (size_t)H.nbody;
    ok = ok && devAlloc(d, &o.cacc, nb6 * S, "cudaMalloc(cacc)");
    ok = ok && devAlloc(d, &o.cfrc_int, nb6 * S, "cudaMalloc(cfrc_int)");
    ok = ok && devAlloc(d, &o.cfrc_ext, nb6 * S, "cudaMalloc(cfrc_ext)");
    setField(d, mjbF_CACC, o.cacc, 6 * H.nbody, 0);
    setField(d, mjbF_CFRC_INT, o.cfrc_int, 6 * H.nbody, 0);
    setField(d, mjbF_CFRC_EXT, o.cfrc_ext, 6 * H.nbody, 0);
  }
  if (outmask & mjbOUT_TRANSMISSION) {
    // mj_transmission + actuator_velocity (engine_core_smooth.c:865-1346, engine_forward.c:216)
    const size_t nu = (size_t)(H.nu > 0 ? H.nu : 1), nvv = (size_t)(H.nv > 0 ? H.nv : 1);
    ok = ok && devAlloc(d, &o.actuator_length, nu * S, "cudaMalloc(actuator_length)");
    ok = ok && devAlloc(d, &o.actuator_moment, nu * nvv * S, "cudaMalloc(actuator_moment)");
    ok = ok && devAlloc(d, &o.actuator_velocity, nu * S, "cudaMalloc(actuator_velocity)");
    setField(d, mjbF_ACTUATOR_LENGTH, o.actuator_length, H.nu, 0);
    setField(d, mjbF_ACTUATOR_MOMENT, o.actuator_moment, H.nu * H.nv, 0);
    setField(d, mjbF_ACTUATOR_VELOCITY, o.actuator_velocity, H.nu, 0);
  }
  if (outmask & mjbOUT_CAMLIGHT) {
    // mj_camlight (engine_core_smooth.c:275-389); one slot is kept for models without cameras or lights
    // so that the four pointers are set together
    const size_t nc = (size_t)(H.ncam > 0 ? H.ncam : 1), nl = (size_t)(H.nlight > 0 ? H.nlight : 1);
    ok = ok && devAlloc(d, &o.cam_xpos, 3 * nc * S, "cudaMalloc(cam_xpos)");
    ok = ok && devAlloc(d, &o.cam_xmat, 9 * nc * S, "cudaMalloc(cam_xmat)");
    ok = ok && devAlloc(d, &o.light_xpos, 3 * nl * S, "cudaMalloc(light_xpos)");
    ok = ok && devAlloc(d, &o.light_xdir, 3 * nl * S, "cudaMalloc(light_xdir)");
    setField(d, mjbF_CAM_XPOS, o.cam_xpos, 3 * H.ncam, 0);
    setField(d, mjbF_CAM_XMAT, o.cam_xmat, 9 * H.ncam, 0);
    setField(d, mjbF_LIGHT_XPOS, o.light_xpos, 3 * H.nlight, 0);
    setField(d, mjbF_LIGHT_XDIR, o.light_xdir, 3 * H.nlight, 0);
  }
  if ((H.enableflags & MJB_ENBL_ENERGY) || H.sensor_energy) {
    // d->energy is part of mj_inverse's output contract when mjENBL_ENERGY is set (engine_inverse.c:210-223)
    ok = ok && devAlloc(d, &o.energy, (size_t)2 * S, "cudaMalloc(energy)");
    setField(d, mjbF_ENERGY, o.energy, 2, 0);
  }
  if (H.nsensordata > 0) {
    // sensordata is part of mj_inverse's output contract (engine_inverse.c:206-246): always produced
    ok = ok && devAlloc(d, &o.sensordata, (size_t)H.nsensordata * S, "cudaMalloc(sensordata)");
    ok = ok && check(d, cudaMemset(o.sensordata, 0, (size_t)H.nsensordata * S * sizeof(double)), "cudaMemset(sensordata)");
    setField(d, mjbF_SENSORDATA, o.sensordata, H.nsensordata, 0);
  }
  if (!ok) {
    std::string e = d->error;
    mjb_deleteData(d);
    return fail("mjb_makeData: " + e);
  }
  // MJB_JIT=1: specialise every eligible model at creation (otherwise on request, mjb_specialize).
  // A model that cannot be specialised keeps the generic kernels; the reason stays in mjb_lastError.
  if (const char* jit = std::getenv("MJB_JIT")) {
    if (jit[0] && jit[0] != '0') mjb_specialize(d, nullptr, 0);
  }
  return d;
}

mjbData* mjb_makeDataMulti(const mjModel* m, int nbatch_max, const int* devices, int ndevice, unsigned outmask,
                           int nconmax, int njmax, char* err, int err_sz) {
  auto fail = [&](const std::string& msg) -> mjbData* {
    if (err && err_sz > 0) std::snprintf(err, err_sz, "%s", msg.c_str());
    return nullptr;
  };
  if (!devices || ndevice <= 0) return fail("mjb_makeDataMulti: at least one device is required");
  if (nbatch_max <= 0) return fail("mjb_makeDataMulti: nbatch_max must be positive");
  mjbData* d = new mjbData_;
  std::memset(&d->out, 0, sizeof(d->out));
  for (int f = 0; f < mjbF_COUNT; f++) { setField(d, f, nullptr, 0, 0); d->field_set[f] = false; }
  d->nbatch_max = nbatch_max;
  d->outmask = outmask;
  const int per = (nbatch_max + ndevice - 1) / ndevice;
  for (int g = 0; g < ndevice; g++) {
    mjbData* sh = mjb_makeData(m, per, devices[g], outmask, nconmax, njmax, err, err_sz);
    if (!sh) { mjb_deleteData(d); return nullptr; }
    d->shards.push_back(sh);
  }
  d->device = devices[0];
  d->hdr = d->shards[0]->hdr;
  d->nconmax = d->shards[0]->nconmax; d->njmax = d->shards[0]->njmax;
  for (int f = 0; f < mjbF_COUNT; f++) { d->field_rows[f] = d->shards[0]->field_rows[f]; d->field_isint[f] = d->shards[0]->field_isint[f]; }
  return d;
}

int mjb_ndevice(const mjbData* d) { return d->shards.empty() ? 1 : (int)d->shards.size(); }

int mjb_specialize(mjbData* d, char* err, int err_sz) {
  if (!d->shards.empty()) {
    int rc = 0;
    for (mjbData* sh : d->shards) if (mjb_specialize(sh, err, err_sz)) rc = -1;
    d->spec_on = rc == 0;
    return rc;
  }
  d->spec_wanted = true;
  if (d->spec_on) return 0;
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  std::string msg;
  if (!mjb::jitSpecialize(d->blob, d->spec, msg)) {
    d->error = "mjb_specialize: " + msg;
    if (err && err_sz > 0) std::snprintf(err, err_sz, "%s", d->error.c_str());
    return -1;
  }
  d->spec_on = true;
  return 0;
}

int mjb_specialized(const mjbData* d) { return d->spec_on ? 1 : 0; }

int mjb_specializeInfo(const mjbData* d, char* key, int key_sz, int* from_cache, double* compile_seconds) {
  if (!d->spec_on) return -1;
  if (key && key_sz > 0) std::snprintf(key, key_sz, "%s", d->spec.key.c_str());
  if (from_cache) *from_cache = d->spec.from_cache ? 1 : 0;
  if (compile_seconds) *compile_seconds = d->spec.compile_seconds;
  return 0;
}

int mjb_precompile(const mjModel* m, char* info, int info_sz) {
  std::vector<unsigned char> blob;
  std::string msg, key;
  bool cached = false;
  double seconds = 0;
  if (!m || !mjb::buildModelBlob(m, blob, msg) || !mjb::jitCompileToCache(blob, key, cached, seconds, msg)) {
    if (info && info_sz > 0) std::snprintf(info, info_sz, "%s", msg.c_str());
    return -1;
  }
  if (info && info_sz > 0) {
    std::snprintf(info, info_sz, "%s %s %.1f s [%s]", key.c_str(), cached ? "cached" : "compiled", seconds,
                  mjb::jitStagePlan(blob).c_str());
  }
  return 0;
}

void mjb_deleteData(mjbData* d) {
  if (!d) return;
  if (!d->shards.empty()) {
    for (mjbData* sh : d->shards) mjb_deleteData(sh);
    delete d;
    return;
  }
  cudaSetDevice(d->device);
  if (d->fd) mjb_deleteData(d->fd);
  if (d->spec_on) mjb::jitUnload(d->spec);
  cudaFree(d->d_fd_out);
  cudaFree(d->d_model); cudaFree(d->d_scratch); cudaFree(d->d_iscratch);
  cudaFree(d->d_cq); cudaFree(d->d_items); cudaFree(d->d_item_con); cudaFree(d->d_contacts);
  cudaFree(d->d_slot_rec);
  cudaFree(d->d_scan_buf);
  cudaFree(d->d_pair_ci);
  cudaFree(d->d_cmask);
  cudaFree(d->d_qpos); cudaFree(d->d_qvel); cudaFree(d->d_qacc); cudaFree(d->d_qacc_discrete); cudaFree(d->d_xfrc_applied); cudaFree(d->d_eq_active);
  cudaFree(d->d_stage); cudaFree(d->d_counter);
  cudaFree(d->d_mocap_pos); cudaFree(d->d_mocap_quat);
  cudaFree(d->d_fwd_qforce); cudaFree(d->d_fwd_xfrc); cudaFree(d->d_fwd_qc); cudaFree(d->d_fwdinv);
  for (int b = 0; b < 2; b++) {
    cudaFree(d->pipe_in[b]); cudaFree(d->pipe_out[b]);
    if (d->ev_in[b]) cudaEventDestroy(d->ev_in[b]);
    if (d->ev_tr[b]) cudaEventDestroy(d->ev_tr[b]);
    if (d->ev_comp[b]) cudaEventDestroy(d->ev_comp[b]);
    if (d->ev_out[b]) cudaEventDestroy(d->ev_out[b]);
  }
  for (cudaEvent_t e : d->ev_pool) cudaEventDestroy(e);
  if (d->s_comp) cudaStreamDestroy(d->s_comp);
  if (d->ev_start) cudaEventDestroy(d->ev_start);
  if (d->s_in) cudaStreamDestroy(d->s_in);
  if (d->s_out) cudaStreamDestroy(d->s_out);
  mjb::Outputs& o = d->out;
  cudaFree(o.qfrc_inverse); cudaFree(o.qfrc_constraint); cudaFree(o.qfrc_passive);
  cudaFree(o.counts); cudaFree(o.status); cudaFree(o.contact_geom); cudaFree(o.contact_info);
  cudaFree(o.contact_num); cudaFree(o.efc_int); cudaFree(o.efc_num); cudaFree(o.qM);
  cudaFree(o.qLD); cudaFree(o.qLDiagInv); cudaFree(o.scratch_dump);
  cudaFree(o.cacc); cudaFree(o.cfrc_int); cudaFree(o.cfrc_ext); cudaFree(o.sensordata); cudaFree(o.energy);
  cudaFree(o.cam_xpos); cudaFree(o.cam_xmat); cudaFree(o.light_xpos); cudaFree(o.light_xdir);
  cudaFree(o.actuator_length); cudaFree(o.actuator_moment); cudaFree(o.actuator_velocity);
  cudaFree(o.qfrc_bias);
  delete d;
}

void mjb_setStream(mjbData* d, void* cuda_stream) { d->stream = (cudaStream_t)cuda_stream; d->stream_dirty = true; }

int mjb_setState(mjbData* d, int nbatch, const mjtNum* qpos, const mjtNum* qvel, const mjtNum* qacc) {
  if (!d->shards.empty()) {
    if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_setState: nbatch out of range"; return -1; }
    const int nq = d->hdr.nq, nv = d->hdr.nv;
    return forShards(d, nbatch, false, [&](mjbData* sh, ShardRange r) {
      return mjb_setState(sh, r.n, qpos + r.first * nq, qvel + r.first * nv, qacc + r.first * nv);
    }) ? 0 : -1;
  }
  d->stream_dirty = true;
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_setState: nbatch out of range"; return -1; }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const mjbHdr& H = d->hdr;
  const size_t n = (size_t)nbatch;
  const size_t bq = n * H.nq * sizeof(double), bv = n * H.nv * sizeof(double);
  if (!ensureStage(d, bq + 2*bv)) return -1;
  char* st = (char*)d->d_stage;
  bool ok = check(d, cudaMemcpyAsync(st, qpos, bq, cudaMemcpyHostToDevice, d->stream), "H2D qpos");
  ok = ok && check(d, cudaMemcpyAsync(st + bq, qvel, bv, cudaMemcpyHostToDevice, d->stream), "H2D qvel");
  ok = ok && check(d, cudaMemcpyAsync(st + bq + bv, qacc, bv, cudaMemcpyHostToDevice, d->stream), "H2D qacc");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)st, d->d_qpos, nbatch, H.nq, d->stride, d->stream), "transpose qpos");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)(st + bq), d->d_qvel, nbatch, H.nv, d->stride, d->stream), "transpose qvel");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)(st + bq + bv), d->d_qacc, nbatch, H.nv, d->stride, d->stream), "transpose qacc");
  d->in_qpos = d->d_qpos; d->in_qvel = d->d_qvel; d->in_qacc = d->d_qacc; d->in_stride = d->stride;
  return ok ? 0 : -1;
}

// d->mocap_pos / d->mocap_quat per state (HOST, nbatch x nmocap x 3 | 4); NULL returns to the model pose
int mjb_setMocap(mjbData* d, int nbatch, const mjtNum* mocap_pos, const mjtNum* mocap_quat) {
  if (!d->shards.empty()) {
    const int nm = d->hdr.nmocap;
    return forShards(d, nbatch, false, [&](mjbData* sh, ShardRange r) {
      return mjb_setMocap(sh, r.n, mocap_pos ? mocap_pos + r.first * 3 * nm : nullptr,
                          mocap_quat ? mocap_quat + r.first * 4 * nm : nullptr);
    }) ? 0 : -1;
  }
  d->stream_dirty = true;
  const mjbHdr& H = d->hdr;
  mjb::Outputs& o = d->out;
  if (!mocap_pos || !mocap_quat || H.nmocap == 0) { o.mocap_pos = nullptr; o.mocap_quat = nullptr; return 0; }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_setMocap: nbatch out of range"; return -1; }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const size_t S = (size_t)d->stride, n = (size_t)nbatch, nm = (size_t)H.nmocap;
  bool ok = true;
  if (!d->d_mocap_pos) {
    ok = ok && devAlloc(d, &d->d_mocap_pos, 3 * nm * S, "cudaMalloc(mocap_pos)");
    ok = ok && devAlloc(d, &d->d_mocap_quat, 4 * nm * S, "cudaMalloc(mocap_quat)");
  }
  if (!ok || !ensureStage(d, n * 7 * nm * sizeof(double))) return -1;
  char* st = (char*)d->d_stage;
  const size_t bp = n * 3 * nm * sizeof(double), bq = n * 4 * nm * sizeof(double);
  ok = ok && check(d, cudaMemcpyAsync(st, mocap_pos, bp, cudaMemcpyHostToDevice, d->stream), "H2D mocap_pos");
  ok = ok && check(d, cudaMemcpyAsync(st + bp, mocap_quat, bq, cudaMemcpyHostToDevice, d->stream), "H2D mocap_quat");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)st, d->d_mocap_pos, nbatch, (int)(3 * nm), d->stride, d->stream), "transpose mocap_pos");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)(st + bp), d->d_mocap_quat, nbatch, (int)(4 * nm), d->stride, d->stream), "transpose mocap_quat");
  ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_setMocap");   // the staging buffer is shared with mjb_setState
  if (!ok) return -1;
  o.mocap_pos = d->d_mocap_pos; o.mocap_quat = d->d_mocap_quat;
  return 0;
}

int mjb_setXfrcApplied(mjbData* d, int nbatch, const mjtNum* xfrc_applied) {
  if (!d->shards.empty()) {
    const int nb = d->hdr.nbody;
    return forShards(d, nbatch, false, [&](mjbData* sh, ShardRange r) {
      return mjb_setXfrcApplied(sh, r.n, xfrc_applied ? xfrc_applied + r.first * 6 * nb : nullptr);
    }) ? 0 : -1;
  }
  d->stream_dirty = true;
  const mjbHdr& H = d->hdr;
  mjb::Outputs& o = d->out;
  if (!xfrc_applied) { o.xfrc_applied = nullptr; return 0; }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_setXfrcApplied: nbatch out of range"; return -1; }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const size_t S = (size_t)d->stride, n = (size_t)nbatch, rows = 6 * (size_t)H.nbody;
  bool ok = true;
  if (!d->d_xfrc_applied) ok = ok && devAlloc(d, &d->d_xfrc_applied, rows * S, "cudaMalloc(xfrc_applied)");
  if (!ok || !ensureStage(d, n * rows * sizeof(double))) return -1;
  ok = ok && check(d, cudaMemcpyAsync(d->d_stage, xfrc_applied, n * rows * sizeof(double), cudaMemcpyHostToDevice, d->stream), "H2D xfrc_applied");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)d->d_stage, d->d_xfrc_applied, nbatch, (int)rows, d->stride, d->stream), "transpose xfrc_applied");
  ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_setXfrcApplied");   // the staging buffer is shared with mjb_setState
  if (!ok) return -1;
  o.xfrc_applied = d->d_xfrc_applied;
  return 0;
}

int mjb_setEqActive(mjbData* d, int nbatch, const unsigned char* eq_active) {
  if (!d->shards.empty()) {
    const int ne = d->hdr.neq;
    return forShards(d, nbatch, false, [&](mjbData* sh, ShardRange r) {
      return mjb_setEqActive(sh, r.n, eq_active ? eq_active + r.first * ne : nullptr);
    }) ? 0 : -1;
  }
  d->stream_dirty = true;
  const mjbHdr& H = d->hdr;
  mjb::Outputs& o = d->out;
  if (!eq_active || H.neq == 0) { o.eq_active = nullptr; return 0; }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_setEqActive: nbatch out of range"; return -1; }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const size_t S = (size_t)d->stride, n = (size_t)nbatch, rows = (size_t)H.neq;
  bool ok = true;
  if (!d->d_eq_active) ok = ok && devAlloc(d, &d->d_eq_active, rows * S, "cudaMalloc(eq_active)");
  if (!ok || !ensureStage(d, n * rows * sizeof(double))) return -1;
  std::vector<double> flags(n * rows);                     // mjtByte -> 0 / 1 as doubles (the transposes are fp64)
  for (size_t k = 0; k < flags.size(); k++) flags[k] = eq_active[k] ? 1.0 : 0.0;
  ok = ok && check(d, cudaMemcpy(d->d_stage, flags.data(), flags.size() * sizeof(double), cudaMemcpyHostToDevice), "H2D eq_active");
  ok = ok && check(d, mjb::launch_aos_to_soa((const double*)d->d_stage, d->d_eq_active, nbatch, (int)rows, d->stride, d->stream), "transpose eq_active");
  ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_setEqActive");
  if (!ok) return -1;
  o.eq_active = d->d_eq_active;
  return 0;
}

int mjb_setStateDevice(mjbData* d, const mjtNum* qpos, const mjtNum* qvel, const mjtNum* qacc,
                       long long stride) {
  if (!d->shards.empty()) { d->error = "mjb_setStateDevice: not available on a multi-device mjbData (use the shard of one device)"; return -1; }
  if (!qpos || !qvel || !qacc) {
    d->in_qpos = d->d_qpos; d->in_qvel = d->d_qvel; d->in_qacc = d->d_qacc; d->in_stride = d->stride;
    return 0;
  }
  if (stride != d->stride) {
    d->error = "mjb_setStateDevice: stride must equal mjb_stride(d) (outputs share the row stride)";
    return -1;
  }
  d->in_qpos = qpos; d->in_qvel = qvel; d->in_qacc = qacc; d->in_stride = stride;
  return 0;
}

}  // extern "C"

namespace {

// launch the phase kernels for states [first, first + count) on the compute stream
// mj_inverse reads m->opt on every call; the batched path flattened the model at mjb_makeData. When
// the caller changed m->opt in between (flags, cone, timestep, gravity ...), the blob is rebuilt and
// uploaded again if its layout is unchanged; a change that alters the layout (e.g. contacts switched
// on for a model uploaded without candidate pairs) is an error asking for a new mjbData.
bool refreshModel(mjbData* d, const mjModel* m) {
  if (!m) return true;
  const unsigned long long h = optHash(m);
  if (h == d->opt_hash) return true;
  std::vector<unsigned char> blob;
  std::string msg;
  if (!mjb::buildModelBlob(m, blob, msg)) { d->error = "mjb_inverse: m->opt changed: " + msg; return false; }
  const mjbHdr* H = reinterpret_cast<const mjbHdr*>(blob.data());
  const mjbHdr& O = d->hdr;
  if (H->bytes != O.bytes || H->staged_bytes != O.staged_bytes || H->discrete_trn != O.discrete_trn ||
      H->sensor_cam != O.sensor_cam || H->sensor_camlight != O.sensor_camlight || H->sensor_trn != O.sensor_trn || H->sensor_energy != O.sensor_energy ||
      H->nscratch != O.nscratch || H->ncand != O.ncand || H->nsensordata != O.nsensordata ||
      H->discrete_acc != O.discrete_acc || H->sensor_post != O.sensor_post || H->sensor_touch != O.sensor_touch ||
      ((H->ncand > 0 && !(H->disableflags & (MJB_DSBL_CONSTRAINT | MJB_DSBL_CONTACT))) && !d->d_cq &&
       !(std::getenv("MJB_CONTACT_PATH")))) {
    d->error = "mjb_inverse: m->opt changed since mjb_makeData in a way that changes the device layout; "
               "create a new mjbData";
    return false;
  }
  // kernels already queued on the stream still read the old blob: order the upload after them
  if (!check(d, cudaStreamSynchronize(d->stream), "cudaStreamSynchronize")) return false;
  if (!check(d, cudaMemcpy(d->d_model, blob.data(), blob.size(), cudaMemcpyHostToDevice), "cudaMemcpy(model)")) return false;
  std::memcpy(&d->hdr, blob.data(), sizeof(mjbHdr));
  d->blob = blob;
  d->opt_hash = h;
  if (d->spec_on) {           // the specialised kernels were compiled for the old options
    mjb::jitUnload(d->spec);
    d->spec_on = false;
    if (d->spec_wanted && mjb_specialize(d, nullptr, 0)) d->error.clear();   // generic kernels otherwise
  }
  if (d->fd) { mjb_deleteData(d->fd); d->fd = nullptr; d->fd_tile = 0; }
  return true;
}

bool launchRange(mjbData* d, long long first, long long count) {
  d->stream_dirty = true;      // mjb_inverseHost clears it again after its own launches
  mjb::LaunchArgs a;
  a.model = d->d_model;
  a.model_bytes = d->model_bytes;
  a.model_in_smem = d->model_in_smem;
  a.sensor_cold = d->hdr.sensor_cam;
  a.qpos = d->in_qpos; a.qvel = d->in_qvel; a.qacc = d->in_qacc;
  a.qacc_discrete = d->d_qacc_discrete;
  a.scratch = d->d_scratch;
  a.iscratch = d->d_iscratch;
  a.chunk_stride = d->chunk_stride;
  a.nscratch = d->hdr.nscratch;
  a.niscratch = mjb::isc_rows(d->hdr);
  a.stride = d->stride;
  a.nconmax = d->nconmax;
  a.njmax = d->njmax;
  a.has_contacts = d->hdr.ncand > 0 &&
                   !(d->hdr.disableflags & (MJB_DSBL_CONSTRAINT | MJB_DSBL_CONTACT));
  a.max_pair_contacts = d->hdr.max_pair_contacts;
  a.simple_pairs = d->hdr.simple_pairs;
  a.has_convex = d->hdr.has_convex;
  a.sensor_ccd = d->hdr.sensor_ccd;
  a.cq = d->d_cq; a.items = d->d_items; a.item_con = d->d_item_con; a.contacts = d->d_contacts;
  a.items_cap = d->items_cap; a.contacts_cap = d->contacts_cap;
  a.slot_rec = d->d_slot_rec;
  a.has_gravcomp = d->hdr.passive_wrench;
  a.has_spatial = d->hdr.has_spatial || d->hdr.has_fluid;    // selects the smooth kernel with the rarer features
  a.skip_sensors = d->skip_sensors;
  a.scan_ngeom = d->hdr.ngeom;
  a.pair_ci = d->d_pair_ci;
  a.scan_buf = d->d_scan_buf; a.scan_buf_cap = mjb::scan_wide_buf_cap(d->hdr.ncand);
  a.sub_nv = d->hdr.nv; a.sub_nbody = d->hdr.nbody; a.sub_nC = d->hdr.nC;
  {
    // MJB_INERTIA=thread | subwarp overrides the choice (A/B measurements)
    const char* env = std::getenv("MJB_INERTIA");
    const bool fits = mjb::inertia_subwarp_fits(d->hdr.nv, d->hdr.nbody, d->hdr.nC);
    a.inertia_subwarp = (fits && env && !std::strcmp(env, "subwarp")) ? 1 : 0;
  }
  a.scan_wide = (std::getenv("MJB_SCAN_FLAT") || !d->d_scan_buf) ? 0 : mjb::scan_wide_states(d->hdr.ncand, d->hdr.ngeom);
  a.cmask = a.scan_wide > 0 ? d->d_cmask : nullptr;
  a.out = d->out;
  // chunks reuse the same intermediates; kernels of consecutive chunks serialise on the stream
  for (long long start = first; start < first + count; start += d->chunk_stride) {
    a.chunk_start = start;
    a.chunk_n = (int)((first + count - start) < d->chunk_stride ? (first + count - start) : d->chunk_stride);
    int launches = 0;
    mjb::PhaseTimer timer;
    timer.ctx = d;
    timer.next_event = [](void* ctx) -> cudaEvent_t {
      mjbData* dd = static_cast<mjbData*>(ctx);
      if (dd->ev_used == dd->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        dd->ev_pool.push_back(e);
      }
      return dd->ev_pool[dd->ev_used++];
    };
    timer.mark = [](void* ctx, int phase, cudaEvent_t b, cudaEvent_t e) {
      static_cast<mjbData*>(ctx)->marks.push_back({phase, b, e});
    };
    if (!check(d, mjb::launch_inverse(a, d->stream, &launches, d->phase_timing ? &timer : nullptr,
                                      d->spec_on ? &d->spec : nullptr),
               "launch mj_inverse kernels")) return false;
    d->kernel_launches += launches;
  }
  return true;
}

}  // namespace

extern "C" {

int mjb_inverseAsync(const mjModel* m, mjbData* d, int nbatch) {
  if (!d->shards.empty()) {
    if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_inverse: nbatch out of range"; return -1; }
    d->last_nbatch = nbatch;
    return forShards(d, nbatch, false, [&](mjbData* sh, ShardRange r) { return mjb_inverseAsync(m, sh, r.n); }) ? 0 : -1;
  }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_inverse: nbatch out of range"; return -1; }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  if (!refreshModel(d, m)) return -1;
  d->last_nbatch = nbatch;
  return launchRange(d, 0, nbatch) ? 0 : -1;
}

// Host-to-host mj_inverse over the batch: for i: copy state i -> mj_inverse -> copy qfrc_inverse,
// as one call. The batch is cut into pieces that flow through a three-stage pipeline -- H2D copy,
// kernels, D2H copy -- on three streams with double-buffered staging, so PCIe transfers in both
// directions overlap the compute of neighbouring pieces. Host buffers should be pinned.
int mjb_inverseHost(const mjModel* m, mjbData* d, int nbatch, const mjtNum* qpos, const mjtNum* qvel,
                    const mjtNum* qacc, mjtNum* qfrc_inverse) {
  if (!d->shards.empty()) {
    // one host thread per device, each driving its own three-stream pipeline on its slice of the
    // caller's arrays (python/mujoco/rollout.cc:180-210 chunks the batch over a thread pool the same way)
    if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_inverseHost: nbatch out of range"; return -1; }
    d->last_nbatch = nbatch;
    const int nq = d->hdr.nq, nv = d->hdr.nv;
    return forShards(d, nbatch, true, [&](mjbData* sh, ShardRange r) {
      return mjb_inverseHost(m, sh, r.n, qpos + r.first * nq, qvel + r.first * nv, qacc + r.first * nv,
                             qfrc_inverse + r.first * nv);
    }) ? 0 : -1;
  }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_inverseHost: nbatch out of range"; return -1; }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  if (!refreshModel(d, m)) return -1;
  const mjbHdr& H = d->hdr;
  d->last_nbatch = nbatch;
  d->in_qpos = d->d_qpos; d->in_qvel = d->d_qvel; d->in_qacc = d->d_qacc; d->in_stride = d->stride;
  if (nbatch == 0) return 0;
  // states per pipeline piece: large enough for the kernels to fill the GPU, small enough that the
  // first copy-in and the last copy-out of a call stay short (MJB_HOST_PIECE overrides, for tuning)
  size_t piece = d->pipe_piece ? d->pipe_piece : 262144;   // measured best of 2^16 .. 2^19 (humanoid, 2^20 states)
  if (!d->pipe_piece) {
    const char* env = std::getenv("MJB_HOST_PIECE");
    if (env && std::atol(env) >= 1024) piece = (size_t)std::atol(env) & ~(size_t)127;
  }
  const size_t in_doubles = piece * (size_t)(H.nq + 2*H.nv), out_doubles = piece * (size_t)H.nv;
  bool ok = true;
  if (!d->s_in) {
    ok = ok && check(d, cudaStreamCreateWithFlags(&d->s_in, cudaStreamNonBlocking), "cudaStreamCreate");
    ok = ok && check(d, cudaStreamCreateWithFlags(&d->s_out, cudaStreamNonBlocking), "cudaStreamCreate");
    ok = ok && check(d, cudaStreamCreateWithFlags(&d->s_comp, cudaStreamNonBlocking), "cudaStreamCreate");
    ok = ok && check(d, cudaEventCreateWithFlags(&d->ev_start, cudaEventDisableTiming), "cudaEventCreate");
    for (int b = 0; b < 2 && ok; b++) {
      ok = ok && check(d, cudaMalloc(&d->pipe_in[b], in_doubles * sizeof(double)), "cudaMalloc(pipe_in)");
      ok = ok && check(d, cudaMalloc(&d->pipe_out[b], out_doubles * sizeof(double)), "cudaMalloc(pipe_out)");
      ok = ok && check(d, cudaEventCreateWithFlags(&d->ev_in[b], cudaEventDisableTiming), "cudaEventCreate");
      ok = ok && check(d, cudaEventCreateWithFlags(&d->ev_tr[b], cudaEventDisableTiming), "cudaEventCreate");
      ok = ok && check(d, cudaEventCreateWithFlags(&d->ev_comp[b], cudaEventDisableTiming), "cudaEventCreate");
      ok = ok && check(d, cudaEventCreateWithFlags(&d->ev_out[b], cudaEventDisableTiming), "cudaEventCreate");
    }
    d->pipe_piece = piece;
    if (!ok) return -1;
  }
  // The host arrays are read from the moment of the call (like cudaMemcpy from host memory): the
  // copy-in stream does NOT wait for the caller's stream, so the copies of this call overlap the
  // kernels of the previous one; the staging buffers are handed over through their own events.
  // Transposes and kernels run on the pipeline's own compute stream, so that the kernels of this
  // call follow those of the previous call directly; the caller's stream only joins at the end
  // (it completes when the last results have landed) and is waited for only if something else was
  // queued on it in between.
  cudaStream_t user = d->stream;
  if (d->stream_dirty) {
    ok = ok && check(d, cudaEventRecord(d->ev_start, user), "cudaEventRecord");
    ok = ok && check(d, cudaStreamWaitEvent(d->s_comp, d->ev_start, 0), "cudaStreamWaitEvent");
  }
  d->stream = d->s_comp;
  int p = 0;
  for (size_t first = 0; first < (size_t)nbatch && ok; first += piece, p++) {
    const int b = (int)(d->pipe_seq++ & 1u);
    const size_t n = ((size_t)nbatch - first) < piece ? ((size_t)nbatch - first) : piece;
    const size_t bq = n * H.nq * sizeof(double), bv = n * H.nv * sizeof(double);
    char* st = (char*)d->pipe_in[b];
    // stage 1: host -> device (waits until the transposes of piece p-2 released this buffer)
    if (d->pipe_used[b]) ok = ok && check(d, cudaStreamWaitEvent(d->s_in, d->ev_tr[b], 0), "wait tr");
    ok = ok && check(d, cudaMemcpyAsync(st, qpos + first * H.nq, bq, cudaMemcpyHostToDevice, d->s_in), "H2D qpos");
    ok = ok && check(d, cudaMemcpyAsync(st + bq, qvel + first * H.nv, bv, cudaMemcpyHostToDevice, d->s_in), "H2D qvel");
    ok = ok && check(d, cudaMemcpyAsync(st + bq + bv, qacc + first * H.nv, bv, cudaMemcpyHostToDevice, d->s_in), "H2D qacc");
    ok = ok && check(d, cudaEventRecord(d->ev_in[b], d->s_in), "record in");
    // stage 2: transposes + kernels on the compute stream
    ok = ok && check(d, cudaStreamWaitEvent(d->stream, d->ev_in[b], 0), "wait in");
    ok = ok && check(d, mjb::launch_aos_to_soa((const double*)st, d->d_qpos + first, (int)n, H.nq, d->stride, d->stream), "transpose qpos");
    ok = ok && check(d, mjb::launch_aos_to_soa((const double*)(st + bq), d->d_qvel + first, (int)n, H.nv, d->stride, d->stream), "transpose qvel");
    ok = ok && check(d, mjb::launch_aos_to_soa((const double*)(st + bq + bv), d->d_qacc + first, (int)n, H.nv, d->stride, d->stream), "transpose qacc");
    ok = ok && check(d, cudaEventRecord(d->ev_tr[b], d->stream), "record tr");
    ok = ok && launchRange(d, (long long)first, (long long)n);
    if (d->pipe_used[b]) ok = ok && check(d, cudaStreamWaitEvent(d->stream, d->ev_out[b], 0), "wait out");
    ok = ok && check(d, mjb::launch_soa_to_aos(d->out.qfrc_inverse + first, (double*)d->pipe_out[b], (int)n, H.nv, d->stride, d->stream), "transpose out");
    ok = ok && check(d, cudaEventRecord(d->ev_comp[b], d->stream), "record comp");
    // stage 3: device -> host
    ok = ok && check(d, cudaStreamWaitEvent(d->s_out, d->ev_comp[b], 0), "wait comp");
    ok = ok && check(d, cudaMemcpyAsync(qfrc_inverse + first * H.nv, d->pipe_out[b], bv, cudaMemcpyDeviceToHost, d->s_out), "D2H qfrc");
    ok = ok && check(d, cudaEventRecord(d->ev_out[b], d->s_out), "record out");
    d->pipe_used[b] = true;
  }
  // the caller's stream completes when the last copies have landed
  d->stream = user;
  d->stream_dirty = false;
  for (int b = 0; b < 2 && ok; b++) {
    if (d->pipe_used[b]) ok = ok && check(d, cudaStreamWaitEvent(d->stream, d->ev_out[b], 0), "join");
  }
  return ok ? 0 : -1;
}

int mjb_inverse(const mjModel* m, mjbData* d, int nbatch) {
  if (!d->shards.empty()) {
    if (mjb_inverseAsync(m, d, nbatch)) return -1;       // queued on every device first, then counted
    int total = 0;
    const bool ok = forShards(d, nbatch, false, [&](mjbData* sh, ShardRange r) {
      int count = 0;
      bool k = check(sh, cudaSetDevice(sh->device), "cudaSetDevice");
      k = k && check(sh, cudaMemsetAsync(sh->d_counter, 0, sizeof(int), sh->stream), "memset counter");
      k = k && check(sh, mjb::launch_count_nonzero(sh->out.status, r.n, sh->d_counter, sh->stream), "count status");
      k = k && check(sh, cudaMemcpyAsync(&count, sh->d_counter, sizeof(int), cudaMemcpyDeviceToHost, sh->stream), "D2H counter");
      k = k && check(sh, cudaStreamSynchronize(sh->stream), "mjb_inverse");
      total += count;
      return k ? 0 : -1;
    });
    return ok ? total : -1;
  }
  if (mjb_inverseAsync(m, d, nbatch)) return -1;
  int count = 0;
  bool ok = check(d, cudaMemsetAsync(d->d_counter, 0, sizeof(int), d->stream), "memset counter");
  ok = ok && check(d, mjb::launch_count_nonzero(d->out.status, nbatch, d->d_counter, d->stream), "count status");
  ok = ok && check(d, cudaMemcpyAsync(&count, d->d_counter, sizeof(int), cudaMemcpyDeviceToHost, d->stream), "D2H counter");
  ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_inverse");
  return ok ? count : -1;
}

// mj_inverseSkip(m, d, skipstage, skipsensor) over the batch (include/mujoco/mujoco.h:137,
// engine_inverse.c:197-261). skipstage tells the CPU engine which stages it may REUSE from the
// previous call on the same mjData; the results equal a full evaluation whenever the caller kept
// the corresponding inputs unchanged, which is the function's contract. Here every stage is
// recomputed (one fused sweep: nothing is cached per state between calls), so the value is
// validated and otherwise ignored. skipsensor != 0 leaves sensordata as the last evaluation wrote
// it, like the reference.
int mjb_inverseSkip(const mjModel* m, mjbData* d, int nbatch, int skipstage, int skipsensor) {
  if (skipstage < mjSTAGE_NONE || skipstage > mjSTAGE_VEL) {
    d->error = "mjb_inverseSkip: skipstage must be mjSTAGE_NONE, mjSTAGE_POS or mjSTAGE_VEL";
    return -1;
  }
  d->skip_sensors = skipsensor != 0;
  const int r = mjb_inverse(m, d, nbatch);
  d->skip_sensors = 0;
  return r;
}

// mjd_inverseFD over the batch (src/engine/engine_derivative_fd.c:611, flg_actuation = 0, no
// sensors): forward differences of qfrc_inverse (and of qM) with respect to qacc, qvel and qpos
// (the latter through mj_integratePos), 1 + 3 nv evaluations per state. The perturbed states are
// generated on the device and evaluated as one large batch by the same phase kernels.
int mjb_inverseFD(const mjModel* m, mjbData* d, int nbatch, mjtNum eps, mjtNum* DfDq, mjtNum* DfDv,
                  mjtNum* DfDa, mjtNum* DmDq) {
  return mjb_inverseFDSensor(m, d, nbatch, eps, 0, DfDq, DfDv, DfDa, nullptr, nullptr, nullptr, DmDq);
}

int mjb_inverseFDSensor(const mjModel* m, mjbData* d, int nbatch, mjtNum eps, int flg_actuation,
                        mjtNum* DfDq, mjtNum* DfDv, mjtNum* DfDa,
                        mjtNum* DsDq, mjtNum* DsDv, mjtNum* DsDa, mjtNum* DmDq) {
  if (flg_actuation) { d->error = "mjb_inverseFD: flg_actuation is not supported (no actuation on the inverse path)"; return -1; }
  if (!d->shards.empty()) { d->error = "mjb_inverseFD: not available on a multi-device mjbData"; return -1; }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_inverseFD: nbatch out of range"; return -1; }
  if (!(eps > 0)) { d->error = "mjb_inverseFD: eps must be positive"; return -1; }
  // the reference's own refusals (engine_derivative_fd.c:618-626)
  if (m->opt.integrator == mjINT_RK4) { d->error = "mjb_inverseFD: RK4 integrator is not supported"; return -1; }
  if (m->opt.noslip_iterations) { d->error = "mjb_inverseFD: noslip solver is not supported"; return -1; }
  // the perturbed copies are evaluated in an inner batch that holds no per-state mocap poses
  if (d->d_mocap_pos || d->d_mocap_quat) {
    d->error = "mjb_inverseFD: per-state mocap poses (mjb_setMocap) are not carried into the perturbed batch";
    return -1;
  }
  if (d->out.xfrc_applied && (DsDq || DsDv || DsDa)) {
    d->error = "mjb_inverseFD: per-state applied wrenches (mjb_setXfrcApplied) are not carried into the perturbed batch";
    return -1;
  }
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const mjbHdr& H = d->hdr;
  const int nv = H.nv, nM = H.nM, nvar = 1 + 3*nv, ns = H.nsensordata;
  if (nbatch == 0 || nv == 0) return 0;
  const bool want_mass = DmDq != nullptr;
  const bool want_sensors = (DsDq || DsDv || DsDa) && ns > 0;
  if (want_sensors && !d->out.sensordata) {
    d->error = "mjb_inverseFD: sensor Jacobians requested but the batch evaluates no sensors (mjDSBL_SENSOR)";
    return -1;
  }
  int tile = (1 << 20) / nvar;
  if (tile < 1) tile = 1;
  if (tile > nbatch) tile = nbatch;
  if (!d->fd || d->fd_tile < tile || (want_mass && !d->fd->out.qM)) {
    if (d->fd) mjb_deleteData(d->fd);
    char err[512];
    d->fd = mjb_makeData(m, tile * nvar, d->device, want_mass ? mjbOUT_INERTIA : 0, 1, 1, err, sizeof(err));
    if (!d->fd) { d->error = std::string("mjb_inverseFD: ") + err; return -1; }
    d->fd_tile = tile;
  }
  mjbData* x = d->fd;
  x->stream = d->stream;
  x->skip_sensors = want_sensors ? 0 : 1;      // skipsensor = !DsDq && !DsDv && !DsDa (engine_derivative_fd.c:629)
  size_t widest = (size_t)nv;
  if (want_mass && (size_t)nM > widest) widest = (size_t)nM;
  if (want_sensors && (size_t)ns > widest) widest = (size_t)ns;
  const size_t need = (size_t)tile * nv * widest;
  if (need > d->fd_out_doubles) {
    cudaFree(d->d_fd_out);
    d->d_fd_out = nullptr; d->fd_out_doubles = 0;
    if (!check(d, cudaMalloc((void**)&d->d_fd_out, need * sizeof(double)), "cudaMalloc(fd out)")) return -1;
    d->fd_out_doubles = need;
  }
  bool ok = true;
  for (int first = 0; first < nbatch && ok; first += tile) {
    const int n = (nbatch - first) < tile ? (nbatch - first) : tile;
    ok = ok && check(d, mjb::launch_fd_expand(d->d_model, d->in_qpos, d->in_qvel, d->in_qacc, d->in_stride,
                                              first, n, eps, x->d_qpos, x->d_qvel, x->d_qacc, x->stride,
                                              d->stream), "fd expand");
    x->in_qpos = x->d_qpos; x->in_qvel = x->d_qvel; x->in_qacc = x->d_qacc; x->in_stride = x->stride;
    if (ok && mjb_inverseAsync(m, x, n * nvar)) { d->error = "mjb_inverseFD: " + x->error; return -1; }
    struct Job { mjtNum* host; const double* field; int v0; int ncol; };
    const Job jobs[7] = {{DfDa, x->out.qfrc_inverse, 1, nv}, {DfDv, x->out.qfrc_inverse, 1 + nv, nv},
                         {DfDq, x->out.qfrc_inverse, 1 + 2*nv, nv}, {DmDq, x->out.qM, 1 + 2*nv, nM},
                         {want_sensors ? DsDa : nullptr, x->out.sensordata, 1, ns},
                         {want_sensors ? DsDv : nullptr, x->out.sensordata, 1 + nv, ns},
                         {want_sensors ? DsDq : nullptr, x->out.sensordata, 1 + 2*nv, ns}};
    for (const Job& j : jobs) {
      if (!j.host) continue;
      ok = ok && check(d, mjb::launch_fd_diff(j.field, x->stride, n, nvar, j.v0, nv, j.ncol, eps,
                                              d->d_fd_out, d->stream), "fd diff");
      const size_t bytes = (size_t)n * nv * j.ncol * sizeof(double);
      ok = ok && check(d, cudaMemcpyAsync(j.host + (size_t)first * nv * j.ncol, d->d_fd_out, bytes,
                                          cudaMemcpyDeviceToHost, d->stream), "D2H fd");
      ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_inverseFD");   // d_fd_out is reused
    }
  }
  return ok ? 0 : -1;
}

// mj_compareFwdInv over the batch (include/mujoco/mujoco.h mj_compareFwdInv, engine_inverse.c:275-316).
// The caller ran the forward dynamics; the states given to mjb_setState carry its qacc. Host arrays,
// row-major per state: qfrc_applied, qfrc_actuator nbatch x nv (either may be NULL = zero),
// xfrc_applied nbatch x nbody x 6 (force, torque; may be NULL), qfrc_constraint nbatch x nv of the
// forward pass; fwdinv out nbatch x 2. Sensors are not re-evaluated (skipsensor = 1 in the reference).
int mjb_compareFwdInv(const mjModel* m, mjbData* d, int nbatch, const mjtNum* qfrc_applied,
                      const mjtNum* qfrc_actuator, const mjtNum* xfrc_applied,
                      const mjtNum* qfrc_constraint, mjtNum* fwdinv) {
  if (!d->shards.empty()) { d->error = "mjb_compareFwdInv: not available on a multi-device mjbData"; return -1; }
  if (nbatch < 0 || nbatch > d->nbatch_max) { d->error = "mjb_compareFwdInv: nbatch out of range"; return -1; }
  if (!qfrc_constraint || !fwdinv) { d->error = "mjb_compareFwdInv: qfrc_constraint and fwdinv are required"; return -1; }
  if (nbatch == 0) return 0;
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const mjbHdr& H = d->hdr;
  const size_t S = (size_t)d->stride, n = (size_t)nbatch, nv = (size_t)H.nv, nb6 = 6 * (size_t)H.nbody;
  mjb::Outputs& o = d->out;
  bool ok = true;
  if (!d->d_fwd_qforce) {
    ok = ok && devAlloc(d, &d->d_fwd_qforce, nv * S, "cudaMalloc(fwd qforce)");
    ok = ok && devAlloc(d, &d->d_fwd_xfrc, nb6 * S, "cudaMalloc(fwd xfrc)");
    ok = ok && devAlloc(d, &d->d_fwd_qc, nv * S, "cudaMalloc(fwd qfrc_constraint)");
    ok = ok && devAlloc(d, &d->d_fwdinv, 2 * S, "cudaMalloc(fwdinv)");
  }
  if (ok && !o.qfrc_constraint) {
    ok = devAlloc(d, &o.qfrc_constraint, nv * S, "cudaMalloc(qfrc_constraint)");
    d->own_qfrc_constraint = true;
  }
  if (!ok) return -1;
  // qforce = qfrc_applied + qfrc_actuator on the host (a validation utility, not a hot path)
  std::vector<double> qforce(n * nv, 0.0);
  for (size_t k = 0; k < n * nv; k++) {
    qforce[k] = (qfrc_applied ? qfrc_applied[k] : 0.0) + (qfrc_actuator ? qfrc_actuator[k] : 0.0);
  }
  const size_t bmax = n * (nb6 > nv ? nb6 : nv) * sizeof(double);
  if (!ensureStage(d, bmax)) return -1;
  struct Up { const double* host; double* dev; size_t rows; };
  const Up ups[3] = {{qforce.data(), d->d_fwd_qforce, nv}, {qfrc_constraint, d->d_fwd_qc, nv},
                     {xfrc_applied, d->d_fwd_xfrc, nb6}};
  for (const Up& u : ups) {
    if (!u.host) continue;
    ok = ok && check(d, cudaMemcpyAsync(d->d_stage, u.host, n * u.rows * sizeof(double), cudaMemcpyHostToDevice, d->stream), "H2D fwd");
    ok = ok && check(d, mjb::launch_aos_to_soa((const double*)d->d_stage, u.dev, nbatch, (int)u.rows, d->stride, d->stream), "transpose fwd");
    ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_compareFwdInv");   // the staging buffer is reused
  }
  if (!ok) return -1;
  o.fwd_qforce = d->d_fwd_qforce; o.fwd_qfrc_constraint = d->d_fwd_qc;
  o.fwd_xfrc = xfrc_applied ? d->d_fwd_xfrc : nullptr;
  o.fwdinv = d->d_fwdinv;
  d->skip_sensors = 1;
  const int r = mjb_inverseAsync(m, d, nbatch);
  d->skip_sensors = 0;
  o.fwd_qforce = nullptr; o.fwd_qfrc_constraint = nullptr; o.fwd_xfrc = nullptr; o.fwdinv = nullptr;
  if (r) return -1;
  ok = ok && check(d, mjb::launch_soa_to_aos(d->d_fwdinv, (double*)d->d_stage, nbatch, 2, d->stride, d->stream), "transpose fwdinv");
  ok = ok && check(d, cudaMemcpyAsync(fwdinv, d->d_stage, n * 2 * sizeof(double), cudaMemcpyDeviceToHost, d->stream), "D2H fwdinv");
  ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_compareFwdInv");
  return ok ? 0 : -1;
}

int mjb_get(mjbData* d, int field, void* host_out) {
  if (!d->shards.empty()) {
    if (field < 0 || field >= mjbF_COUNT || !d->shards[0]->field_set[field]) {
      d->error = "mjb_get: field was not requested in outmask";
      return -1;
    }
    const size_t row_bytes = (size_t)d->field_rows[field] * (d->field_isint[field] ? sizeof(int) : sizeof(double));
    return forShards(d, d->last_nbatch, true, [&](mjbData* sh, ShardRange r) {
      return r.n ? mjb_get(sh, field, (char*)host_out + (size_t)r.first * row_bytes) : 0;
    }) ? 0 : -1;
  }
  if (field < 0 || field >= mjbF_COUNT || !d->field_set[field]) {
    d->error = "mjb_get: field was not requested in outmask";
    return -1;
  }
  if (!d->field_ptr[field]) return 0;       // requested, but the model has no rows of it (nv == 0 ...)
  if (!check(d, cudaSetDevice(d->device), "cudaSetDevice")) return -1;
  const int n = d->last_nbatch, rows = d->field_rows[field];
  const size_t esz = d->field_isint[field] ? sizeof(int) : sizeof(double);
  const size_t bytes = (size_t)n * rows * esz;
  if (bytes == 0) return 0;
  if (!ensureStage(d, bytes)) return -1;
  bool ok;
  if (d->field_isint[field]) {
    ok = check(d, mjb::launch_soa_to_aos_int((const int*)d->field_ptr[field], (int*)d->d_stage, n, rows, d->stride, d->stream), "transpose out");
  } else {
    ok = check(d, mjb::launch_soa_to_aos((const double*)d->field_ptr[field], (double*)d->d_stage, n, rows, d->stride, d->stream), "transpose out");
  }
  ok = ok && check(d, cudaMemcpyAsync(host_out, d->d_stage, bytes, cudaMemcpyDeviceToHost, d->stream), "D2H field");
  ok = ok && check(d, cudaStreamSynchronize(d->stream), "mjb_get");
  return ok ? 0 : -1;
}

int mjb_getQfrcInverse(mjbData* d, mjtNum* qfrc_inverse) { return mjb_get(d, mjbF_QFRC_INVERSE, qfrc_inverse); }

const void* mjb_devicePtr(mjbData* d, int field) {
  if (!d->shards.empty()) return nullptr;    // per-device views: use one mjbData per device
  return (field >= 0 && field < mjbF_COUNT) ? d->field_ptr[field] : nullptr;
}

int mjb_fieldRows(const mjbData* d, int field) {
  return (field >= 0 && field < mjbF_COUNT) ? d->field_rows[field] : -1;
}

long long mjb_stride(const mjbData* d) { return d->stride; }

int mjb_internalSlot(const mjbData* d, const char* name, int* offset, int* size) {
  for (int s = 0; s < MJB_SC_COUNT; s++) {
    if (!std::strcmp(mjb::scratchSlotName(s), name)) {
      *offset = d->hdr.scoff[s];
      *size = (s + 1 < MJB_SC_COUNT ? d->hdr.scoff[s + 1] : d->hdr.nscratch) - d->hdr.scoff[s];
      return 0;
    }
  }
  return -1;
}

int mjb_internalSize(const mjbData* d) { return d->hdr.nscratch; }

int mjb_ncandidate(const mjbData* d) { return d->hdr.ncand; }

void mjb_candidate(const mjbData* d, int i, int* geom1, int* geom2, int* func) {
  *geom1 = d->cand[3*i]; *geom2 = d->cand[3*i + 1]; *func = d->cand[3*i + 2];
}

const char* mjb_lastError(const mjbData* d) { return d->error.c_str(); }

long long mjb_kernelLaunches(const mjbData* d) {
  if (!d->shards.empty()) {
    long long t = 0;
    for (const mjbData* sh : d->shards) t += sh->kernel_launches;
    return t;
  } return d->kernel_launches; }

int mjb_debugQueue(mjbData* d, int* out4) {
  if (!d->d_cq) return -1;
  cudaSetDevice(d->device);
  int q[5];
  if (cudaMemcpy(q, d->d_cq, sizeof q, cudaMemcpyDeviceToHost) != cudaSuccess) return -2;
  out4[0] = q[0]; out4[1] = q[1]; out4[3] = q[3];
  out4[2] = (q[2] ? 1 : 0) | (q[4] ? 2 : 0);     // bit0: item list overflowed, bit1: contact list overflowed
  return 0;
}

int mjb_lastBatch(const mjbData* d) { return d->last_nbatch; }

void mjb_phaseTiming(mjbData* d, int enable) {
  d->phase_timing = enable != 0;
  d->marks.clear();
  d->ev_used = 0;
}

int mjb_phaseTimes(mjbData* d, double* ms, int n) {
  cudaSetDevice(d->device);
  if (!check(d, cudaStreamSynchronize(d->stream), "mjb_phaseTimes")) return -1;
  for (int i = 0; i < n; i++) ms[i] = 0;
  for (const auto& m : d->marks) {
    float t = 0;
    if (cudaEventElapsedTime(&t, m.b, m.e) == cudaSuccess && m.phase < n) ms[m.phase] += t;
  }
  d->marks.clear();
  d->ev_used = 0;
  return 0;
}

int mjb_synchronize(mjbData* d) {
  if (!d->shards.empty()) {
    int rc = 0;
    for (mjbData* sh : d->shards) if (mjb_synchronize(sh)) rc = -1;
    return rc;
  }
  cudaSetDevice(d->device);
  return check(d, cudaStreamSynchronize(d->stream), "cudaStreamSynchronize") ? 0 : -1;
}

double mjb_fp64PeakTflops(int device) {
  if (cudaSetDevice(device) != cudaSuccess) return -1;
  float ms = 0;
  double flops = 0, best = 0;
  for (int rep = 0; rep < 3; rep++) {
    if (mjb::dfma_peak_probe(20000, &ms, &flops, 0) != cudaSuccess || ms <= 0) return -1;
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (tf > best) best = tf;
  }
  return best;
}

}  // extern "C"
