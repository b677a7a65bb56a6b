// Host-side model flattening for libmjb (see mjb_upload.cc).
#ifndef MJB_UPLOAD_H_
#define MJB_UPLOAD_H_

#include <string>
#include <vector>

struct mjModel_;

namespace mjb {

// Validate `m` and build the device model blob (mjb_model.h). Returns false and fills `err`
// when the model uses a feature outside the supported path.
bool buildModelBlob(const mjModel_* m, std::vector<unsigned char>& blob, std::string& err);

// mj_isSparse
bool isSparseJacobian(const mjModel_* m);

// name of a per-thread scratch slot (MJB_SC_*), nullptr when out of range
const char* scratchSlotName(int slot);

}  // namespace mjb

#endif  // MJB_UPLOAD_H_
