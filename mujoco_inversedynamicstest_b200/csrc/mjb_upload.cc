// Model upload: validate an mjModel and flatten what the batched mj_inverse kernels read into the
// single device blob described in mjb_model.h.
//
// Host-only code; the only file of libmjb.so that sees the reference's struct layouts
// (<mujoco/mujoco.h>). Runs once per mjb_makeData.
#include "mjb_upload.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <set>
#include <string>
#include <vector>

#include <mujoco/mujoco.h>

#include "mjb_model.h"

namespace mjb {
namespace {

// engine_collision_driver.c:100
bool filterBitmask(int contype1, int conaffinity1, int contype2, int conaffinity2) {
  return !(contype1 & conaffinity2) && !(contype2 & conaffinity1);
}

// engine_collision_driver.c:168
bool filterBodyPair(int weldbody1, int weldparent1, int weldbody2, int weldparent2,
                    bool dsbl_filterparent) {
  if (weldbody1 == weldbody2) return true;
  if ((!dsbl_filterparent && weldbody1 != 0 && weldbody2 != 0) &&
      (weldbody1 == weldparent2 || weldbody2 == weldparent1)) return true;
  return false;
}

bool canCollide(const mjModel* m, int b) {  // engine_collision_driver.c:187
  return m->body_contype[b] || m->body_conaffinity[b];
}

bool hasPlane(const mjModel* m, int b) {
  for (int g = m->body_geomadr[b]; g < m->body_geomadr[b] + m->body_geomnum[b]; g++) {
    if (m->geom_type[g] == mjGEOM_PLANE) return true;
  }
  return false;
}

// add_pair's geom-level contype/conaffinity OR test (engine_collision_driver.c:937-990)
bool bodyGeomMasksCompatible(const mjModel* m, int b1, int b2) {
  int ct1 = 0, ca1 = 0, ct2 = 0, ca2 = 0;
  for (int g = m->body_geomadr[b1]; g < m->body_geomadr[b1] + m->body_geomnum[b1]; g++) {
    ct1 |= m->geom_contype[g]; ca1 |= m->geom_conaffinity[g];
  }
  for (int g = m->body_geomadr[b2]; g < m->body_geomadr[b2] + m->body_geomnum[b2]; g++) {
    ct2 |= m->geom_contype[g]; ca2 |= m->geom_conaffinity[g];
  }
  return (ct1 & ca2) || (ct2 & ca1);
}

// narrow-phase id of a type-ordered geom pair; -1: no function in the reference's table (pair
// silently skipped there too); -2: a function this library does not implement
int narrowphaseId(int t1, int t2) {
  switch (t1) {
    case mjGEOM_PLANE:
      switch (t2) {
        case mjGEOM_PLANE: case mjGEOM_HFIELD: return -1;
        case mjGEOM_SPHERE: return MJB_FN_PLANE_SPHERE;
        case mjGEOM_CAPSULE: return MJB_FN_PLANE_CAPSULE;
        case mjGEOM_CYLINDER: return MJB_FN_PLANE_CYLINDER;
        case mjGEOM_BOX: return MJB_FN_PLANE_BOX;
        case mjGEOM_ELLIPSOID: return MJB_FN_PLANE_ELLIPSOID;
        default: return -2;
      }
    case mjGEOM_HFIELD:
      return (t2 == mjGEOM_HFIELD) ? -1 : -2;
    case mjGEOM_SPHERE:
      switch (t2) {
        case mjGEOM_SPHERE: return MJB_FN_SPHERE_SPHERE;
        case mjGEOM_CAPSULE: return MJB_FN_SPHERE_CAPSULE;
        case mjGEOM_ELLIPSOID: return MJB_FN_CONVEX;
        case mjGEOM_CYLINDER: return MJB_FN_SPHERE_CYLINDER;
        case mjGEOM_BOX: return MJB_FN_SPHERE_BOX;
        default: return -2;
      }
    case mjGEOM_CAPSULE:
      switch (t2) {
        case mjGEOM_CAPSULE: return MJB_FN_CAPSULE_CAPSULE;
        case mjGEOM_ELLIPSOID: case mjGEOM_CYLINDER: return MJB_FN_CONVEX;
        case mjGEOM_BOX: return MJB_FN_CAPSULE_BOX;
        default: return -2;
      }
    case mjGEOM_ELLIPSOID:
      return (t2 == mjGEOM_ELLIPSOID || t2 == mjGEOM_CYLINDER || t2 == mjGEOM_BOX) ? MJB_FN_CONVEX : -2;
    case mjGEOM_CYLINDER:
      return (t2 == mjGEOM_CYLINDER || t2 == mjGEOM_BOX) ? MJB_FN_CONVEX : -2;
    case mjGEOM_BOX:
      return (t2 == mjGEOM_BOX) ? MJB_FN_BOX_BOX : -2;
    default:
      return -2;
  }
}

const char* geomTypeName(int t) {
  static const char* names[] = {"plane", "hfield", "sphere", "capsule", "ellipsoid", "cylinder",
                                "box", "mesh", "sdf"};
  return (t >= 0 && t < 9) ? names[t] : "?";
}

// getsolparam's reference checks and solimp clamps (engine_core_constraint.c:1362-1384) and the
// K, B of mj_makeImpedance (:1523-1545), all functions of model constants only.
void makeSolParam(const mjModel* m, const mjtNum* solref_in, const mjtNum* solimp_in, double* sp) {
  mjtNum solref[mjNREF] = {solref_in[0], solref_in[1]};
  mjtNum solimp[mjNIMP];
  for (int i = 0; i < mjNIMP; i++) solimp[i] = solimp_in[i];
  if ((solref[0] > 0) ^ (solref[1] > 0)) {   // mixed format: replaced by the default
    solref[0] = 0.02; solref[1] = 1;         // mj_defaultSolRefImp (engine_io.c)
  }
  if (!(m->opt.disableflags & mjDSBL_REFSAFE) && solref[0] > 0) {
    solref[0] = std::max(solref[0], 2*m->opt.timestep);
  }
  solimp[0] = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[0]));
  solimp[1] = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[1]));
  solimp[2] = std::max(0.0, solimp[2]);
  solimp[3] = std::min(mjMAXIMP, std::max(mjMINIMP, solimp[3]));
  solimp[4] = std::max(1.0, solimp[4]);
  sp[MJB_SP_D0] = solimp[0];
  sp[MJB_SP_D1] = solimp[1];
  sp[MJB_SP_WIDTH] = solimp[2];
  sp[MJB_SP_MID] = solimp[3];
  sp[MJB_SP_POWER] = solimp[4];
  if (solref[0] > 0) {
    sp[MJB_SP_K] = 1 / std::max(mjMINVAL, solimp[1]*solimp[1] * solref[0]*solref[0] * solref[1]*solref[1]);
  } else {
    sp[MJB_SP_K] = -solref[0] / std::max(mjMINVAL, solimp[1]*solimp[1]);
  }
  if (solref[1] > 0) {
    sp[MJB_SP_B] = 2 / std::max(mjMINVAL, solimp[1]*solref[0]);
  } else {
    sp[MJB_SP_B] = -solref[1] / std::max(mjMINVAL, solimp[1]);
  }
}

struct Candidate {
  int g1, g2;      // as passed to mj_collideGeoms (before the type swap)
  int ipair;       // explicit pair index or -1
};

// mj_contactParam (engine_collision_driver.c:1289-1382) for a dynamic geom pair
void contactParam(const mjModel* m, int g1, int g2, int* condim, mjtNum* gap, mjtNum* solref,
                  mjtNum* solimp, mjtNum* friction) {
  mjtNum fri[3];
  const int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
  *gap = std::max(m->geom_gap[g1], m->geom_gap[g2]);
  if (p1 != p2) {
    const int g = p1 > p2 ? g1 : g2;
    *condim = m->geom_condim[g];
    for (int i = 0; i < mjNREF; i++) solref[i] = m->geom_solref[mjNREF*g + i];
    for (int i = 0; i < mjNIMP; i++) solimp[i] = m->geom_solimp[mjNIMP*g + i];
    for (int i = 0; i < 3; i++) fri[i] = m->geom_friction[3*g + i];
  } else {
    *condim = std::max(m->geom_condim[g1], m->geom_condim[g2]);
    const mjtNum s1 = m->geom_solmix[g1], s2 = m->geom_solmix[g2];
    mjtNum mix;
    if (s1 >= mjMINVAL && s2 >= mjMINVAL) mix = s1 / (s1 + s2);
    else if (s1 < mjMINVAL && s2 < mjMINVAL) mix = 0.5;
    else if (s1 < mjMINVAL) mix = 0.0;
    else mix = 1.0;
    const mjtNum* r1 = m->geom_solref + mjNREF*g1;
    const mjtNum* r2 = m->geom_solref + mjNREF*g2;
    if (r1[0] > 0 && r2[0] > 0) {
      for (int i = 0; i < mjNREF; i++) solref[i] = mix*r1[i] + (1 - mix)*r2[i];
    } else {
      for (int i = 0; i < mjNREF; i++) solref[i] = std::min(r1[i], r2[i]);
    }
    for (int i = 0; i < mjNIMP; i++) {
      solimp[i] = mix*m->geom_solimp[mjNIMP*g1 + i] + (1 - mix)*m->geom_solimp[mjNIMP*g2 + i];
    }
    for (int i = 0; i < 3; i++) fri[i] = std::max(m->geom_friction[3*g1 + i], m->geom_friction[3*g2 + i]);
  }
  friction[0] = fri[0]; friction[1] = fri[0]; friction[2] = fri[1]; friction[3] = fri[2];
  friction[4] = fri[2];
}

template <typename... T>
void setError(std::string& err, const char* fmt, T... a) {
  char buf[512];
  std::snprintf(buf, sizeof(buf), fmt, a...);
  err = buf;
}

}  // namespace

bool isSparseJacobian(const mjModel* m) {  // mj_isSparse (engine_core_constraint.c:99-106)
  return m->opt.jacobian == mjJAC_SPARSE || (m->opt.jacobian == mjJAC_AUTO && m->nv >= 60);
}

bool buildModelBlob(const mjModel* m, std::vector<unsigned char>& blob, std::string& err) {
  const int dsbl = m->opt.disableflags, enbl = m->opt.enableflags;
  const bool constraints = !(dsbl & mjDSBL_CONSTRAINT);
  const bool contacts = constraints && !(dsbl & mjDSBL_CONTACT) && m->nconmax != 0 && m->nbody >= 2;

  // ---- validation: everything outside the supported path is refused here, never approximated
  if (m->nflex) { err = "flex objects are not supported (nflex > 0)"; return false; }
  if (m->nplugin) { err = "engine plugins are not supported (nplugin > 0)"; return false; }
  // mj_fluid (engine_passive.c:403-431) runs iff this holds
  const bool fluid = !(dsbl & mjDSBL_PASSIVE) && (m->opt.density > 0 || m->opt.viscosity > 0);
  // gravity compensation (engine_passive.c:381-401) runs iff this holds
  const bool gravcomp = !(dsbl & mjDSBL_PASSIVE) && m->ngravcomp && !(dsbl & mjDSBL_GRAVITY) &&
      (m->opt.gravity[0] != 0 || m->opt.gravity[1] != 0 || m->opt.gravity[2] != 0);
  // mjENBL_INVDISCRETE (mj_discreteAcc, engine_inverse.c:81-164): Euler (1), implicitfast (2) and
  // implicit (3); RK4 is an error in the reference too
  int discrete = 0;
  std::vector<double> act_biasvel(m->nu, 0.0);     // d force / d velocity of every actuator (mjd_actuator_vel)
  bool discrete_trn = false;
  if (enbl & mjENBL_INVDISCRETE) {
    if (m->opt.integrator == mjINT_RK4) {
      err = "mjENBL_INVDISCRETE: discrete inverse dynamics is not supported by RK4 (an error in the reference too)";
      return false;
    }
    if (m->opt.integrator == mjINT_EULER) {
      if (!(dsbl & mjDSBL_EULERDAMP)) {
        for (int i = 0; i < m->nv; i++) discrete = (discrete || m->dof_damping[i] > 0) ? 1 : 0;
      }
    } else {
      // implicitfast: qfrc = (M - h*qDeriv) qacc with qDeriv = mjd_actuator_vel + mjd_passive_vel
      // (engine_derivative.c:812-872, 1432-1505) restricted to M's sparsity; implicit adds mjd_rne_vel
      discrete = m->opt.integrator == mjINT_IMPLICIT ? 3 : 2;
      if (!(dsbl & mjDSBL_ACTUATION)) {
        for (int i = 0; i < m->nu; i++) {
          const int group = m->actuator_group[i];
          if (group >= 0 && group <= 30 && (m->opt.disableactuator & (1 << group))) continue;
          double gain_vel = 0;
          if (m->actuator_gaintype[i] == mjGAIN_AFFINE) gain_vel = m->actuator_gainprm[mjNGAIN*i + 2];
          if (m->actuator_gaintype[i] == mjGAIN_MUSCLE || gain_vel != 0) {
            setError(err, "mjENBL_INVDISCRETE with implicit / implicitfast: actuator %d has a velocity-dependent gain "
                     "(needs ctrl / act, which are not inputs of the batched inverse)", i);
            return false;
          }
          if (m->actuator_biastype[i] == mjBIAS_AFFINE) act_biasvel[i] = m->actuator_biasprm[mjNBIAS*i + 2];
          if (act_biasvel[i] != 0) {
            if (m->actuator_trntype[i] == mjTRN_BODY) {
              setError(err, "mjENBL_INVDISCRETE with implicit / implicitfast: adhesion actuator %d has a velocity bias", i);
              return false;
            }
            discrete_trn = true;
          }
        }
      }
      if (!(dsbl & mjDSBL_PASSIVE)) {
        for (int t = 0; t < m->ntendon; t++) {
          if (m->tendon_damping[t] > 0 && m->wrap_type[m->tendon_adr[t]] != mjWRAP_JOINT) {
            setError(err, "mjENBL_INVDISCRETE with implicit / implicitfast: damped spatial tendon %d (its Jacobian row is not formed)", t);
            return false;
          }
        }
      }
    }
  }
  // Sensors (mj_sensorPos / mj_sensorVel / mj_sensorAcc, engine_sensor.c:222,527,708): the types
  // whose inputs exist on this path are evaluated by the sensor kernel; the others (rangefinder,
  // geom distances through mjc_ccd, actuator forces, user / plugin) are refused unless mjDSBL_SENSOR is set.
  const bool sensors = m->nsensor > 0 && !(dsbl & mjDSBL_SENSOR);
  bool sensor_post = false, sensor_subtreevel = false, sensor_touch = false;
  bool sensor_cam = false, sensor_trn = false, sensor_energy = false, sensor_ray = false;
  bool sensor_ccd = false;
  std::vector<int> sensor_int, sensor_pairs;
  std::vector<double> sensor_cutoff;
  static_assert((int)mjSENS_GEOMDIST == (int)MJB_SENS_GEOMDIST && (int)mjSENS_GEOMFROMTO == (int)MJB_SENS_GEOMFROMTO, "mjtSensor");
  for (int i = 0; i < m->nsensor && sensors; i++) {
    const int t = m->sensor_type[i], ot = m->sensor_objtype[i], rt = m->sensor_reftype[i];
    const int rid = m->sensor_refid[i];
    auto frame_obj = [](int o) { return o == mjOBJ_BODY || o == mjOBJ_XBODY || o == mjOBJ_GEOM || o == mjOBJ_SITE; };
    bool ok = false;
    int pair_first = -1, pair_count = 0;
    switch (t) {
      case mjSENS_GEOMDIST: case mjSENS_GEOMNORMAL: case mjSENS_GEOMFROMTO:
        // mj_geomDistance (engine_support.c:1406-1452) over every geom pair of the two objects: pairs whose
        // entry of the collision table is a primitive function run that function with the cutoff as margin;
        // pairs that go through mjc_ccd there (mjc_Convex, mjc_BoxBox) run GJK / EPA with the cutoff
        ok = (ot == mjOBJ_BODY || ot == mjOBJ_GEOM) && (rt == mjOBJ_BODY || rt == mjOBJ_GEOM);
        if (ok) {
          const int oid = m->sensor_objid[i];
          const int n1 = ot == mjOBJ_BODY ? m->body_geomnum[oid] : 1, id1 = ot == mjOBJ_BODY ? m->body_geomadr[oid] : oid;
          const int n2 = rt == mjOBJ_BODY ? m->body_geomnum[rid] : 1, id2 = rt == mjOBJ_BODY ? m->body_geomadr[rid] : rid;
          pair_first = (int)sensor_pairs.size() / 4;
          for (int ga = id1; ga < id1 + n1 && ok; ga++) {
            for (int gb = id2; gb < id2 + n2 && ok; gb++) {
              const int flip = m->geom_type[ga] > m->geom_type[gb];
              const int g1 = flip ? gb : ga, g2 = flip ? ga : gb;
              const int fn = narrowphaseId(m->geom_type[g1], m->geom_type[g2]);
              if (fn == -1) continue;                 // no collision function: the distance stays at the cutoff
              if (fn < 0) { ok = false; break; }
              if (fn == MJB_FN_BOX_BOX || fn == MJB_FN_CONVEX) {       // measured by mjc_ccd in the reference
                if ((dsbl & mjDSBL_NATIVECCD) || m->opt.ccd_iterations > MJB_CVX_MAXIT) { ok = false; break; }
                sensor_ccd = true;
              }
              const int rec4[4] = {g1, g2, fn, flip};
              sensor_pairs.insert(sensor_pairs.end(), rec4, rec4 + 4);
              pair_count++;
            }
          }
        }
        break;
      case mjSENS_JOINTPOS: case mjSENS_JOINTVEL: case mjSENS_BALLQUAT: case mjSENS_BALLANGVEL:
      case mjSENS_SUBTREECOM:
        ok = true; break;
      case mjSENS_JOINTLIMITPOS: case mjSENS_JOINTLIMITVEL: case mjSENS_JOINTLIMITFRC:
        {
          const int jt = m->jnt_type[m->sensor_objid[i]];
          ok = jt == mjJNT_HINGE || jt == mjJNT_SLIDE;
        }
        break;
      case mjSENS_TENDONPOS: case mjSENS_TENDONVEL:
        // fixed or spatial: the sensor kernel walks the path of a spatial tendon that carries no force itself
        ok = true;
        break;
      case mjSENS_TENDONLIMITPOS: case mjSENS_TENDONLIMITVEL: case mjSENS_TENDONLIMITFRC:
        // fixed tendons only
        ok = true;
        {
          const int tid = m->sensor_objid[i];
          for (int j = 0; j < m->tendon_num[tid]; j++) ok = ok && m->wrap_type[m->tendon_adr[tid] + j] == mjWRAP_JOINT;
        }
        break;
      case mjSENS_VELOCIMETER: case mjSENS_GYRO: case mjSENS_MAGNETOMETER:
        ok = true; break;
      case mjSENS_CLOCK:          // d->time is not a batched input: 0, as after mj_resetData
        ok = true; break;
      case mjSENS_CAMPROJECTION:
        ok = true; sensor_cam = true; break;
      case mjSENS_RANGEFINDER:
        // mj_ray over the primitive geoms; height fields, meshes and SDFs in the ray's way are not restated
        ok = true;
        for (int g = 0; g < m->ngeom; g++) {
          const bool invisible = m->geom_matid[g] < 0 ? m->geom_rgba[4*g + 3] == 0
                                                      : m->mat_rgba[4*m->geom_matid[g] + 3] == 0;
          const int gt = m->geom_type[g];
          if (!invisible && (gt == mjGEOM_HFIELD || gt == mjGEOM_MESH || gt == mjGEOM_SDF)) ok = false;
        }
        sensor_ray = true;
        break;
      case mjSENS_ACTUATORPOS: case mjSENS_ACTUATORVEL:
        ok = true; sensor_trn = true; break;
      case mjSENS_ACTUATORFRC: case mjSENS_JOINTACTFRC:
        // d->actuator_force / d->qfrc_actuator are not computed on the inverse path and are not batched
        // inputs: 0, what a fresh mjData holds (the sensor kernel leaves these readings at zero)
        ok = true; break;
      case mjSENS_E_POTENTIAL: case mjSENS_E_KINETIC:
        ok = true; sensor_energy = true; break;
      case mjSENS_TOUCH:
        {
          const int st = m->site_type[m->sensor_objid[i]];
          ok = st == mjGEOM_SPHERE || st == mjGEOM_CAPSULE || st == mjGEOM_ELLIPSOID || st == mjGEOM_CYLINDER ||
               st == mjGEOM_BOX;
          sensor_touch = true;
        }
        break;
      case mjSENS_SUBTREELINVEL: case mjSENS_SUBTREEANGMOM:
        ok = true; sensor_subtreevel = true; break;
      case mjSENS_ACCELEROMETER:
        ok = true; sensor_post = true; break;
      case mjSENS_FORCE: case mjSENS_TORQUE:
        ok = m->site_bodyid[m->sensor_objid[i]] != 0; sensor_post = true; break;   // cfrc_int of the world body is a mixed-frame sum
      case mjSENS_FRAMELINACC: case mjSENS_FRAMEANGACC:
        ok = frame_obj(ot); sensor_post = true; break;
      case mjSENS_FRAMEPOS: case mjSENS_FRAMEQUAT: case mjSENS_FRAMEXAXIS: case mjSENS_FRAMEYAXIS:
      case mjSENS_FRAMEZAXIS: case mjSENS_FRAMELINVEL: case mjSENS_FRAMEANGVEL:
        ok = frame_obj(ot) && (rid < 0 || frame_obj(rt)); break;
      default: break;
    }
    if (!ok) {
      setError(err, "sensor %d (mjtSensor %d) is not evaluated on the device: set mjDSBL_SENSOR or remove it", i, t);
      return false;
    }
    // geom-distance sensors: the object columns hold the range of their rows in sensor_pairs
    const int rec[MJB_SEN_NI] = {t, m->sensor_datatype[i], ot, pair_first >= 0 ? pair_first : m->sensor_objid[i], rt,
                                 pair_first >= 0 ? pair_count : rid, m->sensor_dim[i], m->sensor_adr[i]};
    sensor_int.insert(sensor_int.end(), rec, rec + MJB_SEN_NI);
    sensor_cutoff.push_back(m->sensor_cutoff[i]);
  }
  const bool equalities = constraints && !(dsbl & mjDSBL_EQUALITY) && m->nemax != 0;
  std::vector<char> tendon_in_equality(m->ntendon, 0);
  for (int i = 0; i < m->neq && equalities; i++) {
    const int t = m->eq_type[i];
    if (t != mjEQ_CONNECT && t != mjEQ_WELD && t != mjEQ_JOINT && t != mjEQ_TENDON) {
      setError(err, "equality constraint %d has an unsupported type (flex / distance)", i); return false;
    }
    if (t == mjEQ_TENDON) {
      tendon_in_equality[m->eq_obj1id[i]] = 1;
      if (m->eq_obj2id[i] >= 0) tendon_in_equality[m->eq_obj2id[i]] = 1;
    }
    if (t == mjEQ_JOINT) {
      for (int j : {m->eq_obj1id[i], m->eq_obj2id[i]}) {
        if (j >= 0 && m->jnt_type[j] != mjJNT_HINGE && m->jnt_type[j] != mjJNT_SLIDE) {
          setError(err, "joint equality %d couples a non-scalar joint", i); return false;
        }
      }
    }
  }
  for (int t = 0; t < m->ntendon; t++) {
    const int adr = m->tendon_adr[t];
    if (m->wrap_type[adr] != mjWRAP_JOINT) {
      // spatial tendon: its path is walked on the device when it carries a force (limit, friction
      // loss, spring, damper, equality constraint)
      for (int j = 0; j < m->tendon_num[t]; j++) {
        const int wt = m->wrap_type[adr + j];
        if (wt != mjWRAP_SITE && wt != mjWRAP_SPHERE && wt != mjWRAP_CYLINDER && wt != mjWRAP_PULLEY) {
          setError(err, "spatial tendon %d has an unknown wrap object type", t); return false;
        }
      }
      continue;
    }
    std::set<int> seen;
    for (int j = 0; j < m->tendon_num[t]; j++) {
      if (!seen.insert(m->wrap_objid[adr + j]).second) {
        setError(err, "fixed tendon %d lists the same joint twice", t); return false;
      }
      const int jt = m->jnt_type[m->wrap_objid[adr + j]];
      if (jt != mjJNT_HINGE && jt != mjJNT_SLIDE) {
        setError(err, "fixed tendon %d uses a non-scalar joint", t); return false;
      }
    }
  }
  if (m->nbody >= 65536) { err = "too many bodies for 16-bit pair signatures"; return false; }

  // ---- candidate geom pairs: mj_collision with the broadphase/midphase replaced by the static
  // superset of body pairs they can return (engine_collision_driver.c:1148-1282, :265-484)
  std::vector<Candidate> cands;
  if (contacts) {
    const bool dsbl_filterparent = dsbl & mjDSBL_FILTERPARENT;
    std::set<unsigned> sigs;
    auto add_pair = [&](int b1, int b2) {
      if (!bodyGeomMasksCompatible(m, b1, b2)) return;
      sigs.insert(b1 < b2 ? ((unsigned)b1 << 16) + b2 : ((unsigned)b2 << 16) + b1);
    };
    for (int b1 = 0; b1 < m->nbody; b1++) {
      if (!canCollide(m, b1)) continue;
      if ((b1 == 0 && m->body_geomnum[b1] > 0) || (m->body_weldid[b1] == 0 && hasPlane(m, b1))) {
        for (int b2 = 0; b2 < m->nbody; b2++) {
          if (!canCollide(m, b2)) continue;
          const int weld2 = m->body_weldid[b2];
          const int parent_weld2 = m->body_weldid[m->body_parentid[weld2]];
          if (filterBodyPair(0, 0, weld2, parent_weld2, dsbl_filterparent)) continue;
          add_pair(b1, b2);
        }
      }
    }
    for (int b1 = 1; b1 < m->nbody; b1++) {
      if (!canCollide(m, b1)) continue;
      for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
        if (!canCollide(m, b2)) continue;
        const int weld1 = m->body_weldid[b1], weld2 = m->body_weldid[b2];
        const int pw1 = m->body_weldid[m->body_parentid[weld1]];
        const int pw2 = m->body_weldid[m->body_parentid[weld2]];
        if (filterBodyPair(weld1, pw1, weld2, pw2, dsbl_filterparent)) continue;
        add_pair(b1, b2);
      }
    }

    int pairadr = 0;
    const int npair = m->npair;
    for (unsigned signature : sigs) {
      const int bf1 = (signature >> 16) & 0xFFFF, bf2 = signature & 0xFFFF;
      bool merged = false;
      const int startadr = pairadr;
      while (pairadr < npair && (unsigned)m->pair_signature[pairadr] <= signature) {
        if ((unsigned)m->pair_signature[pairadr] == signature) merged = true;
        cands.push_back({pairadr, -1, pairadr});
        pairadr++;
      }
      if (filterBitmask(m->body_contype[bf1], m->body_conaffinity[bf1], m->body_contype[bf2],
                        m->body_conaffinity[bf2])) continue;
      bool excluded = false;
      for (int e = 0; e < m->nexclude; e++) {
        if ((unsigned)m->exclude_signature[e] == signature) excluded = true;
      }
      if (excluded) continue;

      auto collideGeomPair = [&](int g1, int g2, std::vector<Candidate>& out) {
        if (merged) {
          for (int k = startadr; k < pairadr; k++) {
            if ((m->pair_geom1[k] == g1 && m->pair_geom2[k] == g2) ||
                (m->pair_geom1[k] == g2 && m->pair_geom2[k] == g1)) return;
          }
        }
        out.push_back({g1, g2, -1});
      };
      const int ga1 = m->body_geomadr[bf1], gn1 = m->body_geomnum[bf1];
      const int ga2 = m->body_geomadr[bf2], gn2 = m->body_geomnum[bf2];
      if (gn1 == 1 && gn2 == 1) {
        collideGeomPair(ga1, ga2, cands);
      } else if (!(dsbl & mjDSBL_MIDPHASE) && m->body_bvhadr[bf1] >= 0 && m->body_bvhadr[bf2] >= 0) {
        // mj_collideTree visits leaf pairs in BVH order, then contacts are stably sorted by their
        // STORED (geom[0], geom[1]), i.e. after the type swap of mj_collideGeoms (:227-257,:367)
        std::vector<Candidate> local;
        for (int g1 = ga1; g1 < ga1 + gn1; g1++) {
          for (int g2 = ga2; g2 < ga2 + gn2; g2++) collideGeomPair(g1, g2, local);
        }
        auto key = [&](const Candidate& c) {
          int a = c.g1, b = c.g2;
          if (m->geom_type[a] > m->geom_type[b]) std::swap(a, b);
          return std::make_pair(a, b);
        };
        std::stable_sort(local.begin(), local.end(),
                         [&](const Candidate& x, const Candidate& y) { return key(x) < key(y); });
        cands.insert(cands.end(), local.begin(), local.end());
      } else {
        for (int g1 = ga1; g1 < ga1 + gn1; g1++) {
          for (int g2 = ga2; g2 < ga2 + gn2; g2++) collideGeomPair(g1, g2, cands);
        }
      }
    }
    while (pairadr < npair) { cands.push_back({pairadr, -1, pairadr}); pairadr++; }
  }

  // ---- per-candidate static part of mj_collideGeoms (engine_collision_driver.c:1440-1632)
  const bool override_ = enbl & mjENBL_OVERRIDE;
  const bool sparse = isSparseJacobian(m);
  std::vector<int> cand_int;
  std::vector<double> cand_num;
  int ncand = 0, max_pair_contacts = 1;
  bool simple_pairs = true, has_convex = false;
  for (const Candidate& cd : cands) {
    int g1 = cd.g1, g2 = cd.g2;
    const int ipair = cd.ipair;
    if (ipair >= 0) { g1 = m->pair_geom1[ipair]; g2 = m->pair_geom2[ipair]; }
    if (m->geom_type[g1] > m->geom_type[g2]) std::swap(g1, g2);
    const int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    const int fn = narrowphaseId(t1, t2);
    if (fn == -1) continue;
    if (ipair < 0 && filterBitmask(m->geom_contype[g1], m->geom_conaffinity[g1],
                                   m->geom_contype[g2], m->geom_conaffinity[g2])) continue;
    if (fn == -2) {
      char buf[256];
      std::snprintf(buf, sizeof(buf),
                    "geom pair (%d:%s, %d:%s) needs a collision function outside the supported "
                    "primitive set; filter it with contype/conaffinity/<exclude> or disable contacts",
                    g1, geomTypeName(t1), g2, geomTypeName(t2));
      err = buf;
      return false;
    }

    if (fn == MJB_FN_CONVEX) {
      // mjc_Convex (engine_collision_convex.c:911-1003) as the reference runs it by default: its own GJK / EPA,
      // one contact per pair
      if (dsbl & mjDSBL_NATIVECCD) { err = "convex geom pairs with mjDSBL_NATIVECCD (libccd) are not supported"; return false; }
      if (enbl & mjENBL_MULTICCD) { err = "convex geom pairs with mjENBL_MULTICCD are not supported"; return false; }
      if (m->opt.ccd_iterations > MJB_CVX_MAXIT) {
        setError(err, "opt.ccd_iterations = %d exceeds %d", m->opt.ccd_iterations, MJB_CVX_MAXIT); return false;
      }
      has_convex = true;
    }

    int condim;
    mjtNum gap, solref[mjNREF], solimp[mjNIMP], friction[5], solreffriction[mjNREF] = {0, 0};
    mjtNum margin;
    if (ipair < 0) {
      margin = std::max(m->geom_margin[g1], m->geom_margin[g2]);
      contactParam(m, g1, g2, &condim, &gap, solref, solimp, friction);
    } else {
      margin = m->pair_margin[ipair];
      condim = m->pair_dim[ipair];
      gap = m->pair_gap[ipair];
      for (int i = 0; i < mjNREF; i++) solref[i] = m->pair_solref[mjNREF*ipair + i];
      for (int i = 0; i < mjNIMP; i++) solimp[i] = m->pair_solimp[mjNIMP*ipair + i];
      for (int i = 0; i < 5; i++) friction[i] = m->pair_friction[5*ipair + i];
      if (m->pair_solreffriction[mjNREF*ipair] || m->pair_solreffriction[mjNREF*ipair + 1]) {
        solreffriction[0] = m->pair_solreffriction[mjNREF*ipair];
        solreffriction[1] = m->pair_solreffriction[mjNREF*ipair + 1];
      }
    }
    if (condim < 1 || condim > 6) {
      setError(err, "invalid condim %d", condim); return false;
    }
    if (override_) {   // mj_assignMargin / Ref / Imp / Friction (engine_core_constraint.c:122-165)
      margin = m->opt.o_margin;
      for (int i = 0; i < mjNREF; i++) { solref[i] = m->opt.o_solref[i]; solreffriction[i] = m->opt.o_solref[i]; }
      for (int i = 0; i < mjNIMP; i++) solimp[i] = m->opt.o_solimp[i];
      for (int i = 0; i < 5; i++) friction[i] = m->opt.o_friction[i];
    }
    for (int i = 0; i < 5; i++) friction[i] = std::max(mjMINMU, friction[i]);

    const int b1 = m->geom_bodyid[g1], b2 = m->geom_bodyid[g2];
    int ci[MJB_CAND_NI];
    double cn[MJB_CAND_NN];
    ci[MJB_CI_G1] = g1; ci[MJB_CI_G2] = g2; ci[MJB_CI_FUNC] = fn; ci[MJB_CI_DIM] = condim;
    {
      const int per = fn == MJB_FN_BOX_BOX ? 24   /* MJB_MAXCON_PAIR */
                      : (fn == MJB_FN_PLANE_CYLINDER || fn == MJB_FN_PLANE_BOX) ? 4
                      : (fn == MJB_FN_PLANE_CAPSULE || fn == MJB_FN_CAPSULE_CAPSULE ||
                         fn == MJB_FN_CAPSULE_BOX) ? 2 : 1;
      max_pair_contacts = std::max(max_pair_contacts, per);
      simple_pairs = simple_pairs && (fn == MJB_FN_PLANE_SPHERE || fn == MJB_FN_PLANE_CAPSULE || fn == MJB_FN_SPHERE_SPHERE ||
                                      fn == MJB_FN_SPHERE_CAPSULE || fn == MJB_FN_CAPSULE_CAPSULE);
    }
    ci[MJB_CI_B1] = b1; ci[MJB_CI_B2] = b2;
    // NV == 0 only arises for the sparse Jacobian's merged chain (engine_support.c:659-690)
    const bool static1 = m->body_weldid[b1] == 0, static2 = m->body_weldid[b2] == 0;
    ci[MJB_CI_FLAGS] = (sparse && static1 && static2) ? 1 : 0;
    // mj_filterSphere variant (:146-163)
    const mjtNum rb1 = m->geom_rbound[g1], rb2 = m->geom_rbound[g2];
    if (rb1 > 0 && rb2 > 0) {
      ci[MJB_CI_PLANE] = 0;
      cn[MJB_CN_RBOUND] = rb1 + rb2 + margin;
    } else if (t1 == mjGEOM_PLANE && rb2 > 0) {
      ci[MJB_CI_PLANE] = 1;
      cn[MJB_CN_RBOUND] = margin + rb2;
    } else {
      ci[MJB_CI_PLANE] = 2;   // no bounding-sphere test
      cn[MJB_CN_RBOUND] = 0;
    }
    cn[MJB_CN_MARGIN] = margin;
    cn[MJB_CN_INCLUDEMARGIN] = margin - gap;
    for (int i = 0; i < 5; i++) cn[MJB_CN_FRICTION + i] = friction[i];
    makeSolParam(m, solref, solimp, cn + MJB_CN_SP);
    {
      // elliptic friction rows use solreffriction when non-zero (:1517-1545); K = 0 there
      mjtNum ref[mjNREF] = {solref[0], solref[1]};
      mjtNum srf[mjNREF] = {solreffriction[0], solreffriction[1]};
      if ((srf[0] > 0) ^ (srf[1] > 0)) { srf[0] = 0; srf[1] = 0; }
      if (srf[0] || srf[1]) { ref[0] = srf[0]; ref[1] = srf[1]; }
      double sp2[MJB_SP_N];
      makeSolParam(m, ref, solimp, sp2);
      cn[MJB_CN_BFRIC] = sp2[MJB_SP_B];
    }
    cn[MJB_CN_DA_TRAN] = m->body_invweight0[2*b1] + m->body_invweight0[2*b2];
    cn[MJB_CN_DA_ROT] = m->body_invweight0[2*b1 + 1] + m->body_invweight0[2*b2 + 1];
    for (int i = 0; i < mjNREF; i++) cn[MJB_CN_SOLREF + i] = solref[i];
    for (int i = 0; i < mjNIMP; i++) cn[MJB_CN_SOLIMP + i] = solimp[i];
    cand_int.insert(cand_int.end(), ci, ci + MJB_CAND_NI);
    cand_num.insert(cand_num.end(), cn, cn + MJB_CAND_NN);
    ncand++;
  }

  if (ncand >= (1 << 27)) { err = "too many candidate geom pairs (>= 2^27)"; return false; }

  // ---- compact tables of the bounding-sphere scan, and the tree-level broadphase
  // The scan (mj_filterSphere per candidate) reads 16 bytes per candidate -- geom ids, filter kind,
  // bound -- instead of the 256-byte candidate record, so that even the ~90 K candidates of the
  // 22-humanoid scene stay resident in L2. For scenes with several kinematic trees the candidate
  // list is cut into runs of equal (tree of body 1, tree of body 2); the scan skips a run when the
  // bounding spheres of the two trees do not overlap (what mj_broadphase's sweep achieves with body
  // AABBs, engine_collision_driver.c:1148-1282; conservative, so the survivors are unchanged).
  std::vector<int> scan_int((size_t)ncand * 2), scan_run, tree_int;
  std::vector<double> scan_bound((size_t)ncand), scan_misc(1, 0.0);
  int nrun = 0, ntree = 0;
  {
    if (m->ngeom >= (1 << 28)) { err = "too many geoms"; return false; }
    for (int i = 0; i < ncand; i++) {
      const int* ci = cand_int.data() + (size_t)i * MJB_CAND_NI;
      scan_int[2*i] = ci[MJB_CI_G1] | (ci[MJB_CI_PLANE] << 28);
      scan_int[2*i + 1] = ci[MJB_CI_G2];
      scan_bound[i] = cand_num[(size_t)i * MJB_CAND_NN + MJB_CN_RBOUND];
      scan_misc[0] = std::max(scan_misc[0], cand_num[(size_t)i * MJB_CAND_NN + MJB_CN_MARGIN]);
    }
    // trees: maximal sets of moving bodies with the same root; static bodies (weldid 0) have none
    std::vector<int> tree_of(m->nbody, -1);
    for (int b = 1; b < m->nbody; b++) {
      if (m->body_weldid[b] == 0) continue;
      const int r = m->body_rootid[b];
      if (tree_of[r] < 0) { tree_of[r] = ntree++; tree_int.push_back(r); tree_int.push_back(m->ngeom); tree_int.push_back(0); }
      tree_of[b] = tree_of[r];
    }
    for (int g = 0; g < m->ngeom; g++) {
      const int t = tree_of[m->geom_bodyid[g]];
      if (t < 0) continue;
      tree_int[3*t + 1] = std::min(tree_int[3*t + 1], g);
      tree_int[3*t + 2] = std::max(tree_int[3*t + 2], g + 1);
    }
    for (int t = 0; t < ntree; t++) if (tree_int[3*t + 1] > tree_int[3*t + 2]) tree_int[3*t + 1] = tree_int[3*t + 2] = 0;
    // runs are worth their bookkeeping only for multi-tree scenes with long candidate lists
    if (ntree >= 2 && ncand >= 1024) {
      int i = 0;
      while (i < ncand) {
        const int* ci = cand_int.data() + (size_t)i * MJB_CAND_NI;
        const int t1 = tree_of[ci[MJB_CI_B1]], t2 = tree_of[ci[MJB_CI_B2]];
        int j = i + 1;
        while (j < ncand) {
          const int* cj = cand_int.data() + (size_t)j * MJB_CAND_NI;
          if (tree_of[cj[MJB_CI_B1]] != t1 || tree_of[cj[MJB_CI_B2]] != t2) break;
          j++;
        }
        scan_run.push_back(i); scan_run.push_back(j - i); scan_run.push_back(t1); scan_run.push_back(t2);
        nrun++;
        i = j;
      }
    } else {
      ntree = 0;
      tree_int.clear();
    }
  }

  // ---- sparse structure of qLD: row i = ancestors of dof i ascending, then i
  // (makeDofDofSparse with reduced/upper flags as used for C, engine_io.c:929-1018; mapM2C :1135)
  const int nv = m->nv;
  std::vector<int> C_rownnz(nv), C_rowadr(nv), C_colind, mapM2C;
  for (int i = 0; i < nv; i++) {
    std::vector<int> chain;
    // reduced layout: a simple dof keeps only its diagonal (engine_io.c:952-962)
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      chain.push_back(j);
      if (m->dof_simplenum[i]) break;
    }
    C_rownnz[i] = (int)chain.size();
    C_rowadr[i] = (int)C_colind.size();
    for (int s = 0; s < (int)chain.size(); s++) {
      C_colind.push_back(chain[chain.size() - 1 - s]);
      mapM2C.push_back(m->dof_Madr[i] + (int)chain.size() - 1 - s);
    }
  }
  if ((int)C_colind.size() != m->nC) {
    err = "unexpected sparse inertia structure (nC mismatch)"; return false;
  }

  std::vector<int> body_static(m->nbody), jnt_dofnum(m->njnt);
  for (int b = 0; b < m->nbody; b++) body_static[b] = (m->body_weldid[b] == 0);
  for (int j = 0; j < m->njnt; j++) {
    const int t = m->jnt_type[j];
    jnt_dofnum[j] = t == mjJNT_FREE ? 6 : (t == mjJNT_BALL ? 3 : 1);
  }

  // solver parameter tables
  std::vector<double> sp_jnt(MJB_SP_N*m->njnt), sp_tl(MJB_SP_N*m->ntendon),
      sp_df(MJB_SP_N*nv), sp_tf(MJB_SP_N*m->ntendon);
  for (int j = 0; j < m->njnt; j++) {
    makeSolParam(m, m->jnt_solref + mjNREF*j, m->jnt_solimp + mjNIMP*j, sp_jnt.data() + MJB_SP_N*j);
  }
  for (int i = 0; i < nv; i++) {
    makeSolParam(m, m->dof_solref + mjNREF*i, m->dof_solimp + mjNIMP*i, sp_df.data() + MJB_SP_N*i);
  }
  for (int t = 0; t < m->ntendon; t++) {
    makeSolParam(m, m->tendon_solref_lim + mjNREF*t, m->tendon_solimp_lim + mjNIMP*t,
                 sp_tl.data() + MJB_SP_N*t);
    makeSolParam(m, m->tendon_solref_fri + mjNREF*t, m->tendon_solimp_fri + mjNIMP*t,
                 sp_tf.data() + MJB_SP_N*t);
  }

  // ---- equality constraints (mj_instantiateEquality, engine_core_constraint.c:493-763)
  const int neq = equalities ? m->neq : 0;
  std::vector<int> eq_int((size_t)neq * MJB_EQ_NI, 0);
  std::vector<double> eq_num((size_t)neq * MJB_EQ_NN, 0.0), sp_eq((size_t)neq * MJB_SP_N, 0.0);
  for (int i = 0; i < neq; i++) {
    int* ei = eq_int.data() + (size_t)i * MJB_EQ_NI;
    double* en = eq_num.data() + (size_t)i * MJB_EQ_NN;
    const mjtNum* data = m->eq_data + mjNEQDATA * i;
    const int type = m->eq_type[i], id0 = m->eq_obj1id[i], id1 = m->eq_obj2id[i];
    ei[MJB_EQI_TYPE] = type;
    ei[MJB_EQI_ACTIVE] = m->eq_active0[i];
    makeSolParam(m, m->eq_solref + mjNREF*i, m->eq_solimp + mjNIMP*i, sp_eq.data() + (size_t)i * MJB_SP_N);
    en[MJB_EQN_Q0] = 1; en[MJB_EQN_Q1] = 1;
    if (type == mjEQ_CONNECT || type == mjEQ_WELD) {
      const bool site = m->eq_objtype[i] == mjOBJ_SITE;
      const int b0 = site ? m->site_bodyid[id0] : id0, b1 = site ? m->site_bodyid[id1] : id1;
      ei[MJB_EQI_B0] = b0; ei[MJB_EQI_B1] = b1; ei[MJB_EQI_SITE] = site;
      ei[MJB_EQI_SKIP] = (m->body_weldid[b0] == 0 && m->body_weldid[b1] == 0);
      for (int k = 0; k < 3; k++) {
        if (site) {
          en[MJB_EQN_ANCHOR0 + k] = m->site_pos[3*id0 + k];
          en[MJB_EQN_ANCHOR1 + k] = m->site_pos[3*id1 + k];
        } else if (type == mjEQ_CONNECT) {
          en[MJB_EQN_ANCHOR0 + k] = data[k];
          en[MJB_EQN_ANCHOR1 + k] = data[3 + k];
        } else {                       // weld: anchor of body j is data + 3*(1-j) (:565)
          en[MJB_EQN_ANCHOR0 + k] = data[3 + k];
          en[MJB_EQN_ANCHOR1 + k] = data[k];
        }
      }
      if (type == mjEQ_WELD) {
        for (int k = 0; k < 4; k++) {
          en[MJB_EQN_Q0 + k] = site ? m->site_quat[4*id0 + k] : data[6 + k];
          en[MJB_EQN_Q1 + k] = site ? m->site_quat[4*id1 + k] : (k == 0 ? 1.0 : 0.0);
        }
        en[MJB_EQN_TORQUESCALE] = data[10];
      }
      en[MJB_EQN_DA_TRAN] = m->body_invweight0[2*b0] + m->body_invweight0[2*b1];
      en[MJB_EQN_DA_ROT] = m->body_invweight0[2*b0 + 1] + m->body_invweight0[2*b1 + 1];
    } else {
      ei[MJB_EQI_B0] = id0; ei[MJB_EQI_B1] = id1;
      if (type == mjEQ_JOINT) {
        en[MJB_EQN_DA_TRAN] = m->dof_invweight0[m->jnt_dofadr[id0]] +
                              (id1 >= 0 ? m->dof_invweight0[m->jnt_dofadr[id1]] : 0.0);
      } else {
        en[MJB_EQN_DA_TRAN] = m->tendon_invweight0[id0] + (id1 >= 0 ? m->tendon_invweight0[id1] : 0.0);
      }
    }
  }

  // tendons that carry a force; spatial ones with springs/dampers need the passive wrench carrier
  // A tendon whose path touches no dof has an empty Jacobian row: mj_addConstraint drops its friction
  // and limit rows (engine_core_constraint.c:265-330 -- dense: every entry of the row is zero,
  // sparse: the dof chain is empty; reference models core_constraint/dofless_tendon_*.xml). That is
  // a property of the model: spatial tendons between sites / geoms of static bodies, and (dense
  // only) fixed tendons whose coefficients are all zero.
  std::vector<int> tendon_empty(m->ntendon, 0);
  for (int t = 0; t < m->ntendon; t++) {
    bool empty = true;
    for (int j = 0; j < m->tendon_num[t] && empty; j++) {
      const int w = m->tendon_adr[t] + j, wt = m->wrap_type[w], id = m->wrap_objid[w];
      if (wt == mjWRAP_JOINT) empty = !isSparseJacobian(m) && m->wrap_prm[w] == 0;
      else if (wt == mjWRAP_SITE) empty = m->body_weldid[m->site_bodyid[id]] == 0;
      else if (wt == mjWRAP_SPHERE || wt == mjWRAP_CYLINDER) empty = m->body_weldid[m->geom_bodyid[id]] == 0;
    }
    tendon_empty[t] = empty;
  }
  std::vector<int> tendon_active(m->ntendon, 0);
  bool spatial_passive = false, spatial_active = false;
  for (int t = 0; t < m->ntendon; t++) {
    const bool lim = m->tendon_limited[t] && !(dsbl & mjDSBL_LIMIT) && constraints && !tendon_empty[t];
    const bool fric = m->tendon_frictionloss[t] > 0 && !(dsbl & mjDSBL_FRICTIONLOSS) && constraints && !tendon_empty[t];
    const bool pas = (m->tendon_stiffness[t] != 0 || m->tendon_damping[t] != 0) && !(dsbl & mjDSBL_PASSIVE);
    const bool spatial = m->wrap_type[m->tendon_adr[t]] != mjWRAP_JOINT;
    if (spatial && tendon_in_equality[t] && tendon_empty[t]) {
      // the reference drops such a row in sparse mode and keeps a zero row in dense mode (mj_addConstraint)
      setError(err, "spatial tendon %d between static bodies is used by an equality constraint (not supported)", t);
      return false;
    }
    tendon_active[t] = lim || fric || pas || (spatial && tendon_in_equality[t]);
    if (pas && m->wrap_type[m->tendon_adr[t]] != mjWRAP_JOINT) spatial_passive = true;
    if (tendon_active[t] && m->wrap_type[m->tendon_adr[t]] != mjWRAP_JOINT) spatial_active = true;
  }
  if ((spatial_passive || fluid) && gravcomp) {
    for (int j = 0; j < m->njnt; j++) {
      if (m->jnt_actgravcomp[j]) {
        err = "spatial-tendon springs/dampers or fluid forces together with actuator-routed gravity "
              "compensation (jnt_actgravcomp) are not supported";
        return false;
      }
    }
  }
  // Fluid forces (mj_fluid, engine_passive.c:403-431): per body either the inertia-box model or, if one
  // of its geoms has geom_fluid[0] > 0, the ellipsoid model on its geoms. Everything that does not
  // depend on the state is formed here: the box sides (:530-536) and the semi-axes (mju_geomSemiAxes).
  std::vector<double> fluid_body, fluid_geom;
  bool fluid_ellipsoid = false;
  if (fluid) {
    if (discrete >= 2) {
      err = "mjENBL_INVDISCRETE with implicit / implicitfast: fluid forces (their velocity derivatives are not formed)";
      return false;
    }
    fluid_body.assign((size_t)4 * m->nbody, 0.0);
    for (int i = 1; i < m->nbody; i++) {
      if (m->body_mass[i] < mjMINVAL) continue;
      bool ell = false;
      for (int j = 0; j < m->body_geomnum[i] && !ell; j++) ell = m->geom_fluid[mjNFLUID*(m->body_geomadr[i] + j)] > 0;
      fluid_ellipsoid = fluid_ellipsoid || ell;
      fluid_body[4*i] = ell ? MJB_FLUID_ELLIPSOID : MJB_FLUID_BOX;
      const mjtNum* inertia = m->body_inertia + 3*i;
      fluid_body[4*i + 1] = std::sqrt(std::max((mjtNum)mjMINVAL, (inertia[1] + inertia[2] - inertia[0])) / m->body_mass[i] * 6.0);
      fluid_body[4*i + 2] = std::sqrt(std::max((mjtNum)mjMINVAL, (inertia[0] + inertia[2] - inertia[1])) / m->body_mass[i] * 6.0);
      fluid_body[4*i + 3] = std::sqrt(std::max((mjtNum)mjMINVAL, (inertia[0] + inertia[1] - inertia[2])) / m->body_mass[i] * 6.0);
    }
    if (fluid_ellipsoid) {
      fluid_geom.assign((size_t)MJB_FLUID_NG * m->ngeom, 0.0);
      for (int g = 0; g < m->ngeom; g++) {
        double* r = fluid_geom.data() + (size_t)MJB_FLUID_NG * g;
        for (int k = 0; k < mjNFLUID; k++) r[k] = m->geom_fluid[mjNFLUID*g + k];
        const mjtNum* size = m->geom_size + 3*g;
        switch (m->geom_type[g]) {
          case mjGEOM_SPHERE: r[12] = size[0]; r[13] = size[0]; r[14] = size[0]; break;
          case mjGEOM_CAPSULE: r[12] = size[0]; r[13] = size[0]; r[14] = size[1] + size[0]; break;
          case mjGEOM_CYLINDER: r[12] = size[0]; r[13] = size[0]; r[14] = size[1]; break;
          default: r[12] = size[0]; r[13] = size[1]; r[14] = size[2];
        }
      }
    }
  }

  // ---- assemble
  mjbHdr H;
  std::memset(&H, 0, sizeof(H));
  H.magic = MJB_MAGIC;
  H.nq = m->nq; H.nv = nv; H.nbody = m->nbody; H.njnt = m->njnt; H.ngeom = m->ngeom;
  H.ntendon = m->ntendon; H.nwrap = m->nwrap; H.neq = neq; H.nM = m->nM; H.nC = m->nC;
  H.ncand = ncand;
  H.nrun = nrun; H.ntree = ntree;
  H.max_pair_contacts = max_pair_contacts;
  H.simple_pairs = simple_pairs ? 1 : 0;
  H.has_convex = has_convex ? 1 : 0;
  H.sensor_ccd = sensor_ccd ? 1 : 0;
  H.ccd_iterations = m->opt.ccd_iterations;
  H.ccd_tolerance = m->opt.ccd_tolerance;
  H.disableflags = dsbl; H.enableflags = enbl; H.cone = m->opt.cone;
  H.has_gravcomp = gravcomp ? 1 : 0;
  H.passive_wrench = (gravcomp || spatial_passive || fluid) ? 1 : 0;
  H.has_fluid = fluid ? (fluid_ellipsoid ? 2 : 1) : 0;
  H.density = m->opt.density; H.viscosity = m->opt.viscosity;
  for (int k = 0; k < 3; k++) H.wind[k] = m->opt.wind[k];
  H.has_spatial = spatial_active ? 1 : 0;
  H.discrete_acc = discrete;
  H.discrete_trn = discrete_trn ? 1 : 0;
  H.nsensor = sensors ? m->nsensor : 0;
  H.nsensordata = sensors ? m->nsensordata : 0;
  H.sensor_post = sensor_post ? 1 : 0;
  H.nsite = m->nsite;
  H.nmocap = m->nmocap;
  H.ncam = m->ncam; H.nlight = m->nlight;
  H.nu = m->nu;
  H.sensor_subtreevel = sensor_subtreevel ? 1 : 0;
  H.sensor_touch = sensor_touch ? 1 : 0;
  H.sensor_cam = (sensor_cam || sensor_ray) ? 1 : 0;     // both read tables behind the staged part of the blob
  H.sensor_camlight = sensor_cam ? 1 : 0; H.sensor_trn = sensor_trn ? 1 : 0; H.sensor_energy = sensor_energy ? 1 : 0;
  for (int i = 0; i < 3; i++) H.magnetic[i] = m->opt.magnetic[i];
  H.timestep = m->opt.timestep; H.impratio = m->opt.impratio;
  for (int i = 0; i < 3; i++) H.gravity[i] = m->opt.gravity[i];

  std::vector<int> ints;
  std::vector<double> nums;
  // Every table starts on a 16-byte boundary; the tables whose records the contact kernels gather
  // per candidate / per item (32-byte integer records, scan rows) start on a 32-byte sector, so that
  // the sector alignment of a record does not depend on which other tables a model happens to have
  // (both sections start on 128-byte boundaries of the blob, see below).
  auto hotInt = [](int id) {
    return id == MJB_I_cand_int || id == MJB_I_scan_int || id == MJB_I_scan_run || id == MJB_I_eq_int ||
           id == MJB_I_tree_int || id == MJB_I_sensor_int;
  };
  auto hotNum = [](int id) { return id == MJB_N_cand_num || id == MJB_N_scan_bound || id == MJB_N_eq_num; };
  // Tables read only by the output-only kernels (mj_camlight, mj_transmission, the implicit
  // mj_discreteAcc, cam_project) live behind the part of the blob a CTA stages into shared memory:
  // the phase kernels' shared-memory footprint, and with it their occupancy, does not pay for them.
  // Their offsets are, like every other, relative to the section starts (cold_int / cold_num below).
  auto coldInt = [](int id) {
    return id == MJB_I_cam_mode || id == MJB_I_cam_bodyid || id == MJB_I_cam_targetbodyid ||
           id == MJB_I_light_mode || id == MJB_I_light_bodyid || id == MJB_I_light_targetbodyid ||
           id == MJB_I_actuator_trntype || id == MJB_I_actuator_trn || id == MJB_I_ray_geom;
  };
  auto coldNum = [](int id) {
    return id == MJB_N_cam_pos || id == MJB_N_cam_quat || id == MJB_N_cam_poscom0 || id == MJB_N_cam_pos0 ||
           id == MJB_N_cam_mat0 || id == MJB_N_light_pos || id == MJB_N_light_dir || id == MJB_N_light_poscom0 ||
           id == MJB_N_light_pos0 || id == MJB_N_light_dir0 || id == MJB_N_actuator_gear ||
           id == MJB_N_actuator_cranklength || id == MJB_N_act_biasvel || id == MJB_N_cam_proj;
  };
  std::vector<int> cold_ints;
  std::vector<double> cold_nums;
  std::vector<int> cold_int_ids, cold_num_ids;
  auto pushInts = [&](int id, const int* src, size_t n) {
    if (coldInt(id)) {
      cold_ints.resize((cold_ints.size() + 3) / 4 * 4, 0);
      H.ioff[id] = (int)cold_ints.size();          // rebased once the hot part is complete
      cold_ints.insert(cold_ints.end(), src, src + n);
      cold_int_ids.push_back(id);
      return;
    }
    const size_t al = hotInt(id) ? 8 : 4;
    ints.resize((ints.size() + al - 1) / al * al, 0);
    H.ioff[id] = (int)ints.size();
    ints.insert(ints.end(), src, src + n);
  };
  auto pushNums = [&](int id, const double* src, size_t n) {
    if (coldNum(id)) {
      cold_nums.resize((cold_nums.size() + 1) / 2 * 2, 0.0);
      H.noff[id] = (int)cold_nums.size();
      cold_nums.insert(cold_nums.end(), src, src + n);
      cold_num_ids.push_back(id);
      return;
    }
    const size_t al = hotNum(id) ? 4 : 2;
    nums.resize((nums.size() + al - 1) / al * al, 0.0);
    H.noff[id] = (int)nums.size();
    nums.insert(nums.end(), src, src + n);
  };
#define X(name, rows) pushInts(MJB_I_##name, m->name, (size_t)m->rows);
  MJB_INT_ARRAYS(X)
#undef X
#define X(name, rows) { std::vector<int> w(m->name, m->name + m->rows); \
                        pushInts(MJB_I_##name, w.data(), w.size()); }
  MJB_BYTE_ARRAYS(X)
#undef X
  pushInts(MJB_I_C_rownnz, C_rownnz.data(), C_rownnz.size());
  pushInts(MJB_I_C_rowadr, C_rowadr.data(), C_rowadr.size());
  pushInts(MJB_I_C_colind, C_colind.data(), C_colind.size());
  pushInts(MJB_I_mapM2C, mapM2C.data(), mapM2C.size());
  pushInts(MJB_I_cand_int, cand_int.data(), cand_int.size());
  pushInts(MJB_I_scan_int, scan_int.data(), scan_int.size());
  pushInts(MJB_I_scan_run, scan_run.data(), scan_run.size());
  pushInts(MJB_I_tree_int, tree_int.data(), tree_int.size());
  pushInts(MJB_I_eq_int, eq_int.data(), eq_int.size());
  pushInts(MJB_I_sensor_int, sensor_int.data(), sensor_int.size());
  pushInts(MJB_I_body_static, body_static.data(), body_static.size());
  pushInts(MJB_I_tendon_active, tendon_active.data(), tendon_active.size());
  pushInts(MJB_I_sensor_pairs, sensor_pairs.data(), sensor_pairs.size());
  pushInts(MJB_I_actuator_trn, m->actuator_trnid, (size_t)2 * m->nu);
  {
    // rangefinder sensors: the state-independent part of ray_eliminate (engine_ray.c:69-100; flg_static = 1,
    // no geom groups: only invisible geoms are skipped, the site's own body is tested per sensor)
    std::vector<int> ray_geom(sensor_ray ? m->ngeom : 0, 1);
    for (size_t g = 0; g < ray_geom.size(); g++) {
      ray_geom[g] = m->geom_matid[g] < 0 ? m->geom_rgba[4*g + 3] != 0 : m->mat_rgba[4*m->geom_matid[g] + 3] != 0;
    }
    pushInts(MJB_I_ray_geom, ray_geom.data(), ray_geom.size());
  }
  pushInts(MJB_I_jnt_dofnum_tab, jnt_dofnum.data(), jnt_dofnum.size());
  {
    // static row numbering (mj_makeConstraint order: equality, dof friction, tendon friction, ...)
    const bool rows = !(dsbl & mjDSBL_CONSTRAINT);
    std::vector<int> frow(nv, -1);
    int nfd = 0, nft = 0, ne_rows = 0;
    if (rows && !(dsbl & mjDSBL_FRICTIONLOSS)) {
      for (int i = 0; i < nv; i++) if (m->dof_frictionloss[i] > 0) frow[i] = nfd++;
      for (int t = 0; t < m->ntendon; t++) if (m->tendon_frictionloss[t] > 0 && !tendon_empty[t]) nft++;
    }
    if (rows) {
      for (int i = 0; i < neq; i++) {
        const int* ei = eq_int.data() + (size_t)i * MJB_EQ_NI;
        if (!ei[MJB_EQI_ACTIVE]) continue;
        if (ei[MJB_EQI_TYPE] == mjEQ_CONNECT || ei[MJB_EQI_TYPE] == mjEQ_WELD) {
          if (!ei[MJB_EQI_SKIP]) ne_rows += ei[MJB_EQI_TYPE] == mjEQ_CONNECT ? 3 : 6;
        } else {
          ne_rows++;
        }
      }
    }
    H.ne_rows = ne_rows; H.nf_dof_rows = nfd; H.nf_rows = nfd + nft;
    pushInts(MJB_I_dof_frow, frow.data(), frow.size());
  }
  {
    std::vector<int> flags(m->nbody, 0), seen(m->nbody, 0);
    for (int b = m->nbody - 1; b > 0; b--) {
      const int p = m->body_parentid[b];
      flags[p] |= 1;
      if (!seen[p]) { flags[b] |= 2; seen[p] = 1; }
      if (b != p + 1) flags[p] |= 4;
    }
    // bodies whose pose is read back from scratch after the sweep
    for (int i = 0; i < neq; i++) {
      const int* ei = eq_int.data() + (size_t)i * MJB_EQ_NI;
      if (ei[MJB_EQI_TYPE] == mjEQ_CONNECT || ei[MJB_EQI_TYPE] == mjEQ_WELD) {
        flags[ei[MJB_EQI_B0]] |= 8; flags[ei[MJB_EQI_B1]] |= 8;
      }
    }
    for (int t = 0; t < m->ntendon; t++) {
      if (!tendon_active[t]) continue;
      for (int j = 0; j < m->tendon_num[t]; j++) {
        const int wt = m->wrap_type[m->tendon_adr[t] + j], id = m->wrap_objid[m->tendon_adr[t] + j];
        if (wt == mjWRAP_SITE) flags[m->site_bodyid[id]] |= 8;
        if (wt == mjWRAP_SPHERE || wt == mjWRAP_CYLINDER) {
          const int side = (int)std::lround(m->wrap_prm[m->tendon_adr[t] + j]);
          if (side >= 0 && side < m->nsite) flags[m->site_bodyid[side]] |= 8;
        }
      }
    }
    // bodies whose velocity / acceleration carriers (cvel, cacc_lin) a later phase reads: the
    // bodies flagged above (equality rows, tendon sites) read the carrier ROWS (bit 4), the two
    // bodies of every candidate pair (contact rows) read the carrier RECORD (bit 5, MJB_SC_crec)
    for (int i = 0; i < ncand; i++) {
      const int* ci = cand_int.data() + (size_t)i * MJB_CAND_NI;
      flags[ci[MJB_CI_B1]] |= 32; flags[ci[MJB_CI_B2]] |= 32;
    }
    for (int b = 0; b < m->nbody; b++) if (flags[b] & 8) flags[b] |= 16;
    pushInts(MJB_I_body_tree_flags, flags.data(), flags.size());
    // geom frames a later phase reads (see body_geoms)
    std::vector<int> geom_store(m->ngeom, 0);
    for (int i = 0; i < ncand; i++) {
      const int* ci = cand_int.data() + (size_t)i * MJB_CAND_NI;
      const int fn = ci[MJB_CI_FUNC];
      const bool light = fn == MJB_FN_PLANE_SPHERE || fn == MJB_FN_PLANE_CAPSULE || fn == MJB_FN_SPHERE_SPHERE ||
                         fn == MJB_FN_SPHERE_CAPSULE || fn == MJB_FN_CAPSULE_CAPSULE;
      geom_store[ci[MJB_CI_G1]] |= light ? 1 : 3;
      geom_store[ci[MJB_CI_G2]] |= light ? 1 : 3;
    }
    // tendons that drive an actuator: mj_transmission walks their path when mjbOUT_TRANSMISSION is requested
    std::vector<int> tendon_trn(m->ntendon, 0);
    for (int i = 0; i < m->nu; i++) {
      if (m->actuator_trntype[i] == mjTRN_TENDON) tendon_trn[m->actuator_trnid[2*i]] = 1;
    }
    // ... or are read by a tendonpos / tendonvel sensor
    for (int i = 0; i < m->nsensor && sensors; i++) {
      if (m->sensor_type[i] == mjSENS_TENDONPOS || m->sensor_type[i] == mjSENS_TENDONVEL) tendon_trn[m->sensor_objid[i]] = 1;
    }
    for (int t = 0; t < m->ntendon; t++) {
      if (!tendon_active[t] && !tendon_trn[t]) continue;
      for (int j = 0; j < m->tendon_num[t]; j++) {
        const int wt = m->wrap_type[m->tendon_adr[t] + j];
        // bit 2: only the output-only stages (mj_transmission, tendon sensors) walk this tendon -- the frame is
        // kept only in runs that produce those outputs
        if (wt == mjWRAP_SPHERE || wt == mjWRAP_CYLINDER) geom_store[m->wrap_objid[m->tendon_adr[t] + j]] |= tendon_active[t] ? 3 : 4;
      }
    }
    pushInts(MJB_I_geom_store, geom_store.data(), geom_store.size());
  }
#define X(name, rows, cols) pushNums(MJB_N_##name, m->name, (size_t)m->rows * (cols));
  MJB_NUM_ARRAYS(X)
#undef X
  for (int t = 0; t < m->ntendon; t++) {          // rows of empty tendons are never instantiated
    if (!tendon_empty[t]) continue;
    nums[H.noff[MJB_N_tendon_frictionloss] + t] = 0;
    ints[H.ioff[MJB_I_tendon_limited] + t] = 0;
  }
  pushNums(MJB_N_sp_jnt_limit, sp_jnt.data(), sp_jnt.size());
  pushNums(MJB_N_sp_tendon_limit, sp_tl.data(), sp_tl.size());
  pushNums(MJB_N_sp_dof_friction, sp_df.data(), sp_df.size());
  pushNums(MJB_N_sp_tendon_friction, sp_tf.data(), sp_tf.size());
  pushNums(MJB_N_sp_eq, sp_eq.data(), sp_eq.size());
  pushNums(MJB_N_eq_num, eq_num.data(), eq_num.size());
  pushNums(MJB_N_cand_num, cand_num.data(), cand_num.size());
  pushNums(MJB_N_scan_bound, scan_bound.data(), scan_bound.size());
  pushNums(MJB_N_scan_misc, scan_misc.data(), scan_misc.size());
  pushNums(MJB_N_sensor_cutoff, sensor_cutoff.data(), sensor_cutoff.size());
  pushNums(MJB_N_act_biasvel, act_biasvel.data(), act_biasvel.size());
  pushNums(MJB_N_fluid_body, fluid_body.data(), fluid_body.size());
  pushNums(MJB_N_fluid_geom, fluid_geom.data(), fluid_geom.size());
  {
    // focal lengths in pixels exactly as cam_project forms them (engine_sensor.c:155-160: float
    // arithmetic for the intrinsic form, the host's tan for the field-of-view form)
    std::vector<double> cam_proj((size_t)4 * m->ncam);
    for (int i = 0; i < m->ncam; i++) {
      const float* cam_intrinsic = m->cam_intrinsic + 4*i;
      const float* cam_sensorsize = m->cam_sensorsize + 2*i;
      const int* cam_res = m->cam_resolution + 2*i;
      mjtNum fx, fy;
      if (cam_sensorsize[0] && cam_sensorsize[1]) {
        fx = cam_intrinsic[0] / cam_sensorsize[0] * cam_res[0];
        fy = cam_intrinsic[1] / cam_sensorsize[1] * cam_res[1];
      } else {
        fx = fy = .5 / std::tan(m->cam_fovy[i] * mjPI / 360.) * cam_res[1];
      }
      cam_proj[4*i] = fx; cam_proj[4*i + 1] = fy;
      cam_proj[4*i + 2] = (mjtNum)cam_res[0] / 2.0; cam_proj[4*i + 3] = (mjtNum)cam_res[1] / 2.0;
    }
    pushNums(MJB_N_cam_proj, cam_proj.data(), cam_proj.size());
  }

  // scratch layout
  {
    const int nb = m->nbody, ng = m->ngeom, nt = m->ntendon;
    int sizes[MJB_SC_COUNT];
    sizes[MJB_SC_xpos] = 3*nb; sizes[MJB_SC_xquat] = 4*nb; sizes[MJB_SC_origin] = 3*nb;
    sizes[MJB_SC_geom_xpos] = 4*ng; sizes[MJB_SC_geom_xmat] = 9*ng; sizes[MJB_SC_geom_zaxis] = 4*ng;
    sizes[MJB_SC_cinert] = 10*nb; sizes[MJB_SC_cdof] = 6*nv; sizes[MJB_SC_cvel] = 6*nb;
    sizes[MJB_SC_cacc_lin] = 6*nb; sizes[MJB_SC_cacc] = 6*nb;
    sizes[MJB_SC_cfrc] = 6*nb; sizes[MJB_SC_cfrc_ext] = 6*nb; sizes[MJB_SC_cfrc_ext1] = 6*nb; sizes[MJB_SC_qfrc_c] = nv;
    sizes[MJB_SC_qfrc_passive] = nv;
    sizes[MJB_SC_ten_length] = nt; sizes[MJB_SC_ten_velocity] = nt; sizes[MJB_SC_ten_acc] = nt;
    sizes[MJB_SC_crb] = 10*nb; sizes[MJB_SC_ia] = 21*nb;
    sizes[MJB_SC_cfrc_gc] = (gravcomp || spatial_passive || fluid) ? 6*nb : 0;
    {
      bool has_weld = false;
      for (int i = 0; i < m->neq; i++) has_weld = has_weld || m->eq_type[i] == mjEQ_WELD;
      sizes[MJB_SC_weld_dt] = has_weld ? 3*m->neq : 0;
    }
    sizes[MJB_SC_tree_sphere] = 4*ntree;
    sizes[MJB_SC_crec] = ncand > 0 ? 16*nb : 0;
    int off = 0;
    for (int s = 0; s < MJB_SC_COUNT; s++) { H.scoff[s] = off; off += sizes[s]; }
    H.nscratch = off;
  }

  const size_t hdr_bytes = (sizeof(mjbHdr) + 127) & ~(size_t)127;
  const size_t int_bytes = (ints.size()*sizeof(int) + 127) & ~(size_t)127;
  const size_t num_bytes = (nums.size()*sizeof(double) + 127) & ~(size_t)127;
  const size_t cold_int_bytes = (cold_ints.size()*sizeof(int) + 127) & ~(size_t)127;
  const size_t cold_num_bytes = (cold_nums.size()*sizeof(double) + 15) & ~(size_t)15;
  H.int_section = (int)hdr_bytes;
  H.num_section = (int)(hdr_bytes + int_bytes);
  H.staged_bytes = (int)(hdr_bytes + int_bytes + num_bytes);
  const size_t cold_int_start = (size_t)H.staged_bytes, cold_num_start = cold_int_start + cold_int_bytes;
  for (int id : cold_int_ids) H.ioff[id] += (int)((cold_int_start - (size_t)H.int_section) / sizeof(int));
  for (int id : cold_num_ids) H.noff[id] += (int)((cold_num_start - (size_t)H.num_section) / sizeof(double));
  H.bytes = (int)(cold_num_start + cold_num_bytes);
  blob.assign((size_t)H.bytes, 0);
  std::memcpy(blob.data(), &H, sizeof(H));
  if (!ints.empty()) std::memcpy(blob.data() + H.int_section, ints.data(), ints.size()*sizeof(int));
  if (!nums.empty()) std::memcpy(blob.data() + H.num_section, nums.data(), nums.size()*sizeof(double));
  if (!cold_ints.empty()) std::memcpy(blob.data() + cold_int_start, cold_ints.data(), cold_ints.size()*sizeof(int));
  if (!cold_nums.empty()) std::memcpy(blob.data() + cold_num_start, cold_nums.data(), cold_nums.size()*sizeof(double));
  return true;
}

const char* scratchSlotName(int slot) {
  static const char* names[MJB_SC_COUNT] = {
    "xpos", "xquat", "origin", "geom_xpos", "geom_xmat",
    "cinert", "cdof", "cvel", "cacc_lin", "cacc", "cfrc",
    "cfrc_ext", "cfrc_ext1", "qfrc_c", "qfrc_passive", "ten_length", "ten_velocity", "ten_acc", "crb", "ia",
    "cfrc_gc", "weld_dt", "tree_sphere"};
  return (slot >= 0 && slot < MJB_SC_COUNT) ? names[slot] : nullptr;
}

}  // namespace mjb
