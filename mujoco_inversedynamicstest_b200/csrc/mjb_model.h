// Device-side model image ("model blob") for the batched inverse-dynamics kernels.
//
// mjb_makeData() flattens the constant tables of an mjModel that the mj_inverse pipeline reads
// (reference: include/mujoco/mjmodel.h:593-1155) into ONE contiguous buffer:
//
//     [ mjbHdr | int section | double section ]          (16-byte aligned, size % 16 == 0)
//
// so a CTA can stage it into shared memory with a single TMA bulk copy (cp.async.bulk) and every
// thread then reads topology through warp-uniform shared-memory broadcasts. Arrays copied verbatim
// from mjModel keep their reference names; derived tables (candidate geom pairs with pre-mixed
// contact parameters, pre-clamped solver parameters, the sparse C layout of qLD) are built once on
// the host by mjb_upload.cc following the reference functions cited there.
//
// This header is shared by host C++ (upload) and device code; it has no dependency on the
// reference headers.
#ifndef MJB_MODEL_H_
#define MJB_MODEL_H_

#include <stdint.h>

// ---- integer arrays copied 1:1 from mjModel (name, rows-expression evaluated on mjModel* m)
#define MJB_INT_ARRAYS(X)        \
  X(body_parentid, nbody)        \
  X(body_rootid, nbody)          \
  X(body_weldid, nbody)          \
  X(body_mocapid, nbody)         \
  X(body_jntnum, nbody)          \
  X(body_jntadr, nbody)          \
  X(body_dofnum, nbody)          \
  X(body_dofadr, nbody)          \
  X(body_geomnum, nbody)         \
  X(body_geomadr, nbody)         \
  X(jnt_type, njnt)              \
  X(jnt_qposadr, njnt)           \
  X(jnt_dofadr, njnt)            \
  X(jnt_bodyid, njnt)            \
  X(dof_bodyid, nv)              \
  X(dof_jntid, nv)               \
  X(dof_parentid, nv)            \
  X(dof_Madr, nv)                \
  X(dof_simplenum, nv)           \
  X(site_bodyid, nsite)          \
  X(site_type, nsite)            \
  X(geom_type, ngeom)            \
  X(geom_bodyid, ngeom)          \
  X(tendon_adr, ntendon)         \
  X(tendon_num, ntendon)         \
  X(wrap_type, nwrap)            \
  X(wrap_objid, nwrap)           \
  X(cam_mode, ncam)              \
  X(cam_bodyid, ncam)            \
  X(cam_targetbodyid, ncam)      \
  X(light_mode, nlight)          \
  X(light_bodyid, nlight)        \
  X(light_targetbodyid, nlight)  \
  X(actuator_trntype, nu)

// ---- byte arrays of mjModel widened to int
#define MJB_BYTE_ARRAYS(X)       \
  X(body_sameframe, nbody)       \
  X(jnt_limited, njnt)           \
  X(jnt_actgravcomp, njnt)       \
  X(geom_sameframe, ngeom)       \
  X(site_sameframe, nsite)       \
  X(tendon_limited, ntendon)

// ---- derived integer tables (built by mjb_upload.cc)
#define MJB_DERIVED_INT_ARRAYS(X) \
  X(C_rownnz)  /* nv   : qLD row lengths   (engine_io.c:929-1018)              */ \
  X(C_rowadr)  /* nv   : qLD row starts                                         */ \
  X(C_colind)  /* nC   : qLD column indices (ancestors ascending, self last)    */ \
  X(mapM2C)    /* nC   : qLD[k] = qM[mapM2C[k]] (engine_io.c:1135-1188)         */ \
  X(cand_int)  /* ncand*MJB_CAND_NI : candidate geom pairs, see MJB_CI_*        */ \
  X(eq_int)    /* neq*MJB_EQ_NI : equality constraints, see MJB_EQI_*           */ \
  X(body_static) /* nbody: 1 if no dof on the chain to the world (jac == 0)     */ \
  X(jnt_dofnum_tab) /* njnt : dofs of this joint                                */ \
  X(ray_geom) /* ngeom (empty without rangefinder sensors): 1 if mj_ray tests the geom (visible: ray_eliminate) */ \
  X(actuator_trn) /* nu*2 : actuator_trnid (object, reference / slider site) */ \
  X(tendon_active) /* ntendon: 1 if the tendon carries a force (limit, friction loss, spring, damper) */ \
  X(geom_store) /* ngeom: bit0 position + z axis read by a later phase, bit1 full frame, bit2 full frame in runs with transmission / sensor outputs only */ \
  X(dof_frow)  /* nv   : friction-loss row of the dof within the friction block, -1 if none   */ \
  X(sensor_int) /* nsensor*MJB_SEN_NI : sensors evaluated on the device, see MJB_SEN_*       */ \
  X(sensor_pairs) /* 4 per geom pair of the geom-distance sensors: geom 1, geom 2 (type-ordered like        \
                     mj_geomDistance, engine_support.c:1412-1417), narrow-phase function, 1 if flipped */ \
  X(scan_int)  /* ncand*2 : compact rows of the bounding-sphere scan: geom 1 | filter kind << 28, geom 2      */ \
  X(scan_run)  /* nrun*4  : runs of consecutive candidates with the same (tree of body 1, tree of body 2):   \
                            first candidate, count, tree 1, tree 2 (-1: static body, never culled)           */ \
  X(tree_int)  /* ntree*3 : kinematic trees: root body, first geom, one past the last geom                   */ \
  X(body_tree_flags) /* nbody: bit0 has child bodies, bit1 highest-index child of its parent, \
                                 bit2 has a child other than body+1 (forward-sweep carry must be stored), \
                                 bit3 pose read by an equality constraint or a tendon site, \
                                 bit4 velocity carrier rows read by constraint rows (equality, tendon site), \
                                 bit5 body of a candidate pair: its carrier record (MJB_SC_crec) is read by contact rows */

// ---- double arrays copied 1:1 from mjModel (name, rows, cols)
#define MJB_NUM_ARRAYS(X)         \
  X(qpos0, nq, 1)                 \
  X(qpos_spring, nq, 1)           \
  X(body_pos, nbody, 3)           \
  X(body_quat, nbody, 4)          \
  X(body_ipos, nbody, 3)          \
  X(body_iquat, nbody, 4)         \
  X(body_mass, nbody, 1)          \
  X(body_subtreemass, nbody, 1)   \
  X(body_inertia, nbody, 3)       \
  X(body_invweight0, nbody, 2)    \
  X(body_gravcomp, nbody, 1)      \
  X(jnt_pos, njnt, 3)             \
  X(jnt_axis, njnt, 3)            \
  X(jnt_stiffness, njnt, 1)       \
  X(jnt_range, njnt, 2)           \
  X(jnt_margin, njnt, 1)          \
  X(dof_armature, nv, 1)          \
  X(dof_damping, nv, 1)           \
  X(dof_frictionloss, nv, 1)      \
  X(dof_invweight0, nv, 1)        \
  X(dof_M0, nv, 1)                \
  X(geom_size, ngeom, 3)          \
  X(geom_rbound, ngeom, 1)        \
  X(geom_pos, ngeom, 3)           \
  X(geom_quat, ngeom, 4)          \
  X(site_pos, nsite, 3)           \
  X(site_quat, nsite, 4)          \
  X(site_size, nsite, 3)          \
  X(tendon_range, ntendon, 2)     \
  X(tendon_margin, ntendon, 1)    \
  X(tendon_stiffness, ntendon, 1) \
  X(tendon_damping, ntendon, 1)   \
  X(tendon_frictionloss, ntendon, 1) \
  X(tendon_lengthspring, ntendon, 2) \
  X(tendon_invweight0, ntendon, 1)   \
  X(tendon_length0, ntendon, 1)      \
  X(eq_data, neq, 11)                \
  X(wrap_prm, nwrap, 1)              \
  X(cam_pos, ncam, 3)                \
  X(cam_quat, ncam, 4)               \
  X(cam_poscom0, ncam, 3)            \
  X(cam_pos0, ncam, 3)               \
  X(cam_mat0, ncam, 9)               \
  X(light_pos, nlight, 3)            \
  X(light_dir, nlight, 3)            \
  X(light_poscom0, nlight, 3)        \
  X(light_pos0, nlight, 3)           \
  X(light_dir0, nlight, 3)           \
  X(actuator_gear, nu, 6)            \
  X(actuator_cranklength, nu, 1)

// ---- derived double tables: pre-clamped solver parameters, MJB_SP_N doubles per record
#define MJB_DERIVED_NUM_ARRAYS(X) \
  X(sp_jnt_limit)     /* njnt    */ \
  X(sp_tendon_limit)  /* ntendon */ \
  X(sp_dof_friction)  /* nv      */ \
  X(sp_tendon_friction) /* ntendon */ \
  X(sp_eq)            /* neq     */ \
  X(eq_num)           /* neq*MJB_EQ_NN: site offsets / quaternions of site-defined constraints */ \
  X(cand_num)         /* ncand*MJB_CAND_NN, see MJB_CN_* */ \
  X(scan_bound)       /* ncand: MJB_CN_RBOUND of every candidate, contiguous (the scan reads nothing else) */ \
  X(scan_misc)        /* 1: largest contact margin of any candidate (tree-level culling) */ \
  X(sensor_cutoff)    /* nsensor */ \
  X(act_biasvel)      /* nu: d force / d velocity of the actuator's affine bias (implicitfast mjENBL_INVDISCRETE) */ \
  X(cam_proj)         /* ncam*4: fx, fy, half width, half height in pixels (cam_project, engine_sensor.c:126-215) */ \
  X(fluid_body)       /* nbody*4 (empty without fluid forces): model kind MJB_FLUID_* and the three sides of the \
                         equivalent inertia box (mj_inertiaBoxFluidModel, engine_passive.c:527-537) */ \
  X(fluid_geom)       /* ngeom*MJB_FLUID_NG (empty unless some geom uses the ellipsoid model): geom_fluid's 12 \
                         coefficients, then the semi-axes of mju_geomSemiAxes (engine_util_misc.c:425) */

enum {
#define X(name, rows) MJB_I_##name,
  MJB_INT_ARRAYS(X)
  MJB_BYTE_ARRAYS(X)
#undef X
#define X(name) MJB_I_##name,
  MJB_DERIVED_INT_ARRAYS(X)
#undef X
  MJB_NI
};

enum {
#define X(name, rows, cols) MJB_N_##name,
  MJB_NUM_ARRAYS(X)
#undef X
#define X(name) MJB_N_##name,
  MJB_DERIVED_NUM_ARRAYS(X)
#undef X
  MJB_NN
};

// solver-parameter record (doubles): solimp after the clamps of getsolparam
// (engine_core_constraint.c:1379-1384) and K, B of mj_makeImpedance (:1523-1545), which depend
// only on model constants and opt.timestep.
enum { MJB_SP_D0 = 0, MJB_SP_D1, MJB_SP_WIDTH, MJB_SP_MID, MJB_SP_POWER, MJB_SP_K, MJB_SP_B, MJB_SP_N };

// candidate geom pair: integer columns
enum {
  MJB_CI_G1 = 0,   // geom ids, already ordered so that geom_type[g1] <= geom_type[g2]
  MJB_CI_G2,
  MJB_CI_FUNC,     // narrow-phase function id, MJB_FN_*
  MJB_CI_DIM,      // condim after mj_contactParam / pair_dim
  MJB_CI_B1,       // geom_bodyid[g1], geom_bodyid[g2]
  MJB_CI_B2,
  MJB_CI_FLAGS,    // bit0: both bodies static (contact gets exclude=3, engine_core_constraint.c:1072)
  MJB_CI_PLANE,    // bit0: g1 is a plane (sphere filter uses plane distance)
  MJB_CAND_NI
};

// candidate geom pair: double columns
enum {
  MJB_CN_MARGIN = 0,      // mj_assignMargin(max geom_margin) or pair_margin
  MJB_CN_INCLUDEMARGIN,   // margin - gap
  MJB_CN_RBOUND,          // geom_rbound[g1] + geom_rbound[g2] + margin (0-plane case: rbound[g2] + margin)
  MJB_CN_FRICTION,        // 5 doubles after mj_assignFriction
  MJB_CN_SP = MJB_CN_FRICTION + 5,        // MJB_SP_N doubles for the normal direction
  MJB_CN_BFRIC = MJB_CN_SP + MJB_SP_N,    // B for elliptic friction rows (solreffriction or solref)
  MJB_CN_DA_TRAN,         // body_invweight0[2*b1] + body_invweight0[2*b2]
  MJB_CN_DA_ROT,          // body_invweight0[2*b1+1] + body_invweight0[2*b2+1]
  MJB_CN_SOLREF,          // 2: contact.solref as stored in mjContact (for output)
  MJB_CN_SOLIMP = MJB_CN_SOLREF + 2,      // 5: contact.solimp as stored (unclamped)
  MJB_CAND_NN = MJB_CN_SOLIMP + 5
};

// sensor: integer columns (mjModel sensor_* arrays, include/mujoco/mjmodel.h)
enum { MJB_SEN_TYPE = 0, MJB_SEN_DATATYPE, MJB_SEN_OBJTYPE, MJB_SEN_OBJID, MJB_SEN_REFTYPE, MJB_SEN_REFID,
       MJB_SEN_DIM, MJB_SEN_ADR, MJB_SEN_NI };
// mjtSensor / mjtObj / mjtDataType values restated (include/mujoco/mjmodel.h)
enum { MJB_SENS_TOUCH = 0, MJB_SENS_ACCELEROMETER = 1, MJB_SENS_VELOCIMETER = 2, MJB_SENS_GYRO = 3, MJB_SENS_FORCE = 4,
       MJB_SENS_TORQUE = 5, MJB_SENS_MAGNETOMETER = 6, MJB_SENS_RANGEFINDER = 7, MJB_SENS_CAMPROJECTION = 8, MJB_SENS_JOINTPOS = 9, MJB_SENS_JOINTVEL = 10, MJB_SENS_TENDONPOS = 11,
       MJB_SENS_TENDONVEL = 12, MJB_SENS_ACTUATORPOS = 13, MJB_SENS_ACTUATORVEL = 14, MJB_SENS_BALLQUAT = 17, MJB_SENS_BALLANGVEL = 18, MJB_SENS_JOINTLIMITPOS = 19,
       MJB_SENS_JOINTLIMITVEL = 20, MJB_SENS_JOINTLIMITFRC = 21, MJB_SENS_TENDONLIMITPOS = 22,
       MJB_SENS_TENDONLIMITVEL = 23, MJB_SENS_TENDONLIMITFRC = 24, MJB_SENS_FRAMEPOS = 25,
       MJB_SENS_FRAMEQUAT = 26, MJB_SENS_FRAMEXAXIS = 27, MJB_SENS_FRAMEYAXIS = 28, MJB_SENS_FRAMEZAXIS = 29,
       MJB_SENS_FRAMELINVEL = 30, MJB_SENS_FRAMEANGVEL = 31, MJB_SENS_FRAMELINACC = 32,
       MJB_SENS_FRAMEANGACC = 33, MJB_SENS_SUBTREECOM = 34, MJB_SENS_SUBTREELINVEL = 35,
       MJB_SENS_SUBTREEANGMOM = 36, MJB_SENS_GEOMDIST = 37, MJB_SENS_GEOMNORMAL = 38, MJB_SENS_GEOMFROMTO = 39,
       MJB_SENS_E_POTENTIAL = 40, MJB_SENS_E_KINETIC = 41, MJB_SENS_CLOCK = 42 };
enum { MJB_OBJ_BODY = 1, MJB_OBJ_XBODY = 2, MJB_OBJ_GEOM = 5, MJB_OBJ_SITE = 6 };
enum { MJB_DATATYPE_REAL = 0, MJB_DATATYPE_POSITIVE = 1 };

// equality constraint: integer columns (mj_instantiateEquality, engine_core_constraint.c:493-763)
enum {
  MJB_EQI_TYPE = 0,   // mjtEq: 0 connect, 1 weld, 2 joint, 3 tendon
  MJB_EQI_ACTIVE,     // eq_active0
  MJB_EQI_B0,         // body of object 1 / joint id / tendon id
  MJB_EQI_B1,         // body of object 2 / joint id / tendon id (-1: none)
  MJB_EQI_SITE,       // 1: site semantics (anchors and orientations come from MJB_EQ_NN numbers)
  MJB_EQI_SKIP,       // 1: both bodies static -> all-zero Jacobian, constraint dropped (:284-336)
  MJB_EQ_NI
};
// equality constraint: double columns
enum {
  MJB_EQN_ANCHOR0 = 0,   // 3: point on object-1 body, body frame
  MJB_EQN_ANCHOR1 = 3,   // 3: point on object-2 body, body frame
  MJB_EQN_Q0 = 6,        // 4: orientation offset applied to xquat of body 0 (relpose / site_quat)
  MJB_EQN_Q1 = 10,       // 4: orientation offset applied to xquat of body 1 (identity / site_quat)
  MJB_EQN_TORQUESCALE = 14,
  MJB_EQN_DA_TRAN = 15,  // body_invweight0 sums (translation, rotation) / dof or tendon invweight sum
  MJB_EQN_DA_ROT = 16,
  MJB_EQ_NN = 17
};

// narrow-phase function ids (engine_collision_driver.c:41-52, only primitive pairs)
enum {
  MJB_FN_PLANE_SPHERE = 0,
  MJB_FN_PLANE_CAPSULE,
  MJB_FN_PLANE_CYLINDER,
  MJB_FN_PLANE_BOX,
  MJB_FN_PLANE_ELLIPSOID,
  MJB_FN_SPHERE_SPHERE,
  MJB_FN_SPHERE_CAPSULE,
  MJB_FN_SPHERE_CYLINDER,
  MJB_FN_SPHERE_BOX,
  MJB_FN_CAPSULE_CAPSULE,
  MJB_FN_CAPSULE_BOX,
  MJB_FN_BOX_BOX,
  MJB_FN_CONVEX,     // mjc_Convex: GJK / EPA on the two support mappings (csrc/mjb_convex.h)
  MJB_FN_COUNT
};

// per-thread scratch slots (doubles), offsets in units of one double per thread
enum {
  MJB_SC_xpos = 0,     // nbody*3
  MJB_SC_xquat,        // nbody*4
  MJB_SC_origin,       // nbody*3   origin of the spatial frame of a kinematic tree, stored at its ROOT body
  MJB_SC_geom_xpos,    // ngeom*4   geom position as a 4-double VECTOR per (geom, state): lane l of geom g at (4*g*32 + 4*l);
                       //           one coalesced 256-bit store per geom, one whole sector per gather (mjb_pipeline.h geom_vec)
  MJB_SC_geom_xmat,    // ngeom*9   full frame, rows; only geoms whose pairs need more than the z axis (geom_store bit 1)
  MJB_SC_cinert,       // nbody*10
  MJB_SC_cdof,         // nv*6
  MJB_SC_cvel,         // nbody*6
  MJB_SC_cacc_lin,     // nbody*6   sum cdof*qacc over the dof chain (J*qacc carrier)
  MJB_SC_cacc,         // nbody*6   rne accelerations (forward-sweep carry; stored where a later child reads it)
  MJB_SC_cfrc,         // nbody*6   rne body forces
  MJB_SC_cfrc_ext,     // nbody*6   constraint wrenches on bodies, '+' side (body 2 of a pair)
  MJB_SC_cfrc_ext1,    // nbody*6   '-' side (body 1 of a pair); kept apart so that each sum runs in contact order
  MJB_SC_qfrc_c,       // nv        joint-space constraint force (limits, friction loss, tendons)
  MJB_SC_qfrc_passive, // nv
  MJB_SC_ten_length,   // ntendon
  MJB_SC_ten_velocity, // ntendon
  MJB_SC_ten_acc,      // ntendon   ten_J * qacc
  MJB_SC_crb,          // nbody*10  composite rigid-body inertias
  MJB_SC_ia,           // nbody*21  articulated-body inertias (symmetric 6x6, upper triangle)
  MJB_SC_cfrc_gc,      // nbody*6   passive body wrenches: gravcomp, spatial-tendon springs/dampers (only if needed)
  MJB_SC_weld_dt,      // neq*3     weld rows: (raw rotational efc_force) - (torque J'f), for the cfrc_ext output (only models with welds)
  MJB_SC_tree_sphere,  // ntree*4   bounding sphere of every kinematic tree (tree-level broadphase; only with scan runs)
  MJB_SC_geom_zaxis,   // ngeom*4   z axis of the geom frame (plane normal, capsule axis), vector layout like geom_xpos
  MJB_SC_crec,         // nbody*16  contact carrier RECORDS (only with candidate pairs): unlike every other array this
                       //           one is state-major inside the warp block -- body b, lane l at (16*b*32 + 16*l) --
                       //           so that the 16 doubles [cvel 6 | cacc_lin 6 | tree origin 3 | 0] a contact row
                       //           gathers for one (state, body) are ONE 128-byte line instead of 15 sectors
  MJB_SC_COUNT
};

typedef struct mjbHdr_ {
  uint32_t magic;
  int32_t bytes;            // total blob size (multiple of 16)
  int32_t staged_bytes;     // leading part of the blob a CTA stages into shared memory (everything but the tables
                            // of the output-only kernels: mj_camlight, mj_transmission, implicit mj_discreteAcc)
  // sizes (reference names)
  int32_t nq, nv, nbody, njnt, ngeom, ntendon, nwrap, neq, nM, nC;
  int32_t ncand;            // candidate geom pairs
  int32_t disableflags, enableflags, cone;
  int32_t has_gravcomp;     // ngravcomp>0 && gravity enabled && |gravity|>0 (engine_passive.c:383)
  int32_t has_fixed_tendon_only;
  int32_t nscratch;         // doubles of scratch per state
  int32_t max_pair_contacts; // most contacts one candidate pair can yield (1, 2 or 4)
  // constraint-row counts that depend only on the model: equality rows, friction-loss rows of
  // dofs, friction-loss rows in total (dofs then tendons)
  int32_t ne_rows, nf_dof_rows, nf_rows;
  int32_t has_spatial;      // some spatial tendon carries a force (its path is walked on the device)
  int32_t passive_wrench;   // the passive body-wrench carrier exists (gravcomp or spatial-tendon springs/dampers)
  int32_t discrete_acc;     // mjENBL_INVDISCRETE: qacc is converted first; 1 Euler with damped dofs, 2 implicitfast, 3 implicit
  int32_t discrete_trn;     // implicitfast with velocity-biased actuators: the conversion reads actuator_moment
  int32_t nsensor;          // sensors evaluated on the device (0 with mjDSBL_SENSOR), nsensordata their rows
  int32_t nsensordata;
  int32_t sensor_post;      // some sensor reads cacc / cfrc_int (mj_rnePostConstraint, engine_sensor.c:727-740)
  int32_t nsite;
  int32_t nmocap;
  int32_t ncam, nlight;     // cameras / lights (mj_camlight outputs, mjbOUT_CAMLIGHT)
  int32_t nu;               // actuators (mj_transmission outputs, mjbOUT_TRANSMISSION)
  int32_t sensor_subtreevel; // some sensor reads subtree_linvel / subtree_angmom (mj_subtreeVel)
  int32_t sensor_cam;        // the sensor kernel reads tables behind the staged part of the blob (cam_proj, ray_geom)
  int32_t sensor_camlight;   // some camprojection sensor reads cam_xpos / cam_xmat (mj_camlight runs for it)
  int32_t sensor_trn;        // some actuatorpos / actuatorvel sensor reads mj_transmission's outputs
  int32_t sensor_energy;     // some potential / kinetic energy sensor (mj_energyPos / mj_energyVel run for it)
  int32_t sensor_touch;      // some touch sensor reads the contact list and the contact rows' forces
  int32_t has_fluid;         // mj_fluid runs (opt.density / opt.viscosity > 0, passive forces enabled): 1 inertia-box
                             // model only, 2 some body uses the ellipsoid model (engine_passive.c:403-431)
  int32_t sensor_ccd;        // some geom-distance sensor measures a box-box or convex pair (mjc_ccd with a cut-off)
  int32_t has_convex;        // some candidate pair goes through mjc_Convex (GJK / EPA, csrc/mjb_convex.h)
  int32_t ccd_iterations;    // opt.ccd_iterations (<= MJB_CVX_MAXIT)
  int32_t simple_pairs;      // every candidate pair is plane/sphere/capsule against sphere/capsule (<= 2 contacts, z axes only)
  double timestep, impratio;
  double gravity[3];
  double magnetic[3];        // opt.magnetic (magnetometer sensors)
  double density, viscosity; // opt.density, opt.viscosity (mj_fluid)
  double wind[3];            // opt.wind
  double ccd_tolerance;      // opt.ccd_tolerance
  int32_t nrun;             // runs of the candidate list for the tree-level broadphase (0: flat scan)
  int32_t ntree;            // kinematic trees with collidable geoms
  int32_t ioff[MJB_NI];     // element offsets into the int section
  int32_t noff[MJB_NN];     // element offsets into the double section
  int32_t scoff[MJB_SC_COUNT];  // scratch slot offsets (doubles per thread)
  int32_t int_section;      // byte offset of the int section from blob start
  int32_t num_section;      // byte offset of the double section from blob start
} mjbHdr;

#define MJB_MAGIC 0x6d6a6231u  /* "mjb1" */

// reference enum values restated (include/mujoco/mjmodel.h:49-82, :85-118, :160-170, :273-292, :379-385)
enum { MJB_DSBL_CONSTRAINT = 1<<0, MJB_DSBL_EQUALITY = 1<<1, MJB_DSBL_FRICTIONLOSS = 1<<2,
       MJB_DSBL_LIMIT = 1<<3, MJB_DSBL_CONTACT = 1<<4, MJB_DSBL_PASSIVE = 1<<5,
       MJB_DSBL_GRAVITY = 1<<6, MJB_DSBL_FILTERPARENT = 1<<9, MJB_DSBL_REFSAFE = 1<<11,
       MJB_DSBL_SENSOR = 1<<12, MJB_DSBL_MIDPHASE = 1<<13 };
enum { MJB_ENBL_OVERRIDE = 1<<0, MJB_ENBL_ENERGY = 1<<1, MJB_ENBL_INVDISCRETE = 1<<3 };
enum { MJB_JNT_FREE = 0, MJB_JNT_BALL, MJB_JNT_SLIDE, MJB_JNT_HINGE };
enum { MJB_GEOM_PLANE = 0, MJB_GEOM_HFIELD, MJB_GEOM_SPHERE, MJB_GEOM_CAPSULE, MJB_GEOM_ELLIPSOID,
       MJB_GEOM_CYLINDER, MJB_GEOM_BOX, MJB_GEOM_MESH, MJB_GEOM_SDF };
enum { MJB_CNSTR_EQUALITY = 0, MJB_CNSTR_FRICTION_DOF, MJB_CNSTR_FRICTION_TENDON,
       MJB_CNSTR_LIMIT_JOINT, MJB_CNSTR_LIMIT_TENDON, MJB_CNSTR_CONTACT_FRICTIONLESS,
       MJB_CNSTR_CONTACT_PYRAMIDAL, MJB_CNSTR_CONTACT_ELLIPTIC };
enum { MJB_STATE_SATISFIED = 0, MJB_STATE_QUADRATIC, MJB_STATE_LINEARNEG, MJB_STATE_LINEARPOS,
       MJB_STATE_CONE };
enum { MJB_TRN_JOINT = 0, MJB_TRN_JOINTINPARENT, MJB_TRN_SLIDERCRANK, MJB_TRN_TENDON, MJB_TRN_SITE, MJB_TRN_BODY };
enum { MJB_CAMLIGHT_FIXED = 0, MJB_CAMLIGHT_TRACK, MJB_CAMLIGHT_TRACKCOM, MJB_CAMLIGHT_TARGETBODY, MJB_CAMLIGHT_TARGETBODYCOM };
enum { MJB_SAMEFRAME_NONE = 0, MJB_SAMEFRAME_BODY, MJB_SAMEFRAME_INERTIA, MJB_SAMEFRAME_BODYROT,
       MJB_SAMEFRAME_INERTIAROT };
enum { MJB_WRAP_NONE = 0, MJB_WRAP_JOINT, MJB_WRAP_PULLEY, MJB_WRAP_SITE, MJB_WRAP_SPHERE, MJB_WRAP_CYLINDER };
enum { MJB_FLUID_NONE = 0, MJB_FLUID_BOX, MJB_FLUID_ELLIPSOID };   // fluid_body[4*b]
#define MJB_FLUID_NG 16   // doubles per geom in fluid_geom: geom_fluid[mjNFLUID = 12], semi-axes[3], 0
#define MJB_CVX_MAXIT 64   // most GJK / EPA iterations a model may ask for (opt.ccd_iterations; default 50)
#define MJB_MINVAL 1E-15
#define MJB_MINMU 1E-5
#define MJB_MINIMP 0.0001
#define MJB_MAXIMP 0.9999
#define MJB_PI 3.14159265358979323846
#define MJB_MAXVAL 1E+10

#endif  // MJB_MODEL_H_
