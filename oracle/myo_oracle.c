/* ORACLE / TEST INFRASTRUCTURE -- not part of the product path.
 *
 * fp64, single-world, plain-C restatement of the MuJoCo 2.1.0 forward dynamics subset that the
 * reference's hot path executes:
 *
 *   env.step -> MyoSuite Robot.step -> mujoco-py MjSim.step -> mj_step
 *   reference call sites: /root/reference/src/envs/baoding.py:183,625 (self.step),
 *   :179,608 (self.robot.reset), :206,632 (self.set_state); /root/reference/src/envs/pose.py:102.
 *
 * The arithmetic lives in a third-party dependency that is NOT vendored in the reference:
 *   MuJoCo 2.1.0 (libmujoco210 via free-mujoco-py==2.1.6, /root/reference/requirements.txt:41).
 * Its published algorithm is restated here stage by stage (file names are MuJoCo's):
 *   engine_core_smooth.c   mj_kinematics, mj_comPos, mj_tendon, mj_transmission, mj_crb,
 *                          mj_factorM/mj_solveM, mj_comVel, mj_rne
 *   engine_util_misc.c     mju_wrap (sphere / cylinder, side sites), mju_muscle{Gain,Bias,Dynamics}
 *   engine_passive.c       mj_passive
 *   engine_collision_*.c   broad phase filters, plane/sphere/capsule primitives
 *   engine_core_constraint.c  limits, pyramidal contacts, impedance, reference acceleration
 *   engine_solver.c        primal Newton on the convex constraint cost (solved to 1e-14 here)
 *   engine_forward.c       mj_fwdActuation, mj_fwdAcceleration, mj_Euler (implicit in joint
 *                          damping), mj_advance
 *   engine_setconst.c      mj_setConst (dof_M0, *_invweight0, tendon_length0, actuator_acc0 ...)
 *
 * PINNING: the derived constants MuJoCo itself stored in the reference's .mjb files
 * (dof_M0, dof_invweight0, body_invweight0, body_subtreemass, tendon_length0,
 * tendon_invweight0, actuator_length0, actuator_acc0) are reproduced by o_set_const(); see
 * tests/test_oracle_golden.py.  One-step qpos/qvel/act from MuJoCo are NOT available in this
 * container (no MuJoCo binary): the dynamics stages are "parity unpinned" beyond those constants.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference arm may load this.
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define PI 3.14159265358979323846
#define MINIMP 0.0001
#define MAXIMP 0.9999

enum { JNT_FREE = 0, JNT_BALL = 1, JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { GEOM_PLANE = 0, GEOM_HFIELD = 1, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_ELLIPSOID = 4,
       GEOM_CYLINDER = 5, GEOM_BOX = 6, GEOM_MESH = 7 };
enum { WRAP_JOINT = 1, WRAP_PULLEY = 2, WRAP_SITE = 3, WRAP_SPHERE = 4, WRAP_CYLINDER = 5 };
enum { TRN_JOINT = 0, TRN_TENDON = 3 };
enum { DYN_NONE = 0, DYN_INTEGRATOR = 1, DYN_FILTER = 2, DYN_MUSCLE = 3 };
enum { GAIN_FIXED = 0, GAIN_MUSCLE = 1 };
enum { BIAS_NONE = 0, BIAS_AFFINE = 1, BIAS_MUSCLE = 2 };
enum { CNSTR_LIMIT_JOINT = 3, CNSTR_LIMIT_TENDON = 4, CNSTR_CONTACT_FRICTIONLESS = 5,
       CNSTR_CONTACT_PYRAMIDAL = 6 };

/* ------------------------------------------------------------------ model / data --------- */
#define MODEL_FIELDS(X)                                                                        \
  X(double, qpos0, m->nq) X(double, qpos_spring, m->nq)                                        \
  X(int, body_parentid, m->nbody) X(int, body_rootid, m->nbody) X(int, body_weldid, m->nbody)  \
  X(int, body_jntnum, m->nbody) X(int, body_jntadr, m->nbody) X(int, body_dofnum, m->nbody)    \
  X(int, body_dofadr, m->nbody) X(int, body_geomnum, m->nbody) X(int, body_geomadr, m->nbody)  \
  X(int, body_simple, m->nbody)                                                                \
  X(double, body_pos, 3 * m->nbody) X(double, body_quat, 4 * m->nbody)                         \
  X(double, body_ipos, 3 * m->nbody) X(double, body_iquat, 4 * m->nbody)                       \
  X(double, body_mass, m->nbody) X(double, body_subtreemass, m->nbody)                         \
  X(double, body_inertia, 3 * m->nbody) X(double, body_invweight0, 2 * m->nbody)               \
  X(int, jnt_type, m->njnt) X(int, jnt_qposadr, m->njnt) X(int, jnt_dofadr, m->njnt)           \
  X(int, jnt_bodyid, m->njnt) X(int, jnt_limited, m->njnt)                                     \
  X(double, jnt_solref, 2 * m->njnt) X(double, jnt_solimp, 5 * m->njnt)                        \
  X(double, jnt_pos, 3 * m->njnt) X(double, jnt_axis, 3 * m->njnt)                             \
  X(double, jnt_stiffness, m->njnt) X(double, jnt_range, 2 * m->njnt)                          \
  X(double, jnt_margin, m->njnt)                                                               \
  X(int, dof_bodyid, m->nv) X(int, dof_jntid, m->nv) X(int, dof_parentid, m->nv)               \
  X(int, dof_Madr, m->nv) X(int, dof_simplenum, m->nv)                                         \
  X(double, dof_frictionloss, m->nv) X(double, dof_armature, m->nv)                            \
  X(double, dof_damping, m->nv) X(double, dof_invweight0, m->nv) X(double, dof_M0, m->nv)      \
  X(int, geom_type, m->ngeom) X(int, geom_contype, m->ngeom) X(int, geom_conaffinity, m->ngeom)\
  X(int, geom_condim, m->ngeom) X(int, geom_bodyid, m->ngeom) X(int, geom_priority, m->ngeom)  \
  X(double, geom_solmix, m->ngeom) X(double, geom_solref, 2 * m->ngeom)                        \
  X(double, geom_solimp, 5 * m->ngeom) X(double, geom_size, 3 * m->ngeom)                      \
  X(double, geom_rbound, m->ngeom) X(double, geom_pos, 3 * m->ngeom)                           \
  X(double, geom_quat, 4 * m->ngeom) X(double, geom_friction, 3 * m->ngeom)                    \
  X(double, geom_margin, m->ngeom) X(double, geom_gap, m->ngeom)                               \
  X(int, site_bodyid, m->nsite) X(double, site_pos, 3 * m->nsite)                              \
  X(double, site_quat, 4 * m->nsite)                                                           \
  X(int, tendon_adr, m->ntendon) X(int, tendon_num, m->ntendon)                                \
  X(int, tendon_limited, m->ntendon) X(double, tendon_solref_lim, 2 * m->ntendon)              \
  X(double, tendon_solimp_lim, 5 * m->ntendon) X(double, tendon_range, 2 * m->ntendon)         \
  X(double, tendon_margin, m->ntendon) X(double, tendon_stiffness, m->ntendon)                 \
  X(double, tendon_damping, m->ntendon) X(double, tendon_frictionloss, m->ntendon)             \
  X(double, tendon_lengthspring, m->ntendon) X(double, tendon_length0, m->ntendon)             \
  X(double, tendon_invweight0, m->ntendon)                                                     \
  X(int, wrap_type, m->nwrap) X(int, wrap_objid, m->nwrap) X(double, wrap_prm, m->nwrap)       \
  X(int, actuator_trntype, m->nu) X(int, actuator_dyntype, m->nu)                              \
  X(int, actuator_gaintype, m->nu) X(int, actuator_biastype, m->nu)                            \
  X(int, actuator_trnid, 2 * m->nu) X(int, actuator_ctrllimited, m->nu)                        \
  X(int, actuator_forcelimited, m->nu) X(double, actuator_dynprm, 10 * m->nu)                  \
  X(double, actuator_gainprm, 10 * m->nu) X(double, actuator_biasprm, 10 * m->nu)              \
  X(double, actuator_ctrlrange, 2 * m->nu) X(double, actuator_forcerange, 2 * m->nu)           \
  X(double, actuator_gear, 6 * m->nu) X(double, actuator_acc0, m->nu)                          \
  X(double, actuator_length0, m->nu) X(double, actuator_lengthrange, 2 * m->nu)

typedef struct OModel {
  int nq, nv, nu, na, nbody, njnt, ngeom, nsite, ntendon, nwrap, nM, njmax, nconmax;
  double timestep, gravity[3], impratio;
  int cone, disableflags;
  double stat_meaninertia;
  int nexclude; int* exclude_signature;      /* <contact><exclude>: (body1 << 16) + body2, body1 < body2 (mj_collision skips these body pairs) */
#define X(T, n, c) T* n;
  MODEL_FIELDS(X)
#undef X
} OModel;

#define MAXCONDIM 3
#define DATA_FIELDS(X)                                                                         \
  X(double, qpos, m->nq) X(double, qvel, m->nv) X(double, act, m->na) X(double, ctrl, m->nu)   \
  X(double, qacc_warmstart, m->nv) X(double, qfrc_applied, m->nv)                              \
  X(double, xpos, 3 * m->nbody) X(double, xquat, 4 * m->nbody) X(double, xmat, 9 * m->nbody)   \
  X(double, xipos, 3 * m->nbody) X(double, ximat, 9 * m->nbody)                                \
  X(double, xanchor, 3 * m->njnt) X(double, xaxis, 3 * m->njnt)                                \
  X(double, geom_xpos, 3 * m->ngeom) X(double, geom_xmat, 9 * m->ngeom)                        \
  X(double, site_xpos, 3 * m->nsite) X(double, site_xmat, 9 * m->nsite)                        \
  X(double, subtree_com, 3 * m->nbody) X(double, cdof, 6 * m->nv)                              \
  X(double, cinert, 10 * m->nbody) X(double, crb, 10 * m->nbody)                               \
  X(double, ten_length, m->ntendon) X(double, ten_velocity, m->ntendon)                        \
  X(double, ten_J, m->ntendon * m->nv)                                                         \
  X(double, actuator_length, m->nu) X(double, actuator_velocity, m->nu)                        \
  X(double, actuator_moment, m->nu * m->nv) X(double, actuator_force, m->nu)                   \
  X(double, qM, m->nM) X(double, Mdense, m->nv * m->nv) X(double, Lchol, m->nv * m->nv)        \
  X(double, cvel, 6 * m->nbody) X(double, cdof_dot, 6 * m->nv)                                 \
  X(double, qfrc_bias, m->nv) X(double, qfrc_passive, m->nv) X(double, qfrc_actuator, m->nv)   \
  X(double, qfrc_smooth, m->nv) X(double, qacc_smooth, m->nv) X(double, qacc, m->nv)           \
  X(double, qfrc_constraint, m->nv) X(double, act_dot, m->na)                                  \
  X(int, contact_geom1, m->nconmax) X(int, contact_geom2, m->nconmax)                          \
  X(double, contact_dist, m->nconmax) X(double, contact_pos, 3 * m->nconmax)                   \
  X(double, contact_frame, 9 * m->nconmax) X(double, contact_friction, 5 * m->nconmax)         \
  X(double, contact_solref, 2 * m->nconmax) X(double, contact_solimp, 5 * m->nconmax)          \
  X(double, contact_includemargin, m->nconmax) X(int, contact_dim, m->nconmax)                 \
  X(int, contact_efc_address, m->nconmax)                                                      \
  X(int, efc_type, m->njmax) X(int, efc_id, m->njmax)                                          \
  X(double, efc_J, m->njmax * m->nv) X(double, efc_pos, m->njmax)                              \
  X(double, efc_margin, m->njmax) X(double, efc_diagApprox, m->njmax)                          \
  X(double, efc_R, m->njmax) X(double, efc_D, m->njmax) X(double, efc_KBIP, 4 * m->njmax)      \
  X(double, efc_vel, m->njmax) X(double, efc_aref, m->njmax) X(double, efc_force, m->njmax)

typedef struct OData {
  double time;
  int ncon, nefc, solver_iter, unsupported_pairs, warn_overflow;
#define X(T, n, c) T* n;
  DATA_FIELDS(X)
#undef X
} OData;

OModel* o_model_new(const int* sz) {
  OModel* m = (OModel*)calloc(1, sizeof(OModel));
  m->nq = sz[0]; m->nv = sz[1]; m->nu = sz[2]; m->na = sz[3]; m->nbody = sz[4]; m->njnt = sz[5];
  m->ngeom = sz[6]; m->nsite = sz[7]; m->ntendon = sz[8]; m->nwrap = sz[9]; m->nM = sz[10];
  m->njmax = sz[11]; m->nconmax = sz[12];
#define X(T, n, c) m->n = (T*)calloc((size_t)((c) > 0 ? (c) : 1), sizeof(T));
  MODEL_FIELDS(X)
#undef X
  return m;
}
void o_model_set_excludes(OModel* m, int n, const int* sig) {
  free(m->exclude_signature);
  m->nexclude = n;
  m->exclude_signature = (int*)calloc((size_t)(n > 0 ? n : 1), sizeof(int));
  for (int i = 0; i < n; i++) m->exclude_signature[i] = sig[i];
}
void o_model_free(OModel* m) {
  free(m->exclude_signature);
#define X(T, n, c) free(m->n);
  MODEL_FIELDS(X)
#undef X
  free(m);
}
void o_model_set_opt(OModel* m, double timestep, const double* gravity, double impratio, int cone,
                     int disableflags, double meaninertia) {
  m->timestep = timestep; memcpy(m->gravity, gravity, 3 * sizeof(double));
  m->impratio = impratio; m->cone = cone; m->disableflags = disableflags;
  m->stat_meaninertia = meaninertia;
}
/* returns pointer to a named model array; *count = elements, *is_int = element kind */
void* o_model_field(OModel* m, const char* name, int* count, int* is_int) {
#define X(T, n, c) if (!strcmp(name, #n)) { *count = (c); *is_int = (sizeof(T) == sizeof(int)); return m->n; }
  MODEL_FIELDS(X)
#undef X
  return NULL;
}
OData* o_data_new(const OModel* m) {
  OData* d = (OData*)calloc(1, sizeof(OData));
#define X(T, n, c) d->n = (T*)calloc((size_t)((c) > 0 ? (c) : 1), sizeof(T));
  DATA_FIELDS(X)
#undef X
  return d;
}
void o_data_free(OData* d) {
#define X(T, n, c) free(d->n);
  DATA_FIELDS(X)
#undef X
  free(d);
}
void* o_data_field(const OModel* m, OData* d, const char* name, int* count, int* is_int) {
#define X(T, n, c) if (!strcmp(name, #n)) { *count = (c); *is_int = (sizeof(T) == sizeof(int)); return d->n; }
  DATA_FIELDS(X)
#undef X
  return NULL;
}
double o_data_time(const OData* d) { return d->time; }
void o_data_set_time(OData* d, double t) { d->time = t; }
int o_data_ncon(const OData* d) { return d->ncon; }
int o_data_nefc(const OData* d) { return d->nefc; }
int o_data_solver_iter(const OData* d) { return d->solver_iter; }
int o_data_unsupported(const OData* d) { return d->unsupported_pairs; }
int o_data_overflow(const OData* d) { return d->warn_overflow; }

/* ------------------------------------------------------------------ small math ----------- */
static void zero(double* a, int n) { memset(a, 0, (size_t)n * sizeof(double)); }
static void cpy(double* a, const double* b, int n) { memcpy(a, b, (size_t)n * sizeof(double)); }
static double dot3(const double* a, const double* b) { return a[0]*b[0] + a[1]*b[1] + a[2]*b[2]; }
static double dotn(const double* a, const double* b, int n) { double s = 0; for (int i = 0; i < n; i++) s += a[i]*b[i]; return s; }
static void cross(double* r, const double* a, const double* b) {
  double x = a[1]*b[2] - a[2]*b[1], y = a[2]*b[0] - a[0]*b[2], z = a[0]*b[1] - a[1]*b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static void sub3(double* r, const double* a, const double* b) { r[0] = a[0]-b[0]; r[1] = a[1]-b[1]; r[2] = a[2]-b[2]; }
static void add3(double* r, const double* a, const double* b) { r[0] = a[0]+b[0]; r[1] = a[1]+b[1]; r[2] = a[2]+b[2]; }
static void addscl3(double* r, const double* a, const double* b, double s) { r[0] = a[0]+s*b[0]; r[1] = a[1]+s*b[1]; r[2] = a[2]+s*b[2]; }
static double norm3(const double* a) { return sqrt(dot3(a, a)); }
static double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static void normalize4(double* q) {
  double n = sqrt(q[0]*q[0] + q[1]*q[1] + q[2]*q[2] + q[3]*q[3]);
  if (n < MINVAL) { q[0] = 1; q[1] = q[2] = q[3] = 0; } else { for (int i = 0; i < 4; i++) q[i] /= n; }
}
static void mulquat(double* r, const double* a, const double* b) {
  double t[4] = { a[0]*b[0] - a[1]*b[1] - a[2]*b[2] - a[3]*b[3],
                  a[0]*b[1] + a[1]*b[0] + a[2]*b[3] - a[3]*b[2],
                  a[0]*b[2] - a[1]*b[3] + a[2]*b[0] + a[3]*b[1],
                  a[0]*b[3] + a[1]*b[2] - a[2]*b[1] + a[3]*b[0] };
  cpy(r, t, 4);
}
static void quat2mat(double* R, const double* q) {
  double q00 = q[0]*q[0], q01 = q[0]*q[1], q02 = q[0]*q[2], q03 = q[0]*q[3], q11 = q[1]*q[1],
         q12 = q[1]*q[2], q13 = q[1]*q[3], q22 = q[2]*q[2], q23 = q[2]*q[3], q33 = q[3]*q[3];
  R[0] = q00 + q11 - q22 - q33; R[4] = q00 - q11 + q22 - q33; R[8] = q00 - q11 - q22 + q33;
  R[1] = 2*(q12 - q03); R[2] = 2*(q13 + q02); R[3] = 2*(q12 + q03);
  R[5] = 2*(q23 - q01); R[6] = 2*(q13 - q02); R[7] = 2*(q23 + q01);
}
static void rotvecquat(double* r, const double* v, const double* q) {
  double R[9]; quat2mat(R, q);
  double t[3] = { R[0]*v[0] + R[1]*v[1] + R[2]*v[2], R[3]*v[0] + R[4]*v[1] + R[5]*v[2],
                  R[6]*v[0] + R[7]*v[1] + R[8]*v[2] };
  cpy(r, t, 3);
}
static void mulmatvec3(double* r, const double* R, const double* v) {
  double t[3] = { R[0]*v[0] + R[1]*v[1] + R[2]*v[2], R[3]*v[0] + R[4]*v[1] + R[5]*v[2],
                  R[6]*v[0] + R[7]*v[1] + R[8]*v[2] };
  cpy(r, t, 3);
}
static void mulmatTvec3(double* r, const double* R, const double* v) {
  double t[3] = { R[0]*v[0] + R[3]*v[1] + R[6]*v[2], R[1]*v[0] + R[4]*v[1] + R[7]*v[2],
                  R[2]*v[0] + R[5]*v[1] + R[8]*v[2] };
  cpy(r, t, 3);
}
static void axisangle2quat(double* q, const double* axis, double angle) {
  if (angle == 0) { q[0] = 1; q[1] = q[2] = q[3] = 0; return; }
  double s = sin(angle * 0.5);
  q[0] = cos(angle * 0.5); q[1] = axis[0]*s; q[2] = axis[1]*s; q[3] = axis[2]*s;
}
static double clip(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* spatial algebra: motion/force vectors are [rotation(3); translation(3)] */
static void cross_motion(double* r, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross(a, vel, v); cross(b, vel, v + 3); cross(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; add3(r + 3, b, c);
}
static void cross_force(double* r, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross(a, vel, f); cross(b, vel + 3, f + 3); cross(c, vel, f + 3);
  add3(r, a, b); r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
/* 10-number inertia about an arbitrary point: Ixx Iyy Izz Ixy Ixz Iyz m*cx m*cy m*cz m */
static void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0]*v[0] + i[3]*v[1] + i[4]*v[2] - i[8]*v[4] + i[7]*v[5];
  r[1] = i[3]*v[0] + i[1]*v[1] + i[5]*v[2] + i[8]*v[3] - i[6]*v[5];
  r[2] = i[4]*v[0] + i[5]*v[1] + i[2]*v[2] - i[7]*v[3] + i[6]*v[4];
  r[3] = i[8]*v[1] - i[7]*v[2] + i[9]*v[3];
  r[4] = i[6]*v[2] - i[8]*v[0] + i[9]*v[4];
  r[5] = i[7]*v[0] - i[6]*v[1] + i[9]*v[5];
}
static void inert_com(double* res, const double* inert, const double* mat, const double* dif, double mass) {
  double t[9];
  for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++)
    t[3*r + c] = mat[3*r]*inert[0]*mat[3*c] + mat[3*r + 1]*inert[1]*mat[3*c + 1] + mat[3*r + 2]*inert[2]*mat[3*c + 2];
  res[0] = t[0] + mass*(dif[1]*dif[1] + dif[2]*dif[2]);
  res[1] = t[4] + mass*(dif[0]*dif[0] + dif[2]*dif[2]);
  res[2] = t[8] + mass*(dif[0]*dif[0] + dif[1]*dif[1]);
  res[3] = t[1] - mass*dif[0]*dif[1];
  res[4] = t[2] - mass*dif[0]*dif[2];
  res[5] = t[5] - mass*dif[1]*dif[2];
  res[6] = mass*dif[0]; res[7] = mass*dif[1]; res[8] = mass*dif[2]; res[9] = mass;
}

/* ------------------------------------------------------------------ kinematics ----------- */
/* mj_kinematics (engine_core_smooth.c) */
void o_kinematics(const OModel* m, OData* d) {
  d->xpos[0] = d->xpos[1] = d->xpos[2] = 0;
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  cpy(d->xipos, d->xpos, 3); cpy(d->ximat, d->xmat, 9);
  for (int i = 1; i < m->nbody; i++) {
    double* xpos = d->xpos + 3*i; double* xquat = d->xquat + 4*i;
    int jadr = m->body_jntadr[i], jnum = m->body_jntnum[i];
    if (jnum == 1 && m->jnt_type[jadr] == JNT_FREE) {
      int qa = m->jnt_qposadr[jadr];
      cpy(xpos, d->qpos + qa, 3); cpy(xquat, d->qpos + qa + 3, 4); normalize4(xquat);
      cpy(d->xanchor + 3*jadr, xpos, 3); cpy(d->xaxis + 3*jadr, m->jnt_axis + 3*jadr, 3);
    } else {
      int pid = m->body_parentid[i];
      double v[3];
      mulmatvec3(v, d->xmat + 9*pid, m->body_pos + 3*i); add3(xpos, v, d->xpos + 3*pid);
      mulquat(xquat, d->xquat + 4*pid, m->body_quat + 4*i);
      for (int j = jadr; j < jadr + jnum; j++) {
        int qa = m->jnt_qposadr[j];
        double* xaxis = d->xaxis + 3*j; double* xanchor = d->xanchor + 3*j;
        rotvecquat(xaxis, m->jnt_axis + 3*j, xquat);
        rotvecquat(xanchor, m->jnt_pos + 3*j, xquat); add3(xanchor, xanchor, xpos);
        if (m->jnt_type[j] == JNT_SLIDE) {
          addscl3(xpos, xpos, xaxis, d->qpos[qa] - m->qpos0[qa]);
        } else if (m->jnt_type[j] == JNT_HINGE) {
          double ql[4], vec[3];
          axisangle2quat(ql, m->jnt_axis + 3*j, d->qpos[qa] - m->qpos0[qa]);
          mulquat(xquat, xquat, ql);
          rotvecquat(vec, m->jnt_pos + 3*j, xquat); sub3(xpos, xanchor, vec);
        } else if (m->jnt_type[j] == JNT_BALL) {
          double ql[4], vec[3];
          cpy(ql, d->qpos + qa, 4); normalize4(ql); mulquat(xquat, xquat, ql);
          rotvecquat(vec, m->jnt_pos + 3*j, xquat); sub3(xpos, xanchor, vec);
        }
      }
    }
    normalize4(xquat);
    quat2mat(d->xmat + 9*i, xquat);
    double v[3], q[4];
    mulmatvec3(v, d->xmat + 9*i, m->body_ipos + 3*i); add3(d->xipos + 3*i, v, xpos);
    mulquat(q, xquat, m->body_iquat + 4*i); quat2mat(d->ximat + 9*i, q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; double v[3], q[4];
    mulmatvec3(v, d->xmat + 9*b, m->geom_pos + 3*g); add3(d->geom_xpos + 3*g, v, d->xpos + 3*b);
    mulquat(q, d->xquat + 4*b, m->geom_quat + 4*g); quat2mat(d->geom_xmat + 9*g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; double v[3], q[4];
    mulmatvec3(v, d->xmat + 9*b, m->site_pos + 3*s); add3(d->site_xpos + 3*s, v, d->xpos + 3*b);
    mulquat(q, d->xquat + 4*b, m->site_quat + 4*s); quat2mat(d->site_xmat + 9*s, q);
  }
}

/* mj_comPos: subtree COM (normalised by the *stored* body_subtreemass, which goes stale when
 * body_mass is randomised without mj_setConst -- /root/reference/src/envs/baoding.py:560-566),
 * com-based inertias and motion axes. */
void o_com_pos(const OModel* m, OData* d) {
  zero(d->subtree_com, 3 * m->nbody);
  for (int i = 0; i < m->nbody; i++) addscl3(d->subtree_com + 3*i, d->subtree_com + 3*i, d->xipos + 3*i, m->body_mass[i]);
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    add3(d->subtree_com + 3*p, d->subtree_com + 3*p, d->subtree_com + 3*i);
  }
  for (int i = 0; i < m->nbody; i++) {
    if (m->body_subtreemass[i] < MINVAL) cpy(d->subtree_com + 3*i, d->xipos + 3*i, 3);
    else for (int k = 0; k < 3; k++) d->subtree_com[3*i + k] /= m->body_subtreemass[i];
  }
  zero(d->cinert, 10);
  for (int i = 1; i < m->nbody; i++) {
    double off[3];
    sub3(off, d->xipos + 3*i, d->subtree_com + 3*m->body_rootid[i]);
    inert_com(d->cinert + 10*i, m->body_inertia + 3*i, d->ximat + 9*i, off, m->body_mass[i]);
  }
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j], b = m->jnt_bodyid[j];
    double off[3];
    sub3(off, d->subtree_com + 3*m->body_rootid[b], d->xanchor + 3*j);
    int skip = 0;
    switch (m->jnt_type[j]) {
      case JNT_FREE:
        for (int k = 0; k < 3; k++) { zero(d->cdof + 6*(da + k), 6); d->cdof[6*(da + k) + 3 + k] = 1; }
        skip = 3; /* fallthrough */
      case JNT_BALL:
        for (int k = 0; k < 3; k++) {
          double ax[3] = { d->xmat[9*b + k], d->xmat[9*b + 3 + k], d->xmat[9*b + 6 + k] };
          double* c = d->cdof + 6*(da + k + skip);
          cpy(c, ax, 3); cross(c + 3, ax, off);
        }
        break;
      case JNT_SLIDE:
        zero(d->cdof + 6*da, 3); cpy(d->cdof + 6*da + 3, d->xaxis + 3*j, 3);
        break;
      case JNT_HINGE:
        cpy(d->cdof + 6*da, d->xaxis + 3*j, 3); cross(d->cdof + 6*da + 3, d->xaxis + 3*j, off);
        break;
    }
  }
}

/* mj_jac: translational (jacp 3 x nv) and rotational (jacr 3 x nv) Jacobian of a world point
 * rigidly attached to `body`. Either output may be NULL. */
static void o_jac(const OModel* m, const OData* d, double* jacp, double* jacr, const double* point, int body) {
  int nv = m->nv;
  if (jacp) zero(jacp, 3 * nv);
  if (jacr) zero(jacr, 3 * nv);
  double off[3];
  sub3(off, point, d->subtree_com + 3*m->body_rootid[body]);
  while (body && !m->body_dofnum[body]) body = m->body_parentid[body];
  if (!body) return;
  int i = m->body_dofadr[body] + m->body_dofnum[body] - 1;
  while (i >= 0) {
    const double* c = d->cdof + 6*i;
    if (jacr) { jacr[i] = c[0]; jacr[nv + i] = c[1]; jacr[2*nv + i] = c[2]; }
    if (jacp) {
      double t[3]; cross(t, c, off);
      jacp[i] = c[3] + t[0]; jacp[nv + i] = c[4] + t[1]; jacp[2*nv + i] = c[5] + t[2];
    }
    i = m->dof_parentid[i];
  }
}

/* ------------------------------------------------------------------ tendon wrapping ------ */
static int is_intersect(const double* p1, const double* p2, const double* p3, const double* p4) {
  double det = (p4[1]-p3[1])*(p2[0]-p1[0]) - (p4[0]-p3[0])*(p2[1]-p1[1]);
  if (fabs(det) < MINVAL) return 0;
  double a = ((p4[0]-p3[0])*(p1[1]-p3[1]) - (p4[1]-p3[1])*(p1[0]-p3[0])) / det;
  double b = ((p2[0]-p1[0])*(p1[1]-p3[1]) - (p2[1]-p1[1])*(p1[0]-p3[0])) / det;
  return (a >= 0 && a <= 1 && b >= 0 && b <= 1);
}
static double length_circle(const double* p0, const double* p1, int ind, double rad) {
  double n0 = sqrt(p0[0]*p0[0] + p0[1]*p0[1]), n1 = sqrt(p1[0]*p1[0] + p1[1]*p1[1]);
  double c = (p0[0]*p1[0] + p0[1]*p1[1]) / (n0 * n1);
  double angle = acos(clip(c, -1, 1));
  double cr = p0[1]*p1[0] - p0[0]*p1[1];
  if ((cr > 0 && ind) || (cr < 0 && !ind)) angle = 2*PI - angle;
  return rad * angle;
}
/* 2-D wrap around a circle: d = (x0,y0,x1,y1) end points, sd = optional side point */
static double wrap_circle(double* pnt, const double* d, const double* sd, double rad) {
  double sqlen0 = d[0]*d[0] + d[1]*d[1], sqlen1 = d[2]*d[2] + d[3]*d[3], sqrad = rad*rad;
  double dif[2] = { d[2]-d[0], d[3]-d[1] };
  double dd = dif[0]*dif[0] + dif[1]*dif[1];
  if (sqlen0 < sqrad || sqlen1 < sqrad || rad < MINVAL) return -1;
  if (dd < MINVAL) return -1;
  double a = -(dif[0]*d[0] + dif[1]*d[1]) / dd;
  a = clip(a, 0, 1);
  double tmp[2] = { a*dif[0] + d[0], a*dif[1] + d[1] };
  if (tmp[0]*tmp[0] + tmp[1]*tmp[1] > sqrad && (!sd || sd[0]*tmp[0] + sd[1]*tmp[1] >= 0)) return -1;
  double sol[2][4], good[2];
  double sqrt0 = sqrt(sqlen0 - sqrad), sqrt1 = sqrt(sqlen1 - sqrad);
  for (int i = 0; i < 2; i++) {
    int sgn = (i == 0 ? 1 : -1);
    sol[i][0] = (d[0]*sqrad + sgn*rad*d[1]*sqrt0) / sqlen0;
    sol[i][1] = (d[1]*sqrad - sgn*rad*d[0]*sqrt0) / sqlen0;
    sol[i][2] = (d[2]*sqrad - sgn*rad*d[3]*sqrt1) / sqlen1;
    sol[i][3] = (d[3]*sqrad + sgn*rad*d[2]*sqrt1) / sqlen1;
    if (sd) {
      double t[2] = { sol[i][0] + sol[i][2], sol[i][1] + sol[i][3] };
      double n = sqrt(t[0]*t[0] + t[1]*t[1]);
      if (n < MINVAL) { t[0] = 1; t[1] = 0; } else { t[0] /= n; t[1] /= n; }
      good[i] = t[0]*sd[0] + t[1]*sd[1];
    } else {
      double t[2] = { sol[i][0] - sol[i][2], sol[i][1] - sol[i][3] };
      good[i] = -(t[0]*t[0] + t[1]*t[1]);
    }
    if (is_intersect(d, sol[i], d + 2, sol[i] + 2)) good[i] = -10000;
  }
  int i = (good[0] > good[1] ? 0 : 1);
  cpy(pnt, sol[i], 4);
  if (is_intersect(d, pnt, d + 2, pnt + 2)) return -1;
  return length_circle(sol[i], sol[i] + 2, i, rad);
}
/* mju_wrap (engine_util_misc.c). Inside-wrap (side site inside the geom) is not restated: none of
 * the reference's shipped models use it; returns -2 so the caller can flag it. */
static double o_wrap(double* wpnt, const double* x0, const double* x1, const double* xpos, const double* xmat,
                     double radius, int type, const double* side) {
  double p0[3], p1[3], dif[3], axis0[3], axis1[3], normal[3];
  sub3(dif, x0, xpos); mulmatTvec3(p0, xmat, dif);
  sub3(dif, x1, xpos); mulmatTvec3(p1, xmat, dif);
  if (norm3(p0) < MINVAL || norm3(p1) < MINVAL) return -1;
  if (type == WRAP_SPHERE) {
    cpy(axis0, p0, 3); normalize3(axis0);
    cross(normal, p0, p1);
    double nrm = norm3(normal);
    if (nrm < MINVAL) {
      int im = 0;
      for (int k = 1; k < 3; k++) if (fabs(axis0[k]) > fabs(axis0[im])) im = k;
      double a1[3] = { 1, 1, 1 }; a1[im] = 0;
      cross(normal, axis0, a1);
    }
    normalize3(normal);
    cross(axis1, normal, axis0); normalize3(axis1);
  } else {
    axis0[0] = 1; axis0[1] = 0; axis0[2] = 0; axis1[0] = 0; axis1[1] = 1; axis1[2] = 0;
  }
  double dd[4] = { dot3(p0, axis0), dot3(p0, axis1), dot3(p1, axis0), dot3(p1, axis1) };
  double sd[2]; const double* sdp = NULL;
  if (side) {
    double s[3];
    sub3(dif, side, xpos); mulmatTvec3(s, xmat, dif);
    double inside_norm = (type == WRAP_SPHERE) ? norm3(s) : sqrt(s[0]*s[0] + s[1]*s[1]);
    if (inside_norm < radius) return -2;   /* inside wrap: unsupported */
    sd[0] = dot3(s, axis0); sd[1] = dot3(s, axis1);
    double n = sqrt(sd[0]*sd[0] + sd[1]*sd[1]);
    if (n < MINVAL) { sd[0] = 1; sd[1] = 0; } else { sd[0] /= n; sd[1] /= n; }
    sd[0] *= radius; sd[1] *= radius;
    sdp = sd;
  }
  double pnt[4];
  double wlen = wrap_circle(pnt, dd, sdp, radius);
  if (wlen < 0) return -1;
  double res[6];
  for (int k = 0; k < 3; k++) {
    res[k] = axis0[k]*pnt[0] + axis1[k]*pnt[1];
    res[3 + k] = axis0[k]*pnt[2] + axis1[k]*pnt[3];
  }
  if (type == WRAP_CYLINDER) {
    double L0 = sqrt((p0[0]-res[0])*(p0[0]-res[0]) + (p0[1]-res[1])*(p0[1]-res[1]));
    double L1 = sqrt((p1[0]-res[3])*(p1[0]-res[3]) + (p1[1]-res[4])*(p1[1]-res[4]));
    res[2] = p0[2] + (p1[2] - p0[2]) * L0 / (L0 + wlen + L1);
    res[5] = p0[2] + (p1[2] - p0[2]) * (L0 + wlen) / (L0 + wlen + L1);
    double h = fabs(res[5] - res[2]);
    wlen = sqrt(wlen*wlen + h*h);
  }
  mulmatvec3(wpnt, xmat, res); add3(wpnt, wpnt, xpos);
  mulmatvec3(wpnt + 3, xmat, res + 3); add3(wpnt + 3, wpnt + 3, xpos);
  return wlen;
}

/* mj_tendon: spatial tendons (site / sphere / cylinder / pulley) and fixed (joint) tendons */
void o_tendon(const OModel* m, OData* d) {
  int nv = m->nv;
  double* jac1 = (double*)malloc(sizeof(double) * 3 * nv);
  double* jac2 = (double*)malloc(sizeof(double) * 3 * nv);
  zero(d->ten_length, m->ntendon); zero(d->ten_J, m->ntendon * nv);
  for (int i = 0; i < m->ntendon; i++) {
    int adr = m->tendon_adr[i], num = m->tendon_num[i];
    double* L = d->ten_length + i; double* J = d->ten_J + i*nv;
    if (m->wrap_type[adr] == WRAP_JOINT) {
      for (int j = 0; j < num; j++) {
        int k = m->wrap_objid[adr + j];
        *L += m->wrap_prm[adr + j] * d->qpos[m->jnt_qposadr[k]];
        J[m->jnt_dofadr[k]] = m->wrap_prm[adr + j];
      }
      continue;
    }
    double divisor = 1;
    int j = 0;
    while (j < num - 1) {
      int tp0 = m->wrap_type[adr + j], tp1 = m->wrap_type[adr + j + 1];
      int id0 = m->wrap_objid[adr + j], id1 = m->wrap_objid[adr + j + 1];
      if (tp0 == WRAP_PULLEY || tp1 == WRAP_PULLEY) {
        if (tp0 == WRAP_PULLEY) divisor = m->wrap_prm[adr + j];
        j++; continue;
      }
      double wpnt[12], wlen = -1; int wbody[4];
      cpy(wpnt, d->site_xpos + 3*id0, 3); wbody[0] = m->site_bodyid[id0];
      int isgeom = (tp1 == WRAP_SPHERE || tp1 == WRAP_CYLINDER);
      int idw = -1;
      if (isgeom) {
        idw = id1;
        int tpw = tp1;
        id1 = m->wrap_objid[adr + j + 2];
        int side = (int)lround(m->wrap_prm[adr + j + 1]);
        wlen = o_wrap(wpnt + 3, d->site_xpos + 3*id0, d->site_xpos + 3*id1, d->geom_xpos + 3*idw,
                      d->geom_xmat + 9*idw, m->geom_size[3*idw], tpw,
                      (side >= 0 && side < m->nsite) ? d->site_xpos + 3*side : NULL);
        if (wlen == -2) { d->unsupported_pairs++; wlen = -1; }
      }
      int nseg;
      if (wlen < 0) {
        cpy(wpnt + 3, d->site_xpos + 3*id1, 3); wbody[1] = m->site_bodyid[id1];
        double dif[3]; sub3(dif, wpnt + 3, wpnt);
        *L += norm3(dif) / divisor;
        nseg = 1;
      } else {
        cpy(wpnt + 9, d->site_xpos + 3*id1, 3);
        wbody[1] = wbody[2] = m->geom_bodyid[idw]; wbody[3] = m->site_bodyid[id1];
        double a[3], b[3]; sub3(a, wpnt + 3, wpnt); sub3(b, wpnt + 9, wpnt + 6);
        *L += (norm3(a) + wlen + norm3(b)) / divisor;
        nseg = 3;
      }
      for (int k = 0; k < nseg; k++) {
        if (wbody[k] == wbody[k + 1]) continue;
        double dif[3]; sub3(dif, wpnt + 3*k + 3, wpnt + 3*k); normalize3(dif);
        o_jac(m, d, jac1, NULL, wpnt + 3*k, wbody[k]);
        o_jac(m, d, jac2, NULL, wpnt + 3*k + 3, wbody[k + 1]);
        for (int c = 0; c < nv; c++) {
          double s = 0;
          for (int r = 0; r < 3; r++) s += (jac2[r*nv + c] - jac1[r*nv + c]) * dif[r];
          J[c] += s / divisor;
        }
      }
      j += (isgeom ? 2 : 1);
    }
  }
  free(jac1); free(jac2);
}

/* mj_transmission: joint and tendon transmissions, scalar gear */
void o_transmission(const OModel* m, OData* d) {
  int nv = m->nv;
  zero(d->actuator_moment, m->nu * nv);
  for (int i = 0; i < m->nu; i++) {
    double gear = m->actuator_gear[6*i];
    int id = m->actuator_trnid[2*i];
    if (m->actuator_trntype[i] == TRN_TENDON) {
      d->actuator_length[i] = gear * d->ten_length[id];
      for (int c = 0; c < nv; c++) d->actuator_moment[i*nv + c] = gear * d->ten_J[id*nv + c];
    } else { /* TRN_JOINT, hinge/slide only */
      d->actuator_length[i] = gear * d->qpos[m->jnt_qposadr[id]];
      d->actuator_moment[i*nv + m->jnt_dofadr[id]] = gear;
    }
  }
}

/* ------------------------------------------------------------------ inertia -------------- */
/* mj_crb: composite rigid body; sparse qM in MuJoCo's dof_Madr layout and a dense copy.
 * "Simple" dofs take the stored dof_M0 (so a randomised body_mass does not reach M). */
void o_crb(const OModel* m, OData* d) {
  int nv = m->nv;
  cpy(d->crb, d->cinert, 10 * m->nbody);
  for (int i = m->nbody - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int k = 0; k < 10; k++) d->crb[10*p + k] += d->crb[10*i + k];
  }
  zero(d->qM, m->nM); zero(d->Mdense, nv * nv);
  for (int i = 0; i < nv; i++) {
    if (m->dof_simplenum[i]) {
      d->qM[m->dof_Madr[i]] = m->dof_M0[i];
      d->Mdense[i*nv + i] = m->dof_M0[i];
      continue;
    }
    int adr = m->dof_Madr[i];
    double buf[6];
    mul_inert_vec(buf, d->crb + 10*m->dof_bodyid[i], d->cdof + 6*i);
    d->qM[adr] = m->dof_armature[i];
    int j = i;
    while (j >= 0) {
      d->qM[adr] += dotn(d->cdof + 6*j, buf, 6);
      d->Mdense[i*nv + j] = d->Mdense[j*nv + i] = d->qM[adr];
      adr++;
      j = m->dof_parentid[j];
    }
  }
}
/* dense Cholesky A = L L^T (lower), returns 0 on success */
static int chol_factor(double* L, const double* A, int n) {
  for (int i = 0; i < n; i++) for (int j = 0; j <= i; j++) {
    double s = A[i*n + j];
    for (int k = 0; k < j; k++) s -= L[i*n + k] * L[j*n + k];
    if (i == j) { if (s < MINVAL) s = MINVAL; L[i*n + i] = sqrt(s); }
    else L[i*n + j] = s / L[j*n + j];
  }
  return 0;
}
static void chol_solve(const double* L, double* x, int n) {
  for (int i = 0; i < n; i++) { double s = x[i]; for (int k = 0; k < i; k++) s -= L[i*n + k]*x[k]; x[i] = s / L[i*n + i]; }
  for (int i = n - 1; i >= 0; i--) { double s = x[i]; for (int k = i + 1; k < n; k++) s -= L[k*n + i]*x[k]; x[i] = s / L[i*n + i]; }
}
void o_factor_m(const OModel* m, OData* d) { chol_factor(d->Lchol, d->Mdense, m->nv); }
static void o_solve_m(const OModel* m, const OData* d, double* x) { chol_solve(d->Lchol, x, m->nv); }

/* ------------------------------------------------------------------ collision ------------ */
static void make_frame(double* f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) { f[3] = f[4] = f[5] = 0; if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1; }
  double t = dot3(f, f + 3);
  addscl3(f + 3, f + 3, f, -t); normalize3(f + 3);
  cross(f + 6, f, f + 3);
}
/* returns 1 when a contact (dist < margin) is written to dist/pos/frame[0:3] */
static int raw_sphere_sphere(double* dist, double* pos, double* frame, double margin, const double* pos1, double r1,
                             const double* pos2, double r2) {
  double dif[3]; sub3(dif, pos2, pos1);
  double cd2 = dot3(dif, dif), mind = margin + r1 + r2;
  if (cd2 > mind*mind) return 0;
  double n = normalize3(dif);
  *dist = n - r1 - r2;
  cpy(frame, dif, 3);
  addscl3(pos, pos1, dif, r1 + 0.5 * (*dist));
  return 1;
}
/* ---- capsule - box. NOT a restatement of MuJoCo's mjc_CapsuleBox (that source is not in the container and its case
 * analysis cannot be reproduced from memory): an own closest-point collider, the same algorithm in the kernel
 * (csrc/myo_phys.cuh capsule_box). The capsule is its core segment p(t) = c + t d, |t| <= h, in the box frame; the
 * squared distance to the box f(t) = sum_i max(0, |c_i + t d_i| - s_i)^2 is convex and piecewise quadratic, so f' is
 * monotone and piecewise linear with breakpoints where a coordinate crosses a face: the minimiser is found exactly
 * by walking the sorted breakpoints. Contacts: if both end points of the segment are within reach of the box
 * (capsule lying along a face) one contact per end point, else one contact at the closest point. */
static double box_point_dist(const double* p, const double* s, double* q) {      /* q: closest point of the box to p */
  double f = 0;
  for (int i = 0; i < 3; i++) { q[i] = clip(p[i], -s[i], s[i]); f += (p[i] - q[i]) * (p[i] - q[i]); }
  return sqrt(f);
}
static double seg_box_dfdt(const double* c, const double* d, const double* s, double t) {
  double g = 0;
  for (int i = 0; i < 3; i++) {
    double x = c[i] + t * d[i], e = fabs(x) - s[i];
    if (e > 0) g += 2 * e * (x > 0 ? d[i] : -d[i]);
  }
  return g;
}
static double seg_box_closest_t(const double* c, const double* d, double h, const double* s) {
  double bp[8]; int nb = 0;
  bp[nb++] = -h;
  for (int i = 0; i < 3; i++) if (fabs(d[i]) > 1e-12) for (int sg = -1; sg <= 1; sg += 2) {
    double t = (sg * s[i] - c[i]) / d[i];
    if (t > -h && t < h) bp[nb++] = t;
  }
  bp[nb++] = h;
  for (int i = 1; i < nb; i++) { double v = bp[i]; int j = i - 1; while (j >= 0 && bp[j] > v) { bp[j + 1] = bp[j]; j--; } bp[j + 1] = v; }
  /* f' at every breakpoint (monotone): last negative, first positive, exact zeros in between form a flat stretch */
  double g[8]; int lo = -1, hi = nb;
  for (int k = 0; k < nb; k++) g[k] = seg_box_dfdt(c, d, s, bp[k]);
  for (int k = 0; k < nb; k++) if (g[k] < 0) lo = k;
  for (int k = nb - 1; k >= 0; k--) if (g[k] > 0) hi = k;
  if (hi == 0) return bp[0];
  if (lo == nb - 1) return bp[nb - 1];
  if (hi == lo + 1) return bp[lo] + (bp[hi] - bp[lo]) * (-g[lo]) / (g[hi] - g[lo]);
  return 0.5 * (bp[lo + 1] + bp[hi - 1]);
}
/* one contact between the capsule's point p (box frame) and the box; returns 0 when out of reach */
static int capsule_box_point(const double* p, const double* s, double r, double margin, double* dist, double* pos_b, double* nrm_b) {
  double q[3];
  double dd = box_point_dist(p, s, q);
  if (dd > r + margin) return 0;
  if (dd > 1e-10) {      /* outside: normal from the capsule toward the box */
    for (int i = 0; i < 3; i++) nrm_b[i] = (q[i] - p[i]) / dd;
    *dist = dd - r;
  } else {               /* the core segment is inside the box: leave through the nearest face */
    int ax = 0; double best = 1e300;
    for (int i = 0; i < 3; i++) { double dep = s[i] - fabs(p[i]); if (dep < best) { best = dep; ax = i; } }
    nrm_b[0] = nrm_b[1] = nrm_b[2] = 0;
    nrm_b[ax] = p[ax] > 0 ? -1 : 1;
    *dist = -best - r;
  }
  for (int i = 0; i < 3; i++) pos_b[i] = p[i] + nrm_b[i] * (r + 0.5 * (*dist));
  return 1;
}
/* capsule (pos pc, axis = column 2 of Rc, radius r, half length h) vs box (pos pb, frame Rb, half sizes s):
 * up to two contacts (dist, world pos, world normal capsule -> box) */
static int capsule_box(double margin, const double* pc, const double* Rc, double r, double h, const double* pb, const double* Rb,
                       const double* s, double* dist, double* pos, double* nrm) {
  double ax[3] = { Rc[2], Rc[5], Rc[8] }, dif[3], c[3], d[3];
  sub3(dif, pc, pb); mulmatTvec3(c, Rb, dif); mulmatTvec3(d, Rb, ax);
  double pe[2][3], de[2], pb2[2][3], nb2[2][3]; int he[2];
  for (int e = 0; e < 2; e++) {
    double t = e ? h : -h;
    for (int i = 0; i < 3; i++) pe[e][i] = c[i] + t * d[i];
    he[e] = capsule_box_point(pe[e], s, r, margin, &de[e], pb2[e], nb2[e]);
  }
  int n = 0;
  if (he[0] && he[1]) {
    for (int e = 0; e < 2; e++) { dist[n] = de[e]; mulmatvec3(pos + 3*n, Rb, pb2[e]); add3(pos + 3*n, pos + 3*n, pb); mulmatvec3(nrm + 3*n, Rb, nb2[e]); n++; }
    return n;
  }
  double t = seg_box_closest_t(c, d, h, s), p[3], posb[3], nrmb[3];
  for (int i = 0; i < 3; i++) p[i] = c[i] + t * d[i];
  if (capsule_box_point(p, s, r, margin, &dist[0], posb, nrmb)) {
    mulmatvec3(pos, Rb, posb); add3(pos, pos, pb); mulmatvec3(nrm, Rb, nrmb);
    n = 1;
  }
  return n;
}

/* returns the number of candidate contacts written (dist / pos / frame normal per contact), -1 for an unsupported pair */
static int collide_pair(const OModel* m, const OData* d, int g1, int g2, double margin, double* dist, double* pos, double* frame) {
  int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
  const double *p1 = d->geom_xpos + 3*g1, *p2 = d->geom_xpos + 3*g2;
  const double *R1 = d->geom_xmat + 9*g1, *R2 = d->geom_xmat + 9*g2;
  const double *s1 = m->geom_size + 3*g1, *s2 = m->geom_size + 3*g2;
  zero(frame, 18);
  if (t1 == GEOM_SPHERE && t2 == GEOM_SPHERE) return raw_sphere_sphere(dist, pos, frame, margin, p1, s1[0], p2, s2[0]);
  if (t1 == GEOM_SPHERE && t2 == GEOM_CAPSULE) {
    double axis[3] = { R2[2], R2[5], R2[8] }, vec[3];
    sub3(vec, p1, p2);
    double x = clip(dot3(axis, vec), -s2[1], s2[1]);
    addscl3(vec, p2, axis, x);
    return raw_sphere_sphere(dist, pos, frame, margin, p1, s1[0], vec, s2[0]);
  }
  if (t1 == GEOM_PLANE && t2 == GEOM_SPHERE) {
    double nrm[3] = { R1[2], R1[5], R1[8] }, dif[3];
    sub3(dif, p2, p1);
    double cd = dot3(dif, nrm);
    if (cd > margin + s2[0]) return 0;
    *dist = cd - s2[0];
    cpy(frame, nrm, 3);
    addscl3(pos, p2, nrm, -(*dist)/2 - s2[0]);
    return 1;
  }
  if (t1 == GEOM_CAPSULE && t2 == GEOM_BOX) {
    double nrm[6];
    int n = capsule_box(margin, p1, R1, s1[0], s1[1], p2, R2, s2, dist, pos, nrm);
    for (int k = 0; k < n; k++) cpy(frame + 9*k, nrm + 3*k, 3);
    return n;
  }
  if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) return -1;   /* never in range for shipped models */
  return -1; /* unsupported pair type */
}
static double mixf(double a, double b, double mix) { return mix*a + (1 - mix)*b; }

/* mj_collision: all body pairs b1<b2 (world geoms included), geoms of b1 x geoms of b2 */
void o_collision(const OModel* m, OData* d) {
  d->ncon = 0;
  for (int b1 = 0; b1 < m->nbody; b1++) for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
    if (!m->body_geomnum[b1] || !m->body_geomnum[b2]) continue;
    int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
    int wp1 = m->body_weldid[m->body_parentid[w1]], wp2 = m->body_weldid[m->body_parentid[w2]];
    if (w1 == w2) continue;
    if (w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
    { int ex = 0; for (int e = 0; e < m->nexclude; e++) if (m->exclude_signature[e] == (b1 << 16) + b2) ex = 1; if (ex) continue; }
    for (int ga = m->body_geomadr[b1]; ga < m->body_geomadr[b1] + m->body_geomnum[b1]; ga++)
      for (int gb = m->body_geomadr[b2]; gb < m->body_geomadr[b2] + m->body_geomnum[b2]; gb++) {
        int g1 = ga, g2 = gb;
        if (m->geom_type[g1] > m->geom_type[g2]) { int t = g1; g1 = g2; g2 = t; }
        if (!((m->geom_contype[g1] & m->geom_conaffinity[g2]) || (m->geom_contype[g2] & m->geom_conaffinity[g1]))) continue;
        double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
        double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
        double rb1 = m->geom_rbound[g1], rb2 = m->geom_rbound[g2];
        if (rb1 > 0 && rb2 > 0) {
          double dif[3]; sub3(dif, d->geom_xpos + 3*g1, d->geom_xpos + 3*g2);
          double bound = rb1 + rb2 + margin;
          if (dot3(dif, dif) > bound*bound) continue;
        } else if (m->geom_type[g1] == GEOM_PLANE && rb2 > 0) {
          const double* R = d->geom_xmat + 9*g1;
          double nrm[3] = { R[2], R[5], R[8] }, dif[3];
          sub3(dif, d->geom_xpos + 3*g2, d->geom_xpos + 3*g1);
          if (dot3(dif, nrm) > margin + rb2) continue;
        }
        double dists[2], poss[6], frames[18];
        int r = collide_pair(m, d, g1, g2, margin, dists, poss, frames);
        if (r < 0) { d->unsupported_pairs++; continue; }
        for (int kc = 0; kc < r; kc++) {
        double dist = dists[kc]; double* pos = poss + 3*kc; double* frame = frames + 9*kc;
        if (dist >= margin) continue;
        if (d->ncon >= m->nconmax) { d->warn_overflow++; continue; }
        int c = d->ncon++;
        make_frame(frame);
        d->contact_geom1[c] = g1; d->contact_geom2[c] = g2; d->contact_dist[c] = dist;
        cpy(d->contact_pos + 3*c, pos, 3); cpy(d->contact_frame + 9*c, frame, 9);
        d->contact_includemargin[c] = margin - gap;
        /* mj_contactParam */
        double fri[3]; double* sr = d->contact_solref + 2*c; double* si = d->contact_solimp + 5*c;
        int p1 = m->geom_priority[g1], p2 = m->geom_priority[g2];
        if (p1 != p2) {
          int g = (p1 > p2) ? g1 : g2;
          d->contact_dim[c] = m->geom_condim[g];
          cpy(sr, m->geom_solref + 2*g, 2); cpy(si, m->geom_solimp + 5*g, 5); cpy(fri, m->geom_friction + 3*g, 3);
        } else {
          d->contact_dim[c] = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
          double mix, sm1 = m->geom_solmix[g1], sm2 = m->geom_solmix[g2];
          if (sm1 >= MINVAL && sm2 >= MINVAL) mix = sm1 / (sm1 + sm2);
          else if (sm1 < MINVAL && sm2 < MINVAL) mix = 0.5;
          else if (sm1 < MINVAL) mix = 0; else mix = 1;
          const double *r1 = m->geom_solref + 2*g1, *r2 = m->geom_solref + 2*g2;
          if (r1[0] > 0 && r2[0] > 0) { sr[0] = mixf(r1[0], r2[0], mix); sr[1] = mixf(r1[1], r2[1], mix); }
          else { sr[0] = fmin(r1[0], r2[0]); sr[1] = fmin(r1[1], r2[1]); }
          for (int k = 0; k < 5; k++) si[k] = mixf(m->geom_solimp[5*g1 + k], m->geom_solimp[5*g2 + k], mix);
          for (int k = 0; k < 3; k++) fri[k] = fmax(m->geom_friction[3*g1 + k], m->geom_friction[3*g2 + k]);
        }
        double* f5 = d->contact_friction + 5*c;
        f5[0] = fri[0]; f5[1] = fri[0]; f5[2] = fri[1]; f5[3] = fri[2]; f5[4] = fri[2];
        }
      }
  }
}

/* ------------------------------------------------------------------ constraints ---------- */
static int add_row(const OModel* m, OData* d, const double* jac, double pos, double margin, int type, int id) {
  if (d->nefc >= m->njmax) { d->warn_overflow++; return -1; }
  int r = d->nefc++;
  cpy(d->efc_J + r*m->nv, jac, m->nv);
  d->efc_pos[r] = pos; d->efc_margin[r] = margin; d->efc_type[r] = type; d->efc_id[r] = id;
  return r;
}
/* mj_makeConstraint for the row types the shipped models can produce: joint limits, tendon
 * limits, frictionless and pyramidal (condim 3) contacts. */
void o_make_constraint(const OModel* m, OData* d) {
  int nv = m->nv;
  d->nefc = 0;
  double* jac = (double*)calloc((size_t)nv, sizeof(double));
  for (int i = 0; i < m->njnt; i++) {
    if (!m->jnt_limited[i]) continue;
    int t = m->jnt_type[i];
    if (t != JNT_SLIDE && t != JNT_HINGE) continue;
    double value = d->qpos[m->jnt_qposadr[i]], margin = m->jnt_margin[i];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2*i + (side + 1)/2] - value);
      if (dist < margin) {
        zero(jac, nv); jac[m->jnt_dofadr[i]] = -(double)side;
        add_row(m, d, jac, dist, margin, CNSTR_LIMIT_JOINT, i);
      }
    }
  }
  for (int i = 0; i < m->ntendon; i++) {
    if (!m->tendon_limited[i]) continue;
    double value = d->ten_length[i], margin = m->tendon_margin[i];
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->tendon_range[2*i + (side + 1)/2] - value);
      if (dist < margin) {
        for (int c = 0; c < nv; c++) jac[c] = -side * d->ten_J[i*nv + c];
        add_row(m, d, jac, dist, margin, CNSTR_LIMIT_TENDON, i);
      }
    }
  }
  double* j1 = (double*)malloc(sizeof(double) * 3 * nv);
  double* j2 = (double*)malloc(sizeof(double) * 3 * nv);
  double* jc = (double*)malloc(sizeof(double) * 3 * nv);
  for (int c = 0; c < d->ncon; c++) {
    int b1 = m->geom_bodyid[d->contact_geom1[c]], b2 = m->geom_bodyid[d->contact_geom2[c]];
    const double* fr = d->contact_frame + 9*c;
    double margin = d->contact_includemargin[c], pos = d->contact_dist[c];
    d->contact_efc_address[c] = -1;
    if (pos >= margin) continue;   /* excluded by gap */
    o_jac(m, d, j1, NULL, d->contact_pos + 3*c, b1);
    o_jac(m, d, j2, NULL, d->contact_pos + 3*c, b2);
    for (int k = 0; k < 3; k++) for (int q = 0; q < nv; q++) {
      double s = 0;
      for (int r = 0; r < 3; r++) s += fr[3*k + r] * (j2[r*nv + q] - j1[r*nv + q]);
      jc[k*nv + q] = s;
    }
    int dim = d->contact_dim[c];
    d->contact_efc_address[c] = d->nefc;
    if (dim == 1) {
      add_row(m, d, jc, pos, margin, CNSTR_CONTACT_FRICTIONLESS, c);
    } else if (dim == 3) {
      for (int k = 1; k < dim; k++) {
        double mu = d->contact_friction[5*c + k - 1];
        for (int q = 0; q < nv; q++) jac[q] = jc[q] + mu * jc[k*nv + q];
        add_row(m, d, jac, pos, margin, CNSTR_CONTACT_PYRAMIDAL, c);
        for (int q = 0; q < nv; q++) jac[q] = jc[q] - mu * jc[k*nv + q];
        add_row(m, d, jac, pos, margin, CNSTR_CONTACT_PYRAMIDAL, c);
      }
    } else {
      d->unsupported_pairs++;   /* condim 4/6 (torsional / rolling) not restated */
      d->contact_efc_address[c] = -1;
    }
  }
  free(jac); free(j1); free(j2); free(jc);

  /* mj_makeImpedance: diagApprox, R, D, K/B/imp */
  for (int r = 0; r < d->nefc; r++) {
    int id = d->efc_id[r];
    const double *solref, *solimp0;
    double diag;
    switch (d->efc_type[r]) {
      case CNSTR_LIMIT_JOINT:
        solref = m->jnt_solref + 2*id; solimp0 = m->jnt_solimp + 5*id;
        diag = m->dof_invweight0[m->jnt_dofadr[id]];
        break;
      case CNSTR_LIMIT_TENDON:
        solref = m->tendon_solref_lim + 2*id; solimp0 = m->tendon_solimp_lim + 5*id;
        diag = m->tendon_invweight0[id];
        break;
      default: {
        solref = d->contact_solref + 2*id; solimp0 = d->contact_solimp + 5*id;
        int b1 = m->geom_bodyid[d->contact_geom1[id]], b2 = m->geom_bodyid[d->contact_geom2[id]];
        double tran = m->body_invweight0[2*b1] + m->body_invweight0[2*b2];
        double rot = m->body_invweight0[2*b1 + 1] + m->body_invweight0[2*b2 + 1];
        if (d->efc_type[r] == CNSTR_CONTACT_FRICTIONLESS) diag = tran;
        else {
          int k = (r - d->contact_efc_address[id]) / 2;
          double fri = d->contact_friction[5*id + k];
          diag = tran + fri*fri*(k < 2 ? tran : rot);
        }
      }
    }
    d->efc_diagApprox[r] = diag;
    double si[5] = { clip(solimp0[0], MINIMP, MAXIMP), clip(solimp0[1], MINIMP, MAXIMP), fmax(0, solimp0[2]),
                     clip(solimp0[3], MINIMP, MAXIMP), fmax(1, solimp0[4]) };
    double imp, x = (d->efc_pos[r] - d->efc_margin[r]);
    if (si[0] == si[1] || si[2] <= MINVAL) imp = 0.5*(si[0] + si[1]);
    else {
      x = fabs(x / si[2]);
      if (x >= 1) imp = si[1];
      else if (x <= 0) imp = si[0];
      else {
        double y;
        if (si[4] == 1) y = x;
        else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
        else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
        imp = si[0] + y*(si[1] - si[0]);
      }
    }
    d->efc_R[r] = fmax(MINVAL, (1 - imp) * diag / imp);
    double K, B, dmax = si[1];
    if (solref[0] > 0) {
      double tc = fmax(solref[0], 2*m->timestep), dr = solref[1];
      K = 1 / fmax(MINVAL, dmax*dmax*tc*tc*dr*dr);
      B = 2 / fmax(MINVAL, dmax*tc);
    } else { K = -solref[0] / fmax(MINVAL, dmax*dmax); B = -solref[1] / fmax(MINVAL, dmax); }
    d->efc_KBIP[4*r] = K; d->efc_KBIP[4*r + 1] = B; d->efc_KBIP[4*r + 2] = imp; d->efc_KBIP[4*r + 3] = 0;
  }
  /* pyramidal contacts: one common R for all rows of the contact */
  for (int c = 0; c < d->ncon; c++) {
    int a = d->contact_efc_address[c];
    if (a < 0 || d->contact_dim[c] < 3) continue;
    double mu = d->contact_friction[5*c] / sqrt(m->impratio);
    double Rpy = 2*mu*mu*d->efc_R[a];
    for (int k = 0; k < 2*(d->contact_dim[c] - 1); k++) d->efc_R[a + k] = Rpy;
  }
  for (int r = 0; r < d->nefc; r++) d->efc_D[r] = 1 / d->efc_R[r];
}
/* mj_referenceConstraint */
void o_reference_constraint(const OModel* m, OData* d) {
  for (int r = 0; r < d->nefc; r++) {
    d->efc_vel[r] = dotn(d->efc_J + r*m->nv, d->qvel, m->nv);
    d->efc_aref[r] = -d->efc_KBIP[4*r + 1]*d->efc_vel[r]
                     - d->efc_KBIP[4*r]*d->efc_KBIP[4*r + 2]*(d->efc_pos[r] - d->efc_margin[r]);
  }
}

/* ------------------------------------------------------------------ velocity stage ------- */
void o_com_vel(const OModel* m, OData* d) {
  zero(d->cvel, 6);
  for (int i = 1; i < m->nbody; i++) {
    double cvel[6]; cpy(cvel, d->cvel + 6*m->body_parentid[i], 6);
    int bda = m->body_dofadr[i];
    for (int j = m->body_jntadr[i]; j < m->body_jntadr[i] + m->body_jntnum[i]; j++) {
      switch (m->jnt_type[j]) {
        case JNT_FREE:
          zero(d->cdof_dot + 6*bda, 18);
          for (int k = 0; k < 3; k++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6*(bda + k) + c] * d->qvel[bda + k];
          bda += 3; /* fallthrough */
        case JNT_BALL:
          for (int k = 0; k < 3; k++) cross_motion(d->cdof_dot + 6*(bda + k), cvel, d->cdof + 6*(bda + k));
          for (int k = 0; k < 3; k++) for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6*(bda + k) + c] * d->qvel[bda + k];
          bda += 3;
          break;
        default:
          cross_motion(d->cdof_dot + 6*bda, cvel, d->cdof + 6*bda);
          for (int c = 0; c < 6; c++) cvel[c] += d->cdof[6*bda + c] * d->qvel[bda];
          bda++;
      }
    }
    cpy(d->cvel + 6*i, cvel, 6);
  }
}
void o_passive(const OModel* m, OData* d) {
  int nv = m->nv;
  zero(d->qfrc_passive, nv);
  for (int j = 0; j < m->njnt; j++) {
    double k = m->jnt_stiffness[j];
    if (k == 0) continue;
    if (m->jnt_type[j] == JNT_HINGE || m->jnt_type[j] == JNT_SLIDE) {
      int qa = m->jnt_qposadr[j];
      d->qfrc_passive[m->jnt_dofadr[j]] = -k * (d->qpos[qa] - m->qpos_spring[qa]);
    }
  }
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] -= m->dof_damping[i] * d->qvel[i];
  for (int t = 0; t < m->ntendon; t++) {
    double k = m->tendon_stiffness[t], b = m->tendon_damping[t];
    if (k == 0 && b == 0) continue;
    double frc = -k * (d->ten_length[t] - m->tendon_lengthspring[t]) - b * d->ten_velocity[t];
    for (int c = 0; c < nv; c++) d->qfrc_passive[c] += d->ten_J[t*nv + c] * frc;
  }
}
/* mj_rne with flg_acc = 0: Coriolis, centrifugal and gravity */
void o_rne(const OModel* m, OData* d) {
  int nb = m->nbody;
  double* cacc = (double*)calloc((size_t)6*nb, sizeof(double));
  double* cfrc = (double*)calloc((size_t)6*nb, sizeof(double));
  cacc[3] = -m->gravity[0]; cacc[4] = -m->gravity[1]; cacc[5] = -m->gravity[2];
  for (int i = 1; i < nb; i++) {
    int bda = m->body_dofadr[i];
    cpy(cacc + 6*i, cacc + 6*m->body_parentid[i], 6);
    for (int k = 0; k < m->body_dofnum[i]; k++)
      for (int c = 0; c < 6; c++) cacc[6*i + c] += d->cdof_dot[6*(bda + k) + c] * d->qvel[bda + k];
    double t[6], t1[6];
    mul_inert_vec(cfrc + 6*i, d->cinert + 10*i, cacc + 6*i);
    mul_inert_vec(t, d->cinert + 10*i, d->cvel + 6*i);
    cross_force(t1, d->cvel + 6*i, t);
    for (int c = 0; c < 6; c++) cfrc[6*i + c] += t1[c];
  }
  for (int i = nb - 1; i > 0; i--) {
    int p = m->body_parentid[i];
    if (p > 0) for (int c = 0; c < 6; c++) cfrc[6*p + c] += cfrc[6*i + c];
  }
  for (int i = 0; i < m->nv; i++) d->qfrc_bias[i] = dotn(d->cdof + 6*i, cfrc + 6*m->dof_bodyid[i], 6);
  free(cacc); free(cfrc);
}

/* ------------------------------------------------------------------ actuation ------------ */
static double muscle_gain_length(double L, double lmin, double lmax) {
  if (L < lmin || L > lmax) return 0;
  double a = 0.5*(lmin + 1), b = 0.5*(1 + lmax), x;
  if (L <= a) { x = (L - lmin) / fmax(MINVAL, a - lmin); return 0.5*x*x; }
  if (L <= 1) { x = (1 - L) / fmax(MINVAL, 1 - a); return 1 - 0.5*x*x; }
  if (L <= b) { x = (L - 1) / fmax(MINVAL, b - 1); return 1 - 0.5*x*x; }
  x = (lmax - L) / fmax(MINVAL, lmax - b); return 0.5*x*x;
}
double o_muscle_gain(double len, double vel, const double* lr, double acc0, const double* prm) {
  double range0 = prm[0], range1 = prm[1], force = prm[2], scale = prm[3], lmin = prm[4], lmax = prm[5],
         vmax = prm[6], fvmax = prm[8];
  if (force < 0) force = scale / fmax(MINVAL, acc0);
  double L0 = (lr[1] - lr[0]) / fmax(MINVAL, range1 - range0);
  double L = range0 + (len - lr[0]) / fmax(MINVAL, L0);
  double V = vel / fmax(MINVAL, L0*vmax);
  double FL = muscle_gain_length(L, lmin, lmax);
  double FV, y = fvmax - 1;
  if (V <= -1) FV = 0;
  else if (V <= 0) FV = (V + 1)*(V + 1);
  else if (V <= y) FV = fvmax - (y - V)*(y - V) / fmax(MINVAL, y);
  else FV = fvmax;
  return -force*FL*FV;
}
double o_muscle_bias(double len, const double* lr, double acc0, const double* prm) {
  double range0 = prm[0], range1 = prm[1], force = prm[2], scale = prm[3], lmax = prm[5], fpmax = prm[7];
  if (force < 0) force = scale / fmax(MINVAL, acc0);
  double L0 = (lr[1] - lr[0]) / fmax(MINVAL, range1 - range0);
  double L = range0 + (len - lr[0]) / fmax(MINVAL, L0);
  double b = 0.5*(1 + lmax), x;
  if (L <= 1) return 0;
  if (L <= b) { x = (L - 1) / fmax(MINVAL, b - 1); return -force*fpmax*0.5*x*x; }
  x = (L - b) / fmax(MINVAL, b - 1); return -force*fpmax*(0.5 + x);
}
double o_muscle_dynamics(double ctrl, double act, const double* prm) {
  double c = clip(ctrl, 0, 1), a = clip(act, 0, 1), tau;
  if (c > act) tau = prm[0] * (0.5 + 1.5*a); else tau = prm[1] / (0.5 + 1.5*a);
  return (c - act) / fmax(MINVAL, tau);
}
void o_fwd_actuation(const OModel* m, OData* d) {
  int nv = m->nv, nu = m->nu, na = m->na;
  zero(d->qfrc_actuator, nv);
  for (int i = 0; i < nu; i++) {
    double ctrl = d->ctrl[i];
    if (m->actuator_ctrllimited[i]) ctrl = clip(ctrl, m->actuator_ctrlrange[2*i], m->actuator_ctrlrange[2*i + 1]);
    int ai = i - (nu - na);
    const double* dp = m->actuator_dynprm + 10*i;
    switch (m->actuator_dyntype[i]) {
      case DYN_INTEGRATOR: d->act_dot[ai] = ctrl; break;
      case DYN_FILTER: d->act_dot[ai] = (ctrl - d->act[ai]) / fmax(MINVAL, dp[0]); break;
      case DYN_MUSCLE: d->act_dot[ai] = o_muscle_dynamics(ctrl, d->act[ai], dp); break;
      default: break;
    }
    double gain, bias = 0;
    const double *gp = m->actuator_gainprm + 10*i, *bp = m->actuator_biasprm + 10*i;
    const double* lr = m->actuator_lengthrange + 2*i;
    if (m->actuator_gaintype[i] == GAIN_MUSCLE)
      gain = o_muscle_gain(d->actuator_length[i], d->actuator_velocity[i], lr, m->actuator_acc0[i], gp);
    else gain = gp[0];
    double force = (m->actuator_dyntype[i] == DYN_NONE) ? gain*ctrl : gain*d->act[ai];
    if (m->actuator_biastype[i] == BIAS_AFFINE) bias = bp[0] + bp[1]*d->actuator_length[i] + bp[2]*d->actuator_velocity[i];
    else if (m->actuator_biastype[i] == BIAS_MUSCLE) bias = o_muscle_bias(d->actuator_length[i], lr, m->actuator_acc0[i], bp);
    force += bias;
    if (m->actuator_forcelimited[i]) force = clip(force, m->actuator_forcerange[2*i], m->actuator_forcerange[2*i + 1]);
    d->actuator_force[i] = force;
    for (int c = 0; c < nv; c++) d->qfrc_actuator[c] += d->actuator_moment[i*nv + c] * force;
  }
}

/* ------------------------------------------------------------------ constraint solve ----- */
/* Primal Newton on  cost(a) = 1/2 (a-a0)' M (a-a0) + sum_i 1/2 D_i min(0, J_i a - aref_i)^2
 * (all rows this subset produces are inequality rows), exact line search, run to a gradient
 * norm of 1e-14 * scale so the oracle is the converged answer MuJoCo's Newton (tolerance 1e-8)
 * approximates. */
static int cmp_double(const void* a, const void* b) { double x = *(const double*)a, y = *(const double*)b; return (x > y) - (x < y); }
/* gradient tolerance of the Newton solve: 1e-14 (converged to rounding) for the parity tests; bench.py's CPU arm sets MuJoCo's
 * own default, opt.tolerance = 1e-8, so that the baseline is not slower than the reference's solver would be */
static double g_solver_tol = 1e-14;
void o_set_solver_tol(double tol) { g_solver_tol = tol; }

void o_fwd_constraint(const OModel* m, OData* d) {
  int nv = m->nv, ne = d->nefc;
  zero(d->qfrc_constraint, nv);
  d->solver_iter = 0;
  if (!ne) { cpy(d->qacc, d->qacc_smooth, nv); cpy(d->qacc_warmstart, d->qacc_smooth, nv); return; }
  double* a = d->qacc;
  /* work arrays on the stack (nv <= 64, ne <= njmax): no allocator traffic in the solve */
  double jar[ne], jp[ne], grad[nv], p[nv], Ma[nv], Mp[nv], H[nv*nv], L[nv*nv], bp[ne + 1];
  /* warm start: better of qacc_warmstart and qacc_smooth */
  double cost[2];
  for (int w = 0; w < 2; w++) {
    const double* x = w ? d->qacc_smooth : d->qacc_warmstart;
    double c = 0;
    for (int r = 0; r < ne; r++) { double j = dotn(d->efc_J + r*nv, x, nv) - d->efc_aref[r]; if (j < 0) c += 0.5*d->efc_D[r]*j*j; }
    for (int i = 0; i < nv; i++) { double s = dotn(d->Mdense + i*nv, x, nv); c += 0.5*(s - d->qfrc_smooth[i])*(x[i] - d->qacc_smooth[i]); }
    cost[w] = c;
  }
  cpy(a, cost[0] > cost[1] ? d->qacc_smooth : d->qacc_warmstart, nv);
  double scale = 1 / (m->stat_meaninertia * (nv > 1 ? nv : 1));
  for (int it = 0; it < 200; it++) {
    for (int r = 0; r < ne; r++) jar[r] = dotn(d->efc_J + r*nv, a, nv) - d->efc_aref[r];
    for (int i = 0; i < nv; i++) { Ma[i] = dotn(d->Mdense + i*nv, a, nv); grad[i] = Ma[i] - d->qfrc_smooth[i]; }
    for (int r = 0; r < ne; r++) if (jar[r] < 0) for (int i = 0; i < nv; i++) grad[i] += d->efc_D[r]*jar[r]*d->efc_J[r*nv + i];
    double gn = sqrt(dotn(grad, grad, nv));
    if (gn * scale < g_solver_tol) break;
    d->solver_iter = it + 1;
    cpy(H, d->Mdense, nv*nv);
    for (int r = 0; r < ne; r++) if (jar[r] < 0) {
      const double* J = d->efc_J + r*nv;
      for (int i = 0; i < nv; i++) if (J[i] != 0) for (int k = 0; k < nv; k++) H[i*nv + k] += d->efc_D[r]*J[i]*J[k];
    }
    chol_factor(L, H, nv);
    for (int i = 0; i < nv; i++) p[i] = -grad[i];
    chol_solve(L, p, nv);
    /* exact line search: phi'(alpha) = g.p + alpha p'Mp + sum_i D_i min(0, jar_i + alpha jp_i) jp_i */
    for (int r = 0; r < ne; r++) jp[r] = dotn(d->efc_J + r*nv, p, nv);
    for (int i = 0; i < nv; i++) Mp[i] = dotn(d->Mdense + i*nv, p, nv);
    double pMp = dotn(p, Mp, nv), gp = 0;
    for (int i = 0; i < nv; i++) gp += (Ma[i] - d->qfrc_smooth[i]) * p[i];
    int nb = 0;
    for (int r = 0; r < ne; r++) if (jp[r] != 0) { double t = -jar[r] / jp[r]; if (t > 0) bp[nb++] = t; }
    qsort(bp, (size_t)nb, sizeof(double), cmp_double);
    bp[nb] = INFINITY;
    double alpha = 0, lo = 0;
    for (int s = 0; s <= nb; s++) {
      double hi = bp[s], mid = isinf(hi) ? lo + 1 : 0.5*(lo + hi);
      double c0 = gp, c1 = pMp;   /* derivative = c0 + c1*alpha on this segment */
      for (int r = 0; r < ne; r++) if (jar[r] + mid*jp[r] < 0) { c0 += d->efc_D[r]*jar[r]*jp[r]; c1 += d->efc_D[r]*jp[r]*jp[r]; }
      double root = -c0 / c1;
      if (root <= hi) { alpha = root < lo ? lo : root; break; }
      lo = hi; alpha = hi;
    }
    for (int i = 0; i < nv; i++) a[i] += alpha * p[i];
    if (alpha == 0) break;
  }
  for (int r = 0; r < ne; r++) {
    double j = dotn(d->efc_J + r*nv, a, nv) - d->efc_aref[r];
    d->efc_force[r] = j < 0 ? -d->efc_D[r]*j : 0;
    for (int i = 0; i < nv; i++) d->qfrc_constraint[i] += d->efc_J[r*nv + i] * d->efc_force[r];
  }
  cpy(d->qacc_warmstart, a, nv);
}

/* ------------------------------------------------------------------ pipeline ------------- */
void o_fwd_position(const OModel* m, OData* d) {
  d->unsupported_pairs = 0; d->warn_overflow = 0;
  o_kinematics(m, d); o_com_pos(m, d); o_tendon(m, d); o_transmission(m, d);
  o_crb(m, d); o_factor_m(m, d); o_collision(m, d); o_make_constraint(m, d);
}
void o_fwd_velocity(const OModel* m, OData* d) {
  int nv = m->nv;
  for (int t = 0; t < m->ntendon; t++) d->ten_velocity[t] = dotn(d->ten_J + t*nv, d->qvel, nv);
  for (int i = 0; i < m->nu; i++) d->actuator_velocity[i] = dotn(d->actuator_moment + i*nv, d->qvel, nv);
  o_com_vel(m, d); o_passive(m, d); o_reference_constraint(m, d); o_rne(m, d);
}
void o_fwd_acceleration(const OModel* m, OData* d) {
  for (int i = 0; i < m->nv; i++)
    d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_applied[i] + d->qfrc_actuator[i];
  cpy(d->qacc_smooth, d->qfrc_smooth, m->nv);
  o_solve_m(m, d, d->qacc_smooth);
}
void o_forward(const OModel* m, OData* d) {
  o_fwd_position(m, d); o_fwd_velocity(m, d); o_fwd_actuation(m, d); o_fwd_acceleration(m, d);
  o_fwd_constraint(m, d);
}
/* mj_Euler + mj_advance */
void o_euler(const OModel* m, OData* d) {
  int nv = m->nv;
  double h = m->timestep;
  double qacc[nv];
  int damp = 0;
  for (int i = 0; i < nv; i++) if (m->dof_damping[i] > 0) { damp = 1; break; }
  if (!damp) cpy(qacc, d->qacc, nv);
  else {
    double A[nv*nv], L[nv*nv];
    cpy(A, d->Mdense, nv*nv);
    for (int i = 0; i < nv; i++) { A[i*nv + i] += h * m->dof_damping[i]; qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i]; }
    chol_factor(L, A, nv); chol_solve(L, qacc, nv);
  }
  for (int i = 0; i < m->na; i++) {
    int u = i + (m->nu - m->na);
    d->act[i] += h * d->act_dot[i];
    if (m->actuator_dyntype[u] == DYN_MUSCLE) d->act[i] = clip(d->act[i], 0, 1);
  }
  for (int i = 0; i < nv; i++) d->qvel[i] += h * qacc[i];
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], da = m->jnt_dofadr[j];
    switch (m->jnt_type[j]) {
      case JNT_FREE:
        for (int k = 0; k < 3; k++) d->qpos[qa + k] += h * d->qvel[da + k];
        qa += 3; da += 3; /* fallthrough */
      case JNT_BALL: {
        double w[3] = { d->qvel[da], d->qvel[da + 1], d->qvel[da + 2] };
        double ang = h * normalize3(w), ql[4];
        axisangle2quat(ql, w, ang);
        mulquat(d->qpos + qa, d->qpos + qa, ql); normalize4(d->qpos + qa);
        break;
      }
      default: d->qpos[qa] += h * d->qvel[da];
    }
  }
  d->time += h;
}
void o_step(const OModel* m, OData* d) { o_forward(m, d); o_euler(m, d); }
void o_step_n(const OModel* m, OData* d, int n) { for (int i = 0; i < n; i++) o_step(m, d); }
/* mj_resetData */
void o_reset(const OModel* m, OData* d) {
  cpy(d->qpos, m->qpos0, m->nq); zero(d->qvel, m->nv); zero(d->act, m->na); zero(d->ctrl, m->nu);
  zero(d->qacc_warmstart, m->nv); zero(d->qfrc_applied, m->nv); d->time = 0;
}

/* ------------------------------------------------------------------ mj_setConst ---------- */
/* Derived constants at qpos0.  Writes into the model: body_subtreemass, dof_M0, dof_invweight0,
 * body_invweight0, tendon_length0, tendon_invweight0, actuator_length0, actuator_acc0. */
void o_set_const(OModel* m, OData* d) {
  int nv = m->nv, nb = m->nbody;
  for (int i = 0; i < nb; i++) m->body_subtreemass[i] = m->body_mass[i];
  for (int i = nb - 1; i > 0; i--) m->body_subtreemass[m->body_parentid[i]] += m->body_subtreemass[i];
  cpy(d->qpos, m->qpos0, m->nq);
  /* simple dofs read dof_M0 inside o_crb; compute M generally for this pass */
  int* simple = (int*)malloc(sizeof(int)*(nv > 0 ? nv : 1));
  memcpy(simple, m->dof_simplenum, sizeof(int)*nv);
  for (int i = 0; i < nv; i++) m->dof_simplenum[i] = 0;
  o_kinematics(m, d); o_com_pos(m, d); o_tendon(m, d); o_transmission(m, d); o_crb(m, d); o_factor_m(m, d);
  memcpy(m->dof_simplenum, simple, sizeof(int)*nv); free(simple);
  for (int i = 0; i < nv; i++) m->dof_M0[i] = d->Mdense[i*nv + i];
  double* x = (double*)malloc(sizeof(double)*(nv > 0 ? nv : 1));
  double* A = (double*)malloc(sizeof(double)*(nv > 0 ? nv : 1));
  for (int i = 0; i < nv; i++) { zero(x, nv); x[i] = 1; o_solve_m(m, d, x); A[i] = x[i]; }
  for (int j = 0; j < m->njnt; j++) {
    int da = m->jnt_dofadr[j];
    if (m->jnt_type[j] == JNT_FREE) {
      double t = (A[da] + A[da + 1] + A[da + 2]) / 3, r = (A[da + 3] + A[da + 4] + A[da + 5]) / 3;
      for (int k = 0; k < 3; k++) { m->dof_invweight0[da + k] = t; m->dof_invweight0[da + 3 + k] = r; }
    } else if (m->jnt_type[j] == JNT_BALL) {
      double r = (A[da] + A[da + 1] + A[da + 2]) / 3;
      for (int k = 0; k < 3; k++) m->dof_invweight0[da + k] = r;
    } else m->dof_invweight0[da] = A[da];
  }
  double* jp = (double*)malloc(sizeof(double)*3*(nv > 0 ? nv : 1));
  double* jr = (double*)malloc(sizeof(double)*3*(nv > 0 ? nv : 1));
  m->body_invweight0[0] = m->body_invweight0[1] = 0;
  for (int b = 1; b < nb; b++) {
    o_jac(m, d, jp, jr, d->xipos + 3*b, b);
    double tr = 0, rt = 0;
    for (int r = 0; r < 3; r++) {
      cpy(x, jp + r*nv, nv); o_solve_m(m, d, x); tr += dotn(jp + r*nv, x, nv);
      cpy(x, jr + r*nv, nv); o_solve_m(m, d, x); rt += dotn(jr + r*nv, x, nv);
    }
    m->body_invweight0[2*b] = tr / 3;
    m->body_invweight0[2*b + 1] = rt / 3;
    if (m->body_weldid[b] == 0) { m->body_invweight0[2*b] = m->body_invweight0[2*b + 1] = 0; }
  }
  for (int t = 0; t < m->ntendon; t++) {
    m->tendon_length0[t] = d->ten_length[t];
    cpy(x, d->ten_J + t*nv, nv); o_solve_m(m, d, x);
    m->tendon_invweight0[t] = fmax(MINVAL, dotn(d->ten_J + t*nv, x, nv));
  }
  for (int i = 0; i < m->nu; i++) {
    m->actuator_length0[i] = d->actuator_length[i];
    cpy(x, d->actuator_moment + i*nv, nv); o_solve_m(m, d, x);
    m->actuator_acc0[i] = fmax(MINVAL, sqrt(dotn(x, x, nv)));
  }
  free(x); free(A); free(jp); free(jr);
}

/* ------------------------------------------------------------------ batched CPU stepping -- */
/* Used by bench.py's cpu_baseline / --impl reference arm: advance n independent worlds by
 * nsub substeps each; worlds [w0, w1) of caller-provided state arrays (double). */
/* One Baoding env step for worlds [w0, w1) - the work a SubprocVecEnv worker does per step in the reference: BaodingEnvV1.step's
 * target update, BaseV0.step's muscle remap, frame_skip mj_steps, get_obs (kinematics at the new state), the reward terms of
 * CustomBaodingP2Env.get_reward_dict (/root/reference/src/envs/baoding.py:403-467) with the winning run's weights.
 * task[w] = {which_task, angle1, angle2, x_radius, y_radius, period, counter}; ids = {ball1_site, ball2_site, target1_site,
 * target2_site, ball1_dofadr, ball2_dofadr}. Used only by bench.py's CPU baseline legs. */
void o_batch_env_step(OModel* m, OData* d, int w0, int w1, int nsub, double* qpos, double* qvel, double* act, double* warm,
                      const double* action, double* task, const int* ids, const double* weights, double drop_th, double proximity_th,
                      double* obs, double* reward, int* done) {
  const double cx = -0.0125, cy = -0.07, dt = nsub * m->timestep;
  const int nh = m->nq - 14, nobs = nh + 24 + m->na;
  for (int w = w0; w < w1; w++) {
    double* tk = task + 7*(size_t)w;
    cpy(d->qpos, qpos + (size_t)w*m->nq, m->nq); cpy(d->qvel, qvel + (size_t)w*m->nv, m->nv);
    cpy(d->act, act + (size_t)w*m->na, m->na); cpy(d->qacc_warmstart, warm + (size_t)w*m->nv, m->nv);
    if (tk[0] > 0.5) {
      const double sign = tk[0] < 1.5 ? -1.0 : 1.0, ang = sign * 2 * M_PI * (tk[6] * dt / tk[5]);
      for (int k = 0; k < 2; k++) {
        m->site_pos[3*ids[2 + k]] = tk[3] * cos(ang + tk[1 + k]) + cx;
        m->site_pos[3*ids[2 + k] + 1] = tk[4] * sin(ang + tk[1 + k]) + cy;
      }
    }
    tk[6] += 1;
    for (int i = 0; i < m->nu; i++) d->ctrl[i] = 1.0 / (1.0 + exp(-5.0 * (action[(size_t)w*m->nu + i] - 0.5)));
    for (int s = 0; s < nsub; s++) o_step(m, d);
    o_kinematics(m, d);
    double* o = obs + (size_t)w*nobs;
    cpy(o, d->qpos, nh);
    double d1 = 0, d2 = 0, a2 = 0;
    for (int k = 0; k < 2; k++) {
      const double* ob = d->site_xpos + 3*ids[k]; const double* tg = d->site_xpos + 3*ids[2 + k];
      for (int e = 0; e < 3; e++) {
        o[nh + 6*k + e] = ob[e]; o[nh + 6*k + 3 + e] = d->qvel[ids[4 + k] + e] * dt;
        o[nh + 12 + 3*k + e] = tg[e]; o[nh + 18 + 3*k + e] = tg[e] - ob[e];
        if (k == 0) d1 += (tg[e] - ob[e])*(tg[e] - ob[e]); else d2 += (tg[e] - ob[e])*(tg[e] - ob[e]);
      }
    }
    cpy(o + nh + 24, d->act, m->na);
    for (int i = 0; i < m->na; i++) a2 += d->act[i]*d->act[i];
    d1 = sqrt(d1); d2 = sqrt(d2);
    const int fall = d->site_xpos[3*ids[0] + 2] < drop_th || d->site_xpos[3*ids[1] + 2] < drop_th;
    const double terms[7] = { -d1, -d2, -sqrt(a2) / m->na, fall ? 0.0 : 1.0, -(d1 + d2),
                              (d1 < proximity_th && d2 < proximity_th && !fall) ? 1.0 : 0.0, fall ? 1.0 : 0.0 };
    double r = 0;
    for (int k = 0; k < 7; k++) r += weights[k] * terms[k];
    reward[w] = r; done[w] = fall;
    cpy(qpos + (size_t)w*m->nq, d->qpos, m->nq); cpy(qvel + (size_t)w*m->nv, d->qvel, m->nv);
    cpy(act + (size_t)w*m->na, d->act, m->na); cpy(warm + (size_t)w*m->nv, d->qacc_warmstart, m->nv);
  }
}

void o_batch_step(const OModel* m, OData* d, int w0, int w1, int nsub, double* qpos, double* qvel, double* act,
                  double* warm, const double* ctrl) {
  for (int w = w0; w < w1; w++) {
    cpy(d->qpos, qpos + (size_t)w*m->nq, m->nq); cpy(d->qvel, qvel + (size_t)w*m->nv, m->nv);
    cpy(d->act, act + (size_t)w*m->na, m->na); cpy(d->qacc_warmstart, warm + (size_t)w*m->nv, m->nv);
    cpy(d->ctrl, ctrl + (size_t)w*m->nu, m->nu);
    for (int s = 0; s < nsub; s++) o_step(m, d);
    cpy(qpos + (size_t)w*m->nq, d->qpos, m->nq); cpy(qvel + (size_t)w*m->nv, d->qvel, m->nv);
    cpy(act + (size_t)w*m->na, d->act, m->na); cpy(warm + (size_t)w*m->nv, d->qacc_warmstart, m->nv);
  }
}
