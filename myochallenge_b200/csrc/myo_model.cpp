// MuJoCo 2.1.0 .mjb reader and the host half of the C ABI (myo_model_*).
// File layout: int32 header {54321, sizeof(mjtNum)=8, NINT=57, NPOINTER=266}, 57 sizes, mjOption
// (25 doubles + 10 ints), mjVisual (552 B, skipped), mjStatistic (7 doubles), then one buffer of
// nbuffer bytes holding the 266 model arrays, each aligned to its own element size relative to
// the buffer start (verified to end exactly at EOF on every shipped file).
#include "myo_model.hpp"

#include <cstdio>
#include <cstring>

#include "../../include/myo_b200.h"

namespace myo {

static thread_local std::string g_err;
void set_error(const std::string& msg) { g_err = msg; }

static const char* kSizeNames[57] = {
    "nq", "nv", "nu", "na", "nbody", "njnt", "ngeom", "nsite", "ncam", "nlight", "nmesh", "nmeshvert",
    "nmeshtexvert", "nmeshface", "nmeshgraph", "nskin", "nskinvert", "nskintexvert", "nskinface", "nskinbone",
    "nskinbonevert", "nhfield", "nhfielddata", "ntex", "ntexdata", "nmat", "npair", "nexclude", "neq", "ntendon",
    "nwrap", "nsensor", "nnumeric", "nnumericdata", "ntext", "ntextdata", "ntuple", "ntupledata", "nkey", "nmocap",
    "nuser_body", "nuser_jnt", "nuser_geom", "nuser_site", "nuser_cam", "nuser_tendon", "nuser_actuator",
    "nuser_sensor", "nnames", "nM", "nemax", "njmax", "nconmax", "nstack", "nuserdata", "nsensordata", "nbuffer"};

static const char* kOptDoubles[25] = {
    "timestep", "apirate", "impratio", "tolerance", "noslip_tolerance", "mpr_tolerance", "gravity0", "gravity1",
    "gravity2", "wind0", "wind1", "wind2", "magnetic0", "magnetic1", "magnetic2", "density", "viscosity", "o_margin",
    "o_solref0", "o_solref1", "o_solimp0", "o_solimp1", "o_solimp2", "o_solimp3", "o_solimp4"};
static const char* kOptInts[10] = {"integrator", "collision", "cone", "jacobian", "solver", "iterations",
                                   "noslip_iterations", "mpr_iterations", "disableflags", "enableflags"};
static const char* kStat[7] = {"meaninertia", "meanmass", "meansize", "extent", "center0", "center1", "center2"};

struct PtrSpec { const char* name; int dtype; const char* rows; const char* cols; };
#define D DT_F64
#define I DT_I32
#define B DT_U8
#define F DT_F32
#define C DT_CHAR
// MJMODEL_POINTERS of MuJoCo 2.1.0, in file order. cols: literal count or a size name.
static const PtrSpec kPointers[] = {
    {"qpos0", D, "nq", "1"}, {"qpos_spring", D, "nq", "1"},
    {"body_parentid", I, "nbody", "1"}, {"body_rootid", I, "nbody", "1"}, {"body_weldid", I, "nbody", "1"},
    {"body_mocapid", I, "nbody", "1"}, {"body_jntnum", I, "nbody", "1"}, {"body_jntadr", I, "nbody", "1"},
    {"body_dofnum", I, "nbody", "1"}, {"body_dofadr", I, "nbody", "1"}, {"body_geomnum", I, "nbody", "1"},
    {"body_geomadr", I, "nbody", "1"}, {"body_simple", B, "nbody", "1"}, {"body_sameframe", B, "nbody", "1"},
    {"body_pos", D, "nbody", "3"}, {"body_quat", D, "nbody", "4"}, {"body_ipos", D, "nbody", "3"},
    {"body_iquat", D, "nbody", "4"}, {"body_mass", D, "nbody", "1"}, {"body_subtreemass", D, "nbody", "1"},
    {"body_inertia", D, "nbody", "3"}, {"body_invweight0", D, "nbody", "2"}, {"body_user", D, "nbody", "nuser_body"},
    {"jnt_type", I, "njnt", "1"}, {"jnt_qposadr", I, "njnt", "1"}, {"jnt_dofadr", I, "njnt", "1"},
    {"jnt_bodyid", I, "njnt", "1"}, {"jnt_group", I, "njnt", "1"}, {"jnt_limited", B, "njnt", "1"},
    {"jnt_solref", D, "njnt", "2"}, {"jnt_solimp", D, "njnt", "5"}, {"jnt_pos", D, "njnt", "3"},
    {"jnt_axis", D, "njnt", "3"}, {"jnt_stiffness", D, "njnt", "1"}, {"jnt_range", D, "njnt", "2"},
    {"jnt_margin", D, "njnt", "1"}, {"jnt_user", D, "njnt", "nuser_jnt"},
    {"dof_bodyid", I, "nv", "1"}, {"dof_jntid", I, "nv", "1"}, {"dof_parentid", I, "nv", "1"},
    {"dof_Madr", I, "nv", "1"}, {"dof_simplenum", I, "nv", "1"}, {"dof_solref", D, "nv", "2"},
    {"dof_solimp", D, "nv", "5"}, {"dof_frictionloss", D, "nv", "1"}, {"dof_armature", D, "nv", "1"},
    {"dof_damping", D, "nv", "1"}, {"dof_invweight0", D, "nv", "1"}, {"dof_M0", D, "nv", "1"},
    {"geom_type", I, "ngeom", "1"}, {"geom_contype", I, "ngeom", "1"}, {"geom_conaffinity", I, "ngeom", "1"},
    {"geom_condim", I, "ngeom", "1"}, {"geom_bodyid", I, "ngeom", "1"}, {"geom_dataid", I, "ngeom", "1"},
    {"geom_matid", I, "ngeom", "1"}, {"geom_group", I, "ngeom", "1"}, {"geom_priority", I, "ngeom", "1"},
    {"geom_sameframe", B, "ngeom", "1"}, {"geom_solmix", D, "ngeom", "1"}, {"geom_solref", D, "ngeom", "2"},
    {"geom_solimp", D, "ngeom", "5"}, {"geom_size", D, "ngeom", "3"}, {"geom_rbound", D, "ngeom", "1"},
    {"geom_pos", D, "ngeom", "3"}, {"geom_quat", D, "ngeom", "4"}, {"geom_friction", D, "ngeom", "3"},
    {"geom_margin", D, "ngeom", "1"}, {"geom_gap", D, "ngeom", "1"}, {"geom_user", D, "ngeom", "nuser_geom"},
    {"geom_rgba", F, "ngeom", "4"},
    {"site_type", I, "nsite", "1"}, {"site_bodyid", I, "nsite", "1"}, {"site_matid", I, "nsite", "1"},
    {"site_group", I, "nsite", "1"}, {"site_sameframe", B, "nsite", "1"}, {"site_size", D, "nsite", "3"},
    {"site_pos", D, "nsite", "3"}, {"site_quat", D, "nsite", "4"}, {"site_user", D, "nsite", "nuser_site"},
    {"site_rgba", F, "nsite", "4"},
    {"cam_mode", I, "ncam", "1"}, {"cam_bodyid", I, "ncam", "1"}, {"cam_targetbodyid", I, "ncam", "1"},
    {"cam_pos", D, "ncam", "3"}, {"cam_quat", D, "ncam", "4"}, {"cam_poscom0", D, "ncam", "3"},
    {"cam_pos0", D, "ncam", "3"}, {"cam_mat0", D, "ncam", "9"}, {"cam_fovy", D, "ncam", "1"},
    {"cam_ipd", D, "ncam", "1"}, {"cam_user", D, "ncam", "nuser_cam"},
    {"light_mode", I, "nlight", "1"}, {"light_bodyid", I, "nlight", "1"}, {"light_targetbodyid", I, "nlight", "1"},
    {"light_directional", B, "nlight", "1"}, {"light_castshadow", B, "nlight", "1"}, {"light_active", B, "nlight", "1"},
    {"light_pos", D, "nlight", "3"}, {"light_dir", D, "nlight", "3"}, {"light_poscom0", D, "nlight", "3"},
    {"light_pos0", D, "nlight", "3"}, {"light_dir0", D, "nlight", "3"}, {"light_attenuation", F, "nlight", "3"},
    {"light_cutoff", F, "nlight", "1"}, {"light_exponent", F, "nlight", "1"}, {"light_ambient", F, "nlight", "3"},
    {"light_diffuse", F, "nlight", "3"}, {"light_specular", F, "nlight", "3"},
    {"mesh_vertadr", I, "nmesh", "1"}, {"mesh_vertnum", I, "nmesh", "1"}, {"mesh_texcoordadr", I, "nmesh", "1"},
    {"mesh_faceadr", I, "nmesh", "1"}, {"mesh_facenum", I, "nmesh", "1"}, {"mesh_graphadr", I, "nmesh", "1"},
    {"mesh_vert", F, "nmeshvert", "3"}, {"mesh_normal", F, "nmeshvert", "3"}, {"mesh_texcoord", F, "nmeshtexvert", "2"},
    {"mesh_face", I, "nmeshface", "3"}, {"mesh_graph", I, "nmeshgraph", "1"},
    {"skin_matid", I, "nskin", "1"}, {"skin_rgba", F, "nskin", "4"}, {"skin_inflate", F, "nskin", "1"},
    {"skin_vertadr", I, "nskin", "1"}, {"skin_vertnum", I, "nskin", "1"}, {"skin_texcoordadr", I, "nskin", "1"},
    {"skin_faceadr", I, "nskin", "1"}, {"skin_facenum", I, "nskin", "1"}, {"skin_boneadr", I, "nskin", "1"},
    {"skin_bonenum", I, "nskin", "1"}, {"skin_vert", F, "nskinvert", "3"}, {"skin_texcoord", F, "nskintexvert", "2"},
    {"skin_face", I, "nskinface", "3"}, {"skin_bonevertadr", I, "nskinbone", "1"},
    {"skin_bonevertnum", I, "nskinbone", "1"}, {"skin_bonebindpos", F, "nskinbone", "3"},
    {"skin_bonebindquat", F, "nskinbone", "4"}, {"skin_bonebodyid", I, "nskinbone", "1"},
    {"skin_bonevertid", I, "nskinbonevert", "1"}, {"skin_bonevertweight", F, "nskinbonevert", "1"},
    {"hfield_size", D, "nhfield", "4"}, {"hfield_nrow", I, "nhfield", "1"}, {"hfield_ncol", I, "nhfield", "1"},
    {"hfield_adr", I, "nhfield", "1"}, {"hfield_data", F, "nhfielddata", "1"},
    {"tex_type", I, "ntex", "1"}, {"tex_height", I, "ntex", "1"}, {"tex_width", I, "ntex", "1"},
    {"tex_adr", I, "ntex", "1"}, {"tex_rgb", B, "ntexdata", "1"},
    {"mat_texid", I, "nmat", "1"}, {"mat_texuniform", B, "nmat", "1"}, {"mat_texrepeat", F, "nmat", "2"},
    {"mat_emission", F, "nmat", "1"}, {"mat_specular", F, "nmat", "1"}, {"mat_shininess", F, "nmat", "1"},
    {"mat_reflectance", F, "nmat", "1"}, {"mat_rgba", F, "nmat", "4"},
    {"pair_dim", I, "npair", "1"}, {"pair_geom1", I, "npair", "1"}, {"pair_geom2", I, "npair", "1"},
    {"pair_signature", I, "npair", "1"}, {"pair_solref", D, "npair", "2"}, {"pair_solimp", D, "npair", "5"},
    {"pair_margin", D, "npair", "1"}, {"pair_gap", D, "npair", "1"}, {"pair_friction", D, "npair", "5"},
    {"exclude_signature", I, "nexclude", "1"},
    {"eq_type", I, "neq", "1"}, {"eq_obj1id", I, "neq", "1"}, {"eq_obj2id", I, "neq", "1"},
    {"eq_active", B, "neq", "1"}, {"eq_solref", D, "neq", "2"}, {"eq_solimp", D, "neq", "5"}, {"eq_data", D, "neq", "7"},
    {"tendon_adr", I, "ntendon", "1"}, {"tendon_num", I, "ntendon", "1"}, {"tendon_matid", I, "ntendon", "1"},
    {"tendon_group", I, "ntendon", "1"}, {"tendon_limited", B, "ntendon", "1"}, {"tendon_width", D, "ntendon", "1"},
    {"tendon_solref_lim", D, "ntendon", "2"}, {"tendon_solimp_lim", D, "ntendon", "5"},
    {"tendon_solref_fri", D, "ntendon", "2"}, {"tendon_solimp_fri", D, "ntendon", "5"},
    {"tendon_range", D, "ntendon", "2"}, {"tendon_margin", D, "ntendon", "1"}, {"tendon_stiffness", D, "ntendon", "1"},
    {"tendon_damping", D, "ntendon", "1"}, {"tendon_frictionloss", D, "ntendon", "1"},
    {"tendon_lengthspring", D, "ntendon", "1"}, {"tendon_length0", D, "ntendon", "1"},
    {"tendon_invweight0", D, "ntendon", "1"}, {"tendon_user", D, "ntendon", "nuser_tendon"},
    {"tendon_rgba", F, "ntendon", "4"},
    {"wrap_type", I, "nwrap", "1"}, {"wrap_objid", I, "nwrap", "1"}, {"wrap_prm", D, "nwrap", "1"},
    {"actuator_trntype", I, "nu", "1"}, {"actuator_dyntype", I, "nu", "1"}, {"actuator_gaintype", I, "nu", "1"},
    {"actuator_biastype", I, "nu", "1"}, {"actuator_trnid", I, "nu", "2"}, {"actuator_group", I, "nu", "1"},
    {"actuator_ctrllimited", B, "nu", "1"}, {"actuator_forcelimited", B, "nu", "1"},
    {"actuator_dynprm", D, "nu", "10"}, {"actuator_gainprm", D, "nu", "10"}, {"actuator_biasprm", D, "nu", "10"},
    {"actuator_ctrlrange", D, "nu", "2"}, {"actuator_forcerange", D, "nu", "2"}, {"actuator_gear", D, "nu", "6"},
    {"actuator_cranklength", D, "nu", "1"}, {"actuator_acc0", D, "nu", "1"}, {"actuator_length0", D, "nu", "1"},
    {"actuator_lengthrange", D, "nu", "2"}, {"actuator_user", D, "nu", "nuser_actuator"},
    {"sensor_type", I, "nsensor", "1"}, {"sensor_datatype", I, "nsensor", "1"}, {"sensor_needstage", I, "nsensor", "1"},
    {"sensor_objtype", I, "nsensor", "1"}, {"sensor_objid", I, "nsensor", "1"}, {"sensor_dim", I, "nsensor", "1"},
    {"sensor_adr", I, "nsensor", "1"}, {"sensor_cutoff", D, "nsensor", "1"}, {"sensor_noise", D, "nsensor", "1"},
    {"sensor_user", D, "nsensor", "nuser_sensor"},
    {"numeric_adr", I, "nnumeric", "1"}, {"numeric_size", I, "nnumeric", "1"}, {"numeric_data", D, "nnumericdata", "1"},
    {"text_adr", I, "ntext", "1"}, {"text_size", I, "ntext", "1"}, {"text_data", C, "ntextdata", "1"},
    {"tuple_adr", I, "ntuple", "1"}, {"tuple_size", I, "ntuple", "1"}, {"tuple_objtype", I, "ntupledata", "1"},
    {"tuple_objid", I, "ntupledata", "1"}, {"tuple_objprm", D, "ntupledata", "1"},
    {"key_time", D, "nkey", "1"}, {"key_qpos", D, "nkey", "nq"}, {"key_qvel", D, "nkey", "nv"},
    {"key_act", D, "nkey", "na"}, {"key_mpos", D, "nkey", "nmocap*3"}, {"key_mquat", D, "nkey", "nmocap*4"},
    {"name_bodyadr", I, "nbody", "1"}, {"name_jntadr", I, "njnt", "1"}, {"name_geomadr", I, "ngeom", "1"},
    {"name_siteadr", I, "nsite", "1"}, {"name_camadr", I, "ncam", "1"}, {"name_lightadr", I, "nlight", "1"},
    {"name_meshadr", I, "nmesh", "1"}, {"name_skinadr", I, "nskin", "1"}, {"name_hfieldadr", I, "nhfield", "1"},
    {"name_texadr", I, "ntex", "1"}, {"name_matadr", I, "nmat", "1"}, {"name_pairadr", I, "npair", "1"},
    {"name_excludeadr", I, "nexclude", "1"}, {"name_eqadr", I, "neq", "1"}, {"name_tendonadr", I, "ntendon", "1"},
    {"name_actuatoradr", I, "nu", "1"}, {"name_sensoradr", I, "nsensor", "1"}, {"name_numericadr", I, "nnumeric", "1"},
    {"name_textadr", I, "ntext", "1"}, {"name_tupleadr", I, "ntuple", "1"}, {"name_keyadr", I, "nkey", "1"},
    {"names", C, "nnames", "1"},
};
#undef D
#undef I
#undef B
#undef F
#undef C
static const int kNPointers = (int)(sizeof(kPointers) / sizeof(kPointers[0]));
static_assert(sizeof(kPointers) / sizeof(kPointers[0]) == 266, "MuJoCo 2.1.0 has 266 model pointers");

static int dtype_size(int dt) { return dt == DT_F64 ? 8 : (dt == DT_I32 || dt == DT_F32) ? 4 : 1; }

static int resolve_dim(const Model& m, const char* spec) {
  if (spec[0] >= '0' && spec[0] <= '9') return atoi(spec);
  std::string s(spec);
  size_t star = s.find('*');
  if (star != std::string::npos) return m.sz(s.substr(0, star).c_str()) * atoi(s.c_str() + star + 1);
  return m.sz(spec);
}

std::string load_mjb(const uint8_t* raw, size_t len, Model& m, int& status) {
  status = MYO_E_FORMAT;
  const size_t kHeader = 16 + 4 * 57 + 240 + 552 + 56;
  if (len < kHeader) return "file too short for an MJB header";
  int32_t hdr[4];
  memcpy(hdr, raw, 16);
  if (hdr[0] != 54321 || hdr[1] != 8 || hdr[2] != 57 || hdr[3] != 266) {
    char buf[160];
    snprintf(buf, sizeof buf, "not a MuJoCo 2.1.0 double-precision MJB (header %d %d %d %d)", hdr[0], hdr[1], hdr[2], hdr[3]);
    return buf;
  }
  int32_t ints[57];
  memcpy(ints, raw + 16, sizeof ints);
  for (int k = 0; k < 57; k++) { m.sizes[kSizeNames[k]] = ints[k]; m.size_order.push_back(kSizeNames[k]); }
  size_t off = 16 + 4 * 57;
  double od[25]; int32_t oi[10];
  memcpy(od, raw + off, sizeof od); memcpy(oi, raw + off + sizeof od, sizeof oi);
  for (int k = 0; k < 25; k++) m.opt[kOptDoubles[k]] = od[k];
  for (int k = 0; k < 10; k++) m.opt[kOptInts[k]] = (double)oi[k];
  off += 240 + 552;
  double st[7];
  memcpy(st, raw + off, sizeof st);
  for (int k = 0; k < 7; k++) m.opt[kStat[k]] = st[k];
  off += 56;
  size_t nbuffer = (size_t)m.sz("nbuffer");
  if (len - off != nbuffer) {
    char buf[160];
    snprintf(buf, sizeof buf, "MJB buffer size mismatch: file has %zu bytes, header says %zu", len - off, nbuffer);
    return buf;
  }
  const uint8_t* buf = raw + off;
  size_t p = 0;
  m.arrays.resize(kNPointers);
  for (int k = 0; k < kNPointers; k++) {
    const PtrSpec& s = kPointers[k];
    Array& a = m.arrays[k];
    a.name = s.name; a.dtype = s.dtype;
    a.rows = resolve_dim(m, s.rows); a.cols = resolve_dim(m, s.cols);
    size_t item = (size_t)dtype_size(s.dtype), n = a.count();
    if (n) p = (p + item - 1) / item * item;
    if (p + n * item > nbuffer) return std::string("MJB truncated inside array ") + s.name;
    a.bytes.assign(buf + p, buf + p + n * item);
    if (a.bytes.empty()) a.bytes.resize(8);  // keep data() non-null
    p += n * item;
    m.index[a.name] = k;
  }
  if (p != nbuffer) {
    char b2[160];
    snprintf(b2, sizeof b2, "MJB parse ended at byte %zu of a %zu byte buffer", p, nbuffer);
    return b2;
  }
  status = MYO_OK;
  return "";
}

static std::string group_key(const char* group) {
  std::string g(group);
  if (g == "joint") g = "jnt";
  return g;
}
static const char* group_count(const std::string& g) {
  if (g == "body") return "nbody";
  if (g == "jnt") return "njnt";
  if (g == "geom") return "ngeom";
  if (g == "site") return "nsite";
  if (g == "tendon") return "ntendon";
  if (g == "actuator") return "nu";
  return nullptr;
}
std::string Model::name_of(const char* group, int id) const {
  std::string g = group_key(group);
  const char* cnt = group_count(g);
  if (!cnt || id < 0 || id >= sz(cnt)) return "";
  const Array* adr = arr("name_" + g + "adr");
  const Array* names = arr("names");
  int a = adr->as<int>()[id];
  if (a < 0 || a >= names->rows) return "";
  return std::string(names->as<char>() + a);
}
int Model::name2id(const char* group, const char* name) const {
  std::string g = group_key(group);
  const char* cnt = group_count(g);
  if (!cnt) return -1;
  for (int k = 0; k < sz(cnt); k++) if (name_of(g.c_str(), k) == name) return k;
  return -1;
}

}  // namespace myo

// ------------------------------------------------------------------------------------------ C ABI
struct myo_model { myo::Model m; };

extern "C" {

const char* myo_last_error(void) { return myo::g_err.c_str(); }
const char* myo_version(void) { return "myo_b200 0.1 (sm_100a)"; }

int myo_model_load_mjb_mem(const void* bytes, size_t len, myo_model** out) {
  if (!bytes || !out) { myo::set_error("null argument"); return MYO_E_ARG; }
  myo_model* h = new myo_model();
  int status = MYO_OK;
  std::string err = myo::load_mjb(static_cast<const uint8_t*>(bytes), len, h->m, status);
  if (!err.empty()) { delete h; myo::set_error(err); return status; }
  *out = h;
  return MYO_OK;
}

int myo_model_load_mjb(const char* path, myo_model** out) {
  if (!path || !out) { myo::set_error("null argument"); return MYO_E_ARG; }
  FILE* f = fopen(path, "rb");
  if (!f) { myo::set_error(std::string("cannot open ") + path); return MYO_E_IO; }
  std::vector<uint8_t> raw;
  uint8_t chunk[65536];
  size_t n;
  while ((n = fread(chunk, 1, sizeof chunk, f)) > 0) raw.insert(raw.end(), chunk, chunk + n);
  fclose(f);
  return myo_model_load_mjb_mem(raw.data(), raw.size(), out);
}

void myo_model_free(myo_model* m) { delete m; }

int myo_model_size(const myo_model* m, const char* name, int* out) {
  if (!m || !name || !out) { myo::set_error("null argument"); return MYO_E_ARG; }
  auto it = m->m.sizes.find(name);
  if (it == m->m.sizes.end()) { myo::set_error(std::string("unknown size ") + name); return MYO_E_ARG; }
  *out = it->second;
  return MYO_OK;
}

int myo_model_opt(const myo_model* m, const char* name, double* out) {
  if (!m || !name || !out) { myo::set_error("null argument"); return MYO_E_ARG; }
  auto it = m->m.opt.find(name);
  if (it == m->m.opt.end()) { myo::set_error(std::string("unknown option ") + name); return MYO_E_ARG; }
  *out = it->second;
  return MYO_OK;
}

int myo_model_array(myo_model* m, const char* name, void** ptr, int* rows, int* cols, int* dtype) {
  if (!m || !name || !ptr) { myo::set_error("null argument"); return MYO_E_ARG; }
  myo::Array* a = m->m.arr(name);
  if (!a) { myo::set_error(std::string("unknown model array ") + name); return MYO_E_ARG; }
  *ptr = a->bytes.data();
  if (rows) *rows = a->rows;
  if (cols) *cols = a->cols;
  if (dtype) *dtype = a->dtype;
  return MYO_OK;
}

int myo_model_name2id(const myo_model* m, const char* group, const char* name) {
  if (!m || !group || !name) { myo::set_error("null argument"); return MYO_E_ARG; }
  int id = m->m.name2id(group, name);
  if (id < 0) myo::set_error(std::string("no ") + group + " named '" + name + "'");
  return id;
}

const char* myo_model_id2name(const myo_model* m, const char* group, int id) {
  if (!m || !group) return "";
  myo_model* mm = const_cast<myo_model*>(m);
  mm->m.id2name_cache.push_back(m->m.name_of(group, id));
  return mm->m.id2name_cache.back().c_str();
}

}  // extern "C"

// accessor for the CUDA translation unit
const myo::Model& myo_model_host(const myo_model* m) { return m->m; }
