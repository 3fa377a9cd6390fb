"""ORACLE / TEST INFRASTRUCTURE -- not part of the product path.

Pure-numpy reader *and writer* for MuJoCo 2.1.0 binary models (``.mjb``).

This is the checker-side restatement of ``mj_loadModel``/``mj_saveModel`` (MuJoCo 2.1.0
``engine_io.c``; third-party, not vendored in the reference -- layout recovered from the
shipped artefacts).  The product's loader is the C++ one in
``myochallenge_b200/csrc/mjb_loader.cpp``; tests compare the two field by field.

Reference call site that consumes these files: ``model_path`` kwargs in
/root/reference/src/envs/__init__.py:17,29,44,62 (MyoSuite ``BaseV0.__init__`` ->
``MjSim(load_model_from_mjb(path))``).
"""
from __future__ import annotations

import struct
from collections import OrderedDict

import numpy as np

MJB_MAGIC = 54321
NINT = 57
NPOINTER = 266

SIZE_NAMES = (
    "nq nv nu na nbody njnt ngeom nsite ncam nlight nmesh nmeshvert nmeshtexvert nmeshface "
    "nmeshgraph nskin nskinvert nskintexvert nskinface nskinbone nskinbonevert nhfield "
    "nhfielddata ntex ntexdata nmat npair nexclude neq ntendon nwrap nsensor nnumeric "
    "nnumericdata ntext ntextdata ntuple ntupledata nkey nmocap nuser_body nuser_jnt nuser_geom "
    "nuser_site nuser_cam nuser_tendon nuser_actuator nuser_sensor nnames nM nemax njmax nconmax "
    "nstack nuserdata nsensordata nbuffer"
).split()
assert len(SIZE_NAMES) == NINT

OPT_DOUBLES = (
    "timestep apirate impratio tolerance noslip_tolerance mpr_tolerance "
    "gravity0 gravity1 gravity2 wind0 wind1 wind2 magnetic0 magnetic1 magnetic2 density "
    "viscosity o_margin o_solref0 o_solref1 o_solimp0 o_solimp1 o_solimp2 o_solimp3 o_solimp4"
).split()
OPT_INTS = (
    "integrator collision cone jacobian solver iterations noslip_iterations mpr_iterations "
    "disableflags enableflags"
).split()
OPT_BYTES = 8 * len(OPT_DOUBLES) + 4 * len(OPT_INTS)  # 240
VIS_BYTES = 552
STAT_NAMES = "meaninertia meanmass meansize extent center0 center1 center2".split()
STAT_BYTES = 56
HEADER_BYTES = 16 + 4 * NINT + OPT_BYTES + VIS_BYTES + STAT_BYTES  # 1092

mjNEQDATA = 7
mjNDYN = mjNGAIN = mjNBIAS = 10

# (name, dtype, rows-size-name, cols)  -- cols may be an int or a size name.
_D, _I, _B, _F, _C = "<f8", "<i4", "u1", "<f4", "S1"


def _pointer_table():
    t = []
    a = t.append
    a(("qpos0", _D, "nq", 1)); a(("qpos_spring", _D, "nq", 1))
    for n in "parentid rootid weldid mocapid jntnum jntadr dofnum dofadr geomnum geomadr".split():
        a(("body_" + n, _I, "nbody", 1))
    a(("body_simple", _B, "nbody", 1)); a(("body_sameframe", _B, "nbody", 1))
    a(("body_pos", _D, "nbody", 3)); a(("body_quat", _D, "nbody", 4))
    a(("body_ipos", _D, "nbody", 3)); a(("body_iquat", _D, "nbody", 4))
    a(("body_mass", _D, "nbody", 1)); a(("body_subtreemass", _D, "nbody", 1))
    a(("body_inertia", _D, "nbody", 3)); a(("body_invweight0", _D, "nbody", 2))
    a(("body_user", _D, "nbody", "nuser_body"))
    for n in "type qposadr dofadr bodyid group".split():
        a(("jnt_" + n, _I, "njnt", 1))
    a(("jnt_limited", _B, "njnt", 1)); a(("jnt_solref", _D, "njnt", 2))
    a(("jnt_solimp", _D, "njnt", 5)); a(("jnt_pos", _D, "njnt", 3)); a(("jnt_axis", _D, "njnt", 3))
    a(("jnt_stiffness", _D, "njnt", 1)); a(("jnt_range", _D, "njnt", 2))
    a(("jnt_margin", _D, "njnt", 1)); a(("jnt_user", _D, "njnt", "nuser_jnt"))
    for n in "bodyid jntid parentid Madr simplenum".split():
        a(("dof_" + n, _I, "nv", 1))
    a(("dof_solref", _D, "nv", 2)); a(("dof_solimp", _D, "nv", 5))
    for n in "frictionloss armature damping invweight0 M0".split():
        a(("dof_" + n, _D, "nv", 1))
    for n in "type contype conaffinity condim bodyid dataid matid group priority".split():
        a(("geom_" + n, _I, "ngeom", 1))
    a(("geom_sameframe", _B, "ngeom", 1)); a(("geom_solmix", _D, "ngeom", 1))
    a(("geom_solref", _D, "ngeom", 2)); a(("geom_solimp", _D, "ngeom", 5))
    a(("geom_size", _D, "ngeom", 3)); a(("geom_rbound", _D, "ngeom", 1))
    a(("geom_pos", _D, "ngeom", 3)); a(("geom_quat", _D, "ngeom", 4))
    a(("geom_friction", _D, "ngeom", 3)); a(("geom_margin", _D, "ngeom", 1))
    a(("geom_gap", _D, "ngeom", 1)); a(("geom_user", _D, "ngeom", "nuser_geom"))
    a(("geom_rgba", _F, "ngeom", 4))
    for n in "type bodyid matid group".split():
        a(("site_" + n, _I, "nsite", 1))
    a(("site_sameframe", _B, "nsite", 1)); a(("site_size", _D, "nsite", 3))
    a(("site_pos", _D, "nsite", 3)); a(("site_quat", _D, "nsite", 4))
    a(("site_user", _D, "nsite", "nuser_site")); a(("site_rgba", _F, "nsite", 4))
    # cameras (11)
    for n in "mode bodyid targetbodyid".split():
        a(("cam_" + n, _I, "ncam", 1))
    a(("cam_pos", _D, "ncam", 3)); a(("cam_quat", _D, "ncam", 4)); a(("cam_poscom0", _D, "ncam", 3))
    a(("cam_pos0", _D, "ncam", 3)); a(("cam_mat0", _D, "ncam", 9)); a(("cam_fovy", _D, "ncam", 1))
    a(("cam_ipd", _D, "ncam", 1)); a(("cam_user", _D, "ncam", "nuser_cam"))
    # lights (17)
    for n in "mode bodyid targetbodyid".split():
        a(("light_" + n, _I, "nlight", 1))
    for n in "directional castshadow active".split():
        a(("light_" + n, _B, "nlight", 1))
    for n in "pos dir poscom0 pos0 dir0".split():
        a(("light_" + n, _D, "nlight", 3))
    a(("light_attenuation", _F, "nlight", 3)); a(("light_cutoff", _F, "nlight", 1))
    a(("light_exponent", _F, "nlight", 1))
    for n in "ambient diffuse specular".split():
        a(("light_" + n, _F, "nlight", 3))
    # meshes (11)
    for n in "vertadr vertnum texcoordadr faceadr facenum graphadr".split():
        a(("mesh_" + n, _I, "nmesh", 1))
    a(("mesh_vert", _F, "nmeshvert", 3)); a(("mesh_normal", _F, "nmeshvert", 3))
    a(("mesh_texcoord", _F, "nmeshtexvert", 2)); a(("mesh_face", _I, "nmeshface", 3))
    a(("mesh_graph", _I, "nmeshgraph", 1))
    # skins (20)
    a(("skin_matid", _I, "nskin", 1)); a(("skin_rgba", _F, "nskin", 4)); a(("skin_inflate", _F, "nskin", 1))
    for n in "vertadr vertnum texcoordadr faceadr facenum boneadr bonenum".split():
        a(("skin_" + n, _I, "nskin", 1))
    a(("skin_vert", _F, "nskinvert", 3)); a(("skin_texcoord", _F, "nskintexvert", 2))
    a(("skin_face", _I, "nskinface", 3)); a(("skin_bonevertadr", _I, "nskinbone", 1))
    a(("skin_bonevertnum", _I, "nskinbone", 1)); a(("skin_bonebindpos", _F, "nskinbone", 3))
    a(("skin_bonebindquat", _F, "nskinbone", 4)); a(("skin_bonebodyid", _I, "nskinbone", 1))
    a(("skin_bonevertid", _I, "nskinbonevert", 1)); a(("skin_bonevertweight", _F, "nskinbonevert", 1))
    # hfields (5)
    a(("hfield_size", _D, "nhfield", 4))
    for n in "nrow ncol adr".split():
        a(("hfield_" + n, _I, "nhfield", 1))
    a(("hfield_data", _F, "nhfielddata", 1))
    # textures (5)
    for n in "type height width adr".split():
        a(("tex_" + n, _I, "ntex", 1))
    a(("tex_rgb", _B, "ntexdata", 1))
    # materials (8)
    a(("mat_texid", _I, "nmat", 1)); a(("mat_texuniform", _B, "nmat", 1)); a(("mat_texrepeat", _F, "nmat", 2))
    for n in "emission specular shininess reflectance".split():
        a(("mat_" + n, _F, "nmat", 1))
    a(("mat_rgba", _F, "nmat", 4))
    # pairs (9)
    for n in "dim geom1 geom2 signature".split():
        a(("pair_" + n, _I, "npair", 1))
    a(("pair_solref", _D, "npair", 2)); a(("pair_solimp", _D, "npair", 5)); a(("pair_margin", _D, "npair", 1))
    a(("pair_gap", _D, "npair", 1)); a(("pair_friction", _D, "npair", 5))
    a(("exclude_signature", _I, "nexclude", 1))
    # equality (7)
    for n in "type obj1id obj2id".split():
        a(("eq_" + n, _I, "neq", 1))
    a(("eq_active", _B, "neq", 1)); a(("eq_solref", _D, "neq", 2)); a(("eq_solimp", _D, "neq", 5))
    a(("eq_data", _D, "neq", mjNEQDATA))
    # tendons
    for n in "adr num matid group".split():
        a(("tendon_" + n, _I, "ntendon", 1))
    a(("tendon_limited", _B, "ntendon", 1)); a(("tendon_width", _D, "ntendon", 1))
    a(("tendon_solref_lim", _D, "ntendon", 2)); a(("tendon_solimp_lim", _D, "ntendon", 5))
    a(("tendon_solref_fri", _D, "ntendon", 2)); a(("tendon_solimp_fri", _D, "ntendon", 5))
    a(("tendon_range", _D, "ntendon", 2))
    for n in "margin stiffness damping frictionloss lengthspring length0 invweight0".split():
        a(("tendon_" + n, _D, "ntendon", 1))
    a(("tendon_user", _D, "ntendon", "nuser_tendon")); a(("tendon_rgba", _F, "ntendon", 4))
    a(("wrap_type", _I, "nwrap", 1)); a(("wrap_objid", _I, "nwrap", 1)); a(("wrap_prm", _D, "nwrap", 1))
    # actuators
    for n in "trntype dyntype gaintype biastype".split():
        a(("actuator_" + n, _I, "nu", 1))
    a(("actuator_trnid", _I, "nu", 2)); a(("actuator_group", _I, "nu", 1))
    a(("actuator_ctrllimited", _B, "nu", 1)); a(("actuator_forcelimited", _B, "nu", 1))
    a(("actuator_dynprm", _D, "nu", mjNDYN)); a(("actuator_gainprm", _D, "nu", mjNGAIN))
    a(("actuator_biasprm", _D, "nu", mjNBIAS)); a(("actuator_ctrlrange", _D, "nu", 2))
    a(("actuator_forcerange", _D, "nu", 2)); a(("actuator_gear", _D, "nu", 6))
    a(("actuator_cranklength", _D, "nu", 1)); a(("actuator_acc0", _D, "nu", 1))
    a(("actuator_length0", _D, "nu", 1)); a(("actuator_lengthrange", _D, "nu", 2))
    a(("actuator_user", _D, "nu", "nuser_actuator"))
    # sensors (10)
    for n in "type datatype needstage objtype objid dim adr".split():
        a(("sensor_" + n, _I, "nsensor", 1))
    a(("sensor_cutoff", _D, "nsensor", 1)); a(("sensor_noise", _D, "nsensor", 1))
    a(("sensor_user", _D, "nsensor", "nuser_sensor"))
    a(("numeric_adr", _I, "nnumeric", 1)); a(("numeric_size", _I, "nnumeric", 1))
    a(("numeric_data", _D, "nnumericdata", 1))
    a(("text_adr", _I, "ntext", 1)); a(("text_size", _I, "ntext", 1)); a(("text_data", _C, "ntextdata", 1))
    a(("tuple_adr", _I, "ntuple", 1)); a(("tuple_size", _I, "ntuple", 1))
    a(("tuple_objtype", _I, "ntupledata", 1)); a(("tuple_objid", _I, "ntupledata", 1))
    a(("tuple_objprm", _D, "ntupledata", 1))
    a(("key_time", _D, "nkey", 1)); a(("key_qpos", _D, "nkey", "nq")); a(("key_qvel", _D, "nkey", "nv"))
    a(("key_act", _D, "nkey", "na")); a(("key_mpos", _D, "nkey", "nmocap*3")); a(("key_mquat", _D, "nkey", "nmocap*4"))
    for n, cnt in (("body", "nbody"), ("jnt", "njnt"), ("geom", "ngeom"), ("site", "nsite"), ("cam", "ncam"),
                   ("light", "nlight"), ("mesh", "nmesh"), ("skin", "nskin"), ("hfield", "nhfield"),
                   ("tex", "ntex"), ("mat", "nmat"), ("pair", "npair"), ("exclude", "nexclude"),
                   ("eq", "neq"), ("tendon", "ntendon"), ("actuator", "nu"), ("sensor", "nsensor"),
                   ("numeric", "nnumeric"), ("text", "ntext"), ("tuple", "ntuple"), ("key", "nkey")):
        a(("name_%sadr" % n, _I, cnt, 1))
    a(("names", _C, "nnames", 1))
    assert len(t) == NPOINTER, len(t)
    return t


POINTERS = _pointer_table()
NAME_GROUPS = ("body jnt geom site cam light mesh skin hfield tex mat pair exclude eq tendon "
               "actuator sensor numeric text tuple key").split()


def _dim(sizes, spec):
    if isinstance(spec, int):
        return spec
    if "*" in spec:
        n, k = spec.split("*")
        return sizes[n] * int(k)
    return sizes[spec]


class MjbModel:
    """Parsed model: ``sizes`` (dict), ``opt`` (dict), ``stat`` (dict), ``vis`` (raw bytes) and one
    numpy array per MJMODEL pointer as attributes (2-D arrays squeezed when cols == 1)."""

    def __init__(self):
        self.sizes = OrderedDict()
        self.opt = OrderedDict()
        self.stat = OrderedDict()
        self.vis = bytes(VIS_BYTES)
        self.arrays = OrderedDict()

    def __getattr__(self, k):
        d = self.__dict__
        if "arrays" in d and k in d["arrays"]:
            return d["arrays"][k]
        if "sizes" in d and k in d["sizes"]:
            return d["sizes"][k]
        raise AttributeError(k)

    # -- names ---------------------------------------------------------------------------
    def name(self, group, i):
        adr = int(self.arrays["name_%sadr" % group][i])
        raw = self.arrays["names"].tobytes()
        return raw[adr:raw.index(b"\0", adr)].decode()

    def name2id(self, group, name):
        n = len(self.arrays["name_%sadr" % group])
        for i in range(n):
            if self.name(group, i) == name:
                return i
        raise KeyError("%s '%s' not found" % (group, name))

    def names_of(self, group):
        return [self.name(group, i) for i in range(len(self.arrays["name_%sadr" % group]))]


def load(path_or_bytes) -> MjbModel:
    raw = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray)) else open(path_or_bytes, "rb").read()
    hdr = struct.unpack_from("<4i", raw, 0)
    if hdr[0] != MJB_MAGIC or hdr[1] != 8 or hdr[2] != NINT or hdr[3] != NPOINTER:
        raise ValueError("not a MuJoCo 2.1.0 double-precision MJB: header=%r" % (hdr,))
    m = MjbModel()
    ints = struct.unpack_from("<%di" % NINT, raw, 16)
    for k, v in zip(SIZE_NAMES, ints):
        m.sizes[k] = v
    off = 16 + 4 * NINT
    dbl = struct.unpack_from("<%dd" % len(OPT_DOUBLES), raw, off)
    oi = struct.unpack_from("<%di" % len(OPT_INTS), raw, off + 8 * len(OPT_DOUBLES))
    for k, v in zip(OPT_DOUBLES, dbl):
        m.opt[k] = v
    for k, v in zip(OPT_INTS, oi):
        m.opt[k] = v
    off += OPT_BYTES
    m.vis = bytes(raw[off:off + VIS_BYTES]); off += VIS_BYTES
    st = struct.unpack_from("<7d", raw, off); off += STAT_BYTES
    for k, v in zip(STAT_NAMES, st):
        m.stat[k] = v
    assert off == HEADER_BYTES
    buf = memoryview(raw)[off:]
    if len(buf) != m.sizes["nbuffer"]:
        raise ValueError("buffer size mismatch: file has %d, header says %d" % (len(buf), m.sizes["nbuffer"]))
    p = 0
    for name, dt, rows, cols in POINTERS:
        r, c = _dim(m.sizes, rows), _dim(m.sizes, cols)
        item = np.dtype(dt).itemsize
        n = r * c
        if n:
            p = (p + item - 1) // item * item
        arr = np.frombuffer(buf, dtype=dt, count=n, offset=p).copy()
        p += n * item
        m.arrays[name] = arr.reshape(r, c) if c != 1 else arr
    if p != len(buf):
        raise ValueError("parse ended at %d, buffer is %d bytes" % (p, len(buf)))
    return m


def buffer_size(m: MjbModel) -> int:
    p = 0
    for name, dt, rows, cols in POINTERS:
        n = _dim(m.sizes, rows) * _dim(m.sizes, cols)
        item = np.dtype(dt).itemsize
        if n:
            p = (p + item - 1) // item * item
        p += n * item
    return p


def dump(m: MjbModel) -> bytes:
    """Serialise back to the 2.1.0 layout (byte-exact round trip for files this reader loads)."""
    m.sizes["nbuffer"] = buffer_size(m)
    out = bytearray()
    out += struct.pack("<4i", MJB_MAGIC, 8, NINT, NPOINTER)
    out += struct.pack("<%di" % NINT, *[m.sizes[k] for k in SIZE_NAMES])
    out += struct.pack("<%dd" % len(OPT_DOUBLES), *[m.opt[k] for k in OPT_DOUBLES])
    out += struct.pack("<%di" % len(OPT_INTS), *[m.opt[k] for k in OPT_INTS])
    out += m.vis
    out += struct.pack("<7d", *[m.stat[k] for k in STAT_NAMES])
    assert len(out) == HEADER_BYTES
    body = bytearray()
    for name, dt, rows, cols in POINTERS:
        r, c = _dim(m.sizes, rows), _dim(m.sizes, cols)
        n = r * c
        item = np.dtype(dt).itemsize
        if n:
            while len(body) % item:
                body.append(0)
        arr = np.ascontiguousarray(m.arrays[name], dtype=dt).reshape(-1)
        if arr.size != n:
            raise ValueError("%s has %d elements, sizes imply %d" % (name, arr.size, n))
        body += arr.tobytes()
    assert len(body) == m.sizes["nbuffer"]
    return bytes(out + body)
