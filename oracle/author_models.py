"""ORACLE / TEST INFRASTRUCTURE -- offline model authoring (not part of the product path).

The reference's hand / arm models are absent from its repository
(/root/reference/.MISSING_LARGE_BLOBS:1-8: ``hand/myo_hand_baoding.mjb``, ``hand/myo_hand_pose.mjb``,
``arm/myo_elbow_1dof6muscles.mjb`` ...).  This script re-authors stand-ins with the sizes and names the
reference's env code relies on:

* 23 hand joints in the order of ``jnt_namesHand`` (/root/reference/src/envs/__init__.py:170),
  39 muscles, ``nq = 23 + 2*7`` for baoding (/root/reference/src/envs/baoding.py:183,187-190),
* bodies/geoms ``ball1``/``ball2``, sites ``ball{1,2}_site`` / ``target{1,2}_site``
  (/root/reference/src/envs/baoding.py:372-381), finger tip sites ``THtip IFtip MFtip RFtip LFtip``
  (/root/reference/src/envs/__init__.py:162), elbow joint ``r_elbow_flex`` + site ``wrist`` (:110-111).

Geometry, masses and muscle routes are synthetic (anthropometric guesses), so physics parity for
these models is oracle-vs-kernel only ("parity unpinned" against MuJoCo).  The derived constants
MuJoCo's compiler would store (dof_M0, *_invweight0, tendon_length0, actuator_acc0, subtree mass)
are filled with the oracle's ``o_set_const`` -- which reproduces MuJoCo's own values on the shipped
finger model to 1e-14 (tests/test_oracle_golden.py) -- and ``actuator_lengthrange`` by sampling the
joint ranges.

Run:  python -m oracle.author_models   (writes myochallenge_b200/assets/{hand,arm}/*.mjb)
"""
from __future__ import annotations

import os
from collections import OrderedDict

import numpy as np

from . import mjb, oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEMPLATE = os.path.join(ROOT, "myochallenge_b200", "assets", "finger", "myo_finger_v0.mjb")

FREE, SLIDE, HINGE = 0, 2, 3
PLANE, SPHERE, CAPSULE, CYLINDER, BOX = 0, 2, 3, 5, 6
W_PULLEY, W_SITE, W_SPHERE, W_CYL = 2, 3, 4, 5


def _unit(v):
    v = np.asarray(v, float)
    return v / np.linalg.norm(v)


def quat_z_to(d):
    """Quaternion rotating the z axis onto unit vector d."""
    d = _unit(d)
    z = np.array([0.0, 0.0, 1.0])
    c = float(np.dot(z, d))
    if c > 1 - 1e-12:
        return np.array([1.0, 0, 0, 0])
    if c < -1 + 1e-12:
        return np.array([0.0, 1, 0, 0])
    ax = _unit(np.cross(z, d))
    ang = np.arccos(c)
    return np.concatenate([[np.cos(ang / 2)], ax * np.sin(ang / 2)])


class Builder:
    """All body frames are axis-aligned with the world at the design pose, so positions are given in
    design-frame world coordinates and converted to body-local by subtracting the body origin."""

    def __init__(self):
        self.bodies = [dict(name="world", parent=0, origin=np.zeros(3), mass=0.0, inertia=np.zeros(3), ipos=np.zeros(3),
                            iquat=np.array([1.0, 0, 0, 0]), joints=[], simple=0)]
        self.geoms, self.sites, self.tendons, self.acts = [], [], [], []

    def bid(self, name):
        return next(i for i, b in enumerate(self.bodies) if b["name"] == name)

    def body(self, name, parent, origin, mass, inertia=(1e-6, 1e-6, 1e-6), icenter=None, iaxis=None, simple=0):
        origin = np.asarray(origin, float)
        ipos = np.zeros(3) if icenter is None else np.asarray(icenter, float) - origin
        iquat = np.array([1.0, 0, 0, 0]) if iaxis is None else quat_z_to(iaxis)
        self.bodies.append(dict(name=name, parent=self.bid(parent), origin=origin, mass=float(mass),
                                inertia=np.asarray(inertia, float), ipos=ipos, iquat=iquat, joints=[], simple=simple))
        return len(self.bodies) - 1

    def rod_body(self, name, parent, p0, p1, radius, mass):
        """Body with origin p0 whose inertia is a solid cylinder from p0 to p1."""
        p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
        L = np.linalg.norm(p1 - p0)
        it = mass * (3 * radius**2 + L**2) / 12
        return self.body(name, parent, p0, mass, (it, it, 0.5 * mass * radius**2), icenter=0.5 * (p0 + p1), iaxis=p1 - p0)

    def joint(self, body, name, jtype, axis=(0, 0, 1), anchor=None, rng=(0, 0), ref=0.0, damping=0.0, armature=0.0,
              limited=True, stiffness=0.0):
        b = self.bodies[self.bid(body)]
        pos = np.zeros(3) if anchor is None else np.asarray(anchor, float) - b["origin"]
        b["joints"].append(dict(name=name, type=jtype, axis=_unit(axis), pos=pos, range=rng, ref=ref, damping=damping,
                                armature=armature, limited=limited, stiffness=stiffness))

    def capsule(self, body, name, p0, p1, radius, contype=0, conaffinity=2):
        o = self.bodies[self.bid(body)]["origin"]
        p0, p1 = np.asarray(p0, float), np.asarray(p1, float)
        self.geoms.append(dict(name=name, body=self.bid(body), type=CAPSULE, size=(radius, 0.5 * np.linalg.norm(p1 - p0), 0),
                               pos=0.5 * (p0 + p1) - o, quat=quat_z_to(p1 - p0), contype=contype, conaffinity=conaffinity,
                               rbound=radius + 0.5 * np.linalg.norm(p1 - p0)))

    def sphere(self, body, name, center, radius, contype=0, conaffinity=0):
        o = self.bodies[self.bid(body)]["origin"]
        self.geoms.append(dict(name=name, body=self.bid(body), type=SPHERE, size=(radius, 0, 0), pos=np.asarray(center, float) - o,
                               quat=np.array([1.0, 0, 0, 0]), contype=contype, conaffinity=conaffinity, rbound=radius))

    def box(self, body, name, center, half, contype=0, conaffinity=0):
        o = self.bodies[self.bid(body)]["origin"]
        half = np.asarray(half, float)
        self.geoms.append(dict(name=name, body=self.bid(body), type=BOX, size=tuple(half), pos=np.asarray(center, float) - o,
                               quat=np.array([1.0, 0, 0, 0]), contype=contype, conaffinity=conaffinity, rbound=float(np.linalg.norm(half))))

    def cylinder(self, body, name, center, axis, radius, halflen):
        """Wrapping cylinder (never collides)."""
        o = self.bodies[self.bid(body)]["origin"]
        self.geoms.append(dict(name=name, body=self.bid(body), type=CYLINDER, size=(radius, halflen, 0), pos=np.asarray(center, float) - o,
                               quat=quat_z_to(axis), contype=0, conaffinity=0, rbound=float(np.hypot(radius, halflen))))

    def plane(self, name):
        self.geoms.append(dict(name=name, body=0, type=PLANE, size=(1, 1, 1), pos=np.zeros(3), quat=np.array([1.0, 0, 0, 0]),
                               contype=1, conaffinity=0, rbound=0.0))

    def site(self, body, name, p):
        o = self.bodies[self.bid(body)]["origin"]
        self.sites.append(dict(name=name, body=self.bid(body), pos=np.asarray(p, float) - o))
        return name

    def tendon(self, name, path):
        """path: list of ('site', name) | ('sphere'|'cyl', geom name, side-site name or None) | ('pulley', divisor)"""
        self.tendons.append(dict(name=name, path=path))

    def muscle(self, name, tendon, force):
        self.acts.append(dict(name=name, tendon=tendon, force=force))

    # ------------------------------------------------------------------ compile to MjbModel
    def compile(self, njmax=300, nconmax=64) -> mjb.MjbModel:
        tpl = mjb.load(TEMPLATE)
        m = mjb.MjbModel()
        m.opt = OrderedDict(tpl.opt)
        m.vis = tpl.vis
        m.opt.update(timestep=0.002, impratio=1.0, tolerance=1e-8, iterations=100, integrator=0, cone=0, jacobian=2, solver=2,
                     gravity0=0.0, gravity1=0.0, gravity2=-9.81, disableflags=0, enableflags=0)
        # geoms must be grouped by body
        order = sorted(range(len(self.geoms)), key=lambda g: (self.geoms[g]["body"], g))
        geoms = [self.geoms[g] for g in order]
        nbody, ngeom, nsite = len(self.bodies), len(geoms), len(self.sites)
        joints = []
        for bi, b in enumerate(self.bodies):
            for j in b["joints"]:
                joints.append(dict(j, body=bi))
        njnt = len(joints)
        nq = sum(7 if j["type"] == FREE else 1 for j in joints)
        nv = sum(6 if j["type"] == FREE else 1 for j in joints)
        ntendon, nu = len(self.tendons), len(self.acts)
        gname = {g["name"]: i for i, g in enumerate(geoms) if g["name"]}
        sname = {s["name"]: i for i, s in enumerate(self.sites)}
        tname = {t["name"]: i for i, t in enumerate(self.tendons)}
        wraps = []
        tadr, tnum = [], []
        for t in self.tendons:
            tadr.append(len(wraps))
            for e in t["path"]:
                if e[0] == "site":
                    wraps.append((W_SITE, sname[e[1]], 0.0))
                elif e[0] == "pulley":
                    wraps.append((W_PULLEY, -1, float(e[1])))
                else:
                    side = -1.0 if e[2] is None else float(sname[e[2]])
                    wraps.append((W_SPHERE if e[0] == "sphere" else W_CYL, gname[e[1]], side))
            tnum.append(len(wraps) - tadr[-1])
        nwrap = len(wraps)

        S = OrderedDict((k, 0) for k in mjb.SIZE_NAMES)
        S.update(nq=nq, nv=nv, nu=nu, na=nu, nbody=nbody, njnt=njnt, ngeom=ngeom, nsite=nsite, ntendon=ntendon, nwrap=nwrap,
                 njmax=njmax, nconmax=nconmax, nstack=200000)
        m.sizes = S
        A = OrderedDict()
        # body tables
        parent = np.array([b["parent"] for b in self.bodies], np.int32)
        jntnum = np.array([len(b["joints"]) for b in self.bodies], np.int32)
        jntadr = np.full(nbody, -1, np.int32); dofnum = np.zeros(nbody, np.int32); dofadr = np.full(nbody, -1, np.int32)
        jnt_qposadr, jnt_dofadr = [], []
        ja = qa = da = 0
        for bi, b in enumerate(self.bodies):
            if b["joints"]:
                jntadr[bi] = ja; dofadr[bi] = da
            for j in b["joints"]:
                jnt_qposadr.append(qa); jnt_dofadr.append(da)
                nqj, nvj = (7, 6) if j["type"] == FREE else (1, 1)
                qa += nqj; da += nvj; dofnum[bi] += nvj; ja += 1
        rootid = np.zeros(nbody, np.int32); weldid = np.zeros(nbody, np.int32)
        for bi in range(1, nbody):
            rootid[bi] = bi if parent[bi] == 0 else rootid[parent[bi]]
            weldid[bi] = bi if jntnum[bi] else weldid[parent[bi]]
        geomnum = np.zeros(nbody, np.int32); geomadr = np.full(nbody, -1, np.int32)
        for gi, g in enumerate(geoms):
            if geomnum[g["body"]] == 0:
                geomadr[g["body"]] = gi
            geomnum[g["body"]] += 1
        A["qpos0"] = np.zeros(nq); A["qpos_spring"] = np.zeros(nq)
        A["body_parentid"] = parent; A["body_rootid"] = rootid; A["body_weldid"] = weldid
        A["body_mocapid"] = np.full(nbody, -1, np.int32); A["body_jntnum"] = jntnum; A["body_jntadr"] = jntadr
        A["body_dofnum"] = dofnum; A["body_dofadr"] = dofadr; A["body_geomnum"] = geomnum; A["body_geomadr"] = geomadr
        A["body_simple"] = np.array([b["simple"] for b in self.bodies], np.uint8)
        A["body_sameframe"] = np.array([int(np.allclose(b["ipos"], 0) and np.allclose(b["iquat"], [1, 0, 0, 0])) for b in self.bodies], np.uint8)
        A["body_pos"] = np.array([b["origin"] - self.bodies[b["parent"]]["origin"] for b in self.bodies])
        A["body_quat"] = np.tile([1.0, 0, 0, 0], (nbody, 1))
        A["body_ipos"] = np.array([b["ipos"] for b in self.bodies]); A["body_iquat"] = np.array([b["iquat"] for b in self.bodies])
        A["body_mass"] = np.array([b["mass"] for b in self.bodies]); A["body_subtreemass"] = np.zeros(nbody)
        A["body_inertia"] = np.array([b["inertia"] for b in self.bodies]); A["body_invweight0"] = np.zeros((nbody, 2))
        # joints / dofs
        A["jnt_type"] = np.array([j["type"] for j in joints], np.int32)
        A["jnt_qposadr"] = np.array(jnt_qposadr, np.int32); A["jnt_dofadr"] = np.array(jnt_dofadr, np.int32)
        A["jnt_bodyid"] = np.array([j["body"] for j in joints], np.int32); A["jnt_group"] = np.zeros(njnt, np.int32)
        A["jnt_limited"] = np.array([int(j["limited"] and j["type"] != FREE) for j in joints], np.uint8)
        A["jnt_solref"] = np.tile([0.02, 1.0], (njnt, 1)); A["jnt_solimp"] = np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (njnt, 1))
        A["jnt_pos"] = np.array([j["pos"] for j in joints]); A["jnt_axis"] = np.array([j["axis"] for j in joints])
        A["jnt_stiffness"] = np.array([j["stiffness"] for j in joints], float); A["jnt_range"] = np.array([j["range"] for j in joints], float)
        A["jnt_margin"] = np.zeros(njnt)
        dof_body, dof_jnt, dof_parent, dof_damp, dof_arm, dof_simple = [], [], [], [], [], []
        last_dof_of_body = {}
        for ji, j in enumerate(joints):
            n = 6 if j["type"] == FREE else 1
            for k in range(n):
                i = len(dof_body)
                if i > 0 and dof_body and dof_body[-1] == j["body"]:
                    par = i - 1
                else:
                    pb = parent[j["body"]]
                    while pb and pb not in last_dof_of_body:
                        pb = parent[pb]
                    par = last_dof_of_body.get(pb, -1) if pb else -1
                dof_body.append(j["body"]); dof_jnt.append(ji); dof_parent.append(par)
                dof_damp.append(0.0 if j["type"] == FREE else j["damping"]); dof_arm.append(0.0 if j["type"] == FREE else j["armature"])
                dof_simple.append((n - k) if self.bodies[j["body"]]["simple"] else 0)
                last_dof_of_body[j["body"]] = i
            if j["type"] == FREE:
                A["qpos0"][jnt_qposadr[ji]:jnt_qposadr[ji] + 3] = self.bodies[j["body"]]["origin"]
                A["qpos0"][jnt_qposadr[ji] + 3] = 1.0
            else:
                A["qpos0"][jnt_qposadr[ji]] = j["ref"]
        A["qpos_spring"] = A["qpos0"].copy()
        depth = []
        for i in range(nv):
            depth.append(0 if dof_parent[i] < 0 else depth[dof_parent[i]] + 1)
        madr = np.concatenate([[0], np.cumsum([d + 1 for d in depth])]).astype(np.int32)
        S["nM"] = int(madr[-1])
        A["dof_bodyid"] = np.array(dof_body, np.int32); A["dof_jntid"] = np.array(dof_jnt, np.int32)
        A["dof_parentid"] = np.array(dof_parent, np.int32); A["dof_Madr"] = madr[:-1]; A["dof_simplenum"] = np.array(dof_simple, np.int32)
        A["dof_solref"] = np.tile([0.02, 1.0], (nv, 1)); A["dof_solimp"] = np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (nv, 1))
        A["dof_frictionloss"] = np.zeros(nv); A["dof_armature"] = np.array(dof_arm); A["dof_damping"] = np.array(dof_damp)
        A["dof_invweight0"] = np.zeros(nv); A["dof_M0"] = np.zeros(nv)
        # geoms
        A["geom_type"] = np.array([g["type"] for g in geoms], np.int32)
        A["geom_contype"] = np.array([g["contype"] for g in geoms], np.int32)
        A["geom_conaffinity"] = np.array([g["conaffinity"] for g in geoms], np.int32)
        A["geom_condim"] = np.full(ngeom, 3, np.int32); A["geom_bodyid"] = np.array([g["body"] for g in geoms], np.int32)
        A["geom_dataid"] = np.full(ngeom, -1, np.int32); A["geom_matid"] = np.full(ngeom, -1, np.int32)
        A["geom_group"] = np.zeros(ngeom, np.int32); A["geom_priority"] = np.zeros(ngeom, np.int32)
        A["geom_sameframe"] = np.zeros(ngeom, np.uint8); A["geom_solmix"] = np.ones(ngeom)
        A["geom_solref"] = np.tile([0.02, 1.0], (ngeom, 1)); A["geom_solimp"] = np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (ngeom, 1))
        A["geom_size"] = np.array([g["size"] for g in geoms], float); A["geom_rbound"] = np.array([g["rbound"] for g in geoms], float)
        A["geom_pos"] = np.array([g["pos"] for g in geoms]); A["geom_quat"] = np.array([g["quat"] for g in geoms])
        A["geom_friction"] = np.tile([1.0, 0.005, 0.0001], (ngeom, 1)); A["geom_margin"] = np.zeros(ngeom); A["geom_gap"] = np.zeros(ngeom)
        A["geom_rgba"] = np.tile(np.array([0.8, 0.7, 0.6, 1.0], np.float32), (ngeom, 1))
        # sites
        A["site_type"] = np.full(nsite, 2, np.int32); A["site_bodyid"] = np.array([s["body"] for s in self.sites], np.int32)
        A["site_matid"] = np.full(nsite, -1, np.int32); A["site_group"] = np.zeros(nsite, np.int32)
        A["site_sameframe"] = np.zeros(nsite, np.uint8); A["site_size"] = np.tile([0.002, 0.002, 0.002], (nsite, 1))
        A["site_pos"] = np.array([s["pos"] for s in self.sites]).reshape(nsite, 3); A["site_quat"] = np.tile([1.0, 0, 0, 0], (nsite, 1))
        A["site_rgba"] = np.tile(np.array([0.5, 0.5, 0.5, 1.0], np.float32), (nsite, 1))
        # tendons
        A["tendon_adr"] = np.array(tadr, np.int32); A["tendon_num"] = np.array(tnum, np.int32)
        A["tendon_matid"] = np.full(ntendon, -1, np.int32); A["tendon_group"] = np.zeros(ntendon, np.int32)
        A["tendon_limited"] = np.zeros(ntendon, np.uint8); A["tendon_width"] = np.full(ntendon, 0.001)
        A["tendon_solref_lim"] = np.tile([0.02, 1.0], (ntendon, 1)); A["tendon_solimp_lim"] = np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (ntendon, 1))
        A["tendon_solref_fri"] = np.tile([0.02, 1.0], (ntendon, 1)); A["tendon_solimp_fri"] = np.tile([0.9, 0.95, 0.001, 0.5, 2.0], (ntendon, 1))
        A["tendon_range"] = np.zeros((ntendon, 2))
        for k in "margin stiffness damping frictionloss lengthspring length0 invweight0".split():
            A["tendon_" + k] = np.zeros(ntendon)
        A["tendon_rgba"] = np.tile(np.array([0.95, 0.3, 0.3, 1.0], np.float32), (ntendon, 1))
        A["wrap_type"] = np.array([w[0] for w in wraps], np.int32); A["wrap_objid"] = np.array([w[1] for w in wraps], np.int32)
        A["wrap_prm"] = np.array([w[2] for w in wraps], float)
        # actuators: MuJoCo <muscle> defaults, explicit peak force
        A["actuator_trntype"] = np.full(nu, 3, np.int32); A["actuator_dyntype"] = np.full(nu, 3, np.int32)
        A["actuator_gaintype"] = np.full(nu, 1, np.int32); A["actuator_biastype"] = np.full(nu, 2, np.int32)
        A["actuator_trnid"] = np.array([[tname[a["tendon"]], -1] for a in self.acts], np.int32).reshape(nu, 2)
        A["actuator_group"] = np.zeros(nu, np.int32); A["actuator_ctrllimited"] = np.ones(nu, np.uint8)
        A["actuator_forcelimited"] = np.zeros(nu, np.uint8)
        dyn = np.zeros((nu, 10)); dyn[:, 0] = 0.01; dyn[:, 1] = 0.04
        gain = np.zeros((nu, 10))
        for i, a in enumerate(self.acts):
            gain[i, :9] = [0.75, 1.05, a["force"], 200.0, 0.5, 1.6, 1.5, 1.3, 1.2]
        A["actuator_dynprm"] = dyn; A["actuator_gainprm"] = gain; A["actuator_biasprm"] = gain.copy()
        A["actuator_ctrlrange"] = np.tile([0.0, 1.0], (nu, 1)); A["actuator_forcerange"] = np.zeros((nu, 2))
        gear = np.zeros((nu, 6)); gear[:, 0] = 1.0
        A["actuator_gear"] = gear
        for k in "cranklength acc0 length0".split():
            A["actuator_" + k] = np.zeros(nu)
        A["actuator_lengthrange"] = np.zeros((nu, 2))
        # names
        groups = dict(body=[b["name"] for b in self.bodies], jnt=[j["name"] for j in joints], geom=[g["name"] for g in geoms],
                      site=[s["name"] for s in self.sites], tendon=[t["name"] for t in self.tendons],
                      actuator=[a["name"] for a in self.acts])
        blob = bytearray()
        for g in mjb.NAME_GROUPS:
            adr = []
            for nm in groups.get(g, []):
                adr.append(len(blob)); blob += (nm or "").encode() + b"\0"
            A["name_%sadr" % g] = np.array(adr, np.int32)
        S["nnames"] = len(blob)
        A["names"] = np.frombuffer(bytes(blob), dtype="S1")
        # everything else: zero-sized / zero-filled
        for name, dt, rows, cols in mjb.POINTERS:
            if name not in A:
                A[name] = np.zeros((mjb._dim(S, rows), mjb._dim(S, cols)), dtype=dt)
        m.arrays = OrderedDict((name, A[name]) for name, _, _, _ in mjb.POINTERS)
        m.stat = OrderedDict(meaninertia=1.0, meanmass=float(np.mean(A["body_mass"][1:])), meansize=0.03, extent=0.5,
                             center0=float(self.bodies[-1]["origin"][0]), center1=float(self.bodies[-1]["origin"][1]),
                             center2=float(self.bodies[-1]["origin"][2]))
        return m


def finalize(m: mjb.MjbModel, seed=0, nsample=4000) -> mjb.MjbModel:
    """Fill the constants MuJoCo's compiler derives: mj_setConst quantities, muscle length ranges, acc0."""
    om = oracle.OracleModel(m)
    od = oracle.OracleData(om)
    # length ranges: extreme tendon lengths over the joint ranges (MuJoCo finds them by simulation)
    rng = np.random.default_rng(seed)
    jt = m.arrays["jnt_type"]; qadr = m.arrays["jnt_qposadr"]; jr = m.arrays["jnt_range"]
    lo = np.full(m.sizes["ntendon"], np.inf); hi = np.full(m.sizes["ntendon"], -np.inf)
    hinge = [j for j in range(m.sizes["njnt"]) if jt[j] != FREE]
    for s in range(nsample):
        od.qpos[:] = m.arrays["qpos0"]
        for j in hinge:
            u = rng.uniform() if s >= 2 * len(hinge) else (0.0 if s % 2 == 0 else 1.0) if s // 2 == hinge.index(j) else 0.5
            od.qpos[qadr[j]] = jr[j, 0] + u * (jr[j, 1] - jr[j, 0])
        od.call("o_kinematics"); od.call("o_com_pos"); od.call("o_tendon")
        L = np.array(od.ten_length)
        lo = np.minimum(lo, L); hi = np.maximum(hi, L)
    assert od.unsupported == 0, "authored tendon route hits an inside-wrap"
    for i in range(m.sizes["nu"]):
        t = m.arrays["actuator_trnid"][i, 0]
        pad = 0.02 * (hi[t] - lo[t]) + 1e-4
        om.actuator_lengthrange[i] = (lo[t] - pad, hi[t] + pad)
    od.call("o_set_const")
    for k in ("body_subtreemass", "body_invweight0", "dof_invweight0", "dof_M0", "tendon_length0", "tendon_invweight0",
              "actuator_length0", "actuator_acc0", "actuator_lengthrange"):
        m.arrays[k] = np.array(getattr(om, k)).reshape(m.arrays[k].shape).copy()
    m.arrays["tendon_lengthspring"] = m.arrays["tendon_length0"].copy()
    m.stat["meaninertia"] = float(np.mean(m.arrays["dof_M0"]))
    return m


# ---------------------------------------------------------------------------------------------- hand
HAND_JOINTS = ['pro_sup', 'deviation', 'flexion', 'cmc_abduction', 'cmc_flexion', 'mp_flexion', 'ip_flexion', 'mcp2_flexion',
               'mcp2_abduction', 'pm2_flexion', 'md2_flexion', 'mcp3_flexion', 'mcp3_abduction', 'pm3_flexion', 'md3_flexion',
               'mcp4_flexion', 'mcp4_abduction', 'pm4_flexion', 'md4_flexion', 'mcp5_flexion', 'mcp5_abduction', 'pm5_flexion',
               'md5_flexion']
HAND_MUSCLES = ['ECRL', 'ECRB', 'ECU', 'FCR', 'FCU', 'PL', 'PT', 'PQ', 'FDS5', 'FDS4', 'FDS3', 'FDS2', 'FDP5', 'FDP4', 'FDP3',
                'FDP2', 'EDC5', 'EDC4', 'EDC3', 'EDC2', 'EDM', 'EIP', 'EPL', 'EPB', 'FPL', 'APL', 'OP', 'RI2', 'LU_RB2', 'UI_UB2',
                'RI3', 'LU_RB3', 'UI_UB3', 'RI4', 'LU_RB4', 'UI_UB4', 'RI5', 'LU_RB5', 'UI_UB5']
# The one MuJoCo-produced hand observation in the reference (trained_models/phase_1/phase1_final.zip:_last_original_obs, a
# just-reset Baoding P1 env) pins where the balls start and where the target sites are in the world at the
# initial pose: the stand-in is placed so that its reset observation reproduces those numbers (tests/test_envs_host.py).
REF_BALL_POS = {1: np.array([-0.227, -0.511, 1.452]), 2: np.array([-0.256, -0.552, 1.442])}
REF_TARGET_POS = {1: np.array([-0.21591, -0.51066, 1.44507]), 2: np.array([-0.25508, -0.54608, 1.45052])}
P0 = np.array([-0.2357, -0.4734, 1.408])          # wrist centre (design pose = palm up, pro_sup = -1.57)
ORBIT_C = P0 + np.array([0.0, -0.055, 0.0])      # centre of the palm under the target orbit


def _rot_between(a, b):
    """Smallest rotation matrix taking direction a onto direction b."""
    a, b = _unit(a), _unit(b)
    v, c = np.cross(a, b), float(np.dot(a, b))
    K = np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])
    return np.eye(3) + K + K @ K / (1.0 + c)


def _mat2quat(R):
    w = np.sqrt(max(0.0, 1.0 + R[0, 0] + R[1, 1] + R[2, 2])) / 2.0
    return np.array([w, (R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w)])


def build_hand(kind: str) -> mjb.MjbModel:
    """kind: 'pose' (hand only), 'baoding' (+ two balls and the target sites), 'die' (+ the die and its target)."""
    balls = kind == "baoding"
    B = Builder()
    P = lambda x, y, z: P0 + np.array([x, y, z], float)
    B.body("forearm", "world", P(0, 0.25, 0), 1.0, (4e-3, 4e-4, 4e-3))
    B.rod_body("radius", "forearm", P(0, 0.03, 0), P(0, 0.0, 0), 0.02, 0.10)
    B.joint("radius", "pro_sup", HINGE, (0, 1, 0), rng=(-1.75, 1.57), ref=-1.57, damping=0.2, armature=0.005, stiffness=1.0)
    B.body("carpus", "radius", P(0, 0, 0), 0.05, (8e-6, 8e-6, 8e-6))
    B.joint("carpus", "deviation", HINGE, (0, 0, 1), rng=(-0.35, 0.35), damping=0.2, armature=0.005, stiffness=3.0)
    B.body("palm", "carpus", P(0, 0, 0), 0.30, (2.6e-4, 2.2e-4, 4.5e-4), icenter=P(0, -0.05, 0.002))
    B.joint("palm", "flexion", HINGE, (-1, 0, 0), rng=(-1.0, 1.0), damping=0.2, armature=0.005, stiffness=3.0)
    # palm surface: metacarpal capsules with a transverse arch + raised rims (thenar, hypothenar, heel)
    for k, (x, z) in enumerate([(-0.034, 0.010), (-0.017, 0.003), (0.0, 0.0), (0.017, 0.003), (0.034, 0.010)]):
        B.capsule("palm", f"palm_mc{k}", P(x, -0.012, z), P(x, -0.098, z), 0.010)
    B.capsule("palm", "palm_thenar", P(-0.047, -0.018, 0.020), P(-0.047, -0.070, 0.020), 0.012)
    B.capsule("palm", "palm_hypothenar", P(0.047, -0.015, 0.019), P(0.047, -0.095, 0.019), 0.011)
    B.capsule("palm", "palm_heel", P(-0.035, -0.006, 0.017), P(0.035, -0.006, 0.017), 0.011)
    # thumb
    dt_ = _unit([-0.55, -0.65, 0.45])
    a_ab = _unit(np.cross(dt_, [0, 0, 1]))
    a_fl = -_unit(np.cross(a_ab, dt_))
    t0 = P(-0.038, -0.022, 0.008); t1 = t0 + 0.040 * dt_; t2 = t1 + 0.030 * dt_; t3 = t2 + 0.025 * dt_
    B.rod_body("thumb_mc", "palm", t0, t1, 0.011, 0.030)
    B.joint("thumb_mc", "cmc_abduction", HINGE, a_ab, rng=(-0.4, 0.8), damping=0.05, armature=0.002)
    B.joint("thumb_mc", "cmc_flexion", HINGE, a_fl, rng=(-0.5, 0.9), damping=0.05, armature=0.002)
    B.rod_body("thumb_prox", "thumb_mc", t1, t2, 0.010, 0.020)
    B.joint("thumb_prox", "mp_flexion", HINGE, a_fl, rng=(-0.8, 0.4), damping=0.05, armature=0.002)
    B.rod_body("thumb_dist", "thumb_prox", t2, t3, 0.009, 0.012)
    B.joint("thumb_dist", "ip_flexion", HINGE, a_fl, rng=(-1.3, 0.3), damping=0.05, armature=0.002)
    B.capsule("thumb_mc", "thumb_mc_g", t0, t1, 0.011); B.capsule("thumb_prox", "thumb_prox_g", t1, t2, 0.010)
    B.capsule("thumb_dist", "thumb_dist_g", t2, t3 - 0.006 * dt_, 0.009)
    B.site("thumb_dist", "THtip", t3)
    B.site("world", "THtip_target", t3)        # PoseEnvV0 visualisation targets: <tip>_target, moved by update_target
    # fingers 2..5 (index .. little); rest pose tilted 15 degrees toward the palm
    tilt = np.deg2rad(15.0)
    df = np.array([0.0, -np.cos(tilt), np.sin(tilt)]); nf = np.array([0.0, np.sin(tilt), np.cos(tilt)])
    fx = {2: -0.030, 3: -0.010, 4: 0.010, 5: 0.030}
    flen = {2: (0.040, 0.025, 0.020), 3: (0.044, 0.028, 0.020), 4: (0.041, 0.026, 0.020), 5: (0.033, 0.020, 0.018)}
    tipname = {2: "IFtip", 3: "MFtip", 4: "RFtip", 5: "LFtip"}
    fpts = {}
    for k in (2, 3, 4, 5):
        L1, L2, L3 = flen[k]
        p0 = P(fx[k], -0.100, 0.002); p1 = p0 + L1 * df; p2 = p1 + L2 * df; p3 = p2 + L3 * df
        fpts[k] = (p0, p1, p2, p3)
        B.rod_body(f"prox{k}", "palm", p0, p1, 0.009, 0.020)
        B.joint(f"prox{k}", f"mcp{k}_flexion", HINGE, (-1, 0, 0), rng=(-0.35, 1.57), damping=0.05, armature=0.002)
        B.joint(f"prox{k}", f"mcp{k}_abduction", HINGE, nf, rng=(-0.26, 0.26), damping=0.05, armature=0.002)
        B.rod_body(f"mid{k}", f"prox{k}", p1, p2, 0.008, 0.012)
        B.joint(f"mid{k}", f"pm{k}_flexion", HINGE, (-1, 0, 0), rng=(0.0, 1.57), damping=0.05, armature=0.002)
        B.rod_body(f"dist{k}", f"mid{k}", p2, p3, 0.007, 0.008)
        B.joint(f"dist{k}", f"md{k}_flexion", HINGE, (-1, 0, 0), rng=(0.0, 1.57), damping=0.05, armature=0.002)
        B.capsule(f"prox{k}", f"prox{k}_g", p0, p1, 0.009); B.capsule(f"mid{k}", f"mid{k}_g", p1, p2, 0.008)
        B.capsule(f"dist{k}", f"dist{k}_g", p2, p3 - 0.005 * df, 0.007)
        B.site(f"dist{k}", tipname[k], p3)
        B.site("world", tipname[k] + "_target", p3)
        # extensor wrapping cylinders at MCP (on the palm) and PIP (on the proximal phalanx), axis = flexion axis
        B.cylinder("palm", f"mcp{k}_wrap", p0, (1, 0, 0), 0.0075, 0.006)
        B.cylinder(f"prox{k}", f"pip{k}_wrap", p1, (1, 0, 0), 0.0060, 0.005)
        B.site("palm", f"mcp{k}_side", p0 - 0.015 * nf)
        B.site(f"prox{k}", f"pip{k}_side", p1 - 0.012 * nf)

    # ---- muscle routes -------------------------------------------------------------------------
    def route(name, pts, force):
        path = []
        for i, e in enumerate(pts):
            if e[0] in ("sphere", "cyl"):
                path.append(e)
            else:
                body, p = e
                path.append(("site", B.site(body, f"{name}-P{i + 1}", p)))
        B.tendon(name + "_tendon", path)
        B.muscle(name, name + "_tendon", force)

    wrist = dict(ECRL=(-0.022, -0.013, 300), ECRB=(-0.010, -0.015, 250), ECU=(0.024, -0.013, 200),
                 FCR=(-0.018, 0.014, 250), FCU=(0.024, 0.013, 300), PL=(0.0, 0.016, 100))
    routes = {}
    for nm, (x, z, f) in wrist.items():
        routes[nm] = ([("forearm", P(0.8 * x, 0.20, 1.1 * z)), ("radius", P(x, 0.034, z)), ("palm", P(x, -0.016, 0.8 * z))], f)
    routes["PT"] = ([("forearm", P(0.022, 0.15, 0.0)), ("radius", P(-0.018, 0.05, 0.006))], 150)
    routes["PQ"] = ([("forearm", P(0.018, 0.046, 0.004)), ("radius", P(-0.016, 0.040, 0.009))], 80)
    for k in (2, 3, 4, 5):
        p0, p1, p2, p3 = fpts[k]
        L1, L2, L3 = flen[k]
        x = fx[k]
        vol = [("forearm", P(0.5 * x, 0.18, 0.012)), ("radius", P(0.5 * x, 0.032, 0.013)), ("palm", P(0.7 * x, -0.020, 0.006)),
               ("palm", P(x, -0.090, 0.011)), (f"prox{k}", p0 + 0.5 * L1 * df + 0.0075 * nf)]
        routes[f"FDS{k}"] = (vol + [(f"mid{k}", p1 + 0.4 * L2 * df + 0.0065 * nf)], 80)
        routes[f"FDP{k}"] = (vol[:4] + [(f"prox{k}", p0 + 0.55 * L1 * df + 0.0065 * nf), (f"mid{k}", p1 + 0.5 * L2 * df + 0.006 * nf),
                                         (f"dist{k}", p2 + 0.4 * L3 * df + 0.005 * nf)], 90)
        ext = [("forearm", P(0.5 * x, 0.18, -0.012)), ("radius", P(0.5 * x, 0.032, -0.014)), ("palm", P(x, -0.020, -0.012)),
               ("palm", P(x, -0.080, -0.010)), ("cyl", f"mcp{k}_wrap", f"mcp{k}_side"), (f"prox{k}", p0 + 0.5 * L1 * df - 0.008 * nf),
               ("cyl", f"pip{k}_wrap", f"pip{k}_side"), (f"mid{k}", p1 + 0.5 * L2 * df - 0.007 * nf),
               (f"dist{k}", p2 + 0.3 * L3 * df - 0.006 * nf)]
        routes[f"EDC{k}"] = (ext, 40)
    def shifted(pts, dx):
        out = []
        for e in pts:
            out.append(e if e[0] in ("sphere", "cyl") else (e[0], np.asarray(e[1]) + np.array([dx, 0, 0])))
        return out
    routes["EDM"] = (shifted(routes["EDC5"][0][:6], 0.004), 25)
    routes["EIP"] = (shifted(routes["EDC2"][0][:6], -0.004), 25)
    nt = _unit(np.cross(a_fl, dt_))       # thumb "volar" normal (flexion moves the tip along +nt)
    if np.dot(nt, [0.6, 0, 0.6]) < 0:
        nt = -nt
    routes["FPL"] = ([("forearm", P(-0.010, 0.18, 0.010)), ("radius", P(-0.020, 0.032, 0.012)), ("palm", P(-0.030, -0.012, 0.010)),
                      ("thumb_mc", t0 + 0.020 * dt_ + 0.010 * nt), ("thumb_prox", t1 + 0.015 * dt_ + 0.009 * nt),
                      ("thumb_dist", t2 + 0.010 * dt_ + 0.008 * nt)], 80)
    routes["EPL"] = ([("forearm", P(-0.005, 0.18, -0.012)), ("radius", P(-0.016, 0.032, -0.013)), ("palm", P(-0.034, -0.010, -0.006)),
                      ("thumb_mc", t0 + 0.020 * dt_ - 0.011 * nt), ("thumb_prox", t1 + 0.015 * dt_ - 0.010 * nt),
                      ("thumb_dist", t2 + 0.010 * dt_ - 0.009 * nt)], 40)
    routes["EPB"] = ([("forearm", P(-0.008, 0.16, -0.012)), ("radius", P(-0.020, 0.032, -0.011)), ("palm", P(-0.038, -0.010, -0.004)),
                      ("thumb_mc", t0 + 0.022 * dt_ - 0.011 * nt), ("thumb_prox", t1 + 0.010 * dt_ - 0.010 * nt)], 30)
    routes["APL"] = ([("forearm", P(-0.012, 0.15, -0.008)), ("radius", P(-0.026, 0.032, -0.004)),
                      ("thumb_mc", t0 + 0.008 * dt_ - 0.010 * a_ab * np.sign(a_ab[2] if abs(a_ab[2]) > 1e-9 else 1.0))], 60)
    routes["OP"] = ([("palm", P(0.0, -0.030, 0.012)), ("thumb_mc", t0 + 0.028 * dt_ + 0.009 * nt)], 50)
    for k in (2, 3, 4, 5):
        p0, p1, p2, p3 = fpts[k]
        L1, L2, L3 = flen[k]
        x = fx[k]
        ex = np.array([1.0, 0, 0])
        routes[f"RI{k}"] = ([("palm", P(x - 0.007, -0.060, 0.000)), (f"prox{k}", p0 + 0.3 * L1 * df - 0.0075 * ex + 0.003 * nf)], 25)
        routes[f"LU_RB{k}"] = ([("palm", P(x - 0.004, -0.070, 0.009)), (f"prox{k}", p0 + 0.5 * L1 * df - 0.005 * ex + 0.006 * nf),
                                (f"mid{k}", p1 + 0.3 * L2 * df - 0.006 * nf)], 20)
        routes[f"UI_UB{k}"] = ([("palm", P(x + 0.007, -0.060, 0.000)), (f"prox{k}", p0 + 0.3 * L1 * df + 0.0075 * ex + 0.003 * nf)], 25)
    for nm in HAND_MUSCLES:
        route(nm, *routes[nm])

    if balls:
        B.plane("floor")
        # target sites live on a fixed frame rotated -90 deg about z w.r.t. the world (as the reference's
        # observations imply: local (x, y) -> world (y, -x)); centre_pos = (-0.0125, -0.07) maps onto ORBIT_C
        for k in (1, 2):
            c = REF_BALL_POS[k]
            B.body(f"ball{k}", "world", c, 0.043, (8.3248e-6,) * 3, simple=1)
            B.joint(f"ball{k}", f"ball{k}_free", FREE, limited=False)
            B.sphere(f"ball{k}", f"ball{k}", c, 0.022, contype=3, conaffinity=1)
            B.site(f"ball{k}", f"ball{k}_site", c)
        # The target sites ride on the palm ("desired_positions_wrt_palm" in BaodingEnvV1.step; the reference's obs_rms shows
        # the targets' world z moving with the wrist: std 1 cm). They sit in a massless frame fixed to the palm whose pose is
        # solved from the reference observation: local (x, y) = radius * (cos, sin)(start angle) + center_pos must land on
        # REF_TARGET_POS at the initial pose. The frame is Rz(-90 deg) (local (x, y) -> world (y, -x)) followed by the
        # smallest rotation that aligns the two sites' separation with the observed one.
        B.body("target_frame", "palm", P0, 0.0, (0, 0, 0))
        for k in (1, 2):
            B.site("target_frame", f"target{k}_site", P0)   # position set below (rotated frame)
    if kind == "die":
        # Die: the reference's reset (/root/reference/src/envs/reorient.py:100-160) treats the LAST THREE geoms of body "Object" as
        # boxes that grow / shrink with obj_size_change in all three half sizes (earlier geoms, if any, only in size[1]): here
        # the body carries exactly three boxes - the colliding cube and two thinner, non-colliding slabs inside it (face
        # markings). The target die is a non-colliding copy on the body "target" (moved / turned by reset through body_pos /
        # body_quat), shown offset from the hand: goal_obj_offset = target_o - object_o at the initial pose.
        h, k = 0.015, 0.012
        c = ORBIT_C + np.array([0.0, 0.0, 0.038])
        mass = 0.05
        B.body("Object", "world", c, mass, (mass * (2 * h) ** 2 / 6.0,) * 3, simple=1)
        B.joint("Object", "OBJT_free", FREE, limited=False)
        B.box("Object", "dice", c, (h, h, h), contype=3, conaffinity=1)
        for nm, half in (("dice_marks_y", (h, k, h)), ("dice_marks_x", (k, h, h))):
            B.box("Object", nm, c, half)
        B.site("Object", "object_o", c)
        tc = c + np.array([0.0, 0.0, 0.15])
        B.body("target", "world", tc, 0.0, (0, 0, 0))
        B.box("target", "target_dice", tc, (h, h, h))
        B.site("target", "target_o", tc)
        B.site("target", "target_ball", tc + np.array([0.0, 0.0, 0.03]))
    m = B.compile()
    if balls:
        tb = B.bid("target_frame")
        cx, cy = -0.0125, -0.07
        loc = {k: np.array([0.025 * np.cos(ang) + cx, 0.028 * np.sin(ang) + cy, 0.0]) for k, ang in ((1, 0.75 * np.pi), (2, -0.25 * np.pi))}
        Rz = np.array([[0.0, 1.0, 0.0], [-1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
        R = _rot_between(Rz @ (loc[2] - loc[1]), REF_TARGET_POS[2] - REF_TARGET_POS[1]) @ Rz
        origin = 0.5 * (REF_TARGET_POS[1] + REF_TARGET_POS[2]) - R @ (0.5 * (loc[1] + loc[2]))
        m.arrays["body_quat"][tb] = _mat2quat(R)
        m.arrays["body_pos"][tb] = origin - P0          # the palm's frame is axis-aligned with origin P0 at the design pose
        for k in (1, 2):
            m.arrays["site_pos"][m.name2id("site", f"target{k}_site")] = loc[k]
    return finalize(m)


def build_elbow() -> mjb.MjbModel:
    """1 hinge (r_elbow_flex), 6 muscles (3 flexors, 3 extensors), wrist site -- config 1 stand-in."""
    B = Builder()
    S0 = np.array([0.0, 0.0, 1.2])
    B.body("humerus", "world", S0, 2.0, (0.02, 0.02, 0.002))
    E = S0 + np.array([0, 0, -0.30])
    B.rod_body("forearm", "humerus", E, E + np.array([0, 0, -0.27]), 0.025, 1.2)
    B.joint("forearm", "r_elbow_flex", HINGE, (0, 1, 0), rng=(0.0, 2.27), damping=0.3, armature=0.01)
    B.capsule("forearm", "forearm_g", E, E + np.array([0, 0, -0.27]), 0.025, contype=0, conaffinity=0)
    B.site("forearm", "wrist", E + np.array([0, 0, -0.27]))
    B.site("world", "wrist_target", E + np.array([0, 0, -0.27]))      # PoseEnvV0 visualisation target
    B.cylinder("humerus", "elbow_wrap", E, (0, 1, 0), 0.018, 0.03)
    B.site("humerus", "elbow_side_post", E + np.array([0.05, 0, 0.0]))
    specs = [("TRIlong", -1, 0.012, 800), ("TRIlat", -1, 0.0, 600), ("TRImed", -1, -0.012, 600),
             ("BIClong", 1, 0.012, 600), ("BICshort", 1, -0.012, 450), ("BRA", 1, 0.0, 1000)]
    for nm, side, y, force in specs:
        if side > 0:   # flexors: anterior (-x), straight line to the forearm
            path = [("site", B.site("humerus", nm + "-P1", S0 + np.array([-0.02, y, -0.08]))),
                    ("site", B.site("forearm", nm + "-P2", E + np.array([-0.012, y, -0.05])))]
        else:          # extensors: posterior (+x), wrap over the elbow cylinder
            path = [("site", B.site("humerus", nm + "-P1", S0 + np.array([0.025, y, -0.10]))),
                    ("cyl", "elbow_wrap", "elbow_side_post"),
                    ("site", B.site("forearm", nm + "-P2", E + np.array([0.022, y, -0.03])))]
        B.tendon(nm + "_tendon", path)
        B.muscle(nm, nm + "_tendon", force)
    return finalize(B.compile())


def main():
    out = os.path.join(ROOT, "myochallenge_b200", "assets")
    os.makedirs(os.path.join(out, "hand"), exist_ok=True)
    os.makedirs(os.path.join(out, "arm"), exist_ok=True)
    for rel, m in (("hand/myo_hand_baoding.mjb", build_hand("baoding")), ("hand/myo_hand_pose.mjb", build_hand("pose")),
                   ("hand/myo_hand_die.mjb", build_hand("die")), ("arm/myo_elbow_1dof6muscles.mjb", build_elbow())):
        raw = mjb.dump(m)
        with open(os.path.join(out, rel), "wb") as f:
            f.write(raw)
        s = m.sizes
        print(f"{rel}: {len(raw)} B  nq={s['nq']} nv={s['nv']} nu={s['nu']} nbody={s['nbody']} ngeom={s['ngeom']} nsite={s['nsite']} "
              f"ntendon={s['ntendon']} nwrap={s['nwrap']} nM={s['nM']}")


if __name__ == "__main__":
    main()
