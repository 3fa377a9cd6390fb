"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol include/myo_b200.h
declares, the C++ MJB loader agrees with the checker-side reader field by field, and error paths behave."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ELBOW, FINGER, HAND_BAODING, HAND_POSE, MOTOR_FINGER, MYO_LOAD, ROOT
from myochallenge_b200 import _capi, sim
from oracle import mjb


def test_library_exports_every_declared_symbol(product_lib):
    header = open(os.path.join(ROOT, "include", "myo_b200.h")).read()
    declared = set(re.findall(r"\b(myo_[a-z0-9_]+)\s*\(", header))
    declared -= {"myo_status", "myo_task_kind", "myo_param_kind", "myo_stage"}
    assert declared, "no declarations found"
    assert declared == set(_capi.SIGNATURES), declared ^ set(_capi.SIGNATURES)
    for name in declared:
        assert hasattr(product_lib, name), name


def test_task_cfg_struct_matches_header_size(product_lib):
    # sizeof(myo_task_cfg) as the C compiler sees it vs the ctypes mirror
    import subprocess, tempfile
    src = '#include <stdio.h>\n#include "myo_b200.h"\nint main(){printf("%zu %zu", sizeof(myo_task_cfg), sizeof(myo_policy_cfg));return 0;}'
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "t.c"), "w").write(src)
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), os.path.join(d, "t.c"), "-o", os.path.join(d, "t")])
        a, b = subprocess.check_output([os.path.join(d, "t")]).split()
    assert int(a) == C.sizeof(_capi.TaskCfg) and int(b) == C.sizeof(_capi.PolicyCfg)


@pytest.mark.parametrize("path", [FINGER, MOTOR_FINGER, MYO_LOAD, HAND_BAODING, HAND_POSE, ELBOW])
def test_cpp_loader_matches_checker_reader(product_lib, path):
    ref = mjb.load(path)
    m = sim.Model(path)
    for k, v in ref.sizes.items():
        assert m.size(k) == v, k
    for k in ("timestep", "gravity2", "tolerance", "iterations", "meaninertia"):
        want = ref.opt[k] if k in ref.opt else ref.stat[k]
        assert m.opt(k) == want
    for name, arr in ref.arrays.items():
        if name == "names" or arr.size == 0:
            continue
        got = m.array(name)
        assert got.shape == np.asarray(arr).shape or got.size == np.asarray(arr).size, name
        assert np.array_equal(got.reshape(-1), np.asarray(arr).reshape(-1).view(got.dtype) if arr.dtype.kind == "S" else np.asarray(arr).reshape(-1)), name
    for g in ("body", "jnt", "geom", "site", "tendon", "actuator"):
        for i, nm in enumerate(ref.names_of(g)):
            if nm:
                assert m.id2name(g, i) == nm
                assert m.name2id(g, nm) == ref.name2id(g, nm)


def test_loader_rejects_garbage(product_lib, tmp_path):
    p = tmp_path / "bad.mjb"
    p.write_bytes(b"\0" * 2000)
    with pytest.raises(_capi.MyoError, match="not a MuJoCo 2.1.0"):
        sim.Model(str(p))
    raw = open(FINGER, "rb").read()
    p.write_bytes(raw[:-7])
    with pytest.raises(_capi.MyoError, match="size mismatch"):
        sim.Model(str(p))
    with pytest.raises(_capi.MyoError, match="cannot open"):
        sim.Model(str(tmp_path / "missing.mjb"))


def test_model_arrays_are_writable_views(product_lib):
    m = sim.Model(FINGER)
    a = m.array("body_mass")
    a[1] = 0.5
    assert m.array("body_mass")[1] == 0.5


def test_no_cpu_fallback(product_lib):
    """Without a CUDA device every batch call fails loudly (MYO_E_CUDA); with one, device='cpu' is refused."""
    import torch
    m = sim.Model(FINGER)
    cfg = m.default_task_cfg(_capi.TASK_POSE)
    if not torch.cuda.is_available():
        with pytest.raises(_capi.MyoError):
            sim.BatchSim(m, 4, cfg, device="cuda:0")
    with pytest.raises(_capi.MyoError):
        sim.BatchSim(m, 4, cfg, device="cpu")


def test_baoding_default_cfg_resolves_names(product_lib):
    m = sim.Model(HAND_BAODING)
    cfg = m.default_task_cfg(_capi.TASK_BAODING)
    assert m.id2name("body", cfg.ball_body[0]) == "ball1" and m.id2name("site", cfg.target_site[1]) == "target2_site"
    assert list(cfg.ball_qposadr) == [23, 30] and list(cfg.ball_dofadr) == [23, 29]      # /root/reference/src/envs/baoding.py:187-190
    with pytest.raises(_capi.MyoError, match="baoding names"):
        sim.Model(FINGER).default_task_cfg(_capi.TASK_BAODING)
