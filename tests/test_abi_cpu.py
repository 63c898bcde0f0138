"""The C-ABI library builds, loads and exports every entry point include/zbphys.h declares (no compute calls: no GPU here)."""
import ctypes
import os
import re

import pytest

import ksim_gym_zbot_b200  # noqa: F401
from ksim_gym_zbot_b200 import build as zb_build
from ksim_gym_zbot_b200 import model as M

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    path = zb_build.build_extension()
    return ctypes.CDLL(path)


def _declared():
    hdr = open(os.path.join(ROOT, "include", "zbphys.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(zb_[a-z0-9_]+)\s*\(", hdr)))


def test_every_declared_symbol_is_exported(lib):
    names = _declared()
    assert {"zb_engine_create", "zb_step_f32", "zb_forward_f32", "zb_reset_f32", "zb_terminations_f32"} <= set(names)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/zbphys.h but not exported by libzbphys.so"
    from ksim_gym_zbot_b200.engine import EXPORTS

    assert set(EXPORTS) == set(names)


def test_struct_layouts_match_header(lib):
    assert lib.zb_abi_version() == 5
    assert lib.zb_sizeof_model() == ctypes.sizeof(M.ZbModel)
    assert lib.zb_sizeof_opts() == ctypes.sizeof(M.ZbOpts)


def test_no_cpu_fallback(lib, zmodel):
    """Without a CUDA device the engine refuses to exist (ZB_ENODEV) — there is no CPU path behind the ABI."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib.zb_last_error.restype = ctypes.c_char_p
    h = ctypes.c_void_p()
    opts = M.make_opts()
    rc = lib.zb_engine_create(ctypes.byref(zmodel), ctypes.byref(opts), 0, ctypes.byref(h))
    assert rc == -2 and not h.value
    assert b"no CUDA device" in lib.zb_last_error()
    from ksim_gym_zbot_b200.engine import ZbotCudaEngine

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ZbotCudaEngine(zmodel, 4)


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ksim-gym-zbot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "zb_oracle" not in src and "oracle/" not in src, f"{f} references the oracle"


def test_model_compile_constants(zmodel):
    """mj_setConst equivalents used by the constraint rows."""
    import numpy as np

    q0 = zmodel.field("qpos0").astype(np.float64)
    Mq, _ = M._mass_matrix_np(zmodel, q0)
    np.testing.assert_allclose(zmodel.field("dof_invweight0")[6:], np.diag(np.linalg.inv(Mq))[6:], rtol=1e-5)
    assert abs(zmodel.meaninertia - np.diag(Mq).mean()) < 1e-6
    assert 4.0 < M.total_mass(zmodel) < 7.0
    assert M.foot_bottom_z(zmodel, M.default_qpos(zmodel)) < 0  # standing pose starts in light contact
    p = M.randomize_env_params(zmodel, 64, 0)
    assert p["floor_friction"].min() >= 0.3 and p["floor_friction"].max() <= 1.5  # train.py:1446-1448
    scale = p["body_mass"][:, 1] / zmodel.field("body_mass")[1]
    assert scale.min() >= 0.95 - 1e-6 and scale.max() <= 1.15 + 1e-6  # train.py:1443
    assert np.abs(p["qpos0"][:, 7:] - zmodel.field("qpos0")[7:]).max() <= np.radians(2) + 1e-6  # train.py:1445


def test_cta_shape_policy(lib):
    """pick_shape: sub-wave batches are spread over the SMs with small 128-register CTAs, larger ones run in the fewest waves of
    (balanced) 72-register CTAs of up to 28 warps."""
    import ctypes

    def shape(n, sms=148):
        w, k = ctypes.c_int(), ctypes.c_int()
        assert lib.zb_cta_shape(n, sms, ctypes.byref(w), ctypes.byref(k)) == 0
        return w.value, k.value

    assert shape(1) == (0, 1) and shape(148) == (0, 1) and shape(1024) == (0, 7) and shape(16 * 148) == (0, 16)
    assert shape(16 * 148 + 1) == (1, 17)
    assert shape(4096) == (1, 28)            # ONE wave of 147 CTAs
    assert shape(5000) == (1, 17)            # two balanced waves instead of 28 + 6
    assert shape(8192) == (1, 28) and shape(16384) == (1, 28) and shape(65536) == (1, 28)
    for n in (2369, 3000, 4144, 4145, 10000, 12345, 262144):
        wide, k = shape(n)
        waves = -(-n // (28 * 148))
        assert wide == 1 and 1 <= k <= 28 and -(-n // k) <= waves * 148  # fits the minimal number of waves
    assert shape(4096, sms=132) == (1, 16)   # another SM count: two waves of 16 (4096 / 264 = 15.5)
    w, k = ctypes.c_int(), ctypes.c_int()
    assert lib.zb_cta_shape(0, 148, ctypes.byref(w), ctypes.byref(k)) == -1
