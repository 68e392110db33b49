"""The C-ABI library loads and exports every symbol include/wbc_b200.h declares (no compute calls: CPU only)."""
import ctypes as C
import os
import re

from wbc_quadruped_dob_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    txt = open(os.path.join(ROOT, "include", "wbc_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(wbc_[a-z0-9_]+)\s*\(", txt)))


def test_header_declares_expected_entry_points():
    names = declared_functions()
    assert set(api.EXPORTS) == set(names), (names, api.EXPORTS)


def test_library_exports_every_declared_symbol():
    lib = api.load()
    for name in declared_functions():
        assert hasattr(lib, name), "libwbc_b200.so does not export " + name


def test_params_layout_and_defaults():
    p = api.default_params()
    assert C.sizeof(api.Params) == 176
    assert (p.kcom, p.dcom, p.q1_weight, p.slack_weight, p.mu, p.tau_max) == (2500.0, 50.0, 50.0, 1e8, 0.6, 60.0)
    assert (p.joint_dt, p.kp_sw, p.kd_sw, p.g_acc, p.obs_gain, p.obs_dt) == (0.025, 300.0, 20.0, 9.81, 10.0, 0.0025)
    assert tuple(p.gravity) == (0.0, 0.0, -9.8)
    assert (p.qp_epsx, p.qp_rho, p.qp_outerits) == (1e-2, 1e4, 5)
    assert (p.obs_gain2, p.obs_order, p.obs_form) == (1.0, 1, 0)
    assert api.load().wbc_version().startswith(b"wbc_b200")


def test_no_cpu_fallback_in_product_sources():
    """The product package must never import or link the oracle."""
    pkg = os.path.join(ROOT, "wbc_quadruped_dob_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                txt = open(os.path.join(dp, f), errors="ignore").read()
                assert "oracle_py" not in txt and "wbc_oracle" not in txt and "libref_alglib" not in txt, f
