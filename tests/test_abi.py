"""CPU: the C-ABI shared library loads, exports every symbol include/eepacc.h declares, answers the
metadata calls of the CasADi convention without a GPU, and fails loudly (no fallback) on compute."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from eepacc_mpc_casadi_matlab_b200 import api, settings as S, usecase as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    h = open(os.path.join(ROOT, "include", "eepacc.h")).read()
    fams = re.findall(r"^EEPACC_DECLARE\((\w+)\)", h, flags=re.M)
    syms = [f + s for f in fams for s in api.FAMILY_SUFFIXES]
    plain = re.findall(r"^(?:int|long long|double|void\*?|const char\*)\s+\**(eepacc_\w+)\(", h, flags=re.M)
    return fams, syms + plain


def test_header_families_match_binding_table():
    fams, _ = declared_symbols()
    assert sorted(fams) == sorted(api.FAMILIES)


def test_library_exports_every_declared_symbol():
    L = api.load_library()
    _, syms = declared_symbols()
    assert len(syms) > 200
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_pod_layouts_match_the_header():
    L = api.load_library()
    assert L.eepacc_sizeof_params() == C.sizeof(S.Params)
    assert L.eepacc_sizeof_route() == C.sizeof(S.Route)
    assert [L.eepacc_qp_nvar(c) for c in range(4)] == [100, 120, 40, 40]
    pm, pa = S.default_params("mc"), S.default_params("abo")
    assert [L.eepacc_qp_ncon(c, C.byref(pm)) for c in range(4)] == [362, 522, 262, 220]
    assert L.eepacc_qp_ncon(S.AB, C.byref(pa)) == 282


def test_casadi_style_metadata_without_gpu():
    """casadi_c_{n_in,n_out,name_in,sparsity_in,work}_id equivalents (casadi_fun.c:83-120)."""
    L = api.load_library()
    assert L.eepacc_qpsolve_n_in() == 10 and L.eepacc_qpsolve_n_out() == 6
    names = [L.eepacc_qpsolve_name_in(i).decode() for i in range(10)]
    assert names == ["h", "g", "a", "lba", "uba", "lbx", "ubx", "x0", "lam_x0", "lam_a0"]     # conic.hpp:170-208
    assert [L.eepacc_qpsolve_name_out(i).decode() for i in range(4)] == ["x", "cost", "lam_a", "lam_x"]
    sp = L.eepacc_qpbuild_fb_sparsity_out(2)                       # "a": dense 522 x 120 in CCS
    assert (sp[0], sp[1]) == (522, 120) and sp[2] == 0 and sp[2 + 120] == 522 * 120
    assert sp[2 + 121] == 0 and sp[2 + 121 + 521] == 521 and sp[2 + 121 + 522] == 0
    sp = L.eepacc_F_jac_sparsity_out(2)
    assert (sp[0], sp[1]) == (5, 10)
    sz = [C.c_longlong() for _ in range(4)]
    assert L.eepacc_step_fb_work(*[C.byref(s) for s in sz]) == 0
    assert [s.value for s in sz] == [1, 4, 2, 0]
    assert L.eepacc_f_name_in(5) is None


def test_mem_handles_and_loud_failure_without_device():
    import torch
    L = api.load_library()
    mem = L.eepacc_step_fb_checkout()
    assert mem >= 0
    iw = (C.c_longlong * 2)(1, 0)
    rc = L.eepacc_step_fb(None, None, iw, None, mem)
    assert rc != 0 and b"not set up" in L.eepacc_last_error()
    one = (C.c_double * 16)()
    rc = L.eepacc_mem_synth_lead(mem, one, 16, None, 0, one, 10, 4.0)         # scenario synthesis needs a set-up mem too
    assert rc != 0 and b"not set up" in L.eepacc_last_error()
    if not torch.cuda.is_available():
        p, r = S.default_params("mc"), U.default_route()
        rc = L.eepacc_mem_setup(mem, S.FB, C.byref(p), C.byref(r), 1, -1)
        assert rc == 2 and b"no CPU fallback" in L.eepacc_last_error()         # EEPACC_ERR_NO_DEVICE
        with pytest.raises(api.EepaccError):
            api.Engine("fb", p, r, B=1)
    L.eepacc_step_fb_release(mem)
    assert L.eepacc_step_fb(None, None, iw, None, mem) != 0                      # released handle is invalid


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "eepacc_mpc_casadi_matlab_b200")
    for dp_, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp_, f)).read()
                assert "oracle" not in src.replace("qp_oracle.c", "").replace("oracle/", "").lower() or f in ("k3_qp.cu",), f


def test_nlp_header_symbols_and_loud_failure_without_device():
    """include/eepacc_nlp.h: every declared entry point is exported; metadata answers without a GPU; set-up refuses to
    run without a device (no CPU fallback)."""
    import torch
    from eepacc_mpc_casadi_matlab_b200 import nlp
    L = api.load_library()
    h = open(os.path.join(ROOT, "include", "eepacc_nlp.h")).read()
    syms = re.findall(r"^(?:int|long long|double|void|const char\*)\s+\**(eepacc_nlp\w*)\(", h, flags=re.M)
    assert len(syms) >= 16, syms
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    L.eepacc_nlp_name_in.restype = C.c_char_p
    L.eepacc_nlp_name_out.restype = C.c_char_p
    L.eepacc_nlp_nz.restype = C.c_longlong
    assert [L.eepacc_nlp_name_in(i) for i in range(3)] == [b"x0", b"p", None]                  # nlpsol scheme
    assert [L.eepacc_nlp_name_out(i) for i in range(4)] == [b"x", b"f", b"stats", None]
    assert L.eepacc_nlp_nz(870) == 8704                                                         # RunOpt_NLP.m: 4 + 10 N
    assert L.eepacc_nlp_sizeof_tables() == C.sizeof(nlp.NlpTables)
    o = nlp.NlpOpts()
    L.eepacc_nlp_default_opts(C.byref(o))
    assert (o.tol, o.max_iter) == (1e-8, 500) and o.mu_min == 1e-9
    if not torch.cuda.is_available():
        with pytest.raises(api.EepaccError, match="no CPU fallback"):
            nlp.NlpSolver(S.default_params("mc"), U.default_route(), 10)
