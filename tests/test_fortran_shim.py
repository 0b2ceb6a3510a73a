"""integration/*.f90 (the Fortran shim a MagIC maintainer would add) cannot be compiled here -- the image has no Fortran
compiler -- so its C-facing half is checked textually against include/magic_sht.h and the built library: every
bind(C, name=...) must be an exported symbol, the bind(C) derived types must list the members of the C structs in the same
order and with matching types, and module sht must export exactly the public list of sht_native.f90:16-20."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _read(*p):
    return open(os.path.join(ROOT, *p)).read()


def _fortran_type(src, name):
    """[(kind, member), ...] of `type, bind(C) :: name`, continuation lines joined."""
    body = re.search(r"type, bind\(C\) :: %s\n(.*?)end type %s" % (name, name), src, re.S).group(1)
    body = re.sub(r"&\s*\n\s*&", " ", body)
    out = []
    for line in body.splitlines():
        line = line.split("!")[0].strip()
        if not line:
            continue
        decl, names = line.split("::")
        kind = {"integer(c_int)": "int", "real(c_double)": "double", "type(c_ptr)": "ptr"}[decl.strip()]
        out += [(kind, n.strip()) for n in names.split(",") if n.strip()]
    return out


def _c_struct(src, name):
    body = re.search(r"typedef struct \{((?:(?!typedef struct).)*?)\} %s;" % name, src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for stmt in body.split(";"):
        stmt = " ".join(stmt.split())
        if not stmt:
            continue
        m = re.match(r"(const )?(int|double) (.*)", stmt)
        base = m.group(2)
        for item in m.group(3).split(","):
            item = item.strip()
            out.append(("ptr" if item.startswith("*") else base, item.lstrip("*").strip()))
    return out


def test_every_bound_name_is_exported_by_the_library():
    from magic_b200.lib import SYMBOLS
    src = _read("integration", "magic_b200_c.f90")
    names = re.findall(r"bind\(C, name='(\w+)'\)", src)
    assert len(names) >= 30 and len(set(names)) == len(names)
    for n in names:
        assert n == "strlen" or n in SYMBOLS, n
    # what the three shims need is all there: the 17 procedures, the loop, the transposer
    for n in ("magic_scal_to_spat", "magic_toraxi_to_spat", "magic_torpol_to_curl_spat_IC", "magic_rloop_run",
              "magic_rloop_get_br_v_bcs", "magic_transp_lm2r", "magic_transp_unique_id"):
        assert n in names


def test_bind_c_types_mirror_the_c_structs():
    f = _read("integration", "magic_b200_c.f90")
    h = _read("include", "magic_sht.h")
    for name in ("magic_params", "magic_radial", "magic_fields_in", "magic_fields_out", "magic_lm_in", "magic_lm_out"):
        assert _fortran_type(f, name) == _c_struct(h, name), name
    # and the Python mirror agrees with both
    from magic_b200.riter import Params
    assert [n for n, _ in Params._fields_] == [n for _, n in _c_struct(h, "magic_params")]


def test_module_sht_has_the_public_list_of_the_reference():
    src = _read("integration", "sht_cuda.f90")
    public = re.search(r"public :: (initialize_sht.*?)\n\ncontains", src, re.S).group(1)
    public = {n.strip() for n in re.sub(r"&", " ", public).split(",") if n.strip()}
    expected = {"initialize_sht", "finalize_sht", "scal_to_spat", "scal_to_grad_spat", "pol_to_grad_spat", "torpol_to_spat",
                "sphtor_to_spat", "torpol_to_curl_spat_IC", "torpol_to_spat_IC", "torpol_to_dphspat", "pol_to_curlr_spat",
                "torpol_to_curl_spat", "scal_to_SH", "spat_to_qst", "spat_to_sphertor", "axi_to_spat", "toraxi_to_spat"}
    assert public == expected                                            # sht_native.f90:16-20 == shtns.f90:22-26
    for n in expected:
        assert len(re.findall(r"^   subroutine %s\(" % n, src, re.M)) == 1, n
        assert len(re.findall(r"end subroutine %s$" % n, src, re.M)) == 1, n
    # each wrapper forwards to the C entry point of the same name
    for n in expected - {"initialize_sht", "finalize_sht"}:
        assert "magic_%s(sht_h" % n in src, n


def test_overriding_procedures_keep_the_reference_argument_lists():
    """rIteration.f90:34-44 and mpi_transpose.f90:40-52: same dummy names in the same order, no TARGET on dummies."""
    r = _read("integration", "rIter_cuda.f90")
    args = re.search(r"subroutine radialLoop\((.*?)\)\n", r, re.S).group(1)
    args = [a.strip() for a in re.sub(r"&", " ", args).split(",")]
    assert args == ["this", "l_graph", "l_frame", "time", "timeStage", "tscheme", "dtLast", "lTOCalc", "lTONext", "lTONext2",
                    "lHelCalc", "lPowerCalc", "lRmsCalc", "lPressCalc", "lPressNext", "lViscBcCalc", "lFluxProfCalc",
                    "lPerpParCalc", "lGeosCalc", "lHemiCalc", "lPhaseCalc", "l_probe_out", "dsdt", "dwdt", "dzdt", "dpdt",
                    "dxidt", "dphidt", "dbdt", "djdt", "dVxVhLM", "dVxBhLM", "dVSrLM", "dVXirLM", "lorentz_torque_ic",
                    "lorentz_torque_ma", "br_vt_lm_cmb", "br_vp_lm_cmb", "br_vt_lm_icb", "br_vp_lm_icb", "dtrkc", "dthkc"]
    assert "target, intent" not in r
    t = _read("integration", "mpi_transp_cuda.f90")
    for proc in ("create_comm", "destroy_comm", "transp_lm2r", "transp_r2lm"):
        assert re.search(r"procedure :: %s\s+=> %s_cuda" % (proc, proc), t), proc
    assert "arr_LMloc(llm:ulm,1:n_r_max,*)" in t and "arr_Rloc(1:lm_max,nRstart:nRstop,*)" in t


def _crack(path):
    """numpy.f2py's Fortran parser: good enough for modules without type-bound procedures (it returns nothing for the
    reference's own rIter.f90 either), i.e. for magic_b200_c.f90 and sht_cuda.f90."""
    import contextlib
    import io
    import numpy.f2py.crackfortran as cf
    cf.verbose = 0
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf), contextlib.redirect_stderr(buf):
        blocks = cf.crackfortran([path])
    out = {}

    def walk(bs):
        for b in bs:
            if b.get("block") in ("subroutine", "function"):
                out[b["name"]] = list(b.get("args", []))
            walk(b.get("body", []))
    walk(blocks)
    return out


def test_a_fortran_parser_accepts_the_two_plain_modules():
    """Parsed with numpy.f2py.crackfortran: every interface of magic_b200_c.f90 takes as many arguments as the C prototype of
    the same name in include/magic_sht.h, and the 17 procedures of sht_cuda.f90 have the reference's argument counts."""
    h = _read("include", "magic_sht.h")
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = {}
    for m in re.finditer(r"\b(?:int|double|void \*|const char \*|long long)\s*\*?\s*(magic_\w+)\(([^;{]*?)\);", h, re.S):
        args = " ".join(m.group(2).split())
        protos[m.group(1)] = 0 if args in ("", "void") else args.count(",") + 1
    procs = _crack(os.path.join(ROOT, "integration", "magic_b200_c.f90"))
    bound = [n for n in procs if n.startswith("magic_") and n != "magic_check"]
    assert len(bound) >= 30
    for n in bound:
        cname = [k for k in protos if k.lower() == n][0]
        assert len(procs[n]) == protos[cname], (n, procs[n], protos[cname])
    sht = _crack(os.path.join(ROOT, "integration", "sht_cuda.f90"))
    expected = {"initialize_sht": 1, "finalize_sht": 0, "scal_to_spat": 3, "scal_to_grad_spat": 4, "pol_to_grad_spat": 4,
                "torpol_to_spat": 7, "sphtor_to_spat": 5, "torpol_to_curl_spat_ic": 9, "torpol_to_spat_ic": 8,
                "torpol_to_dphspat": 5, "pol_to_curlr_spat": 3, "torpol_to_curl_spat": 9, "scal_to_sh": 3, "spat_to_qst": 7,
                "spat_to_sphertor": 5, "axi_to_spat": 2, "toraxi_to_spat": 4}        # sht_native.f90:24-405
    assert {k: len(v) for k, v in sht.items()} == expected


def test_log_step_batches_are_bound():
    names = re.findall(r"bind\(C, name='(\w+)'\)", _read("integration", "magic_b200_c.f90"))
    for n in ("magic_rloop_diagnostics", "magic_rloop_dtb", "magic_rloop_to_next", "magic_rloop_to", "magic_rloop_rms_keep",
              "magic_rloop_rms", "magic_rloop_pin_host"):
        assert n in names, n
    src = _read("integration", "rIter_cuda.f90")
    for n in ("magic_rloop_diagnostics", "magic_rloop_dtb", "magic_rloop_to_next", "magic_rloop_to", "magic_rloop_rms_keep",
              "magic_rloop_rms"):
        assert re.search(r"\b%s\(" % n, src), n


def test_names_imported_from_the_reference_exist_there():
    """No Fortran compiler here: at least every entity the shims `use` from a reference module must occur in that module's
    source (catches misspelt or renamed imports).  Needs the reference tree; skipped where it is absent (GPU box)."""
    import glob
    import pytest
    ref = "/root/reference/src"
    if not os.path.isdir(ref):
        pytest.skip("reference tree not present")
    modules = {}
    for path in glob.glob(os.path.join(ref, "*.f90")) + glob.glob(os.path.join(ref, "*.F90")):
        text = open(path, errors="replace").read()
        for m in re.findall(r"^\s*module\s+(\w+)\s*$", text, re.M | re.I):
            modules.setdefault(m.lower(), "")
            modules[m.lower()] += text.lower()     # a module may have several flavours (sht_native / shtns, fft variants)
    own = {"magic_b200_c", "sht", "mpi_transp_cuda_mod", "riter_cuda_mod", "iso_c_binding"}
    checked = 0
    for f in ("rIter_cuda.f90", "mpi_transp_cuda.f90", "sht_cuda.f90"):
        src = re.sub(r"&\s*\n\s*&?", " ", _read("integration", f))
        src = "\n".join(l.split("!")[0] for l in src.splitlines())
        for mod, names in re.findall(r"^\s*use\s+(\w+)\s*,\s*only\s*:\s*(.*)$", src, re.M | re.I):
            if mod.lower() in own:
                continue
            assert mod.lower() in modules, f"{f}: module {mod} is not in the reference"
            for n in names.split(","):
                n = n.split("=>")[-1].strip()
                if n:
                    assert re.search(r"\b%s\b" % re.escape(n.lower()), modules[mod.lower()]), f"{f}: {mod} has no {n}"
                    checked += 1
    assert checked > 150


def test_bind_c_interfaces_have_the_arity_of_the_c_prototypes():
    """Every bind(C) function interface lists as many dummies as its C prototype has parameters, and every call of a bound
    function in the shims passes that many actual arguments."""
    h = re.sub(r"/\*.*?\*/", "", _read("include", "magic_sht.h"), flags=re.S)
    cproto = {}
    for m in re.finditer(r"\b(?:int|void\s*\*|long long|double|const char\s*\*)\s*(magic_\w+)\s*\(([^;]*?)\)\s*;", h, re.S):
        cproto[m.group(1)] = [p for p in (q.strip() for q in m.group(2).replace("\n", " ").split(",")) if p and p != "void"]
    iface = re.sub(r"&\s*\n\s*&?", " ", _read("integration", "magic_b200_c.f90"))
    arity = {}
    for m in re.finditer(r"function\s+(magic_\w+)\s*\(([^)]*)\)\s*bind\(C,\s*name='(\w+)'\)", iface, re.I):
        n = len([a for a in m.group(2).split(",") if a.strip()])
        assert m.group(3) in cproto, m.group(3)
        assert n == len(cproto[m.group(3)]), (m.group(3), n, cproto[m.group(3)])
        arity[m.group(1).lower()] = n
    assert len(arity) >= 35
    calls = 0
    for f in ("rIter_cuda.f90", "mpi_transp_cuda.f90", "sht_cuda.f90"):
        src = "\n".join(re.sub(r"'[^']*'", "''", l).split("!")[0] for l in _read("integration", f).splitlines())
        src = re.sub(r"&\s*\n\s*&?", " ", src)
        for m in re.finditer(r"\b(magic_\w+)\s*\(", src):
            name = m.group(1).lower()
            if name not in arity:
                continue
            depth, j, args, cur = 1, m.end(), 0, ""
            while depth > 0:
                ch = src[j]
                depth += ch == "("
                depth -= ch == ")"
                if (ch == "," and depth == 1) or depth == 0:
                    args += bool(cur.strip())
                    cur = ""
                else:
                    cur += ch
                j += 1
            assert args == arity[name], (f, name, args, arity[name])
            calls += 1
    assert calls >= 40


def test_free_form_line_length():
    """gfortran truncates free-form lines at 132 characters unless told otherwise; the shims stay below."""
    for f in ("rIter_cuda.f90", "mpi_transp_cuda.f90", "sht_cuda.f90", "magic_b200_c.f90"):
        for n, line in enumerate(_read("integration", f).splitlines(), 1):
            assert len(line) <= 132, (f, n, len(line))
