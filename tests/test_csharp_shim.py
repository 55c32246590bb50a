"""The C# shim (csharp/*.cs) cannot be compiled here (no dotnet / mono in the image), so it is checked textually against
the header it binds: every [DllImport] name is declared in include/pf_abi.h and exported by libpfasr.so, and every
[StructLayout] mirror lists the C struct's fields in the same order."""
import os
import re

from aliparaformerasr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = open(os.path.join(ROOT, "csharp", "PfAsr.cs"), encoding="utf-8").read()
HDR = open(os.path.join(ROOT, "include", "pf_abi.h"), encoding="utf-8").read()


def _c_fields(name):
    body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (name, name), HDR, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            out.append(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", decl)[-1])
    return out


def _cs_fields(name):
    body = re.search(r"struct %s\b[^{]*\{(.*?)\n    \}" % name, CS, re.S).group(1)
    body = re.sub(r"//.*", "", body)
    out = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl.startswith("public"):
            names = decl.split(None, 2)[2]
            out += [n.strip() for n in names.split(",")]
    return out


def test_dllimports_are_declared_and_exported():
    lib = _lib.load()
    names = re.findall(r"static extern \w+ (pf_\w+)\(", CS)
    assert len(names) >= 25 and len(set(names)) == len(names)
    for n in names:
        assert re.search(r"\b%s\(" % n, HDR), f"{n} is not declared in pf_abi.h"
        assert hasattr(lib, n), f"{n} is not exported by libpfasr.so"
    for other in ("OfflineProjOfCuda.cs", "OnlineRecognizerOfCuda.cs"):
        src = open(os.path.join(ROOT, "csharp", other), encoding="utf-8").read()
        for n in set(re.findall(r"PfAsr\.(pf_\w+)\(", src)):
            assert n in names, f"{other} calls {n}, which PfAsr.cs does not import"


def test_struct_mirrors_list_the_c_fields_in_order():
    for cs, c in (("PfConfig", "pf_config"), ("PfResult", "pf_result"), ("PfOnlineResult", "pf_online_result"), ("PfAudio", "pf_audio"),
                  ("PfTextResult", "pf_text_result")):
        assert _cs_fields(cs) == _c_fields(c), (cs, _cs_fields(cs), _c_fields(c))
    assert int(re.search(r"ABI version (\d+)", CS).group(1)) == int(re.search(r"#define\s+PF_ABI_VERSION\s+(\d+)", HDR).group(1))
