"""Extracts what a UG4 registration source registers -- class name, typedefs (T / TBase), constructors, methods with their
overload signatures, class group -- from the text of the registration functions. Run on the reference's registration files it
produces tests/golden/registry_names.json (committed: /root/reference does not travel); tests/test_binding.py runs the same
extraction on include/register_navier_stokes_b200.cpp and on the mock registry's run-time dump and compares.

    python tests/golden/make_registry_names.py            # regenerate the fixture from /root/reference
"""
import json
import os
import re
import sys

REF = "/root/reference"
REF_FILES = ["register_navier_stokes.cpp", "incompressible/incompressible_navier_stokes_plugin.cpp",
             "incompressible/fv1/register_fv1.cpp", "incompressible/fvcr/register_fvcr.cpp"]
# the classes of the assembly path; the data exports of IncompressibleNavierStokesBase are outside it
CLASSES = ["NavierStokesBase", "IncompressibleNavierStokesBase", "NavierStokesFV1", "NavierStokesFVCR",
           "INavierStokesUpwind", "NavierStokesNoUpwind", "NavierStokesFullUpwind", "NavierStokesSkewedUpwind",
           "NavierStokesLinearProfileSkewedUpwind", "NavierStokesPositiveUpwind", "NavierStokesRegularUpwind",
           "INavierStokesFV1Stabilization", "INavierStokesSRFV1Stabilization", "NavierStokesFIELDSStabilization",
           "NavierStokesFLOWStabilization", "NavierStokesFV1WithoutStabilization"]
NOT_PROVIDED = {"IncompressibleNavierStokesBase": ["velocity", "velocity_ip", "velocity_grad", "pressure", "pressure_grad"]}


def _norm(s):
    return re.sub(r"\s+", "", s)


def extract(text):
    """{class: {typedefs, ctors, methods, group, smart}} for every registration block `string name = string("X").append(suffix)`"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    text = re.sub(r"//[^\n]*", "", text)
    out = {}
    for m in re.finditer(r'string\s+name\s*=\s*string\("(\w+)"\)\.append\(suffix\);', text):
        cls = m.group(1)
        head = text[max(0, m.start() - 400):m.start()]
        head = head[head.rfind("{"):]
        typedefs = [_norm(t) for t in re.findall(r"typedef\s+([^;]+?)\s+(?:T|TBase|TBase2);", head)]
        tail = text[m.end():]
        end = tail.find("reg.add_class_to_group")
        chain = tail[:end]
        g = re.search(r'reg\.add_class_to_group\(name,\s*"(\w+)"', tail)
        ctors = [_norm(c) for c in re.findall(r"add_constructor\s*<([^;]*?)>\s*\(", chain)]
        ctors += ["void(*)()"] * len(re.findall(r"\.add_constructor\(\)", chain))
        methods = []
        for mm in re.finditer(r'\.add_method\("(\w+)",\s*(static_cast<(.*?)>\s*\(&T::\w+\)|&T::\w+)', chain, flags=re.S):
            methods.append([mm.group(1), _norm(mm.group(3)) if mm.group(3) else ""])
        out[cls] = {"typedefs": typedefs, "ctors": ctors, "methods": methods, "group": g.group(1) if g else None,
                    "smart": "set_construct_as_smart_pointer(true)" in _norm(chain)}
    return out


def reference_registry():
    reg = {}
    for f in REF_FILES:
        for k, v in extract(open(os.path.join(REF, f)).read()).items():
            if k in CLASSES and k not in reg:
                reg[k] = v
    for cls, names in NOT_PROVIDED.items():
        reg[cls]["methods"] = [m for m in reg[cls]["methods"] if m[0] not in names]
    return reg


if __name__ == "__main__":
    reg = reference_registry()
    missing = [c for c in CLASSES if c not in reg]
    assert not missing, missing
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "registry_names.json")
    json.dump({"source": REF_FILES, "not_provided": NOT_PROVIDED, "classes": reg}, open(out, "w"), indent=1, sort_keys=True)
    print("wrote", out, len(reg), "classes")
