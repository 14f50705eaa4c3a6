"""Records the Python surface the reference's SWIG interface declares (python/constvplugin.i) as a small
fixture, so that tests can compare this repo's constvplugin.i and the ctypes constvplugin module with it on
machines where /root/reference does not exist.

    python tests/golden/make_swig_surface.py   ->  tests/golden/reference_swig_surface.json
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def swig_surface(text: str) -> dict:
    """{"module", "templates": {name: type}, "unit_getters": {"Class::getter": "unit expression"},
    "classes": {name: {"base", "constructor": [[type, name, default]...], "methods": {name: [ret, [arg types]]}}}}"""
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    out = {"module": re.search(r"%module\s+(\w+)", text).group(1), "templates": {}, "unit_getters": {}, "classes": {}}
    for name, typ in re.findall(r"%template\((\w+)\)\s*([\w<>:]+)\s*;", text):
        out["templates"][name] = typ
    for cls, fn, body in re.findall(r"%pythonappend\s+OpenMM::(\w+)::(\w+)\(\)\s*const\s*%\{(.*?)%\}", text, flags=re.S):
        m = re.search(r"unit\.Quantity\(val,\s*(.*?)\)\s*(?:if|$)", body.strip(), flags=re.S)
        out["unit_getters"][f"{cls}::{fn}"] = re.sub(r"\s+", "", m.group(1)) if m else re.sub(r"\s+", "", body)
    for cls, base, body in re.findall(r"class\s+(\w+)\s*:\s*public\s+(\w+)\s*\{(.*?)\};", text, flags=re.S):
        info = {"base": base, "constructor": [], "methods": {}}
        for line in body.split(";"):
            line = " ".join(line.replace("public:", "").split())
            if not line:
                continue
            m = re.match(r"(?:virtual\s+)?(?:(\w[\w\s\*&:<>]*?)\s+)?(\w+)\s*\((.*?)\)\s*(const)?$", line)
            if not m:
                continue
            ret, name, args = m.group(1), m.group(2), m.group(3).strip()
            parsed = []
            for a in filter(None, [x.strip() for x in args.split(",")]):
                am = re.match(r"(.+?)\s+(\w+)\s*(?:=\s*(.+))?$", a)
                parsed.append([am.group(1), am.group(2), am.group(3)])
            if name == cls:
                info["constructor"] = parsed
            else:
                info["methods"][name] = [ret, [p[0] for p in parsed]]
        out["classes"][cls] = info
    return out


if __name__ == "__main__":
    src = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/python/constvplugin.i"
    surface = swig_surface(open(src).read())
    with open(os.path.join(HERE, "reference_swig_surface.json"), "w") as fh:
        json.dump(surface, fh, indent=1, sort_keys=True)
    print(json.dumps(surface, indent=1, sort_keys=True))
