"""Extract the mesh golden vectors held by the reference's own unit tests into JSON.

Run in the build container (needs /root/reference); the JSON it writes is committed
so the tests never read /root/reference at run time.

Sources (reference @ v0.5.4):
  tests/mesh/cartesianmesh2d_dirichlet.cpp:21-72   (mesh config), :171-284 (coord, dL, UN, pN)
  tests/mesh/cartesianmesh2d_yperiodic.cpp          (same layout, y periodic)
"""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))


def brace_block(text, start):
    """Return the substring of the balanced {...} that opens at or after `start`."""
    i = text.index("{", start)
    depth, j = 0, i
    while True:
        if text[j] == "{":
            depth += 1
        elif text[j] == "}":
            depth -= 1
            if depth == 0:
                return text[i: j + 1]
        j += 1


def parse_nested(block):
    return json.loads(block.replace("{", "[").replace("}", "]"))


def extract(path):
    src = open(path).read()
    out = {"source": os.path.relpath(path, REF)}
    # mesh config: config["mesh"][a]["subDomains"][s][key] = value
    axes = {}
    for m in re.finditer(r'config\["mesh"\]\[(\d)\]\["(direction|start)"\]\s*=\s*"([^"]+)"', src):
        axes.setdefault(int(m.group(1)), {"subDomains": {}})[m.group(2)] = m.group(3)
    for m in re.finditer(r'config\["mesh"\]\[(\d)\]\["subDomains"\]\[(\d)\]\["(end|cells|stretchRatio)"\]\s*=\s*([-0-9.eE]+)', src):
        axes[int(m.group(1))]["subDomains"].setdefault(int(m.group(2)), {})[m.group(3)] = float(m.group(4))
    mesh = []
    for a in sorted(axes):
        subs = [axes[a]["subDomains"][s] for s in sorted(axes[a]["subDomains"])]
        for s in subs:
            s["cells"] = int(s["cells"])
        mesh.append({"direction": axes[a]["direction"], "start": float(axes[a]["start"]), "subDomains": subs})
    out["mesh"] = mesh
    out["coordTrue"] = parse_nested(brace_block(src, src.index("RealVec3D coordTrue")))
    out["dLTrue"] = parse_nested(brace_block(src, src.index("RealVec3D dLTrue")))
    out["UN"] = int(re.search(r"ASSERT_EQ\((\d+), mesh->UN\)", src).group(1))
    out["pN"] = int(re.search(r"ASSERT_EQ\((\d+), mesh->pN\)", src).group(1))
    out["tolerance"] = 1e-12  # ASSERT_NEAR(..., 1e-12) in the reference tests
    return out


if __name__ == "__main__":
    for name, per in (("cartesianmesh2d_dirichlet", [0, 0]), ("cartesianmesh2d_yperiodic", [0, 1])):
        g = extract(os.path.join(REF, "tests", "mesh", name + ".cpp"))
        g["periodic"] = per
        with open(os.path.join(HERE, name + ".json"), "w") as f:
            json.dump(g, f, indent=1)
        print(name, "UN", g["UN"], "pN", g["pN"], "axes", [len(a["subDomains"]) for a in g["mesh"]])
