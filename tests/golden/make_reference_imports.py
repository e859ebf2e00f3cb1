"""Generates tests/golden/reference_imports.json: every name the reference scripts import from pytorch3d
(module -> sorted names, with the scripts that use them).  Run in the authoring container, where /root/reference
exists; the test (tests/test_host.py::test_compat_resolves_every_reference_import) only reads the JSON."""
import ast, glob, json, os, re
out = {}
for path in sorted(glob.glob("/root/reference/*.py")):
    src = open(path).read()
    try:
        tree = ast.parse(src)
        nodes = [n for n in ast.walk(tree) if isinstance(n, ast.ImportFrom) and n.module and n.module.startswith("pytorch3d")]
        pairs = [(n.module, a.name) for n in nodes for a in n.names]
    except SyntaxError:
        # batch_rendering_test.py does not parse (upstream syntax error at :258): take its import block textually
        pairs = []
        for m in re.finditer(r"from\s+(pytorch3d[\w.]*)\s+import\s+(\([^)]*\)|[^\n]+)", src):
            names = re.sub(r"[()\\]", " ", m.group(2)).split(",")
            pairs += [(m.group(1), n.split("#")[0].strip()) for n in names if n.split("#")[0].strip()]
    for mod, name in pairs:
        out.setdefault(mod, {}).setdefault(name, []).append(os.path.basename(path))
json.dump({m: {k: sorted(set(v)) for k, v in sorted(d.items())} for m, d in sorted(out.items())},
          open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_imports.json"), "w"), indent=1)
print({m: len(d) for m, d in out.items()})
