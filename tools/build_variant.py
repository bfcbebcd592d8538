"""Build a VARIANT of csrc/librtb200.so with extra -D flags (kernel experiments), selected at run time with
RTB200_LIB=<path>.  usage: python tools/build_variant.py <name> [-DWF_MIN_CTAS=3 ...] -> tools/_variants/librtb200_<name>.so"""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
b = importlib.import_module("html5-canvas-raytracer_b200.build")
name, defs = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "tools", "_variants")
obj_dir = os.path.join(out_dir, "_obj_" + name)
os.makedirs(obj_dir, exist_ok=True)
objs = []
procs = []
for src, extra in b.SOURCES.items():
    o = os.path.join(obj_dir, src.replace(".cu", ".o"))
    objs.append(o)
    procs.append(subprocess.Popen([b._nvcc(), *b.ARCH, *b.COMMON, *extra, *defs, "-c", os.path.join(b.CSRC, src), "-o", o]))
for p in procs:
    if p.wait() != 0:
        raise SystemExit("compile failed")
lib = os.path.join(out_dir, f"librtb200_{name}.so")
subprocess.check_call([b._nvcc(), *b.ARCH, "-shared", "-o", lib, *objs, "-ldl"])
print(lib)
