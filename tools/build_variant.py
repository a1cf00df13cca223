"""Build a variant of libb200rec.so with extra nvcc flags into variants/<name>.so (development tool for A/B timing).
usage: python tools/build_variant.py name -DFLAG ..."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from b200rec import build as B
name, flags = sys.argv[1], sys.argv[2:]
out_dir = os.path.join(ROOT, "variants"); os.makedirs(out_dir, exist_ok=True)
obj_dir = os.path.join("/tmp", "b200rec_variant_" + name); os.makedirs(obj_dir, exist_ok=True)
objs = []
procs = []
for src in B._sources():
    obj = os.path.join(obj_dir, src[:-3] + ".o"); objs.append(obj)
    procs.append(subprocess.Popen([B.NVCC, *B.FLAGS, *flags, "-c", os.path.join(B.CSRC, src), "-o", obj]))
for p in procs:
    assert p.wait() == 0
out = os.path.join(out_dir, name + ".so")
subprocess.check_call([B.NVCC, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a"])
print(out)
