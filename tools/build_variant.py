"""Build an experimental variant of libnerftiny.so for same-box A/B timing: recompiles the named sources with extra -D flags
and links them with the regular objects into nerf_tiny_b200/build/variants/lib<name>.so (select it with NT_LIB_PATH).
    python tools/build_variant.py <name> "<-DFLAG ...>" mlp_tc.cu [more.cu]"""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nerf_tiny_b200 import build as B
name, flags, srcs = sys.argv[1], sys.argv[2].split(), sys.argv[3:]
B.build()
vdir = os.path.join(B.HERE, "build", "variants")
os.makedirs(vdir, exist_ok=True)
objs = []
procs = []
for src in B.SOURCES:
    obj = os.path.join(B.HERE, "build", src.replace(".cu", ".o"))
    if src in srcs:
        obj = os.path.join(vdir, f"{name}_{src.replace('.cu', '.o')}")
        procs.append(subprocess.Popen([B.NVCC, *[f for f in B.FLAGS if f not in ("-Xptxas", "-v")], *flags, "-c", os.path.join(B.CSRC, src), "-o", obj]))
    objs.append(obj)
for p in procs:
    assert p.wait() == 0
out = os.path.join(vdir, f"lib{name}.so")
subprocess.check_call([B.NVCC, "-shared", "-o", out, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcuda"])
print(out)
