"""Build a variant of libcpnest_b200.so with extra -D flags (kernel experiments): only api.cu and the iid-Gaussian unit
are compiled with the flags, the other objects are reused.  usage: python scripts/build_variant.py NAME -DFOO [-DBAR ...]
-> cpnest_b200/lib_exp/NAME/libcpnest_b200.so (select with CPNEST_B200_LIB)."""
import os, subprocess, sys, concurrent.futures as cf
sys.path.insert(0, ".")
from cpnest_b200 import build as B
name, defs = sys.argv[1], sys.argv[2:]
B.build()
out = os.path.join(B.HERE, "lib_exp", name)
os.makedirs(out, exist_ok=True)
units = ["api.cu", "model_gauss_iid.cu"]
def comp(u):
    o = os.path.join(out, u[:-3] + ".o")
    r = subprocess.run([B.NVCC] + B.FLAGS + defs + ["-c", os.path.join(B.CSRC, u), "-o", o], capture_output=True, text=True)
    if r.returncode: raise RuntimeError(r.stderr)
    return o
with cf.ThreadPoolExecutor(2) as ex:
    objs = list(ex.map(comp, units))
others = [os.path.join(B.OBJ, f) for f in sorted(os.listdir(B.OBJ)) if f.endswith(".o") and f[:-2] + ".cu" not in units]
lib = os.path.join(out, "libcpnest_b200.so")
r = subprocess.run([B.NVCC, "-shared", "-o", lib] + objs + others + ["-ldl"], capture_output=True, text=True)
if r.returncode: raise RuntimeError(r.stderr)
print("wrote", lib)
