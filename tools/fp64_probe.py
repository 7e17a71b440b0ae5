"""FP64 pipe probe table: issue-slot fraction by operand shape and resident warps per SM sub-partition."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from cmtj_b200 import _abi
lib = _abi.load_library()
f = lib.cmtjb_fp64_pipe_probe
f.restype = C.c_int
f.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]
names = {0: "dfma_1reg_2const", 1: "dfma_3reg", 2: "mix_12dfma_3dmul_1dadd_3reg", 3: "rhs_shape_ilp3", 4: "rhs_shape_two_vectors", 5: "dfma_3reg_rotating_operands"}
out = {}
for kind in range(6):
    for w in (1, 2, 3, 4, 5, 6, 8, 12, 16):
        frac, mhz = C.c_double(), C.c_double()
        rc = f(0, kind, w, 100000 if kind not in (3, 4) else 200000, C.byref(frac), C.byref(mhz))
        assert rc == 0, rc
        out.setdefault(names[kind], {})[w] = round(frac.value, 4)
        print(names[kind], w, f"{frac.value:.4f}", f"{mhz.value:.0f} MHz", flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/fp64_probe.json", "w"), indent=1)
