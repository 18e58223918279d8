import ctypes as C, os, sys
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tools")
from mug_diffusion_b200 import lib as L_
lib = L_.load()
lib.mugd_debug_set_tc_plain_store.argtypes = [C.c_int]
lib.mugd_debug_set_tc_plain_store(int(os.environ.get("PLAIN", "0")))
sys.argv = ["bench_gemm.py"]
import runpy
runpy.run_path("/root/repo/tools/bench_gemm.py", run_name="__main__")
