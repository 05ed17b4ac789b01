"""Run bench.py against another build of the library (A/B runs on ONE box, one process per build):
    python tools/probe/bench_with_lib.py build/libscm_prev.so --steps 5 --warmup 3 --no-cpu --no-e2e | python tools/probe/bench_with_lib.py --show
"""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, root)
if sys.argv[1] == "--show":
    d = json.loads(sys.stdin.read().strip().splitlines()[-1])
    st = d["stages_ms"]
    print(f"{d['value'] / 1e6:7.3f} M cells/s  {d['ms_per_step']:8.3f} ms/step  gram {st['gram_syrk_kernel']:8.3f}  adjust {st['adjust']:6.3f}  "
          f"seg {st['seg_colsum']:6.3f}  eig {st['eig']:6.3f}")
    sys.exit(0)
lib = os.path.abspath(sys.argv[1])
sys.argv = [os.path.join(root, "bench.py")] + sys.argv[2:]
from scmerge_b200 import _lib as L
L.LIB_PATH = lib
import bench
bench.main()
