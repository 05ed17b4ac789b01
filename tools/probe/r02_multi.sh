#!/bin/bash
# Multi-GPU bench line under torchrun: bash tools/probe/r02_multi.sh N [extra bench flags]
cd "${GRAFT_REPO_ROOT:-/root/repo}"
N=$1; shift
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29711 bench.py --gpus $N --steps 10 --warmup 3 "$@" \
    > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo rc=$?
python - <<P
import json
d=json.load(open("gpurun_out/r02_bench_n$N.json"))
print("value", d["value"], "ms", d["ms_per_step"], d["scaling"]); print(d["stages_ms"]); print("parity", d["parity"]["ok"], d["parity"]["rel_frobenius"], d["parity"]["identical_fullalpha_across_ranks"])
print("e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
for k,v in d["extra"].items(): print(k, {a:b for a,b in v.items() if a in ("ms_per_step","value","error","gram_tflops_per_gpu")}, v.get("stages_ms"))
P
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/r02_bench_n$N.err | tail -5
