#!/bin/bash
# 1 -> N GPU scaling of the default bench (run through gpurun --gpus N): bash scripts/scale_run.sh <tag> <N...>
TAG=$1; shift
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 > gpurun_out/scale_${TAG}_g$n.json 2> gpurun_out/scale_${TAG}_g$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n \
      bench.py --gpus $n > gpurun_out/scale_${TAG}_g$n.json 2> gpurun_out/scale_${TAG}_g$n.err
  fi
  echo "n=$n rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/scale_${TAG}_g$n.json").read().strip().splitlines()[-1])
    a = d.get("also", {})
    print("  C2 %.4g pairs/s (%.3f ms)  e2e %.4g | cloud %.4g (%.1f ms) | mixed %.4g (%.2f ms) | decay %.4g" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], a["cloud"]["value"], a["cloud"]["ms_per_step"],
        a["mixed"]["value"], a["mixed"]["ms_per_step"], a["decay"]["value"]))
except Exception as e:
    print("  parse failed:", e)
PY
done
