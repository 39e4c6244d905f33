#!/bin/bash
# Round 2, fifteenth GPU visit (8 GPUs): the scaling line of the final tree at N = 8, launched like the driver does (both modes on one index set).
tag=${1:-r2o}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_film_io.py -m gpu -q 2>&1 | tail -1
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-path-tracer > gpurun_out/${tag}_n1.json 2> gpurun_out/${tag}_n1.err
python -c "
import json
try:
    d = json.load(open('gpurun_out/${tag}_n1.json')); print('n1', round(d['value'], 3), 'e2e', round(d['e2e']['value'], 3), d['clocks'])
except Exception as e: print('n1 failed', e)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --parallelism iteration > gpurun_out/${tag}_n8.json 2> gpurun_out/${tag}_n8.err
python - <<P
import json
try:
    d = json.load(open("gpurun_out/${tag}_n8.json"))
    print("clocks", d["clocks"])
    print("n8", d["config"]["mode"], round(d["value"], 3), "Msamples/s e2e", round(d["e2e"]["value"], 3), "ms/step", round(d["ms_per_step"], 2), {k: (round(v["value"], 2), round(v["e2e"], 2)) for k, v in d["modes"].items()})
except Exception as e:
    print("n8 failed", e)
P
tail -3 gpurun_out/${tag}_n8.err
exit 0
