#!/bin/bash
# One-GPU gate, what the driver runs at round end:  gpurun --timeout 500 -- 'bash tools/gpu_final1.sh'
export PYTHONPATH=$PWD
O=gpurun_out/final1
mkdir -p $O
timeout 300 python -m pytest tests -m gpu -q --durations=8 -p no:cacheprovider > $O/pytest.log 2>&1; echo rc=$? >> $O/pytest.log
UCC_B200_RUN_RAB_EMU=1 timeout 60 python -m pytest tests/test_nvl_gpu.py -m gpu -q -k "cl_hier and rab" -p no:cacheprovider > $O/pytest_rab_emu.log 2>&1; echo rc=$? >> $O/pytest_rab_emu.log; tail -3 $O/pytest_rab_emu.log | cut -c1-200
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo rc=$? >> $O/smoke.log
timeout 100 python bench.py --impl reference --steps 10 --warmup 3 --out $O/bench_ref.json > $O/bench_ref.log 2>&1
timeout 120 python bench.py --steps 10 --warmup 3 --out $O/bench_ours.json > $O/bench_ours.log 2>&1
tail -14 $O/pytest.log | cut -c1-300; tail -3 $O/smoke.log
python - <<PY
import json
for f in ("ref", "ours"):
    try:
        d = json.load(open("$O/bench_%s.json" % f))
        print(f, "busbw", d["busbw_per_gpu_GBps"], "us", d["latency_us"], "ok", d["correct"], "e2e us", d["e2e"]["us_per_step"], d["config"]["algorithm"][-40:],
              "bad sweep rows", [(r["bytes"], r.get("ok")) for r in d.get("sweep", []) if not r.get("ok", True)])
    except Exception as e:
        print(f, "no json", e)
PY
