#!/bin/bash
# One-GPU verification + measurement set of a round: GPU tests, smoke, the three bench configurations, the reference arm.
# usage (on the GPU box): bash tools/final_run.sh TAG   -> gpurun_out/TAG_*.json
TAG=${1:-final}
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || tail -c 2000 gpurun_out/${TAG}_bench.err
python bench.py --impl reference > gpurun_out/${TAG}_reference_arm.json 2> gpurun_out/${TAG}_reference_arm.err || tail -c 2000 gpurun_out/${TAG}_reference_arm.err
python bench.py --config infer > gpurun_out/${TAG}_infer.json 2> gpurun_out/${TAG}_infer.err || tail -c 2000 gpurun_out/${TAG}_infer.err
python bench.py --config scaled > gpurun_out/${TAG}_scaled.json 2> gpurun_out/${TAG}_scaled.err || tail -c 2000 gpurun_out/${TAG}_scaled.err
python - "$TAG" <<'PY'
import json, sys
tag = sys.argv[1]
for n in ("bench", "reference_arm", "infer", "scaled"):
    try:
        d = json.load(open(f"gpurun_out/{tag}_{n}.json"))
        gb = d.get("gpu_baseline") or {}
        print(n, round(d["value"], 1), d["unit"], round(d.get("ms_per_step", 0), 3), "ms | e2e", round((d.get("e2e") or {}).get("value", 0), 1),
              "| roofline", round((d.get("roofline") or {}).get("frac", 0) or 0, 3), "| launches", d.get("gpu_launches"), "| clocks", (d.get("clocks") or {}).get("sm_mhz"),
              (d.get("clocks") or {}).get("reasons"), "| stock bf16", round((gb.get("bf16_autocast") or {}).get("value", 0), 1), "| cpu", round((d.get("cpu_baseline") or {}).get("value", 0), 1))
    except Exception as e:
        print(n, "FAILED", e)
PY
