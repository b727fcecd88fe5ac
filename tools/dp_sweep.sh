#!/bin/bash
# usage: tools/dp_sweep.sh N  -- bench.py --gpus N under a few settings (one summary line each)
N=$1
run() {
  name=$1; shift
  env "$@" timeout -k 10 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/dp_$name.json 2> gpurun_out/dp_$name.err
  python - "$name" <<'PY'
import json, sys
name = sys.argv[1]
try:
    d = json.load(open(f"gpurun_out/dp_{name}.json"))
    b = d["breakdown"]
    print(name, "value", round(d["value"]), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"]), "ffn ms", round(b["ffn"]["ms_per_step"], 4),
          "dp_check", d.get("dp_grad_check", {}).get("rel_err_sharded_vs_global_batch"))
except Exception as e:
    print(name, "FAILED", e)
    print(open(f"gpurun_out/dp_{name}.err").read()[-1500:])
PY
}
if [ "$2" = "nvls" ]; then
run nvls_r4 NCCL_ALGO=NVLS KIT_COLLECTIVE_SMS=4
run nvls_r8 NCCL_ALGO=NVLS KIT_COLLECTIVE_SMS=8
exit 0
fi
if [ "$2" = "b3" ]; then
run b3_n$N KIT_BUCKET_LAYERS=3
exit 0
fi
if [ "$2" = "train" ]; then
run train_n$N
exit 0
fi
if [ "$2" = "final" ]; then
run final_r8 KIT_COLLECTIVE_SMS=8
run final_r16 KIT_COLLECTIVE_SMS=16
env timeout -k 10 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --config scaled > gpurun_out/r02_scaled_dp$N.json 2> gpurun_out/r02_scaled_dp$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_scaled_dp$N.json')); print('scaled', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -c 1500 gpurun_out/r02_scaled_dp$N.err
exit 0
fi
if [ "$2" = "sg" ]; then
run chain KIT_DP_SINGLE_GRAPH=0
run single KIT_DP_SINGLE_GRAPH=1
exit 0
fi
if [ "$2" = "v2" ]; then
run bf16_b0 KIT_BUCKET_LAYERS=0
run bf16_b2 KIT_BUCKET_LAYERS=2
run fp32_b0 KIT_BUCKET_LAYERS=0 KIT_DP_COMPRESS=none
run bf16_b0_r16 KIT_BUCKET_LAYERS=0 KIT_COLLECTIVE_SMS=16
exit 0
fi
if [ "$2" = "short" ]; then
run r8_b0 KIT_BUCKET_LAYERS=0
run r16_b0 KIT_BUCKET_LAYERS=0 KIT_COLLECTIVE_SMS=16
run r8_b2 KIT_BUCKET_LAYERS=2
exit 0
fi
run r8_b2 KIT_BUCKET_LAYERS=2
run r8_b3 KIT_BUCKET_LAYERS=3
run r8_b0 KIT_BUCKET_LAYERS=0
run r16_b2 KIT_BUCKET_LAYERS=2 KIT_COLLECTIVE_SMS=16
run r4_b2 KIT_BUCKET_LAYERS=2 KIT_COLLECTIVE_SMS=4
run r0_b2 KIT_BUCKET_LAYERS=2 KIT_SM_RESERVE=0
