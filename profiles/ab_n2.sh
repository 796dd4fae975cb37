mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 2 --steps 20 --warmup 5 --no-overhead --no-parity > gpurun_out/ab_$tag.log 2>&1; python - <<PY
import json
txt=open("gpurun_out/ab_$tag.log").read().strip().splitlines()
js=[l for l in txt if l.startswith("{")]
d=json.loads(js[-1]) if js else None
print("$tag", d and (round(d["ms_per_step"],4), round(d["sustained"]["ms_per_step"],4), {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items() if k in ("prepare","bands","k3a","k4")}))
PY
}
run base TML_DUMMY=1
run csum0 TML_XS_K3A_CSUM=0
run cap8 TML_K4_CTAS=8
run cap4 TML_K4_CTAS=4
run csum0cap8 TML_XS_K3A_CSUM=0 TML_K4_CTAS=8
