mkdir -p gpurun_out
NP=${NP:-2}
run() { tag=$1; shift; env TML_TIMELINE=1 "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $NP --steps 30 --warmup 5 --no-overhead --no-cpu-baseline $PAR > gpurun_out/ab_$tag.log 2>&1; python - <<PY
import json,re,statistics as st
txt=open("gpurun_out/ab_$tag.log").read().strip().splitlines()
js=[l for l in txt if l.startswith("{")]
d=json.loads(js[-1]) if js else None
print("$tag", d and (round(d["ms_per_step"],4), round(d["sustained"]["ms_per_step"],4), {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items() if k in ("prepare","bands","k3a","k4")}, (d.get("parity") or {}).get("ok")))
rows=[[float(x) for x in re.findall(r"-?\d+\.\d+",l)] for l in txt if l.startswith("[tml timeline]")]
rows=[r for r in rows if r[4]<0.26][5:35]
if rows: print("    ", dict(zip(["k3a_end","k4_launch","k3e_end","k4_end","prepare","align","reduce","bands"],[round(st.median(c),3) for c in zip(*rows)])))
PY
}
PAR=""
run base TML_DUMMY=1
PAR="--no-parity"
run static TML_K4_TICKET=0
run k8 TML_K4_CTAS=8
run c5 TML_XS_COMPOSE_CTAS=5
run c5k8 TML_XS_COMPOSE_CTAS=5 TML_K4_CTAS=8
run c4k8 TML_XS_COMPOSE_CTAS=4 TML_K4_CTAS=8
