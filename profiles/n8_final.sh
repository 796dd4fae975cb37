mkdir -p gpurun_out
NP=8
python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29531 tests/multi_gpu_check.py > gpurun_out/r02_multi_gpu_check_n8.log 2>&1; echo "mgc rc=$?"; tail -3 gpurun_out/r02_multi_gpu_check_n8.log
TML_TIMELINE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NP --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus $NP --steps 100 --warmup 5 --no-overhead --no-cpu-baseline > gpurun_out/r02_bench_n8.log 2>&1
python - <<PY
import json,re,statistics as st
txt=open("gpurun_out/r02_bench_n8.log").read().strip().splitlines()
js=[l for l in txt if l.startswith("{")]
d=json.loads(js[-1]) if js else None
print("n8", d and (round(d["ms_per_step"],4), round(d["sustained"]["ms_per_step"],4), round(d["value"],1), {k:round(v,3) for k,v in d["roofline"]["stage_ms"].items() if k in ("prepare","bands","k3a","k4")}, (d.get("parity") or {}).get("ok"), d["e2e"]["value"]))
rows=[[float(x) for x in re.findall(r"-?\d+\.\d+",l)] for l in txt if l.startswith("[tml timeline]")]
rows=[r for r in rows if r[4]<0.3][5:105]
if rows: print("    ", dict(zip(["k3a_end","k4_launch","k3e_end","k4_end","prepare","align","reduce","bands"],[round(st.median(c),3) for c in zip(*rows)])))
PY
