mkdir -p gpurun_out
(time timeout 1500 python -m pytest tests/test_gpu_sharded.py -m gpu -x -q) > gpurun_out/r02m_sharded.log 2>&1; tail -4 gpurun_out/r02m_sharded.log
for n in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 50 --warmup 5 > gpurun_out/r02m_bench_${n}gpu.json 2> gpurun_out/r02m_bench_${n}gpu.err; echo "N=$n rc=$?"
done
python bench.py --steps 50 --warmup 5 > gpurun_out/r02m_bench_1gpu.json 2> gpurun_out/r02m_bench_1gpu.err; echo "N=1 rc=$?"
OMP_NUM_THREADS=1 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r02m_ref.json 2>/dev/null
python - <<'PY'
import json
for n in (1,2,4,8):
    try:
        d=json.loads(open(f"gpurun_out/r02m_bench_{n}gpu.json").read().splitlines()[-1])
    except Exception as e:
        print(n, "unreadable", e); continue
    print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"], d.get("sharded_parity"), d["config"]["best_agent_exchange"])
    for w in d["workloads"]: print("   ", w["name"], w["config"]["agents"], w["value"], w["ms_per_step"], w["e2e"]["value"], w["roofline"]["frac"])
r=json.loads(open("gpurun_out/r02m_ref.json").read()); print("ref", r["value"], r["cpu_baseline"]["cores"])
PY
nproc
