# end-of-round measurements on one GPU: default bench (C2 fp32), bf16, the other configurations
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/r01n_bench_n1.json 2> gpurun_out/r01n_bench_n1.err
python bench.py --dtype bf16 --steps 500 --no-cpu-baseline > gpurun_out/r01n_bench_n1_bf16.json 2>/dev/null
for c in 3 4 5; do python bench.py --config $c --steps 200 --no-cpu-baseline > gpurun_out/r01n_bench_c$c.json 2>/dev/null; done
python bench.py --config 5 --dtype bf16 --steps 200 --no-cpu-baseline > gpurun_out/r01n_bench_c5_bf16.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r01n_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    print(f.split("/")[-1], round(j["ms_per_step"], 4), "ms", round(j["value"] / 1e6, 2), "M hm/s step_frac", round(j["roofline"]["step"]["frac"], 3),
          {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in j["kernels"].items()},
          "e2e", round(j["e2e"]["value"] / 1e6, 2), "cpu", (j.get("cpu_baseline") or {}).get("value"))
PY
