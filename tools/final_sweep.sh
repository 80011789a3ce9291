# end-of-session measurements on one GPU (label $1, default r01s): GPU tests, smoke, default bench (C2 fp32),
# bf16, the other configurations, the reference arm, then -- only after those exited -- the ncu launch list
# and one full capture of the step's kernels
L=${1:-r01s}
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${L}_pytest_gpu.log 2>&1; tail -2 gpurun_out/${L}_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/${L}_smoke.log 2>&1; tail -1 gpurun_out/${L}_smoke.log
python bench.py > gpurun_out/${L}_bench_n1.json 2> gpurun_out/${L}_bench_n1.err
python bench.py --dtype bf16 --steps 500 --no-cpu-baseline > gpurun_out/${L}_bench_n1_bf16.json 2>/dev/null
for c in 3 4 5; do python bench.py --config $c --steps 200 --no-cpu-baseline > gpurun_out/${L}_bench_c$c.json 2>/dev/null; done
python bench.py --config 5 --dtype bf16 --steps 200 --no-cpu-baseline > gpurun_out/${L}_bench_c5_bf16.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${L}_bench_reference.json 2>/dev/null
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${L}_bench_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    if j.get("impl") == "reference":
        print(f.split("/")[-1], j["value"], j["unit"], j.get("cpu_baseline")); continue
    print(f.split("/")[-1], round(j["ms_per_step"], 4), "ms", round(j["value"] / 1e6, 2), "M hm/s step_frac", round(j["roofline"]["step"]["frac"], 3),
          "dominant:", j["roofline"]["kernel"][:40], round(j["roofline"]["frac"], 3),
          {k: (round(v["ms"] * 1e3, 1), round(v["frac"], 3)) for k, v in j["kernels"].items()},
          "e2e", round(j["e2e"]["value"] / 1e6, 2), "cpu", (j.get("cpu_baseline") or {}).get("value"))
PY
K='regex:encode_kernel|decode_expected|oks_loss_fast|finalize_kernel|pack_records|mailbox'
ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/${L}_launches_bench.csv \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph --serial > gpurun_out/${L}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k "$K" -s 30 -c 4 -o gpurun_out/${L}_full -f \
    python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-graph --serial > gpurun_out/${L}_ncu2.log 2>&1
tail -2 gpurun_out/${L}_ncu2.log
