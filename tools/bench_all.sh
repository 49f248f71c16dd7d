# every BASELINE config + variants on one GPU -> gpurun_out/bench_r2_<name>.json
for c in "" "--march voxel" "--config 1" "--config 3" "--config 4" "--config 5" "--dd"; do
  n=$(echo "$c" | tr -d " -"); n=${n:-config2}
  timeout 900 python bench.py --steps 50 --warmup 5 $c > gpurun_out/bench_r2_$n.json 2> gpurun_out/bench_r2_$n.err || tail -5 gpurun_out/bench_r2_$n.err
done
PAGNERF_L2_WINDOW=1 timeout 600 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r2_l2window.json 2> gpurun_out/bench_r2_l2window.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r2_reference.json 2> gpurun_out/bench_r2_reference.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_r2_*.json")):
    try:
        j = json.load(open(f))
        print(f.split("bench_r2_")[1][:-5], round(j["ms_per_step"], 4), round(j["value"], 1), j["unit"], "e2e", round(j["e2e"]["value"], 1), "roof", (j.get("roofline") or {}).get("kernel"), round((j.get("roofline") or {}).get("frac") or 0, 3))
    except Exception as e:
        print(f, "ERR", e)
PY
