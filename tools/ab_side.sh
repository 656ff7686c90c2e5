# A/B of the side-stream weight gradients x halo boxes (run under gpurun, 1 GPU)
F="--steps 20 --warmup 5 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-dropin --no-fixed-global"
for h in 0 1; do for v in 0 2 3; do
  SEDB200_CONV_HALO=$h SEDB200_WGRAD_SIDE=$v timeout 300 python bench.py $F > gpurun_out/ab_${h}_$v.json 2> gpurun_out/ab_${h}_$v.err; echo "halo=$h side=$v rc=$?"
done; done
python - <<'PY'
import json
for h in (0, 1):
  for v in (0, 2, 3):
    try:
        d = json.loads([l for l in open(f"gpurun_out/ab_{h}_{v}.json") if l.startswith("{")][-1])
        print("halo", h, "side", v, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"] / 1e6, 3), "sum(phases)", round(sum(d["phases_ms"].values()), 4))
    except Exception as e:
        print(h, v, "failed", e)
PY
