# usage: tools/ab_env.sh "<bench args>" VAR "v1 v2 ..."   -- the same bench line under each value of an environment knob
F="--steps 10 --warmup 5 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-dropin --no-fixed-global"
for v in $3; do
  env $2=$v timeout 300 python bench.py $1 $F > gpurun_out/abenv_$v.json 2> gpurun_out/abenv_$v.err
  python - "$2" "$v" <<'PY'
import json, sys
k, v = sys.argv[1:3]
try:
    d = json.loads([l for l in open(f"gpurun_out/abenv_{v}.json") if l.startswith("{")][-1])
    ph = d["phases_ms"]
    print(f"{k}={v} ms_per_step {d['ms_per_step']:.4f} sum(phases) {sum(ph.values()):.4f}", {a: ph[a] for a in ph if "wgrad" in a or "bwd_sums" in a or "bwd_dy" in a or "dgrad" in a or "bwd_lean" in a})
except Exception as e:
    print(v, "failed", e, open(f"gpurun_out/abenv_{v}.err").read()[-300:])
PY
done
