# same-box A/B of two builds of the library: sed_crnn_b200/libsedb200_prev.so (SEDB200_LIB_PATH) against the current one
# usage: tools/ab_lib.sh "c2 c1 c5"
F="--warmup 5 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-dropin --no-fixed-global"
for cfg in ${1:-c2 c1 c5}; do for lib in prev cur prev cur; do
  if [ $lib = prev ]; then export SEDB200_LIB_PATH=$PWD/sed_crnn_b200/libsedb200_prev.so; else unset SEDB200_LIB_PATH; fi
  S=20; [ $cfg = c5 ] && S=5
  timeout 300 python bench.py --config $cfg $F --steps $S > gpurun_out/ablib.json 2> gpurun_out/ablib.err
  python - $cfg $lib <<'PY'
import json, sys
cfg, lib = sys.argv[1:3]
try:
    d = json.loads([l for l in open("gpurun_out/ablib.json") if l.startswith("{")][-1])
    print(cfg, lib, "ms_per_step", round(d["ms_per_step"], 4), "sum(phases)", round(sum(d["phases_ms"].values()), 4), "frac", round(d["roofline"]["frac"], 3))
except Exception as e:
    print(cfg, lib, "failed", e, open("gpurun_out/ablib.err").read()[-300:])
PY
done; done
