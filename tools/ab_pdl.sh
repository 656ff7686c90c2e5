F="--steps 20 --warmup 5 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-fixed-global"
timeout 900 python -m pytest tests/test_crnn_gpu.py tests/test_dropin_gpu.py tests/test_p2p_gpu.py -x -q > gpurun_out/pdl3_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/pdl3_tests.log
i=0
for cfg in c2 fork; do for extra in "" "--no-cuda-graph"; do for v in 0 1; do
  i=$((i+1))
  SEDB200_PDL=$v timeout 300 python bench.py --config $cfg $F $extra > gpurun_out/pdl3_$i.json 2> gpurun_out/pdl3_$i.err
  python - "$i" "$cfg" "$v" "$extra" <<'PY'
import json, sys
i, cfg, v, extra = sys.argv[1:5]
try:
    d = json.loads([l for l in open(f"gpurun_out/pdl3_{i}.json") if l.startswith("{")][-1])
    di = d.get("dropin_e2e", {})
    print(f"cfg={cfg} graph={'no' if extra else 'yes'} pdl={v} ms_per_step {d['ms_per_step']:.4f} e2e {d['e2e']['value']/1e6:.3f} dropin", {k: {kk: round(vv['ms_per_step'], 3) for kk, vv in di[k].items() if isinstance(vv, dict)} for k in di})
except Exception as e:
    print(i, "failed", e)
PY
done; done; done
