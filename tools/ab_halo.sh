# A/B of conv_tc_kernel's halo mode (run under gpurun, 1 GPU): conv / parity tests, then the same bench line with the
# halo boxes off and on
F="--steps 20 --warmup 5 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-dropin --no-fixed-global"
timeout 900 python -m pytest tests/test_tc_conv_gpu.py tests/test_crnn_gpu.py tests/test_fullsize_gpu.py -x -q > gpurun_out/halo_tests.log 2>&1
echo "tests rc=$?"; tail -5 gpurun_out/halo_tests.log
SEDB200_CONV_HALO=0 timeout 300 python bench.py $F > gpurun_out/halo_off.json 2> gpurun_out/halo_off.err; echo "off rc=$?"
timeout 300 python bench.py $F > gpurun_out/halo_on.json 2> gpurun_out/halo_on.err; echo "on rc=$?"
python - <<'PY'
import json
for f in ("halo_off", "halo_on"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][-1])
        ph = d["phases_ms"]
        print(f, d["ms_per_step"], {k: ph[k] for k in ph if k.startswith("conv1") or k.startswith("conv2")}, d["roofline"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
