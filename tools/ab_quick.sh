# parity of the tensor-core kernels + one quick bench line (run under gpurun, 1 GPU)
F="--steps 20 --warmup 5 --no-logmel --no-other-configs --no-cpu-baseline --no-library-baseline --no-dropin --no-fixed-global"
timeout 1200 python -m pytest tests/test_tc_conv_gpu.py tests/test_tc_gemm_gpu.py tests/test_crnn_gpu.py tests/test_fullsize_gpu.py tests/test_dropin_gpu.py -x -q > gpurun_out/quick_tests.log 2>&1
echo "tests rc=$?"; tail -4 gpurun_out/quick_tests.log
for h in 0 1; do
SEDB200_CONV_HALO=$h timeout 300 python bench.py $F > gpurun_out/quick_$h.json 2> gpurun_out/quick_$h.err; echo "bench halo=$h rc=$?"
python - $h <<'PY'
import json, sys
h = sys.argv[1]
d = json.loads([l for l in open(f"gpurun_out/quick_{h}.json") if l.startswith("{")][-1])
ph = d["phases_ms"]
print("halo", h, "ms_per_step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"] / 1e6, 3), "frac", round(d["roofline"]["frac"], 3), "sum(phases)", round(sum(ph.values()), 4))
print({k: v for k, v in ph.items() if k.startswith("conv") or "proj" in k or "dx" in k})
PY
done
