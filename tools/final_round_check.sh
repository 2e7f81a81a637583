# round-end verification on one B200: GPU test suite, smoke, the headline bench and the supplementary bench lines
set -x
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 400 python bench.py 2>gpurun_out/final_err.log | tail -1 > gpurun_out/final_base.json
timeout 200 python bench.py --batch 512 --no-cpu-baseline 2>>gpurun_out/final_err.log | tail -1 > gpurun_out/final_base_b512.json
timeout 200 python bench.py --lengths alpaca --no-cpu-baseline 2>>gpurun_out/final_err.log | tail -1 > gpurun_out/final_base_alpaca.json
timeout 200 python bench.py --model flan-t5-small --no-cpu-baseline 2>>gpurun_out/final_err.log | tail -1 > gpurun_out/final_small.json
timeout 300 python bench.py --model flan-t5-large --no-cpu-baseline --steps 2 2>>gpurun_out/final_err.log | tail -1 > gpurun_out/final_large.json
timeout 200 python bench.py --dtype fp16 --no-cpu-baseline 2>>gpurun_out/final_err.log | tail -1 > gpurun_out/final_fp16.json
for f in base base_b512 base_alpaca small large fp16; do python - <<PY
import json
try:
    d=json.load(open("gpurun_out/final_$f.json"))
    print("$f", round(d["ms_per_step"],1), "ms", round(d["value"]), "tok/s e2e", round(d["e2e"]["value"]), "roofline", round(d["roofline"]["frac"],3), "decode", round(d["decode_loop"]["ms"],1), round(d["decode_loop"]["frac_of_hbm_peak"],3), "enc", round(d["encoder"]["ms"],1), round(d["encoder"]["frac_of_bf16_sustained"],3), d["clocks"]["reasons"])
except Exception as e:
    print("$f FAILED", e)
PY
done
