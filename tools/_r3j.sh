timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_pytest.log 2>&1; tail -5 gpurun_out/r3j_pytest.log
B="timeout 300 python bench.py --steps 200 --warmup 5 --no-extras --no-cpu-baseline"
run() { name=$1; shift
  for cfg in "base:" "zoomo:--zoom 4 --steps 50" "zoomt:--zoom 4 --regime translucent --steps 20"; do
    tag=${cfg%%:*}; extra=${cfg#*:}
    env "$@" $B $extra > gpurun_out/r3j_${name}_${tag}.json 2> gpurun_out/r3j_${name}_${tag}.err
    python - <<P
import json
try:
    d=json.load(open("gpurun_out/r3j_${name}_${tag}.json"))
    print("${name} ${tag}: ms/step %.4f march_ms %.4f frac %.3f samples %.0f e2e_fps %.1f" % (d["ms_per_step"], d["roofline"]["kernel_ms_per_launch"], d["roofline"]["frac"], d["samples_per_frame"], d["e2e"]["fps"]))
except Exception as e: print("${name} ${tag} failed", e)
P
  done
}
run shared X=1
run plain NMR_NO_SHARED_ENCODE=1
run shared2 X=1
