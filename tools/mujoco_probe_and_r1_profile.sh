#!/bin/bash
# Round-2 call 1: MuJoCo probe on the GPU box, GPU test-suite, bench of the shipped binary, sanitizer runs,
# latency scan, ncu of the shipped binary.  Outputs: gpurun_out/probe_*
set -u
O=gpurun_out
mkdir -p $O
python - > $O/r2c1_mujoco_probe.txt 2>&1 <<'PY'
import importlib.util as u, subprocess, sys, glob
for m in ("mujoco", "mujoco_py", "dm_control", "gym", "gymnasium", "pyquaternion", "brax", "mujoco_mjx"):
    print(m, "->", u.find_spec(m))
print(subprocess.run([sys.executable, "-m", "pip", "list"], capture_output=True, text=True).stdout.lower().count("mujoco"), "pip packages mention mujoco")
print("wheelhouse:", [p for p in glob.glob("/opt/wheelhouse/*") if "mujoco" in p.lower() or "gym" in p.lower()])
print("libmujoco on disk:", subprocess.run("find / -xdev \\( -name 'libmujoco*' -o -name 'mjpro*' -o -name '.mujoco' \\) 2>/dev/null | head", shell=True, capture_output=True, text=True).stdout.strip() or "none")
print("/root/reference exists:", __import__("os").path.exists("/root/reference"))
PY
(time python -m pytest tests -m gpu -q -x) > $O/r2c1_pytest.log 2>&1
python bench.py > $O/r2c1_bench_1gpu.json 2> $O/r2c1_bench_1gpu.err
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/gpu_small_rollout.py 64 6 > $O/r2c1_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/gpu_small_rollout.py 64 3 > $O/r2c1_racecheck.log 2>&1
python tools/gpu_latency_scan.py > $O/r2c1_latency_scan.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_step -s 60 -c 1 -f -o $O/r2c1_kstep \
    python bench.py --steps 80 --warmup 20 --no-cpu-baseline > $O/r2c1_ncu.log 2>&1
DMB_TRACE=1 python tools/gpu_cta_trace.py 4096 > $O/r2c1_cta_trace.txt 2>&1
nvidia-smi > $O/r2c1_nvidia_smi.txt 2>&1
cat $O/r2c1_mujoco_probe.txt; tail -3 $O/r2c1_pytest.log; cut -c1-300 $O/r2c1_bench_1gpu.json; tail -5 $O/r2c1_memcheck.log; tail -5 $O/r2c1_racecheck.log; cat $O/r2c1_latency_scan.txt
