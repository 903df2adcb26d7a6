#!/bin/bash
# One parametrised GPU session script (replaces the one-off tools/gpu_*.sh of round 1).
#   tools/gpu_run.sh tests            # pytest -m gpu (all, not -x) + smoke
#   tools/gpu_run.sh bench [args...]  # python bench.py args
#   tools/gpu_run.sh py <script> [args...]
#   tools/gpu_run.sh ncu <kernel-regex> -- <command...>   # one --set full capture of the first matching launch
#   tools/gpu_run.sh launches -- <command...>             # launch list (gpu__time_duration.sum)
# Several steps can be chained with ':::'  e.g.  tools/gpu_run.sh tests ::: bench --steps 5
# Every step logs to gpurun_out/<tag>_<step>.log; TAG env names the session (default r02).
set -u
mkdir -p gpurun_out
TAG=${TAG:-r02}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/${TAG}_gpu.txt 2>&1
nproc >> gpurun_out/${TAG}_gpu.txt
step=0
run_step() {
  local kind=$1; shift
  step=$((step+1))
  local log=gpurun_out/${TAG}_${step}_${kind}.log
  case "$kind" in
    tests)  timeout ${T:-1500} python -m pytest tests -m gpu -q "$@" > "$log" 2>&1; tail -25 "$log"
            timeout 300 python -c "import __graft_entry__ as g; g.smoke()" >> "$log" 2>&1; tail -2 "$log" ;;
    bench)  timeout ${T:-1500} python bench.py "$@" > "$log" 2>&1; tail -5 "$log" ;;
    py)     timeout ${T:-1500} python "$@" > "$log" 2>&1; tail -40 "$log" ;;
    sh)     timeout ${T:-1500} bash -c "$*" > "$log" 2>&1; tail -40 "$log" ;;
    ncu)    local rx=$1; shift; [ "$1" = "--" ] && shift
            timeout ${T:-1500} ncu --set full --clock-control none --import-source on -k "regex:$rx" -s ${NCU_SKIP:-0} -c ${NCU_COUNT:-1} \
              -f -o gpurun_out/${TAG}_${step}_ncu "$@" > "$log" 2>&1; tail -5 "$log" ;;
    launches) [ "$1" = "--" ] && shift
            timeout ${T:-1500} ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_COUNT:-400} --csv \
              --log-file gpurun_out/${TAG}_${step}_launches.csv "$@" > "$log" 2>&1; tail -5 "$log" ;;
    *) echo "unknown step $kind"; return 1 ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" = ":::" ]; then run_step "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_step "${args[@]}"
exit 0
