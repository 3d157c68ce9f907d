#!/bin/bash
# compute-sanitizer over the hot path (SURVEY.md section 5 commitment).  Run on a GPU box:
#   bash scripts/sanitize.sh            -> gpurun_out/sanitize_*.log + a summary line per tool
# memcheck / racecheck / synccheck / initcheck over __graft_entry__.smoke() (build, search, fused update, Add_Points) and over
# the streaming test (delete boxes + update + device map_incremental per scan).
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
run() {  # tool, name, command...
  tool=$1; name=$2; shift 2
  timeout 900 $CS --tool $tool --print-limit 20 "$@" > gpurun_out/sanitize_${tool}_${name}.log 2>&1
  rc=$?
  echo "$tool $name rc=$rc: $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|errors' gpurun_out/sanitize_${tool}_${name}.log | tail -1)"
}
for tool in memcheck racecheck synccheck initcheck; do
  run $tool smoke python -c "import __graft_entry__ as g; g.smoke()"
done
run memcheck stream python -m pytest tests/test_gpu_stream.py -x -q
run racecheck stream python -m pytest tests/test_gpu_stream.py -x -q
