#!/bin/bash
# compute-sanitizer over smoke-size inputs of every kernel family: memcheck on the smoke run, the
# WMF / NCF / MF / evaluation / graph tests, racecheck on the shared-memory heavy ones.  Summaries -> gpurun_out/sanitizer_*.log
#   bash tools/sanitize.sh            (1 GPU)
#   bash tools/sanitize.sh dist       (2 GPUs: memcheck of the peer-memory epoch, both ranks)
#   bash tools/sanitize.sh ncf        (NCF epoch strategies + graph dropout; NOT yet run: added after the round's GPU budget was spent)
set -u
OUT=gpurun_out
mkdir -p $OUT
CS="compute-sanitizer --print-limit 20 --error-exitcode 1"
run() {  # name, tool, command...
  local name=$1 tool=$2; shift 2
  timeout 900 $CS --tool $tool "$@" > $OUT/sanitizer_${name}_${tool}.log 2>&1
  echo "$name $tool rc=$? : $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' $OUT/sanitizer_${name}_${tool}.log | tail -1)"
}
if [ "${1:-}" = "dist" ]; then
  # --report-api-errors no: torch's own fabric-handle probe (c10::cuda::isFabricSupported -> cuMemCreate) is refused on this
  # box and would be counted as two errors that have nothing to do with a kernel
  run dist memcheck --target-processes all --report-api-errors no python -m pytest tests/test_gpu_dist.py -x -q -k sharded
  exit 0
fi
if [ "${1:-}" = "aush" ]; then
  run aush memcheck python -m pytest tests/test_gpu_aush.py -x -q -k "bit_stable or train_step"
  run aush racecheck python -m pytest tests/test_gpu_aush.py -x -q -k "bit_stable"
  exit 0
fi
if [ "${1:-}" = "ncf" ]; then
  # the NCF epoch strategies (graph replay, lazy embedding Adam, duplicate links, out-of-range ids) and LightGCN graph dropout
  run ncf_epoch memcheck python -m pytest tests/test_gpu_models.py -x -q -k "bit_for_bit or out_of_range or dropout"
  run ncf_epoch racecheck python -m pytest tests/test_gpu_models.py -x -q -k "out_of_range"
  exit 0
fi
run smoke memcheck python -c "import __graft_entry__ as g; g.smoke()"
run wmf memcheck python -m pytest tests/test_gpu_wmf.py -x -q -k "unrolled or plain"
run models memcheck python -m pytest tests/test_gpu_models.py -x -q -k "small or mf or bit_stable"
run eval memcheck python -m pytest tests/test_gpu_eval_workflow.py -x -q -k "fullrank or candidate or recall"
run graph memcheck python -m pytest tests/test_gpu_graph_spmm.py -x -q -k "not scale"
run smoke racecheck python -c "import __graft_entry__ as g; g.smoke()"
run wmf racecheck python -m pytest tests/test_gpu_wmf.py -x -q -k "unrolled"
run models racecheck python -m pytest tests/test_gpu_models.py -x -q -k "small and fp32"
run aush memcheck python -m pytest tests/test_gpu_aush.py -x -q -k "bit_stable or train_step"
run aush racecheck python -m pytest tests/test_gpu_aush.py -x -q -k "bit_stable"
