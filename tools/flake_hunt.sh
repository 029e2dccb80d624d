#!/bin/bash
# Fresh-process repetitions of the bit-exactness test that failed 2 of ~60 runs in round 1, plus
# compute-sanitizer passes over it. Usage: tools/flake_hunt.sh [runs] [parallel]   (run on the GPU box under gpurun)
N=${1:-40}; PAR=${2:-4}
OUT=gpurun_out/flake; mkdir -p $OUT; rm -f $OUT/*
{
  lscpu | grep -E 'Model name|^CPU\(s\)|Thread|Socket'
  python -m pytest tests/test_oracle_golden.py -q -k golden_inputs -p no:cacheprovider 2>&1 | tail -2; python - <<'PY'
import torch, hashlib
from tests.golden import cases
print("torch cpu capability", torch.backends.cpu.get_cpu_capability(), "threads", torch.get_num_threads())
f, c = cases.dense_inputs(64, 96)
sha = lambda t: hashlib.sha256(t.numpy().tobytes()).hexdigest()[:12]
print("dense_inputs sha", sha(f), sha(c), "(build container: b00239533f11 c7c1072af39e)")
PY
} > $OUT/host.txt 2>&1
run_one() { python -m pytest tests/test_gpu_kernels.py -x -q -k "golden_inputs or corr or lookup or dense_postproc" -p no:cacheprovider > gpurun_out/flake/run_$1.log 2>&1 && echo "ok $1" || echo "FAIL $1"; }
export -f run_one
seq 1 $N | xargs -P $PAR -I{} bash -c 'run_one {}' > $OUT/summary.txt 2>&1
echo "runs $N failures $(grep -c FAIL $OUT/summary.txt)" >> $OUT/summary.txt
# keep only the failing logs (+ one passing sample)
for f in $OUT/run_*.log; do n=${f##*run_}; n=${n%.log}; grep -q "FAIL $n\$" $OUT/summary.txt || { [ "$n" = 1 ] || rm -f $f; }; done
for tool in memcheck initcheck racecheck; do
  timeout 600 compute-sanitizer --tool $tool --print-limit 20 python -m pytest tests/test_gpu_kernels.py -x -q -k "dense_postproc" -p no:cacheprovider > $OUT/sanitizer_$tool.log 2>&1
  echo "$tool exit $?" >> $OUT/summary.txt
done
tail -5 $OUT/summary.txt
