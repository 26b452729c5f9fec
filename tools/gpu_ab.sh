#!/bin/bash
# tools/gpu_ab.sh TAG "ENV1" "ENV2" … — GPU parity tests once, then one short bench line per environment setting
# (A/B runs of switches and of library variants built with tools/build_variant.sh; WL_LIB_PATH selects a variant).
TAG=$1; shift
O=gpurun_out; mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest_rc=$?"; tail -3 $O/${TAG}_pytest.log
i=0
for E in "$@"; do
  env $E python bench.py --steps 10 --warmup 3 --no-cpu > $O/${TAG}_ab$i.json 2> $O/${TAG}_ab$i.err
  echo "[$i] $E rc=$? $(python tools/show_bench.py $O/${TAG}_ab$i.json | head -1)"
  i=$((i+1))
done
