#!/bin/bash
# tools/gpu_dist.sh N TAG [ENV…] — on N GPUs of one box: the slab path against the single-GPU path (tests/dist_check.py), then one
# bench line per environment setting (first: defaults).
N=$1; TAG=$2; shift 2
O=gpurun_out; mkdir -p $O
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ -z "$SKIP_CHECK" ]; then
timeout 400 $RUN --master-port 29511 tests/dist_check.py > $O/${TAG}_dist_check.log 2>&1; echo "dist_rc=$?"
grep -c " ok" $O/${TAG}_dist_check.log
fi
if [ -n "$DIST_CHECK_ENV" ]; then   # the same check under another level plan (e.g. WL_DIST_REPL_CELLS=…)
  timeout 400 env $DIST_CHECK_ENV $RUN --master-port 29512 tests/dist_check.py > $O/${TAG}_dist_check_b.log 2>&1; echo "dist_b_rc=$? ($DIST_CHECK_ENV)"
  grep -c " ok" $O/${TAG}_dist_check_b.log
fi
i=0
for E in "WL_NOP=1" "$@"; do
  timeout 300 env $E $RUN --master-port $((29520 + i)) bench.py --gpus $N --steps 10 --warmup 3 > $O/${TAG}_bench$i.json 2> $O/${TAG}_bench$i.err
  echo "[$i] $E rc=$? $(python tools/show_bench.py $O/${TAG}_bench$i.json | head -1)"
  i=$((i+1))
done
