#!/bin/bash
# tools/gpu_round.sh TAG — the standard 1-GPU evidence pass on a B200 box: GPU parity tests, the default bench line,
# the ncu launch list of the same command, and one `ncu --set full` capture of the dominant kernels.
TAG=${1:-x}
O=gpurun_out
mkdir -p $O
# (the merged gpurun_out/ must stay below 64 MiB in total: delete old .ncu-rep files locally before calling this)
if [ "$2" != "notest" ]; then
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest_rc=$?"
tail -3 $O/${TAG}_pytest.log
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo "smoke_rc=$?"
fi
python bench.py --steps 20 --warmup 5 > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench_rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>> $O/${TAG}_bench.err; echo "ref_rc=$?"
if [ "$2" != "noprof" ]; then
python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file $O/${TAG}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/${TAG}_ncu_list.log 2>&1; echo "list_rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'k_conv_diff3_tma2|k_resjac|k_conv_diff_shell|k_prolong_increment_init_m|k_pcg_update|k_pcg_mult_s|k_project_correct|k_project_pre|k_vcycle_tail' \
  --launch-skip 30 -c 18 -f -o $O/${TAG}_full python bench.py --steps 1 --warmup 1 --no-cpu > $O/${TAG}_ncu_full.log 2>&1; echo "full_rc=$?"
fi
