O=gpurun_out
python -m pytest tests -m gpu -x -q > $O/r02f_pytest.log 2>&1; echo "pytest_rc=$?"; tail -5 $O/r02f_pytest.log
python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02f_bench.json 2> $O/r02f_bench.err; echo rc=$?
WL_CFL_FUSE=0 python bench.py --steps 10 --warmup 3 --no-cpu > $O/r02f_bench_nocfl.json 2>> $O/r02f_bench.err; echo rc=$?
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_resjac|k_project_pre|k_project_correct_cfl|k_cfl_upper' --launch-skip 4 -c 6 -f -o $O/r02f_full python bench.py --steps 1 --warmup 1 --no-cpu > $O/r02f_ncu_full.log 2>&1; echo "full_rc=$?"
