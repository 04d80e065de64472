#!/bin/bash
# Round-end evidence on one B200 (run under gpurun): full GPU test suite, smoke, bench (both arms), the ncu launch list of the
# bench command, --set full captures of every kernel family (exported to CSV on the box: the .ncu-rep files are too big to
# bring back through gpurun_out/).  Outputs under gpurun_out/<tag>_*.
tag=${1:-r02_final}
python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${tag}_tests.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.log 2>&1
python bench.py > gpurun_out/${tag}_bench.log 2> gpurun_out/${tag}_bench.err
K='vm_cycles|vm_link|vm_sponge_kernel|vm_sponge_trace|vm_check|vm_gadgets|ram_rows|ram_inverse|ram_check|ev_rows|ev_check|st_rows|st_push_rows|st_check|rq_push|lh_rows|lh_chain|check_columns'
ncu --set full --clock-control none -k regex:"$K" -c 48 -o /tmp/${tag}_full python tools/profile_all.py 18 > gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_ncu_full_raw.csv 2>> gpurun_out/${tag}_ncu_full.log
python tools/ncu_raw_summary.py gpurun_out/${tag}_ncu_full_raw.csv > gpurun_out/${tag}_ncu_summary.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches_bench.csv python bench.py --steps 2 --warmup 1 > gpurun_out/${tag}_launches_bench.out 2>&1
ncu --set full --clock-control none --import-source on -k regex:'vm_cycles' -c 1 -o /tmp/${tag}_vmc python tools/profile_all.py 18 >> gpurun_out/${tag}_ncu_full.log 2>&1
ncu -i /tmp/${tag}_vmc.ncu-rep --page source --print-source cuda,sass --csv > /tmp/${tag}_vmc_source.csv 2>> gpurun_out/${tag}_ncu_full.log
python tools/ncu_source_top.py /tmp/${tag}_vmc_source.csv 60 > gpurun_out/${tag}_ncu_vm_cycles_source_top.txt 2>&1
ls -la /tmp/${tag}_*.ncu-rep /tmp/${tag}_vmc_source.csv >> gpurun_out/${tag}_ncu_full.log
tail -3 gpurun_out/${tag}_tests.log; tail -2 gpurun_out/${tag}_smoke.log; tail -c 600 gpurun_out/${tag}_bench.log; tail -3 gpurun_out/${tag}_ncu_full.log; du -sh gpurun_out
