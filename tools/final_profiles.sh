#!/bin/bash
# Round-end measurement set (one B200): bench line, reference arm, launch list, ncu --set full captures.  Outputs under gpurun_out/.
TAG=${1:-r2}
python bench.py > gpurun_out/${TAG}_bench_final.json 2> gpurun_out/${TAG}_bench_final.err
python tools/dialogue_rnn_step.py > gpurun_out/${TAG}_dialogue_rnn_step.json 2> gpurun_out/${TAG}_dialogue_rnn_step.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_step_final.csv python profiles/profile_step.py > gpurun_out/${TAG}_ncu_launches.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 30 --launch-count 14 -f -o gpurun_out/${TAG}_full_gemm python profiles/profile_net.py > gpurun_out/${TAG}_ncu_full_gemm.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"attention|layernorm|reduce_ln" --launch-skip 8 --launch-count 8 -f -o gpurun_out/${TAG}_full_attn python profiles/profile_net.py > gpurun_out/${TAG}_ncu_full_attn.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"graph_build" --launch-skip 3 --launch-count 1 -f -o gpurun_out/${TAG}_full_graph_build python tools/graph_bench.py > /dev/null 2> gpurun_out/${TAG}_ncu_full_graph.log
ncu --set full --clock-control none --import-source on -k regex:"graph_gather" --launch-skip 8 --launch-count 14 -f -o gpurun_out/${TAG}_full_graph python tools/graph_bench.py > /dev/null 2>> gpurun_out/${TAG}_ncu_full_graph.log
python -m pytest tests -m gpu -q -s > gpurun_out/${TAG}_pytest_full.log 2>&1; tail -2 gpurun_out/${TAG}_pytest_full.log; grep PARITY gpurun_out/${TAG}_pytest_full.log > gpurun_out/${TAG}_parity_full_size.txt
# keep the merged output small (gpurun copies back at most 64 MiB): compact CSV exports instead of the raw reports
for n in gemm attn graph graph_build; do
  if [ -f gpurun_out/${TAG}_full_$n.ncu-rep ]; then
    ncu -i gpurun_out/${TAG}_full_$n.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_summary.py > gpurun_out/${TAG}_ncu_full_$n.csv
    rm -f gpurun_out/${TAG}_full_$n.ncu-rep
  fi
done
ls -la gpurun_out/${TAG}_*
