#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): bench line, ncu launch list of one pass, ncu --set full of the RA-pair kernel
# forms and of the ps_shout phase kernel.  Nothing measured under ncu is a bench number.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
(time timeout 900 python bench.py --steps 5 --warmup 3 --no-sweep) > gpurun_out/r2_bench_d.json 2> gpurun_out/r2_bench_d.err
export JA_NO_AHEAD=1 JA_NO_PERSIST=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/r2_launches_nanogpt_pass.csv python scripts/ncu_pass.py > gpurun_out/r2_ncu_pass.out 2>&1
for form in wide small large; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_round_prod_bool --launch-skip 3 --launch-count 1 -f -o /tmp/prof_pair_$form python scripts/ncu_pair.py $form > gpurun_out/r2_ncu_pair_$form.out 2>&1
  ncu -i /tmp/prof_pair_$form.ncu-rep --page raw --csv > gpurun_out/r2_ncu_pair_${form}_raw.csv 2>/dev/null
done
for skip in 72 100; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ps_phase --launch-skip $skip --launch-count 1 -f -o /tmp/prof_ps_$skip python scripts/ps_probe.py > gpurun_out/r2_ncu_ps_$skip.out 2>&1
  ncu -i /tmp/prof_ps_$skip.ncu-rep --page raw --csv > gpurun_out/r2_ncu_ps_${skip}_raw.csv 2>/dev/null
done
tail -2 gpurun_out/r2_bench_d.err; tail -2 gpurun_out/r2_ncu_pass.out; wc -l gpurun_out/r2_launches_nanogpt_pass.csv; ls -la gpurun_out/r2_ncu_*_raw.csv
