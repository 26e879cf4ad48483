#!/bin/bash
# Round-2 measurement pass on one B200 (run under gpurun): the default bench line, the ncu launch list of one pass, ncu --set full of the
# RA-pair kernel forms and of the ps_shout phase kernel.  Nothing measured under ncu is a bench number.
cd "${GRAFT_REPO_ROOT:-.}"; mkdir -p gpurun_out
(time timeout 1500 python bench.py) > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err
export JA_NO_AHEAD=1 JA_NO_PERSIST=1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/r2_launches_nanogpt_pass.csv python scripts/ncu_pass.py > gpurun_out/r2_ncu_pass.out 2>&1
for form in wide small large; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_round_prod_bool --launch-skip 3 --launch-count 1 -f -o /tmp/prof_pair_$form python scripts/ncu_pair.py $form > gpurun_out/r2_ncu_pair_$form.out 2>&1
  ncu -i /tmp/prof_pair_$form.ncu-rep --page raw --csv > gpurun_out/r2_ncu_pair_${form}_raw.csv 2>/dev/null
done
# ps_shout phase pass: scripts/ps_phase_probe.py launches 60 phase kernels at T = 2^12, then clamp (48) and remainder (12) at T = 2^14
for skip in 73 111; do
  timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ps_phase --launch-skip $skip --launch-count 1 -f -o /tmp/prof_ps_$skip python scripts/ps_phase_probe.py > gpurun_out/r2_ncu_ps_$skip.out 2>&1
  ncu -i /tmp/prof_ps_$skip.ncu-rep --page raw --csv > gpurun_out/r2_ncu_ps_${skip}_raw.csv 2>/dev/null
done
unset JA_NO_AHEAD JA_NO_PERSIST
JA_SC_TRACE=1 timeout 300 python scripts/pass_times.py > gpurun_out/r2_trace_final.log 2>&1       # -> scripts/trace_summary.py
tail -2 gpurun_out/r2_bench_final.err; tail -2 gpurun_out/r2_ncu_pass.out; wc -l gpurun_out/r2_launches_nanogpt_pass.csv; ls -la gpurun_out/r2_ncu_*_raw.csv
