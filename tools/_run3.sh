timeout 600 python -m pytest tests/test_gpu_decoders.py -m gpu -x -q -k "beam" > gpurun_out/t4_tests.log 2>&1; echo tests_rc=$?; tail -3 gpurun_out/t4_tests.log
for r in 1 0; do ICD_BEAM_ATT_RING=$r timeout 300 python bench.py --workload beam --steps 5 --warmup 3 --no-cpu-baseline 2>gpurun_out/t4_beam_$r.err | tail -1 > gpurun_out/t4_beam_$r.json; python -c "
import json;d=json.loads(open('gpurun_out/t4_beam_$r.json').read());print('ring=$r', d.get('value'))"; done
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:att_step_fwd_grouped -c 26 --csv --log-file gpurun_out/t4_launches.csv python bench.py --workload beam --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/t4_ncu.log 2>&1
python - <<'P'
import csv
rows=[r for r in csv.reader(open('gpurun_out/t4_launches.csv')) if len(r)>5 and r[0].isdigit()]
print([round(float(r[-1])/1000,1) for r in rows])
P
