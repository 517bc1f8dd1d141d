timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s5f_bench.json 2> gpurun_out/s5f_bench.err; echo rc=$?
python -c "
import json;d=json.load(open('gpurun_out/s5f_bench.json'));print({k:d[k] for k in ('value','ms_per_step','frac_of_bf16_peak','clocks')});print(d['parity']['gate'], d['checksum']['bits']);print({k:(v['ms_per_step'],v.get('frac_of_bf16_peak'),v.get('checksum_bits')) for k,v in d['secondary'].items()})"
