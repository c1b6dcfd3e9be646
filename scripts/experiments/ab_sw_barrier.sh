for v in 0 1 0 1 0 1; do
  PNP_SW_BARRIER=$v python bench.py --sections= --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('sw_barrier=$v', round(d['us_per_inner_iteration'],2), {k:round(v,2) for k,v in d['kernel_us'].items()}, d['psnr_first_last'])"
done
python -m pytest tests/test_gpu_epoch.py -q -m gpu -x 2>&1 | tail -2
PNP_SW_BARRIER=1 python -m pytest tests/test_gpu_epoch.py -q -m gpu -x 2>&1 | tail -2
